"""Run ONE eager UNet forward (or VAE encode / seg decode) between cudaProfilerStart/Stop, for
`ncu --profile-from-start off ...` launch lists and full captures.  Also prints a CUDA-event per-kernel-family
breakdown when run without ncu (--events)."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "latent-diffusion-segmentation_b200"))

import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=1)
    ap.add_argument("--size", type=int, default=64)
    ap.add_argument("--what", default="unet", choices=["unet", "encode", "decode"])
    ap.add_argument("--events", action="store_true")
    ap.add_argument("--iters", type=int, default=1)
    args = ap.parse_args()
    from bench import build_models
    from ldmseg import _native as nat
    dev = torch.device("cuda:0")
    torch.cuda.set_device(dev)
    unet, vae_image, vae_semseg, sched = build_models(dev)
    if args.what == "unet":
        plan = unet._get_engine().plan(args.batch, args.size)
        plan.x_in.normal_()
        run = plan.run
    elif args.what == "encode":
        plan = vae_image._get_engine().plan(args.batch, args.size * 8)
        plan.x_in.normal_()
        run = plan.run
    else:
        plan = vae_semseg._get_engine().dec_plan(args.batch, args.size)
        plan.z_in.normal_()
        run = plan.run
    for _ in range(2):
        run()
    torch.cuda.synchronize()
    if args.events:
        fams = {}
        names = ["igemm", "groupnorm", "layernorm", "attention", "im2col_s2", "upsample2x", "convt_shuffle_ln",
                 "softmax_rows"]
        real = {n: getattr(nat, n) for n in names}
        evs = []

        def wrap(n):
            def f(*a, **k):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                real[n](*a, **k)
                e1.record()
                evs.append((n, e0, e1, a))
            return f
        for n in names:
            setattr(nat, n, wrap(n))
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        run()
        t1.record()
        torch.cuda.synchronize()
        for n in names:
            setattr(nat, n, real[n])
        tot = t0.elapsed_time(t1)
        for n, e0, e1, a in evs:
            fams.setdefault(n, []).append(e0.elapsed_time(e1))
        print(f"total eager forward: {tot:.3f} ms, {len(evs)} ops")
        for n, v in sorted(fams.items(), key=lambda kv: -sum(kv[1])):
            print(f"  {n:18s} n={len(v):4d} sum={sum(v):8.3f} ms  avg={1e3 * sum(v) / len(v):8.1f} us  max={1e3 * max(v):8.1f} us")
        # igemm detail: slowest launches with their shapes
        ig = [(e0.elapsed_time(e1), a[0]) for n, e0, e1, a in evs if n == "igemm"]
        ig.sort(key=lambda t: -t[0])
        print("  slowest igemm launches (ms, M, N, K, block_n, split_k):")
        for ms, p in ig[:25]:
            k = sum(p.seg_taps[i] * p.src_c[p.seg_src[i]] for i in range(p.nseg))
            m = p.nb * p.h * p.w
            tf = 2.0 * m * p.n * k / (ms * 1e-3) / 1e12
            print(f"    {ms:7.4f}  M={m:6d} N={p.n:5d} K={k:6d} bn={p.block_n:3d} split={p.split_k:2d}  {tf:7.1f} TFLOP/s")
        # graph replay time of the same plan
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            run()
        for _ in range(3):
            g.replay()
        t0.record()
        for _ in range(10):
            g.replay()
        t1.record()
        torch.cuda.synchronize()
        print(f"graph replay: {t0.elapsed_time(t1) / 10:.3f} ms per forward")
        return
    torch.cuda.profiler.start()
    for _ in range(args.iters):
        run()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()


if __name__ == "__main__":
    main()
