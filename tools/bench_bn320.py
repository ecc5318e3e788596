"""Pair tilings of the batch-2/4/8 convolution shapes, timed back to back in one CUDA graph: block_n 160 / 256 / 320,
whole tiles and stream-K tail.  Calibrates the planner's per-k-block cost of the 320-wide pair tile (plan.py).

Usage: python tools/bench_bn320.py [--batches 8 4 2]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "latent-diffusion-segmentation_b200"))
import torch  # noqa: E402
from ldmseg import _native as nat  # noqa: E402
from ldmseg import _pack as pk  # noqa: E402


def run_case(nb, h, cin, n, bn, tail, iters=20, residual=True, stats=True, out_f32=False):
    dev = "cuda"
    m = nb * h * h
    x = torch.randn(m, cin, device=dev).to(torch.bfloat16)
    wt = pk.to_bf16(pk.tile_pack(pk.pack_conv3x3(torch.randn(n, cin, 3, 3, device=dev) * 0.02)))
    out = torch.empty(m, n, device=dev, dtype=torch.float32 if out_f32 else torch.bfloat16)
    bias = torch.randn(n, device=dev)
    res = torch.randn(m, n, device=dev).to(torch.bfloat16) if residual else None
    st = torch.zeros(nb, n, 2, device=dev) if stats else None
    ws = torch.zeros(16 * 1024 * 1024, device=dev)
    cnt = torch.zeros(8192, device=dev, dtype=torch.int32)
    p = nat.make_igemm_params([x], [cin], nb, h, h, [(0, 9)], wt, n, out, n, bias=bias, residual=res, res_ld=n,
                              block_n=bn, workspace=ws, counters=cnt, stats=st, pdl=True, weight_tiled=True, pair=True,
                              weight_static=True, stream_k=tail)
    for _ in range(3):
        nat.igemm(p)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(iters):
            nat.igemm(p)
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (2 * iters)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batches", nargs="+", type=int, default=[8, 4, 2])
    ap.add_argument("--profile", nargs=5, type=int, default=None, help="nb h cin n bn")
    ap.add_argument("--stages", action="store_true",
                    help="where the time of one launch goes: the same case with the epilogue dropped, without fused "
                         "statistics, without operand loads (igemm debug switches; results are garbage)")
    args = ap.parse_args()
    torch.cuda.set_device(0)
    if args.profile:   # a few eager launches of one case for an ncu capture (ncu -k regex:igemm -s 3 -c 1)
        nb, h, cin, n, bn = args.profile
        run_case(nb, h, cin, n, bn, False, iters=2)
        return
    if args.stages:
        for (nb, h, cin, n) in [(8, 64, 320, 320), (8, 64, 960, 320), (8, 32, 640, 640)]:
            for bn, tail in ((320, False), (160, False)):
                line = f"nb={nb} {h}x{h} {cin}->{n} bn{bn}{'t' if tail else ''}:"
                for name, kw in (("full", {}), ("no residual", dict(residual=False)), ("no stats", dict(stats=False)),
                                 ("bias only", dict(residual=False, stats=False)),
                                 ("bias only, f32 out", dict(residual=False, stats=False, out_f32=True)),
                                 ("f32 out", dict(out_f32=True))):
                    line += f"  {name} {run_case(nb, h, cin, n, bn, tail, **kw):6.1f}"
                nat.load().ldmseg_set_debug(32)
                line += f"  no epilogue {run_case(nb, h, cin, n, bn, tail):6.1f}"
                nat.load().ldmseg_set_debug(0)
                print(line, flush=True)
        return
    shapes = [(64, 320, 320), (64, 640, 320), (64, 960, 320), (32, 320, 640), (32, 640, 640), (32, 1280, 640),
              (32, 1920, 640), (16, 640, 1280), (16, 1280, 1280), (16, 2560, 1280)]
    for nb in args.batches:
        for (h, cin, n) in shapes:
            kb = 9 * cin // 64
            line = f"nb={nb} {h}x{h} {cin}->{n} kb={kb} m={nb * h * h}:"
            best, best_us = None, 1e30
            for bn in (160, 256, 320):
                for tail in (False, True):
                    try:
                        us = run_case(nb, h, cin, n, bn, tail)
                    except Exception as e:  # noqa: BLE001
                        line += f"  bn{bn}{'t' if tail else ' '}:    ERR"
                        print(repr(e)[:200], flush=True)
                        continue
                    line += f"  bn{bn}{'t' if tail else ' '}: {us:7.1f}"
                    if us < best_us:
                        best, best_us = f"bn{bn}{'t' if tail else ''}", us
            fl = 2.0 * nb * h * h * n * 9 * cin
            print(line + f"   best {best} {fl / best_us / 1e6:5.0f} TF/s", flush=True)


if __name__ == "__main__":
    main()
