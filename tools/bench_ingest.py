"""Which operand stream bounds the implicit GEMM?  Times one igemm (graph of back-to-back launches, CUDA events)
with the A and/or B loads of all but the first k-block of every work item left out (ldmseg_set_debug bits 8 / 16;
results are garbage, timing is what matters), for the 1-CTA and the CTA-pair kernel."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "latent-diffusion-segmentation_b200"))
import torch  # noqa: E402
from ldmseg import _native as nat  # noqa: E402
from ldmseg import _pack as pk  # noqa: E402


def run_case(nb, h, w, cin, n, taps, bn, pair, debug, iters=20, act=0):
    dev = "cuda"
    m = nb * h * w
    x = torch.randn(m, cin, device=dev).to(torch.bfloat16)
    kp = taps * ((cin + 63) // 64 * 64)
    wt = pk.to_bf16(pk.tile_pack(torch.randn(n, kp, device=dev) * 0.02))
    n_out = n // 2 if act == nat.ACT_GEGLU else n
    out = torch.empty(m, n_out, device=dev, dtype=torch.bfloat16)
    bias = torch.randn(n, device=dev)
    p = nat.make_igemm_params([x], [cin], nb, h, w, [(0, taps)], wt, n, out, n_out, bias=bias, block_n=bn,
                              weight_tiled=True, pair=pair, act=act)
    lib = nat.load()
    old = lib.ldmseg_set_debug(debug)
    try:
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
        e1.record()
        torch.cuda.synchronize()
    finally:
        lib.ldmseg_set_debug(old)
    us = e0.elapsed_time(e1) * 1e3 / iters
    return us, 2.0 * m * n * taps * cin / us / 1e6


def sweep_bn():
    """MMA-only time (no operand loads after the first k-block) per tile width: does N = 160 cost N/2 cycles?"""
    for nb, h, w, cin, n, taps in ((1, 1, 8192, 4096, 3840, 1), (8, 64, 64, 320, 320, 9), (8, 64, 64, 320, 640, 9),
                                   (8, 32, 32, 640, 640, 9), (8, 32, 32, 640, 1280, 9)):
        for pair in (False, True):
            line = f"nb={nb} {h}x{w} cin={cin} n={n} taps={taps} pair={int(pair)}:"
            for bn in (128, 160, 256):
                us0, tf0 = run_case(nb, h, w, cin, n, taps, bn, pair, 0)
                us1, tf1 = run_case(nb, h, w, cin, n, taps, bn, pair, 24)
                line += f"  bn{bn}: full {us0:6.1f}us {tf0:5.0f}TF  none {us1:6.1f}us {tf1:5.0f}TF |"
            print(line, flush=True)


def main():
    torch.cuda.set_device(0)
    if len(sys.argv) > 1 and sys.argv[1] == "bn":
        return sweep_bn()
    if len(sys.argv) > 1 and sys.argv[1] == "ff":
        # transformer feed-forward shapes: short K, wide N -- epilogue / store bound?
        for nb, h, w, cin, n, taps, bn, act in ((8, 1, 4096, 320, 2560, 1, 256, 2), (8, 1, 4096, 320, 2560, 1, 256, 0),
                                                (8, 1, 4096, 1280, 320, 1, 160, 0), (8, 1, 1024, 640, 5120, 1, 256, 2),
                                                (8, 1, 1024, 2560, 640, 1, 160, 0), (8, 1, 4096, 320, 960, 1, 256, 0)):
            for pair in (False, True):
                line = f"nb={nb} {h}x{w} cin={cin} n={n} bn={bn} act={act} pair={int(pair)}:"
                for dbg, nm in ((0, "full"), (24, "noAB"), (32, "noEpi"), (56, "mmaOnly")):
                    us, tf = run_case(nb, h, w, cin, n, taps, bn, pair, dbg, act=act)
                    line += f"  {nm} {us:7.1f}us {tf:5.0f}TF"
                print(line, flush=True)
        return
    cases = [(8, 64, 64, 320, 320, 9, 160), (8, 32, 32, 640, 640, 9, 256), (8, 16, 16, 1280, 1280, 9, 256),
             (8, 1, 4096, 320, 2560, 1, 256), (8, 1, 4096, 1280, 320, 1, 160), (1, 64, 64, 320, 320, 9, 160),
             (1, 1, 8192, 4096, 4096, 1, 256)]
    for nb, h, w, cin, n, taps, bn in cases:
        for pair in (False, True):
            line = f"nb={nb} {h}x{w} cin={cin} n={n} taps={taps} bn={bn} pair={int(pair)}:"
            for dbg, nm in ((0, "full"), (8, "noA"), (16, "noB"), (24, "none")):
                us, tf = run_case(nb, h, w, cin, n, taps, bn, pair, dbg)
                line += f"  {nm} {us:7.1f}us {tf:5.0f}TF"
            print(line, flush=True)


if __name__ == "__main__":
    main()
