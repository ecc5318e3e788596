"""Micro-benchmark of the tcgen05 implicit-GEMM kernel on UNet shapes (CUDA events, warm, back-to-back)."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "latent-diffusion-segmentation_b200"))
import torch  # noqa: E402
from ldmseg import _native as nat  # noqa: E402


def run_case(nb, h, w, cin, n, taps, bn, split, iters=30, residual=False):
    dev = "cuda"
    m = nb * h * w
    x = torch.randn(m, cin, device=dev).to(torch.bfloat16)
    kp = taps * ((cin + 63) // 64 * 64)
    wt = (torch.randn(n, kp, device=dev) * 0.02).to(torch.bfloat16)
    out = torch.empty(m, n, device=dev, dtype=torch.bfloat16)
    bias = torch.randn(n, device=dev)
    res = torch.randn(m, n, device=dev).to(torch.bfloat16) if residual else None
    ws = torch.zeros(24 * 1024 * 1024, device=dev)
    cnt = torch.zeros(8192, device=dev, dtype=torch.int32)
    p = nat.make_igemm_params([x], [cin], nb, h, w, [(0, taps)], wt, n, out, n, bias=bias, residual=res,
                              res_ld=n, block_n=bn, split_k=split, workspace=ws, counters=cnt)
    for _ in range(3):
        nat.igemm(p)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(iters):
            nat.igemm(p)
    g.replay()
    torch.cuda.synchronize()
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / iters
    fl = 2.0 * m * n * taps * cin
    return us, fl / us / 1e6


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--one", nargs=8, type=int, default=None, help="nb h w cin n taps bn split")
    args = ap.parse_args()
    torch.cuda.set_device(0)
    if args.one:
        nb, h, w, cin, n, taps, bn, split = args.one
        us, tf = run_case(nb, h, w, cin, n, taps, bn, split, iters=5)
        print(f"{us:.2f} us {tf:.1f} TFLOP/s")
        return
    cases = [
        # (nb,h,w,cin,n,taps)
        (1, 64, 64, 320, 320, 9), (1, 64, 64, 320, 320, 1), (1, 1, 4096, 320, 960, 1), (1, 1, 4096, 320, 2560, 1),
        (1, 1, 4096, 1280, 320, 1), (1, 32, 32, 640, 640, 9), (1, 1, 1024, 640, 640, 1), (1, 16, 16, 1280, 1280, 9),
        (1, 8, 8, 1280, 1280, 9), (1, 1, 256, 1280, 1280, 1), (8, 64, 64, 320, 320, 9), (8, 32, 32, 640, 640, 9),
        (8, 16, 16, 1280, 1280, 9), (8, 1, 4096, 320, 2560, 1), (1, 1, 8192, 4096, 4096, 1),
    ]
    for c in cases:
        nb, h, w, cin, n, taps = c
        line = f"nb={nb} {h}x{w} cin={cin} n={n} taps={taps}:"
        for bn in (64, 128, 160, 256):
            for split in (1, 4):
                if split > 1 and nb * h * w > 1024:
                    continue
                try:
                    us, tf = run_case(nb, h, w, cin, n, taps, bn, split)
                    line += f"  bn{bn}/s{split}: {us:7.1f}us {tf:6.0f}TF"
                except Exception as e:  # noqa: BLE001
                    line += f"  bn{bn}/s{split}: ERR"
        print(line, flush=True)


if __name__ == "__main__":
    main()
