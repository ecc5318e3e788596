"""Weight-streaming regime (small M, huge K): cold weights (cycled through > L2), row-major vs block-tiled."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "latent-diffusion-segmentation_b200"))
import torch
from ldmseg import _native as nat
from ldmseg import _pack as pk
torch.cuda.set_device(0)
dev = "cuda"

def case(nb, h, cin, n, bn, split, tiled, nw=8, iters=16):
    m = nb * h * h
    x = torch.randn(m, cin, device=dev).to(torch.bfloat16)
    kp = 9 * cin
    ws_ = []
    for _ in range(nw):
        w = (torch.randn(n, kp, device=dev) * 0.02).to(torch.bfloat16)
        ws_.append(pk.tile_pack(w) if tiled else w)
    out = torch.empty(m, n, device=dev, dtype=torch.bfloat16)
    wsp = torch.zeros(24 * 1024 * 1024, device=dev)
    cnt = torch.zeros(8192, device=dev, dtype=torch.int32)
    ps = [nat.make_igemm_params([x], [cin], nb, h, h, [(0, 9)], w, n, out, n, block_n=bn, split_k=split, workspace=wsp,
                                counters=cnt, weight_tiled=tiled) for w in ws_]
    for p in ps:
        nat.igemm(p)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(iters):
            nat.igemm(ps[i % nw])
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / iters
    return us, n * kp * 2 / us / 1e6  # TB/s of weight bytes

for (nb, h, cin, n) in [(1, 16, 2560, 1280), (1, 8, 2560, 1280), (1, 16, 1280, 1280), (8, 8, 2560, 1280)]:
    for tiled in (False, True):
        line = f"nb={nb} {h}x{h} cin={cin} n={n} tiled={int(tiled)}:"
        for (bn, split) in ((256, 14), (256, 7), (128, 7), (128, 3), (64, 3), (160, 9)):
            m_tiles = (nb * h * h + 127) // 128
            if m_tiles * ((n + bn - 1) // bn) * split > 148:
                continue
            us, tbs = case(nb, h, cin, n, bn, split, tiled)
            line += f"  bn{bn}/s{split}: {us:6.1f}us {tbs:4.2f}TB/s"
        print(line, flush=True)
