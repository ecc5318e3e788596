"""Attention kernel timing on the UNet shapes (graph of repeated launches, CUDA events)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "latent-diffusion-segmentation_b200"))
import torch
from ldmseg import _native as nat
torch.cuda.set_device(0)
for (nb, ntok, heads, d) in [(1, 4096, 8, 40), (1, 1024, 8, 80), (1, 256, 8, 160), (1, 64, 8, 160), (8, 4096, 8, 40), (8, 1024, 8, 80)]:
    qkv = torch.randn(nb * ntok, 3 * heads * d, device="cuda").to(torch.bfloat16)
    out = torch.empty(nb * ntok, heads * d, device="cuda", dtype=torch.bfloat16)
    for _ in range(2):
        nat.attention(qkv, nb, ntok, heads, d, out)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(10):
            nat.attention(qkv, nb, ntok, heads, d, out)
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 100
    fl = 4.0 * nb * heads * ntok * ntok * d
    print(f"attn nb={nb} ntok={ntok} d={d}: {us:8.1f} us  {fl / us / 1e6:7.1f} TFLOP/s", flush=True)
