"""Does programmatic dependent launch shorten a chain of small dependent igemm launches inside a CUDA graph?"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "latent-diffusion-segmentation_b200"))
import torch
from ldmseg import _native as nat
torch.cuda.set_device(0)
dev = "cuda"
m, c = 4096, 320
x = [torch.randn(m, c, device=dev).to(torch.bfloat16) for _ in range(2)]
wt = (torch.randn(c, 320, device=dev) * 0.05).to(torch.bfloat16)
bias = torch.zeros(c, device=dev)
for pdl in (False, True):
    ps = []
    for i in range(40):   # ping-pong chain: each launch reads the previous launch's output
        ps.append(nat.make_igemm_params([x[i % 2]], [c], 1, 1, m, [(0, 1)], wt, c, x[(i + 1) % 2], c, bias=bias,
                                        block_n=128, pdl=pdl))
    old = nat.set_pdl(pdl)
    for p in ps[:4]:
        nat.igemm(p)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for p in ps:
            nat.igemm(p)
    nat.set_pdl(old)
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    print(f"pdl={pdl}: {e0.elapsed_time(e1) * 1e3 / 40:.2f} us per dependent igemm (graph)")
    e0.record()
    old = nat.set_pdl(pdl)
    for p in ps:
        nat.igemm(p)
    nat.set_pdl(old)
    e1.record(); torch.cuda.synchronize()
    print(f"pdl={pdl}: {e0.elapsed_time(e1) * 1e3 / 40:.2f} us per dependent igemm (stream)")
