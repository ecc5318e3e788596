"""Which epilogue feature costs what: conv 320->320 @64x64 (bn160, split 2 and bn128 split 1) with toggles."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "latent-diffusion-segmentation_b200"))
import torch
from ldmseg import _native as nat
torch.cuda.set_device(0)
dev = "cuda"

def case(nb, h, w, cin, n, taps, bn, split, residual, rowbias, stats, act=0, iters=30):
    m = nb * h * w
    x = torch.randn(m, cin, device=dev).to(torch.bfloat16)
    kp = taps * ((cin + 63) // 64 * 64)
    wt = (torch.randn(n, kp, device=dev) * 0.02).to(torch.bfloat16)
    nout = n // 2 if act == nat.ACT_GEGLU else n
    out = torch.empty(m, nout, device=dev, dtype=torch.bfloat16)
    bias = torch.randn(n, device=dev)
    res = torch.randn(m, n, device=dev).to(torch.bfloat16) if residual else None
    rb = torch.randn(nb, n, device=dev) if rowbias else None
    st = torch.zeros(nb, n, 2, device=dev) if stats else None
    ws = torch.zeros(24 * 1024 * 1024, device=dev)
    cnt = torch.zeros(8192, device=dev, dtype=torch.int32)
    p = nat.make_igemm_params([x], [cin], nb, h, w, [(0, taps)], wt, n, out, nout, bias=bias, residual=res, res_ld=n,
                              rowbias=rb, rowbias_ld=n, act=act, block_n=bn, split_k=split, workspace=ws, counters=cnt,
                              stats=st, stats_hw=h * w)
    for _ in range(3):
        nat.igemm(p)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(iters):
            nat.igemm(p)
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / iters

for (bn, split) in ((160, 2), (128, 1), (160, 1)):
    line = f"conv 320@64x64 bn{bn}/s{split}:"
    for name, kw in (("plain", {}), ("res", dict(residual=True)), ("rowbias", dict(rowbias=True)), ("stats", dict(stats=True)),
                     ("all", dict(residual=True, rowbias=True, stats=True))):
        a = dict(residual=False, rowbias=False, stats=False); a.update(kw)
        line += f"  {name}: {case(1, 64, 64, 320, 320, 9, bn, split, **a):6.1f}us"
    print(line, flush=True)
for act, nm in ((0, "none"), (1, "silu"), (2, "geglu")):
    print(f"linear 4096x320->2560 bn256 act={nm}: {case(1, 1, 4096, 320, 2560, 1, 256, 1, False, False, False, act=act):6.1f}us", flush=True)
print(f"linear 4096x320->320 +res+stats bn128: {case(1, 1, 4096, 320, 320, 1, 128, 1, True, False, True):6.1f}us")
