"""Per-phase clock() attribution of the attention kernel (needs the -DLDMSEG_ATTN_TIMING build:
python build_native.py --variant tim -DLDMSEG_ATTN_TIMING; LDMSEG_LIB=.../libldmseg_b200_tim.so)."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "latent-diffusion-segmentation_b200"))
import torch
from ldmseg import _native as nat
torch.cuda.set_device(0)
lib = ctypes.CDLL(os.environ["LDMSEG_LIB"])
names = ["wait S", "S->regs", "row max", "-", "exp + sums + pack", "wait P.V", "P store", "-"]
for (nb, ntok, heads, d) in [(1, 4096, 8, 40), (8, 4096, 8, 40), (8, 1024, 8, 80)]:
    qkv = torch.randn(nb * ntok, 3 * heads * d, device="cuda").to(torch.bfloat16)
    out = torch.empty(nb * ntok, heads * d, device="cuda", dtype=torch.bfloat16)
    for _ in range(3):
        nat.attention(qkv, nb, ntok, heads, d, out)
    torch.cuda.synchronize()
    buf = (ctypes.c_uint32 * 32)()
    lib.ldmseg_attn_timing_read(buf)
    v = list(buf)
    iters = ntok // (128 if d <= 40 else 64)
    print(f"nb={nb} ntok={ntok} d={d}: cycles per iteration (CTA 0), {iters} iterations")
    for t in range(2):
        tot = sum(v[t * 8:t * 8 + 8])
        print(f"  softmax tile {t}: " + ", ".join(f"{names[i]} {v[t * 8 + i] / iters:.0f}" for i in range(8)) + f" | total {tot / iters:.0f}")
    for t in range(2):
        print(f"  MMA issuer {t}: waiting {v[16 + 4 * t] / iters:.0f}, issue S {v[17 + 4 * t] / iters:.0f}, issue P.V {v[18 + 4 * t] / iters:.0f}")
