"""Measured tiling choice for the igemm launches of a plan (block_n, split_k, pair) on the current GPU.

The planner's cycle model (ldmseg/engine/plan.py choose_tiling) is a fit; this tool times every legal candidate for
each distinct (M, N, k-blocks) of a forward plan -- a CUDA graph of back-to-back launches with PDL, CUDA events --
and writes the winners to ldmseg/engine/tuned_b200.json, which choose_tiling consults first.

    python tools/tune_tiling.py --log profiles/r01_ablate_unet_b1_v15.log        (shapes from the op tags of a plan)
"""
import argparse
import json
import math
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "latent-diffusion-segmentation_b200"))
import torch  # noqa: E402
from ldmseg import _native as nat  # noqa: E402
from ldmseg import _pack as pk  # noqa: E402
from ldmseg.engine import plan as planmod  # noqa: E402

SMS = 148


def shapes_from_log(path):
    seen = {}
    for line in open(path):
        mt = re.search(r"igemm:(\d+):([^:]*):n(\d+):kb(\d+):bn(\d+):s(\d+):p(\d)", line)
        if mt:
            m, name, n, kb = int(mt.group(1)), mt.group(2), int(mt.group(3)), int(mt.group(4))
            seen.setdefault((m, n, kb), name)
    return seen


def time_case(m, n, kb, bn, split, pair, geglu, ws, cnt, iters=10):
    dev = "cuda"
    side = int(math.isqrt(m))
    conv = kb % 9 == 0 and kb >= 45 and side * side == m
    if conv:
        cin, nb, h, w, taps = kb // 9 * 64, 1, side, side, 9
    else:
        cin, nb, h, w, taps = kb * 64, 1, 1, m, 1
    x = torch.randn(m, cin, device=dev).to(torch.bfloat16)
    wt = pk.to_bf16(pk.tile_pack(torch.randn(n, kb * 64, device=dev) * 0.02))
    n_out = n // 2 if geglu else n
    out = torch.empty(m, n_out, device=dev, dtype=torch.bfloat16)
    bias = torch.randn(n, device=dev)
    p = nat.make_igemm_params([x], [cin], nb, h, w, [(0, taps)], wt, n, out, n_out, bias=bias, block_n=bn,
                              split_k=split, workspace=ws, counters=cnt, weight_tiled=True, pair=pair, pdl=True,
                              act=nat.ACT_GEGLU if geglu else nat.ACT_NONE)
    old = nat.set_pdl(True)
    try:
        nat.igemm(p)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(iters):
                nat.igemm(p)
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        best = 1e9
        for _ in range(2):
            e0.record()
            g.replay()
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) * 1e3 / iters)
    finally:
        nat.set_pdl(old)
    return best


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--log", required=True)
    ap.add_argument("--out", default=os.path.join(ROOT, "latent-diffusion-segmentation_b200", "ldmseg", "engine",
                                                  "tuned_b200.json"))
    ap.add_argument("--min-gain", type=float, default=0.04, help="keep an entry only if it beats the model's pick by this")
    ap.add_argument("--merge", default=None, help="existing table whose entries are kept (other plans' shapes)")
    args = ap.parse_args()
    torch.cuda.set_device(0)
    ws = torch.zeros(16 * 1024 * 1024, device="cuda")
    cnt = torch.zeros(8192, device="cuda", dtype=torch.int32)
    table = {}
    if args.merge and os.path.exists(args.merge):
        table = json.load(open(args.merge)).get("entries", {})
    total_model = total_best = 0.0
    for (m, n, kb), name in sorted(shapes_from_log(args.log).items()):
        if f"{m},{n},{kb}" in table:
            continue
        geglu = name.endswith("ff1")
        m_tiles = (m + 127) // 128
        cands = []
        for bn in (64, 128, 160, 256):
            tiles = m_tiles * ((n + bn - 1) // bn)
            for s in (1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 12, 14, 16):
                if s > 1 and (tiles * s > SMS or kb // s < 2 or tiles * s * 128 * bn > ws.numel()):
                    continue
                if geglu and n % 32:
                    continue
                cands.append((bn, s, False))
                if bn != 64 and m_tiles >= 2 and m_tiles % 2 == 0 and (s == 1 or tiles * s <= SMS):
                    cands.append((bn, s, True))
        model = planmod.choose_tiling(m, n, kb, allow_pair=True, use_tuned=False)
        res = {}
        for c in cands:
            try:
                res[c] = time_case(m, n, kb, c[0], c[1], c[2], geglu, ws, cnt)
            except RuntimeError as e:  # noqa: PERF203
                print(f"  {c}: {str(e)[:80]}")
        if tuple(model) not in res:
            res[tuple(model)] = time_case(m, n, kb, model[0], model[1], model[2], geglu, ws, cnt)
        best = min(res, key=res.get)
        t_model, t_best = res[tuple(model)], res[best]
        total_model += t_model
        total_best += t_best
        keep = t_best < t_model * (1.0 - args.min_gain)
        print(f"M={m:5d} N={n:5d} kb={kb:3d} {name:22s} model {tuple(model)} {t_model:6.1f} us | best {best} {t_best:6.1f} us"
              f"{'  <- tuned' if keep else ''}", flush=True)
        if keep:
            table[f"{m},{n},{kb}"] = [best[0], best[1], int(best[2]), round(t_best, 2), round(t_model, 2)]
    print(f"sum over distinct shapes: model {total_model:.1f} us, best {total_best:.1f} us")
    with open(args.out, "w") as f:
        json.dump({"device": torch.cuda.get_device_name(0), "source": "op tags of profiles/r01_ablate_unet_b{1,8}_v15.log",
                   "columns": ["block_n", "split_k", "pair", "us_tuned", "us_model"], "entries": table}, f, indent=1)
    print(f"wrote {len(table)} entries to {args.out}")


if __name__ == "__main__":
    main()
