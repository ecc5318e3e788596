"""In-graph cost attribution for one UNet forward plan.

ncu launch lists are cold-cache and serialised; CUDA events around eager launches include launch gaps.  Neither
says what a kernel family costs INSIDE the captured graph (warm L2, PDL overlap).  This tool measures

  * the full graph replay time,
  * for every op family (and every latent level) the replay time of the graph WITHOUT those ops (ablation:
    results are garbage, timing is what matters), so full - ablated = what the family really costs,
  * per-op warm time: a graph holding 20 back-to-back copies of one op (includes per-launch latency, no overlap).

Usage: python tools/ablate_unet.py --batch 1 [--per-op]
"""
import argparse
import collections
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "latent-diffusion-segmentation_b200"))

import torch  # noqa: E402


def time_graph(fn, reps=10):
    from ldmseg import _native as nat
    g = torch.cuda.CUDAGraph()
    fn()
    torch.cuda.synchronize()
    with torch.cuda.graph(g):
        fn()
    for _ in range(3):
        g.replay()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(reps):
        g.replay()
    t1.record()
    torch.cuda.synchronize()
    return t0.elapsed_time(t1) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=1)
    ap.add_argument("--size", type=int, default=64)
    ap.add_argument("--per-op", action="store_true")
    ap.add_argument("--full-only", action="store_true", help="time the whole graph only (A/B of two builds)")
    args = ap.parse_args()
    from bench import build_models
    from ldmseg import _native as nat
    if os.environ.get("LDMSEG_DEBUG_FLAGS"):   # igemm timing experiments (garbage results), see IgemmKParams.debug
        nat.load().ldmseg_set_debug(int(os.environ["LDMSEG_DEBUG_FLAGS"]))
    dev = torch.device("cuda:0")
    torch.cuda.set_device(dev)
    unet, _, _, _ = build_models(dev)
    plan = unet._get_engine().plan(args.batch, args.size)
    plan.x_in.normal_()
    ops, tags = plan.ops, plan.tags

    def runner(keep):
        def f():
            old = nat.set_pdl(plan.pdl)
            try:
                for op, k in zip(ops, keep):
                    if k:
                        op()
            finally:
                nat.set_pdl(old)
        return f

    full = time_graph(runner([True] * len(ops)))
    print(f"full graph: {full * 1e3:.1f} us, {len(ops)} ops")
    if args.full_only:
        more = [time_graph(runner([True] * len(ops)), reps=20) for _ in range(3)]
        print("full graph again: " + " ".join(f"{t * 1e3:.1f}" for t in more) + " us")
        return
    fam = lambda t: t.split(":")[0]
    rows = lambda t: int(t.split(":")[1])
    groups = collections.OrderedDict()
    for f in sorted(set(fam(t) for t in tags)):
        groups[f"family {f}"] = [fam(t) != f for t in tags]
    for r in sorted(set(rows(t) for t in tags), reverse=True):
        groups[f"rows {r}"] = [rows(t) != r for t in tags]
    for f in ("igemm", "gn", "attn"):
        for r in sorted(set(rows(t) for t in tags if fam(t) == f), reverse=True):
            groups[f"{f} rows {r}"] = [not (fam(t) == f and rows(t) == r) for t in tags]
    # igemm sub-kinds by layer role
    roles = ("conv1", "conv2", "proj_in", "qkv", "to_out", "ff1", "ff2", "proj_out")
    for role in roles:
        groups[f"igemm role {role}"] = [not (fam(t) == "igemm" and t.split(":")[2].endswith(role)) for t in tags]
    for name, keep in groups.items():
        n = keep.count(False)
        if n == 0:
            continue
        t = time_graph(runner(keep))
        print(f"without {name:24s} ({n:3d} ops): {t * 1e3:8.1f} us   delta {1e3 * (full - t):8.1f} us   "
              f"per-op {1e3 * (full - t) / n:6.1f} us")
    if args.per_op:
        print("per-op warm time (20 back-to-back copies in one graph):")
        for i, (op, tag) in enumerate(zip(ops, tags)):
            def f(op=op):
                old = nat.set_pdl(plan.pdl)
                try:
                    for _ in range(20):
                        op()
                finally:
                    nat.set_pdl(old)
            t = time_graph(f, reps=5) / 20
            print(f"  {i:3d} {t * 1e3:7.1f} us  {tag}")


if __name__ == "__main__":
    main()
