"""Hot spots of an ncu source page (`ncu -i X.ncu-rep --page source --csv > f.csv`): per kernel launch, total
stall-reason mix and the N most-sampled SASS instructions.  Usage: python tools/ncu_source_hot.py f.csv [N]"""
import csv
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    topn = int(sys.argv[2]) if len(sys.argv) > 2 else 30
    blocks, cur = [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "hdr": None, "data": []}
            blocks.append(cur)
        elif cur is not None and cur["hdr"] is None:
            cur["hdr"] = r
        elif cur is not None and len(r) == len(cur["hdr"]):
            cur["data"].append(r)
    for k, b in enumerate(blocks):
        hdr, data = b["hdr"], b["data"]
        ci = {h: i for i, h in enumerate(hdr)}
        S, E = ci["# Samples"], ci["Instructions Executed"]
        stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
        tot = sum(int(r[S]) for r in data)
        print(f"== launch {k}: {b['name'][:60]}  SASS instrs {len(data)}  samples {tot}  "
              f"warp-instr executed {sum(int(r[E]) for r in data)}")
        agg = {h: sum(int(r[ci[h]]) for r in data) for h in stalls}
        print("   stalls:", ", ".join(f"{h[6:]} {100 * v / max(tot, 1):.1f}%" for h, v in
                                       sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
        for i, r in sorted(enumerate(data), key=lambda t: -int(t[1][S]))[:topn]:
            st = sorted(((h[6:], int(r[ci[h]])) for h in stalls), key=lambda kv: -kv[1])[:2]
            print(f"   {i:6d} smp={r[S]:>6s} exec={r[E]:>8s} {r[1].strip()[:64]:64s} {st}")


if __name__ == "__main__":
    main()
