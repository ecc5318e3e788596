"""profiles/igemm_dram_traffic.json from an ncu --set full summary (tools/ncu_summary.py output): the average
dram__bytes_read.sum + dram__bytes_write.sum per captured igemm launch.  bench.py reports it as roofline.traffic.
Usage: python tools/make_traffic_json.py profiles/r02_ncu_igemm_full_b1.txt [igemm_dram_traffic_b8.json]"""
import json
import os
import sys

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def main():
    path = sys.argv[1]
    rows = [l.split(" | ") for l in open(path) if "igemm_kernel" in l and " | " in l]
    tot, durs = [], []
    for r in rows:
        rd, wr = r[6].split(), r[7].split()
        tot.append(float(rd[0]) * UNIT[rd[1]] + float(wr[0]) * UNIT[wr[1]])
        durs.append(float(r[2].split()[0]))
    out = {"traffic": round(sum(tot) / len(tot)), "unit": "bytes per launch (dram read + write, mean of the captured launches)",
           "launches_captured": len(tot), "mean_duration_us_under_ncu": round(sum(durs) / len(durs), 2),
           "source": os.path.relpath(path, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))}
    name = sys.argv[2] if len(sys.argv) > 2 else "igemm_dram_traffic.json"   # igemm_dram_traffic_b8.json for batch 8
    dst = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", name)
    with open(dst, "w") as f:
        json.dump(out, f)
    print(out)


if __name__ == "__main__":
    main()
