"""Summarise an `ncu --set full` capture (.ncu-rep) per launch: duration, tensor-pipe activity, DRAM traffic and
rates, L2 traffic, registers.  Usage: python tools/ncu_summary.py file.ncu-rep [more.ncu-rep ...]"""
import csv
import io
import subprocess
import sys

COLS = [
    ("gpu__time_duration.sum", "dur_us"),
    ("sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pct"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
    ("dram__bytes_read.sum", "dram_rd"),
    ("dram__bytes_write.sum", "dram_wr"),
    ("lts__t_bytes.sum", "l2_bytes"),
    ("sm__cycles_elapsed.max", "cycles"),
    ("launch__registers_per_thread", "regs"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_pct"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2_pct"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "mufu_pct"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "fma_pct"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_pct"),
]


def main():
    for path in sys.argv[1:]:
        out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        hdr, units = rows[0], rows[1]
        col = {h: i for i, h in enumerate(hdr)}
        print(f"== {path}")
        print("kernel | grid | " + " | ".join(n for _, n in COLS))
        for r in rows[2:]:
            name = r[col["Kernel Name"]].split("(")[0]
            vals = []
            for m, _ in COLS:
                if m in col:
                    vals.append(f"{r[col[m]]} {units[col[m]]}".strip())
                else:
                    vals.append("-")
            print(f"{name} | {r[col['Grid Size']]} | " + " | ".join(vals))
        # top warp-stall reasons, first launch of every kernel name
        seen = set()
        stall_cols = [h for h in hdr if "issue_stalled" in h and h.endswith("_per_warp_active.pct")]
        for r in rows[2:]:
            name = r[col["Kernel Name"]].split("(")[0]
            if name in seen or not stall_cols:
                continue
            seen.add(name)
            vals = []
            for h in stall_cols:
                try:
                    vals.append((float(r[col[h]].replace(",", "")), h.split("issue_stalled_")[1].split("_per_warp")[0]))
                except ValueError:
                    pass
            vals.sort(reverse=True)
            print(f"  stalls {name}: " + ", ".join(f"{n} {v:.0f}%" for v, n in vals[:7]))


if __name__ == "__main__":
    main()
