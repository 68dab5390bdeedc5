#!/usr/bin/env python
"""Compact per-kernel summary of an .ncu-rep (run here, no GPU needed):
   python tools/ncu_summary.py gpurun_out/prof.ncu-rep [--md profiles/out.md]"""
import csv
import io
import subprocess
import sys

COLS = [
    ("time_us", "gpu__time_duration.sum"),
    ("dram_rd_MB", "dram__bytes_read.sum"),
    ("dram_wr_MB", "dram__bytes_write.sum"),
    ("dram_pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
    ("l2_pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("l1_pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("sm_pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("issue_pct", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
    ("warps_pct", "sm__warps_active.avg.pct_of_peak_sustained_active"),
    ("l1_hit", "l1tex__t_sector_hit_rate.pct"),
    ("l2_hit", "lts__t_sector_hit_rate.pct"),
    ("regs", "launch__registers_per_thread"),
    ("grid", "launch__grid_size"),
    ("smem_KB", "launch__shared_mem_per_block_dynamic"),
    ("tensor_pct", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
]
SCALE = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3,
         "second": 1e6}


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    out = []
    for r in data:
        name = r[idx["Kernel Name"]].split("(")[0].replace("void ", "").replace("gatres::", "")
        vals = []
        for label, key in COLS:
            if key not in idx:
                vals.append("-")
                continue
            v, u = r[idx[key]], units[idx[key]]
            try:
                f = float(v.replace(",", ""))
                if label == "smem_KB":
                    f *= SCALE.get(u, 1.0) * 1e3
                else:
                    f *= SCALE.get(u, 1.0)
                vals.append(f"{f:.1f}" if abs(f) < 1e5 else f"{f:.3g}")
            except ValueError:
                vals.append(v[:10])
        out.append((name, vals))
    head = "| kernel | " + " | ".join(l for l, _ in COLS) + " |"
    sep = "|" + "---|" * (len(COLS) + 1)
    lines = [head, sep] + ["| " + n + " | " + " | ".join(v) + " |" for n, v in out]
    text = "\n".join(lines)
    print(text)
    if "--md" in sys.argv:
        path = sys.argv[sys.argv.index("--md") + 1]
        with open(path, "a") as f:
            f.write(f"\n### {rep}\n\n{text}\n")


if __name__ == "__main__":
    main()
