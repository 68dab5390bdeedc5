#!/usr/bin/env python
"""Aggregate an ncu `--metrics gpu__time_duration.sum --csv` launch list by kernel:
   python tools/launch_shares.py gpurun_out/launches.csv [skip_first_n]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr, data = rows[hi], rows[hi + 1:]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg, tot = collections.OrderedDict(), 0.0
for r in data[skip:]:
    if len(r) <= vi:
        continue
    n = r[ki].split("(")[0].replace("void ", "").replace("gatres::", "")
    v = float(r[vi].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[ui], 1e-3)
    a = agg.setdefault(n, [0, 0.0])
    a[0] += 1
    a[1] += v
    tot += v
print(f"{len(data) - skip} launches, {tot:.1f} us total")
print("| kernel | launches | us total | share | us avg |\n|---|---|---|---|---|")
for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:22]:
    print(f"| {n[:64]} | {c} | {t:.1f} | {100 * t / tot:.1f}% | {t / c:.2f} |")
