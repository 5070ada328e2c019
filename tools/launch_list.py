#!/usr/bin/env python3
"""Condense an `ncu --metrics gpu__time_duration.sum --csv` log into a per-kernel launch list.

    python tools/launch_list.py gpurun_out/launches.csv > profiles/rNN_bench_launch_list.txt
"""
import csv
import sys
from collections import OrderedDict

rows = []
with open(sys.argv[1]) as fh:
    lines = [l for l in fh if l.startswith('"')]
rd = csv.reader(lines)
hdr = next(rd)
ix = {h: i for i, h in enumerate(hdr)}
agg = OrderedDict()
for r in rd:
    if len(r) <= ix["Metric Value"] or r[ix["Metric Name"]] != "gpu__time_duration.sum":
        continue
    key = (r[ix["Kernel Name"]][:60], r[ix["Grid Size"]])
    v = float(r[ix["Metric Value"]].replace(",", ""))
    unit = r[ix["Metric Unit"]]
    us = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
    a = agg.setdefault(key, [0, 0.0])
    a[0] += 1
    a[1] += us
print("%-62s %-18s %6s %10s %12s" % ("kernel", "grid", "count", "avg_us", "total_us"))
for (k, g), (n, t) in agg.items():
    print("%-62s %-18s %6d %10.1f %12.1f" % (k, g, n, t / n, t))
