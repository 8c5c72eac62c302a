#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (launches, total time, share).

    python scripts/launch_list_summary.py gpurun_out/launches.csv "title" > profiles/rN_x_launches.txt
"""
import collections
import csv
import sys

UNIT = {"nsecond": 1e-3, "ns": 1e-3, "usecond": 1.0, "us": 1.0, "msecond": 1e3, "ms": 1e3, "second": 1e6, "s": 1e6}
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
h = rows[0]
agg = collections.OrderedDict()
for r in rows[1:]:
    d = dict(zip(h, r))
    if d.get("Metric Name") != "gpu__time_duration.sum":
        continue
    k = d["Kernel Name"].split("(")[0]
    t = float(d["Metric Value"].replace(",", "")) * UNIT[d["Metric Unit"]]
    a = agg.setdefault(k, [0, 0.0, 0.0])
    a[0] += 1
    a[1] += t
    a[2] = max(a[2], t)
tot = sum(a[1] for a in agg.values())
print("# %s" % (sys.argv[2] if len(sys.argv) > 2 else sys.argv[1]))
print("# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised launches: read the shares, not the absolutes)")
print("# %-38s %8s %14s %8s %12s" % ("kernel", "launches", "total us", "share", "longest us"))
for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
    print("%-40s %8d %14.1f %7.1f%% %12.1f" % (k, a[0], a[1], 100 * a[1] / tot, a[2]))
