#!/usr/bin/env python
"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list.  usage: launch_summary.py launches.csv [header text]"""
import csv
import re
import sys
from collections import OrderedDict

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
h = rows[hi]
name, val, unit = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
agg = OrderedDict()
tot = 0.0
for r in rows[hi + 1:]:
    if len(r) <= val:
        continue
    v = float(r[val].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[unit], 1e-6)
    a = agg.setdefault(r[name], [0, 0.0])
    a[0] += 1
    a[1] += v
    tot += v
if len(sys.argv) > 2:
    print(sys.argv[2])
    print()
for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
    print(f"{re.sub(r'^void ', '', k)[:100]:100s} n={n:5d} {v:10.3f} ms {100 * v / tot:5.1f}%")
