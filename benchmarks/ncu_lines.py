#!/usr/bin/env python
"""Development helper: per-source-line summary of an `ncu --page source --csv --print-source sass,cuda` export
(samples, instructions executed, dominant stall reasons).  usage: ncu_lines.py export.csv [top_n]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
h = None
lines = []
for r in rows:
    if r and r[0] == "Line No":
        h = r
        samp_i = h.index("# Samples")
        continue
    if h is None or len(r) < len(h):
        continue
    if r[0] == "":
        continue            # SASS rows
    if not r[samp_i].isdigit():
        continue
    lines.append(r)
ix = {n: i for i, n in enumerate(h)}
samp = ix["# Samples"]
inst = ix["Instructions Executed"]
stall_cols = [i for i, n in enumerate(h) if n.startswith("stall_") and "Not Issued" not in n]
tot = sum(int(r[samp]) for r in lines)
toti = sum(int(r[inst]) for r in lines)
print(f"total samples {tot}, warp instructions {toti}")
lines.sort(key=lambda r: -int(r[samp]))
for r in lines[:top]:
    st = sorted(((int(r[i]), h[i][6:]) for i in stall_cols), reverse=True)[:3]
    print(f"{r[0]:>5} {100 * int(r[samp]) / max(tot, 1):5.1f}% inst {100 * int(r[inst]) / max(toti, 1):5.1f}%  "
          f"{' '.join(f'{n}:{c}' for c, n in st if c)}  | {r[1].strip()[:110]}")
