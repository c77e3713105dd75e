#!/usr/bin/env python
"""Per-instruction view of an ncu `--page source --csv --print-source sass` export: where the stall samples sit.
Usage: ncu_hot.py src.csv [lo hi]   (lo/hi: instruction index range to list in full)"""
import csv, sys
from collections import Counter
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
data = rows[2:]
col = {h: i for i, h in enumerate(hdr)}
S, X = col["# Samples"], col["Instructions Executed"]
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[S] or 0) for r in data)
print("instructions", len(data), "samples", tot)
agg = Counter()
for r in data:
    for h in stall_cols:
        agg[h] += int(r[col[h]] or 0)
print("stall totals:", ", ".join(f"{h[6:]} {100*v/tot:.1f}%" for h, v in agg.most_common(8)))
# blocks of 50 instructions
print("samples by block of 50 instructions (>=1%):")
for b in range(0, len(data), 50):
    s = sum(int(r[S] or 0) for r in data[b:b + 50])
    ex = max(int(r[X] or 0) for r in data[b:b + 50])
    if s >= 0.01 * tot:
        print(f"  {b:6d}: {100*s/tot:5.1f}%  max-exec {ex:.3g}")
if len(sys.argv) >= 4:
    lo, hi = int(sys.argv[2]), int(sys.argv[3])
    for n, r in enumerate(data[lo:hi], lo):
        s = int(r[S] or 0)
        top = sorted(((int(r[col[h]] or 0), h[6:]) for h in stall_cols), reverse=True)[:2]
        print(f"{n:6d} {100*s/tot:5.2f}% {int(r[X] or 0):>11d} {r[col['Source']][:70]:70s} {top[0][1]}:{top[0][0]} {top[1][1]}:{top[1][0]}")
