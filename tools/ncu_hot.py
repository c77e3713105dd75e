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

def by_opcode(lo, hi):
    from collections import defaultdict
    agg = defaultdict(lambda: [0, 0, Counter()])
    iters = max(int(r[X] or 0) for r in data[lo:hi])
    for r in data[lo:hi]:
        t = r[col["Source"]].split()
        op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
        a = agg[op]
        a[0] += int(r[S] or 0)
        a[1] += int(r[X] or 0)
        for h in stall_cols:
            a[2][h[6:]] += int(r[col[h]] or 0)
    sel = sum(int(r[col["stall_selected"]] or 0) for r in data[lo:hi])
    cyc_per_sample = (sum(int(r[X] or 0) for r in data[lo:hi]) / iters) / sel   # issue cycles per step / selected samples
    print(f"by opcode, instructions {lo}..{hi}: (cycles per loop trip per warp, assuming 1 cycle per issued instruction)")
    for op, (s, x, st) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:16]:
        top = ", ".join(f"{k}:{v*cyc_per_sample:.0f}" for k, v in st.most_common(4))
        print(f"  {op:10s} {x/iters:7.1f} instr/trip  {s*cyc_per_sample:7.0f} cycles  ({top})")
    print(f"  total {sum(a[0] for a in agg.values())*cyc_per_sample:.0f} cycles per trip")

if len(sys.argv) >= 5 and sys.argv[4] == "ops":
    by_opcode(int(sys.argv[2]), int(sys.argv[3]))
