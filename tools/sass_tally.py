#!/usr/bin/env python
"""Static SASS view of one kernel in a cubin/.so: opcode tally, FP64 / spill density per 200 instructions, and the
instruction mix of the hottest loop (largest backward-branch span that contains DFMA).  Usage:
  sass_tally.py file.cubin kernel-substring [--dump out.sass]"""
import re, subprocess, sys
from collections import Counter

def load(path, pat):
    txt = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True, check=True).stdout
    on, rows = False, []
    for line in txt.splitlines():
        if "Function :" in line:
            on = pat in line
            continue
        m = re.match(r"\s+/\*([0-9a-f]+)\*/\s+(.*?)\s*;?\s*/\*", line)
        if on and m:
            rows.append((int(m.group(1), 16), m.group(2).strip()))
    return rows

def opcode(ins):
    t = ins.split()
    op = t[1] if t[0].startswith("@") else t[0]
    return op.split(".")[0]

def main():
    path, pat = sys.argv[1], sys.argv[2]
    rows = load(path, pat)
    if "--dump" in sys.argv:
        open(sys.argv[sys.argv.index("--dump") + 1], "w").write("\n".join(f"{a:05x} {i}" for a, i in rows))
    print(f"{len(rows)} instructions")
    tally = Counter(opcode(i) for _, i in rows)
    print("; ".join(f"{o} {c}" for o, c in tally.most_common(28)))
    addr_index = {a: n for n, (a, _) in enumerate(rows)}
    loops = []
    for n, (a, ins) in enumerate(rows):
        m = re.search(r"BRA(?:\.\w+)*\s+(?:U?P\d,\s*|!?U?P\d+,\s*)*0x([0-9a-f]+)", ins)
        if m and "BRA" in opcode(ins):
            tgt = int(m.group(1), 16)
            if tgt <= a and tgt in addr_index:
                loops.append((addr_index[tgt], n))
    fp = {"DFMA", "DMUL", "DADD", "DSETP"}
    best = None
    for lo, hi in loops:
        body = [opcode(i) for _, i in rows[lo:hi + 1]]
        nfp = sum(b in fp for b in body)
        if best is None or nfp > best[0]:
            best = (nfp, lo, hi, Counter(body))
    if best:
        nfp, lo, hi, c = best
        print(f"hottest loop: instructions {lo}..{hi} ({hi - lo + 1}), FP64 {nfp}")
        print("  " + "; ".join(f"{o} {n}" for o, n in c.most_common(24)))
    # all loops with DFMA, innermost first
    shown = 0
    for lo, hi in sorted(loops, key=lambda x: x[1] - x[0]):
        body = Counter(opcode(i) for _, i in rows[lo:hi + 1])
        nfp = sum(body[b] for b in fp)
        if nfp >= 150 and shown < 4:
            shown += 1
            print(f"  loop {lo}..{hi} len {hi-lo+1}: FP64 {nfp} SHFL {body['SHFL']} LDS {body['LDS']} STS {body['STS']} LDL {body['LDL']} STL {body['STL']} "
                  f"MOV {body['MOV']+body['IMAD']} MUFU {body['MUFU']} VOTE {body['VOTE']} BRA {body['BRA']}")

main()
