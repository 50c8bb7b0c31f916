#!/usr/bin/env python
"""Hot-spot view of an `ncu --page source --csv` export (SASS view): opcode histogram weighted by executed
instructions and the top SASS lines. usage: sass_hot.py file.csv [top_n]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ia, isrc, iex, ismp = hdr.index("Address"), hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
ops = collections.Counter(); tot = 0; lines = []
for n, r in enumerate(rows[2:]):
    if len(r) <= iex: continue
    ex = int(r[iex] or 0); tot += ex
    src = r[isrc].strip()
    toks = src.split()
    op = toks[1] if toks and toks[0].startswith("@") and len(toks) > 1 else (toks[0] if toks else "?")
    ops[op.split(".")[0]] += ex
    lines.append((ex, n, int(r[ismp] or 0), src))
print("total warp instructions", tot)
for op, c in ops.most_common(25):
    print(f"{op:12s} {c:14d} {c/tot:6.3f}")
top = int(sys.argv[2]) if len(sys.argv) > 2 else 0
if top:
    print("--- SASS in program order with exec count (only lines >= 0.2% )")
    for ex, n, smp, src in lines:
        if ex / tot >= 0.002: print(f"{n:5d} {ex:12d} {smp:6d}  {src}")
