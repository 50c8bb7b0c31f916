#!/usr/bin/env python
"""Basic-block view of `ncu --page source --csv` (SASS): consecutive instructions with the same executed count are
merged; prints count, #instr, share and an opcode summary. usage: sass_blocks.py file.csv [min_share]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
isrc, iex = hdr.index("Source"), hdr.index("Instructions Executed")
minshare = float(sys.argv[2]) if len(sys.argv) > 2 else 0.004
ins = []
for r in rows[2:]:
    if len(r) <= iex: continue
    toks = r[isrc].split()
    op = toks[1] if toks and toks[0].startswith("@") and len(toks) > 1 else (toks[0] if toks else "?")
    ins.append((int(r[iex] or 0), op.split(".")[0], r[isrc].strip()))
tot = sum(i[0] for i in ins)
blocks = []
for i, (ex, op, src) in enumerate(ins):
    if blocks and blocks[-1][0] == ex: blocks[-1][2].append(op); blocks[-1][3] = i
    else: blocks.append([ex, i, [op], i])
print("total", tot)
for ex, start, ops, end in blocks:
    share = ex * len(ops) / tot
    if share >= minshare:
        c = collections.Counter(ops)
        print(f"[{start:5d}-{end:5d}] x{ex:10d} n={len(ops):4d} share={share:6.3f}  " + " ".join(f"{k}:{v}" for k, v in c.most_common(12)))
