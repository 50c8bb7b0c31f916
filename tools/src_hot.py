#!/usr/bin/env python
"""Per-CUDA-source-line executed warp instructions from `ncu --page source --csv --print-source cuda,sass`.
usage: src_hot.py file.csv [min_share]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
minshare = float(sys.argv[2]) if len(sys.argv) > 2 else 0.005
cur_file = None
out = []
hdr = None
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]; continue
    if r and r[0] == "Line No":
        hdr = r; iex = hdr.index("Instructions Executed"); ismp = hdr.index("# Samples"); continue
    if hdr and r and r[0].isdigit() and len(r) > iex and r[iex] not in ("", "-"):
        try:
            out.append((cur_file, int(r[0]), int(r[iex]), int(r[ismp] or 0), r[1].strip()))
        except ValueError:
            pass
tot = sum(o[2] for o in out); stot = sum(o[3] for o in out)
print("total warp instr", tot, "samples", stot)
for f, ln, ex, smp, src in out:
    if ex / tot >= minshare or smp / max(stot, 1) >= minshare:
        print(f"{f:22s}:{ln:4d} {ex/tot:6.3f} {smp/max(stot,1):6.3f}  {src[:110]}")
