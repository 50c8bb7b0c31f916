#!/usr/bin/env python
"""Per-kernel SASS opcode census of poppy_b200/libpoppy_cuda.so: the instructions that show which Blackwell features each
kernel uses (B200_PROFILING.md: TMA = UTMALDG / UBLKCP, mbarrier = SYNCS, texture gather = TLD4, cp.async = LDGSTS, packed
fp32 = FFMA2 / FADD2 / FMUL2, SIMD video = VABSDIFF4 / VIMNMX).
usage: python tools/sass_census.py [lib.so] > profiles/r2_sass_census.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "poppy_b200", "libpoppy_cuda.so")
WATCH = ["UTMALDG", "UBLKCP", "SYNCS", "TLD4", "LDGSTS", "FFMA2", "FADD2", "FMUL2", "VABSDIFF4", "VIMNMX", "VIMNMX3", "FFMA", "FADD", "FMUL",
         "DFMA", "DADD", "DMUL", "ATOMS", "SHFL", "BAR", "LDS", "STS", "LDG", "STG"]
sass = subprocess.run(["cuobjdump", "-sass", lib], stdout=subprocess.PIPE, text=True, check=True).stdout
demangle = lambda n: subprocess.run(["cu++filt", n], stdout=subprocess.PIPE, text=True).stdout.strip() or n
arch = re.findall(r"arch = (sm_\w+)", sass)
cur, counts, total = None, collections.OrderedDict(), {}
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        total[cur] = 0
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\w+\s+)?([A-Z][A-Z0-9_]*)", line)
    if m and cur:
        counts[cur][m.group(1)] += 1
        total[cur] += 1
print(f"# SASS opcode census of {os.path.relpath(lib, ROOT)} (static instruction counts; architectures: {sorted(set(arch))})")
print("# " + " ".join(f"{w:>9s}" for w in ["total"] + WATCH) + "  kernel")
for k, c in counts.items():
    name = demangle(k).replace("poppy::", "").replace("(anonymous namespace)::", "")
    if name.endswith(")"):                      # drop the parameter list (the parenthesis group that closes the name)
        depth = 0
        for i in range(len(name) - 1, -1, -1):
            depth += name[i] == ")"
            depth -= name[i] == "("
            if depth == 0:
                name = name[:i]
                break
    name = name.replace("(bool)", "").replace("(int)", "").replace("void ", "")
    print("  " + " ".join(f"{v:9d}" for v in [total[k]] + [c.get(w, 0) for w in WATCH]) + "  " + name)
agg = collections.Counter()
for c in counts.values():
    agg.update(c)
print("  " + " ".join(f"{v:9d}" for v in [sum(total.values())] + [agg.get(w, 0) for w in WATCH]) + "  ALL")
