#!/usr/bin/env python
"""Render the same batch repeatedly and compare per-frame checksums run against run (GPU only).
usage: python tools/stress_determinism.py [frames] [runs] [workload]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from poppy_b200 import host, shard, synth
from poppy_b200.renderer import MorphRenderer


F = int(sys.argv[1]) if len(sys.argv) > 1 else 96
runs = int(sys.argv[2]) if len(sys.argv) > 2 else 8
wl = dict(synth.WORKLOADS[sys.argv[3] if len(sys.argv) > 3 else "4k"])
W, H, L = wl["w"], wl["h"], wl["levels"]
inp = synth.make_inputs(W, H, wl["n_points"], wl["jitter"], wl["seed"])
phases = np.ascontiguousarray(shard.phase_schedule(F))
plan = host.SequencePlan(inp.pts1, inp.pts2, W, H, phases, chain=False, threads=os.cpu_count() or 1)
r = MorphRenderer(W, H, L, len(inp.pts1), plan.max_triangles, F)
r.set_pair(inp.bgr1, inp.bgr2, inp.gabor2)
r.set_points(inp.pts1, inp.pts2)
ref_sums, ref_frames = None, None
bad = 0
for run in range(runs):
    r.render(phases, phases.astype(np.float64), plan.tri_idx, plan.tri_offsets, chain=False)
    r.sync()
    sums = [r.checksum(i, 1) for i in range(F)]
    if ref_sums is None:
        ref_sums, ref_frames = sums, r.download(0, F)
        continue
    diff = [i for i in range(F) if sums[i] != ref_sums[i]]
    for i in diff:
        a = r.download(i, 1)[0]
        d = np.argwhere((a != ref_frames[i]).any(axis=2))
        bad += 1
        print(f"run {run} frame {i}: {len(d)} px differ, rows {d[:,0].min()}..{d[:,0].max()} cols {d[:,1].min()}..{d[:,1].max()}"
              f" max abs {np.abs(a.astype(int) - ref_frames[i].astype(int)).max()}", flush=True)
    print(f"run {run}: {len(diff)} of {F} frames differ from run 0", flush=True)
print("DETERMINISTIC" if bad == 0 else f"NONDETERMINISTIC ({bad} frame instances)")
