#!/usr/bin/env python
"""Full-frame parity of many phases of a bench workload against the unmodified reference library (GPU box only):
renders `count` evenly spread phases of the 600-phase schedule on the GPU, the same phases with the reference on K
single-threaded processes, and compares every byte. usage: parity_sweep.py [workload] [count] [procs]"""
import json
import multiprocessing as mp
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np


def worker(args):
    name, phases = args
    from oracle import ref
    from poppy_b200 import synth
    ref.set_threads(1)
    c = synth.WORKLOADS[name]
    inp = synth.make_inputs(c["w"], c["h"], c["n_points"], c["jitter"], c["seed"])
    out = []
    for s in phases:
        dst, _ = ref.morph_images(inp.bgr1, inp.bgr2, inp.gabor2, inp.pts1, inp.pts2, float(s), float(s), c["levels"])
        out.append(dst)
    return out


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "4k"
    count = int(sys.argv[2]) if len(sys.argv) > 2 else 96
    procs = int(sys.argv[3]) if len(sys.argv) > 3 else min(os.cpu_count() or 1, 16)
    from poppy_b200 import host, shard, synth
    from poppy_b200.renderer import MorphRenderer
    c = synth.WORKLOADS[name]
    w, h, L = c["w"], c["h"], c["levels"]
    inp = synth.make_inputs(w, h, c["n_points"], c["jitter"], c["seed"])
    sched = shard.phase_schedule(600)
    phases = np.ascontiguousarray(sched[np.linspace(0, 599, count).round().astype(int)])
    t0 = time.perf_counter()
    with mp.get_context("spawn").Pool(procs) as pool:
        parts = pool.map_async(worker, [(name, phases[i::procs]) for i in range(procs)])
        plan = host.SequencePlan(inp.pts1, inp.pts2, w, h, phases, threads=2)
        with MorphRenderer(w, h, L, len(inp.pts1), plan.max_triangles, count) as r:
            r.set_pair(inp.bgr1, inp.bgr2, inp.gabor2)
            r.set_points(inp.pts1, inp.pts2)
            r.render(phases, phases.astype(np.float64), plan.tri_idx, plan.tri_offsets)
            got = r.download(0, count)
        parts = parts.get()
    differing, worst = 0, 0
    for i in range(procs):
        for k, want in enumerate(parts[i]):
            d = np.abs(got[i + k * procs].astype(np.int16) - want.astype(np.int16))
            differing += int((d != 0).sum())
            worst = max(worst, int(d.max()))
    print(json.dumps({"workload": name, "frames_compared": count, "phases": [round(float(p), 4) for p in phases[:3]] + ["..."] +
                      [round(float(phases[-1]), 4)], "bytes_compared": int(got.size), "differing_bytes": differing, "max_abs": worst,
                      "reference_processes": procs, "seconds": round(time.perf_counter() - t0, 1)}))


if __name__ == "__main__":
    main()
