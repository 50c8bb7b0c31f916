#!/usr/bin/env python
"""Benchmark of the morph-render hot path (BASELINE.json metric: morphed frames/sec; HBM GB/s vs peak).

    python bench.py --gpus N --steps K --warmup W            # the CUDA path, one process per GPU (torchrun for N>1)
    python bench.py --impl reference --gpus N --steps K ...   # the reference's own CPU morph_images on host cores

A *step* is one pass of the hot path over one batch of synthetic input: `--frames` independent frame phases
(direct mode) of the named workload, rendered into the HBM frame ring with the image pair, the point sets and the
per-frame triangle lists prepared beforehand. Default workload: BASELINE.json configs[3] — synthetic 3840x2160
pair, 20k matched points (~40k triangles), 6-level pyramid; 600 phases per step and rank.

JSON line (rank 0):
  value       whole-job frames/s with inputs resident in HBM, device time (CUDA events on the context's stream),
              max over ranks
  e2e         frames/s through the public API call (poppy_b200.api.render_phases semantics: host topology planning
              on host threads + H2D of the pair from pinned memory + render + D2H of every frame to pinned memory)
  roofline    whole-path algorithmic bytes (23 B x P0 + 132 B x sum P1..P(L-1) + 104 B x PL
              per frame) x frames/s against the measured HBM copy peak, plus the per-kernel-class time shares
  cpu_baseline  the reference library (oracle/_ref, unmodified reference code) timed on a bounded sample of the same
              workload on the host cores of rank 0, with the parity of those frames against the GPU frames
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


# ------------------------------------------------------------------------------------------------------------------
def algorithmic_bytes_per_frame(w: int, h: int, levels: int) -> int:
    """SURVEY.md section 8(d) / BASELINE.md section 3."""
    px = []
    lw, lh = w, h
    for _ in range(levels + 1):
        px.append(lw * lh)
        lw, lh = (lw + 1) // 2, (lh + 1) // 2
    return 23 * px[0] + 132 * sum(px[1:levels]) + 104 * px[levels]


def ncu_traffic():
    """DRAM bytes per frame of every kernel class from the committed `ncu --set full` capture
    (profiles/ncu_traffic.json, written by tools/ncu_summary.py traffic): dram__bytes_read.sum + dram__bytes_write.sum."""
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        return json.load(open(path))
    except Exception:
        return None


def dominant_kernel_roofline(kern, W, H, F, chunk_frames, peak, traffic, levels=6):
    """roofline of the kernel class with the largest share of the step, per launch: algorithmic bytes of that stage
     / its CUDA-event launch time measured in this run."""
    if not kern:
        return None
    P, lw, lh = [], W, H
    for _ in range(levels + 1):
        P.append(lw * lh)
        lw, lh = (lw + 1) // 2, (lh + 1) // 2
    px, mid = P[0], sum(P[1:levels])
    # compulsory bytes per frame of each kernel class as designed: every input read once, every output written once
    # 
    alg = {
        "raster_warp": 8 * px + 8 * px / max(chunk_frames, 1),     # warped pair written; both sources read once per chunk
        "unsharp_store": 15 * px,                                   # lapBlend 32FC3 read, 8UC3 frame written
        # level 0: warped pair 8 + mask 4 + coarse (24 + 12)/4 read, out 12 written; level k: 28 + 9 read, 12 written;
        # coarsest: 28 read, 12 written
        "blend_collapse": 33 * px + 49 * mid + 40 * P[levels],
        # level 0->1: warped 8 + basis 4 read, mask0 4 + 7 planes/4 written; level k->k+1: 28 read, 7 written
        "pyr_down": 23 * px + 35 * mid,
    }
    k = kern[0]
    name = k["kernel"]
    per_chunk = {"raster_warp": 1, "unsharp_store": 1, "blend_collapse": levels + 1, "pyr_down": levels}
    frames_per_launch = F * per_chunk[name] / k["launches_per_step"] if name in per_chunk else None
    if name in ("blend_collapse", "pyr_down") and frames_per_launch:
        # the class is a sequence of per-level launches: rate it over the whole sequence of one chunk
        k = dict(k, us_per_launch=k["us_per_launch"] * per_chunk[name])
        out_note = f"{per_chunk[name]} per-level launches of one chunk rated together"
    else:
        out_note = None
    out = {"kernel": name, "us_per_launch": k["us_per_launch"], "share_of_step": k["share"]}
    if out_note:
        out["note"] = out_note
    if name in alg and frames_per_launch:
        b = alg[name] * frames_per_launch
        ach = b / (k["us_per_launch"] * 1e-6) / 1e9
        out.update({"algorithmic_bytes_per_launch": int(b), "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak})
    if traffic and name in traffic.get("bytes_per_frame", {}) and frames_per_launch:
        out["traffic"] = int(traffic["bytes_per_frame"][name] * frames_per_launch)
        out["traffic_source"] = traffic.get("source")
    return out


def issue_roofline(traffic, fps_per_gpu, clocks, sms=148, slots_per_sm=4):
    """The other roofline of this path: warp instructions per frame (ncu smsp__inst_executed.sum of every kernel,
    profiles/ncu_traffic.json) against the issue slots of the GPU at the SM clock sampled during the run."""
    if not traffic or "warp_instructions_per_frame" not in traffic:
        return None
    wi = float(sum(traffic["warp_instructions_per_frame"].values()))
    mhz = float((clocks or {}).get("sm_mhz") or 1965.0)
    peak = sms * slots_per_sm * mhz * 1e6
    return {"warp_instructions_per_frame": int(wi), "peak_warp_instructions_per_s": peak,
            "roofline_frames_per_s": peak / wi, "frac": fps_per_gpu * wi / peak}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows = []
        self.proc = None
        self.gpu = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.gpu)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except Exception:
                continue
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7), ("sw_power_cap", 8)):
                if len(r) > col and r[col].lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------------------------
def reference_arm(args, wl, rank, world):
    """The reference's own CPU implementation of the path (oracle/_ref/libpoppy_ref.so = unmodified reference
    sources + vendored OpenCV 4.6.0), all host threads, one frame phase per step (a bounded sample)."""
    if rank != 0:
        return
    from oracle import ref
    from poppy_b200 import synth, shard
    if not ref.available():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libpoppy_ref.so not built"}))
        return
    inp = synth.make_inputs(wl["w"], wl["h"], wl["n_points"], wl["jitter"], wl["seed"])
    sched = shard.phase_schedule(wl["frames"])
    total = args.warmup + args.steps
    picks = [int(round(i * (len(sched) - 1) / max(total - 1, 1))) for i in range(total)]
    times = []
    for i, k in enumerate(picks):
        s = float(sched[k])
        t0 = time.perf_counter()
        ref.morph_images(inp.bgr1, inp.bgr2, inp.gabor2, inp.pts1, inp.pts2, s, s, wl["levels"])
        dt = time.perf_counter() - t0
        if i >= args.warmup:
            times.append(dt)
    fps = len(times) / sum(times)
    cores = ref.get_threads()
    sample = f"{len(times)} frames of {wl['name']} (phases spread over [0,1]), 1 frame per step, after {args.warmup} warm-up frames"
    print(json.dumps({
        "impl": "reference", "metric": "morphed frames/sec", "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * sum(times) / len(times),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8/f32", "data": "synthetic",
        "config": {"workload": wl["desc"], "frames_per_step": 1, "mode": "direct"},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": "reference", "sample": sample},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


# ------------------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--workload", default="4k", choices=["1080p", "4k", "8k"])
    ap.add_argument("--frames", type=int, default=0, help="frame phases per step and rank (default: workload's)")
    ap.add_argument("--chunk", type=int, default=0, help="frames per kernel batch (0 = library default)")
    ap.add_argument("--cpu-frames", type=int, default=4, help="reference frames timed for cpu_baseline (0 = skip)")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--e2e-slice", type=int, default=64, help="frames per planned/rendered/downloaded slice of the e2e pipeline")
    ap.add_argument("--no-stage-pass", action="store_true",
                    help="profiling runs (under ncu): skip the per-kernel-class timing pass and the e2e leg")
    args = ap.parse_args()

    from poppy_b200 import synth, shard
    wl = dict(synth.WORKLOADS[args.workload])
    wl["name"] = args.workload
    if args.frames:
        wl["frames"] = args.frames
    if args.workload == "8k" and not args.frames:
        wl["frames"] = 300            # 2400 8K frames (239 GB) do not fit one GPU's ring; 300 per rank do
    wl["desc"] = (f"synthetic {wl['w']}x{wl['h']} BGR pair, {wl['n_points']}+4 matched points, {wl['levels']}-level "
                  f"pyramid, {wl['frames']} independent phases per rank (BASELINE.json configs[3] shape)")

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        reference_arm(args, wl, rank, world)
        return

    import torch
    import torch.distributed as dist
    from poppy_b200 import build, host
    from poppy_b200.renderer import MorphRenderer

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the morph renderer has no CPU path")
    if rank == 0:
        build.build()
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        dist.barrier()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- inputs (identical on every rank), this rank's phase range, host topology plan ---------------------------
    F = wl["frames"]
    W, H, L = wl["w"], wl["h"], wl["levels"]
    inp = synth.make_inputs(W, H, wl["n_points"], wl["jitter"], wl["seed"])
    sched_all = shard.phase_schedule(F * world)
    lo, hi = shard.phase_range(rank, world, F * world)
    phases = np.ascontiguousarray(sched_all[lo:hi])
    masks = phases.astype(np.float64)
    ncores = os.cpu_count() or 1
    plan_threads = max(1, ncores // world)
    t0 = time.perf_counter()
    plan = host.SequencePlan(inp.pts1, inp.pts2, W, H, phases, chain=False, threads=plan_threads)
    plan_s = time.perf_counter() - t0

    r = MorphRenderer(W, H, L, len(inp.pts1), plan.max_triangles, F, device=local_rank,
                      chunk_frames=args.chunk or None)
    # pinned host copies of the pair (source of the e2e H2D)
    pin = lambda a: torch.from_numpy(a).pin_memory()
    h_bgr1, h_bgr2, h_gab = pin(inp.bgr1), pin(inp.bgr2), pin(inp.gabor2)
    r.set_pair(h_bgr1.numpy(), h_bgr2.numpy(), h_gab.numpy())
    r.set_points(inp.pts1, inp.pts2)
    stream = torch.cuda.ExternalStream(r.stream(), device=torch.device("cuda", local_rank))

    def step():
        r.render(phases, masks, plan.tri_idx, plan.tri_offsets, chain=False)

    for _ in range(args.warmup):
        step()
    r.sync()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = r.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    e1.synchronize()
    barrier()
    dev_ms = e0.elapsed_time(e1)
    launches = r.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    checksum = r.checksum(0, F)

    # ---- per-kernel-class shares: one more step with CUDA-event stage timing -------------------------------------
    stage, stage_total_ms = {}, None
    if not args.no_stage_pass:
        r._check(r._lib.poppy_cuda_set_stage_timing(r._ctx, 1))
        step()
        stage = r.stage_times()
        stage_total_ms = r.last_render_ms()
        r._check(r._lib.poppy_cuda_set_stage_timing(r._ctx, 0))

    # ---- e2e: the public-API path with host buffers --------------------------------------------------------------
    ring_frames = min(F, max(64, args.e2e_slice))        # pinned host ring: at least one slice
    h_ring = torch.empty((ring_frames, H, W, 3), dtype=torch.uint8).pin_memory()
    frame_bytes = H * W * 3
    e2e_times, e2e_parts = [], None
    barrier()
    # The sequence streams through the renderer slice by slice: a planner thread triangulates slice k+1 on the host
    # cores while slice k renders and slice k-1 is copied to pinned memory (copy stream) - the reference's frame loop
    # (src/poppy.hpp:172-243) with its three stages overlapped instead of run back to back.
    import queue
    import threading
    slice_frames = max(1, min(F, args.e2e_slice))
    slices = [(a, min(a + slice_frames, F)) for a in range(0, F, slice_frames)]
    for it in range(0 if args.no_stage_pass else max(args.e2e_steps, 1) + 1):
        plans = queue.Queue(maxsize=3)
        plan_busy = [0.0]

        def planner():
            for a, b in slices:
                t0 = time.perf_counter()
                p = host.SequencePlan(inp.pts1, inp.pts2, W, H, phases[a:b], chain=False, threads=plan_threads)
                plan_busy[0] += time.perf_counter() - t0
                plans.put((a, b, p))

        t_a = time.perf_counter()
        th = threading.Thread(target=planner, daemon=True)
        th.start()
        r.set_pair(h_bgr1.numpy(), h_bgr2.numpy(), h_gab.numpy())
        r.set_points(inp.pts1, inp.pts2)
        t_c = time.perf_counter()
        for _ in slices:
            a, b, p = plans.get()
            r.render(phases[a:b], masks[a:b], p.tri_idx, p.tri_offsets, chain=False, first_slot=a)
            slot = a % ring_frames
            if slot + (b - a) > ring_frames:
                slot = 0
            r.download_async(a, b - a, h_ring.data_ptr() + slot * frame_bytes, W * 3, frame_bytes)
            p.close()
        r.sync()
        t_d = time.perf_counter()
        th.join()
        if it > 0:        # first pass is warm-up
            e2e_times.append(t_d - t_a)
            e2e_parts = {"total_s": t_d - t_a, "h2d_s": t_c - t_a, "planner_busy_s": plan_busy[0], "plan_threads": plan_threads,
                         "slice_frames": slice_frames, "slices": len(slices)}
    e2e_s = statistics.median(e2e_times) if e2e_times else float("inf")

    # ---- the drop-in call itself: one morph_images() per frame, exactly the reference's signature and calling pattern
    # (src/poppy.hpp:215: images, points and ratios in, dst and morphedPoints out; every call uploads the pair,
    # triangulates on ONE host thread, renders one frame and downloads it)
    single_call = None
    if rank == 0 and world == 1 and not args.no_stage_pass and args.e2e_steps > 0:
        from poppy_b200 import api
        api.Settings.instance().pyramid_levels = L
        ts = []
        for k in range(5):
            sk = float(phases[(k * 131) % F])
            t0 = time.perf_counter()
            api.morph_images(inp.bgr1, inp.bgr2, inp.bgr1, inp.bgr2, inp.gabor2, inp.pts1, inp.pts2, sk, sk)
            ts.append(time.perf_counter() - t0)
        api.release()
        single_call = {"value": 1.0 / statistics.median(ts[1:]), "unit": "frames/s",
                       "what": "poppy_b200.api.morph_images() called once per frame (pair H2D + single-thread Delaunay + "
                               "render + D2H per call), the reference's own calling pattern"}
    h2d_bytes = inp.bgr1.nbytes + inp.bgr2.nbytes + inp.gabor2.nbytes + inp.pts1.nbytes + inp.pts2.nbytes + \
        plan.tri_idx.nbytes + plan.tri_offsets.nbytes + phases.nbytes + masks.nbytes
    d2h_bytes = F * frame_bytes

    # ---- reduce over ranks ---------------------------------------------------------------------------------------
    if world > 1:
        t = torch.tensor([dev_ms, e2e_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms, e2e_s = float(t[0]), float(t[1])
        l = torch.tensor([launches], dtype=torch.int64, device="cuda")
        dist.all_reduce(l, op=dist.ReduceOp.SUM)
        launches = int(l[0])
        sums = [torch.zeros(1, dtype=torch.int64, device="cuda") for _ in range(world)]
        dist.all_gather(sums, torch.tensor([checksum & 0x7FFFFFFFFFFFFFFF], dtype=torch.int64, device="cuda"))
        checksum_all = shard.combine_checksums([int(s[0]) for s in sums])
    else:
        checksum_all = shard.combine_checksums([checksum & 0x7FFFFFFFFFFFFFFF])

    if rank == 0:
        total_frames = F * world * args.steps
        fps = total_frames / (dev_ms / 1000.0)
        alg = algorithmic_bytes_per_frame(W, H, L)
        peak, peak_src = measured_peaks()
        traffic = ncu_traffic()
        chunk_used = args.chunk or 32
        achieved = alg * (fps / world) / 1e9            # per GPU
        kern = []
        tot = sum(v["ms"] for v in stage.values()) or 1.0
        for name, v in sorted(stage.items(), key=lambda kv: -kv[1]["ms"]):
            if v["launches"]:
                kern.append({"kernel": name, "ms_per_step": round(v["ms"], 3), "share": round(v["ms"] / tot, 4),
                             "launches_per_step": v["launches"], "us_per_launch": round(1000 * v["ms"] / v["launches"], 2)})
        out = {
            "metric": "morphed frames/sec", "value": fps, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8/f32", "data": "synthetic",
            "config": {"workload": wl["desc"], "mode": "direct (independent phases, reference '-f 1 -p s')",
                       "frames_per_step_per_gpu": F, "parallelism": f"phase-sharded x{world}, no collective",
                       "l2_policy": "inputs larger than L2: each step streams >100 GB through HBM scratch + frame ring",
                       "timing": "CUDA events on the renderer's stream, max over ranks"},
            "e2e": {"value": F * world / e2e_s, "unit": "frames/s", "h2d_bytes_per_step": int(h2d_bytes),
                    "d2h_bytes_per_step": int(d2h_bytes), "breakdown": e2e_parts, "single_call": single_call,
                    "what": "H2D pair/points + per slice: host Delaunay planning (threads), H2D triangles, render, D2H of every frame to "
                            "pinned memory; the three stages of consecutive slices overlap"},
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": (int(sum(traffic["bytes_per_frame"].values()) * F) if traffic else None),
                         "traffic_what": "ncu dram__bytes_read+write of all kernels of one step (profiles/ncu_traffic.json)",
                         "dominant_kernel": dominant_kernel_roofline(kern, W, H, F, chunk_used, peak, traffic, L),
                         "issue": issue_roofline(traffic, fps / world, clocks),
                         "peak_source": peak_src,
                         "scope": "whole render path per GPU (SURVEY.md 8(d) algorithmic bytes/frame x frames/s)",
                         "algorithmic_bytes_per_frame": alg, "kernels": kern,
                         "stage_timed_step_ms": stage_total_ms},
            "clocks": clocks,
            "host_plan_s": plan_s, "frames_checksum": f"{checksum_all:016x}",
        }
        # ---- CPU baseline + parity on the sampled frames (rank 0, N=1 only) --------------------------------------
        if world == 1 and args.cpu_frames > 0:
            from oracle import ref
            if ref.available():
                idx = [int(round(i * (F - 1) / max(args.cpu_frames - 1, 1))) for i in range(args.cpu_frames)]
                ref.morph_images(inp.bgr1, inp.bgr2, inp.gabor2, inp.pts1, inp.pts2, 0.5, 0.5, L)   # warm-up
                times, worst, diff_bytes, within1 = [], 0, 0, 1.0
                for k in idx:
                    s = float(phases[k])
                    t0 = time.perf_counter()
                    want, _ = ref.morph_images(inp.bgr1, inp.bgr2, inp.gabor2, inp.pts1, inp.pts2, s, s, L)
                    times.append(time.perf_counter() - t0)
                    got = r.download(k, 1)[0]
                    d = np.abs(got.astype(np.int16) - want.astype(np.int16))
                    worst = max(worst, int(d.max())); diff_bytes += int((d != 0).sum())
                    within1 = min(within1, float((d <= 1).mean()))
                out["cpu_baseline"] = {"value": len(times) / sum(times), "unit": "frames/s", "cores": ref.get_threads(),
                                       "kind": "reference",
                                       "sample": f"{len(times)} frames of the same workload (phases {[round(float(phases[k]), 3) for k in idx]}), "
                                                 "unmodified reference morph_images() on all host threads"}
                out["parity_vs_reference"] = {"frames": len(idx), "differing_bytes": diff_bytes, "max_abs": worst,
                                              "min_fraction_within_1": within1}
            else:
                out["cpu_baseline"] = {"value": None, "unit": "frames/s", "cores": 0, "kind": "reference",
                                       "sample": "oracle/_ref not built"}
        print(json.dumps(out))
    r.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
