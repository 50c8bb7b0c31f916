#!/usr/bin/env python
"""Summarise ncu output into small text files that can be committed under profiles/.

  python tools/ncu_summary.py launches gpurun_out/launches_TAG.csv   > profiles/rN_launches_TAG.txt
  python tools/ncu_summary.py full     gpurun_out/prof_TAG.ncu-rep   > profiles/rN_ncu_full_TAG.txt
"""
import csv
import io
import re
import subprocess
import sys
from collections import OrderedDict

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static",
    "launch__grid_size", "launch__block_size", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
    "smsp__inst_executed.sum", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fmalite.avg.pct_of_peak_sustained_active",
    "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_cbu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_adu.avg.pct_of_peak_sustained_active",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__thread_inst_executed_per_inst_executed.pct",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
]


def short(name: str) -> str:
    name = re.sub(r"^void\s+", "", name)
    name = re.sub(r"\((?!bool|int).*$", "", name)
    return name.replace("poppy::", "").replace("(bool)", "").replace("(int)", "")


def launches(path):
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr = rows[0]
    ki, mi, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = OrderedDict()
    for r in rows[1:]:
        if r[mi] != "gpu__time_duration.sum":
            continue
        v = float(r[vi].replace(",", ""))
        v = v / 1000.0 if r[ui] == "ns" else v
        a = agg.setdefault(short(r[ki]), [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print(f"# ncu --metrics gpu__time_duration.sum --clock-control none  ({path}); cold-cache, serialised: compare SHARES")
    print(f"{'kernel':40s} {'launches':>8s} {'total_us':>12s} {'us/launch':>10s} {'share':>7s}")
    for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:40s} {n:8d} {us:12.1f} {us / n:10.1f} {us / tot:7.3f}")
    print(f"{'TOTAL':40s} {sum(a[0] for a in agg.values()):8d} {tot:12.1f}")


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    print(f"# ncu --set full --clock-control none ({path}); per launch")
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        print(f"== {short(d['Kernel Name'])}  grid {d.get('Grid Size')} block {d.get('Block Size')}")
        for k in KEYS:
            if k in d and d[k] != "":
                print(f"   {k:85s} {d[k]:>16s} {u.get(k, '')}")


CLASS_OF = [("k_raster_warp", "raster_warp"), ("k_pyr_down", "pyr_down"), ("k_collapse_roll", "blend_collapse"),
            ("k_collapse_tma", "blend_collapse"), ("k_pyramid_tail", "blend_collapse"), ("k_calm", "calm_analysis"),
            ("k_blend_coarsest", "blend_collapse"), ("k_unsharp", "unsharp_store"), ("k_tri_geometry", "tri_geometry"),
            ("k_bin_", "bin_triangles"), ("k_lerp_points", "lerp_points")]


def traffic(path, frames):
    """JSON for profiles/ncu_traffic.json: DRAM bytes per frame of each kernel class of one chunk of `frames` frames."""
    import json
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    agg = {}
    inst = {}
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    for r in rows[2:]:
        d = dict(zip(hdr, r)); u = dict(zip(hdr, units))
        name = short(d["Kernel Name"])
        cls = next((c for k, c in CLASS_OF if k in name), None)
        if not cls:
            continue
        b = sum(float(d[k].replace(",", "")) * scale[u[k]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
        agg[cls] = agg.get(cls, 0.0) + b / frames
        inst[cls] = inst.get(cls, 0.0) + float(d["smsp__inst_executed.sum"].replace(",", "")) / frames
    print(json.dumps({"source": f"ncu --set full --clock-control none, one {frames}-frame chunk ({path})",
                      "bytes_per_frame": {k: round(v) for k, v in agg.items()},
                      "warp_instructions_per_frame": {k: round(v) for k, v in inst.items()}}, indent=1))


if __name__ == "__main__":
    if sys.argv[1] == "traffic":
        traffic(sys.argv[2], int(sys.argv[3]))
        sys.exit(0)
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
