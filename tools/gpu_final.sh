#!/bin/bash
# final validation: all GPU tests, smoke, the default bench line and the reference arm, as the driver runs them
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv,noheader
timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -6
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -2
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/final_reference.json 2> gpurun_out/final_reference.err; tail -c 600 gpurun_out/final_reference.json
T0=$(date +%s); timeout 1200 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; echo "bench wall $(( $(date +%s) - T0 )) s"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/final_bench.json").read().strip().splitlines()[-1])
print("value", d["value"], "frac", d["roofline"]["frac"], "e2e", d["e2e"]["value"], "single", (d["e2e"].get("single_call") or {}).get("value"))
print("parity", d["cpu_baseline"].get("parity_vs_reference"), "cpu", d["cpu_baseline"]["value"], "launches", d["gpu_launches"], "clocks", d["clocks"], "unsharp", d["unsharp"]["exact_chunk_share"])
print("breakdown", d["e2e"].get("breakdown"))
PY
