#!/bin/bash
# e2e A/B of the host planner: unguided vs guided point-location walks (POPPY_PLAN_GUIDE)
mkdir -p gpurun_out
for g in 0 1; do
  POPPY_PLAN_GUIDE=$g python bench.py --steps 3 --warmup 3 --cpu-frames 0 --kprocs 0 --e2e-steps 2 ${BENCH_ARGS} > gpurun_out/e2e_guide$g.json 2> gpurun_out/e2e_guide$g.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/e2e_guide$g.json").read().strip().splitlines()[-1])
print("guide=$g", "value", round(d["value"],1), "e2e", d["e2e"]["value"], "single_call", (d["e2e"].get("single_call") or {}).get("value"), "host_plan", d.get("host_plan"), "breakdown", d["e2e"].get("breakdown"))
PY
done
for g in 0 1; do POPPY_PLAN_GUIDE=$g python tools/plan_timing.py 64 16; POPPY_PLAN_GUIDE=$g python tools/plan_timing.py 16 1; done
