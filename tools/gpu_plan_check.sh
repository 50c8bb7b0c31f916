#!/bin/bash
python -m pytest tests/test_gpu_api.py tests/test_gpu_fullsize.py -q 2>&1 | tail -3
for i in 1 2; do python bench.py --steps 5 --warmup 3 --cpu-frames 2 --kprocs 0 --e2e-steps 1 > gpurun_out/plan_check.json 2> gpurun_out/plan_check.err; python - <<'PY'
import json
d=json.loads(open("gpurun_out/plan_check.json").read().strip().splitlines()[-1])
print("value", round(d["value"],1), "frac", round(d["roofline"]["frac"],4), "e2e", round(d["e2e"]["value"],1), "parity", d.get("parity_vs_reference"), "sum", d["frames_checksum"])
PY
done
