#!/bin/bash
python -m pytest tests/test_gpu_api.py tests/test_gpu_dropin.py tests/test_gpu_calm.py -q 2>&1 | tail -3
python bench.py --steps 3 --warmup 3 --cpu-frames 2 --kprocs 0 --e2e-steps 2 > gpurun_out/api_check.json 2> gpurun_out/api_check.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/api_check.json").read().strip().splitlines()[-1])
print("value", d["value"], "e2e", d["e2e"]["value"], "single", d["e2e"]["single_call"]["value"], "parity", d.get("parity_vs_reference"))
print(d["unsharp"])
PY
