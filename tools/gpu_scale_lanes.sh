#!/bin/bash
# N-GPU device-timed value (no e2e, no CPU baseline), per-rank step times. usage: LANES_LIST="4 3" tools/gpu_scale_lanes.sh <N>
N=${1:-8}
mkdir -p gpurun_out
for l in ${LANES_LIST:-4}; do
  POPPY_CUDA_LANES=$l timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 \
      bench.py --gpus $N --steps 5 --warmup 3 --cpu-frames 0 --kprocs 0 --e2e-steps 0 --no-stage-pass ${BENCH_ARGS} > gpurun_out/scale_lanes${l}.json 2> gpurun_out/scale_lanes${l}.err || echo FAILED
  python - <<PY
import json
d = json.loads(open("gpurun_out/scale_lanes${l}.json").read().strip().splitlines()[-1])
print("N", d["n_gpus"], "lanes ${l} value", round(d["value"], 1), "per GPU", round(d["value"] / d["n_gpus"], 1), "ms/step", round(d["ms_per_step"], 2),
      d["clocks"]["reasons"], "by rank", d.get("ms_per_step_by_rank"))
PY
done
nvidia-smi --query-gpu=index,power.draw,temperature.gpu,clocks.sm --format=csv,noheader | head -8
