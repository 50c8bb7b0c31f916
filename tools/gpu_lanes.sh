#!/bin/bash
for l in ${LANES_LIST:-3 2 4 1 3}; do
  POPPY_CUDA_LANES=$l python bench.py --steps 4 --warmup 3 --cpu-frames 0 --kprocs 0 --e2e-steps 0 --no-stage-pass ${BENCH_ARGS} > gpurun_out/lanes$l.json 2>/dev/null
  python -c "
import json; d=json.loads(open('gpurun_out/lanes$l.json').read().strip().splitlines()[-1]); print('lanes $l value', round(d['value'],1), 'ms/step', round(d['ms_per_step'],2), d['clocks'])"
done
nvidia-smi --query-gpu=power.draw,temperature.gpu,clocks.sm,clocks.mem --format=csv,noheader
