#!/bin/bash
# N-GPU sanity of every bench mode, launched the way the driver launches it. usage: tools/gpu_multi.sh <N>
N=${1:-2}
mkdir -p gpurun_out
run() {  # tag, args...
  tag=$1; shift
  timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus $N "$@" > gpurun_out/multi${N}_${tag}.json 2> gpurun_out/multi${N}_${tag}.err || echo "FAILED $tag"
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/multi${N}_${tag}.json").read().strip().splitlines()[-1])
    print("${tag}", "n", d.get("n_gpus"), "fps", d.get("value"), "scaling", d.get("scaling"), "frac", (d.get("roofline") or {}).get("frac"), "e2e", (d.get("e2e") or {}).get("value"), "sum", d.get("frames_checksum"), "impl", d.get("impl"))
except Exception as e:
    print("${tag}", "no json", e)
PY
  tail -1 gpurun_out/multi${N}_${tag}.err
}
run weak4k --frames 192 --steps 3 --warmup 3 --cpu-frames 0 --kprocs 0 --e2e-steps 1
run strong8k --workload 8k --total-frames 192 --ring 32 --steps 2 --warmup 3 --cpu-frames 0 --kprocs 0 --e2e-steps 0
run chain --mode chain --steps 3 --warmup 3 --e2e-steps 1
run ref --impl reference --steps 1 --warmup 1 --kprocs 0
python tools/stress_determinism.py > gpurun_out/stress_determinism.log 2>&1; tail -3 gpurun_out/stress_determinism.log
