#!/bin/bash
# N-GPU default bench line incl. e2e (the driver's scaling run). usage: tools/gpu_multi_e2e.sh <N>
N=${1:-4}
mkdir -p gpurun_out
nproc
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/scale${N}.json 2> gpurun_out/scale${N}.err || echo FAILED
python - <<PY
import json
d=json.loads(open("gpurun_out/scale${N}.json").read().strip().splitlines()[-1])
print("n", d["n_gpus"], "value", d["value"], "e2e", d["e2e"]["value"], "breakdown", d["e2e"].get("breakdown"), "host_plan", d.get("host_plan"))
PY
tail -2 gpurun_out/scale${N}.err
