#!/bin/bash
# secondary bench lines: photographic content (the reference's demo pair tiled to 4K), adaptive and dense unsharp routes
mkdir -p gpurun_out
lscpu > gpurun_out/box_lscpu.txt 2>&1
python bench.py --content photo --steps 3 --warmup 3 --cpu-frames 2 --kprocs 0 --e2e-steps 1 > gpurun_out/bench_photo_adaptive.json 2> gpurun_out/bench_photo_adaptive.err
python bench.py --content photo --unsharp-mode 1 --steps 3 --warmup 3 --cpu-frames 0 --kprocs 0 --e2e-steps 0 > gpurun_out/bench_photo_dense.json 2> gpurun_out/bench_photo_dense.err
python - <<'PY'
import json
for n in ("adaptive","dense"):
    try:
        d=json.loads(open(f"gpurun_out/bench_photo_{n}.json").read().strip().splitlines()[-1])
        print(n, d["value"], d["roofline"]["frac"], d["e2e"]["value"], d.get("cpu_baseline",{}).get("parity_vs_reference"), d.get("unsharp"))
    except Exception as e: print(n, "failed", e)
PY
grep -i "model name\|^CPU(s)\|L2" gpurun_out/box_lscpu.txt
