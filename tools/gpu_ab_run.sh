timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 600 python bench.py --mode chain --steps 5 --warmup 3 --e2e-steps 2 > gpurun_out/bench_r2_chain.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/bench_r2_chain.json')); print('CHAIN', d['value'], 'e2e', d['e2e']['value'], d['unsharp']['exact_chunk_share'], d['parity_vs_reference'], d['cpu_baseline']['value'], 'launches', d['gpu_launches'])"
BENCH_ARGS="--mode chain" NCU_SKIP=65 NCU_COUNT=3 bash tools/gpu_ncu_quick.sh chain | grep tail | cut -c1-200
