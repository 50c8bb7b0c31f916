timeout 300 python -m pytest tests/test_gpu_stages.py tests/test_gpu_api.py tests/test_gpu_calm.py tests/test_gpu_configs.py -x -q 2>&1 | tail -3
for m in 0 1; do
timeout 900 python bench.py --unsharp-mode $m --kprocs 0 --cpu-frames 2 > gpurun_out/bench_r2_mode$m.json 2> gpurun_out/bench_r2_mode$m.err
python - <<EOF
import json
d=json.load(open("gpurun_out/bench_r2_mode$m.json"))
print("MODE $m fps", d["value"], "frac", d["roofline"]["frac"], "e2e", d["e2e"]["value"], d["parity_vs_reference"], d["unsharp"]["exact_chunk_share"], "staged", d["roofline"]["stage_timed_step_ms"], "ms/step", d["ms_per_step"])
EOF
done
timeout 600 python bench.py --mode chain --steps 5 --warmup 3 --e2e-steps 1 > gpurun_out/bench_r2_chain.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/bench_r2_chain.json')); print('CHAIN', d['value'], d['e2e']['value'], d['unsharp'], d['parity_vs_reference'])"
