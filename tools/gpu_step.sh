mkdir -p gpurun_out; timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15; timeout 600 python bench.py --frames 96 --steps 3 --warmup 3 --cpu-frames 2 --e2e-steps 1 > gpurun_out/bench_r2_calm2.json 2> gpurun_out/bench_r2_calm2.err; python - <<EOF
import json
d=json.load(open("gpurun_out/bench_r2_calm2.json"))
print(d["value"], d["roofline"]["frac"], d["parity_vs_reference"], d["frames_checksum"])
for k in d["roofline"]["kernels"]: print(k)
EOF
tail -3 gpurun_out/bench_r2_calm2.err
