# one GPU iteration: quick smoke of the newest kernel under a short timeout (a hang must not eat the box), then the GPU
# tests, then a short bench with the per-class stage times
mkdir -p gpurun_out
timeout 180 python -m pytest tests/test_gpu_stages.py -x -q 2>&1 | tail -4 || { echo "SMOKE FAILED"; exit 1; }
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
TAG=${1:-step}
timeout 600 python bench.py --frames 192 --steps 3 --warmup 3 --cpu-frames 2 --e2e-steps 1 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
python - <<EOF
import json
d=json.load(open("gpurun_out/bench_${TAG}.json"))
print("fps", d["value"], "frac", d["roofline"]["frac"], d["parity_vs_reference"], d["frames_checksum"], "e2e", d["e2e"]["value"], d.get("unsharp"))
for k in d["roofline"]["kernels"]: print(k)
EOF
tail -3 gpurun_out/bench_${TAG}.err
