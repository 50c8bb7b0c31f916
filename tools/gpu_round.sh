set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -3
timeout 900 python bench.py --frames 96 --steps 3 --warmup 3 --cpu-frames 2 --e2e-steps 1 > gpurun_out/bench_small.json 2> gpurun_out/bench_small.err; tail -c 3000 gpurun_out/bench_small.json; tail -5 gpurun_out/bench_small.err
