# One GPU round: parity tests, smoke, short bench, ncu launch list (+ optional full capture of one kernel).
# usage: bash tools/gpu_round.sh [tag] [ncu-kernel-regex]
set -x
TAG=${1:-run}
KRE=${2:-}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -25
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -3
timeout 900 python bench.py --frames 96 --steps 3 --warmup 3 --cpu-frames 2 --e2e-steps 1 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; tail -c 3500 gpurun_out/bench_${TAG}.json; tail -5 gpurun_out/bench_${TAG}.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_${TAG}.csv \
   python bench.py --frames 16 --steps 1 --warmup 3 --cpu-frames 0 --e2e-steps 0 --no-stage-pass > gpurun_out/ncu_list_${TAG}.log 2>&1
tail -3 gpurun_out/ncu_list_${TAG}.log
if [ -n "$KRE" ]; then
  # one chunk of 16 frames launches 14 tiled kernels (raster_warp, down0, 5 down, 5 collapse, collapse0, unsharp)
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$KRE" -s ${NCU_SKIP:-14} -c ${NCU_COUNT:-14} -f -o gpurun_out/prof_${TAG} \
     python bench.py --frames 16 --steps 1 --warmup 1 --cpu-frames 0 --e2e-steps 0 --no-stage-pass > gpurun_out/ncu_full_${TAG}.log 2>&1
  tail -3 gpurun_out/ncu_full_${TAG}.log
fi
