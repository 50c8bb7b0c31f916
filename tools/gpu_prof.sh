# ncu --set full capture (with source) of selected kernels of one 16-frame chunk; no tests, no bench.
# usage: bash tools/gpu_prof.sh tag kernel-regex skip count   [env POPPY_* switches pass through]
set -x
TAG=${1:-prof}; KRE=${2:-k_raster_warp}; SKIP=${3:-1}; COUNT=${4:-1}
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$KRE" -s $SKIP -c $COUNT -f -o gpurun_out/prof_${TAG} \
   python bench.py --frames 16 --steps 1 --warmup 1 --cpu-frames 0 --e2e-steps 0 --no-stage-pass > gpurun_out/ncu_full_${TAG}.log 2>&1
tail -3 gpurun_out/ncu_full_${TAG}.log | cut -c1-300
ls -la gpurun_out/
