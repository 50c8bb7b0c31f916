# launch list + ncu --set full of every tiled kernel of one 32-frame chunk (current defaults)
set -x
TAG=${1:-v10}
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_${TAG}.csv \
   python bench.py --frames 96 --steps 1 --warmup 3 --cpu-frames 0 --e2e-steps 0 --no-stage-pass > gpurun_out/ncu_list_${TAG}.log 2>&1
tail -2 gpurun_out/ncu_list_${TAG}.log | cut -c1-200
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_collapse_roll|k_pyr_down|k_unsharp|k_raster|k_blend_coarsest|k_tri_geometry" -s 16 -c 16 -f -o gpurun_out/prof_${TAG} \
   python bench.py --frames 32 --steps 1 --warmup 1 --cpu-frames 0 --e2e-steps 0 --no-stage-pass > gpurun_out/ncu_full_${TAG}.log 2>&1
tail -2 gpurun_out/ncu_full_${TAG}.log | cut -c1-200
ls -la gpurun_out | tail -5
