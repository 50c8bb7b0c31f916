# launch list + ncu --set full of every tiled kernel of one 32-frame chunk (current defaults, dense unsharp route), then the
# same for the calm route on the chain workload (the kernels that route adds)
set -x
TAG=${1:-r2}
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_${TAG}.csv \
   python bench.py --frames 96 --steps 1 --warmup 3 --cpu-frames 0 --e2e-steps 0 --no-stage-pass --unsharp-mode 1 > gpurun_out/ncu_list_${TAG}.log 2>&1
tail -2 gpurun_out/ncu_list_${TAG}.log | cut -c1-200
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"k_collapse|k_pyr_down|k_unsharp|k_raster|k_tri_geometry" -s 15 -c 15 -f -o gpurun_out/prof_${TAG} \
   python bench.py --frames 32 --steps 1 --warmup 1 --cpu-frames 0 --e2e-steps 0 --no-stage-pass --unsharp-mode 1 > gpurun_out/ncu_full_${TAG}.log 2>&1
tail -2 gpurun_out/ncu_full_${TAG}.log | cut -c1-200
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_calm|k_collapse_tma<1, 1>|k_collapse_tma<\(bool\)1, \(bool\)1>" -s 0 -c 8 -f -o gpurun_out/prof_${TAG}_calm \
   python bench.py --frames 16 --steps 1 --warmup 0 --cpu-frames 0 --e2e-steps 0 --no-stage-pass --unsharp-mode 2 > gpurun_out/ncu_full_${TAG}_calm.log 2>&1
tail -2 gpurun_out/ncu_full_${TAG}_calm.log | cut -c1-200
ls -la gpurun_out | tail -6
