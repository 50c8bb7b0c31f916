set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
./tools/texprobe
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25
bash tools/gpu_ab.sh "occ5|POPPY_CUDA_RW_CTAS=5|" "occ4|POPPY_CUDA_RW_CTAS=4|" "occ6|POPPY_CUDA_RW_CTAS=6|" "chunk32|POPPY_CUDA_RW_CTAS=5|--chunk 32"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_v6.csv \
   python bench.py --frames 16 --steps 1 --warmup 3 --cpu-frames 0 --e2e-steps 0 --no-stage-pass > gpurun_out/ncu_list_v6.log 2>&1
tail -2 gpurun_out/ncu_list_v6.log | cut -c1-300
