timeout 180 python -m pytest tests/test_gpu_stages.py tests/test_gpu_calm.py -x -q 2>&1 | tail -4
bash tools/gpu_ab.sh \
 "rw6|X=1|--frames 192 --unsharp-mode 1" \
 "rw8|POPPY_CUDA_RW_CTAS=8|--frames 192 --unsharp-mode 1" \
 "rw5|POPPY_CUDA_RW_CTAS=5|--frames 192 --unsharp-mode 1" \
 "rw4|POPPY_CUDA_RW_CTAS=4|--frames 192 --unsharp-mode 1"
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
