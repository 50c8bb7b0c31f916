# quick GPU check: parity tests, then A/B bench runs given as gpu_ab.sh specs
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
bash tools/gpu_ab.sh "$@"
