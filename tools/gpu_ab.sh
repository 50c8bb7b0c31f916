# A/B bench runs: each argument is "ENV=.. ENV=.. -- bench args"; prints fps + per-class ms for each.
# usage: bash tools/gpu_ab.sh "label|env assignments|bench args" ...
mkdir -p gpurun_out
for spec in "$@"; do
  label=${spec%%|*}; rest=${spec#*|}; envs=${rest%%|*}; args=${rest#*|}
  env $envs timeout 600 python bench.py --frames 96 --steps 3 --warmup 3 --cpu-frames 0 --e2e-steps 0 $args > gpurun_out/ab_${label}.json 2> gpurun_out/ab_${label}.err
  python - "$label" gpurun_out/ab_${label}.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
    ks = " ".join(f"{k['kernel']}={k['ms_per_step']:.2f}" for k in d["roofline"]["kernels"][:6])
    print(f"[{sys.argv[1]}] fps={d['value']:.1f} ms/step={d['ms_per_step']:.2f} frac={d['roofline']['frac']:.4f} staged={d['roofline']['stage_timed_step_ms']:.2f} | {ks} | sum={d['frames_checksum']}")
except Exception as e:
    print(f"[{sys.argv[1]}] FAILED {e}"); print(open(sys.argv[2].replace('.json', '.err')).read()[-1500:])
PY
done
