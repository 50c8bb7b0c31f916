# A/B on the full 600-frame workload (no cpu baseline / e2e): "label|env|bench args" ...
mkdir -p gpurun_out
for spec in "$@"; do
  label=${spec%%|*}; rest=${spec#*|}; envs=${rest%%|*}; args=${rest#*|}
  env $envs timeout 600 python bench.py --steps 3 --warmup 3 --cpu-frames 0 --e2e-steps 0 --no-stage-pass $args > gpurun_out/ab_${label}.json 2> gpurun_out/ab_${label}.err
  python - "$label" gpurun_out/ab_${label}.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
    print(f"[{sys.argv[1]}] fps={d['value']:.1f} ms/step={d['ms_per_step']:.2f} frac={d['roofline']['frac']:.4f} sum={d['frames_checksum']}")
except Exception as e:
    print(f"[{sys.argv[1]}] FAILED {e}"); print(open(sys.argv[2].replace('.json', '.err')).read()[-1500:])
PY
done
