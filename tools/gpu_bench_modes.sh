# the bench lines of the other workloads / modes (short runs): 1080p, chain, 8k strong-scaling slices, reference arm
mkdir -p gpurun_out
for spec in "1080p:--workload 1080p --frames 192 --steps 3 --warmup 3 --cpu-frames 2 --kprocs 0 --e2e-steps 1" \
            "chain:--mode chain --steps 5 --warmup 3 --e2e-steps 1" \
            "8k:--workload 8k --total-frames 96 --ring 32 --steps 2 --warmup 3 --cpu-frames 1 --kprocs 0 --e2e-steps 1" \
            "refchain:--impl reference --mode chain --steps 2 --warmup 1" \
            "ref4k:--impl reference --steps 2 --warmup 1 --kprocs 4"; do
  tag=${spec%%:*}; args=${spec#*:}
  timeout 900 python bench.py $args > gpurun_out/bench_modes_${tag}.json 2> gpurun_out/bench_modes_${tag}.err || echo "FAILED $tag"
  python - <<EOF
import json
try:
    d=json.load(open("gpurun_out/bench_modes_${tag}.json"))
    print("${tag}", "fps", d.get("value"), "frac", (d.get("roofline") or {}).get("frac"), "e2e", (d.get("e2e") or {}).get("value"), "parity", d.get("parity_vs_reference"), "cpu", d.get("cpu_baseline"), "unsharp", d.get("unsharp"))
except Exception as e:
    print("${tag}", "no json", e)
EOF
  tail -2 gpurun_out/bench_modes_${tag}.err
done
