# quick per-kernel counters of one 32-frame chunk: time, warp instructions, issue utilisation, DRAM bytes
mkdir -p gpurun_out
TAG=${1:-q}
timeout 900 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread \
  --clock-control none ${NCU_KERNEL:+-k regex:$NCU_KERNEL} -s ${NCU_SKIP:-40} -c ${NCU_COUNT:-40} --csv --log-file gpurun_out/ncu_quick_${TAG}.csv \
  python bench.py --frames 32 --steps 1 --warmup 1 --cpu-frames 0 --e2e-steps 0 --no-stage-pass ${BENCH_ARGS:---unsharp-mode 1} > gpurun_out/ncu_quick_${TAG}.log 2>&1
tail -2 gpurun_out/ncu_quick_${TAG}.log
python - <<EOF
import csv, collections
rows = list(csv.reader(open("gpurun_out/ncu_quick_${TAG}.csv")))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
H = rows[hdr]
k = collections.OrderedDict()
for r in rows[hdr + 1:]:
    if len(r) < len(H): continue
    d = dict(zip(H, r))
    key = (d["ID"], d["Kernel Name"][:40], d.get("Grid Size", ""))
    k.setdefault(key, {})[d["Metric Name"]] = float(d["Metric Value"].replace(",", ""))
for (i, name, grid), m in k.items():
    print(f"{i:>4} {name:40s} {grid:>22s} t={m.get('gpu__time_duration.sum',0)/1e3:8.1f}us inst={m.get('smsp__inst_executed.sum',0)/1e6:8.1f}M issue={m.get('smsp__issue_active.avg.pct_of_peak_sustained_active',0):5.1f}% warps={m.get('sm__warps_active.avg.pct_of_peak_sustained_active',0):5.1f}% rd={m.get('dram__bytes_read.sum',0)/1e6:8.1f}MB wr={m.get('dram__bytes_write.sum',0)/1e6:8.1f}MB regs={m.get('launch__registers_per_thread',0):.0f}")
EOF
