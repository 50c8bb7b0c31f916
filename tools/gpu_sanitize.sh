#!/bin/bash
# compute-sanitizer over small renders of every route (dense / calm / chain / 64 levels / odd sizes) and the conditioning kernels
mkdir -p gpurun_out
cat > san_tmp.py <<'PY'
import numpy as np
from poppy_b200 import api, host, synth
from poppy_b200.renderer import MorphRenderer
def run(w, h, levels, n, frames, chain, mode):
    inp = synth.block_inputs(w, h, n, seed=w + h)
    ph = np.linspace(0.1, 0.9, frames).astype(np.float32)
    plan = host.SequencePlan(inp.pts1, inp.pts2, w, h, ph, chain=chain, threads=2)
    r = MorphRenderer(w, h, levels, len(inp.pts1), plan.max_triangles, frames)
    r.set_unsharp_mode(mode)
    r.set_pair(inp.bgr1, inp.bgr2, inp.gabor2); r.set_points(inp.pts1, inp.pts2)
    r.render(ph, ph.astype(np.float64), plan.tri_idx, plan.tri_offsets, chain=chain)
    out = r.download(0, frames); r.close()
    return int(out.sum())
print(run(640, 384, 6, 300, 5, False, 1), run(640, 384, 6, 300, 5, False, 2), run(333, 217, 64, 60, 4, True, 0), run(1024, 512, 3, 500, 3, False, 0))
img = np.random.default_rng(0).integers(0, 256, (150, 211, 3), dtype=np.uint8)
print(int(api.blur_margin(img, (260, 180)).sum()), float(api.gabor_filter(img.astype(np.float32) / 255).sum()))
PY
for tool in memcheck racecheck; do
  timeout 1500 compute-sanitizer --tool $tool --print-limit 5 python san_tmp.py > gpurun_out/sanitizer_$tool.log 2>&1
  echo "== $tool: $(grep -c 'ERROR SUMMARY' gpurun_out/sanitizer_$tool.log) summaries"; grep "ERROR SUMMARY\|RACECHECK SUMMARY\|hazard" gpurun_out/sanitizer_$tool.log | head -5; tail -3 gpurun_out/sanitizer_$tool.log
done
rm -f san_tmp.py
