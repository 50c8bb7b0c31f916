#!/bin/bash
python -m pytest tests/test_margin.py -q 2>&1 | tail -3
python - <<'PY'
import time, numpy as np
from poppy_b200 import api
from oracle import ref
rng = np.random.default_rng(0)
img = rng.integers(0, 256, (2000, 3600, 3), dtype=np.uint8)
api.blur_margin(img, (3840, 2160))
t = time.perf_counter(); g = api.blur_margin(img, (3840, 2160)); tg = time.perf_counter() - t
t = time.perf_counter(); r = ref.blur_margin(img, (3840, 2160)); tr = time.perf_counter() - t
print(f"blur_margin 3600x2000 -> 3840x2160: CUDA entry (H2D + 8 kernels + D2H) {tg*1e3:.1f} ms, reference CPU {tr*1e3:.1f} ms, identical={bool((g == r).all())}")
PY
python - <<'PY'
import time, numpy as np
from poppy_b200 import api, synth
from oracle import ref
src = synth.noise_image(3840, 2160, 5).astype(np.float32) / np.float32(255)
api.gabor_filter(src[:64, :64].copy())
t = time.perf_counter(); g = api.gabor_filter(src); tg = time.perf_counter() - t
t = time.perf_counter(); r = ref.gabor_filter(src); tr = time.perf_counter() - t
d = np.abs(g - r)
print(f"gabor_filter 3840x2160: CUDA entry (H2D + kernel + D2H) {tg*1e3:.1f} ms, reference CPU {tr*1e3:.1f} ms, max abs diff {d.max():.3g}, values differing {(g.view(np.uint32) != r.view(np.uint32)).sum()} of {g.size}")
PY
