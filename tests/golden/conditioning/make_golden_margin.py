#!/usr/bin/env python
"""Writes tests/golden/conditioning/blur_margin.npz from the UNMODIFIED reference library (oracle/_ref/libpoppy_ref.so ->
poppy::blur_margin, reference src/util.cpp:574-602). Run in the build container; the fixture travels, the reference does not."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))))
from oracle import ref  # noqa: E402

rng = np.random.default_rng(2024)
out = {}
cases = [(60, 90, 80, 120), (48, 48, 48, 48), (120, 40, 120, 100), (75, 131, 140, 131)]
for i, (h, w, uh, uw) in enumerate(cases):
    src = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    src[h // 4:h // 2, w // 4:w // 2] = (255, 0, 128)          # flat saturated block: rounding at the extremes
    out[f"src{i}"] = src
    out[f"union{i}"] = np.array([uw, uh], np.int32)
    out[f"dst{i}"] = ref.blur_margin(src, (uw, uh))
out["n"] = np.int32(len(cases))
path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "blur_margin.npz")
np.savez_compressed(path, **out)
print("wrote", path, os.path.getsize(path), "bytes")
