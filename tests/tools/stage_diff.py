"""Debug aid: print per-stage mismatch counts of the CUDA path against the reference library for a few cases.
Run on a GPU box:  python tests/tools/stage_diff.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from poppy_b200 import build, synth  # noqa: E402
from poppy_b200 import renderer as R  # noqa: E402
from oracle import ref  # noqa: E402
from tests.util import bits_differ, flat_tri  # noqa: E402

build.build()
cases = [(320, 240, 200, 0.37, 6), (203, 157, 60, 0.5, 64), (257, 131, 100, 0.9, 4), (641, 479, 300, 0.37, 6),
         (100, 75, 20, 0.0, 6), (64, 48, 12, 0.5, 2), (1920, 1080, 2000, 0.5, 6)]
for (w, h, n, s, L) in cases:
    inp = synth.make_inputs(w, h, n, 8.0, seed=7)
    want = ref.stages(inp.bgr1, inp.bgr2, inp.gabor2, inp.pts1, inp.pts2, s, s, L)
    tri = want.tri_idx
    with R.MorphRenderer(w, h, L, len(inp.pts1), max(len(tri), 1), 1, keep_stages=True) as r:
        r.set_pair(inp.bgr1, inp.bgr2, inp.gabor2)
        r.set_points(inp.pts1, inp.pts2)
        cat, offs = flat_tri([tri])
        r.render([s], [s], cat, offs)
        dst = r.download(0, 1)[0]
        rep = {
            "points": bits_differ(r.read_stage(R.STAGE_MORPHED_POINTS, 0), want.morphed_points),
            "tri_map": bits_differ(r.read_stage(R.STAGE_TRI_MAP, 0), want.tri_map),
            "warped1": bits_differ(r.read_stage(R.STAGE_WARPED1, 0), want.warped1),
            "warped2": bits_differ(r.read_stage(R.STAGE_WARPED2, 0), want.warped2),
            "mask": bits_differ(r.read_stage(R.STAGE_MASK, 0), want.mask),
            "lap_blend": bits_differ(r.read_stage(R.STAGE_LAP_BLEND, 0), want.lap_blend),
            "dst": bits_differ(dst, want.dst),
        }
        lb = r.read_stage(R.STAGE_LAP_BLEND, 0)
        print((w, h, n, s, L), "T=%d" % len(tri), rep, "lap maxabs %.3e" % np.abs(lb - want.lap_blend).max(),
              "render ms %.3f" % r.last_render_ms(), flush=True)
