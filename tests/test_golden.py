"""The CPU restatement against the committed golden vectors (tests/golden/*.npz, generated from the reference
library by tests/golden/make_golden.py). Runs everywhere, needs neither the reference tree nor a GPU."""
import glob
import os

import numpy as np
import pytest

from oracle import port
from tests.util import bits_differ

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FRAME_FILES = sorted(f for f in glob.glob(os.path.join(GOLDEN, "*.npz"))
                     if not os.path.basename(f).startswith(("chain_", "topology")))
STAGE_KEYS = ["morphed_points", "tri_map", "hom", "m1", "m2", "mapx1", "mapy1", "mapx2", "mapy2", "warped1", "warped2",
              "mask", "lap_blend", "dst"]


def test_fixtures_present():
    assert len(FRAME_FILES) >= 8
    assert os.path.exists(os.path.join(GOLDEN, "chain_shapes_80x64_N12_L64.npz"))
    assert os.path.exists(os.path.join(GOLDEN, "topology.npz"))


@pytest.mark.parametrize("path", FRAME_FILES, ids=[os.path.basename(f)[:-4] for f in FRAME_FILES])
def test_port_reproduces_golden_stages(path):
    g = np.load(path)
    got = port.morph_frame(g["bgr1"], g["bgr2"], g["gabor2"], g["pts1"], g["pts2"], g["tri_idx"], float(g["shape"]),
                           float(g["mask_ratio"]), int(g["levels"]))
    for k in STAGE_KEYS:
        assert bits_differ(getattr(got, k), g[k]) == 0, k


def test_port_reproduces_golden_chain():
    """Frame-loop recurrence (reference src/poppy.hpp:177-243) restated around the port + the host topology."""
    from poppy_b200 import host
    g = np.load(os.path.join(GOLDEN, "chain_shapes_80x64_N12_L64.npz"))
    n_frames, levels = int(g["n_frames"]), int(g["levels"])
    h, w = g["bgr1"].shape[:2]
    cur_img, cur_pts = g["bgr1"], g["pts1"]
    for j in range(n_frames):
        s = host.chain_ratio(j, n_frames)
        mp = host.morph_points(cur_pts, g["pts2"], s, w, h)
        tri = host.triangulate(mp, w, h)
        st = port.morph_frame(cur_img, g["bgr2"], g["gabor2"], cur_pts, g["pts2"], tri, s, s, levels)
        assert bits_differ(st.morphed_points, g["points"][j]) == 0, j
        assert (st.dst == g["frames"][j]).all(), j
        cur_img, cur_pts = st.dst, st.morphed_points


def test_fixtures_come_from_the_avx2_fma_dispatch_of_the_reference():
    """Bit parity is defined against the reference's AVX2/FMA-dispatched OpenCV build: the whole-pipeline
    fixtures record the CPU features the reference ran with when they were made."""
    import os
    import numpy as np
    full = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "full")
    for c in (1, 2, 3):
        feats = str(np.load(os.path.join(full, f"c{c}.npz"))["cpu_features"])
        assert "AVX2" in feats and "FP16" in feats, feats
