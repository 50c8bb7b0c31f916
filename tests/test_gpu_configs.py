"""BASELINE.json configs 1-3: the reference's own demo pairs driven through its own front end (blur_margin, extractor,
matcher / face landmarks / autoalign, gabor_filter) and frame loop (reference src/poppy.hpp:46-248) at the default 64
pyramid levels. The fixtures are what that pipeline handed to morph_images() and got back, frame by frame
(tests/golden/make_golden_full.py); here the whole chain runs on the GPU with frame j-1 resident as frame j's source and
every frame must come out byte-identical (the north-star tolerance is asserted as well)."""
import os

import numpy as np
import pytest

from poppy_b200 import api, host
from tests.util import assert_frame_parity, bits_differ, frame_checksum

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
FULL = os.path.join(HERE, "golden", "full")
DUMPS = os.path.join(os.path.dirname(HERE), "oracle", "_ref", "full")


@pytest.fixture(autouse=True)
def _fresh_api(native_lib):
    yield
    api.release()


def _check_chain(frames, hashes, what):
    bad = [j for j in range(len(hashes)) if frame_checksum(frames[j]) != int(hashes[j])]
    assert not bad, f"{what}: frames {bad[:8]} differ from the reference pipeline ({len(bad)} of {len(hashes)})"


@pytest.mark.parametrize("config", [1, 2])
def test_config_chain_is_bit_identical_to_the_reference_pipeline(config):
    g = np.load(os.path.join(FULL, f"c{config}.npz"))
    N = int(g["n_frames"])
    assert N == 60 and int(g["levels"]) == 64
    # the reference's schedule (src/poppy.hpp:181-210) is reproduced by poppy_host_chain_ratio
    want_ratio = g["ratios"]
    got_ratio = np.array([host.chain_ratio(j, N) for j in range(N)])
    assert (got_ratio == want_ratio[:, 0]).all() and (got_ratio == want_ratio[:, 1]).all()
    api.Settings.instance().pyramid_levels = int(g["levels"])
    frames = api.morph_sequence(g["corrected1"], g["corrected2"], g["gabor2"], g["pts1"], g["pts2"], number_of_frames=N)
    for j, want in zip(g["kept"], g["kept_frames"]):
        assert_frame_parity(frames[j], want, (config, int(j)))
        assert bits_differ(frames[j], want) == 0, (config, int(j))
    _check_chain(frames, g["hashes"], f"config {config}")
    # the point recurrence (lastMorphedPoints, src/poppy.hpp:178-179,218) as the host planner sees it
    plan = host.SequencePlan(g["pts1"], g["pts2"], int(g["width"]), int(g["height"]), got_ratio.astype(np.float32), chain=True)
    try:
        for j in (0, 1, N // 2, N - 1):
            assert bits_differ(plan.points(j), g["morphed"][j]) == 0, (config, j)
    finally:
        plan.close()


def test_config3_1080p_chain_is_bit_identical_to_the_reference_pipeline():
    """cat -> dog with --autoalign on a 1920x1080 canvas, 120 chained frames, 64 levels. The 1080p inputs are too large for
    the repository: they come from the reference dump shipped under oracle/_ref/full/c3."""
    d = os.path.join(DUMPS, "c3")
    if not os.path.exists(os.path.join(d, "meta.txt")):
        pytest.skip("oracle/_ref/full/c3 not shipped (run oracle/build_ref_full.sh + the dump, see tests/golden/make_golden_full.py)")
    import sys
    sys.path.insert(0, os.path.join(HERE, "golden"))
    from make_golden_full import load_dump
    D = load_dump(d)
    g = np.load(os.path.join(FULL, "c3.npz"))
    assert (D["hashes"] == g["hashes"]).all(), "the shipped dump is not the one the committed fixture was made from"
    N = D["frames"]
    assert (D["w"], D["h"], N, D["levels"]) == (1920, 1080, 120, 64)
    api.Settings.instance().pyramid_levels = D["levels"]
    frames = api.morph_sequence(D["image"]("corrected1"), D["image"]("corrected2"), D["gabor2"](), D["pts1"], D["pts2"],
                                number_of_frames=N)
    for j in D["kept"]:
        want = D["frame"](j)
        assert_frame_parity(frames[j], want, ("config 3", j))
        assert bits_differ(frames[j], want) == 0, ("config 3", j)
    _check_chain(frames, D["hashes"], "config 3")
