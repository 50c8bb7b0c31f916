"""SURVEY.md 8(f-3), blur_margin (reference src/util.cpp:574-602): the numpy restatement (oracle/margin.py) is pinned against
the reference library and the committed fixture; the CUDA entry point poppy_cuda_blur_margin is compared with both."""
import os

import numpy as np
import pytest

from oracle import margin, ref

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "blur_margin.npz")
# (rows, cols, union_h, union_w): both axes padded, one axis, none (1-pixel margins), tall/wide, margins wider than the padding
CASES = [(100, 150, 120, 180), (100, 150, 100, 150), (200, 100, 200, 160), (90, 90, 150, 90), (333, 211, 400, 300),
         (64, 48, 70, 300), (500, 500, 512, 512)]


def _img(rng, h, w):
    return rng.integers(0, 256, (h, w, 3), dtype=np.uint8)


def test_fixed_point_kernel_of_blur_margin():
    k = margin.gaussian_kernel_fixed(127, 6.0)
    assert sum(k) == 256 and k[63] == 18 and k == k[::-1]
    assert [i for i, v in enumerate(k) if v][0] == 46 and k[46:50] == [1, 0, 1, 1]       # error diffusion leaves a gap


@pytest.mark.skipif(not ref.available(), reason="reference library not built")
def test_kernel_matches_reference_impulse_response():
    # a 1-row image shrinks the vertical kernel to [1]; 255 * tap / 256 rounds back to the tap for taps < 128
    img = np.zeros((1, 301, 3), np.uint8)
    img[0, 150] = 255
    out = ref.gaussian_blur_u8(img, 127, 6.0)
    assert out[0, 150 - 63:150 + 64, 1].astype(int).tolist() == margin.gaussian_kernel_fixed(127, 6.0)


@pytest.mark.skipif(not ref.available(), reason="reference library not built")
@pytest.mark.parametrize("shape", [(40, 70), (1, 50), (50, 1), (3, 3), (130, 20), (20, 200), (2, 2), (127, 127), (300, 190)])
def test_oracle_gaussian_blur_u8_matches_reference(shape):
    img = _img(np.random.default_rng(shape[0] * 1000 + shape[1]), *shape)
    assert (ref.gaussian_blur_u8(img, 127, 6.0) == margin.gaussian_blur_u8(img, 127, 6.0)).all()


@pytest.mark.skipif(not ref.available(), reason="reference library not built")
@pytest.mark.parametrize("case", CASES)
def test_oracle_blur_margin_matches_reference(case):
    h, w, uh, uw = case
    img = _img(np.random.default_rng(h + 7 * w), h, w)
    assert (ref.blur_margin(img, (uw, uh)) == margin.blur_margin(img, (uw, uh))).all()


def test_oracle_blur_margin_matches_golden():
    g = np.load(GOLDEN)
    for i in range(int(g["n"])):
        got = margin.blur_margin(g[f"src{i}"], tuple(g[f"union{i}"]))
        assert (got == g[f"dst{i}"]).all(), i


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES + [(1080, 1920, 1200, 2100)])
def test_cuda_blur_margin_matches_oracle_and_reference(native_lib, case):
    from poppy_b200 import api
    h, w, uh, uw = case
    img = _img(np.random.default_rng(h + 7 * w), h, w)
    got = api.blur_margin(img, (uw, uh))
    want = ref.blur_margin(img, (uw, uh)) if ref.available() else margin.blur_margin(img, (uw, uh))
    assert got.shape == want.shape and (got == want).all()
    # strided input rows (a cv::Mat ROI of a wider image)
    wide = np.zeros((h, w + 9, 3), np.uint8)
    wide[:, :w] = img
    assert (api.blur_margin(wide[:, :w], (uw, uh)) == want).all()


@pytest.mark.gpu
def test_cuda_blur_margin_matches_golden_and_rejects_bad_geometry(native_lib):
    from poppy_b200 import api
    g = np.load(GOLDEN)
    for i in range(int(g["n"])):
        assert (api.blur_margin(g[f"src{i}"], tuple(g[f"union{i}"])) == g[f"dst{i}"]).all(), i
    with pytest.raises(RuntimeError):          # the source is larger than the union: cv::Mat ROI assertion in the reference
        api.blur_margin(np.zeros((50, 80, 3), np.uint8), (60, 50))
