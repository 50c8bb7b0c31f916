"""Host stages of the product (C++ part of libpoppy_cuda.so, include/poppy_host.h): point hygiene, the
cv::Subdiv2D-exact Delaunay + vertex-index lookup, the chain schedule and the threaded sequence planner."""
import os

import numpy as np
import pytest

from oracle import port, ref
from poppy_b200 import host, synth
from tests.util import bits_differ

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_triangulation_matches_golden_subdiv2d(native_lib):
    g = np.load(os.path.join(GOLDEN, "topology.npz"))
    for i in range(int(g["n_cases"])):
        w, h = (int(v) for v in g[f"size_{i}"])
        got = host.triangulate(g[f"pts_{i}"], w, h)
        assert got.shape == g[f"tri_{i}"].shape and (got == g[f"tri_{i}"]).all(), i


def test_triangulation_matches_golden_frames(native_lib):
    import glob
    for path in glob.glob(os.path.join(GOLDEN, "*_L*.npz")):
        if os.path.basename(path).startswith("chain_"):
            continue
        g = np.load(path)
        h, w = g["bgr1"].shape[:2]
        mp = host.morph_points(g["pts1"], g["pts2"], float(g["shape"]), w, h)
        assert bits_differ(mp, g["morphed_points"]) == 0
        assert (host.triangulate(mp, w, h) == g["tri_idx"]).all(), path


@pytest.mark.skipif(not ref.available(), reason="reference library not built")
def test_triangulation_matches_reference_random(native_lib):
    rng = np.random.default_rng(17)
    for trial in range(40):
        w, h, n = int(rng.integers(30, 700)), int(rng.integers(30, 500)), int(rng.integers(3, 300))
        if trial % 3 == 0:
            p = np.stack([rng.integers(0, w, n), rng.integers(0, h, n)], 1)
        elif trial % 3 == 1:
            p = np.stack([rng.uniform(-3, w - 0.01, n), rng.uniform(-3, h - 0.01, n)], 1)
        else:
            p = np.stack([rng.integers(0, 6, n) * (w // 6), rng.integers(0, 6, n) * (h // 6)], 1)
        p = p.astype(np.float32)
        a, b = ref.triangulate(w, h, p), host.triangulate(p, w, h)
        assert a.shape == b.shape and (a == b).all(), (trial, w, h, n)


def test_out_of_range_point_is_an_error_like_subdiv2d(native_lib):
    # clip_points lets x == cols through (src/util.cpp:455); cv::Subdiv2D then throws (subdivision2d.cpp:287)
    p = np.array([[0, 0], [10, 0], [0, 10], [20, 5]], np.float32)
    with pytest.raises(Exception):
        host.triangulate(p, 20, 20)


def test_morph_points_matches_port(native_lib):
    rng = np.random.default_rng(4)
    p1 = rng.uniform(-10, 330, (500, 2)).astype(np.float32)
    p2 = rng.uniform(-10, 330, (500, 2)).astype(np.float32)
    for s in (0.0, 1 / 3, 0.5, 0.999, 1.0, 1 / 59):
        assert bits_differ(host.morph_points(p1, p2, s, 320, 240), port.morph_points(p1, p2, np.float32(s), 320, 240)) == 0


def test_chain_schedule(native_lib):
    for n in (1, 2, 12, 60, 120):
        r = [host.chain_ratio(j, n) for j in range(n)]
        assert r[0] == 0.0
        for j in range(1, n):
            lin = j / float(n)
            assert r[j] == min(1.0, (1.0 / (1.0 - lin)) / n)
        if n > 1:
            assert abs(r[-1] - 1.0) < 1e-12


def test_sequence_plan_direct_and_chain(native_lib):
    inp = synth.make_inputs(200, 150, 80, 6.0, seed=3)
    phases = np.linspace(0, 1, 9).astype(np.float32)
    plan1 = host.SequencePlan(inp.pts1, inp.pts2, 200, 150, phases, chain=False, threads=1)
    plan4 = host.SequencePlan(inp.pts1, inp.pts2, 200, 150, phases, chain=False, threads=4)
    assert (plan1.tri_idx == plan4.tri_idx).all() and (plan1.tri_offsets == plan4.tri_offsets).all()
    for f, s in enumerate(phases):
        mp = host.morph_points(inp.pts1, inp.pts2, float(s), 200, 150)
        assert bits_differ(plan1.points(f), mp) == 0
        assert (plan1.triangles(f) == host.triangulate(mp, 200, 150)).all()
    ratios = np.array([host.chain_ratio(j, 9) for j in range(9)], np.float32)
    chain = host.SequencePlan(inp.pts1, inp.pts2, 200, 150, ratios, chain=True, threads=3)
    cur = inp.pts1
    for f in range(9):
        cur = host.morph_points(cur, inp.pts2, float(ratios[f]), 200, 150)
        assert bits_differ(chain.points(f), cur) == 0
        assert (chain.triangles(f) == host.triangulate(cur, 200, 150)).all()


@pytest.mark.skipif(not ref.available(), reason="reference library not built")
@pytest.mark.parametrize("kind", ["uniform", "lattice", "clustered"])
def test_triangulation_matches_reference_at_scale(native_lib, kind):
    """Long point-location walks (thousands of points): the dependence-cut walk loop (DelaunayMesh::walk_run) with its
    hand-over to the generic step at degenerate predicates must reproduce cv::Subdiv2D's triangle list exactly."""
    rng = np.random.default_rng(123)
    w, h, n = 1920, 1080, 6000
    if kind == "uniform":
        p = np.stack([rng.uniform(2, w - 3, n), rng.uniform(2, h - 3, n)], 1)
    elif kind == "lattice":        # collinear runs, points on edges, duplicates
        p = np.stack([rng.integers(0, 96, n) * 20, rng.integers(0, 54, n) * 20], 1)
    else:
        c = rng.uniform(100, 900, (12, 2))
        p = c[rng.integers(0, 12, n)] + rng.normal(0, 25, (n, 2))
        p = np.clip(p, 0, [w - 1, h - 1])
    p = p.astype(np.float32)
    a, b = ref.triangulate(w, h, p), host.triangulate(p, w, h)
    assert a.shape == b.shape and (a == b).all(), kind
