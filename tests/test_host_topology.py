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


def _points(kind, rng, w, h, n):
    if kind == "uniform":
        p = np.stack([rng.uniform(2, w - 3, n), rng.uniform(2, h - 3, n)], 1)
    elif kind == "lattice":        # collinear runs, points on edges, duplicates: the degenerate predicates
        p = np.stack([rng.integers(0, w // 20, n) * 20, rng.integers(0, h // 20, n) * 20], 1)
    else:
        c = rng.uniform(0.1 * w, 0.9 * w, (12, 2)) * [1, h / w]
        p = np.clip(c[rng.integers(0, 12, n)] + rng.normal(0, 25, (n, 2)), 0, [w - 1, h - 1])
    return p.astype(np.float32)


@pytest.mark.parametrize("kind", ["uniform", "lattice", "clustered"])
@pytest.mark.parametrize("jitter", [0.02, 1.5, 40.0])
def test_guided_walks_give_the_unguided_triangulation(native_lib, kind, jitter):
    """The sequence planner predicts each frame's point-location walks from the previous frame's (poppy::WalkTrace,
    DelaunayMesh::walk_guided): whatever the guide holds - a neighbouring frame, a frame whose points moved far, the
    previous CALL's unrelated point set - every frame's triangle list must be the one an unguided insertion gives."""
    rng = np.random.default_rng(7)
    w, h, n = 1280, 720, 3000
    p1 = _points(kind, rng, w, h, n)
    p2 = np.clip(p1 + rng.uniform(-jitter, jitter, p1.shape), 0, [w - 1, h - 1]).astype(np.float32)
    phases = np.linspace(0, 1, 7).astype(np.float32)
    # one thread: frames 1.. are guided by their predecessor; the first frame by what the previous test case left in the pool
    plan = host.SequencePlan(p1, p2, w, h, phases, chain=False, threads=1)
    for f, s in enumerate(phases):
        mp = host.morph_points(p1, p2, float(s), w, h)
        want = host.triangulate(mp, w, h)
        got = plan.triangles(f)
        assert got.shape == want.shape and (got == want).all(), (kind, jitter, f)
    plan.close()
    # a stale guide of a different size: a second, smaller, unrelated sequence planned right after
    q1 = _points("uniform", rng, w, h, 700)
    plan = host.SequencePlan(q1, q1[::-1].copy(), w, h, phases[:3], chain=False, threads=1)
    for f, s in enumerate(phases[:3]):
        mp = host.morph_points(q1, q1[::-1].copy(), float(s), w, h)
        assert (plan.triangles(f) == host.triangulate(mp, w, h)).all(), (kind, jitter, f, "stale guide")
    plan.close()


@pytest.mark.parametrize("threads", [2, 5, 16])
def test_paced_planner_threads_follow_each_other(native_lib, threads):
    """With several workers frame f is triangulated a few points behind frame f - 1 on another thread and predicted by the
    trace that thread is still recording (poppy::WalkPace). Repeated to give a race a chance to show."""
    rng = np.random.default_rng(threads)
    w, h, n = 960, 540, 1500
    for rep in range(3):
        p1 = _points(["uniform", "lattice", "clustered"][rep], rng, w, h, n)
        p2 = np.clip(p1 + rng.uniform(-3, 3, p1.shape), 0, [w - 1, h - 1]).astype(np.float32)
        phases = np.linspace(0, 1, 37).astype(np.float32)
        plan = host.SequencePlan(p1, p2, w, h, phases, chain=False, threads=threads)
        for f in range(0, 37, 3):
            mp = host.morph_points(p1, p2, float(phases[f]), w, h)
            want = host.triangulate(mp, w, h)
            assert plan.triangles(f).shape == want.shape and (plan.triangles(f) == want).all(), (threads, rep, f)
        plan.close()


def test_sequential_triangulate_matches_plain(native_lib):
    rng = np.random.default_rng(5)
    w, h = 800, 600
    p1 = _points("uniform", rng, w, h, 900)
    p2 = np.clip(p1 + rng.uniform(-5, 5, p1.shape), 0, [w - 1, h - 1]).astype(np.float32)
    for s in np.linspace(0, 1, 6):
        mp = host.morph_points(p1, p2, float(s), w, h)
        assert (host.triangulate(mp, w, h, sequential=True) == host.triangulate(mp, w, h)).all()
    other = _points("lattice", rng, w, h, 400)           # an unrelated point set right after: stale guide
    assert (host.triangulate(other, w, h, sequential=True) == host.triangulate(other, w, h)).all()


@pytest.mark.skipif(not ref.available(), reason="reference library not built")
def test_guided_planner_matches_reference_subdiv2d(native_lib):
    rng = np.random.default_rng(11)
    w, h, n = 1920, 1080, 5000
    p1 = _points("uniform", rng, w, h, n)
    p2 = np.clip(p1 + rng.uniform(-6, 6, p1.shape), 0, [w - 1, h - 1]).astype(np.float32)
    phases = np.linspace(0.3, 0.32, 5).astype(np.float32)
    plan = host.SequencePlan(p1, p2, w, h, phases, chain=False, threads=1)
    for f, s in enumerate(phases):
        mp = host.morph_points(p1, p2, float(s), w, h)
        a = ref.triangulate(w, h, mp)
        assert a.shape == plan.triangles(f).shape and (a == plan.triangles(f)).all(), f
