"""Deterministic synthetic inputs for the morph-render path (SURVEY.md §8(d) configs 4/5 and the 1080p point).

Pure numpy/scipy so the same bytes are produced in the build container and on the GPU box; the reference oracle
and the CUDA path are always fed the *same arrays*, so the generator only needs to be deterministic, not identical
to cv::RNG. Images are uniform u8 noise blurred with a Gaussian (smooth but textured), gabor2 is a smooth field
in [0, 1] (stand-in for the reference's Gabor response, src/util.cpp:40-60), points are uniform in the frame with
a bounded jitter for set 2 plus the four frame corners (reference add_corners, src/util.cpp:268-279).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np


@dataclass
class MorphInputs:
    bgr1: np.ndarray      # H x W x 3 uint8
    bgr2: np.ndarray      # H x W x 3 uint8
    gabor2: np.ndarray    # H x W x 3 float32 in [0, 1]
    pts1: np.ndarray      # N x 2 float32 (x, y)
    pts2: np.ndarray      # N x 2 float32

    @property
    def width(self):
        return self.bgr1.shape[1]

    @property
    def height(self):
        return self.bgr1.shape[0]


def _blur(a: np.ndarray, sigma: float) -> np.ndarray:
    from scipy.ndimage import gaussian_filter
    out = np.empty_like(a, dtype=np.float32)
    for c in range(a.shape[2]):
        out[..., c] = gaussian_filter(a[..., c].astype(np.float32), sigma, mode="mirror")
    return out


def noise_image(w: int, h: int, seed: int, sigma: float = 3.0) -> np.ndarray:
    rng = np.random.Generator(np.random.PCG64(seed))
    raw = rng.integers(0, 256, size=(h, w, 3), dtype=np.uint8)
    sm = _blur(raw, sigma)
    # stretch the blurred noise back over the 8-bit range so the frames carry texture
    sm = (sm - 127.5) * (sigma * 2.5) + 127.5
    return np.clip(np.rint(sm), 0, 255).astype(np.uint8)


def smooth_field(w: int, h: int, seed: int, sigma: float = 6.0) -> np.ndarray:
    rng = np.random.Generator(np.random.PCG64(seed))
    raw = rng.random(size=(h, w, 3), dtype=np.float32)
    sm = _blur(raw, sigma)
    sm = (sm - 0.5) * (sigma * 2.0) + 0.5
    return np.clip(sm, 0.0, 1.0).astype(np.float32)


def matched_points(w: int, h: int, n: int, jitter: float, seed: int):
    rng = np.random.Generator(np.random.PCG64(seed))
    p1 = np.empty((n, 2), np.float32)
    p1[:, 0] = rng.uniform(2, w - 3, n).astype(np.float32)
    p1[:, 1] = rng.uniform(2, h - 3, n).astype(np.float32)
    p2 = p1 + rng.uniform(-jitter, jitter, size=(n, 2)).astype(np.float32)
    p2[:, 0] = np.clip(p2[:, 0], 0, w - 1)
    p2[:, 1] = np.clip(p2[:, 1], 0, h - 1)
    corners = np.array([[0, 0], [w - 1, 0], [0, h - 1], [w - 1, h - 1]], np.float32)
    return (np.concatenate([p1, corners]).astype(np.float32), np.concatenate([p2.astype(np.float32), corners]))


def make_inputs(w: int, h: int, n_points: int, jitter: float = 8.0, seed: int = 1234) -> MorphInputs:
    pts1, pts2 = matched_points(w, h, n_points, jitter, seed + 3)
    return MorphInputs(bgr1=noise_image(w, h, seed), bgr2=noise_image(w, h, seed + 1),
                       gabor2=smooth_field(w, h, seed + 2), pts1=pts1, pts2=pts2)


# Named workloads (BASELINE.json configs 4/5 and the 1080p point of the metric)
WORKLOADS = {
    "1080p": dict(w=1920, h=1080, n_points=5000, jitter=8.0, seed=1080, frames=600, levels=6),
    "4k": dict(w=3840, h=2160, n_points=20000, jitter=8.0, seed=1234, frames=600, levels=6),
    "8k": dict(w=7680, h=4320, n_points=50000, jitter=16.0, seed=4321, frames=2400, levels=6),
}


def workload_inputs(name: str) -> MorphInputs:
    c = WORKLOADS[name]
    return make_inputs(c["w"], c["h"], c["n_points"], c["jitter"], c["seed"])


def shape_image(w: int, h: int, kind: str, seed: int = 0) -> np.ndarray:
    """Hard-edged synthetic picture (a filled square or disc on a flat background, mild noise) — the analogue of
    the reference's square.png / circle.png demo pair; sharp edges drive the unsharp-mask threshold branch."""
    rng = np.random.Generator(np.random.PCG64(seed))
    yy, xx = np.mgrid[0:h, 0:w]
    cx, cy, r = w / 2.0, h / 2.0, min(w, h) * 0.3
    if kind == "square":
        inside = (np.abs(xx - cx) <= r) & (np.abs(yy - cy) <= r)
        fg, bg = (30, 40, 220), (235, 235, 235)
    else:
        inside = (xx - cx) ** 2 + (yy - cy) ** 2 <= r * r
        fg, bg = (220, 60, 30), (20, 20, 20)
    img = np.where(inside[..., None], np.array(fg, np.float32), np.array(bg, np.float32))
    img += rng.normal(0, 2.0, size=img.shape).astype(np.float32)
    return np.clip(np.rint(img), 0, 255).astype(np.uint8)


def shape_inputs(w: int, h: int, n_points: int = 40, seed: int = 11) -> MorphInputs:
    """square -> disc pair with points on the two outlines (matched by angle) plus random interior points."""
    rng = np.random.Generator(np.random.PCG64(seed))
    cx, cy, r = w / 2.0, h / 2.0, min(w, h) * 0.3
    ang = np.linspace(0, 2 * np.pi, n_points, endpoint=False)
    circ = np.stack([cx + r * np.cos(ang), cy + r * np.sin(ang)], 1)
    t = np.maximum(np.abs(np.cos(ang)), np.abs(np.sin(ang)))
    sq = np.stack([cx + r * np.cos(ang) / t, cy + r * np.sin(ang) / t], 1)
    extra = np.stack([rng.uniform(1, w - 2, n_points // 2), rng.uniform(1, h - 2, n_points // 2)], 1)
    corners = np.array([[0, 0], [w - 1, 0], [0, h - 1], [w - 1, h - 1]], np.float32)
    p1 = np.concatenate([sq, extra, corners]).astype(np.float32)
    p2 = np.concatenate([circ, extra + rng.uniform(-3, 3, extra.shape), corners]).astype(np.float32)
    p2[:, 0] = np.clip(p2[:, 0], 0, w - 1)
    p2[:, 1] = np.clip(p2[:, 1], 0, h - 1)
    return MorphInputs(bgr1=shape_image(w, h, "square", seed), bgr2=shape_image(w, h, "disc", seed + 1),
                       gabor2=smooth_field(w, h, seed + 2, sigma=4.0), pts1=p1, pts2=p2)


def block_inputs(w: int, h: int, n_points: int = 30, block: int = 5, seed: int = 51, jitter: float = 2.0) -> MorphInputs:
    """Random saturated colour blocks: after the pyramid blend many pixels differ from their Gaussian blur by more
    than the unsharp threshold, so the sharpening branch of unsharp_mask (src/util.cpp:135-145) is exercised."""
    rng = np.random.Generator(np.random.PCG64(seed))

    def blocks(s):
        r = np.random.Generator(np.random.PCG64(s))
        small = r.integers(0, 2, size=((h + block - 1) // block, (w + block - 1) // block, 3), dtype=np.uint8) * 255
        return np.ascontiguousarray(np.kron(small, np.ones((block, block, 1), np.uint8))[:h, :w])

    pts1, pts2 = matched_points(w, h, n_points, jitter, seed + 3)
    g = rng.random(size=(h, w, 3), dtype=np.float32)
    return MorphInputs(bgr1=blocks(seed), bgr2=blocks(seed + 1), gabor2=g, pts1=pts1, pts2=pts2)
