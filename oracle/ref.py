"""TEST INFRASTRUCTURE ONLY — ctypes binding of oracle/_ref/libpoppy_ref.so.

libpoppy_ref.so is the *unmodified* reference morph path (reference src/algo.cpp, util.cpp, draw.cpp, settings.cpp
plus the vendored OpenCV 4.6.0 core+imgproc), compiled in place by oracle/build_ref.sh. It is the parity pin for the
whole repo. Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
module; nothing under poppy_b200/ does.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libpoppy_ref.so")

_lib = None


class _StageDump(C.Structure):
    _fields_ = [
        ("morphed_points", C.c_void_p), ("tri_idx", C.c_void_p), ("max_tri", C.c_int32), ("n_tri", C.c_int32),
        ("tri_map", C.c_void_p), ("hom", C.c_void_p), ("m1", C.c_void_p), ("m2", C.c_void_p),
        ("mapx1", C.c_void_p), ("mapy1", C.c_void_p), ("mapx2", C.c_void_p), ("mapy2", C.c_void_p),
        ("warped1", C.c_void_p), ("warped2", C.c_void_p), ("mask", C.c_void_p), ("lap_blend", C.c_void_p),
        ("dst", C.c_void_p),
    ]


def available() -> bool:
    return os.path.exists(LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise RuntimeError(f"{LIB_PATH} missing: run oracle/build_ref.sh where /root/reference exists")
        _lib = C.CDLL(LIB_PATH)
        _lib.poppy_ref_last_error.restype = C.c_char_p
        _lib.poppy_ref_opencv_version.restype = C.c_char_p
    return _lib


def _check(rc):
    if rc != 0:
        raise RuntimeError("reference oracle failed: " + lib().poppy_ref_last_error().decode(errors="replace"))


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _u8(a):
    return np.ascontiguousarray(a, dtype=np.uint8)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def opencv_version() -> str:
    return lib().poppy_ref_opencv_version().decode()


def set_threads(n: int):
    lib().poppy_ref_set_threads(C.c_int(n))


def get_threads() -> int:
    return int(lib().poppy_ref_get_threads())


def morph_images(bgr1, bgr2, gabor2, pts1, pts2, shape, mask, levels):
    """poppy::morph_images() itself (reference src/algo.cpp:178-273). Returns (dst u8 HxWx3, morphed points Nx2)."""
    bgr1, bgr2, gabor2, pts1, pts2 = _u8(bgr1), _u8(bgr2), _f32(gabor2), _f32(pts1), _f32(pts2)
    h, w = bgr1.shape[:2]
    n = pts1.shape[0]
    dst = np.empty((h, w, 3), np.uint8)
    mp = np.empty((n, 2), np.float32)
    _check(lib().poppy_ref_morph_images(w, h, _p(bgr1), _p(bgr2), _p(gabor2), _p(pts1), _p(pts2), n,
                                        C.c_double(shape), C.c_double(mask), int(levels), _p(dst), _p(mp)))
    return dst, mp


def chain(bgr1, bgr2, gabor2, pts1, pts2, n_frames, levels):
    """The frame-loop recurrence of reference src/poppy.hpp:177-243 (phase < 0). Returns (frames, points)."""
    bgr1, bgr2, gabor2, pts1, pts2 = _u8(bgr1), _u8(bgr2), _f32(gabor2), _f32(pts1), _f32(pts2)
    h, w = bgr1.shape[:2]
    n = pts1.shape[0]
    frames = np.empty((n_frames, h, w, 3), np.uint8)
    pts = np.empty((n_frames, n, 2), np.float32)
    _check(lib().poppy_ref_chain(w, h, _p(bgr1), _p(bgr2), _p(gabor2), _p(pts1), _p(pts2), n, int(n_frames),
                                 int(levels), _p(frames), _p(pts)))
    return frames, pts


def triangulate(w, h, pts):
    """clip -> make_uniq -> cv::Subdiv2D (4.6.0) -> get_triangle_indices (reference src/algo.cpp:205-213)."""
    pts = _f32(pts)
    n = pts.shape[0]
    cap = 2 * n + 16
    idx = np.empty((cap, 3), np.int32)
    nt = C.c_int(0)
    _check(lib().poppy_ref_triangulate(w, h, _p(pts), n, _p(idx), cap, C.byref(nt)))
    assert nt.value <= cap
    return idx[: nt.value].copy()


@dataclass
class Stages:
    morphed_points: np.ndarray
    tri_idx: np.ndarray
    tri_map: np.ndarray
    hom: np.ndarray
    m1: np.ndarray
    m2: np.ndarray
    mapx1: np.ndarray
    mapy1: np.ndarray
    mapx2: np.ndarray
    mapy2: np.ndarray
    warped1: np.ndarray
    warped2: np.ndarray
    mask: np.ndarray
    lap_blend: np.ndarray
    dst: np.ndarray


def stages(bgr1, bgr2, gabor2, pts1, pts2, shape, mask, levels) -> Stages:
    """Every stage boundary of SURVEY.md §3.2, produced by the reference's own helpers called in order."""
    bgr1, bgr2, gabor2, pts1, pts2 = _u8(bgr1), _u8(bgr2), _f32(gabor2), _f32(pts1), _f32(pts2)
    h, w = bgr1.shape[:2]
    n = pts1.shape[0]
    cap = 2 * n + 16
    f = np.float32
    out = dict(
        morphed_points=np.empty((n, 2), f), tri_idx=np.empty((cap, 3), np.int32), tri_map=np.empty((h, w), np.int32),
        hom=np.empty((cap, 3, 3), f), m1=np.empty((cap, 3, 3), f), m2=np.empty((cap, 3, 3), f),
        mapx1=np.empty((h, w), f), mapy1=np.empty((h, w), f), mapx2=np.empty((h, w), f), mapy2=np.empty((h, w), f),
        warped1=np.empty((h, w, 3), np.uint8), warped2=np.empty((h, w, 3), np.uint8), mask=np.empty((h, w), f),
        lap_blend=np.empty((h, w, 3), f), dst=np.empty((h, w, 3), np.uint8))
    d = _StageDump()
    for k, v in out.items():
        setattr(d, k, v.ctypes.data)
    d.max_tri = cap
    _check(lib().poppy_ref_stages(w, h, _p(bgr1), _p(bgr2), _p(gabor2), _p(pts1), _p(pts2), n, C.c_double(shape),
                                  C.c_double(mask), int(levels), C.byref(d)))
    t = d.n_tri
    for k in ("tri_idx", "hom", "m1", "m2"):
        out[k] = out[k][:t].copy()
    return Stages(**out)


def fill_triangles(w, h, tri_xy):
    """paint_triangles (reference src/algo.cpp:95-106): cv::fillConvexPoly of triangle i with value i+1, in order."""
    tri_xy = np.ascontiguousarray(tri_xy, dtype=np.int32).reshape(-1, 6)
    img = np.zeros((h, w), np.int32)
    _check(lib().poppy_ref_fill_triangles(_p(img), w, h, _p(tri_xy), tri_xy.shape[0]))
    return img


def remap_u8c3(src, mapx, mapy):
    src, mapx, mapy = _u8(src), _f32(mapx), _f32(mapy)
    h, w = src.shape[:2]
    dh, dw = mapx.shape
    dst = np.empty((dh, dw, 3), np.uint8)
    _check(lib().poppy_ref_remap_u8c3(_p(src), w, h, _p(mapx), _p(mapy), dw, dh, _p(dst)))
    return dst


def _cn(a):
    return 1 if a.ndim == 2 else a.shape[2]


def pyr_down(src, dsize=None):
    src = _f32(src)
    h, w = src.shape[:2]
    dw, dh = dsize if dsize else ((w + 1) // 2, (h + 1) // 2)
    dst = np.empty((dh, dw) + src.shape[2:], np.float32)
    _check(lib().poppy_ref_pyr_down(_p(src), w, h, _cn(src), _p(dst), dw, dh))
    return dst


def pyr_up(src, dsize):
    src = _f32(src)
    h, w = src.shape[:2]
    dw, dh = dsize
    dst = np.empty((dh, dw) + src.shape[2:], np.float32)
    _check(lib().poppy_ref_pyr_up(_p(src), w, h, _cn(src), _p(dst), dw, dh))
    return dst


def lap_blend(l, r, mask, levels):
    l, r, mask = _f32(l), _f32(r), _f32(mask)
    h, w = mask.shape
    out = np.empty((h, w, 3), np.float32)
    _check(lib().poppy_ref_lap_blend(_p(l), _p(r), _p(mask), w, h, int(levels), _p(out)))
    return out


def unsharp(src, radius, amount, threshold):
    src = _f32(src)
    h, w = src.shape[:2]
    out = np.empty_like(src)
    _check(lib().poppy_ref_unsharp(_p(src), w, h, C.c_float(radius), C.c_float(amount), C.c_float(threshold), _p(out)))
    return out


def gaussian_blur(src, sigma):
    src = _f32(src)
    h, w = src.shape[:2]
    out = np.empty_like(src)
    _check(lib().poppy_ref_gaussian_blur(_p(src), w, h, _cn(src), C.c_double(sigma), _p(out)))
    return out


def median3(src):
    src = _f32(src)
    h, w = src.shape[:2]
    out = np.empty_like(src)
    _check(lib().poppy_ref_median3(_p(src), w, h, _cn(src), _p(out)))
    return out


def mask(gabor2, mask_ratio):
    gabor2 = _f32(gabor2)
    h, w = gabor2.shape[:2]
    out = np.empty((h, w), np.float32)
    _check(lib().poppy_ref_mask(_p(gabor2), w, h, C.c_double(mask_ratio), _p(out)))
    return out


def blur_margin(src, union_size):
    """poppy::blur_margin (reference src/util.cpp:574-602); union_size = (width, height)."""
    src = _u8(src)
    h, w = src.shape[:2]
    uw, uh = int(union_size[0]), int(union_size[1])
    out = np.empty((uh, uw, 3), np.uint8)
    _check(lib().poppy_ref_blur_margin(_p(src), w, h, uw, uh, _p(out)))
    return out


def gaussian_blur_u8(src, ksize, sigma):
    src = _u8(src)
    h, w = src.shape[:2]
    out = np.empty_like(src)
    _check(lib().poppy_ref_gaussian_blur_u8(_p(src), w, h, int(ksize), C.c_double(sigma), _p(out)))
    return out


def gabor_filter(src):
    """poppy::gabor_filter(src, dst) with the reference's defaults (src/util.cpp:40-60, called at src/poppy.hpp:122)."""
    src = _f32(src)
    h, w = src.shape[:2]
    out = np.empty_like(src)
    _check(lib().poppy_ref_gabor_filter(_p(src), w, h, _p(out)))
    return out


def gabor_kernel(ksize, sigma, theta, lambd, gamma, psi):
    out = np.empty((ksize, ksize), np.float32)
    _check(lib().poppy_ref_gabor_kernel(int(ksize), C.c_double(sigma), C.c_double(theta), C.c_double(lambd), C.c_double(gamma),
                                        C.c_double(psi), _p(out)))
    return out
