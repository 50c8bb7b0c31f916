"""Python mirror of the host stages (include/poppy_host.h): point hygiene, Delaunay topology in
cv::Subdiv2D::getTriangleList order, the chain schedule and the multi-threaded sequence planner. All of it runs in
the C++ part of libpoppy_cuda.so; nothing here computes."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import PoppyCudaError


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def _host_check(rc):
    if rc != 0:
        raise PoppyCudaError(rc, _lib.load().poppy_host_last_error().decode())


def morph_points(pts1, pts2, shape_ratio, width, height) -> np.ndarray:
    """clip_points on both sets, morph_points, clip_points (reference src/algo.cpp:185,191,202,205)."""
    pts1 = np.ascontiguousarray(pts1, np.float32)
    pts2 = np.ascontiguousarray(pts2, np.float32)
    out = np.empty_like(pts1)
    _host_check(_lib.load().poppy_host_morph_points(_ptr(pts1), _ptr(pts2), pts1.shape[0], float(shape_ratio),
                                                    width, height, _ptr(out)))
    return out


def triangulate(pts, width, height, sequential=False) -> np.ndarray:
    """Triangle vertex indices (T x 3) of the Delaunay mesh of `pts` in cv::Subdiv2D::getTriangleList order
    (reference src/algo.cpp:205-213). sequential: the caller walks through the frames of a sequence call by call; the
    previous call's point-location walks then predict this call's (same result, faster)."""
    pts = np.ascontiguousarray(pts, np.float32)
    cap = 2 * pts.shape[0] + 16
    tri = np.empty((cap, 3), np.int32)
    nt = C.c_int(0)
    lib = _lib.load()
    fn = lib.poppy_host_triangulate_next if sequential else lib.poppy_host_triangulate
    _host_check(fn(_ptr(pts), pts.shape[0], width, height, _ptr(tri), cap, C.byref(nt)))
    return tri[: nt.value].copy()


def chain_ratio(j, n_frames) -> float:
    """shape = color of frame j in the reference frame loop (src/poppy.hpp:181-210, phase < 0)."""
    return float(_lib.load().poppy_host_chain_ratio(j, n_frames))


class SequencePlan:
    """Morphed points + triangle lists of every frame of a sequence, planned on host threads."""

    def __init__(self, pts1, pts2, width, height, shape_ratio, chain=False, threads=0, copy=True):
        """copy=False: tri_idx / tri_offsets are views of the plan's own memory, valid until close() (a 600-frame plan at 4K
        holds 290 MB of triangle indices; the streaming callers hand them to the renderer and close the plan right after)."""
        self._lib = _lib.load()
        pts1 = np.ascontiguousarray(pts1, np.float32)
        pts2 = np.ascontiguousarray(pts2, np.float32)
        self.shape_ratio = np.ascontiguousarray(shape_ratio, np.float32)
        self.n = pts1.shape[0]
        self.frames = self.shape_ratio.shape[0]
        self._plan = C.c_void_p()
        _host_check(self._lib.poppy_host_plan_create(C.byref(self._plan), _ptr(pts1), _ptr(pts2), self.n, width, height,
                                                     self.frames, _ptr(self.shape_ratio), 1 if chain else 0, threads))
        tri, off, mx = C.c_void_p(), C.c_void_p(), C.c_int(0)
        _host_check(self._lib.poppy_host_plan_triangles(self._plan, C.byref(tri), C.byref(off), C.byref(mx)))
        self.max_triangles = mx.value
        self.tri_offsets = np.ctypeslib.as_array(C.cast(off, C.POINTER(C.c_int32)), shape=(self.frames + 1,)).copy()
        total = int(self.tri_offsets[-1])
        if total:
            view = np.ctypeslib.as_array(C.cast(tri, C.POINTER(C.c_int32)), shape=(total * 3,))
            self.tri_idx = (view.copy() if copy else view).reshape(-1, 3)
        else:
            self.tri_idx = np.zeros((0, 3), np.int32)

    def points(self, frame) -> np.ndarray:
        p = C.c_void_p()
        _host_check(self._lib.poppy_host_plan_points(self._plan, frame, C.byref(p)))
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_float)), shape=(self.n * 2,)).copy().reshape(-1, 2)

    def triangles(self, frame) -> np.ndarray:
        return self.tri_idx[self.tri_offsets[frame]: self.tri_offsets[frame + 1]]

    def close(self):
        if self._plan:
            self._lib.poppy_host_plan_destroy(self._plan)
            self._plan = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
