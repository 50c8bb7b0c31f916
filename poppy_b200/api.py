"""Python mirror of the reference interface for the morph path.

  Settings                the fields of poppy::Settings (reference src/settings.hpp:14-27); the path reads
                          pyramid_levels (src/algo.cpp:261)
  init(...)               poppy::init (reference src/poppy.hpp:30-44), same argument order
  morph_images(...)       poppy::morph_images (reference src/algo.hpp:26, src/algo.cpp:178-273): one frame, host
                          arrays in, (dst, morphedPoints) out
  morph_sequence(...)     the frame loop of poppy::morph<Twriter>() (reference src/poppy.hpp:177-243, phase < 0):
                          the chain recurrence kept resident on the GPU, frames handed to writer.write() in order
  render_phases(...)      direct mode: N independent phases of one pair ("-f 1 -p s", src/poppy.hpp:186-200)

Host stages (clip / uniq / Delaunay / index lookup) run in the C++ part of libpoppy_cuda.so, everything else in its
CUDA kernels. Nothing in this module computes pixels; without the native library or a GPU every call raises.
"""
from __future__ import annotations

from dataclasses import dataclass

import threading

import numpy as np

from . import host
from .renderer import MorphRenderer


@dataclass
class _Settings:
    show_gui: bool = False
    enable_wait: bool = False
    number_of_frames: float = 60
    frame_rate: float = 30
    match_tolerance: float = 1
    max_keypoints: int = 300
    pyramid_levels: int = 64
    enable_auto_align: bool = False
    enable_radial_mask: bool = False
    enable_face_detection: bool = False
    enable_denoise: bool = False
    enable_src_scaling: bool = False
    face_neighbors: int = 8
    fourcc: str = "FFV1"
    cuda_device: int = 0          # addition: the GPU the renderer runs on


class Settings:
    """Process-wide singleton, as poppy::Settings::instance()."""
    _instance = None

    @classmethod
    def instance(cls) -> _Settings:
        if cls._instance is None:
            cls._instance = _Settings()
        return cls._instance


def init(show_gui, number_of_frames, match_tolerance, auto_align, radial_mask, face_detect, denoise, src_scaling,
         frame_rate, pyramid_levels, fourcc, enable_wait, face_neighbors):
    s = Settings.instance()
    s.show_gui, s.enable_wait, s.number_of_frames, s.frame_rate = show_gui, enable_wait, number_of_frames, frame_rate
    s.match_tolerance, s.enable_auto_align, s.enable_radial_mask = match_tolerance, auto_align, radial_mask
    s.enable_denoise, s.enable_src_scaling, s.enable_face_detection = denoise, src_scaling, face_detect
    s.pyramid_levels, s.fourcc, s.face_neighbors = pyramid_levels, fourcc, face_neighbors


_cache: dict = {}


def _renderer(w, h, n_points, n_tri, n_frames) -> MorphRenderer:
    s = Settings.instance()
    key = (w, h, int(s.pyramid_levels), s.cuda_device)
    r = _cache.get(key)
    if r is None or r.max_points < n_points or r.max_triangles < n_tri or r.max_batch_frames < n_frames:
        if r is not None:
            r.close()
        r = MorphRenderer(w, h, int(s.pyramid_levels), max(n_points, 64), max(n_tri, 2 * n_points + 16), n_frames,
                          device=s.cuda_device)
        _cache.clear()
        _cache[key] = r
    return r


def release():
    for r in _cache.values():
        r.close()
    _cache.clear()


def morph_images(img1, img2, corrected1, corrected2, gabor2, src_points1, src_points2, shape_ratio, mask_ratio,
                 linear=0.0):
    """One frame. img1 supplies only the frame size; img2 and linear are unused, as in the reference.
    Returns (dst HxWx3 uint8, morphedPoints Nx2 float32)."""
    h, w = img1.shape[:2]
    p1 = np.ascontiguousarray(src_points1, np.float32)
    p2 = np.ascontiguousarray(src_points2, np.float32)
    if p1.shape != p2.shape:
        raise ValueError("point sets differ in size")          # assert at src/algo.cpp:51
    morphed = host.morph_points(p1, p2, shape_ratio, w, h)     # host copy, needed for the topology
    r = _renderer(w, h, len(p1), 2 * len(p1) + 16, 1)          # sized for the most triangles these points can give
    # the pair travels to the device (150 MB at 4K) while this thread triangulates; both are C calls that release the GIL
    failure = []

    def upload():
        try:
            r.set_pair(np.ascontiguousarray(corrected1), np.ascontiguousarray(corrected2), np.ascontiguousarray(gabor2))
            r.set_points(p1, p2)
        except Exception as e:      # re-raised on the calling thread
            failure.append(e)
    t = threading.Thread(target=upload)
    t.start()
    try:
        tri = host.triangulate(morphed, w, h, sequential=True)
    finally:
        t.join()
    if failure:
        raise failure[0]
    r.render([shape_ratio], [mask_ratio], tri, [0, len(tri)])
    dst = r.download(0, 1)[0]
    return dst, r.morphed_points(0)


def blur_margin(src, union_size):
    """poppy::blur_margin (reference src/util.cpp:574-602) on the GPU: `src` (H x W x 3 uint8) centred on a black canvas of
    union_size = (width, height) with its four margins Gaussian-blurred. Returns the canvas."""
    from . import _lib
    import ctypes as C
    src = np.ascontiguousarray(src, np.uint8)
    if src.ndim != 3 or src.shape[2] != 3:
        raise ValueError("blur_margin: expected an H x W x 3 uint8 image")
    uw, uh = int(union_size[0]), int(union_size[1])
    dst = np.empty((max(uh, 0), max(uw, 0), 3), np.uint8)
    lib = _lib.load()
    rc = lib.poppy_cuda_blur_margin(Settings.instance().cuda_device, src.ctypes.data_as(C.c_void_p), src.strides[0], src.shape[1],
                                    src.shape[0], uw, uh, dst.ctypes.data_as(C.c_void_p), dst.strides[0] if dst.size else 0)
    if rc != 0:
        raise RuntimeError(f"poppy_cuda_blur_margin failed ({rc}): {lib.poppy_cuda_blur_margin_last_error().decode()}")
    return dst


def gabor_filter(src):
    """poppy::gabor_filter(src, dst) with the reference's defaults (src/util.cpp:40-60) on the GPU; src: H x W x 3 float32
    (the reference passes corrected2 / 255). Agrees with the reference to float rounding (see include/poppy_cuda.h)."""
    from . import _lib
    import ctypes as C
    src = np.ascontiguousarray(src, np.float32)
    if src.ndim != 3 or src.shape[2] != 3:
        raise ValueError("gabor_filter: expected an H x W x 3 float32 image")
    dst = np.empty_like(src)
    lib = _lib.load()
    rc = lib.poppy_cuda_gabor_filter(Settings.instance().cuda_device, src.ctypes.data_as(C.c_void_p), src.strides[0], src.shape[1],
                                     src.shape[0], dst.ctypes.data_as(C.c_void_p), dst.strides[0])
    if rc != 0:
        raise RuntimeError(f"poppy_cuda_gabor_filter failed ({rc}): {lib.poppy_cuda_blur_margin_last_error().decode()}")
    return dst


def morph_sequence(corrected1, corrected2, gabor2, src_points1, src_points2, writer=None, number_of_frames=None,
                   threads=0):
    """Chain mode: frame j = morph_images(previous frame, previous morphed points, shape = color = 1/(N-j)).
    Returns the frames (N x H x W x 3) and calls writer.write(frame) per frame if a writer is given."""
    n_frames = int(number_of_frames if number_of_frames is not None else Settings.instance().number_of_frames)
    h, w = corrected1.shape[:2]
    ratio = np.array([host.chain_ratio(j, n_frames) for j in range(n_frames)], np.float64)
    plan = host.SequencePlan(src_points1, src_points2, w, h, ratio.astype(np.float32), chain=True, threads=threads)
    try:
        r = _renderer(w, h, plan.n, plan.max_triangles, n_frames)
        r.set_pair(np.ascontiguousarray(corrected1), np.ascontiguousarray(corrected2), np.ascontiguousarray(gabor2))
        r.set_points(src_points1, src_points2)
        r.render(ratio.astype(np.float32), ratio, plan.tri_idx, plan.tri_offsets, chain=True)
        frames = r.download(0, n_frames)
    finally:
        plan.close()
    if writer is not None:
        for f in frames:
            writer.write(f)
    return frames


def render_phases(corrected1, corrected2, gabor2, src_points1, src_points2, phases, mask_ratios=None, threads=0):
    """Direct mode: independent phases s_k of one pair, shape = s_k and mask = mask_ratios[k] (default s_k)."""
    phases = np.ascontiguousarray(phases, np.float32)
    masks = np.ascontiguousarray(mask_ratios if mask_ratios is not None else phases.astype(np.float64), np.float64)
    h, w = corrected1.shape[:2]
    plan = host.SequencePlan(src_points1, src_points2, w, h, phases, chain=False, threads=threads)
    try:
        r = _renderer(w, h, plan.n, plan.max_triangles, len(phases))
        r.set_pair(np.ascontiguousarray(corrected1), np.ascontiguousarray(corrected2), np.ascontiguousarray(gabor2))
        r.set_points(src_points1, src_points2)
        r.render(phases, masks, plan.tri_idx, plan.tri_offsets, chain=False)
        return r.download(0, len(phases))
    finally:
        plan.close()
