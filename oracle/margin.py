"""TEST INFRASTRUCTURE — CPU restatement (numpy, integer arithmetic) of the input conditioning that precedes the morph path
once per pair: poppy::blur_margin (reference src/util.cpp:574-602) and the 8-bit cv::GaussianBlur it is made of.

OpenCV 4.6.0 runs an 8-bit GaussianBlur of an isolated image through its bit-exact fixed-point path
(OCV imgproc/src/smooth.dispatch.cpp:654-684):
  * kernel: getGaussianKernelBitExact (:82-198, softdouble = IEEE double operations) scaled to 8 fractional bits with error
    diffusion, the centre tap taking what is left of 256 (getGaussianKernelFixedPoint_ED, :224-259);
  * rows:    h[x] = sum_k m[k] * src[reflect101(x + k - r)]      16-bit (hlineSmoothONa_yzy_a, smooth.simd.hpp:1136-1199)
  * columns: v = sum_k m[k] * h[reflect101(y + k - r)]           32-bit, out = (v + 2^15) >> 16 (vlineSmoothONa_yzy_a :1780-1866,
             ufixedpoint32 -> uint8_t rounding of fixedpoint.inl.hpp)
The sums are integers, so their order is immaterial; no tap sum can overflow (the taps add up to 256).
Pinned against the reference library (tests/test_margin.py): kernel taps by impulse response, images of every
awkward size, and blur_margin itself. Only tests may import this module."""
from __future__ import annotations

import math

import numpy as np


def gaussian_kernel_fixed(n: int, sigma: float, bits: int = 8) -> list[int]:
    """Taps of the n-tap Gaussian in `bits`-bit fixed point (sigma > 0, n odd)."""
    scale2x = -0.125 / (sigma * sigma)
    n2 = (n - 1) // 2
    vals = [math.exp(float(x * x) * scale2x) for x in range(1 - n, 1 - n + 2 * n2, 2)]
    s = 0.0
    for t in vals:
        s += t
    s *= 2.0
    s += 1.0
    mul1 = 1.0 / s
    kb = [t * mul1 for t in vals] + [1.0 * mul1]
    res = [0] * n
    err, total = 0.0, 0
    for i in range(n2):
        adj = kb[i] * float(1 << bits) + err
        v0 = int(np.rint(adj))          # cvRound: half to even
        err = adj - v0
        res[i] = res[n - 1 - i] = v0
        total += v0
    res[n2] = (1 << bits) - 2 * total
    return res


def reflect101(p: int, length: int) -> int:
    """cv::borderInterpolate(p, length, BORDER_REFLECT_101) (OCV core/src/copy.cpp)."""
    if length == 1:
        return 0
    while p < 0 or p >= length:
        p = -p if p < 0 else 2 * (length - 1) - p
    return p


def gaussian_blur_u8(img: np.ndarray, ksize: int, sigma: float) -> np.ndarray:
    """cv::GaussianBlur(img, (ksize, ksize), sigma) of an isolated 8-bit image (H x W x C)."""
    h, w = img.shape[:2]
    kx = gaussian_kernel_fixed(ksize, sigma) if w > 1 else [256]       # a 1-pixel axis shrinks its kernel to [1]
    ky = gaussian_kernel_fixed(ksize, sigma) if h > 1 else [256]
    src = img.astype(np.uint32)
    rx, ry = len(kx) // 2, len(ky) // 2
    ix = np.array([reflect101(x, w) for x in range(-rx, w + rx)])
    iy = np.array([reflect101(y, h) for y in range(-ry, h + ry)])
    rows = np.zeros_like(src)
    for k, m in enumerate(kx):
        if m:
            rows += m * src[:, ix[k:k + w]]
    assert rows.max(initial=0) <= 0xFFFF
    out = np.zeros_like(src)
    for k, m in enumerate(ky):
        if m:
            out += m * rows[iy[k:k + h]]
    return np.minimum((out + (1 << 15)) >> 16, 255).astype(np.uint8)


def margin_rects(cols: int, rows: int, union_w: int, union_h: int):
    """The source ROI and the four margin ROIs (x, y, w, h) of blur_margin, with its double -> int truncations."""
    margin_factor = 1.3
    margin = (cols + rows) / 100.0
    dx = abs(cols - union_w) / 2.0
    dy = abs(rows - union_h) / 2.0
    roi = (int(dx), int(dy), cols, rows)
    dx = margin_factor if dx == 0 else dx + margin
    dy = margin_factor if dy == 0 else dy + margin
    return roi, [(0, 0, int(dx), union_h), (int(union_w - dx), 0, int(dx), union_h),
                 (0, 0, union_w, int(dy)), (0, int(union_h - dy), union_w, int(dy))]


def blur_margin(src: np.ndarray, union_size) -> np.ndarray:
    """poppy::blur_margin: src centred on a black union-sized canvas; left, right, top, bottom margins blurred (all four from
    the unblurred canvas, written in that order)."""
    union_w, union_h = int(union_size[0]), int(union_size[1])
    rows, cols = src.shape[:2]
    (rx, ry, rw, rh), margins = margin_rects(cols, rows, union_w, union_h)
    canvas = np.zeros((union_h, union_w, 3), np.uint8)
    canvas[ry:ry + rh, rx:rx + rw] = src
    out = canvas.copy()
    for x, y, w, h in margins:
        out[y:y + h, x:x + w] = gaussian_blur_u8(canvas[y:y + h, x:x + w], 127, 6.0)
    return out


# ---- gabor_filter (reference src/util.cpp:40-60, called with its defaults at src/poppy.hpp:122) ------------------------------
# dst = (1/16) * sum_i clamp01(filter2D(src, getGaborKernel(13 x 13, sigma 5, theta_i, lambda 10, gamma 0.04, psi pi/4)))
# with theta_i = i * float(180 / 16) = 11 i (degrees handed to a function that takes radians - reproduced as is).
# cv::filter2D routes a 13 x 13 float kernel over a float image through crossCorr (OCV imgproc/src/filter.dispatch.cpp:1291:
# kernel area 169 >= 130), which promotes 32-bit float images to DOUBLE (templmatch.cpp:592 maxDepth), correlates block-wise by
# DFT and rounds the result back to float. This restatement evaluates the same correlation directly in double
# (BORDER_REFLECT_101, anchor at the centre): it agrees with the DFT evaluation to ~1e-15 relative, i.e. the float results are
# equal except where the exact value lies within that distance of a float rounding boundary (a few values per hundred
# million, off by one float ulp). FLOATING-POINT PARITY WITH TOLERANCE, stated in tests/test_margin.py - not bit-exact.
GABOR = dict(angles=16, ksize=13, sigma=5.0, lambd=10.0, gamma=0.04, psi=math.pi / 4)


def gabor_kernel(ksize: int, sigma: float, theta: float, lambd: float, gamma: float, psi: float) -> np.ndarray:
    """cv::getGaborKernel(..., CV_32F) (OCV imgproc/src/gabor.cpp:51-95)."""
    sigma_x, sigma_y = sigma, sigma / gamma
    c, s = math.cos(theta), math.sin(theta)
    xmax = ymax = ksize // 2
    ex, ey = -0.5 / (sigma_x * sigma_x), -0.5 / (sigma_y * sigma_y)
    cscale = math.pi * 2 / lambd
    k = np.empty((2 * ymax + 1, 2 * xmax + 1), np.float32)
    for y in range(-ymax, ymax + 1):
        for x in range(-xmax, xmax + 1):
            xr = x * c + y * s
            yr = -x * s + y * c
            k[ymax - y, xmax - x] = np.float32(1.0 * math.exp(ex * xr * xr + ey * yr * yr) * math.cos(cscale * xr + psi))
    return k


def gabor_thetas() -> list[float]:
    step = float(np.float32(180 // GABOR["angles"]))          # float step = (180 / numAngles): integer division
    return [i * step for i in range(GABOR["angles"])]


def gabor_filter(src: np.ndarray) -> np.ndarray:
    from scipy.ndimage import correlate
    src = np.ascontiguousarray(src, np.float32)
    dst = np.zeros_like(src)
    for theta in gabor_thetas():
        k = gabor_kernel(GABOR["ksize"], GABOR["sigma"], theta, GABOR["lambd"], GABOR["gamma"], GABOR["psi"]).astype(np.float64)
        plane = np.empty_like(src)
        for c in range(src.shape[2]):
            plane[..., c] = correlate(src[..., c].astype(np.float64), k, mode="mirror").astype(np.float32)
        plane[plane > 1.0] = 1.0
        plane[plane < 0.0] = 0.0
        dst += plane
    return dst / np.float32(GABOR["angles"])
