// Kernel group 4 — unsharp_mask(lapBlend, 1, amount, 0.3) fused with the final convertTo(CV_8U, 255).
// reference src/util.cpp:113-148, src/algo.cpp:263-265
//
//   GaussianBlur(sigma 1) -> 9 taps, separable, BORDER_REFLECT_101. Row filter and symmetric column filter both
//   run as FMA chains in OpenCV's FMA-dispatched translation unit:
//     row:    s = k[-4]*x[-4]; s = fma(x[j], k[j], s), j = -3..4       (OCV imgproc/src/filter.simd.hpp:1663-1681,2477-2487)
//     column: s = fma(k0, t0, 0); s = fma(k[j], t[+j] + t[-j], s), j = 1..4   (filter.simd.hpp:1909-1960)
//   except for the last (3*w) % 8 interleaved elements of every row, which fall to the generic scalar loops
//   (filter.simd.hpp:2477-2487, 2753-2759); the reference build does not contract those: products and sums are
//   rounded separately there.
//   diff = x - blur; 3x3 per-channel median with replicated border (median_blur.simd.hpp:677-745);
//   where ||diff||_2 >= 0.3 (cv::norm of a Vec3f: double accumulation + sqrt): x + amount*diff, else x
//   (src/util.cpp:135-145, unfused); then cvRound(v*255) saturated to 8 bits (convert_scale.simd.hpp, saturate.hpp:105).
//
// One CTA produces a 32x16 output tile: the row pass is written to shared memory for the tile plus a
// 1-pixel (median) + 4-row (Gaussian) halo, the column pass and the difference stay in shared memory, and the
// packed BGR bytes are staged so that global stores are 32-bit and coalesced.
#include "common.cuh"
#include "kernels.cuh"

namespace poppy {

namespace {

constexpr int TW = 32, TH = 16;
constexpr int CW = TW + 2;            // cells per row: tile + median halo
constexpr int TR = TH + 10;           // row-pass rows: tile + median halo + Gaussian halo
constexpr int DR = TH + 2;

// getGaussianKernel(9, 1, CV_32F) (bit-exact kernel, OCV imgproc/src/smooth.dispatch.cpp:81-198): centre .. edge
__device__ __forceinline__ float gk(int j) {
    j = j < 0 ? -j : j;
    const unsigned bits = j == 0 ? 0x3ecc4252u : j == 1 ? 0x3e77c75du : j == 2 ? 0x3d5d25cdu : j == 3 ? 0x3b913926u : 0x390c54e2u;
    return __uint_as_float(bits);
}

#define POPPY_SORT2(a, b) { float lo_ = fminf(a, b); b = fmaxf(a, b); a = lo_; }
__device__ __forceinline__ float median9(float p0, float p1, float p2, float p3, float p4, float p5, float p6, float p7,
                                         float p8) {
    POPPY_SORT2(p1, p2) POPPY_SORT2(p4, p5) POPPY_SORT2(p7, p8) POPPY_SORT2(p0, p1) POPPY_SORT2(p3, p4)
    POPPY_SORT2(p6, p7) POPPY_SORT2(p1, p2) POPPY_SORT2(p4, p5) POPPY_SORT2(p7, p8) POPPY_SORT2(p0, p3)
    POPPY_SORT2(p5, p8) POPPY_SORT2(p4, p7) POPPY_SORT2(p3, p6) POPPY_SORT2(p1, p4) POPPY_SORT2(p2, p5)
    POPPY_SORT2(p4, p7) POPPY_SORT2(p4, p2) POPPY_SORT2(p6, p4) POPPY_SORT2(p4, p2)
    return p4;
}
#undef POPPY_SORT2

}  // namespace

// block 256; grid (ceil(w/32), ceil(h/16), frames). norm_thr2: smallest double whose sqrt is >= (double)0.3f.
__global__ void __launch_bounds__(256)
k_unsharp_store(const float* __restrict__ lap, int w, int h, int pitch, size_t stride,
                const FrameParams* __restrict__ fp, double norm_thr2, uint8_t* __restrict__ frames_base,
                size_t frame_bytes) {
    __shared__ float s_row[TR][3][CW];
    __shared__ float s_diff[DR][3][CW];
    __shared__ __align__(4) uint8_t s_out[TH][TW * 3];

    const int tx0 = blockIdx.x * TW, ty0 = blockIdx.y * TH, f = blockIdx.z;
    const int tid = threadIdx.x;
    const float* img = lap + (size_t)f * 3 * stride;
    const int tail_from = 3 * w - (3 * w) % 8;      // first interleaved element handled by the scalar filter loops

    // 1) row pass for every cell of the halo'd tile
    for (int i = tid; i < TR * CW; i += 256) {
        const int rj = i / CW, ci = i - rj * CW;
        const int ry = ty0 - 5 + rj;
        if (ry < 0 || ry >= h) continue;
        const int gx = min(max(tx0 - 1 + ci, 0), w - 1);
        int cx[9];
#pragma unroll
        for (int j = 0; j < 9; ++j) cx[j] = reflect101(gx - 4 + j, w);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float* row = img + (size_t)c * stride + (size_t)ry * pitch;
            float s = __fmul_rn(gk(-4), __ldg(row + cx[0]));
            if (w == 1) {
                s = __ldg(row);          // GaussianBlur shrinks the kernel to [1] along a 1-pixel axis
            } else if (3 * gx + c < tail_from) {
#pragma unroll
                for (int j = 1; j < 9; ++j) s = fmaf(__ldg(row + cx[j]), gk(j - 4), s);
            } else {
#pragma unroll
                for (int j = 1; j < 9; ++j) s = __fadd_rn(s, __fmul_rn(gk(j - 4), __ldg(row + cx[j])));
            }
            s_row[rj][c][ci] = s;
        }
    }
    __syncthreads();

    // 2) column pass + difference at the (clamped) cell position
    for (int i = tid; i < DR * CW; i += 256) {
        const int dj = i / CW, ci = i - dj * CW;
        const int gy = min(max(ty0 - 1 + dj, 0), h - 1);
        const int gx = min(max(tx0 - 1 + ci, 0), w - 1);
        const int base = ty0 - 5;
        int rp[5], rm[5];
#pragma unroll
        for (int j = 1; j <= 4; ++j) { rp[j] = reflect101(gy + j, h) - base; rm[j] = reflect101(gy - j, h) - base; }
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            float s;
            if (h == 1) {
                s = s_row[gy - base][c][ci];
            } else if (3 * gx + c < tail_from) {
                s = fmaf(gk(0), s_row[gy - base][c][ci], 0.f);
#pragma unroll
                for (int j = 1; j <= 4; ++j) s = fmaf(gk(j), __fadd_rn(s_row[rp[j]][c][ci], s_row[rm[j]][c][ci]), s);
            } else {
                s = __fadd_rn(__fmul_rn(gk(0), s_row[gy - base][c][ci]), 0.f);
#pragma unroll
                for (int j = 1; j <= 4; ++j)
                    s = __fadd_rn(s, __fmul_rn(gk(j), __fadd_rn(s_row[rp[j]][c][ci], s_row[rm[j]][c][ci])));
            }
            const float x = __ldg(img + (size_t)c * stride + (size_t)gy * pitch + gx);
            s_diff[dj][c][ci] = __fsub_rn(x, s);
        }
    }
    __syncthreads();

    // 3) median, threshold, sharpen, 8-bit pack
    const FrameParams P = fp[f];
    for (int i = tid; i < TH * TW; i += 256) {
        const int oy = i / TW, ox = i - oy * TW;
        const int gx = tx0 + ox, gy = ty0 + oy;
        if (gx >= w || gy >= h) continue;
        float med[3];
        double nn = 0.0;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float* d0 = &s_diff[oy][c][ox];
            const float* d1 = &s_diff[oy + 1][c][ox];
            const float* d2 = &s_diff[oy + 2][c][ox];
            med[c] = median9(d0[0], d0[1], d0[2], d1[0], d1[1], d1[2], d2[0], d2[1], d2[2]);
            nn = __dadd_rn(nn, __dmul_rn((double)med[c], (double)med[c]));
        }
        const bool sharpen = nn >= norm_thr2;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            float v = __ldg(img + (size_t)c * stride + (size_t)gy * pitch + gx);
            if (sharpen) v = __fadd_rn(v, __fmul_rn(P.amount, med[c]));
            const int q = cv_round(__fmul_rn(v, 255.f));
            s_out[oy][ox * 3 + c] = (uint8_t)min(max(q, 0), 255);
        }
    }
    __syncthreads();

    // 4) coalesced store of the packed BGR rows
    uint8_t* dst = frames_base + (size_t)P.dst_slot * frame_bytes;
    const int cols = min(TW, w - tx0), rows = min(TH, h - ty0), row_bytes = cols * 3;
    const size_t dst_pitch = (size_t)w * 3;
    if ((dst_pitch & 3) == 0 && ((size_t)dst & 3) == 0) {
        const int words = row_bytes >> 2;
        for (int i = tid; i < rows * (TW * 3 / 4); i += 256) {
            const int r = i / (TW * 3 / 4), wd = i - r * (TW * 3 / 4);
            if (wd < words)
                *reinterpret_cast<uint32_t*>(dst + (size_t)(ty0 + r) * dst_pitch + (size_t)tx0 * 3 + wd * 4) =
                    *reinterpret_cast<const uint32_t*>(&s_out[r][wd * 4]);
        }
        const int tail = row_bytes & 3;
        if (tail)
            for (int i = tid; i < rows * tail; i += 256) {
                const int r = i / tail, b = (row_bytes & ~3) + (i - r * tail);
                dst[(size_t)(ty0 + r) * dst_pitch + (size_t)tx0 * 3 + b] = s_out[r][b];
            }
    } else {
        for (int i = tid; i < rows * row_bytes; i += 256) {
            const int r = i / row_bytes, b = i - r * row_bytes;
            dst[(size_t)(ty0 + r) * dst_pitch + (size_t)tx0 * 3 + b] = s_out[r][b];
        }
    }
}

// Order-dependent 64-bit checksum: per-block FNV-1a over 8-byte words folded with a position weight, atomically
// xor-added. Used for "checksum of checksums" parity at full size without copying frames to the host.
__global__ void k_checksum(const uint8_t* __restrict__ data, size_t bytes, unsigned long long* __restrict__ out) {
    unsigned long long acc = 0;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < bytes; i += stride) {
        unsigned long long v = (unsigned long long)data[i] + 1ull;
        unsigned long long k = (i + 0x9E3779B97F4A7C15ull) * 0xBF58476D1CE4E5B9ull;
        k ^= k >> 29;
        acc += v * (k | 1ull);
    }
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(out, acc);
}

void launch_unsharp_store(cudaStream_t st, const float* lap_blend, LevelDesc l, const FrameParams* fp,
                          uint8_t* frames_base, size_t frame_bytes, int frames) {
    // smallest double s with sqrt(s) >= (double)0.3f: lets the kernel compare the squared norm (exactly
    // equivalent to "cv::norm(diff) >= threshold" with a correctly rounded sqrt)
    static const double thr2 = [] {
        const double t = (double)0.3f;
        double s = t * t;
        while (__builtin_sqrt(s) >= t) s = __builtin_nextafter(s, 0.0);
        while (__builtin_sqrt(s) < t) s = __builtin_nextafter(s, 1.0);
        return s;
    }();
    k_unsharp_store<<<dim3(div_up(l.w, TW), div_up(l.h, TH), frames), 256, 0, st>>>(
        lap_blend, l.w, l.h, l.pitch, l.plane_stride, fp, thr2, frames_base, frame_bytes);
}

void launch_checksum(cudaStream_t st, const uint8_t* data, size_t bytes, unsigned long long* out) {
    int blocks = (int)((bytes + 255) / 256);
    if (blocks > 148 * 16) blocks = 148 * 16;
    if (blocks < 1) blocks = 1;
    k_checksum<<<blocks, 256, 0, st>>>(data, bytes, out);
}

}  // namespace poppy
