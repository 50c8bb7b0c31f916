// Kernel group 3 — Laplacian-pyramid blend (reference src/blend.hpp:11-91) on planar float32 levels.
//
// Per frame and level k >= 1 the Gaussian pyramids are stored as 7 planes (left B,G,R, right B,G,R, mask); level 0
// is never materialised in float: it is the warped 8-bit pair converted on the fly (src/algo.cpp:247-248) plus the
// frame's mask plane. The collapsed result out[k] has 3 planes.
//
//   k_pyr_down<L0>   cv::pyrDown of all 7 planes          OCV imgproc/src/pyramids.cpp:745-900 (+344-402, 503-521)
//   k_blend_coarsest resultSmallest                         reference src/blend.hpp:68-69
//   k_collapse<L0>   lap = G - pyrUp(G_coarse) for both images, per-level blend, and
//                    out = pyrUp(out_coarse) + blended      reference src/blend.hpp:45-77; pyramids.cpp:903-1005
//
// Bit-exactness: OpenCV's SSE-baseline vector bodies associate the 5-tap sums differently from the scalar code
// that handles row borders and loop tails, so the association is selected per element position exactly as the
// reference loops do (see h_vec3/h_vec1/v_vec below); all adds and multiplies are individually rounded.
#include "common.cuh"
#include "kernels.cuh"

namespace poppy {

namespace {

constexpr float kInv255 = (float)(1.0 / 255.0);

// horizontal 1-4-6-4-1, vector-body association: r2*6 + ((r1+r3)*4 + (r0+r4))
__device__ __forceinline__ float h5_vec(float t0, float t1, float t2, float t3, float t4) {
    return __fadd_rn(__fmul_rn(t2, 6.f), __fadd_rn(__fmul_rn(__fadd_rn(t1, t3), 4.f), __fadd_rn(t0, t4)));
}
// scalar association: ((r2*6 + (r1+r3)*4) + r0) + r4
__device__ __forceinline__ float h5_sca(float t0, float t1, float t2, float t3, float t4) {
    return __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(t2, 6.f), __fmul_rn(__fadd_rn(t1, t3), 4.f)), t0), t4);
}
// vertical, vector body: ((r1+r3+r2)*4 + (r0+r4+(r2+r2))) * 1/256
__device__ __forceinline__ float v5_vec(float r0, float r1, float r2, float r3, float r4) {
    float a = __fmul_rn(__fadd_rn(__fadd_rn(r1, r3), r2), 4.f);
    float b = __fadd_rn(__fadd_rn(r0, r4), __fadd_rn(r2, r2));
    return __fmul_rn(__fadd_rn(a, b), 1.f / 256);
}
__device__ __forceinline__ float v5_sca(float r0, float r1, float r2, float r3, float r4) {
    return __fmul_rn(h5_sca(r0, r1, r2, r3, r4), 1.f / 256);
}

struct Taps7 { float v[7]; };

// level-0 texel: six 8-bit channels -> float * (1/255), plus the mask plane
__device__ __forceinline__ Taps7 load0(const uint2* __restrict__ warped, const float* __restrict__ mask0, size_t idx) {
    uint2 p = __ldg(warped + idx);
    Taps7 t;
    t.v[0] = __fmul_rn((float)(p.x & 255u), kInv255);
    t.v[1] = __fmul_rn((float)((p.x >> 8) & 255u), kInv255);
    t.v[2] = __fmul_rn((float)((p.x >> 16) & 255u), kInv255);
    t.v[3] = __fmul_rn((float)(p.y & 255u), kInv255);
    t.v[4] = __fmul_rn((float)((p.y >> 8) & 255u), kInv255);
    t.v[5] = __fmul_rn((float)((p.y >> 16) & 255u), kInv255);
    t.v[6] = __ldg(mask0 + idx);
    return t;
}

}  // namespace

// block (32, 8); grid (ceil(dw/32), ceil(dh/8), frames)
template <bool L0>
__global__ void __launch_bounds__(256)
k_pyr_down(const uint2* __restrict__ warped, const float* __restrict__ mask0, const float* __restrict__ src, int sw,
           int sh, int spitch, size_t sstride, float* __restrict__ dst, int dw, int dh, int dpitch, size_t dstride) {
    const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y, f = blockIdx.z;
    if (x >= dw || y >= dh) return;
    const int width0 = min((sw - 3) / 2 + 1, dw);
    // positions handled by the reference's vector bodies (pyramids.cpp:380-402 cn=3, 344-360 cn=1, 503-521)
    const bool h_vec3 = x >= 1 && x <= width0 - 2;
    const int k1 = width0 >= 5 ? (width0 - 5) / 4 + 1 : 0;
    const bool h_vec1 = x >= 1 && x < 1 + 4 * k1;
    const int v_end3 = (dw * 3) & ~3, v_end1 = dw & ~3;

    int cx[5];
#pragma unroll
    for (int j = 0; j < 5; ++j) cx[j] = reflect101(2 * x - 2 + j, sw);

    float rows[5][7];
#pragma unroll
    for (int k = 0; k < 5; ++k) {
        const int sy = reflect101(2 * y - 2 + k, sh);
        Taps7 t[5];
        if (L0) {
            const size_t rowbase = ((size_t)f * sh + sy) * sw;
#pragma unroll
            for (int j = 0; j < 5; ++j) t[j] = load0(warped, mask0, rowbase + cx[j]);
        } else {
            const float* base = src + (size_t)f * 7 * sstride + (size_t)sy * spitch;
#pragma unroll
            for (int j = 0; j < 5; ++j)
#pragma unroll
                for (int p = 0; p < 7; ++p) t[j].v[p] = __ldg(base + (size_t)p * sstride + cx[j]);
        }
#pragma unroll
        for (int p = 0; p < 7; ++p) {
            const bool vec = p < 6 ? h_vec3 : h_vec1;
            rows[k][p] = vec ? h5_vec(t[0].v[p], t[1].v[p], t[2].v[p], t[3].v[p], t[4].v[p])
                             : h5_sca(t[0].v[p], t[1].v[p], t[2].v[p], t[3].v[p], t[4].v[p]);
        }
    }
    float* out = dst + (size_t)f * 7 * dstride + (size_t)y * dpitch + x;
#pragma unroll
    for (int p = 0; p < 7; ++p) {
        const bool vec = p < 6 ? (3 * x + (p % 3)) < v_end3 : x < v_end1;
        out[(size_t)p * dstride] = vec ? v5_vec(rows[0][p], rows[1][p], rows[2][p], rows[3][p], rows[4][p])
                                       : v5_sca(rows[0][p], rows[1][p], rows[2][p], rows[3][p], rows[4][p]);
    }
}

// block 256; grid (ceil(w*h/256), frames)
__global__ void k_blend_coarsest(const float* __restrict__ g, int w, int h, int pitch, size_t stride,
                                 float* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x, f = blockIdx.y;
    if (i >= w * h) return;
    const size_t o = (size_t)(i / w) * pitch + (i % w);
    const float* gf = g + (size_t)f * 7 * stride + o;
    const float m = gf[6 * stride], anti = __fsub_rn(1.0f, m);
#pragma unroll
    for (int c = 0; c < 3; ++c)
        out[((size_t)f * 3 + c) * stride + o] =
            __fadd_rn(__fmul_rn(gf[(size_t)c * stride], m), __fmul_rn(gf[(size_t)(3 + c) * stride], anti));
}

namespace {

// horizontal pass of cv::pyrUp at fine column x of one coarse row (pyramids.cpp:945-978)
__device__ __forceinline__ float up_h(const float* __restrict__ row, int n, int x) {
    const int sx = x >> 1;
    if (n == 1) return __fmul_rn(__ldg(row), 8.f);
    if (x & 1) {
        if (sx == n - 1) return __fmul_rn(__ldg(row + n - 1), 8.f);
        return __fmul_rn(__fadd_rn(__ldg(row + sx), __ldg(row + sx + 1)), 4.f);
    }
    if (sx == 0) return __fadd_rn(__fmul_rn(__ldg(row), 6.f), __fmul_rn(__ldg(row + 1), 2.f));
    if (sx == n - 1) return __fadd_rn(__ldg(row + n - 2), __fmul_rn(__ldg(row + n - 1), 7.f));
    return __fadd_rn(__fadd_rn(__ldg(row + sx - 1), __fmul_rn(__ldg(row + sx), 6.f)), __ldg(row + sx + 1));
}

// cv::pyrUp value at fine pixel (x, y) of a coarse plane (pyramids.cpp:929-993)
__device__ __forceinline__ float up_at(const float* __restrict__ plane, int cw, int ch, int cpitch, int x, int y) {
    const int sy = y >> 1;
    const float* r1 = plane + (size_t)sy * cpitch;
    const float* r2 = plane + (size_t)(reflect101(2 * (sy + 1), 2 * ch) >> 1) * cpitch;
    if (y & 1) return __fmul_rn(__fmul_rn(__fadd_rn(up_h(r1, cw, x), up_h(r2, cw, x)), 4.f), 1.f / 64);
    const float* r0 = plane + (size_t)(reflect101(2 * (sy - 1), 2 * ch) >> 1) * cpitch;
    return __fmul_rn(__fadd_rn(__fadd_rn(up_h(r0, cw, x), __fmul_rn(up_h(r1, cw, x), 6.f)), up_h(r2, cw, x)), 1.f / 64);
}

}  // namespace

// block (32, 8); grid (ceil(w/32), ceil(h/8), frames)
template <bool L0>
__global__ void __launch_bounds__(256)
k_collapse(const uint2* __restrict__ warped, const float* __restrict__ mask0, const float* __restrict__ g_fine, int w,
           int h, int fpitch, size_t fstride, const float* __restrict__ g_coarse, const float* __restrict__ out_coarse,
           int cw, int ch, int cpitch, size_t cstride, float* __restrict__ out_fine, int opitch, size_t ostride) {
    const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y, f = blockIdx.z;
    if (x >= w || y >= h) return;
    float gl[3], gr[3], m;
    if (L0) {
        Taps7 t = load0(warped, mask0, ((size_t)f * h + y) * w + x);
#pragma unroll
        for (int c = 0; c < 3; ++c) { gl[c] = t.v[c]; gr[c] = t.v[3 + c]; }
        m = t.v[6];
    } else {
        const float* p = g_fine + (size_t)f * 7 * fstride + (size_t)y * fpitch + x;
#pragma unroll
        for (int c = 0; c < 3; ++c) { gl[c] = __ldg(p + (size_t)c * fstride); gr[c] = __ldg(p + (size_t)(3 + c) * fstride); }
        m = __ldg(p + 6 * fstride);
    }
    const float anti = __fsub_rn(1.0f, m);
    const float* gc = g_coarse + (size_t)f * 7 * cstride;
    const float* oc = out_coarse + (size_t)f * 3 * cstride;
    float* o = out_fine + (size_t)f * 3 * ostride + (size_t)y * opitch + x;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const float lap_l = __fsub_rn(gl[c], up_at(gc + (size_t)c * cstride, cw, ch, cpitch, x, y));
        const float lap_r = __fsub_rn(gr[c], up_at(gc + (size_t)(3 + c) * cstride, cw, ch, cpitch, x, y));
        const float blended = __fadd_rn(__fmul_rn(lap_l, m), __fmul_rn(lap_r, anti));
        o[(size_t)c * ostride] = __fadd_rn(up_at(oc + (size_t)c * cstride, cw, ch, cpitch, x, y), blended);
    }
}

// ------------------------------------------------------------------------------------------------------------------
void launch_pyr_down0(cudaStream_t st, const uint2* warped, const float* mask0, int w, int h, float* dst, LevelDesc dl,
                      int frames) {
    k_pyr_down<true><<<dim3(div_up(dl.w, 32), div_up(dl.h, 8), frames), dim3(32, 8), 0, st>>>(
        warped, mask0, nullptr, w, h, 0, 0, dst, dl.w, dl.h, dl.pitch, dl.plane_stride);
}

void launch_pyr_down(cudaStream_t st, const float* src, LevelDesc sl, float* dst, LevelDesc dl, int frames) {
    k_pyr_down<false><<<dim3(div_up(dl.w, 32), div_up(dl.h, 8), frames), dim3(32, 8), 0, st>>>(
        nullptr, nullptr, src, sl.w, sl.h, sl.pitch, sl.plane_stride, dst, dl.w, dl.h, dl.pitch, dl.plane_stride);
}

void launch_blend_coarsest(cudaStream_t st, const float* g, LevelDesc l, float* out, int frames) {
    k_blend_coarsest<<<dim3(div_up(l.w * l.h, 256), frames), 256, 0, st>>>(g, l.w, l.h, l.pitch, l.plane_stride, out);
}

void launch_collapse(cudaStream_t st, const float* g_fine, LevelDesc fl, const float* g_coarse, const float* out_coarse,
                     LevelDesc cl, float* out_fine, int frames) {
    k_collapse<false><<<dim3(div_up(fl.w, 32), div_up(fl.h, 8), frames), dim3(32, 8), 0, st>>>(
        nullptr, nullptr, g_fine, fl.w, fl.h, fl.pitch, fl.plane_stride, g_coarse, out_coarse, cl.w, cl.h, cl.pitch,
        cl.plane_stride, out_fine, fl.pitch, fl.plane_stride);
}

void launch_collapse0(cudaStream_t st, const uint2* warped, const float* mask0, int w, int h, const float* g_coarse,
                      const float* out_coarse, LevelDesc cl, float* out_fine, LevelDesc ol, int frames) {
    k_collapse<true><<<dim3(div_up(w, 32), div_up(h, 8), frames), dim3(32, 8), 0, st>>>(
        warped, mask0, nullptr, w, h, 0, 0, g_coarse, out_coarse, cl.w, cl.h, cl.pitch, cl.plane_stride, out_fine,
        ol.pitch, ol.plane_stride);
}

}  // namespace poppy
