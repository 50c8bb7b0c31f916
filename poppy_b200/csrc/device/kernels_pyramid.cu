// Kernel group 3 — Laplacian-pyramid blend (reference src/blend.hpp:11-91) on planar float32 levels.
//
// Per frame and level k >= 1 the Gaussian pyramids are stored as 7 planes (left B,G,R, right B,G,R, mask); level 0
// is never materialised in float: it is the warped 8-bit pair converted on the fly (src/algo.cpp:247-248) plus the
// frame's mask evaluated from the pair's mask basis (src/algo.cpp:250-258). The collapsed result out[k] has 3 planes.
//
//   k_pyr_down_tile   cv::pyrDown of one plane, 64x16 output tile per CTA: separable 1-4-6-4-1, the row pass goes
//   k_pyr_down0_tile  through shared memory once             OCV imgproc/src/pyramids.cpp:745-900 (+344-402, 503-521)
//   k_blend_coarsest  resultSmallest                         reference src/blend.hpp:68-69
//   k_collapse_tile   lap = G - pyrUp(G_coarse) for both images, per-level blend, and
//                     out = pyrUp(out_coarse) + blended, 128x32 fine tile per CTA; the polyphase row pass of the nine
//                     coarse planes is staged in shared memory  reference src/blend.hpp:45-77; pyramids.cpp:903-1005
//
// Bit-exactness: OpenCV's SSE-baseline vector bodies associate the 5-tap sums differently from the scalar code
// that handles row borders and loop tails, so the association is selected per element position exactly as the
// reference loops do (hv/vv below); all adds and multiplies are individually rounded (no FMA contraction).
// Tile interiors take vectorised fast paths; everything that touches an image border takes a scalar path with
// cv::borderInterpolate semantics, so any size (down to 1x1 levels) is handled by the same kernels.
#include "common.cuh"
#include "kernels.cuh"
#include "pixel_ops.cuh"

namespace poppy {

namespace {

// horizontal 1-4-6-4-1, vector-body association: r2*6 + ((r1+r3)*4 + (r0+r4))
__device__ __forceinline__ float h5_vec(float t0, float t1, float t2, float t3, float t4) {
    return __fadd_rn(__fmul_rn(t2, 6.f), __fadd_rn(__fmul_rn(__fadd_rn(t1, t3), 4.f), __fadd_rn(t0, t4)));
}
// scalar association: ((r2*6 + (r1+r3)*4) + r0) + r4
__device__ __forceinline__ float h5_sca(float t0, float t1, float t2, float t3, float t4) {
    return __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(t2, 6.f), __fmul_rn(__fadd_rn(t1, t3), 4.f)), t0), t4);
}
__device__ __forceinline__ float h5(bool vec, float t0, float t1, float t2, float t3, float t4) {
    return vec ? h5_vec(t0, t1, t2, t3, t4) : h5_sca(t0, t1, t2, t3, t4);
}
// vertical, vector body: ((r1+r3+r2)*4 + (r0+r4+(r2+r2))) * 1/256
__device__ __forceinline__ float v5_vec(float r0, float r1, float r2, float r3, float r4) {
    const float a = __fmul_rn(__fadd_rn(__fadd_rn(r1, r3), r2), 4.f);
    const float b = __fadd_rn(__fadd_rn(r0, r4), __fadd_rn(r2, r2));
    return __fmul_rn(__fadd_rn(a, b), 1.f / 256);
}
__device__ __forceinline__ float v5_sca(float r0, float r1, float r2, float r3, float r4) {
    return __fmul_rn(h5_sca(r0, r1, r2, r3, r4), 1.f / 256);
}
__device__ __forceinline__ float v5(bool vec, float r0, float r1, float r2, float r3, float r4) {
    return vec ? v5_vec(r0, r1, r2, r3, r4) : v5_sca(r0, r1, r2, r3, r4);
}

// Which output columns the reference's vector bodies produce (pyramids.cpp:380-402 cn=3, 344-360 cn=1, 503-521);
// all other positions come from the scalar loops.
struct DownSel {
    int width0, h1_end, v_end3, v_end1;
    __device__ DownSel(int sw, int dw) {
        width0 = min((sw - 3) / 2 + 1, dw);
        const int k1 = width0 >= 5 ? (width0 - 5) / 4 + 1 : 0;
        h1_end = 1 + 4 * k1;
        v_end3 = (dw * 3) & ~3;
        v_end1 = dw & ~3;
    }
    __device__ __forceinline__ bool h3(int x) const { return x >= 1 && x <= width0 - 2; }
    __device__ __forceinline__ bool h1(int x) const { return x >= 1 && x < h1_end; }
    __device__ __forceinline__ bool v3(int x, int ch) const { return 3 * x + ch < v_end3; }
    __device__ __forceinline__ bool v1(int x) const { return x < v_end1; }
};

constexpr int PD_OW = 64, PD_OH = 16;        // output tile of the pyrDown kernels
constexpr int PD_IR = 2 * PD_OH + 3;         // source rows a tile needs

// Column pass of a pyrDown tile for one plane: thread (t, q) turns row-pass rows 4q..4q+6 of columns 2t, 2t+1 into
// output rows oy0+2q, oy0+2q+1.
__device__ __forceinline__ void down_column_pass(const float (*hs)[PD_OW], int tid, int ox0, int oy0, int dw, int dh,
                                                 bool vec_a, bool vec_b, float* __restrict__ dplane, int dpitch) {
    const int t = tid & 31, q = tid >> 5;
    const int x = ox0 + 2 * t, y = oy0 + 2 * q;
    if (x >= dw || y >= dh) return;
    float2 hrow[7];
    const int nrows = (y + 1 < dh) ? 7 : 5;
#pragma unroll
    for (int j = 0; j < 7; ++j)
        if (j < nrows) hrow[j] = *reinterpret_cast<const float2*>(&hs[4 * q + j][2 * t]);
    float2 o;
    o.x = v5(vec_a, hrow[0].x, hrow[1].x, hrow[2].x, hrow[3].x, hrow[4].x);
    o.y = v5(vec_b, hrow[0].y, hrow[1].y, hrow[2].y, hrow[3].y, hrow[4].y);
    *reinterpret_cast<float2*>(dplane + (size_t)y * dpitch + x) = o;
    if (nrows == 7) {
        o.x = v5(vec_a, hrow[2].x, hrow[3].x, hrow[4].x, hrow[5].x, hrow[6].x);
        o.y = v5(vec_b, hrow[2].y, hrow[3].y, hrow[4].y, hrow[5].y, hrow[6].y);
        *reinterpret_cast<float2*>(dplane + (size_t)(y + 1) * dpitch + x) = o;
    }
}

}  // namespace

// ------------------------------------------------------------------------------------------------------------------
// level k -> k+1 (k >= 1) of one plane. block 256; grid (ceil(dw/64), ceil(dh/16), frames * 7): blockIdx.z is the
// plane job f*7+p, whose planes start at job * stride in both levels.
__global__ void __launch_bounds__(256)
k_pyr_down_tile(const float* __restrict__ src, int sw, int sh, int spitch, size_t sstride, float* __restrict__ dst, int dw,
                int dh, int dpitch, size_t dstride) {
    __shared__ __align__(16) float hs[PD_IR][PD_OW];
    const int job = blockIdx.z, p = job % 7;
    const float* __restrict__ sp = src + (size_t)job * sstride;
    float* __restrict__ dp = dst + (size_t)job * dstride;
    const int ox0 = blockIdx.x * PD_OW, oy0 = blockIdx.y * PD_OH;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const DownSel sel(sw, dw);
    const bool three = p < 6;

    // row pass: one warp per source row, lane -> output columns x, x+1 (taps 2x-2 .. 2x+4)
    const int last_row = 2 * (min(oy0 + PD_OH, dh) - 1) + 2 - (2 * oy0 - 2);      // last needed row-pass row
    const int x = ox0 + 2 * lane;
    const bool va = three ? sel.h3(x) : sel.h1(x), vb = three ? sel.h3(x + 1) : sel.h1(x + 1);
    for (int r = warp; r <= last_row; r += 8) {
        const float* __restrict__ row = sp + (size_t)reflect101(2 * oy0 - 2 + r, sh) * spitch;
        float2 o = make_float2(0.f, 0.f);
        if (x < dw) {
            const int c0 = 2 * x - 2;
            if (c0 >= 0 && c0 + 6 < sw) {
                const float2 a = __ldg(reinterpret_cast<const float2*>(row + c0));
                const float4 b = __ldg(reinterpret_cast<const float4*>(row + c0 + 2));
                const float c = __ldg(row + c0 + 6);
                o.x = h5(va, a.x, a.y, b.x, b.y, b.z);
                o.y = h5(vb, b.x, b.y, b.z, b.w, c);
            } else {
                float t[5];
#pragma unroll
                for (int j = 0; j < 5; ++j) t[j] = __ldg(row + reflect101(c0 + j, sw));
                o.x = h5(va, t[0], t[1], t[2], t[3], t[4]);
                if (x + 1 < dw) {
#pragma unroll
                    for (int j = 0; j < 5; ++j) t[j] = __ldg(row + reflect101(c0 + 2 + j, sw));
                    o.y = h5(vb, t[0], t[1], t[2], t[3], t[4]);
                }
            }
        }
        *reinterpret_cast<float2*>(&hs[r][2 * lane]) = o;
    }
    __syncthreads();
    const int xo = ox0 + 2 * (tid & 31);
    const bool wa = three ? sel.v3(xo, p % 3) : sel.v1(xo), wb = three ? sel.v3(xo + 1, p % 3) : sel.v1(xo + 1);
    down_column_pass(hs, tid, ox0, oy0, dw, dh, wa, wb, dp, dpitch);
}

// level 0 -> 1, all seven planes of a frame: the source is the warped 8-bit pair (one uint2 per pixel: image 1 BGR
// in .x, image 2 BGR in .y) converted on the fly, and the frame's blend mask evaluated from the mask basis.
// block 256; grid (ceil(dw/64), ceil(dh/16), frames); dynamic shared memory 7 * PD_IR * PD_OW floats.
__global__ void __launch_bounds__(256)
k_pyr_down0_tile(const uint2* __restrict__ warped, int wpitch, const float* __restrict__ basis, int bpitch,
                 const FrameParams* __restrict__ fp, int sw, int sh, float* __restrict__ dst, int dw, int dh, int dpitch,
                 size_t dstride) {
    extern __shared__ __align__(16) float smem_dyn[];
    float (*hs)[PD_IR][PD_OW] = reinterpret_cast<float (*)[PD_IR][PD_OW]>(smem_dyn);
    const int f = blockIdx.z;
    const int ox0 = blockIdx.x * PD_OW, oy0 = blockIdx.y * PD_OH;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const DownSel sel(sw, dw);
    const double alpha = fp[f].mask_alpha, beta = fp[f].mask_beta;
    const uint2* __restrict__ wframe = warped + (size_t)f * sh * wpitch;

    const int last_row = 2 * (min(oy0 + PD_OH, dh) - 1) + 2 - (2 * oy0 - 2);
    const int x = ox0 + 2 * lane;
    const bool va3 = sel.h3(x), vb3 = sel.h3(x + 1), va1 = sel.h1(x), vb1 = sel.h1(x + 1);
    for (int r = warp; r <= last_row; r += 8) {
        const int sy = reflect101(2 * oy0 - 2 + r, sh);
        const uint2* __restrict__ wrow = wframe + (size_t)sy * wpitch;
        const float* __restrict__ brow = basis + (size_t)sy * bpitch;
        if (x >= dw) continue;          // columns past the level: never read by the column pass
        const int c0 = 2 * x - 2;
        if (c0 >= 0 && c0 + 6 < sw) {
            uint32_t lo[7], hi[7];
            float mk[7];
            {
                const uint4 a = __ldg(reinterpret_cast<const uint4*>(wrow + c0));
                const uint4 b = __ldg(reinterpret_cast<const uint4*>(wrow + c0 + 2));
                const uint4 c = __ldg(reinterpret_cast<const uint4*>(wrow + c0 + 4));
                const uint2 d = __ldg(wrow + c0 + 6);
                lo[0] = a.x; hi[0] = a.y; lo[1] = a.z; hi[1] = a.w;
                lo[2] = b.x; hi[2] = b.y; lo[3] = b.z; hi[3] = b.w;
                lo[4] = c.x; hi[4] = c.y; lo[5] = c.z; hi[5] = c.w;
                lo[6] = d.x; hi[6] = d.y;
                const float2 ma = __ldg(reinterpret_cast<const float2*>(brow + c0));
                const float4 mb = __ldg(reinterpret_cast<const float4*>(brow + c0 + 2));
                const float mc = __ldg(brow + c0 + 6);
                mk[0] = blend_mask(ma.x, alpha, beta); mk[1] = blend_mask(ma.y, alpha, beta);
                mk[2] = blend_mask(mb.x, alpha, beta); mk[3] = blend_mask(mb.y, alpha, beta);
                mk[4] = blend_mask(mb.z, alpha, beta); mk[5] = blend_mask(mb.w, alpha, beta);
                mk[6] = blend_mask(mc, alpha, beta);
            }
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                float t[7];
#pragma unroll
                for (int j = 0; j < 7; ++j) t[j] = unit_from_byte(lo[j], c);
                *reinterpret_cast<float2*>(&hs[c][r][2 * lane]) =
                    make_float2(h5(va3, t[0], t[1], t[2], t[3], t[4]), h5(vb3, t[2], t[3], t[4], t[5], t[6]));
#pragma unroll
                for (int j = 0; j < 7; ++j) t[j] = unit_from_byte(hi[j], c);
                *reinterpret_cast<float2*>(&hs[3 + c][r][2 * lane]) =
                    make_float2(h5(va3, t[0], t[1], t[2], t[3], t[4]), h5(vb3, t[2], t[3], t[4], t[5], t[6]));
            }
            *reinterpret_cast<float2*>(&hs[6][r][2 * lane]) =
                make_float2(h5(va1, mk[0], mk[1], mk[2], mk[3], mk[4]), h5(vb1, mk[2], mk[3], mk[4], mk[5], mk[6]));
        } else {
#pragma unroll 1
            for (int i = 0; i < 2; ++i) {
                if (x + i >= dw) break;
                const bool v3 = i ? vb3 : va3, v1 = i ? vb1 : va1;
                uint32_t lo[5], hi[5];
                float mk[5];
#pragma unroll
                for (int j = 0; j < 5; ++j) {
                    const int cx = reflect101(c0 + 2 * i + j, sw);
                    const uint2 px = __ldg(wrow + cx);
                    lo[j] = px.x; hi[j] = px.y;
                    mk[j] = blend_mask(__ldg(brow + cx), alpha, beta);
                }
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    hs[c][r][2 * lane + i] = h5(v3, unit_from_byte(lo[0], c), unit_from_byte(lo[1], c), unit_from_byte(lo[2], c),
                                                unit_from_byte(lo[3], c), unit_from_byte(lo[4], c));
                    hs[3 + c][r][2 * lane + i] = h5(v3, unit_from_byte(hi[0], c), unit_from_byte(hi[1], c), unit_from_byte(hi[2], c),
                                                    unit_from_byte(hi[3], c), unit_from_byte(hi[4], c));
                }
                hs[6][r][2 * lane + i] = h5(v1, mk[0], mk[1], mk[2], mk[3], mk[4]);
            }
        }
    }
    __syncthreads();
    const int xo = ox0 + 2 * (tid & 31);
#pragma unroll 1
    for (int p = 0; p < 7; ++p) {
        const bool wa = p < 6 ? sel.v3(xo, p % 3) : sel.v1(xo), wb = p < 6 ? sel.v3(xo + 1, p % 3) : sel.v1(xo + 1);
        down_column_pass(hs[p], tid, ox0, oy0, dw, dh, wa, wb, dst + ((size_t)f * 7 + p) * dstride, dpitch);
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

// ------------------------------------------------------------------------------------------------------------------
namespace {

constexpr int CT_FW = 128, CT_FH = 32;                 // fine tile of the collapse kernel
constexpr int CT_HR = CT_FH / 2 + 2;                   // coarse rows cy0-1 .. cy0+CT_FH/2 take part
constexpr size_t CT_SMEM = (size_t)9 * CT_HR * CT_FW * sizeof(float);

// horizontal pass of cv::pyrUp at fine column x of one coarse row of n pixels (pyramids.cpp:945-978)
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

// vertical pass of cv::pyrUp (pyramids.cpp:929-993): even fine row (r0 + 6 r1 + r2)/64, odd fine row 4 (r1 + r2)/64
__device__ __forceinline__ float up_even(float h0, float h1, float h2) {
    return __fmul_rn(__fadd_rn(__fadd_rn(h0, __fmul_rn(h1, 6.f)), h2), 1.f / 64);
}
__device__ __forceinline__ float up_odd(float h1, float h2) {
    return __fmul_rn(__fmul_rn(__fadd_rn(h1, h2), 4.f), 1.f / 64);
}

struct Up4x4 { float4 row[4]; };     // pyrUp values of a 4-column x 4-row fine block

// Fine rows 4q..4q+3 of the tile come from coarse rows sy0 = cy0+2q and sy0+1; ja..jd are the shared-memory rows of
// (sy0-1 | reflected), sy0, sy0+1 (clamped), sy0+2 (clamped) for this plane.
__device__ __forceinline__ Up4x4 up_block(const float* __restrict__ hp, int col, int ja, int jb, int jc, int jd) {
    const float4 a = *reinterpret_cast<const float4*>(hp + ja * CT_FW + col);
    const float4 b = *reinterpret_cast<const float4*>(hp + jb * CT_FW + col);
    const float4 c = *reinterpret_cast<const float4*>(hp + jc * CT_FW + col);
    const float4 d = *reinterpret_cast<const float4*>(hp + jd * CT_FW + col);
    Up4x4 u;
    u.row[0] = make_float4(up_even(a.x, b.x, c.x), up_even(a.y, b.y, c.y), up_even(a.z, b.z, c.z), up_even(a.w, b.w, c.w));
    u.row[1] = make_float4(up_odd(b.x, c.x), up_odd(b.y, c.y), up_odd(b.z, c.z), up_odd(b.w, c.w));
    u.row[2] = make_float4(up_even(b.x, c.x, d.x), up_even(b.y, c.y, d.y), up_even(b.z, c.z, d.z), up_even(b.w, c.w, d.w));
    u.row[3] = make_float4(up_odd(c.x, d.x), up_odd(c.y, d.y), up_odd(c.z, d.z), up_odd(c.w, d.w));
    return u;
}

}  // namespace

// block 256; grid (ceil(w/128), ceil(h/32), frames); dynamic shared memory CT_SMEM.
// L0: the fine Gaussian level is the warped 8-bit pair + mask basis (see k_pyr_down0_tile); else g_fine (7 planes).
template <bool L0>
__global__ void __launch_bounds__(256, 2)
k_collapse_tile(const uint2* __restrict__ warped, int wpitch, const float* __restrict__ basis, int bpitch,
                const FrameParams* __restrict__ fp, const float* __restrict__ g_fine, int w, int h, int fpitch,
                size_t fstride, const float* __restrict__ g_coarse, const float* __restrict__ out_coarse, int cw, int ch,
                int cpitch, size_t cstride, float* __restrict__ out_fine, int opitch, size_t ostride) {
    extern __shared__ __align__(16) float smem_dyn[];
    float* hs = smem_dyn;                                   // [9][CT_HR][CT_FW]
    const int f = blockIdx.z;
    const int fx0 = blockIdx.x * CT_FW, fy0 = blockIdx.y * CT_FH;
    const int cx0 = fx0 >> 1, cy0 = fy0 >> 1;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const float* __restrict__ gc = g_coarse + (size_t)f * 7 * cstride;
    const float* __restrict__ oc = out_coarse + (size_t)f * 3 * cstride;

    // 1) polyphase row pass of the nine coarse planes: one warp per (plane, coarse row); lane -> 4 fine columns
    {
        const int fx = fx0 + 4 * lane, a = cx0 + 2 * lane;
        const bool interior = a >= 1 && a + 2 <= cw - 1;
        for (int task = warp; task < 9 * CT_HR; task += 8) {
            const int p = task / CT_HR, j = task - p * CT_HR;
            const int cy = cy0 - 1 + j;
            if (cy < 0 || cy >= ch || fx >= w) continue;
            const float* __restrict__ row = (p < 6 ? gc + (size_t)p * cstride : oc + (size_t)(p - 6) * cstride) + (size_t)cy * cpitch;
            float4 o;
            if (interior) {
                const float cm = __ldg(row + a - 1);
                const float2 c01 = __ldg(reinterpret_cast<const float2*>(row + a));
                const float cp = __ldg(row + a + 2);
                o.x = __fadd_rn(__fadd_rn(cm, __fmul_rn(c01.x, 6.f)), c01.y);
                o.y = __fmul_rn(__fadd_rn(c01.x, c01.y), 4.f);
                o.z = __fadd_rn(__fadd_rn(c01.x, __fmul_rn(c01.y, 6.f)), cp);
                o.w = __fmul_rn(__fadd_rn(c01.y, cp), 4.f);
            } else {
                o.x = up_h(row, cw, fx);
                o.y = fx + 1 < w ? up_h(row, cw, fx + 1) : 0.f;
                o.z = fx + 2 < w ? up_h(row, cw, fx + 2) : 0.f;
                o.w = fx + 3 < w ? up_h(row, cw, fx + 3) : 0.f;
            }
            *reinterpret_cast<float4*>(hs + ((size_t)p * CT_HR + j) * CT_FW + 4 * lane) = o;
        }
    }
    __syncthreads();

    // 2) column pass + blend: thread -> fine columns fx..fx+3, fine rows fy..fy+3 (coarse rows sy0, sy0+1)
    const int q = warp;
    const int fx = fx0 + 4 * lane, fy = fy0 + 4 * q;
    if (fx >= w || fy >= h) return;
    const int sy0 = cy0 + 2 * q;
    // shared-memory row of coarse row r is r - cy0 + 1
    const int jb = 2 * q + 1;                                                   // sy0
    const int ja = sy0 >= 1 ? jb - 1 : (ch > 1 ? 1 - cy0 + 1 : jb);              // borderInterpolate(2(sy0-1), 2ch)/2
    const int jc = min(sy0 + 1, ch - 1) - cy0 + 1;
    const int jd = min(sy0 + 2, ch - 1) - cy0 + 1;
    const int rows = min(4, h - fy);

    float4 mk[4];                      // the mask is shared by the three channels
    uint4 w01[4], w23[4];
    if (L0) {
        const double alpha = fp[f].mask_alpha, beta = fp[f].mask_beta;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            if (i < rows) {
                const uint2* __restrict__ wrow = warped + ((size_t)f * h + fy + i) * wpitch + fx;
                w01[i] = __ldg(reinterpret_cast<const uint4*>(wrow));
                w23[i] = __ldg(reinterpret_cast<const uint4*>(wrow + 2));
                const float4 b = __ldg(reinterpret_cast<const float4*>(basis + (size_t)(fy + i) * bpitch + fx));
                mk[i] = make_float4(blend_mask(b.x, alpha, beta), blend_mask(b.y, alpha, beta),
                                    blend_mask(b.z, alpha, beta), blend_mask(b.w, alpha, beta));
            }
        }
    } else {
        const float* __restrict__ mp = g_fine + ((size_t)f * 7 + 6) * fstride + (size_t)fy * fpitch + fx;
#pragma unroll
        for (int i = 0; i < 4; ++i)
            if (i < rows) mk[i] = __ldg(reinterpret_cast<const float4*>(mp + (size_t)i * fpitch));
    }

#pragma unroll 1
    for (int c = 0; c < 3; ++c) {
        // acc = (gl - up_l) * m, then += (gr - up_r) * (1 - m), then out = up_o + acc      (blend.hpp:52-53,70-72,62-63)
        float4 acc[4];
        const float* __restrict__ lp = L0 ? nullptr : g_fine + ((size_t)f * 7 + c) * fstride + (size_t)fy * fpitch + fx;
        {
            const Up4x4 u = up_block(hs + (size_t)c * CT_HR * CT_FW, 4 * lane, ja, jb, jc, jd);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                if (i < rows) {
                    float4 g;
                    if (L0) g = make_float4(unit_from_byte(w01[i].x, c), unit_from_byte(w01[i].z, c), unit_from_byte(w23[i].x, c),
                                            unit_from_byte(w23[i].z, c));
                    else g = __ldg(reinterpret_cast<const float4*>(lp + (size_t)i * fpitch));
                    acc[i] = make_float4(__fmul_rn(__fsub_rn(g.x, u.row[i].x), mk[i].x), __fmul_rn(__fsub_rn(g.y, u.row[i].y), mk[i].y),
                                         __fmul_rn(__fsub_rn(g.z, u.row[i].z), mk[i].z), __fmul_rn(__fsub_rn(g.w, u.row[i].w), mk[i].w));
                }
            }
        }
        {
            const Up4x4 u = up_block(hs + (size_t)(3 + c) * CT_HR * CT_FW, 4 * lane, ja, jb, jc, jd);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                if (i < rows) {
                    float4 g;
                    if (L0) g = make_float4(unit_from_byte(w01[i].y, c), unit_from_byte(w01[i].w, c), unit_from_byte(w23[i].y, c),
                                            unit_from_byte(w23[i].w, c));
                    else g = __ldg(reinterpret_cast<const float4*>(lp + 3 * fstride + (size_t)i * fpitch));
                    acc[i].x = __fadd_rn(acc[i].x, __fmul_rn(__fsub_rn(g.x, u.row[i].x), __fsub_rn(1.f, mk[i].x)));
                    acc[i].y = __fadd_rn(acc[i].y, __fmul_rn(__fsub_rn(g.y, u.row[i].y), __fsub_rn(1.f, mk[i].y)));
                    acc[i].z = __fadd_rn(acc[i].z, __fmul_rn(__fsub_rn(g.z, u.row[i].z), __fsub_rn(1.f, mk[i].z)));
                    acc[i].w = __fadd_rn(acc[i].w, __fmul_rn(__fsub_rn(g.w, u.row[i].w), __fsub_rn(1.f, mk[i].w)));
                }
            }
        }
        {
            const Up4x4 u = up_block(hs + (size_t)(6 + c) * CT_HR * CT_FW, 4 * lane, ja, jb, jc, jd);
            float* __restrict__ op = out_fine + ((size_t)f * 3 + c) * ostride + (size_t)fy * opitch + fx;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                if (i < rows)
                    *reinterpret_cast<float4*>(op + (size_t)i * opitch) =
                        make_float4(__fadd_rn(u.row[i].x, acc[i].x), __fadd_rn(u.row[i].y, acc[i].y),
                                    __fadd_rn(u.row[i].z, acc[i].z), __fadd_rn(u.row[i].w, acc[i].w));
            }
        }
    }
}

// frame's blend mask as a plane (stage dumps only; the render path evaluates it inside the pyramid kernels)
__global__ void k_mask_plane(const float* __restrict__ basis, int bpitch, const FrameParams* __restrict__ fp, int f,
                             float* __restrict__ out, int w, int h) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= w) return;
    out[(size_t)y * w + x] = blend_mask(basis[(size_t)y * bpitch + x], fp[f].mask_alpha, fp[f].mask_beta);
}

// ------------------------------------------------------------------------------------------------------------------
void launch_pyr_down0(cudaStream_t st, const uint2* warped, int wpitch, const float* basis, int bpitch,
                      const FrameParams* fp, int w, int h, float* dst, LevelDesc dl, int frames) {
    static const size_t smem = (size_t)7 * PD_IR * PD_OW * sizeof(float);
    static bool once = false;
    if (!once) {
        cudaFuncSetAttribute(k_pyr_down0_tile, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        once = true;
    }
    k_pyr_down0_tile<<<dim3(div_up(dl.w, PD_OW), div_up(dl.h, PD_OH), frames), 256, smem, st>>>(
        warped, wpitch, basis, bpitch, fp, w, h, dst, dl.w, dl.h, dl.pitch, dl.plane_stride);
}

void launch_pyr_down(cudaStream_t st, const float* src, LevelDesc sl, float* dst, LevelDesc dl, int frames) {
    k_pyr_down_tile<<<dim3(div_up(dl.w, PD_OW), div_up(dl.h, PD_OH), frames * 7), 256, 0, st>>>(
        src, sl.w, sl.h, sl.pitch, sl.plane_stride, dst, dl.w, dl.h, dl.pitch, dl.plane_stride);
}

void launch_blend_coarsest(cudaStream_t st, const float* g, LevelDesc l, float* out, int frames) {
    k_blend_coarsest<<<dim3(div_up(l.w * l.h, 256), frames), 256, 0, st>>>(g, l.w, l.h, l.pitch, l.plane_stride, out);
}

void launch_collapse(cudaStream_t st, const float* g_fine, LevelDesc fl, const float* g_coarse, const float* out_coarse,
                     LevelDesc cl, float* out_fine, int frames) {
    static bool once = false;
    if (!once) {
        cudaFuncSetAttribute(k_collapse_tile<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CT_SMEM);
        once = true;
    }
    k_collapse_tile<false><<<dim3(div_up(fl.w, CT_FW), div_up(fl.h, CT_FH), frames), 256, CT_SMEM, st>>>(
        nullptr, 0, nullptr, 0, nullptr, g_fine, fl.w, fl.h, fl.pitch, fl.plane_stride, g_coarse, out_coarse, cl.w, cl.h,
        cl.pitch, cl.plane_stride, out_fine, fl.pitch, fl.plane_stride);
}

void launch_collapse0(cudaStream_t st, const uint2* warped, int wpitch, const float* basis, int bpitch,
                      const FrameParams* fp, int w, int h, const float* g_coarse, const float* out_coarse, LevelDesc cl,
                      float* out_fine, LevelDesc ol, int frames) {
    static bool once = false;
    if (!once) {
        cudaFuncSetAttribute(k_collapse_tile<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CT_SMEM);
        once = true;
    }
    k_collapse_tile<true><<<dim3(div_up(w, CT_FW), div_up(h, CT_FH), frames), 256, CT_SMEM, st>>>(
        warped, wpitch, basis, bpitch, fp, nullptr, w, h, 0, 0, g_coarse, out_coarse, cl.w, cl.h, cl.pitch,
        cl.plane_stride, out_fine, ol.pitch, ol.plane_stride);
}

void launch_mask_plane(cudaStream_t st, const float* basis, int bpitch, const FrameParams* fp, int frame, float* out, int w,
                       int h) {
    k_mask_plane<<<dim3(div_up(w, 256), h), 256, 0, st>>>(basis, bpitch, fp, frame, out, w, h);
}

}  // namespace poppy
