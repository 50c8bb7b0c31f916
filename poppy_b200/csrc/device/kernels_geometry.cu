// Kernel group 1 — vertex lerp, per-triangle affine solve, and the exact triangle-ID rasteriser.
//
//   k_clip_points      clip_points()                        reference src/util.cpp:453-460
//   k_lerp_points      morph_points() + clip_points()       reference src/algo.cpp:50-58,202-205
//   k_tri_geometry     make_triangler_points, solve_homography, morph_homography, the Mat::inv of create_map
//                                                           reference src/algo.cpp:83-93,108-144,154-157
//                      + closed form of cv::FillConvexPoly's edge walkers (OCV drawing.cpp:1093-1255)
//   k_raster_triangles paint_triangles(): Bresenham outline + 16.16 DDA span fill, "later triangle wins"
//                      resolved with atomicMax             reference src/algo.cpp:95-106
#include "common.cuh"
#include "kernels.cuh"

namespace poppy {

// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float2 clip_point(float2 p, int cols, int rows) {
    p.x = p.x > (float)cols ? (float)(cols - 1) : p.x;
    p.y = p.y > (float)rows ? (float)(rows - 1) : p.y;
    p.x = p.x < 0.f ? 0.f : p.x;
    p.y = p.y < 0.f ? 0.f : p.y;
    return p;
}

__global__ void k_clip_points(const float2* __restrict__ in, float2* __restrict__ out, int n, int cols, int rows) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = clip_point(in[i], cols, rows);
}

// x = (float)((1.0 - (double)s) * (double)a + (double)(float)(s * b))      
__device__ __forceinline__ float lerp_coord(float a, float b, float s) {
    float sb = __fmul_rn(s, b);
    double t = __dadd_rn(__dmul_rn(__dsub_rn(1.0, (double)s), (double)a), (double)sb);
    return __double2float_rn(t);
}

// grid (ceil(n/256), frames). Frame f reads p1 + f * p1_frame_stride (0: every frame starts from the same set)
// and writes out + f * out_frame_stride.
__global__ void k_lerp_points(const float2* __restrict__ p1, size_t p1_frame_stride, const float2* __restrict__ p2,
                              const FrameParams* __restrict__ fp, float2* __restrict__ out, size_t out_frame_stride, int n,
                              int cols, int rows) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int f = blockIdx.y;
    if (i >= n) return;
    float s = fp[f].shape;
    float2 a = p1[(size_t)f * p1_frame_stride + i], b = p2[i], r;
    r.x = lerp_coord(a.x, b.x, s);
    r.y = lerp_coord(a.y, b.y, s);
    out[(size_t)f * out_frame_stride + i] = clip_point(r, cols, rows);
}

// ------------------------------------------------------------------------------------------------------------------
// 3x3 float inverse via double cofactors, zero matrix when singular — OCV core/src/lapack.cpp:760-763,965-995,1045
__device__ void inv3(const float* m, float* o) {
#define M(r, c) ((double)m[(r) * 3 + (c)])
#define COF(a, b, c, d) __dsub_rn(__dmul_rn(a, b), __dmul_rn(c, d))
    double d = __dadd_rn(__dsub_rn(__dmul_rn(M(0, 0), COF(M(1, 1), M(2, 2), M(1, 2), M(2, 1))),
                                   __dmul_rn(M(0, 1), COF(M(1, 0), M(2, 2), M(1, 2), M(2, 0)))),
                         __dmul_rn(M(0, 2), COF(M(1, 0), M(2, 1), M(1, 1), M(2, 0))));
    if (d == 0.) {
#pragma unroll
        for (int i = 0; i < 9; ++i) o[i] = 0.f;
        return;
    }
    d = __ddiv_rn(1., d);
    o[0] = __double2float_rn(__dmul_rn(COF(M(1, 1), M(2, 2), M(1, 2), M(2, 1)), d));
    o[1] = __double2float_rn(__dmul_rn(COF(M(0, 2), M(2, 1), M(0, 1), M(2, 2)), d));
    o[2] = __double2float_rn(__dmul_rn(COF(M(0, 1), M(1, 2), M(0, 2), M(1, 1)), d));
    o[3] = __double2float_rn(__dmul_rn(COF(M(1, 2), M(2, 0), M(1, 0), M(2, 2)), d));
    o[4] = __double2float_rn(__dmul_rn(COF(M(0, 0), M(2, 2), M(0, 2), M(2, 0)), d));
    o[5] = __double2float_rn(__dmul_rn(COF(M(0, 2), M(1, 0), M(0, 0), M(1, 2)), d));
    o[6] = __double2float_rn(__dmul_rn(COF(M(1, 0), M(2, 1), M(1, 1), M(2, 0)), d));
    o[7] = __double2float_rn(__dmul_rn(COF(M(0, 1), M(2, 0), M(0, 0), M(2, 1)), d));
    o[8] = __double2float_rn(__dmul_rn(COF(M(0, 0), M(1, 1), M(0, 1), M(1, 0)), d));
#undef COF
#undef M
}

// 3x3 float product as the FMA-dispatched gemm evaluates it: fma(a2,b2, fma(a0,b0, a1*b1))
// (OCV core/src/matmul.simd.hpp:827-858)
__device__ void mul3(const float* a, const float* b, float* o) {
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            float p1 = __fmul_rn(a[i * 3 + 1], b[3 + j]);
            o[i * 3 + j] = fmaf(a[i * 3 + 2], b[6 + j], fmaf(a[i * 3 + 0], b[j], p1));
        }
}

// Closed form of the scan-fill state machine of cv::FillConvexPoly for a triangle (drawing.cpp:1163-1252):
// which rows are span-filled and where the two edge walkers are on each of them.
__device__ void make_fill_record(TriRaster& r, int img_h) {
    const int n = 3;
    int imin = 0, ymin = r.vy[0], ymax = r.vy[0];
#pragma unroll
    for (int i = 1; i < n; ++i) {
        if (r.vy[i] < ymin) { ymin = r.vy[i]; imin = i; }
        ymax = max(ymax, r.vy[i]);
    }
    ymax = min(ymax, img_h - 1);
    int idx[2] = {imin, imin}, ye[2] = {ymin, ymin}, nseg[2] = {0, 0};
    const int di[2] = {1, n - 1};
    r.sw[0] = r.sw[1] = 32767;
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) { r.x0[i][j] = -65536; r.dx[i][j] = 0; }
    int y = ymin, edges = n, yend = ymax + 1;
    for (;;) {
        for (int i = 0; i < 2; ++i) {
            if (y < ye[i]) continue;
            int from = idx[i], to = from + di[i];
            if (to >= n) to -= n;
            while (edges-- > 0) {
                int ty = r.vy[to];
                if (ty > y) {
                    long long xs = (long long)r.vx[from] << 16, xe = (long long)r.vx[to] << 16;
                    int j = min(nseg[i], 1);
                    r.x0[i][j] = (int)xs;
                    r.dx[i][j] = (int)(((xe - xs) * 2 + (ty - y)) / (2 * (ty - y)));
                    if (j == 1) r.sw[i] = (short)y;
                    nseg[i]++;
                    ye[i] = ty;
                    idx[i] = to;
                    break;
                }
                from = to;
                to += di[i];
                if (to >= n) to -= n;
            }
        }
        if (edges < 0) { yend = y; break; }
        int ynext = min(ye[0], ye[1]);
        if (ynext > ymax) break;           // both walkers run to the last row
        y = ynext;
    }
    r.ymin = (short)ymin;
    r.yend = (short)yend;
}

// grid (ceil(max_tri/128), frames)
__global__ void k_tri_geometry(const int3* __restrict__ tri_idx, const FrameParams* __restrict__ fp,
                               const float2* __restrict__ p1, size_t p1_frame_stride, const float2* __restrict__ p2,
                               const float2* __restrict__ morphed, size_t morphed_frame_stride, int max_tri, int img_h,
                               TriInverse* __restrict__ inv_out, TriRaster* __restrict__ rast_out) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    int f = blockIdx.y;
    FrameParams P = fp[f];
    if (t >= P.n_tri) return;
    int3 id = tri_idx[P.tri_base + t];
    const float2* s1 = p1 + (size_t)f * p1_frame_stride;
    const float2* mp = morphed + (size_t)f * morphed_frame_stride;
    int v[3] = {id.x, id.y, id.z};
    float P1[9], P2[9];
    TriRaster R;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        float2 a = s1[v[k]], b = p2[v[k]], m = mp[v[k]];
        // cv::Point(float, float): truncation toward zero (reference src/algo.cpp:89)
        P1[k] = (float)__float2int_rz(a.x); P1[3 + k] = (float)__float2int_rz(a.y); P1[6 + k] = 1.f;
        P2[k] = (float)__float2int_rz(b.x); P2[3 + k] = (float)__float2int_rz(b.y); P2[6 + k] = 1.f;
        R.vx[k] = __float2int_rz(m.x); R.vy[k] = __float2int_rz(m.y);
    }
    make_fill_record(R, img_h);
    rast_out[(size_t)f * max_tri + t] = R;

    float iP1[9], H[9], iH[9], M1[9], M2[9];
    inv3(P1, iP1);
    mul3(P2, iP1, H);          // H = P2 * inv(P1)                       (algo.cpp:110)
    inv3(H, iH);
    const float r = P.shape, omr = P.one_minus_r;
#pragma unroll
    for (int i = 0; i < 9; ++i) {
        float eye = (i % 4 == 0) ? 1.f : 0.f;
        M1[i] = fmaf(H[i], r, __fmul_rn(eye, omr));      // eye*(1-r) + H*r    via scaleAdd_32f (algo.cpp:130)
        M2[i] = fmaf(iH[i], omr, __fmul_rn(eye, r));     // eye*r + inv(H)*(1-r)               (algo.cpp:131)
    }
    TriInverse out;
    inv3(M1, out.a);
    inv3(M2, out.b);
    out.pad[0] = out.pad[1] = 0.f;
    inv_out[(size_t)f * max_tri + t] = out;
}

// ------------------------------------------------------------------------------------------------------------------
// Exact fillConvexPoly(img32S, tri, i+1) for all triangles of a frame chunk, one warp per triangle.
// 8-connected Bresenham outline (closed form of LineIterator, OCV imgproc.hpp:4956-4970 / drawing.cpp:159-260):
// step i of an edge sits at major = start + i, minor = start + sign * ((2*minor_len*i + major_len - 1) / (2*major_len)).
// grid.x covers warps over max_tri, grid.y = frames.
__global__ void k_raster_triangles(const TriRaster* __restrict__ rast, const FrameParams* __restrict__ fp, int max_tri,
                                   int* __restrict__ tri_map, int w, int h) {
    const int lane = threadIdx.x & 31;
    const int t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int f = blockIdx.y;
    if (t >= fp[f].n_tri) return;
    const TriRaster R = rast[(size_t)f * max_tri + t];
    int* map = tri_map + (size_t)f * w * h;
    const int color = t + 1;

#pragma unroll
    for (int e = 0; e < 3; ++e) {
        int x0 = R.vx[(e + 2) % 3], y0 = R.vy[(e + 2) % 3], x1 = R.vx[e], y1 = R.vy[e];
        int dx = x1 - x0, dy = y1 - y0, sy = 1;
        if (dx < 0) { dx = -dx; dy = -dy; x0 = x1; y0 = y1; }
        if (dy < 0) { dy = -dy; sy = -1; }
        const bool steep = dy > dx;
        const int major = steep ? dy : dx, minor = steep ? dx : dy;
        for (int i = lane; i <= major; i += 32) {
            int m = major > 0 ? (2 * minor * i + major - 1) / (2 * major) : 0;
            int x = steep ? x0 + m : x0 + i;
            int y = steep ? y0 + sy * i : y0 + sy * m;
            if ((unsigned)x < (unsigned)w && (unsigned)y < (unsigned)h) atomicMax(&map[(size_t)y * w + x], color);
        }
    }
    for (int y = R.ymin + lane; y < R.yend; y += 32) {
        if (y < 0) continue;
        long long xa, xb;
        {
            int j = y >= R.sw[0] ? 1 : 0, ys = j ? R.sw[0] : R.ymin;
            xa = (long long)R.x0[0][j] + (long long)(y - ys) * R.dx[0][j];
            j = y >= R.sw[1] ? 1 : 0; ys = j ? R.sw[1] : R.ymin;
            xb = (long long)R.x0[1][j] + (long long)(y - ys) * R.dx[1][j];
        }
        long long xl = xa > xb ? xb : xa, xr = xa > xb ? xa : xb;
        int xx1 = (int)((xl + 32768) >> 16), xx2 = (int)((xr + 32768) >> 16);
        if (xx2 >= 0 && xx1 < w) {
            xx1 = max(xx1, 0);
            xx2 = min(xx2, w - 1);
            for (int x = xx1; x <= xx2; ++x) atomicMax(&map[(size_t)y * w + x], color);
        }
    }
}

// ------------------------------------------------------------------------------------------------------------------
void launch_clip_points(cudaStream_t st, const float2* in, float2* out, int n, int cols, int rows) {
    if (n > 0) k_clip_points<<<div_up(n, 256), 256, 0, st>>>(in, out, n, cols, rows);
}

void launch_lerp_points(cudaStream_t st, const float2* p1, size_t p1_frame_stride, const float2* p2,
                        const FrameParams* fp, float2* out, size_t out_frame_stride, int n, int frames, int cols,
                        int rows) {
    if (n > 0)
        k_lerp_points<<<dim3(div_up(n, 256), frames), 256, 0, st>>>(p1, p1_frame_stride, p2, fp, out, out_frame_stride,
                                                                    n, cols, rows);
}

void launch_tri_geometry(cudaStream_t st, const int3* tri_idx, const FrameParams* fp, const float2* p1,
                         size_t p1_frame_stride, const float2* p2, const float2* morphed, size_t morphed_frame_stride,
                         int max_tri, int tri_in_chunk_max, int frames, int img_h, TriInverse* inv_out,
                         TriRaster* rast_out) {
    if (tri_in_chunk_max > 0)
        k_tri_geometry<<<dim3(div_up(tri_in_chunk_max, 128), frames), 128, 0, st>>>(
            tri_idx, fp, p1, p1_frame_stride, p2, morphed, morphed_frame_stride, max_tri, img_h, inv_out, rast_out);
}

void launch_raster_triangles(cudaStream_t st, const TriRaster* rast, const FrameParams* fp, int max_tri,
                             int tri_in_chunk_max, int frames, int* tri_map, int w, int h) {
    if (tri_in_chunk_max > 0)
        k_raster_triangles<<<dim3(div_up(tri_in_chunk_max * 32, 256), frames), 256, 0, st>>>(rast, fp, max_tri, tri_map, w, h);
}

}  // namespace poppy
