// TEST INFRASTRUCTURE ONLY — CPU restatement ("port") of the reference morph-render path.
//
// Plain C++ (no OpenCV), one straightforward whole-image pass per reference stage, written from the reference
// sources and from the arithmetic of the OpenCV 4.6.0 primitives they call. It exists to
// (1) check the CUDA path stage by stage and (2) serve as a CPU baseline of kind "port". It is pinned against
// the real reference (oracle/_ref/libpoppy_ref.so, built from the unmodified reference TUs) by
// tests/test_oracle_vs_reference.py and against tests/golden/*.npz, which were generated from that library.
// Nothing under poppy_b200/ may call into this file.
//
// Topology (cv::Subdiv2D + get_triangle_indices, reference src/algo.cpp:205-213) is an *input* here: the
// triangle index list comes from the host stage under test or from the reference library.
//
// Build: g++ -O2 -std=c++17 -mfma -ffp-contract=off -fPIC -shared   (fmaf must be a single-rounding FMA; no
// implicit contraction anywhere else).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

namespace {

// ---------------------------------------------------------------------------------------------------------------
// helpers
// ---------------------------------------------------------------------------------------------------------------
// cv::borderInterpolate(BORDER_REFLECT_101)  — OCV core/src/copy.cpp:748-793
int reflect101(int p, int len) {
    if ((unsigned)p < (unsigned)len) return p;
    if (len == 1) return 0;
    do {
        if (p < 0) p = -p;               // -p - 1 + delta, delta = 1
        else p = 2 * len - 2 - p;        // len - 1 - (p - len) - delta
    } while ((unsigned)p >= (unsigned)len);
    return p;
}

// cvRound(float) on x86: cvtss2si, round-half-even, INT_MIN for NaN / out of range — OCV core fast_math.hpp:311
int cv_round(float v) {
    if (!(std::fabs(v) < 2147483648.0f)) return INT32_MIN;
    return (int)std::nearbyintf(v);
}

struct Tri2i { int x[3], y[3]; };

// ---------------------------------------------------------------------------------------------------------------
// a2/a3: clip_points (reference src/util.cpp:453-460) and morph_points (src/algo.cpp:50-58)
// ---------------------------------------------------------------------------------------------------------------
void clip_point(float& x, float& y, int cols, int rows) {
    x = x > cols ? cols - 1 : x;
    y = y > rows ? rows - 1 : y;
    x = x < 0 ? 0 : x;
    y = y < 0 ? 0 : y;
}

float lerp_coord(float a, float b, float s) {
    float sb = s * b;                                   // float x float product, rounded to float
    return (float)((1.0 - (double)s) * (double)a + (double)sb);
}

// ---------------------------------------------------------------------------------------------------------------
// a6: cv::fillConvexPoly on a 32SC1 image, shift 0, LINE_8 — OCV imgproc/src/drawing.cpp:1093-1255 (scan fill),
// :262-297 + :159-260 + imgproc.hpp:4956-4970 (8-connected line, left-to-right). Vertices are inside the image
// (they are clipped points that cv::Subdiv2D accepted), so clipLine is the identity and is not restated.
// ---------------------------------------------------------------------------------------------------------------
void draw_line8(int32_t* img, int w, int h, int x0, int y0, int x1, int y1, int32_t color) {
    int dx = x1 - x0, dy = y1 - y0, sx = 1, sy = 1;
    int px = x0, py = y0;
    if (dx < 0) { dx = -dx; dy = -dy; px = x1; py = y1; }         // always walk left to right
    if (dy < 0) { dy = -dy; sy = -1; }
    bool steep = dy > dx;
    int major = steep ? dy : dx, minor = steep ? dx : dy;
    int err = major - 2 * minor;
    for (int i = 0; i <= major; ++i) {
        if ((unsigned)px < (unsigned)w && (unsigned)py < (unsigned)h) img[(size_t)py * w + px] = color;
        bool neg = err < 0;
        err += -2 * minor + (neg ? 2 * major : 0);
        if (steep) { py += sy; if (neg) px += sx; }
        else       { px += sx; if (neg) py += sy; }
    }
}

void fill_convex_tri(int32_t* img, int w, int h, const Tri2i& t, int32_t color) {
    const int n = 3;
    const int64_t ONE = 1 << 16;
    int imin = 0;
    int64_t ymin = t.y[0], ymax = t.y[0], xmin = t.x[0], xmax = t.x[0];
    int px = t.x[n - 1], py = t.y[n - 1];
    for (int i = 0; i < n; ++i) {
        if (t.y[i] < ymin) { ymin = t.y[i]; imin = i; }
        ymax = std::max<int64_t>(ymax, t.y[i]);
        xmax = std::max<int64_t>(xmax, t.x[i]);
        xmin = std::min<int64_t>(xmin, t.x[i]);
        draw_line8(img, w, h, px, py, t.x[i], t.y[i], color);
        px = t.x[i]; py = t.y[i];
    }
    if (xmax < 0 || ymax < 0 || xmin >= w || ymin >= h) return;
    ymax = std::min<int64_t>(ymax, h - 1);

    struct Walker { int idx, di, ye; int64_t x, dx; } e[2];
    int y = (int)ymin, edges = n;
    for (int i = 0; i < 2; ++i) { e[i].idx = imin; e[i].ye = y; e[i].x = -ONE; e[i].dx = 0; }
    e[0].di = 1; e[1].di = n - 1;
    do {
        for (int i = 0; i < 2; ++i) {
            if (y < e[i].ye) continue;
            int from = e[i].idx, to = from + e[i].di;
            if (to >= n) to -= n;
            while (edges-- > 0) {
                int ty = t.y[to];
                if (ty > y) {
                    int64_t xs = (int64_t)t.x[from] << 16, xe = (int64_t)t.x[to] << 16;
                    e[i].ye = ty;
                    e[i].dx = ((xe - xs) * 2 + (ty - y)) / (2 * (ty - y));
                    e[i].x = xs;
                    e[i].idx = to;
                    break;
                }
                from = to;
                to += e[i].di;
                if (to >= n) to -= n;
            }
        }
        if (edges < 0) break;
        if (y >= 0) {
            int l = e[0].x > e[1].x ? 1 : 0, r = 1 - l;
            int xx1 = (int)((e[l].x + (ONE >> 1)) >> 16), xx2 = (int)((e[r].x + (ONE >> 1)) >> 16);
            if (xx2 >= 0 && xx1 < w) {
                xx1 = std::max(xx1, 0);
                xx2 = std::min(xx2, w - 1);
                for (int x = xx1; x <= xx2; ++x) img[(size_t)y * w + x] = color;
            }
        }
        e[0].x += e[0].dx;
        e[1].x += e[1].dx;
    } while (++y <= (int)ymax);
}

// ---------------------------------------------------------------------------------------------------------------
// a7/a8: 3x3 float inverse (OCV core/src/lapack.cpp:760-763,965-995,1045), 3x3 product
// (core/src/matmul.simd.hpp:827-858 as evaluated by the FMA-dispatched TU), and the (1-r)I + rH blends
// (core/src/matmul.simd.hpp:1934-1948 scaleAdd_32f)
// ---------------------------------------------------------------------------------------------------------------
void inv3(const float* m, float* o) {
    auto M = [&](int r, int c) { return m[r * 3 + c]; };
    double d = M(0, 0) * ((double)M(1, 1) * M(2, 2) - (double)M(1, 2) * M(2, 1)) -
               M(0, 1) * ((double)M(1, 0) * M(2, 2) - (double)M(1, 2) * M(2, 0)) +
               M(0, 2) * ((double)M(1, 0) * M(2, 1) - (double)M(1, 1) * M(2, 0));
    if (d == 0.) { std::fill(o, o + 9, 0.f); return; }
    d = 1. / d;
    o[0] = (float)(((double)M(1, 1) * M(2, 2) - (double)M(1, 2) * M(2, 1)) * d);
    o[1] = (float)(((double)M(0, 2) * M(2, 1) - (double)M(0, 1) * M(2, 2)) * d);
    o[2] = (float)(((double)M(0, 1) * M(1, 2) - (double)M(0, 2) * M(1, 1)) * d);
    o[3] = (float)(((double)M(1, 2) * M(2, 0) - (double)M(1, 0) * M(2, 2)) * d);
    o[4] = (float)(((double)M(0, 0) * M(2, 2) - (double)M(0, 2) * M(2, 0)) * d);
    o[5] = (float)(((double)M(0, 2) * M(1, 0) - (double)M(0, 0) * M(1, 2)) * d);
    o[6] = (float)(((double)M(1, 0) * M(2, 1) - (double)M(1, 1) * M(2, 0)) * d);
    o[7] = (float)(((double)M(0, 1) * M(2, 0) - (double)M(0, 0) * M(2, 1)) * d);
    o[8] = (float)(((double)M(0, 0) * M(1, 1) - (double)M(0, 1) * M(1, 0)) * d);
}

void mul3(const float* a, const float* b, float* o) {
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            float p1 = a[i * 3 + 1] * b[3 + j];
            o[i * 3 + j] = std::fmaf(a[i * 3 + 2], b[6 + j], std::fmaf(a[i * 3 + 0], b[j], p1));
        }
}

void tri_to_mat(const Tri2i& t, float* p) {
    for (int i = 0; i < 3; ++i) { p[i] = (float)t.x[i]; p[3 + i] = (float)t.y[i]; p[6 + i] = 1.f; }
}

// H, M1, M2 and the two inverses create_map() applies (reference src/algo.cpp:108-157)
void triangle_matrices(const Tri2i& t1, const Tri2i& t2, float r, float* H, float* M1, float* M2, float* iM1,
                       float* iM2) {
    float P1[9], P2[9], iP1[9], iH[9];
    tri_to_mat(t1, P1);
    tri_to_mat(t2, P2);
    inv3(P1, iP1);
    mul3(P2, iP1, H);
    inv3(H, iH);
    const float one_minus_r = (float)(1.0 - (double)r);
    for (int i = 0; i < 9; ++i) {
        float eye = (i % 4 == 0) ? 1.f : 0.f;
        M1[i] = std::fmaf(H[i], r, eye * one_minus_r);
        M2[i] = std::fmaf(iH[i], one_minus_r, eye * r);
    }
    inv3(M1, iM1);
    inv3(M2, iM2);
}

// ---------------------------------------------------------------------------------------------------------------
// a9: create_map (reference src/algo.cpp:159-175; first-party code, compiled without FMA)
// ---------------------------------------------------------------------------------------------------------------
void map_pixel(const float* h, int x, int y, float& mx, float& my) {
    float fx = (float)x, fy = (float)y;
    float z = h[6] * fx + h[7] * fy + h[8];
    if (z == 0) z = (float)0.00001;
    mx = (h[0] * fx + h[1] * fy + h[2]) / z;
    my = (h[3] * fx + h[4] * fy + h[5]) / z;
}

// ---------------------------------------------------------------------------------------------------------------
// a10: cv::remap 8UC3, INTER_LINEAR, BORDER_CONSTANT(0) — OCV imgproc/src/imgwarp.cpp:1197-1234 (map split),
// :213-287 (weight table), :648-856 (sampling and borders)
// ---------------------------------------------------------------------------------------------------------------
void bilinear_weights(int fx, int fy, int* wt) {
    if (fx == 0 && fy == 0) { wt[0] = 32767; wt[1] = 0; wt[2] = 0; wt[3] = 1; return; }   // saturated + fix-up
    wt[0] = 32 * (32 - fy) * (32 - fx);
    wt[1] = 32 * (32 - fy) * fx;
    wt[2] = 32 * fy * (32 - fx);
    wt[3] = 32 * fy * fx;
}

void remap_pixel(const uint8_t* src, int w, int h, float mx, float my, uint8_t* out) {
    int sx = cv_round(mx * 32.f), sy = cv_round(my * 32.f);
    int X = std::clamp(sx >> 5, -32768, 32767), Y = std::clamp(sy >> 5, -32768, 32767);
    int wt[4];
    bilinear_weights(sx & 31, sy & 31, wt);
    if (X >= w || X + 1 < 0 || Y >= h || Y + 1 < 0) { out[0] = out[1] = out[2] = 0; return; }
    for (int c = 0; c < 3; ++c) {
        auto tap = [&](int xx, int yy) -> int {
            return ((unsigned)xx < (unsigned)w && (unsigned)yy < (unsigned)h) ? src[((size_t)yy * w + xx) * 3 + c] : 0;
        };
        int v = tap(X, Y) * wt[0] + tap(X + 1, Y) * wt[1] + tap(X, Y + 1) * wt[2] + tap(X + 1, Y + 1) * wt[3];
        out[c] = (uint8_t)std::clamp((v + 16384) >> 15, 0, 255);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// a12: blend mask (reference src/algo.cpp:250-258): gray of gabor2 (OCV imgproc/src/color_rgb.simd.hpp:594-642),
// 1 - gray, addWeighted (OCV core/src/arithm.simd.hpp:1161-1204, 1705-1770), clamp to [0,1]
// ---------------------------------------------------------------------------------------------------------------
void mask_basis(const float* gabor, int w, int h, float* m2) {
    // RGB2Gray<float> runs per row: 8-lane vector body fma(r,cr, fma(g,cg, b*cb)); the scalar tail
    // "b*cb + g*cg + r*cr" is contracted by the compiler of the FMA-dispatched TU into fma(r,cr, fma(b,cb, g*cg))
    // (established against the reference library, tests/test_oracle_vs_reference.py::test_mask_tail_columns).
    const int vec_end = w & ~7;
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            const float* p = gabor + ((size_t)y * w + x) * 3;
            float g = x < vec_end ? std::fmaf(p[2], 0.299f, std::fmaf(p[1], 0.587f, p[0] * 0.114f))
                                  : std::fmaf(p[2], 0.299f, std::fmaf(p[0], 0.114f, p[1] * 0.587f));
            m2[(size_t)y * w + x] = 1.0f - g;
        }
}

void blend_mask(const float* m2, size_t n, double mask_ratio, float* out) {
    // addWeighted(ones, 1-mr, m2, -mr, 0) on 32F data runs in double: the vector body widens to f64 and evaluates
    // fma(1.0, alpha, fma(m2, beta, 0.0)) (arithm.simd.hpp:1161-1204,1721-1730), i.e. alpha + round(m2*beta), then
    // narrows. The <16-element scalar tail of the (continuous) matrix may contract the product into the sum; the
    // two differ only below double precision, so one formula is restated.
    const double alpha = 1.0 - mask_ratio, beta = -mask_ratio;
    for (size_t i = 0; i < n; ++i) {
        double prod = (double)m2[i] * beta;
        float v = (float)(alpha + prod);
        out[i] = v < 0 ? 0.f : (v > 1 ? 1.f : v);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// a13: cv::pyrDown / cv::pyrUp, 32F — OCV imgproc/src/pyramids.cpp:745-900 (+344-402 H, 503-521 V) and 903-1005
// (+700-718). The SSE-baseline SIMD bodies associate differently from the scalar tails; both are restated and
// selected by element position exactly as the reference loops do.
// ---------------------------------------------------------------------------------------------------------------
struct Img {
    int w = 0, h = 0, c = 0;
    std::vector<float> d;
    Img() {}
    Img(int w_, int h_, int c_) : w(w_), h(h_), c(c_), d((size_t)w_ * h_ * c_) {}
    float& at(int x, int y, int k) { return d[((size_t)y * w + x) * c + k]; }
    float at(int x, int y, int k) const { return d[((size_t)y * w + x) * c + k]; }
};

Img pyr_down(const Img& s, int dw, int dh) {
    const int cn = s.c;
    Img o(dw, dh, cn);
    const int width0 = std::min((s.w - 3) / 2 + 1, dw);       // C++ division truncates toward zero, as in OpenCV
    // which destination pixels the horizontal SIMD body covers
    auto h_simd = [&](int px) {
        if (cn == 3) return px >= 1 && px <= width0 - 2;
        if (cn == 1) { int k = width0 >= 5 ? (width0 - 5) / 4 + 1 : 0; return px >= 1 && px < 1 + 4 * k; }
        return false;
    };
    std::vector<float> rows((size_t)5 * dw * cn);
    const int vec_end = (dw * cn) & ~3;                         // vertical SIMD body: 4 floats per step
    for (int y = 0; y < dh; ++y) {
        for (int k = 0; k < 5; ++k) {
            int sy = reflect101(2 * y - 2 + k, s.h);
            float* row = rows.data() + (size_t)k * dw * cn;
            for (int px = 0; px < dw; ++px)
                for (int ch = 0; ch < cn; ++ch) {
                    float t[5];
                    for (int j = 0; j < 5; ++j) t[j] = s.at(reflect101(2 * px - 2 + j, s.w), sy, ch);
                    float v;
                    if (h_simd(px)) v = t[2] * 6.f + ((t[1] + t[3]) * 4.f + (t[0] + t[4]));
                    else v = t[2] * 6.f + (t[1] + t[3]) * 4.f + t[0] + t[4];
                    row[px * cn + ch] = v;
                }
        }
        const float *r0 = rows.data(), *r1 = r0 + (size_t)dw * cn, *r2 = r1 + (size_t)dw * cn,
                    *r3 = r2 + (size_t)dw * cn, *r4 = r3 + (size_t)dw * cn;
        for (int e = 0; e < dw * cn; ++e) {
            float v;
            if (e < vec_end) v = ((r1[e] + r3[e] + r2[e]) * 4.f + (r0[e] + r4[e] + (r2[e] + r2[e]))) * (1.f / 256);
            else v = (r2[e] * 6.f + (r1[e] + r3[e]) * 4.f + r0[e] + r4[e]) * (1.f / 256);
            o.d[(size_t)y * dw * cn + e] = v;
        }
    }
    return o;
}

Img pyr_up(const Img& s, int dw, int dh) {
    const int cn = s.c;
    Img o(dw, dh, cn);
    // horizontal pass of one source row into 2*sw (+1) entries
    auto hrow = [&](int sy, std::vector<float>& row) {
        row.assign((size_t)(2 * s.w + 1) * cn, 0.f);
        for (int ch = 0; ch < cn; ++ch) {
            auto S = [&](int x) { return s.at(x, sy, ch); };
            auto R = [&](int x) -> float& { return row[(size_t)x * cn + ch]; };
            if (s.w == 1) { R(0) = R(1) = S(0) * 8.f; continue; }
            R(0) = S(0) * 6.f + S(1) * 2.f;
            R(1) = (S(0) + S(1)) * 4.f;
            int n = s.w;
            R(2 * n - 2) = S(n - 2) + S(n - 1) * 7.f;
            R(2 * n - 1) = S(n - 1) * 8.f;
            if (dw > 2 * n) R(dw - 1) = R(2 * n - 1);
            for (int x = 1; x < n - 1; ++x) {
                R(2 * x) = S(x - 1) + S(x) * 6.f + S(x + 1);
                R(2 * x + 1) = (S(x) + S(x + 1)) * 4.f;
            }
        }
    };
    std::vector<float> r0, r1, r2;
    for (int y = 0; y < s.h; ++y) {
        hrow(reflect101(2 * (y - 1), 2 * s.h) / 2, r0);
        hrow(y, r1);
        hrow(reflect101(2 * (y + 1), 2 * s.h) / 2, r2);
        int y0 = 2 * y, y1 = std::min(2 * y + 1, dh - 1);
        for (int e = 0; e < dw * cn; ++e) {
            float t1 = ((r1[e] + r2[e]) * 4.f) * (1.f / 64);
            float t0 = (r0[e] + r1[e] * 6.f + r2[e]) * (1.f / 64);
            o.d[(size_t)y1 * dw * cn + e] = t1;      // odd row first, the even row wins when they alias
            o.d[(size_t)y0 * dw * cn + e] = t0;
        }
    }
    if (dh > 2 * s.h)
        for (int e = 0; e < dw * cn; ++e) o.d[(size_t)(2 * s.h) * dw * cn + e] = o.d[(size_t)(2 * s.h - 2) * dw * cn + e];
    return o;
}

// reference src/blend.hpp:11-91
Img laplacian_blend(const Img& l, const Img& r, const Img& mask, int levels) {
    std::vector<Img> gl{l}, gr{r}, gm{mask};
    for (int k = 0; k < levels; ++k) {
        int dw = (gl[k].w + 1) / 2, dh = (gl[k].h + 1) / 2;
        gl.push_back(pyr_down(gl[k], dw, dh));
        gr.push_back(pyr_down(gr[k], dw, dh));
        gm.push_back(pyr_down(gm[k], dw, dh));
    }
    auto blend = [&](const Img& a, const Img& b, const Img& m) {
        Img o(a.w, a.h, 3);
        for (int y = 0; y < a.h; ++y)
            for (int x = 0; x < a.w; ++x) {
                float mv = m.at(x, y, 0), anti = 1.0f - mv;
                for (int c = 0; c < 3; ++c) o.at(x, y, c) = a.at(x, y, c) * mv + b.at(x, y, c) * anti;
            }
        return o;
    };
    auto lap = [&](const Img& fine, const Img& coarse) {
        Img up = pyr_up(coarse, fine.w, fine.h), o(fine.w, fine.h, fine.c);
        for (size_t i = 0; i < o.d.size(); ++i) o.d[i] = fine.d[i] - up.d[i];
        return o;
    };
    Img cur = blend(gl[levels], gr[levels], gm[levels]);
    for (int k = levels - 1; k >= 0; --k) {
        Img res = blend(lap(gl[k], gl[k + 1]), lap(gr[k], gr[k + 1]), gm[k]);
        Img up = pyr_up(cur, res.w, res.h);
        for (size_t i = 0; i < up.d.size(); ++i) up.d[i] = up.d[i] + res.d[i];
        cur = std::move(up);
    }
    return cur;
}

// ---------------------------------------------------------------------------------------------------------------
// a14: unsharp_mask (reference src/util.cpp:113-148): GaussianBlur sigma 1 -> 9 taps (OCV imgproc/src/
// smooth.dispatch.cpp:81-198,289; row filter filter.simd.hpp:1625-1735 + 2446-2488, symmetric column filter
// :2712-2760, both FMA chains in the dispatched TU), subtract, medianBlur 3 (median_blur.simd.hpp:677-745,
// replicate border), norm threshold, scaled add.
// ---------------------------------------------------------------------------------------------------------------
const uint32_t kGauss9Bits[5] = {0x3ecc4252u, 0x3e77c75du, 0x3d5d25cdu, 0x3b913926u, 0x390c54e2u};  // centre .. edge
float gauss_tap(int k) { float f; std::memcpy(&f, &kGauss9Bits[std::abs(k)], 4); return f; }

Img gaussian9(const Img& s) {
    // Both filters run 8 floats at a time as FMA chains; the last (w*cn) % 8 elements of every row fall to the
    // generic scalar loops (filter.simd.hpp:2477-2487, 2753-2759), which the reference build does not contract:
    // there each product and each sum is rounded separately (established against the reference library).
    Img t(s.w, s.h, s.c), o(s.w, s.h, s.c);
    const int n = s.w * s.c, tail_from = n - n % 8;
    for (int y = 0; y < s.h; ++y)
        for (int x = 0; x < s.w; ++x)
            for (int c = 0; c < s.c; ++c) {
                const bool fused = x * s.c + c < tail_from;
                if (s.w == 1) { t.at(x, y, c) = s.at(x, y, c); continue; }   // 1-pixel-wide: kernel shrinks to [1]
                float acc = gauss_tap(-4) * s.at(reflect101(x - 4, s.w), y, c);
                for (int k = -3; k <= 4; ++k) {
                    float v = s.at(reflect101(x + k, s.w), y, c);
                    acc = fused ? std::fmaf(v, gauss_tap(k), acc) : acc + gauss_tap(k) * v;
                }
                t.at(x, y, c) = acc;
            }
    for (int y = 0; y < s.h; ++y)
        for (int x = 0; x < s.w; ++x)
            for (int c = 0; c < s.c; ++c) {
                const bool fused = x * s.c + c < tail_from;
                if (s.h == 1) { o.at(x, y, c) = t.at(x, y, c); continue; }   // smooth.dispatch.cpp:621-628
                float acc = fused ? std::fmaf(gauss_tap(0), t.at(x, y, c), 0.f) : gauss_tap(0) * t.at(x, y, c) + 0.f;
                for (int k = 1; k <= 4; ++k) {
                    float pair = t.at(x, reflect101(y + k, s.h), c) + t.at(x, reflect101(y - k, s.h), c);
                    acc = fused ? std::fmaf(gauss_tap(k), pair, acc) : acc + gauss_tap(k) * pair;
                }
                o.at(x, y, c) = acc;
            }
    return o;
}

Img median3(const Img& s) {
    Img o(s.w, s.h, s.c);
    for (int y = 0; y < s.h; ++y)
        for (int x = 0; x < s.w; ++x)
            for (int c = 0; c < s.c; ++c) {
                float v[9];
                int n = 0;
                for (int dy = -1; dy <= 1; ++dy)
                    for (int dx = -1; dx <= 1; ++dx)
                        v[n++] = s.at(std::clamp(x + dx, 0, s.w - 1), std::clamp(y + dy, 0, s.h - 1), c);
                std::nth_element(v, v + 4, v + 9);
                o.at(x, y, c) = v[4];
            }
    return o;
}

Img unsharp(const Img& s, float amount, float threshold) {
    Img blur = gaussian9(s), diff(s.w, s.h, s.c);
    for (size_t i = 0; i < diff.d.size(); ++i) diff.d[i] = s.d[i] - blur.d[i];
    Img med = median3(diff), o = s;
    for (size_t p = 0; p < (size_t)s.w * s.h; ++p) {
        const float* dv = &med.d[p * 3];
        double nn = 0;
        for (int c = 0; c < 3; ++c) nn += (double)dv[c] * dv[c];
        if (std::sqrt(nn) >= threshold)
            for (int c = 0; c < 3; ++c) o.d[p * 3 + c] = s.d[p * 3 + c] + amount * dv[c];
    }
    return o;
}

uint8_t to_u8(float v) {                                        // convertTo(CV_8U, 255): cvRound(v*255) saturated
    int r = cv_round(v * 255.f);
    return (uint8_t)std::clamp(r, 0, 255);
}

Img wrap(const float* p, int w, int h, int c) {
    Img i(w, h, c);
    std::memcpy(i.d.data(), p, i.d.size() * sizeof(float));
    return i;
}

}  // namespace

// =================================================================================================================
// C ABI (loaded by oracle/port.py)
// =================================================================================================================
extern "C" {

struct poppy_oracle_stage_dump {          // same layout as poppy_ref_stage_dump (oracle/ref_shim.cpp)
    float* morphed_points; int32_t* tri_idx; int32_t max_tri; int32_t n_tri; int32_t* tri_map;
    float* hom; float* m1; float* m2; float* mapx1; float* mapy1; float* mapx2; float* mapy2;
    uint8_t* warped1; uint8_t* warped2; float* mask; float* lap_blend; uint8_t* dst;
};

void poppy_oracle_clip_points(float* xy, int n, int w, int h) {
    for (int i = 0; i < n; ++i) clip_point(xy[2 * i], xy[2 * i + 1], w, h);
}

// clip both sets, lerp, clip the result (reference src/algo.cpp:185,191,202,205)
void poppy_oracle_morph_points(const float* p1, const float* p2, int n, float s, int w, int h, float* out) {
    for (int i = 0; i < n; ++i) {
        float ax = p1[2 * i], ay = p1[2 * i + 1], bx = p2[2 * i], by = p2[2 * i + 1];
        clip_point(ax, ay, w, h);
        clip_point(bx, by, w, h);
        float x = lerp_coord(ax, bx, s), y = lerp_coord(ay, by, s);
        clip_point(x, y, w, h);
        out[2 * i] = x; out[2 * i + 1] = y;
    }
}

void poppy_oracle_fill_triangles(int32_t* img, int w, int h, const int32_t* tri_xy, int n_tri) {
    for (int i = 0; i < n_tri; ++i) {
        Tri2i t;
        for (int k = 0; k < 3; ++k) { t.x[k] = tri_xy[6 * i + 2 * k]; t.y[k] = tri_xy[6 * i + 2 * k + 1]; }
        fill_convex_tri(img, w, h, t, i + 1);
    }
}

void poppy_oracle_triangle_matrices(const int32_t* tri1_xy, const int32_t* tri2_xy, int n_tri, float r, float* H,
                                    float* M1, float* M2, float* iM1, float* iM2) {
    for (int i = 0; i < n_tri; ++i) {
        Tri2i a, b;
        for (int k = 0; k < 3; ++k) {
            a.x[k] = tri1_xy[6 * i + 2 * k]; a.y[k] = tri1_xy[6 * i + 2 * k + 1];
            b.x[k] = tri2_xy[6 * i + 2 * k]; b.y[k] = tri2_xy[6 * i + 2 * k + 1];
        }
        triangle_matrices(a, b, r, H + 9 * i, M1 + 9 * i, M2 + 9 * i, iM1 + 9 * i, iM2 + 9 * i);
    }
}

void poppy_oracle_remap(const uint8_t* src, int w, int h, const float* mapx, const float* mapy, int dw, int dh,
                        uint8_t* dst) {
    for (size_t i = 0; i < (size_t)dw * dh; ++i) remap_pixel(src, w, h, mapx[i], mapy[i], dst + 3 * i);
}

void poppy_oracle_mask(const float* gabor2, int w, int h, double mask_ratio, float* out) {
    std::vector<float> m2((size_t)w * h);
    mask_basis(gabor2, w, h, m2.data());
    blend_mask(m2.data(), m2.size(), mask_ratio, out);
}

void poppy_oracle_pyr_down(const float* src, int w, int h, int cn, float* dst, int dw, int dh) {
    Img o = pyr_down(wrap(src, w, h, cn), dw, dh);
    std::memcpy(dst, o.d.data(), o.d.size() * 4);
}

void poppy_oracle_pyr_up(const float* src, int w, int h, int cn, float* dst, int dw, int dh) {
    Img o = pyr_up(wrap(src, w, h, cn), dw, dh);
    std::memcpy(dst, o.d.data(), o.d.size() * 4);
}

void poppy_oracle_lap_blend(const float* l, const float* r, const float* mask, int w, int h, int levels, float* out) {
    Img o = laplacian_blend(wrap(l, w, h, 3), wrap(r, w, h, 3), wrap(mask, w, h, 1), levels);
    std::memcpy(out, o.d.data(), o.d.size() * 4);
}

void poppy_oracle_gaussian9(const float* src, int w, int h, int cn, float* out) {
    Img o = gaussian9(wrap(src, w, h, cn));
    std::memcpy(out, o.d.data(), o.d.size() * 4);
}

void poppy_oracle_median3(const float* src, int w, int h, int cn, float* out) {
    Img o = median3(wrap(src, w, h, cn));
    std::memcpy(out, o.d.data(), o.d.size() * 4);
}

void poppy_oracle_unsharp(const float* src, int w, int h, float amount, float threshold, float* out) {
    Img o = unsharp(wrap(src, w, h, 3), amount, threshold);
    std::memcpy(out, o.d.data(), o.d.size() * 4);
}

// One frame of the path, reference src/algo.cpp:178-265, given the triangle index list of the morphed points.
int poppy_oracle_morph_frame(int w, int h, const uint8_t* bgr1, const uint8_t* bgr2, const float* gabor2,
                             const float* pts1, const float* pts2, int n, const int32_t* tri_idx, int n_tri,
                             double shape, double mask_ratio, int levels, poppy_oracle_stage_dump* d) {
    const float s = (float)shape;
    std::vector<float> c1(pts1, pts1 + 2 * n), c2(pts2, pts2 + 2 * n), mp(2 * n);
    poppy_oracle_clip_points(c1.data(), n, w, h);
    poppy_oracle_clip_points(c2.data(), n, w, h);
    poppy_oracle_morph_points(c1.data(), c2.data(), n, s, w, h, mp.data());
    if (d->morphed_points) std::memcpy(d->morphed_points, mp.data(), mp.size() * 4);
    d->n_tri = n_tri;
    if (d->tri_idx) std::memcpy(d->tri_idx, tri_idx, (size_t)std::min(n_tri, d->max_tri) * 12);

    auto gather = [&](const std::vector<float>& p, int i) {
        Tri2i t;
        for (int k = 0; k < 3; ++k) {
            int v = tri_idx[3 * i + k];
            if (v < 0 || v >= n) { t.x[k] = t.y[k] = 0; continue; }
            t.x[k] = (int)p[2 * v]; t.y[k] = (int)p[2 * v + 1];           // Point(float, float): truncation
        }
        return t;
    };
    std::vector<int32_t> tri_map((size_t)w * h, 0);
    std::vector<float> iM1((size_t)n_tri * 9), iM2((size_t)n_tri * 9), H(9), M1(9), M2(9);
    for (int i = 0; i < n_tri; ++i) {
        fill_convex_tri(tri_map.data(), w, h, gather(mp, i), i + 1);
        triangle_matrices(gather(c1, i), gather(c2, i), s, H.data(), M1.data(), M2.data(), &iM1[9 * i], &iM2[9 * i]);
        if (d->hom && i < d->max_tri) std::memcpy(d->hom + 9 * i, H.data(), 36);
        if (d->m1 && i < d->max_tri) std::memcpy(d->m1 + 9 * i, M1.data(), 36);
        if (d->m2 && i < d->max_tri) std::memcpy(d->m2 + 9 * i, M2.data(), 36);
    }
    if (d->tri_map) std::memcpy(d->tri_map, tri_map.data(), tri_map.size() * 4);

    std::vector<uint8_t> w1((size_t)w * h * 3), w2((size_t)w * h * 3);
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            size_t p = (size_t)y * w + x;
            int id = tri_map[p] - 1;
            float ax = (float)x, ay = (float)y, bx = ax, by = ay;
            if (id >= 0) {
                map_pixel(&iM1[9 * id], x, y, ax, ay);
                map_pixel(&iM2[9 * id], x, y, bx, by);
            }
            if (d->mapx1) { d->mapx1[p] = ax; d->mapy1[p] = ay; d->mapx2[p] = bx; d->mapy2[p] = by; }
            remap_pixel(bgr1, w, h, ax, ay, &w1[p * 3]);
            remap_pixel(bgr2, w, h, bx, by, &w2[p * 3]);
        }
    if (d->warped1) std::memcpy(d->warped1, w1.data(), w1.size());
    if (d->warped2) std::memcpy(d->warped2, w2.data(), w2.size());

    Img l(w, h, 3), r(w, h, 3), m(w, h, 1);
    const float inv255 = (float)(1.0 / 255.0);
    for (size_t i = 0; i < l.d.size(); ++i) { l.d[i] = (float)w1[i] * inv255; r.d[i] = (float)w2[i] * inv255; }
    poppy_oracle_mask(gabor2, w, h, mask_ratio, m.d.data());
    if (d->mask) std::memcpy(d->mask, m.d.data(), m.d.size() * 4);
    Img blended = laplacian_blend(l, r, m, levels);
    if (d->lap_blend) std::memcpy(d->lap_blend, blended.d.data(), blended.d.size() * 4);
    const float amount = (float)(1.0 - std::sin(mask_ratio * M_PI));
    Img sharp = unsharp(blended, amount, 0.3f);
    if (d->dst)
        for (size_t i = 0; i < sharp.d.size(); ++i) d->dst[i] = to_u8(sharp.d[i]);
    return 0;
}

}  // extern "C"
