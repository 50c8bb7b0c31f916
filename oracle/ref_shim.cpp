// TEST INFRASTRUCTURE ONLY — C-ABI veneer over the *unmodified* reference morph path.
//
// Linked by oracle/build_ref.sh against the reference's own src/{algo,util,draw,settings}.cpp and the
// vendored OpenCV 4.6.0 core+imgproc into oracle/_ref/libpoppy_ref.so. Nothing under poppy_b200/ may link
// or load this library; only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline / --impl reference
// legs do (through oracle/ref.py).
//
// Entry points:
//   poppy_ref_morph_images   -> poppy::morph_images()            (reference src/algo.cpp:178-273), verbatim call
//   poppy_ref_stages         -> the same path called stage by stage through the reference's exported helpers
//                               (src/algo.hpp:14-25, src/blend.hpp:11-91, src/util.hpp:98) so every stage boundary
//                               of SURVEY.md §3.2 can be dumped
//   poppy_ref_triangulate    -> clip/uniq/Subdiv2D/get_triangle_indices (src/algo.cpp:205-213)
//   poppy_ref_chain          -> the frame-loop recurrence of src/poppy.hpp:177-243 around morph_images()
//   poppy_ref_<primitive>    -> thin wrappers of the OpenCV primitives the path uses (kernel-level parity tests)
#include <cstdint>
#include <cstring>
#include <vector>
#include <string>
#include <exception>

#include <opencv2/core.hpp>
#include <opencv2/imgproc.hpp>

#include "algo.hpp"
#include "util.hpp"
#include "blend.hpp"
#include "settings.hpp"

namespace {
thread_local std::string g_err;

template <class F> int guarded(F&& f) {
    try { f(); return 0; }
    catch (const cv::Exception& e) { g_err = e.what(); return -1; }
    catch (const std::exception& e) { g_err = e.what(); return -2; }
    catch (...) { g_err = "unknown exception"; return -3; }
}

std::vector<cv::Point2f> to_points(const float* xy, int n) {
    std::vector<cv::Point2f> v(n);
    for (int i = 0; i < n; ++i) v[i] = cv::Point2f(xy[2 * i], xy[2 * i + 1]);
    return v;
}
void from_points(const std::vector<cv::Point2f>& v, float* xy) {
    for (size_t i = 0; i < v.size(); ++i) { xy[2 * i] = v[i].x; xy[2 * i + 1] = v[i].y; }
}
cv::Mat wrap_u8c3(const uint8_t* p, int w, int h) { return cv::Mat(h, w, CV_8UC3, const_cast<uint8_t*>(p)); }
cv::Mat wrap_f32(const float* p, int w, int h, int cn) { return cv::Mat(h, w, CV_32FC(cn), const_cast<float*>(p)); }
void copy_out(const cv::Mat& m, void* dst) {
    if (!dst) return;
    cv::Mat c = m.isContinuous() ? m : m.clone();
    std::memcpy(dst, c.data, c.total() * c.elemSize());
}
void mats_out(const std::vector<cv::Mat>& ms, float* dst) {
    if (!dst) return;
    for (size_t i = 0; i < ms.size(); ++i)
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c) dst[i * 9 + r * 3 + c] = ms[i].at<float>(r, c);
}
}  // namespace

extern "C" {

struct poppy_ref_stage_dump {
    float* morphed_points;   // n x 2
    int32_t* tri_idx;        // capacity max_tri x 3
    int32_t max_tri;
    int32_t n_tri;           // out
    int32_t* tri_map;        // w*h
    float* hom;              // T x 9
    float* m1;               // T x 9
    float* m2;               // T x 9
    float* mapx1; float* mapy1; float* mapx2; float* mapy2;   // w*h each
    uint8_t* warped1; uint8_t* warped2;                       // w*h*3
    float* mask;             // w*h
    float* lap_blend;        // w*h*3
    uint8_t* dst;            // w*h*3
};

const char* poppy_ref_last_error(void) { return g_err.c_str(); }

const char* poppy_ref_opencv_version(void) { return CV_VERSION; }

void poppy_ref_set_threads(int n) { cv::setNumThreads(n); }
int poppy_ref_get_threads(void) { return cv::getNumThreads(); }

int poppy_ref_morph_images(int w, int h, const uint8_t* bgr1, const uint8_t* bgr2, const float* gabor2,
                           const float* pts1, const float* pts2, int n, double shape, double mask, int levels,
                           uint8_t* dst_out, float* morphed_out) {
    return guarded([&] {
        poppy::Settings::instance().pyramid_levels = (size_t)levels;
        poppy::Settings::instance().show_gui = false;
        cv::Mat c1 = wrap_u8c3(bgr1, w, h), c2 = wrap_u8c3(bgr2, w, h), g2 = wrap_f32(gabor2, w, h, 3);
        cv::Mat gf1, gf2, dst, last;
        std::vector<cv::Point2f> morphed;
        poppy::morph_images(c1, c2, c1, c2, g2, gf1, gf2, dst, last, morphed, to_points(pts1, n), to_points(pts2, n),
                            shape, mask, 0.0);
        copy_out(dst, dst_out);
        if (morphed_out) from_points(morphed, morphed_out);
    });
}

// Frame-loop recurrence of the reference driver (src/poppy.hpp:177-243, phase < 0 i.e. the CLI default):
// frame j sources frame j-1's pixels and morphed points, shape = color = (j == 0 ? 0 : 1/(N-j)).
int poppy_ref_chain(int w, int h, const uint8_t* bgr1, const uint8_t* bgr2, const float* gabor2, const float* pts1,
                    const float* pts2, int n, int n_frames, int levels, uint8_t* frames_out /* n_frames*w*h*3 */,
                    float* points_out /* n_frames*n*2 or null */) {
    return guarded([&] {
        poppy::Settings::instance().pyramid_levels = (size_t)levels;
        poppy::Settings::instance().show_gui = false;
        poppy::Settings::instance().number_of_frames = n_frames;
        cv::Mat img1 = wrap_u8c3(bgr1, w, h).clone();
        cv::Mat corrected1 = img1.clone(), corrected2 = wrap_u8c3(bgr2, w, h).clone();
        cv::Mat g2 = wrap_f32(gabor2, w, h, 3), gf1, gf2, morphed;
        std::vector<cv::Point2f> p1 = to_points(pts1, n), p2 = to_points(pts2, n), cur, prev;
        const double N = poppy::Settings::instance().number_of_frames;
        for (int j = 0; j < n_frames; ++j) {
            if (!prev.empty()) p1 = prev;
            cur.clear();
            double linear = j / N, progress;
            if (linear == 0) progress = 0;
            else if (linear == 1) progress = 1;
            else progress = (1.0 / (1.0 - linear)) / N;
            double shape = progress > 1 ? 1 : progress;
            poppy::morph_images(img1, corrected2, corrected1, corrected2, g2, gf1, gf2, morphed, morphed.clone(), cur,
                                p1, p2, shape, shape, linear);
            corrected1 = morphed.clone();
            prev = cur;
            copy_out(morphed, frames_out + (size_t)j * w * h * 3);
            if (points_out) from_points(cur, points_out + (size_t)j * n * 2);
        }
    });
}

int poppy_ref_triangulate(int w, int h, const float* pts, int n, int32_t* tri_idx, int max_tri, int* n_tri) {
    return guarded([&] {
        std::vector<cv::Point2f> p = to_points(pts, n), uniq;
        poppy::clip_points(p, w, h);
        poppy::make_uniq(p, uniq);
        cv::Subdiv2D sd(cv::Rect(0, 0, w, h));
        sd.insert(uniq);
        std::vector<cv::Vec3i> idx;
        poppy::get_triangle_indices(sd, p, idx);
        *n_tri = (int)idx.size();
        for (int i = 0; i < (int)idx.size() && i < max_tri; ++i)
            for (int k = 0; k < 3; ++k) tri_idx[3 * i + k] = idx[i][k];
    });
}

// The path of src/algo.cpp:178-265 issued stage by stage through the reference's own exported helpers.
int poppy_ref_stages(int w, int h, const uint8_t* bgr1, const uint8_t* bgr2, const float* gabor2, const float* pts1,
                     const float* pts2, int n, double shape, double mask, int levels, poppy_ref_stage_dump* d) {
    return guarded([&] {
        using namespace poppy;
        std::vector<cv::Point2f> s1 = to_points(pts1, n), s2 = to_points(pts2, n), morphed, uniq;
        cv::Mat c1 = wrap_u8c3(bgr1, w, h), c2 = wrap_u8c3(bgr2, w, h), g2 = wrap_f32(gabor2, w, h, 3);
        clip_points(s1, w, h);
        clip_points(s2, w, h);
        morph_points(s1, s2, morphed, shape);
        clip_points(morphed, w, h);
        make_uniq(morphed, uniq);
        cv::Subdiv2D sd(cv::Rect(0, 0, w, h));
        sd.insert(uniq);
        std::vector<cv::Vec3i> idx;
        get_triangle_indices(sd, morphed, idx);
        d->n_tri = (int)idx.size();
        if (d->morphed_points) from_points(morphed, d->morphed_points);
        if (d->tri_idx)
            for (int i = 0; i < (int)idx.size() && i < d->max_tri; ++i)
                for (int k = 0; k < 3; ++k) d->tri_idx[3 * i + k] = idx[i][k];
        std::vector<std::vector<cv::Point>> t1, t2, tm;
        make_triangler_points(idx, s1, t1);
        make_triangler_points(idx, s2, t2);
        make_triangler_points(idx, morphed, tm);
        cv::Mat tri_map = cv::Mat::zeros(h, w, CV_32SC1);
        paint_triangles(tri_map, tm);
        copy_out(tri_map, d->tri_map);
        std::vector<cv::Mat> hom, m1, m2;
        solve_homography(t1, t2, hom);
        morph_homography(hom, m1, m2, shape);
        mats_out(hom, d->hom); mats_out(m1, d->m1); mats_out(m2, d->m2);
        cv::Mat mx1, my1, mx2, my2, w1, w2;
        create_map(tri_map, m1, mx1, my1);
        cv::remap(c1, w1, mx1, my1, cv::INTER_LINEAR);
        create_map(tri_map, m2, mx2, my2);
        cv::remap(c2, w2, mx2, my2, cv::INTER_LINEAR);
        copy_out(mx1, d->mapx1); copy_out(my1, d->mapy1); copy_out(mx2, d->mapx2); copy_out(my2, d->mapy2);
        copy_out(w1, d->warped1); copy_out(w2, d->warped2);
        cv::Mat_<cv::Vec3f> l, r;
        w1.convertTo(l, CV_32F, 1.0 / 255.0);
        w2.convertTo(r, CV_32F, 1.0 / 255.0);
        cv::Mat m;
        cv::cvtColor(g2, m, cv::COLOR_BGR2GRAY);
        m = 1.0 - m;
        cv::Mat ones = cv::Mat::ones(m.size(), m.type());
        cv::Mat lbmask = (ones * (1.0 - mask)) - (m * mask);
        lbmask.setTo(0.0, lbmask < 0);
        lbmask.setTo(1.0, lbmask > 1);
        copy_out(lbmask, d->mask);
        LaplacianBlending lb(l, r, lbmask, levels);
        cv::Mat_<cv::Vec3f> blended = lb.blend();
        copy_out(blended, d->lap_blend);
        double amount = sin(mask * M_PI);
        cv::Mat dst = unsharp_mask(blended, 1, 1.0 - amount, 0.3);
        dst.convertTo(dst, CV_8U, 255);
        copy_out(dst, d->dst);
    });
}

// ---- primitives ------------------------------------------------------------------------------------------------
int poppy_ref_fill_triangles(int32_t* img, int w, int h, const int32_t* tri_xy /* T x 6 */, int n_tri) {
    return guarded([&] {
        cv::Mat m(h, w, CV_32SC1, img);
        std::vector<std::vector<cv::Point>> tris(n_tri);
        for (int i = 0; i < n_tri; ++i)
            for (int k = 0; k < 3; ++k) tris[i].push_back(cv::Point(tri_xy[6 * i + 2 * k], tri_xy[6 * i + 2 * k + 1]));
        poppy::paint_triangles(m, tris);
    });
}

int poppy_ref_remap_u8c3(const uint8_t* src, int w, int h, const float* mapx, const float* mapy, int dw, int dh,
                         uint8_t* dst) {
    return guarded([&] {
        cv::Mat out;
        cv::remap(wrap_u8c3(src, w, h), out, wrap_f32(mapx, dw, dh, 1), wrap_f32(mapy, dw, dh, 1), cv::INTER_LINEAR);
        copy_out(out, dst);
    });
}

int poppy_ref_pyr_down(const float* src, int w, int h, int cn, float* dst, int dw, int dh) {
    return guarded([&] {
        cv::Mat out;
        cv::pyrDown(wrap_f32(src, w, h, cn), out, cv::Size(dw, dh));
        copy_out(out, dst);
    });
}

int poppy_ref_pyr_up(const float* src, int w, int h, int cn, float* dst, int dw, int dh) {
    return guarded([&] {
        cv::Mat out;
        cv::pyrUp(wrap_f32(src, w, h, cn), out, cv::Size(dw, dh));
        copy_out(out, dst);
    });
}

int poppy_ref_lap_blend(const float* l, const float* r, const float* mask, int w, int h, int levels, float* out) {
    return guarded([&] {
        cv::Mat_<cv::Vec3f> lm = wrap_f32(l, w, h, 3), rm = wrap_f32(r, w, h, 3);
        cv::Mat_<float> mm = wrap_f32(mask, w, h, 1);
        poppy::LaplacianBlending lb(lm, rm, mm, levels);
        cv::Mat_<cv::Vec3f> b = lb.blend();
        copy_out(b, out);
    });
}

int poppy_ref_unsharp(const float* src, int w, int h, float radius, float amount, float threshold, float* out) {
    return guarded([&] {
        cv::Mat res = poppy::unsharp_mask(wrap_f32(src, w, h, 3), radius, amount, threshold);
        copy_out(res, out);
    });
}

int poppy_ref_gaussian_blur(const float* src, int w, int h, int cn, double sigma, float* out) {
    return guarded([&] {
        cv::Mat res;
        cv::GaussianBlur(wrap_f32(src, w, h, cn), res, cv::Size(0, 0), sigma);
        copy_out(res, out);
    });
}

int poppy_ref_median3(const float* src, int w, int h, int cn, float* out) {
    return guarded([&] {
        cv::Mat res;
        cv::medianBlur(wrap_f32(src, w, h, cn), res, 3);
        copy_out(res, out);
    });
}

// mask basis as the reference builds it (src/algo.cpp:250-258)
int poppy_ref_mask(const float* gabor2, int w, int h, double mask_ratio, float* out) {
    return guarded([&] {
        cv::Mat m;
        cv::cvtColor(wrap_f32(gabor2, w, h, 3), m, cv::COLOR_BGR2GRAY);
        m = 1.0 - m;
        cv::Mat ones = cv::Mat::ones(m.size(), m.type());
        cv::Mat lbmask = (ones * (1.0 - mask_ratio)) - (m * mask_ratio);
        lbmask.setTo(0.0, lbmask < 0);
        lbmask.setTo(1.0, lbmask > 1);
        copy_out(lbmask, out);
    });
}

// input generators of SURVEY.md §8(d): cv::RNG noise + GaussianBlur, used for synthetic configs 4/5
int poppy_ref_synth_image(int w, int h, uint64_t seed, double sigma, uint8_t* out_bgr) {
    return guarded([&] {
        cv::RNG rng(seed);
        cv::Mat img(h, w, CV_8UC3);
        rng.fill(img, cv::RNG::UNIFORM, 0, 256);
        cv::Mat blurred;
        cv::GaussianBlur(img, blurred, cv::Size(0, 0), sigma);
        copy_out(blurred, out_bgr);
    });
}

// input conditioning ahead of the path: poppy::blur_margin (src/util.cpp:574-602) and the 8-bit
// GaussianBlur it is made of (OpenCV's fixed-point path)
int poppy_ref_blur_margin(const uint8_t* src, int w, int h, int union_w, int union_h, uint8_t* out) {
    return guarded([&] {
        cv::Mat s(h, w, CV_8UC3, const_cast<uint8_t*>(src));
        cv::Mat dst;
        poppy::blur_margin(s, cv::Size(union_w, union_h), dst);
        if (dst.cols != union_w || dst.rows != union_h || dst.type() != CV_8UC3) throw std::runtime_error("blur_margin: unexpected result");
        for (int y = 0; y < union_h; ++y) std::memcpy(out + (size_t)y * union_w * 3, dst.ptr(y), (size_t)union_w * 3);
    });
}

// poppy::gabor_filter with its defaults, as src/poppy.hpp:122 calls it (16 angles, 13 x 13, sigma 5, lambda 10, gamma 0.04, psi pi/4)
int poppy_ref_gabor_filter(const float* src, int w, int h, float* out) {
    return guarded([&] {
        cv::Mat dst;
        poppy::gabor_filter(wrap_f32(src, w, h, 3), dst);
        copy_out(dst, out);
    });
}

// cv::getGaborKernel as gabor_filter requests it (CV_32F), row-major ksize x ksize
int poppy_ref_gabor_kernel(int ksize, double sigma, double theta, double lambd, double gamma, double psi, float* out) {
    return guarded([&] {
        cv::Mat k = cv::getGaborKernel(cv::Size(ksize, ksize), sigma, theta, lambd, gamma, psi, CV_32F);
        copy_out(k, out);
    });
}

int poppy_ref_gaussian_blur_u8(const uint8_t* src, int w, int h, int ksize, double sigma, uint8_t* out) {
    return guarded([&] {
        cv::Mat s(h, w, CV_8UC3, const_cast<uint8_t*>(src));
        cv::Mat dst;
        cv::GaussianBlur(s.clone(), dst, cv::Size(ksize, ksize), sigma);
        for (int y = 0; y < h; ++y) std::memcpy(out + (size_t)y * w * 3, dst.ptr(y), (size_t)w * 3);
    });
}

}  // extern "C"
