// integration/algo_b200.cpp — the drop-in: poppy::morph_images() with the reference's own signature (cv::Mat in and
// out), compiled against the reference's unmodified src/algo.hpp and linked with libpoppy_cuda.so. It replaces the body
// at reference src/algo.cpp:178-273; nothing else of the reference changes (src/poppy.hpp:215 keeps calling it).
//
// This is the file a reference maintainer adds. It is compiled and run by
// oracle/build_ref_full.sh -> oracle/_ref/poppy_dropin, where the reference's own frame loop poppy::morph<Sink>()
// (src/poppy.hpp:46-248) is driven once with the stock body and once with this one (tests/test_gpu_dropin.py).
//
// Host stages (clip_points / make_uniq / cv::Subdiv2D-order Delaunay / vertex-index lookup, src/algo.cpp:184-213) run
// inside poppy_morph_images() on the CPU; everything from morph_points() to the 8-bit frame runs on the GPU. There is
// no CPU rendering path: without a CUDA device the call raises cv::Exception.
#include "algo.hpp"
#include "settings.hpp"

#include <algorithm>

#include <poppy_host.h>   // include/poppy_host.h (+ poppy_cuda.h) of this repository

namespace poppy {

namespace {
// One device context per (frame size, pyramid levels); Settings::pyramid_levels is read on every call like the
// reference does (src/algo.cpp:261), so changing it between calls rebuilds the context.
poppy_cuda_ctx* context_for(int cols, int rows, int n_points) {
    static poppy_cuda_ctx* ctx = nullptr;
    static int w = 0, h = 0, lv = 0, cap = 0;
    const int levels = (int)Settings::instance().pyramid_levels;
    if (!ctx || w != cols || h != rows || lv != levels || cap < n_points) {
        if (ctx) poppy_cuda_destroy(ctx);
        ctx = nullptr;
        cap = std::max(n_points, 64);
        if (poppy_cuda_create(&ctx, /*device*/ 0, cols, rows, levels, cap, 2 * cap + 16, /*frames in flight*/ 1) != 0)
            CV_Error(cv::Error::GpuApiCallError, poppy_cuda_last_error(nullptr));
        w = cols; h = rows; lv = levels;
    }
    return ctx;
}
}  // namespace

double morph_images(const Mat& img1, const Mat& img2, const Mat& corrected1, const Mat& corrected2, const Mat& gabor2,
                    Mat& goodFeatures1, Mat& goodFeatures2, Mat& dst, const Mat& last, vector<Point2f>& morphedPoints,
                    vector<Point2f> srcPoints1, vector<Point2f> srcPoints2, double shapeRatio, double maskRatio,
                    double linear) {
    (void)img2; (void)goodFeatures1; (void)goodFeatures2; (void)last; (void)linear;      // unused by the reference body too
    CV_Assert(srcPoints1.size() == srcPoints2.size());
    CV_Assert(corrected1.type() == CV_8UC3 && corrected2.type() == CV_8UC3 && gabor2.type() == CV_32FC3);
    CV_Assert(corrected1.size() == img1.size() && corrected2.size() == img1.size() && gabor2.size() == img1.size());
    const int n = (int)srcPoints1.size();
    poppy_cuda_ctx* c = context_for(img1.cols, img1.rows, n);
    // the reference assigns a fresh Mat to dst (src/algo.cpp:264-265): never write through a caller's alias
    Mat out(img1.rows, img1.cols, CV_8UC3);
    morphedPoints.resize(n);
    static const float none[2] = {0.f, 0.f};
    const float* p1 = n ? &srcPoints1.data()->x : none;        // cv::Point2f is two packed floats
    const float* p2 = n ? &srcPoints2.data()->x : none;
    float* mp = n ? &morphedPoints.data()->x : nullptr;
    const int rc = poppy_morph_images(c, corrected1.data, corrected1.step, corrected2.data, corrected2.step,
                                      gabor2.ptr<float>(), gabor2.step, p1, p2, n, shapeRatio, maskRatio, out.data,
                                      out.step, mp);
    // where cv::Subdiv2D would throw (a clipped point with x == cols, subdivision2d.cpp:287) so does this
    if (rc != 0) CV_Error(cv::Error::StsError, poppy_host_last_error());
    dst = out;
    return 0;
}

}  // namespace poppy
