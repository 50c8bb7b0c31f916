// integration/util_b200.cpp — the drop-in for the input conditioning that runs once per image ahead of the morph path:
// poppy::blur_margin (reference src/util.cpp:574-602, called at src/poppy.cpp:239,296,307) and poppy::gabor_filter with the
// arguments src/poppy.hpp:122 passes (the defaults of src/util.hpp:95). Same signatures over cv::Mat, compiled against the
// reference's unmodified src/util.hpp and linked with libpoppy_cuda.so in place of those two functions
// (oracle/build_ref_full.sh -> oracle/_ref/poppy_dropin; `poppy_dropin full` runs the reference's whole pipeline with them).
//
// blur_margin is bit-exact; gabor_filter agrees with the reference to float rounding (include/poppy_cuda.h). The GPU entry
// point implements the default parameter set only: the extractor's own call (src/extractor.cpp:65-66, 31 x 31 kernels) is not
// part of the morph path, so any other parameter set is an error here and the caller keeps the reference body for it.
#include "util.hpp"

#include <stdexcept>

#include <poppy_cuda.h>

namespace poppy {

void blur_margin(const Mat& src, const Size& szUnion, Mat& dst) {
    CV_Assert(src.type() == CV_8UC3);
    Mat s = src.isContinuous() ? src : src.clone();
    Mat out(szUnion.height, szUnion.width, CV_8UC3);
    if (poppy_cuda_blur_margin(/*device*/ 0, s.data, s.step, s.cols, s.rows, szUnion.width, szUnion.height, out.data, out.step) != 0)
        CV_Error(cv::Error::StsError, poppy_cuda_blur_margin_last_error());
    dst = out;
}

void gabor_filter(const Mat& src, Mat& dst, size_t numAngles, int kernel_size, double sig, double lm, double gm, double ps) {
    CV_Assert(src.type() == CV_32FC3);
    if (numAngles != 16 || kernel_size != 13 || sig != 5 || lm != 10 || gm != 0.04 || ps != CV_PI / 4)
        CV_Error(cv::Error::StsBadArg, "gabor_filter: the GPU entry point implements the default parameter set");
    Mat s = src.isContinuous() ? src : src.clone();
    Mat out(src.rows, src.cols, CV_32FC3);
    if (poppy_cuda_gabor_filter(/*device*/ 0, s.ptr<float>(), s.step, s.cols, s.rows, out.ptr<float>(), out.step) != 0)
        CV_Error(cv::Error::StsError, poppy_cuda_blur_margin_last_error());
    dst = out;
}

}  // namespace poppy
