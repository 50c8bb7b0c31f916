// TEST INFRASTRUCTURE ONLY — harness around the *unmodified* reference pipeline for BASELINE.json configs 1-3.
//
// Built by oracle/build_ref_full.sh (see there) from the reference translation units where they lie. This file
// replaces exactly the two things of the reference's front end that are absent from this image:
// SDL image loading (-> cv::imread) and the boost command line (-> poppy::init with the CLI's defaults,
// src/poppy.cpp:344-360). Everything else - blur_margin, extractor, matcher, autoalign, face landmarks, gabor_filter,
// the frame loop with its recurrence (src/poppy.hpp:46-248) - is the reference's own code, driven through
// poppy::morph<Sink>() exactly as run() drives it (src/poppy.cpp:239,307,324).
//
// src/algo.cpp is compiled with -Dmorph_images=morph_images_reference, so the call at src/poppy.hpp:215 lands in the
// interposer below, which records the call's inputs/outputs and forwards to the reference body or (poppy_dropin
// only, -DPOPPY_WITH_B200) to integration/algo_b200.cpp compiled against the reference's src/algo.hpp.
//
//   poppy_ref_full dump <config 1|2|3> <images dir> <out dir> [all|some]     (cwd = reference src/: face assets)
//        runs the reference, writes the fixture files read by tests/golden/make_golden_full.py
//   poppy_ref_full raw <config> <images dir> <dump dir>          adds the pre-canvas inputs to a dump (for `full`)
//   poppy_dropin full <dump dir> [frames]
//        the WHOLE pipeline from the raw inputs, all-reference and with blur_margin + gabor_filter + morph_images served by
//        integration/{util,algo}_b200.cpp; compares the union canvases and every frame
//   poppy_dropin replay <dump dir> <ref|b200|both> [frames]
//        re-runs poppy::morph<Sink>() from the dumped union images (no /root/reference needed, configs without
//        face detection) with the chosen morph_images body; `both` runs the chain twice and compares every frame
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <string>
#include <vector>

#include <opencv2/core/utility.hpp>
#include <opencv2/imgcodecs.hpp>

#include "poppy.hpp"

namespace poppy {
double morph_images_reference(const Mat& img1, const Mat& img2, const Mat& corrected1, const Mat& corrected2, const Mat& gabor2,
                              Mat& goodFeatures1, Mat& goodFeatures2, Mat& dst, const Mat& last, vector<Point2f>& morphedPoints,
                              vector<Point2f> srcPoints1, vector<Point2f> srcPoints2, double shapeRatio, double maskRatio,
                              double linear);
#ifdef POPPY_WITH_B200
double morph_images_b200(const Mat& img1, const Mat& img2, const Mat& corrected1, const Mat& corrected2, const Mat& gabor2,
                         Mat& goodFeatures1, Mat& goodFeatures2, Mat& dst, const Mat& last, vector<Point2f>& morphedPoints,
                         vector<Point2f> srcPoints1, vector<Point2f> srcPoints2, double shapeRatio, double maskRatio,
                         double linear);
// poppy_dropin only: src/util.cpp is compiled with -Dblur_margin=blur_margin_reference -Dgabor_filter=gabor_filter_reference
// and integration/util_b200.cpp with the _b200 renames; the interposers below route every call of the pipeline
void blur_margin_reference(const Mat& src, const Size& szUnion, Mat& dst);
void gabor_filter_reference(const Mat& src, Mat& dst, size_t numAngles, int kernel_size, double sig, double lm, double gm, double ps);
void blur_margin_b200(const Mat& src, const Size& szUnion, Mat& dst);
void gabor_filter_b200(const Mat& src, Mat& dst, size_t numAngles, int kernel_size, double sig, double lm, double gm, double ps);
#endif
}  // namespace poppy

namespace {

int g_cond = 0;                     // 0 reference blur_margin / gabor_filter, 1 integration/util_b200.cpp
int g_cond_calls[2] = {0, 0};       // blur_margin / gabor_filter calls served by the B200 bodies

struct Call {                       // one morph_images() call as seen at src/poppy.hpp:215
    double shape, mask, linear;
    std::vector<cv::Point2f> morphed;
    cv::Mat dst;
    uint64_t hash;
};
struct Record {
    cv::Mat corrected1, corrected2, gabor2;         // arguments of the first call
    std::vector<cv::Point2f> pts1, pts2;
    std::vector<Call> calls;
    bool keep_frames = true;
    double seconds_in_morph_images = 0;
} g_rec;
int g_impl = 0;                     // 0 reference body, 1 integration/algo_b200.cpp

// position-weighted byte checksum, the same function as k_checksum (poppy_b200/csrc/device/kernels_unsharp.cu) and
// tests/util.py::frame_checksum: sum over bytes of (b + 1) * (mix(i) | 1) mod 2^64
uint64_t fnv1a(const cv::Mat& m) {
    uint64_t acc = 0, i = 0;
    for (int y = 0; y < m.rows; ++y) {
        const uint8_t* p = m.ptr<uint8_t>(y);
        for (size_t x = 0; x < (size_t)m.cols * m.elemSize(); ++x, ++i) {
            uint64_t k = (i + 0x9E3779B97F4A7C15ull) * 0xBF58476D1CE4E5B9ull;
            k ^= k >> 29;
            acc += ((uint64_t)p[x] + 1ull) * (k | 1ull);
        }
    }
    return acc;
}

void write_file(const std::string& path, const void* data, size_t bytes) {
    std::ofstream f(path, std::ios::binary);
    f.write((const char*)data, (std::streamsize)bytes);
}
void write_mat(const std::string& path, const cv::Mat& m) {
    cv::Mat c = m.isContinuous() ? m : m.clone();
    write_file(path, c.data, c.total() * c.elemSize());
}
cv::Mat read_mat(const std::string& path, int rows, int cols, int type) {
    cv::Mat m(rows, cols, type);
    std::ifstream f(path, std::ios::binary);
    f.read((char*)m.data, (std::streamsize)(m.total() * m.elemSize()));
    if (!f) { fprintf(stderr, "cannot read %s\n", path.c_str()); exit(4); }
    return m;
}

struct Sink {                       // the Twriter of poppy::morph<Twriter>() (src/poppy.hpp:46,219)
    std::vector<cv::Mat> frames;
    bool keep = true;
    std::vector<uint64_t> hashes;
    void write(cv::Mat& m) {
        hashes.push_back(fnv1a(m));
        if (keep) frames.push_back(m.clone());
    }
};

struct Config { const char* a; const char* b; int frames; bool face, autoalign; int canvas_w, canvas_h, scale_to; };
const Config kConfigs[4] = {
    {},
    {"square.png", "circle.png", 60, false, false, 0, 0, 0},                       // BASELINE.json configs[0]
    {"subject01.normal.png", "subject01.happy.png", 60, true, false, 0, 0, 0},     // configs[1]
    {"cat.png", "dog.png", 120, false, true, 1920, 1080, 1080},                    // configs[2] (SURVEY.md 8(d) row 3 reading)
};

void init_settings(const Config& c) {
    // the CLI's defaults (src/poppy.cpp:347-360, src/settings.hpp:14-27); frames / face / autoalign from the config
    poppy::init(false, (size_t)c.frames, 1.0, c.autoalign, false, c.face, false, false, 30, 64, "FFV1", false, 8);
}

// what run() does to each input for phase = -1 (src/poppy.cpp:239,307): centre on the union canvas, blur the margins
cv::Mat to_union(const cv::Mat& img, cv::Size sz) {
    cv::Mat out(sz.height, sz.width, img.type(), cv::Scalar::all(0));
    poppy::blur_margin(img, sz, out);
    return out.clone();
}

}  // namespace

#ifdef POPPY_WITH_B200
namespace poppy {
void blur_margin(const Mat& src, const Size& szUnion, Mat& dst) {
    if (g_cond) { ++g_cond_calls[0]; blur_margin_b200(src, szUnion, dst); }
    else blur_margin_reference(src, szUnion, dst);
}
void gabor_filter(const Mat& src, Mat& dst, size_t numAngles, int kernel_size, double sig, double lm, double gm, double ps) {
    // the morph path's call (src/poppy.hpp:122) uses the defaults; the extractor's 31 x 31 call is not on the path
    const bool defaults = numAngles == 16 && kernel_size == 13 && sig == 5 && lm == 10 && gm == 0.04 && ps == CV_PI / 4;
    if (g_cond && defaults) { ++g_cond_calls[1]; gabor_filter_b200(src, dst, numAngles, kernel_size, sig, lm, gm, ps); }
    else gabor_filter_reference(src, dst, numAngles, kernel_size, sig, lm, gm, ps);
}
}  // namespace poppy
#endif

namespace poppy {
// the interposer: src/poppy.hpp:215 calls this
double morph_images(const Mat& img1, const Mat& img2, const Mat& corrected1, const Mat& corrected2, const Mat& gabor2,
                    Mat& goodFeatures1, Mat& goodFeatures2, Mat& dst, const Mat& last, vector<Point2f>& morphedPoints,
                    vector<Point2f> srcPoints1, vector<Point2f> srcPoints2, double shapeRatio, double maskRatio, double linear) {
    if (g_rec.calls.empty()) {
        g_rec.corrected1 = corrected1.clone(); g_rec.corrected2 = corrected2.clone(); g_rec.gabor2 = gabor2.clone();
        g_rec.pts1 = srcPoints1; g_rec.pts2 = srcPoints2;
    }
    const auto t0 = std::chrono::steady_clock::now();
    double r;
#ifdef POPPY_WITH_B200
    if (g_impl == 1)
        r = morph_images_b200(img1, img2, corrected1, corrected2, gabor2, goodFeatures1, goodFeatures2, dst, last, morphedPoints,
                              srcPoints1, srcPoints2, shapeRatio, maskRatio, linear);
    else
#endif
        r = morph_images_reference(img1, img2, corrected1, corrected2, gabor2, goodFeatures1, goodFeatures2, dst, last,
                                   morphedPoints, srcPoints1, srcPoints2, shapeRatio, maskRatio, linear);
    g_rec.seconds_in_morph_images += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    Call c;
    c.shape = shapeRatio; c.mask = maskRatio; c.linear = linear; c.morphed = morphedPoints;
    c.hash = fnv1a(dst);
    if (g_rec.keep_frames) c.dst = dst.clone();
    g_rec.calls.push_back(std::move(c));
    return r;
}
}  // namespace poppy

namespace {

int cmd_dump(int argc, char** argv) {
    if (argc < 5) { fprintf(stderr, "usage: dump <config> <images dir> <out dir> [all|some]\n"); return 2; }
    const int cfg = atoi(argv[2]);
    if (cfg < 1 || cfg > 3) return 2;
    const Config& C = kConfigs[cfg];
    const std::string img_dir = argv[3], out = argv[4];
    const bool all = argc < 6 || std::string(argv[5]) == "all";
    init_settings(C);
    cv::Mat a = cv::imread(img_dir + "/" + C.a, cv::IMREAD_COLOR), b = cv::imread(img_dir + "/" + C.b, cv::IMREAD_COLOR);
    if (a.empty() || b.empty()) { fprintf(stderr, "cannot read the input images\n"); return 3; }
    if (C.scale_to) {   // "upscaled to 1920x1080": aspect-preserving cv::resize, then the union canvas
        cv::Mat t;
        cv::resize(a, t, cv::Size(C.scale_to, C.scale_to * a.rows / a.cols), 0, 0, cv::INTER_LINEAR); a = t.clone();
        cv::resize(b, t, cv::Size(C.scale_to, C.scale_to * b.rows / b.cols), 0, 0, cv::INTER_LINEAR); b = t.clone();
    }
    cv::Size sz(std::max(a.cols, b.cols), std::max(a.rows, b.rows));
    if (C.canvas_w) sz = cv::Size(C.canvas_w, C.canvas_h);
    cv::Mat u1 = to_union(a, sz), u2 = to_union(b, sz), c1, c2;
    Sink sink;
    sink.keep = false;
    g_rec.keep_frames = true;
    const auto t0 = std::chrono::steady_clock::now();
    poppy::morph(u1, u2, c1, c2, -1.0, false, sink);
    const double total_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    const int N = (int)g_rec.calls.size(), n = (int)g_rec.pts1.size();
    if (N == 0) { fprintf(stderr, "the reference rendered no frame (no matches?)\n"); return 5; }
    write_mat(out + "/img1.u8", u1); write_mat(out + "/img2.u8", u2);
    write_mat(out + "/corrected1.u8", g_rec.corrected1); write_mat(out + "/corrected2.u8", g_rec.corrected2);
    write_mat(out + "/gabor2.f32", g_rec.gabor2);
    write_file(out + "/pts1.f32", g_rec.pts1.data(), (size_t)n * 8);
    write_file(out + "/pts2.f32", g_rec.pts2.data(), (size_t)n * 8);
    std::vector<double> ratios;
    std::vector<float> morphed;
    std::vector<uint64_t> hashes;
    for (const Call& c : g_rec.calls) {
        ratios.push_back(c.shape); ratios.push_back(c.mask); ratios.push_back(c.linear);
        for (const auto& p : c.morphed) { morphed.push_back(p.x); morphed.push_back(p.y); }
        hashes.push_back(c.hash);
    }
    write_file(out + "/ratios.f64", ratios.data(), ratios.size() * 8);
    write_file(out + "/morphed.f32", morphed.data(), morphed.size() * 4);
    write_file(out + "/hashes.u64", hashes.data(), hashes.size() * 8);
    std::string kept;
    for (int j = 0; j < N; ++j) {
        const bool keep = all || j < 2 || j == N / 2 || j >= N - 2 || j % 16 == 0;
        if (!keep) continue;
        char name[64];
        snprintf(name, sizeof name, "/frame_%04d.u8", j);
        write_mat(out + name, g_rec.calls[j].dst);
        kept += (kept.empty() ? "" : " ") + std::to_string(j);
    }
    std::ofstream meta(out + "/meta.txt");
    meta << "config " << cfg << "\nimage_a " << C.a << "\nimage_b " << C.b << "\nwidth " << sz.width << "\nheight " << sz.height
         << "\nframes " << N << "\npoints " << n << "\nlevels " << poppy::Settings::instance().pyramid_levels << "\nface " << C.face
         << "\nautoalign " << C.autoalign << "\nkept_frames " << kept << "\ntotal_seconds " << total_s << "\nmorph_images_seconds "
         << g_rec.seconds_in_morph_images << "\nthreads " << cv::getNumThreads() << "\ncpu_features " << cv::getCPUFeaturesLine() << "\n";
    fprintf(stderr, "config %d: %dx%d, %d points, %d frames, %.2f s total, %.2f s in morph_images\n", cfg, sz.width, sz.height, n, N,
            total_s, g_rec.seconds_in_morph_images);
    return 0;
}

struct Meta { int w = 0, h = 0, frames = 0, face = 0, autoalign = 0; };
Meta read_meta(const std::string& dir) {
    Meta m;
    std::ifstream f(dir + "/meta.txt");
    std::string k;
    while (f >> k) {
        std::string rest;
        std::getline(f, rest);
        const int v = atoi(rest.c_str());
        if (k == "width") m.w = v; else if (k == "height") m.h = v; else if (k == "frames") m.frames = v;
        else if (k == "face") m.face = v; else if (k == "autoalign") m.autoalign = v;
    }
    return m;
}

// one run of the reference's frame loop from the dumped union images with the chosen morph_images body
double run_chain(const cv::Mat& u1, const cv::Mat& u2, const Meta& M, int frames, int impl, Sink& sink, double* in_morph) {
    Config C{nullptr, nullptr, frames, M.face != 0, M.autoalign != 0, 0, 0, 0};
    init_settings(C);
    g_impl = impl;
    g_rec = Record();
    g_rec.keep_frames = false;
    cv::Mat c1, c2;
    const auto t0 = std::chrono::steady_clock::now();
    poppy::morph(u1, u2, c1, c2, -1.0, false, sink);
    *in_morph = g_rec.seconds_in_morph_images;
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

int cmd_replay(int argc, char** argv) {
    if (argc < 4) { fprintf(stderr, "usage: replay <dump dir> <ref|b200|both> [frames]\n"); return 2; }
    const std::string dir = argv[2], which = argv[3];
    const Meta M = read_meta(dir);
    if (M.w <= 0 || M.face) { fprintf(stderr, "bad dump (or a face config: its assets live in the reference tree)\n"); return 3; }
    const int frames = argc > 4 ? atoi(argv[4]) : M.frames;
    cv::Mat u1 = read_mat(dir + "/img1.u8", M.h, M.w, CV_8UC3), u2 = read_mat(dir + "/img2.u8", M.h, M.w, CV_8UC3);
#ifndef POPPY_WITH_B200
    if (which != "ref") { fprintf(stderr, "this binary has no B200 body (build poppy_dropin)\n"); return 3; }
#endif
    Sink ref_sink, gpu_sink;
    double ref_s = 0, gpu_s = 0, ref_mi = 0, gpu_mi = 0;
    if (which == "ref" || which == "both") ref_s = run_chain(u1, u2, M, frames, 0, ref_sink, &ref_mi);
    if (which == "b200" || which == "both") gpu_s = run_chain(u1, u2, M, frames, 1, gpu_sink, &gpu_mi);
    long long differing = 0, total = 0;
    int max_abs = 0, first_bad = -1;
    double min_psnr = 1e9, min_within1 = 1.0;
    if (which == "both") {
        if (ref_sink.frames.size() != gpu_sink.frames.size()) { printf("{\"error\": \"frame counts differ\"}\n"); return 1; }
        for (size_t j = 0; j < ref_sink.frames.size(); ++j) {
            const cv::Mat &a = ref_sink.frames[j], &b = gpu_sink.frames[j];
            long long bad = 0, within1 = 0;
            double se = 0;
            for (int y = 0; y < a.rows; ++y) {
                const uint8_t *pa = a.ptr<uint8_t>(y), *pb = b.ptr<uint8_t>(y);
                for (int i = 0; i < a.cols * 3; ++i) {
                    const int d = abs((int)pa[i] - (int)pb[i]);
                    bad += d != 0; within1 += d <= 1; se += (double)d * d;
                    max_abs = std::max(max_abs, d);
                }
            }
            const long long cnt = (long long)a.rows * a.cols * 3;
            differing += bad; total += cnt;
            min_within1 = std::min(min_within1, (double)within1 / cnt);
            if (se > 0) min_psnr = std::min(min_psnr, 10.0 * log10(255.0 * 255.0 * cnt / se));
            if (bad && first_bad < 0) first_bad = (int)j;
        }
    }
    const Sink& any = which == "b200" ? gpu_sink : ref_sink;
    uint64_t chain_hash = 1469598103934665603ull;
    for (uint64_t h : any.hashes) { chain_hash ^= h; chain_hash *= 1099511628211ull; }
    printf("{\"frames\": %zu, \"width\": %d, \"height\": %d, \"mode\": \"%s\", \"differing_bytes\": %lld, \"bytes\": %lld, "
           "\"max_abs\": %d, \"min_fraction_within_1\": %.6f, \"min_psnr_db\": %s, \"first_differing_frame\": %d, "
           "\"reference_s\": %.3f, \"reference_morph_images_s\": %.3f, \"b200_s\": %.3f, \"b200_morph_images_s\": %.3f, "
           "\"chain_hash\": \"%016llx\"}\n",
           any.hashes.size(), M.w, M.h, which.c_str(), differing, total, max_abs, min_within1,
           min_psnr > 1e8 ? "null" : std::to_string(min_psnr).c_str(), first_bad, ref_s, ref_mi, gpu_s, gpu_mi,
           (unsigned long long)chain_hash);
    return differing == 0 ? 0 : 1;
}

// the inputs of a config before the union canvas (what run() reads from disk, resized for config 3)
bool load_raw(int cfg, const std::string& img_dir, cv::Mat& a, cv::Mat& b, cv::Size& sz) {
    const Config& C = kConfigs[cfg];
    a = cv::imread(img_dir + "/" + C.a, cv::IMREAD_COLOR);
    b = cv::imread(img_dir + "/" + C.b, cv::IMREAD_COLOR);
    if (a.empty() || b.empty()) return false;
    if (C.scale_to) {
        cv::Mat t;
        cv::resize(a, t, cv::Size(C.scale_to, C.scale_to * a.rows / a.cols), 0, 0, cv::INTER_LINEAR); a = t.clone();
        cv::resize(b, t, cv::Size(C.scale_to, C.scale_to * b.rows / b.cols), 0, 0, cv::INTER_LINEAR); b = t.clone();
    }
    sz = cv::Size(std::max(a.cols, b.cols), std::max(a.rows, b.rows));
    if (C.canvas_w) sz = cv::Size(C.canvas_w, C.canvas_h);
    return true;
}

// adds the pre-canvas inputs to an existing dump (raw1.u8, raw2.u8, raw.txt), for `full`
int cmd_raw(int argc, char** argv) {
    if (argc < 5) { fprintf(stderr, "usage: raw <config> <images dir> <dump dir>\n"); return 2; }
    const int cfg = atoi(argv[2]);
    if (cfg < 1 || cfg > 3) return 2;
    cv::Mat a, b;
    cv::Size sz;
    if (!load_raw(cfg, argv[3], a, b, sz)) { fprintf(stderr, "cannot read the input images\n"); return 3; }
    const std::string out = argv[4];
    write_mat(out + "/raw1.u8", a); write_mat(out + "/raw2.u8", b);
    std::ofstream f(out + "/raw.txt");
    f << a.cols << " " << a.rows << " " << b.cols << " " << b.rows << " " << sz.width << " " << sz.height << "\n";
    return 0;
}

// The reference's WHOLE pipeline from the raw inputs - blur_margin, extractor, matcher, gabor_filter, frame loop - run twice:
// all-reference, and with blur_margin + gabor_filter (default-parameter call) + morph_images served by the B200 library.
int cmd_full(int argc, char** argv) {
    if (argc < 3) { fprintf(stderr, "usage: full <dump dir> [frames]\n"); return 2; }
#ifndef POPPY_WITH_B200
    fprintf(stderr, "this binary has no B200 bodies (build poppy_dropin)\n");
    return 3;
#else
    const std::string dir = argv[2];
    const Meta M = read_meta(dir);
    int aw = 0, ah = 0, bw = 0, bh = 0, uw = 0, uh = 0;
    { std::ifstream f(dir + "/raw.txt"); f >> aw >> ah >> bw >> bh >> uw >> uh; }
    if (M.w <= 0 || M.face || aw <= 0) { fprintf(stderr, "bad dump (run `raw` first; face configs need the reference tree)\n"); return 3; }
    const int frames = argc > 3 ? atoi(argv[3]) : M.frames;
    cv::Mat a = read_mat(dir + "/raw1.u8", ah, aw, CV_8UC3), b = read_mat(dir + "/raw2.u8", bh, bw, CV_8UC3);
    Sink sinks[2];
    double secs[2] = {0, 0}, in_morph[2] = {0, 0};
    long long canvas_diff = 0;
    cv::Mat canvases[2][2];
    for (int impl = 0; impl < 2; ++impl) {
        g_cond = impl;
        const auto t0 = std::chrono::steady_clock::now();
        cv::Mat u1 = to_union(a, cv::Size(uw, uh)), u2 = to_union(b, cv::Size(uw, uh));
        canvases[impl][0] = u1; canvases[impl][1] = u2;
        secs[impl] = run_chain(u1, u2, M, frames, impl, sinks[impl], &in_morph[impl]);
        secs[impl] = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    }
    g_cond = 0;
    for (int k = 0; k < 2; ++k) canvas_diff += cv::norm(canvases[0][k], canvases[1][k], cv::NORM_L1) != 0;
    long long differing = 0, total = 0;
    int max_abs = 0, first_bad = -1;
    if (sinks[0].frames.size() != sinks[1].frames.size()) { printf("{\"error\": \"frame counts differ\"}\n"); return 1; }
    for (size_t j = 0; j < sinks[0].frames.size(); ++j) {
        const cv::Mat &x = sinks[0].frames[j], &y = sinks[1].frames[j];
        long long bad = 0;
        for (int r = 0; r < x.rows; ++r) {
            const uint8_t *px = x.ptr<uint8_t>(r), *py = y.ptr<uint8_t>(r);
            for (int i = 0; i < x.cols * 3; ++i) {
                const int d = abs((int)px[i] - (int)py[i]);
                bad += d != 0;
                max_abs = std::max(max_abs, d);
            }
        }
        differing += bad; total += (long long)x.rows * x.cols * 3;
        if (bad && first_bad < 0) first_bad = (int)j;
    }
    printf("{\"frames\": %zu, \"width\": %d, \"height\": %d, \"mode\": \"full pipeline: reference vs blur_margin + gabor_filter + morph_images on B200\", "
           "\"differing_bytes\": %lld, \"bytes\": %lld, \"max_abs\": %d, \"first_differing_frame\": %d, \"canvases_differing\": %lld, "
           "\"b200_blur_margin_calls\": %d, \"b200_gabor_filter_calls\": %d, \"reference_s\": %.3f, \"b200_s\": %.3f}\n",
           sinks[0].frames.size(), uw, uh, differing, total, max_abs, first_bad, canvas_diff, g_cond_calls[0], g_cond_calls[1], secs[0], secs[1]);
    return differing == 0 && canvas_diff == 0 ? 0 : 1;
#endif
}

}  // namespace

int main(int argc, char** argv) {
    if (argc >= 2 && std::string(argv[1]) == "raw") return cmd_raw(argc, argv);
    if (argc >= 2 && std::string(argv[1]) == "full") return cmd_full(argc, argv);
    if (argc >= 2 && std::string(argv[1]) == "dump") return cmd_dump(argc, argv);
    if (argc >= 2 && std::string(argv[1]) == "replay") return cmd_replay(argc, argv);
    fprintf(stderr, "usage: %s dump|raw|replay|full ...\n", argv[0]);
    return 2;
}
