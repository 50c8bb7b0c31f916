#include "morph_images.hpp"

#include <cmath>
#include <cstring>
#include <memory>
#include <mutex>
#include <thread>

#include "../../../include/poppy_host.h"

namespace poppy {

Settings* Settings::instance_ = nullptr;

namespace {

struct CachedContext {
    poppy_cuda_ctx* ctx = nullptr;
    int w = 0, h = 0, levels = 0, max_points = 0, max_tri = 0, max_frames = 0, device = 0;
    ~CachedContext() { if (ctx) poppy_cuda_destroy(ctx); }
};

std::mutex g_mu;
std::unique_ptr<CachedContext> g_cached;

void cu(poppy_cuda_ctx* c, int rc) {
    if (rc != 0) throw MorphError(std::string("poppy_cuda: ") + poppy_cuda_last_error(c));
}

poppy_cuda_ctx* context_for(int w, int h, int levels, int n_points, int n_tri, int n_frames) {
    const int device = Settings::instance().cuda_device;
    CachedContext* k = g_cached.get();
    if (k && k->w == w && k->h == h && k->levels == levels && k->device == device && k->max_points >= n_points &&
        k->max_tri >= n_tri && k->max_frames >= n_frames)
        return k->ctx;
    g_cached.reset(new CachedContext());
    k = g_cached.get();
    k->w = w; k->h = h; k->levels = levels; k->device = device;
    k->max_points = std::max(n_points, 64);
    k->max_tri = std::max(2 * n_points + 16, n_tri);
    k->max_frames = n_frames;
    int rc = poppy_cuda_create(&k->ctx, device, w, h, levels, k->max_points, k->max_tri, k->max_frames);
    if (rc != 0) {
        std::string msg = std::string("poppy_cuda_create: ") + poppy_cuda_last_error(nullptr);
        g_cached.reset();
        throw MorphError(msg);
    }
    // both entry points render one frame per kernel batch (a single frame, or a chain = a recurrence): size the scratch for
    // that instead of the 32-frame batches of direct-mode sequences
    poppy_cuda_set_chunk_frames(k->ctx, 1);
    return k->ctx;
}

}  // namespace

void release_cached_contexts() {
    std::lock_guard<std::mutex> lock(g_mu);
    g_cached.reset();
}

double morph_images(const Image8& img1, const Image8& /*img2*/, const Image8& corrected1, const Image8& corrected2,
                    const Image32F& gabor2, Image8& /*goodFeatures1*/, Image8& /*goodFeatures2*/, Image8& dst,
                    const Image8& /*last*/, std::vector<Point2f>& morphedPoints, std::vector<Point2f> srcPoints1,
                    std::vector<Point2f> srcPoints2, double shapeRatio, double maskRatio, double /*linear*/) {
    const int w = img1.cols, h = img1.rows;
    if (corrected1.cols != w || corrected1.rows != h || corrected2.cols != w || corrected2.rows != h ||
        gabor2.cols != w || gabor2.rows != h)
        throw MorphError("morph_images: image sizes differ");
    if (srcPoints1.size() != srcPoints2.size()) throw MorphError("morph_images: point sets differ in size");
    const int n = (int)srcPoints1.size();

    // host stages, reference src/algo.cpp:184-213 (subDiv1/subDiv2 feed only the GUI analysis and are skipped)
    clip_points(srcPoints1, w, h);
    clip_points(srcPoints2, w, h);
    morphedPoints.resize(n);
    static const float none[2] = {0.f, 0.f};
    const float* p1 = n ? &srcPoints1.data()->x : none;
    const float* p2 = n ? &srcPoints2.data()->x : none;
    float* mp = n ? &morphedPoints.data()->x : nullptr;
    if (n) poppy_host_morph_points(p1, p2, n, shapeRatio, w, h, mp);
    std::lock_guard<std::mutex> lock(g_mu);
    // sized for the most triangles n points can give, so that the pair can travel to the device while this thread triangulates
    poppy_cuda_ctx* c = context_for(w, h, (int)Settings::instance().pyramid_levels, n, 2 * n + 16, 1);
    int rc_up = 0;
    std::thread upload([&] {
        rc_up = poppy_cuda_set_pair(c, corrected1.data, corrected1.step, corrected2.data, corrected2.step, gabor2.data, gabor2.step);
        if (rc_up == 0) rc_up = poppy_cuda_set_points(c, p1, p2, n);
    });
    std::vector<int32_t> tri;
    std::string err;
    const bool tri_ok = triangulate_points_next(morphedPoints, w, h, tri, &err);
    upload.join();
    if (!tri_ok) throw MorphError("morph_images: " + err);
    cu(c, rc_up);
    const int n_tri = (int)tri.size() / 3;
    const float s = (float)shapeRatio;
    const int32_t offs[2] = {0, n_tri};
    cu(c, poppy_cuda_render(c, 1, &s, &maskRatio, tri.data(), offs, 0));
    if (dst.cols != w || dst.rows != h || dst.data == nullptr) dst.create(w, h);
    cu(c, poppy_cuda_download(c, 0, 1, dst.data, dst.step, dst.step * h));
    // the device recomputes morph_points(); hand back its copy (bit-identical to the host's, see tests)
    if (n) cu(c, poppy_cuda_get_morphed_points(c, 0, mp));
    cu(c, poppy_cuda_sync(c));
    return 0;
}

void morph_sequence(const Image8& corrected1, const Image8& corrected2, const Image32F& gabor2,
                    std::vector<Point2f> srcPoints1, std::vector<Point2f> srcPoints2, int number_of_frames,
                    const std::function<void(const Image8&)>& write) {
    const int w = corrected1.cols, h = corrected1.rows, n = (int)srcPoints1.size(), N = number_of_frames;
    if (N < 1) return;
    if (srcPoints1.size() != srcPoints2.size()) throw MorphError("morph_sequence: point sets differ in size");
    std::vector<float> ratio(N);
    std::vector<double> mask(N);
    for (int j = 0; j < N; ++j) { mask[j] = poppy_host_chain_ratio(j, N); ratio[j] = (float)mask[j]; }
    poppy_host_plan* plan = nullptr;
    static const float none[2] = {0.f, 0.f};
    const float* p1 = n ? &srcPoints1.data()->x : none;
    const float* p2 = n ? &srcPoints2.data()->x : none;
    if (poppy_host_plan_create(&plan, p1, p2, n, w, h, N, ratio.data(), 1, 0) != 0)
        throw MorphError(std::string("morph_sequence: ") + poppy_host_last_error());
    struct PlanGuard { poppy_host_plan* p; ~PlanGuard() { poppy_host_plan_destroy(p); } } guard{plan};
    const int32_t *tri = nullptr, *offs = nullptr;
    int max_tri = 0;
    poppy_host_plan_triangles(plan, &tri, &offs, &max_tri);

    std::lock_guard<std::mutex> lock(g_mu);
    poppy_cuda_ctx* c = context_for(w, h, (int)Settings::instance().pyramid_levels, n, max_tri, N);
    cu(c, poppy_cuda_set_pair(c, corrected1.data, corrected1.step, corrected2.data, corrected2.step, gabor2.data, gabor2.step));
    cu(c, poppy_cuda_set_points(c, p1, p2, n));
    // The chain is rendered slice by slice (frame j-1 stays in HBM as frame j's source across the calls) and every finished
    // slice is handed to the writer ring: slice k+1 renders while slice k downloads and slice k-1 is being written.
    struct Sink { const std::function<void(const Image8&)>* write; } sink{&write};
    auto deliver = [](void* u, int, uint8_t* bgr, int fw, int fh, size_t step) {
        Image8 f;
        f.data = bgr; f.cols = fw; f.rows = fh; f.step = step;
        (*static_cast<Sink*>(u)->write)(f);
    };
    poppy_host_writer* wr = nullptr;
    if (poppy_host_writer_create(&wr, c, std::min(N, 8), deliver, nullptr, 0, &sink) != 0)
        throw MorphError("morph_sequence: cannot create the writer ring");
    struct WriterGuard { poppy_host_writer* w; ~WriterGuard() { poppy_host_writer_destroy_cuda(w); } } wguard{wr};
    const int S = 4;
    std::vector<int32_t> so;
    for (int a = 0; a < N; a += S) {
        const int cnt = std::min(S, N - a);
        so.assign(cnt + 1, 0);
        for (int i = 0; i <= cnt; ++i) so[i] = offs[a + i] - offs[a];
        cu(c, poppy_cuda_render_range(c, a, cnt, ratio.data() + a, mask.data() + a, tri + (size_t)offs[a] * 3, so.data(), 1));
        if (poppy_host_writer_submit(wr, a, cnt, a) != 0) throw MorphError(std::string("morph_sequence: ") + poppy_cuda_last_error(c));
    }
    if (poppy_host_writer_flush(wr) != 0) throw MorphError("morph_sequence: a frame download failed");
}

}  // namespace poppy
