// Writer hand-off (include/poppy_host.h, poppy_host_writer_*): the reference's frame loop passes every finished frame to
// Twriter::write(Mat&) (reference src/poppy.hpp:219: cv::VideoWriter on native builds, src/poppy.cpp:249; the GIF writer
// on WASM, src/poppy.cpp:57-84) and waits for it. Here the frames leave the GPU through a ring of page-locked buffers:
// the caller's thread enqueues the download of rendered ring slots and goes on planning / rendering, a delivery thread
// waits for each copy and calls `write` strictly in frame order (video encoders are order dependent), and an optional
// worker pool runs a per-frame `convert` step (pixel-format conversion, compression ...) ahead of delivery.
#include <atomic>
#include <condition_variable>
#include <cstring>
#include <deque>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../../include/poppy_host.h"

struct poppy_host_writer {
    poppy_host_writer_io io{};
    int width = 0, height = 0, ring = 0, workers = 0;
    size_t frame_bytes = 0;
    std::vector<uint8_t*> buffers;              // page-locked frame buffers
    struct Item { int frame_index; int buffer; uint64_t ticket; bool converted; };
    std::mutex mu;
    std::condition_variable cv_free, cv_work, cv_done;
    std::deque<int> free_buffers;
    std::deque<Item> pending;                   // submitted, in frame order
    size_t next_convert = 0;                    // first entry of `pending` no worker has taken yet
    uint64_t submitted = 0, written = 0;
    bool stop = false;
    int error = 0;
    std::thread deliverer;
    std::vector<std::thread> pool;

    void deliver_loop() {
        for (;;) {
            Item it;
            {
                std::unique_lock<std::mutex> lk(mu);
                cv_work.wait(lk, [&] { return stop || (!pending.empty() && (workers == 0 || pending.front().converted)); });
                if (pending.empty()) { if (stop) return; continue; }
                if (workers > 0 && !pending.front().converted) { if (stop) return; continue; }
                it = pending.front();
            }
            if (workers == 0) {
                if (io.wait && io.wait(io.user, it.ticket) != 0) { std::lock_guard<std::mutex> lk(mu); error = POPPY_CUDA_ERR_CUDA; }
            }
            if (io.write) io.write(io.user, it.frame_index, buffers[it.buffer], width, height, (size_t)width * 3);
            {
                std::lock_guard<std::mutex> lk(mu);
                pending.pop_front();
                if (next_convert > 0) --next_convert;
                free_buffers.push_back(it.buffer);
                ++written;
            }
            cv_free.notify_all();
            cv_done.notify_all();
        }
    }

    void convert_loop() {
        for (;;) {
            size_t idx;
            Item it;
            {
                std::unique_lock<std::mutex> lk(mu);
                cv_work.wait(lk, [&] { return stop || next_convert < pending.size(); });
                if (next_convert >= pending.size()) { if (stop) return; continue; }
                idx = next_convert++;
                it = pending[idx];
            }
            if (io.wait && io.wait(io.user, it.ticket) != 0) { std::lock_guard<std::mutex> lk(mu); error = POPPY_CUDA_ERR_CUDA; }
            if (io.convert) io.convert(io.user, it.frame_index, buffers[it.buffer], width, height, (size_t)width * 3);
            {
                std::lock_guard<std::mutex> lk(mu);
                // entries only leave at the front, and only after they are converted: find ours by frame index
                for (auto& p : pending)
                    if (p.frame_index == it.frame_index && p.buffer == it.buffer) { p.converted = true; break; }
            }
            cv_work.notify_all();
        }
    }
};

extern "C" {

int poppy_host_writer_create_io(poppy_host_writer** out, const poppy_host_writer_io* io, int width, int height, int ring_frames,
                                int workers) {
    if (!out) return POPPY_CUDA_ERR_INVALID;
    *out = nullptr;
    if (!io || !io->download || !io->write || !io->alloc || !io->release || width <= 0 || height <= 0 || ring_frames < 1 || workers < 0)
        return POPPY_CUDA_ERR_INVALID;
    poppy_host_writer* w = new poppy_host_writer();
    w->io = *io;
    w->width = width; w->height = height; w->ring = ring_frames; w->workers = io->convert ? workers : 0;
    w->frame_bytes = (size_t)width * height * 3;
    for (int i = 0; i < ring_frames; ++i) {
        void* p = nullptr;
        if (io->alloc(io->user, w->frame_bytes, &p) != 0 || !p) {
            for (uint8_t* b : w->buffers) io->release(io->user, b);
            delete w;
            return POPPY_CUDA_ERR_CUDA;
        }
        w->buffers.push_back((uint8_t*)p);
        w->free_buffers.push_back(i);
    }
    w->deliverer = std::thread([w] { w->deliver_loop(); });
    for (int i = 0; i < w->workers; ++i) w->pool.emplace_back([w] { w->convert_loop(); });
    *out = w;
    return 0;
}

int poppy_host_writer_submit(poppy_host_writer* w, int first_slot, int count, int first_frame_index) {
    if (!w || count < 0) return POPPY_CUDA_ERR_INVALID;
    for (int i = 0; i < count; ++i) {
        int buf;
        {
            std::unique_lock<std::mutex> lk(w->mu);
            w->cv_free.wait(lk, [&] { return !w->free_buffers.empty(); });       // back-pressure: the ring is full
            buf = w->free_buffers.front();
            w->free_buffers.pop_front();
        }
        uint64_t ticket = 0;
        const int rc = w->io.download(w->io.user, first_slot + i, w->buffers[buf], (size_t)w->width * 3, &ticket);
        if (rc != 0) {
            std::lock_guard<std::mutex> lk(w->mu);
            w->free_buffers.push_back(buf);
            return rc;
        }
        {
            std::lock_guard<std::mutex> lk(w->mu);
            w->pending.push_back({first_frame_index + i, buf, ticket, false});
            ++w->submitted;
        }
        w->cv_work.notify_all();
    }
    return 0;
}

int poppy_host_writer_flush(poppy_host_writer* w) {
    if (!w) return POPPY_CUDA_ERR_INVALID;
    std::unique_lock<std::mutex> lk(w->mu);
    w->cv_done.wait(lk, [&] { return w->written == w->submitted; });
    return w->error;
}

void poppy_host_writer_destroy(poppy_host_writer* w) {
    if (!w) return;
    poppy_host_writer_flush(w);
    {
        std::lock_guard<std::mutex> lk(w->mu);
        w->stop = true;
    }
    w->cv_work.notify_all();
    if (w->deliverer.joinable()) w->deliverer.join();
    for (auto& t : w->pool) t.join();
    for (uint8_t* b : w->buffers) w->io.release(w->io.user, b);
    delete w;
}

// ---- the CUDA transport: ring slots of a poppy_cuda context ---------------------------------------------------------------
namespace {
struct CudaTransport { poppy_cuda_ctx* ctx; poppy_write_fn write; poppy_write_fn convert; void* user; int w, h; };
int ct_download(void* u, int slot, uint8_t* dst, size_t step, uint64_t* ticket) {
    auto* t = (CudaTransport*)u;
    return poppy_cuda_download_async(t->ctx, slot, 1, dst, step, step * (size_t)t->h, ticket);
}
int ct_wait(void* u, uint64_t ticket) { return poppy_cuda_download_wait(((CudaTransport*)u)->ctx, ticket); }
int ct_alloc(void*, size_t bytes, void** out) { return poppy_cuda_alloc_pinned(bytes, out); }
void ct_release(void*, void* p) { poppy_cuda_free_pinned(p); }
void ct_write(void* u, int idx, uint8_t* bgr, int w, int h, size_t step) { auto* t = (CudaTransport*)u; t->write(t->user, idx, bgr, w, h, step); }
void ct_convert(void* u, int idx, uint8_t* bgr, int w, int h, size_t step) { auto* t = (CudaTransport*)u; t->convert(t->user, idx, bgr, w, h, step); }
}  // namespace

int poppy_host_writer_create(poppy_host_writer** out, poppy_cuda_ctx* ctx, int ring_frames, poppy_write_fn write,
                             poppy_write_fn convert, int workers, void* user) {
    if (!out || !ctx || !write) return POPPY_CUDA_ERR_INVALID;
    int w = 0, h = 0;
    if (int rc = poppy_cuda_get_info(ctx, &w, &h, nullptr, nullptr, nullptr, nullptr)) return rc;
    // at most 64 download tickets may be outstanding (poppy_cuda.h)
    if (ring_frames > 48) ring_frames = 48;
    auto* t = new CudaTransport{ctx, write, convert, user, w, h};
    poppy_host_writer_io io{};
    io.user = t; io.download = ct_download; io.wait = ct_wait; io.alloc = ct_alloc; io.release = ct_release; io.write = ct_write;
    io.convert = convert ? ct_convert : nullptr;
    const int rc = poppy_host_writer_create_io(out, &io, w, h, ring_frames, workers);
    if (rc != 0) { delete t; return rc; }
    (*out)->io.owned_transport = t;
    return 0;
}

void poppy_host_writer_destroy_cuda(poppy_host_writer* w) {
    if (!w) return;
    auto* t = (CudaTransport*)w->io.owned_transport;
    poppy_host_writer_destroy(w);
    delete t;
}

}  // extern "C"
