// link stubs for the symbols host_abi.cpp references but the planner test never calls
#include <cstddef>
#include <cstdint>
#include <functional>
#include <string>
#include <vector>
#include "morph_images.hpp"
extern "C" {
struct poppy_cuda_ctx;
int poppy_cuda_download(poppy_cuda_ctx*, int, int, uint8_t*, size_t, size_t) { return -1; }
int poppy_cuda_get_info(poppy_cuda_ctx*, int*, int*, int*, int*, int*, int*) { return -1; }
int poppy_cuda_get_morphed_points(poppy_cuda_ctx*, int, float*) { return -1; }
const char* poppy_cuda_last_error(const poppy_cuda_ctx*) { return "stub"; }
int poppy_cuda_render(poppy_cuda_ctx*, int, const float*, const double*, const int32_t*, const int32_t*, int) { return -1; }
int poppy_cuda_set_pair(poppy_cuda_ctx*, const uint8_t*, size_t, const uint8_t*, size_t, const float*, size_t) { return -1; }
int poppy_cuda_set_points(poppy_cuda_ctx*, const float*, const float*, int) { return -1; }
int poppy_cuda_sync(poppy_cuda_ctx*) { return -1; }
}
namespace poppy {
Settings* Settings::instance_ = nullptr;
double morph_images(const Image8&, const Image8&, const Image8&, const Image8&, const Image32F&, Image8&, Image8&, Image8&, const Image8&,
                    std::vector<Point2f>&, std::vector<Point2f>, std::vector<Point2f>, double, double, double) { return 0; }
void morph_sequence(const Image8&, const Image8&, const Image32F&, std::vector<Point2f>, std::vector<Point2f>, int,
                    const std::function<void(const Image8&)>&) {}
void release_cached_contexts() {}
}
