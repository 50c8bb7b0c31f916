/* poppy_host.h — C ABI of the host stages that stay on the CPU (north-star): point hygiene, the Delaunay
 * topology of the morphed points, the vertex-index lookup and the frame schedule. Together with poppy_cuda.h this
 * is everything poppy::morph_images() does (reference src/algo.cpp:178-273); the C++ shim in
 * poppy_b200/csrc/host/morph_images.hpp keeps the reference's C++ signature on top of it.
 *
 * Replaced reference interfaces:
 *   clip_points / make_uniq / check_points          src/util.cpp:453-471,541-548
 *   morph_points                                    src/algo.cpp:50-58   (host copy; the device recomputes it bit-identically)
 *   cv::Subdiv2D::insert / getTriangleList          OCV imgproc/src/subdivision2d.cpp:412-490,756-785
 *   get_triangle_indices                            src/algo.cpp:60-81
 *   frame schedule of morph<Twriter>()              src/poppy.hpp:177-210
 */
#ifndef POPPY_HOST_H_
#define POPPY_HOST_H_

#include <stddef.h>
#include <stdint.h>

#include "poppy_cuda.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct poppy_host_plan poppy_host_plan;

/* clip both sets, lerp with (float)shape_ratio, clip the result: the morphedPoints of morph_images. */
int poppy_host_morph_points(const float* pts1_xy, const float* pts2_xy, int n, double shape_ratio, int width,
                            int height, float* out_xy);

/* clip -> make_uniq -> Delaunay -> triangle vertex indices into pts (first exact-equal occurrence), in
 * cv::Subdiv2D::getTriangleList order. *n_tri receives the triangle count even when it exceeds `cap`
 * (then POPPY_CUDA_ERR_CAPACITY is returned). POPPY_CUDA_ERR_INVALID where cv::Subdiv2D would throw
 * (a point with x >= width or y >= height after clipping). */
int poppy_host_triangulate(const float* pts_xy, int n, int width, int height, int32_t* tri_idx, int cap, int* n_tri);

/* shape/mask ratio of frame j of an n_frames segment, reference src/poppy.hpp:181-210 with phase < 0:
 * 0 for j == 0, 1/(N-j) afterwards, capped at 1. */
double poppy_host_chain_ratio(int j, int n_frames);

/* Topology plan of a whole sequence: morphed points and triangle lists of every frame, computed on `threads`
 * host threads (0 = hardware concurrency).
 * chain == 0: frame f lerps (pts1, pts2) with shape_ratio[f]  (direct mode)
 * chain == 1: frame f lerps (morphed points of frame f-1, pts2) with shape_ratio[f]  (src/poppy.hpp:178-179) */
int poppy_host_plan_create(poppy_host_plan** out, const float* pts1_xy, const float* pts2_xy, int n, int width,
                           int height, int n_frames, const float* shape_ratio, int chain, int threads);
int poppy_host_plan_triangles(const poppy_host_plan* plan, const int32_t** tri_idx, const int32_t** tri_offsets,
                              int* max_triangles);
int poppy_host_plan_points(const poppy_host_plan* plan, int frame, const float** xy);
void poppy_host_plan_destroy(poppy_host_plan* plan);

/* C flavour of poppy::morph_images(): one frame, host buffers in and out. The context must have been created for
 * the same frame size; pyramid_levels is the context's (Settings::pyramid_levels at creation).
 * dst_bgr (rows `dst_step` bytes apart) and morphed_xy (n x 2) receive the outputs. Always returns 0.0 in the
 * reference; here 0 / negative status. */
int poppy_morph_images(poppy_cuda_ctx* ctx, const uint8_t* corrected1, size_t step1, const uint8_t* corrected2,
                       size_t step2, const float* gabor2, size_t gstep, const float* src_points1_xy,
                       const float* src_points2_xy, int n, double shape_ratio, double mask_ratio, uint8_t* dst_bgr,
                       size_t dst_step, float* morphed_xy);

const char* poppy_host_last_error(void);

#ifdef __cplusplus
}
#endif
#endif /* POPPY_HOST_H_ */
