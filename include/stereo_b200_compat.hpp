// stereo_b200_compat.hpp -- the reference's own C++ stage signatures on top of the C ABI.
//
// A maintainer of hamza1030/stereo_matching_cuda who wants the B200 path keeps main.cu as it
// is, drops the reference's .cu stage files from the build, includes this header instead of the
// six .cuh headers and links libstereo_b200.so.  Every function below has the signature declared
// in the reference header named beside it and forwards to the matching sb200_* entry point on a
// process-wide context (device 0, like main.cu:44-48).  Differences that remain visible:
//   * errors throw std::runtime_error instead of printing and calling exit(0)
//     (SystemIncludes.h:46-52);
//   * the disparity range, radius, eps and thresholds are no longer macros: set them once with
//     stereo_b200::params() before the first call (defaults = SystemIncludes.h:6-24).
#pragma once
#include <cstdlib>
#include <stdexcept>
#include <string>

#include "stereo_b200.h"

namespace stereo_b200 {
inline sb200_params& params() {
    static sb200_params p = [] {
        sb200_params q;
        sb200_default_params(&q);
        q.box_mode = SB200_BOX_SAT;  // the reference's float32 integral-image arithmetic, bit for bit
        return q;
    }();
    return p;
}
inline sb200_ctx* context() {
    static sb200_ctx* ctx = [] {
        sb200_ctx* c = nullptr;
        if (sb200_ctx_create(0, &c) != SB200_OK) throw std::runtime_error(sb200_last_error(nullptr));
        return c;
    }();
    return ctx;
}
inline void check(int rc) {
    if (rc != SB200_OK) throw std::runtime_error(std::string("stereo_b200: ") + sb200_last_error(context()));
}
}  // namespace stereo_b200

// rgb_to_grayscale.cuh:7 -- returns a malloc'd buffer the caller frees (main.cu:189-190)
inline unsigned char* rgb_to_grayscale(unsigned char* h_rgb, const int n, int channels, bool /*host_gpu_compare*/) {
    unsigned char* gray = static_cast<unsigned char*>(std::malloc(n));
    stereo_b200::check(sb200_rgb_to_grayscale(stereo_b200::context(), &stereo_b200::params(), h_rgb, n, channels, gray));
    return gray;
}
// costVolume.cuh:7
inline void compute_cost(unsigned char* i1, unsigned char* i2, float* cost, int w1, int w2, int h1, int h2, int dmin,
                         bool /*host_gpu_compare*/) {
    stereo_b200::check(sb200_compute_cost(stereo_b200::context(), &stereo_b200::params(), i1, i2, cost, w1, w2, h1, h2, dmin));
}
// integral.cuh:3
inline void integral(float* image, float* integral_out, int width, int height) {
    stereo_b200::check(sb200_integral(stereo_b200::context(), image, integral_out, width, height));
}
// filter.cuh:12 (live semantics: guide mean + variance, guidedFilter.cu:58-123)
inline void filter(unsigned char* image, int width, int height, unsigned char* mean, float* var, bool /*cuda*/) {
    stereo_b200::check(sb200_filter(stereo_b200::context(), &stereo_b200::params(), image, width, height, mean, var));
}
// guidedFilter.cuh:7
inline void compute_guided_filter(unsigned char* i, float* cost, float* filter_cost, float* disp_map, unsigned char* mean,
                                  const int w, const int h, const int size_d, int dmin, bool /*host_gpu_compare*/) {
    stereo_b200::check(sb200_compute_guided_filter(stereo_b200::context(), &stereo_b200::params(), i, cost, filter_cost,
                                                   disp_map, mean, w, h, size_d, dmin));
}
// guidedFilter.cuh:8 is a __global__ kernel in the reference; as a host call on host arrays:
inline void dispSelect(float* q, float* filter_cost, float* dmap, const int n, int label) {
    stereo_b200::check(sb200_winner_take_all(stereo_b200::context(), q, filter_cost, dmap, n, label));
}
// occlusion.cuh:8
inline void detect_occlusion(float* disparityLeft, float* disparityRight, const int dOcclusion, unsigned char* dmapl,
                             unsigned char* dmapr, const int w, const int h) {
    stereo_b200::check(sb200_detect_occlusion(stereo_b200::context(), &stereo_b200::params(), disparityLeft, disparityRight,
                                              dOcclusion, dmapl, dmapr, w, h));
}
// occlusion.cuh:14
inline void fill_occlusion(float* disparity, const int w, const int h, const float vMin) {
    stereo_b200::check(sb200_fill_occlusion(stereo_b200::context(), disparity, w, h, vMin));
}
