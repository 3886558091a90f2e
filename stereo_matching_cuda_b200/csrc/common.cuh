// common.cuh -- context, workspace arena and launch helpers shared by the .cu files.
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "../../include/stereo_b200.h"

struct sb200_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    int sm_count = 148;
    size_t smem_optin = 0;
    // workspace arena: one device allocation, bump-allocated per entry point, grown on demand
    char* ws = nullptr;
    size_t ws_cap = 0;
    size_t ws_off = 0;
    uint64_t launches = 0;
    char err[512] = {0};
    // optional timing of the pipeline's phases
    int timing = 0;
    cudaEvent_t ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t ev_x[2] = {nullptr, nullptr};  // around the NCCL halo exchange of sb200_pipeline_strips_nccl
    bool ev_valid = false;
    bool fused_attr_set = false;
    bool mma_attr_set = false;
    bool rgb_mma_attr_set = false;
    int gray_kernel = 1;  // gray-guide fused kernel: 0 = warp-shuffle box sums (fused_cvf.cu), 1 = tensor-core box sums (fused_mma.cu)
    cudaEvent_t ev_stream = nullptr;  // orders a new stream after the work queued on the previous one (set_stream)
    // overlapped host-pointer batch entry (sb200_pipeline_batch): copy streams, per-slot events, device staging
    cudaStream_t s_in = nullptr, s_out = nullptr;
    cudaEvent_t ev_in[2] = {nullptr, nullptr}, ev_comp[2] = {nullptr, nullptr}, ev_out[2] = {nullptr, nullptr};
    bool batch_ready = false;
    char* stage = nullptr;
    size_t stage_cap = 0;
    // optional: the tensor-core gray kernel also stores the filtered cost volume q of view v here (D-major planes of
    // rows_out x w floats, costVolume.cu:178's layout); set by the entry points that want it for ONE call, then cleared
    float* qvol[2] = {nullptr, nullptr};
    int rgb_kernel = 4;  // RGB-guide fused kernel: 4 = tensor-core (fused_mma_rgb.cu), 3 = three-stage shuffle kernel
                         // (fused_cvf_rgb3.cu), 2 = its two-stage predecessor (fused_cvf_rgb.cu)
};

extern char g_sb200_global_err[512];

// Every extern "C" entry point runs on the context's device and leaves the caller's current device as it found it
// (a process may hold several contexts, or have torch's current device elsewhere).
struct DevGuard {
    int prev = -1;
    bool switched = false;
    explicit DevGuard(const sb200_ctx* c) {
        if (c && cudaGetDevice(&prev) == cudaSuccess && prev != c->device) switched = (cudaSetDevice(c->device) == cudaSuccess);
    }
    explicit DevGuard(int device) {
        if (cudaGetDevice(&prev) == cudaSuccess && prev != device) switched = (cudaSetDevice(device) == cudaSuccess);
    }
    ~DevGuard() {
        if (switched) cudaSetDevice(prev);
    }
    DevGuard(const DevGuard&) = delete;
    DevGuard& operator=(const DevGuard&) = delete;
};

inline int sb_fail(sb200_ctx* ctx, int code, const char* fmt, ...) {
    char* dst = ctx ? ctx->err : g_sb200_global_err;
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(dst, 512, fmt, ap);
    va_end(ap);
    return code;
}

#define SB_CUDA(ctx, call)                                                                          \
    do {                                                                                            \
        cudaError_t e__ = (call);                                                                   \
        if (e__ != cudaSuccess)                                                                     \
            return sb_fail(ctx, SB200_ERR_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #call,        \
                           cudaGetErrorString(e__));                                                \
    } while (0)

#define SB_TRY(expr)                    \
    do {                                \
        int rc__ = (expr);              \
        if (rc__ != SB200_OK) return rc__; \
    } while (0)

// launch + count + immediate launch-error check (asynchronous execution errors surface at
// the next synchronising call, which also goes through SB_CUDA)
#define SB_LAUNCH(ctx, kernel, grid, block, smem, ...)                                              \
    do {                                                                                            \
        kernel<<<grid, block, smem, (ctx)->stream>>>(__VA_ARGS__);                                   \
        (ctx)->launches++;                                                                          \
        cudaError_t e__ = cudaGetLastError();                                                       \
        if (e__ != cudaSuccess)                                                                     \
            return sb_fail(ctx, SB200_ERR_CUDA, "%s:%d launch %s -> %s", __FILE__, __LINE__,        \
                           #kernel, cudaGetErrorString(e__));                                       \
    } while (0)

inline size_t sb_align(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

// Make sure the arena holds `bytes`; resets the bump pointer.  Growing synchronises the
// stream (only happens while shapes are still growing).
int sb_ws_reserve(sb200_ctx* ctx, size_t bytes);
template <typename T>
inline T* sb_ws_alloc(sb200_ctx* ctx, size_t count) {
    size_t off = sb_align(ctx->ws_off);
    size_t end = off + count * sizeof(T);
    if (end > ctx->ws_cap) return nullptr;
    ctx->ws_off = end;
    return reinterpret_cast<T*>(ctx->ws + off);
}

inline int sb_div_up(long long a, long long b) { return (int)((a + b - 1) / b); }

// ---- stage kernels (stage_kernels.cu) ------------------------------------------------
int sbk_rgb_to_gray(sb200_ctx* ctx, const sb200_params* p, const uint8_t* rgb, int n, int ch, uint8_t* gray);
int sbk_x_derivative(sb200_ctx* ctx, const uint8_t* img, float* grad, int w, int h);
int sbk_cost_volume(sb200_ctx* ctx, const sb200_params* p, const uint8_t* i1, const uint8_t* i2, const float* g1,
                    const float* g2, float* cost, int w, int h, int size_d, int dmin);
int sbk_integral(sb200_ctx* ctx, const float* in, float* tmp, float* out, int w, int h);
int sbk_box_from_sat(sb200_ctx* ctx, const float* sat, float* mean, int w, int h, int r);
int sbk_box_sliding(sb200_ctx* ctx, const float* in, double* tmp, float* mean, int w, int h, int r);
int sbk_u8_to_float(sb200_ctx* ctx, const uint8_t* in, float* out, size_t n);
int sbk_float_to_u8(sb200_ctx* ctx, const float* in, uint8_t* out, size_t n);
int sbk_mul(sb200_ctx* ctx, const float* a, const float* b, float* out, size_t n);
int sbk_sub(sb200_ctx* ctx, const float* a, const float* b, float* out, size_t n);
int sbk_ak_bk(sb200_ctx* ctx, const float* mean, const float* var, const float* mIp, const float* mp, float* a,
              float* b, size_t n, double eps);
int sbk_q(sb200_ctx* ctx, const float* im, const float* a, const float* b, float* q, size_t n);
int sbk_disp_select(sb200_ctx* ctx, const float* q, float* best, float* dmap, size_t n, int label);
int sbk_detect_occlusion(sb200_ctx* ctx, float* dL, const float* dR, int dOcc, int d_lr, int w, int h);
int sbk_fill_occlusion(sb200_ctx* ctx, float* disp, int w, int h, float vMin);
// weighted median of the pixels the L/R check marked; ws: (radius+1)^2 spatial weights, wc: 256 colour weights (device)
int sbk_weighted_median(sb200_ctx* ctx, const uint8_t* gray, const float* occ, const float* filled, float* out, int w, int h,
                        int dmin, int size_d, int radius, const unsigned* ws, const unsigned* wc);
int sbk_lr_check_fill(sb200_ctx* ctx, const float* dL, const float* dR, int w, int h, int dOcc, int d_lr, float vMin,
                      float* occ, float* filled);
int sbk_fill_f32(sb200_ctx* ctx, float* dst, float v, size_t n);
int sbk_labels_i16(sb200_ctx* ctx, const float* const src[4], int16_t* const dst[4], size_t n);  // NULL entries are skipped
int sbk_subpixel(sb200_ctx* ctx, const float* vol, const float* disp, const float* occ, const float* filled, float* out,
                 size_t n, int dmin, int size_d);
int sbk_write_mat(sb200_ctx* ctx, const float* mat, uint8_t* out, size_t n, float* scratch);  // scratch: 2*ceil(n/1024)+2 words
int sbk_fl_to_ch2(sb200_ctx* ctx, const float* image, uint8_t* result, int mn, int mx, size_t len);
int sbk_rgb_split(sb200_ctx* ctx, const uint8_t* rgb, int ch, float* r, float* g, float* b, size_t n);
int sbk_rgb_inverse(sb200_ctx* ctx, float* const mu[3], float* const cov[6], size_t n, double eps);
int sbk_rgb_ab(sb200_ctx* ctx, float* const mu[3], float* const M[6], float* const mIp[3], const float* mp,
               float* const a[3], float* b, size_t n);
int sbk_rgb_q(sb200_ctx* ctx, const float* mb, float* const ma[3], float* const I[3], float* q, size_t n);

// ---- fused path (fused_cvf.cu) --------------------------------------------------------
struct SbFusedGeom {
    int w, h;         // image width, rows held in the input buffers
    int y_out0;       // first output row (index into the held rows)
    int rows_out;     // number of output rows
    int y_global0;    // frame row index of held row 0 (for clipped window areas)
    int frame_h;      // rows of the whole frame
};
// 1 if the fused kernels cover these parameters (radius 9, exact integer cost lattice)
int sbf_fused_supported(const sb200_params* p);
size_t sbf_workspace_bytes(const sb200_ctx* ctx, int w, int h_held, int rows_out, int dabs, int size_d, int n_views);
// gray images (held rows) -> per-view best cost + labels (+ mean debug image) for BOTH views
int sbf_pair_disparity(sb200_ctx* ctx, const sb200_params* p, const uint8_t* gray_l, const uint8_t* gray_r,
                       const SbFusedGeom& g, float* bestL, float* dispL, float* bestR, float* dispR,
                       uint8_t* meanL, uint8_t* meanR);
// single view (guide, other, dmin)
int sbf_view_disparity(sb200_ctx* ctx, const sb200_params* p, const uint8_t* guide, const uint8_t* other,
                       const SbFusedGeom& g, int dmin, int size_d, float* best, float* disp, uint8_t* mean);
// tensor-core variant (fused_mma.cu): same contract as run_fused of fused_cvf.cu
int sbf_mma_supported(const sb200_params* p);
size_t sbf_mma_workspace_bytes(const sb200_ctx* ctx, int w, int h_held, int rows_out, int dabs, int size_d, int n_views);
int sbf_run_fused_mma(sb200_ctx* ctx, const sb200_params* p, const uint8_t* const gray[2], const SbFusedGeom& g,
                      const int dmin[2], int size_d, int n_views, float* const best[2], float* const disp[2],
                      uint8_t* const mean[2]);
// RGB guide, fused (fused_cvf_rgb.cu): colour images (interleaved, `channels` bytes per pixel) + their gray versions
size_t sbf_rgb_workspace_bytes(const sb200_ctx* ctx, int w, int h_held, int rows_out, int dabs, int size_d);
int sbf_pair_disparity_rgb(sb200_ctx* ctx, const sb200_params* p, const uint8_t* rgb_l, const uint8_t* rgb_r, int channels,
                           const uint8_t* gray_l, const uint8_t* gray_r, const SbFusedGeom& g, float* bestL, float* dispL,
                           float* bestR, float* dispR);
// RGB guide, three-stage kernel (fused_cvf_rgb3.cu): same contract
// tensor-core RGB-guide kernel (fused_mma_rgb.cu); same contract as the rgb3 pair below
int sbf_rgb_mma_supported(const sb200_params* p);
size_t sbf_rgb_mma_workspace_bytes(const sb200_ctx* ctx, int w, int h_held, int rows_out, int dabs, int size_d);
int sbf_pair_disparity_rgb_mma(sb200_ctx* ctx, const sb200_params* p, const uint8_t* rgb_l, const uint8_t* rgb_r, int channels,
                               const uint8_t* gray_l, const uint8_t* gray_r, const SbFusedGeom& g, float* bestL, float* dispL,
                               float* bestR, float* dispR);
size_t sbf_rgb3_workspace_bytes(const sb200_ctx* ctx, int w, int h_held, int rows_out, int dabs, int size_d);
int sbf_pair_disparity_rgb3(sb200_ctx* ctx, const sb200_params* p, const uint8_t* rgb_l, const uint8_t* rgb_r, int channels,
                            const uint8_t* gray_l, const uint8_t* gray_r, const SbFusedGeom& g, float* bestL, float* dispL,
                            float* bestR, float* dispR);
