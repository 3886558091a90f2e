// api.cu -- the C ABI of include/stereo_b200.h: context, workspace, stage entry points and the
// fused pipeline.  Host C++ driving CUDA; no torch types, no CPU fallback.
#include <vector>
#include <dlfcn.h>

#include <new>

#include "common.cuh"

char g_sb200_global_err[512] = {0};

extern "C" {

const char* sb200_version(void) { return "stereo_b200 0.1 (sm_100a)"; }

void sb200_default_params(sb200_params* p) {
    if (!p) return;
    p->dmin = -15;
    p->dmax = 0;
    p->radius = 9;
    p->d_lr = 0;
    p->eps = 6.5025;
    p->alpha = (float)0.9;
    p->th_color = 7.0f;
    p->th_grad = 2.0f;
    p->r_w = 0.299;
    p->g_w = 0.587;
    p->b_w = 0.0721;
    p->guide_mode = SB200_GUIDE_GRAY;
    p->box_mode = SB200_BOX_SLIDING;
}

int sb200_ctx_create(int device, sb200_ctx** out) {
    if (!out) return sb_fail(nullptr, SB200_ERR_INVALID, "ctx_create: out is NULL");
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return sb_fail(nullptr, SB200_ERR_CUDA, "no CUDA device: %s (this library has no CPU fallback)",
                       e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    if (device < 0 || device >= ndev) return sb_fail(nullptr, SB200_ERR_INVALID, "device %d of %d", device, ndev);
    DevGuard dev_guard__(device);
    sb200_ctx* ctx = new (std::nothrow) sb200_ctx();
    if (!ctx) return sb_fail(nullptr, SB200_ERR_NOMEM, "ctx alloc");
    ctx->device = device;
    int v = 0;
    cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, device);
    ctx->sm_count = v > 0 ? v : 148;
    cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
    ctx->smem_optin = (size_t)v;
    int major = 0;
    cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device);
    if (major != 10) {
        delete ctx;
        return sb_fail(nullptr, SB200_ERR_UNSUPPORTED, "device %d is sm_%d0; this library is built for sm_100a only", device,
                       major);
    }
    bool ok = true;
    for (int i = 0; i < 5; i++) ok = ok && (cudaEventCreate(&ctx->ev[i]) == cudaSuccess);
    for (int i = 0; i < 2; i++) ok = ok && (cudaEventCreate(&ctx->ev_x[i]) == cudaSuccess);
    ctx->ev_valid = ok;
    if (cudaEventCreateWithFlags(&ctx->ev_stream, cudaEventDisableTiming) != cudaSuccess) ctx->ev_stream = nullptr;
    // A/B switch for the RGB-guide kernel: "4" = tensor-core (fused_mma_rgb.cu, the default), "3" = three-stage shuffle
    // kernel (fused_cvf_rgb3.cu), "2" = two-stage (fused_cvf_rgb.cu)
    if (const char* e = getenv("SB200_RGB_KERNEL")) ctx->rgb_kernel = (e[0] == '4') ? 4 : (e[0] == '3') ? 3 : (e[0] == '2') ? 2 : ctx->rgb_kernel;
    // A/B switch for the gray-guide kernel: "shfl" = warp-shuffle box sums (fused_cvf.cu), "mma" = tensor-core box sums
    if (const char* e = getenv("SB200_GRAY_KERNEL")) ctx->gray_kernel = (e[0] == 's') ? 0 : (e[0] == 'm') ? 1 : ctx->gray_kernel;
    *out = ctx;
    return SB200_OK;
}

void sb200_ctx_destroy(sb200_ctx* ctx) {
    DevGuard dev_guard__(ctx);
    if (!ctx) return;
    cudaStreamSynchronize(ctx->stream);
    if (ctx->ev_stream) cudaEventDestroy(ctx->ev_stream);
    if (ctx->ws) cudaFree(ctx->ws);
    if (ctx->stage) cudaFree(ctx->stage);
    if (ctx->batch_ready) {
        cudaStreamSynchronize(ctx->s_in);
        cudaStreamSynchronize(ctx->s_out);
        cudaStreamDestroy(ctx->s_in);
        cudaStreamDestroy(ctx->s_out);
        for (int i = 0; i < 2; i++) {
            cudaEventDestroy(ctx->ev_in[i]);
            cudaEventDestroy(ctx->ev_comp[i]);
            cudaEventDestroy(ctx->ev_out[i]);
        }
    }
    if (ctx->ev_valid) {
        for (int i = 0; i < 5; i++) cudaEventDestroy(ctx->ev[i]);
        for (int i = 0; i < 2; i++) cudaEventDestroy(ctx->ev_x[i]);
    }
    delete ctx;
}

int sb200_ctx_set_stream(sb200_ctx* ctx, void* stream) {
    DevGuard dev_guard__(ctx);
    if (!ctx) return SB200_ERR_INVALID;
    cudaStream_t next = (cudaStream_t)stream;
    if (next != ctx->stream) {
        // the arena is shared by all calls and rewound by each: work queued on the old stream must be done with it
        // before work on the new stream overwrites it.  One in-flight stream per context.
        if (ctx->ev_stream) {
            SB_CUDA(ctx, cudaEventRecord(ctx->ev_stream, ctx->stream));
            SB_CUDA(ctx, cudaStreamWaitEvent(next, ctx->ev_stream, 0));
        } else {
            SB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        }
    }
    ctx->stream = next;
    return SB200_OK;
}

int sb200_ctx_synchronize(sb200_ctx* ctx) {
    DevGuard dev_guard__(ctx);
    if (!ctx) return SB200_ERR_INVALID;
    SB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SB200_OK;
}

const char* sb200_last_error(const sb200_ctx* ctx) { return ctx ? ctx->err : g_sb200_global_err; }
uint64_t sb200_launch_count(const sb200_ctx* ctx) { return ctx ? ctx->launches : 0; }
int sb200_ctx_rgb_kernel(const sb200_ctx* ctx) { return ctx ? ctx->rgb_kernel : 0; }
int sb200_ctx_gray_kernel(const sb200_ctx* ctx) { return ctx ? ctx->gray_kernel : -1; }
int sb200_ctx_set_gray_kernel(sb200_ctx* ctx, int which) {
    if (!ctx || (which != 0 && which != 1)) return SB200_ERR_INVALID;
    ctx->gray_kernel = which;
    return SB200_OK;
}

int sb200_ctx_enable_timing(sb200_ctx* ctx, int on) {
    DevGuard dev_guard__(ctx);
    if (!ctx) return SB200_ERR_INVALID;
    ctx->timing = on;
    return SB200_OK;
}

int sb200_last_timing(sb200_ctx* ctx, float* ms_prep, float* ms_fused, float* ms_merge, float* ms_occl) {
    DevGuard dev_guard__(ctx);
    if (!ctx || !ctx->ev_valid) return SB200_ERR_INVALID;
    SB_CUDA(ctx, cudaEventSynchronize(ctx->ev[4]));
    float* outs[4] = {ms_prep, ms_fused, ms_merge, ms_occl};
    for (int i = 0; i < 4; i++)
        if (outs[i]) SB_CUDA(ctx, cudaEventElapsedTime(outs[i], ctx->ev[i], ctx->ev[i + 1]));
    return SB200_OK;
}

int sb200_last_exchange_ms(sb200_ctx* ctx, float* ms) {
    DevGuard dev_guard__(ctx);
    if (!ctx || !ctx->ev_valid || !ms) return SB200_ERR_INVALID;
    SB_CUDA(ctx, cudaEventSynchronize(ctx->ev_x[1]));
    SB_CUDA(ctx, cudaEventElapsedTime(ms, ctx->ev_x[0], ctx->ev_x[1]));
    return SB200_OK;
}

}  // extern "C"

int sb_ws_reserve(sb200_ctx* ctx, size_t bytes) {  // (called under the entry point's DevGuard)
    bytes = sb_align(bytes, 1 << 20);
    if (bytes > ctx->ws_cap) {
        SB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        if (ctx->ws) SB_CUDA(ctx, cudaFree(ctx->ws));
        ctx->ws = nullptr;
        ctx->ws_cap = 0;
        cudaError_t e = cudaMalloc(&ctx->ws, bytes);
        if (e != cudaSuccess) return sb_fail(ctx, SB200_ERR_NOMEM, "workspace of %zu bytes: %s", bytes, cudaGetErrorString(e));
        ctx->ws_cap = bytes;
    }
    ctx->ws_off = 0;
    return SB200_OK;
}

namespace {

#define REQUIRE(ctx, cond, msg)                                              \
    do {                                                                     \
        if (!(ctx)) return SB200_ERR_INVALID;                                \
        if (!(cond)) return sb_fail(ctx, SB200_ERR_INVALID, "%s: %s", __func__, msg); \
    } while (0)

template <typename T>
int ws_get(sb200_ctx* ctx, T** out, size_t count) {
    *out = sb_ws_alloc<T>(ctx, count);
    if (!*out) return sb_fail(ctx, SB200_ERR_NOMEM, "workspace arena exhausted (internal sizing bug)");
    return SB200_OK;
}

int h2d(sb200_ctx* ctx, void* dst, const void* src, size_t bytes) {
    SB_CUDA(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
    return SB200_OK;
}
int d2h(sb200_ctx* ctx, void* dst, const void* src, size_t bytes) {
    SB_CUDA(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    return SB200_OK;
}

// box mean in the requested mode; scratch: tmp_f (n floats), sat (n floats), tmp_d (n doubles)
struct BoxScratch {
    float* tmp_f;
    float* sat;
    double* tmp_d;
};
int box_mean(sb200_ctx* ctx, const sb200_params* p, const float* in, float* out, int w, int h, const BoxScratch& s) {
    if (p->box_mode == SB200_BOX_SAT) {
        SB_TRY(sbk_integral(ctx, in, s.tmp_f, s.sat, w, h));
        return sbk_box_from_sat(ctx, s.sat, out, w, h, p->radius);
    }
    return sbk_box_sliding(ctx, in, s.tmp_d, out, w, h, p->radius);
}
size_t box_scratch_bytes(size_t n) { return 2 * sb_align(n * 4) + sb_align(n * 8); }
int box_scratch_get(sb200_ctx* ctx, BoxScratch* s, size_t n) {
    SB_TRY(ws_get(ctx, &s->tmp_f, n));
    SB_TRY(ws_get(ctx, &s->sat, n));
    SB_TRY(ws_get(ctx, &s->tmp_d, n));
    return SB200_OK;
}

// guide statistics as the live reference path computes them (guidedFilter.cu:58-123)
int guide_stats(sb200_ctx* ctx, const sb200_params* p, const uint8_t* d_i, float* I, float* mean_I, float* var_I,
                float* t0, float* t1, uint8_t* d_mean, int w, int h, const BoxScratch& bs) {
    const size_t n = (size_t)w * h;
    SB_TRY(sbk_u8_to_float(ctx, d_i, I, n));
    SB_TRY(box_mean(ctx, p, I, mean_I, w, h, bs));
    SB_TRY(sbk_mul(ctx, I, I, t0, n));
    SB_TRY(box_mean(ctx, p, t0, t1, w, h, bs));
    SB_TRY(sbk_mul(ctx, mean_I, mean_I, t0, n));
    SB_TRY(sbk_sub(ctx, t1, t0, var_I, n));
    if (d_mean) SB_TRY(sbk_float_to_u8(ctx, mean_I, d_mean, n));
    return SB200_OK;
}

size_t gf_ws_bytes(size_t n) { return 15 * sb_align(n * 4) + box_scratch_bytes(n) + 4096; }

// the body of sb200_compute_guided_filter_dev on an already reserved arena.  d_cost == NULL: the slices
// are generated one at a time from (d_i, d_other) instead of being read from a materialised volume.
int guided_filter_staged(sb200_ctx* ctx, const sb200_params* p, const uint8_t* d_i, const float* d_cost, float* d_best,
                         float* d_dmap, uint8_t* d_mean, int w, int h, int size_d, int dmin,
                         const uint8_t* d_other = nullptr) {
    const size_t n = (size_t)w * h;
    float *I, *mean_I, *var_I, *t0, *t1, *mp, *mIp, *a, *b, *ma, *mb, *q;
    float *pkbuf = nullptr, *g1 = nullptr, *g2 = nullptr;
    if (!d_cost) {
        SB_TRY(ws_get(ctx, &pkbuf, n));
        SB_TRY(ws_get(ctx, &g1, n));
        SB_TRY(ws_get(ctx, &g2, n));
        SB_TRY(sbk_x_derivative(ctx, d_i, g1, w, h));
        SB_TRY(sbk_x_derivative(ctx, d_other, g2, w, h));
    }
    SB_TRY(ws_get(ctx, &I, n));
    SB_TRY(ws_get(ctx, &mean_I, n));
    SB_TRY(ws_get(ctx, &var_I, n));
    SB_TRY(ws_get(ctx, &t0, n));
    SB_TRY(ws_get(ctx, &t1, n));
    SB_TRY(ws_get(ctx, &mp, n));
    SB_TRY(ws_get(ctx, &mIp, n));
    SB_TRY(ws_get(ctx, &a, n));
    SB_TRY(ws_get(ctx, &b, n));
    SB_TRY(ws_get(ctx, &ma, n));
    SB_TRY(ws_get(ctx, &mb, n));
    SB_TRY(ws_get(ctx, &q, n));
    BoxScratch bs;
    SB_TRY(box_scratch_get(ctx, &bs, n));
    SB_TRY(guide_stats(ctx, p, d_i, I, mean_I, var_I, t0, t1, d_mean, w, h, bs));
    for (int s = 0; s < size_d; s++) {  // guidedFilter.cu:171-238
        const float* pk = d_cost ? d_cost + (size_t)s * n : pkbuf;
        if (!d_cost) SB_TRY(sbk_cost_volume(ctx, p, d_i, d_other, g1, g2, pkbuf, w, h, 1, dmin + s));
        SB_TRY(box_mean(ctx, p, pk, mp, w, h, bs));
        SB_TRY(sbk_mul(ctx, I, pk, t0, n));
        SB_TRY(box_mean(ctx, p, t0, mIp, w, h, bs));
        SB_TRY(sbk_ak_bk(ctx, mean_I, var_I, mIp, mp, a, b, n, p->eps));
        SB_TRY(box_mean(ctx, p, a, ma, w, h, bs));
        SB_TRY(box_mean(ctx, p, b, mb, w, h, bs));
        SB_TRY(sbk_q(ctx, I, ma, mb, q, n));
        SB_TRY(sbk_disp_select(ctx, q, d_best, d_dmap, n, dmin + s));
    }
    return SB200_OK;
}

// One view with an RGB guide, staged: cost slices are generated one at a time, every
// box filter is a pair of streaming kernels.  Correct, HBM-bound, ~50x slower than the fused gray path;
// a fused RGB kernel is the next step.
size_t rgb_ws_bytes(size_t n) { return 30 * sb_align(n * 4) + box_scratch_bytes(n) + 4096; }
int view_rgb_staged(sb200_ctx* ctx, const sb200_params* p, const uint8_t* d_rgb, int channels, const uint8_t* d_gray,
                    const uint8_t* d_other, float* d_best, float* d_disp, int w, int h, int size_d, int dmin) {
    const size_t n = (size_t)w * h;
    float *I[3], *mu[3], *M[6], *mIp[3], *a[3], *ma[3], *mp, *b, *mb, *t, *q, *pk, *g1, *g2;
    for (int c = 0; c < 3; c++) {
        SB_TRY(ws_get(ctx, &I[c], n));
        SB_TRY(ws_get(ctx, &mu[c], n));
        SB_TRY(ws_get(ctx, &mIp[c], n));
        SB_TRY(ws_get(ctx, &a[c], n));
        SB_TRY(ws_get(ctx, &ma[c], n));
    }
    for (int k = 0; k < 6; k++) SB_TRY(ws_get(ctx, &M[k], n));
    SB_TRY(ws_get(ctx, &mp, n));
    SB_TRY(ws_get(ctx, &b, n));
    SB_TRY(ws_get(ctx, &mb, n));
    SB_TRY(ws_get(ctx, &t, n));
    SB_TRY(ws_get(ctx, &q, n));
    SB_TRY(ws_get(ctx, &pk, n));
    SB_TRY(ws_get(ctx, &g1, n));
    SB_TRY(ws_get(ctx, &g2, n));
    BoxScratch bs;
    SB_TRY(box_scratch_get(ctx, &bs, n));
    SB_TRY(sbk_rgb_split(ctx, d_rgb, channels, I[0], I[1], I[2], n));
    for (int c = 0; c < 3; c++) SB_TRY(box_mean(ctx, p, I[c], mu[c], w, h, bs));
    static const int A[6] = {0, 0, 0, 1, 1, 2}, B[6] = {0, 1, 2, 1, 2, 2};
    for (int k = 0; k < 6; k++) {
        SB_TRY(sbk_mul(ctx, I[A[k]], I[B[k]], t, n));
        SB_TRY(box_mean(ctx, p, t, M[k], w, h, bs));
    }
    SB_TRY(sbk_rgb_inverse(ctx, mu, M, n, p->eps));
    SB_TRY(sbk_x_derivative(ctx, d_gray, g1, w, h));
    SB_TRY(sbk_x_derivative(ctx, d_other, g2, w, h));
    SB_TRY(sbk_fill_f32(ctx, d_best, 3.3961514e38f, n));  // 0x7F7F7F7F, main.cu:112
    SB_TRY(sbk_fill_f32(ctx, d_disp, 0.0f, n));
    for (int s = 0; s < size_d; s++) {
        SB_TRY(sbk_cost_volume(ctx, p, d_gray, d_other, g1, g2, pk, w, h, 1, dmin + s));
        SB_TRY(box_mean(ctx, p, pk, mp, w, h, bs));
        for (int c = 0; c < 3; c++) {
            SB_TRY(sbk_mul(ctx, I[c], pk, t, n));
            SB_TRY(box_mean(ctx, p, t, mIp[c], w, h, bs));
        }
        SB_TRY(sbk_rgb_ab(ctx, mu, M, mIp, mp, a, b, n));
        for (int c = 0; c < 3; c++) SB_TRY(box_mean(ctx, p, a[c], ma[c], w, h, bs));
        SB_TRY(box_mean(ctx, p, b, mb, w, h, bs));
        SB_TRY(sbk_rgb_q(ctx, mb, ma, I, q, n));
        SB_TRY(sbk_disp_select(ctx, q, d_best, d_disp, n, dmin + s));
    }
    return SB200_OK;
}

int check_params(sb200_ctx* ctx, const sb200_params* p) {
    if (!ctx) return SB200_ERR_INVALID;
    if (!p) return sb_fail(ctx, SB200_ERR_INVALID, "params is NULL");
    if (p->dmax < p->dmin) return sb_fail(ctx, SB200_ERR_INVALID, "dmax < dmin");
    if (p->radius < 1 || p->radius > 64) return sb_fail(ctx, SB200_ERR_INVALID, "radius out of range");
    if (p->box_mode != SB200_BOX_SLIDING && p->box_mode != SB200_BOX_SAT) return sb_fail(ctx, SB200_ERR_INVALID, "box_mode");
    return SB200_OK;
}

// the tensor-core kernel needs a slightly smaller cost lattice than the shuffle kernels (sbf_mma_supported)
static inline bool rgb_uses_mma(const sb200_ctx* ctx, const sb200_params* p) { return ctx->rgb_kernel == 4 && sbf_rgb_mma_supported(p); }
size_t rgb_fused_ws_bytes(const sb200_ctx* ctx, const sb200_params* p, int w, int h_held, int rows_out, int dabs, int size_d) {
    if (rgb_uses_mma(ctx, p)) return sbf_rgb_mma_workspace_bytes(ctx, w, h_held, rows_out, dabs, size_d);
    return ctx->rgb_kernel == 2 ? sbf_rgb_workspace_bytes(ctx, w, h_held, rows_out, dabs, size_d)
                                : sbf_rgb3_workspace_bytes(ctx, w, h_held, rows_out, dabs, size_d);
}

// arena bytes pipeline_core needs for one call
size_t pipeline_ws_bytes(const sb200_ctx* ctx, const sb200_params* p, int w, const SbFusedGeom& g, const sb200_outputs* o = nullptr) {
    const int size_d = p->dmax - p->dmin + 1;
    const size_t n_held = (size_t)w * g.h, n_out = (size_t)w * g.rows_out;
    const int dabs = max(abs(p->dmin), abs(p->dmax));
    size_t bytes = p->guide_mode != SB200_GUIDE_RGB
                       ? (sbf_fused_supported(p) ? sbf_workspace_bytes(ctx, w, g.h, g.rows_out, dabs, size_d, 2) : gf_ws_bytes(n_held) + 2 * sb_align(n_held))
                       : (p->box_mode == SB200_BOX_SAT ? rgb_ws_bytes(n_held) : rgb_fused_ws_bytes(ctx, p, w, g.h, g.rows_out, dabs, size_d));
    bytes += 2 * sb_align(n_held) + 4 * sb_align(n_out * 4) + 4096;
    if (g.rows_out != g.h) bytes += 2 * sb_align(n_held);  // strips stage the mean images on held rows
    if (o && o->subpixel_left) bytes += sb_align((size_t)size_d * n_out * 4) + 2 * sb_align(n_out * 4);  // the left view's filtered volume
    return bytes;
}

// pipeline core on device buffers; `held` geometry for strips
int pipeline_core(sb200_ctx* ctx, const sb200_params* p, const uint8_t* d_left, const uint8_t* d_right, int channels,
                  int w, const SbFusedGeom& g, const sb200_outputs* o, bool reserve) {
    const int size_d = p->dmax - p->dmin + 1;
    const size_t n_held = (size_t)w * g.h;
    const size_t n_out = (size_t)w * g.rows_out;
    if (reserve) SB_TRY(sb_ws_reserve(ctx, pipeline_ws_bytes(ctx, p, w, g, o)));
    const bool full = (g.rows_out == g.h);
    const bool rgb_guide = (p->guide_mode == SB200_GUIDE_RGB);
    // RGB guide: fused kernel (fused_cvf_rgb.cu); box_mode SAT selects the staged, materialising path instead
    const bool rgb_staged = rgb_guide && p->box_mode == SB200_BOX_SAT;
    // parameter sets the fused kernels are not built for (radius != 9, no exact cost lattice) run through the
    // stage kernels with double-accumulated sliding boxes: slower, still on the GPU, same results contract
    const bool gray_staged = !rgb_guide && !sbf_fused_supported(p);
    if (gray_staged && !full) return sb_fail(ctx, SB200_ERR_UNSUPPORTED, "these parameters need the staged path, which takes whole frames only");
    if (rgb_guide && !rgb_staged && !sbf_fused_supported(p))
        return sb_fail(ctx, SB200_ERR_UNSUPPORTED, "RGB guide: the fused kernel needs radius 9 and a lattice cost; use box_mode SAT for the staged path");
    if (rgb_guide && channels < 3) return sb_fail(ctx, SB200_ERR_UNSUPPORTED, "RGB guide needs a colour input");
    if (rgb_staged && !full) return sb_fail(ctx, SB200_ERR_UNSUPPORTED, "staged RGB guide needs a whole frame");
    if (rgb_guide && (o->mean_left || o->mean_right))
        return sb_fail(ctx, SB200_ERR_UNSUPPORTED, "mean_left/mean_right are gray-guide outputs");
    // sub-pixel refinement keeps the left view's filtered volume: only the tensor-core gray kernel stores it
    const bool subpix = o->subpixel_left != nullptr;
    if (subpix && (rgb_guide ? (rgb_staged || !rgb_uses_mma(ctx, p)) : (gray_staged || ctx->gray_kernel != 1 || !sbf_mma_supported(p))))
        return sb_fail(ctx, SB200_ERR_UNSUPPORTED, "subpixel_left: the tensor-core fused kernels only (they keep the filtered volume)");
    float* vol = nullptr;
    if (subpix) SB_TRY(ws_get(ctx, &vol, (size_t)size_d * n_out));
    const uint8_t* gl = d_left;
    const uint8_t* gr = d_right;
    if (channels != 1) {
        uint8_t *tl, *tr;
        // gray outputs cover output rows only; write straight into them when the frame is whole
        if (full && o->gray_left) tl = o->gray_left; else SB_TRY(ws_get(ctx, &tl, n_held));
        if (full && o->gray_right) tr = o->gray_right; else SB_TRY(ws_get(ctx, &tr, n_held));
        SB_TRY(sbk_rgb_to_gray(ctx, p, d_left, (int)n_held, channels, tl));
        SB_TRY(sbk_rgb_to_gray(ctx, p, d_right, (int)n_held, channels, tr));
        gl = tl;
        gr = tr;
    }
    const size_t out_off = (size_t)g.y_out0 * w;
    if (!(full && channels != 1)) {
        if (o->gray_left) SB_CUDA(ctx, cudaMemcpyAsync(o->gray_left, gl + out_off, n_out, cudaMemcpyDeviceToDevice, ctx->stream));
        if (o->gray_right) SB_CUDA(ctx, cudaMemcpyAsync(o->gray_right, gr + out_off, n_out, cudaMemcpyDeviceToDevice, ctx->stream));
    }
    float *dL = o->disp_left, *dR = o->disp_right;
    if (!dL) SB_TRY(ws_get(ctx, &dL, n_out));
    if (!dR) SB_TRY(ws_get(ctx, &dR, n_out));
    uint8_t* mL = o->mean_left;
    uint8_t* mR = o->mean_right;
    uint8_t *mLh = nullptr, *mRh = nullptr;
    if (!full) {  // mean images are produced on held rows; stage and crop
        if (mL) SB_TRY(ws_get(ctx, &mLh, n_held));
        if (mR) SB_TRY(ws_get(ctx, &mRh, n_held));
    }
    if (rgb_guide && !rgb_staged) {
        if (rgb_uses_mma(ctx, p)) {
            ctx->qvol[0] = vol;  // for this launch only
            const int rc = sbf_pair_disparity_rgb_mma(ctx, p, d_left, d_right, channels, gl, gr, g, o->best_left, dL, o->best_right, dR);
            ctx->qvol[0] = nullptr;
            SB_TRY(rc);
        } else if (ctx->rgb_kernel != 2)
            SB_TRY(sbf_pair_disparity_rgb3(ctx, p, d_left, d_right, channels, gl, gr, g, o->best_left, dL, o->best_right, dR));
        else
            SB_TRY(sbf_pair_disparity_rgb(ctx, p, d_left, d_right, channels, gl, gr, g, o->best_left, dL, o->best_right, dR));
    } else if (rgb_guide) {
        float *bL = o->best_left, *bR = o->best_right;
        if (!bL) SB_TRY(ws_get(ctx, &bL, n_out));
        if (!bR) SB_TRY(ws_get(ctx, &bR, n_out));
        const size_t mark = ctx->ws_off;
        SB_TRY(view_rgb_staged(ctx, p, d_left, channels, gl, gr, bL, dL, w, g.h, size_d, p->dmin));
        ctx->ws_off = mark;  // the second view reuses the first view's scratch (same stream)
        SB_TRY(view_rgb_staged(ctx, p, d_right, channels, gr, gl, bR, dR, w, g.h, size_d, -p->dmax));
    } else if (gray_staged) {
        sb200_params ps = *p;
        ps.box_mode = SB200_BOX_SLIDING;
        float *bL = o->best_left, *bR = o->best_right;
        if (!bL) SB_TRY(ws_get(ctx, &bL, n_out));
        if (!bR) SB_TRY(ws_get(ctx, &bR, n_out));
        SB_TRY(sbk_fill_f32(ctx, bL, 3.3961514e38f, n_out));  // 0x7F7F7F7F, main.cu:112
        SB_TRY(sbk_fill_f32(ctx, bR, 3.3961514e38f, n_out));
        SB_TRY(sbk_fill_f32(ctx, dL, 0.0f, n_out));
        SB_TRY(sbk_fill_f32(ctx, dR, 0.0f, n_out));
        const size_t mark = ctx->ws_off;
        SB_TRY(guided_filter_staged(ctx, &ps, gl, nullptr, bL, dL, mL, w, g.h, size_d, p->dmin, gr));
        ctx->ws_off = mark;
        SB_TRY(guided_filter_staged(ctx, &ps, gr, nullptr, bR, dR, mR, w, g.h, size_d, -p->dmax, gl));
    } else {
        ctx->qvol[0] = vol;  // for this launch only
        const int rc = sbf_pair_disparity(ctx, p, gl, gr, g, o->best_left, dL, o->best_right, dR, full ? mL : mLh, full ? mR : mRh);
        ctx->qvol[0] = nullptr;
        SB_TRY(rc);
    }
    if (!full) {
        if (mL) SB_CUDA(ctx, cudaMemcpyAsync(mL, mLh + out_off, n_out, cudaMemcpyDeviceToDevice, ctx->stream));
        if (mR) SB_CUDA(ctx, cudaMemcpyAsync(mR, mRh + out_off, n_out, cudaMemcpyDeviceToDevice, ctx->stream));
    }
    if (subpix) {
        float *occ = o->occlusion, *fil = o->filled;
        if (!occ) SB_TRY(ws_get(ctx, &occ, n_out));
        if (!fil) SB_TRY(ws_get(ctx, &fil, n_out));
        SB_TRY(sbk_lr_check_fill(ctx, dL, dR, w, g.rows_out, p->dmin - 100, p->d_lr, (float)p->dmin, occ, fil));
        SB_TRY(sbk_subpixel(ctx, vol, dL, occ, fil, o->subpixel_left, n_out, p->dmin, size_d));
    } else if (o->occlusion || o->filled)
        SB_TRY(sbk_lr_check_fill(ctx, dL, dR, w, g.rows_out, p->dmin - 100, p->d_lr, (float)p->dmin, o->occlusion,
                                 o->filled));
    if (ctx->timing && ctx->ev_valid) SB_CUDA(ctx, cudaEventRecord(ctx->ev[4], ctx->stream));
    return SB200_OK;
}

}  // namespace

extern "C" {

// ---- device-pointer stage entry points -------------------------------------------------
int sb200_rgb_to_grayscale_dev(sb200_ctx* ctx, const sb200_params* p, const uint8_t* d_rgb, int n, int channels,
                               uint8_t* d_gray) {
    DevGuard dev_guard__(ctx);
    SB_TRY(check_params(ctx, p));
    REQUIRE(ctx, d_rgb && d_gray && n > 0, "null pointer or n <= 0");
    REQUIRE(ctx, channels >= 3, "channels must be >= 3 (the reference reads image[ch*i+{0,1,2}])");
    return sbk_rgb_to_gray(ctx, p, d_rgb, n, channels, d_gray);
}

int sb200_compute_cost_dev(sb200_ctx* ctx, const sb200_params* p, const uint8_t* d_i1, const uint8_t* d_i2,
                           float* d_cost, int w, int h, int dmin) {
    DevGuard dev_guard__(ctx);
    SB_TRY(check_params(ctx, p));
    REQUIRE(ctx, d_i1 && d_i2 && d_cost && w > 0 && h > 0, "null pointer or empty image");
    const size_t n = (size_t)w * h;
    SB_TRY(sb_ws_reserve(ctx, 2 * sb_align(n * 4) + 4096));
    float *g1, *g2;
    SB_TRY(ws_get(ctx, &g1, n));
    SB_TRY(ws_get(ctx, &g2, n));
    SB_TRY(sbk_x_derivative(ctx, d_i1, g1, w, h));
    SB_TRY(sbk_x_derivative(ctx, d_i2, g2, w, h));
    return sbk_cost_volume(ctx, p, d_i1, d_i2, g1, g2, d_cost, w, h, p->dmax - p->dmin + 1, dmin);
}

int sb200_integral_dev(sb200_ctx* ctx, const float* d_image, float* d_integral, int w, int h) {
    DevGuard dev_guard__(ctx);
    REQUIRE(ctx, d_image && d_integral && w > 0 && h > 0, "null pointer or empty image");
    const size_t n = (size_t)w * h;
    SB_TRY(sb_ws_reserve(ctx, sb_align(n * 4) + 4096));
    float* tmp;
    SB_TRY(ws_get(ctx, &tmp, n));
    return sbk_integral(ctx, d_image, tmp, d_integral, w, h);
}

int sb200_box_filter_dev(sb200_ctx* ctx, const sb200_params* p, const float* d_image, float* d_mean, int w, int h) {
    DevGuard dev_guard__(ctx);
    SB_TRY(check_params(ctx, p));
    REQUIRE(ctx, d_image && d_mean && w > 0 && h > 0, "null pointer or empty image");
    const size_t n = (size_t)w * h;
    SB_TRY(sb_ws_reserve(ctx, box_scratch_bytes(n) + 4096));
    BoxScratch bs;
    SB_TRY(box_scratch_get(ctx, &bs, n));
    return box_mean(ctx, p, d_image, d_mean, w, h, bs);
}

int sb200_compute_guided_filter_dev(sb200_ctx* ctx, const sb200_params* p, const uint8_t* d_i, const float* d_cost,
                                    float* d_filter_cost, float* d_disp_map, uint8_t* d_mean, int w, int h,
                                    int size_d, int dmin) {
    DevGuard dev_guard__(ctx);
    SB_TRY(check_params(ctx, p));
    REQUIRE(ctx, d_i && d_cost && d_filter_cost && d_disp_map && w > 0 && h > 0 && size_d > 0, "null pointer or empty");
    SB_TRY(sb_ws_reserve(ctx, gf_ws_bytes((size_t)w * h)));
    return guided_filter_staged(ctx, p, d_i, d_cost, d_filter_cost, d_disp_map, d_mean, w, h, size_d, dmin);
}

int sb200_winner_take_all_dev(sb200_ctx* ctx, const float* d_q, float* d_filter_cost, float* d_dmap, int n, int label) {
    DevGuard dev_guard__(ctx);
    REQUIRE(ctx, d_q && d_filter_cost && d_dmap && n > 0, "null pointer or n <= 0");
    return sbk_disp_select(ctx, d_q, d_filter_cost, d_dmap, (size_t)n, label);
}

int sb200_detect_occlusion_dev(sb200_ctx* ctx, const sb200_params* p, float* d_dL, const float* d_dR, int dOcclusion,
                               int w, int h) {
    DevGuard dev_guard__(ctx);
    SB_TRY(check_params(ctx, p));
    REQUIRE(ctx, d_dL && d_dR && w > 0 && h > 0, "null pointer or empty image");
    return sbk_detect_occlusion(ctx, d_dL, d_dR, dOcclusion, p->d_lr, w, h);
}

int sb200_fill_occlusion_dev(sb200_ctx* ctx, float* d_disparity, int w, int h, float vMin) {
    DevGuard dev_guard__(ctx);
    REQUIRE(ctx, d_disparity && w > 0 && h > 0, "null pointer or empty image");
    return sbk_fill_occlusion(ctx, d_disparity, w, h, vMin);
}

int sb200_lr_check_fill_dev(sb200_ctx* ctx, const sb200_params* p, const float* d_dL, const float* d_dR, int w, int h,
                            int dOcclusion, float vMin, float* d_occlusion, float* d_filled) {
    DevGuard dev_guard__(ctx);
    SB_TRY(check_params(ctx, p));
    REQUIRE(ctx, d_dL && d_dR && w > 0 && h > 0 && (d_occlusion || d_filled), "null pointer or empty image");
    return sbk_lr_check_fill(ctx, d_dL, d_dR, w, h, dOcclusion, p->d_lr, vMin, d_occlusion, d_filled);
}

int sb200_view_disparity_dev(sb200_ctx* ctx, const sb200_params* p, const uint8_t* d_guide, const uint8_t* d_other,
                             int w, int h, int dmin, int size_d, float* d_best, float* d_disp, uint8_t* d_mean) {
    DevGuard dev_guard__(ctx);
    SB_TRY(check_params(ctx, p));
    REQUIRE(ctx, d_guide && d_other && w > 1 && h > 0 && size_d > 0 && (d_best || d_disp), "null pointer or empty");
    const int dabs = max(abs(dmin), abs(dmin + size_d - 1));
    SB_TRY(sb_ws_reserve(ctx, sbf_workspace_bytes(ctx, w, h, h, dabs, size_d, 1)));
    SbFusedGeom g{w, h, 0, h, 0, h};
    if (ctx->timing && ctx->ev_valid) {
        int rc = sbf_view_disparity(ctx, p, d_guide, d_other, g, dmin, size_d, d_best, d_disp, d_mean);
        if (rc == SB200_OK) SB_CUDA(ctx, cudaEventRecord(ctx->ev[4], ctx->stream));
        return rc;
    }
    return sbf_view_disparity(ctx, p, d_guide, d_other, g, dmin, size_d, d_best, d_disp, d_mean);
}

// the same kernel, keeping the filtered volume (tensor-core gray kernel only)
int sb200_view_volume_dev(sb200_ctx* ctx, const sb200_params* p, const uint8_t* d_guide, const uint8_t* d_other, int w, int h,
                          int dmin, int size_d, float* d_volume, float* d_best, float* d_disp) {
    DevGuard dev_guard__(ctx);
    SB_TRY(check_params(ctx, p));
    REQUIRE(ctx, d_guide && d_other && d_volume && w > 1 && h > 0 && size_d > 0, "null pointer or empty");
    if (ctx->gray_kernel != 1 || !sbf_fused_supported(p) || !sbf_mma_supported(p))
        return sb_fail(ctx, SB200_ERR_UNSUPPORTED, "filtered volume: tensor-core fused kernel only (radius 9, small exact cost lattice)");
    const int dabs = max(abs(dmin), abs(dmin + size_d - 1));
    SB_TRY(sb_ws_reserve(ctx, sbf_workspace_bytes(ctx, w, h, h, dabs, size_d, 1) + 2 * sb_align((size_t)w * h * 4) + 4096));
    float* best = d_best;
    if (!best && !d_disp) SB_TRY(ws_get(ctx, &best, (size_t)w * h));  // the kernel's merge wants one of the two
    SbFusedGeom g{w, h, 0, h, 0, h};
    ctx->qvol[0] = d_volume;
    const int rc = sbf_view_disparity(ctx, p, d_guide, d_other, g, dmin, size_d, best, d_disp, nullptr);
    ctx->qvol[0] = nullptr;
    return rc;
}

int sb200_subpixel_refine_dev(sb200_ctx* ctx, const float* d_volume, const float* d_disp, const float* d_occlusion,
                              const float* d_filled, float* d_out, int w, int h, int dmin, int size_d) {
    DevGuard dev_guard__(ctx);
    REQUIRE(ctx, d_volume && d_disp && d_out && w > 0 && h > 0 && size_d > 0, "null pointer or empty image");
    return sbk_subpixel(ctx, d_volume, d_disp, d_occlusion, d_filled, d_out, (size_t)w * h, dmin, size_d);
}

// ---- 8-bit visualisation on the device ----------------------------------------
int sb200_write_mat_dev(sb200_ctx* ctx, const float* d_mat, uint8_t* d_out, int w, int h) {
    DevGuard dev_guard__(ctx);
    REQUIRE(ctx, d_mat && d_out && w > 0 && h > 0, "null pointer or empty image");
    const size_t n = (size_t)w * h;
    const size_t nb = (n + 1023) / 1024;
    SB_TRY(sb_ws_reserve(ctx, sb_align((2 * nb + 2) * 4) + 4096));
    float* scratch;
    SB_TRY(ws_get(ctx, &scratch, 2 * nb + 2));
    return sbk_write_mat(ctx, d_mat, d_out, n, scratch);
}
int sb200_write_mat(sb200_ctx* ctx, const float* h_mat, uint8_t* h_out, int w, int h) {
    DevGuard dev_guard__(ctx);
    REQUIRE(ctx, h_mat && h_out && w > 0 && h > 0, "null pointer or empty image");
    const size_t n = (size_t)w * h;
    const size_t nb = (n + 1023) / 1024;
    SB_TRY(sb_ws_reserve(ctx, sb_align(n * 4) + sb_align(n) + sb_align((2 * nb + 2) * 4) + 4096));
    float *d_m, *scratch;
    uint8_t* d_o;
    SB_TRY(ws_get(ctx, &d_m, n));
    SB_TRY(ws_get(ctx, &d_o, n));
    SB_TRY(ws_get(ctx, &scratch, 2 * nb + 2));
    SB_TRY(h2d(ctx, d_m, h_mat, n * 4));
    SB_TRY(sbk_write_mat(ctx, d_m, d_o, n, scratch));
    SB_TRY(d2h(ctx, h_out, d_o, n));
    SB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SB200_OK;
}
int sb200_fl_to_ch2_dev(sb200_ctx* ctx, const float* d_image, uint8_t* d_result, int vmin, int vmax, int len) {
    DevGuard dev_guard__(ctx);
    REQUIRE(ctx, d_image && d_result && len > 0 && vmax != vmin, "null pointer, empty image or max == min");
    return sbk_fl_to_ch2(ctx, d_image, d_result, vmin, vmax, (size_t)len);
}

// ---- fused pipeline ---------------------------------------------------------------------
int sb200_strip_halo_rows(const sb200_params* p) { return p ? 2 * p->radius : 0; }

// which implementation sb200_pipeline* uses for these parameters (the selection logic of pipeline_core, made visible)
int sb200_pipeline_path(const sb200_ctx* ctx, const sb200_params* p) {
    if (!ctx || !p) return -SB200_ERR_INVALID;
    if (p->guide_mode == SB200_GUIDE_RGB) {
        if (p->box_mode == SB200_BOX_SAT) return SB200_PATH_STAGED;
        if (!sbf_fused_supported(p)) return -SB200_ERR_UNSUPPORTED;
        return rgb_uses_mma(ctx, p) ? SB200_PATH_FUSED_TENSOR : SB200_PATH_FUSED_SHUFFLE;
    }
    if (!sbf_fused_supported(p)) return SB200_PATH_STAGED;
    return (ctx->gray_kernel == 1 && sbf_mma_supported(p)) ? SB200_PATH_FUSED_TENSOR : SB200_PATH_FUSED_SHUFFLE;
}

int sb200_pipeline_dev(sb200_ctx* ctx, const sb200_params* p, const uint8_t* d_left, const uint8_t* d_right,
                       int channels, int w, int h, const sb200_outputs* d_out) {
    DevGuard dev_guard__(ctx);
    SB_TRY(check_params(ctx, p));
    REQUIRE(ctx, d_left && d_right && d_out && w > 1 && h > 0, "null pointer or empty image");
    REQUIRE(ctx, channels == 1 || channels >= 3, "channels must be 1 (gray) or >= 3");
    SbFusedGeom g{w, h, 0, h, 0, h};
    return pipeline_core(ctx, p, d_left, d_right, channels, w, g, d_out, true);
}

int sb200_pipeline_strip_dev(sb200_ctx* ctx, const sb200_params* p, const uint8_t* d_left, const uint8_t* d_right,
                             int channels, int w, const sb200_strip* s, const sb200_outputs* d_out) {
    DevGuard dev_guard__(ctx);
    SB_TRY(check_params(ctx, p));
    REQUIRE(ctx, d_left && d_right && d_out && s && w > 1, "null pointer or empty image");
    REQUIRE(ctx, channels == 1 || channels >= 3, "channels must be 1 (gray) or >= 3");
    REQUIRE(ctx, s->rows > 0 && s->halo_top >= 0 && s->halo_bot >= 0 && s->y0 >= s->halo_top &&
                     s->y0 + s->rows + s->halo_bot <= s->frame_h,
            "strip geometry outside the frame");
    const int need = 2 * p->radius;
    REQUIRE(ctx, s->halo_top >= (s->y0 < need ? s->y0 : need), "halo_top smaller than min(y0, 2*radius)");
    const int below = s->frame_h - (s->y0 + s->rows);
    REQUIRE(ctx, s->halo_bot >= (below < need ? below : need), "halo_bot smaller than min(rows below, 2*radius)");
    SbFusedGeom g{w, s->halo_top + s->rows + s->halo_bot, s->halo_top, s->rows, s->y0 - s->halo_top, s->frame_h};
    return pipeline_core(ctx, p, d_left, d_right, channels, w, g, d_out, true);
}

int sb200_pipeline_batch_dev(sb200_ctx* ctx, const sb200_params* p, const uint8_t* d_left, const uint8_t* d_right,
                             int channels, int w, int h, int n_pairs, const sb200_outputs* d_out) {
    DevGuard dev_guard__(ctx);
    SB_TRY(check_params(ctx, p));
    REQUIRE(ctx, d_left && d_right && d_out && w > 1 && h > 0 && n_pairs > 0, "null pointer or empty batch");
    REQUIRE(ctx, channels == 1 || channels >= 3, "channels must be 1 (gray) or >= 3");
    const size_t n = (size_t)w * h;
    SbFusedGeom g{w, h, 0, h, 0, h};
    for (int i = 0; i < n_pairs; i++) {
        sb200_outputs o = *d_out;
        float** fp[] = {&o.disp_left, &o.disp_right, &o.occlusion, &o.filled, &o.best_left, &o.best_right, &o.subpixel_left};
        for (float** f : fp)
            if (*f) *f += (size_t)i * n;
        uint8_t** up[] = {&o.gray_left, &o.gray_right, &o.mean_left, &o.mean_right};
        for (uint8_t** u : up)
            if (*u) *u += (size_t)i * n;
        // pairs run back to back on one stream and share the arena: stream order makes reuse safe
        SB_TRY(pipeline_core(ctx, p, d_left + (size_t)i * n * channels, d_right + (size_t)i * n * channels, channels, w, g,
                             &o, true));
    }
    return SB200_OK;
}

int sb200_pipeline(sb200_ctx* ctx, const sb200_params* p, const uint8_t* h_left, const uint8_t* h_right, int channels,
                   int w, int h, const sb200_outputs* h_out) {
    DevGuard dev_guard__(ctx);
    SB_TRY(check_params(ctx, p));
    REQUIRE(ctx, h_left && h_right && h_out && w > 1 && h > 0, "null pointer or empty image");
    REQUIRE(ctx, channels == 1 || channels >= 3, "channels must be 1 (gray) or >= 3");
    const size_t n = (size_t)w * h;
    const int size_d = p->dmax - p->dmin + 1;
    const int dabs = max(abs(p->dmin), abs(p->dmax));
    size_t bytes = (p->guide_mode != SB200_GUIDE_RGB ? (sbf_fused_supported(p) ? sbf_workspace_bytes(ctx, w, h, h, dabs, size_d, 2) : gf_ws_bytes(n) + 2 * sb_align(n))
                    : (p->box_mode == SB200_BOX_SAT ? rgb_ws_bytes(n) : rgb_fused_ws_bytes(ctx, p, w, h, h, dabs, size_d))) +
                   2 * sb_align(n) + 4 * sb_align(n * 4) + 4096;
    bytes += 2 * sb_align(n * channels) + 7 * sb_align(n * 4) + 4 * sb_align(n) + 4096;
    if (h_out->subpixel_left) bytes += sb_align((size_t)size_d * n * 4) + 2 * sb_align(n * 4);
    SB_TRY(sb_ws_reserve(ctx, bytes));
    uint8_t *dl, *dr;
    SB_TRY(ws_get(ctx, &dl, n * channels));
    SB_TRY(ws_get(ctx, &dr, n * channels));
    SB_TRY(h2d(ctx, dl, h_left, n * channels));
    SB_TRY(h2d(ctx, dr, h_right, n * channels));
    sb200_outputs d{};
    float* const hf[] = {h_out->disp_left, h_out->disp_right, h_out->occlusion, h_out->filled, h_out->best_left, h_out->best_right, h_out->subpixel_left};
    float** const df[] = {&d.disp_left, &d.disp_right, &d.occlusion, &d.filled, &d.best_left, &d.best_right, &d.subpixel_left};
    for (int i = 0; i < 7; i++)
        if (hf[i]) SB_TRY(ws_get(ctx, df[i], n));
    uint8_t* const hu[] = {h_out->gray_left, h_out->gray_right, h_out->mean_left, h_out->mean_right};
    uint8_t** const du[] = {&d.gray_left, &d.gray_right, &d.mean_left, &d.mean_right};
    for (int i = 0; i < 4; i++)
        if (hu[i]) SB_TRY(ws_get(ctx, du[i], n));
    SbFusedGeom g{w, h, 0, h, 0, h};
    SB_TRY(pipeline_core(ctx, p, dl, dr, channels, w, g, &d, false));
    for (int i = 0; i < 7; i++)
        if (hf[i]) SB_TRY(d2h(ctx, hf[i], *df[i], n * 4));
    for (int i = 0; i < 4; i++)
        if (hu[i]) SB_TRY(d2h(ctx, hu[i], *du[i], n));
    SB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SB200_OK;
}

// ---- row strips over NCCL (what main.cu:44-48's single device becomes on an 8-GPU box) ------------------
// The library talks to NCCL through dlopen/dlsym: nothing links against it, a process that never calls this entry
// point never loads it, and the communicator is whatever the host application already has (ncclCommInitRank in a C++
// host, or the raw communicator sharding.py builds in Python).
namespace {
struct NcclApi {
    void* lib = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    int (*Send)(const void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*Recv)(void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    bool ok = false;
};
NcclApi g_nccl;
int nccl_load(sb200_ctx* ctx) {
    if (g_nccl.ok) return SB200_OK;
    const char* names[] = {getenv("SB200_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    for (const char* nm : names) {
        if (!nm) continue;
        g_nccl.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        if (g_nccl.lib) break;
    }
    if (!g_nccl.lib) return sb_fail(ctx, SB200_ERR_UNSUPPORTED, "libnccl.so.2 not found (set SB200_NCCL_LIB): %s", dlerror());
    g_nccl.GroupStart = (int (*)())dlsym(g_nccl.lib, "ncclGroupStart");
    g_nccl.GroupEnd = (int (*)())dlsym(g_nccl.lib, "ncclGroupEnd");
    g_nccl.Send = (int (*)(const void*, size_t, int, int, void*, cudaStream_t))dlsym(g_nccl.lib, "ncclSend");
    g_nccl.Recv = (int (*)(void*, size_t, int, int, void*, cudaStream_t))dlsym(g_nccl.lib, "ncclRecv");
    g_nccl.GetErrorString = (const char* (*)(int))dlsym(g_nccl.lib, "ncclGetErrorString");
    if (!g_nccl.GroupStart || !g_nccl.GroupEnd || !g_nccl.Send || !g_nccl.Recv)
        return sb_fail(ctx, SB200_ERR_UNSUPPORTED, "libnccl lacks ncclSend/ncclRecv/ncclGroupStart/ncclGroupEnd");
    g_nccl.ok = true;
    return SB200_OK;
}
#define SB_NCCL(ctx, call)                                                                                   \
    do {                                                                                                     \
        int r__ = (call);                                                                                    \
        if (r__ != 0)                                                                                        \
            return sb_fail(ctx, SB200_ERR_CUDA, "%s -> NCCL error %d (%s)", #call, r__,                      \
                           g_nccl.GetErrorString ? g_nccl.GetErrorString(r__) : "?");                        \
    } while (0)
}  // namespace

int sb200_strip_rows(int frame_h, int rank, int world, int* y0, int* rows) {
    // balanced split, multiples of 4 rows (the fused kernel's row step) except for the last strip
    if (frame_h < 1 || world < 1 || rank < 0 || rank >= world || !y0 || !rows) return SB200_ERR_INVALID;
    const int units = (frame_h + 3) / 4;
    const int base = units / world, extra = units % world;
    const int u0 = rank * base + (rank < extra ? rank : extra);
    const int un = base + (rank < extra ? 1 : 0);
    *y0 = min(frame_h, 4 * u0);
    *rows = min(frame_h, 4 * (u0 + un)) - *y0;
    return SB200_OK;
}

int sb200_pipeline_strips_nccl(sb200_ctx* ctx, const sb200_params* p, void* nccl_comm, int rank, int world,
                               const uint8_t* d_own_left, const uint8_t* d_own_right, int channels, int w, int frame_h,
                               int y0, int rows, const sb200_outputs* d_out) {
    DevGuard dev_guard__(ctx);
    SB_TRY(check_params(ctx, p));
    REQUIRE(ctx, d_own_left && d_own_right && d_out && w > 1, "null pointer or empty image");
    REQUIRE(ctx, channels == 1 || channels >= 3, "channels must be 1 (gray) or >= 3");
    REQUIRE(ctx, world >= 1 && rank >= 0 && rank < world, "rank outside the communicator");
    REQUIRE(ctx, rows > 0 && y0 >= 0 && y0 + rows <= frame_h, "strip outside the frame");
    const int halo = 2 * p->radius;
    REQUIRE(ctx, world == 1 || rows >= halo, "a strip must be at least 2*radius rows tall (its neighbours need that many)");
    REQUIRE(ctx, (rank == 0) == (y0 == 0) && (rank == world - 1) == (y0 + rows == frame_h), "strips must tile the frame in rank order");
    REQUIRE(ctx, world == 1 || nccl_comm, "communicator is NULL");
    REQUIRE(ctx, !d_out->subpixel_left, "subpixel_left: not in the NCCL strip entry (use sb200_pipeline_strip_dev)");
    const int top = rank > 0 ? halo : 0, bot = rank < world - 1 ? halo : 0;
    if (world > 1) SB_TRY(nccl_load(ctx));
    const size_t row_bytes = (size_t)w * channels;
    const int held = top + rows + bot;
    SbFusedGeom g{w, held, top, rows, y0 - top, frame_h};
    SB_TRY(sb_ws_reserve(ctx, pipeline_ws_bytes(ctx, p, w, g) + 2 * sb_align(row_bytes * held) + 4096));
    uint8_t *hl, *hr;
    SB_TRY(ws_get(ctx, &hl, row_bytes * held));
    SB_TRY(ws_get(ctx, &hr, row_bytes * held));
    const uint8_t* own[2] = {d_own_left, d_own_right};
    uint8_t* hld[2] = {hl, hr};
    for (int i = 0; i < 2; i++)
        SB_CUDA(ctx, cudaMemcpyAsync(hld[i] + row_bytes * top, own[i], row_bytes * rows, cudaMemcpyDeviceToDevice, ctx->stream));
    if (ctx->timing && ctx->ev_valid) SB_CUDA(ctx, cudaEventRecord(ctx->ev_x[0], ctx->stream));
    if (world > 1) {
        // one grouped exchange: 2*radius INPUT rows of both images with each neighbour (ncclUint8 = 1)
        SB_NCCL(ctx, g_nccl.GroupStart());
        for (int i = 0; i < 2; i++) {
            if (top) {
                SB_NCCL(ctx, g_nccl.Recv(hld[i], row_bytes * top, 1, rank - 1, nccl_comm, ctx->stream));
                SB_NCCL(ctx, g_nccl.Send(own[i], row_bytes * halo, 1, rank - 1, nccl_comm, ctx->stream));
            }
            if (bot) {
                SB_NCCL(ctx, g_nccl.Send(own[i] + row_bytes * (rows - halo), row_bytes * halo, 1, rank + 1, nccl_comm, ctx->stream));
                SB_NCCL(ctx, g_nccl.Recv(hld[i] + row_bytes * (top + rows), row_bytes * bot, 1, rank + 1, nccl_comm, ctx->stream));
            }
        }
        SB_NCCL(ctx, g_nccl.GroupEnd());
    }
    if (ctx->timing && ctx->ev_valid) SB_CUDA(ctx, cudaEventRecord(ctx->ev_x[1], ctx->stream));
    return pipeline_core(ctx, p, hl, hr, channels, w, g, d_out, false);
}

// Host pointers, n_pairs pairs, OVERLAPPED: the upload of pair i+1 and the download of pair i-1 run on their own
// streams under the kernels of pair i (two device staging slots, events between the three streams).  This is what the
// reference's per-stage cudaMemcpy round trips (guidedFilter.cu:39-56, costVolume.cu:23-53) become on a bus that is
// 100x slower than HBM.  Host buffers should be page-locked (cudaHostAlloc / cudaHostRegister / torch pin_memory): with
// pageable memory the copies still work but the driver stages them and the overlap is lost.
}  // extern "C"
namespace {
// h_i16 != NULL: the four label maps leave as int16 (h_out then only carries the non-label outputs, or is NULL)
int batch_host(sb200_ctx* ctx, const sb200_params* p, const uint8_t* h_left, const uint8_t* h_right, int channels, int w, int h,
               int n_pairs, const sb200_outputs* h_out_in, const sb200_labels_i16* h_i16) {
    DevGuard dev_guard__(ctx);
    SB_TRY(check_params(ctx, p));
    static const sb200_outputs no_outputs{};
    const sb200_outputs* h_out = h_out_in ? h_out_in : &no_outputs;
    REQUIRE(ctx, h_left && h_right && (h_out_in || h_i16) && w > 1 && h > 0 && n_pairs > 0, "null pointer or empty batch");
    REQUIRE(ctx, channels == 1 || channels >= 3, "channels must be 1 (gray) or >= 3");
    REQUIRE(ctx, !h_out->subpixel_left, "subpixel_left: not in the overlapped batch entry (use sb200_pipeline / sb200_pipeline_batch_dev)");
    int16_t* const hi[4] = {h_i16 ? h_i16->disp_left : nullptr, h_i16 ? h_i16->disp_right : nullptr,
                            h_i16 ? h_i16->occlusion : nullptr, h_i16 ? h_i16->filled : nullptr};
    if (h_i16) {
        REQUIRE(ctx, p->dmin - 100 >= -32768 && p->dmax <= 32767 && -p->dmax >= -32768 && -p->dmin <= 32767, "labels do not fit int16");
        REQUIRE(ctx, !(hi[0] && h_out->disp_left) && !(hi[1] && h_out->disp_right) && !(hi[2] && h_out->occlusion) && !(hi[3] && h_out->filled),
                "a label map is requested both as float and as int16");
    }
    const size_t n = (size_t)w * h;
    if (!ctx->batch_ready) {
        SB_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->s_in, cudaStreamNonBlocking));
        SB_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->s_out, cudaStreamNonBlocking));
        for (int i = 0; i < 2; i++) {
            SB_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_in[i], cudaEventDisableTiming));
            SB_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_comp[i], cudaEventDisableTiming));
            SB_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_out[i], cudaEventDisableTiming));
        }
        ctx->batch_ready = true;
    }
    float* const hf[] = {h_out->disp_left, h_out->disp_right, h_out->occlusion, h_out->filled, h_out->best_left, h_out->best_right};
    uint8_t* const hu[] = {h_out->gray_left, h_out->gray_right, h_out->mean_left, h_out->mean_right};
    size_t slot_bytes = 2 * sb_align(n * channels);
    for (int i = 0; i < 6; i++)
        if (hf[i] || (i < 4 && hi[i])) slot_bytes += sb_align(n * 4);
    for (int i = 0; i < 4; i++)
        if (hi[i]) slot_bytes += sb_align(n * 2);
    for (int i = 0; i < 4; i++)
        if (hu[i]) slot_bytes += sb_align(n);
    if (2 * slot_bytes > ctx->stage_cap) {
        SB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        SB_CUDA(ctx, cudaStreamSynchronize(ctx->s_in));
        SB_CUDA(ctx, cudaStreamSynchronize(ctx->s_out));
        if (ctx->stage) SB_CUDA(ctx, cudaFree(ctx->stage));
        ctx->stage = nullptr;
        ctx->stage_cap = 0;
        cudaError_t e = cudaMalloc(&ctx->stage, 2 * slot_bytes);
        if (e != cudaSuccess) return sb_fail(ctx, SB200_ERR_NOMEM, "staging of %zu bytes: %s", 2 * slot_bytes, cudaGetErrorString(e));
        ctx->stage_cap = 2 * slot_bytes;
    }
    // the copy streams start after whatever the caller queued on the compute stream
    if (ctx->ev_stream) {
        SB_CUDA(ctx, cudaEventRecord(ctx->ev_stream, ctx->stream));
        SB_CUDA(ctx, cudaStreamWaitEvent(ctx->s_in, ctx->ev_stream, 0));
        SB_CUDA(ctx, cudaStreamWaitEvent(ctx->s_out, ctx->ev_stream, 0));
    }
    SbFusedGeom g{w, h, 0, h, 0, h};
    for (int i = 0; i < n_pairs; i++) {
        const int slot = i & 1;
        char* base = ctx->stage + (size_t)slot * slot_bytes;
        uint8_t* dl = reinterpret_cast<uint8_t*>(base);
        uint8_t* dr = dl + sb_align(n * channels);
        char* o = reinterpret_cast<char*>(dr) + sb_align(n * channels);
        sb200_outputs d{};
        float** const df[] = {&d.disp_left, &d.disp_right, &d.occlusion, &d.filled, &d.best_left, &d.best_right};
        uint8_t** const du[] = {&d.gray_left, &d.gray_right, &d.mean_left, &d.mean_right};
        for (int k = 0; k < 6; k++)
            if (hf[k] || (k < 4 && hi[k])) {
                *df[k] = reinterpret_cast<float*>(o);
                o += sb_align(n * 4);
            }
        for (int k = 0; k < 4; k++)
            if (hu[k]) {
                *du[k] = reinterpret_cast<uint8_t*>(o);
                o += sb_align(n);
            }
        int16_t* di[4] = {nullptr, nullptr, nullptr, nullptr};
        const float* fsrc[4] = {nullptr, nullptr, nullptr, nullptr};
        for (int k = 0; k < 4; k++)
            if (hi[k]) {
                di[k] = reinterpret_cast<int16_t*>(o);
                fsrc[k] = *df[k];
                o += sb_align(n * 2);
            }
        // upload pair i once the kernels of pair i-2 are done with this slot's inputs
        if (i >= 2) SB_CUDA(ctx, cudaStreamWaitEvent(ctx->s_in, ctx->ev_comp[slot], 0));
        SB_CUDA(ctx, cudaMemcpyAsync(dl, h_left + (size_t)i * n * channels, n * channels, cudaMemcpyHostToDevice, ctx->s_in));
        SB_CUDA(ctx, cudaMemcpyAsync(dr, h_right + (size_t)i * n * channels, n * channels, cudaMemcpyHostToDevice, ctx->s_in));
        SB_CUDA(ctx, cudaEventRecord(ctx->ev_in[slot], ctx->s_in));
        // kernels of pair i: after its upload, and after pair i-2's results have left this slot
        SB_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_in[slot], 0));
        if (i >= 2) SB_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_out[slot], 0));
        SB_TRY(pipeline_core(ctx, p, dl, dr, channels, w, g, &d, true));
        if (h_i16) SB_TRY(sbk_labels_i16(ctx, fsrc, di, n));
        SB_CUDA(ctx, cudaEventRecord(ctx->ev_comp[slot], ctx->stream));
        // download the results of pair i
        SB_CUDA(ctx, cudaStreamWaitEvent(ctx->s_out, ctx->ev_comp[slot], 0));
        for (int k = 0; k < 4; k++)
            if (hi[k]) SB_CUDA(ctx, cudaMemcpyAsync(hi[k] + (size_t)i * n, di[k], n * 2, cudaMemcpyDeviceToHost, ctx->s_out));
        for (int k = 0; k < 6; k++)
            if (hf[k]) SB_CUDA(ctx, cudaMemcpyAsync(hf[k] + (size_t)i * n, *df[k], n * 4, cudaMemcpyDeviceToHost, ctx->s_out));
        for (int k = 0; k < 4; k++)
            if (hu[k]) SB_CUDA(ctx, cudaMemcpyAsync(hu[k] + (size_t)i * n, *du[k], n, cudaMemcpyDeviceToHost, ctx->s_out));
        SB_CUDA(ctx, cudaEventRecord(ctx->ev_out[slot], ctx->s_out));
    }
    SB_CUDA(ctx, cudaStreamSynchronize(ctx->s_out));
    SB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SB200_OK;
}
}  // namespace
extern "C" {
int sb200_pipeline_batch(sb200_ctx* ctx, const sb200_params* p, const uint8_t* h_left, const uint8_t* h_right, int channels,
                         int w, int h, int n_pairs, const sb200_outputs* h_out) {
    if (ctx && !h_out) return sb_fail(ctx, SB200_ERR_INVALID, "outputs is NULL");
    return batch_host(ctx, p, h_left, h_right, channels, w, h, n_pairs, h_out, nullptr);
}
int sb200_pipeline_batch_i16(sb200_ctx* ctx, const sb200_params* p, const uint8_t* h_left, const uint8_t* h_right, int channels,
                             int w, int h, int n_pairs, const sb200_labels_i16* h_labels, const sb200_outputs* h_other) {
    if (ctx && !h_labels) return sb_fail(ctx, SB200_ERR_INVALID, "labels is NULL");
    return batch_host(ctx, p, h_left, h_right, channels, w, h, n_pairs, h_other, h_labels);
}

// ---- host-pointer stage entry points (blocking, reference calling convention) -------------
int sb200_rgb_to_grayscale(sb200_ctx* ctx, const sb200_params* p, const uint8_t* h_rgb, int n, int channels,
                           uint8_t* h_gray) {
    DevGuard dev_guard__(ctx);
    SB_TRY(check_params(ctx, p));
    REQUIRE(ctx, h_rgb && h_gray && n > 0, "null pointer or n <= 0");
    REQUIRE(ctx, channels >= 3, "channels must be >= 3 (the reference reads image[ch*i+{0,1,2}])");
    SB_TRY(sb_ws_reserve(ctx, sb_align((size_t)n * channels) + sb_align(n) + 4096));
    uint8_t *d_rgb, *d_gray;
    SB_TRY(ws_get(ctx, &d_rgb, (size_t)n * channels));
    SB_TRY(ws_get(ctx, &d_gray, (size_t)n));
    SB_TRY(h2d(ctx, d_rgb, h_rgb, (size_t)n * channels));
    SB_TRY(sbk_rgb_to_gray(ctx, p, d_rgb, n, channels, d_gray));
    SB_TRY(d2h(ctx, h_gray, d_gray, n));
    SB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SB200_OK;
}

int sb200_x_derivative(sb200_ctx* ctx, const uint8_t* img, float* grad, int w, int h) {
    DevGuard dev_guard__(ctx);
    REQUIRE(ctx, img && grad && w > 0 && h > 0, "null pointer or empty image");
    const size_t n = (size_t)w * h;
    SB_TRY(sb_ws_reserve(ctx, sb_align(n) + sb_align(n * 4) + 4096));
    uint8_t* d_i;
    float* d_g;
    SB_TRY(ws_get(ctx, &d_i, n));
    SB_TRY(ws_get(ctx, &d_g, n));
    SB_TRY(h2d(ctx, d_i, img, n));
    SB_TRY(sbk_x_derivative(ctx, d_i, d_g, w, h));
    SB_TRY(d2h(ctx, grad, d_g, n * 4));
    SB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SB200_OK;
}

int sb200_compute_cost(sb200_ctx* ctx, const sb200_params* p, const uint8_t* i1, const uint8_t* i2, float* cost, int w1,
                       int w2, int h1, int h2, int dmin) {
    DevGuard dev_guard__(ctx);
    SB_TRY(check_params(ctx, p));
    REQUIRE(ctx, i1 && i2 && cost && w1 > 0 && h1 > 0, "null pointer or empty image");
    REQUIRE(ctx, w1 == w2 && h1 == h2, "the two images must have the same size (costVolume.cu assumes it)");
    const size_t n = (size_t)w1 * h1;
    const int size_d = p->dmax - p->dmin + 1;
    SB_TRY(sb_ws_reserve(ctx, 2 * sb_align(n) + 2 * sb_align(n * 4) + sb_align(n * size_d * 4) + 4096));
    uint8_t *d1, *d2;
    float *g1, *g2, *dc;
    SB_TRY(ws_get(ctx, &d1, n));
    SB_TRY(ws_get(ctx, &d2, n));
    SB_TRY(ws_get(ctx, &g1, n));
    SB_TRY(ws_get(ctx, &g2, n));
    SB_TRY(ws_get(ctx, &dc, n * size_d));
    SB_TRY(h2d(ctx, d1, i1, n));
    SB_TRY(h2d(ctx, d2, i2, n));
    SB_TRY(sbk_x_derivative(ctx, d1, g1, w1, h1));
    SB_TRY(sbk_x_derivative(ctx, d2, g2, w1, h1));
    SB_TRY(sbk_cost_volume(ctx, p, d1, d2, g1, g2, dc, w1, h1, size_d, dmin));
    SB_TRY(d2h(ctx, cost, dc, n * size_d * 4));
    SB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SB200_OK;
}

int sb200_integral(sb200_ctx* ctx, const float* image, float* integral, int width, int height) {
    DevGuard dev_guard__(ctx);
    REQUIRE(ctx, image && integral && width > 0 && height > 0, "null pointer or empty image");
    const size_t n = (size_t)width * height;
    SB_TRY(sb_ws_reserve(ctx, 3 * sb_align(n * 4) + 4096));
    float *d_in, *d_tmp, *d_out;
    SB_TRY(ws_get(ctx, &d_in, n));
    SB_TRY(ws_get(ctx, &d_tmp, n));
    SB_TRY(ws_get(ctx, &d_out, n));
    SB_TRY(h2d(ctx, d_in, image, n * 4));
    SB_TRY(sbk_integral(ctx, d_in, d_tmp, d_out, width, height));
    SB_TRY(d2h(ctx, integral, d_out, n * 4));
    SB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SB200_OK;
}

int sb200_box_filter_sat(sb200_ctx* ctx, const sb200_params* p, const float* integral, float* mean, int w, int h) {
    DevGuard dev_guard__(ctx);
    SB_TRY(check_params(ctx, p));
    REQUIRE(ctx, integral && mean && w > 0 && h > 0, "null pointer or empty image");
    const size_t n = (size_t)w * h;
    SB_TRY(sb_ws_reserve(ctx, 2 * sb_align(n * 4) + 4096));
    float *d_s, *d_m;
    SB_TRY(ws_get(ctx, &d_s, n));
    SB_TRY(ws_get(ctx, &d_m, n));
    SB_TRY(h2d(ctx, d_s, integral, n * 4));
    SB_TRY(sbk_box_from_sat(ctx, d_s, d_m, w, h, p->radius));
    SB_TRY(d2h(ctx, mean, d_m, n * 4));
    SB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SB200_OK;
}

int sb200_box_filter(sb200_ctx* ctx, const sb200_params* p, const float* image, float* mean, int w, int h) {
    DevGuard dev_guard__(ctx);
    SB_TRY(check_params(ctx, p));
    REQUIRE(ctx, image && mean && w > 0 && h > 0, "null pointer or empty image");
    const size_t n = (size_t)w * h;
    SB_TRY(sb_ws_reserve(ctx, 2 * sb_align(n * 4) + box_scratch_bytes(n) + 4096));
    float *d_i, *d_m;
    SB_TRY(ws_get(ctx, &d_i, n));
    SB_TRY(ws_get(ctx, &d_m, n));
    BoxScratch bs;
    SB_TRY(box_scratch_get(ctx, &bs, n));
    SB_TRY(h2d(ctx, d_i, image, n * 4));
    SB_TRY(box_mean(ctx, p, d_i, d_m, w, h, bs));
    SB_TRY(d2h(ctx, mean, d_m, n * 4));
    SB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SB200_OK;
}

int sb200_filter(sb200_ctx* ctx, const sb200_params* p, const uint8_t* image, int width, int height, uint8_t* mean,
                 float* var) {
    DevGuard dev_guard__(ctx);
    SB_TRY(check_params(ctx, p));
    REQUIRE(ctx, image && (mean || var) && width > 0 && height > 0, "null pointer or empty image");
    const size_t n = (size_t)width * height;
    SB_TRY(sb_ws_reserve(ctx, 2 * sb_align(n) + 5 * sb_align(n * 4) + box_scratch_bytes(n) + 4096));
    uint8_t *d_i, *d_m;
    float *I, *mean_I, *var_I, *t0, *t1;
    SB_TRY(ws_get(ctx, &d_i, n));
    SB_TRY(ws_get(ctx, &d_m, n));
    SB_TRY(ws_get(ctx, &I, n));
    SB_TRY(ws_get(ctx, &mean_I, n));
    SB_TRY(ws_get(ctx, &var_I, n));
    SB_TRY(ws_get(ctx, &t0, n));
    SB_TRY(ws_get(ctx, &t1, n));
    BoxScratch bs;
    SB_TRY(box_scratch_get(ctx, &bs, n));
    SB_TRY(h2d(ctx, d_i, image, n));
    SB_TRY(guide_stats(ctx, p, d_i, I, mean_I, var_I, t0, t1, d_m, width, height, bs));
    if (mean) SB_TRY(d2h(ctx, mean, d_m, n));
    if (var) SB_TRY(d2h(ctx, var, var_I, n * 4));
    SB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SB200_OK;
}

int sb200_compute_guided_filter(sb200_ctx* ctx, const sb200_params* p, const uint8_t* i, const float* cost,
                                float* filter_cost, float* disp_map, uint8_t* mean, int w, int h, int size_d, int dmin) {
    DevGuard dev_guard__(ctx);
    SB_TRY(check_params(ctx, p));
    REQUIRE(ctx, i && cost && filter_cost && disp_map && w > 0 && h > 0 && size_d > 0, "null pointer or empty");
    const size_t n = (size_t)w * h;
    SB_TRY(sb_ws_reserve(ctx, gf_ws_bytes(n) + 2 * sb_align(n) + 2 * sb_align(n * 4) + sb_align(n * size_d * 4) + 4096));
    uint8_t *d_i, *d_m;
    float *d_c, *d_b, *d_d;
    SB_TRY(ws_get(ctx, &d_i, n));
    SB_TRY(ws_get(ctx, &d_m, n));
    SB_TRY(ws_get(ctx, &d_c, n * size_d));
    SB_TRY(ws_get(ctx, &d_b, n));
    SB_TRY(ws_get(ctx, &d_d, n));
    SB_TRY(h2d(ctx, d_i, i, n));
    SB_TRY(h2d(ctx, d_c, cost, n * size_d * 4));
    SB_TRY(h2d(ctx, d_b, filter_cost, n * 4));
    SB_TRY(h2d(ctx, d_d, disp_map, n * 4));
    SB_TRY(guided_filter_staged(ctx, p, d_i, d_c, d_b, d_d, d_m, w, h, size_d, dmin));
    SB_TRY(d2h(ctx, filter_cost, d_b, n * 4));
    SB_TRY(d2h(ctx, disp_map, d_d, n * 4));
    if (mean) SB_TRY(d2h(ctx, mean, d_m, n));
    SB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SB200_OK;
}

int sb200_winner_take_all(sb200_ctx* ctx, const float* q, float* filter_cost, float* dmap, int n, int label) {
    DevGuard dev_guard__(ctx);
    REQUIRE(ctx, q && filter_cost && dmap && n > 0, "null pointer or n <= 0");
    const size_t nn = (size_t)n;
    SB_TRY(sb_ws_reserve(ctx, 3 * sb_align(nn * 4) + 4096));
    float *d_q, *d_b, *d_d;
    SB_TRY(ws_get(ctx, &d_q, nn));
    SB_TRY(ws_get(ctx, &d_b, nn));
    SB_TRY(ws_get(ctx, &d_d, nn));
    SB_TRY(h2d(ctx, d_q, q, nn * 4));
    SB_TRY(h2d(ctx, d_b, filter_cost, nn * 4));
    SB_TRY(h2d(ctx, d_d, dmap, nn * 4));
    SB_TRY(sbk_disp_select(ctx, d_q, d_b, d_d, nn, label));
    SB_TRY(d2h(ctx, filter_cost, d_b, nn * 4));
    SB_TRY(d2h(ctx, dmap, d_d, nn * 4));
    SB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SB200_OK;
}

int sb200_detect_occlusion(sb200_ctx* ctx, const sb200_params* p, float* disparityLeft, const float* disparityRight,
                           int dOcclusion, uint8_t* dmapl, uint8_t* dmapr, int w, int h) {
    DevGuard dev_guard__(ctx);
    (void)dmapl;  // passed through untouched by the reference (occlusion.cu:40-41,55-56)
    (void)dmapr;
    SB_TRY(check_params(ctx, p));
    REQUIRE(ctx, disparityLeft && disparityRight && w > 0 && h > 0, "null pointer or empty image");
    const size_t n = (size_t)w * h;
    SB_TRY(sb_ws_reserve(ctx, 2 * sb_align(n * 4) + 4096));
    float *d_l, *d_r;
    SB_TRY(ws_get(ctx, &d_l, n));
    SB_TRY(ws_get(ctx, &d_r, n));
    SB_TRY(h2d(ctx, d_l, disparityLeft, n * 4));
    SB_TRY(h2d(ctx, d_r, disparityRight, n * 4));
    SB_TRY(sbk_detect_occlusion(ctx, d_l, d_r, dOcclusion, p->d_lr, w, h));
    SB_TRY(d2h(ctx, disparityLeft, d_l, n * 4));
    SB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SB200_OK;
}

int sb200_fill_occlusion(sb200_ctx* ctx, float* disparity, int w, int h, float vMin) {
    DevGuard dev_guard__(ctx);
    REQUIRE(ctx, disparity && w > 0 && h > 0, "null pointer or empty image");
    const size_t n = (size_t)w * h;
    SB_TRY(sb_ws_reserve(ctx, sb_align(n * 4) + 4096));
    float* d_d;
    SB_TRY(ws_get(ctx, &d_d, n));
    SB_TRY(h2d(ctx, d_d, disparity, n * 4));
    SB_TRY(sbk_fill_occlusion(ctx, d_d, w, h, vMin));
    SB_TRY(d2h(ctx, disparity, d_d, n * 4));
    SB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SB200_OK;
}


// ---- post-processing beyond the reference: weighted median of the marked pixels ----
void sb200_default_wmedian_params(sb200_wmedian_params* wp) {
    if (!wp) return;
    wp->radius = 19;
    wp->sigma_space = 9.0f;
    wp->sigma_color = 25.5f;
}

static int wmedian_tables(sb200_ctx* ctx, const sb200_wmedian_params* wp, unsigned** d_ws, unsigned** d_wc) {
    // the integer weight tables of the definition in stereo_b200.h (double arithmetic on the host, a few KB)
    const int R = wp->radius;
    std::vector<unsigned> t((size_t)(R + 1) * (R + 1) + 256);
    const double ss = (double)wp->sigma_space * (double)wp->sigma_space, sc = (double)wp->sigma_color * (double)wp->sigma_color;
    for (int dy = 0; dy <= R; dy++)
        for (int dx = 0; dx <= R; dx++) t[(size_t)dy * (R + 1) + dx] = (unsigned)floor(1024.0 * exp(-(double)(dx * dx + dy * dy) / ss) + 0.5);
    for (int d = 0; d < 256; d++) t[(size_t)(R + 1) * (R + 1) + d] = (unsigned)floor(1024.0 * exp(-(double)(d * d) / sc) + 0.5);
    unsigned* d_t;
    SB_TRY(ws_get(ctx, &d_t, t.size()));
    // pageable source: the copy is staged by the runtime before the call returns, so the vector may go out of scope
    SB_CUDA(ctx, cudaMemcpyAsync(d_t, t.data(), t.size() * sizeof(unsigned), cudaMemcpyHostToDevice, ctx->stream));
    *d_ws = d_t;
    *d_wc = d_t + (size_t)(R + 1) * (R + 1);
    return SB200_OK;
}

static int wmedian_check(sb200_ctx* ctx, const sb200_params* p, const sb200_wmedian_params* wp) {
    SB_TRY(check_params(ctx, p));
    REQUIRE(ctx, wp && wp->radius >= 1 && wp->radius <= 32 && wp->sigma_space > 0.0f && wp->sigma_color > 0.0f,
            "weighted median: radius in [1, 32], positive sigmas");
    return SB200_OK;
}

int sb200_weighted_median_dev(sb200_ctx* ctx, const sb200_params* p, const sb200_wmedian_params* wp, const uint8_t* d_gray,
                              const float* d_occlusion, const float* d_filled, float* d_out, int w, int h) {
    DevGuard dev_guard__(ctx);
    SB_TRY(wmedian_check(ctx, p, wp));
    REQUIRE(ctx, d_gray && d_occlusion && d_filled && d_out && w > 0 && h > 0, "null pointer or empty image");
    REQUIRE(ctx, d_out != d_filled, "weighted median: out must not alias filled (the windows read filled)");
    SB_TRY(sb_ws_reserve(ctx, 64 * 1024));
    unsigned *ws, *wc;
    SB_TRY(wmedian_tables(ctx, wp, &ws, &wc));
    return sbk_weighted_median(ctx, d_gray, d_occlusion, d_filled, d_out, w, h, p->dmin, p->dmax - p->dmin + 1, wp->radius, ws, wc);
}

int sb200_weighted_median(sb200_ctx* ctx, const sb200_params* p, const sb200_wmedian_params* wp, const uint8_t* gray,
                          const float* occlusion, const float* filled, float* out, int w, int h) {
    DevGuard dev_guard__(ctx);
    SB_TRY(wmedian_check(ctx, p, wp));
    REQUIRE(ctx, gray && occlusion && filled && out && w > 0 && h > 0, "null pointer or empty image");
    const size_t n = (size_t)w * h;
    SB_TRY(sb_ws_reserve(ctx, sb_align(n) + 3 * sb_align(n * 4) + 64 * 1024 + 4096));
    uint8_t* d_g;
    float *d_o, *d_f, *d_r;
    SB_TRY(ws_get(ctx, &d_g, n));
    SB_TRY(ws_get(ctx, &d_o, n));
    SB_TRY(ws_get(ctx, &d_f, n));
    SB_TRY(ws_get(ctx, &d_r, n));
    SB_TRY(h2d(ctx, d_g, gray, n));
    SB_TRY(h2d(ctx, d_o, occlusion, n * 4));
    SB_TRY(h2d(ctx, d_f, filled, n * 4));
    unsigned *ws, *wc;
    SB_TRY(wmedian_tables(ctx, wp, &ws, &wc));
    SB_TRY(sbk_weighted_median(ctx, d_g, d_o, d_f, d_r, w, h, p->dmin, p->dmax - p->dmin + 1, wp->radius, ws, wc));
    SB_TRY(d2h(ctx, out, d_r, n * 4));
    SB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SB200_OK;
}

}  // extern "C"
