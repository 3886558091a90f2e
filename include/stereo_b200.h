/*
 * stereo_b200.h -- C ABI of the B200-native local stereo pipeline.
 *
 * Drop-in boundary for hamza1030/stereo_matching_cuda's host stage functions (the free
 * functions declared in its .cuh headers and called from main.cu:65-155).  Every entry
 * point below names the reference interface it replaces (paths relative to
 * stereo_matching_cuda/ in the reference).  Differences that are deliberate:
 *   - extern "C", int status return (0 = ok), never exit(): the reference's CHECK macro
 *     prints and calls exit(0) (SystemIncludes.h:46-52);
 *   - an explicit context owns the device, the stream and a workspace arena, so no entry
 *     point cudaMallocs per call once warm (the reference allocates and frees inside every
 *     stage: e.g. guidedFilter.cu:39-56, integral.cu:12-16);
 *   - the reference's compile-time macros (SystemIncludes.h:6-24) are the runtime
 *     sb200_params; sb200_default_params() yields exactly the macro values;
 *   - `_dev` variants take device pointers and run asynchronously on the context's stream
 *     so stages compose without PCIe round trips.  The plain variants take HOST pointers and
 *     block, like the reference's.
 * There is no CPU fallback: every entry point fails with SB200_ERR_CUDA if no device works.
 *
 * Layouts (same as the reference): images row-major, RGB interleaved with `channels` bytes
 * per pixel, cost volumes planar D-major cost[k*w*h + y*w + x] (costVolume.cu:178),
 * disparity labels stored as float (guidedFilter.cu:406).
 */
#ifndef STEREO_B200_H
#define STEREO_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SB200_OK 0
#define SB200_ERR_INVALID 1     /* bad argument (null pointer, non-positive size, ...) */
#define SB200_ERR_CUDA 2        /* a CUDA runtime call or kernel failed; see sb200_last_error */
#define SB200_ERR_UNSUPPORTED 3 /* valid request outside what the kernels implement */
#define SB200_ERR_NOMEM 4

#define SB200_GUIDE_GRAY 0 /* the reference's only mode (main.cu:65-66) */
#define SB200_GUIDE_RGB 1  /* colour guided filter, SURVEY.md A.8 (not in the reference) */

#define SB200_BOX_SLIDING 0 /* separable sliding-window sums (exact for the cost stage) */
#define SB200_BOX_SAT 1     /* float32 summed-area table, bit-faithful to integral.cu */

typedef struct sb200_ctx sb200_ctx;

/* SystemIncludes.h:6-24 as data.  sb200_default_params() fills the macro values. */
typedef struct sb200_params {
    int dmin;        /* D_MIN  -15  :12 */
    int dmax;        /* D_MAX    0  :11   size_d = dmax - dmin + 1 */
    int radius;      /* RADIUS   9  :21 */
    int d_lr;        /* D_LR     0  :24 */
    double eps;      /* EPS 6.5025  :23 (double literal in the reference) */
    float alpha;     /* ALPHA  0.9  :10 */
    float th_color;  /* TH_color 7  :14 */
    float th_grad;   /* TH_grad  2  :13 */
    double r_w;      /* R_W  0.299  :7 */
    double g_w;      /* G_W  0.587  :8 */
    double b_w;      /* B_W 0.0721  :9 */
    int guide_mode;  /* SB200_GUIDE_* */
    int box_mode;    /* SB200_BOX_*   (stage entry points; the fused gray pipeline ignores it; with guide_mode RGB,
                        SB200_BOX_SAT selects the staged colour path, which materialises every slice and is ~50x slower) */
} sb200_params;

void sb200_default_params(sb200_params* p);

/* ---- context ----------------------------------------------------------------------- */
/* replaces main.cu:44-48 (cudaGetDeviceProperties / cudaSetDevice(0)) */
int sb200_ctx_create(int device, sb200_ctx** out);
void sb200_ctx_destroy(sb200_ctx* ctx);
/* `stream` is a cudaStream_t (NULL = the legacy default stream).  Not owned. */
int sb200_ctx_set_stream(sb200_ctx* ctx, void* stream);
int sb200_ctx_synchronize(sb200_ctx* ctx);
/* message of the last failure on this context (or of ctx_create when ctx == NULL) */
const char* sb200_last_error(const sb200_ctx* ctx);
/* kernels launched by this context since creation (bench.py's gpu_launches) */
uint64_t sb200_launch_count(const sb200_ctx* ctx);
/* which fused RGB-guide kernel this context runs: 2 = k_fused_cvf_rgb (two warp roles), 3 = k_fused_cvf_rgb3 (four
 * roles + TMA operand ring; the default).  Results are bit-identical; the environment variable SB200_RGB_KERNEL=2|3, read by
 * sb200_ctx_create, overrides the default (A/B measurements). */
int sb200_ctx_rgb_kernel(const sb200_ctx* ctx);
/* which fused gray-guide kernel this context runs: 1 = k_fused_mma (horizontal window sums on the tensor cores; the
 * default), 0 = k_fused_cvf (warp-shuffle window sums).  Same contract, results equal within the parity tolerance (their
 * second-stage float sums are ordered differently).  SB200_GRAY_KERNEL=mma|shfl, read by sb200_ctx_create, overrides. */
int sb200_ctx_gray_kernel(const sb200_ctx* ctx);
int sb200_ctx_set_gray_kernel(sb200_ctx* ctx, int which);
const char* sb200_version(void);

/* ---- stage drop-ins, HOST pointers, blocking ----------------------------------------- */
/* rgb_to_grayscale.cuh:7  unsigned char* rgb_to_grayscale(h_rgb, n, channels, compare).
 * The reference returns a malloc'd buffer; here the caller provides gray[n]. channels >= 3. */
int sb200_rgb_to_grayscale(sb200_ctx* ctx, const sb200_params* p, const uint8_t* h_rgb, int n, int channels,
                           uint8_t* h_gray);
/* costVolume.cuh:7  compute_cost(i1, i2, cost, w1, w2, h1, h2, dmin, compare); size_d is the
 * reference's macro D_MAX-D_MIN+1, here p->dmax - p->dmin + 1; `dmin` stays an argument because
 * main.cu:80-82 passes D_MIN for the left view and -D_MAX for the right view. w1==w2, h1==h2. */
int sb200_compute_cost(sb200_ctx* ctx, const sb200_params* p, const uint8_t* i1, const uint8_t* i2, float* cost,
                       int w1, int w2, int h1, int h2, int dmin);
/* costVolume.cuh:16  x_derivativeOnGPU as a host-callable stage */
int sb200_x_derivative(sb200_ctx* ctx, const uint8_t* img, float* grad, int w, int h);
/* integral.cuh:3  integral(image, integral, width, height): float32 SAT, sequential add order */
int sb200_integral(sb200_ctx* ctx, const float* image, float* integral, int width, int height);
/* guidedFilter.cuh:23  computeBoxFilterOnGPU(image, integral, mean, w, h): 4-tap SAT lookup */
int sb200_box_filter_sat(sb200_ctx* ctx, const sb200_params* p, const float* integral, float* mean, int w, int h);
/* box mean of a float image in p->box_mode (SAT: the two calls above; SLIDING: double-accumulated) */
int sb200_box_filter(sb200_ctx* ctx, const sb200_params* p, const float* image, float* mean, int w, int h);
/* filter.cuh:12  filter(image, width, height, mean, var, cuda): guide mean (uchar) + variance.
 * Semantics are those of the live path (guidedFilter.cu:58-123), not filter.cu's dead tile kernel. */
int sb200_filter(sb200_ctx* ctx, const sb200_params* p, const uint8_t* image, int width, int height, uint8_t* mean,
                 float* var);
/* guidedFilter.cuh:7  compute_guided_filter(i, cost, filter_cost, disp_map, mean, w, h, size_d, dmin, compare):
 * guide statistics + per-slice guided filter + running WTA.  filter_cost / disp_map are IN/OUT
 * and must be pre-initialised by the caller (main.cu:112-120). */
int sb200_compute_guided_filter(sb200_ctx* ctx, const sb200_params* p, const uint8_t* i, const float* cost,
                                float* filter_cost, float* disp_map, uint8_t* mean, int w, int h, int size_d,
                                int dmin);
/* guidedFilter.cuh:8  dispSelectOnGPU(q, filter_cost, dmap, n, label): the live winner-take-all
 * (winner_take_all.cuh is commented out in the reference).  Update on best >= q. */
int sb200_winner_take_all(sb200_ctx* ctx, const float* q, float* filter_cost, float* dmap, int n, int label);
/* occlusion.cuh:8  detect_occlusion(dL, dR, dOcclusion, dmapl, dmapr, w, h); the two uchar maps
 * are passed through untouched by the reference (occlusion.cu:40-41,55-56) and may be NULL. */
int sb200_detect_occlusion(sb200_ctx* ctx, const sb200_params* p, float* disparityLeft, const float* disparityRight,
                           int dOcclusion, uint8_t* dmapl, uint8_t* dmapr, int w, int h);
/* occlusion.cuh:14  fill_occlusion(disparity, w, h, vMin) */
int sb200_fill_occlusion(sb200_ctx* ctx, float* disparity, int w, int h, float vMin);

/* ---- stage drop-ins, DEVICE pointers, asynchronous on the context's stream -------------- */
int sb200_rgb_to_grayscale_dev(sb200_ctx* ctx, const sb200_params* p, const uint8_t* d_rgb, int n, int channels,
                               uint8_t* d_gray);
int sb200_compute_cost_dev(sb200_ctx* ctx, const sb200_params* p, const uint8_t* d_i1, const uint8_t* d_i2,
                           float* d_cost, int w, int h, int dmin);
int sb200_integral_dev(sb200_ctx* ctx, const float* d_image, float* d_integral, int w, int h);
int sb200_box_filter_dev(sb200_ctx* ctx, const sb200_params* p, const float* d_image, float* d_mean, int w, int h);
int sb200_compute_guided_filter_dev(sb200_ctx* ctx, const sb200_params* p, const uint8_t* d_i, const float* d_cost,
                                    float* d_filter_cost, float* d_disp_map, uint8_t* d_mean, int w, int h,
                                    int size_d, int dmin);
int sb200_winner_take_all_dev(sb200_ctx* ctx, const float* d_q, float* d_filter_cost, float* d_dmap, int n,
                              int label);
int sb200_detect_occlusion_dev(sb200_ctx* ctx, const sb200_params* p, float* d_dL, const float* d_dR, int dOcclusion,
                               int w, int h);
int sb200_fill_occlusion_dev(sb200_ctx* ctx, float* d_disparity, int w, int h, float vMin);

/* ---- 8-bit visualisation (main.cu:13-35 write_mat, occlusion.cu:230-237 flToCh2OnGPU) ---------------------------
 * write_mat: out = (uchar)(int)((v - min) * 255 / (max - min)) with the reference's sequential min/max scan, whose
 * `else if` never considers a value that raises the running maximum as a minimum; the device version reproduces that
 * with a prefix-max, so a driver can download 8-bit images instead of float maps.  The reference writes the PNG itself
 * (stb); here the caller gets the 8-bit image. */
int sb200_write_mat(sb200_ctx* ctx, const float* h_mat, uint8_t* h_out, int w, int h);
int sb200_write_mat_dev(sb200_ctx* ctx, const float* d_mat, uint8_t* d_out, int w, int h);
/* flToCh2OnGPU: result = min(255, (uchar)(160 * (pix - vmin) / (vmax - vmin))) */
int sb200_fl_to_ch2_dev(sb200_ctx* ctx, const float* d_image, uint8_t* d_result, int vmin, int vmax, int len);

/* ---- fused pipeline (replaces main.cu:65-155 as one call) -------------------------------- */
/* Outputs; any pointer may be NULL.  All arrays are w*h (w*rows for a row strip, n_pairs*w*h for a batch). */
typedef struct sb200_outputs {
    float* disp_left;      /* dmapl  main.cu:133 : WTA labels of the left view, in [dmin, dmax]  */
    float* disp_right;     /* dmapr  main.cu:134 : WTA labels of the right view, in [-dmax, -dmin] */
    float* occlusion;      /* main.cu:150 : disp_left with occluded pixels set to dmin-100       */
    float* filled;         /* main.cu:155 : occlusion map after scan-line fill (vMin = dmin)     */
    float* best_left;      /* best_costl main.cu:133 : min over d of the filtered cost            */
    float* best_right;     /* best_costr main.cu:134                                              */
    uint8_t* gray_left;    /* I_l main.cu:65 */
    uint8_t* gray_right;   /* I_r main.cu:66 */
    uint8_t* mean_left;    /* mean1 main.cu:133 : (uchar)min((int)mean_I,255) debug image        */
    uint8_t* mean_right;   /* mean2 main.cu:134 */
    float* subpixel_left;  /* NOT in the reference: disp_left refined by sb200_subpixel_refine_dev's parabola,
                              pixels the L/R check marked keep their filled label.  Tensor-core fused kernels (either
                              guide); the call keeps the left view's filtered volume in the arena (size_d*w*h floats). */
} sb200_outputs;

/* Row-strip geometry for multi-GPU sharding: the images passed in are rows
 * [y0 - halo_top, y0 + rows + halo_bot) of a frame_h-row frame; outputs cover rows
 * [y0, y0 + rows).  Zero-initialise for a whole frame (rows = h, y0 = 0, frame_h = h). */
typedef struct sb200_strip {
    int y0;
    int rows;
    int halo_top;
    int halo_bot;
    int frame_h;
} sb200_strip;

/* Device pointers, asynchronous.  left/right: channels==1 (gray) or >=3 (interleaved RGB).
 * cost + guided-filter box sums + running argmin are one fused kernel per pair (both views);
 * L/R check + fill are a second kernel.  The D-deep volume is never materialised.
 * guide_mode RGB uses the colour guided filter (fused; box_mode SAT selects the staged colour path).
 * Parameter sets the fused kernels are not built for (radius != 9, alpha/thresholds without an exact
 * integer cost lattice) run through the stage kernels on the GPU instead (whole frames only). */
int sb200_pipeline_dev(sb200_ctx* ctx, const sb200_params* p, const uint8_t* d_left, const uint8_t* d_right,
                       int channels, int w, int h, const sb200_outputs* d_out);
/* same with HOST pointers, blocking (pageable or pinned; copies happen inside) */
int sb200_pipeline(sb200_ctx* ctx, const sb200_params* p, const uint8_t* h_left, const uint8_t* h_right, int channels,
                   int w, int h, const sb200_outputs* h_out);
/* n_pairs pairs with HOST pointers (left + i*w*h*channels, outputs + i*w*h), blocking, OVERLAPPED: the upload of pair
 * i+1 and the download of pair i-1 run on their own streams under the kernels of pair i.  Use page-locked host buffers
 * (cudaHostAlloc / cudaHostRegister); pageable buffers work but the driver stages their copies and the overlap is lost.
 * Replaces the reference's per-stage cudaMemcpy round trips (guidedFilter.cu:39-56, costVolume.cu:23-53). */
int sb200_pipeline_batch(sb200_ctx* ctx, const sb200_params* p, const uint8_t* h_left, const uint8_t* h_right, int channels,
                         int w, int h, int n_pairs, const sb200_outputs* h_out);
/* The same call with the four LABEL maps delivered as int16 (any pointer may be NULL): the reference keeps its labels in
 * floats (guidedFilter.cu:409), but they are integers in [dmin-100, dmax], and halving the bytes of the device->host copy
 * is what keeps an end-to-end batch at the device rate on a busy PCIe root.  h_other (may be NULL) carries any further
 * outputs (best costs, gray/mean images); a map must not be requested in both forms. */
typedef struct sb200_labels_i16 {
    int16_t* disp_left;
    int16_t* disp_right;
    int16_t* occlusion;
    int16_t* filled;
} sb200_labels_i16;
int sb200_pipeline_batch_i16(sb200_ctx* ctx, const sb200_params* p, const uint8_t* h_left, const uint8_t* h_right, int channels,
                             int w, int h, int n_pairs, const sb200_labels_i16* h_labels, const sb200_outputs* h_other);
/* n_pairs pairs of identical shape, contiguous: left + i*w*h*channels, outputs + i*w*h */
int sb200_pipeline_batch_dev(sb200_ctx* ctx, const sb200_params* p, const uint8_t* d_left, const uint8_t* d_right,
                             int channels, int w, int h, int n_pairs, const sb200_outputs* d_out);
/* one row strip of a taller frame; h = halo_top + rows + halo_bot rows are passed in */
int sb200_pipeline_strip_dev(sb200_ctx* ctx, const sb200_params* p, const uint8_t* d_left, const uint8_t* d_right,
                             int channels, int w, const sb200_strip* strip, const sb200_outputs* d_out);
/* Row strips across the GPUs of a box in ONE call per rank: `d_own_*` are this rank's rows [y0, y0+rows) of the frame
 * (device pointers); the 2*radius input rows each neighbouring rank holds are fetched with one grouped
 * ncclSend/ncclRecv on the context's stream, then the strip runs as sb200_pipeline_strip_dev does.  `nccl_comm` is
 * the caller's ncclComm_t (ranks hold consecutive strips in rank order, as sb200_strip_rows deals them out); libnccl
 * is loaded with dlopen on first use, the library does not link against it.  Asynchronous.  Outputs cover the own
 * rows.  Replaces the single-device assumption of main.cu:44-48 for frames too large for one GPU's time budget. */
int sb200_pipeline_strips_nccl(sb200_ctx* ctx, const sb200_params* p, void* nccl_comm, int rank, int world,
                               const uint8_t* d_own_left, const uint8_t* d_own_right, int channels, int w, int frame_h,
                               int y0, int rows, const sb200_outputs* d_out);
/* Which implementation sb200_pipeline* runs for these parameters on this context -- the fused kernels are built for the
 * reference's RADIUS 9 (SystemIncludes.h:9) and for cost parameters with an exact integer lattice (its ALPHA / TH_* are);
 * anything else still runs on the GPU, through the stage kernels: same results contract, whole frames only, and of the
 * order of a second per 1080p D=256 pair instead of milliseconds.  A caller that changes RADIUS can ask first.
 * Returns SB200_PATH_* (>= 0) or -SB200_ERR_* (e.g. -SB200_ERR_UNSUPPORTED: RGB guide with parameters no fused kernel
 * covers and box_mode not SAT). */
#define SB200_PATH_STAGED 0
#define SB200_PATH_FUSED_SHUFFLE 1
#define SB200_PATH_FUSED_TENSOR 2
int sb200_pipeline_path(const sb200_ctx* ctx, const sb200_params* p);
/* the balanced split sb200_pipeline_strips_nccl expects: rows [y0, y0+rows) of rank `rank` of `world` */
int sb200_strip_rows(int frame_h, int rank, int world, int* y0, int* rows);
/* halo rows each side that sb200_pipeline_strip_dev needs: 2*radius (two cascaded boxes) */
int sb200_strip_halo_rows(const sb200_params* p);

/* Fused view kernel alone (for measurement and for callers that only want one view):
 * guide/other are gray images on the device; best/disp are outputs. */
int sb200_view_disparity_dev(sb200_ctx* ctx, const sb200_params* p, const uint8_t* d_guide, const uint8_t* d_other,
                             int w, int h, int dmin, int size_d, float* d_best, float* d_disp, uint8_t* d_mean);
/* The same kernel, keeping what the reference drops after the winner-take-all: d_volume[k*w*h + y*w + x] = the filtered cost
 * q (compute_q, guidedFilter.cu:363-369) of disparity dmin + k, in the layout of the reference's cost volume
 * (costVolume.cu:178).  d_best/d_disp may be NULL.  Tensor-core gray kernel only (SB200_ERR_UNSUPPORTED otherwise). */
int sb200_view_volume_dev(sb200_ctx* ctx, const sb200_params* p, const uint8_t* d_guide, const uint8_t* d_other, int w, int h,
                          int dmin, int size_d, float* d_volume, float* d_best, float* d_disp);
/* Sub-pixel refinement, NOT in the reference.  With k = (int)disp[i] - dmin and q-, q0, q+ the volume at
 * k-1, k, k+1 of pixel i:  den = (q- - q0) + (q+ - q0);  out[i] = disp[i] + clamp(0.5 (q- - q+) / den, -0.5, 0.5) when
 * 0 < k < size_d-1 and den > 0, else disp[i]; float operations in exactly this order (the oracle's so_subpixel_refine agrees
 * bit for bit).  With d_occlusion != NULL, pixels the L/R check marked ((int)occlusion[i] < dmin, occlusion.cu:139)
 * get d_filled[i] (disp[i] when d_filled is NULL) instead. */
int sb200_subpixel_refine_dev(sb200_ctx* ctx, const float* d_volume, const float* d_disp, const float* d_occlusion,
                              const float* d_filled, float* d_out, int w, int h, int dmin, int size_d);
/* L/R check + fill alone: occlusion/filled outputs from the two label maps */
int sb200_lr_check_fill_dev(sb200_ctx* ctx, const sb200_params* p, const float* d_dL, const float* d_dR, int w, int h,
                            int dOcclusion, float vMin, float* d_occlusion, float* d_filled);

/* device timing of the last pipeline call's kernels, in milliseconds (needs
 * sb200_ctx_enable_timing(ctx,1) before the call; adds event records, no syncs) */
int sb200_ctx_enable_timing(sb200_ctx* ctx, int on);
int sb200_last_timing(sb200_ctx* ctx, float* ms_prep, float* ms_fused, float* ms_merge, float* ms_occl);
/* device time of the last sb200_pipeline_strips_nccl call's halo exchange (timing enabled) */
int sb200_last_exchange_ms(sb200_ctx* ctx, float* ms);

/* ---- Post-processing the reference stops short of -------------------------------------------------
 * NOT in the reference: its pipeline ends at fill_occlusion (occlusion.cu:111-132).  The IPOL pipeline it mirrors goes
 * on with a weighted median of the filled pixels; this entry provides it with a definition that is exact in integers,
 * so that the CUDA path and the oracle (oracle/stereo_oracle.c: so_weighted_median) agree bit for bit:
 *   for every pixel the L/R check marked ((int)occlusion[i] < dmin, the test of fill_occlusion, occlusion.cu:139):
 *     out[i] = the smallest label l with 2 * sum_{q in window, filled[q] <= l} W(i,q) >= sum_{q in window} W(i,q),
 *     window = [x-radius, x+radius] x [y-radius, y+radius] clipped to the image,
 *     W(i,q) = ws[|dy|][|dx|] * wc[|gray[i] - gray[q]|], integer tables
 *     ws = round(1024 exp(-(dx^2+dy^2)/sigma_space^2)), wc = round(1024 exp(-dI^2/sigma_color^2)) (computed in double);
 *   every other pixel: out[i] = filled[i].
 * Labels are the integral floats of the disparity maps, dmin <= label <= dmax.  radius <= 32. */
typedef struct sb200_wmedian_params {
    int radius;          /* 19 (IPOL) */
    float sigma_space;   /* 9 */
    float sigma_color;   /* 25.5 = 0.1 * 255 */
} sb200_wmedian_params;
void sb200_default_wmedian_params(sb200_wmedian_params* wp);
int sb200_weighted_median(sb200_ctx* ctx, const sb200_params* p, const sb200_wmedian_params* wp, const uint8_t* gray,
                          const float* occlusion, const float* filled, float* out, int w, int h);
int sb200_weighted_median_dev(sb200_ctx* ctx, const sb200_params* p, const sb200_wmedian_params* wp, const uint8_t* d_gray,
                              const float* d_occlusion, const float* d_filled, float* d_out, int w, int h);

#ifdef __cplusplus
}
#endif
#endif /* STEREO_B200_H */
