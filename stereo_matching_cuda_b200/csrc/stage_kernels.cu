// stage_kernels.cu -- stage-materialising kernels behind the per-stage drop-in entry points.
//
// These write every intermediate to HBM exactly where the reference's stage boundaries are,
// and follow its float32 arithmetic operation by operation (explicit __f*_rn intrinsics, so
// nvcc cannot contract a*b+c: the reference's goldens are reproduced by the non-contracted
// evaluation).  They are HBM-bound streaming kernels; the
// headline path is the fused kernel in fused_cvf.cu, which never materialises the volume.
#include <climits>

#include "common.cuh"

// ---------------------------------------------------------------------------------------
// rgb_to_grayscale.cu:14-23 (sumArraysOnGPU): double luma, truncation toward zero.
__global__ void k_rgb_to_gray(const uint8_t* __restrict__ rgb, uint8_t* __restrict__ gray, int n, int ch, double rw,
                              double gw, double bw) {
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n) return;
    size_t i = (size_t)ch * idx;
    double val = __dadd_rn(__dadd_rn(__dmul_rn(rw, (double)rgb[i]), __dmul_rn(gw, (double)rgb[i + 1])),
                           __dmul_rn(bw, (double)rgb[i + 2]));
    gray[idx] = (unsigned char)val;
}

int sbk_rgb_to_gray(sb200_ctx* ctx, const sb200_params* p, const uint8_t* rgb, int n, int ch, uint8_t* gray) {
    SB_LAUNCH(ctx, k_rgb_to_gray, sb_div_up(n, 256), 256, 0, rgb, gray, n, ch, p->r_w, p->g_w, p->b_w);
    return SB200_OK;
}

// costVolume.cu:358-381 (x_derivativeOnGPU): (left - right)/2 with one-sided (still halved) ends.
__global__ void k_x_derivative(const uint8_t* __restrict__ in, float* __restrict__ out, int w, int h) {
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    int y = blockIdx.y;
    if (x >= w) return;
    const uint8_t* row = in + (size_t)y * w;
    int xl = x - 1 >= 0 ? x - 1 : x;
    int xr = x + 1 < w ? x + 1 : x;
    if (w == 1) { xl = xr = 0; }
    out[(size_t)y * w + x] = 1.0f * (float)((int)row[xl] - (int)row[xr]) / 2;
}

int sbk_x_derivative(sb200_ctx* ctx, const uint8_t* img, float* grad, int w, int h) {
    dim3 grid(sb_div_up(w, 256), h);
    SB_LAUNCH(ctx, k_x_derivative, grid, 256, 0, img, grad, w, h);
    return SB200_OK;
}

// costVolume.cu:163-190 (costVolumOnGPU2); grid.y = slice so any size_d works (the
// reference's block (16,size_d) caps size_d at 64, costVolume.cu:39).
__global__ void k_cost_volume(const uint8_t* __restrict__ i1, const uint8_t* __restrict__ i2,
                              const float* __restrict__ g1, const float* __restrict__ g2, float* __restrict__ cost,
                              int w, int h, int dmin, float alpha, float th_color, float th_grad) {
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    int y = blockIdx.y;
    int k = blockIdx.z;
    if (x >= w) return;
    int d = dmin + k;
    size_t idx = (size_t)y * w + x;
    float oma = __fsub_rn(1.0f, alpha);
    float c = __fadd_rn(__fmul_rn(oma, th_color), __fmul_rn(alpha, th_grad));
    if (x + d < w && x + d >= 0) {
        float ci = fminf((float)abs((int)i1[idx] - (int)i2[idx + d]), th_color);
        float cg = fminf(fabsf(__fsub_rn(g1[idx], g2[idx + d])), th_grad);
        c = __fadd_rn(__fmul_rn(oma, ci), __fmul_rn(alpha, cg));
    }
    cost[(size_t)k * w * h + idx] = c;
}

int sbk_cost_volume(sb200_ctx* ctx, const sb200_params* p, const uint8_t* i1, const uint8_t* i2, const float* g1,
                    const float* g2, float* cost, int w, int h, int size_d, int dmin) {
    // grid.z is limited to 65535 and grid.y likewise: loop over slabs if ever needed
    if (h > 65535 || size_d > 65535) return sb_fail(ctx, SB200_ERR_UNSUPPORTED, "cost_volume: h or size_d > 65535");
    dim3 grid(sb_div_up(w, 256), h, size_d);
    SB_LAUNCH(ctx, k_cost_volume, grid, 256, 0, i1, i2, g1, g2, cost, w, h, dmin, p->alpha, p->th_color, p->th_grad);
    return SB200_OK;
}

// ---------------------------------------------------------------------------------------
// integral.cu:78-90 (rowSum): out[x] = in[x] + out[x-1], strictly sequential per row.  The
// reference runs one thread per row over global memory; here a warp owns 32 rows, stages
// 32x32 tiles through shared memory with coalesced loads/stores, and lane l walks row l of
// the tile -- the add ORDER per row is unchanged, so the result is bit-identical.
#define SCAN_WARPS 4
__global__ void k_row_scan(const float* __restrict__ in, float* __restrict__ out, int w, int h) {
    __shared__ float tile[SCAN_WARPS][32][33];
    int warp = threadIdx.y, lane = threadIdx.x;
    int row0 = (blockIdx.x * SCAN_WARPS + warp) * 32;
    if (row0 >= h) return;
    float carry = 0.0f;
    for (int c0 = 0; c0 < w; c0 += 32) {
        int x = c0 + lane;
#pragma unroll 8
        for (int r = 0; r < 32; r++) {
            int y = row0 + r;
            tile[warp][r][lane] = (y < h && x < w) ? in[(size_t)y * w + x] : 0.0f;
        }
        __syncwarp();
#pragma unroll 8
        for (int c = 0; c < 32; c++) {
            float v = tile[warp][lane][c];
            carry = (c0 == 0 && c == 0) ? v : __fadd_rn(v, carry);
            tile[warp][lane][c] = carry;
        }
        __syncwarp();
#pragma unroll 8
        for (int r = 0; r < 32; r++) {
            int y = row0 + r;
            if (y < h && x < w) out[(size_t)y * w + x] = tile[warp][r][lane];
        }
        __syncwarp();
    }
}

// integral.cu:121-131 (colSum): out[y] = in[y] + out[y-1] per column, one thread per column
// (coalesced across the warp); loads are batched 8 rows ahead to hide latency.
__global__ void k_col_scan(const float* __restrict__ in, float* __restrict__ out, int w, int h) {
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= w) return;
    float carry = in[x];
    out[x] = carry;
    int y = 1;
    for (; y + 8 <= h; y += 8) {
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; i++) v[i] = in[(size_t)(y + i) * w + x];
#pragma unroll
        for (int i = 0; i < 8; i++) {
            carry = __fadd_rn(v[i], carry);
            out[(size_t)(y + i) * w + x] = carry;
        }
    }
    for (; y < h; y++) {
        carry = __fadd_rn(in[(size_t)y * w + x], carry);
        out[(size_t)y * w + x] = carry;
    }
}

int sbk_integral(sb200_ctx* ctx, const float* in, float* tmp, float* out, int w, int h) {
    dim3 blk(32, SCAN_WARPS);
    SB_LAUNCH(ctx, k_row_scan, sb_div_up(h, 32 * SCAN_WARPS), blk, 0, in, tmp, w, h);
    SB_LAUNCH(ctx, k_col_scan, sb_div_up(w, 64), 64, 0, tmp, out, w, h);
    return SB200_OK;
}

// guidedFilter.cu:297-318 (computeBoxFilterOnGPU / computeMeanOnGPU): same tap order, IEEE divide.
__global__ void k_box_from_sat(const float* __restrict__ S, float* __restrict__ mean, int w, int h, int r) {
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    int idy = blockIdx.y * blockDim.y + threadIdx.y;
    if (idx >= w || idy >= h) return;
    int ymin = max(-1, idy - r - 1);
    int ymax = min(h - 1, idy + r);
    int xmin = max(-1, idx - r - 1);
    int xmax = min(w - 1, idx + r);
    float val = S[(size_t)ymax * w + xmax];
    if (xmin >= 0) val = __fsub_rn(val, S[(size_t)ymax * w + xmin]);
    if (ymin >= 0) val = __fsub_rn(val, S[(size_t)ymin * w + xmax]);
    if (xmin >= 0 && ymin >= 0) val = __fadd_rn(val, S[(size_t)ymin * w + xmin]);
    mean[(size_t)idy * w + idx] = __fdiv_rn(val, (float)((xmax - xmin) * (ymax - ymin)));
}

int sbk_box_from_sat(sb200_ctx* ctx, const float* sat, float* mean, int w, int h, int r) {
    dim3 blk(32, 8), grid(sb_div_up(w, 32), sb_div_up(h, 8));
    SB_LAUNCH(ctx, k_box_from_sat, grid, blk, 0, sat, mean, w, h, r);
    return SB200_OK;
}

// Sliding-window box mean with double accumulation (the "exact" box of the tests): a
// vertical running sum per column, then the clipped horizontal window per pixel.
__global__ void k_box_v_f64(const float* __restrict__ in, double* __restrict__ tmp, int w, int h, int r) {
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= w) return;
    double s = 0.0;
    for (int y = 0; y < min(r, h); y++) s += (double)in[(size_t)y * w + x];
    for (int y = 0; y < h; y++) {
        if (y + r < h) s += (double)in[(size_t)(y + r) * w + x];
        if (y - r - 1 >= 0) s -= (double)in[(size_t)(y - r - 1) * w + x];
        tmp[(size_t)y * w + x] = s;
    }
}
__global__ void k_box_h_f64(const double* __restrict__ tmp, float* __restrict__ mean, int w, int h, int r) {
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    int y = blockIdx.y;
    if (x >= w) return;
    int x0 = max(0, x - r), x1 = min(w - 1, x + r);
    int y0 = max(0, y - r), y1 = min(h - 1, y + r);
    double s = 0.0;
    for (int xx = x0; xx <= x1; xx++) s += tmp[(size_t)y * w + xx];
    double area = (double)(x1 - x0 + 1) * (double)(y1 - y0 + 1);
    mean[(size_t)y * w + x] = (float)(s / area);
}

int sbk_box_sliding(sb200_ctx* ctx, const float* in, double* tmp, float* mean, int w, int h, int r) {
    SB_LAUNCH(ctx, k_box_v_f64, sb_div_up(w, 64), 64, 0, in, tmp, w, h, r);
    dim3 grid(sb_div_up(w, 128), h);
    SB_LAUNCH(ctx, k_box_h_f64, grid, 128, 0, tmp, mean, w, h, r);
    return SB200_OK;
}

// ---------------------------------------------------------------------------------------
// elementwise helpers of guidedFilter.cu:442-474 and the per-slice kernels :345-369,:403-411
__global__ void k_u8_to_float(const uint8_t* __restrict__ in, float* __restrict__ out, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = 1.0f * (float)(int)in[i];
}
__global__ void k_float_to_u8(const float* __restrict__ in, uint8_t* __restrict__ out, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        int c = (int)in[i];
        out[i] = (c > 255) ? 255 : (unsigned char)c;
    }
}
__global__ void k_mul(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = __fmul_rn(a[i], b[i]);
}
__global__ void k_sub(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = __fsub_rn(a[i], b[i]);
}
// compute_ak_and_bk, guidedFilter.cu:345-354: c = 1.0f/(var+EPS) evaluated in double (EPS is a
// double literal) and narrowed; a = (mIp - mean*mp)*c; b = mp - mean*a
__global__ void k_ak_bk(const float* __restrict__ mean, const float* __restrict__ var, const float* __restrict__ mIp,
                        const float* __restrict__ mp, float* __restrict__ a, float* __restrict__ b, size_t n,
                        double eps) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float c = (float)(1.0 / ((double)var[i] + eps));
    float av = __fmul_rn(__fsub_rn(mIp[i], __fmul_rn(mean[i], mp[i])), c);
    a[i] = av;
    b[i] = __fsub_rn(mp[i], __fmul_rn(mean[i], av));
}
// compute_q, guidedFilter.cu:363-369
__global__ void k_q(const float* __restrict__ im, const float* __restrict__ a, const float* __restrict__ b,
                    float* __restrict__ q, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) q[i] = __fadd_rn(__fmul_rn(a[i], im[i]), b[i]);
}
// dispSelectOnGPU, guidedFilter.cu:403-411: the live winner-take-all; >= so the last slice wins ties
__global__ void k_disp_select(const float* __restrict__ q, float* __restrict__ best, float* __restrict__ dmap, size_t n,
                              int label) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        float qv = q[i];
        if (best[i] >= qv) {
            dmap[i] = (float)label;
            best[i] = qv;
        }
    }
}
__global__ void k_fill_f32(float* dst, float v, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = v;
}

#define EW_LAUNCH(kernel, n, ...) SB_LAUNCH(ctx, kernel, sb_div_up((long long)(n), 256), 256, 0, __VA_ARGS__)
int sbk_u8_to_float(sb200_ctx* ctx, const uint8_t* in, float* out, size_t n) { EW_LAUNCH(k_u8_to_float, n, in, out, n); return SB200_OK; }
int sbk_float_to_u8(sb200_ctx* ctx, const float* in, uint8_t* out, size_t n) { EW_LAUNCH(k_float_to_u8, n, in, out, n); return SB200_OK; }
int sbk_mul(sb200_ctx* ctx, const float* a, const float* b, float* out, size_t n) { EW_LAUNCH(k_mul, n, a, b, out, n); return SB200_OK; }
int sbk_sub(sb200_ctx* ctx, const float* a, const float* b, float* out, size_t n) { EW_LAUNCH(k_sub, n, a, b, out, n); return SB200_OK; }
int sbk_ak_bk(sb200_ctx* ctx, const float* mean, const float* var, const float* mIp, const float* mp, float* a,
              float* b, size_t n, double eps) {
    EW_LAUNCH(k_ak_bk, n, mean, var, mIp, mp, a, b, n, eps);
    return SB200_OK;
}
int sbk_q(sb200_ctx* ctx, const float* im, const float* a, const float* b, float* q, size_t n) { EW_LAUNCH(k_q, n, im, a, b, q, n); return SB200_OK; }
int sbk_disp_select(sb200_ctx* ctx, const float* q, float* best, float* dmap, size_t n, int label) {
    EW_LAUNCH(k_disp_select, n, q, best, dmap, n, label);
    return SB200_OK;
}
int sbk_fill_f32(sb200_ctx* ctx, float* dst, float v, size_t n) { EW_LAUNCH(k_fill_f32, n, dst, v, n); return SB200_OK; }

// ---------------------------------------------------------------------------------------
// occlusion.cu:3-15 (detect_occlusionOnGPU), in place on dL.  Reads of dL and writes of dL are
// per-pixel (no cross-pixel dependency), dR is read-only, so in place is race free.
__global__ void k_detect_occlusion(float* __restrict__ dL, const float* __restrict__ dR, int dOcc, int d_lr, int w,
                                   int h) {
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    int y = blockIdx.y;
    if (x >= w) return;
    size_t id = (size_t)y * w + x;
    int d = (int)dL[id];
    if (x + d < 0 || x + d >= w || fabsf((float)d + dR[id + d]) > (float)d_lr) dL[id] = (float)dOcc;
}

int sbk_detect_occlusion(sb200_ctx* ctx, float* dL, const float* dR, int dOcc, int d_lr, int w, int h) {
    dim3 grid(sb_div_up(w, 256), h);
    SB_LAUNCH(ctx, k_detect_occlusion, grid, 256, 0, dL, dR, dOcc, d_lr, w, h);
    return SB200_OK;
}

// L/R consistency check (occlusion.cu:3-15) + scan-line fill (occlusion.cu:134-176), one image row per block.
// The reference walks left and right from every occluded pixel; here the nearest valid value at-or-left and at-or-right
// of every pixel comes from a segmented scan of the row held in shared memory (a gather from the unfilled row, which
// SURVEY.md A.6 shows is what every schedule of the reference's in-place kernel computes):
//   1. the rows of dL and dR arrive with coalesced 128-bit loads; the check gathers dR[x+d] from shared memory;
//   2. every thread owns one contiguous segment of the row (odd length: conflict-free) and finds its right-most and
//      left-most valid pixel; ONE block scan (two barriers) turns those into the carries from all segments to the left
//      and to the right; a forward and a backward pass over the own segment then produce the filled values;
//   3. the filled row leaves with coalesced 128-bit stores.
// Six block barriers per row whatever its width (the previous version: three per 256-pixel tile and pass, 180 at 8K).
//   do_check = 0: `dL` is taken as the already-checked map.  occ_out / filled_out may be NULL; dL may alias filled_out.
#ifndef FILL_THREADS
#define FILL_THREADS 256
#endif
__global__ void __launch_bounds__(FILL_THREADS)
k_lr_check_fill(const float* dL, const float* __restrict__ dR, int w, int seg, int dOcc, int d_lr, float vMin,
                float* occ_out, float* filled_out, int do_check) {
    extern __shared__ __align__(16) float srow[];  // [0,w4): the checked row; [w4, 2 w4): dR, later the filled row
    const int w4 = (w + 3) & ~3;
    float* sv = srow;
    float* so = srow + w4;
    __shared__ int s_hi[FILL_THREADS / 32], s_lo[FILL_THREADS / 32];
    const int y = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const size_t base = (size_t)y * w;
    const bool vec = (w & 3) == 0;
    // 1. load
    if (vec) {
        const float4* l4 = reinterpret_cast<const float4*>(dL + base);
        const float4* r4 = reinterpret_cast<const float4*>(dR + base);
        for (int i = tid; i < w / 4; i += FILL_THREADS) {
            reinterpret_cast<float4*>(sv)[i] = l4[i];
            if (do_check) reinterpret_cast<float4*>(so)[i] = __ldg(r4 + i);
        }
    } else {
        for (int x = tid; x < w; x += FILL_THREADS) {
            sv[x] = dL[base + x];
            if (do_check) so[x] = dR[base + x];
        }
    }
    __syncthreads();
    if (do_check) {
        // a pixel's check reads its own dL and gathers dR[x+d]: the checked value can replace dL[x] at once
        for (int x = tid; x < w; x += FILL_THREADS) {
            float v = sv[x];
            const int d = (int)v;
            if (x + d < 0 || x + d >= w || fabsf((float)d + so[x + d]) > (float)d_lr) v = (float)dOcc;
            sv[x] = v;
        }
        __syncthreads();
        if (occ_out) {
            if (vec) {
                float4* o4 = reinterpret_cast<float4*>(occ_out + base);
                for (int i = tid; i < w / 4; i += FILL_THREADS) o4[i] = reinterpret_cast<float4*>(sv)[i];
            } else {
                for (int x = tid; x < w; x += FILL_THREADS) occ_out[base + x] = sv[x];
            }
        }
    }
    if (!filled_out) return;
    // 2. segmented scan: thread t owns [t*seg, min(w, (t+1)*seg))
    const int x0 = tid * seg, x1 = min(w, x0 + seg);
    int hi = -1, lo = INT_MAX;  // right-most / left-most valid pixel of the segment
    for (int x = x0; x < x1; x++)
        if (sv[x] >= vMin) {
            hi = x;
            lo = min(lo, x);
        }
    // exclusive scans over the threads: max of hi from the left, min of lo from the right
    int ph = hi, pl = lo;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const int a = __shfl_up_sync(0xffffffffu, ph, off), b2 = __shfl_down_sync(0xffffffffu, pl, off);
        if (lane >= off) ph = max(ph, a);
        if (lane + off < 32) pl = min(pl, b2);
    }
    if (lane == 31) s_hi[wid] = ph;
    if (lane == 0) s_lo[wid] = pl;
    __syncthreads();
    int carry_l = -1, carry_r = INT_MAX;
    for (int k = 0; k < FILL_THREADS / 32; k++) {
        if (k < wid) carry_l = max(carry_l, s_hi[k]);
        if (k > wid) carry_r = min(carry_r, s_lo[k]);
    }
    {   // inclusive -> exclusive inside the warp
        const int a = __shfl_up_sync(0xffffffffu, ph, 1), b2 = __shfl_down_sync(0xffffffffu, pl, 1);
        if (lane > 0) carry_l = max(carry_l, a);
        if (lane < 31) carry_r = min(carry_r, b2);
    }
    const float left0 = carry_l >= 0 ? sv[carry_l] : vMin;
    const float right0 = carry_r != INT_MAX ? sv[carry_r] : vMin;
    __syncthreads();  // `so` (dR) is free: every thread is past the check
    float cur = left0;
    for (int x = x0; x < x1; x++) {  // nearest valid value at-or-left
        const float v = sv[x];
        if (v >= vMin) cur = v;
        so[x] = cur;
    }
    cur = right0;
    for (int x = x1 - 1; x >= x0; x--) {  // nearest valid value at-or-right, and the fill
        const float v = sv[x];
        if (v >= vMin) cur = v;
        const int dX = (int)v;  // occlusion.cu:140-142: the self test uses the truncated int
        so[x] = ((float)dX >= vMin) ? v : fmaxf(so[x], cur);
    }
    __syncthreads();
    // 3. store
    if (vec) {
        float4* o4 = reinterpret_cast<float4*>(filled_out + base);
        for (int i = tid; i < w / 4; i += FILL_THREADS) o4[i] = reinterpret_cast<float4*>(so)[i];
    } else {
        for (int x = tid; x < w; x += FILL_THREADS) filled_out[base + x] = so[x];
    }
}

static int launch_lr(sb200_ctx* ctx, const float* dL, const float* dR, int w, int h, int dOcc, int d_lr, float vMin,
                     float* occ, float* filled, int do_check) {
    const size_t smem = (size_t)((w + 3) & ~3) * 8;
    if (smem > ctx->smem_optin)
        return sb_fail(ctx, SB200_ERR_UNSUPPORTED, "row of %d pixels does not fit shared memory", w);
    if (smem > 48 * 1024)
        SB_CUDA(ctx, cudaFuncSetAttribute(k_lr_check_fill, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int seg = (w + FILL_THREADS - 1) / FILL_THREADS;
    seg |= 1;  // odd segment length: the threads' strided shared-memory accesses fall on different banks
    SB_LAUNCH(ctx, k_lr_check_fill, h, FILL_THREADS, smem, dL, dR, w, seg, dOcc, d_lr, vMin, occ, filled, do_check);
    return SB200_OK;
}

// fill_occlusion (occlusion.cu:111-132) in place: every block reads its whole row into shared
// memory before writing any of it, and rows are independent.
int sbk_fill_occlusion(sb200_ctx* ctx, float* disp, int w, int h, float vMin) {
    return launch_lr(ctx, disp, nullptr, w, h, 0, 0, vMin, nullptr, disp, 0);
}

int sbk_lr_check_fill(sb200_ctx* ctx, const float* dL, const float* dR, int w, int h, int dOcc, int d_lr, float vMin,
                      float* occ, float* filled) {
    return launch_lr(ctx, dL, dR, w, h, dOcc, d_lr, vMin, occ, filled, 1);
}

// ---------------------------------------------------------------------------------------
// Weighted median of the pixels the L/R check marked (beyond the reference, defined in stereo_b200.h).
// One warp per marked pixel: the window's integer weights go into a per-warp histogram over the labels in shared
// memory (integer atomics: the sums do not depend on the order), then the warp finds the first label whose cumulated
// weight reaches half of the total.  Unmarked pixels are copied by their warp's lane 0.
#define WMED_WARPS 8
__global__ void __launch_bounds__(32 * WMED_WARPS) k_weighted_median(const uint8_t* __restrict__ gray, const float* __restrict__ occ,
                                                                      const float* __restrict__ filled, float* __restrict__ out,
                                                                      int w, int h, int dmin, int size_d, int radius,
                                                                      const unsigned* __restrict__ ws, const unsigned* __restrict__ wc) {
    extern __shared__ unsigned wmed_hist[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    unsigned* hist = wmed_hist + (size_t)wid * size_d;
    const long long n = (long long)w * h;
    for (long long i = (long long)blockIdx.x * WMED_WARPS + wid; i < n; i += (long long)gridDim.x * WMED_WARPS) {
        const float fo = filled[i];
        if (!((float)(int)occ[i] < (float)dmin)) {  // not marked (the test of fill_occlusion, occlusion.cu:139)
            if (lane == 0) out[i] = fo;
            continue;
        }
        const int y = (int)(i / w), x = (int)(i - (long long)y * w);
        for (int b = lane; b < size_d; b += 32) hist[b] = 0u;
        __syncwarp();
        const int x0 = max(0, x - radius), x1 = min(w - 1, x + radius), y0 = max(0, y - radius), y1 = min(h - 1, y + radius);
        const int ww = x1 - x0 + 1, taps = ww * (y1 - y0 + 1);
        const int gi = gray[i];
        unsigned total = 0;
        for (int t = lane; t < taps; t += 32) {
            const int ty = t / ww, tx = t - ty * ww;
            const int qx = x0 + tx, qy = y0 + ty;
            const size_t q = (size_t)qy * w + qx;
            const unsigned wgt = ws[abs(qy - y) * (radius + 1) + abs(qx - x)] * wc[abs(gi - (int)gray[q])];
            int b = (int)filled[q] - dmin;
            b = min(max(b, 0), size_d - 1);
            atomicAdd(&hist[b], wgt);
            total += wgt;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) total += __shfl_xor_sync(0xffffffffu, total, o);
        __syncwarp();
        // lane l owns bins [l*per, (l+1)*per): its sum, an exclusive warp scan of the sums, then a walk inside the chunk
        const int per = (size_d + 31) / 32;
        unsigned mine = 0;
        for (int b = lane * per; b < min(size_d, (lane + 1) * per); b++) mine += hist[b];
        unsigned incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        const unsigned long long half2 = total;  // compare 2 * cum >= total in 64 bits
        const bool here = 2ull * incl >= half2;
        const unsigned ballot = __ballot_sync(0xffffffffu, here);
        const int owner = ballot ? __ffs(ballot) - 1 : 31;
        if (lane == owner) {
            unsigned cum = incl - mine;
            int b = lane * per;
            const int bend = min(size_d, (lane + 1) * per);
            for (; b < bend; b++) {
                cum += hist[b];
                if (2ull * cum >= half2) break;
            }
            if (b >= size_d) b = size_d - 1;
            out[i] = (float)(dmin + b);
        }
        __syncwarp();
    }
}

int sbk_weighted_median(sb200_ctx* ctx, const uint8_t* gray, const float* occ, const float* filled, float* out, int w, int h,
                        int dmin, int size_d, int radius, const unsigned* ws, const unsigned* wc) {
    const size_t smem = (size_t)WMED_WARPS * size_d * sizeof(unsigned);
    if (smem > 48 * 1024) return sb_fail(ctx, SB200_ERR_UNSUPPORTED, "weighted median: more than %d labels", 48 * 1024 / 4 / WMED_WARPS);
    const long long n = (long long)w * h;
    const int blocks = (int)min((long long)ctx->sm_count * 32, (n + WMED_WARPS - 1) / WMED_WARPS);
    SB_LAUNCH(ctx, k_weighted_median, blocks, 32 * WMED_WARPS, smem, gray, occ, filled, out, w, h, dmin, size_d, radius, ws, wc);
    return SB200_OK;
}

// ---------------------------------------------------------------------------------------
// Label maps as int16 for the host-pointer batch entry: the reference keeps its labels in floats
// (guidedFilter.cu:409 `dmap[i] = (float)label`), but they are integers of a few hundred at most, and the device->host copy
// of four float maps is what an end-to-end batch spends its PCIe time on.  (short)float is exact for them.
struct LabelMaps {
    const float* src[4];
    int16_t* dst[4];
};
__global__ void k_labels_i16(const LabelMaps M, size_t n) {
    const float* __restrict__ s = M.src[blockIdx.y];
    int16_t* __restrict__ d = M.dst[blockIdx.y];
    if (!s || !d) return;
    const size_t i4 = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (i4 + 3 < n) {
        const float4 v = *reinterpret_cast<const float4*>(s + i4);
        short4 o;
        o.x = (short)v.x;
        o.y = (short)v.y;
        o.z = (short)v.z;
        o.w = (short)v.w;
        *reinterpret_cast<short4*>(d + i4) = o;
    } else {
        for (size_t i = i4; i < n; i++) d[i] = (int16_t)s[i];
    }
}
int sbk_labels_i16(sb200_ctx* ctx, const float* const src[4], int16_t* const dst[4], size_t n) {
    LabelMaps M;
    for (int k = 0; k < 4; k++) {
        M.src[k] = src[k];
        M.dst[k] = dst[k];
    }
    SB_LAUNCH(ctx, k_labels_i16, dim3((unsigned)sb_div_up((long long)((n + 3) / 4), 256), 4), 256, 0, M, n);
    return SB200_OK;
}

// ---------------------------------------------------------------------------------------
// Sub-pixel refinement (beyond the reference): the vertex of the parabola through the filtered costs at
// label-1, label, label+1 of the kept volume.  Float operations in a fixed order (no contraction), so the oracle's
// so_subpixel_refine agrees bit for bit on the same volume.
__global__ void k_subpixel(const float* __restrict__ vol, const float* __restrict__ disp, const float* __restrict__ occ,
                           const float* __restrict__ filled, float* __restrict__ out, size_t n, int dmin, int size_d) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float d = disp[i];
    if (occ && (int)occ[i] < dmin) {  // marked by the L/R check (the test of fill_occlusion, occlusion.cu:139): keeps its filled label
        out[i] = filled ? filled[i] : d;
        return;
    }
    const int k = (int)d - dmin;
    float r = d;
    if (k > 0 && k < size_d - 1) {
        const float qm = vol[(size_t)(k - 1) * n + i], q0 = vol[(size_t)k * n + i], qp = vol[(size_t)(k + 1) * n + i];
        const float den = __fadd_rn(__fsub_rn(qm, q0), __fsub_rn(qp, q0));
        if (den > 0.0f) {
            float t = __fdiv_rn(__fmul_rn(0.5f, __fsub_rn(qm, qp)), den);
            t = fminf(0.5f, fmaxf(-0.5f, t));
            r = __fadd_rn(d, t);
        }
    }
    out[i] = r;
}
int sbk_subpixel(sb200_ctx* ctx, const float* vol, const float* disp, const float* occ, const float* filled, float* out,
                 size_t n, int dmin, int size_d) {
    EW_LAUNCH(k_subpixel, n, vol, disp, occ, filled, out, n, dmin, size_d);
    return SB200_OK;
}

// ---------------------------------------------------------------------------------------
// 8-bit visualisation on the device: write_mat's normalisation (main.cu:13-35) and flToCh2OnGPU
// (occlusion.cu:230-237), so that a driver downloads 8-bit images instead of float maps.
// write_mat scans sequentially with `if (v > max) max = v; else if (v <= min) min = v;`: a value that RAISES the running
// maximum is never a candidate for the minimum.  In parallel: max = the global maximum; min = the minimum over the
// elements that are not strict prefix maxima (element i with v_i > max(init, v_0..v_{i-1})), found with a block-wise
// prefix-max: per-block maxima, their exclusive scan, then a block scan inside every block.
#define WM_BLOCK 1024
__device__ __forceinline__ int wm_key(float v) {  // order-preserving float -> int
    int i = __float_as_int(v);
    return i >= 0 ? i : i ^ 0x7FFFFFFF;
}
__device__ __forceinline__ float wm_unkey(int k) { return __int_as_float(k >= 0 ? k : k ^ 0x7FFFFFFF); }
__device__ __forceinline__ float wm_block_scan_max(float v, float* swarp, float* total) {  // inclusive, 1024 threads
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        float n = __shfl_up_sync(0xffffffffu, v, off);
        if (lane >= off) v = fmaxf(v, n);
    }
    if (lane == 31) swarp[wid] = v;
    __syncthreads();
    if (wid == 0) {
        float t = swarp[lane];
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            float n = __shfl_up_sync(0xffffffffu, t, off);
            if (lane >= off) t = fmaxf(t, n);
        }
        swarp[lane] = t;
    }
    __syncthreads();
    if (wid > 0) v = fmaxf(v, swarp[wid - 1]);
    *total = swarp[31];
    return v;
}
__global__ void __launch_bounds__(WM_BLOCK) k_wm_blockmax(const float* __restrict__ mat, size_t n, float* __restrict__ bmax) {
    __shared__ float swarp[32];
    const size_t i = (size_t)blockIdx.x * WM_BLOCK + threadIdx.x;
    float tot;
    wm_block_scan_max(i < n ? mat[i] : -INFINITY, swarp, &tot);
    if (threadIdx.x == 0) bmax[blockIdx.x] = tot;
}
// one block: bpre[b] = max(init, bmax[0..b-1]); mm[0] = key(global max), mm[1] = key(min init)
__global__ void __launch_bounds__(WM_BLOCK) k_wm_scan(const float* __restrict__ bmax, int nb, float* __restrict__ bpre,
                                                       int* __restrict__ mm) {
    __shared__ float swarp[32];
    float carry = -150000000.0f;
    for (int b0 = 0; b0 < nb; b0 += WM_BLOCK) {
        const int b = b0 + threadIdx.x;
        const float v = b < nb ? bmax[b] : -INFINITY;
        float tot;
        const float inc = wm_block_scan_max(v, swarp, &tot);
        // exclusive = max of the inclusive value of the previous thread and the carry
        __shared__ float sinc[WM_BLOCK];
        sinc[threadIdx.x] = inc;
        __syncthreads();
        const float prev = threadIdx.x > 0 ? sinc[threadIdx.x - 1] : -INFINITY;
        if (b < nb) bpre[b] = fmaxf(carry, prev);
        carry = fmaxf(carry, tot);
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        mm[0] = wm_key(carry);
        mm[1] = wm_key(150000000.0f);
    }
}
__global__ void __launch_bounds__(WM_BLOCK) k_wm_min(const float* __restrict__ mat, size_t n, const float* __restrict__ bpre,
                                                      int* __restrict__ mm) {
    __shared__ float swarp[32];
    __shared__ float sinc[WM_BLOCK];
    __shared__ int smin;
    if (threadIdx.x == 0) smin = 0x7FFFFFFF;
    const size_t i = (size_t)blockIdx.x * WM_BLOCK + threadIdx.x;
    const float v = i < n ? mat[i] : -INFINITY;
    float tot;
    sinc[threadIdx.x] = wm_block_scan_max(v, swarp, &tot);
    __syncthreads();
    const float before = fmaxf(bpre[blockIdx.x], threadIdx.x > 0 ? sinc[threadIdx.x - 1] : -INFINITY);
    int key = 0x7FFFFFFF;
    if (i < n && !(v > before)) key = wm_key(v);  // not a strict prefix maximum: a candidate for the minimum
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) key = min(key, __shfl_xor_sync(0xffffffffu, key, off));
    if ((threadIdx.x & 31) == 0 && key != 0x7FFFFFFF) atomicMin(&smin, key);
    __syncthreads();
    if (threadIdx.x == 0 && smin != 0x7FFFFFFF) atomicMin(&mm[1], smin);
}
__global__ void k_wm_scale(const float* __restrict__ mat, size_t n, const int* __restrict__ mm, uint8_t* __restrict__ out) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float mx = wm_unkey(mm[0]), mn = wm_unkey(mm[1]);
    const int c = (int)__fdiv_rn(__fmul_rn(__fsub_rn(mat[i], mn), 255.0f), __fsub_rn(mx, mn));
    out[i] = (unsigned char)c;
}
// scratch: (nb + nb) floats + 2 ints, nb = ceil(n / 1024)
int sbk_write_mat(sb200_ctx* ctx, const float* mat, uint8_t* out, size_t n, float* scratch) {
    const int nb = sb_div_up((long long)n, WM_BLOCK);
    float* bmax = scratch;
    float* bpre = scratch + nb;
    int* mm = reinterpret_cast<int*>(scratch + 2 * (size_t)nb);
    SB_LAUNCH(ctx, k_wm_blockmax, nb, WM_BLOCK, 0, mat, n, bmax);
    SB_LAUNCH(ctx, k_wm_scan, 1, WM_BLOCK, 0, bmax, nb, bpre, mm);
    SB_LAUNCH(ctx, k_wm_min, nb, WM_BLOCK, 0, mat, n, bpre, mm);
    SB_LAUNCH(ctx, k_wm_scale, sb_div_up((long long)n, 256), 256, 0, mat, n, mm, out);
    return SB200_OK;
}
// flToCh2OnGPU, occlusion.cu:230-237
__global__ void k_fl_to_ch2(const float* __restrict__ image, uint8_t* __restrict__ result, int mn, int mx, size_t len) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= len) return;
    const float pix = image[i];
    const float c = __fdiv_rn(__fmul_rn(160.0f, __fsub_rn(pix, (float)mn)), (float)(mx - mn));
    result[i] = (c > 255.0f) ? 255 : (unsigned char)c;
}
int sbk_fl_to_ch2(sb200_ctx* ctx, const float* image, uint8_t* result, int mn, int mx, size_t len) {
    EW_LAUNCH(k_fl_to_ch2, len, image, result, mn, mx, len);
    return SB200_OK;
}

// ---------------------------------------------------------------------------------------
// RGB guide (not in the reference, whose guide is always gray): staged kernels.
__global__ void k_rgb_split(const uint8_t* __restrict__ rgb, int ch, float* __restrict__ r, float* __restrict__ g,
                            float* __restrict__ b, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    r[i] = (float)rgb[i * ch];
    g[i] = (float)rgb[i * ch + 1];
    b[i] = (float)rgb[i * ch + 2];
}
// cov[k] holds box(I_a I_b) on entry; M[k] = ((Sigma + eps U)^-1)[k] on exit, order xx xy xz yy yz zz
__global__ void k_rgb_inverse(const float* __restrict__ mr, const float* __restrict__ mg, const float* __restrict__ mb,
                              float* __restrict__ c0, float* __restrict__ c1, float* __restrict__ c2,
                              float* __restrict__ c3, float* __restrict__ c4, float* __restrict__ c5, size_t n,
                              double eps) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float mx = mr[i], my = mg[i], mz = mb[i];
    double xx = (double)__fsub_rn(c0[i], __fmul_rn(mx, mx)) + eps, xy = __fsub_rn(c1[i], __fmul_rn(mx, my));
    double xz = __fsub_rn(c2[i], __fmul_rn(mx, mz)), yy = (double)__fsub_rn(c3[i], __fmul_rn(my, my)) + eps;
    double yz = __fsub_rn(c4[i], __fmul_rn(my, mz)), zz = (double)__fsub_rn(c5[i], __fmul_rn(mz, mz)) + eps;
    double a00 = yy * zz - yz * yz, a01 = xz * yz - xy * zz, a02 = xy * yz - xz * yy;
    double a11 = xx * zz - xz * xz, a12 = xy * xz - xx * yz, a22 = xx * yy - xy * xy;
    double id = 1.0 / (xx * a00 + xy * a01 + xz * a02);
    c0[i] = (float)(a00 * id); c1[i] = (float)(a01 * id); c2[i] = (float)(a02 * id);
    c3[i] = (float)(a11 * id); c4[i] = (float)(a12 * id); c5[i] = (float)(a22 * id);
}
struct RgbPlanes {
    const float* mu[3];
    const float* M[6];
    const float* mIp[3];
    const float* mp;
    float* a[3];
    float* b;
};
__global__ void k_rgb_ab(RgbPlanes P, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float mp = P.mp[i], mx = P.mu[0][i], my = P.mu[1][i], mz = P.mu[2][i];
    float cx = __fsub_rn(P.mIp[0][i], __fmul_rn(mx, mp));
    float cy = __fsub_rn(P.mIp[1][i], __fmul_rn(my, mp));
    float cz = __fsub_rn(P.mIp[2][i], __fmul_rn(mz, mp));
    float ax = __fadd_rn(__fadd_rn(__fmul_rn(P.M[0][i], cx), __fmul_rn(P.M[1][i], cy)), __fmul_rn(P.M[2][i], cz));
    float ay = __fadd_rn(__fadd_rn(__fmul_rn(P.M[1][i], cx), __fmul_rn(P.M[3][i], cy)), __fmul_rn(P.M[4][i], cz));
    float az = __fadd_rn(__fadd_rn(__fmul_rn(P.M[2][i], cx), __fmul_rn(P.M[4][i], cy)), __fmul_rn(P.M[5][i], cz));
    P.a[0][i] = ax; P.a[1][i] = ay; P.a[2][i] = az;
    P.b[i] = __fsub_rn(mp, __fadd_rn(__fadd_rn(__fmul_rn(ax, mx), __fmul_rn(ay, my)), __fmul_rn(az, mz)));
}
__global__ void k_rgb_q(const float* __restrict__ mb, const float* __restrict__ ma0, const float* __restrict__ ma1,
                        const float* __restrict__ ma2, const float* __restrict__ r, const float* __restrict__ g,
                        const float* __restrict__ b, float* __restrict__ q, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float v = mb[i];
    v = __fadd_rn(v, __fmul_rn(ma0[i], r[i]));
    v = __fadd_rn(v, __fmul_rn(ma1[i], g[i]));
    v = __fadd_rn(v, __fmul_rn(ma2[i], b[i]));
    q[i] = v;
}
int sbk_rgb_split(sb200_ctx* ctx, const uint8_t* rgb, int ch, float* r, float* g, float* b, size_t n) {
    EW_LAUNCH(k_rgb_split, n, rgb, ch, r, g, b, n);
    return SB200_OK;
}
int sbk_rgb_inverse(sb200_ctx* ctx, float* const mu[3], float* const cov[6], size_t n, double eps) {
    EW_LAUNCH(k_rgb_inverse, n, mu[0], mu[1], mu[2], cov[0], cov[1], cov[2], cov[3], cov[4], cov[5], n, eps);
    return SB200_OK;
}
int sbk_rgb_ab(sb200_ctx* ctx, float* const mu[3], float* const M[6], float* const mIp[3], const float* mp,
               float* const a[3], float* b, size_t n) {
    RgbPlanes P;
    for (int c = 0; c < 3; c++) { P.mu[c] = mu[c]; P.mIp[c] = mIp[c]; P.a[c] = a[c]; }
    for (int k = 0; k < 6; k++) P.M[k] = M[k];
    P.mp = mp;
    P.b = b;
    EW_LAUNCH(k_rgb_ab, n, P, n);
    return SB200_OK;
}
int sbk_rgb_q(sb200_ctx* ctx, const float* mb, float* const ma[3], float* const I[3], float* q, size_t n) {
    EW_LAUNCH(k_rgb_q, n, mb, ma[0], ma[1], ma[2], I[0], I[1], I[2], q, n);
    return SB200_OK;
}
