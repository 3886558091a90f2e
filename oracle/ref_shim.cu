// ref_shim.cu -- extern "C" doorway into the UNMODIFIED reference sources.
//
// TEST INFRASTRUCTURE ONLY (see stereo_oracle.c header).  This file contains none of
// the reference's code: it #includes the reference's own headers from where they lie
// (-I/root/reference/stereo_matching_cuda) and is linked against objects compiled by
// oracle/Makefile directly from /root/reference/stereo_matching_cuda/*.cu.  The output
// (oracle/_ref/libref.so) is git-ignored but travels to the GPU box.
//
// Two families of entry points:
//   ref_*_cpu  call the reference's CPU twins (run anywhere, no GPU needed)
//   ref_*_gpu  call the reference's own host stage functions, which cudaMalloc/launch
//              their sm_100a-recompiled kernels (GPU box only; size_d is fixed to
//              D_MAX-D_MIN+1 = 16 by the macros in SystemIncludes.h:11-12)
#include "costVolume.cuh"
#include "guidedFilter.cuh"
#include "integral.cuh"
#include "occlusion.cuh"
#include "rgb_to_grayscale.cuh"

#include <cstdint>
#include <cstring>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

extern "C" {

int ref_size_d_macro() { return D_MAX - D_MIN + 1; }
int ref_dmin_macro() { return D_MIN; }
int ref_max_threads() {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

// ---- CPU twins -------------------------------------------------------------------
void ref_rgb_to_gray_cpu(unsigned char* rgb, unsigned char* gray, int n, int ch) { sumArraysOnHost(rgb, gray, n, ch); }
void ref_x_derivative_cpu(unsigned char* in, float* out, int w, int h) { x_derivativeOnCpu(in, out, w, h); }
void ref_cost_volume_cpu(unsigned char* i1, unsigned char* i2, float* cost, int w, int h, int size_d, int dmin) {
    std::vector<float> g1((size_t)w * h), g2((size_t)w * h);
    x_derivativeOnCpu(i1, g1.data(), w, h);
    x_derivativeOnCpu(i2, g2.data(), w, h);
    compute_costVolumeOnCpu(i1, i2, cost, g1.data(), g2.data(), w, w, h, h, size_d, dmin);
}
void ref_integral_cpu(float* in, float* out, int w, int h) { integralOnCPU(in, out, w, h); }
void ref_box_filter_cpu(float* image, float* integral, float* mean, int w, int h) {
    computeBoxFilterOnCPU(image, integral, mean, w, h);
}
void ref_disp_select_cpu(float* q, float* best, float* dmap, int n, int label) { dispSelectOnCPU(q, best, dmap, n, label); }
// guided_filter_onCpu (guidedFilter.cu:540-653): best-cost output is usable, dmap is not (its :622 bug)
void ref_guided_filter_cpu(unsigned char* im, float* cost, float* best, float* dmap, unsigned char* mean, int w, int h,
                           int size_d, int dmin) {
    guided_filter_onCpu(im, cost, best, dmap, mean, w, h, size_d, dmin);
}
void ref_fill_occlusion_cpu(float* disp, int w, int h, float vMin) { fill_occlusionOnCPU(disp, w, h, vMin); }
// detect_occlusionOnCPU reads dR[x+d] before its range test (occlusion.cu:98-99): give it slack on both sides.
void ref_detect_occlusion_cpu(float* dL, float* dR, int dOcclusion, int w, int h, int slack) {
    size_t n = (size_t)w * h;
    std::vector<float> padded(n + 2 * (size_t)slack, 0.0f);
    memcpy(padded.data() + slack, dR, n * sizeof(float));
    detect_occlusionOnCPU(dL, padded.data() + slack, dOcclusion, w, h);
}

// One view, chained from the reference's CPU functions in main.cu order:
// guide statistics as guided_filter_onCpu :558-567, per slice the body of :588-627 with
// dispSelectOnCPU for the labels (because :622 writes a wrong dmap).  The two lines that
// are inline code in the reference (a_k, b_k at :608-611) are restated here.  Cost slices
// are produced one at a time by compute_costVolumeOnCpu(size_d=1) so D*n floats are never
// held.  Slices are independent, so `nthreads` OpenMP threads each run whole slices with
// the reference's functions; WTA is applied in slice order.
void ref_view_disparity_cpu(unsigned char* guide, unsigned char* other, float* best, float* dmap,
                            unsigned char* mean, int w, int h, int size_d, int dmin, int nthreads) {
    const int n = w * h;
    std::vector<float> im(n), imXim(n), int_im(n), int_imXim(n), mean_im(n), mean_imXim(n), mean_im2(n), var_im(n);
    std::vector<float> g1(n), g2(n);
    chToFlOnCPU(guide, im.data(), n);
    pixelMultOnCPU(im.data(), im.data(), imXim.data(), n);
    integralOnCPU(im.data(), int_im.data(), w, h);
    integralOnCPU(imXim.data(), int_imXim.data(), w, h);
    computeBoxFilterOnCPU(im.data(), int_im.data(), mean_im.data(), w, h);
    computeBoxFilterOnCPU(imXim.data(), int_imXim.data(), mean_imXim.data(), w, h);
    pixelMultOnCPU(mean_im.data(), mean_im.data(), mean_im2.data(), n);
    pixelSousOnCPU(mean_imXim.data(), mean_im2.data(), var_im.data(), n);
    if (mean) flToChOnCPU(mean_im.data(), mean, n);
    x_derivativeOnCpu(guide, g1.data(), w, h);
    x_derivativeOnCpu(other, g2.data(), w, h);
    int nt = nthreads > 0 ? nthreads : ref_max_threads();
    if (nt > size_d) nt = size_d;
    if (nt < 1) nt = 1;
    std::vector<float> qs((size_t)nt * n);
    for (int s0 = 0; s0 < size_d; s0 += nt) {
        int cnt = size_d - s0 < nt ? size_d - s0 : nt;
#pragma omp parallel for num_threads(nt) schedule(static, 1)
        for (int t = 0; t < cnt; t++) {
            std::vector<float> pk(n), pk_int(n), pk_mean(n), conv(n), conv_int(n), conv_mean(n), ak(n), bk(n), ak_int(n),
                bk_int(n), ak_mean(n), bk_mean(n);
            compute_costVolumeOnCpu(guide, other, pk.data(), g1.data(), g2.data(), w, w, h, h, 1, dmin + s0 + t);
            pixelMultOnCPU(pk.data(), im.data(), conv.data(), n);
            integralOnCPU(conv.data(), conv_int.data(), w, h);
            computeBoxFilterOnCPU(conv.data(), conv_int.data(), conv_mean.data(), w, h);
            integralOnCPU(pk.data(), pk_int.data(), w, h);
            computeBoxFilterOnCPU(pk.data(), pk_int.data(), pk_mean.data(), w, h);
            for (int k = 0; k < n; k++) {  // guidedFilter.cu:608-611
                ak[k] = 1.0f * (conv_mean[k] - pk_mean[k] * mean_im[k]) / (var_im[k] + EPS);
                bk[k] = 1.0f * (pk_mean[k] - ak[k] * mean_im[k]);
            }
            integralOnCPU(ak.data(), ak_int.data(), w, h);
            computeBoxFilterOnCPU(ak.data(), ak_int.data(), ak_mean.data(), w, h);
            integralOnCPU(bk.data(), bk_int.data(), w, h);
            computeBoxFilterOnCPU(bk.data(), bk_int.data(), bk_mean.data(), w, h);
            float* q = qs.data() + (size_t)t * n;
            for (int k = 0; k < n; k++) q[k] = (ak_mean[k] * im[k] + bk_mean[k]) * 1.0f;  // :619
        }
        for (int t = 0; t < cnt; t++) dispSelectOnCPU(qs.data() + (size_t)t * n, best, dmap, n, dmin + s0 + t);
    }
}

// Whole pair on the CPU from gray inputs, main.cu:79-155 order.
void ref_pipeline_gray_cpu(unsigned char* gl, unsigned char* gr, int w, int h, int dmin, int size_d, float* dL, float* dR,
                           float* occ, float* filled, float* bestL, float* bestR, int nthreads) {
    const int n = w * h;
    memset(bestL, 9999999.0f, n * sizeof(float));  // main.cu:112-113 (byte fill 0x7F)
    memset(bestR, 9999999.0f, n * sizeof(float));
    memset(dL, 0, n * sizeof(float));
    memset(dR, 0, n * sizeof(float));
    int dmax = dmin + size_d - 1;
    ref_view_disparity_cpu(gl, gr, bestL, dL, nullptr, w, h, size_d, dmin, nthreads);
    ref_view_disparity_cpu(gr, gl, bestR, dR, nullptr, w, h, size_d, -dmax, nthreads);
    memcpy(occ, dL, n * sizeof(float));
    ref_detect_occlusion_cpu(occ, dR, dmin - 100, w, h, size_d + 128);
    memcpy(filled, occ, n * sizeof(float));
    fill_occlusionOnCPU(filled, w, h, (float)dmin);
}

// ---- the reference's own GPU host functions (GPU box only) -------------------------
void ref_rgb_to_gray_gpu(unsigned char* rgb, unsigned char* gray_out, int n, int ch) {
    unsigned char* g = rgb_to_grayscale(rgb, n, ch, false);
    memcpy(gray_out, g, n);
    free(g);
}
void ref_compute_cost_gpu(unsigned char* i1, unsigned char* i2, float* cost, int w, int h, int dmin) {
    compute_cost(i1, i2, cost, w, w, h, h, dmin, false);
}
void ref_integral_gpu(float* image, float* out, int w, int h) { integral(image, out, w, h); }
void ref_compute_guided_filter_gpu(unsigned char* im, float* cost, float* best, float* dmap, unsigned char* mean, int w,
                                   int h, int size_d, int dmin) {
    compute_guided_filter(im, cost, best, dmap, mean, w, h, size_d, dmin, false);
}
void ref_detect_occlusion_gpu(float* dL, float* dR, int dOcclusion, int w, int h) {
    std::vector<unsigned char> a((size_t)w * h, 0), b((size_t)w * h, 0);
    detect_occlusion(dL, dR, dOcclusion, a.data(), b.data(), w, h);
}
void ref_fill_occlusion_gpu(float* disp, int w, int h, float vMin) { fill_occlusion(disp, w, h, vMin); }

}  // extern "C"
