// fma_chain.cu -- measures the FP32 pipe's peak on this GPU: every thread runs 8 independent FFMA chains, enough warps
// per SM to saturate the issue slots.  The roofline denominator bench.py uses is 148 SMs x 128 lanes x sm_max_mhz; this
// is the measured counterpart (SURVEY 8d asks for it next to MEASURED_PEAKS.json).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/fma_chain tools/fma_chain.cu && ./tools/fma_chain
#include <cuda_runtime.h>
#include <cstdio>
__global__ void __launch_bounds__(1024) k_fma(float* out, int iters, float a, float b) {
    float x[8];
#pragma unroll
    for (int i = 0; i < 8; i++) x[i] = threadIdx.x * 1e-3f + i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 16; u++)
#pragma unroll
            for (int i = 0; i < 8; i++) x[i] = fmaf(x[i], a, b);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const int blocks = p.multiProcessorCount * 2, threads = 1024, iters = 4000;
    float* out;
    cudaMalloc(&out, (size_t)blocks * threads * 4);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    double best = 0;
    for (int rep = 0; rep < 5; rep++) {
        cudaEventRecord(e0);
        k_fma<<<blocks, threads>>>(out, iters, 1.0000001f, 1e-7f);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        const double instr = (double)blocks * threads * iters * 128.0;  // FFMA lane-instructions
        const double rate = instr / (ms * 1e-3);
        if (rate > best) best = rate;
        printf("rep %d: %.3f ms, %.2f T FFMA lane-instr/s (%.1f TFLOP/s)\n", rep, ms, rate / 1e12, 2 * rate / 1e12);
    }
    int clk = 0;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("peak measured: %.2f T lane-instr/s = %.1f %% of %d SMs x 128 lanes x %.0f MHz (device attribute clock)\n", best / 1e12,
           100.0 * best / (p.multiProcessorCount * 128.0 * clk * 1e3), p.multiProcessorCount, clk / 1e3);
    return 0;
}
