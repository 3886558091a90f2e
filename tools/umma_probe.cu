// umma_probe.cu -- stand-alone B200 probe for the tensor-core box filter of the MMA fused kernel.
// Answers, on the hardware, the questions the design rests on:
//   1. layout: A (the 0/1 band matrix) in Tensor Memory, B (data) in shared memory as N-major, un-swizzled
//      8x16-byte core matrices, D in Tensor Memory -- does D[m][n] = sum_{k=m..m+18} B[k][n] come out exactly?
//   2. accumulation: are sums of integer-valued halves exact; how is the fp32 accumulator rounded?
//   3. throughput: clocks per tcgen05.mma at M=128, N=16/32/64, K=16 with A in TMEM; tcgen05.ld/st bytes per clock.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/umma_probe tools/umma_probe.cu
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x)                                                                                  \
    do {                                                                                       \
        cudaError_t e = (x);                                                                   \
        if (e != cudaSuccess) {                                                                \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__);     \
            exit(1);                                                                           \
        }                                                                                      \
    } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t b, unsigned n) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(b), "r"(n) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t b, unsigned parity) {
    asm volatile(
        "{\n.reg .pred p;\nW1:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D1;\nbra W1;\nD1:\n}\n" ::"r"(b),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tm_alloc(uint32_t* dst, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst)), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tm_dealloc(uint32_t a, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(a), "r"(cols) : "memory");
}
__device__ __forceinline__ void fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tm_st1(uint32_t a, uint32_t v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(a), "r"(v) : "memory");
}
__device__ __forceinline__ void tm_ld16(uint32_t a, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(a)
        : "memory");
}
__device__ __forceinline__ void tm_st16(uint32_t a, const uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(a),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}
__device__ __forceinline__ void tm_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tm_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n.reg .pred P1;\nelect.sync _|P1, 0xffffffff;\nselp.u32 %0, 1, 0, P1;\n}\n" : "=r"(pred));
    return pred != 0;
}
// D[tmem] (+)= A[tmem] * B[smem desc]; issued by ONE thread
__device__ __forceinline__ void umma_ts(uint32_t d, uint32_t a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}\n" ::"r"(d),
        "r"(a), "l"(bdesc), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t mbar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(mbar) : "memory");
}
// un-swizzled shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start>>4 [0,14), LBO>>4 [16,30),
// SBO>>4 [32,46), version 1 at [46,48), layout type 0 at [61,64)
__host__ __device__ inline uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) |
           (1ull << 46);
}
// instruction descriptor (cute::UMMA::InstrDescriptor), kind::f16: D f32, A/B f16, A K-major (TMEM), B major as given
__host__ __device__ inline uint32_t make_idesc(int M, int N, int b_mn_major, int b_neg) {
    return (1u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)b_neg << 14) | (0u << 15) | ((uint32_t)b_mn_major << 16) |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

constexpr int KB = 160;   // K of the band product (10 MMA K-steps)
constexpr int WIN = 19;

// ---- test 1: band product, layout variants --------------------------------------------------------------------
// variant 0: B N-major, LBO = 128 (K groups contiguous), SBO = KB*16 (N groups)      <- the design
// variant 1: same memory, LBO/SBO fields swapped
// variant 2: B K-major: element (k,n) at (n%8)*16 + (n/8)*SBO + (k%8)*2 + (k/8)*LBO, LBO = 128, SBO = (KB/8)*128
// variant 3: as 2 with the fields swapped
struct T1 {
    const __half* B;  // [KB][N] row-major on the host side
    float* D;         // [128][N]
    int N, variant, decode;
};
__global__ void __launch_bounds__(128) k_band(const T1 t) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint32_t tbase_s;
    __shared__ __align__(8) uint64_t bar;
    const int tid = threadIdx.x, warp = tid >> 5;
    const int N = t.N;
    if (warp == 0) tm_alloc(&tbase_s, 512);
    if (tid == 0) mbar_init(smem_u32(&bar), 1);
    // B into shared memory
    const bool kmaj = t.variant >= 2;
    const uint32_t lbo = 128, sbo = kmaj ? (KB / 8) * 128 : KB * 16;
    __half* sB = reinterpret_cast<__half*>(smem);
    for (int i = tid; i < KB * N; i += 128) {
        const int k = i / N, n = i % N;
        uint32_t off;
        if (!kmaj) off = (k % 8) * 16 + (k / 8) * lbo + (n / 8) * sbo + (n % 8) * 2;
        else off = (n % 8) * 16 + (n / 8) * sbo + (k % 8) * 2 + (k / 8) * lbo;
        sB[off / 2] = t.B[i];
    }
    fence_async_smem();
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tb = tbase_s;
    const uint32_t A0 = 256, D0 = 0;
    // band rows: lane m, K element k = 2*col + (half index)
    {
        const int m = tid;
        const uint32_t ta = tb + ((uint32_t)(warp * 32) << 16) + A0;
        for (int j = 0; j < KB / 2; j++) {
            float lo, hi;
            if (t.decode) {
                lo = ((2 * j) % 16 == m % 16 && (2 * j) / 16 == 0) ? 1.0f : 0.0f;
                hi = ((2 * j + 1) % 16 == m % 16 && (2 * j + 1) / 16 == 0) ? 1.0f : 0.0f;
            } else {
                lo = (2 * j >= m && 2 * j < m + WIN) ? 1.0f : 0.0f;
                hi = (2 * j + 1 >= m && 2 * j + 1 < m + WIN) ? 1.0f : 0.0f;
            }
            __half2 v = __floats2half2_rn(lo, hi);
            tm_st1(ta + j, *reinterpret_cast<uint32_t*>(&v));
        }
        tm_wait_st();
    }
    fence_before();
    __syncthreads();
    fence_after();
    if (tid == 0) {
        const uint32_t idesc = make_idesc(128, N, kmaj ? 0 : 1, 0);
        const int ksteps = t.decode ? 1 : KB / 16;
        for (int j = 0; j < ksteps; j++) {
            const uint32_t start = smem_u32(smem) + j * 2 * lbo;
            uint64_t desc = (t.variant & 1) ? make_desc(start, sbo, lbo) : make_desc(start, lbo, sbo);
            umma_ts(tb + D0, tb + A0 + 8 * j, desc, idesc, j > 0);
        }
        umma_commit(smem_u32(&bar));
    }
    mbar_wait(smem_u32(&bar), 0);
    fence_after();
    for (int c = 0; c < N; c += 16) {
        uint32_t r[16];
        tm_ld16(tb + ((uint32_t)(warp * 32) << 16) + D0 + c, r);
        tm_wait_ld();
        for (int i = 0; i < 16; i++) t.D[tid * N + c + i] = __uint_as_float(r[i]);
    }
    fence_before();
    __syncthreads();
    if (warp == 0) {
        fence_after();
        tm_dealloc(tb, 512);
    }
}

// ---- test 2: accumulator rounding ------------------------------------------------------------------------------
// A = all ones (K=16), B column n holds a probe vector; repeated accumulation.
struct T2 {
    const __half* B;  // [reps][16][16]
    float* D;         // [16]
    int reps;
};
__global__ void __launch_bounds__(128) k_round(const T2 t) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint32_t tbase_s;
    __shared__ __align__(8) uint64_t bar;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (warp == 0) tm_alloc(&tbase_s, 512);
    if (tid == 0) mbar_init(smem_u32(&bar), 1);
    __half* sB = reinterpret_cast<__half*>(smem);
    for (int i = tid; i < t.reps * 256; i += 128) {
        const int rep = i / 256, k = (i % 256) / 16, n = i % 16;
        const uint32_t off = rep * 512 + (k % 8) * 16 + (k / 8) * 128 + (n / 8) * 256 + (n % 8) * 2;
        sB[off / 2] = t.B[i];
    }
    fence_async_smem();
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tb = tbase_s;
    {
        const uint32_t ta = tb + ((uint32_t)(warp * 32) << 16) + 256;
        __half2 one = __floats2half2_rn(1.0f, 1.0f);
        for (int j = 0; j < 8; j++) tm_st1(ta + j, *reinterpret_cast<uint32_t*>(&one));
        tm_wait_st();
    }
    fence_before();
    __syncthreads();
    fence_after();
    if (tid == 0) {
        const uint32_t idesc = make_idesc(128, 16, 1, 0);
        for (int r = 0; r < t.reps; r++) umma_ts(tb, tb + 256, make_desc(smem_u32(smem) + r * 512, 128, 256), idesc, r > 0);
        umma_commit(smem_u32(&bar));
    }
    mbar_wait(smem_u32(&bar), 0);
    fence_after();
    uint32_t r[16];
    tm_ld16(tb + ((uint32_t)(warp * 32) << 16), r);
    tm_wait_ld();
    if (tid == 0)
        for (int i = 0; i < 16; i++) t.D[i] = __uint_as_float(r[i]);
    fence_before();
    __syncthreads();
    if (warp == 0) {
        fence_after();
        tm_dealloc(tb, 512);
    }
}

// ---- test 3: throughput ------------------------------------------------------------------------------------------
// mode 0: `issuers` threads (one per warp) each issue `n` MMAs (M=128, N, K=16; A in TMEM, B in shared memory)
// mode 1: every warp loops tcgen05.ld x16;  mode 2: tcgen05.st x16;  mode 3: ld x16 + st x16 alternating
struct T3 {
    long long* clk;  // [blocks]
    int mode, N, n, issuers;
};
__global__ void __launch_bounds__(512) k_tput(const T3 t) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint32_t tbase_s;
    __shared__ __align__(8) uint64_t bar[8];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (warp == 0) tm_alloc(&tbase_s, 512);
    if (tid == 0)
        for (int i = 0; i < 8; i++) mbar_init(smem_u32(&bar[i]), 1);
    for (int i = tid; i < 16384; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    fence_async_smem();
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tb = tbase_s;
    const uint32_t tl = tb + ((uint32_t)((warp & 3) * 32) << 16);
    if (warp < 4) {
        uint32_t z[16];
        for (int i = 0; i < 16; i++) z[i] = 0;
        for (int c = 0; c < 512; c += 16) tm_st16(tl + c, z);
        tm_wait_st();
    }
    fence_before();
    __syncthreads();
    fence_after();
    long long t0 = clock64();
    if (t.mode == 0) {
        // warp-uniform issue: descriptors from uniform values, one elected lane issues (UTCHMMA with uniform registers,
        // no per-instruction R2UR waterfall)
        const int wu = __shfl_sync(0xffffffffu, warp, 0);
        if (wu < t.issuers) {
            const uint32_t idesc = make_idesc(128, t.N, 1, 0);
            const uint32_t sb = smem_u32(smem) + wu * 8192;
            const uint64_t d0 = make_desc(sb, 128, 2560);
            for (int i = 0; i < t.n; i += 10) {
                if (elect_one()) {
#pragma unroll
                    for (int j = 0; j < 10; j++) umma_ts(tb + wu * 64, tb + 384 + 8 * j, d0 + (uint64_t)(j * 16), idesc, 1);
                }
                __syncwarp();
            }
            if (elect_one()) umma_commit(smem_u32(&bar[wu]));
            __syncwarp();
            mbar_wait(smem_u32(&bar[wu]), 0);
        }
    } else {
        uint32_t r[16];
        for (int i = 0; i < 16; i++) r[i] = i;
        const uint32_t col = (warp >> 2) * 32;  // warps sharing a lane quarter use different columns
        for (int i = 0; i < t.n; i++) {
            if (t.mode == 1 || t.mode == 3) tm_ld16(tl + col + ((i & 1) * 16), r);
            if (t.mode == 2 || t.mode == 3) tm_st16(tl + col + 256 + ((i & 1) * 16), r);
            if ((i & 3) == 3) {
                if (t.mode != 2) tm_wait_ld();
                if (t.mode != 1) tm_wait_st();
            }
        }
        tm_wait_ld();
        tm_wait_st();
        if (r[3] == 0xdeadbeef) t.clk[0] = 1;
    }
    fence_before();
    __syncthreads();
    long long t1 = clock64();
    if (tid == 0) t.clk[blockIdx.x] = t1 - t0;
    if (warp == 0) {
        fence_after();
        tm_dealloc(tb, 512);
    }
}

int main() {
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    printf("device: %s, %d SMs, cc %d.%d\n", prop.name, prop.multiProcessorCount, prop.major, prop.minor);

    // ---- test 1
    for (int decode = 1; decode >= 0; decode--) {
        for (int N : {16, 32}) {
            std::vector<__half> hB(KB * N);
            std::vector<float> fB(KB * N);
            srand(7);
            for (int i = 0; i < KB * N; i++) {
                float v = decode ? (float)(100 * (i / N) + (i % N)) : (float)(rand() % 751);
                if (decode && v > 2000) v = 0;
                fB[i] = v;
                hB[i] = __float2half(v);
            }
            __half* dB;
            float* dD;
            CK(cudaMalloc(&dB, hB.size() * 2));
            CK(cudaMalloc(&dD, 128 * N * 4));
            CK(cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice));
            CK(cudaFuncSetAttribute(k_band, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
            for (int variant = 0; variant < 4; variant++) {
                CK(cudaMemset(dD, 0xff, 128 * N * 4));
                T1 t{dB, dD, N, variant, decode};
                k_band<<<1, 128, 65536>>>(t);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) {
                    printf("test1 decode=%d N=%d variant=%d: CUDA error %s\n", decode, N, variant, cudaGetErrorString(e));
                    return 1;
                }
                std::vector<float> hD(128 * N);
                CK(cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost));
                double maxerr = 0;
                int bad = 0;
                for (int m = 0; m < 128; m++)
                    for (int n = 0; n < N; n++) {
                        double ex = 0;
                        if (decode) ex = fB[(m % 16) * N + n];
                        else
                            for (int k = m; k < m + WIN; k++) ex += fB[k * N + n];
                        double err = fabs(ex - hD[m * N + n]);
                        if (!(err <= 0.0)) bad++;
                        if (err > maxerr || err != err) maxerr = err;
                    }
                printf("test1 decode=%d N=%d variant=%d: mismatches %d / %d, max err %g\n", decode, N, variant, bad, 128 * N, maxerr);
                if (bad && decode && N == 16) {
                    for (int m : {0, 1, 2, 17, 33, 127}) {
                        printf("   D[%3d][0..15]:", m);
                        for (int n = 0; n < 16; n++) printf(" %g", hD[m * N + n]);
                        printf("\n");
                    }
                }
            }
            cudaFree(dB);
            cudaFree(dD);
        }
    }

    // ---- test 2: rounding probes (column n of D = sum over reps and k of B[rep][k][n])
    {
        const int reps = 33;
        std::vector<__half> hB(reps * 256, __float2half(0.0f));
        auto B = [&](int rep, int k, int n) -> __half& { return hB[rep * 256 + k * 16 + n]; };
        // col 0: 32 x (16 x 32768) = 2^24, then +3     -> RN 16777220, RZ 16777218
        // col 1: 2^24 then +1                          -> RN(even)/RZ 16777216, RU 16777218
        // col 2: 2^24 then -3 ... as col 0 with sign    -> RN -16777220 ... uses negative values
        // col 3: inside one MMA: 32768 + 15 x 2^-9      (exact 32768.029296875; fp32 ulp 2^-8 -> RN 32768.03125, trunc-each 32768)
        // col 4: inside one MMA: 15 x 65504 + 0.5       (exact 982560.5, representable)
        // col 5: hi/lo: 19 x (1000 + 2^-3)  hi and lo in the same MMA: exact 19002.375
        // col 6: 2^24 then +1 then +1 (two reps)        -> sticky?
        // col 7: integer sums 16 x 750 x 33 reps = 396000 exact
        for (int r = 0; r < 32; r++)
            for (int k = 0; k < 16; k++) {
                B(r, k, 0) = __float2half(32768.0f);
                B(r, k, 1) = __float2half(32768.0f);
                B(r, k, 2) = __float2half(-32768.0f);
                B(r, k, 6) = __float2half(k < 15 ? 32768.0f : 32768.0f);
            }
        B(32, 0, 0) = __float2half(3.0f);
        B(32, 0, 1) = __float2half(1.0f);
        B(32, 0, 2) = __float2half(-3.0f);
        B(0, 0, 3) = __float2half(32768.0f);
        for (int k = 1; k < 16; k++) B(0, k, 3) = __float2half(0.001953125f);
        for (int k = 0; k < 15; k++) B(0, k, 4) = __float2half(65504.0f);
        B(0, 15, 4) = __float2half(0.5f);
        for (int k = 0; k < 8; k++) {
            B(0, k, 5) = __float2half(1000.0f);
            B(0, k + 8, 5) = __float2half(0.125f);
        }
        B(32, 0, 6) = __float2half(1.0f);
        B(32, 1, 6) = __float2half(1.0f);
        for (int r = 0; r < 33; r++)
            for (int k = 0; k < 16; k++) B(r, k, 7) = __float2half(750.0f);
        __half* dB;
        float* dD;
        CK(cudaMalloc(&dB, hB.size() * 2));
        CK(cudaMalloc(&dD, 64));
        CK(cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice));
        CK(cudaFuncSetAttribute(k_round, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
        T2 t{dB, dD, reps};
        k_round<<<1, 128, 65536>>>(t);
        CK(cudaDeviceSynchronize());
        float hD[16];
        CK(cudaMemcpy(hD, dD, 64, cudaMemcpyDeviceToHost));
        printf("test2 col0 (2^24+3: RN 16777220, RZ 16777218): %.1f\n", hD[0]);
        printf("test2 col1 (2^24+1: RN/RZ 16777216, RU 16777218): %.1f\n", hD[1]);
        printf("test2 col2 (-(2^24+3): RN -16777220, RZ -16777218, RD -16777220): %.1f\n", hD[2]);
        printf("test2 col3 (32768 + 15*2^-9 in one MMA: exact 32768.029297, RN 32768.03125): %.6f\n", hD[3]);
        printf("test2 col4 (15*65504+0.5 = 982560.5): %.2f\n", hD[4]);
        printf("test2 col5 (8*1000 + 8*0.125 = 8001): %.4f\n", hD[5]);
        printf("test2 col6 (2^24 + (1+1) in one MMA: 16777218): %.1f\n", hD[6]);
        printf("test2 col7 (33*16*750 = 396000): %.1f\n", hD[7]);
        cudaFree(dB);
        cudaFree(dD);
    }

    // ---- test 3
    {
        long long* dclk;
        const int nb = prop.multiProcessorCount;
        CK(cudaMalloc(&dclk, nb * 8));
        CK(cudaFuncSetAttribute(k_tput, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
        std::vector<long long> h(nb);
        auto run = [&](int mode, int N, int n, int issuers, int threads, const char* what, double unit_per_iter) {
            T3 t{dclk, mode, N, n, issuers};
            for (int rep = 0; rep < 2; rep++) {
                k_tput<<<nb, threads, 65536>>>(t);
                CK(cudaDeviceSynchronize());
            }
            CK(cudaMemcpy(h.data(), dclk, nb * 8, cudaMemcpyDeviceToHost));
            long long mx = 0;
            for (auto v : h) mx = v > mx ? v : mx;
            printf("test3 %-44s: %8lld clk, %.2f clk/op, %.1f %s/clk/SM\n", what, mx, (double)mx / n, unit_per_iter * n / mx,
                   mode == 0 ? "MAC" : "B");
        };
        const int n = 4000;
        run(0, 16, n, 1, 128, "mma M128 N16 K16, 1 issuer", 128.0 * 16 * 16);
        run(0, 16, n, 2, 128, "mma M128 N16 K16, 2 issuers", 2 * 128.0 * 16 * 16);
        run(0, 16, n, 4, 128, "mma M128 N16 K16, 4 issuers", 4 * 128.0 * 16 * 16);
        run(0, 32, n, 1, 128, "mma M128 N32 K16, 1 issuer", 128.0 * 32 * 16);
        run(0, 32, n, 2, 128, "mma M128 N32 K16, 2 issuers", 2 * 128.0 * 32 * 16);
        run(0, 64, n, 1, 128, "mma M128 N64 K16, 1 issuer", 128.0 * 64 * 16);
        run(0, 64, n, 2, 128, "mma M128 N64 K16, 2 issuers", 2 * 128.0 * 64 * 16);
        for (int threads : {128, 256, 512}) {
            char buf[64];
            const double bytes = threads * 16.0 * 4;
            snprintf(buf, 64, "tcgen05.ld x16, %d warps", threads / 32);
            run(1, 0, n, 0, threads, buf, bytes);
            snprintf(buf, 64, "tcgen05.st x16, %d warps", threads / 32);
            run(2, 0, n, 0, threads, buf, bytes);
            snprintf(buf, 64, "tcgen05.ld+st x16, %d warps", threads / 32);
            run(3, 0, n, 0, threads, buf, 2 * bytes);
        }
        cudaFree(dclk);
    }
    printf("probe done\n");
    return 0;
}
