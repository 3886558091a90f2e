// mma_dev.cuh -- device helpers of the tensor-core fused kernels (fused_mma.cu: gray guide, fused_mma_rgb.cu: RGB guide):
// tcgen05.mma with the A operand in Tensor Memory, un-swizzled shared-memory descriptors, TMEM loads/stores of the shapes
// the kernels use, small shared-memory accessors.  Encodings follow cute::UMMA::{SmemDescriptor, InstrDescriptor}; they
// were checked on the B200 with tools/umma_probe.cu (profiles/r2_umma_probe.txt).
#pragma once
#include "fused_dev.cuh"

#if defined(RGBM_TIMELINE) || defined(MMA_TIMELINE)
// developer instrumentation (tools/timeline_rgb.py, tools/timeline_gray.py): block 0 records, per role and iteration, the start clock and the cycles
// spent in each kind of barrier wait
#define TL_N 4096
static __device__ long long g_tl[6][TL_N][4];
#define TL_DECL long long tl_acc[3] = {0, 0, 0}; long long tl_t0 = 0, tl_m = 0;
#define TL_BEGIN() do { tl_acc[0] = tl_acc[1] = tl_acc[2] = 0; tl_t0 = clock64(); } while (0)
#define TL_WAIT(f, call) do { const long long t_ = clock64(); call; const long long d_ = clock64() - t_; tl_acc[f] += d_; tl_m += d_; } while (0)
#define TL_MARK(seg) do { const long long t_ = clock64(); if ((seg) == 1) tl_m = t_; else if ((seg) == 2) { tl_acc[1] += t_ - tl_m; tl_m = t_; } else { tl_acc[2] += t_ - tl_m; } } while (0)
#define TL_END(role, K) do { if (blockIdx.x == 0 && lane == 0 && (K) < TL_N) { g_tl[role][K][0] = tl_t0; g_tl[role][K][1] = tl_acc[0]; g_tl[role][K][2] = tl_acc[1]; g_tl[role][K][3] = tl_acc[2]; } } while (0)
#else
#define TL_DECL
#define TL_BEGIN() do { } while (0)
#define TL_WAIT(f, call) call
#define TL_MARK(seg) do { } while (0)
#define TL_END(role, K) do { } while (0)
#endif


namespace {

__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n.reg .pred P1;\nelect.sync _|P1, 0xffffffff;\nselp.u32 %0, 1, 0, P1;\n}\n" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// D[tmem] (+)= A[tmem] x B[smem]; one thread issues for the block
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
// un-swizzled shared-memory matrix descriptor (the fields of cute::UMMA::SmemDescriptor): N-major B operand as 8 x 16-byte
// core matrices; LBO = distance of consecutive 8-row K groups, SBO = distance of consecutive 8-column N groups
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) |
           (1ull << 46);
}
// instruction descriptor (cute::UMMA::InstrDescriptor), kind::f16: D f32, A f16 K-major (TMEM), B f16 N-major, M = 128
__host__ __device__ constexpr uint32_t instr_desc(int N, int b_neg) {
    return (1u << 4) | ((uint32_t)b_neg << 14) | (1u << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
__device__ __forceinline__ void tm_st16u(uint32_t taddr, const uint32_t* r) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}
__device__ __forceinline__ void tm_ld16u(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, "
        "[%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tm_ld8(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ void tm_st8(uint32_t taddr, const uint32_t (&r)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]),
                 "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
}
template <int N>
__device__ __forceinline__ void tm_ldN(uint32_t taddr, uint32_t (&r)[N]) {
    static_assert(N == 4 || N == 8 || N == 16, "tcgen05.ld shape");
    if constexpr (N == 4) tm_ld4(taddr, r);
    else if constexpr (N == 8) tm_ld8(taddr, r);
    else tm_ld16u(taddr, r);
}
template <int N>
__device__ __forceinline__ void tm_stN(uint32_t taddr, const uint32_t (&r)[N]) {
    static_assert(N == 8 || N == 16, "tcgen05.st shape");
    if constexpr (N == 8) tm_st8(taddr, r);
    else tm_st16u(taddr, r);
}
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void sts32(uint32_t addr, uint32_t a) {
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(a) : "memory");
}
__device__ __forceinline__ uint2 lds64(uint32_t addr) {
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint32_t lds16(uint32_t addr) {
    unsigned short v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ float2 ld_early_f2(const void* p) {  // coherent: written by this thread one group earlier
    float2 v;
    asm volatile("ld.global.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p) : "memory");
    return v;
}
// f16 + f32 -> f32 with the half taken from either half of a packed pair
__device__ __forceinline__ float fhadd_lo(unsigned h2, float c) { return fhadd(__low2half(u2h2(h2)), c); }
__device__ __forceinline__ float fhadd_hi(unsigned h2, float c) { return fhadd(__high2half(u2h2(h2)), c); }



// ---- geometry and preparation shared by the gray (fused_mma.cu) and RGB (fused_mma_rgb.cu) kernels ----------------
constexpr int M_TW = 128;       // a/b and output lanes per strip = Tensor-Memory lanes
constexpr int M_VW = 110;       // valid output columns per strip (128 - 2*RAD)
constexpr int M_KB = 160;       // cost columns per strip = K of MMA 1 (146 used, 10 K-steps)
constexpr int M_KC = M_TW + 2 * RAD;  // 146 cost columns that feed the 128 a/b lanes
constexpr int M_K2 = 128;       // K of MMA 2 (8 K-steps)
constexpr int M_MTC = 42;       // match chunks (4 pixels x 4 shifted copies, 64 B) per row in an operand slot

struct PrepM {
    const uint8_t* gray;  // held rows, pitch w
    int w, h_held, y_global0, frame_h;
    int n_strips, rows_pad;
    uint2* GA;
    float2* GB;
    __half* GC;
    uint4* MT;
    int n_chunk, padm;
    uint8_t* mean_u8;
    double eps;
    float S, scale;
};

__device__ __forceinline__ bool in_frame(const PrepM& P, int x, int y) {
    const int yg = y + P.y_global0;
    return x >= 0 && x < P.w && y >= 0 && y < P.h_held && yg >= 0 && yg < P.frame_h;
}
// (I, G = I[x-1] - I[x+1]) with the reference's border rule (costVolume.cu:364-378); (1024, 1024) outside: an
// out-of-range match saturates both truncated terms, i.e. gives the reference's constant cost (costVolume.cu:184)
__device__ __forceinline__ void pix_ig(const PrepM& P, int x, int y, float& I, float& G) {
    if (!in_frame(P, x, y)) {
        I = 1024.0f;
        G = 1024.0f;
        return;
    }
    const uint8_t* row = P.gray + (size_t)y * P.w;
    const int ic = row[x];
    const int il = (x - 1 >= 0) ? row[x - 1] : ic;
    const int ir = (x + 1 < P.w) ? row[x + 1] : ic;
    I = (float)ic;
    G = (float)(il - ir);
}

// MT: thread per (padded row, chunk): the chunk's four shifted copies are 7 consecutive pixels dealt out four times
// (the preparation kernels take both images of a pair in one launch: blockIdx.z selects the image)
constexpr int PREP_ROWS = 8;  // rows per block of the gather kernels (k_prep_ga, k_prep_ga3)
__global__ void __launch_bounds__(256) k_prep_mt(const PrepM P0, const PrepM P1) {
    const PrepM& P = blockIdx.z ? P1 : P0;
    const int i = blockIdx.x * 256 + threadIdx.x;
    const int yrow = blockIdx.y;
    if (i >= P.n_chunk) return;
    const int y = yrow - PADY;
    float I[7], G[7];
#pragma unroll
    for (int j = 0; j < 7; j++) pix_ig(P, 4 * i + j - P.padm, y, I[j], G[j]);
    uint4* dst = P.MT + ((size_t)yrow * P.n_chunk + i) * 4;
#pragma unroll
    for (int s = 0; s < 4; s++) {
        uint4 v;
        v.x = h22u(__floats2half2_rn(I[s], I[s + 1]));
        v.y = h22u(__floats2half2_rn(I[s + 2], I[s + 3]));
        v.z = h22u(__floats2half2_rn(G[s], G[s + 1]));
        v.w = h22u(__floats2half2_rn(G[s + 2], G[s + 3]));
        dst[s] = v;
    }
}

Plan make_plan_mma(int w, int rows_out, int size_d, int sm_count, int n_views, int nd) {
    Plan best{};
    double best_cost = 1e300;
    const int n_strips = (w + M_VW - 1) / M_VW;
    const int groups = (size_d + nd - 1) / nd;
    for (int n_chunks = 1; n_chunks <= groups; n_chunks++) {
        const int gpc = (groups + n_chunks - 1) / n_chunks;
        if ((groups + gpc - 1) / gpc != n_chunks) continue;
        for (int n_bands = 1; n_bands <= 64; n_bands++) {
            const int band_rows = (rows_out + n_bands - 1) / n_bands;
            if (n_bands > 1 && band_rows < 64) break;
            if ((rows_out + band_rows - 1) / band_rows != n_bands) continue;
            const long blocks = (long)n_strips * n_bands * n_chunks * n_views;
            const long waves = (blocks + sm_count - 1) / sm_count;
            const double t = (double)waves * gpc * (band_rows + 4.0 * RAD) + 0.02 * n_chunks * rows_out / 64.0;
            if (t < best_cost) {
                best_cost = t;
                best = Plan{n_strips, n_bands, band_rows, n_chunks, gpc * nd};
            }
        }
    }
    return best;
}

struct MmaGeom {
    int padm, n_chunk;
};
MmaGeom mma_geom(int n_strips, int dlo_all, int dhi_all) {
    // padded match columns X = x + padm cover x from (first cost column + lowest d) to (last cost column + highest d + 7),
    // rounded out to whole chunks of the operand slot
    const int xmin = -2 * RAD + dlo_all, xmax = (n_strips - 1) * M_VW - 2 * RAD + M_KB + dhi_all + 8;
    const int padm = (max(0, -xmin) + 4 + 3) / 4 * 4;
    const int n_chunk = (padm + xmax + 4 * (M_MTC + 2) + 3) / 4;
    return MmaGeom{padm, n_chunk};
}

}  // namespace
