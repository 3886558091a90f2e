// fused_dev.cuh -- device helpers and host planning shared by the fused kernels
// (fused_cvf.cu: gray guide, fused_cvf_rgb.cu: RGB guide).
#pragma once
#include <type_traits>

#include "common.cuh"

namespace {

constexpr int KPX = 8;            // pixels per lane
constexpr int SW = 32 * KPX;      // columns per warp strip
constexpr int HALO = 20;          // left halo (>= 2*radius, multiple of 4)
constexpr int VALID_W = 216;      // valid output columns per strip: local [20, 236)
constexpr int RAD = 9;
constexpr int WIN = 2 * RAD + 1;  // 19
constexpr int NWARP = 4;          // disparities in flight per block (= warp pairs)
constexpr int NTHREADS = 2 * NWARP * 32;
constexpr int ROWS = 2;           // image rows per pipeline iteration: independent horizontal work for ILP
constexpr int PADY = 40;          // padding rows above/below the prepared planes
constexpr float BEST_INIT_BITS_F = 3.3961514e38f;  // 0x7F7F7F7F, main.cu:112
// A filtered row in the q ring: pixels 0..3 and 4..7 of every lane as two planes of float4.  The planes are 36 float4
// apart (4 of padding): a merging thread reads 8 bytes of one plane, and with 32 float4 per plane the two planes of a
// half-warp's reads fell on the same 16 banks (ncu: twice the ideal wavefronts on these loads)
#ifndef FUSED_QV
#define FUSED_QV 36
#endif
constexpr int QV = FUSED_QV;

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void named_bar_arrive(int id, int nthreads) {
    asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// mbarrier (shared-memory barrier object with phase parity) -- used where a producer/consumer
// ring is deeper than the 16 hardware named barriers allow
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
// (the barriers are addressed by their 32-bit shared-window address, computed once per kernel)
__device__ __forceinline__ void mbar_init(uint32_t b, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(b), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t b) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(b) : "memory");
}
// arrive from lane 0 only, predicated instead of branched (the warp must be converged)
__device__ __forceinline__ void mbar_arrive_lane0(uint32_t b, int lane) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.eq.s32 p, %1, 0;\n"
        "@p mbarrier.arrive.shared::cta.b64 _, [%0];\n"
        "}\n" ::"r"(b),
        "r"(lane)
        : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t b, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "MB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra MB_DONE;\n"
        "bra MB_WAIT;\n"
        "MB_DONE:\n"
        "}\n" ::"r"(b),
        "r"(parity)
        : "memory");
}
// non-blocking: has the phase with this parity completed?
__device__ __forceinline__ bool mbar_test(uint32_t b, unsigned parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(b), "r"(parity)
        : "memory");
    return ok != 0;
}

// ---- Tensor Memory as per-thread scratch -------------------------------------------------
// The 19-row rings and the producer->consumer hand-off live in TMEM (256 KB per SM, idle in a
// kernel without tensor-core work) instead of shared memory: tcgen05.ld/st have their own data
// path, so this traffic (72 of ~205 wavefronts per row, round-1 ncu: L1TEX data pipe at 78 %)
// leaves the shared-memory/shuffle/L1 pipe.  With the 32x32b shape every thread of a warp owns
// one TMEM lane (warp w may touch lanes 32*(w%4)..+31) and N consecutive 32-bit columns, i.e.
// private scratch; warps p and p+4 (a producer/consumer pair) share a lane quarter, which is
// what lets the producer hand its rows to the consumer through TMEM.
// Column map of one lane (512 allocated): [0,304) ring of (a[8],b[8]) x 19 slots,
// [304,380) ring of the lattice cost (8 halfs = 4 words) x 19 slots, [384,416) hand-off rows
// (the gray kernel adds [416,480): two slots of rows between its second and third stage).
constexpr uint32_t TM_COLS = 512;
constexpr uint32_t TM_RING_AB = 0, TM_RING_P = 304, TM_HAND = 384;

__device__ __forceinline__ void tm_alloc(uint32_t* smem_dst) {  // one converged warp
    uint32_t dst = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst), "r"(TM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tm_dealloc(uint32_t taddr) {  // the warp that allocated
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(TM_COLS) : "memory");
}
__device__ __forceinline__ void tm_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tm_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tm_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tm_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tm_ld4(uint32_t taddr, uint32_t (&r)[4]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ void tm_st4(uint32_t taddr, const uint32_t (&r)[4]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
                 "r"(r[2]), "r"(r[3])
                 : "memory");
}
__device__ __forceinline__ void tm_ld16(uint32_t taddr, float (&a)[8], float (&b)[8]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, "
        "[%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
#pragma unroll
    for (int j = 0; j < 8; j++) {
        a[j] = __uint_as_float(r[j]);
        b[j] = __uint_as_float(r[8 + j]);
    }
}
__device__ __forceinline__ void tm_st16(uint32_t taddr, const float (&a)[8], const float (&b)[8]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16};" ::"r"(taddr),
        "r"(__float_as_uint(a[0])), "r"(__float_as_uint(a[1])), "r"(__float_as_uint(a[2])), "r"(__float_as_uint(a[3])),
        "r"(__float_as_uint(a[4])), "r"(__float_as_uint(a[5])), "r"(__float_as_uint(a[6])), "r"(__float_as_uint(a[7])),
        "r"(__float_as_uint(b[0])), "r"(__float_as_uint(b[1])), "r"(__float_as_uint(b[2])), "r"(__float_as_uint(b[3])),
        "r"(__float_as_uint(b[4])), "r"(__float_as_uint(b[5])), "r"(__float_as_uint(b[6])), "r"(__float_as_uint(b[7]))
        : "memory");
}

// 19-wide horizontal window sums for the 8 consecutive pixels this lane holds.
// Lane L holds columns 8L..8L+7.  With LP_L[i] the lane-local prefix sums and T_L the lane
// total, the window [x-9, x+9] for pixel j of lane L is
//   (T_{L-1} - LP_{L-1}[j-2]) + T_L + LP_{L+1}[j+1]            for 2 <= j <= 6
// with the obvious edge forms for j = 0, 1, 7 (one pixel of lanes L-2 / L+2 is needed).
// 16 shuffles + 26 adds per 8 pixels.  Results are valid for lanes whose neighbours exist,
// i.e. for strip-local columns [9, 247).
__device__ __forceinline__ void hsum19(const float (&v)[KPX], float (&h)[KPX]) {
    const unsigned F = 0xffffffffu;
    float lp[KPX];
    lp[0] = v[0];
#pragma unroll
    for (int i = 1; i < KPX; i++) lp[i] = lp[i - 1] + v[i];
    const float T = lp[7];
    const float Tm1 = __shfl_up_sync(F, T, 1);
    const float Tp1 = __shfl_down_sync(F, T, 1);
    const float C = T + Tm1;
    float r[7];
#pragma unroll
    for (int j = 0; j < 6; j++) r[j] = __shfl_down_sync(F, lp[j + 1], 1);
    r[6] = Tp1;
    float l[6];
#pragma unroll
    for (int i = 0; i < 6; i++) l[i] = __shfl_up_sync(F, lp[i], 1);
    const float vm2 = __shfl_up_sync(F, v[7], 2);
    const float vp2 = __shfl_down_sync(F, v[0], 2);
    h[0] = (vm2 + C) + r[0];
    h[1] = C + r[1];
#pragma unroll
    for (int j = 2; j <= 6; j++) h[j] = (C + r[j]) - l[j - 2];
    h[7] = ((C - l[5]) + Tp1) + vp2;
}

// ---- strip-tiled planes ------------------------------------------------------------------
// Strip s covers frame columns [VALID_W*s - HALO, VALID_W*s - HALO + SW); the 2*HALO columns two neighbouring strips
// share are stored in both.  A record = one padded row of one strip, laid out as [16-byte chunk][lane][16 B] so that
// the c-th 128-bit load of lane L (pixels 8L .. 8L+7 of the strip) is 16-byte unit c*32 + L of the record: one bulk
// copy fills a shared-memory slot and every read of it is conflict-free.  The functions below say where strip-local
// column cl (lane cl>>3, pixel j = cl&7) of record rec lives in each plane; the prep kernels store through them and
// tests/layout_check.cu checks them on the CPU against the pixel order the fused kernels assume.
__device__ __host__ __forceinline__ int strips_of_column(int x, int n_strips, int (&strip)[2], int (&cl)[2]) {
    const int xa = x + HALO;
    if (xa < 0) return 0;
    const int s1 = xa / VALID_W, c1 = xa - s1 * VALID_W;
    int n = 0;
    for (int k = 0; k < 2; k++) {
        const int s = s1 - k, c = c1 + k * VALID_W;
        if (s < 0 || s >= n_strips || c >= SW) continue;
        strip[n] = s;
        cl[n] = c;
        n++;
    }
    return n;
}
// (I,G) words: chunk = j>>2 (pixels 0..3 | 4..7), 64 16-byte units per record; index in 4-byte words
__device__ __host__ __forceinline__ size_t tg_index(size_t rec, int cl) {
    const int L = cl >> 3, j = cl & 7;
    return (rec * 64 + (j >> 2) * 32 + L) * 4 + (j & 3);
}
// I as half: one chunk of 8 halves per lane, 32 units per record; index in halves
__device__ __host__ __forceinline__ size_t ti_index(size_t rec, int cl) { return (rec * 32 + (cl >> 3)) * 8 + (cl & 7); }
// (mean_I, c2) float2: chunk = j>>1, 128 units per record; index in float2
__device__ __host__ __forceinline__ size_t tst_index(size_t rec, int cl) {
    const int L = cl >> 3, j = cl & 7;
    return (rec * 128 + (j >> 1) * 32 + L) * 2 + (j & 1);
}
// colour as halves: chunk = channel, 96 units per record; index in halves
__device__ __host__ __forceinline__ size_t tc_index(size_t rec, int ch, int cl) {
    return rec * 768 + (size_t)((ch * 32 + (cl >> 3)) * 8 + (cl & 7));
}
// RGB statistics: chunk 2j = (mu_r, mu_g, mu_b, M_rr), 2j+1 = (M_rg, M_rb, M_gg, M_gb) of pixel j (index in float4),
// chunks 16, 17 = M_bb of pixels 0..3, 4..7 (index in floats); 576 units per record
__device__ __host__ __forceinline__ size_t ts_s1_index(size_t rec, int cl) {
    return rec * 576 + (size_t)((2 * (cl & 7)) * 32 + (cl >> 3));
}
__device__ __host__ __forceinline__ size_t ts_s2_index(size_t rec, int cl) {
    return rec * 576 + (size_t)((2 * (cl & 7) + 1) * 32 + (cl >> 3));
}
__device__ __host__ __forceinline__ size_t ts_s3_index(size_t rec, int cl) {
    const int L = cl >> 3, j = cl & 7;
    return rec * 2304 + (size_t)(((16 + (j >> 2)) * 32 + L) * 4 + (j & 3));
}

// Match operands.  A warp reads, per row, the 8 pixels x 32 lanes at columns x + d: 32 B per lane.  From a linear
// plane that is two 128-bit loads whose lanes are 32 B apart (half of every sector fetched by each, 8.75 L1 tag
// requests per load in ncu).  The shifted copies of the (I,G) plane are therefore stored DE-INTERLEAVED: the even
// groups of 4 pixels in the first half of a copy, the odd groups in the second half, so that the first 4 pixels of all
// lanes are 512 contiguous bytes, and the next 4 likewise.  match_ptrs() returns the two pointers for linear element
// index e (a multiple of 4) of a copy; both advance by pitch/2 elements per row.
__device__ __host__ __forceinline__ size_t deint_index(size_t i, size_t half_plane) {  // linear element -> stored element
    const size_t g4 = i >> 2;
    return (g4 & 1 ? half_plane : 0) + (g4 >> 1) * 4 + (i & 3);
}
__device__ __host__ __forceinline__ void match_ptrs(const unsigned* copy, long long e, size_t half_plane, const unsigned*& p0,
                                                    const unsigned*& p1) {
    if ((e >> 2) & 1) {
        p0 = copy + half_plane + (e - 4) / 2;
        p1 = copy + (e + 4) / 2;
    } else {
        p0 = copy + e / 2;
        p1 = copy + half_plane + e / 2;
    }
}

__device__ __forceinline__ __half2 u2h2(unsigned u) { return *reinterpret_cast<__half2*>(&u); }
__device__ __forceinline__ unsigned h22u(__half2 h) { return *reinterpret_cast<unsigned*>(&h); }

// ---- helpers of the three-stage kernels (fused_cvf.cu, fused_cvf_rgb3.cu) ----------------------
// f16 x f16 + f32 -> f32 and f16 + f32 -> f32 in one FMA-pipe instruction (FHFMA / FHADD, PTX 8.6, sm_100+)
__device__ __forceinline__ float fhfma(__half a, __half b, float c) {
    float d;
    asm("fma.rn.f32.f16 %0, %1, %2, %3;" : "=f"(d) : "h"(__half_as_ushort(a)), "h"(__half_as_ushort(b)), "f"(c));
    return d;
}
__device__ __forceinline__ float fhadd(__half a, float c) {
    float d;
    asm("add.rn.f32.f16 %0, %1, %2;" : "=f"(d) : "h"(__half_as_ushort(a)), "f"(c));
    return d;
}
// a load the compiler may not sink towards its use
__device__ __forceinline__ float4 ld_early_f4(const void* p) {  // coherent: written by this block one group earlier
    float4 v;
    asm volatile("ld.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}

// TMA bulk copy global -> shared, completion counted in bytes on an mbarrier
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t mbar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(mbar)
                 : "memory");
}
// ask L2 for `bytes` (multiple of 16) at a 16-byte aligned global address; no destination, no completion
__device__ __forceinline__ void l2_prefetch(const void* src, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t b, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}

__device__ __forceinline__ float inv_rows(const float* lut, int y, int y_global0, int frame_h) {
    // 1 / (scale * clipped window height) at held row y, 0 outside the frame: the height takes
    // 19 values, the reciprocals come from a table in shared memory
    int yg = y + y_global0;
    int ay = min(frame_h - 1, yg + RAD) - max(0, yg - RAD) + 1;
    if (yg < 0 || yg >= frame_h) ay = 0;
    return lut[ay];
}

// Fold the 4 disparities of a group into the running (best,label) of two columns: ascending d, `best >= q` (the last
// slice wins ties, guidedFilter.cu:406).  "Minimum, the later index on a tie" is associative, so a 2-level tournament
// gives what the reference's sequential scan gives, with a shorter dependency chain.  qp: this thread's two columns
// in the q row of the group's first warp (the other warps follow at `stride` floats).
__device__ __forceinline__ float4 merge4(const float* qp, int stride, const float (&lab)[NWARP], const float4 pb) {
    static_assert(NWARP == 4, "tournament written for 4 disparities per group");
    float b0 = pb.x, l0 = pb.y, b1 = pb.z, l1 = pb.w;
    float2 qv[NWARP];
#pragma unroll
    for (int wv = 0; wv < NWARP; wv++) qv[wv] = *reinterpret_cast<const float2*>(qp + wv * stride);
    {
        const bool t01 = qv[0].x >= qv[1].x, t23 = qv[2].x >= qv[3].x;
        const float m01 = t01 ? qv[1].x : qv[0].x, m23 = t23 ? qv[3].x : qv[2].x;
        const float a01 = t01 ? lab[1] : lab[0], a23 = t23 ? lab[3] : lab[2];
        const bool t = m01 >= m23;
        const float m = t ? m23 : m01, a = t ? a23 : a01;
        if (b0 >= m) { b0 = m; l0 = a; }
    }
    {
        const bool t01 = qv[0].y >= qv[1].y, t23 = qv[2].y >= qv[3].y;
        const float m01 = t01 ? qv[1].y : qv[0].y, m23 = t23 ? qv[3].y : qv[2].y;
        const float a01 = t01 ? lab[1] : lab[0], a23 = t23 ? lab[3] : lab[2];
        const bool t = m01 >= m23;
        const float m = t ? m23 : m01, a = t ? a23 : a01;
        if (b1 >= m) { b1 = m; l1 = a; }
    }
    return make_float4(b0, l0, b1, l1);
}

// register budget of a warp role: one warpgroup (4 consecutive warps) grows or shrinks its per-thread register count;
// ptxas allocates the code dominated by the instruction against the new count
template <int N>
__device__ __forceinline__ void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N>
__device__ __forceinline__ void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }

// Chunk merge, interleaved (best,label) planes: fold the per-chunk (best,label) planes in chunk order with the same rule.
__global__ void k_merge_chunks_bl(const float2* __restrict__ BL, int n_chunks, int view, int rows, int w, int pitchS,
                                  float* __restrict__ best, float* __restrict__ disp) {
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    int y = blockIdx.y;
    if (x >= w) return;
    const size_t plane = (size_t)rows * pitchS;
    float b = BEST_INIT_BITS_F, l = 0.0f;
    for (int c = 0; c < n_chunks; c++) {
        float2 q = BL[(size_t)(c * 2 + view) * plane + (size_t)y * pitchS + x];
        if (b >= q.x) { b = q.x; l = q.y; }
    }
    if (best) best[(size_t)y * w + x] = b;
    if (disp) disp[(size_t)y * w + x] = l;
}

// ---------------------------------------------------------------------------------------
// Chunk merge: fold the per-chunk (best,label) planes in chunk order with the same rule.
__global__ void k_merge_chunks(const float* __restrict__ bestS, const float* __restrict__ labS, int n_chunks, int view,
                               int rows, int w, int pitchS, float* __restrict__ best, float* __restrict__ disp) {
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    int y = blockIdx.y;
    if (x >= w) return;
    const size_t plane = (size_t)rows * pitchS;
    float b = BEST_INIT_BITS_F, l = 0.0f;
    for (int c = 0; c < n_chunks; c++) {
        size_t off = (size_t)(c * 2 + view) * plane + (size_t)y * pitchS + x;
        float q = bestS[off];
        if (b >= q) { b = q; l = labS[off]; }
    }
    if (best) best[(size_t)y * w + x] = b;
    if (disp) disp[(size_t)y * w + x] = l;
}

// Find integers (nI, nG, S) with (1-alpha) ~= nI/S and alpha/2 ~= nG/S (relative 1e-6) such
// that the lattice cost P = nI*cI + nG*cG stays exactly representable through both box sums.
bool find_lattice(const sb200_params* p, int* nI, int* nG, int* S) {
    const double wI = (double)(1.0f - p->alpha), wG = (double)p->alpha / 2.0;
    const double tc = p->th_color, tg2 = 2.0 * p->th_grad;
    if (tc != (double)(int)tc || tg2 != (double)(int)tg2 || tc < 0 || tg2 < 0 || tc > 255 || tg2 > 1020) return false;
    for (int s = 1; s <= 4096; s++) {
        double a = wI * s, b = wG * s;
        long ia = lround(a), ib = lround(b);
        if (ia < 0 || ib < 0 || (ia == 0 && ib == 0)) continue;
        if (fabs(a - ia) > 1e-6 * fmax(a, 1e-3) * 1.0 + 1e-9 * s && fabs(a - ia) > 2e-6 * a) continue;
        if (fabs(b - ib) > 2e-6 * b && fabs(b - ib) > 1e-9 * s) continue;
        double pmax = ia * tc + ib * tg2;
        // half-exact lattice values and fp32-exact 361-element sums of I*P
        if (pmax > 2047.0 || 361.0 * 255.0 * pmax >= 16777216.0) return false;
        *nI = (int)ia;
        *nG = (int)ib;
        *S = s;
        return true;
    }
    return false;
}

struct Plan {
    int n_strips, n_bands, band_rows, n_chunks, chunk_d;
};

// Tile (strip x band x chunk x 2 views) so that the block count is close to a multiple of the
// SM count with as little warm-up (36 extra rows per band) and group padding as possible.
Plan make_plan(int w, int rows_out, int size_d, int sm_count, int n_views) {
    Plan best{};
    double best_cost = 1e300;
    const int n_strips = (w + VALID_W - 1) / VALID_W;
    const int groups = (size_d + NWARP - 1) / NWARP;
    for (int n_chunks = 1; n_chunks <= groups; n_chunks++) {
        int gpc = (groups + n_chunks - 1) / n_chunks;  // groups per chunk
        int real_chunks = (groups + gpc - 1) / gpc;
        if (real_chunks != n_chunks) continue;
        for (int n_bands = 1; n_bands <= 64; n_bands++) {
            int band_rows = (rows_out + n_bands - 1) / n_bands;
            if (n_bands > 1 && band_rows < 64) break;
            int real_bands = (rows_out + band_rows - 1) / band_rows;
            if (real_bands != n_bands) continue;
            long blocks = (long)n_strips * n_bands * n_chunks * n_views;
            long waves = (blocks + sm_count - 1) / sm_count;
            // time ~ waves * per-block work; per-block work ~ groups/chunk * (band_rows + 36)
            double t = (double)waves * gpc * (band_rows + 4.0 * RAD) + 0.02 * n_chunks * rows_out / 64.0;
            if (t < best_cost) {
                best_cost = t;
                best = Plan{n_strips, n_bands, band_rows, n_chunks, gpc * NWARP};
            }
        }
    }
    return best;
}

int pad_x(int dabs) { return (dabs + SW + 8 + 7) / 8 * 8; }

}  // namespace
