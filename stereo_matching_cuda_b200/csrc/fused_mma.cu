// fused_mma.cu -- the hot path, second generation: the same fused pipeline as fused_cvf.cu (cost -> guided-filter
// aggregation -> running argmin for both views, the D-deep volume never leaves the SM), with the four HORIZONTAL
// 19-column window sums per cell taken by the tensor cores as a product with a 0/1 band matrix.
//
// Replaces, per disparity slice, the same reference chain as fused_cvf.cu:
//   costVolumOnGPU2 (costVolume.cu:163-190) -> pixelMultOnGPU -> 4x { rowSum + colSum (integral.cu:78-131) +
//   computeBoxFilterOnGPU (guidedFilter.cu:297-318) } -> compute_ak_and_bk (:345-354) -> compute_q (:363-369) ->
//   dispSelectOnGPU (:403-411).
//
// Why: ncu put fused_cvf.cu's wall at the L1TEX data pipe (72 %), 43 % of it the 16 warp shuffles per 8 pixels of every
// horizontal window sum, and 21 of its 67 instructions per cell were those sums.  A box sum along x is the product
// H = Band x V with Band[m][k] = 1 for m <= k <= m+18: tcgen05.mma does it off the SM's issue slots and data pipe.
//
// Decomposition.  A block owns a strip of 128 columns (110 valid after the two cascaded radius-9 filters) and a group
// of 8 consecutive disparities, and marches down the rows two at a time.  Tensor Memory lane = image column, so after
// an MMA every thread holds ITS column for all 8 disparities and 2 rows: guide operands are loaded once per column and
// row (not per disparity), and the winner-take-all over the 8 disparities is a register tournament.
//   role A (10 warps)  lattice cost P (packed half, 2 disparities per instruction) for 160 cost columns x 8 d x 2 rows;
//                      EXACT fp16 pieces of P and I*P:  P, (I&15)*P, (I&240)*P;  what enters the MMA is the vertical
//                      difference  piece(y) - piece(y-19)  (exact in fp16), the 19-row ring of P lives in shared memory
//   MMA 1 (tensor)     dH[x][n] = sum_k Band[x][k] * dPiece[k][n]   (K = 160, fp32 accumulate: exact integers)
//   role B (4 warps)   S_p += dH_p, S_Ip += dH_Ip  (the 2-D box sums, exact);  a, b (guidedFilter.cu:345-354);
//                      split into fp16 hi + lo (22 significant bits), scaled by a power of two
//   MMA 2 (tensor)     H_a, H_b = Band x (hi) + Band x (lo)
//   role C (4 warps)   vertical running sums of H with a 19-slot ring in Tensor Memory (re-summed every 16 iterations,
//                      which bounds the float drift), q = mean_a*I + mean_b, tournament over the 8 disparities with the
//                      reference's `best >= q` rule, read-modify-write of the chunk's (best,label) plane: once per
//                      8 disparities (fused_cvf.cu: once per 4)
//   1 warp issues the MMAs (warp-uniform descriptors, one elected lane), 1 warp runs the TMA producer.
// Operands come from strip-tiled planes written by three small preparation kernels (k_prep_ga/gb/mt) through
// cp.async.bulk into shared-memory rings; everything between the roles is mbarrier-synchronised, double-buffered.
// Layout, descriptor encodings, exactness and the accumulator's rounding (truncation) were measured first with
// tools/umma_probe.cu on the B200 (profiles/r2_umma_probe.txt).
#include "fused_dev.cuh"

namespace {

constexpr int M_TW = 128;       // a/b and output lanes per strip = Tensor-Memory lanes
constexpr int M_VW = 110;       // valid output columns per strip (128 - 2*RAD)
constexpr int M_KB = 160;       // cost columns per strip = K of MMA 1 (146 used, 10 K-steps)
constexpr int M_KC = M_TW + 2 * RAD;  // 146 cost columns that feed the 128 a/b lanes
constexpr int M_K2 = 128;       // K of MMA 2 (8 K-steps)
constexpr int M_ND = 8;         // disparities per group
constexpr int M_NWA = 10;       // role-A warps: 5 per image row of an iteration
constexpr int M_THREADS = 640;  // 4 B + 4 C + 10 A + MMA + TMA warps
constexpr int M_NGA = 16;       // guide ring: iterations kept in shared memory (history of 11 + prefetch)
constexpr int M_NOP = 4;        // operand ring (a/b statistics + match rows): iterations in flight
constexpr int M_MTC = 42;       // match chunks (4 pixels x 4 shifted copies, 64 B) per row in an operand slot
constexpr int M_RESUM = 16;     // role C re-sums its ring every M_RESUM iterations

constexpr uint32_t GA_ROW = M_KB * 8;             // (I, G, I&15, I&240) as 4 halves per cost column
constexpr uint32_t GB_ROW = M_TW * 8;             // (mean_I, c2 * scale) per a/b lane
constexpr uint32_t MT_ROW = M_MTC * 64;
constexpr uint32_t OP_GB = 0, OP_MT = ROWS * GB_ROW, OP_BYTES = OP_MT + ROWS * MT_ROW;
constexpr uint32_t B1_GROUP = M_KB * 16;          // one N-group (8 columns) of B1: 160 rows x 16 B
constexpr uint32_t B1_BYTES = 6 * B1_GROUP;       // dP (2 rows), d(lo) (2), d(hi) (2)
constexpr uint32_t B2_GROUP = M_K2 * 16;
constexpr uint32_t B2_BYTES = 8 * B2_GROUP;       // hi: (row, d-quad) x 4, lo likewise
constexpr uint32_t PR_SLOT = M_KB * 16;           // P of 8 disparities (halves) per cost column

struct MSmem {
    unsigned char b1[2][B1_BYTES];
    unsigned char b2[2][B2_BYTES];
    unsigned char pring[WIN][PR_SLOT];
    unsigned char ga[M_NGA][ROWS * GA_ROW];
    unsigned char op[M_NOP][OP_BYTES];
    float ry_lut[2][WIN + 1];  // [0][n] = scale/(S*n), [1][n] = 1/(scale*n); [.][0] = 0
    uint64_t ga_full[M_NGA], ga_empty[M_NGA], op_full[M_NOP], op_empty[M_NOP];
    uint64_t b1_full[2], b1_empty[2], d1_full[2], d1_empty[2], b2_full[2], b2_empty[2], d2_full[2], d2_empty[2];
    uint32_t tmem_base;
};
static_assert(sizeof(MSmem) <= 227 * 1024, "shared memory budget");

// Tensor-Memory column map of a lane (512 columns)
constexpr uint32_t TC_BAND = 0;     // 80 columns: Band[lane][0..159] as packed halves
constexpr uint32_t TC_D1 = 80;      // 2 x 32: dH_p (row, d) | dH_Ip (row, d)
constexpr uint32_t TC_D2 = 144;     // 2 x 32: (row, d, {a,b})
constexpr uint32_t TC_RING = 208;   // 19 x 16: (H_a, H_b) x 8 d of the last 19 rows

struct MmaArgs {
    const uint2* GA[2];    // per IMAGE: [strip][rows_pad][160] (I, G, I&15, I&240) halves; pad (1024,1024,0,0)
    const float2* GB[2];   // per IMAGE: [strip][rows_pad][128] (mean_I, scale * c/(S*area)); 0 outside
    const uint4* MT[2];    // per IMAGE: [rows_pad][n_chunk][4] chunk i, copy s: I halves of X = 4i+s..+3, then G halves
    int rows_pad, n_chunk, padm;
    int w;
    int y_out0, rows_out, y_global0, frame_h;
    int dmin[2], size_d;
    int n_strips, n_bands, band_rows, n_chunks, chunk_d, n_views;
    float2* BL;
    int pitchS;
    float S, scale, inv_scale;
    unsigned wI2, wG2, tc2, tg2;  // half2 broadcasts: lattice weights nI, nG; thresholds th_color, 2*th_grad
};

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
__device__ __forceinline__ void tm_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
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
__device__ __forceinline__ void tm_st16u(uint32_t taddr, const uint32_t* r) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
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
__device__ __forceinline__ float2 ld_early_f2(const void* p) {  // coherent: written by this thread one group earlier
    float2 v;
    asm volatile("ld.global.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p) : "memory");
    return v;
}
// f16 + f32 -> f32 with the half taken from either half of a packed pair
__device__ __forceinline__ float fhadd_lo(unsigned h2, float c) { return fhadd(__low2half(u2h2(h2)), c); }
__device__ __forceinline__ float fhadd_hi(unsigned h2, float c) { return fhadd(__high2half(u2h2(h2)), c); }

// Register budgets of the warpgroups.  setmaxnreg only moves registers inside the pool the block got at launch: 640 threads
// x 96 registers (what __launch_bounds__(640, 1) lets ptxas use) = 61440, NOT the whole 64 K file -- budgets that add up
// to more leave the last setmaxnreg.inc waiting forever.
constexpr int M_REGS_LAUNCH = 96;
constexpr int M_REGS_B = 160, M_REGS_C = 152, M_REGS_A = 56;
static_assert(128 * M_REGS_B + 128 * M_REGS_C + 384 * M_REGS_A <= M_THREADS * M_REGS_LAUNCH, "register pool of the block");

__global__ void __launch_bounds__(M_THREADS, 1) k_fused_mma(const MmaArgs A) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    MSmem& sm = *reinterpret_cast<MSmem*>(smem_raw);
    const int lane = threadIdx.x & 31;
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);  // warp-uniform for the compiler

    int bid = blockIdx.x;
    const int view = bid % A.n_views;
    bid /= A.n_views;
    const int chunk = bid % A.n_chunks;
    bid /= A.n_chunks;
    const int band = bid % A.n_bands;
    const int strip = bid / A.n_bands;

    const int xc0 = strip * M_VW - 2 * RAD;  // frame column of cost column k = 0
    const int xa0 = strip * M_VW - RAD;      // frame column of a/b lane 0
    const int xo0 = strip * M_VW;            // frame column of output lane 0
    const int dlo = A.dmin[view] + chunk * A.chunk_d;
    const int dcnt = min(A.chunk_d, A.size_d - chunk * A.chunk_d);
    const int ngroups = (dcnt + M_ND - 1) / M_ND;
    const int yb0 = A.y_out0 + band * A.band_rows;
    const int yb1 = min(yb0 + A.band_rows, A.y_out0 + A.rows_out);
    const int y_first = yb0 - 2 * RAD;
    const int niter = ((yb1 - yb0) + 4 * RAD + ROWS - 1) / ROWS;
    constexpr int WARM_IT = 4 * RAD / ROWS;
    const int Ktotal = ngroups * niter;

    // ---- set-up: Tensor Memory, barriers, tables, a finite guide ring ----
    if (warp == 0) tm_alloc(&sm.tmem_base);
    if (threadIdx.x == 32) {
        for (int i = 0; i < M_NGA; i++) {
            mbar_init(smem_addr(&sm.ga_full[i]), 1);
            mbar_init(smem_addr(&sm.ga_empty[i]), M_NWA + 4);  // role A at K+10, role C at K+9
        }
        for (int i = 0; i < M_NOP; i++) {
            mbar_init(smem_addr(&sm.op_full[i]), 1);
            mbar_init(smem_addr(&sm.op_empty[i]), M_NWA + 4);  // role A (match rows), role B (statistics)
        }
        for (int i = 0; i < 2; i++) {
            mbar_init(smem_addr(&sm.b1_full[i]), M_NWA);
            mbar_init(smem_addr(&sm.b1_empty[i]), 1);
            mbar_init(smem_addr(&sm.d1_full[i]), 1);
            mbar_init(smem_addr(&sm.d1_empty[i]), 4);
            mbar_init(smem_addr(&sm.b2_full[i]), 4);
            mbar_init(smem_addr(&sm.b2_empty[i]), 1);
            mbar_init(smem_addr(&sm.d2_full[i]), 1);
            mbar_init(smem_addr(&sm.d2_empty[i]), 4);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x >= 64 && threadIdx.x < 64 + 2 * (WIN + 1)) {
        const int t = threadIdx.x - 64, n = t % (WIN + 1);
        float v = 0.0f;
        if (n > 0) v = (t < WIN + 1) ? A.scale * __frcp_rn(A.S * (float)n) : A.inv_scale * __frcp_rn((float)n);
        sm.ry_lut[t / (WIN + 1)][n] = v;
    }
    {   // the guide ring is read 10 iterations back before those rows exist (times a zero cost): keep it finite;
        // cost columns 146..159 of B1 and everything else start from zero as well
        uint4* z = reinterpret_cast<uint4*>(smem_raw);
        const int n16 = (int)(offsetof(MSmem, ry_lut) / 16);
        for (int i = threadIdx.x; i < n16; i += M_THREADS) z[i] = make_uint4(0u, 0u, 0u, 0u);
    }
    fence_async_smem();
    tm_fence_before();
    __syncthreads();
    tm_fence_after();
    const uint32_t tmem = sm.tmem_base;

    auto bar = [&](const uint64_t* b) { return smem_addr(b); };

    if (warp < 4) {
        // ================= role B: 2-D box sums of P and I*P, a and b, fp16 hi/lo split =================
        reg_inc<M_REGS_B>();
        const int l = warp * 32 + lane;
        const uint32_t tl = tmem + ((uint32_t)(warp * 32) << 16);
        {   // this lane's row of the band matrix: Band[l][k] = 1 for l <= k <= l + 18
            for (int j0 = 0; j0 < M_KB / 2; j0 += 16) {
                uint32_t v[16];
#pragma unroll
                for (int j = 0; j < 16; j++) {
                    const int k0 = 2 * (j0 + j), k1 = k0 + 1;
                    const float e0 = (k0 >= l && k0 <= l + 2 * RAD) ? 1.0f : 0.0f;
                    const float e1 = (k1 >= l && k1 <= l + 2 * RAD) ? 1.0f : 0.0f;
                    v[j] = h22u(__floats2half2_rn(e0, e1));
                }
                tm_st16u(tl + TC_BAND + j0, v);
            }
            tm_wait_st();
            tm_fence_before();
        }
        named_bar_arrive(1, 128 + 32);  // the MMA warp waits for the band
        const int x = xa0 + l;
        const bool xin = x >= 0 && x < A.w;
        const float rx = xin ? __frcp_rn((float)(min(A.w - 1, x + RAD) - max(0, x - RAD) + 1)) : 0.0f;
        const uint32_t b2_lane = (uint32_t)((l >> 3) * 128 + (l & 7) * 16);
        for (int g = 0; g < ngroups; g++) {
            float Sp[M_ND], Sip[M_ND];
#pragma unroll
            for (int d = 0; d < M_ND; d++) Sp[d] = Sip[d] = 0.0f;
#pragma unroll 1
            for (int it = 0; it < niter; it++) {
                const int K = g * niter + it;
                const int yi0 = y_first + it * ROWS;
                mbar_wait(bar(&sm.d1_full[K & 1]), (unsigned)(K >> 1) & 1u);
                tm_fence_after();
                uint32_t dh[32];
                tm_ld32(tl + TC_D1 + 32 * (K & 1), dh);
                mbar_wait(bar(&sm.op_full[K & (M_NOP - 1)]), (unsigned)(K / M_NOP) & 1u);
                const uint32_t opa = smem_addr(&sm.op[K & (M_NOP - 1)][0]) + OP_GB + (uint32_t)l * 8;
                uint2 st[ROWS];
#pragma unroll
                for (int r = 0; r < ROWS; r++) st[r] = lds64(opa + r * GB_ROW);
                tm_wait_ld();
                tm_fence_before();
                __syncwarp();
                mbar_arrive_lane0(bar(&sm.d1_empty[K & 1]), lane);
                uint32_t hi[ROWS][M_ND], lo[ROWS][M_ND];
#pragma unroll
                for (int r = 0; r < ROWS; r++) {
                    const float mI = __uint_as_float(st[r].x), c2 = __uint_as_float(st[r].y);
                    const float r1 = rx * inv_rows(sm.ry_lut[0], yi0 + r - RAD, A.y_global0, A.frame_h);
#pragma unroll
                    for (int d = 0; d < M_ND; d++) {
                        Sp[d] += __uint_as_float(dh[r * 8 + d]);
                        Sip[d] += __uint_as_float(dh[16 + r * 8 + d]);
                        const float cov = fmaf(-mI, Sp[d], Sip[d]);
                        const float a = cov * c2;
                        const float b = fmaf(-mI, a, Sp[d] * r1);
                        const unsigned h2 = h22u(__floats2half2_rn(a, b));
                        // hi - value = -(lo part); MMA 2 takes the lo pass with B negated
                        const float na = fhadd_lo(h2, -a), nb = fhadd_hi(h2, -b);
                        hi[r][d] = h2;
                        lo[r][d] = h22u(__floats2half2_rn(na, nb));
                    }
                }
                if (K >= 2) mbar_wait(bar(&sm.b2_empty[K & 1]), (unsigned)((K >> 1) - 1) & 1u);
                const uint32_t b2a = smem_addr(&sm.b2[K & 1][0]) + b2_lane;
#pragma unroll
                for (int r = 0; r < ROWS; r++) {
#pragma unroll
                    for (int q4 = 0; q4 < 2; q4++) {
                        sts128(b2a + (uint32_t)(r * 2 + q4) * B2_GROUP, hi[r][4 * q4], hi[r][4 * q4 + 1], hi[r][4 * q4 + 2],
                               hi[r][4 * q4 + 3]);
                        sts128(b2a + (uint32_t)(4 + r * 2 + q4) * B2_GROUP, lo[r][4 * q4], lo[r][4 * q4 + 1],
                               lo[r][4 * q4 + 2], lo[r][4 * q4 + 3]);
                    }
                }
                fence_async_smem();
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive(bar(&sm.b2_full[K & 1]));
                    mbar_arrive(bar(&sm.op_empty[K & (M_NOP - 1)]));
                }
            }
        }
    } else if (warp < 8) {
        // ================= role C: vertical sums of H_a, H_b, q, winner-take-all =================
        reg_inc<M_REGS_C>();
        const int w4 = warp - 4;
        const int l = w4 * 32 + lane;
        const uint32_t tl = tmem + ((uint32_t)(w4 * 32) << 16);
        const int x = xo0 + l;
        const bool valid = l < M_VW && x < A.w;
        const float rx = valid ? __frcp_rn((float)(min(A.w - 1, x + RAD) - max(0, x - RAD) + 1)) : 0.0f;
        const size_t planeS = (size_t)A.rows_out * A.pitchS;
        float2* const bl0 = A.BL + (size_t)(chunk * 2 + view) * planeS + (size_t)(yb0 - A.y_out0) * A.pitchS + x;
        const int band_rows = yb1 - yb0;
        const uint32_t ga_I = smem_addr(&sm.ga[0][0]) + (uint32_t)(l + 2 * RAD) * 8;  // I of this thread's column in a guide row
        for (int g = 0; g < ngroups; g++) {
            float Sa[M_ND], Sb[M_ND];
#pragma unroll
            for (int d = 0; d < M_ND; d++) Sa[d] = Sb[d] = 0.0f;
            {
                uint32_t z[16];
#pragma unroll
                for (int i = 0; i < 16; i++) z[i] = 0u;
                for (int s = 0; s < WIN; s++) tm_st16u(tl + TC_RING + 16 * s, z);
                tm_wait_st();
            }
            const int dbase = dlo + g * M_ND;
            const int dact = min(M_ND, dcnt - g * M_ND);  // disparities of this group that exist
            const bool ld_ok = (g > 0) && valid;
            auto prefetch = [&](int e, float2 (&pb)[ROWS]) {
#pragma unroll
                for (int r = 0; r < ROWS; r++) {
                    pb[r] = make_float2(BEST_INIT_BITS_F, 0.0f);
                    if (ld_ok && e >= 0 && e * ROWS + r < band_rows) pb[r] = ld_early_f2(bl0 + (size_t)(e * ROWS + r) * A.pitchS);
                }
            };
            float2 pbA[ROWS], pbB[ROWS];
            prefetch(0, pbA);
            prefetch(1, pbB);
            int slot = 0;
#pragma unroll 1
            for (int it = 0; it < niter; it++) {
                const int K = g * niter + it;
                const int e = it - WARM_IT;
                float2 pb[ROWS];
                if (e >= 0) {
#pragma unroll
                    for (int r = 0; r < ROWS; r++) {
                        pb[r] = pbA[r];
                        pbA[r] = pbB[r];
                    }
                    prefetch(e + 2, pbB);
                }
                int slots[ROWS];
#pragma unroll
                for (int r = 0; r < ROWS; r++) {
                    slots[r] = slot;
                    slot = (slot + 1 == WIN) ? 0 : slot + 1;
                }
                mbar_wait(bar(&sm.d2_full[K & 1]), (unsigned)(K >> 1) & 1u);
                tm_fence_after();
                uint32_t h[32], o[ROWS][16];
                tm_ld32(tl + TC_D2 + 32 * (K & 1), h);
#pragma unroll
                for (int r = 0; r < ROWS; r++) tm_ld16u(tl + TC_RING + 16 * slots[r], o[r]);
                tm_wait_ld();
                tm_fence_before();
                __syncwarp();
                mbar_arrive_lane0(bar(&sm.d2_empty[K & 1]), lane);
#pragma unroll
                for (int r = 0; r < ROWS; r++) tm_st16u(tl + TC_RING + 16 * slots[r], &h[16 * r]);
                // I of the output rows: guide rows of iteration K - 9
                uint32_t iw[ROWS];
                if (e >= 0) {
                    const uint32_t ga = ga_I + (uint32_t)((K - 9) & (M_NGA - 1)) * (ROWS * GA_ROW);
#pragma unroll
                    for (int r = 0; r < ROWS; r++) iw[r] = lds32(ga + r * GA_ROW);
                }
#pragma unroll
                for (int r = 0; r < ROWS; r++) {
#pragma unroll
                    for (int d = 0; d < M_ND; d++) {
                        Sa[d] += __uint_as_float(h[16 * r + 2 * d]) - __uint_as_float(o[r][2 * d]);
                        Sb[d] += __uint_as_float(h[16 * r + 2 * d + 1]) - __uint_as_float(o[r][2 * d + 1]);
                    }
                    if (e >= 0) {
                        const int yq = yb0 + e * ROWS + r;
                        const float rxy = rx * inv_rows(sm.ry_lut[1], yq, A.y_global0, A.frame_h);
                        const float I = __low2float(u2h2(iw[r]));
                        float q[M_ND];
#pragma unroll
                        for (int d = 0; d < M_ND; d++) {
                            q[d] = fmaf(Sa[d], I, Sb[d]) * rxy;
                            if (d >= dact) q[d] = __int_as_float(0x7f800000);
                        }
                        // ascending d, `best >= q`: minimum, the later index on a tie (guidedFilter.cu:406), as a tournament
                        float m1[4], a1[4];
#pragma unroll
                        for (int i = 0; i < 4; i++) {
                            const bool t = q[2 * i] >= q[2 * i + 1];
                            m1[i] = t ? q[2 * i + 1] : q[2 * i];
                            a1[i] = t ? (float)(2 * i + 1) : (float)(2 * i);
                        }
                        const bool t01 = m1[0] >= m1[1], t23 = m1[2] >= m1[3];
                        const float m01 = t01 ? m1[1] : m1[0], a01 = t01 ? a1[1] : a1[0];
                        const float m23 = t23 ? m1[3] : m1[2], a23 = t23 ? a1[3] : a1[2];
                        const bool tt = m01 >= m23;
                        const float m = tt ? m23 : m01, am = tt ? a23 : a01;
                        float2 nb = pb[r];
                        if (nb.x >= m) {
                            nb.x = m;
                            nb.y = am + (float)dbase;
                        }
                        if (valid && e * ROWS + r < band_rows) bl0[(size_t)(e * ROWS + r) * A.pitchS] = nb;
                    }
                }
                tm_wait_st();
                if ((it & (M_RESUM - 1)) == M_RESUM - 1) {
                    // re-sum the ring: bounds the rounding drift of the running sums to M_RESUM iterations
#pragma unroll
                    for (int d = 0; d < M_ND; d++) Sa[d] = Sb[d] = 0.0f;
                    for (int s = 0; s < WIN; s++) {
                        uint32_t v[16];
                        tm_ld16u(tl + TC_RING + 16 * s, v);
                        tm_wait_ld();
#pragma unroll
                        for (int d = 0; d < M_ND; d++) {
                            Sa[d] += __uint_as_float(v[2 * d]);
                            Sb[d] += __uint_as_float(v[2 * d + 1]);
                        }
                    }
                }
                if (K >= 9) {
                    __syncwarp();
                    mbar_arrive_lane0(bar(&sm.ga_empty[(K - 9) & (M_NGA - 1)]), lane);
                }
            }
        }
    } else {
    reg_dec<M_REGS_A>();  // warpgroups 2..4 (role A, MMA issue, TMA producer) give their registers to roles B and C
    if (warp < 8 + M_NWA) {
        // ================= role A: lattice cost, exact fp16 pieces, vertical differences =================
        const int wa = warp - 8;
        const int r = wa / 5;                   // image row of the iteration this warp works on
        const int k = (wa % 5) * 32 + lane;     // cost column
        const int x = xc0 + k;
        const bool used = k < M_KC && x >= 0 && x < A.w;
        const __half2 wI = used ? u2h2(A.wI2) : __float2half2_rn(0.0f);
        const __half2 wG = used ? u2h2(A.wG2) : __float2half2_rn(0.0f);
        const __half2 tc = u2h2(A.tc2), tg = u2h2(A.tg2);
        const uint32_t b1_lane = (uint32_t)((k >> 3) * 128 + (k & 7) * 16);
        const uint32_t pr0 = smem_addr(&sm.pring[0][0]) + (uint32_t)k * 16;
        const uint32_t ga0 = smem_addr(&sm.ga[0][0]) + (uint32_t)k * 8;
        for (int g = 0; g < ngroups; g++) {
            // ring of P: zero at group start (the two threads of a column -- one per row of an iteration -- share it, so
            // both must have left the previous group before it is cleared, and see it cleared before they go on)
            named_bar_sync(2, M_NWA * 32);
            for (int s = r; s < WIN; s += 2) sts128(pr0 + (uint32_t)s * PR_SLOT, 0u, 0u, 0u, 0u);
            named_bar_sync(2, M_NWA * 32);
            // match operands of this thread: padded column X0 = x + d0 + padm, 8 consecutive pixels from copy X0 & 3
            const int d0 = dlo + g * M_ND;
            const int X0 = x + d0 + A.padm;
            const int i0 = (xc0 + d0 + A.padm) >> 2;  // first chunk of the slot (warp-uniform)
            const uint32_t mt_off = OP_MT + (uint32_t)r * MT_ROW + (uint32_t)((X0 >> 2) - i0) * 64 + (uint32_t)(X0 & 3) * 16;
            int slot = r;  // ring slot of row (2*it + r) mod 19
#pragma unroll 1
            for (int it = 0; it < niter; it++) {
                const int K = g * niter + it;
                mbar_wait(bar(&sm.ga_full[K & (M_NGA - 1)]), (unsigned)(K / M_NGA) & 1u);
                mbar_wait(bar(&sm.op_full[K & (M_NOP - 1)]), (unsigned)(K / M_NOP) & 1u);
                // guide (I, G | I&15, I&240) of the entering row; (I&15, I&240) of the row that leaves (19 rows up):
                // row r = 0: second row of iteration K-10, r = 1: first row of iteration K-9
                const uint2 gn = lds64(ga0 + (uint32_t)(K & (M_NGA - 1)) * (ROWS * GA_ROW) + (uint32_t)r * GA_ROW);
                const uint32_t go = lds32(ga0 + (uint32_t)((K - 10 + r) & (M_NGA - 1)) * (ROWS * GA_ROW) +
                                          (uint32_t)(1 - r) * GA_ROW + 4);
                const uint32_t opa = smem_addr(&sm.op[K & (M_NOP - 1)][0]) + mt_off;
                const uint4 m0 = lds128(opa), m1 = lds128(opa + 64);
                const uint4 po = lds128(pr0 + (uint32_t)slot * PR_SLOT);
                const __half2 gI = __low2half2(u2h2(gn.x)), gG = __high2half2(u2h2(gn.x));
                const __half2 gl = __low2half2(u2h2(gn.y)), gh = __high2half2(u2h2(gn.y));
                const __half2 ol = __hneg2(__low2half2(u2h2(go))), oh = __hneg2(__high2half2(u2h2(go)));
                const unsigned mi[4] = {m0.x, m0.y, m1.x, m1.y}, mg[4] = {m0.z, m0.w, m1.z, m1.w};
                const unsigned pold[4] = {po.x, po.y, po.z, po.w};
                unsigned pn[4], dp[4], dl[4], dh[4];
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const __half2 cI = __hmin2(__habs2(__hsub2(u2h2(mi[j]), gI)), tc);
                    const __half2 cG = __hmin2(__habs2(__hsub2(u2h2(mg[j]), gG)), tg);
                    const __half2 P = __hfma2(cG, wG, __hmul2(cI, wI));
                    const __half2 Po = u2h2(pold[j]);
                    pn[j] = h22u(P);
                    dp[j] = h22u(__hsub2(P, Po));
                    dl[j] = h22u(__hfma2(Po, ol, __hmul2(P, gl)));
                    dh[j] = h22u(__hfma2(Po, oh, __hmul2(P, gh)));
                }
                sts128(pr0 + (uint32_t)slot * PR_SLOT, pn[0], pn[1], pn[2], pn[3]);
                slot += 2;
                if (slot >= WIN) slot -= WIN;
                if (K >= 2) mbar_wait(bar(&sm.b1_empty[K & 1]), (unsigned)((K >> 1) - 1) & 1u);
                const uint32_t b1a = smem_addr(&sm.b1[K & 1][0]) + b1_lane + (uint32_t)r * B1_GROUP;
                sts128(b1a, dp[0], dp[1], dp[2], dp[3]);
                sts128(b1a + 2 * B1_GROUP, dl[0], dl[1], dl[2], dl[3]);
                sts128(b1a + 4 * B1_GROUP, dh[0], dh[1], dh[2], dh[3]);
                fence_async_smem();
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive(bar(&sm.b1_full[K & 1]));
                    mbar_arrive(bar(&sm.op_empty[K & (M_NOP - 1)]));
                    if (K >= 10) mbar_arrive(bar(&sm.ga_empty[(K - 10) & (M_NGA - 1)]));
                }
            }
        }
    } else if (warp == 8 + M_NWA) {
        // ================= MMA issue: warp-uniform descriptors, one elected lane =================
        named_bar_sync(1, 128 + 32);  // the band matrix is in Tensor Memory
        tm_fence_after();
        constexpr uint32_t ID32 = instr_desc(32, 0), ID16 = instr_desc(16, 0), ID32N = instr_desc(32, 1);
        const uint64_t db1 = smem_desc(smem_addr(&sm.b1[0][0]), 128, B1_GROUP);
        const uint64_t db2 = smem_desc(smem_addr(&sm.b2[0][0]), 128, B2_GROUP);
        const uint32_t ta = tmem + TC_BAND;
        auto stage2 = [&](int K) {
            mbar_wait(bar(&sm.b2_full[K & 1]), (unsigned)(K >> 1) & 1u);
            if (K >= 2) mbar_wait(bar(&sm.d2_empty[K & 1]), (unsigned)((K >> 1) - 1) & 1u);
            tm_fence_after();
            if (elect_one()) {
                const uint64_t dh = db2 + (uint64_t)((K & 1) * (B2_BYTES >> 4));
                const uint64_t dl = dh + (uint64_t)((4 * B2_GROUP) >> 4);
                const uint32_t td = tmem + TC_D2 + 32 * (K & 1);
#pragma unroll
                for (int j = 0; j < M_K2 / 16; j++) umma_ts(td, ta + 8 * j, dh + (uint64_t)(j * 16), ID32, j > 0);
#pragma unroll
                for (int j = 0; j < M_K2 / 16; j++) umma_ts(td, ta + 8 * j, dl + (uint64_t)(j * 16), ID32N, 1);
                umma_commit(bar(&sm.d2_full[K & 1]));
                umma_commit(bar(&sm.b2_empty[K & 1]));
            }
            __syncwarp();
        };
#pragma unroll 1
        for (int K = 0; K < Ktotal; K++) {
            mbar_wait(bar(&sm.b1_full[K & 1]), (unsigned)(K >> 1) & 1u);
            if (K >= 2) mbar_wait(bar(&sm.d1_empty[K & 1]), (unsigned)((K >> 1) - 1) & 1u);
            tm_fence_after();
            if (elect_one()) {
                const uint64_t dp = db1 + (uint64_t)((K & 1) * (B1_BYTES >> 4));
                const uint64_t dhi = dp + (uint64_t)((4 * B1_GROUP) >> 4);
                const uint32_t td = tmem + TC_D1 + 32 * (K & 1);
#pragma unroll
                for (int j = 0; j < M_KB / 16; j++) umma_ts(td, ta + 8 * j, dp + (uint64_t)(j * 16), ID32, j > 0);
#pragma unroll
                for (int j = 0; j < M_KB / 16; j++) umma_ts(td + 16, ta + 8 * j, dhi + (uint64_t)(j * 16), ID16, 1);
                umma_commit(bar(&sm.d1_full[K & 1]));
                umma_commit(bar(&sm.b1_empty[K & 1]));
            }
            __syncwarp();
            if (K >= 1) stage2(K - 1);
        }
        if (Ktotal > 0) stage2(Ktotal - 1);
    } else {
        // ================= TMA producer =================
        if (lane == 0) {
            const uint2* GAp = A.GA[view] + (size_t)strip * A.rows_pad * M_KB;
            const float2* GBp = A.GB[view] + (size_t)strip * A.rows_pad * M_TW;
            const uint4* MTp = A.MT[1 - view];
            int K = 0;
            for (int g = 0; g < ngroups; g++) {
                const int d0 = dlo + g * M_ND;
                const int i0 = (xc0 + d0 + A.padm) >> 2;
                for (int it = 0; it < niter; it++, K++) {
                    const long long row = (long long)PADY + y_first + it * ROWS;  // padded row of the entering rows
                    {
                        const int sl = K & (M_NGA - 1);
                        if (K >= M_NGA) mbar_wait(bar(&sm.ga_empty[sl]), (unsigned)(K / M_NGA - 1) & 1u);
                        const uint32_t full = bar(&sm.ga_full[sl]);
                        mbar_expect_tx(full, ROWS * GA_ROW);
                        bulk_g2s(smem_addr(&sm.ga[sl][0]), GAp + row * M_KB, ROWS * GA_ROW, full);
                    }
                    {
                        const int sl = K & (M_NOP - 1);
                        if (K >= M_NOP) mbar_wait(bar(&sm.op_empty[sl]), (unsigned)(K / M_NOP - 1) & 1u);
                        const uint32_t full = bar(&sm.op_full[sl]);
                        const uint32_t dst = smem_addr(&sm.op[sl][0]);
                        mbar_expect_tx(full, OP_BYTES);
                        bulk_g2s(dst + OP_GB, GBp + (row - RAD) * M_TW, ROWS * GB_ROW, full);
#pragma unroll
                        for (int r = 0; r < ROWS; r++)
                            bulk_g2s(dst + OP_MT + r * MT_ROW, MTp + ((row + r) * A.n_chunk + i0) * 4, MT_ROW, full);
                    }
                }
            }
        }
    }
    }
    tm_fence_before();
    __syncthreads();
    if (warp == 0) {
        tm_fence_after();
        tm_dealloc(tmem);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Preparation: three streaming kernels per image (replace chToFlOnGPU, x_derivativeOnGPU and the guide statistics of
// compute_guided_filter, guidedFilter.cu:58-123, like k_prep of fused_cvf.cu).
struct PrepM {
    const uint8_t* gray;  // held rows, pitch w
    int w, h_held, y_global0, frame_h;
    int n_strips, rows_pad;
    uint2* GA;
    float2* GB;
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

// GA: thread per (strip, padded row, cost column)
__global__ void __launch_bounds__(160) k_prep_ga(const PrepM P) {
    const int k = threadIdx.x, yrow = blockIdx.x, strip = blockIdx.y;
    const int x = strip * M_VW - 2 * RAD + k, y = yrow - PADY;
    float I, G;
    pix_ig(P, x, y, I, G);
    const bool in = in_frame(P, x, y);
    const int ii = in ? (int)I : 0;
    uint2 v;
    v.x = h22u(__floats2half2_rn(I, G));
    v.y = h22u(__floats2half2_rn((float)(ii & 15), (float)(ii & 240)));
    P.GA[((size_t)strip * P.rows_pad + yrow) * M_KB + k] = v;
}

// MT: thread per (padded row, chunk, copy)
__global__ void __launch_bounds__(256) k_prep_mt(const PrepM P) {
    const int idx = blockIdx.x * 256 + threadIdx.x;
    const int yrow = blockIdx.y;
    if (idx >= P.n_chunk * 4) return;
    const int i = idx >> 2, s = idx & 3;
    const int y = yrow - PADY;
    float I[4], G[4];
#pragma unroll
    for (int j = 0; j < 4; j++) pix_ig(P, 4 * i + s + j - P.padm, y, I[j], G[j]);
    uint4 v;
    v.x = h22u(__floats2half2_rn(I[0], I[1]));
    v.y = h22u(__floats2half2_rn(I[2], I[3]));
    v.z = h22u(__floats2half2_rn(G[0], G[1]));
    v.w = h22u(__floats2half2_rn(G[2], G[3]));
    P.MT[((size_t)yrow * P.n_chunk + i) * 4 + s] = v;
}

// GB: block = (strip, 32 padded rows), thread per a/b lane; exact integer window sums of I and I^2
constexpr int GB_TR = 16;
__global__ void __launch_bounds__(M_TW) k_prep_gb(const PrepM P) {
    __shared__ unsigned short sI[GB_TR + 2 * RAD][M_TW + 2 * RAD + 2];
    __shared__ int h1[GB_TR + 2 * RAD][M_TW], h2[GB_TR + 2 * RAD][M_TW];
    const int l = threadIdx.x, strip = blockIdx.y;
    const int yrow0 = blockIdx.x * GB_TR;
    const int xa0 = strip * M_VW - RAD;
    for (int i = l; i < (GB_TR + 2 * RAD) * (M_TW + 2 * RAD); i += M_TW) {
        const int py = i / (M_TW + 2 * RAD), px = i - py * (M_TW + 2 * RAD);
        const int x = xa0 + px - RAD, y = yrow0 - PADY + py - RAD;
        sI[py][px] = in_frame(P, x, y) ? P.gray[(size_t)y * P.w + x] : 0;
    }
    __syncthreads();
    for (int py = 0; py < GB_TR + 2 * RAD; py++) {
        int s1 = 0, s2 = 0;
#pragma unroll
        for (int t = 0; t < WIN; t++) {
            const int v = sI[py][l + t];
            s1 += v;
            s2 += v * v;
        }
        h1[py][l] = s1;
        h2[py][l] = s2;
    }
    __syncthreads();
    const int x = xa0 + l;
    int v1 = 0, v2 = 0;
    for (int t = 0; t < WIN - 1; t++) {
        v1 += h1[t][l];
        v2 += h2[t][l];
    }
    for (int ty = 0; ty < GB_TR; ty++) {
        v1 += h1[ty + WIN - 1][l];
        v2 += h2[ty + WIN - 1][l];
        const int yrow = yrow0 + ty;
        if (yrow < P.rows_pad) {
            const int y = yrow - PADY, yg = y + P.y_global0;
            float2 out = make_float2(0.0f, 0.0f);
            if (in_frame(P, x, y)) {
                const int ax = min(P.w - 1, x + RAD) - max(0, x - RAD) + 1;
                const int ay = min(P.frame_h - 1, yg + RAD) - max(0, yg - RAD) + 1;
                const float area = (float)(ax * ay);
                const float mI = __fdiv_rn((float)v1, area);
                const float mII = __fdiv_rn((float)v2, area);
                const float var = __fsub_rn(mII, __fmul_rn(mI, mI));
                const float c = (float)(1.0 / ((double)var + P.eps));  // guidedFilter.cu:350
                const float rxy = __fmul_rn(__frcp_rn((float)ax), __frcp_rn(P.S * (float)ay));
                out = make_float2(mI, __fmul_rn(__fmul_rn(c, rxy), P.scale));
                if (P.mean_u8 && l >= RAD && l < RAD + M_VW) {
                    const int m = (int)mI;
                    P.mean_u8[(size_t)y * P.w + x] = (m > 255) ? 255 : (unsigned char)m;
                }
            }
            P.GB[((size_t)strip * P.rows_pad + yrow) * M_TW + l] = out;
        }
        v1 -= h1[ty][l];
        v2 -= h2[ty][l];
    }
}

Plan make_plan_mma(int w, int rows_out, int size_d, int sm_count, int n_views) {
    Plan best{};
    double best_cost = 1e300;
    const int n_strips = (w + M_VW - 1) / M_VW;
    const int groups = (size_d + M_ND - 1) / M_ND;
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
                best = Plan{n_strips, n_bands, band_rows, n_chunks, gpc * M_ND};
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

// the MMA kernel needs, beyond what the shuffle kernel needs, (I & 15) * P and (I & 240) * P exact in fp16
int sbf_mma_supported(const sb200_params* p) {
    int nI, nG, S;
    if (p->radius != RAD || !find_lattice(p, &nI, &nG, &S)) return 0;
    const double pmax = nI * (double)p->th_color + nG * 2.0 * (double)p->th_grad;
    return 15.0 * pmax <= 2047.0;
}

size_t sbf_mma_workspace_bytes(const sb200_ctx* ctx, int w, int h_held, int rows_out, int dabs, int size_d, int n_views) {
    Plan plan = make_plan_mma(w, rows_out, size_d, ctx->sm_count, n_views);
    const int rows_pad = h_held + 2 * PADY;
    const MmaGeom mg = mma_geom(plan.n_strips, -dabs - M_ND, dabs + M_ND);
    const int pitchS = (w + 3) / 4 * 4;
    size_t bytes = 0;
    bytes += 2 * sb_align((size_t)plan.n_strips * rows_pad * M_KB * 8);
    bytes += 2 * sb_align((size_t)plan.n_strips * rows_pad * M_TW * 8);
    bytes += 2 * sb_align((size_t)rows_pad * mg.n_chunk * 64);
    bytes += sb_align((size_t)plan.n_chunks * 2 * rows_out * pitchS * 8);
    return bytes + 4096;
}

int sbf_run_fused_mma(sb200_ctx* ctx, const sb200_params* p, const uint8_t* const gray[2], const SbFusedGeom& g,
                      const int dmin[2], int size_d, int n_views, float* const best[2], float* const disp[2],
                      uint8_t* const mean[2]) {
    int nI, nG, S;
    if (!sbf_mma_supported(p) || !find_lattice(p, &nI, &nG, &S))
        return sb_fail(ctx, SB200_ERR_UNSUPPORTED, "MMA fused kernel: radius 9 and a small exact cost lattice only");
    if (size_d < 1 || g.w < 2 || g.h < 1 || g.rows_out < 1) return sb_fail(ctx, SB200_ERR_INVALID, "fused: bad shape");
    Plan plan = make_plan_mma(g.w, g.rows_out, size_d, ctx->sm_count, n_views);
    const int rows_pad = g.h + 2 * PADY;
    int dabs = 0;
    for (int v = 0; v < n_views; v++) dabs = max(dabs, max(abs(dmin[v]), abs(dmin[v] + size_d - 1)));
    const MmaGeom mg = mma_geom(plan.n_strips, -dabs - M_ND, dabs + M_ND);
    const int pitchS = (g.w + 3) / 4 * 4;

    // power-of-two scale that keeps a and b inside fp16: |a| <= pmax / (2 sqrt(eps)), |b| <= pmax + 255 |a|
    const double pmax = (nI * (double)p->th_color + nG * 2.0 * (double)p->th_grad) / S;
    const double amax = 0.5 * pmax / sqrt(p->eps > 1e-12 ? p->eps : 1e-12);
    const double bmax = pmax + 255.0 * amax;
    int sh = (int)floor(log2(30000.0 / bmax));
    if (sh > 24) sh = 24;
    if (sh < -24) sh = -24;
    const float scale = ldexpf(1.0f, sh);

    uint2* GA[2];
    float2* GB[2];
    uint4* MT[2];
    for (int i = 0; i < 2; i++) {
        GA[i] = sb_ws_alloc<uint2>(ctx, (size_t)plan.n_strips * rows_pad * M_KB);
        GB[i] = sb_ws_alloc<float2>(ctx, (size_t)plan.n_strips * rows_pad * M_TW);
        MT[i] = sb_ws_alloc<uint4>(ctx, (size_t)rows_pad * mg.n_chunk * 4);
    }
    const size_t planeS = (size_t)g.rows_out * pitchS;
    float2* BL = sb_ws_alloc<float2>(ctx, planeS * 2 * plan.n_chunks);
    if (!GA[0] || !GA[1] || !GB[0] || !GB[1] || !MT[0] || !MT[1] || !BL)
        return sb_fail(ctx, SB200_ERR_NOMEM, "fused (mma): workspace arena too small (internal)");

    if (ctx->timing && ctx->ev_valid) SB_CUDA(ctx, cudaEventRecord(ctx->ev[0], ctx->stream));
    for (int i = 0; i < 2; i++) {
        PrepM P;
        P.gray = gray[i];
        P.w = g.w;
        P.h_held = g.h;
        P.y_global0 = g.y_global0;
        P.frame_h = g.frame_h;
        P.n_strips = plan.n_strips;
        P.rows_pad = rows_pad;
        P.GA = GA[i];
        P.GB = GB[i];
        P.MT = MT[i];
        P.n_chunk = mg.n_chunk;
        P.padm = mg.padm;
        P.mean_u8 = mean[i];
        P.eps = p->eps;
        P.S = (float)S;
        P.scale = scale;
        SB_LAUNCH(ctx, k_prep_ga, dim3(rows_pad, plan.n_strips), M_KB, 0, P);
        SB_LAUNCH(ctx, k_prep_gb, dim3(sb_div_up(rows_pad, GB_TR), plan.n_strips), M_TW, 0, P);
        SB_LAUNCH(ctx, k_prep_mt, dim3(sb_div_up(mg.n_chunk * 4, 256), rows_pad), 256, 0, P);
    }
    if (ctx->timing && ctx->ev_valid) SB_CUDA(ctx, cudaEventRecord(ctx->ev[1], ctx->stream));

    MmaArgs A;
    for (int i = 0; i < 2; i++) {
        A.GA[i] = GA[i];
        A.GB[i] = GB[i];
        A.MT[i] = MT[i];
        A.dmin[i] = dmin[i];
    }
    A.rows_pad = rows_pad;
    A.n_chunk = mg.n_chunk;
    A.padm = mg.padm;
    A.w = g.w;
    A.y_out0 = g.y_out0;
    A.rows_out = g.rows_out;
    A.y_global0 = g.y_global0;
    A.frame_h = g.frame_h;
    A.size_d = size_d;
    A.n_strips = plan.n_strips;
    A.n_bands = plan.n_bands;
    A.band_rows = plan.band_rows;
    A.n_chunks = plan.n_chunks;
    A.chunk_d = plan.chunk_d;
    A.n_views = n_views;
    A.BL = BL;
    A.pitchS = pitchS;
    A.S = (float)S;
    A.scale = scale;
    A.inv_scale = 1.0f / scale;
    auto pack2 = [](float v) {
        __half2 h = __floats2half2_rn(v, v);
        return *reinterpret_cast<unsigned*>(&h);
    };
    A.wI2 = pack2((float)nI);
    A.wG2 = pack2((float)nG);
    A.tc2 = pack2(p->th_color);
    A.tg2 = pack2(2.0f * p->th_grad);

    const size_t smem = sizeof(MSmem);
    if (!ctx->mma_attr_set) {
        SB_CUDA(ctx, cudaFuncSetAttribute(k_fused_mma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        ctx->mma_attr_set = true;
    }
    const int nblocks = plan.n_strips * plan.n_bands * plan.n_chunks * n_views;
    SB_LAUNCH(ctx, k_fused_mma, nblocks, M_THREADS, smem, A);
    if (ctx->timing && ctx->ev_valid) SB_CUDA(ctx, cudaEventRecord(ctx->ev[2], ctx->stream));
    for (int v = 0; v < n_views; v++) {
        if (!best[v] && !disp[v]) continue;
        dim3 grid(sb_div_up(g.w, 256), g.rows_out);
        SB_LAUNCH(ctx, k_merge_chunks_bl, grid, 256, 0, BL, plan.n_chunks, v, g.rows_out, g.w, pitchS, best[v], disp[v]);
    }
    if (ctx->timing && ctx->ev_valid) SB_CUDA(ctx, cudaEventRecord(ctx->ev[3], ctx->stream));
    return SB200_OK;
}
