// fused_mma_rgb.cu -- the hot path with the RGB guide of BASELINE configs[2] (colour guided filter), on the
// tensor cores: the design of fused_mma.cu (cost -> guided-filter aggregation -> running argmin for both views, the volume
// never leaves the SM, every HORIZONTAL 19-column window sum a product with a 0/1 band matrix) carried over to the four
// first-stage sums (p, R p, G p, B p) and the four second-stage sums (a_r, a_g, a_b, b) of the colour filter.
//
// The reference has no colour guide; what this replaces per disparity slice is its gray chain
//   costVolumOnGPU2 (costVolume.cu:163-190) -> pixelMultOnGPU -> 4x box (integral.cu:78-131, guidedFilter.cu:297-318) ->
//   compute_ak_and_bk (:345-354) -> compute_q (:363-369) -> dispSelectOnGPU (:403-411)
// generalised as in He et al.'s colour guided filter: cov_c = mean(I_c p) - mu_c mean(p), a = (Sigma + eps U)^-1 cov,
// b = mean(p) - a.mu, q = mean(a).I + mean(b); the cost stays the reference's gray cost (A.3).  Parity tests:
// tests/test_rgb_guide.py (against the CPU checker's port of the same definition).
//
// Decomposition (differences to fused_mma.cu): a block owns a strip of 128 columns (110 valid) and a group of FOUR
// consecutive disparities -- the 19-row ring of (a_r, a_g, a_b, b) takes 16 Tensor-Memory columns per row like gray's
// (a, b) of eight disparities, and Tensor Memory is full either way.
//   role A (5 warps)   lattice cost P (packed half, 2 disparities per instruction) for 160 cost columns x 4 d x 4 rows; the
//                      exact fp16 pieces P, (C&15) P, (C&240) P for C = R, G, B; what enters MMA 1 is piece(y) - piece(y-19)
//   MMA 1 (tensor)     dH = Band x dPiece: pass 1 = (lo_r, lo_g, lo_b, P), pass 2 = (hi_r, hi_g, hi_b, 0) accumulated onto it
//   role B (4 warps)   S += dH (2-D box sums, exact integers); cov, a = M cov (3x3, symmetric, from the preparation), b;
//                      their vertical 19-row sums (ring in Tensor Memory, re-summed every 64 rows); fp16 hi + lo.  One thread
//                      per a/b lane and ONE barrier wait per hand-off: d1_full collects MMA 1's commit, MMA 2's commit of the
//                      previous iteration ("your half of B2 is free") and the bulk copy of the half's statistics rows
//                      (8 warps of 2 disparities with three waits: 6.55 ms; this: 6.14 ms, bit-identical but for the re-sum)
//   MMA 2 (tensor)     H = Band x hi - Band x (hi - value)
//   role C (4 warps)   q = (H_ar (R-128) + H_ag (G-128) + H_ab (B-128) + H_b') / area, tournament over the 4 disparities with
//                      `best >= q`, read-modify-write of the chunk's (best,label) plane (prefetched by the TMA producer)
// Hand-offs are per 2 rows (two halves of every buffer), MMA 1 and MMA 2 have their own issuing warps, one warp runs TMA.
#include "mma_dev.cuh"

namespace {

constexpr int R_ND = 4;          // disparities per group
constexpr int MR = 4;            // image rows per pipeline iteration
constexpr int HR = 2;            // rows per hand-off
constexpr int NH = MR / HR;
constexpr int R_NWA = 5;         // role-A warps: one thread per cost column
#ifndef RGBM_BSPLIT
#define RGBM_BSPLIT 1            // role B: 1 = one thread per a/b lane (4 disparities), 2 = two threads (2 disparities each)
#endif
constexpr int NWB = 4 * RGBM_BSPLIT, NWC = 4;
constexpr int BD = R_ND / RGBM_BSPLIT;   // disparities per role-B thread
constexpr int NBC = 4 * BD;              // Tensor-Memory columns of a row that a role-B thread owns
constexpr int R_WARPS = (NWB + NWC + R_NWA + 3 + 3) / 4 * 4;
constexpr int R_THREADS = 32 * R_WARPS;
#ifndef RGBM_NA
#define RGBM_NA 4
#endif
#ifndef RGBM_NB
#define RGBM_NB 2
#endif
constexpr int R_NA = RGBM_NA;          // role A's operand ring (guide + match rows): slots of HR rows
constexpr int R_NB = RGBM_NB;          // role B's operand ring (statistics rows): slots of HR rows
#ifndef RGBM_PF
#define RGBM_PF 0
#endif
#ifndef RGBM_SLEEP
#define RGBM_SLEEP 0
#endif
constexpr int R_PF = RGBM_PF;    // L2 prefetch distance of the operand rows, in half-steps
constexpr int R_NGC = 2;         // role C's ring (colours of the output rows, previous (best,label))
#ifndef RGBM_RESUM
#define RGBM_RESUM 16
#endif
#ifndef RGBM_MERGE_BAR
#define RGBM_MERGE_BAR 1         // 1 = role B waits once per half: d1_full collects MMA 1's commit and MMA 2's of the previous iteration
#endif
#ifndef RGBM_OSPLIT
#define RGBM_OSPLIT 1            // 1 = role B loads the ring rows that leave the window AFTER the D1 rows and under the a/b arithmetic
#endif
#ifndef RGBM_A_MERGE0
#define RGBM_A_MERGE0 1          // 1 = role A's b1_empty wait of half 0 rides on its operand slot's full barrier (slots 0 and 2), as in
#endif                           // the gray kernel: -0.4 %, bit-identical (both halves merged: +4 %, the wait then precedes the loads)
#ifndef RGBM_S_MERGE
#define RGBM_S_MERGE 1           // 1 = the statistics rows' bulk copy completes on d1_full as well (slot == half): role B has ONE wait per half
#endif
constexpr int R_RESUM = RGBM_RESUM;
static_assert(!RGBM_A_MERGE0 || (RGBM_NA % NH == 0 && RGBM_NA >= NH), "A_MERGE0: half 0 must always land on the even operand slots");
static_assert(!RGBM_S_MERGE || (RGBM_MERGE_BAR && RGBM_NB == 2), "S_MERGE needs the merged d1_full barrier and one statistics slot per half");
constexpr float I_CENTER = 128.0f;
static_assert((4 * RAD) % MR == 0, "MR must divide the warm-up length");

constexpr uint32_t GA_ROW = M_KB * 16;            // (I, G | R&15, R&240 | G&15, G&240 | B&15, B&240) halves per cost column
constexpr uint32_t GB_ROW = M_TW * 36;            // [128 x (mu_r, mu_g, mu_b, Mrr)] [128 x (Mrg, Mrb, Mgg, Mgb)] [128 x Mbb]
constexpr uint32_t GB_S2 = M_TW * 16, GB_S3 = M_TW * 32;
constexpr uint32_t GC_ROW = M_TW * 6;             // [3][128] halves: C - 128 of the output lanes
constexpr uint32_t PB_ROW = M_TW * 8;
constexpr uint32_t MT_ROW = M_MTC * 64;
constexpr uint32_t A_MT = HR * GA_ROW, A_SLOT = HR * (GA_ROW + MT_ROW);  // role A's slot: HR guide rows, HR match rows
constexpr uint32_t B_SLOT = HR * GB_ROW;
constexpr uint32_t GC_SLOT = MR * (GC_ROW + PB_ROW), GC_PB = MR * GC_ROW;
constexpr uint32_t B1_GROUP = M_KB * 16;          // one N-group (8 columns) of B1: 160 rows x 16 B
constexpr uint32_t B1_HALF = 8 * B1_GROUP;        // pass 1: (row, d pair) x 4 = 4 groups; pass 2: 4 groups
constexpr uint32_t B1_BYTES = NH * B1_HALF;
constexpr uint32_t B2_GROUP = M_K2 * 16;
constexpr uint32_t B2_HALF = 8 * B2_GROUP;        // hi: (row, d pair) x 4 groups; lo likewise
constexpr uint32_t B2_BYTES = NH * B2_HALF;
constexpr uint32_t PR_P = M_KB * 8;               // a ring slot: P of 4 disparities per cost column, then the row's six
constexpr uint32_t PR_SLOT = PR_P + M_KB * 12;    // colour pieces per cost column (needed again when the row leaves)

struct RSmem {
    unsigned char b1[B1_BYTES];
    unsigned char b2[B2_BYTES];
    unsigned char pring[WIN][PR_SLOT];
    unsigned char ar[R_NA][A_SLOT];
    unsigned char br[R_NB][B_SLOT];
    unsigned char gc[R_NGC][GC_SLOT];
    float ry_lut[2][WIN + 1];
    uint64_t a_full[R_NA], a_empty[R_NA], s_full[R_NB], s_empty[R_NB], gc_full[R_NGC], gc_empty[R_NGC];
    uint64_t b1_full[NH], b1_empty[NH], d1_full[NH], d1_empty[NH], b2_full[NH], b2_empty[NH], d2_full[NH], d2_empty[NH];
    uint32_t tmem_base;
};
static_assert(sizeof(RSmem) <= 227 * 1024, "shared memory budget");

// Tensor-Memory column map of a lane (512 columns).  A row of D1 / D2 / the ring is 16 columns [d pair][value][d in pair];
// values of D1: (sum R p, sum G p, sum B p, sum p); of D2 and the ring: (a_r, a_g, a_b, b').
constexpr uint32_t TC_BAND = 0;     // 80
constexpr uint32_t TC_D1 = 80;      // 2 halves x 2 rows x 16
constexpr uint32_t TC_D2 = 144;     // likewise
constexpr uint32_t TC_RING = 208;   // 19 x 16

struct RgbMmaArgs {
    const uint4* GA[2];
    const float* GB[2];
    const __half* GC[2];
    const uint4* MT[2];
    int rows_pad, n_chunk, padm;
    int w;
    int y_out0, rows_out, y_global0, frame_h;
    int dmin[2], size_d;
    int n_strips, n_bands, band_rows, n_chunks, chunk_d, n_views;
    float2* BL;
    float* QV[2];          // optional, per VIEW: filtered cost volume q, [size_d][rows_out][w]; NULL = not stored
    int pitchS;
    float S, scale, inv_scale;
    unsigned wI2, wG2, tc2, tg2;
};

constexpr int R_REGS_LAUNCH = (65536 / R_THREADS) / 8 * 8;
constexpr int R_REGS_B = RGBM_BSPLIT == 2 ? 120 : 184;
constexpr int R_REGS_C = RGBM_BSPLIT == 2 ? 96 : 112;
constexpr int R_REGS_A = RGBM_BSPLIT == 2 ? 72 : 104;
static_assert(32 * NWB * R_REGS_B + 32 * NWC * R_REGS_C + 32 * (R_WARPS - NWB - NWC) * R_REGS_A <= R_THREADS * R_REGS_LAUNCH,
              "register pool of the block");

template <int N>
__device__ __forceinline__ void reg_set() {
    if (N < R_REGS_LAUNCH) reg_dec<N>();
    else if (N > R_REGS_LAUNCH) reg_inc<N>();
}

// QVOL: role C also stores q (the filtered cost volume; see k_fused_mma in fused_mma.cu) -- its own instantiation
template <bool QVOL>
__global__ void __launch_bounds__(R_THREADS, 1) k_fused_mma_rgb(const RgbMmaArgs A) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    RSmem& sm = *reinterpret_cast<RSmem*>(smem_raw);
    const int lane = threadIdx.x & 31;
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);

    int bid = blockIdx.x;
    const int view = bid % A.n_views;
    bid /= A.n_views;
    const int chunk = bid % A.n_chunks;
    bid /= A.n_chunks;
    const int band = bid % A.n_bands;
    const int strip = bid / A.n_bands;

    const int xc0 = strip * M_VW - 2 * RAD;
    const int xa0 = strip * M_VW - RAD;
    const int xo0 = strip * M_VW;
    const int dlo = A.dmin[view] + chunk * A.chunk_d;
    const int dcnt = min(A.chunk_d, A.size_d - chunk * A.chunk_d);
    const int ngroups = (dcnt + R_ND - 1) / R_ND;
    const int yb0 = A.y_out0 + band * A.band_rows;
    const int yb1 = min(yb0 + A.band_rows, A.y_out0 + A.rows_out);
    const int y_first = yb0 - 2 * RAD;
    const int niter = ((yb1 - yb0) + 4 * RAD + MR - 1) / MR;
    constexpr int WARM_IT = 4 * RAD / MR;
    const int Ktotal = ngroups * niter;

    if (warp == 0) tm_alloc(&sm.tmem_base);
    auto bar = [&](const uint64_t* b) { return smem_addr(b); };
    if (threadIdx.x == 32) {
        for (int i = 0; i < R_NA; i++) {
            mbar_init(bar(&sm.a_full[i]), 1 + ((RGBM_A_MERGE0 && (i & 1) == 0) ? 1 : 0));
            mbar_init(bar(&sm.a_empty[i]), R_NWA);
        }
        for (int i = 0; i < R_NB; i++) {
            mbar_init(bar(&sm.s_full[i]), 1);
            mbar_init(bar(&sm.s_empty[i]), NWB);
        }
        for (int i = 0; i < R_NGC; i++) {
            mbar_init(bar(&sm.gc_full[i]), 1);
            mbar_init(bar(&sm.gc_empty[i]), NWC);
        }
        for (int i = 0; i < NH; i++) {
            mbar_init(bar(&sm.b1_full[i]), R_NWA);
            mbar_init(bar(&sm.b1_empty[i]), 1);
            mbar_init(bar(&sm.d1_full[i]), 1 + RGBM_MERGE_BAR + RGBM_S_MERGE);
            mbar_init(bar(&sm.d1_empty[i]), NWB);
            mbar_init(bar(&sm.b2_full[i]), NWB);
            mbar_init(bar(&sm.b2_empty[i]), 1);
            mbar_init(bar(&sm.d2_full[i]), 1);
            mbar_init(bar(&sm.d2_empty[i]), NWC);
        }
        if (RGBM_A_MERGE0) mbar_arrive(bar(&sm.a_full[0]));  // iteration 0 has no previous MMA 1
        if (RGBM_MERGE_BAR)  // phase 0 has no previous MMA 2
            for (int i = 0; i < NH; i++) mbar_arrive(bar(&sm.d1_full[i]));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x >= 64 && threadIdx.x < 64 + 2 * (WIN + 1)) {
        const int t = threadIdx.x - 64, n = t % (WIN + 1);
        float v = 0.0f;
        if (n > 0) v = (t < WIN + 1) ? A.scale * __frcp_rn(A.S * (float)n) : A.inv_scale * __frcp_rn((float)n);
        sm.ry_lut[t / (WIN + 1)][n] = v;
    }
    {   // cost columns 146..159 of B1 are never written and meet zero band entries: they must be finite
        uint4* z = reinterpret_cast<uint4*>(smem_raw);
        const int n16 = (int)(offsetof(RSmem, ry_lut) / 16);
        for (int i = threadIdx.x; i < n16; i += R_THREADS) z[i] = make_uint4(0u, 0u, 0u, 0u);
    }
    fence_async_smem();
    tm_fence_before();
    __syncthreads();
    tm_fence_after();
    const uint32_t tmem = sm.tmem_base;

    if (warp < NWB) {
        // ================= role B =================
        reg_set<R_REGS_B>();
        const int q4 = warp & 3, h = warp >> 2;  // lane quarter, disparity part
        const int l = q4 * 32 + lane;
        const uint32_t tl = tmem + ((uint32_t)(q4 * 32) << 16);
        if (h == 0) {  // this lane's row of the band matrix: Band[l][k] = 1 for l <= k <= l + 18
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
            named_bar_arrive(1, 128 + 64);
        }
        const int x = xa0 + l;
        const bool xin = x >= 0 && x < A.w;
        const float rx = xin ? __frcp_rn((float)(min(A.w - 1, x + RAD) - max(0, x - RAD) + 1)) : 0.0f;
        const float r1_whole = rx * sm.ry_lut[0][WIN];
        const uint32_t td1 = tl + TC_D1 + NBC * h, tring = tl + TC_RING + NBC * h;
        const uint32_t b2a = smem_addr(&sm.b2[0]) + (uint32_t)((l >> 3) * 128 + (l & 7) * 16) + (uint32_t)h * (BD / 2) * B2_GROUP;
        const uint32_t ops = smem_addr(&sm.br[0][0]) + (uint32_t)l * 16;
        const uint32_t mb_d1f = bar(&sm.d1_full[0]), mb_d1e = bar(&sm.d1_empty[0]), mb_b2f = bar(&sm.b2_full[0]),
                       mb_b2e = bar(&sm.b2_empty[0]), mb_sf = bar(&sm.s_full[0]), mb_se = bar(&sm.s_empty[0]);
        int K = 0;
        TL_DECL
        for (int g = 0; g < ngroups; g++) {
            float Sp[BD], Sr[BD], Sg[BD], Sb[BD], V[4][BD];
#pragma unroll
            for (int d = 0; d < BD; d++) {
                Sp[d] = Sr[d] = Sg[d] = Sb[d] = 0.0f;
                V[0][d] = V[1][d] = V[2][d] = V[3][d] = 0.0f;
            }
            {
                uint32_t z[NBC];
#pragma unroll
                for (int i = 0; i < NBC; i++) z[i] = 0u;
                for (int s = 0; s < WIN; s++) tm_stN(tring + 16 * s, z);
                tm_wait_st();
            }
            int slot = 0;
#pragma unroll 1
            for (int it = 0; it < niter; it++, K++) {
                const int yi0 = y_first + it * MR;
                TL_BEGIN();
                float r1[MR];
                {
                    const int yg = yi0 - RAD + A.y_global0;
                    if (yg >= RAD && yg + MR - 1 + RAD < A.frame_h) {
#pragma unroll
                        for (int r = 0; r < MR; r++) r1[r] = r1_whole;
                    } else {
#pragma unroll
                        for (int r = 0; r < MR; r++) r1[r] = rx * inv_rows(sm.ry_lut[0], yi0 + r - RAD, A.y_global0, A.frame_h);
                    }
                }
#pragma unroll
                for (int half = 0; half < NH; half++) {
                    uint32_t dd1[2][NBC], o[2][NBC];
                    int slots[2];
                    // the statistics of the half's rows: into registers, and the slot goes back to the producer
                    const int hs = NH * K + half, ks = hs % R_NB;
                    if (RGBM_S_MERGE) { TL_WAIT(1, mbar_wait(mb_d1f + 8 * half, (unsigned)K & 1u)); }
                    else TL_WAIT(0, mbar_wait(mb_sf + 8 * ks, (unsigned)(hs / R_NB) & 1u));
                    uint4 st1[HR], st2[HR];
                    uint32_t st3[HR];
#pragma unroll
                    for (int j = 0; j < HR; j++) {
                        const uint32_t sa = ops + ks * B_SLOT + j * GB_ROW;
                        st1[j] = lds128(sa);
                        st2[j] = lds128(sa + GB_S2);
                        st3[j] = lds32(sa + GB_S3 - (uint32_t)l * 12);  // the third plane is 4 B per lane
                    }
                    if (!RGBM_S_MERGE) TL_WAIT(1, mbar_wait(mb_d1f + 8 * half, (unsigned)K & 1u));
                    tm_fence_after();
#pragma unroll
                    for (int j = 0; j < 2; j++) {
                        slots[j] = slot;
                        slot = (slot + 1 == WIN) ? 0 : slot + 1;
                        tm_ldN(td1 + 16 * HR * half + 16 * j, dd1[j]);
                        if (!RGBM_OSPLIT) tm_ldN(tring + 16 * slots[j], o[j]);  // the (a, b') row that leaves the vertical window
                    }
                    tm_wait_ld();
                    tm_fence_before();
                    __syncwarp();
                    if (lane == 0) {
                        mbar_arrive(mb_d1e + 8 * half);
                        mbar_arrive(mb_se + 8 * ks);  // the statistics are in registers
                    }
                    uint32_t abr[2][NBC];
                    if (RGBM_OSPLIT) {
#pragma unroll
                        for (int j = 0; j < 2; j++) tm_ldN(tring + 16 * slots[j], o[j]);
                    }
#pragma unroll
                    for (int pass = 0; pass < (RGBM_OSPLIT ? 2 : 1); pass++) {
                    if (RGBM_OSPLIT && pass == 1) tm_wait_ld();
#pragma unroll
                    for (int j = 0; j < 2; j++) {
                        const int r = 2 * half + j;
                        const uint4 s1 = st1[j], s2 = st2[j];
                        const uint32_t s3 = st3[j];
                        const float mr = __uint_as_float(s1.x), mg = __uint_as_float(s1.y), mb = __uint_as_float(s1.z);
                        const float Mrr = __uint_as_float(s1.w), Mrg = __uint_as_float(s2.x), Mrb = __uint_as_float(s2.y);
                        const float Mgg = __uint_as_float(s2.z), Mgb = __uint_as_float(s2.w), Mbb = __uint_as_float(s3);
                        const float cr = I_CENTER - mr, cg = I_CENTER - mg, cb = I_CENTER - mb;
                        uint32_t (&ab)[NBC] = abr[j];
                        if (!RGBM_OSPLIT || pass == 0) {
#pragma unroll
                        for (int d = 0; d < BD; d++) {
                            const int e = (d >> 1) * 8 + (d & 1);
                            Sr[d] += __uint_as_float(dd1[j][e]);
                            Sg[d] += __uint_as_float(dd1[j][e + 2]);
                            Sb[d] += __uint_as_float(dd1[j][e + 4]);
                            Sp[d] += __uint_as_float(dd1[j][e + 6]);
                            const float cx = fmaf(-mr, Sp[d], Sr[d]);
                            const float cy = fmaf(-mg, Sp[d], Sg[d]);
                            const float cz = fmaf(-mb, Sp[d], Sb[d]);
                            const float ar = fmaf(Mrr, cx, fmaf(Mrg, cy, Mrb * cz));
                            const float ag = fmaf(Mrg, cx, fmaf(Mgg, cy, Mgb * cz));
                            const float ab_ = fmaf(Mrb, cx, fmaf(Mgb, cy, Mbb * cz));
                            // b' = b + 128 (a_r + a_g + a_b): role C evaluates q = mean(a).(C - 128) + mean(b'), which halves the
                            // magnitudes that cancel in mean(a).C + mean(b)
                            const float bb = fmaf(ar, cr, fmaf(ag, cg, fmaf(ab_, cb, Sp[d] * r1[r])));
                            ab[e] = __float_as_uint(ar);
                            ab[e + 2] = __float_as_uint(ag);
                            ab[e + 4] = __float_as_uint(ab_);
                            ab[e + 6] = __float_as_uint(bb);
                            if (!RGBM_OSPLIT) {
                                V[0][d] += ar - __uint_as_float(o[j][e]);
                                V[1][d] += ag - __uint_as_float(o[j][e + 2]);
                                V[2][d] += ab_ - __uint_as_float(o[j][e + 4]);
                                V[3][d] += bb - __uint_as_float(o[j][e + 6]);
                            }
                        }
                        }
                        if (RGBM_OSPLIT && pass == 0) continue;
                        if (RGBM_OSPLIT) {
#pragma unroll
                            for (int d = 0; d < BD; d++) {
                                const int e = (d >> 1) * 8 + (d & 1);
#pragma unroll
                                for (int c = 0; c < 4; c++) V[c][d] += __uint_as_float(ab[e + 2 * c]) - __uint_as_float(o[j][e + 2 * c]);
                            }
                        }
                        tm_stN(tring + 16 * slots[j], ab);
                        if (r == MR - 1 && (it & (R_RESUM - 1)) == R_RESUM - 1) {
                            // re-sum the ring (it now ends with this row): bounds the drift of the running sums
                            tm_wait_st();
#pragma unroll
                            for (int d = 0; d < BD; d++) V[0][d] = V[1][d] = V[2][d] = V[3][d] = 0.0f;
                            for (int s = 0; s < WIN; s++) {
                                uint32_t v[NBC];
                                tm_ldN(tring + 16 * s, v);
                                tm_wait_ld();
#pragma unroll
                                for (int d = 0; d < BD; d++) {
                                    const int e = (d >> 1) * 8 + (d & 1);
#pragma unroll
                                    for (int c = 0; c < 4; c++) V[c][d] += __uint_as_float(v[e + 2 * c]);
                                }
                            }
                        }
                        if (!RGBM_MERGE_BAR && j == 0 && K >= 1) TL_WAIT(2, mbar_wait(mb_b2e + 8 * half, (unsigned)(K - 1) & 1u));
                        // fp16 hi + lo of the vertical sums: hi - value = -(lo part); MMA 2 takes the lo pass with B negated.
                        // A B2 group (8 columns of D2) = (a_r, a_g, a_b, b') of one disparity pair.
#pragma unroll
                        for (int q = 0; q < BD / 2; q++) {
                            uint32_t hi[4], lo[4];
#pragma unroll
                            for (int c = 0; c < 4; c++) {
                                const float v0 = V[c][2 * q], v1 = V[c][2 * q + 1];
                                const unsigned hv = h22u(__floats2half2_rn(v0, v1));
                                hi[c] = hv;
                                lo[c] = h22u(__floats2half2_rn(fhadd_lo(hv, -v0), fhadd_hi(hv, -v1)));
                            }
                            sts128(b2a + half * B2_HALF + (uint32_t)(j * 2 + q) * B2_GROUP, hi[0], hi[1], hi[2], hi[3]);
                            sts128(b2a + half * B2_HALF + (uint32_t)(4 + j * 2 + q) * B2_GROUP, lo[0], lo[1], lo[2], lo[3]);
                        }
                    }
                    }  // pass
                    tm_wait_st();
                    fence_async_smem();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(mb_b2f + 8 * half);
                }
                if (warp == 0) TL_END(0, K);
            }
        }
    } else if (warp < NWB + NWC) {
        // ================= role C: q, winner-take-all =================
        reg_set<R_REGS_C>();
        const int q4 = warp & 3;
        const int l = q4 * 32 + lane;
        const uint32_t tl = tmem + ((uint32_t)(q4 * 32) << 16);
        const int x = xo0 + l;
        const bool valid = l < M_VW && x < A.w;
        const float rx = valid ? __frcp_rn((float)(min(A.w - 1, x + RAD) - max(0, x - RAD) + 1)) : 0.0f;
        const float rxy_whole = rx * sm.ry_lut[1][WIN];
        const int pitchS = A.pitchS;
        const size_t planeS = (size_t)A.rows_out * pitchS;
        float2* const bl0 = A.BL + (size_t)(chunk * 2 + view) * planeS + (size_t)(yb0 - A.y_out0) * pitchS + x;
        const int band_rows = yb1 - yb0;
        const uint32_t td2 = tl + TC_D2;
        const uint32_t gcs = smem_addr(&sm.gc[0][0]) + (uint32_t)l * 2;
        const uint32_t pbs = smem_addr(&sm.gc[0][0]) + GC_PB + (uint32_t)l * 8;
        const size_t qplane = (size_t)A.rows_out * A.w;
        float* const qv0 = (QVOL && A.QV[view]) ? A.QV[view] + (size_t)(chunk * A.chunk_d) * qplane + (size_t)(yb0 - A.y_out0) * A.w + x : nullptr;
        const uint32_t mb_d2f = bar(&sm.d2_full[0]), mb_d2e = bar(&sm.d2_empty[0]), mb_gcf = bar(&sm.gc_full[0]), mb_gce = bar(&sm.gc_empty[0]);
        int K = 0;
        TL_DECL
        for (int g = 0; g < ngroups; g++) {
            const float dbase = (float)(dlo + g * R_ND);
            const int dact = min(R_ND, dcnt - g * R_ND);
            const bool ld_ok = g > 0;
            const float2 binit = make_float2(BEST_INIT_BITS_F, 0.0f);
            float2* blp = bl0;
            float* qvp = (QVOL && qv0) ? qv0 + (size_t)(g * R_ND) * qplane : nullptr;
#pragma unroll 1
            for (int it = 0; it < niter; it++, K++) {
                const int e = it - WARM_IT;
                const int kc = K & (R_NGC - 1);
                const int row0 = e * MR;
                TL_BEGIN();
                TL_WAIT(0, mbar_wait(mb_gcf + 8 * kc, (unsigned)(K / R_NGC) & 1u));
                if (e < 0) {
#pragma unroll
                    for (int hf = 0; hf < NH; hf++) mbar_wait(mb_d2f + 8 * hf, (unsigned)K & 1u);
                    __syncwarp();
                    if (lane == 0) {
#pragma unroll
                        for (int hf = 0; hf < NH; hf++) mbar_arrive(mb_d2e + 8 * hf);
                        mbar_arrive(mb_gce + 8 * kc);
                    }
                    continue;
                }
                float2 pb[MR];
                float rxy[MR], Cc[MR][3];
                bool st_ok[MR];
                {
                    const int yg = yb0 + row0 + A.y_global0;
                    const bool whole = yg >= RAD && yg + MR - 1 + RAD < A.frame_h;
#pragma unroll
                    for (int j = 0; j < MR; j++) {
                        pb[j] = binit;
                        if (ld_ok) {
                            const uint2 v = lds64(pbs + kc * GC_SLOT + j * PB_ROW);
                            pb[j] = make_float2(__uint_as_float(v.x), __uint_as_float(v.y));
                        }
                        rxy[j] = whole ? rxy_whole : rx * inv_rows(sm.ry_lut[1], yb0 + row0 + j, A.y_global0, A.frame_h);
#pragma unroll
                        for (int c = 0; c < 3; c++)
                            Cc[j][c] = __half2float(__ushort_as_half((unsigned short)lds16(gcs + kc * GC_SLOT + j * GC_ROW + c * (M_TW * 2))));
                        st_ok[j] = valid && row0 + j < band_rows;
                    }
                }
#pragma unroll
                for (int half = 0; half < NH; half++) {
                    uint32_t hh[2][16];
                    TL_WAIT(1, mbar_wait(mb_d2f + 8 * half, (unsigned)K & 1u));
                    tm_fence_after();
#pragma unroll
                    for (int j = 0; j < 2; j++) tm_ld16u(td2 + 16 * (HR * half + j), hh[j]);
                    tm_wait_ld();
                    tm_fence_before();
                    __syncwarp();
                    mbar_arrive_lane0(mb_d2e + 8 * half, lane);
#pragma unroll
                    for (int j2 = 0; j2 < 2; j2++) {
                        const int j = HR * half + j2;
                        float q[R_ND];
#pragma unroll
                        for (int d = 0; d < R_ND; d++) {
                            const int ee = (d >> 1) * 8 + (d & 1);
                            const float sr = __uint_as_float(hh[j2][ee]), sg = __uint_as_float(hh[j2][ee + 2]);
                            const float sb = __uint_as_float(hh[j2][ee + 4]), s0 = __uint_as_float(hh[j2][ee + 6]);
                            q[d] = fmaf(sr, Cc[j][0], fmaf(sg, Cc[j][1], fmaf(sb, Cc[j][2], s0))) * rxy[j];
                        }
                        if (dact < R_ND) {
#pragma unroll
                            for (int d = 0; d < R_ND; d++)
                                if (d >= dact) q[d] = __int_as_float(0x7f800000);
                        }
                        if (QVOL && qvp && st_ok[j]) {
                            float* qd = qvp + (size_t)j * A.w;
#pragma unroll
                            for (int d = 0; d < R_ND; d++) {
                                if (d < dact) *qd = q[d];
                                qd += qplane;
                            }
                        }
                        // ascending d, `best >= q`: minimum, the later index on a tie (guidedFilter.cu:406), as a tournament
                        const bool t01 = q[0] >= q[1], t23 = q[2] >= q[3];
                        const float m01 = t01 ? q[1] : q[0], a01 = t01 ? 1.0f : 0.0f;
                        const float m23 = t23 ? q[3] : q[2], a23 = t23 ? 3.0f : 2.0f;
                        const bool tt = m01 >= m23;
                        const float m = tt ? m23 : m01, am = tt ? a23 : a01;
                        float2 nb = pb[j];
                        if (nb.x >= m) {
                            nb.x = m;
                            nb.y = am + dbase;
                        }
                        if (st_ok[j]) blp[(size_t)j * pitchS] = nb;
                    }
                }
                blp += (size_t)MR * pitchS;
                if (QVOL && qvp) qvp += (size_t)MR * A.w;
                asm volatile("fence.proxy.async.global;" ::: "memory");  // the TMA producer reads these rows one group later
                __syncwarp();
                mbar_arrive_lane0(mb_gce + 8 * kc, lane);
                if (warp == NWB) TL_END(1, K);
            }
        }
    } else {
    reg_set<R_REGS_A>();
    if (warp < NWB + NWC + R_NWA) {
        // ================= role A: lattice cost, exact fp16 pieces, vertical differences =================
        const int k = (warp - NWB - NWC) * 32 + lane;
        const int x = xc0 + k;
        const bool used = k < M_KC && x >= 0 && x < A.w;
        const __half2 wI = used ? u2h2(A.wI2) : __float2half2_rn(0.0f);
        const __half2 wG = used ? u2h2(A.wG2) : __float2half2_rn(0.0f);
        const __half2 tc = u2h2(A.tc2), tg = u2h2(A.tg2);
        const uint32_t b1a = smem_addr(&sm.b1[0]) + (uint32_t)((k >> 3) * 128 + (k & 7) * 16);
        const uint32_t pr0 = smem_addr(&sm.pring[0][0]) + (uint32_t)k * 8;
        const uint32_t pi0 = smem_addr(&sm.pring[0][0]) + PR_P + (uint32_t)k * 12;
        const uint32_t ops = smem_addr(&sm.ar[0][0]);
        const uint32_t mb_af = bar(&sm.a_full[0]), mb_ae = bar(&sm.a_empty[0]), mb_b1f = bar(&sm.b1_full[0]), mb_b1e = bar(&sm.b1_empty[0]);
        int K = 0;
        TL_DECL
        for (int g = 0; g < ngroups; g++) {
            for (int s = 0; s < WIN; s++) {
                const uint32_t so = (uint32_t)s * PR_SLOT;
                asm volatile("st.shared.v2.u32 [%0], {%1, %1};" ::"r"(pr0 + so), "r"(0u) : "memory");
                sts32(pi0 + so, 0u);
                sts32(pi0 + so + 4, 0u);
                sts32(pi0 + so + 8, 0u);
            }
            const int d0 = dlo + g * R_ND;
            const int X0 = x + d0 + A.padm;
            const int i0 = (xc0 + d0 + A.padm) >> 2;
            const uint32_t mt_off = A_MT + (uint32_t)((X0 >> 2) - i0) * 64 + (uint32_t)(X0 & 3) * 16;
            int slot = 0;
#pragma unroll 1
            for (int it = 0; it < niter; it++, K++) {
                TL_BEGIN();
#pragma unroll
                for (int r = 0; r < MR; r++) {
                    const int half = r / HR, rr = r % HR;
                    const int hs = NH * K + half, ka = hs % R_NA;
                    if (rr == 0) TL_WAIT(0, mbar_wait(mb_af + 8 * ka, (unsigned)(hs / R_NA) & 1u));
                    const uint32_t opa = ops + ka * A_SLOT;
                    const uint32_t so = (uint32_t)slot * PR_SLOT;
                    TL_MARK(1);
                    const uint4 gn = lds128(opa + rr * GA_ROW + (uint32_t)k * 16);
                    const uint4 mt = lds128(opa + mt_off + rr * MT_ROW);
                    const uint2 po = lds64(pr0 + so);
                    const uint32_t go[3] = {lds32(pi0 + so), lds32(pi0 + so + 4), lds32(pi0 + so + 8)};
                    const __half2 gI = __low2half2(u2h2(gn.x)), gG = __high2half2(u2h2(gn.x));
                    const unsigned gcn[3] = {gn.y, gn.z, gn.w};
                    const unsigned mi[2] = {mt.x, mt.y}, mg[2] = {mt.z, mt.w};
                    const unsigned pold[2] = {po.x, po.y};
                    unsigned pn[2], g1[2][4], g2[2][4];  // per disparity pair: pass-1 group (lo_r, lo_g, lo_b, dP), pass-2 group
#pragma unroll
                    for (int j = 0; j < 2; j++) {
                        const __half2 cI = __hmin2(__habs2(__hsub2(u2h2(mi[j]), gI)), tc);
                        const __half2 cG = __hmin2(__habs2(__hsub2(u2h2(mg[j]), gG)), tg);
                        const __half2 P = __hfma2(cG, wG, __hmul2(cI, wI));
                        const __half2 Po = u2h2(pold[j]);
                        pn[j] = h22u(P);
                        g1[j][3] = h22u(__hsub2(P, Po));
                        g2[j][3] = 0u;
#pragma unroll
                        for (int c = 0; c < 3; c++) {
                            const __half2 gl = __low2half2(u2h2(gcn[c])), gh = __high2half2(u2h2(gcn[c]));
                            const __half2 ol = __hneg2(__low2half2(u2h2(go[c]))), oh = __hneg2(__high2half2(u2h2(go[c])));
                            g1[j][c] = h22u(__hfma2(Po, ol, __hmul2(P, gl)));
                            g2[j][c] = h22u(__hfma2(Po, oh, __hmul2(P, gh)));
                        }
                    }
                    asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(pr0 + so), "r"(pn[0]), "r"(pn[1]) : "memory");
                    sts32(pi0 + so, gn.y);
                    sts32(pi0 + so + 4, gn.z);
                    sts32(pi0 + so + 8, gn.w);
                    slot = (slot + 1 == WIN) ? 0 : slot + 1;
                    if (rr == 0 && K >= 1 && !(RGBM_A_MERGE0 && half == 0)) TL_WAIT(0, mbar_wait(mb_b1e + 8 * half, (unsigned)(K - 1) & 1u));
                    TL_MARK(2);
                    const uint32_t bh = b1a + half * B1_HALF + (uint32_t)(rr * 2) * B1_GROUP;
#pragma unroll
                    for (int j = 0; j < 2; j++) {
                        sts128(bh + (uint32_t)j * B1_GROUP, g1[j][0], g1[j][1], g1[j][2], g1[j][3]);
                        sts128(bh + (uint32_t)(4 + j) * B1_GROUP, g2[j][0], g2[j][1], g2[j][2], g2[j][3]);
                    }
                    if (rr == HR - 1) {
                        fence_async_smem();
                        __syncwarp();
                        if (lane == 0) {
                            mbar_arrive(mb_b1f + 8 * half);
                            mbar_arrive(mb_ae + 8 * ka);
                        }
                    }
                    TL_MARK(3);
                }
                if (warp == NWB + NWC) TL_END(2, K);
            }
        }
    } else if (warp == NWB + NWC + R_NWA) {
        // ================= MMA 1 issue =================
        named_bar_sync(1, 128 + 64);
        tm_fence_after();
        constexpr uint32_t ID = instr_desc(32, 0);
        const uint64_t db1 = smem_desc(smem_addr(&sm.b1[0]), 128, B1_GROUP);
        const uint32_t ta = tmem + TC_BAND, td1 = tmem + TC_D1;
        const uint32_t mb_b1f = bar(&sm.b1_full[0]), mb_b1e = bar(&sm.b1_empty[0]), mb_d1f = bar(&sm.d1_full[0]), mb_d1e = bar(&sm.d1_empty[0]);
#pragma unroll 1
        TL_DECL
        for (int K = 0; K < Ktotal; K++) {
            TL_BEGIN();
#pragma unroll
            for (int half = 0; half < NH; half++) {
                TL_WAIT(0, mbar_wait(mb_b1f + 8 * half, (unsigned)K & 1u));
                if (K >= 1) TL_WAIT(1, mbar_wait(mb_d1e + 8 * half, (unsigned)(K - 1) & 1u));
                tm_fence_after();
                if (elect_one()) {
                    const uint64_t p1 = db1 + (uint64_t)((half * B1_HALF) >> 4), p2 = p1 + (uint64_t)((4 * B1_GROUP) >> 4);
                    const uint32_t td = td1 + 16 * HR * half;
#pragma unroll
                    for (int j = 0; j < M_KB / 16; j++) umma_ts(td, ta + 8 * j, p1 + (uint64_t)(j * 16), ID, j > 0);
#pragma unroll
                    for (int j = 0; j < M_KB / 16; j++) umma_ts(td, ta + 8 * j, p2 + (uint64_t)(j * 16), ID, 1);
                    umma_commit(mb_d1f + 8 * half);
                    if (RGBM_A_MERGE0 && half == 0) umma_commit(bar(&sm.a_full[0]) + 8 * ((NH * (K + 1)) % R_NA));
                    else umma_commit(mb_b1e + 8 * half);
                }
                __syncwarp();
            }
            TL_END(3, K);
        }
    } else if (warp == NWB + NWC + R_NWA + 2) {
        // ================= MMA 2 issue =================
        named_bar_sync(1, 128 + 64);
        tm_fence_after();
        constexpr uint32_t ID = instr_desc(32, 0), IDN = instr_desc(32, 1);
        const uint64_t db2 = smem_desc(smem_addr(&sm.b2[0]), 128, B2_GROUP);
        const uint32_t ta = tmem + TC_BAND, td2 = tmem + TC_D2;
        const uint32_t mb_b2f = bar(&sm.b2_full[0]), mb_b2e = bar(&sm.b2_empty[0]), mb_d2f = bar(&sm.d2_full[0]), mb_d2e = bar(&sm.d2_empty[0]);
#pragma unroll 1
        TL_DECL
        for (int K = 0; K < Ktotal; K++) {
            TL_BEGIN();
#pragma unroll
            for (int half = 0; half < NH; half++) {
                TL_WAIT(0, mbar_wait(mb_b2f + 8 * half, (unsigned)K & 1u));
                if (K >= 1) TL_WAIT(1, mbar_wait(mb_d2e + 8 * half, (unsigned)(K - 1) & 1u));
                tm_fence_after();
                if (elect_one()) {
                    const uint64_t dhi = db2 + (uint64_t)((half * B2_HALF) >> 4), dlo2 = dhi + (uint64_t)((4 * B2_GROUP) >> 4);
                    const uint32_t td = td2 + 16 * HR * half;
#pragma unroll
                    for (int j = 0; j < M_K2 / 16; j++) umma_ts(td, ta + 8 * j, dhi + (uint64_t)(j * 16), ID, j > 0);
#pragma unroll
                    for (int j = 0; j < M_K2 / 16; j++) umma_ts(td, ta + 8 * j, dlo2 + (uint64_t)(j * 16), IDN, 1);
                    umma_commit(mb_d2f + 8 * half);
                    umma_commit((RGBM_MERGE_BAR ? bar(&sm.d1_full[0]) : mb_b2e) + 8 * half);
                }
                __syncwarp();
            }
            TL_END(4, K);
        }
    } else if (warp == NWB + NWC + R_NWA + 1) {
        // ================= TMA producer =================
        // Three rings with different consumers (role A runs ahead of role B, role B ahead of role C): the warp polls the
        // rings' empty barriers (mbarrier.test_wait) and refills whichever has a free slot, so a ring whose consumer lags
        // does not hold back the others.
        const char* GAp = reinterpret_cast<const char*>(A.GA[view]) + (size_t)strip * A.rows_pad * GA_ROW;
        const char* GBp = reinterpret_cast<const char*>(A.GB[view]) + (size_t)strip * A.rows_pad * GB_ROW;
        const char* GCp = reinterpret_cast<const char*>(A.GC[view]) + (size_t)strip * A.rows_pad * GC_ROW;
        const char* MTp = reinterpret_cast<const char*>(A.MT[1 - view]);
        const size_t mt_pitch = (size_t)A.n_chunk * 64;
        const long long row0 = (long long)PADY + y_first;
        const uint32_t a_s = smem_addr(&sm.ar[0][0]), b_s = smem_addr(&sm.br[0][0]), gc_s = smem_addr(&sm.gc[0][0]);
        const uint32_t a_f = bar(&sm.a_full[0]), a_e = bar(&sm.a_empty[0]), s_f = bar(&sm.s_full[0]), s_e = bar(&sm.s_empty[0]);
        const uint32_t gc_f = bar(&sm.gc_full[0]), gc_e = bar(&sm.gc_empty[0]);
        const char* pb_src = reinterpret_cast<const char*>(A.BL + (size_t)(chunk * 2 + view) * ((size_t)A.rows_out * A.pitchS) +
                                                           (size_t)(yb0 - A.y_out0) * A.pitchS + xo0);
        const size_t pb_pitch = (size_t)A.pitchS * 8;
        const int nhalf = NH * niter;  // half-steps per group
        // per ring: next step to issue, its position inside the group (g, step in group)
        int sa = 0, ga = 0, ha = 0;
        int sb = 0, hb = 0;
        int sc = 0, gcg = 0, itc = 0;
        const int SA = NH * Ktotal, SC = Ktotal;
        // The operand planes of one pass over the rows (about 11 MB per block, 1.6 GB per launch wave) do not stay in L2 from
        // one disparity group to the next: every bulk load would pay the DRAM latency, which the rings' few slots cannot
        // cover.  Each refill therefore also asks L2 for the rows R_PF half-steps ahead (cp.async.bulk.prefetch.L2).
        auto pf_a = [&](int g, int hh) {  // role A's rows of half-step hh of group g
            if (hh >= nhalf) { hh -= nhalf; g++; }
            if (g >= ngroups) return;
            const int i0 = (xc0 + dlo + g * R_ND + A.padm) >> 2;
            const long long row = row0 + (long long)hh * HR;
            l2_prefetch(GAp + row * GA_ROW, HR * GA_ROW);
#pragma unroll
            for (int r = 0; r < HR; r++) l2_prefetch(MTp + (row + r) * mt_pitch + (size_t)i0 * 64, MT_ROW);
        };
        auto pf_b = [&](int hh) {
            if (hh >= nhalf) hh -= nhalf;
            l2_prefetch(GBp + (row0 - RAD + (long long)hh * HR) * GB_ROW, B_SLOT);
        };
        if (elect_one()) {
            for (int i = 0; i < R_PF && i < nhalf; i++) {
                pf_a(0, i);
                pf_b(i);
            }
        }
        __syncwarp();
        while (sa < SA || sb < SA || sc < SC) {
            bool any = false;
            if (sa < SA) {
                const int slot = sa % R_NA;
                if (sa < R_NA || mbar_test(a_e + 8 * slot, (unsigned)(sa / R_NA - 1) & 1u)) {
                    if (elect_one()) {
                        const int d0 = dlo + ga * R_ND;
                        const int i0 = (xc0 + d0 + A.padm) >> 2;
                        const long long row = row0 + (long long)ha * HR;
                        const uint32_t dst = a_s + slot * A_SLOT;
                        mbar_expect_tx(a_f + 8 * slot, A_SLOT);
                        bulk_g2s(dst, GAp + row * GA_ROW, HR * GA_ROW, a_f + 8 * slot);
#pragma unroll
                        for (int r = 0; r < HR; r++)
                            bulk_g2s(dst + A_MT + r * MT_ROW, MTp + (row + r) * mt_pitch + (size_t)i0 * 64, MT_ROW, a_f + 8 * slot);
                        if (R_PF > 0) pf_a(ga, ha + R_PF);
                    }
                    __syncwarp();
                    sa++;
                    if (++ha == nhalf) { ha = 0; ga++; }
                    any = true;
                }
            }
            if (sb < SA) {
                const int slot = sb % R_NB;
                if (sb < R_NB || mbar_test(s_e + 8 * slot, (unsigned)(sb / R_NB - 1) & 1u)) {
                    if (elect_one()) {
                        const long long row = row0 - RAD + (long long)hb * HR;
                        const uint32_t sbar = RGBM_S_MERGE ? bar(&sm.d1_full[0]) + 8 * slot : s_f + 8 * slot;
                        mbar_expect_tx(sbar, B_SLOT);
                        bulk_g2s(b_s + slot * B_SLOT, GBp + row * GB_ROW, B_SLOT, sbar);
                        if (R_PF > 0 && sb + R_PF < SA) pf_b(hb + R_PF);
                    }
                    __syncwarp();
                    sb++;
                    if (++hb == nhalf) hb = 0;
                    any = true;
                }
            }
            if (sc < SC) {
                const int slot = sc & (R_NGC - 1);
                if (sc < R_NGC || mbar_test(gc_e + 8 * slot, (unsigned)(sc / R_NGC - 1) & 1u)) {
                    if (elect_one()) {
                        const bool pb = gcg > 0 && itc >= WARM_IT;  // rows past the band's end read the plane's padding; not used
                        const long long row = row0 - 2 * RAD + (long long)itc * MR;
                        mbar_expect_tx(gc_f + 8 * slot, MR * GC_ROW + (pb ? MR * PB_ROW : 0u));
                        bulk_g2s(gc_s + slot * GC_SLOT, GCp + row * GC_ROW, MR * GC_ROW, gc_f + 8 * slot);
                        if (pb) {
#pragma unroll
                            for (int r = 0; r < MR; r++)
                                bulk_g2s(gc_s + slot * GC_SLOT + GC_PB + r * PB_ROW, pb_src + (size_t)((itc - WARM_IT) * MR + r) * pb_pitch,
                                         PB_ROW, gc_f + 8 * slot);
                        }
                    }
                    __syncwarp();
                    sc++;
                    if (++itc == niter) { itc = 0; gcg++; }
                    any = true;
                }
            }
#if RGBM_SLEEP > 0
            if (!any) __nanosleep(RGBM_SLEEP);
#endif
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
// Preparation (per image): GA / GC by gather, GB from exact integer window sums of the colours and their products with the
// 3x3 inverse in double (the arithmetic of k_prep_rgb3, fused_cvf_rgb3.cu), MT by k_prep_mt (mma_dev.cuh).
struct PrepR {
    const uint8_t* gray;  // held rows, pitch w
    const uint8_t* rgb;   // held rows, interleaved, ch bytes per pixel
    int ch, w, h_held, y_global0, frame_h;
    int n_strips, rows_pad;
    uint4* GA;
    float* GB;
    __half* GC;
    double eps;
    float S, scale;
};

__device__ __forceinline__ bool in_frame_r(const PrepR& P, int x, int y) {
    const int yg = y + P.y_global0;
    return x >= 0 && x < P.w && y >= 0 && y < P.h_held && yg >= 0 && yg < P.frame_h;
}

// GA, GC: thread per (strip, padded row, cost column)
__global__ void __launch_bounds__(160) k_prep_ga3(const PrepR P0, const PrepR P1) {
    const PrepR& P = blockIdx.z ? P1 : P0;
    const int k = threadIdx.x, strip = blockIdx.y;
    const int x = strip * M_VW - 2 * RAD + k;
    for (int yrow = blockIdx.x * PREP_ROWS; yrow < min(P.rows_pad, (int)(blockIdx.x + 1) * PREP_ROWS); yrow++) {
    const int y = yrow - PADY;
    const bool in = in_frame_r(P, x, y);
    float I = 1024.0f, G = 1024.0f;
    int c[3] = {0, 0, 0};
    if (in) {
        const uint8_t* row = P.gray + (size_t)y * P.w;
        const int ic = row[x];
        const int il = (x - 1 >= 0) ? row[x - 1] : ic;
        const int ir = (x + 1 < P.w) ? row[x + 1] : ic;
        I = (float)ic;
        G = (float)(il - ir);
        const uint8_t* q = P.rgb + ((size_t)y * P.w + x) * P.ch;
        c[0] = q[0];
        c[1] = q[1];
        c[2] = q[2];
    }
    uint4 v;
    v.x = h22u(__floats2half2_rn(I, G));
    v.y = h22u(__floats2half2_rn((float)(c[0] & 15), (float)(c[0] & 240)));
    v.z = h22u(__floats2half2_rn((float)(c[1] & 15), (float)(c[1] & 240)));
    v.w = h22u(__floats2half2_rn((float)(c[2] & 15), (float)(c[2] & 240)));
    const size_t rec = (size_t)strip * P.rows_pad + yrow;
    P.GA[rec * M_KB + k] = v;
    if (k >= 2 * RAD && k < 2 * RAD + M_TW) {  // output lane l is cost column l + 18
        const int l = k - 2 * RAD;
#pragma unroll
        for (int ch = 0; ch < 3; ch++) P.GC[rec * (3 * M_TW) + ch * M_TW + l] = __float2half(in ? (float)c[ch] - I_CENTER : 0.0f);
    }
    }
}

// GB: block = (strip, GB3_TR padded rows), thread per a/b lane.  The block's colour tile (GB3_TR + 18 rows x 146 columns)
// is loaded once; a thread then walks down its column: the window sums are vertical running sums of the rows' horizontal
// 19-tap sums of the nine quantities (the row that enters adds, the row that leaves is recomputed and subtracted).
constexpr int GB3_TR = 32;
constexpr int GB3_T = 160;  // threads: one per tile column (146 used: the 128 a/b lanes and 9 columns on either side)
// The nine window sums (R, G, B and their six products; exact integers) of every a/b lane: each thread slides the VERTICAL
// 19-row sums of its tile column down the block's rows in registers (+ entering pixel, - leaving pixel), a warp prefix sum
// turns a row of them into prefix sums along x, and a lane's 19-column window is the difference of two prefixes (plus the
// total of the warp to its left when the window straddles two warps).  The first version took the 19 x 9 taps of a row's
// horizontal sums twice per output row (0.11 ms per image; this: see DESIGN 6).
__global__ void __launch_bounds__(GB3_T) k_prep_gb3(const PrepR P0, const PrepR P1) {
    const PrepR& P = blockIdx.z ? P1 : P0;
    __shared__ int sP[2][9][GB3_T];       // warp-local inclusive prefix sums along x of the current row (double-buffered)
    __shared__ int sT[2][9][GB3_T / 32];  // the warps' totals
    const int t = threadIdx.x, lane = t & 31, wid = t >> 5, strip = blockIdx.y;
    const int yrow0 = blockIdx.x * GB3_TR;
    const int xa0 = strip * M_VW - RAD;
    const int xc = xa0 + t - RAD;  // frame column of this thread's tile column
    auto pixel = [&](int y, int& r, int& g, int& b) {  // held row y of the thread's column; 0 outside the frame
        r = g = b = 0;
        if (t < M_TW + 2 * RAD && in_frame_r(P, xc, y)) {
            const uint8_t* q = P.rgb + ((size_t)y * P.w + xc) * P.ch;
            r = q[0];
            g = q[1];
            b = q[2];
        }
    };
    int v[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    auto add = [&](int y, int sgn) {
        int r, g, b;
        pixel(y, r, g, b);
        v[0] += sgn * r; v[1] += sgn * g; v[2] += sgn * b;
        v[3] += sgn * r * r; v[4] += sgn * r * g; v[5] += sgn * r * b;
        v[6] += sgn * g * g; v[7] += sgn * g * b; v[8] += sgn * b * b;
    };
    const int y0 = yrow0 - PADY;  // held row of the block's first output row
    for (int dy = -RAD; dy < RAD; dy++) add(y0 + dy, 1);  // rows y0-9 .. y0+8
    const int x = xa0 + t;  // frame column of a/b lane t (threads 0..127)
    for (int ty = 0; ty < GB3_TR; ty++) {
        const int yrow = yrow0 + ty, y = yrow - PADY, yg = y + P.y_global0;
        const int buf = ty & 1;
        add(y + RAD, 1);
        // warp-local inclusive prefix sums of the nine vertical sums
        int pre[9];
#pragma unroll
        for (int q = 0; q < 9; q++) {
            int p = v[q];
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const int n = __shfl_up_sync(0xffffffffu, p, off);
                if (lane >= off) p += n;
            }
            pre[q] = p;
            sP[buf][q][t] = p;
            if (lane == 31) sT[buf][q][wid] = p;
        }
        __syncthreads();  // (one barrier per row: the buffers alternate)
        if (t < M_TW && yrow < P.rows_pad) {
            // window of lane t = tile columns t .. t+18: prefix(t+18) - prefix(t-1)
            const int b = t + 2 * RAD, wb = b >> 5;
            int s[9];
#pragma unroll
            for (int q = 0; q < 9; q++) {
                int acc = sP[buf][q][b];
                if (wb != wid) acc += sT[buf][q][wid];       // the window ends in the next warp: this warp's total
                s[q] = acc - (pre[q] - v[q]);                // minus the exclusive prefix at t (0 at a warp's first lane)
            }
            float o[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            if (in_frame_r(P, x, y)) {
                const int ax = min(P.w - 1, x + RAD) - max(0, x - RAD) + 1;
                const int ay = min(P.frame_h - 1, yg + RAD) - max(0, yg - RAD) + 1;
                const float area = (float)(ax * ay);
                const float mr = __fdiv_rn((float)s[0], area), mg = __fdiv_rn((float)s[1], area), mb = __fdiv_rn((float)s[2], area);
                const double xx = (double)__fsub_rn(__fdiv_rn((float)s[3], area), __fmul_rn(mr, mr)) + P.eps;
                const double xy = __fsub_rn(__fdiv_rn((float)s[4], area), __fmul_rn(mr, mg));
                const double xz = __fsub_rn(__fdiv_rn((float)s[5], area), __fmul_rn(mr, mb));
                const double yy = (double)__fsub_rn(__fdiv_rn((float)s[6], area), __fmul_rn(mg, mg)) + P.eps;
                const double yz = __fsub_rn(__fdiv_rn((float)s[7], area), __fmul_rn(mg, mb));
                const double zz = (double)__fsub_rn(__fdiv_rn((float)s[8], area), __fmul_rn(mb, mb)) + P.eps;
                const double a00 = yy * zz - yz * yz, a01 = xz * yz - xy * zz, a02 = xy * yz - xz * yy;
                const double a11 = xx * zz - xz * xz, a12 = xy * xz - xx * yz, a22 = xx * yy - xy * xy;
                const double id = 1.0 / (xx * a00 + xy * a01 + xz * a02);
                const float rxy = __fmul_rn(__fmul_rn(__frcp_rn((float)ax), __frcp_rn(P.S * (float)ay)), P.scale);
                o[0] = mr; o[1] = mg; o[2] = mb;
                o[3] = __fmul_rn((float)(a00 * id), rxy);
                o[4] = __fmul_rn((float)(a01 * id), rxy);
                o[5] = __fmul_rn((float)(a02 * id), rxy);
                o[6] = __fmul_rn((float)(a11 * id), rxy);
                o[7] = __fmul_rn((float)(a12 * id), rxy);
                o[8] = __fmul_rn((float)(a22 * id), rxy);
            }
            float* dst = P.GB + ((size_t)strip * P.rows_pad + yrow) * (GB_ROW / 4);
            *reinterpret_cast<float4*>(dst + t * 4) = make_float4(o[0], o[1], o[2], o[3]);
            *reinterpret_cast<float4*>(dst + M_TW * 4 + t * 4) = make_float4(o[4], o[5], o[6], o[7]);
            dst[M_TW * 8 + t] = o[8];
        }
        add(y - RAD, -1);
    }
}

}  // namespace

int sbf_rgb_mma_supported(const sb200_params* p) { return sbf_mma_supported(p); }

size_t sbf_rgb_mma_workspace_bytes(const sb200_ctx* ctx, int w, int h_held, int rows_out, int dabs, int size_d) {
    Plan plan = make_plan_mma(w, rows_out, size_d, ctx->sm_count, 2, R_ND);
    const int rows_pad = h_held + 2 * PADY;
    const MmaGeom mg = mma_geom(plan.n_strips, -dabs - 8, dabs + 8);
    const int pitchS = (w + 3) / 4 * 4;
    const size_t recs = (size_t)plan.n_strips * rows_pad;
    size_t bytes = 0;
    bytes += 2 * (sb_align(recs * GA_ROW) + sb_align(recs * GB_ROW) + sb_align(recs * GC_ROW));
    bytes += 2 * sb_align((size_t)rows_pad * mg.n_chunk * 64);
    bytes += sb_align(((size_t)plan.n_chunks * 2 * rows_out + MR + 1) * pitchS * 8 + PB_ROW);
    return bytes + 8192;
}

int sbf_pair_disparity_rgb_mma(sb200_ctx* ctx, const sb200_params* p, const uint8_t* rgb_l, const uint8_t* rgb_r, int channels,
                               const uint8_t* gray_l, const uint8_t* gray_r, const SbFusedGeom& g, float* bestL, float* dispL,
                               float* bestR, float* dispR) {
    int nI, nG, S;
    if (!sbf_mma_supported(p) || !find_lattice(p, &nI, &nG, &S))
        return sb_fail(ctx, SB200_ERR_UNSUPPORTED, "MMA fused RGB kernel: radius 9 and a small exact cost lattice only");
    if (g.w < 2 || g.h < 1 || g.rows_out < 1) return sb_fail(ctx, SB200_ERR_INVALID, "fused RGB: bad shape");
    const int size_d = p->dmax - p->dmin + 1;
    const int dmin[2] = {p->dmin, -p->dmax};
    const int dabs = max(abs(p->dmin), abs(p->dmax));
    Plan plan = make_plan_mma(g.w, g.rows_out, size_d, ctx->sm_count, 2, R_ND);
    const int rows_pad = g.h + 2 * PADY;
    const MmaGeom mg = mma_geom(plan.n_strips, -dabs - 8, dabs + 8);
    const int pitchS = (g.w + 3) / 4 * 4;
    const size_t recs = (size_t)plan.n_strips * rows_pad;

    // power-of-two scale that keeps the vertical sums of a and b' inside fp16 (see fused_mma.cu; |a| is bounded per
    // principal direction of the colour covariance like the gray a)
    const double pmax = (nI * (double)p->th_color + nG * 2.0 * (double)p->th_grad) / S;
    const double amax = 0.5 * pmax / sqrt(p->eps > 1e-12 ? p->eps : 1e-12);
    const double bmax = pmax + 255.0 * 1.7320508 * amax;
    int sh = (int)floor(log2(30000.0 / bmax));
    if (sh > 24) sh = 24;
    if (sh < -24) sh = -24;
    const float scale = ldexpf(1.0f, sh);

    const uint8_t* gray[2] = {gray_l, gray_r};
    const uint8_t* rgb[2] = {rgb_l, rgb_r};
    uint4 *GA[2], *MT[2];
    float* GB[2];
    __half* GC[2];
    for (int i = 0; i < 2; i++) {
        GA[i] = sb_ws_alloc<uint4>(ctx, recs * M_KB);
        GB[i] = sb_ws_alloc<float>(ctx, recs * (GB_ROW / 4));
        GC[i] = sb_ws_alloc<__half>(ctx, recs * 3 * M_TW);
        MT[i] = sb_ws_alloc<uint4>(ctx, (size_t)rows_pad * mg.n_chunk * 4);
    }
    const size_t planeS = (size_t)g.rows_out * pitchS;
    float2* BL = sb_ws_alloc<float2>(ctx, planeS * 2 * plan.n_chunks + (size_t)(MR + 1) * pitchS + M_TW);
    if (!GA[0] || !GA[1] || !GB[0] || !GB[1] || !GC[0] || !GC[1] || !MT[0] || !MT[1] || !BL)
        return sb_fail(ctx, SB200_ERR_NOMEM, "fused RGB (mma): workspace arena too small (internal)");

    if (ctx->timing && ctx->ev_valid) SB_CUDA(ctx, cudaEventRecord(ctx->ev[0], ctx->stream));
    PrepR PR[2];
    PrepM QM[2];
    for (int i = 0; i < 2; i++) {
        PrepR& P = PR[i];
        P.gray = gray[i];
        P.rgb = rgb[i];
        P.ch = channels;
        P.w = g.w;
        P.h_held = g.h;
        P.y_global0 = g.y_global0;
        P.frame_h = g.frame_h;
        P.n_strips = plan.n_strips;
        P.rows_pad = rows_pad;
        P.GA = GA[i];
        P.GB = GB[i];
        P.GC = GC[i];
        P.eps = p->eps;
        P.S = (float)S;
        P.scale = scale;
        PrepM& Q = QM[i];
        Q.gray = gray[i];
        Q.w = g.w;
        Q.h_held = g.h;
        Q.y_global0 = g.y_global0;
        Q.frame_h = g.frame_h;
        Q.n_strips = plan.n_strips;
        Q.rows_pad = rows_pad;
        Q.GA = nullptr;
        Q.GB = nullptr;
        Q.GC = nullptr;
        Q.MT = MT[i];
        Q.n_chunk = mg.n_chunk;
        Q.padm = mg.padm;
        Q.mean_u8 = nullptr;
        Q.eps = p->eps;
        Q.S = (float)S;
        Q.scale = scale;
    }
    SB_LAUNCH(ctx, k_prep_ga3, dim3(sb_div_up(rows_pad, PREP_ROWS), plan.n_strips, 2), M_KB, 0, PR[0], PR[1]);
    SB_LAUNCH(ctx, k_prep_gb3, dim3(sb_div_up(rows_pad, GB3_TR), plan.n_strips, 2), GB3_T, 0, PR[0], PR[1]);
    SB_LAUNCH(ctx, k_prep_mt, dim3(sb_div_up(mg.n_chunk, 256), rows_pad, 2), 256, 0, QM[0], QM[1]);
    if (ctx->timing && ctx->ev_valid) SB_CUDA(ctx, cudaEventRecord(ctx->ev[1], ctx->stream));

    RgbMmaArgs A;
    for (int i = 0; i < 2; i++) {
        A.GA[i] = GA[i];
        A.GB[i] = GB[i];
        A.GC[i] = GC[i];
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
    A.n_views = 2;
    A.BL = BL;
    A.QV[0] = ctx->qvol[0];
    A.QV[1] = ctx->qvol[1];
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

    const size_t smem = sizeof(RSmem);
    if (!ctx->rgb_mma_attr_set) {
        SB_CUDA(ctx, cudaFuncSetAttribute(k_fused_mma_rgb<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        SB_CUDA(ctx, cudaFuncSetAttribute(k_fused_mma_rgb<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        ctx->rgb_mma_attr_set = true;
    }
    const int nblocks = plan.n_strips * plan.n_bands * plan.n_chunks * 2;
    if (A.QV[0] || A.QV[1]) SB_LAUNCH(ctx, k_fused_mma_rgb<true>, nblocks, R_THREADS, smem, A);
    else SB_LAUNCH(ctx, k_fused_mma_rgb<false>, nblocks, R_THREADS, smem, A);
    if (ctx->timing && ctx->ev_valid) SB_CUDA(ctx, cudaEventRecord(ctx->ev[2], ctx->stream));
    float* best[2] = {bestL, bestR};
    float* disp[2] = {dispL, dispR};
    for (int v = 0; v < 2; v++) {
        if (!best[v] && !disp[v]) continue;
        dim3 grid(sb_div_up(g.w, 256), g.rows_out);
        SB_LAUNCH(ctx, k_merge_chunks_bl, grid, 256, 0, BL, plan.n_chunks, v, g.rows_out, g.w, pitchS, best[v], disp[v]);
    }
    if (ctx->timing && ctx->ev_valid) SB_CUDA(ctx, cudaEventRecord(ctx->ev[3], ctx->stream));
    return SB200_OK;
}

#ifdef RGBM_TIMELINE
extern "C" int sb200_debug_timeline(long long* host, int n_bytes) {
    return (int)cudaMemcpyFromSymbol(host, g_tl, (size_t)n_bytes < sizeof(g_tl) ? (size_t)n_bytes : sizeof(g_tl));
}
#endif
