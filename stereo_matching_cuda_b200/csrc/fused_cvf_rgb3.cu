// fused_cvf_rgb3.cu -- the RGB-guide fused kernel in the three-stage form of fused_cvf.cu.
//
// Same arithmetic as fused_cvf_rgb.cu (colour guided filter of He et al. with the reference's cost
// costVolume.cu:163-190, box / border rules guidedFilter.cu:297-318 and WTA guidedFilter.cu:403-411),
// restructured the way the gray kernel was:
//  * a block is 4 TRIOS of warps (p, p+4, p+8: one SM sub-partition, one Tensor-Memory lane quarter),
//    one trio per disparity, one image row per pipeline iteration:
//      stage 0  lattice cost, vertical sums of p, R p, G p, B p (exact integers, FHFMA/FHADD accumulation),
//               their four horizontal sums                                         -> (S_p, S_Rp, S_Gp, S_Bp)
//      stage 1  3x3 solve a = M cov, b = mp - a.mu, vertical sums of a_r, a_g, a_b, b    -> (V_ar, V_ag, V_ab, V_b)
//      stage 2  their horizontal sums, q = SS_a . I + SS_b                               -> q ring (shared memory)
//      stage 3  (4 merge warps) merge of the block's 4 disparities into the running (best,label)
//    16 warps x 128 registers at launch; every role trades registers with setmaxnreg (152 / 168 / 152 / 40);
//    hand-offs through Tensor-Memory columns: named barriers (0 -> 1), mbarriers with two slots (1 -> 2);
//  * the guide operands of stages 0 and 1 -- (I,G) of the row and the colour rows entering and leaving the first
//    window; the 72 statistics words per lane (mu, scaled inverse covariance) -- are fetched ONCE per block by
//    TMA bulk copies from strip-tiled planes into a 4-slot shared-memory ring (13 KB per iteration; the stage-1
//    warp of trio p fills slot p) and read with conflict-free 128-bit loads by the 8 warps of the two stages;
//    through L1 go only the match operands (two aligned 128-bit loads from the (I,G) copy shifted by d & 3) and
//    stage 2's colour row (three 128-bit loads, one row ahead);
//  * rings of the marching filters: (a_r, a_g) and the cost in Tensor Memory, (a_b, b) in shared memory;
//  * (best,label) as one interleaved float2 plane per disparity chunk, merged by k_merge_chunks_bl.
#include "fused_dev.cuh"

namespace {

struct Rgb3Args {
    const unsigned* IG[2];  // per IMAGE: padded half2 (I, G); + s*shift_stride = the plane moved left by s elements,
    size_t shift_stride;    // de-interleaved (fused_dev.cuh, match_ptrs)
    size_t ig_origin;       // linear element index of pixel (0,0) in a copy
    // strip-tiled planes, record (strip, padded row) = [16-byte chunk][lane][16 B] (see fused_cvf.cu, k_prep)
    const uint4* Tg[2];     // per IMAGE: (I,G) half2, 2 chunks, 1 KB per record
    const uint4* TC[2];     // per IMAGE: colour as halves, chunk c = channel c of the lane's 8 pixels, 1.5 KB per record
    const uint4* TS[2];     // per IMAGE: chunk 2j = (mu_r, mu_g, mu_b, M_rr), 2j+1 = (M_rg, M_rb, M_gg, M_gb) of pixel j,
                            // chunks 16, 17 = M_bb of pixels 0..3, 4..7; M scaled by 1/(S*area); 9 KB per record
    int rows_pad, pitch, w, y_out0, rows_out, y_global0, frame_h;
    int dmin[2];
    int size_d;
    int n_strips, n_bands, band_rows, n_chunks, chunk_d, n_views;
    float2* BL;             // [chunk][view][rows_out][pitchS] running (best cost, label)
    int pitchS;
    float S;
    unsigned wpack, thpack;
    int zero;               // always 0, opaque to the compiler
};

constexpr int R3_THREADS = 4 * NWARP * 32;  // three pipeline stages per disparity (warps p, p+4, p+8) + 4 merge warps
// Register budget per warp role (16 warps x 128 registers at launch; the roles trade registers with setmaxnreg, one
// warpgroup = the 4 warps of a role): 152 + 168 + 152 + 40 = 512 = 4 x 128
constexpr int R3_REGS0 = 152, R3_REGS1 = 168, R3_REGS2 = 152, R3_REGS3 = 40;
static_assert(R3_REGS0 + R3_REGS1 + R3_REGS2 + R3_REGS3 <= 512, "register file: 64 K registers per SM");
constexpr uint32_t T3_RING_A = 0, T3_RING_P = 304, T3_HAND = 384, T3_HAND2 = 416;  // TMEM columns of a lane
constexpr int NQ3 = 4;     // depth of the ring of filtered rows between stage 2 and the merge warps
// Operand ring of stages 0 and 1: slot K & 3 holds the guide operands of iteration K.  The stage-1 warp of trio p
// fills slot p (iterations K = p mod 4), RGB3_AHEAD iterations ahead of its own position, so the cost of issuing
// the bulk copies (~100 instructions) is shared by the four trios instead of slowing one of them -- the trios run in
// step through the merge ring, so the slowest one sets the pace.  Measured alternatives (1080p D=256): stage 2's
// colour row in the slot 8.56 ms (the ring can then only be refilled as fast as the LAST stage releases it); all
// fills by trio 0's stage-1 warp 8.29; by trio 0's stage-0 warp 9.45; one ring per stage, each filled by its own
// stage's trio-0 warp 9.19.
constexpr int NS3 = 4;
#ifndef RGB3_AHEAD
#define RGB3_AHEAD 2
#endif
constexpr int AHEAD3 = RGB3_AHEAD;
constexpr int WARM3 = 4 * RAD;   // warm-up iterations (one row each) before the first output row

struct Slot3 {          // guide operands of one iteration (one row) of stages 0 and 1, 4 bulk copies (13 KB)
    uint4 g[2][32];     // (I,G) at row yi
    uint4 cn[3][32];    // colour at row yi        (enters the first-stage window)
    uint4 co[3][32];    // colour at row yi-19     (leaves it)
    uint4 st[18][32];   // statistics at row ya = yi-9
};
constexpr uint32_t SLOT3_BYTES = sizeof(Slot3);
// (stage 2 reads the colour of its output row straight from the strip-tiled plane, one row ahead)
struct Smem3 {
    Slot3 slot[NS3];                   // operand ring of stages 0 and 1
    float4 ringB[NWARP][WIN][4][32];   // stage-1 private rings: a_b (planes 0,1) and b (planes 2,3)
    float4 qbuf[NQ3][NWARP][2][QV];    // filtered row of each stage-2 warp
    uint64_t sfull[NS3], sempty[NS3];
    uint64_t qfull[NQ3], qempty[NQ3];
    uint64_t full2[NWARP][2], empty2[NWARP][2];
    float ry_lut[2][WIN + 1];
    uint32_t tmem_base;
};
static_assert(sizeof(Smem3) <= 232448, "Smem3 exceeds the 227 KB a block may opt in to");

__device__ __forceinline__ unsigned word_of(const uint4& v, int k) { return k == 0 ? v.x : k == 1 ? v.y : k == 2 ? v.z : v.w; }

__global__ void __launch_bounds__(R3_THREADS, 1) k_fused_cvf_rgb3(const Rgb3Args A) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Smem3& sm = *reinterpret_cast<Smem3*>(smem_raw);

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int pair = warp & (NWARP - 1);
    const int stage = warp / NWARP;
    int bid = blockIdx.x;
    const int view = bid % A.n_views;
    bid /= A.n_views;
    const int chunk = bid % A.n_chunks;
    bid /= A.n_chunks;
    const int band = bid % A.n_bands;
    const int strip = bid / A.n_bands;

    const int xs = strip * VALID_W - HALO;
    const int xl = xs + KPX * lane;
    const int pitch = A.pitch;
    const int dlo = A.dmin[view] + chunk * A.chunk_d;
    const int dcnt = min(A.chunk_d, A.size_d - chunk * A.chunk_d);
    const int ngroups = (dcnt + NWARP - 1) / NWARP;
    const int yb0 = A.y_out0 + band * A.band_rows;
    const int yb1 = min(yb0 + A.band_rows, A.y_out0 + A.rows_out);
    const int y_first = yb0 - 2 * RAD;
    const int niter = (yb1 - yb0) + WARM3;  // one row per iteration
    const int n_emit = niter - WARM3;
    const int BAR_FULL = 2 + pair, BAR_EMPTY = 2 + NWARP + pair;
    int zero = A.zero;
    asm volatile("" : "+r"(zero));

    if (warp == 0) tm_alloc(&sm.tmem_base);
    const uint32_t mb_qfull = smem_addr(&sm.qfull[0]), mb_qempty = smem_addr(&sm.qempty[0]);
    const uint32_t mb_full2 = smem_addr(&sm.full2[pair][0]), mb_empty2 = smem_addr(&sm.empty2[pair][0]);
    const uint32_t mb_sfull = smem_addr(&sm.sfull[0]), mb_sempty = smem_addr(&sm.sempty[0]);
    const uint32_t slot0 = smem_addr(&sm.slot[0]) + 16 * lane;
    static_assert((NS3 & (NS3 - 1)) == 0 && (NQ3 & (NQ3 - 1)) == 0, "ring indices are taken from the bits of the counters");
    static_assert(AHEAD3 < NS3, "the ring cannot be filled further ahead than it is deep");
    static_assert(NS3 == NWARP, "trio p fills slot p");
    if (threadIdx.x == 0) {
#pragma unroll
        for (int b = 0; b < NS3; b++) {
            mbar_init(mb_sfull + 8 * b, 1);
            mbar_init(mb_sempty + 8 * b, 2 * NWARP);  // released by the stage-0 and stage-1 warps
        }
    }
    if (threadIdx.x == 32) {
#pragma unroll
        for (int b = 0; b < NQ3; b++) {
            mbar_init(mb_qfull + 8 * b, NWARP);
            mbar_init(mb_qempty + 8 * b, NWARP);
        }
#pragma unroll
        for (int b = 0; b < NWARP; b++) {
            for (int k = 0; k < 2; k++) {
                mbar_init(smem_addr(&sm.full2[b][k]), 1);
                mbar_init(smem_addr(&sm.empty2[b][k]), 1);
            }
        }
    }
    if (threadIdx.x >= 64 && threadIdx.x < 64 + 2 * (WIN + 1)) {
        const int t = threadIdx.x - 64, n = t % (WIN + 1);
        sm.ry_lut[t / (WIN + 1)][n] = (n == 0) ? 0.0f : __frcp_rn((t < WIN + 1 ? A.S : 1.0f) * (float)n);
    }
    tm_fence_before();
    __syncthreads();
    tm_fence_after();
    const uint32_t tbase = sm.tmem_base + ((uint32_t)(pair * 32) << 16);
    const uint32_t tA = tbase + T3_RING_A, tP = tbase + T3_RING_P, tH = tbase + T3_HAND, tH2 = tbase + T3_HAND2;

    // operand ring, consumer side: a warp waits for the slot of its iteration K, reads its 16-byte chunks and releases
    // the slot; `t` is a word of the data read, so the release cannot be issued before the reads have completed
    auto slot_wait = [&](int K) -> uint32_t {
        mbar_wait(mb_sfull + 8 * (K & (NS3 - 1)), (unsigned)(K / NS3) & 1u);
        return slot0 + (uint32_t)(K & (NS3 - 1)) * SLOT3_BYTES;
    };
    auto slot_release = [&](int K, unsigned t) {
        mbar_arrive_lane0(mb_sempty + 8 * (K & (NS3 - 1)) + (t & (unsigned)zero), lane);
    };
    constexpr uint32_t OFF_G = offsetof(Slot3, g), OFF_CN = offsetof(Slot3, cn), OFF_CO = offsetof(Slot3, co),
                       OFF_ST = offsetof(Slot3, st);
    // operand ring, producer side (stage-1 warps): fill the slot of iteration Kf, row iteration itf of its group, with
    // 4 TMA bulk copies of the strip-tiled planes
    const int Ktotal = ngroups * niter;
    auto fill_slot = [&](int Kf, int itf) {
        const int sl = Kf & (NS3 - 1);
        if (Kf >= NS3) mbar_wait(mb_sempty + 8 * sl, (unsigned)(Kf / NS3 - 1) & 1u);  // its 8 readers released it
        if (lane == 0) {
            const uint32_t full = mb_sfull + 8 * sl;
            const uint32_t dst = smem_addr(&sm.slot[0]) + (uint32_t)sl * SLOT3_BYTES;
            const long long rec = (long long)strip * A.rows_pad + PADY + y_first + itf;  // record of row yi
            mbar_expect_tx(full, SLOT3_BYTES);
            bulk_g2s(dst + OFF_G, A.Tg[view] + rec * 64, 1024, full);
            bulk_g2s(dst + OFF_CN, A.TC[view] + rec * 96, 1536, full);
            bulk_g2s(dst + OFF_CO, A.TC[view] + (rec - WIN) * 96, 1536, full);
            bulk_g2s(dst + OFF_ST, A.TS[view] + (rec - RAD) * 576, 9216, full);
        }
        __syncwarp();
    };
    // this warp's next fill: iteration fillK (row iteration fill_it of its group); after a fill it moves on by 4
    // iterations (trio p fills the iterations K = p mod 4)
    constexpr int fill_step = NWARP;
    int fillK = pair, fill_it = fillK;
    const bool filler = (stage == 1);  // (the other roles never call fill_due)
    // called by a filling warp at the start of its iteration K: fill what is due up to K + AHEAD3
    auto fill_due = [&](int K) {
        while (fillK <= K + AHEAD3 && fillK < Ktotal) {
            fill_slot(fillK, fill_it);
            fillK += fill_step;
            fill_it += fill_step;
            if (fill_it >= niter) fill_it -= niter;
        }
    };
    if (filler) fill_due(-1);  // the slots of iterations 0 .. AHEAD3-1

    // a = M cov, b = mp - a.mu at row ya = yi - 9 from the four first-stage sums; the statistics come from the operand
    // ring, half a lane's pixels at a time.  On return the arrays hold (a_r, a_g, a_b, b).
    auto solve = [&](int K, float (&SP)[KPX], float (&SR)[KPX], float (&SG)[KPX], float (&SB)[KPX], const float (&rx)[KPX],
                     float ry1, uint32_t sa, unsigned t) {
#pragma unroll
        for (int hq = 0; hq < 2; hq++) {
            uint4 q1[4], q2[4];
#pragma unroll
            for (int jj = 0; jj < 4; jj++) {
                q1[jj] = lds128(sa + OFF_ST + (2 * (4 * hq + jj)) * 512);
                q2[jj] = lds128(sa + OFF_ST + (2 * (4 * hq + jj) + 1) * 512);
                t |= q1[jj].x | q2[jj].x;
            }
            const uint4 q3 = lds128(sa + OFF_ST + (16 + hq) * 512);
            t |= q3.x;
#pragma unroll
            for (int jj = 0; jj < 4; jj++) {
                const int j = 4 * hq + jj;
                const float mr = __uint_as_float(q1[jj].x), mg = __uint_as_float(q1[jj].y), mb = __uint_as_float(q1[jj].z);
                const float Mrr = __uint_as_float(q1[jj].w);
                const float Mrg = __uint_as_float(q2[jj].x), Mrb = __uint_as_float(q2[jj].y);
                const float Mgg = __uint_as_float(q2[jj].z), Mgb = __uint_as_float(q2[jj].w);
                const float Mbb = __uint_as_float(word_of(q3, jj));
                const float cx = fmaf(-mr, SP[j], SR[j]);
                const float cy = fmaf(-mg, SP[j], SG[j]);
                const float cz = fmaf(-mb, SP[j], SB[j]);
                const float ar = fmaf(Mrr, cx, fmaf(Mrg, cy, Mrb * cz));
                const float ag = fmaf(Mrg, cx, fmaf(Mgg, cy, Mgb * cz));
                const float ab = fmaf(Mrb, cx, fmaf(Mgb, cy, Mbb * cz));
                const float mp = SP[j] * (rx[j] * ry1);
                SP[j] = ar;
                SR[j] = ag;
                SG[j] = ab;
                SB[j] = mp - fmaf(ar, mr, fmaf(ag, mg, ab * mb));
            }
        }
        slot_release(K, t);
    };

    if (stage == 0) {
        // ====== STAGE 0: lattice cost; vertical and horizontal window sums of p, R p, G p, B p ======
        reg_inc<R3_REGS0>();
        const unsigned* __restrict__ IGm = A.IG[1 - view];
        __half2 wm[KPX];
#pragma unroll
        for (int j = 0; j < KPX; j++) {
            int x = xl + j;
            wm[j] = (x >= 0 && x < A.w) ? u2h2(A.wpack) : __float2half2_rn(0.0f);
        }
        const __half2 th = u2h2(A.thpack);
        for (int g = 0; g < ngroups; g++) {
            const int dk = g * NWARP + pair;
            const bool active = dk < dcnt;
            const int d = dlo + dk;
            float VP[KPX], VR[KPX], VG[KPX], VB[KPX];
#pragma unroll
            for (int j = 0; j < KPX; j++) VP[j] = VR[j] = VG[j] = VB[j] = 0.0f;
            {
                const uint32_t z[4] = {0u, 0u, 0u, 0u};
                for (int s = 0; s < WIN; s++) tm_st4(tP + 4 * s, z);
                tm_wait_st();
            }
            __syncthreads();  // group start
            if (active) {
                const long long r0 = (long long)y_first * pitch + xl;
                const unsigned *pm0, *pm1;
                match_ptrs(IGm + (size_t)(d & 3) * A.shift_stride, (long long)A.ig_origin + r0 + (d - (d & 3)),
                           A.shift_stride / 2, pm0, pm1);
                uint4 m0 = __ldg(reinterpret_cast<const uint4*>(pm0));
                uint4 m1 = __ldg(reinterpret_cast<const uint4*>(pm1));
                pm0 += pitch / 2;
                pm1 += pitch / 2;
                int slot = 0;
#pragma unroll 1
                for (int it = 0; it < niter; it++) {
                    // next row's match operands, issued behind the wait for this row's (see touch() in fused_cvf.cu)
                    const int dep = (int)(m0.x | m1.x) & zero;
                    const uint4 n0 = __ldg(reinterpret_cast<const uint4*>(pm0 + dep));
                    const uint4 n1 = __ldg(reinterpret_cast<const uint4*>(pm1 + dep));
                    pm0 += pitch / 2;
                    pm1 += pitch / 2;
                    uint4 gq[2], cn[3], co[3];
                    const int K = g * niter + it;
                    const uint32_t sa = slot_wait(K);
                    unsigned t = 0;
                    {
                        gq[0] = lds128(sa + OFF_G);
                        gq[1] = lds128(sa + OFF_G + 512);
                        t |= gq[0].x | gq[1].x;
#pragma unroll
                        for (int c = 0; c < 3; c++) {
                            cn[c] = lds128(sa + OFF_CN + c * 512);
                            co[c] = lds128(sa + OFF_CO + c * 512);
                            t |= cn[c].x | co[c].x;
                        }
                        slot_release(K, t);
                    }
                    uint32_t pold[4];  // the NEGATED lattice costs of row yi-19
                    tm_ld4(tP + 4 * slot, pold);
                    const unsigned gg[KPX] = {gq[0].x, gq[0].y, gq[0].z, gq[0].w, gq[1].x, gq[1].y, gq[1].z, gq[1].w};
                    const unsigned mm[KPX] = {m0.x, m0.y, m0.z, m0.w, m1.x, m1.y, m1.z, m1.w};
                    __half ph[KPX];
#pragma unroll
                    for (int j = 0; j < KPX; j++) {
                        __half2 diff = __hsub2(u2h2(gg[j]), u2h2(mm[j]));
                        __half2 c = __hmin2(__habs2(diff), th);
                        __half2 pr = __hmul2(c, wm[j]);
                        ph[j] = __hadd(__low2half(pr), __high2half(pr));
                    }
                    tm_wait_ld();
                    uint32_t pneg[4];
#pragma unroll
                    for (int k = 0; k < KPX / 2; k++) {
                        const __half2 pn2 = __halves2half2(ph[2 * k], ph[2 * k + 1]);
                        const __half2 po2 = u2h2(pold[k]);
                        pneg[k] = h22u(__hneg2(pn2));
                        const __half2 dp = __hadd2(pn2, po2);  // exact: |.| <= 2^11
                        VP[2 * k] = fhadd(__low2half(dp), VP[2 * k]);
                        VP[2 * k + 1] = fhadd(__high2half(dp), VP[2 * k + 1]);
                        const __half2 rn = u2h2(word_of(cn[0], k)), gn = u2h2(word_of(cn[1], k)), bn = u2h2(word_of(cn[2], k));
                        const __half2 ro = u2h2(word_of(co[0], k)), go = u2h2(word_of(co[1], k)), bo = u2h2(word_of(co[2], k));
                        VR[2 * k] = fhfma(__low2half(rn), ph[2 * k], VR[2 * k]);
                        VR[2 * k + 1] = fhfma(__high2half(rn), ph[2 * k + 1], VR[2 * k + 1]);
                        VG[2 * k] = fhfma(__low2half(gn), ph[2 * k], VG[2 * k]);
                        VG[2 * k + 1] = fhfma(__high2half(gn), ph[2 * k + 1], VG[2 * k + 1]);
                        VB[2 * k] = fhfma(__low2half(bn), ph[2 * k], VB[2 * k]);
                        VB[2 * k + 1] = fhfma(__high2half(bn), ph[2 * k + 1], VB[2 * k + 1]);
                        VR[2 * k] = fhfma(__low2half(ro), __low2half(po2), VR[2 * k]);
                        VR[2 * k + 1] = fhfma(__high2half(ro), __high2half(po2), VR[2 * k + 1]);
                        VG[2 * k] = fhfma(__low2half(go), __low2half(po2), VG[2 * k]);
                        VG[2 * k + 1] = fhfma(__high2half(go), __high2half(po2), VG[2 * k + 1]);
                        VB[2 * k] = fhfma(__low2half(bo), __low2half(po2), VB[2 * k]);
                        VB[2 * k + 1] = fhfma(__high2half(bo), __high2half(po2), VB[2 * k + 1]);
                    }
                    tm_st4(tP + 4 * slot, pneg);
                    slot = (slot + 1 == WIN) ? 0 : slot + 1;
                    float SP[KPX], SR[KPX], SG[KPX], SB[KPX];
                    hsum19(VP, SP);
                    hsum19(VR, SR);
                    hsum19(VG, SG);
                    hsum19(VB, SB);
                    if (it > 0) {  // stage 1 has copied the previous row out of the hand-off columns
                        named_bar_sync(BAR_EMPTY, 64);
                        tm_fence_after();
                    }
                    tm_st16(tH, SP, SR);
                    tm_st16(tH + 16, SG, SB);
                    tm_wait_st();
                    tm_fence_before();
                    named_bar_arrive(BAR_FULL, 64);
                    m0 = n0;
                    m1 = n1;
                }
            } else {
                for (int it = 0; it < niter; it++) {
                    slot_wait(g * niter + it);
                    slot_release(g * niter + it, 0u);
                }
            }
            __syncthreads();  // group end
        }
    } else if (stage == 1) {
        // ====== STAGE 1: a = M cov, b = mp - a.mu; their vertical sums; horizontal sums of a_r, a_g ======
        reg_inc<R3_REGS1>();
        float rx[KPX];
#pragma unroll
        for (int j = 0; j < KPX; j++) {
            int x = xl + j;
            int ax = min(A.w - 1, x + RAD) - max(0, x - RAD) + 1;
            rx[j] = (x >= 0 && x < A.w) ? __frcp_rn((float)ax) : 0.0f;
        }
        for (int g = 0; g < ngroups; g++) {
            const int dk = g * NWARP + pair;
            const bool active = dk < dcnt;
            float Var[KPX], Vag[KPX], Vab[KPX], Vb[KPX];
#pragma unroll
            for (int j = 0; j < KPX; j++) Var[j] = Vag[j] = Vab[j] = Vb[j] = 0.0f;
            for (int s = 0; s < WIN; s++) {
                tm_st16(tA + 16 * s, Var, Vag);  // zeros
#pragma unroll
                for (int v = 0; v < 4; v++) sm.ringB[pair][s][v][lane] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
            tm_wait_st();
            __syncthreads();  // group start

            int slot = 0;
            auto iter = [&](auto emit_tag, int it) {
                constexpr bool EMIT = decltype(emit_tag)::value;
                if (filler) fill_due(g * niter + it);
                const float ry1 = inv_rows(sm.ry_lut[0], y_first + it - RAD, A.y_global0, A.frame_h);
                named_bar_sync(BAR_FULL, 64);  // stage 0 has published row yi
                tm_fence_after();
                float SP[KPX], SR[KPX], SG[KPX], SB[KPX];
                tm_ld16(tH, SP, SR);
                tm_ld16(tH + 16, SG, SB);
                tm_wait_ld();
                if (it + 1 < niter) {
                    tm_fence_before();
                    named_bar_arrive(BAR_EMPTY, 64);
                }
                // ---- (a, b) at row ya = yi - 9
                {
                    const int K = g * niter + it;
                    solve(K, SP, SR, SG, SB, rx, ry1, slot_wait(K), 0u);
                }
                float (&ar)[KPX] = SP, (&ag)[KPX] = SR, (&ab)[KPX] = SG, (&bb)[KPX] = SB;
                // ---- second stage: row ya enters, row ya-19 leaves (read from the rings only now: the solve above
                //      needs the registers)
                float aro[KPX], ago[KPX];
                tm_ld16(tA + 16 * slot, aro, ago);
                const float4 ob0 = sm.ringB[pair][slot][0][lane], ob1 = sm.ringB[pair][slot][1][lane];
                const float4 ob2 = sm.ringB[pair][slot][2][lane], ob3 = sm.ringB[pair][slot][3][lane];
                tm_wait_ld();
                tm_st16(tA + 16 * slot, ar, ag);
                sm.ringB[pair][slot][0][lane] = make_float4(ab[0], ab[1], ab[2], ab[3]);
                sm.ringB[pair][slot][1][lane] = make_float4(ab[4], ab[5], ab[6], ab[7]);
                sm.ringB[pair][slot][2][lane] = make_float4(bb[0], bb[1], bb[2], bb[3]);
                sm.ringB[pair][slot][3][lane] = make_float4(bb[4], bb[5], bb[6], bb[7]);
                slot = (slot + 1 == WIN) ? 0 : slot + 1;
                const float abo[KPX] = {ob0.x, ob0.y, ob0.z, ob0.w, ob1.x, ob1.y, ob1.z, ob1.w};
                const float bbo[KPX] = {ob2.x, ob2.y, ob2.z, ob2.w, ob3.x, ob3.y, ob3.z, ob3.w};
#pragma unroll
                for (int j = 0; j < KPX; j++) {
                    Var[j] += ar[j] - aro[j];
                    Vag[j] += ag[j] - ago[j];
                    Vab[j] += ab[j] - abo[j];
                    Vb[j] += bb[j] - bbo[j];
                }
                if (EMIT) {  // (the four horizontal sums are taken by stage 2: the hand-off carries the vertical sums)
                    const int E = g * n_emit + (it - WARM3);
                    if (E > 1) {  // stage 2 has copied emission E-2 out of this hand-off slot
                        mbar_wait(mb_empty2 + 8 * (E & 1), (unsigned)(E / 2 - 1) & 1u);
                        tm_fence_after();
                    }
                    tm_st16(tH2 + 32 * (E & 1), Var, Vag);
                    tm_st16(tH2 + 32 * (E & 1) + 16, Vab, Vb);
                    tm_wait_st();
                    tm_fence_before();
                    __syncwarp();
                    mbar_arrive_lane0(mb_full2 + 8 * (E & 1), lane);
                } else {
                    tm_wait_st();
                }
            };
            if (active) {
                int it = 0;
#pragma unroll 1
                for (; it < WARM3; it++) iter(std::false_type{}, it);
#pragma unroll 1
                for (; it < niter; it++) iter(std::true_type{}, it);
            } else {
                for (int it = 0; it < niter; it++) {  // a trio without a disparity still fills and releases its share
                    if (filler) fill_due(g * niter + it);
                    slot_wait(g * niter + it);
                    slot_release(g * niter + it, 0u);
                }
            }
            __syncthreads();  // group end
        }
    } else if (stage == 2) {
        // ====== STAGE 2: the four horizontal sums of a_r, a_g, a_b, b; q = SS_a . I + SS_b ======
        reg_inc<R3_REGS2>();
        float rx[KPX];
#pragma unroll
        for (int j = 0; j < KPX; j++) {
            int x = xl + j;
            int ax = min(A.w - 1, x + RAD) - max(0, x - RAD) + 1;
            rx[j] = (x >= 0 && x < A.w) ? __frcp_rn((float)ax) : 0.0f;
        }
        for (int g = 0; g < ngroups; g++) {
            const int dk = g * NWARP + pair;
            const bool active = dk < dcnt;
            if (!active) {  // the previous group's merges have drained (group-end barrier)
                const float inf = __int_as_float(0x7f800000);
#pragma unroll
                for (int b = 0; b < NQ3; b++)
#pragma unroll
                    for (int v = 0; v < 2; v++) sm.qbuf[b][pair][v][lane] = make_float4(inf, inf, inf, inf);
            }
            __syncthreads();  // group start
            if (active) {
                // colour at the output row yq = yb0 + e: three 128-bit loads per lane from the strip-tiled plane, fetched
                // one row ahead (the next loads are issued behind the wait for the current ones, see touch())
                const uint4* pc = A.TC[view] + ((long long)strip * A.rows_pad + PADY + yb0) * 96 + lane;
                uint4 cq[3];
#pragma unroll
                for (int c = 0; c < 3; c++) cq[c] = __ldg(pc + c * 32);
#pragma unroll 1
                for (int e = 0; e < n_emit; e++) {
                    uint4 cqn[3];
                    {
                        pc += 96;
                        const int dep = (int)(cq[0].x | cq[1].x | cq[2].x) & zero;
#pragma unroll
                        for (int c = 0; c < 3; c++) cqn[c] = __ldg(pc + c * 32 + dep);
                    }
                    const int E = g * n_emit + e;
                    mbar_wait(mb_full2 + 8 * (E & 1), (unsigned)(E / 2) & 1u);  // stage 1 has published this emission
                    tm_fence_after();
                    float SAr[KPX], SAg[KPX], Vab[KPX], Vb[KPX];
                    tm_ld16(tH2 + 32 * (E & 1), SAr, SAg);
                    tm_ld16(tH2 + 32 * (E & 1) + 16, Vab, Vb);
                    tm_wait_ld();
                    tm_fence_before();
                    __syncwarp();
                    mbar_arrive_lane0(mb_empty2 + 8 * (E & 1), lane);
                    const int qb = E & (NQ3 - 1);
                    if (E >= NQ3) mbar_wait(mb_qempty + 8 * qb, (unsigned)(E / NQ3 - 1) & 1u);  // merged NQ3 emissions ago
                    float SAb[KPX], SBb[KPX];
                    {
                        float t[KPX];
                        hsum19(SAr, t);
#pragma unroll
                        for (int j = 0; j < KPX; j++) SAr[j] = t[j];
                    }
                    {
                        float t[KPX];
                        hsum19(SAg, t);
#pragma unroll
                        for (int j = 0; j < KPX; j++) SAg[j] = t[j];
                    }
                    hsum19(Vab, SAb);
                    hsum19(Vb, SBb);
                    const float ry2 = inv_rows(sm.ry_lut[1], yb0 + e, A.y_global0, A.frame_h);
                    float q[KPX];
#pragma unroll
                    for (int k = 0; k < KPX / 2; k++) {
                        const float2 r2 = __half22float2(u2h2(word_of(cq[0], k)));
                        const float2 g2 = __half22float2(u2h2(word_of(cq[1], k)));
                        const float2 b2 = __half22float2(u2h2(word_of(cq[2], k)));
                        q[2 * k] = fmaf(SAr[2 * k], r2.x, fmaf(SAg[2 * k], g2.x, fmaf(SAb[2 * k], b2.x, SBb[2 * k]))) *
                                   (rx[2 * k] * ry2);
                        q[2 * k + 1] = fmaf(SAr[2 * k + 1], r2.y, fmaf(SAg[2 * k + 1], g2.y, fmaf(SAb[2 * k + 1], b2.y, SBb[2 * k + 1]))) *
                                       (rx[2 * k + 1] * ry2);
                    }
                    sm.qbuf[qb][pair][0][lane] = make_float4(q[0], q[1], q[2], q[3]);
                    sm.qbuf[qb][pair][1][lane] = make_float4(q[4], q[5], q[6], q[7]);
                    __syncwarp();
                    mbar_arrive_lane0(mb_qfull + 8 * qb, lane);
#pragma unroll
                    for (int c = 0; c < 3; c++) cq[c] = cqn[c];
                }
            } else {
                // no disparity for this trio in the (last, partial) group: its q slots hold +inf
                for (int e = 0; e < n_emit; e++) {
                    const int E = g * n_emit + e;
                    const int qb = E & (NQ3 - 1);
                    if (E >= NQ3) mbar_wait(mb_qempty + 8 * qb, (unsigned)(E / NQ3 - 1) & 1u);
                    if (lane == 0) mbar_arrive(mb_qfull + 8 * qb);
                    __syncwarp();
                }
            }
            __syncthreads();  // group end
        }
    } else {
        // ====== STAGE 3: merge of the block's 4 disparities into the running (best,label) ======
        // Thread t of these 4 warps folds strip-local columns 2t, 2t+1 of the rows the stage-2 warps publish in the q
        // ring; the (best,label) of the previous groups is fetched two rows ahead.
        reg_dec<R3_REGS3>();
        const int mc = 2 * (threadIdx.x - 3 * NWARP * 32);
        const int mx = xs + mc;
        const bool mvalid = (mc >= HALO) && (mc < HALO + VALID_W) && (mx < A.w);
        const int qoff = (((mc & 7) >> 2) * QV + (mc >> 3)) * 4 + (mc & 3);
        const size_t planeS = (size_t)A.rows_out * A.pitchS;
        float2* __restrict__ BL = A.BL + (size_t)(chunk * 2 + view) * planeS;
        const size_t bl_row = (size_t)A.pitchS / 2;  // 16-byte units per row of the plane
        float4* const bl0 = reinterpret_cast<float4*>(BL + (size_t)(yb0 - A.y_out0) * A.pitchS + mx);
        for (int g = 0; g < ngroups; g++) {
            const int dbase = dlo + g * NWARP;
            const float lab[NWARP] = {(float)dbase, (float)(dbase + 1), (float)(dbase + 2), (float)(dbase + 3)};
            const bool ld_ok = (g > 0) && mvalid;
            auto prefetch_at = [&](int r) -> float4 {  // row r of the band
                float4 pb = make_float4(BEST_INIT_BITS_F, 0.0f, BEST_INIT_BITS_F, 0.0f);
                if (ld_ok && r < n_emit) pb = ld_early_f4(bl0 + (size_t)r * bl_row);
                return pb;
            };
            __syncthreads();  // group start
            float4 pbA = prefetch_at(0), pbB = prefetch_at(1);
            float4* blp = bl0;
#pragma unroll 1
            for (int e = 0; e < n_emit; e++) {
                const float4 pb = pbA;
                pbA = pbB;
                pbB = prefetch_at(e + 2);
                const int E = g * n_emit + e;
                const int qb = E & (NQ3 - 1);
                mbar_wait(mb_qfull + 8 * qb, (unsigned)(E / NQ3) & 1u);
                const float4 nb = merge4(reinterpret_cast<const float*>(&sm.qbuf[qb][0][0][0]) + qoff, 2 * QV * 4, lab, pb);
                if (mvalid) *blp = nb;
                __syncwarp();
                mbar_arrive_lane0(mb_qempty + 8 * qb, lane);
                blp += bl_row;
            }
            __syncthreads();  // group end
        }
    }
    tm_fence_before();
    __syncthreads();
    if (warp == 0) {
        tm_fence_after();
        tm_dealloc(sm.tmem_base);
    }
}

// ---------------------------------------------------------------------------------------
// Per-frame preparation: the strip-tiled colour plane TC and statistics plane TS of one image.  Same arithmetic as
// k_prep_rgb (fused_cvf_rgb.cu): exact integer window sums of R, G, B and their six products, covariance in float
// as the staged path forms it, adjugate inverse in double, scaled by 1/(S*area).
constexpr int RPT3 = 16;
constexpr int RPP3 = RPT3 + 2 * RAD;  // 34

struct Rgb3PrepArgs {
    const uint8_t* rgb;  // held rows, interleaved, `ch` bytes per pixel
    int ch, w, h_held, y_global0, frame_h;
    __half* TC;
    float* TS;
    int n_strips, rows_pad;
    int pitch, padx;  // extent of the padded frame the tiles cover
    double eps;
    float S;
};

// (the index arithmetic of tc_index / ts_s1_index / ts_s2_index / ts_s3_index in fused_dev.cuh, which
// tests/layout_check.cu checks against the pixel order the kernel assumes, written out per record)
__device__ __forceinline__ void store_tiled3(const Rgb3PrepArgs& P, int x, int yrow, __half r, __half g, __half b,
                                             float4 s1, float4 s2, float s3) {
    const int xa = x + HALO;
    if (xa < 0) return;
    const int sA = xa / VALID_W;
    const int cA = xa - sA * VALID_W;
#pragma unroll
    for (int k = 0; k < 2; k++) {  // the 40 columns two neighbouring strips share are stored in both
        const int s = sA - k, cl = cA + k * VALID_W;
        if (s < 0 || s >= P.n_strips || cl >= SW) continue;
        const size_t rec = (size_t)s * P.rows_pad + yrow;
        const int L = cl >> 3, j = cl & 7;
        __half* tc = P.TC + rec * 768;  // 3 chunks x 32 lanes x 8 halves
        tc[(0 * 32 + L) * 8 + j] = r;
        tc[(1 * 32 + L) * 8 + j] = g;
        tc[(2 * 32 + L) * 8 + j] = b;
        float* ts = P.TS + rec * 2304;  // 18 chunks x 32 lanes x 4 floats
        reinterpret_cast<float4*>(ts)[(2 * j) * 32 + L] = s1;
        reinterpret_cast<float4*>(ts)[(2 * j + 1) * 32 + L] = s2;
        ts[((16 + (j >> 2)) * 32 + L) * 4 + (j & 3)] = s3;
    }
}

__global__ void __launch_bounds__(256) k_prep_rgb3(const Rgb3PrepArgs P) {
    __shared__ int sC[3][RPP3][RPP3 + 1];
    __shared__ int hS[9][RPP3][RPT3 + 1];
    const int x0 = blockIdx.x * RPT3 - P.padx;
    const int y0 = blockIdx.y * RPT3 - PADY;
    const int tid = threadIdx.x;
    // a tile that lies entirely in the padding (29 % of them at 1080p D=256) only stores zeros
    const bool tile_in = x0 + RPT3 > 0 && x0 < P.w && y0 + RPT3 > 0 && y0 < P.h_held && y0 + RPT3 + P.y_global0 > 0 &&
                         y0 + P.y_global0 < P.frame_h;
    for (int i = tid; tile_in && i < RPP3 * RPP3; i += 256) {
        int py = i / RPP3, px = i - py * RPP3;
        int x = x0 + px - RAD, y = y0 + py - RAD;
        int yg = y + P.y_global0;
        bool in = (x >= 0 && x < P.w && y >= 0 && y < P.h_held && yg >= 0 && yg < P.frame_h);
        const uint8_t* q = P.rgb + ((size_t)(in ? y : 0) * P.w + (in ? x : 0)) * P.ch;
        sC[0][py][px] = in ? (int)q[0] : 0;
        sC[1][py][px] = in ? (int)q[1] : 0;
        sC[2][py][px] = in ? (int)q[2] : 0;
    }
    __syncthreads();
    for (int i = tid; tile_in && i < RPP3 * RPT3; i += 256) {
        int py = i / RPT3, tx = i - py * RPT3;
        int s[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
        for (int k = 0; k < WIN; k++) {
            int r = sC[0][py][tx + k], g = sC[1][py][tx + k], b = sC[2][py][tx + k];
            s[0] += r; s[1] += g; s[2] += b;
            s[3] += r * r; s[4] += r * g; s[5] += r * b;
            s[6] += g * g; s[7] += g * b; s[8] += b * b;
        }
#pragma unroll
        for (int q = 0; q < 9; q++) hS[q][py][tx] = s[q];
    }
    __syncthreads();
    const int n_rows_pad = P.h_held + 2 * PADY;
    for (int i = tid; i < RPT3 * RPT3; i += 256) {
        int ty = i / RPT3, tx = i - ty * RPT3;
        int x = x0 + tx, y = y0 + ty;
        if (x + P.padx >= P.pitch || y + PADY >= n_rows_pad) continue;
        int yg = y + P.y_global0;
        bool in = (x >= 0 && x < P.w && y >= 0 && y < P.h_held && yg >= 0 && yg < P.frame_h);
        const __half hz = __float2half(0.0f);
        if (!in) {
            store_tiled3(P, x, y + PADY, hz, hz, hz, make_float4(0.f, 0.f, 0.f, 0.f), make_float4(0.f, 0.f, 0.f, 0.f), 0.0f);
            continue;
        }
        int s[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
        for (int k = 0; k < WIN; k++)
#pragma unroll
            for (int q = 0; q < 9; q++) s[q] += hS[q][ty + k][tx];
        int ax = min(P.w - 1, x + RAD) - max(0, x - RAD) + 1;
        int ay = min(P.frame_h - 1, yg + RAD) - max(0, yg - RAD) + 1;
        float area = (float)(ax * ay);
        float mr = __fdiv_rn((float)s[0], area), mg = __fdiv_rn((float)s[1], area), mb = __fdiv_rn((float)s[2], area);
        double xx = (double)__fsub_rn(__fdiv_rn((float)s[3], area), __fmul_rn(mr, mr)) + P.eps;
        double xy = __fsub_rn(__fdiv_rn((float)s[4], area), __fmul_rn(mr, mg));
        double xz = __fsub_rn(__fdiv_rn((float)s[5], area), __fmul_rn(mr, mb));
        double yy = (double)__fsub_rn(__fdiv_rn((float)s[6], area), __fmul_rn(mg, mg)) + P.eps;
        double yz = __fsub_rn(__fdiv_rn((float)s[7], area), __fmul_rn(mg, mb));
        double zz = (double)__fsub_rn(__fdiv_rn((float)s[8], area), __fmul_rn(mb, mb)) + P.eps;
        double a00 = yy * zz - yz * yz, a01 = xz * yz - xy * zz, a02 = xy * yz - xz * yy;
        double a11 = xx * zz - xz * xz, a12 = xy * xz - xx * yz, a22 = xx * yy - xy * xy;
        double id = 1.0 / (xx * a00 + xy * a01 + xz * a02);
        float rxy = __fmul_rn(__frcp_rn((float)ax), __frcp_rn(P.S * (float)ay));
        const float4 s1 = make_float4(mr, mg, mb, __fmul_rn((float)(a00 * id), rxy));
        const float4 s2 = make_float4(__fmul_rn((float)(a01 * id), rxy), __fmul_rn((float)(a02 * id), rxy),
                                      __fmul_rn((float)(a11 * id), rxy), __fmul_rn((float)(a12 * id), rxy));
        const float s3 = __fmul_rn((float)(a22 * id), rxy);
        store_tiled3(P, x, y + PADY, __float2half((float)sC[0][ty + RAD][tx + RAD]), __float2half((float)sC[1][ty + RAD][tx + RAD]),
                     __float2half((float)sC[2][ty + RAD][tx + RAD]), s1, s2, s3);
    }
}

}  // namespace

int sbf_prep_gray_tiled(sb200_ctx* ctx, const sb200_params* p, const uint8_t* gray, const SbFusedGeom& g, int pitch, int padx,
                        float S, unsigned* IG, size_t shift_stride, unsigned* Tg, void* TI, float2* Tst, int n_strips);

size_t sbf_rgb3_workspace_bytes(const sb200_ctx* ctx, int w, int h_held, int rows_out, int dabs, int size_d) {
    const int padx = pad_x(dabs);
    const int pitch = (w + 2 * padx + 7) / 8 * 8;
    const size_t plane = (size_t)pitch * (h_held + 2 * PADY);
    const int pitchS = (w + 3) / 4 * 4;
    Plan plan = make_plan(w, rows_out, size_d, ctx->sm_count, 2);
    const size_t recs = (size_t)plan.n_strips * (h_held + 2 * PADY);
    size_t bytes = 0;
    bytes += 2 * sb_align(plane * 4 * 4);                                                   // IG x2, 4 shifted copies
    bytes += 2 * (sb_align(recs * 1024) + sb_align(recs * 512) + sb_align(recs * 2048));  // gray tiles Tg, TI, Tst
    bytes += 2 * (sb_align(recs * 1536) + sb_align(recs * 9216));                         // TC, TS
    bytes += sb_align((size_t)plan.n_chunks * 2 * rows_out * pitchS * 8);                 // BL
    return bytes + 8192;
}

int sbf_pair_disparity_rgb3(sb200_ctx* ctx, const sb200_params* p, const uint8_t* rgb_l, const uint8_t* rgb_r, int channels,
                            const uint8_t* gray_l, const uint8_t* gray_r, const SbFusedGeom& g, float* bestL, float* dispL,
                            float* bestR, float* dispR) {
    if (p->radius != RAD) return sb_fail(ctx, SB200_ERR_UNSUPPORTED, "fused RGB kernel is built for radius %d", RAD);
    int nI, nG, S;
    if (!find_lattice(p, &nI, &nG, &S))
        return sb_fail(ctx, SB200_ERR_UNSUPPORTED, "fused RGB kernel: no exact integer cost lattice for these parameters");
    if (g.w < 2 || g.h < 1 || g.rows_out < 1) return sb_fail(ctx, SB200_ERR_INVALID, "fused RGB: bad shape");
    const int size_d = p->dmax - p->dmin + 1;
    const int dmin[2] = {p->dmin, -p->dmax};
    const int dabs = max(abs(p->dmin), abs(p->dmax));
    const int padx = pad_x(dabs);
    const int pitch = (g.w + 2 * padx + 7) / 8 * 8;
    const int rows_pad = g.h + 2 * PADY;
    const size_t plane = (size_t)pitch * rows_pad;
    const int pitchS = (g.w + 3) / 4 * 4;
    Plan plan = make_plan(g.w, g.rows_out, size_d, ctx->sm_count, 2);
    const size_t recs = (size_t)plan.n_strips * rows_pad;
    const uint8_t* gray[2] = {gray_l, gray_r};
    const uint8_t* rgb[2] = {rgb_l, rgb_r};
    unsigned *IG[2], *Tg[2];
    __half *TI[2], *TC[2];
    float2* Tst[2];
    float* TS[2];
    for (int i = 0; i < 2; i++) {
        IG[i] = sb_ws_alloc<unsigned>(ctx, 4 * plane);
        Tg[i] = sb_ws_alloc<unsigned>(ctx, recs * 256);
        TI[i] = sb_ws_alloc<__half>(ctx, recs * 256);
        Tst[i] = sb_ws_alloc<float2>(ctx, recs * 256);
        TC[i] = sb_ws_alloc<__half>(ctx, recs * 768);
        TS[i] = sb_ws_alloc<float>(ctx, recs * 2304);
        if (!IG[i] || !Tg[i] || !TI[i] || !Tst[i] || !TC[i] || !TS[i])
            return sb_fail(ctx, SB200_ERR_NOMEM, "fused RGB: workspace arena too small (internal)");
    }
    const size_t planeS = (size_t)g.rows_out * pitchS;
    float2* BL = sb_ws_alloc<float2>(ctx, planeS * 2 * plan.n_chunks);
    if (!BL) return sb_fail(ctx, SB200_ERR_NOMEM, "fused RGB: workspace arena too small (internal)");

    if (ctx->timing && ctx->ev_valid) SB_CUDA(ctx, cudaEventRecord(ctx->ev[0], ctx->stream));
    for (int i = 0; i < 2; i++) {
        SB_TRY(sbf_prep_gray_tiled(ctx, p, gray[i], g, pitch, padx, (float)S, IG[i], plane, Tg[i], TI[i], Tst[i], plan.n_strips));
        Rgb3PrepArgs P;
        P.rgb = rgb[i];
        P.ch = channels;
        P.w = g.w;
        P.h_held = g.h;
        P.y_global0 = g.y_global0;
        P.frame_h = g.frame_h;
        P.TC = TC[i];
        P.TS = TS[i];
        P.n_strips = plan.n_strips;
        P.rows_pad = rows_pad;
        P.pitch = pitch;
        P.padx = padx;
        P.eps = p->eps;
        P.S = (float)S;
        dim3 grid(sb_div_up(pitch, RPT3), sb_div_up(rows_pad, RPT3));
        SB_LAUNCH(ctx, k_prep_rgb3, grid, 256, 0, P);
    }
    if (ctx->timing && ctx->ev_valid) SB_CUDA(ctx, cudaEventRecord(ctx->ev[1], ctx->stream));

    Rgb3Args A;
    const size_t origin = (size_t)PADY * pitch + padx;
    for (int i = 0; i < 2; i++) {
        A.IG[i] = IG[i];
        A.Tg[i] = reinterpret_cast<const uint4*>(Tg[i]);
        A.TC[i] = reinterpret_cast<const uint4*>(TC[i]);
        A.TS[i] = reinterpret_cast<const uint4*>(TS[i]);
        A.dmin[i] = dmin[i];
    }
    A.shift_stride = plane;
    A.ig_origin = origin;
    A.rows_pad = rows_pad;
    A.pitch = pitch;
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
    A.pitchS = pitchS;
    A.S = (float)S;
    A.zero = 0;
    __half2 wp = __floats2half2_rn((float)nI, (float)nG);
    __half2 tp = __floats2half2_rn(p->th_color, 2.0f * p->th_grad);
    A.wpack = *reinterpret_cast<unsigned*>(&wp);
    A.thpack = *reinterpret_cast<unsigned*>(&tp);
    const size_t smem = sizeof(Smem3);
    SB_CUDA(ctx, cudaFuncSetAttribute(k_fused_cvf_rgb3, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int nblocks = plan.n_strips * plan.n_bands * plan.n_chunks * 2;
    SB_LAUNCH(ctx, k_fused_cvf_rgb3, nblocks, R3_THREADS, smem, A);
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
