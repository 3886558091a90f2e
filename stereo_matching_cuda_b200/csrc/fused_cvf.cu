// fused_cvf.cu -- the hot path: cost volume + guided-filter aggregation + running argmin in
// one kernel, for both views of a pair, without ever materialising the D-deep volume.
//
// Replaces, per disparity slice, the reference's chain:
//   costVolumOnGPU2 (costVolume.cu:163-190) -> copyFromBigToLittleOnGPU -> pixelMultOnGPU ->
//   4x { rowSum + colSum (integral.cu:78-131) + computeBoxFilterOnGPU (guidedFilter.cu:297-318) }
//   -> compute_ak_and_bk (:345-354) -> compute_q (:363-369) -> dispSelectOnGPU (:403-411)
// and, per frame, chToFlOnGPU / the two guide-statistics box filters (guidedFilter.cu:58-123).
//
// Design (DESIGN.md has the long form):
//  * A warp owns a strip of 256 image columns (8 consecutive pixels per lane) for ONE
//    disparity and marches down the rows.  Vertical 19-row window sums are running sums in
//    registers; the values that leave the window come from 19-slot rings in Tensor Memory.
//  * Horizontal 19-column window sums never go through shared memory: each lane forms the
//    prefix sums of its 8 pixels and the window is assembled from 16 warp shuffles of
//    neighbouring lanes' prefixes.  216 of the 256 columns are valid after two cascaded
//    radius-9 filters (20-column halo each side, kept 4-pixel aligned for 128-bit loads).
//  * The cost is evaluated on an integer lattice: 20*p = 2*min(|dI|,7) + 9*min(|dG|,4) with
//    G = 2*gradient, so the first-stage box sums of p and I*p are sums of integers below 2^24
//    held in fp32 -- exact, order independent, identical on every tiling or GPU split.  The
//    half-precision costs enter the fp32 sums through the f16 x f16 + f32 instructions of
//    sm_100 (FHFMA / FHADD), without conversions.
//  * Warp specialisation: a block is 4 TRIOS of warps (p, p+4, p+8: same SM sub-partition, same Tensor-Memory lane
//    quarter), one trio per disparity, a 3-stage pipeline down the rows, plus 4 MERGE warps (p+12):
//      stage 0  cost, vertical sums of P and I*P, horizontal sums of P          -> (S_P, V_IP)
//      stage 1  horizontal sums of I*P, a, b, their vertical sums and rings     -> (V_a, V_b)
//      stage 2  horizontal sums of a and b, q = mean_a*I + mean_b               -> q ring (shared memory)
//      stage 3  merge of the block's 4 disparities into the running (best,label)
//    Rows travel between stages through Tensor-Memory columns (tcgen05.st / tcgen05.ld) guarded by named barriers
//    (0 -> 1) and mbarriers (1 -> 2, two slots).  The block is launched with 16 warps x 128 registers and every role
//    (one warpgroup) trades registers with setmaxnreg (152 / 176 / 128 / 56); four warps per scheduler fill the issue
//    slots one warp left empty (ncu: 27 % -> 44 % -> 58 % -> 74 % issue utilisation for 1, 2, 3, 4 warps per scheduler).
//  * The guide operands every trio of a block needs -- (I,G), I, (mean_I, c2) of the strip's rows -- are fetched ONCE
//    per block: TMA bulk copies (cp.async.bulk) from strip-tiled planes into a 16-slot shared-memory ring, one slot
//    (8 KB) per pipeline iteration, 5 iterations ahead (the stage-1 warp of trio p fills the iterations K = p mod 4);
//    the 12 stage warps read their 16-byte chunks conflict-free and release the slot through an mbarrier.  Only the
//    match operands (per disparity) are read through L1, as two aligned, fully coalesced 128-bit loads from the
//    de-interleaved copy of the (I,G) plane shifted by d&3.
//  * The stage-2 warps leave their filtered rows in a 4-deep shared-memory ring; each merge thread folds the 4
//    disparities for 2 columns into the running (best,label) with the reference's `best >= q` rule (last slice wins
//    ties; a 2-level tournament, which is the same function), one 16-byte load and store per row.
//  * Blocks tile (column strip x row band x disparity chunk x view); per-chunk (best,label) planes are merged in chunk
//    order by a small second kernel.
// The busiest unit is the L1TEX/shared-memory data pipe (72 %: shuffles + 128-bit shared loads), then the issue slots
// (74 %); DRAM explains ~20 % of the kernel's time.  The FUSED_* macros are the switches of the A/B experiments
// recorded in DESIGN.md 5, at their measured defaults.
#include "fused_dev.cuh"

namespace {


struct FusedArgs {
    const unsigned* IG[2];  // per IMAGE (0 left, 1 right): padded half2 (I, G=I[x-1]-I[x+1]), 4 copies: copy s at
    size_t shift_stride;    // IG[i] + s*shift_stride is the plane moved left by s elements, stored de-interleaved
    size_t ig_origin;       // (fused_dev.cuh, match_ptrs); linear element index of pixel (0,0) in a copy; the match
                            // operands at x+d are two aligned, fully coalesced 16 B loads from copy (d & 3)
    // Guide operands, STRIP-TILED (see k_prep): record (strip, padded row) of each plane holds the 256 columns of the
    // strip as [16-byte chunk c][lane][16 B], i.e. exactly the bytes lane L reads with its c-th 128-bit load, so one
    // bulk copy per plane and row pair fills the shared-memory operand ring without bank conflicts on the read side.
    const uint4* Tg[2];     // per IMAGE: (I,G) half2, 2 chunks per lane, 1 KB per row
    const uint4* TI[2];     // per IMAGE: I as half (exact: 0..255), 1 chunk, 512 B per row
    const uint4* Tst[2];    // per IMAGE: float2 (mean_I, c/(S*area)), 4 chunks, 2 KB per row
    int rows_pad;           // padded rows per strip (held rows + 2*PADY)
    int pitch;              // elements per padded row
    int w;
    int y_out0, rows_out;   // output rows, in held-row coordinates
    int y_global0, frame_h; // frame row of held row 0; rows in the whole frame
    int dmin[2];
    int size_d;
    int n_strips, n_bands, band_rows, n_chunks, chunk_d;
    int n_views;            // 2: both views (view v guides with image v); 1: view 0 only
    float2* BL;             // [chunk][view][rows_out][pitchS] running (best cost, label) of each disparity chunk
    int pitchS;
    float S;                // lattice scale: p = P / S
    unsigned wpack;         // half2 (nI, nG): P = nI*cI + nG*cG
    unsigned thpack;        // half2 (th_color, 2*th_grad)
    int zero;               // always 0; opaque to the compiler (see touch_ops)
};

// three pipeline stages (warps p, p+4, p+8) per disparity and the 4 merge warps (p+12)
constexpr int K3_THREADS = 4 * NWARP * 32;
// Register budgets of the four warp roles: the block is launched with 16 warps x 128 registers and every role (one
// warpgroup) trades with setmaxnreg at the top of its branch; 152 + 176 + 128 + 56 = 512 = the whole register file.
// (Overridable for A/B builds: tools/build_variant.sh ... -DFUSED_REGS1=168)
#ifndef FUSED_REGS0
#define FUSED_REGS0 152
#endif
#ifndef FUSED_REGS1
#define FUSED_REGS1 176
#endif
#ifndef FUSED_REGS2
#define FUSED_REGS2 128
#endif
#ifndef FUSED_REGS3
#define FUSED_REGS3 56
#endif
static_assert(FUSED_REGS0 + FUSED_REGS1 + FUSED_REGS2 + FUSED_REGS3 <= 512, "register file: 64 K registers per SM");
constexpr uint32_t TM_HAND2 = 416;  // TMEM columns [416,480): stage 1 -> stage 2 hand-off rows, 2 slots
constexpr int NQ = 4;               // depth of the ring of filtered rows between stage 2 and the merge warps
constexpr int NS = 16;              // slots of the operand ring (one per pipeline iteration)
constexpr int LOAD_AHEAD = 5;       // the loading warps fill the slot of iteration K + LOAD_AHEAD while they work on K
struct Slot {                 // guide operands of one iteration, filled by 4 bulk copies (8 KB)
    uint4 g[ROWS][2][32];     // (I,G) at rows yi
    uint4 io[ROWS][32];       // I at rows yi-19 (leave the first-stage window)
    uint4 iq[ROWS][32];       // I at rows yq = yi-18 (the output rows of this iteration)
    uint4 st[ROWS][4][32];    // (mean_I, c2) at rows yi-9
};
constexpr uint32_t SLOT_BYTES = sizeof(Slot);
struct SmemLayout {
    Slot slot[NS];
    uint64_t sfull[NS], sempty[NS];       // mbarriers of the operand ring: bulk copies -> 12 warps and back
    float4 qbuf[NQ][NWARP][ROWS][2][QV];  // filtered rows of each consumer warp
    uint64_t qfull[NQ], qempty[NQ];       // mbarriers of the q ring (the 4 stage-2 warps write, the 4 merge warps read)
    uint64_t full2[NWARP][2], empty2[NWARP][2];  // mbarriers of the 2-slot stage 1 -> stage 2 hand-off of each pair
    float ry_lut[2][WIN + 1];             // [0][n] = 1/(S*n), [1][n] = 1/n for a clipped window of n rows; [.][0] = 0
    uint32_t tmem_base;
};

// operands of one row step, fetched one step ahead; the producer and the consumer warp of a
// pair each fetch only what their stage needs
struct ProdOps {
    uint4 m0, m1;     // match (I,G) half2 x8 at row yi, columns x+d (from the copy shifted by d & 3)
};
struct ProdPtrs {
    const unsigned *m0, *m1;  // pixels 0..3 and 4..7 of this lane (see match_ptrs)
};

__device__ __forceinline__ void load_prod(ProdOps& o, const ProdPtrs& p, int dep) {
    o.m0 = __ldg(reinterpret_cast<const uint4*>(p.m0 + dep));
    o.m1 = __ldg(reinterpret_cast<const uint4*>(p.m1 + dep));
}
// One word of every load of `o`, OR-ed together.  The next step's loads are made to depend on
// it (masked to zero by a kernel argument the compiler cannot see through), so the scoreboard
// wait for THIS step's operands is taken before the new loads are issued.  Without it the
// first use of an operand waits on a scoreboard slot that the just-issued prefetch loads
// share, i.e. on a full L2 round trip every step (ncu: 50 % of all stall samples, round 1).
__device__ __forceinline__ int touch(const ProdOps& o) {
    return (int)(o.m0.x | o.m1.x);
}
__device__ __forceinline__ void copy_ops(ProdOps (&a)[ROWS], const ProdOps (&b)[ROWS]) {
#pragma unroll
    for (int r = 0; r < ROWS; r++) a[r] = b[r];
}
__global__ void __launch_bounds__(K3_THREADS, 1) k_fused_cvf(const FusedArgs A) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SmemLayout& sm = *reinterpret_cast<SmemLayout*>(smem_raw);

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int pair = warp & (NWARP - 1);  // warps p, p+4, p+8 share SM sub-partition p and TMEM lane quarter p
    const int stage = warp / NWARP;       // 0: first stage, 1: coefficients + vertical sums, 2: output + merge
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
    // 36 warm-up rows, then one output row per input row; iterations take ROWS rows (the last
    // one may run past the band: those rows read zero padding and are not stored)
    const int niter = ((yb1 - yb0) + 4 * RAD + ROWS - 1) / ROWS;
    constexpr int WARM_IT = 4 * RAD / ROWS;  // 36 warm-up rows fill both windows
    static_assert((4 * RAD) % ROWS == 0, "ROWS must divide the warm-up length");
    const int n_emit = niter - WARM_IT;      // emissions (iterations that output rows) per group
    const int BAR_FULL = 2 + pair, BAR_EMPTY = 2 + NWARP + pair;
    // A.zero in a register the compiler cannot re-read from the constant bank: a uniform constant
    // load inside the row loop shares its scoreboard with the loads issued just before it and
    // then waits for their full L2 round trip (ncu: 9 % of all samples on one LOP3)
    int zero = A.zero;
    asm volatile("" : "+r"(zero));

    // Tensor Memory: all 512 columns of the SM, one block per SM (register-limited)
    if (warp == 0) tm_alloc(&sm.tmem_base);
    const uint32_t mb_qfull = smem_addr(&sm.qfull[0]), mb_qempty = smem_addr(&sm.qempty[0]);
    const uint32_t mb_full2 = smem_addr(&sm.full2[pair][0]), mb_empty2 = smem_addr(&sm.empty2[pair][0]);
    const uint32_t mb_sfull = smem_addr(&sm.sfull[0]), mb_sempty = smem_addr(&sm.sempty[0]);
    const uint32_t slot0 = smem_addr(&sm.slot[0]) + 16 * lane;  // this lane's 16 bytes of a slot's chunk 0
    static_assert((NS & (NS - 1)) == 0, "slot index and phase are taken from the bits of the iteration counter");
    if (threadIdx.x == 0) {
#pragma unroll
        for (int b = 0; b < NS; b++) {
            mbar_init(mb_sfull + 8 * b, 1);
            mbar_init(mb_sempty + 8 * b, 3 * NWARP);  // the 12 stage warps read every slot
        }
    }
    if (threadIdx.x == 32) {
#pragma unroll
        for (int b = 0; b < NQ; b++) {
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
    const uint32_t tAB = tbase + TM_RING_AB, tP = tbase + TM_RING_P, tH = tbase + TM_HAND, tH2 = tbase + TM_HAND2;

    // operand ring, consumer side: every warp waits for the slot of its iteration K, reads its 16-byte chunks and
    // releases the slot.  `t` is a word of the loaded data: the release depends on it, so it cannot be issued before
    // the reads have completed.
    auto slot_wait = [&](int K) -> uint32_t {
        mbar_wait(mb_sfull + 8 * (K & (NS - 1)), (unsigned)(K / NS) & 1u);
        return slot0 + (uint32_t)(K & (NS - 1)) * SLOT_BYTES;
    };
    auto slot_release = [&](int K, unsigned t) {  // (a warp-wide load returns for all lanes at once: lane 0's word is enough)
        mbar_arrive_lane0(mb_sempty + 8 * (K & (NS - 1)) + (t & (unsigned)zero), lane);
    };
    constexpr uint32_t OFF_G = offsetof(Slot, g), OFF_IO = offsetof(Slot, io), OFF_IQ = offsetof(Slot, iq),
                       OFF_ST = offsetof(Slot, st);

    if (stage == 0) {
        // ====== STAGE 0: lattice cost, vertical window sums of P and I*P, their horizontal sums ======
        reg_inc<FUSED_REGS0>();
        const unsigned* __restrict__ IGm = A.IG[1 - view];
        __half2 wm[KPX];  // lattice weights (nI, nG), 0 outside the image (masks the cost)
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
            float VP[KPX], VIP[KPX];
#pragma unroll
            for (int j = 0; j < KPX; j++) VP[j] = VIP[j] = 0.0f;
            {
                const uint32_t z[4] = {0u, 0u, 0u, 0u};
                for (int s = 0; s < WIN; s++) tm_st4(tP + 4 * s, z);
                tm_wait_st();
            }
            __syncthreads();  // group start
            if (active) {
                ProdPtrs rp;
                const long long r0 = (long long)y_first * pitch + xl;
                match_ptrs(IGm + (size_t)(d & 3) * A.shift_stride, (long long)A.ig_origin + r0 + (d - (d & 3)),
                           A.shift_stride / 2, rp.m0, rp.m1);
                int slot = 0;
                ProdOps opsA[ROWS], opsB[ROWS];
#pragma unroll
                for (int r = 0; r < ROWS; r++) {
                    load_prod(opsA[r], rp, 0);
                    rp.m0 += pitch / 2;
                        rp.m1 += pitch / 2;
                }
                auto iter = [&](const ProdOps (&o)[ROWS], ProdOps (&nxt)[ROWS], int it) {
                    int dep = 0;
#pragma unroll
                    for (int r = 0; r < ROWS; r++) dep |= touch(o[r]);
                    dep &= zero;
#pragma unroll
                    for (int r = 0; r < ROWS; r++) {
                        load_prod(nxt[r], rp, dep);
                        rp.m0 += pitch / 2;
                        rp.m1 += pitch / 2;
                    }
                    // guide operands of this iteration from the shared-memory ring
                    uint4 gq[ROWS][2], ioq[ROWS];
                    {
                        const int K = g * niter + it;
                        const uint32_t sa = slot_wait(K);
                        unsigned t = 0;
#pragma unroll
                        for (int r = 0; r < ROWS; r++) {
                            gq[r][0] = lds128(sa + OFF_G + (r * 2 + 0) * 512);
                            gq[r][1] = lds128(sa + OFF_G + (r * 2 + 1) * 512);
                            ioq[r] = lds128(sa + OFF_IO + r * 512);
                            t |= gq[r][0].x | gq[r][1].x | ioq[r].x;
                        }
                        slot_release(K, t);
                    }
                    // the (negated) lattice costs that leave the window (rows yi-19), from the TMEM ring
                    uint32_t pold[ROWS][4];
                    int slots[ROWS];
#pragma unroll
                    for (int r = 0; r < ROWS; r++) {
                        slots[r] = slot;
                        tm_ld4(tP + 4 * slot, pold[r]);
                        slot = (slot + 1 == WIN) ? 0 : slot + 1;
                    }
                    float SP[ROWS][KPX], SIP[ROWS][KPX];
                    uint32_t pneg[ROWS][4];
                    __half ph[ROWS][KPX];
#pragma unroll
                    for (int r = 0; r < ROWS; r++) {
                        const unsigned gg[KPX] = {gq[r][0].x, gq[r][0].y, gq[r][0].z, gq[r][0].w,
                                                  gq[r][1].x, gq[r][1].y, gq[r][1].z, gq[r][1].w};
                        const unsigned mm[KPX] = {o[r].m0.x, o[r].m0.y, o[r].m0.z, o[r].m0.w,
                                                  o[r].m1.x, o[r].m1.y, o[r].m1.z, o[r].m1.w};
#pragma unroll
                        for (int j = 0; j < KPX; j++) {
                            __half2 gv = u2h2(gg[j]);
                            __half2 diff = __hsub2(gv, u2h2(mm[j]));
                            __half2 c = __hmin2(__habs2(diff), th);  // (min(|dI|,Tc), min(|dG|,2Tg))
                            __half2 pr = __hmul2(c, wm[j]);
                            ph[r][j] = __hadd(__low2half(pr), __high2half(pr));
                        }
                    }
                    tm_wait_ld();  // pold is in registers; the slots may be overwritten
#pragma unroll
                    for (int r = 0; r < ROWS; r++) {
                        const unsigned gg[KPX] = {gq[r][0].x, gq[r][0].y, gq[r][0].z, gq[r][0].w,
                                                  gq[r][1].x, gq[r][1].y, gq[r][1].z, gq[r][1].w};
                        const unsigned io[4] = {ioq[r].x, ioq[r].y, ioq[r].z, ioq[r].w};
                        // the ring holds -P: the window update is then P + (-P_old) in packed half
                        // (exact, |.| <= 2^11) and the f16 x f16 + f32 forms FHADD / FHFMA of sm_100
                        // fold every half -> float conversion into the accumulation
#pragma unroll
                        for (int k = 0; k < KPX / 2; k++) {
                            const __half2 pn2 = __halves2half2(ph[r][2 * k], ph[r][2 * k + 1]);
                            const __half2 po2 = u2h2(pold[r][k]);
                            pneg[r][k] = h22u(__hneg2(pn2));
                            const __half2 dp = __hadd2(pn2, po2);
                            const __half2 io2 = u2h2(io[k]);
                            VP[2 * k] = fhadd(__low2half(dp), VP[2 * k]);
                            VP[2 * k + 1] = fhadd(__high2half(dp), VP[2 * k + 1]);
                            VIP[2 * k] = fhfma(__low2half(u2h2(gg[2 * k])), ph[r][2 * k], VIP[2 * k]);
                            VIP[2 * k + 1] = fhfma(__low2half(u2h2(gg[2 * k + 1])), ph[r][2 * k + 1], VIP[2 * k + 1]);
                            VIP[2 * k] = fhfma(__low2half(io2), __low2half(po2), VIP[2 * k]);
                            VIP[2 * k + 1] = fhfma(__high2half(io2), __high2half(po2), VIP[2 * k + 1]);
                        }
                        tm_st4(tP + 4 * slots[r], pneg[r]);
                        hsum19(VP, SP[r]);
                        // (the horizontal sums of I*P are taken by stage 1: this stage was the long pole with them)
#pragma unroll
                        for (int j = 0; j < KPX; j++) SIP[r][j] = VIP[j];
                    }
                    // stage 1 has copied the previous rows out of the hand-off columns
                    if (it > 0) {
                        named_bar_sync(BAR_EMPTY, 64);
                        tm_fence_after();
                    }
#pragma unroll
                    for (int r = 0; r < ROWS; r++) tm_st16(tH + 16 * r, SP[r], SIP[r]);
                    tm_wait_st();
                    tm_fence_before();
                    named_bar_arrive(BAR_FULL, 64);
                };
#pragma unroll 1
                for (int it = 0; it < niter; it++) {
                    iter(opsA, opsB, it);
                    copy_ops(opsA, opsB);
                }
            } else {
                for (int it = 0; it < niter; it++) {  // a pair without a disparity still releases its share of the ring
                    slot_wait(g * niter + it);
                    slot_release(g * niter + it, 0u);
                }
            }
            __syncthreads();  // group end
        }
    } else if (stage == 1) {
        // ====== STAGE 1: a, b; their vertical and horizontal window sums ======
        reg_inc<FUSED_REGS1>();
        float rx[KPX];  // 1 / clipped window width, 0 outside the image
#pragma unroll
        for (int j = 0; j < KPX; j++) {
            int x = xl + j;
            int ax = min(A.w - 1, x + RAD) - max(0, x - RAD) + 1;
            rx[j] = (x >= 0 && x < A.w) ? __frcp_rn((float)ax) : 0.0f;
        }
        // Operand ring, producer side: the stage-1 warps fill the slot of iteration K + LOAD_AHEAD, for the whole block,
        // while they work on K: 4 bulk copies of the strip-tiled guide planes.  The warp of trio p fills the iterations
        // K = p mod 4, so the ~100 instructions of a fill are shared by the four trios instead of slowing trio 0 (the
        // trios run in step through the merge ring: the slowest one sets the pace).
        const int Ktotal = ngroups * niter;
        constexpr int fill_step = NWARP;
        int fillK = pair, fill_it = fillK;  // this warp's next iteration to fill, and its row iteration inside its group
        auto fill_due = [&](int K) {  // called at the start of iteration K: fill what is due up to K + LOAD_AHEAD
            while (fillK <= K + LOAD_AHEAD && fillK < Ktotal) {
                const int sl = fillK & (NS - 1);
                if (fillK >= NS) mbar_wait(mb_sempty + 8 * sl, (unsigned)(fillK / NS - 1) & 1u);  // all 12 warps released it
                if (lane == 0) {
                    const uint32_t full = mb_sfull + 8 * sl;
                    const uint32_t dst = smem_addr(&sm.slot[0]) + (uint32_t)sl * SLOT_BYTES;
                    const long long rec = (long long)strip * A.rows_pad + PADY + y_first + fill_it * ROWS;  // record of row yi0
                    mbar_expect_tx(full, SLOT_BYTES);
                    bulk_g2s(dst + OFF_G, A.Tg[view] + rec * 64, ROWS * 1024, full);
                    bulk_g2s(dst + OFF_IO, A.TI[view] + (rec - WIN) * 32, ROWS * 512, full);
                    bulk_g2s(dst + OFF_IQ, A.TI[view] + (rec - 2 * RAD) * 32, ROWS * 512, full);
                    bulk_g2s(dst + OFF_ST, A.Tst[view] + (rec - RAD) * 128, ROWS * 2048, full);
                }
                __syncwarp();
                fillK += fill_step;
                fill_it += fill_step;
                if (fill_it >= niter) fill_it -= niter;
            }
        };
        fill_due(-1);  // the slots of iterations 0 .. LOAD_AHEAD-1
        for (int g = 0; g < ngroups; g++) {
            const int dk = g * NWARP + pair;
            const bool active = dk < dcnt;
            float Va[KPX], Vb[KPX];
#pragma unroll
            for (int j = 0; j < KPX; j++) Va[j] = Vb[j] = 0.0f;
            for (int s = 0; s < WIN; s++) tm_st16(tAB + 16 * s, Va, Vb);  // zeros
            tm_wait_st();
            __syncthreads();  // group start

            int slot = 0;
            auto iter = [&](auto emit_tag, int it) {
                constexpr bool EMIT = decltype(emit_tag)::value;
                fill_due(g * niter + it);
                // (mean_I, c2) of rows ya from the shared-memory ring
                uint4 sq[ROWS][4];
                {
                    const int K = g * niter + it;
                    const uint32_t sa = slot_wait(K);
                    unsigned t = 0;
#pragma unroll
                    for (int r = 0; r < ROWS; r++)
#pragma unroll
                        for (int c = 0; c < 4; c++) {
                            sq[r][c] = lds128(sa + OFF_ST + (r * 4 + c) * 512);
                            t |= sq[r][c].x;
                        }
                    slot_release(K, t);
                }
                const int yi0 = y_first + it * ROWS;
                float ry1[ROWS];
#pragma unroll
                for (int r = 0; r < ROWS; r++) ry1[r] = inv_rows(sm.ry_lut[0], yi0 + r - RAD, A.y_global0, A.frame_h);
                // the (a,b) rows that leave the second-stage window, from this pair's TMEM ring
                float ao[ROWS][KPX], bo[ROWS][KPX];
                int slots[ROWS];
#pragma unroll
                for (int r = 0; r < ROWS; r++) {
                    slots[r] = slot;
                    tm_ld16(tAB + 16 * slot, ao[r], bo[r]);
                    slot = (slot + 1 == WIN) ? 0 : slot + 1;
                }
                named_bar_sync(BAR_FULL, 64);  // stage 0 has published rows yi0 .. yi0+ROWS-1
                tm_fence_after();
                float SP[ROWS][KPX], SIP[ROWS][KPX];
#pragma unroll
                for (int r = 0; r < ROWS; r++) tm_ld16(tH + 16 * r, SP[r], SIP[r]);
                tm_wait_ld();
                if (it + 1 < niter) {  // hand-off columns are free again
                    tm_fence_before();
                    named_bar_arrive(BAR_EMPTY, 64);
                }
                float SA[ROWS][KPX], SB[ROWS][KPX];
#pragma unroll
                for (int r = 0; r < ROWS; r++) {
                    {   // stage 0 hands the vertical sums of I*P over: their horizontal sums are taken here
                        float hs[KPX];
                        hsum19(SIP[r], hs);
#pragma unroll
                        for (int j = 0; j < KPX; j++) SIP[r][j] = hs[j];
                    }
                    // ---- a, b at row ya = yi - 9
                    float a[KPX], b[KPX];
                    const unsigned stt[16] = {sq[r][0].x, sq[r][0].y, sq[r][0].z, sq[r][0].w, sq[r][1].x, sq[r][1].y, sq[r][1].z, sq[r][1].w,
                                              sq[r][2].x, sq[r][2].y, sq[r][2].z, sq[r][2].w, sq[r][3].x, sq[r][3].y, sq[r][3].z, sq[r][3].w};
#pragma unroll
                    for (int j = 0; j < KPX; j++) {
                        const float mI = __uint_as_float(stt[2 * j]), c2 = __uint_as_float(stt[2 * j + 1]);
                        float cov = fmaf(-mI, SP[r][j], SIP[r][j]);
                        a[j] = cov * c2;
                        float mp = SP[r][j] * (rx[j] * ry1[r]);
                        b[j] = fmaf(-mI, a[j], mp);
                    }
                    // ---- second stage: (a,b) of row ya enter, row ya-19 leaves
                    tm_st16(tAB + 16 * slots[r], a, b);
#pragma unroll
                    for (int j = 0; j < KPX; j++) {
                        Va[j] += a[j] - ao[r][j];
                        Vb[j] += b[j] - bo[r][j];
                    }
                    if (EMIT) {  // (the horizontal sums of a and b are taken by stage 2)
#pragma unroll
                        for (int j = 0; j < KPX; j++) {
                            SA[r][j] = Va[j];
                            SB[r][j] = Vb[j];
                        }
                    }
                }
                if (EMIT) {
                    const int E = g * n_emit + (it - WARM_IT);
                    if (E > 1) {  // stage 2 has copied the rows of emission E-2 out of this hand-off slot
                        mbar_wait(mb_empty2 + 8 * (E & 1), (unsigned)(E / 2 - 1) & 1u);
                        tm_fence_after();
                    }
#pragma unroll
                    for (int r = 0; r < ROWS; r++) tm_st16(tH2 + 32 * (E & 1) + 16 * r, SA[r], SB[r]);
                    tm_wait_st();  // (also: this iteration's ring stores are complete before the next loads)
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
                for (; it < WARM_IT; it++) iter(std::false_type{}, it);
#pragma unroll 1
                for (; it < niter; it++) iter(std::true_type{}, it);
            } else {
                for (int it = 0; it < niter; it++) {  // a trio without a disparity still fills and releases its share
                    fill_due(g * niter + it);
                    slot_wait(g * niter + it);
                    slot_release(g * niter + it, 0u);
                }
            }
            __syncthreads();  // group end
        }
    } else if (stage == 2) {
        // ====== STAGE 2: horizontal sums of a and b, q = mean_a * I + mean_b ======
        reg_inc<FUSED_REGS2>();
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
                for (int b = 0; b < NQ; b++)
#pragma unroll
                    for (int r = 0; r < ROWS; r++)
#pragma unroll
                        for (int v = 0; v < 2; v++) sm.qbuf[b][pair][r][v][lane] = make_float4(inf, inf, inf, inf);
            }
            __syncthreads();  // group start
            // the warm-up iterations only release this warp's share of the operand ring
            for (int it = 0; it < WARM_IT; it++) {
                slot_wait(g * niter + it);
                slot_release(g * niter + it, 0u);
            }
            if (active) {
#pragma unroll 1
                for (int e = 0; e < n_emit; e++) {
                    // I at the output rows yq = yb0 + e*ROWS + r, from the shared-memory ring
                    uint4 iqA[ROWS];
                    {
                        const int K = g * niter + WARM_IT + e;
                        const uint32_t sa = slot_wait(K);
                        unsigned t = 0;
#pragma unroll
                        for (int r = 0; r < ROWS; r++) {
                            iqA[r] = lds128(sa + OFF_IQ + r * 512);
                            t |= iqA[r].x;
                        }
                        slot_release(K, t);
                    }
                    const int E = g * n_emit + e;
                    mbar_wait(mb_full2 + 8 * (E & 1), (unsigned)(E / 2) & 1u);  // stage 1 has published this emission
                    tm_fence_after();
                    float SA[ROWS][KPX], SB[ROWS][KPX];
#pragma unroll
                    for (int r = 0; r < ROWS; r++) tm_ld16(tH2 + 32 * (E & 1) + 16 * r, SA[r], SB[r]);
                    tm_wait_ld();
                    tm_fence_before();
                    __syncwarp();
                    mbar_arrive_lane0(mb_empty2 + 8 * (E & 1), lane);
                    const int qb = E & (NQ - 1);
                    if (E >= NQ) mbar_wait(mb_qempty + 8 * qb, (unsigned)(E / NQ - 1) & 1u);  // merged NQ emissions ago
#pragma unroll
                    for (int r = 0; r < ROWS; r++) {
                        {
                            float ha[KPX];
                            hsum19(SA[r], ha);
#pragma unroll
                            for (int j = 0; j < KPX; j++) SA[r][j] = ha[j];
                        }
                        {
                            float hb[KPX];
                            hsum19(SB[r], hb);
#pragma unroll
                            for (int j = 0; j < KPX; j++) SB[r][j] = hb[j];
                        }
                        const float ry2 = inv_rows(sm.ry_lut[1], yb0 + e * ROWS + r, A.y_global0, A.frame_h);
                        const float2 iq01 = __half22float2(u2h2(iqA[r].x)), iq23 = __half22float2(u2h2(iqA[r].y));
                        const float2 iq45 = __half22float2(u2h2(iqA[r].z)), iq67 = __half22float2(u2h2(iqA[r].w));
                        const float iq[KPX] = {iq01.x, iq01.y, iq23.x, iq23.y, iq45.x, iq45.y, iq67.x, iq67.y};
                        float q[KPX];
#pragma unroll
                        for (int j = 0; j < KPX; j++) q[j] = fmaf(SA[r][j], iq[j], SB[r][j]) * (rx[j] * ry2);
                        sm.qbuf[qb][pair][r][0][lane] = make_float4(q[0], q[1], q[2], q[3]);
                        sm.qbuf[qb][pair][r][1][lane] = make_float4(q[4], q[5], q[6], q[7]);
                    }
                    __syncwarp();
                    mbar_arrive_lane0(mb_qfull + 8 * qb, lane);
                }
            } else {
                // no disparity for this pair in the (last, partial) group: its slots of the q ring hold +inf
                for (int e = 0; e < n_emit; e++) {
                    slot_wait(g * niter + WARM_IT + e);
                    slot_release(g * niter + WARM_IT + e, 0u);
                    const int E = g * n_emit + e;
                    const int qb = E & (NQ - 1);
                    if (E >= NQ) mbar_wait(mb_qempty + 8 * qb, (unsigned)(E / NQ - 1) & 1u);
                    if (lane == 0) mbar_arrive(mb_qfull + 8 * qb);
                    __syncwarp();
                }
            }
            __syncthreads();  // group end
        }
    } else {
        // ====== STAGE 3: merge of the block's 4 disparities into the running (best,label) ======
        // Thread t of these 4 warps folds strip-local columns 2t, 2t+1 of the rows the stage-2 warps publish in the q
        // ring; the (best,label) of the previous groups is fetched two emissions ahead.
        reg_dec<FUSED_REGS3>();
        const int mc = 2 * (threadIdx.x - 3 * NWARP * 32);
        const int mx = xs + mc;
        const bool mvalid = (mc >= HALO) && (mc < HALO + VALID_W) && (mx < A.w);
        const int qoff = (((mc & 7) >> 2) * QV + (mc >> 3)) * 4 + (mc & 3);
        const size_t planeS = (size_t)A.rows_out * A.pitchS;
        float2* __restrict__ BL = A.BL + (size_t)(chunk * 2 + view) * planeS;
        const size_t bl_row = (size_t)A.pitchS / 2;  // 16-byte units per row of the plane
        float4* const bl0 = reinterpret_cast<float4*>(BL + (size_t)(yb0 - A.y_out0) * A.pitchS + mx);
        const int band_rows = yb1 - yb0;
        for (int g = 0; g < ngroups; g++) {
            const int dbase = dlo + g * NWARP;
            const float lab[NWARP] = {(float)dbase, (float)(dbase + 1), (float)(dbase + 2), (float)(dbase + 3)};
            const bool ld_ok = (g > 0) && mvalid;
            auto prefetch_at = [&](int e, float4 (&pb)[ROWS]) {  // the rows of emission e
#pragma unroll
                for (int r = 0; r < ROWS; r++) {
                    pb[r] = make_float4(BEST_INIT_BITS_F, 0.0f, BEST_INIT_BITS_F, 0.0f);
                    if (ld_ok && e * ROWS + r < band_rows) pb[r] = ld_early_f4(bl0 + (size_t)(e * ROWS + r) * bl_row);
                }
            };
            __syncthreads();  // group start
            float4 pbA[ROWS], pbB[ROWS];
            prefetch_at(0, pbA);
            prefetch_at(1, pbB);
            float4* blp = bl0;
#pragma unroll 1
            for (int e = 0; e < n_emit; e++) {
                float4 pb[ROWS];
#pragma unroll
                for (int r = 0; r < ROWS; r++) {
                    pb[r] = pbA[r];
                    pbA[r] = pbB[r];
                }
                prefetch_at(e + 2, pbB);
                const int E = g * n_emit + e;
                const int qb = E & (NQ - 1);
                mbar_wait(mb_qfull + 8 * qb, (unsigned)(E / NQ) & 1u);
#pragma unroll
                for (int r = 0; r < ROWS; r++) {
                    const float4 nb = merge4(reinterpret_cast<const float*>(&sm.qbuf[qb][0][r][0][0]) + qoff, ROWS * 2 * QV * 4, lab, pb[r]);
                    if (mvalid && e * ROWS + r < band_rows) blp[r * bl_row] = nb;
                }
                __syncwarp();
                mbar_arrive_lane0(mb_qempty + 8 * qb, lane);
                blp += ROWS * bl_row;
            }
            __syncthreads();  // group end
        }
    }
    // every tcgen05 operation of this block has completed (waits above); release Tensor Memory
    tm_fence_before();
    __syncthreads();
    if (warp == 0) {
        tm_fence_after();
        tm_dealloc(sm.tmem_base);
    }
}

// ---------------------------------------------------------------------------------------
// Per-frame preparation (replaces chToFlOnGPU, x_derivativeOnGPU and the guide-statistics part
// of compute_guided_filter, guidedFilter.cu:58-123): for one image writes the padded planes
//   IG  half2 (I, G)   G = I[x-1]-I[x+1] = 2*gradient (costVolume.cu:364-378 border rule);
//                      padding = (1024,1024) so an out-of-range match saturates both terms
//   If  I as HALF (exact for 0..255), element x stored at index x+4 of the plane; zero padding
//   st  float2 (mean_I, c/(S*area))  zero padding; mean/variance from EXACT integer window
//       sums (<= 2^25, int32), c = (float)(1/((double)var+eps)) as guidedFilter.cu:350
// One block computes a 32x32 tile; the 50x50 input patch is staged in shared memory.
constexpr int PT = 32;
constexpr int PP = PT + 2 * RAD;  // 50

struct PrepArgs {
    const uint8_t* gray;  // held rows, pitch w
    int w, h_held, y_global0, frame_h;
    unsigned* IG;
    size_t shift_stride;  // != 0: also write the copies moved left by 1..3 elements at IG + s*shift_stride
    float* If;            // optional linear planes (not used by the gray fused kernel)
    float2* st;
    unsigned* Tg;         // optional strip-tiled planes (FusedArgs): (I,G) words, I halfs, (mean_I, c2) pairs
    __half* TI;
    float2* Tst;
    int n_strips, rows_pad;
    uint8_t* mean_u8;  // optional, held rows pitch w
    int pitch, padx;   // padded planes: rows [-PADY, h_held+PADY), cols [-padx, pitch-padx)
    double eps;
    float S;
};

__device__ __forceinline__ void store_ig(const PrepArgs& P, size_t o, unsigned v) {
    if (P.shift_stride) {  // 4 copies moved left by 0..3 elements, each de-interleaved (fused_dev.cuh, match_ptrs)
#pragma unroll
        for (int s = 0; s < 4; s++)
            if (o >= (size_t)s) P.IG[s * P.shift_stride + deint_index(o - s, P.shift_stride / 2)] = v;
    } else {
        P.IG[o] = v;
    }
}

// One pixel of the padded frame into the strip-tiled planes.  Strip s covers columns [216 s - 20, 216 s + 236); the
// 40 columns two strips share are stored in both.  Inside a record the 256 columns are laid out as
// [16-byte chunk c][lane][16 B] with column = 8*lane + j: the c-th 128-bit load of lane L is uint4 c*32 + L.
// One pixel of the padded frame into the strip-tiled planes (fused_dev.cuh: strips_of_column, tg/ti/tst_index).
__device__ __forceinline__ void store_tiled(const PrepArgs& P, int x, int yrow, unsigned ig, __half ih, float2 st) {
    const int xa = x + HALO;
    if (xa < 0) return;
    const int s1 = xa / VALID_W;
    const int c1 = xa - s1 * VALID_W;
#pragma unroll
    for (int k = 0; k < 2; k++) {  // (strips_of_column, unrolled)
        const int s = s1 - k, cl = c1 + k * VALID_W;
        if (s < 0 || s >= P.n_strips || cl >= SW) continue;
        const size_t rec = (size_t)s * P.rows_pad + yrow;
        P.Tg[tg_index(rec, cl)] = ig;
        P.TI[ti_index(rec, cl)] = ih;
        P.Tst[tst_index(rec, cl)] = st;
    }
}

__global__ void __launch_bounds__(256) k_prep(const PrepArgs P) {
    __shared__ int sI[PP][PP + 1];
    __shared__ int hI[PP][PT + 1];
    __shared__ int hII[PP][PT + 1];
    // tile origin in padded coordinates -> image coordinates
    const int x0 = blockIdx.x * PT - P.padx;
    const int y0 = blockIdx.y * PT - PADY;
    const int tid = threadIdx.x;
    // a tile that lies entirely in the padding only stores the padding values
    const bool tile_in = x0 + PT > 0 && x0 < P.w && y0 + PT > 0 && y0 < P.h_held && y0 + PT + P.y_global0 > 0 &&
                         y0 + P.y_global0 < P.frame_h;
    for (int i = tid; tile_in && i < PP * PP; i += 256) {
        int py = i / PP, px = i - py * PP;
        int x = x0 + px - RAD, y = y0 + py - RAD;
        int yg = y + P.y_global0;
        bool in = (x >= 0 && x < P.w && y >= 0 && y < P.h_held && yg >= 0 && yg < P.frame_h);
        sI[py][px] = in ? (int)P.gray[(size_t)y * P.w + x] : 0;
    }
    __syncthreads();
    for (int i = tid; tile_in && i < PP * PT; i += 256) {
        int py = i / PT, tx = i - py * PT;
        int s1 = 0, s2 = 0;
#pragma unroll
        for (int k = 0; k < WIN; k++) {
            int v = sI[py][tx + k];
            s1 += v;
            s2 += v * v;
        }
        hI[py][tx] = s1;
        hII[py][tx] = s2;
    }
    __syncthreads();
    const int n_rows_pad = P.h_held + 2 * PADY;
    for (int i = tid; i < PT * PT; i += 256) {
        int ty = i / PT, tx = i - ty * PT;
        int x = x0 + tx, y = y0 + ty;
        if (x + P.padx >= P.pitch || y + PADY >= n_rows_pad) continue;
        const size_t o = (size_t)(y + PADY) * P.pitch + (x + P.padx);
        int yg = y + P.y_global0;
        bool in = (x >= 0 && x < P.w && y >= 0 && y < P.h_held && yg >= 0 && yg < P.frame_h);
        if (!in) {
            __half2 pad = __floats2half2_rn(1024.0f, 1024.0f);
            store_ig(P, o, *reinterpret_cast<unsigned*>(&pad));
            if (P.If) reinterpret_cast<__half*>(P.If)[o + 4] = __float2half(0.0f);
            if (P.st) P.st[o] = make_float2(0.0f, 0.0f);
            if (P.Tg) store_tiled(P, x, y + PADY, *reinterpret_cast<unsigned*>(&pad), __float2half(0.0f), make_float2(0.0f, 0.0f));
            continue;
        }
        int s1 = 0, s2 = 0;
#pragma unroll
        for (int k = 0; k < WIN; k++) {
            s1 += hI[ty + k][tx];
            s2 += hII[ty + k][tx];
        }
        int ax = min(P.w - 1, x + RAD) - max(0, x - RAD) + 1;
        int ay = min(P.frame_h - 1, yg + RAD) - max(0, yg - RAD) + 1;
        float area = (float)(ax * ay);
        float mI = __fdiv_rn((float)s1, area);
        float mII = __fdiv_rn((float)s2, area);
        float var = __fsub_rn(mII, __fmul_rn(mI, mI));
        float c = (float)(1.0 / ((double)var + P.eps));
        float rxy = __fmul_rn(__frcp_rn((float)ax), __frcp_rn(P.S * (float)ay));
        int ic = sI[ty + RAD][tx + RAD];
        int il = (x - 1 >= 0) ? sI[ty + RAD][tx + RAD - 1] : ic;
        int ir = (x + 1 < P.w) ? sI[ty + RAD][tx + RAD + 1] : ic;
        __half2 ig = __floats2half2_rn((float)ic, (float)(il - ir));
        store_ig(P, o, *reinterpret_cast<unsigned*>(&ig));
        const float2 stv = make_float2(mI, __fmul_rn(c, rxy));
        if (P.If) reinterpret_cast<__half*>(P.If)[o + 4] = __float2half((float)ic);
        if (P.st) P.st[o] = stv;
        if (P.Tg) store_tiled(P, x, y + PADY, *reinterpret_cast<unsigned*>(&ig), __float2half((float)ic), stv);
        if (P.mean_u8) {
            int m = (int)mI;
            P.mean_u8[(size_t)y * P.w + x] = (m > 255) ? 255 : (unsigned char)m;
        }
    }
}

}  // namespace

int sbf_fused_supported(const sb200_params* p) {
    int nI, nG, S;
    return p->radius == RAD && find_lattice(p, &nI, &nG, &S);
}

size_t sbf_workspace_bytes(const sb200_ctx* ctx, int w, int h_held, int rows_out, int dabs, int size_d, int n_views) {
    const int padx = pad_x(dabs);
    const int pitch = (w + 2 * padx + 7) / 8 * 8;
    const size_t plane = (size_t)pitch * (h_held + 2 * PADY);
    const int pitchS = (w + 3) / 4 * 4;
    Plan plan = make_plan(w, rows_out, size_d, ctx->sm_count, n_views);
    size_t bytes = 0;
    const size_t recs = (size_t)plan.n_strips * (h_held + 2 * PADY);
    bytes += 2 * sb_align(plane * 4 * 4);  // IG x2, 4 shifted copies each
    bytes += 2 * (sb_align(recs * 1024) + sb_align(recs * 512) + sb_align(recs * 2048));  // strip-tiled Tg, TI, Tst x2
    bytes += sb_align((size_t)plan.n_chunks * 2 * rows_out * pitchS * 8);  // BL
    const size_t mma = sbf_mma_workspace_bytes(ctx, w, h_held, rows_out, dabs, size_d, n_views);
    return (bytes > mma ? bytes : mma) + 4096;
}

static int launch_prep(sb200_ctx* ctx, const sb200_params* p, const uint8_t* gray, const SbFusedGeom& g, int pitch, int padx,
                       float S, unsigned* IG, size_t shift_stride, float* If, float2* st, uint8_t* mean, unsigned* Tg = nullptr,
                       __half* TI = nullptr, float2* Tst = nullptr, int n_strips = 0) {
    PrepArgs P;
    P.gray = gray;
    P.w = g.w;
    P.h_held = g.h;
    P.y_global0 = g.y_global0;
    P.frame_h = g.frame_h;
    P.IG = IG;
    P.shift_stride = shift_stride;
    P.If = If;
    P.st = st;
    P.Tg = Tg;
    P.TI = TI;
    P.Tst = Tst;
    P.n_strips = n_strips;
    P.rows_pad = g.h + 2 * PADY;
    P.mean_u8 = mean;
    P.pitch = pitch;
    P.padx = padx;
    P.eps = p->eps;
    P.S = S;
    dim3 grid(sb_div_up(pitch, PT), sb_div_up(g.h + 2 * PADY, PT));
    SB_LAUNCH(ctx, k_prep, grid, 256, 0, P);
    return SB200_OK;
}

int sbf_prep_gray_planes(sb200_ctx* ctx, const sb200_params* p, const uint8_t* gray, const SbFusedGeom& g, int pitch, int padx,
                         float S, unsigned* IG, float* If, float2* st) {
    return launch_prep(ctx, p, gray, g, pitch, padx, S, IG, 0, If, st, nullptr);
}

// gray planes for the three-stage RGB kernel (fused_cvf_rgb3.cu): the 4 shifted (I,G) copies and the strip-tiled planes
int sbf_prep_gray_tiled(sb200_ctx* ctx, const sb200_params* p, const uint8_t* gray, const SbFusedGeom& g, int pitch, int padx,
                        float S, unsigned* IG, size_t shift_stride, unsigned* Tg, void* TI, float2* Tst, int n_strips) {
    return launch_prep(ctx, p, gray, g, pitch, padx, S, IG, shift_stride, nullptr, nullptr, nullptr, Tg,
                       reinterpret_cast<__half*>(TI), Tst, n_strips);
}

static int run_fused(sb200_ctx* ctx, const sb200_params* p, const uint8_t* const gray[2], const SbFusedGeom& g,
                     const int dmin[2], int size_d, int n_views, float* const best[2], float* const disp[2],
                     uint8_t* const mean[2]) {
    if (ctx->gray_kernel == 1 && p->guide_mode == SB200_GUIDE_GRAY && sbf_mma_supported(p))
        return sbf_run_fused_mma(ctx, p, gray, g, dmin, size_d, n_views, best, disp, mean);
    if (p->radius != RAD)
        return sb_fail(ctx, SB200_ERR_UNSUPPORTED, "fused kernel is built for radius %d (got %d): use box_mode stages", RAD,
                       p->radius);
    if (p->guide_mode != SB200_GUIDE_GRAY)
        return sb_fail(ctx, SB200_ERR_UNSUPPORTED, "fused kernel: gray guide only (the RGB guide runs through the staged path)");
    int nI, nG, S;
    if (!find_lattice(p, &nI, &nG, &S))
        return sb_fail(ctx, SB200_ERR_UNSUPPORTED,
                       "fused kernel: (alpha=%g, th_color=%g, th_grad=%g) has no exact integer cost lattice", p->alpha,
                       p->th_color, p->th_grad);
    if (size_d < 1 || g.w < 2 || g.h < 1 || g.rows_out < 1) return sb_fail(ctx, SB200_ERR_INVALID, "fused: bad shape");

    int dabs = 0;
    for (int v = 0; v < n_views; v++) dabs = max(dabs, max(abs(dmin[v]), abs(dmin[v] + size_d - 1)));
    const int padx = pad_x(dabs);
    const int pitch = (g.w + 2 * padx + 7) / 8 * 8;
    const int rows_pad = g.h + 2 * PADY;
    const size_t plane = (size_t)pitch * rows_pad;
    const int pitchS = (g.w + 3) / 4 * 4;
    Plan plan = make_plan(g.w, g.rows_out, size_d, ctx->sm_count, n_views);

    unsigned* IG[2];
    unsigned* Tg[2];
    __half* TI[2];
    float2* Tst[2];
    const size_t recs = (size_t)plan.n_strips * rows_pad;
    for (int i = 0; i < 2; i++) {
        IG[i] = sb_ws_alloc<unsigned>(ctx, 4 * plane);
        Tg[i] = sb_ws_alloc<unsigned>(ctx, recs * 256);
        TI[i] = sb_ws_alloc<__half>(ctx, recs * 256);
        Tst[i] = sb_ws_alloc<float2>(ctx, recs * 256);
    }
    const size_t planeS = (size_t)g.rows_out * pitchS;
    float2* BL = sb_ws_alloc<float2>(ctx, planeS * 2 * plan.n_chunks);
    if (!IG[0] || !IG[1] || !Tg[0] || !Tg[1] || !TI[0] || !TI[1] || !Tst[0] || !Tst[1] || !BL)
        return sb_fail(ctx, SB200_ERR_NOMEM, "fused: workspace arena too small (internal)");

    if (ctx->timing && ctx->ev_valid) SB_CUDA(ctx, cudaEventRecord(ctx->ev[0], ctx->stream));
    for (int i = 0; i < 2; i++) SB_TRY(launch_prep(ctx, p, gray[i], g, pitch, padx, (float)S, IG[i], plane, nullptr, nullptr, mean[i], Tg[i], TI[i], Tst[i],
                           plan.n_strips));
    if (ctx->timing && ctx->ev_valid) SB_CUDA(ctx, cudaEventRecord(ctx->ev[1], ctx->stream));

    FusedArgs A;
    const size_t origin = (size_t)PADY * pitch + padx;
    for (int i = 0; i < 2; i++) {
        A.IG[i] = IG[i];
        A.Tg[i] = reinterpret_cast<const uint4*>(Tg[i]);
        A.TI[i] = reinterpret_cast<const uint4*>(TI[i]);
        A.Tst[i] = reinterpret_cast<const uint4*>(Tst[i]);
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
    A.BL = BL;
    A.pitchS = pitchS;
    A.S = (float)S;
    __half2 wp = __floats2half2_rn((float)nI, (float)nG);
    __half2 tp = __floats2half2_rn(p->th_color, 2.0f * p->th_grad);
    A.zero = 0;
    A.wpack = *reinterpret_cast<unsigned*>(&wp);
    A.thpack = *reinterpret_cast<unsigned*>(&tp);

    const size_t smem = sizeof(SmemLayout);
    if (!ctx->fused_attr_set) {
        SB_CUDA(ctx, cudaFuncSetAttribute(k_fused_cvf, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        ctx->fused_attr_set = true;
    }
    A.n_views = n_views;
    const int nblocks = plan.n_strips * plan.n_bands * plan.n_chunks * n_views;
    SB_LAUNCH(ctx, k_fused_cvf, nblocks, K3_THREADS, smem, A);
    if (ctx->timing && ctx->ev_valid) SB_CUDA(ctx, cudaEventRecord(ctx->ev[2], ctx->stream));
    for (int v = 0; v < n_views; v++) {
        if (!best[v] && !disp[v]) continue;
        dim3 grid(sb_div_up(g.w, 256), g.rows_out);
        SB_LAUNCH(ctx, k_merge_chunks_bl, grid, 256, 0, BL, plan.n_chunks, v, g.rows_out, g.w, pitchS, best[v], disp[v]);
    }
    if (ctx->timing && ctx->ev_valid) SB_CUDA(ctx, cudaEventRecord(ctx->ev[3], ctx->stream));
    return SB200_OK;
}

int sbf_pair_disparity(sb200_ctx* ctx, const sb200_params* p, const uint8_t* gray_l, const uint8_t* gray_r,
                       const SbFusedGeom& g, float* bestL, float* dispL, float* bestR, float* dispR, uint8_t* meanL,
                       uint8_t* meanR) {
    const uint8_t* gray[2] = {gray_l, gray_r};
    const int dmin[2] = {p->dmin, -p->dmax};  // main.cu:79-82
    float* best[2] = {bestL, bestR};
    float* disp[2] = {dispL, dispR};
    uint8_t* mean[2] = {meanL, meanR};
    return run_fused(ctx, p, gray, g, dmin, p->dmax - p->dmin + 1, 2, best, disp, mean);
}

int sbf_view_disparity(sb200_ctx* ctx, const sb200_params* p, const uint8_t* guide, const uint8_t* other,
                       const SbFusedGeom& g, int dmin, int size_d, float* best, float* disp, uint8_t* mean) {
    const uint8_t* gray[2] = {guide, other};
    const int dm[2] = {dmin, dmin};
    float* b[2] = {best, nullptr};
    float* d[2] = {disp, nullptr};
    uint8_t* m[2] = {mean, nullptr};
    return run_fused(ctx, p, gray, g, dm, size_d, 1, b, d, m);
}
