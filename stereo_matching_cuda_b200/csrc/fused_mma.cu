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
// of 8 consecutive disparities, and marches down the rows four at a time (hand-offs between the roles every two rows:
// every buffer between two roles is two halves).  Tensor Memory lane = image column, so after an MMA every thread holds
// ITS column for all 8 disparities: guide operands are loaded once per column and row (not per disparity), and the
// winner-take-all over the 8 disparities is a register tournament.
//   role A (5 warps)   lattice cost P (packed half, 2 disparities per instruction) for 160 cost columns x 8 d x 4 rows;
//                      EXACT fp16 pieces of P and I*P:  P, (I&15)*P, (I&240)*P;  what enters the MMA is the vertical
//                      difference  piece(y) - piece(y-19)  (exact in fp16), the 19-row ring of P lives in shared memory
//   MMA 1 (tensor)     dH[x][n] = sum_k Band[x][k] * dPiece[k][n]   (K = 160, fp32 accumulate: exact integers)
//   role B (4 warps)   S_p += dH_p, S_Ip += dH_Ip  (the 2-D box sums, exact);  a, b (guidedFilter.cu:345-354); their
//                      vertical 19-row sums as running sums with a 19-slot ring in Tensor Memory (re-summed every 64
//                      rows, which bounds the float drift); split into fp16 hi + lo (22 significant bits, scaled by a
//                      power of two)
//   MMA 2 (tensor)     H_a, H_b = Band x (hi) - Band x (hi - value): the finished 19 x 19 window sums of a and b
//   role C (4 warps)   q = mean_a*(I-128) + mean(b + 128 a), tournament over the 8 disparities with the reference's
//                      `best >= q` rule, read-modify-write of the chunk's (best,label) plane: once per 8 disparities
//                      (fused_cvf.cu: once per 4); stateless between iterations.  The <true> instantiation also
//                      stores q itself (the filtered cost volume the reference drops after dispSelectOnGPU)
//   2 warps issue the MMAs (warp-uniform descriptors, one elected lane each), 1 warp runs the TMA producer.
// Operands come from strip-tiled planes written by three small preparation kernels (k_prep_ga/gb/mt) through
// cp.async.bulk into shared-memory rings; everything between the roles is mbarrier-synchronised.  Roles A and B wait
// ONCE per hand-off: the barrier they wait on anyway also collects the tensor core's "your output buffer is free"
// commit (see the barrier set-up) -- the critical warp's run time is the sum of its blocking round trips, not of its
// instructions (profiles/r2_ablations_k_fused_mma.txt).
// Layout, descriptor encodings, exactness and the accumulator's rounding (truncation) were measured first with
// tools/umma_probe.cu on the B200 (profiles/r2_umma_probe.txt).
#include "mma_dev.cuh"

namespace {

constexpr int M_ND = 8;         // disparities per group
constexpr int MR = 4;           // image rows per pipeline iteration (one hand-off between the roles per MR rows)
constexpr int M_NWA = 5;        // role-A warps: one thread per cost column, all MR rows
#ifndef MMA_BSPLIT
#define MMA_BSPLIT 1   // role B: 1 = one thread per a/b lane (8 disparities), 2 = two threads (4 disparities each)
#endif
#ifndef MMA_CSPLIT
#define MMA_CSPLIT 1   // role C: 1 = one thread per output lane (MR rows), 2 = two threads (MR/2 rows each)
#endif
constexpr int NWB = 4 * MMA_BSPLIT, NWC = 4 * MMA_CSPLIT;  // warps of roles B and C
constexpr int BD = M_ND / MMA_BSPLIT;                      // disparities per role-B thread
constexpr int CR = MR / MMA_CSPLIT;                        // rows per role-C thread
constexpr int M_WARPS = (NWB + NWC + M_NWA + 3 + 3) / 4 * 4;  // + 2 MMA warps + TMA warp, rounded up to whole warpgroups
constexpr int HR = 2;           // rows per hand-off between the roles: every buffer between two roles is two halves of HR rows
constexpr int NH = MR / HR;
constexpr int M_THREADS = 32 * M_WARPS;
constexpr int M_NOP = 4;        // operand ring (guide rows, a/b statistics, match rows): iterations in flight
constexpr int M_NGC = 4;        // ring of the output rows' intensities and previous (best,label) (role C runs behind the others)
#ifndef MMA_RESUM
#define MMA_RESUM 16
#endif
constexpr int M_RESUM = MMA_RESUM;  // role B re-sums its ring every M_RESUM iterations (64 rows); 8 -> 16: -1.5 % time, error 2.3e-5 -> 3.5e-5 of 1e-4
constexpr float I_CENTER = 128.0f;  // q = mean_a * (I - I_CENTER) + mean(b + I_CENTER * a)
static_assert((4 * RAD) % MR == 0, "MR must divide the warm-up length");

constexpr uint32_t GA_ROW = M_KB * 8;             // (I, G, I&15, I&240) as 4 halves per cost column
constexpr uint32_t GB_ROW = M_TW * 8;             // (mean_I, c2 * scale) per a/b lane
constexpr uint32_t GC_ROW = M_TW * 2;             // I - I_CENTER as half per output lane
constexpr uint32_t PB_ROW = M_TW * 8;             // (best,label) of the output lanes after the chunk's previous group
constexpr uint32_t GC_SLOT = MR * (GC_ROW + PB_ROW), GC_PB = MR * GC_ROW;
constexpr uint32_t MT_ROW = M_MTC * 64;
constexpr uint32_t OP_GA = 0, OP_GB = MR * GA_ROW, OP_MT = OP_GB + MR * GB_ROW, OP_BYTES = OP_MT + MR * MT_ROW;
constexpr uint32_t B1_GROUP = M_KB * 16;          // one N-group (8 columns) of B1: 160 rows x 16 B
constexpr uint32_t B1_HALF = 3 * HR * B1_GROUP;   // per half: dP (HR rows), d(lo) (HR), d(hi) (HR)
constexpr uint32_t B1_BYTES = NH * B1_HALF;
constexpr uint32_t B2_GROUP = M_K2 * 16;
constexpr uint32_t B2_HALF = 4 * HR * B2_GROUP;   // per half: hi (row, d-quad) x 2 HR, lo likewise
constexpr uint32_t B2_BYTES = NH * B2_HALF;
constexpr uint32_t PR_IL = M_KB * 16;             // a ring slot: P of 8 disparities (halves) per cost column, then the
constexpr uint32_t PR_SLOT = PR_IL + M_KB * 4;    // row's (I&15, I&240) per cost column (needed again when the row leaves)

struct MSmem {
    unsigned char b1[B1_BYTES];
    unsigned char b2[B2_BYTES];
    unsigned char pring[WIN][PR_SLOT];
    unsigned char op[M_NOP][OP_BYTES];
    unsigned char gc[M_NGC][GC_SLOT];
    float ry_lut[2][WIN + 1];  // [0][n] = scale/(S*n), [1][n] = 1/(scale*n); [.][0] = 0
    uint64_t op_full[M_NOP], op_empty[M_NOP], gc_full[M_NGC], gc_empty[M_NGC];
    uint64_t b1_full[NH], b1_empty[NH], d1_full[NH], d1_empty[NH], b2_full[NH], d2_full[NH], d2_empty[NH];
    uint32_t tmem_base;
};
static_assert(sizeof(MSmem) <= 227 * 1024, "shared memory budget");

// Tensor-Memory column map of a lane (512 columns)
constexpr uint32_t TC_BAND = 0;     // 80 columns: Band[lane][0..159] as packed halves
constexpr uint32_t TC_D1 = 80;      // 64: dH_p (row, d) | dH_Ip (row, d)
constexpr uint32_t TC_D2 = 144;     // 64: (row, d, {a,b})
constexpr uint32_t TC_RING = 208;   // 19 x 16: (H_a, H_b) x 8 d of the last 19 rows

struct MmaArgs {
    const uint2* GA[2];    // per IMAGE: [strip][rows_pad][160] (I, G, I&15, I&240) halves; pad (1024,1024,0,0)
    const float2* GB[2];   // per IMAGE: [strip][rows_pad][128] (mean_I, scale * c/(S*area)); 0 outside
    const __half* GC[2];   // per IMAGE: [strip][rows_pad][128] I - I_CENTER of the output lanes
    const uint4* MT[2];    // per IMAGE: [rows_pad][n_chunk][4] chunk i, copy s: I halves of X = 4i+s..+3, then G halves
    int rows_pad, n_chunk, padm;
    int w;
    int y_out0, rows_out, y_global0, frame_h;
    int dmin[2], size_d;
    int n_strips, n_bands, band_rows, n_chunks, chunk_d, n_views;
    float2* BL;
    float* QV[2];          // optional, per VIEW: filtered cost volume q, [size_d][rows_out][w]; NULL = not stored
    int pitchS;
    float S, scale, inv_scale;
    unsigned wI2, wG2, tc2, tg2;  // half2 broadcasts: lattice weights nI, nG; thresholds th_color, 2*th_grad
};

// Register budgets of the warpgroups.  setmaxnreg only moves registers inside the pool the block got at launch: the launch-time
// register count x threads (here 128 x 512 = the whole 64 K file; with 640 threads it is 96 x 640 = 61440) -- budgets that add up
// to more leave the last setmaxnreg.inc waiting forever.
constexpr int M_REGS_LAUNCH = (65536 / M_THREADS) / 8 * 8;
constexpr int M_REGS_B = MMA_BSPLIT == 1 ? 168 : 96;
constexpr int M_REGS_C = MMA_CSPLIT == 1 ? 120 : 72;
constexpr int M_REGS_A = (M_WARPS == 16) ? 112 : (M_WARPS == 20 ? 80 : 72);
static_assert(32 * NWB * M_REGS_B + 32 * NWC * M_REGS_C + 32 * (M_WARPS - NWB - NWC) * M_REGS_A <= M_THREADS * M_REGS_LAUNCH,
              "register pool of the block");

// QVOL: role C also stores the filtered cost q of every cell (sb200_view_volume_dev, subpixel_left).  A separate
// instantiation: the extra stores in the default kernel's role-C loop cost 6 % even when switched off at run time.
template <bool QVOL>
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
    // 36 warm-up rows fill both windows, then one output row per input row; iterations take MR rows (the last one may run
    // past the band: those rows read zero padding and are not stored)
    const int niter = ((yb1 - yb0) + 4 * RAD + MR - 1) / MR;
    constexpr int WARM_IT = 4 * RAD / MR;
    const int Ktotal = ngroups * niter;

    // ---- set-up: Tensor Memory, barriers, tables, zeroed shared memory ----
    if (warp == 0) tm_alloc(&sm.tmem_base);
    auto bar = [&](const uint64_t* b) { return smem_addr(b); };
    if (threadIdx.x == 32) {
        for (int i = 0; i < M_NOP; i++) {
            mbar_init(bar(&sm.op_full[i]), 2);
            mbar_init(bar(&sm.op_empty[i]), M_NWA + NWB);  // role A (guide and match rows), role B (statistics)
        }
        for (int i = 0; i < M_NGC; i++) {
            mbar_init(bar(&sm.gc_full[i]), 1);
            mbar_init(bar(&sm.gc_empty[i]), NWC);
        }
        for (int i = 0; i < NH; i++) {
            mbar_init(bar(&sm.b1_full[i]), M_NWA);
            mbar_init(bar(&sm.b1_empty[i]), 1);
            mbar_init(bar(&sm.d1_full[i]), 2);
            mbar_init(bar(&sm.d1_empty[i]), NWB);
            mbar_init(bar(&sm.b2_full[i]), NWB);
            mbar_init(bar(&sm.d2_full[i]), 1);
            mbar_init(bar(&sm.d2_empty[i]), 4);  // one warp per lane quarter finishes a half
        }
        // Two of the barriers collect a second, tensor-core arrival so that roles A and B wait once where they waited twice
        // (-3.5 % and -5.5 % kernel time, outputs bit-identical):
        //   d1_full[h], phase K   = MMA 1's commit of (K, h)  +  MMA 2's commit of (K-1, h)  ("B2 half h may be overwritten")
        //   op_full[s], iteration K = the operand bulk copies +  MMA 1's commit of (K-1, half 0)  ("B1 half 0 may be overwritten")
        // Neither second arrival can land in a later phase than its own: MMA 2 of (K, h) needs role B's output of (K, h), which
        // role B writes after waiting for phase K; MMA 1 of (K, 0) needs role A's rows of K, written after op_full of K.
        // Iteration 0 has no previous MMA: those arrivals are made here.
        for (int i = 0; i < NH; i++) mbar_arrive(bar(&sm.d1_full[i]));
        mbar_arrive(bar(&sm.op_full[0]));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x >= 64 && threadIdx.x < 64 + 2 * (WIN + 1)) {
        const int t = threadIdx.x - 64, n = t % (WIN + 1);
        float v = 0.0f;
        if (n > 0) v = (t < WIN + 1) ? A.scale * __frcp_rn(A.S * (float)n) : A.inv_scale * __frcp_rn((float)n);
        sm.ry_lut[t / (WIN + 1)][n] = v;
    }
    {   // cost columns 146..159 of B1 are never written and are multiplied by zero band entries: they must be finite
        uint4* z = reinterpret_cast<uint4*>(smem_raw);
        const int n16 = (int)(offsetof(MSmem, ry_lut) / 16);
        for (int i = threadIdx.x; i < n16; i += M_THREADS) z[i] = make_uint4(0u, 0u, 0u, 0u);
    }
    fence_async_smem();
    tm_fence_before();
    __syncthreads();
    tm_fence_after();
    const uint32_t tmem = sm.tmem_base;

    if (warp < NWB) {
        // ================= role B: 2-D box sums of P and I*P, a and b, their vertical window sums, fp16 hi/lo split ====
        // warp = (lane quarter, disparity part h): BD disparities x MR rows per thread and iteration.  The vertical 19-row
        // sums of a and b are running sums (ring of the last 19 rows in Tensor Memory, re-summed every M_RESUM iterations:
        // that bounds the rounding drift); MMA 2 then takes their horizontal sums, so role C receives finished window sums.
        reg_inc<M_REGS_B>();
        const int q4 = warp & 3, h = warp >> 2;
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
            named_bar_arrive(1, 128 + 64);  // the two MMA warps wait for the band
        }
        const int x = xa0 + l;
        const bool xin = x >= 0 && x < A.w;
        const float rx = xin ? __frcp_rn((float)(min(A.w - 1, x + RAD) - max(0, x - RAD) + 1)) : 0.0f;
        const float r1_whole = rx * sm.ry_lut[0][WIN];
        const uint32_t td1 = tl + TC_D1 + BD * h, tring = tl + TC_RING + 2 * BD * h;
        const uint32_t b2a = smem_addr(&sm.b2[0]) + (uint32_t)((l >> 3) * 128 + (l & 7) * 16) + (uint32_t)h * (BD / 4) * B2_GROUP;
        const uint32_t ops = smem_addr(&sm.op[0][0]) + OP_GB + (uint32_t)l * 8;
        const uint32_t mb_d1f = bar(&sm.d1_full[0]), mb_d1e = bar(&sm.d1_empty[0]), mb_b2f = bar(&sm.b2_full[0]),
                       mb_ope = bar(&sm.op_empty[0]);
        int K = 0;
        for (int g = 0; g < ngroups; g++) {
            float Sp[BD], Sip[BD], Va[BD], Vb[BD];
#pragma unroll
            for (int d = 0; d < BD; d++) Sp[d] = Sip[d] = Va[d] = Vb[d] = 0.0f;
            {
                uint32_t z[2 * BD];
#pragma unroll
                for (int i = 0; i < 2 * BD; i++) z[i] = 0u;
                for (int s = 0; s < WIN; s++) tm_stN(tring + 16 * s, z);
                tm_wait_st();
            }
            int slot = 0;
#pragma unroll 1
            for (int it = 0; it < niter; it++, K++) {
                const int yi0 = y_first + it * MR;
                const int ko = K & (M_NOP - 1);
                // the slot's statistics rows: role A waited for the slot before it wrote the B1 rows MMA 1 has consumed when
                // d1_full completes, so that barrier orders the slot's bulk copy before these loads as well
                mbar_wait(mb_d1f, (unsigned)K & 1u);
                uint2 st[MR];
#pragma unroll
                for (int r = 0; r < MR; r++) st[r] = lds64(ops + ko * OP_BYTES + r * GB_ROW);
                float r1[MR];
                {   // 1 / (S * clipped window area) of the a/b rows: a constant while the rows' windows are whole
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
                    uint32_t dp[2][BD], dip[2][BD], o[2][2 * BD];
                    int slots[2];
                    if (half > 0) mbar_wait(mb_d1f + 8 * half, (unsigned)K & 1u);
                    tm_fence_after();
#pragma unroll
                    for (int j = 0; j < 2; j++) {
                        slots[j] = slot;
                        slot = (slot + 1 == WIN) ? 0 : slot + 1;
                        tm_ldN(td1 + 16 * HR * half + 8 * j, dp[j]);
                        tm_ldN(td1 + 16 * HR * half + 8 * HR + 8 * j, dip[j]);
                        tm_ldN(tring + 16 * slots[j], o[j]);  // the (a,b) row that leaves the vertical window
                    }
                    tm_wait_ld();
                    // this half of D1 is in registers: MMA 1 of the next iteration may overwrite it
                    tm_fence_before();
                    __syncwarp();
                    mbar_arrive_lane0(mb_d1e + 8 * half, lane);
#pragma unroll
                    for (int j = 0; j < 2; j++) {
                        const int r = 2 * half + j;
                        const float mI = __uint_as_float(st[r].x), c2 = __uint_as_float(st[r].y);
                        const float mIc = I_CENTER - mI;
                        uint32_t ab[2 * BD];
#pragma unroll
                        for (int d = 0; d < BD; d++) {
                            Sp[d] += __uint_as_float(dp[j][d]);
                            Sip[d] += __uint_as_float(dip[j][d]);
                            const float cov = fmaf(-mI, Sp[d], Sip[d]);
                            const float a = cov * c2;
                            // b + 128 a: role C evaluates q = mean_a * (I - 128) + mean(b + 128 a), which halves the magnitudes
                            // that cancel in mean_a * I + mean_b (guidedFilter.cu:363-369) and with them the rounding error
                            const float b = fmaf(mIc, a, Sp[d] * r1[r]);
                            ab[d] = __float_as_uint(a);
                            ab[BD + d] = __float_as_uint(b);
                            Va[d] += a - __uint_as_float(o[j][d]);
                            Vb[d] += b - __uint_as_float(o[j][BD + d]);
                        }
                        tm_stN(tring + 16 * slots[j], ab);
                        if (r == MR - 1 && (it & (M_RESUM - 1)) == M_RESUM - 1) {
                            // re-sum the ring (it now ends with this row): bounds the drift of the running sums
                            tm_wait_st();
#pragma unroll
                            for (int d = 0; d < BD; d++) Va[d] = Vb[d] = 0.0f;
                            for (int s = 0; s < WIN; s++) {
                                uint32_t v[2 * BD];
                                tm_ldN(tring + 16 * s, v);
                                tm_wait_ld();
#pragma unroll
                                for (int d = 0; d < BD; d++) {
                                    Va[d] += __uint_as_float(v[d]);
                                    Vb[d] += __uint_as_float(v[BD + d]);
                                }
                            }
                        }
                        // fp16 hi + lo of the vertical sums: hi - value = -(lo part); MMA 2 takes the lo pass with B negated.
                        // A B2 group (8 columns of D2) = a of 4 disparities, then b of the same 4.
#pragma unroll
                        for (int q = 0; q < BD / 4; q++) {
                            uint32_t hi[4], lo[4];
#pragma unroll
                            for (int d = 0; d < 2; d++) {
                                const float a0 = Va[4 * q + 2 * d], a1 = Va[4 * q + 2 * d + 1], b0 = Vb[4 * q + 2 * d], b1 = Vb[4 * q + 2 * d + 1];
                                const unsigned ha = h22u(__floats2half2_rn(a0, a1)), hb = h22u(__floats2half2_rn(b0, b1));
                                hi[d] = ha;
                                hi[2 + d] = hb;
                                lo[d] = h22u(__floats2half2_rn(fhadd_lo(ha, -a0), fhadd_hi(ha, -a1)));
                                lo[2 + d] = h22u(__floats2half2_rn(fhadd_lo(hb, -b0), fhadd_hi(hb, -b1)));
                            }
                            sts128(b2a + half * B2_HALF + (uint32_t)(j * 2 + q) * B2_GROUP, hi[0], hi[1], hi[2], hi[3]);
                            sts128(b2a + half * B2_HALF + (uint32_t)(2 * HR + j * 2 + q) * B2_GROUP, lo[0], lo[1], lo[2], lo[3]);
                        }
                    }
                    tm_wait_st();
                    fence_async_smem();
                    __syncwarp();
                    if (lane == 0) {
                        mbar_arrive(mb_b2f + 8 * half);
                        if (half == NH - 1) mbar_arrive(mb_ope + 8 * ko);
                    }
                }
            }
        }
    } else if (warp < NWB + NWC) {
        // ================= role C: q = mean_a * I + mean_b, winner-take-all =================
        // warp = (lane quarter, row part c): CR rows of every iteration, all 8 disparities; no state between iterations
        if (M_REGS_C < M_REGS_LAUNCH) reg_dec<M_REGS_C>(); else if (M_REGS_C > M_REGS_LAUNCH) reg_inc<M_REGS_C>();
        const int q4 = warp & 3, c = (warp - NWB) >> 2;
        const int l = q4 * 32 + lane;
        const uint32_t tl = tmem + ((uint32_t)(q4 * 32) << 16);
        const int x = xo0 + l;
        const bool valid = l < M_VW && x < A.w;
        const float rx = valid ? __frcp_rn((float)(min(A.w - 1, x + RAD) - max(0, x - RAD) + 1)) : 0.0f;
        const float rxy_whole = rx * sm.ry_lut[1][WIN];
        const int pitchS = A.pitchS;
        const size_t planeS = (size_t)A.rows_out * pitchS;
        // this thread finishes rows yb0 + MR e + CR c + {0 .. CR-1}
        float2* const bl0 = A.BL + (size_t)(chunk * 2 + view) * planeS + (size_t)(yb0 - A.y_out0 + CR * c) * pitchS + x;
        const int band_rows = yb1 - yb0;
        // optional filtered-cost volume (compute_q per slice, guidedFilter.cu:363-369, kept instead of dropped after the WTA)
        const size_t qplane = (size_t)A.rows_out * A.w;
        float* const qv0 = (QVOL && A.QV[view]) ? A.QV[view] + (size_t)(chunk * A.chunk_d) * qplane + (size_t)(yb0 - A.y_out0 + CR * c) * A.w + x : nullptr;
        const uint32_t td2 = tl + TC_D2 + 16 * CR * c;
        const uint32_t gcs = smem_addr(&sm.gc[0][0]) + (uint32_t)(CR * c) * GC_ROW + (uint32_t)l * 2;
        const uint32_t pbs = smem_addr(&sm.gc[0][0]) + GC_PB + (uint32_t)(CR * c) * PB_ROW + (uint32_t)l * 8;
        const uint32_t mb_d2f = bar(&sm.d2_full[0]), mb_d2e = bar(&sm.d2_empty[0]), mb_gcf = bar(&sm.gc_full[0]), mb_gce = bar(&sm.gc_empty[0]);
        int K = 0;
        for (int g = 0; g < ngroups; g++) {
            const float dbase = (float)(dlo + g * M_ND);
            const int dact = min(M_ND, dcnt - g * M_ND);  // disparities of this group that exist
            // (best,label) after the chunk's previous group: the TMA producer brings the rows into the ring slot, M_NGC
            // iterations ahead (a register prefetch with ld.global put its latency on this role's critical path)
            const bool ld_ok = g > 0;
            const float2 binit = make_float2(BEST_INIT_BITS_F, 0.0f);
            float2* blp = bl0;
            float* qvp = (QVOL && qv0) ? qv0 + (size_t)(g * M_ND) * qplane : nullptr;
#pragma unroll 1
            for (int it = 0; it < niter; it++, K++) {
                const int e = it - WARM_IT;
                const int kc = K & (M_NGC - 1);
                const int row0 = e * MR + CR * c;  // first of this thread's rows in the band
                mbar_wait(mb_gcf + 8 * kc, (unsigned)(K / M_NGC) & 1u);
                if (e < 0) {
                    // warm-up iterations produce no output rows; the barriers still advance in step with the other roles
#pragma unroll
                    for (int jj = 0; jj < CR; jj += HR) mbar_wait(mb_d2f + 8 * ((CR * c + jj) / HR), (unsigned)K & 1u);
                    __syncwarp();
                    if (lane == 0) {
#pragma unroll
                        for (int jj = 0; jj < CR; jj += HR) mbar_arrive(mb_d2e + 8 * ((CR * c + jj) / HR));
                        mbar_arrive(mb_gce + 8 * kc);
                    }
                    continue;
                }
                float2 pb[CR];
                float rxy[CR], Ic[CR];
                bool st_ok[CR];
                {
                    const int yg = yb0 + row0 + A.y_global0;
                    const bool whole = yg >= RAD && yg + CR - 1 + RAD < A.frame_h;  // the rows' windows are not clipped
#pragma unroll
                    for (int j = 0; j < CR; j++) {
                        pb[j] = binit;
                        if (ld_ok) {
                            const uint2 v = lds64(pbs + kc * GC_SLOT + j * PB_ROW);
                            pb[j] = make_float2(__uint_as_float(v.x), __uint_as_float(v.y));
                        }
                        rxy[j] = whole ? rxy_whole : rx * inv_rows(sm.ry_lut[1], yb0 + row0 + j, A.y_global0, A.frame_h);
                        Ic[j] = __half2float(__ushort_as_half((unsigned short)lds16(gcs + kc * GC_SLOT + j * GC_ROW)));
                        st_ok[j] = valid && row0 + j < band_rows;
                    }
                }
#pragma unroll
                for (int jj = 0; jj < CR; jj += HR) {
                    uint32_t hh[2][16];
                    const int half = (CR * c + jj) / HR;  // CSPLIT == 1: compile-time; CSPLIT == 2: this warp's half
                    mbar_wait(mb_d2f + 8 * half, (unsigned)K & 1u);
                    tm_fence_after();
#pragma unroll
                    for (int j = 0; j < 2; j++) tm_ld16u(td2 + 16 * (jj + j), hh[j]);
                    tm_wait_ld();
                    // this half of D2 is in registers
                    tm_fence_before();
                    __syncwarp();
                    mbar_arrive_lane0(mb_d2e + 8 * half, lane);
#pragma unroll
                    for (int j2 = 0; j2 < 2; j2++) {
                        const int j = jj + j2;
                        // columns of a row: per group of 4 disparities, the window sums of a, then of b
                        float q[M_ND];
#pragma unroll
                        for (int d = 0; d < M_ND; d++) {
                            const float sa = __uint_as_float(hh[j2][8 * (d >> 2) + (d & 3)]), sb = __uint_as_float(hh[j2][8 * (d >> 2) + 4 + (d & 3)]);
                            q[d] = fmaf(sa, Ic[j], sb) * rxy[j];
                        }
                        if (dact < M_ND) {  // last, partial group
#pragma unroll
                            for (int d = 0; d < M_ND; d++)
                                if (d >= dact) q[d] = __int_as_float(0x7f800000);
                        }
                        if (QVOL && qvp && st_ok[j]) {
                            float* qd = qvp + (size_t)j * A.w;
#pragma unroll
                            for (int d = 0; d < M_ND; d++) {
                                if (d < dact) *qd = q[d];
                                qd += qplane;
                            }
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
                // the stores above are read back by the TMA producer (async proxy) one group later
                asm volatile("fence.proxy.async.global;" ::: "memory");
                __syncwarp();
                mbar_arrive_lane0(mb_gce + 8 * kc, lane);
            }
        }
    } else {
    if (M_REGS_A < M_REGS_LAUNCH) reg_dec<M_REGS_A>();  // the last warpgroups: role A, MMA issue, TMA producer
    if (warp < NWB + NWC + M_NWA) {
        // ================= role A: lattice cost, exact fp16 pieces, vertical differences =================
        // one thread per cost column: 8 disparities x MR rows per iteration
        const int k = (warp - NWB - NWC) * 32 + lane;
        const int x = xc0 + k;
        const bool used = k < M_KC && x >= 0 && x < A.w;
        const __half2 wI = used ? u2h2(A.wI2) : __float2half2_rn(0.0f);
        const __half2 wG = used ? u2h2(A.wG2) : __float2half2_rn(0.0f);
        const __half2 tc = u2h2(A.tc2), tg = u2h2(A.tg2);
        const uint32_t b1a = smem_addr(&sm.b1[0]) + (uint32_t)((k >> 3) * 128 + (k & 7) * 16);
        const uint32_t pr0 = smem_addr(&sm.pring[0][0]) + (uint32_t)k * 16;
        const uint32_t pi0 = smem_addr(&sm.pring[0][0]) + PR_IL + (uint32_t)k * 4;
        const uint32_t ops = smem_addr(&sm.op[0][0]);
        const uint32_t mb_opf = bar(&sm.op_full[0]), mb_ope = bar(&sm.op_empty[0]), mb_b1f = bar(&sm.b1_full[0]), mb_b1e = bar(&sm.b1_empty[0]);
        int K = 0;
        for (int g = 0; g < ngroups; g++) {
            // ring of P (and of the rows' guide pieces): zero at group start; a thread owns its column of the ring
            for (int s = 0; s < WIN; s++) {
                sts128(pr0 + (uint32_t)s * PR_SLOT, 0u, 0u, 0u, 0u);
                sts32(pi0 + (uint32_t)s * PR_SLOT, 0u);
            }
            // match operands of this thread: padded column X0 = x + d0 + padm, 8 consecutive pixels from copy X0 & 3
            const int d0 = dlo + g * M_ND;
            const int X0 = x + d0 + A.padm;
            const int i0 = (xc0 + d0 + A.padm) >> 2;  // first chunk of the slot (warp-uniform)
            const uint32_t mt_off = OP_MT + (uint32_t)((X0 >> 2) - i0) * 64 + (uint32_t)(X0 & 3) * 16;
            int slot = 0;  // ring slot of the iteration's first row: row (MR*it + r) mod 19
#pragma unroll 1
            for (int it = 0; it < niter; it++, K++) {
                const int ko = K & (M_NOP - 1);
                mbar_wait(mb_opf + 8 * ko, (unsigned)(K / M_NOP) & 1u);
                const uint32_t opa = ops + ko * OP_BYTES;
#pragma unroll
                for (int r = 0; r < MR; r++) {
                    const uint32_t so = (uint32_t)slot * PR_SLOT;
                    // guide (I, G | I&15, I&240) of the entering row; (I&15, I&240) of the row that leaves (19 rows up) and its
                    // costs come from this thread's ring slot, which the entering row then takes
                    const uint2 gn = lds64(opa + OP_GA + r * GA_ROW + (uint32_t)k * 8);
                    const uint4 m0 = lds128(opa + mt_off + r * MT_ROW), m1 = lds128(opa + mt_off + r * MT_ROW + 64);
                    const uint4 po = lds128(pr0 + so);
                    const uint32_t go = lds32(pi0 + so);
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
                    sts128(pr0 + so, pn[0], pn[1], pn[2], pn[3]);
                    sts32(pi0 + so, gn.y);
                    slot = (slot + 1 == WIN) ? 0 : slot + 1;
                    const int half = r / HR, rr = r % HR;
                    // MMA 1 of the previous iteration has read this half of B1
                    if (rr == 0 && K >= 1 && half > 0) mbar_wait(mb_b1e + 8 * half, (unsigned)(K - 1) & 1u);
                    const uint32_t bh = b1a + half * B1_HALF + (uint32_t)rr * B1_GROUP;
                    sts128(bh, dp[0], dp[1], dp[2], dp[3]);
                    sts128(bh + HR * B1_GROUP, dl[0], dl[1], dl[2], dl[3]);
                    sts128(bh + 2 * HR * B1_GROUP, dh[0], dh[1], dh[2], dh[3]);
                    if (rr == HR - 1) {
                        fence_async_smem();
                        __syncwarp();
                        if (lane == 0) {
                            mbar_arrive(mb_b1f + 8 * half);
                            if (r == MR - 1) mbar_arrive(mb_ope + 8 * ko);
                        }
                    }
                }
            }
        }
    } else if (warp == NWB + NWC + M_NWA) {
        // ================= MMA 1 issue: warp-uniform descriptors, one elected lane =================
        // per half: dH_p and dH_lo of HR rows in one N = 16 HR instruction chain, then dH_hi accumulated onto dH_lo
        named_bar_sync(1, 128 + 64);  // the band matrix is in Tensor Memory
        tm_fence_after();
        constexpr uint32_t IDPL = instr_desc(16 * HR, 0), IDH = instr_desc(8 * HR, 0);
        const uint64_t db1 = smem_desc(smem_addr(&sm.b1[0]), 128, B1_GROUP);
        const uint32_t ta = tmem + TC_BAND, td1 = tmem + TC_D1;
        const uint32_t mb_b1f = bar(&sm.b1_full[0]), mb_b1e = bar(&sm.b1_empty[0]), mb_d1f = bar(&sm.d1_full[0]), mb_d1e = bar(&sm.d1_empty[0]);
#pragma unroll 1
        for (int K = 0; K < Ktotal; K++) {
#pragma unroll
            for (int half = 0; half < NH; half++) {
                mbar_wait(mb_b1f + 8 * half, (unsigned)K & 1u);
                if (K >= 1) mbar_wait(mb_d1e + 8 * half, (unsigned)(K - 1) & 1u);
                tm_fence_after();
                if (elect_one()) {
                    const uint64_t dpl = db1 + (uint64_t)((half * B1_HALF) >> 4), dh = dpl + (uint64_t)((2 * HR * B1_GROUP) >> 4);
                    const uint32_t td = td1 + 16 * HR * half;
#pragma unroll
                    for (int j = 0; j < M_KB / 16; j++) umma_ts(td, ta + 8 * j, dpl + (uint64_t)(j * 16), IDPL, j > 0);
#pragma unroll
                    for (int j = 0; j < M_KB / 16; j++) umma_ts(td + 8 * HR, ta + 8 * j, dh + (uint64_t)(j * 16), IDH, 1);
                    umma_commit(mb_d1f + 8 * half);
                    if (half == 0) umma_commit(bar(&sm.op_full[0]) + 8 * ((K + 1) & (M_NOP - 1)));  // see the barrier set-up
                    else umma_commit(mb_b1e + 8 * half);
                }
                __syncwarp();
            }
        }
    } else if (warp == NWB + NWC + M_NWA + 2) {
        // ================= MMA 2 issue (its own warp: never queued behind a wait of MMA 1) =================
        named_bar_sync(1, 128 + 64);
        tm_fence_after();
        constexpr uint32_t ID = instr_desc(16 * HR, 0), IDN = instr_desc(16 * HR, 1);
        const uint64_t db2 = smem_desc(smem_addr(&sm.b2[0]), 128, B2_GROUP);
        const uint32_t ta = tmem + TC_BAND, td2 = tmem + TC_D2;
        const uint32_t mb_b2f = bar(&sm.b2_full[0]), mb_d2f = bar(&sm.d2_full[0]), mb_d2e = bar(&sm.d2_empty[0]);
#pragma unroll 1
        for (int K = 0; K < Ktotal; K++) {
#pragma unroll
            for (int half = 0; half < NH; half++) {
                mbar_wait(mb_b2f + 8 * half, (unsigned)K & 1u);
                if (K >= 1) mbar_wait(mb_d2e + 8 * half, (unsigned)(K - 1) & 1u);
                tm_fence_after();
                if (elect_one()) {
                    const uint64_t dhi = db2 + (uint64_t)((half * B2_HALF) >> 4), dlo = dhi + (uint64_t)((2 * HR * B2_GROUP) >> 4);
                    const uint32_t td = td2 + 16 * HR * half;
#pragma unroll
                    for (int j = 0; j < M_K2 / 16; j++) umma_ts(td, ta + 8 * j, dhi + (uint64_t)(j * 16), ID, j > 0);
#pragma unroll
                    for (int j = 0; j < M_K2 / 16; j++) umma_ts(td, ta + 8 * j, dlo + (uint64_t)(j * 16), IDN, 1);
                    umma_commit(mb_d2f + 8 * half);
                    umma_commit(bar(&sm.d1_full[0]) + 8 * half);  // "B2 half may be overwritten": see the barrier set-up
                }
                __syncwarp();
            }
        }
    } else if (warp == NWB + NWC + M_NWA + 1) {
        // ================= TMA producer (warp-uniform addresses, one elected lane issues) =================
        const char* GAp = reinterpret_cast<const char*>(A.GA[view] + (size_t)strip * A.rows_pad * M_KB);
        const char* GBp = reinterpret_cast<const char*>(A.GB[view] + (size_t)strip * A.rows_pad * M_TW);
        const char* GCp = reinterpret_cast<const char*>(A.GC[view] + (size_t)strip * A.rows_pad * M_TW);
        const char* MTp = reinterpret_cast<const char*>(A.MT[1 - view]);
        const size_t mt_pitch = (size_t)A.n_chunk * 64;
        const long long row0 = (long long)PADY + y_first;  // padded row of the first entering row of a group
        const uint32_t op_s = smem_addr(&sm.op[0][0]), gc_s = smem_addr(&sm.gc[0][0]);
        const uint32_t op_f = bar(&sm.op_full[0]), op_e = bar(&sm.op_empty[0]), gc_f = bar(&sm.gc_full[0]), gc_e = bar(&sm.gc_empty[0]);
        int K = 0;
        for (int g = 0; g < ngroups; g++) {
            const int d0 = dlo + g * M_ND;
            const int i0 = (xc0 + d0 + A.padm) >> 2;
            const char* ga_src = GAp + row0 * GA_ROW;
            const char* gb_src = GBp + (row0 - RAD) * GB_ROW;
            const char* gc_src = GCp + (row0 - 2 * RAD) * GC_ROW;
            const char* pb_src = reinterpret_cast<const char*>(A.BL + (size_t)(chunk * 2 + view) * ((size_t)A.rows_out * A.pitchS) +
                                                               (size_t)(yb0 - A.y_out0) * A.pitchS + xo0);
            const size_t pb_pitch = (size_t)A.pitchS * 8;
            const char* mt_src = MTp + row0 * mt_pitch + (size_t)i0 * 64;
#pragma unroll 1
            for (int it = 0; it < niter; it++, K++) {
                const int so = K & (M_NOP - 1), sc = K & (M_NGC - 1);
                if (K >= M_NOP) mbar_wait(op_e + 8 * so, (unsigned)(K / M_NOP - 1) & 1u);
                if (elect_one()) {
                    const uint32_t dst = op_s + so * OP_BYTES;
                    mbar_expect_tx(op_f + 8 * so, OP_BYTES);
                    bulk_g2s(dst + OP_GA, ga_src, MR * GA_ROW, op_f + 8 * so);
                    bulk_g2s(dst + OP_GB, gb_src, MR * GB_ROW, op_f + 8 * so);
#pragma unroll
                    for (int r = 0; r < MR; r++) bulk_g2s(dst + OP_MT + r * MT_ROW, mt_src + r * mt_pitch, MT_ROW, op_f + 8 * so);
                }
                __syncwarp();
                if (K >= M_NGC) mbar_wait(gc_e + 8 * sc, (unsigned)(K / M_NGC - 1) & 1u);
                if (elect_one()) {
                    const bool pb = g > 0 && it >= WARM_IT;  // rows past the band's end read the plane's padding; they are not used
                    mbar_expect_tx(gc_f + 8 * sc, MR * GC_ROW + (pb ? MR * PB_ROW : 0u));
                    bulk_g2s(gc_s + sc * GC_SLOT, gc_src, MR * GC_ROW, gc_f + 8 * sc);
                    if (pb) {
#pragma unroll
                        for (int r = 0; r < MR; r++)
                            bulk_g2s(gc_s + sc * GC_SLOT + GC_PB + r * PB_ROW, pb_src + (size_t)((it - WARM_IT) * MR + r) * pb_pitch, PB_ROW,
                                     gc_f + 8 * sc);
                    }
                }
                __syncwarp();
                ga_src += MR * GA_ROW;
                gb_src += MR * GB_ROW;
                gc_src += MR * GC_ROW;
                mt_src += MR * mt_pitch;
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

// GA: thread per (strip, padded row, cost column)
__global__ void __launch_bounds__(160) k_prep_ga(const PrepM P0, const PrepM P1) {
    const PrepM& P = blockIdx.z ? P1 : P0;
    const int k = threadIdx.x, strip = blockIdx.y;
    const int x = strip * M_VW - 2 * RAD + k;
    for (int yrow = blockIdx.x * PREP_ROWS; yrow < min(P.rows_pad, (int)(blockIdx.x + 1) * PREP_ROWS); yrow++) {
    const int y = yrow - PADY;
    float I, G;
    pix_ig(P, x, y, I, G);
    const bool in = in_frame(P, x, y);
    const int ii = in ? (int)I : 0;
    uint2 v;
    v.x = h22u(__floats2half2_rn(I, G));
    v.y = h22u(__floats2half2_rn((float)(ii & 15), (float)(ii & 240)));
    P.GA[((size_t)strip * P.rows_pad + yrow) * M_KB + k] = v;
    // the output lanes' intensities, centred (role C): lane l is cost column l + 18
    if (k >= 2 * RAD && k < 2 * RAD + M_TW)
        P.GC[((size_t)strip * P.rows_pad + yrow) * M_TW + (k - 2 * RAD)] = __float2half(in ? I - I_CENTER : 0.0f);
    }
}


// GB: block = (strip, GB_TR padded rows); exact integer window sums of I and I^2.  Horizontal 19-column sums as SLIDING sums
// along x: a task is (tile row, segment of 32 lanes), 19 taps to start and then +entering -leaving per lane (the first
// version took 19 taps per lane and row); vertical sums by one thread per a/b lane, sliding down the tile.
constexpr int GB_TR = 16;
constexpr int GB_ROWS = GB_TR + 2 * RAD;   // tile rows
constexpr int GB_SEG = 32, GB_NSEG = M_TW / GB_SEG;
__global__ void __launch_bounds__(M_TW) k_prep_gb(const PrepM P0, const PrepM P1) {
    const PrepM& P = blockIdx.z ? P1 : P0;
    __shared__ unsigned short sI[GB_ROWS][M_TW + 2 * RAD + 4];  // pitch 150 halves = 75 words: rows fall on different banks
    __shared__ int h1[GB_ROWS][M_TW + 1], h2[GB_ROWS][M_TW + 1];
    const int l = threadIdx.x, strip = blockIdx.y;
    const int yrow0 = blockIdx.x * GB_TR;
    const int xa0 = strip * M_VW - RAD;
    for (int i = l; i < GB_ROWS * (M_TW + 2 * RAD); i += M_TW) {
        const int py = i / (M_TW + 2 * RAD), px = i - py * (M_TW + 2 * RAD);
        const int x = xa0 + px - RAD, y = yrow0 - PADY + py - RAD;
        sI[py][px] = in_frame(P, x, y) ? P.gray[(size_t)y * P.w + x] : 0;
    }
    __syncthreads();
    for (int t = l; t < GB_NSEG * GB_ROWS; t += M_TW) {  // consecutive threads: consecutive rows of one segment
        const int seg = t / GB_ROWS, py = t - seg * GB_ROWS;
        const unsigned short* row = sI[py] + seg * GB_SEG;
        int s1 = 0, s2 = 0;
#pragma unroll
        for (int k = 0; k < WIN; k++) {
            const int v = row[k];
            s1 += v;
            s2 += v * v;
        }
        h1[py][seg * GB_SEG] = s1;
        h2[py][seg * GB_SEG] = s2;
#pragma unroll 8
        for (int i = 1; i < GB_SEG; i++) {
            const int a = row[i + WIN - 1], b = row[i - 1];
            s1 += a - b;
            s2 += a * a - b * b;
            h1[py][seg * GB_SEG + i] = s1;
            h2[py][seg * GB_SEG + i] = s2;
        }
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


}  // namespace

// the MMA kernel needs, beyond what the shuffle kernel needs, (I & 15) * P and (I & 240) * P exact in fp16
int sbf_mma_supported(const sb200_params* p) {
    int nI, nG, S;
    if (p->radius != RAD || !find_lattice(p, &nI, &nG, &S)) return 0;
    const double pmax = nI * (double)p->th_color + nG * 2.0 * (double)p->th_grad;
    return 15.0 * pmax <= 2047.0;
}

size_t sbf_mma_workspace_bytes(const sb200_ctx* ctx, int w, int h_held, int rows_out, int dabs, int size_d, int n_views) {
    Plan plan = make_plan_mma(w, rows_out, size_d, ctx->sm_count, n_views, M_ND);
    const int rows_pad = h_held + 2 * PADY;
    const MmaGeom mg = mma_geom(plan.n_strips, -dabs - M_ND, dabs + M_ND);
    const int pitchS = (w + 3) / 4 * 4;
    size_t bytes = 0;
    bytes += 2 * sb_align((size_t)plan.n_strips * rows_pad * M_KB * 8);
    bytes += 2 * sb_align((size_t)plan.n_strips * rows_pad * M_TW * 8);
    bytes += 2 * sb_align((size_t)plan.n_strips * rows_pad * M_TW * 2);
    bytes += 2 * sb_align((size_t)rows_pad * mg.n_chunk * 64);
    bytes += sb_align(((size_t)plan.n_chunks * 2 * rows_out + MR + 1) * pitchS * 8 + PB_ROW);
    return bytes + 4096;
}

int sbf_run_fused_mma(sb200_ctx* ctx, const sb200_params* p, const uint8_t* const gray[2], const SbFusedGeom& g,
                      const int dmin[2], int size_d, int n_views, float* const best[2], float* const disp[2],
                      uint8_t* const mean[2]) {
    int nI, nG, S;
    if (!sbf_mma_supported(p) || !find_lattice(p, &nI, &nG, &S))
        return sb_fail(ctx, SB200_ERR_UNSUPPORTED, "MMA fused kernel: radius 9 and a small exact cost lattice only");
    if (size_d < 1 || g.w < 2 || g.h < 1 || g.rows_out < 1) return sb_fail(ctx, SB200_ERR_INVALID, "fused: bad shape");
    Plan plan = make_plan_mma(g.w, g.rows_out, size_d, ctx->sm_count, n_views, M_ND);
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
    __half* GC[2];
    uint4* MT[2];
    for (int i = 0; i < 2; i++) {
        GA[i] = sb_ws_alloc<uint2>(ctx, (size_t)plan.n_strips * rows_pad * M_KB);
        GB[i] = sb_ws_alloc<float2>(ctx, (size_t)plan.n_strips * rows_pad * M_TW);
        GC[i] = sb_ws_alloc<__half>(ctx, (size_t)plan.n_strips * rows_pad * M_TW);
        MT[i] = sb_ws_alloc<uint4>(ctx, (size_t)rows_pad * mg.n_chunk * 4);
    }
    const size_t planeS = (size_t)g.rows_out * pitchS;
    float2* BL = sb_ws_alloc<float2>(ctx, planeS * 2 * plan.n_chunks + (size_t)(MR + 1) * pitchS + M_TW);  // + what a bulk read of the last rows may touch
    if (!GA[0] || !GA[1] || !GB[0] || !GB[1] || !GC[0] || !GC[1] || !MT[0] || !MT[1] || !BL)
        return sb_fail(ctx, SB200_ERR_NOMEM, "fused (mma): workspace arena too small (internal)");

    if (ctx->timing && ctx->ev_valid) SB_CUDA(ctx, cudaEventRecord(ctx->ev[0], ctx->stream));
    PrepM PP[2];
    for (int i = 0; i < 2; i++) {
        PrepM& P = PP[i];
        P.gray = gray[i];
        P.w = g.w;
        P.h_held = g.h;
        P.y_global0 = g.y_global0;
        P.frame_h = g.frame_h;
        P.n_strips = plan.n_strips;
        P.rows_pad = rows_pad;
        P.GA = GA[i];
        P.GB = GB[i];
        P.GC = GC[i];
        P.MT = MT[i];
        P.n_chunk = mg.n_chunk;
        P.padm = mg.padm;
        P.mean_u8 = mean[i];
        P.eps = p->eps;
        P.S = (float)S;
        P.scale = scale;
    }
    SB_LAUNCH(ctx, k_prep_ga, dim3(sb_div_up(rows_pad, PREP_ROWS), plan.n_strips, 2), M_KB, 0, PP[0], PP[1]);
    SB_LAUNCH(ctx, k_prep_gb, dim3(sb_div_up(rows_pad, GB_TR), plan.n_strips, 2), M_TW, 0, PP[0], PP[1]);
    SB_LAUNCH(ctx, k_prep_mt, dim3(sb_div_up(mg.n_chunk, 256), rows_pad, 2), 256, 0, PP[0], PP[1]);
    if (ctx->timing && ctx->ev_valid) SB_CUDA(ctx, cudaEventRecord(ctx->ev[1], ctx->stream));

    MmaArgs A;
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
    A.n_views = n_views;
    A.BL = BL;
    for (int v = 0; v < 2; v++) A.QV[v] = v < n_views ? ctx->qvol[v] : nullptr;
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
        SB_CUDA(ctx, cudaFuncSetAttribute(k_fused_mma<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        SB_CUDA(ctx, cudaFuncSetAttribute(k_fused_mma<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        ctx->mma_attr_set = true;
    }
    const int nblocks = plan.n_strips * plan.n_bands * plan.n_chunks * n_views;
    if (A.QV[0] || A.QV[1]) SB_LAUNCH(ctx, k_fused_mma<true>, nblocks, M_THREADS, smem, A);
    else SB_LAUNCH(ctx, k_fused_mma<false>, nblocks, M_THREADS, smem, A);
    if (ctx->timing && ctx->ev_valid) SB_CUDA(ctx, cudaEventRecord(ctx->ev[2], ctx->stream));
    for (int v = 0; v < n_views; v++) {
        if (!best[v] && !disp[v]) continue;
        dim3 grid(sb_div_up(g.w, 256), g.rows_out);
        SB_LAUNCH(ctx, k_merge_chunks_bl, grid, 256, 0, BL, plan.n_chunks, v, g.rows_out, g.w, pitchS, best[v], disp[v]);
    }
    if (ctx->timing && ctx->ev_valid) SB_CUDA(ctx, cudaEventRecord(ctx->ev[3], ctx->stream));
    return SB200_OK;
}
