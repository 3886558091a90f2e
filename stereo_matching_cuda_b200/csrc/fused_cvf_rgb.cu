// fused_cvf_rgb.cu -- the fused cost + guided filter + WTA kernel with an RGB guide.
//
// The reference's guide is always the gray image (main.cu:65-66); BASELINE.json's 1080p D=256
// configuration asks for a colour guide, defined in SURVEY A.8 as He et al.'s colour guided filter
// with the reference's box / border / WTA semantics:
//   cost p: the gray cost of costVolume.cu:163-190 (unchanged);
//   per pixel, once per frame:  mu = box(I), Sigma = box(I I^T) - mu mu^T, M = (Sigma + eps U)^-1;
//   per slice:  mp = box(p), cov_c = box(I_c p) - mu_c mp, a = M cov, b = mp - a.mu,
//               q = box(a).I + box(b);  winner-take-all as guidedFilter.cu:403-411.
// Same machinery as fused_cvf.cu (warp strips of 256 columns, producer/consumer warp pairs,
// shuffle-built 19-wide window sums, rings and hand-off in Tensor Memory, integer cost lattice so
// the four first-stage sums are exact), with 4 + 4 filtered quantities instead of 2 + 2:
//   producer: cost, sums of p, R p, G p, B p, 3x3 solve -> (a, b)   (ring: cost, TMEM)
//   consumer: sums of a_r, a_g, a_b, b, q, merge                   (ring: a_r,a_g in TMEM; a_b,b in shared memory)
// One row per pipeline iteration (the colour path needs about twice the registers per row).
#include "fused_dev.cuh"

namespace {

struct RgbArgs {
    const unsigned* IG[2];   // per IMAGE: padded half2 (I, G) of the gray image (cost)
    const uint2* C2[2];      // per IMAGE: padded {half2(R,G), half2(B,0)}, zero padding
    const float4* S1[2];     // per VIEW: padded (mu_r, mu_g, mu_b, M_rr)   M already scaled by 1/(S*area)
    const float4* S2[2];     // per VIEW: padded (M_rg, M_rb, M_gg, M_gb)
    const float* S3[2];      // per VIEW: padded M_bb
    int pitch, w, y_out0, rows_out, y_global0, frame_h;
    int dmin[2];
    int size_d;
    int n_strips, n_bands, band_rows, n_chunks, chunk_d, n_views;
    float* bestS;
    float* labS;
    int pitchS;
    float S;
    unsigned wpack, thpack;
    int zero;
};

struct RgbSmem {
    float4 ringB[NWARP][WIN][4][32];  // consumer-private: a_b planes 0,1  b planes 2,3
    float4 qbuf[2][NWARP][2][32];     // filtered row of each consumer warp, double buffered
    uint32_t tmem_base;
};
// TMEM columns of one lane: [0,304) ring of (a_r[8], a_g[8]) x 19, [304,380) ring of the cost x 19,
// [384,416) hand-off (a_r, a_g, a_b, b)
constexpr uint32_t TMR_RING_A = 0, TMR_RING_P = 304, TMR_HAND = 384;

struct RProdOps {
    uint4 g0, g1;     // guide gray (I,G) x8, row yi
    unsigned m[KPX];  // match gray (I,G) x8, row yi, columns x+d
    uint4 cn[4];      // guide colour x8, row yi
    uint4 co[4];      // guide colour x8, row yi-19
};
struct RProdPtrs {
    const unsigned* g;
    const unsigned* m;
    const uint2* cn;
    const uint2* co;
};
__device__ __forceinline__ void load_rprod(RProdOps& o, const RProdPtrs& p, int dep) {
    const uint4* pg = reinterpret_cast<const uint4*>(p.g + dep);
    o.g0 = __ldg(pg);
    o.g1 = __ldg(pg + 1);
    const unsigned* pm = p.m + dep;
#pragma unroll
    for (int j = 0; j < KPX; j++) o.m[j] = __ldg(pm + j);
    const uint4* pn = reinterpret_cast<const uint4*>(p.cn + dep);
    const uint4* po = reinterpret_cast<const uint4*>(p.co + dep);
#pragma unroll
    for (int k = 0; k < 4; k++) {
        o.cn[k] = __ldg(pn + k);
        o.co[k] = __ldg(po + k);
    }
}
__device__ __forceinline__ int touch(const RProdOps& o) {
    unsigned t = o.g0.x | o.g1.x;
#pragma unroll
    for (int j = 0; j < KPX; j++) t |= o.m[j];
#pragma unroll
    for (int k = 0; k < 4; k++) t |= o.cn[k].x | o.co[k].x;
    return (int)t;
}
// colour of pixel j out of 4 uint4 = 8 x {half2(R,G), half2(B,0)}
__device__ __forceinline__ void rgb_of(const uint4 (&c)[4], int j, float& r, float& g, float& b) {
    const uint4 q = c[j >> 1];
    const unsigned rg = (j & 1) ? q.z : q.x, b0 = (j & 1) ? q.w : q.y;
    const float2 f = __half22float2(u2h2(rg));
    r = f.x;
    g = f.y;
    b = __low2float(u2h2(b0));
}

__device__ __forceinline__ float inv_rows_rgb(int y, int y_global0, int frame_h, float scale) {
    int yg = y + y_global0;
    if (yg < 0 || yg >= frame_h) return 0.0f;
    int ay = min(frame_h - 1, yg + RAD) - max(0, yg - RAD) + 1;
    return __frcp_rn(scale * (float)ay);
}

__global__ void __launch_bounds__(NTHREADS, 1) k_fused_cvf_rgb(const RgbArgs A) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    RgbSmem& sm = *reinterpret_cast<RgbSmem*>(smem_raw);

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int pair = warp & (NWARP - 1);
    const bool consumer = warp >= NWARP;
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
    const int niter = (yb1 - yb0) + 4 * RAD;  // one row per iteration
    const int BAR_FULL = 2 + pair, BAR_EMPTY = 2 + NWARP + pair;

    if (warp == 0) tm_alloc(&sm.tmem_base);
    tm_fence_before();
    __syncthreads();
    tm_fence_after();
    const uint32_t tbase = sm.tmem_base + ((uint32_t)(pair * 32) << 16);
    const uint32_t tA = tbase + TMR_RING_A, tP = tbase + TMR_RING_P, tH = tbase + TMR_HAND;

    if (!consumer) {
        // ============================ PRODUCER: cost and first-stage sums ============================
        const unsigned* __restrict__ IGg = A.IG[view];
        const unsigned* __restrict__ IGm = A.IG[1 - view];
        const uint2* __restrict__ C2 = A.C2[view];
        __half2 wm[KPX];
#pragma unroll
        for (int j = 0; j < KPX; j++) {
            int x = xl + j;
            wm[j] = (x >= 0 && x < A.w) ? u2h2(A.wpack) : __float2half2_rn(0.0f);
        }
        const __half2 th = u2h2(A.thpack);
        const float4* __restrict__ S1 = A.S1[view];
        const float4* __restrict__ S2 = A.S2[view];
        const float* __restrict__ S3 = A.S3[view];
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
            const int d = dlo + dk;
            float VP[KPX], VR[KPX], VG[KPX], VB[KPX];
#pragma unroll
            for (int j = 0; j < KPX; j++) VP[j] = VR[j] = VG[j] = VB[j] = 0.0f;
            {
                const uint32_t z[4] = {0u, 0u, 0u, 0u};
                for (int s = 0; s < WIN; s++) tm_st4(tP + 4 * s, z);
                tm_wait_st();
            }
            __syncthreads();
            if (active) {
                RProdPtrs rp;
                const long long r0 = (long long)y_first * pitch + xl;
                rp.g = IGg + r0;
                rp.m = IGm + r0 + d;
                rp.cn = C2 + r0;
                rp.co = C2 + r0 - (long long)WIN * pitch;
                int slot = 0;
                // Operands live in ONE register set each and are re-fetched right after their last use:
                // the image rows (used by the cost and the vertical sums) while the four window sums
                // run, the 72 statistics words (used by the 3x3 solve at the end of the iteration) at the
                // end for the next row.  Each group of loads is made to depend on a word of the other
                // group, so its scoreboard wait is taken when the other group has long completed.
                RProdOps cur;
                load_rprod(cur, rp, 0);
                long long rs = (long long)(y_first - RAD) * pitch + xl;  // statistics row ya = yi - 9
                float4 s1[KPX], s2[KPX];
                float4 s3a, s3b;
                auto load_stats = [&](long long off) {
                    const float4* p1 = S1 + off;
                    const float4* p2 = S2 + off;
#pragma unroll
                    for (int j = 0; j < KPX; j++) {
                        s1[j] = __ldg(p1 + j);
                        s2[j] = __ldg(p2 + j);
                    }
                    const float4* p3 = reinterpret_cast<const float4*>(S3 + off);
                    s3a = __ldg(p3);
                    s3b = __ldg(p3 + 1);
                };
                load_stats(rs);
#pragma unroll 1
                for (int it = 0; it < niter; it++, rs += pitch) {
                    const int yi = y_first + it;
                    uint32_t pold[4];
                    tm_ld4(tP + 4 * slot, pold);
                    const unsigned gg[KPX] = {cur.g0.x, cur.g0.y, cur.g0.z, cur.g0.w, cur.g1.x, cur.g1.y, cur.g1.z, cur.g1.w};
                    __half ph[KPX];
                    float pn[KPX];
#pragma unroll
                    for (int j = 0; j < KPX; j++) {
                        __half2 gv = u2h2(gg[j]);
                        __half2 diff = __hsub2(gv, u2h2(cur.m[j]));
                        __half2 c = __hmin2(__habs2(diff), th);
                        __half2 pr = __hmul2(c, wm[j]);
                        ph[j] = __hadd(__low2half(pr), __high2half(pr));
                        pn[j] = __half2float(ph[j]);
                    }
                    uint32_t pnew[4];
                    pnew[0] = h22u(__halves2half2(ph[0], ph[1]));
                    pnew[1] = h22u(__halves2half2(ph[2], ph[3]));
                    pnew[2] = h22u(__halves2half2(ph[4], ph[5]));
                    pnew[3] = h22u(__halves2half2(ph[6], ph[7]));
                    tm_wait_ld();
                    tm_st4(tP + 4 * slot, pnew);
                    slot = (slot + 1 == WIN) ? 0 : slot + 1;
#pragma unroll
                    for (int j = 0; j < KPX; j++) {
                        const float2 f2 = __half22float2(u2h2(pold[j >> 1]));
                        const float po = (j & 1) ? f2.y : f2.x;
                        float rn, gn, bn, ro, go, bo;
                        rgb_of(cur.cn, j, rn, gn, bn);
                        rgb_of(cur.co, j, ro, go, bo);
                        VP[j] += pn[j] - po;
                        VR[j] = fmaf(rn, pn[j], VR[j]);
                        VG[j] = fmaf(gn, pn[j], VG[j]);
                        VB[j] = fmaf(bn, pn[j], VB[j]);
                        VR[j] = fmaf(-ro, po, VR[j]);
                        VG[j] = fmaf(-go, po, VG[j]);
                        VB[j] = fmaf(-bo, po, VB[j]);
                    }
                    {   // image rows of the next iteration, behind the wait on this row's statistics
                        unsigned tw = __float_as_uint(s3a.x) | __float_as_uint(s3b.x);
#pragma unroll
                        for (int j = 0; j < KPX; j++) tw |= __float_as_uint(s1[j].x) | __float_as_uint(s2[j].x);
                        rp.g += pitch;
                        rp.m += pitch;
                        rp.cn += pitch;
                        rp.co += pitch;
                        load_rprod(cur, rp, (int)tw & A.zero);
                    }
                    float SP[KPX], SR[KPX], SG[KPX], SB[KPX];
                    hsum19(VP, SP);
                    hsum19(VR, SR);
                    hsum19(VG, SG);
                    hsum19(VB, SB);
                    // ---- a = M cov, b = mp - a.mu at row ya = yi - 9
                    const float ry1 = inv_rows_rgb(yi - RAD, A.y_global0, A.frame_h, A.S);
                    const float s3[KPX] = {s3a.x, s3a.y, s3a.z, s3a.w, s3b.x, s3b.y, s3b.z, s3b.w};
                    float ar[KPX], ag[KPX], ab[KPX], bb[KPX];
#pragma unroll
                    for (int j = 0; j < KPX; j++) {
                        const float mr = s1[j].x, mg = s1[j].y, mb = s1[j].z;
                        const float cx = fmaf(-mr, SP[j], SR[j]);
                        const float cy = fmaf(-mg, SP[j], SG[j]);
                        const float cz = fmaf(-mb, SP[j], SB[j]);
                        ar[j] = fmaf(s1[j].w, cx, fmaf(s2[j].x, cy, s2[j].y * cz));
                        ag[j] = fmaf(s2[j].x, cx, fmaf(s2[j].z, cy, s2[j].w * cz));
                        ab[j] = fmaf(s2[j].y, cx, fmaf(s2[j].w, cy, s3[j] * cz));
                        const float mp = SP[j] * (rx[j] * ry1);
                        bb[j] = mp - fmaf(ar[j], mr, fmaf(ag[j], mg, ab[j] * mb));
                    }
                    load_stats(rs + pitch + (touch(cur) & A.zero));  // next row's statistics
                    if (it > 0) {
                        named_bar_sync(BAR_EMPTY, 64);
                        tm_fence_after();
                    }
                    tm_st16(tH, ar, ag);
                    tm_st16(tH + 16, ab, bb);
                    tm_wait_st();
                    tm_fence_before();
                    named_bar_arrive(BAR_FULL, 64);
                }
            }
            __syncthreads();
        }
    } else {
        // ============================ CONSUMER: 3x3 solve and second stage ============================
        const uint2* __restrict__ C2 = A.C2[view];
        float rx[KPX];
#pragma unroll
        for (int j = 0; j < KPX; j++) {
            int x = xl + j;
            int ax = min(A.w - 1, x + RAD) - max(0, x - RAD) + 1;
            rx[j] = (x >= 0 && x < A.w) ? __frcp_rn((float)ax) : 0.0f;
        }
        const int mc = 2 * (threadIdx.x - NWARP * 32);
        const int mx = xs + mc;
        const bool mvalid0 = (mc >= HALO) && (mc < HALO + VALID_W) && (mx < A.w);
        const bool mvalid1 = (mc + 1 >= HALO) && (mc + 1 < HALO + VALID_W) && (mx + 1 < A.w);
        const int mlane = mc >> 3, mj = mc & 7;
        const int qoff = ((mj >> 2) * 32 + mlane) * 4 + (mj & 3);
        const size_t planeS = (size_t)A.rows_out * A.pitchS;
        float* __restrict__ bestS = A.bestS + (size_t)(chunk * 2 + view) * planeS;
        float* __restrict__ labS = A.labS + (size_t)(chunk * 2 + view) * planeS;

        for (int g = 0; g < ngroups; g++) {
            const int dk = g * NWARP + pair;
            const bool active = dk < dcnt;
            const int dbase = dlo + g * NWARP;
            float Var[KPX], Vag[KPX], Vab[KPX], Vb[KPX];
#pragma unroll
            for (int j = 0; j < KPX; j++) Var[j] = Vag[j] = Vab[j] = Vb[j] = 0.0f;
            for (int s = 0; s < WIN; s++) {
                tm_st16(tA + 16 * s, Var, Vag);  // zeros
#pragma unroll
                for (int v = 0; v < 4; v++) sm.ringB[pair][s][v][lane] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
            tm_wait_st();
            if (!active) {
                const float inf = __int_as_float(0x7f800000);
#pragma unroll
                for (int b = 0; b < 2; b++)
#pragma unroll
                    for (int v = 0; v < 2; v++) sm.qbuf[b][pair][v][lane] = make_float4(inf, inf, inf, inf);
            }
            __syncthreads();

            int slot = 0, obuf = 0;
            auto merge = [&](int yq, float b0, float b1, float l0, float l1) {
                const float* qb = reinterpret_cast<const float*>(&sm.qbuf[obuf][0][0][0]);
#pragma unroll
                for (int wv = 0; wv < NWARP; wv++) {
                    float2 qv = *reinterpret_cast<const float2*>(qb + wv * 256 + qoff);
                    float lab = (float)(dbase + wv);
                    if (b0 >= qv.x) { b0 = qv.x; l0 = lab; }
                    if (b1 >= qv.y) { b1 = qv.y; l1 = lab; }
                }
                const size_t moff = (size_t)(yq - A.y_out0) * A.pitchS + mx;
                if (mvalid0) { bestS[moff] = b0; labS[moff] = l0; }
                if (mvalid1) { bestS[moff + 1] = b1; labS[moff + 1] = l1; }
                obuf ^= 1;
            };
            auto prefetch_best = [&](int yq, float& b0, float& b1, float& l0, float& l1) {
                b0 = b1 = BEST_INIT_BITS_F;
                l0 = l1 = 0.0f;
                if (g > 0) {
                    const size_t moff = (size_t)(yq - A.y_out0) * A.pitchS + mx;
                    if (mvalid0) { b0 = bestS[moff]; l0 = labS[moff]; }
                    if (mvalid1) { b1 = bestS[moff + 1]; l1 = labS[moff + 1]; }
                }
            };

            if (active) {
                // row pointers of THIS iteration's operands (stats at ya = yi-9, colour at yq = yi-18)
                long long rq = (long long)(y_first - 2 * RAD) * pitch + xl;  // colour row yq = yi - 18
#pragma unroll 1
                for (int it = 0; it < niter; it++, rq += pitch) {
                    const bool emit = it >= 4 * RAD;
                    const int yi = y_first + it;
                    const int yq = yi - 2 * RAD;
                    uint4 cq[4];  // used at the end of the iteration
                    {
                        const uint4* pc = reinterpret_cast<const uint4*>(C2 + rq);
#pragma unroll
                        for (int k = 0; k < 4; k++) cq[k] = __ldg(pc + k);
                    }
                    float pb0, pb1, pl0, pl1;
                    if (emit) prefetch_best(yq, pb0, pb1, pl0, pl1);
                    float aro[KPX], ago[KPX];
                    tm_ld16(tA + 16 * slot, aro, ago);
                    const float4 ob0 = sm.ringB[pair][slot][0][lane], ob1 = sm.ringB[pair][slot][1][lane];
                    const float4 ob2 = sm.ringB[pair][slot][2][lane], ob3 = sm.ringB[pair][slot][3][lane];
                    named_bar_sync(BAR_FULL, 64);
                    tm_fence_after();
                    float ar[KPX], ag[KPX], ab[KPX], bb[KPX];  // (a, b) of row ya = yi - 9, solved by the producer
                    tm_ld16(tH, ar, ag);
                    tm_ld16(tH + 16, ab, bb);
                    tm_wait_ld();
                    if (it + 1 < niter) {
                        tm_fence_before();
                        named_bar_arrive(BAR_EMPTY, 64);
                    }
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
                    if (emit) {
                        float SAr[KPX], SAg[KPX], SAb[KPX], SBb[KPX];
                        hsum19(Var, SAr);
                        hsum19(Vag, SAg);
                        hsum19(Vab, SAb);
                        hsum19(Vb, SBb);
                        const float ry2 = inv_rows_rgb(yq, A.y_global0, A.frame_h, 1.0f);
                        float q[KPX];
#pragma unroll
                        for (int j = 0; j < KPX; j++) {
                            float r, gch, b;
                            rgb_of(cq, j, r, gch, b);
                            q[j] = fmaf(SAr[j], r, fmaf(SAg[j], gch, fmaf(SAb[j], b, SBb[j]))) * (rx[j] * ry2);
                        }
                        sm.qbuf[obuf][pair][0][lane] = make_float4(q[0], q[1], q[2], q[3]);
                        sm.qbuf[obuf][pair][1][lane] = make_float4(q[4], q[5], q[6], q[7]);
                        named_bar_sync(1, NWARP * 32);
                        merge(yq, pb0, pb1, pl0, pl1);
                    }
                    tm_wait_st();
                }
            } else {
                for (int yq = yb0; yq < yb1; yq++) {
                    float pb0, pb1, pl0, pl1;
                    prefetch_best(yq, pb0, pb1, pl0, pl1);
                    named_bar_sync(1, NWARP * 32);
                    merge(yq, pb0, pb1, pl0, pl1);
                }
            }
            __syncthreads();
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
// Per-frame preparation for the colour guide: padded planes C2 (colour as halves), S1/S2/S3
// (mu and the scaled inverse covariance).  Window sums of R,G,B and their 6 products are exact
// integers (<= 361*255^2 < 2^31); the 3x3 inverse is evaluated in double like the oracle.
constexpr int RPT = 16;              // output tile
constexpr int RPP = RPT + 2 * RAD;   // 34

struct RgbPrepArgs {
    const uint8_t* rgb;  // held rows, interleaved, `ch` bytes per pixel
    int ch, w, h_held, y_global0, frame_h;
    uint2* C2;
    float4* S1;
    float4* S2;
    float* S3;
    int pitch, padx;
    double eps;
    float S;
};

__global__ void __launch_bounds__(256) k_prep_rgb(const RgbPrepArgs P) {
    __shared__ int sC[3][RPP][RPP + 1];
    __shared__ int hS[9][RPP][RPT + 1];
    const int x0 = blockIdx.x * RPT - P.padx;
    const int y0 = blockIdx.y * RPT - PADY;
    const int tid = threadIdx.x;
    for (int i = tid; i < RPP * RPP; i += 256) {
        int py = i / RPP, px = i - py * RPP;
        int x = x0 + px - RAD, y = y0 + py - RAD;
        int yg = y + P.y_global0;
        bool in = (x >= 0 && x < P.w && y >= 0 && y < P.h_held && yg >= 0 && yg < P.frame_h);
        const uint8_t* q = P.rgb + ((size_t)y * P.w + x) * P.ch;
        sC[0][py][px] = in ? (int)q[0] : 0;
        sC[1][py][px] = in ? (int)q[1] : 0;
        sC[2][py][px] = in ? (int)q[2] : 0;
    }
    __syncthreads();
    for (int i = tid; i < RPP * RPT; i += 256) {
        int py = i / RPT, tx = i - py * RPT;
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
    for (int i = tid; i < RPT * RPT; i += 256) {
        int ty = i / RPT, tx = i - ty * RPT;
        int x = x0 + tx, y = y0 + ty;
        if (x + P.padx >= P.pitch || y + PADY >= n_rows_pad) continue;
        const size_t o = (size_t)(y + PADY) * P.pitch + (x + P.padx);
        int yg = y + P.y_global0;
        bool in = (x >= 0 && x < P.w && y >= 0 && y < P.h_held && yg >= 0 && yg < P.frame_h);
        if (!in) {
            P.C2[o] = make_uint2(0u, 0u);
            P.S1[o] = make_float4(0.f, 0.f, 0.f, 0.f);
            P.S2[o] = make_float4(0.f, 0.f, 0.f, 0.f);
            P.S3[o] = 0.0f;
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
        P.S1[o] = make_float4(mr, mg, mb, __fmul_rn((float)(a00 * id), rxy));
        P.S2[o] = make_float4(__fmul_rn((float)(a01 * id), rxy), __fmul_rn((float)(a02 * id), rxy),
                              __fmul_rn((float)(a11 * id), rxy), __fmul_rn((float)(a12 * id), rxy));
        P.S3[o] = __fmul_rn((float)(a22 * id), rxy);
        __half2 rg = __floats2half2_rn((float)sC[0][ty + RAD][tx + RAD], (float)sC[1][ty + RAD][tx + RAD]);
        __half2 b0 = __floats2half2_rn((float)sC[2][ty + RAD][tx + RAD], 0.0f);
        P.C2[o] = make_uint2(*reinterpret_cast<unsigned*>(&rg), *reinterpret_cast<unsigned*>(&b0));
    }
}

}  // namespace

// gray prep of fused_cvf.cu (IG planes only are needed here, but it is one launch per image either way)
int sbf_prep_gray_planes(sb200_ctx* ctx, const sb200_params* p, const uint8_t* gray, const SbFusedGeom& g, int pitch, int padx,
                         float S, unsigned* IG, float* If, float2* st);

size_t sbf_rgb_workspace_bytes(const sb200_ctx* ctx, int w, int h_held, int rows_out, int dabs, int size_d) {
    const int padx = pad_x(dabs);
    const int pitch = (w + 2 * padx + 7) / 8 * 8;
    const size_t plane = (size_t)pitch * (h_held + 2 * PADY);
    const int pitchS = (w + 3) / 4 * 4;
    Plan plan = make_plan(w, rows_out, size_d, ctx->sm_count, 2);
    size_t bytes = 0;
    bytes += 2 * (sb_align(plane * 4) * 2 + sb_align(plane * 8));                       // IG, If, st (gray prep)
    bytes += 2 * (sb_align(plane * 8) + 2 * sb_align(plane * 16) + sb_align(plane * 4));  // C2, S1, S2, S3
    bytes += 2 * sb_align((size_t)plan.n_chunks * 2 * rows_out * pitchS * 4);
    return bytes + 8192;
}

int sbf_pair_disparity_rgb(sb200_ctx* ctx, const sb200_params* p, const uint8_t* rgb_l, const uint8_t* rgb_r, int channels,
                           const uint8_t* gray_l, const uint8_t* gray_r, const SbFusedGeom& g, float* bestL, float* dispL,
                           float* bestR, float* dispR) {
    if (p->radius != RAD) return sb_fail(ctx, SB200_ERR_UNSUPPORTED, "fused RGB kernel is built for radius %d", RAD);
    int nI, nG, S;
    if (!find_lattice(p, &nI, &nG, &S))
        return sb_fail(ctx, SB200_ERR_UNSUPPORTED, "fused RGB kernel: no exact integer cost lattice for these parameters");
    const int size_d = p->dmax - p->dmin + 1;
    const int dmin[2] = {p->dmin, -p->dmax};
    const int dabs = max(abs(p->dmin), abs(p->dmax));
    const int padx = pad_x(dabs);
    const int pitch = (g.w + 2 * padx + 7) / 8 * 8;
    const int rows_pad = g.h + 2 * PADY;
    const size_t plane = (size_t)pitch * rows_pad;
    const int pitchS = (g.w + 3) / 4 * 4;
    Plan plan = make_plan(g.w, g.rows_out, size_d, ctx->sm_count, 2);
    const uint8_t* gray[2] = {gray_l, gray_r};
    const uint8_t* rgb[2] = {rgb_l, rgb_r};
    unsigned* IG[2];
    float* If[2];
    float2* st[2];
    uint2* C2[2];
    float4 *S1[2], *S2[2];
    float* S3[2];
    for (int i = 0; i < 2; i++) {
        IG[i] = sb_ws_alloc<unsigned>(ctx, plane);
        If[i] = sb_ws_alloc<float>(ctx, plane);
        st[i] = sb_ws_alloc<float2>(ctx, plane);
        C2[i] = sb_ws_alloc<uint2>(ctx, plane);
        S1[i] = sb_ws_alloc<float4>(ctx, plane);
        S2[i] = sb_ws_alloc<float4>(ctx, plane);
        S3[i] = sb_ws_alloc<float>(ctx, plane);
        if (!IG[i] || !If[i] || !st[i] || !C2[i] || !S1[i] || !S2[i] || !S3[i])
            return sb_fail(ctx, SB200_ERR_NOMEM, "fused RGB: workspace arena too small (internal)");
    }
    const size_t planeS = (size_t)g.rows_out * pitchS;
    float* bestS = sb_ws_alloc<float>(ctx, planeS * 2 * plan.n_chunks);
    float* labS = sb_ws_alloc<float>(ctx, planeS * 2 * plan.n_chunks);
    if (!bestS || !labS) return sb_fail(ctx, SB200_ERR_NOMEM, "fused RGB: workspace arena too small (internal)");

    if (ctx->timing && ctx->ev_valid) SB_CUDA(ctx, cudaEventRecord(ctx->ev[0], ctx->stream));
    for (int i = 0; i < 2; i++) {
        SB_TRY(sbf_prep_gray_planes(ctx, p, gray[i], g, pitch, padx, (float)S, IG[i], If[i], st[i]));
        RgbPrepArgs P;
        P.rgb = rgb[i];
        P.ch = channels;
        P.w = g.w;
        P.h_held = g.h;
        P.y_global0 = g.y_global0;
        P.frame_h = g.frame_h;
        P.C2 = C2[i];
        P.S1 = S1[i];
        P.S2 = S2[i];
        P.S3 = S3[i];
        P.pitch = pitch;
        P.padx = padx;
        P.eps = p->eps;
        P.S = (float)S;
        dim3 grid(sb_div_up(pitch, RPT), sb_div_up(rows_pad, RPT));
        SB_LAUNCH(ctx, k_prep_rgb, grid, 256, 0, P);
    }
    if (ctx->timing && ctx->ev_valid) SB_CUDA(ctx, cudaEventRecord(ctx->ev[1], ctx->stream));

    RgbArgs A;
    const size_t origin = (size_t)PADY * pitch + padx;
    for (int i = 0; i < 2; i++) {
        A.IG[i] = IG[i] + origin;
        A.C2[i] = C2[i] + origin;
        A.S1[i] = S1[i] + origin;
        A.S2[i] = S2[i] + origin;
        A.S3[i] = S3[i] + origin;
        A.dmin[i] = dmin[i];
    }
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
    A.bestS = bestS;
    A.labS = labS;
    A.pitchS = pitchS;
    A.S = (float)S;
    A.zero = 0;
    __half2 wp = __floats2half2_rn((float)nI, (float)nG);
    __half2 tp = __floats2half2_rn(p->th_color, 2.0f * p->th_grad);
    A.wpack = *reinterpret_cast<unsigned*>(&wp);
    A.thpack = *reinterpret_cast<unsigned*>(&tp);
    const size_t smem = sizeof(RgbSmem);
    SB_CUDA(ctx, cudaFuncSetAttribute(k_fused_cvf_rgb, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int nblocks = plan.n_strips * plan.n_bands * plan.n_chunks * 2;
    SB_LAUNCH(ctx, k_fused_cvf_rgb, nblocks, NTHREADS, smem, A);
    if (ctx->timing && ctx->ev_valid) SB_CUDA(ctx, cudaEventRecord(ctx->ev[2], ctx->stream));
    float* best[2] = {bestL, bestR};
    float* disp[2] = {dispL, dispR};
    for (int v = 0; v < 2; v++) {
        if (!best[v] && !disp[v]) continue;
        dim3 grid(sb_div_up(g.w, 256), g.rows_out);
        SB_LAUNCH(ctx, k_merge_chunks, grid, 256, 0, bestS, labS, plan.n_chunks, v, g.rows_out, g.w, pitchS, best[v], disp[v]);
    }
    if (ctx->timing && ctx->ev_valid) SB_CUDA(ctx, cudaEventRecord(ctx->ev[3], ctx->stream));
    return SB200_OK;
}
