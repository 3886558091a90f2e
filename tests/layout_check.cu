// Host-only checks of the data-layout arithmetic the fused kernels rely on (no GPU needed: compiled with nvcc, only
// __host__ code runs).  Built and run by tests/test_layout_host.py.
//   1. strip tiling: every frame column lands in the strip(s) that read it, at the lane/pixel the kernels assume
//      (lane L of a warp holds strip-local columns 8L..8L+7; its c-th 128-bit load is 16-byte unit c*32+L of a record);
//   2. de-interleaved match planes: the two pointers of match_ptrs address exactly the stored positions of the 8
//      consecutive pixels starting at a linear element index e;
//   3. make_plan: the (strip x band x chunk) tiling covers every column, row and disparity exactly once;
//   4. find_lattice: the reference's default cost weights have the exact integer lattice DESIGN.md 4.2 quotes.
#include <cstdio>
#include <vector>

#include "../stereo_matching_cuda_b200/csrc/fused_dev.cuh"

static int fails = 0;
#define CHECK(cond, ...)                  \
    do {                                  \
        if (!(cond)) {                    \
            if (fails < 20) {             \
                printf("FAIL %s:%d: ", __FILE__, __LINE__); \
                printf(__VA_ARGS__);      \
                printf("\n");             \
            }                             \
            fails++;                      \
        }                                 \
    } while (0)

static void check_strip_tiling(int w) {
    const int n_strips = (w + VALID_W - 1) / VALID_W;
    const size_t rows_pad = 3;
    std::vector<int> seen((size_t)n_strips * SW, 0);
    for (int x = -300; x < w + 300; x++) {
        int strip[2], cl[2];
        const int n = strips_of_column(x, n_strips, strip, cl);
        for (int k = 0; k < n; k++) {
            CHECK(strip[k] >= 0 && strip[k] < n_strips && cl[k] >= 0 && cl[k] < SW, "x=%d strip=%d cl=%d", x, strip[k], cl[k]);
            // the kernels' statement of the layout: lane L of the strip's warp holds columns xs + 8L + j
            CHECK(strip[k] * VALID_W - HALO + cl[k] == x, "x=%d strip=%d cl=%d", x, strip[k], cl[k]);
            seen[(size_t)strip[k] * SW + cl[k]]++;
        }
        // valid output columns of a strip are [HALO, HALO+VALID_W): every image column is valid in exactly one strip
        if (x >= 0 && x < w) {
            int nvalid = 0;
            for (int k = 0; k < n; k++) nvalid += (cl[k] >= HALO && cl[k] < HALO + VALID_W);
            CHECK(nvalid == 1, "x=%d valid in %d strips", x, nvalid);
        }
    }
    for (size_t i = 0; i < seen.size(); i++) CHECK(seen[i] == 1, "strip column %zu written %d times", i, seen[i]);
    // index functions against "16-byte unit c*32+L, word/half k inside it"
    for (size_t rec = 0; rec < rows_pad; rec++)
        for (int cl = 0; cl < SW; cl++) {
            const int L = cl >> 3, j = cl & 7;
            // (I,G): 4-byte words, 4 per unit, unit = chunk (j>>2)
            CHECK(tg_index(rec, cl) == (rec * 64 + (size_t)(j >> 2) * 32 + L) * 4 + (j & 3), "tg cl=%d", cl);
            // I: halves, 8 per unit, one chunk
            CHECK(ti_index(rec, cl) == (rec * 32 + L) * 8 + j, "ti cl=%d", cl);
            // (mean_I, c2): float2, 2 per unit, chunk = j>>1
            CHECK(tst_index(rec, cl) == (rec * 128 + (size_t)(j >> 1) * 32 + L) * 2 + (j & 1), "tst cl=%d", cl);
            for (int ch = 0; ch < 3; ch++)  // colour: chunk = channel
                CHECK(tc_index(rec, ch, cl) == (rec * 96 + (size_t)ch * 32 + L) * 8 + j, "tc cl=%d", cl);
            CHECK(ts_s1_index(rec, cl) == rec * 576 + (size_t)(2 * j) * 32 + L, "ts1 cl=%d", cl);
            CHECK(ts_s2_index(rec, cl) == rec * 576 + (size_t)(2 * j + 1) * 32 + L, "ts2 cl=%d", cl);
            CHECK(ts_s3_index(rec, cl) == (rec * 576 + (size_t)(16 + (j >> 2)) * 32 + L) * 4 + (j & 3), "ts3 cl=%d", cl);
        }
    // records do not overlap and are dense: all indices of a record fall inside it, each exactly once
    {
        std::vector<int> g(64 * 4, 0), ih(32 * 8, 0), st(128 * 2, 0), tc(96 * 8, 0), ts(576 * 4, 0);
        for (int cl = 0; cl < SW; cl++) {
            g[tg_index(1, cl) - 64 * 4]++;
            ih[ti_index(1, cl) - 32 * 8]++;
            st[tst_index(1, cl) - 128 * 2]++;
            for (int ch = 0; ch < 3; ch++) tc[tc_index(1, ch, cl) - 96 * 8]++;
            for (int k = 0; k < 4; k++) {
                ts[(ts_s1_index(1, cl) - 576) * 4 + k]++;
                ts[(ts_s2_index(1, cl) - 576) * 4 + k]++;
            }
            ts[ts_s3_index(1, cl) - 576 * 4]++;
        }
        for (int v : g) CHECK(v == 1, "Tg record not dense");
        for (int v : ih) CHECK(v == 1, "TI record not dense");
        for (int v : st) CHECK(v == 1, "Tst record not dense");
        for (int v : tc) CHECK(v == 1, "TC record not dense");
        for (int v : ts) CHECK(v == 1, "TS record not dense");
    }
}

static void check_deinterleave() {
    const size_t half_plane = 4096;  // elements per half of a copy (multiple of 4)
    std::vector<unsigned> copy(2 * half_plane, 0xffffffffu);
    for (size_t i = 0; i < 2 * half_plane; i++) {
        const size_t k = deint_index(i, half_plane);
        CHECK(k < 2 * half_plane && copy[k] == 0xffffffffu, "deint_index(%zu) = %zu collides", i, k);
        copy[k] = (unsigned)i;  // the stored element remembers its linear index
    }
    for (long long e = 4; e + 8 <= (long long)(2 * half_plane); e += 4) {
        const unsigned *p0, *p1;
        match_ptrs(copy.data(), e, half_plane, p0, p1);
        for (int k = 0; k < 4; k++) {
            CHECK(p0[k] == (unsigned)(e + k), "e=%lld first load word %d holds %u", e, k, p0[k]);
            CHECK(p1[k] == (unsigned)(e + 4 + k), "e=%lld second load word %d holds %u", e, k, p1[k]);
        }
        CHECK(((p0 - copy.data()) & 3) == 0 && ((p1 - copy.data()) & 3) == 0, "e=%lld unaligned", e);
        // lanes are 8 elements apart in the frame and 4 elements (16 B) apart in each half: coalesced
        const unsigned *q0, *q1;
        if (e + 16 <= (long long)(2 * half_plane)) {
            match_ptrs(copy.data(), e + 8, half_plane, q0, q1);
            CHECK(q0 - p0 == 4 && q1 - p1 == 4, "e=%lld neighbouring lanes are not 16 B apart", e);
        }
    }
}

static void check_plan(int w, int rows, int size_d, int n_views) {
    const Plan p = make_plan(w, rows, size_d, 148, n_views);
    CHECK(p.n_strips * VALID_W >= w && (p.n_strips - 1) * VALID_W < w, "strips %d for w=%d", p.n_strips, w);
    CHECK(p.n_bands * p.band_rows >= rows && (p.n_bands - 1) * p.band_rows < rows, "bands %d x %d for %d rows", p.n_bands,
          p.band_rows, rows);
    CHECK(p.chunk_d % NWARP == 0 && p.n_chunks * p.chunk_d >= size_d && (p.n_chunks - 1) * p.chunk_d < size_d,
          "chunks %d x %d for D=%d", p.n_chunks, p.chunk_d, size_d);
}

int main() {
    for (int w : {2, 33, 216, 217, 384, 470, 1920, 3052, 7680}) check_strip_tiling(w);
    check_deinterleave();
    for (int w : {33, 384, 1920, 7680})
        for (int rows : {1, 25, 288, 1080, 4320})
            for (int d : {1, 3, 16, 70, 256, 512})
                for (int v : {1, 2}) check_plan(w, rows, d, v);
    {
        sb200_params p;
        sb200_default_params(&p);
        int nI = 0, nG = 0, S = 0;
        CHECK(find_lattice(&p, &nI, &nG, &S) && nI == 2 && nG == 9 && S == 20, "default lattice (%d, %d, %d)", nI, nG, S);
        CHECK(pad_x(255) % 8 == 0 && pad_x(255) >= 255 + SW, "pad_x");
    }
    static_assert(HALO >= 2 * RAD && HALO % 4 == 0 && VALID_W == SW - 2 * HALO, "strip geometry");
    printf(fails ? "layout_check: %d FAILED\n" : "layout_check: ok\n", fails);
    return fails ? 1 : 0;
}
