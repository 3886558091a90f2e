// Compiles main.cu's call sequence (main.cu:65-155) against include/stereo_b200_compat.hpp: the
// reference's signatures, our library.  Run on a B200: prints the occluded fraction of a tiny pair.
#include <cstdio>
#include <cstring>
#include <vector>

#include "stereo_b200_compat.hpp"

int main() {
    const int w = 96, h = 64, ch = 3, n = w * h;
    std::vector<unsigned char> L(n * ch), R(n * ch);
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++)
            for (int c = 0; c < ch; c++) {
                R[(y * w + x) * ch + c] = (unsigned char)((x * 7 + y * 13 + c * 5) * 31 % 251);
                int xs = x - 4 < 0 ? 0 : x - 4;
                L[(y * w + x) * ch + c] = (unsigned char)((xs * 7 + y * 13 + c * 5) * 31 % 251);
            }
    const sb200_params& p = stereo_b200::params();
    const int size_d = p.dmax - p.dmin + 1;
    unsigned char* I_l = rgb_to_grayscale(L.data(), n, ch, false);
    unsigned char* I_r = rgb_to_grayscale(R.data(), n, ch, false);
    std::vector<float> costl((size_t)n * size_d), costr((size_t)n * size_d);
    compute_cost(I_l, I_r, costl.data(), w, w, h, h, p.dmin, false);
    compute_cost(I_r, I_l, costr.data(), w, w, h, h, -p.dmax, false);
    std::vector<float> bestl(n), bestr(n), dmapl(n, 0.f), dmapr(n, 0.f);
    memset(bestl.data(), 0x7F, n * sizeof(float));  // main.cu:112
    memset(bestr.data(), 0x7F, n * sizeof(float));
    std::vector<unsigned char> mean1(n), mean2(n), c1(n, 0), c2(n, 0);
    compute_guided_filter(I_l, costl.data(), bestl.data(), dmapl.data(), mean1.data(), w, h, size_d, p.dmin, false);
    compute_guided_filter(I_r, costr.data(), bestr.data(), dmapr.data(), mean2.data(), w, h, size_d, -p.dmax, false);
    std::vector<float> occ(dmapl), filled;
    detect_occlusion(occ.data(), dmapr.data(), p.dmin - 100, c1.data(), c2.data(), w, h);
    filled = occ;
    fill_occlusion(filled.data(), w, h, (float)p.dmin);
    int nocc = 0, good = 0;
    for (int i = 0; i < n; i++) {
        nocc += occ[i] == (float)(p.dmin - 100);
        good += dmapl[i] == -4.0f;
    }
    printf("compat smoke: %d of %d occluded, %d labelled -4\n", nocc, n, good);
    free(I_l);
    free(I_r);
    return good > n / 2 ? 0 : 1;
}
