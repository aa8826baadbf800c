// C++ host-mirror check: builds a small seeded dataset, runs the reference-named classes
// (MVDeconFFT / MVDeconInput / MVDeconvolution and LRFFT / LRInput / BayesMVDeconvolution) over the
// C-ABI and dumps inputs + psi so that the Python test can compare them with the oracle.
// Linked against the CUDA library on a GPU box or against the CPU emulator in the CPU test suite.
#include "../../include/spim_mvdecon.hpp"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>

using namespace spim_b200;

static void dump(FILE* f, const Image& im) { fwrite(im.data.data(), sizeof(float), im.data.size(), f); }

int main(int argc, char** argv) {
    if (argc < 3) { fprintf(stderr, "usage: host_mirror_test <out.bin> <gen 1|2>\n"); return 2; }
    const int gen = atoi(argv[2]);
    const int X = 18, Y = 14, Z = 12, K = 5, V = 3;
    std::mt19937 rng(1234);
    std::uniform_real_distribution<float> u(0.05f, 1.0f);
    std::vector<Image> imgs, ws, psfs;
    for (int v = 0; v < V; ++v) {
        Image im(X, Y, Z), w(X, Y, Z), k(K, K, K);
        for (auto& t : im.data) t = u(rng);
        for (auto& t : w.data) t = u(rng) / V;
        for (int z = 0; z < K; ++z) for (int y = 0; y < K; ++y) for (int x = 0; x < K; ++x) {
            const float dx = x - 2 + 0.3f * v, dy = y - 2, dz = z - 2 - 0.2f * v;
            k.data[((size_t)z * K + y) * K + x] = std::exp(-0.5f * (dx * dx + dy * dy + dz * dz / 2.f));
        }
        // a hole without data in view 0 (img == 0) exercises the gen-2 quotient rule and the final mask
        if (v == 0) for (int i = 0; i < 40; ++i) im.data[i] = 0.f;
        imgs.push_back(im); ws.push_back(w); psfs.push_back(k);
    }
    FILE* f = fopen(argv[1], "wb");
    if (!f) return 3;
    const int hdr[6] = {X, Y, Z, K, V, gen};
    fwrite(hdr, sizeof(int), 6, f);
    for (int v = 0; v < V; ++v) { dump(f, imgs[v]); dump(f, ws[v]); dump(f, psfs[v]); }
    try {
        Image psi;
        double avg = 0;
        size_t nstats = 0;
        if (gen == 2) {
            MVDeconInput views;
            for (int v = 0; v < V; ++v) views.add(std::make_shared<MVDeconFFT>(imgs[v], ws[v], psfs[v]));
            MVDeconvolution d(views, PSFTYPE::EFFICIENT_BAYESIAN, 3, 0.006, 1.0, 0, "deconvolved");
            psi = d.getPsi(); avg = d.getAvg(); nstats = d.getStatistics().size();
            if (d.getCurrentIteration() != 3 || d.getName() != "deconvolved") return 4;
            // per-view operator through the same library
            Image c1 = views.getViews()[1]->convolve1(psi);
            dump(f, psi); dump(f, views.getViews()[1]->getKernel2()); dump(f, c1);
        } else {
            LRInput views;
            for (int v = 0; v < V; ++v) views.add(std::make_shared<LRFFT>(imgs[v], ws[v], psfs[v], std::vector<int>{0}, false, std::array<int, 3>{{0, 0, 0}}));
            BayesMVDeconvolution d(views, PSFTYPE::OPTIMIZATION_I, 3, 0.006, 1.0, 2, "deconvolved");
            psi = d.getPsi(); avg = d.getAvg(); nstats = d.getStatistics().size();
            Image c1 = views.getViews()[1]->convolve1(psi);
            dump(f, psi); dump(f, views.getViews()[1]->getKernel2()); dump(f, c1);
        }
        fwrite(&avg, sizeof(double), 1, f);
        if (nstats != 9) return 5;
        // error behaviour: a CPU device id is rejected like the reference's CUDA-only path
        bool threw = false;
        try { MVDeconFFT bad(imgs[0], ws[0], psfs[0], std::vector<int>{-1}); } catch (const std::invalid_argument&) { threw = true; }
        if (!threw) return 6;
    } catch (const std::exception& e) {
        fprintf(stderr, "host_mirror_test: %s\n", e.what());
        fclose(f);
        return 1;
    }
    fclose(f);
    printf("HOST_MIRROR_OK\n");
    return 0;
}
