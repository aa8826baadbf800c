// C++ host-mirror check of the fusion pre-step (include/spim_fusion.hpp): raw stacks + registrations + bead locations ->
// ProcessForDeconvolution (device-side transform, blending weights, PSF extraction, weight normalisation) -> deconvolution.
// Inputs and results are dumped so that the Python test can recompute everything with the oracle.
// Linked against the CUDA library on a GPU box or against the CPU emulator in the CPU test suite.
#include "../../include/spim_fusion.hpp"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>

using namespace spim_b200;

static void dump(FILE* f, const Image& im) {
    fwrite(im.dims.data(), sizeof(int), 3, f);
    fwrite(im.data.data(), sizeof(float), im.data.size(), f);
}

int main(int argc, char** argv) {
    if (argc < 2) { fprintf(stderr, "usage: fusion_mirror_test <out.bin>\n"); return 2; }
    const int V = 2, SX = 16, SY = 14, SZ = 8, BX = 18, BY = 14, BZ = 12, NB = 3;
    std::mt19937 rng(4321);
    std::uniform_real_distribution<float> u(10.f, 1000.f);
    std::uniform_real_distribution<double> ub(1.0, 6.0);
    std::vector<Image> stacks;
    std::vector<AffineTransform3D> models;
    std::vector<std::vector<std::array<double, 3>>> beads;
    for (int v = 0; v < V; ++v) {
        Image st(SX, SY, SZ);
        for (auto& t : st.data) t = u(rng);
        stacks.push_back(st);
        const double a = 0.7 * v, c = std::cos(a), s = std::sin(a), zs = 2.0;
        const double cx = (SX - 1) / 2.0, cz = (SZ - 1) / 2.0 * zs;
        std::array<double, 12> m{{c, 0, s * zs, -(c * cx + s * cz) + cx + 1.0, 0, 1, 0, 0.5 * v, -s, 0, c * zs, -(-s * cx + c * cz) + cz - 1.0}};
        models.emplace_back(m);
        std::vector<std::array<double, 3>> b;
        for (int i = 0; i < NB; ++i) b.push_back({{ub(rng) * 2, ub(rng) * 2, ub(rng)}});
        beads.push_back(b);
    }
    FILE* f = fopen(argv[1], "wb");
    if (!f) return 3;
    const int hdr[4] = {V, NB, BX, BY};
    fwrite(hdr, sizeof(int), 4, f);
    const int bz = BZ;
    fwrite(&bz, sizeof(int), 1, f);
    for (int v = 0; v < V; ++v) {
        dump(f, stacks[v]);
        fwrite(models[v].getRowPackedCopy().data(), sizeof(double), 12, f);
        for (auto& b : beads[v]) fwrite(b.data(), sizeof(double), 3, f);
    }
    try {
        ProcessForDeconvolution pfd({{BX, BY, BZ}}, {{-1, 0, 1}}, {{1, 1, 0}}, {{4, 4, 2}}, V, PSFTYPE::INDEPENDENT, 0.006,
                                    /*osemIndex*/ 2, 1.0, /*numThreads*/ 2);
        if (!pfd.fuseStacksAndGetPSFs(stacks, models, WeightType::VIRTUAL_WEIGHTS, {}, beads, {{5, 5, 3}})) return 4;
        dump(f, pfd.getTransformedImg(0));
        dump(f, pfd.getTransformedWeight(1));         // raw blending weight: virtual weights are normalised at init
        dump(f, pfd.getExtractPSF().getTransformedPSF(1));
        dump(f, pfd.getExtractPSF().getInputCalibrationPSF(0));
        const double st[3] = {(double)pfd.getMinOverlappingViews(), pfd.getAvgOverlappingViews(), pfd.getOSEMspeedup()};
        fwrite(st, sizeof(double), 3, f);
        Image psi = pfd.deconvolve(2);
        dump(f, pfd.getTransformedWeight(1));         // now min(1, w / sum * osem)
        dump(f, psi);
        bool threw = false;
        try { AffineTransform3D(std::array<double, 12>{}).inverse(); } catch (const std::runtime_error&) { threw = true; }
        if (!threw) return 6;
    } catch (const std::exception& e) {
        fprintf(stderr, "fusion_mirror_test: %s\n", e.what());
        fclose(f);
        return 1;
    }
    fclose(f);
    printf("FUSION_MIRROR_OK\n");
    return 0;
}
