// Native driver for ThreadSanitizer runs of the FFT-convolution kernels: with SPIM_EMU_THREADS=T the emulator executes every
// block of the x-forward, column and x-inverse kernels as T real threads that split the work items like the threads of a
// CUDA block and meet at real barriers (csrc/hd.h), so a missing or misplaced barrier -- one thread reading a tile element
// another one has not written yet -- is a data race ThreadSanitizer reports.  Compiled together with csrc/spim_b200.cu
// (-DSPIM_HOST_EMU -fsanitize=thread) by tests/test_tsan_kernels.py, once as is and once with -DSPIM_EMU_NO_STAGE_BARRIER
// (negative control).  Results of the threaded run must equal the single-thread run bit for bit.
#include "../../include/spim_mvdecon.h"
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

static unsigned g_seed = 99u;
static float rnd() { g_seed = g_seed * 1664525u + 1013904223u; return (float)(g_seed >> 8) * (1.0f / 16777216.0f); }

static int conv(const int n[3], const int k[3], int ext, std::vector<float>& out) {
    std::vector<float> img((size_t)n[0] * n[1] * n[2]), ker((size_t)k[0] * k[1] * k[2]);
    g_seed = 7u + 13u * (unsigned)(n[0] + 3 * n[1] + 5 * n[2] + ext);
    for (auto& v : img) v = rnd();
    for (auto& v : ker) v = rnd();
    out.assign(img.size(), 0.f);
    return mvd_convolve(img.data(), n, ker.data(), k, ext, 1.0f, out.data(), 0);
}

// mode 0: defaults (rows of 16 voxels: the TMA-fed x-forward kernel with its fix-up phases, de-duplicated mirror halos,
// conv2's constant halo by shift); 1: narrow column tiles; 2: 8-line x-forward tiles; 3: full forward sweeps and the
// literal constant extension; 4: rows that are not 16-byte aligned (the plain-load x-forward kernel)
static const char* kSwitches[] = {"SPIM_COL_NARROW", "SPIM_XFWD_LINES", "SPIM_DEDUP", "SPIM_CONST_SHIFT"};
static int decon(int mode, std::vector<float>& psi) {
    const int n[3] = {10, 12, mode == 4 ? 14 : 16}, kd[3] = {5, 5, 5}, V = 2;
    for (const char* e : kSwitches) unsetenv(e);
    if (mode == 1) setenv("SPIM_COL_NARROW", "1", 1);
    if (mode == 2) setenv("SPIM_XFWD_LINES", "8", 1);
    if (mode == 3) { setenv("SPIM_DEDUP", "0", 1); setenv("SPIM_CONST_SHIFT", "0", 1); }
    mvd_params p;
    mvd_params_default(&p);
    for (int d = 0; d < 3; ++d) p.dims[d] = n[d];
    p.num_views = V; p.iteration_type = MVD_EFFICIENT_BAYESIAN; p.generation = 2;
    mvd_session* s = nullptr;
    if (mvd_session_create(&p, &s)) return 1;
    const size_t N = (size_t)n[0] * n[1] * n[2];
    std::vector<float> img(N), w(N, 0.5f), psf(125);
    g_seed = 4242u;
    for (auto& v : psf) v = 0.1f + rnd();
    int rc = 0;
    for (int v = 0; v < V; ++v) {
        for (auto& x : img) x = 0.05f + 0.95f * rnd();
        rc |= mvd_set_view(s, v, img.data(), w.data(), psf.data(), kd);
    }
    rc |= mvd_init(s);
    rc |= mvd_run(s, 2, nullptr, nullptr);
    rc |= mvd_finish(s);
    psi.assign(N, 0.f);
    rc |= mvd_get_psi(s, psi.data());
    mvd_session_destroy(s);
    return rc;
}

int main() {
    const int shapes[][6] = {{9, 7, 11, 3, 5, 3}, {5, 30, 33, 1, 7, 9}, {6, 6, 6, 4, 2, 6}, {12, 20, 18, 3, 7, 5}, {3, 40, 6, 3, 9, 3},
                             {10, 12, 16, 5, 5, 5}, {7, 9, 24, 3, 5, 7}};     // the last two: 16-byte aligned rows (TMA-fed x-forward)
    int fail = 0;
    for (int pass = 0; pass < 2; ++pass) {
        static std::vector<std::vector<float>> ref;
        size_t idx = 0;
        if (pass == 0) unsetenv("SPIM_EMU_THREADS"); else setenv("SPIM_EMU_THREADS", "4", 1);
        for (auto& sh : shapes)
            for (int ext = 0; ext < 5; ++ext) {
                std::vector<float> out;
                if (conv(sh, sh + 3, ext, out)) { fprintf(stderr, "mvd_convolve failed: %s\n", mvd_last_error()); fail = 1; }
                if (pass == 0) ref.push_back(out);
                else if (memcmp(ref[idx].data(), out.data(), out.size() * 4)) { fprintf(stderr, "threaded result differs (conv %zu)\n", idx); fail = 1; }
                ++idx;
            }
        for (int mode = 0; mode < 5; ++mode) {
            std::vector<float> psi;
            if (decon(mode, psi)) { fprintf(stderr, "deconvolution failed: %s\n", mvd_last_error()); fail = 1; }
            if (pass == 0) ref.push_back(psi);
            else if (memcmp(ref[idx].data(), psi.data(), psi.size() * 4)) { fprintf(stderr, "threaded result differs (decon mode %d)\n", mode); fail = 1; }
            ++idx;
        }
        for (const char* e : kSwitches) unsetenv(e);
    }
    printf(fail ? "KERNEL_DRIVER_FAILED\n" : "KERNEL_DRIVER_OK\n");
    return fail;
}
