// CPU check behind the default-on fast epilogue (csrc/fast_math.h, kernels.h next_value / epi_one):
// the WHOLE update step  next = max(minValue, value > 0 ? (lambda > 0 ? 2v / (1 + sqrt(1 + 2 lambda v)) : v) : minValue)
// evaluated with the branch-free MUFU-seeded division / square root equals the IEEE evaluation for EVERY float `value`
// below 2^126 in magnitude -- zero, denormal, negative and NaN inputs included, because whatever the refinement produces
// outside the normal range (NaN or a tiny number) is removed by the same select / clamp that follows in the reference's
// computeNextValue (FD/MVDeconvolution.java:692-724).  The hardware approximations are modelled with their flush-to-zero
// behaviour and with seeds displaced by up to 2 (rcp) / 3 (rsqrt) ulp, beyond the documented error bounds.
// The ratio step q = img / blur is checked for img in {0} U [1e-4, 1] against blur over the normal range.
//   usage: fast_epilogue_composite [stride]   (stride 1 = exhaustive over all 2^32 bit patterns; default 1)
#include "../../spim_registration_b200/csrc/fast_math.h"
#include <atomic>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

static float from_bits(uint32_t b) { float f; memcpy(&f, &b, 4); return f; }
static uint32_t bits_of(float f) { uint32_t b; memcpy(&b, &f, 4); return b; }
static float nudge(float v, int ulps) {
    if (!(fabsf(v) > 0.f) || isinf(v) || isnan(v)) return v;
    return from_bits(bits_of(v) + (uint32_t)ulps);
}
static float ftz(float v) { return (fabsf(v) < 1.17549435e-38f) ? copysignf(0.f, v) : v; }
// rcp.approx.ftz.f32 / rsqrt.approx.ftz.f32 with a displaced result
static float rcp_seed(float b, int u) { return ftz(nudge((float)(1.0 / (double)ftz(b)), u)); }
static float rsqrt_seed(float x, int u) { return ftz(nudge((float)(1.0 / sqrt((double)ftz(x))), u)); }

static float fmax_nan(float a, float b) { return fmaxf(a, b); }   // fmaxf returns the non-NaN operand, like the device's

static float next_ieee(float value, float two_lambda, bool lam_pos, float min_value) {
    volatile float den = 1.f + sqrtf(fmaf(two_lambda, value, 1.f));
    volatile float tik = (value + value) / den;
    float adj = lam_pos ? (float)tik : value;
    adj = (value > 0.f) ? adj : min_value;
    return fmax_nan(min_value, adj);
}
static float next_fast(float value, float two_lambda, bool lam_pos, float min_value, int us, int ur) {
    const float x = fmaf(two_lambda, value, 1.f);
    const float den = 1.f + spim_sqrt_from_seed(x, rsqrt_seed(x, us));
    const float tik = spim_div_from_seed(value + value, den, rcp_seed(den, ur));
    float adj = lam_pos ? tik : value;
    adj = (value > 0.f) ? adj : min_value;
    return fmax_nan(min_value, adj);
}
static bool same(float a, float b) { return bits_of(a) == bits_of(b) || (isnan(a) && isnan(b)); }

int main(int argc, char** argv) {
    const uint64_t stride = argc > 1 ? strtoull(argv[1], nullptr, 10) : 1;
    const float lambdas[] = {0.006f, 0.0006f, 0.06f, 0.5f, 0.f};
    const float min_value = 1e-4f;
    const unsigned nthreads = std::max(1u, std::thread::hardware_concurrency());
    std::atomic<uint64_t> bad_update(0), bad_big(0), n_update(0);
    std::vector<std::thread> pool;
    for (unsigned t = 0; t < nthreads; ++t)
        pool.emplace_back([&, t]() {
            uint64_t bad = 0, big = 0, cnt = 0;
            for (uint64_t b = (uint64_t)t * stride; b < (1ull << 32); b += (uint64_t)nthreads * stride) {
                const float v = from_bits((uint32_t)b);
                // one displaced-seed combination per value, cycling through all 7 x 5 of them
                const int us = (int)(b % 7) - 3, ur = (int)((b / 7) % 5) - 2;
                for (float lam : lambdas) {
                    const float tl = (float)(2.0 * (double)lam);
                    const bool lp = lam > 0.f;
                    const bool ok = same(next_ieee(v, tl, lp, min_value), next_fast(v, tl, lp, min_value, us, ur));
                    if (!ok) { if (fabsf(v) < 8.5070592e37f) ++bad; else ++big; }
                    ++cnt;
                }
            }
            bad_update += bad; bad_big += big; n_update += cnt;
        });
    for (auto& th : pool) th.join();
    // ratio: img in {0} U [1e-4, 1] (the loaders' min-max normalised range), blur over 2^-100 .. 2^100 and negative
    uint64_t bad_ratio = 0, n_ratio = 0;
    uint64_t s = 88172645463325252ull;
    auto rnd = [&]() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; };
    for (long i = 0; i < 40000000L; ++i) {
        const float a = (i % 16 == 0) ? 0.f : (1e-4f + (float)((rnd() >> 40) * (1.0 / 16777216.0)) * (1.f - 1e-4f));
        uint32_t bb = (uint32_t)(rnd() >> 32);
        const uint32_t e = 27 + (bb >> 23) % 200;                 // biased exponent 27..226
        bb = (bb & 0x807fffffu) | (e << 23);
        const float b = from_bits(bb);
        const float want = a / b;
        if (want != 0.f && fabsf(want) < 1.17549435e-38f) continue;   // denormal quotients: blur > 1e33, never produced
        const int ur = (int)(i % 5) - 2;
        if (!same(spim_div_from_seed(a, b, rcp_seed(b, ur)), want)) ++bad_ratio;
        ++n_ratio;
    }
    printf("update: %llu mismatches of %llu (|value| < 2^126), %llu beyond; ratio: %llu mismatches of %llu\n",
           (unsigned long long)bad_update.load(), (unsigned long long)n_update.load(), (unsigned long long)bad_big.load(),
           (unsigned long long)bad_ratio, (unsigned long long)n_ratio);
    const bool ok = bad_update.load() == 0 && bad_ratio == 0;
    printf(ok ? "FAST_EPILOGUE_OK\n" : "FAST_EPILOGUE_FAIL\n");
    return ok ? 0 : 1;
}
