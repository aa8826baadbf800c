// CPU check of csrc/fast_math.h: with seeds anywhere inside the hardware approximations' error bounds with a margin (rcp.approx:
// 1 ulp, checked to 2; rsqrt.approx: 2 ulp, checked to 3) the refined results equal the correctly rounded IEEE quotient / square root.
//   sqrt: EVERY float in [1, 4) (both exponent parities; other binades differ by an exact power of 4)
//   div : random operand pairs over the ranges the epilogue sees and far beyond (quotients kept in the normal range)
#include "../../spim_registration_b200/csrc/fast_math.h"
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>

static float nudge(float v, int ulps) {
    int32_t b;
    memcpy(&b, &v, 4);
    b += ulps;
    memcpy(&v, &b, 4);
    return v;
}

int main(int argc, char** argv) {
    const long ndiv = argc > 1 ? atol(argv[1]) : 20000000L;
    long bad_sqrt = 0, bad_div = 0, nsqrt = 0;
    for (uint32_t bits = 0x3f800000u; bits < 0x40800000u; ++bits) {
        float x;
        memcpy(&x, &bits, 4);
        const float want = sqrtf(x);
        const float y = (float)(1.0 / sqrt((double)x));
        for (int u = -3; u <= 3; ++u) {
            if (spim_sqrt_from_seed(x, nudge(y, u)) != want) ++bad_sqrt;
            ++nsqrt;
        }
    }
    std::mt19937_64 rng(12345);
    std::uniform_real_distribution<double> ex(-30.0, 30.0), man(1.0, 2.0);
    for (long i = 0; i < ndiv; ++i) {
        const float a = (float)(man(rng) * exp2(floor(ex(rng))));
        const float b = (float)(man(rng) * exp2(floor(ex(rng))));
        const float want = a / b;
        const float r = (float)(1.0 / (double)b);
        for (int u = -2; u <= 2; ++u)
            if (spim_div_from_seed(a, b, nudge(r, u)) != want) ++bad_div;
    }
    printf("sqrt: %ld mismatches of %ld; div: %ld mismatches of %ld\n", bad_sqrt, nsqrt, bad_div, 5 * ndiv);
    printf(bad_sqrt == 0 && bad_div == 0 ? "FAST_MATH_OK\n" : "FAST_MATH_FAIL\n");
    return bad_sqrt == 0 && bad_div == 0 ? 0 : 1;
}
