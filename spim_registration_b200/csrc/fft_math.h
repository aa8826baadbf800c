// In-register radix butterflies for the mixed-radix FFT stages.
// All twiddle constants are compile-time (constexpr Taylor evaluation in double), so every
// dft<R> below compiles to straight-line FADD/FMUL/FFMA code with immediate operands.
#pragma once
#include "hd.h"
#include <utility>
#include <type_traits>

namespace spim {

// ---- constexpr sin/cos of 2*pi*num/den -------------------------------------------------
constexpr double kPi = 3.14159265358979323846264338327950288;

constexpr double c_sin_small(double a) {  // |a| <= pi/2
    double a2 = a * a, term = a, sum = a;
    for (int i = 1; i < 16; ++i) { term *= -a2 / double((2 * i) * (2 * i + 1)); sum += term; }
    return sum;
}
constexpr double c_cos_small(double a) {
    double a2 = a * a, term = 1.0, sum = 1.0;
    for (int i = 1; i < 16; ++i) { term *= -a2 / double((2 * i - 1) * (2 * i)); sum += term; }
    return sum;
}
// returns cos (want_cos) or sin of 2*pi*num/den with exact values on the axes
constexpr double c_sincos_2pi(long long num, long long den, bool want_cos) {
    num %= den;
    if (num < 0) num += den;
    long long q = (4 * num) / den;
    long long r = 4 * num - q * den;
    double a = (kPi / 2.0) * (double(r) / double(den));
    double s = (r == 0) ? 0.0 : c_sin_small(a);
    double c = (r == 0) ? 1.0 : c_cos_small(a);
    double S = 0, C = 0;
    switch (q) {
        case 0: S = s;  C = c;  break;
        case 1: S = c;  C = -s; break;
        case 2: S = -s; C = -c; break;
        default: S = -c; C = s; break;
    }
    return want_cos ? C : S;
}

template <int NUM, int DEN>
struct Tw {  // e^{-2*pi*i*NUM/DEN} = (c, -s)
    static constexpr float c = (float)c_sincos_2pi(NUM, DEN, true);
    static constexpr float s = (float)c_sincos_2pi(NUM, DEN, false);
};

// ---- static_for -----------------------------------------------------------------------
template <class F, int... I>
SPIM_HD void static_for_impl(F&& f, std::integer_sequence<int, I...>) {
    (f(std::integral_constant<int, I>{}), ...);
}
template <int N, class F>
SPIM_HD void static_for(F&& f) {
    static_for_impl(static_cast<F&&>(f), std::make_integer_sequence<int, N>{});
}

// ---- complex helpers ------------------------------------------------------------------
SPIM_HD float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
SPIM_HD float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
SPIM_HD float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
SPIM_HD float2 cmulc(float2 a, float2 b) {  // a * conj(b)
    return make_float2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y);
}
SPIM_HD float2 cconj(float2 a) { return make_float2(a.x, -a.y); }
// multiply by -i (forward quarter turn) or +i (inverse)
template <bool INV> SPIM_HD float2 rot90(float2 a) {
    return INV ? make_float2(-a.y, a.x) : make_float2(a.y, -a.x);
}
// multiply by e^{-+2*pi*i*NUM/DEN} (sign by INV) with compile-time constants
template <int NUM, int DEN, bool INV> SPIM_HD float2 ctw(float2 a) {
    constexpr int n = ((NUM % DEN) + DEN) % DEN;
    if constexpr (n == 0) {
        return a;
    } else if constexpr (4 * n == DEN) {
        return rot90<INV>(a);
    } else if constexpr (2 * n == DEN) {
        return make_float2(-a.x, -a.y);
    } else if constexpr (4 * n == 3 * DEN) {
        return rot90<!INV>(a);
    } else {
        constexpr float c = Tw<n, DEN>::c;
        constexpr float s = INV ? -Tw<n, DEN>::s : Tw<n, DEN>::s;  // w = c - i*s (fwd)
        return make_float2(a.x * c + a.y * s, a.y * c - a.x * s);
    }
}

constexpr bool is_prime(int n) {
    if (n < 2) return false;
    for (int i = 2; i * i <= n; ++i) if (n % i == 0) return false;
    return true;
}
// first factor used to split a composite radix
constexpr int split_factor(int r) {
    if (r == 16) return 4;
    if (r == 8) return 2;
    if (r == 12) return 4;
    for (int i = 2; i * i <= r; ++i) if (r % i == 0) return i;
    return r;
}

template <int R, bool INV> struct Dft;

template <int R, bool INV> SPIM_HD void dft(float2 (&a)[R]) { Dft<R, INV>::run(a); }

template <bool INV> struct Dft<1, INV> { SPIM_HD static void run(float2 (&)[1]) {} };

template <bool INV> struct Dft<2, INV> {
    SPIM_HD static void run(float2 (&a)[2]) {
        float2 t = a[0];
        a[0] = cadd(t, a[1]);
        a[1] = csub(t, a[1]);
    }
};

template <bool INV> struct Dft<4, INV> {
    SPIM_HD static void run(float2 (&a)[4]) {
        float2 t0 = cadd(a[0], a[2]), t1 = csub(a[0], a[2]);
        float2 t2 = cadd(a[1], a[3]), t3 = rot90<INV>(csub(a[1], a[3]));
        a[0] = cadd(t0, t2);
        a[2] = csub(t0, t2);
        a[1] = cadd(t1, t3);
        a[3] = csub(t1, t3);
    }
};

// odd prime radix, symmetric form:  X_k = a0 + sum_q cos(kq) s_q  -+ i sum_q sin(kq) d_q
template <int P, bool INV> struct DftPrime {
    SPIM_HD static void run(float2 (&a)[P]) {
        constexpr int H = (P - 1) / 2;
        float2 s[H], d[H];
        static_for<H>([&](auto qi) {
            constexpr int q = decltype(qi)::value + 1;
            s[q - 1] = cadd(a[q], a[P - q]);
            d[q - 1] = csub(a[q], a[P - q]);
        });
        float2 a0 = a[0];
        float2 x0 = a0;
        static_for<H>([&](auto qi) { x0 = cadd(x0, s[decltype(qi)::value]); });
        a[0] = x0;
        static_for<H>([&](auto ki) {
            constexpr int k = decltype(ki)::value + 1;
            float2 t = a0;
            float2 u = make_float2(0.f, 0.f);
            static_for<H>([&](auto qi) {
                constexpr int q = decltype(qi)::value + 1;
                constexpr float c = Tw<(k * q) % P, P>::c;
                constexpr float sn = Tw<(k * q) % P, P>::s;
                t.x += c * s[q - 1].x;
                t.y += c * s[q - 1].y;
                u.x += sn * d[q - 1].x;
                u.y += sn * d[q - 1].y;
            });
            // forward: X_k = t - i u ; X_{P-k} = t + i u   (inverse swaps the two)
            float2 lo = make_float2(t.x + u.y, t.y - u.x);
            float2 hi = make_float2(t.x - u.y, t.y + u.x);
            a[k] = INV ? hi : lo;
            a[P - k] = INV ? lo : hi;
        });
    }
};

// composite radix R = R1*R2 (Cooley-Tukey in registers, natural-order output)
template <int R, bool INV> struct DftComposite {
    SPIM_HD static void run(float2 (&a)[R]) {
        constexpr int R1 = split_factor(R);
        constexpr int R2 = R / R1;
        float2 y[R];
        static_for<R2>([&](auto n2i) {
            constexpr int n2 = decltype(n2i)::value;
            float2 t[R1];
            static_for<R1>([&](auto n1i) { constexpr int n1 = decltype(n1i)::value; t[n1] = a[R2 * n1 + n2]; });
            dft<R1, INV>(t);
            static_for<R1>([&](auto k1i) {
                constexpr int k1 = decltype(k1i)::value;
                y[k1 * R2 + n2] = ctw<n2 * k1, R, INV>(t[k1]);
            });
        });
        static_for<R1>([&](auto k1i) {
            constexpr int k1 = decltype(k1i)::value;
            float2 u[R2];
            static_for<R2>([&](auto n2i) { constexpr int n2 = decltype(n2i)::value; u[n2] = y[k1 * R2 + n2]; });
            dft<R2, INV>(u);
            static_for<R2>([&](auto k2i) { constexpr int k2 = decltype(k2i)::value; a[k1 + R1 * k2] = u[k2]; });
        });
    }
};

constexpr int cgcd(int a, int b) { return b == 0 ? a : cgcd(b, a % b); }
constexpr int cmodinv(int a, int m) {      // a^-1 mod m (m small)
    for (int x = 1; x < m; ++x) if ((a * x) % m == 1) return x;
    return 1;
}
// coprime factors for the prime-factor (Good-Thomas) butterfly, 0 if the radix has none
constexpr int pfa_factor(int r) {
    if (r == 6) return 2;
    if (r == 10) return 2;
    if (r == 12) return 4;
    if (r == 14) return 2;
    if (r == 15) return 3;
    return 0;
}

// composite radix with coprime factors R = R1*R2: Good-Thomas index maps make the two small DFTs
// independent -- no internal twiddle multiplications, the permutations are register renames.
//   input  n = (R2*n1 + R1*n2) mod R,  output k = (k1*R2*(R2^-1 mod R1) + k2*R1*(R1^-1 mod R2)) mod R
template <int R, bool INV> struct DftPFA {
    SPIM_HD static void run(float2 (&a)[R]) {
        constexpr int R1 = pfa_factor(R);
        constexpr int R2 = R / R1;
        static_assert(cgcd(R1, R2) == 1, "PFA needs coprime factors");
        constexpr int E1 = R2 * cmodinv(R2 % R1, R1);   // == 1 mod R1, == 0 mod R2
        constexpr int E2 = R1 * cmodinv(R1 % R2, R2);   // == 0 mod R1, == 1 mod R2
        float2 y[R];
        static_for<R2>([&](auto n2i) {
            constexpr int n2 = decltype(n2i)::value;
            float2 t[R1];
            static_for<R1>([&](auto n1i) { constexpr int n1 = decltype(n1i)::value; t[n1] = a[(R2 * n1 + R1 * n2) % R]; });
            dft<R1, INV>(t);
            static_for<R1>([&](auto k1i) { constexpr int k1 = decltype(k1i)::value; y[k1 * R2 + n2] = t[k1]; });
        });
        static_for<R1>([&](auto k1i) {
            constexpr int k1 = decltype(k1i)::value;
            float2 u[R2];
            static_for<R2>([&](auto n2i) { constexpr int n2 = decltype(n2i)::value; u[n2] = y[k1 * R2 + n2]; });
            dft<R2, INV>(u);
            static_for<R2>([&](auto k2i) { constexpr int k2 = decltype(k2i)::value; a[(k1 * E1 + k2 * E2) % R] = u[k2]; });
        });
    }
};

template <int R, bool INV> struct Dft {
    SPIM_HD static void run(float2 (&a)[R]) {
        if constexpr (is_prime(R)) DftPrime<R, INV>::run(a);
        else if constexpr (pfa_factor(R) != 0) DftPFA<R, INV>::run(a);
        else DftComposite<R, INV>::run(a);
    }
};

}  // namespace spim
