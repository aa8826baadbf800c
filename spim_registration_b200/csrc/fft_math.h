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


// =========================================================================================
// Packed pairs (Blackwell FADD2 / FMUL2 / FFMA2): the TWO columns (or lines) of a work item go through identical
// arithmetic, so their real parts travel in one 64-bit register pair and their imaginary parts in another, and every
// butterfly operation is ONE packed instruction for both columns -- half the floating-point issue slots of the scalar
// form.  Scalar operands (compile-time twiddle constants, per-row twiddles) are broadcast by the instruction itself.
// The host / emulator build evaluates the same expressions component-wise.
// =========================================================================================
struct C2 { float2 re, im; };      // re = (re_a, re_b), im = (im_a, im_b)

#if defined(__CUDA_ARCH__) && !defined(SPIM_HOST_EMU)
SPIM_HD float2 p2add(float2 a, float2 b) { return __fadd2_rn(a, b); }
SPIM_HD float2 p2sub(float2 a, float2 b) { return __fadd2_rn(a, make_float2(-b.x, -b.y)); }
SPIM_HD float2 p2mul(float2 a, float2 b) { return __fmul2_rn(a, b); }
SPIM_HD float2 p2fma(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
#else
SPIM_HD float2 p2add(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
SPIM_HD float2 p2sub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
SPIM_HD float2 p2mul(float2 a, float2 b) { return make_float2(a.x * b.x, a.y * b.y); }
SPIM_HD float2 p2fma(float2 a, float2 b, float2 c) { return make_float2(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y)); }
#endif
SPIM_HD float2 p2neg(float2 a) { return make_float2(-a.x, -a.y); }
SPIM_HD float2 p2bc(float s) { return make_float2(s, s); }       // scalar broadcast (an operand modifier on the GPU)

SPIM_HD C2 cadd(C2 a, C2 b) { C2 r; r.re = p2add(a.re, b.re); r.im = p2add(a.im, b.im); return r; }
SPIM_HD C2 csub(C2 a, C2 b) { C2 r; r.re = p2sub(a.re, b.re); r.im = p2sub(a.im, b.im); return r; }
// a + i b  /  a - i b
SPIM_HD C2 caddi(C2 a, C2 b) { C2 r; r.re = p2sub(a.re, b.im); r.im = p2add(a.im, b.re); return r; }
SPIM_HD C2 csubi(C2 a, C2 b) { C2 r; r.re = p2add(a.re, b.im); r.im = p2sub(a.im, b.re); return r; }
// a * (c + i s) and a * (c - i s) with scalar c, s (broadcast)
SPIM_HD C2 cmul_s(C2 a, float c, float s) {
    C2 r;
    r.re = p2fma(p2neg(a.im), p2bc(s), p2mul(a.re, p2bc(c)));
    r.im = p2fma(a.re, p2bc(s), p2mul(a.im, p2bc(c)));
    return r;
}
SPIM_HD C2 cmulc_s(C2 a, float c, float s) {
    C2 r;
    r.re = p2fma(a.im, p2bc(s), p2mul(a.re, p2bc(c)));
    r.im = p2fma(p2neg(a.re), p2bc(s), p2mul(a.im, p2bc(c)));
    return r;
}
// a * b with b = two different complex numbers (the kernel-spectrum multiply)
SPIM_HD C2 cmul(C2 a, C2 b) {
    C2 r;
    r.re = p2fma(p2neg(a.im), b.im, p2mul(a.re, b.re));
    r.im = p2fma(a.re, b.im, p2mul(a.im, b.re));
    return r;
}
// interleaved float4 (re_a, im_a, re_b, im_b) <-> packed pair
SPIM_HD C2 c2_from_il(float4 v) { C2 r; r.re = make_float2(v.x, v.z); r.im = make_float2(v.y, v.w); return r; }
SPIM_HD float4 c2_to_il(C2 a) { return make_float4(a.re.x, a.im.x, a.re.y, a.im.y); }
// packed float4 (re_a, re_b, im_a, im_b): the layout of shared-memory tiles between stages
SPIM_HD C2 c2_from_pk(float4 v) { C2 r; r.re = make_float2(v.x, v.y); r.im = make_float2(v.z, v.w); return r; }
SPIM_HD float4 c2_to_pk(C2 a) { return make_float4(a.re.x, a.re.y, a.im.x, a.im.y); }
SPIM_HD C2 c2_from_ab(float2 a, float2 b) { C2 r; r.re = make_float2(a.x, b.x); r.im = make_float2(a.y, b.y); return r; }
SPIM_HD float2 c2_a(C2 v) { return make_float2(v.re.x, v.im.x); }
SPIM_HD float2 c2_b(C2 v) { return make_float2(v.re.y, v.im.y); }

template <bool INV> SPIM_HD C2 rot90(C2 a) {        // multiply by -i (forward) or +i (inverse)
    C2 r;
    if (INV) { r.re = p2neg(a.im); r.im = a.re; } else { r.re = a.im; r.im = p2neg(a.re); }
    return r;
}
template <int NUM, int DEN, bool INV> SPIM_HD C2 ctw(C2 a) {
    constexpr int n = ((NUM % DEN) + DEN) % DEN;
    if constexpr (n == 0) {
        return a;
    } else if constexpr (4 * n == DEN) {
        return rot90<INV>(a);
    } else if constexpr (2 * n == DEN) {
        C2 r; r.re = p2neg(a.re); r.im = p2neg(a.im); return r;
    } else if constexpr (4 * n == 3 * DEN) {
        return rot90<!INV>(a);
    } else {
        constexpr float c = Tw<n, DEN>::c;
        constexpr float s = INV ? -Tw<n, DEN>::s : Tw<n, DEN>::s;  // w = c - i*s (fwd)
        return cmulc_s(a, c, s);
    }
}

template <int R, bool INV> struct Dft2;
template <int R, bool INV> SPIM_HD void dft(C2 (&a)[R]) { Dft2<R, INV>::run(a); }

template <bool INV> struct Dft2<1, INV> { SPIM_HD static void run(C2 (&)[1]) {} };
template <bool INV> struct Dft2<2, INV> {
    SPIM_HD static void run(C2 (&a)[2]) {
        C2 t = a[0];
        a[0] = cadd(t, a[1]);
        a[1] = csub(t, a[1]);
    }
};
template <bool INV> struct Dft2<4, INV> {
    SPIM_HD static void run(C2 (&a)[4]) {
        C2 t0 = cadd(a[0], a[2]), t1 = csub(a[0], a[2]);
        C2 t2 = cadd(a[1], a[3]), d = csub(a[1], a[3]);
        a[0] = cadd(t0, t2);
        a[2] = csub(t0, t2);
        // t1 -+ i d (forward: a[1] = t1 - i d)
        a[1] = INV ? caddi(t1, d) : csubi(t1, d);
        a[3] = INV ? csubi(t1, d) : caddi(t1, d);
    }
};
template <int P, bool INV> struct Dft2Prime {
    SPIM_HD static void run(C2 (&a)[P]) {
        constexpr int H = (P - 1) / 2;
        C2 s[H], d[H];
        static_for<H>([&](auto qi) {
            constexpr int q = decltype(qi)::value + 1;
            s[q - 1] = cadd(a[q], a[P - q]);
            d[q - 1] = csub(a[q], a[P - q]);
        });
        const C2 a0 = a[0];
        C2 x0 = a0;
        static_for<H>([&](auto qi) { x0 = cadd(x0, s[decltype(qi)::value]); });
        a[0] = x0;
        static_for<H>([&](auto ki) {
            constexpr int k = decltype(ki)::value + 1;
            C2 t = a0, u;
            static_for<H>([&](auto qi) {
                constexpr int q = decltype(qi)::value + 1;
                constexpr float c = Tw<(k * q) % P, P>::c;
                constexpr float sn = Tw<(k * q) % P, P>::s;
                t.re = p2fma(s[q - 1].re, p2bc(c), t.re);
                t.im = p2fma(s[q - 1].im, p2bc(c), t.im);
                if constexpr (q == 1) { u.re = p2mul(d[0].re, p2bc(sn)); u.im = p2mul(d[0].im, p2bc(sn)); }
                else { u.re = p2fma(d[q - 1].re, p2bc(sn), u.re); u.im = p2fma(d[q - 1].im, p2bc(sn), u.im); }
            });
            // forward: X_k = t - i u ; X_{P-k} = t + i u   (inverse swaps the two)
            const C2 lo = csubi(t, u), hi = caddi(t, u);
            a[k] = INV ? hi : lo;
            a[P - k] = INV ? lo : hi;
        });
    }
};
template <int R, bool INV> struct Dft2Composite {
    SPIM_HD static void run(C2 (&a)[R]) {
        constexpr int R1 = split_factor(R);
        constexpr int R2 = R / R1;
        C2 y[R];
        static_for<R2>([&](auto n2i) {
            constexpr int n2 = decltype(n2i)::value;
            C2 t[R1];
            static_for<R1>([&](auto n1i) { constexpr int n1 = decltype(n1i)::value; t[n1] = a[R2 * n1 + n2]; });
            dft<R1, INV>(t);
            static_for<R1>([&](auto k1i) {
                constexpr int k1 = decltype(k1i)::value;
                y[k1 * R2 + n2] = ctw<n2 * k1, R, INV>(t[k1]);
            });
        });
        static_for<R1>([&](auto k1i) {
            constexpr int k1 = decltype(k1i)::value;
            C2 u[R2];
            static_for<R2>([&](auto n2i) { constexpr int n2 = decltype(n2i)::value; u[n2] = y[k1 * R2 + n2]; });
            dft<R2, INV>(u);
            static_for<R2>([&](auto k2i) { constexpr int k2 = decltype(k2i)::value; a[k1 + R1 * k2] = u[k2]; });
        });
    }
};
template <int R, bool INV> struct Dft2PFA {
    SPIM_HD static void run(C2 (&a)[R]) {
        constexpr int R1 = pfa_factor(R);
        constexpr int R2 = R / R1;
        constexpr int E1 = R2 * cmodinv(R2 % R1, R1);
        constexpr int E2 = R1 * cmodinv(R1 % R2, R2);
        C2 y[R];
        static_for<R2>([&](auto n2i) {
            constexpr int n2 = decltype(n2i)::value;
            C2 t[R1];
            static_for<R1>([&](auto n1i) { constexpr int n1 = decltype(n1i)::value; t[n1] = a[(R2 * n1 + R1 * n2) % R]; });
            dft<R1, INV>(t);
            static_for<R1>([&](auto k1i) { constexpr int k1 = decltype(k1i)::value; y[k1 * R2 + n2] = t[k1]; });
        });
        static_for<R1>([&](auto k1i) {
            constexpr int k1 = decltype(k1i)::value;
            C2 u[R2];
            static_for<R2>([&](auto n2i) { constexpr int n2 = decltype(n2i)::value; u[n2] = y[k1 * R2 + n2]; });
            dft<R2, INV>(u);
            static_for<R2>([&](auto k2i) { constexpr int k2 = decltype(k2i)::value; a[(k1 * E1 + k2 * E2) % R] = u[k2]; });
        });
    }
};
template <int R, bool INV> struct Dft2 {
    SPIM_HD static void run(C2 (&a)[R]) {
        if constexpr (is_prime(R)) Dft2Prime<R, INV>::run(a);
        else if constexpr (pfa_factor(R) != 0) Dft2PFA<R, INV>::run(a);
        else Dft2Composite<R, INV>::run(a);
    }
};

}  // namespace spim
