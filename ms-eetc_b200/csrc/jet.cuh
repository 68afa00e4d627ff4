// Second-order forward-mode jets in two variables (value, gradient, symmetric Hessian).
// Used by the shooting-interval device functions to get exact first and second sensitivities of
// the RK4 step (replaces CasADi's AD of train.py:294-344 inside IPOPT callbacks, ocp.py:290).
#pragma once
#include <math.h>

#ifndef MS_FAST_RCP
#define MS_FAST_RCP 1
#endif
#if defined(__CUDACC__)
#define MS_HD __host__ __device__ __forceinline__
#else
#define MS_HD inline
#endif

namespace mseetc {

// correctly rounded reciprocal: __drcp_rn on the device (a handful of instructions instead of the full division
// sequence), 1.0/x on the host -- bitwise identical results
MS_HD double rcp(double x) {
#if defined(__CUDA_ARCH__)
    return __drcp_rn(x);
#else
    return 1.0 / x;
#endif
}

// reciprocal of a slack (a positive, normal number): hardware approximation (2^-23) + two Newton steps, 1-2 ulp -- 5 instructions
// against ~12 of the correctly rounded one; used where a last-bit difference is harmless (barrier terms z/s, mu/s: 38 of them
// per cell and iteration)
MS_HD double rcp_slack(double x) {
#if defined(__CUDA_ARCH__) && MS_FAST_RCP
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    r = fma(fma(-x, r, 1.0), r, r);
    r = fma(fma(-x, r, 1.0), r, r);
    return r;
#else
    return rcp(x);
#endif
}

// square root and its reciprocal together: hardware rsqrt approximation + two Newton steps + one correction of the root, 1-2 ulp;
// ~10 instructions against ~30 for the correctly rounded sqrt followed by the correctly rounded reciprocal.  Explicit fma / mul
// only, so that every kernel that inlines it produces the same bits.
MS_HD void sqrt_inv(double x, double& s, double& inv) {
#if defined(__CUDA_ARCH__) && MS_FAST_RCP
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double h = __dmul_rn(0.5, x);
    y = __dmul_rn(y, fma(__dmul_rn(-h, y), y, 1.5));
    y = __dmul_rn(y, fma(__dmul_rn(-h, y), y, 1.5));
    double r = __dmul_rn(x, y);
    r = fma(fma(-r, r, x), __dmul_rn(0.5, y), r);
    s = r; inv = y;
#else
    s = sqrt(x); inv = rcp(s);
#endif
}
MS_HD double fsqrt(double x) { double s, i; sqrt_inv(x, s, i); return s; }

// individually rounded operations that the compiler may not contract into an FMA: used where two kernels must reproduce the
// same value bit for bit (the residual d(x) - w of an active row is a difference of nearly equal numbers that is later
// multiplied by z/s ~ 1e12; a last-bit difference between the kernel that condenses it and the one that recovers the step
// would show up in the multipliers)
MS_HD double mul_rn(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dmul_rn(a, b);
#else
    return a * b;
#endif
}
MS_HD double add_rn(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dadd_rn(a, b);
#else
    return a + b;
#endif
}
MS_HD double sub_rn(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dsub_rn(a, b);
#else
    return a - b;
#endif
}

// variables: x0 (= b at interval start), x1 (= total specific force F)
struct Jet2 {
    double v, g0, g1, h00, h01, h11;
};

MS_HD Jet2 jconst(double c) { return Jet2{c, 0, 0, 0, 0, 0}; }
MS_HD Jet2 jvar0(double x) { return Jet2{x, 1, 0, 0, 0, 0}; }
MS_HD Jet2 jvar1(double x) { return Jet2{x, 0, 1, 0, 0, 0}; }

MS_HD Jet2 operator+(const Jet2& a, const Jet2& b) {
    return Jet2{a.v + b.v, a.g0 + b.g0, a.g1 + b.g1, a.h00 + b.h00, a.h01 + b.h01, a.h11 + b.h11};
}
MS_HD Jet2 operator-(const Jet2& a, const Jet2& b) {
    return Jet2{a.v - b.v, a.g0 - b.g0, a.g1 - b.g1, a.h00 - b.h00, a.h01 - b.h01, a.h11 - b.h11};
}
MS_HD Jet2 operator+(const Jet2& a, double c) { Jet2 r = a; r.v += c; return r; }
MS_HD Jet2 operator-(const Jet2& a, double c) { Jet2 r = a; r.v -= c; return r; }
MS_HD Jet2 operator*(double c, const Jet2& a) {
    return Jet2{c * a.v, c * a.g0, c * a.g1, c * a.h00, c * a.h01, c * a.h11};
}
MS_HD Jet2 operator*(const Jet2& a, const Jet2& b) {
    Jet2 r;
    r.v = a.v * b.v;
    r.g0 = a.g0 * b.v + a.v * b.g0;
    r.g1 = a.g1 * b.v + a.v * b.g1;
    r.h00 = a.h00 * b.v + 2.0 * a.g0 * b.g0 + a.v * b.h00;
    r.h01 = a.h01 * b.v + a.g0 * b.g1 + a.g1 * b.g0 + a.v * b.h01;
    r.h11 = a.h11 * b.v + 2.0 * a.g1 * b.g1 + a.v * b.h11;
    return r;
}
// f(a) for scalar f with derivatives f1, f2 at a.v
MS_HD Jet2 jchain(const Jet2& a, double f0, double f1, double f2) {
    Jet2 r;
    r.v = f0;
    r.g0 = f1 * a.g0;
    r.g1 = f1 * a.g1;
    r.h00 = f1 * a.h00 + f2 * a.g0 * a.g0;
    r.h01 = f1 * a.h01 + f2 * a.g0 * a.g1;
    r.h11 = f1 * a.h11 + f2 * a.g1 * a.g1;
    return r;
}
MS_HD Jet2 jsqrt(const Jet2& a) {
    double s, inv;
    sqrt_inv(a.v, s, inv);
    const double f1 = 0.5 * inv;
    return jchain(a, s, f1, -0.5 * f1 * inv * inv);
}
MS_HD Jet2 jrecip(const Jet2& a) {
    const double r = rcp_slack(a.v);
    return jchain(a, r, -r * r, 2.0 * r * r * r);
}

// ---- second-order jets in three variables (value, gradient, symmetric Hessian 00 01 02 11 12 22): used by the time-domain loss
// integration of integrateLosses=True, whose rows depend on (b_k, Fel_k, Fpb_k)
struct Jet3 {
    double v, g[3], h[6];
};
MS_HD int j3idx(int i, int j) { return (i <= j) ? (i == 0 ? j : (i == 1 ? 2 + j : 5)) : j3idx(j, i); }
MS_HD Jet3 j3const(double c) { return Jet3{c, {0, 0, 0}, {0, 0, 0, 0, 0, 0}}; }
MS_HD Jet3 j3var(double x, int i) { Jet3 r = j3const(x); r.g[i] = 1.0; return r; }
MS_HD Jet3 operator+(const Jet3& a, const Jet3& b) {
    Jet3 r; r.v = a.v + b.v;
    for (int i = 0; i < 3; ++i) r.g[i] = a.g[i] + b.g[i];
    for (int i = 0; i < 6; ++i) r.h[i] = a.h[i] + b.h[i];
    return r;
}
MS_HD Jet3 operator-(const Jet3& a, const Jet3& b) {
    Jet3 r; r.v = a.v - b.v;
    for (int i = 0; i < 3; ++i) r.g[i] = a.g[i] - b.g[i];
    for (int i = 0; i < 6; ++i) r.h[i] = a.h[i] - b.h[i];
    return r;
}
MS_HD Jet3 operator*(double c, const Jet3& a) {
    Jet3 r; r.v = c * a.v;
    for (int i = 0; i < 3; ++i) r.g[i] = c * a.g[i];
    for (int i = 0; i < 6; ++i) r.h[i] = c * a.h[i];
    return r;
}
MS_HD Jet3 operator+(const Jet3& a, double c) { Jet3 r = a; r.v += c; return r; }
MS_HD Jet3 operator*(const Jet3& a, const Jet3& b) {
    Jet3 r; r.v = a.v * b.v;
    for (int i = 0; i < 3; ++i) r.g[i] = a.g[i] * b.v + a.v * b.g[i];
    for (int i = 0; i < 3; ++i)
        for (int j = i; j < 3; ++j) {
            const int q = j3idx(i, j);
            r.h[q] = a.h[q] * b.v + a.g[i] * b.g[j] + a.g[j] * b.g[i] + a.v * b.h[q];
        }
    return r;
}
// f(a) for scalar f with derivatives f1, f2 at a.v
MS_HD Jet3 j3chain(const Jet3& a, double f0, double f1, double f2) {
    Jet3 r; r.v = f0;
    for (int i = 0; i < 3; ++i) r.g[i] = f1 * a.g[i];
    for (int i = 0; i < 3; ++i)
        for (int j = i; j < 3; ++j) { const int q = j3idx(i, j); r.h[q] = f1 * a.h[q] + f2 * a.g[i] * a.g[j]; }
    return r;
}
MS_HD Jet3 j3sqrt(const Jet3& a) {
    double s, inv;
    sqrt_inv(a.v, s, inv);
    const double f1 = 0.5 * inv;
    return j3chain(a, s, f1, -0.5 * f1 * inv * inv);
}
// S(u, w): composition of a two-variable jet S (value and partials up to second order w.r.t. its arguments) with jets u, w
MS_HD Jet3 j3compose(const Jet2& S, const Jet3& u, const Jet3& w) {
    Jet3 r; r.v = S.v;
    for (int i = 0; i < 3; ++i) r.g[i] = S.g0 * u.g[i] + S.g1 * w.g[i];
    for (int i = 0; i < 3; ++i)
        for (int j = i; j < 3; ++j) {
            const int q = j3idx(i, j);
            r.h[q] = S.g0 * u.h[q] + S.g1 * w.h[q] + S.h00 * u.g[i] * u.g[j] + S.h01 * (u.g[i] * w.g[j] + u.g[j] * w.g[i]) + S.h11 * w.g[i] * w.g[j];
        }
    return r;
}

}  // namespace mseetc
