// fastmath64.cuh -- lean double-precision sincos / exp / log / tanh for the environment kernels (sm_100a).
//
// Why: ncu on the UAV step kernel (profiles/) showed 71 % of all executed instructions inside libdevice's
// sincos/tanh/log/exp/pow, and 27 % of ALL instructions being UMOV pairs that materialise fp64 polynomial
// coefficients as immediates (DFMA has no fp64-immediate form on sm_100a).  These versions
//   * keep the coefficients in __constant__ memory, which the compiler fetches with LDCU.128 (two doubles per
//     instruction instead of two UMOVs per double),
//   * drop the special-case ladders the env kernels never need (huge arguments, denormal results), keeping cheap
//     guards so that 0, inf and NaN still behave as in IEEE libm,
//   * stay within ~1-1.5 ulp (coefficients: tools/gen_fastmath_coeffs.py, near-minimax fits made with mpmath;
//     accuracy is re-measured on the GPU by tests/test_fastmath_gpu.py).
// Accuracy note: tanh() is accurate to ~2e-16 ABSOLUTE (not relative) for |x| -> 0, which is what the controllers
// need (it always multiplies an O(1) gain).
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <math_constants.h>
#include <stdint.h>

namespace fm64 {

// S: sin(r) = r + r z S(z);  C: cos(r) = 1 - z/2 + z^2 C(z);  z = r^2, |r| <= pi/4
static __constant__ double kS[6] = {-0x1.5555555555555p-3, 0x1.1111111110b23p-7, -0x1.a01a019e623afp-13,
                                    0x1.71de376b9c575p-19, -0x1.ae5fdeb852566p-26, 0x1.5dfb4dfaf15abp-33};
static __constant__ double kC[6] = {0x1.5555555555555p-5, -0x1.6c16c16c1691fp-10, 0x1.a01a019f3df42p-16,
                                    -0x1.27e4fa025e03dp-22, 0x1.1eeb52bea9072p-29, -0x1.906dd38cb66a1p-37};
// E: exp(r) = 1 + r + r^2 E(r), |r| <= ln2/2
static __constant__ double kE[10] = {0x1.0000000000001p-1, 0x1.5555555555556p-3, 0x1.5555555553b6cp-5,
                                     0x1.1111111110918p-7, 0x1.6c16c1794e1dcp-10, 0x1.a01a01a83ba17p-13,
                                     0x1.a019b621f9d08p-16, 0x1.71de0be2b5e96p-19, 0x1.2894f52583578p-22,
                                     0x1.af3ce42b12b24p-26};
// L: log(m) = 2s + s z L(z), s = (m-1)/(m+1), z = s^2, m in [sqrt(1/2), sqrt(2))
static __constant__ double kL[7] = {0x1.5555555555558p-1, 0x1.99999999949c3p-2, 0x1.2492492ef134dp-2,
                                    0x1.c71c61a265960p-3, 0x1.74630fb47b087p-3, 0x1.39f2ac8e848c3p-3,
                                    0x1.2be78035f90e7p-3};
// A: asin(t) = t + t z A(z), z = t^2 <= 1/4 (degree 11; |x| > 1/2 goes through pi/2 - 2 asin(sqrt((1 - |x|) / 2)))
static __constant__ double kA[12] = {0x1.555555555554fp-3, 0x1.3333333336f25p-4, 0x1.6db6db682abf8p-5,
                                     0x1.f1c71fa66977fp-6, 0x1.6e8b28e7f8608p-6, 0x1.1c596afc9d8acp-6,
                                     0x1.c86de3215c77bp-7, 0x1.8546aebb901aap-7, 0x1.fdf3b5f7fbfb1p-8,
                                     0x1.07d821fcdce60p-6, -0x1.64309bd35e9c9p-7, 0x1.cf09bf99f54b7p-6};
// reduction constants: pi/2 in three pieces, 2/pi, ln2 in two pieces, log2(e), 1.5 * 2^52 (round-to-nearest magic)
static __constant__ double kR[8] = {0x1.921fb54442d18p+0, 0x1.1a62633145c07p-54, -0x1.f1976b7ed8fbcp-110,
                                    0x1.45f306dc9c883p-1, 0x1.62e42fefa39efp-1, 0x1.abc9e3b39803fp-56,
                                    0x1.71547652b82fep+0, 6755399441055744.0};

// 1/y and x/y without the slow-path branches of the compiler's IEEE division: MUFU.RCP64H seed (>= 20 bits), two
// Newton steps, one correction step for the quotient.  Valid for normal, finite y (every call site guarantees it);
// <= 1 ulp.  Branch-free code keeps consecutive function evaluations in ONE basic block, so the scheduler can
// interleave their dependent DFMA chains (the kernels are latency-bound: ncu "stall_wait").
__device__ __forceinline__ double rcp(double y) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(y));
    double e = fma(-y, r, 1.0);
    r = fma(r, e, r);
    e = fma(-y, r, 1.0);
    r = fma(r, e, r);
    return r;
}
__device__ __forceinline__ double div(double x, double y) {
    const double r = rcp(y);
    const double q = x * r;
    return fma(fma(-y, q, x), r, q);
}

// sqrt without the compiler's slow-path branch (which would split the caller's basic block): MUFU.RSQ64H seed, two
// Goldschmidt steps, one FMA correction.  Correctly rounded except in rare half-way cases (<= 1 ulp); 0 -> 0,
// +inf -> +inf, negative / NaN -> NaN like IEEE sqrt; subnormal arguments are treated as 0.
__device__ __forceinline__ double sqrt_nb(double x) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double g = x * y, h = 0.5 * y;
    double r = fma(-g, h, 0.5);
    g = fma(g, r, g);
    h = fma(h, r, h);
    r = fma(-g, h, 0.5);
    g = fma(g, r, g);
    h = fma(h, r, h);
    g = fma(fma(-g, g, x), h, g);
    return (x == 0.0 || x == CUDART_INF) ? x : g;
}

// sin and cos of x for |x| < 1e5 (three-term Cody-Waite reduction by pi/2, two degree-5 kernels in z = r^2).
// No argument check: callers go through sincos() / sincos3() below.
__device__ __forceinline__ void sincos_core(double x, double *sn, double *cs) {
    const double t = fma(x, kR[3], kR[7]); // x * 2/pi rounded to the nearest integer in the low mantissa bits
    const int n = __double2loint(t);
    const double fn = t - kR[7];
    double r = fma(-fn, kR[0], x);
    r = fma(-fn, kR[1], r);
    r = fma(-fn, kR[2], r);
    const double z = r * r;
    double ps = kS[5];
    ps = fma(ps, z, kS[4]); ps = fma(ps, z, kS[3]); ps = fma(ps, z, kS[2]); ps = fma(ps, z, kS[1]); ps = fma(ps, z, kS[0]);
    ps = fma(r * z, ps, r);
    double pc = kC[5];
    pc = fma(pc, z, kC[4]); pc = fma(pc, z, kC[3]); pc = fma(pc, z, kC[2]); pc = fma(pc, z, kC[1]); pc = fma(pc, z, kC[0]);
    pc = fma(z, fma(z, pc, -0.5), 1.0);
    const double s = (n & 1) ? pc : ps;
    const double c = (n & 1) ? ps : pc;
    // quadrant signs straight into the sign bit: bit 1 of n (resp. n + 1) moved to bit 31 of the high word
    *sn = __hiloint2double(__double2hiint(s) ^ ((n & 2) << 30), __double2loint(s));
    *cs = __hiloint2double(__double2hiint(c) ^ (((n + 1) & 2) << 30), __double2loint(c));
}

__device__ __forceinline__ void sincos(double x, double *sn, double *cs) {
    if (!(fabs(x) < 1.0e5)) { // huge / inf / NaN: not on the env kernels' paths
        ::sincos(x, sn, cs);
        return;
    }
    sincos_core(x, sn, cs);
}

// three angles behind ONE range check, so that the six polynomial chains share a basic block
__device__ __forceinline__ void sincos3(double a, double b, double c, double *sa, double *ca, double *sb, double *cb,
                                        double *sc, double *cc) {
    if (!(fabs(a) + fabs(b) + fabs(c) < 1.0e5)) { // huge, inf or NaN in any of the three (the sum propagates NaN)
        ::sincos(a, sa, ca);
        ::sincos(b, sb, cb);
        ::sincos(c, sc, cc);
        return;
    }
    sincos_core(a, sa, ca);
    sincos_core(b, sb, cb);
    sincos_core(c, sc, cc);
}
__device__ __forceinline__ void sincos2(double a, double b, double *sa, double *ca, double *sb, double *cb) {
    if (!(fabs(a) + fabs(b) < 1.0e5)) {
        ::sincos(a, sa, ca);
        ::sincos(b, sb, cb);
        return;
    }
    sincos_core(a, sa, ca);
    sincos_core(b, sb, cb);
}

// exp without the range guards: valid for -708 < x < 709 (tanh calls it on [0, 40])
__device__ __forceinline__ double exp_core(double x) {
    const double t = fma(x, kR[6], kR[7]);
    const int n = __double2loint(t);
    const double fn = t - kR[7];
    double r = fma(-fn, kR[4], x);
    r = fma(-fn, kR[5], r);
    // even/odd split Horner: two chains of depth 5 instead of one of depth 9
    const double r2 = r * r;
    double pe = kE[8], po = kE[9];
    pe = fma(pe, r2, kE[6]); po = fma(po, r2, kE[7]);
    pe = fma(pe, r2, kE[4]); po = fma(po, r2, kE[5]);
    pe = fma(pe, r2, kE[2]); po = fma(po, r2, kE[3]);
    pe = fma(pe, r2, kE[0]); po = fma(po, r2, kE[1]);
    const double p = fma(po, r, pe);
    const double y = fma(r2, p, r) + 1.0;
    return y * __hiloint2double((n + 1023) << 20, 0); // 2^n by exponent construction
}

__device__ __forceinline__ double exp(double x) {
    double y = exp_core(x);
    if (x < -708.0) y = 0.0;       // (denormal results flush to 0)
    if (x > 709.0) y = CUDART_INF;
    return y;                      // NaN in -> NaN out (the comparisons are false, the polynomial is NaN)
}

// natural log, branch-free.  x = 0 (and subnormal x, treated as 0) -> -inf, x < 0 or NaN -> NaN, +inf -> +inf.
__device__ __forceinline__ double log(double x) {
    int hi = __double2hiint(x);
    const int lo = __double2loint(x);
    int k = (hi >> 20) - 1023;
    hi &= 0x000fffff;
    const int i = (hi + 0x95f64) & 0x100000; // m >= sqrt(2): halve it and bump the exponent
    k += i >> 20;
    const double m = __hiloint2double(hi | (i ^ 0x3ff00000), lo);
    const double f = m - 1.0;
    const double s = div(f, 2.0 + f);
    const double z = s * s;
    double R = kL[6];
    R = fma(R, z, kL[5]); R = fma(R, z, kL[4]); R = fma(R, z, kL[3]); R = fma(R, z, kL[2]); R = fma(R, z, kL[1]);
    R = fma(R, z, kL[0]);
    R *= z;
    const double hfsq = 0.5 * f * f;
    const double dk = (double)k;
    double y = dk * kR[4] - ((hfsq - (s * (hfsq + R) + dk * kR[5])) - f);
    if (x < 2.2250738585072014e-308) y = -CUDART_INF; // zero / subnormal
    if (!(x >= 0.0)) y = CUDART_NAN;                   // negative / NaN
    if (x == CUDART_INF) y = x;
    return y;
}

// log of a value that is an absolute value (|e|, |s| of the sliding surfaces): the same evaluation without the two guards
// a magnitude cannot trigger by being negative.  NaN and +inf arguments give an unspecified finite value here; the callers'
// expressions carry the argument itself in a linear term, which propagates NaN / inf whatever this returns.
__device__ __forceinline__ double log_nonneg(double x) {
    int hi = __double2hiint(x);
    const int lo = __double2loint(x);
    int k = (hi >> 20) - 1023;
    hi &= 0x000fffff;
    const int i = (hi + 0x95f64) & 0x100000;
    k += i >> 20;
    const double m = __hiloint2double(hi | (i ^ 0x3ff00000), lo);
    const double f = m - 1.0;
    const double s = div(f, 2.0 + f);
    const double z = s * s;
    double R = kL[6];
    R = fma(R, z, kL[5]); R = fma(R, z, kL[4]); R = fma(R, z, kL[3]); R = fma(R, z, kL[2]); R = fma(R, z, kL[1]);
    R = fma(R, z, kL[0]);
    R *= z;
    const double hfsq = 0.5 * f * f;
    const double dk = (double)k;
    const double y = dk * kR[4] - ((hfsq - (s * (hfsq + R) + dk * kR[5])) - f);
    return x < 2.2250738585072014e-308 ? -CUDART_INF : y; // zero / subnormal
}

// tanh, branch-free; ~2e-16 absolute accuracy
__device__ __forceinline__ double tanh(double x) {
    const double ax = (fabs(x) > 20.0) ? 20.0 : fabs(x); // tanh(20) rounds to 1; NaN stays NaN (comparison false)
    const double t = exp_core(2.0 * ax);
    const double y = fma(-2.0, rcp(t + 1.0), 1.0);
    return copysign(y, x);
}

// asin for |x| <= 1 (NaN outside), branch-free: one degree-11 kernel on t <= 1/2, the upper half folded by
// asin(x) = pi/2 - 2 asin(sqrt((1 - |x|) / 2)).  <= 1.5 ulp.
__device__ __forceinline__ double asin(double x) {
    const double ax = fabs(x);
    const bool big = ax > 0.5;
    const double w = big ? 0.5 * (1.0 - ax) : ax * ax; // z = t^2
    const double t = big ? sqrt_nb(w) : ax;
    double p = kA[11];
    p = fma(p, w, kA[10]); p = fma(p, w, kA[9]); p = fma(p, w, kA[8]); p = fma(p, w, kA[7]); p = fma(p, w, kA[6]);
    p = fma(p, w, kA[5]); p = fma(p, w, kA[4]); p = fma(p, w, kA[3]); p = fma(p, w, kA[2]); p = fma(p, w, kA[1]);
    p = fma(p, w, kA[0]);
    const double a = fma(t * w, p, t);
    // pi/2 - 2a with the low part of pi/2 folded in
    const double r = big ? (kR[0] - 2.0 * a) + kR[1] : a;
    return copysign(r, x);
}

} // namespace fm64
