// ns_core.cuh — arithmetic definitions of MCL_MODE_NS (the north-star formulation), usable on host and device.
//
// NS mode is this engine's own definition (the reference has no likelihood field, no per-particle noise and no
// systematic resampler: SURVEY.md D3-D5), so it is specified to be REPRODUCIBLE BY CONSTRUCTION: every quantity
// that feeds an integer decision is computed with IEEE basic operations and fma only (no libm), in a fixed order.
// A CPU restatement (oracle/mcl_oracle_ns.cpp) and any number of GPUs then agree bit for bit.
//
//   weights      W_i = floor(2^32 * exp(temper * (ll_i - max ll)))  as uint64 (Q32); prefix sums are integers, hence
//                order- and partition-independent
//   resampling   systematic: slot k takes ancestor(k) = min{ i : C_i * (N<<32) > ((k<<32) + u0) * C_total }, compared
//                in 128-bit integers (C = inclusive prefix of W over the GLOBAL particle order)
//   motion noise Philox4x32-10 keyed (seed; global particle index, stream, step) -> Box-Muller with the
//                deterministic log / sincospi below
#pragma once
#include <stdint.h>
#include <math.h>

#ifdef __CUDACC__
#define NS_HD __host__ __device__ __forceinline__
#else
#define NS_HD inline
#endif

namespace mcl {
namespace ns {

// ---- IEEE building blocks that are never contracted -------------------------------------------------------------------
#ifdef __CUDA_ARCH__
NS_HD double mul(double a, double b) { return __dmul_rn(a, b); }
NS_HD double add(double a, double b) { return __dadd_rn(a, b); }
NS_HD double fmad(double a, double b, double c) { return __fma_rn(a, b, c); }
NS_HD float mulf(float a, float b) { return __fmul_rn(a, b); }
NS_HD float addf(float a, float b) { return __fadd_rn(a, b); }
NS_HD float fmaf_(float a, float b, float c) { return __fmaf_rn(a, b, c); }
#else
// host translation units are built with -ffp-contract=off
NS_HD double mul(double a, double b) { return a * b; }
NS_HD double add(double a, double b) { return a + b; }
NS_HD double fmad(double a, double b, double c) { return fma(a, b, c); }
NS_HD float mulf(float a, float b) { return a * b; }
NS_HD float addf(float a, float b) { return a + b; }
NS_HD float fmaf_(float a, float b, float c) { return fmaf(a, b, c); }
#endif

// ---- sin/cos of pi*r for |r| <= 1/4, Taylor in f64 (error < 1e-16) -------------------------------------------------------
NS_HD void sincos_reduced(double x /* radians, |x| <= pi/4 */, double& s, double& c) {
    const double x2 = mul(x, x);
    // sin x = x * (1 - x2/6 + x2^2/120 - ...), cos x = 1 - x2/2 + x2^2/24 - ...
    double ps = -1.0 / 1307674368000.0;                         // -1/15!
    ps = fmad(ps, x2, 1.0 / 6227020800.0);                      //  1/13!
    ps = fmad(ps, x2, -1.0 / 39916800.0);                       // -1/11!
    ps = fmad(ps, x2, 1.0 / 362880.0);                          //  1/9!
    ps = fmad(ps, x2, -1.0 / 5040.0);                           // -1/7!
    ps = fmad(ps, x2, 1.0 / 120.0);                             //  1/5!
    ps = fmad(ps, x2, -1.0 / 6.0);                              // -1/3!
    s = fmad(mul(ps, x2), x, x);
    double pc = 1.0 / 20922789888000.0;                         //  1/16!
    pc = fmad(pc, x2, -1.0 / 87178291200.0);                    // -1/14!
    pc = fmad(pc, x2, 1.0 / 479001600.0);                       //  1/12!
    pc = fmad(pc, x2, -1.0 / 3628800.0);                        // -1/10!
    pc = fmad(pc, x2, 1.0 / 40320.0);                           //  1/8!
    pc = fmad(pc, x2, -1.0 / 720.0);                            // -1/6!
    pc = fmad(pc, x2, 1.0 / 24.0);                              //  1/4!
    pc = fmad(pc, x2, -0.5);
    c = fmad(pc, x2, 1.0);
}

// sin and cos of an angle in radians, |theta| up to ~1e5: Cody-Waite reduction by pi/2 in three parts.
NS_HD void det_sincos(double theta, double& s, double& c) {
    const double TWO_OVER_PI = 0.63661977236758134308;
    const double P1 = 1.57079632673412561417e+00;     // pi/2 high 33 bits
    const double P2 = 6.07710050630396597660e-11;     // next 33 bits
    const double P3 = 2.02226624879595063154e-21;     // remainder
    const double kf = rint(mul(theta, TWO_OVER_PI));
    double r = fmad(-kf, P1, theta);
    r = fmad(-kf, P2, r);
    r = fmad(-kf, P3, r);
    double sr, cr;
    sincos_reduced(r, sr, cr);
    const long long k = (long long)kf;
    switch (k & 3) {
        case 0: s = sr; c = cr; break;
        case 1: s = cr; c = -sr; break;
        case 2: s = -sr; c = -cr; break;
        default: s = -cr; c = sr; break;
    }
}
// fp32 sin/cos of an fp32 angle: evaluate in f64, round once.
NS_HD void det_sincosf(float theta, float& s, float& c) {
    double sd, cd;
    det_sincos((double)theta, sd, cd);
    s = (float)sd;
    c = (float)cd;
}

// natural log of x in (0, 1], f64: x = m * 2^e with m in [sqrt(1/2), sqrt(2)), log m = 2 atanh((m-1)/(m+1)).
NS_HD double det_log(double x) {
    union { double d; uint64_t u; } v;
    v.d = x;
    int e = (int)((v.u >> 52) & 0x7ff) - 1023;
    v.u = (v.u & 0x000fffffffffffffull) | 0x3ff0000000000000ull;     // m in [1,2)
    double m = v.d;
    if (m > 1.4142135623730951) { m = mul(m, 0.5); e += 1; }
    const double t = (m - 1.0) / (m + 1.0);                           // IEEE division
    const double t2 = mul(t, t);
    double p = 1.0 / 23.0;
    p = fmad(p, t2, 1.0 / 21.0);
    p = fmad(p, t2, 1.0 / 19.0);
    p = fmad(p, t2, 1.0 / 17.0);
    p = fmad(p, t2, 1.0 / 15.0);
    p = fmad(p, t2, 1.0 / 13.0);
    p = fmad(p, t2, 1.0 / 11.0);
    p = fmad(p, t2, 1.0 / 9.0);
    p = fmad(p, t2, 1.0 / 7.0);
    p = fmad(p, t2, 1.0 / 5.0);
    p = fmad(p, t2, 1.0 / 3.0);
    p = fmad(mul(p, t2), t, t);                                       // atanh(t)
    return fmad((double)e, 0.69314718055994530942, mul(2.0, p));
}

// Two standard normals from two 32-bit words (Box-Muller): u = (word + 0.5) * 2^-32 in (0,1).
NS_HD void det_normal_pair(uint32_t w1, uint32_t w2, float& z0, float& z1) {
    const double u1 = mul(add((double)w1, 0.5), 2.3283064365386963e-10);
    const double u2 = mul(add((double)w2, 0.5), 2.3283064365386963e-10);
    const double r = sqrt(mul(-2.0, det_log(u1)));
    double s, c;
    det_sincos(mul(6.28318530717958647692, u2), s, c);
    z0 = (float)mul(r, c);
    z1 = (float)mul(r, s);
}

// W = floor(2^32 * exp(t)) for t <= 0 (fp32), as uint64 in [0, 2^32].
NS_HD uint64_t det_exp_q32(float t) {
    if (!(t > -22.5f)) return 0;                                      // below 2^-32 (also catches NaN)
    if (t >= 0.f) return 1ull << 32;
    const double y = mul((double)t, 1.44269504088896340736);          // t * log2(e)
    const double kf = floor(y);
    const double g = mul(y - kf, 0.69314718055994530942);             // in [0, ln 2)
    double p = 1.0 / 6227020800.0;                                    // 1/13!
    p = fmad(p, g, 1.0 / 479001600.0);
    p = fmad(p, g, 1.0 / 39916800.0);
    p = fmad(p, g, 1.0 / 3628800.0);
    p = fmad(p, g, 1.0 / 362880.0);
    p = fmad(p, g, 1.0 / 40320.0);
    p = fmad(p, g, 1.0 / 5040.0);
    p = fmad(p, g, 1.0 / 720.0);
    p = fmad(p, g, 1.0 / 120.0);
    p = fmad(p, g, 1.0 / 24.0);
    p = fmad(p, g, 1.0 / 6.0);
    p = fmad(p, g, 0.5);
    p = fmad(p, g, 1.0);
    p = fmad(p, g, 1.0);                                              // exp(g) in [1, 2)
    const int k = (int)kf;                                            // -33 .. -1
    union { double d; uint64_t u; } sc;
    sc.u = (uint64_t)(1023 + 32 + k) << 52;                           // 2^(32+k), exact
    const double scaled = mul(p, sc.d);
    uint64_t w = (uint64_t)scaled;                                    // truncation = floor for positives
    return w > (1ull << 32) ? (1ull << 32) : w;
}

// ---- systematic resampling in integers --------------------------------------------------------------------------------
struct U128 { uint64_t hi, lo; };
NS_HD U128 mul64(uint64_t a, uint64_t b) {
    U128 r;
#ifdef __CUDA_ARCH__
    r.lo = a * b;
    r.hi = __umul64hi(a, b);
#else
    unsigned __int128 p = (unsigned __int128)a * b;
    r.lo = (uint64_t)p;
    r.hi = (uint64_t)(p >> 64);
#endif
    return r;
}
NS_HD bool gt128(const U128& a, const U128& b) { return a.hi > b.hi || (a.hi == b.hi && a.lo > b.lo); }
// c * (n << 32)   (c < 2^60, n < 2^31  =>  < 2^123)
NS_HD U128 lhs_of(uint64_t c, uint64_t n) {
    U128 p = mul64(c, n);
    U128 r;
    r.hi = (p.hi << 32) | (p.lo >> 32);
    r.lo = p.lo << 32;
    return r;
}
// ((k << 32) + u0) * total
NS_HD U128 rhs_of(uint64_t k, uint32_t u0, uint64_t total) { return mul64((k << 32) + u0, total); }
// slot k selects the first particle whose inclusive prefix c satisfies selects(c, ...)
NS_HD bool selects(uint64_t c, uint64_t n, const U128& rhs) { return gt128(lhs_of(c, n), rhs); }

// theta wrapped into [-pi, pi] with fp32 constants (at most a few turns off)
NS_HD float wrap_pi(float t) {
    const float PI_F = 3.14159274f, TWO_PI_F = 6.28318548f;
    for (int i = 0; i < 4 && t > PI_F; i++) t = addf(t, -TWO_PI_F);
    for (int i = 0; i < 4 && t < -PI_F; i++) t = addf(t, TWO_PI_F);
    return t;
}

}  // namespace ns
}  // namespace mcl
