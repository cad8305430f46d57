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
NS_HD float divf(float a, float b) { return __fdiv_rn(a, b); }
NS_HD float sqrtf_(float a) { return __fsqrt_rn(a); }
#else
// host translation units are built with -ffp-contract=off
NS_HD double mul(double a, double b) { return a * b; }
NS_HD double add(double a, double b) { return a + b; }
NS_HD double fmad(double a, double b, double c) { return fma(a, b, c); }
NS_HD float mulf(float a, float b) { return a * b; }
NS_HD float addf(float a, float b) { return a + b; }
NS_HD float fmaf_(float a, float b, float c) { return fmaf(a, b, c); }
NS_HD float divf(float a, float b) { return a / b; }
NS_HD float sqrtf_(float a) { return sqrtf(a); }
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

// ---- fp32 elementary functions for the motion noise (NS-7): IEEE add/mul/div/sqrt/fma only, fixed order -------------------
// sin/cos of an fp32 angle (|theta| up to a few turns): Cody-Waite reduction by pi/2 in three fp32 parts, degree-9/10
// Taylor polynomials; within ~1 ulp of the exact value.
NS_HD void sincos_poly_f(float x /* |x| <= pi/4 */, float& s, float& c) {
    const float x2 = mulf(x, x);
    float ps = 2.75573192e-06f;                                       //  1/9!
    ps = fmaf_(ps, x2, -1.98412701e-04f);                             // -1/7!
    ps = fmaf_(ps, x2, 8.33333377e-03f);                              //  1/5!
    ps = fmaf_(ps, x2, -1.66666672e-01f);                             // -1/3!
    s = fmaf_(mulf(ps, x2), x, x);
    float pc = -2.75573200e-07f;                                      // -1/10!
    pc = fmaf_(pc, x2, 2.48015876e-05f);                              //  1/8!
    pc = fmaf_(pc, x2, -1.38888892e-03f);                             // -1/6!
    pc = fmaf_(pc, x2, 4.16666679e-02f);                              //  1/4!
    pc = fmaf_(pc, x2, -0.5f);
    c = fmaf_(pc, x2, 1.0f);
}
NS_HD void quadrant_f(int k, float sr, float cr, float& s, float& c) {
    switch (k & 3) {
        case 0: s = sr; c = cr; break;
        case 1: s = cr; c = -sr; break;
        case 2: s = -sr; c = -cr; break;
        default: s = -cr; c = sr; break;
    }
}
NS_HD void det_sincosf32(float theta, float& s, float& c) {
    const float kf = rintf(mulf(theta, 0.636619747f));               // theta * 2/pi
    float r = fmaf_(-kf, 1.5703125f, theta);                          // pi/2 = 1.5703125 + 4.837512969970703125e-4 + 7.549789954891882e-8
    r = fmaf_(-kf, 4.837512969970703125e-4f, r);
    r = fmaf_(-kf, 7.54978995489188216e-8f, r);
    float sr, cr;
    sincos_poly_f(r, sr, cr);
    quadrant_f((int)kf, sr, cr, s, c);
}
// sin/cos of 2*pi*u for u in [0,1): exact quadrant reduction in units of a turn
NS_HD void det_sincos2pif(float u, float& s, float& c) {
    const float q = rintf(mulf(u, 4.0f));
    const float f = fmaf_(q, -0.25f, u);                              // exact, |f| <= 1/8
    const float x = fmaf_(f, 6.28318548f, mulf(f, -1.74845553e-07f)); // 2 pi f (2 pi = 6.28318548 - 1.74845553e-07)
    float sr, cr;
    sincos_poly_f(x, sr, cr);
    quadrant_f((int)q, sr, cr, s, c);
}
// natural log of a normal fp32 x in (0, 1]: x = m * 2^e, m in [sqrt(1/2), sqrt(2)), log m = 2 atanh((m-1)/(m+1))
NS_HD float det_logf(float x) {
    union { float f; uint32_t u; } v;
    v.f = x;
    int e = (int)(v.u >> 23) - 127;
    v.u = (v.u & 0x007fffffu) | 0x3f800000u;                          // m in [1,2)
    float m = v.f;
    if (m > 1.41421354f) { m = mulf(m, 0.5f); e += 1; }
    const float t = divf(addf(m, -1.0f), addf(m, 1.0f));
    const float t2 = mulf(t, t);
    float p = 0.111111112f;                                           // 1/9
    p = fmaf_(p, t2, 0.142857149f);                                   // 1/7
    p = fmaf_(p, t2, 0.200000003f);                                   // 1/5
    p = fmaf_(p, t2, 0.333333343f);                                   // 1/3
    p = fmaf_(mulf(p, t2), t, t);                                     // atanh(t)
    return fmaf_((float)e, 0.693147182f, mulf(2.0f, p));
}
// Two standard normals from two 32-bit words (Box-Muller): u = ((word >> 9) + 0.5) * 2^-23 in (0,1), exact in fp32.
NS_HD void det_normal_pair(uint32_t w1, uint32_t w2, float& z0, float& z1) {
    const float u1 = mulf(addf((float)(w1 >> 9), 0.5f), 1.1920929e-07f);
    const float u2 = mulf(addf((float)(w2 >> 9), 0.5f), 1.1920929e-07f);
    const float r = sqrtf_(mulf(-2.0f, det_logf(u1)));
    float s, c;
    det_sincos2pif(u2, s, c);
    z0 = mulf(r, c);
    z1 = mulf(r, s);
}

// W = trunc(2^32 * e), e = exp(t) evaluated in fp32 (IEEE ops only, 24 significant bits), for t <= 0; uint64 in [0, 2^32].
NS_HD uint64_t det_exp_q32(float t) {
    if (!(t > -22.5f)) return 0;                                      // below 2^-32 (also catches NaN)
    if (t >= 0.f) return 1ull << 32;
    const float kf = rintf(mulf(t, 1.44269502f));                     // -32 .. 0
    float g = fmaf_(-kf, 0.693145752f, t);                            // ln 2 = 0.693145752 + 1.42860677e-06 (Cody-Waite)
    g = fmaf_(-kf, 1.42860677e-06f, g);                               // |g| <= 0.347
    float p = 1.98412701e-04f;                                        // 1/7!
    p = fmaf_(p, g, 1.38888892e-03f);                                 // 1/6!
    p = fmaf_(p, g, 8.33333377e-03f);                                 // 1/5!
    p = fmaf_(p, g, 4.16666679e-02f);                                 // 1/4!
    p = fmaf_(p, g, 1.66666672e-01f);                                 // 1/3!
    p = fmaf_(p, g, 0.5f);
    p = fmaf_(p, g, 1.0f);
    p = fmaf_(p, g, 1.0f);                                            // exp(g) in [0.70, 1.42]
    union { float f; uint32_t u; } sc;
    sc.u = (uint32_t)(127 + 32 + (int)kf) << 23;                      // 2^(32+k), exact
    const uint64_t w = (uint64_t)mulf(p, sc.f);                       // exact scaling, truncation
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

// theta brought back towards [-pi, pi]: one conditional turn each way with fp32 constants (the per-step change of
// theta is far below a turn; det_sincosf32 is accurate over several turns anyway)
NS_HD float wrap_pi(float t) {
    const float PI_F = 3.14159274f, TWO_PI_F = 6.28318548f;
    if (t > PI_F) t = addf(t, -TWO_PI_F);
    if (t < -PI_F) t = addf(t, TWO_PI_F);
    return t;
}

}  // namespace ns
}  // namespace mcl
