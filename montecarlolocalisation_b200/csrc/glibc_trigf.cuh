// glibc_trigf.cuh — sinf / cosf exactly as the reference binary executes them (MC:644-645 call cosf/sinf; the Eigen fp32
// cos/sin of MC:747-748 reach the same libm functions through oracle/shim), so that the engine's float trig is
// bit-identical to the compiled reference instead of "correctly rounded".
//
// Third-party algorithm restated here, not part of /root/reference: GNU libc 2.39 (Ubuntu GLIBC 2.39-0ubuntu8.5, the libm of
// this image), sysdeps/ieee754/flt-32/{s_sinf.c, s_cosf.c, sincosf.h, sincosf_data.c} (the ARM optimized-routines
// sinf/cosf): the argument is widened to f64, reduced by pi/2 (|y| < 120: one multiply by 2^24 * 2/pi, truncation,
// round-to-nearest quadrant; larger: 3 x 32-bit words of 4/pi selected by the exponent), and a degree-7 sine or degree-8
// cosine polynomial in f64 is rounded to float once. On x86-64 libm selects one of two builds of the same source at load
// time (ifunc): the FMA build (CPUs with FMA + AVX2), in which every `a + b * c` of the source is one fused operation, or
// the SSE2 build without fusion. Which operations are fused was read off the disassembly of this image's libm.so.6
// (__sinf_fma / __cosf_fma); FMA = true / false below reproduces the two builds. The engine picks the build the HOST's
// libm uses (Engine::open probes sinf/cosf on arguments where the two differ), because "what the reference computes"
// means what the reference binary computes on this machine.
//
// tests/native/glibc_trigf_check.cpp compares both functions with the host libm over all 2^32 float bit patterns
// (exhaustive mode) or a 1-in-61 sample (the CPU test suite); tests/test_gpu_ref_parity.py compares the device form with
// the golden vectors recorded from the compiled reference.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#ifdef __CUDACC__
#define GT_HD __host__ __device__ __forceinline__
#else
#define GT_HD inline
#endif

namespace mcl {
namespace glibc_trig {

GT_HD double gt_mul(double a, double b) {
#ifdef __CUDA_ARCH__
    return __dmul_rn(a, b);
#else
    return a * b;
#endif
}
// a * b + c: fused in the FMA build, two roundings in the SSE2 build
template <bool FMA>
GT_HD double gt_mad(double a, double b, double c) {
#ifdef __CUDA_ARCH__
    return FMA ? __fma_rn(a, b, c) : __dadd_rn(__dmul_rn(a, b), c);
#else
    if (FMA) return fma(a, b, c);
    volatile double p = a * b;     // volatile: never contracted, whatever the host compiler flags
    return p + c;
#endif
}
GT_HD uint32_t gt_bits(float y) {
#ifdef __CUDA_ARCH__
    return __float_as_uint(y);
#else
    uint32_t b; memcpy(&b, &y, 4); return b;
#endif
}
GT_HD float gt_float(uint32_t b) {
#ifdef __CUDA_ARCH__
    return __uint_as_float(b);
#else
    float y; memcpy(&y, &b, 4); return y;
#endif
}
GT_HD float gt_narrow(double v) {
#ifdef __CUDA_ARCH__
    return __double2float_rn(v);
#else
    return (float)v;
#endif
}

// sine (odd == false) or cosine (odd == true) polynomial on the reduced argument: x = r * sign, x2 = r * r
template <bool FMA>
GT_HD float gt_poly(double x, double x2, bool negated, bool odd) {
    // __sincosf_table[0] / [1]: [1] holds the cosine coefficients negated (quadrants 2 and 3)
    const double c0 = negated ? -0x1p0 : 0x1p0, c1 = negated ? 0x1.ffffffd0c621cp-2 : -0x1.ffffffd0c621cp-2;
    const double c2 = negated ? -0x1.55553e1068f19p-5 : 0x1.55553e1068f19p-5, c3 = negated ? 0x1.6c087e89a359dp-10 : -0x1.6c087e89a359dp-10;
    const double c4 = negated ? -0x1.99343027bf8c3p-16 : 0x1.99343027bf8c3p-16;
    const double s1 = -0x1.555545995a603p-3, s2 = 0x1.1107605230bc4p-7, s3 = -0x1.994eb3774cf24p-13;
    if (!odd) {
        const double x3 = gt_mul(x, x2);
        const double t = gt_mad<FMA>(x2, s3, s2);          // s2 + x2 * s3
        const double x7 = gt_mul(x3, x2);
        const double s = gt_mad<FMA>(x3, s1, x);           // x + x3 * s1
        return gt_narrow(gt_mad<FMA>(x7, t, s));           // s + x7 * t
    }
    const double x4 = gt_mul(x2, x2);
    const double hi = gt_mad<FMA>(x2, c4, c3);             // c3 + x2 * c4
    const double lo = gt_mad<FMA>(x2, c1, c0);             // c0 + x2 * c1
    const double x6 = gt_mul(x4, x2);
    const double c = gt_mad<FMA>(x4, c2, lo);              // lo + x4 * c2
    return gt_narrow(gt_mad<FMA>(x6, hi, c));              // c + x6 * hi
}

// 4/pi as a bit string, in 32-bit windows that start every 8 bits (__inv_pio4)
#define GT_INV_PIO4_WORDS                                                                                                   \
    {0xa2u,       0xa2f9u,     0xa2f983u,   0xa2f9836eu, 0xf9836e4eu, 0x836e4e44u, 0x6e4e4415u, 0x4e441529u,                  \
     0x441529fcu, 0x1529fc27u, 0x29fc2757u, 0xfc2757d1u, 0x2757d1f5u, 0x57d1f534u, 0xd1f534ddu, 0xf534ddc0u,                  \
     0x34ddc0dbu, 0xddc0db62u, 0xc0db6295u, 0xdb629599u, 0x6295993cu, 0x95993c43u, 0x993c4390u, 0x3c439041u}
#ifdef __CUDACC__
static __device__ __constant__ uint32_t gt_inv_pio4_dev[24] = GT_INV_PIO4_WORDS;
#endif
static const uint32_t gt_inv_pio4_host[24] = GT_INV_PIO4_WORDS;

// |y| >= 120: three words of 4/pi selected by the exponent; the product's top two bits are the quadrant
GT_HD double gt_reduce_large(uint32_t xi, int* np) {
#ifdef __CUDA_ARCH__
    const uint32_t* inv_pio4 = gt_inv_pio4_dev;
#else
    const uint32_t* inv_pio4 = gt_inv_pio4_host;
#endif
    const int a = (int)((xi >> 26) & 15u);
    const int shift = (int)((xi >> 23) & 7u);
    uint32_t m = (xi & 0xffffffu) | 0x800000u;
    m <<= shift;
    uint64_t res0 = (uint64_t)(uint32_t)(m * inv_pio4[a]);          // 32-bit product (wraps)
    const uint64_t res1 = (uint64_t)m * inv_pio4[a + 4];
    const uint64_t res2 = (uint64_t)m * inv_pio4[a + 8];
    res0 = (res2 >> 32) | (res0 << 32);
    res0 += res1;
    const uint64_t n = (res0 + (1ull << 61)) >> 62;
    res0 -= n << 62;
    *np = (int)n;
#ifdef __CUDA_ARCH__
    return __dmul_rn(__ll2double_rn((long long)res0), 0x1.921FB54442D18p-62);
#else
    return (double)(int64_t)res0 * 0x1.921FB54442D18p-62;
#endif
}

GT_HD float gt_invalid(float y) {          // __math_invalidf: (y - y) / (y - y) as x86 evaluates it
    const uint32_t b = gt_bits(y);
    if ((b & 0x7fffffffu) > 0x7f800000u) return gt_float(b | 0x00400000u);      // NaN in: the same NaN, quieted
    return gt_float(0xffc00000u);                                               // Inf in: the default NaN
}

// sign of sine in quadrant q & 3: {1, -1, -1, 1}
GT_HD double gt_sign(int q) { return ((q + 1) & 2) ? -1.0 : 1.0; }

// ODD_BASE: 0 for sinf, 1 for cosf (cosf evaluates the polynomial of quadrant n ^ 1)
template <bool FMA, int ODD_BASE>
GT_HD float gt_sincosf(float y) {
    const uint32_t xi = gt_bits(y);
    const uint32_t top = (xi >> 20) & 0x7ffu;                       // abstop12
    const double x = (double)y;
    if (top < 0x3f4u) {                                             // |y| < pi/4
        if (top < 0x398u) return ODD_BASE ? 1.0f : y;               // |y| < 2^-12
        return gt_poly<FMA>(x, gt_mul(x, x), false, ODD_BASE != 0);
    }
    if (top < 0x42fu) {                                             // |y| < 120
        const double r = gt_mul(x, 0x1.45F306DC9C883p+23);         // 2/pi * 2^24: the quadrant lands in bits 24..31
#ifdef __CUDA_ARCH__
        const int n = (__double2int_rz(r) + 0x800000) >> 24;
#else
        const int n = ((int32_t)r + 0x800000) >> 24;
#endif
        const double xr = gt_mad<FMA>(-(double)n, 0x1.921FB54442D18p0, x);     // x - n * pi/2
        return gt_poly<FMA>(gt_mul(xr, gt_sign(n)), gt_mul(xr, xr), (n & 2) != 0, ((n ^ ODD_BASE) & 1) != 0);
    }
    if (top < 0x7f8u) {
        int n;
        const double xr = gt_reduce_large(xi, &n);
        const int q = n + (int)(xi >> 31);                          // the original sign moves the quadrant
        return gt_poly<FMA>(gt_mul(xr, gt_sign(q)), gt_mul(xr, xr), (q & 2) != 0, ((n ^ ODD_BASE) & 1) != 0);
    }
    return gt_invalid(y);
}

template <bool FMA> GT_HD float sinf_as_glibc(float y) { return gt_sincosf<FMA, 0>(y); }
template <bool FMA> GT_HD float cosf_as_glibc(float y) { return gt_sincosf<FMA, 1>(y); }

}  // namespace glibc_trig
}  // namespace mcl
