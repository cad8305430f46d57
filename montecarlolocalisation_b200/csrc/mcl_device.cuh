// mcl_device.cuh — device-side helpers shared by the REF and NS kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "glibc_trigf.cuh"

namespace mcl {

// ---- programmatic dependent launch ---------------------------------------------------------------------------------
// The kernels of a filter tick are short and strictly dependent, so the gap between one kernel's last block and the next
// kernel's first is a visible share of the tick. Kernels launched with LAUNCH_PDL (engine_internal.hpp) may be made
// resident while their predecessor drains; pdl_enter() is their first statement: it waits until every earlier kernel of
// the stream has completed and its writes are visible (a no-op under an ordinary launch), then lets the next kernel of
// the stream be made resident in turn. Nothing may touch global memory before it.
__device__ __forceinline__ void pdl_enter() {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

// ---- float trig -------------------------------------------------------------------------------------
// The reference calls cosf/sinf (MC:644-645) and Eigen's fp32 cos/sin (MC:747-748, which oracle/shim maps to the same libm
// calls). `trig` selects what the kernels evaluate (mcl_config.trig_mode resolved by Engine::open):
//   TRIG_GLIBC_FMA / TRIG_GLIBC_SSE2   glibc's sinf/cosf operation for operation in f64 (glibc_trigf.cuh): bit-identical to
//                                      the reference binary on a host whose libm selected that build;
//   TRIG_CR                            "correctly rounded fp32": evaluate in f64 (error ~1e-16) and round once. Differs from
//                                      a correctly rounded result only when the f64 value falls within ~1e-16 of an fp32
//                                      rounding boundary (probability ~1e-9 per call).
constexpr int TRIG_GLIBC_FMA = 0, TRIG_GLIBC_SSE2 = 1, TRIG_CR = 2;
__device__ __forceinline__ float cr_cosf(float t) { return __double2float_rn(cos((double)t)); }
__device__ __forceinline__ float cr_sinf(float t) { return __double2float_rn(sin((double)t)); }
__device__ __forceinline__ float ref_cosf(float t, int trig) {
    return trig == TRIG_GLIBC_FMA ? glibc_trig::cosf_as_glibc<true>(t) : trig == TRIG_GLIBC_SSE2 ? glibc_trig::cosf_as_glibc<false>(t) : cr_cosf(t);
}
__device__ __forceinline__ float ref_sinf(float t, int trig) {
    return trig == TRIG_GLIBC_FMA ? glibc_trig::sinf_as_glibc<true>(t) : trig == TRIG_GLIBC_SSE2 ? glibc_trig::sinf_as_glibc<false>(t) : cr_sinf(t);
}

// ---- exact f64 building blocks (never contracted into FMA) -----------------------------------------------
__device__ __forceinline__ double dmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double dadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double dsub(double a, double b) { return __dadd_rn(a, -b); }
__device__ __forceinline__ double ddiv(double a, double b) { return __ddiv_rn(a, b); }

// static_cast<int>(double) as x86-64 does it (cvttsd2si): truncate toward zero; NaN and out-of-range give
// INT_MIN, which then fails the reference's bounds test (MC:307-309).
__device__ __forceinline__ int trunc_x86(double v) {
    return (fabs(v) < 2147483648.0) ? __double2int_rz(v) : INT32_MIN;
}

// (int)((a) / res) with the quotient correctly rounded first, as the CPU computes it (MC:304-305).
// Fast path: multiply by the rounded reciprocal (|error| <= 3.4e-16*|q|) and accept its truncation unless
// the product lies within 4.5e-16*|q| of an integer; only then pay for the IEEE division.
__device__ __forceinline__ int cell_of(double a, double res, double inv_res) {
    double q = dmul(a, inv_res);
    double t = trunc(q);
    double f = fabs(q - t);
    double tol = fabs(q) * 4.5e-16;
    if (f > tol && f < 1.0 - tol && fabs(q) < 2147483000.0) return __double2int_rz(t);
    return trunc_x86(ddiv(a, res));
}

// ---- Philox4x32-10 (counter-based; Salmon et al. 2011) ---------------------------------------------------
struct Philox {
    static constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
    __host__ __device__ static inline void round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
        // one widening multiply per product (IMAD.WIDE.U32 on the device) instead of a high and a low one
        const uint64_t p0 = (uint64_t)M0 * c[0], p1 = (uint64_t)M1 * c[2];
        const uint32_t hi0 = (uint32_t)(p0 >> 32), hi1 = (uint32_t)(p1 >> 32);
        const uint32_t lo0 = (uint32_t)p0, lo1 = (uint32_t)p1;
        uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    }
    // counter = (c0,c1,c2,c3), key = (k0,k1); out = 4 x 32 random bits
    __host__ __device__ static inline void gen(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                                               uint32_t (&out)[4]) {
        uint32_t c[4] = {c0, c1, c2, c3};
#pragma unroll
        for (int i = 0; i < 10; i++) { round(c, k0, k1); k0 += W0; k1 += W1; }
        out[0] = c[0]; out[1] = c[1]; out[2] = c[2]; out[3] = c[3];
    }
};
// 53-bit canonical double in [0,1) from two 32-bit words.
__host__ __device__ inline double canonical53(uint32_t hi, uint32_t lo) {
    uint64_t v = (((uint64_t)hi << 32) | lo) >> 11;
    return (double)v * (1.0 / 9007199254740992.0);
}

// ---- block reductions ------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return v;
}

// ---- packed fp32 pairs (sm_100 FFMA2 / FADD2): two independent IEEE fp32 operations on a 64-bit register pair per issue slot
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float lo, float hi) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void upk2(f32x2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ void upk2u(f32x2 v, uint32_t& lo, uint32_t& hi) { asm("mov.b64 {%0, %1}, %2;" : "=r"(lo), "=r"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) { f32x2 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }

}  // namespace mcl
