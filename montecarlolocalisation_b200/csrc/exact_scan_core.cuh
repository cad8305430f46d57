// exact_scan_core.cuh — the arithmetic core of the PARALLEL, BIT-EXACT reproduction of a SEQUENTIAL f64 accumulation
//
//      s_i = RN(s_{i-1} + (double)w_i),   s_{-1} = 0,   w_i >= 0 fp32                 (reference: MC:675 and MC:496-505)
//
// fp addition is not associative, so a tree reduction rounds differently from the reference's left-to-right loop and
// would flip resampled indices. This header reproduces the loop's roundings exactly, in parallel:
//
//   1. An ordinary parallel f64 prefix sum gives P~_i with |P~_i - s_i| <= (i + depth) * 2^-52 * s_i. Unless P~_i lies
//      within that margin of a power of two, it tells the BINADE of s_i (the exponent E with s_i in [2^E, 2^(E+1))).
//   2. While s stays inside one binade, ulp(s) = u = 2^(E-52) is constant and s = A*u with a 53-bit integer A.
//      Adding w = q*u + r (0 <= r < u) and rounding to nearest-even gives A' = A + q + c, where c = [r > u/2], or for the
//      tie r == u/2 exactly, c = (A + q) & 1. So every add is an INTEGER increment that depends on A only through its
//      parity, and only for ties: f(A) = A + (A even ? d_even : d_odd). Such maps compose associatively
//      ("parity monoid") and can be prefix-scanned in parallel with integer adds, which are exact.
//   3. The few elements where the binade changes, or where the prediction is within the margin ("SEQ" elements, ~25-100
//      per million), are added one at a time with the hardware f64 adder, by one thread, after the composites of the
//      parallel runs between them are known.
//
// Everything here is plain integer / bit arithmetic, usable on host and device, so the logic is model-checked on the
// CPU (tests/test_exact_scan_model.py) before the kernels in exact_scan.cuh use it.
#pragma once
#include <stdint.h>
#include <string.h>

#ifdef __CUDACC__
#define XS_HD __host__ __device__ __forceinline__
#else
#define XS_HD inline
#endif

namespace mcl {
namespace xs {

// f(A) = A + (A even ? e : o). `reset` marks a SEQ element: the running composite restarts after it.
struct Par {
    uint64_t e, o;
};
constexpr uint64_t SAT = 1ull << 60;      // saturating cap: a composite this large means the binade prediction was wrong

// Every Par component is <= SAT by construction (par_of_weight clamps, par_compose saturates, the identity is 0), so a + b
// cannot wrap and "either operand saturated" implies s >= SAT: the cap is decided on the high word of the sum alone.
XS_HD uint64_t sat_add(uint64_t a, uint64_t b) { uint64_t s = a + b; return (uint32_t)(s >> 32) >= (uint32_t)(SAT >> 32) ? SAT : s; }

XS_HD Par par_identity() { return Par{0, 0}; }
// apply a first, then b
XS_HD Par par_compose(const Par& a, const Par& b) {
    Par r;
    r.e = sat_add(a.e, (a.e & 1) ? b.o : b.e);             // even input -> parity after a is parity(a.e)
    r.o = sat_add(a.o, ((a.o + 1) & 1) ? b.o : b.e);       // odd input  -> parity after a is parity(1 + a.o)
    return r;
}

XS_HD uint64_t f64_bits(double x) { uint64_t b; memcpy(&b, &x, 8); return b; }
XS_HD double bits_f64(uint64_t b) { double x; memcpy(&x, &b, 8); return x; }
XS_HD uint32_t f32_bits(float x) { uint32_t b; memcpy(&b, &x, 4); return b; }

// Binade (unbiased exponent) of a positive normal double.
XS_HD int f64_exponent(double x) { return (int)((f64_bits(x) >> 52) & 0x7ff) - 1023; }
XS_HD uint64_t f64_fraction(double x) { return f64_bits(x) & ((1ull << 52) - 1); }

// Prediction for one prefix value: usable (finite, positive, not within `margin` ulps of a binade edge) or not.
struct Pred {
    int E;          // binade
    bool ok;        // false: zero, ambiguous, NaN/Inf
    bool zero;      // P~ == 0  (every weight so far is zero, so s is exactly 0)
};
XS_HD Pred predict(double p, uint64_t margin) {
    Pred r;
    r.zero = (p == 0.0);
    uint64_t b = f64_bits(p);
    int be = (int)((b >> 52) & 0x7ff);
    r.E = be - 1023;
    uint64_t frac = b & ((1ull << 52) - 1);
    bool finite_pos_normal = (b >> 63) == 0 && be != 0 && be != 0x7ff;
    r.ok = finite_pos_normal && frac >= margin && frac <= (1ull << 52) - margin;
    return r;
}

// The increment of adding fp32 w >= 0 to an accumulator that sits in binade E (unit u = 2^(E-52)).
// Returns false if w cannot be handled here (negative, NaN, Inf): the caller must fall back.
XS_HD bool par_of_weight(float w, int E, Par& out) {
    uint32_t wb = f32_bits(w);
    if (wb >> 31) return wb == 0x80000000u ? (out = par_identity(), true) : false;     // -0.0 adds nothing; negatives: no
    uint32_t bexp = (wb >> 23) & 0xff, mant = wb & 0x7fffffu;
    if (bexp == 0xff) return false;                            // Inf / NaN
    if (bexp == 0 && mant == 0) { out = par_identity(); return true; }
    uint64_t m = bexp ? (uint64_t)(mant | 0x800000u) : (uint64_t)mant;      // w = m * 2^ew
    int ew = bexp ? (int)bexp - 150 : -149;
    int shift = ew - (E - 52);
    if (shift >= 0) {
        uint64_t q = shift >= 36 ? SAT : (m << shift);         // m < 2^24; >= 2^60 can never stay inside the binade
        out.e = out.o = q >= SAT ? SAT : q;
        return true;
    }
    int k = -shift;
    if (k > 25) { out = par_identity(); return true; }         // w < u/4: rounds away entirely
    uint64_t q = m >> k;
    uint64_t rem = m & ((1ull << k) - 1), half = 1ull << (k - 1);
    if (rem > half) { out.e = out.o = q + 1; }
    else if (rem < half) { out.e = out.o = q; }
    else { out.e = q + (q & 1); out.o = q + ((q + 1) & 1); }   // tie: round so that A + q + c is even
    return true;
}

// Apply a composite to an exact accumulator value s that must lie in binade E. ok=false if s is not in that binade or
// the result leaves it (prediction wrong): the caller falls back to the sequential kernel.
XS_HD double par_apply(double s, const Par& f, int E, bool& ok) {
    if (f.e == 0 && f.o == 0) return s;
    uint64_t b = f64_bits(s);
    int be = (int)((b >> 52) & 0x7ff);
    if ((b >> 63) || be - 1023 != E || be == 0 || be == 0x7ff) { ok = false; return s; }
    uint64_t A = (b & ((1ull << 52) - 1)) | (1ull << 52);
    uint64_t d = (A & 1) ? f.o : f.e;
    if (d >= SAT) { ok = false; return s; }
    uint64_t A2 = A + d;
    if (A2 >= (1ull << 53)) { ok = false; return s; }
    return bits_f64(((uint64_t)be << 52) | (A2 & ((1ull << 52) - 1)));
}

// Error margin (in ulps of P~_i) that separates a safe binade prediction from an ambiguous one: the sequential sum
// drifts from the exact sum by at most i half-ulps, the parallel one by at most `depth` half-ulps; doubled for safety.
XS_HD uint64_t margin_for(uint64_t i, uint64_t depth) { return i + depth + 64; }

}  // namespace xs
}  // namespace mcl
