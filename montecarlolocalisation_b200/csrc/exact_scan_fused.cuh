// exact_scan_fused.cuh — the bit-exact sequential f64 accumulation (exact_scan_core.cuh) as ONE kernel per accumulation.
//
// exact_scan.cuh runs an accumulation as three to five dependent launches (tile sums, classification scan, chain, apply scan,
// fallback), each re-reading the weights; at 1M particles the launches cost more than the arithmetic. Here a tile of 4096
// weights (16 per thread) stays in the registers of its block through all four stages and the stages of DIFFERENT tiles are
// ordered by per-tile published words instead of kernel boundaries:
//
//   1. tile sum -> published
//      wait for the sums of every lower tile; P~ at the tile's two edges = F(t), F(t + 1), a FIXED-association sum of
//      tsum[0 .. t) that every block evaluates identically (so neighbouring tiles agree on the value at their shared edge)
//   2. classify by THREAD (16 consecutive weights): a thread whose first item's predecessor and whose last item lie safely
//      inside one binade E holds only PAR items of that binade (the sums are non-decreasing): their increments of the integer
//      significand are RN(w / 2^(E-52)), a pair (even, odd) only for exact ties - a handful of independent f64 operations
//      per item. A thread that straddles a binade edge or the start of the sum, or holds a negative, NaN, Inf or huge weight,
//      becomes one SEQ block: its 16 weights go through the hardware adder in order (~20 such threads per million weights).
//      Segmented parity-monoid scan of the thread aggregates -> SEQ blocks + tile summary, published
//      wait for the summaries of every lower tile; segmented scan of them -> the composite carried into this tile and
//      the SEQ blocks below it, in order
//   3. one thread walks the SEQ blocks of the lower tiles and of this tile with the hardware adder (every block does this
//      redundantly, and no block waits for another block's walk)
//   4. apply: exact s_i for the tile's own elements from the registers -> CDF (or just the total), and - for the CDF that
//      resampling searches - the guide table of that search, scattered as the values are written (what k_ref_guide did as a
//      launch of its own, re-reading the CDF). What a PAR thread writes is VERIFIED, not predicted: the value entering it must
//      lie in binade E and its last value below 2^(E+1).
//
// A block only ever waits for lower tiles. When the whole grid fits on the device at once (the host checks: 444 tiles = 1.8M
// weights on a B200) a tile is simply its block index; otherwise tiles are taken in ticket order, so that a block only waits
// for tiles that are already running: any number of tiles, no co-residency requirement. What a tile publishes is
// self-validating 64-bit words (data and "ready" in one load, no fence on the polling path), which the last block to finish
// leaves empty for the next launch. The last block to finish also runs the single-chain fallback if any prediction could not
// be trusted, so correctness never rests on the margin analysis, and (optionally) advances the adaptive-injection state from
// the total (what k_ref_ema did as a launch of its own).
//
// The body is the device function xsf_run<CDF, ITEMS>; the kernels around it: k_xs_fused<CDF> below (ITEMS = 16, any number of
// tiles) and kernels_ref.cuh's k_ref_scans_one_tile<ITEMS>, which runs the total and then the CDF of a filter of one tile in a
// single launch by calling the body twice (ITEMS = 4 / 8 / 16 by size: the fewer weights, the fewer per thread).
//
// Where the time goes (profiles/r2_xs_trace.txt, 1M weights, 245 tiles, ~22 us): every stage is a latency, not a throughput:
// load + tile sum 2.8 us, first exchange 1.2, classification 1.5 (3 for the handful of tiles that hold a SEQ block or a tie,
// and every higher tile waits for those), second exchange + scan of 244 summaries ~4, block fetch 1.4, walk 3.7, apply 3.
#pragma once
#include "exact_scan.cuh"

namespace mcl {
namespace xs {

constexpr int XSF_ITEMS = 16;                // per thread: 4 x float4 (a 1M-weight accumulation is 245 tiles: one wave at 3 CTAs per SM, and
                                             // one round of the block-wide scan over the lower tiles' summaries)
constexpr int XSF_TILE = XS_THREADS * XSF_ITEMS;
constexpr int XSF_WALK = 96;                 // SEQ blocks below a tile that its block can walk (more: fallback)
constexpr int XSF_BLOCKS = 48;               // SEQ blocks per tile (more: fallback); the summary words carry the count in 8 bits
constexpr int XSF_BSTRIDE = 64;              // blocks reserved per tile in global memory
constexpr int XSF_PSTRIDE = 16;              // u64 words between published records: one 128-byte line each, so that the polls of
                                             // all tiles (every tile reads every lower tile's words) spread over the L2 slices
constexpr int XSF_MAX_TILES = 1 << 13;       // beyond (16M weights) the multi-launch form takes over

struct FusedWs {
    unsigned long long* pub;     // [16 nt]  one line per tile: word 0 the tile sum, words 2-3 the tile summary (self-validating)
    struct SeqBlock* blocks;     // [nt * 16]  SEQ blocks of every tile, written before the tile's summary words
    unsigned* counters;          // [0] tile ticket, [1] finished blocks (both left at 0 by the last block), [2] == epoch: fall back
    unsigned long long* trace;   // null, or [nt][16] %globaltimer stamps at the stage boundaries of every tile (mcl_debug_exact_scan_trace)
    const int* abort;            // optimistic tick (kernels_ref.cuh: RefParams::abort): non-null and set = return at once
    int by_index;                // 1: tile = blockIdx.x (the host checked that the whole grid fits on the device at once, so every
                                 // block a block waits for is running or done); 0: tiles are taken in ticket order
};
// adaptive injection (MC:469-492) advanced by the accumulation that produces the total (mcl_step); inj == null: not asked
struct FusedEma {
    double* inj;             // {weight_slow, weight_fast, p_inject, cdf_is_monotone, total}
    int* counters;           // resampling counters [0..3], cleared for the resampling that follows
    double n, a_slow, a_fast;
};

// guide table of the CDF search (kernels_ref.cuh: k_ref_guide, which this replaces when the CDF comes from the one-kernel form):
// table[b] = first i with cdf[i] >= b / buckets for b = 0..buckets, n beyond the last CDF value; table == null: not asked
struct FusedGuide {
    int* table;
    int buckets, log2_buckets;   // a power of two
    int force_fallback;          // tests: take the single-chain fallback (which also rebuilds the table) whatever the predictions say
};
constexpr int XSF_GQ = 64;       // bucket-edge ranges too long for one thread, per tile, that the whole block writes (more: the thread loops)
__device__ __forceinline__ int xsf_guide_floor(double c, double B, int buckets) {          // = ref_guide_floor
    const double x = c * B;
    return !(x >= 0.0) ? -1 : (x >= B ? buckets : (int)x);           // NaN owns nothing
}

// edges lo..hi answer `val`; ranges too long for one thread go to the block-wide queue (gq: [XSF_GQ][3], *gq_n entries asked for)
__device__ __noinline__ void xsf_guide_range(int* __restrict__ table, int lo, int hi, int val, int* gq, int* gq_n) {
    if (hi - lo + 1 > 8) {                                 // a particle holding a large share of the weight
        const int slot = atomicAdd(gq_n, 1);
        if (slot < XSF_GQ) { gq[3 * slot] = lo; gq[3 * slot + 1] = hi; gq[3 * slot + 2] = val; return; }
    }
    for (int b = lo; b <= hi; b++) table[b] = val;
}

__device__ __forceinline__ unsigned long long xsf_now() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#define XSF_STAMP(k) do { if (ws.trace != nullptr && tid == 0) ws.trace[(size_t)t * 16 + (k)] = xsf_now(); } while (0)

template <int ITEMS>
__device__ __forceinline__ void load_items12(const float* __restrict__ w, int64_t base, int64_t n, float (&x)[ITEMS]) {
    if (base + ITEMS <= n) {
        const float4* p = reinterpret_cast<const float4*>(w + base);          // base is a multiple of ITEMS (4, 8 or 16) floats
#pragma unroll
        for (int q = 0; q < ITEMS / 4; q++) { const float4 a = __ldg(p + q); x[4 * q] = a.x; x[4 * q + 1] = a.y; x[4 * q + 2] = a.z; x[4 * q + 3] = a.w; }
    } else {
#pragma unroll
        for (int j = 0; j < ITEMS; j++) x[j] = (base + j < n) ? w[base + j] : 0.f;
    }
}

// Deterministic block-wide sum of one double per thread: warp trees (fixed shuffle pattern), then the warp totals in order.
__device__ __forceinline__ double block_sum_fixed(double v, double* smem8) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = dadd(v, __shfl_down_sync(0xffffffffu, v, o));
    __syncthreads();
    if ((threadIdx.x & 31) == 0) smem8[threadIdx.x >> 5] = v;
    __syncthreads();
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < XS_THREADS / 32; k++) s = dadd(s, smem8[k]);
    return s;
}

// The single-chain fallback inside the fused kernel: s_i = s_{i-1} + (double)w_i left to right by one thread while the rest
// of the block stages the next 1024 terms (normalised on the fly when the accumulation is the CDF's).
template <bool CDF>
__device__ __forceinline__ void fused_sequential(const float* __restrict__ w, int64_t n, double divisor, double* __restrict__ cdf,
                                                 double* __restrict__ total_out, float (*tile)[XS_SEQ_TILE], double (*outb)[XS_SEQ_TILE]) {
    double acc = 0.0;
    const int64_t n_tiles = (n + XS_SEQ_TILE - 1) / XS_SEQ_TILE;
    auto term = [&](int64_t g) -> float {
        if (g >= n) return 0.f;
        const float x = w[g];
        return CDF ? __double2float_rn(ddiv((double)x, divisor)) : x;
    };
    for (int i = threadIdx.x; i < XS_SEQ_TILE; i += blockDim.x) tile[0][i] = term(i);
    __syncthreads();
    for (int64_t t = 0; t < n_tiles; t++) {
        const int cur = t & 1;
        if (threadIdx.x == 0) {
            const int64_t cnt = min((int64_t)XS_SEQ_TILE, n - t * XS_SEQ_TILE);
            for (int i = 0; i < cnt; i++) { acc = dadd(acc, (double)tile[cur][i]); if (CDF) outb[cur][i] = acc; }
        } else {
            if (CDF && t > 0) {
                const int64_t base = (t - 1) * XS_SEQ_TILE;
                for (int i = threadIdx.x - 1; i < XS_SEQ_TILE; i += blockDim.x - 1) { const int64_t g = base + i; if (g < n) cdf[g] = outb[cur ^ 1][i]; }
            }
            if (t + 1 < n_tiles) {
                const int64_t base = (t + 1) * XS_SEQ_TILE;
                for (int i = threadIdx.x - 1; i < XS_SEQ_TILE; i += blockDim.x - 1) tile[cur ^ 1][i] = term(base + i);
            }
        }
        __syncthreads();
    }
    if (CDF) {
        const int64_t base = (n_tiles - 1) * XS_SEQ_TILE;
        const int cur = (n_tiles - 1) & 1;
        for (int i = threadIdx.x; i < XS_SEQ_TILE; i += blockDim.x) { const int64_t g = base + i; if (g < n) cdf[g] = outb[cur][i]; }
    }
    if (threadIdx.x == 0 && total_out) *total_out = acc;
}

// A thread's 16 weights that go through the hardware adder in order, with the composite of the PAR items between the previous
// SEQ block of the tile (or the tile's start) and the block, and the predicted binade of the value that composite applies to.
template <int ITEMS>
struct SeqBlockT {
    uint32_t idx0;
    int n_items;
    Par pre;
    int first_in_tile, E_prev;
    float w[ITEMS];
};
struct SeqBlock : SeqBlockT<XSF_ITEMS> {};          // the form that travels through global memory between tiles
constexpr int XSF_BPIECES = (int)sizeof(SeqBlock) / 16;
static_assert(sizeof(SeqBlock) == 32 + 4 * XSF_ITEMS && sizeof(SeqBlock) % 16 == 0, "published as 16-byte words");
static_assert(sizeof(SeqBlockT<4>) % 16 == 0 && sizeof(SeqBlockT<8>) % 16 == 0, "read back as 16-byte words");

// published words, one 128-byte line per tile: pub[16 t] = the bits of tile t's f64 sum (never the EMPTY NaN pattern);
// pub[16 t + 2], pub[16 t + 3] = the tile summary, word = composite component (clamped to 2^58: anything that large means a
// wrong prediction and fails verification) | four bits of the SEQ block count << 59 | valid << 63. The last block of every
// launch leaves all of them EMPTY, so a word is data and "ready" flag in one: polling needs no second load and no fence.
constexpr unsigned long long XSF_EMPTY1 = ~0ull;
constexpr unsigned long long XSF_VALID = 1ull << 63;
constexpr unsigned long long XSF_VMASK = (1ull << 59) - 1, XSF_VCLAMP = 1ull << 58;
__device__ __forceinline__ unsigned long long xsf_pack(unsigned long long v, unsigned cnt4) {
    return (v < XSF_VCLAMP ? v : XSF_VCLAMP) | ((unsigned long long)(cnt4 & 15u) << 59) | XSF_VALID;
}
// (publishing with red.and/red.or instead of plain stores, and polling with ld.volatile or ld.relaxed.sys instead of
// ld.relaxed.gpu, were measured: no difference; a summary is seen by every poller ~0.5 us after it is stored)
__device__ __forceinline__ void ld_relaxed_u64x2(const unsigned long long* p, unsigned long long& a, unsigned long long& b) {
    asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
}
__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// one raw block of shared memory, two lives: [SEQ blocks below this tile | this tile's own | exact values after every item of the
// own blocks] while the tile is processed, the single-chain fallback's staging buffers in the last block to finish
constexpr int XSF_WALK_BYTES = XSF_WALK * (int)sizeof(SeqBlock), XSF_OWN_BYTES = XSF_BLOCKS * (int)sizeof(SeqBlock),
              XSF_OWNS_BYTES = XSF_BLOCKS * XSF_ITEMS * (int)sizeof(double);
constexpr int XSF_FB_BYTES = 2 * XS_SEQ_TILE * (int)sizeof(float) + 2 * XS_SEQ_TILE * (int)sizeof(double);
constexpr int XSF_RAW_BYTES = XSF_WALK_BYTES + XSF_OWN_BYTES + XSF_OWNS_BYTES > XSF_FB_BYTES ? XSF_WALK_BYTES + XSF_OWN_BYTES + XSF_OWNS_BYTES : XSF_FB_BYTES;

// The body of the kernels below. single: the accumulation is ONE tile handled by this block alone (no ticket, nobody to wait
// for, the block is its own "last block to finish"), which lets k_xs_both run the total and the CDF back to back in one launch.
template <bool CDF, int ITEMS>
__device__ __forceinline__ void xsf_run(unsigned char* __restrict__ sm_raw, const bool single, const float* __restrict__ w, const int64_t n, const int nt,
                                        const unsigned epoch, const FusedWs& ws, const double* __restrict__ divisor, double* __restrict__ cdf_out,
                                        double* __restrict__ total_out, const FusedEma& ema, const FusedGuide& guide) {
    // ITEMS weights per thread: 16 in general; 4 or 8 when a lone tile is small (single), so that its few thousand weights spread
    // over more threads and every per-thread chain (increments, SEQ blocks, apply) is that much shorter
    using SB = SeqBlockT<ITEMS>;
    constexpr int TILE = XS_THREADS * ITEMS;
    constexpr int BPIECES = (int)sizeof(SB) / 16;
    __shared__ double sm_d[8];
    __shared__ double sm_last[8];
    __shared__ unsigned long long sm_u[8];
    __shared__ ScanState sm_st[8];
    __shared__ ScanState sm_carry_in, sm_tile_state;
    __shared__ int sm_tile, sm_fail, sm_is_last;
    __shared__ int sm_gq[XSF_GQ * 3];                                  // (first edge, last edge, element) ranges for the whole block
    __shared__ int sm_gq_n;
    __shared__ double sm_own_pre[XSF_BLOCKS];                           // exact value entering every SEQ block of this tile
    __shared__ double sm_start;                                        // exact value after the last SEQ block below this tile
    __shared__ int sm_start_valid;                                     // 0: no SEQ block below this tile (the sum so far is exactly 0)
    SB* const sm_walk = reinterpret_cast<SB*>(sm_raw);
    SB* const sm_own = reinterpret_cast<SB*>(sm_raw + XSF_WALK_BYTES);
    double* const sm_own_s = reinterpret_cast<double*>(sm_raw + XSF_WALK_BYTES + XSF_OWN_BYTES);      // [block][item]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool by_index = single || ws.by_index != 0;
    if (tid == 0) {
        if (!by_index) sm_tile = (int)atomicAdd(ws.counters + 0, 1u);
        sm_fail = 0; sm_gq_n = 0;
        sm_carry_in.v = par_identity(); sm_carry_in.reset = 0; sm_carry_in.cnt = 0;
    }
    if (!by_index) __syncthreads();             // (by index: the block-wide scans below synchronise long before those are read)
    const int t = single ? 0 : (ws.by_index ? (int)blockIdx.x : sm_tile);
    XSF_STAMP(0);
    if (guide.force_fallback && t == 0 && tid == 0) atomicExch(ws.counters + 2, epoch);
    // guide table: element i answers the bucket edges in (cdf[i-1], cdf[i]] (k_ref_guide's rule); `pf` carries floor(cdf[i-1] * buckets)
    const bool want_guide = CDF && guide.table != nullptr;
    const double GB = (double)guide.buckets;
    auto guide_emit = [&](int lo, int hi, int val) {
        if (hi == lo) guide.table[hi] = val;
        else if (hi > lo) xsf_guide_range(guide.table, lo, hi, val, sm_gq, &sm_gq_n);
    };
    auto guide_item = [&](int64_t i, double v, int& pf) {          // (the SEQ-block and all-zero threads; PAR threads shift integers)
        const int hi = xsf_guide_floor(v, GB, guide.buckets);
        guide_emit(pf + 1, hi, (int)i);
        pf = hi;
    };
    const int64_t base = (int64_t)t * TILE + (int64_t)tid * ITEMS;
    float x[ITEMS];
    load_items12(w, base, n, x);
    if (CDF) {
        // x = (float)((double)w / total) (MC:497,503) without a division per element: r = w * (1 / total) is within 1.5 ulp of
        // the quotient, so the quotient and its correctly rounded double both lie in [r (1 - 2^-50), r (1 + 2^-50)]; when the two
        // ends round to the same float that float is the answer, otherwise (~1e-8 of the elements, and NaN) the division decides
        const double tot = single ? __ldcg(divisor) : *divisor;      // (single: written a moment ago by this very block)
        const double inv = ddiv(1.0, tot);
#pragma unroll
        for (int j = 0; j < ITEMS; j++) {
            const double r = dmul((double)x[j], inv);
            const float lo = __double2float_rn(dmul(r, 1.0 - 0x1p-50)), hi = __double2float_rn(dmul(r, 1.0 + 0x1p-50));
            x[j] = (base + j < n) ? (lo == hi ? lo : __double2float_rn(ddiv((double)x[j], tot))) : 0.f;
        }
    }
    // ---- stage 1: tile sum, then P~ at the tile's edges from the lower tiles' sums --------------------------------------------
    double s = 0.0;
#pragma unroll
    for (int j = 0; j < ITEMS; j++) s = dadd(s, (double)x[j]);
    const double incl = block_scan_incl(s, sm_d);
    double excl_thr = __shfl_up_sync(0xffffffffu, incl, 1);
    if (lane == 31) sm_last[warp] = incl;
    __syncthreads();
    if (lane == 0) excl_thr = warp ? sm_last[warp - 1] : 0.0;
    const double tile_sum = sm_last[XS_THREADS / 32 - 1];
    if (tid == 0) {
        unsigned long long b = (unsigned long long)__double_as_longlong(tile_sum);
        if (tile_sum != tile_sum) b = 0x7ff8000000000000ull;          // any NaN: the canonical one (never the EMPTY pattern)
        st_relaxed_u64(ws.pub + (size_t)t * XSF_PSTRIDE, b);
    }
    XSF_STAMP(1);
    double part = 0.0;
    for (int k0 = tid; k0 < t; k0 += 2 * XS_THREADS) {                  // two lower tiles per round trip
        const int k1 = k0 + XS_THREADS;
        const unsigned long long* p0 = ws.pub + (size_t)k0 * XSF_PSTRIDE;
        const unsigned long long* p1 = ws.pub + (size_t)k1 * XSF_PSTRIDE;
        unsigned long long v0 = ld_relaxed_u64(p0), v1 = k1 < t ? ld_relaxed_u64(p1) : 0ull;
        while (v0 == XSF_EMPTY1) v0 = ld_relaxed_u64(p0);              // (plain spinning: __nanosleep oversleeps by microseconds here)
        while (v1 == XSF_EMPTY1) v1 = ld_relaxed_u64(p1);
        part = dadd(part, __longlong_as_double((long long)v0));
        if (k1 < t) part = dadd(part, __longlong_as_double((long long)v1));
    }
    const double toff = block_sum_fixed(part, sm_d);                                        // F(t)
    if (tid == (t % XS_THREADS)) part = dadd(part, tile_sum);
    const double toff_next = block_sum_fixed(part, sm_d);                                   // F(t + 1): the same value tile t + 1 computes as its F
    XSF_STAMP(2);
    const int64_t tile_end = min(n, (int64_t)(t + 1) * TILE);       // one past the last valid element of this tile
    // P~ after this thread's last item, and after the previous thread's (tile edges are shared values)
    double my_last = dadd(toff, dadd(excl_thr, s));
    if (base + ITEMS - 1 >= tile_end - 1 && base <= tile_end - 1) my_last = toff_next;
    double prev_last = __shfl_up_sync(0xffffffffu, my_last, 1);
    __syncthreads();
    if (lane == 31) sm_last[warp] = my_last;
    __syncthreads();
    if (lane == 0) prev_last = warp ? sm_last[warp - 1] : toff;
    const uint64_t depth = 2ull * (uint64_t)((nt + 31) / 32) + 64;
    // ---- stage 2: thread kinds, increments, segmented parity-monoid scan ---------------------------------------------------------
    const int n_mine = (int)max((int64_t)0, min((int64_t)ITEMS, n - base));              // valid items of this thread
    const uint64_t thr_margin = margin_for((uint64_t)min(n, base + ITEMS), depth);
    const Pred q0 = predict(prev_last, thr_margin), q1 = predict(my_last, thr_margin);
    // P~ exactly zero after the thread and its own weights all (+-)0: every weight so far is zero, so are the sums. (Earlier
    // weights that cancel, +1 then -1, sit in a SEQ block - a negative weight is never PAR - and the apply stage checks
    // that no block precedes a zero thread.)
    bool own_zero = true;
#pragma unroll
    for (int j = 0; j < ITEMS; j++) own_zero &= (x[j] == 0.f);
    const bool zero = n_mine == 0 || (my_last == 0.0 && own_zero);
    const int thr_E = q0.E;
    bool easy = !zero && q0.ok && q1.ok && q0.E == q1.E && thr_E >= -900 && thr_E <= 900;
    const double fscale = __longlong_as_double((long long)(1023 + 52 - (easy ? thr_E : 0)) << 52);          // 2^(52 - E)
    // increment of the significand by weight xj in binade E: even = for an even significand, odd = for an odd one (they differ
    // by one for an exact tie, MC's round-to-nearest-even); invalid: negative, NaN, Inf or too large for the binade
    auto increment = [&](float xj, unsigned long long& odd, bool& invalid) -> unsigned long long {
        const double y = dmul((double)xj, fscale);
        const double z = dadd(y, 0x1p52);                                      // RN-even(y) in the low significand bits
        const double r = dsub(y, dsub(z, 0x1p52));                             // y - RN(y), exact
        invalid = !(y >= 0.0) || !(y < 0x1p52);
        const unsigned long long even = (unsigned long long)__double_as_longlong(z) & ((1ull << 52) - 1);
        odd = r == 0.5 ? even + 1 : (r == -0.5 ? even - 1 : even);
        return even;
    };
    unsigned long long excl_d = 0, thr_d = 0;
    bool has_tie = false;
    Par thr_par = par_identity();
    if (easy) {
        bool any_invalid = false;
#pragma unroll
        for (int j = 0; j < ITEMS; j++) {
            unsigned long long o; bool inv;
            const unsigned long long e = increment(x[j], o, inv);              // (items past n are +0: nothing)
            any_invalid |= inv; has_tie |= (o != e);
            thr_d += e;
        }
        easy = !any_invalid;
        if (easy && has_tie) {                                                 // the thread's composite depends on the parity it starts from
            // (rare, but a tile with a tie anywhere publishes its summary only after this and every higher tile waits for it: the
            // weights go through a local array, touched in this branch only, instead of a register select per item. Folding the
            // ties into the pass above was measured: the extra branch there slows every tile by more than this costs one)
            float xl[ITEMS];
#pragma unroll
            for (int j = 0; j < ITEMS; j++) xl[j] = x[j];
#pragma unroll 1
            for (int j = 0; j < ITEMS; j++) {
                unsigned long long o; bool inv;
                Par f;
                f.e = increment(xl[j], o, inv); f.o = o;
                thr_par = par_compose(thr_par, f);
            }
        }
    }
    const bool is_block = !zero && !easy;
    const bool fast = !__syncthreads_or(!easy || has_tie);          // every thread easy and tie-free: the tile is one u64 prefix sum
    ScanState pre;
    pre.v = par_identity(); pre.reset = 0; pre.cnt = 0;
    ScanState agg;
    agg.v = par_identity(); agg.reset = 0; agg.cnt = 0;
    if (fast) {
        unsigned long long incl_d = thr_d;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long up = __shfl_up_sync(0xffffffffu, incl_d, o);
            if (lane >= o) incl_d += up;
        }
        if (lane == 31) sm_u[warp] = incl_d;
        __syncthreads();
        unsigned long long woff = 0, whole = 0;
#pragma unroll
        for (int k = 0; k < XS_THREADS / 32; k++) { if (k < warp) woff += sm_u[k]; whole += sm_u[k]; }
        excl_d = woff + incl_d - thr_d;
        if (tid == XS_THREADS - 1) {
            const unsigned long long D = whole < SAT ? whole : SAT;
            sm_tile_state.v.e = D; sm_tile_state.v.o = D; sm_tile_state.reset = 0; sm_tile_state.cnt = 0;
            st_relaxed_u64(ws.pub + (size_t)t * XSF_PSTRIDE + 2, xsf_pack(D, 0));
            st_relaxed_u64(ws.pub + (size_t)t * XSF_PSTRIDE + 3, xsf_pack(D, 0));
        }
    } else {
        XSF_STAMP(10);
        if (easy) agg.v = has_tie ? thr_par : Par{thr_d < SAT ? thr_d : SAT, thr_d < SAT ? thr_d : SAT};
        else if (is_block) { agg.reset = 1; agg.cnt = 1; }
        // block-wide exclusive scan of the thread aggregates
        ScanState inc = agg;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            ScanState up = st_shfl_up(inc, o);
            if (lane >= o) inc = st_combine(up, inc);
        }
        if (lane == 31) sm_st[warp] = inc;
        __syncthreads();
        for (int k = 0; k < warp; k++) pre = st_combine(pre, sm_st[k]);
        {
            ScanState lane_excl = st_shfl_up(inc, 1);
            if (lane > 0) pre = st_combine(pre, lane_excl);
        }
        // SEQ blocks out: rank = blocks before this thread; composite = the PAR items since the previous block / the tile's start
        XSF_STAMP(11);
        if (is_block) {
            if (pre.cnt < XSF_BLOCKS) {
                SB* const d = sm_own + pre.cnt;              // built in shared memory, copied out by the same thread as 16-byte words
                d->idx0 = (uint32_t)base; d->n_items = n_mine; d->pre = pre.v; d->first_in_tile = pre.reset ? 0 : 1; d->E_prev = q0.E;
#pragma unroll
                for (int j = 0; j < ITEMS / 4; j++) reinterpret_cast<float4*>(d->w)[j] = make_float4(x[4 * j], x[4 * j + 1], x[4 * j + 2], x[4 * j + 3]);
                if (!single) {                                     // (a lone tile's blocks are read by nobody else)
                    int4* g = reinterpret_cast<int4*>(reinterpret_cast<SB*>(ws.blocks) + (size_t)t * XSF_BSTRIDE + pre.cnt);
                    const int4* sb4 = reinterpret_cast<const int4*>(d);
#pragma unroll
                    for (int k = 0; k < BPIECES; k++) g[k] = sb4[k];
                    if (ws.trace != nullptr) ws.trace[(size_t)t * 16 + 12] = xsf_now();
                    __threadfence();                               // the block before the summary words
                    if (ws.trace != nullptr) ws.trace[(size_t)t * 16 + 13] = xsf_now();
                }
            } else sm_fail = 1;
        }
        __syncthreads();
        // state of the whole tile = the last thread's inclusive state; publish it
        if (tid == XS_THREADS - 1) {
            const ScanState whole = st_combine(pre, agg);
            sm_tile_state = whole;
            const unsigned cnt = (unsigned)min(whole.cnt, XSF_BLOCKS);
            if (whole.cnt > XSF_BLOCKS) sm_fail = 1;
            st_relaxed_u64(ws.pub + (size_t)t * XSF_PSTRIDE + 2, xsf_pack(whole.v.e, cnt & 15u));
            st_relaxed_u64(ws.pub + (size_t)t * XSF_PSTRIDE + 3, xsf_pack(whole.v.o, cnt >> 4));
        }
    }
    XSF_STAMP(3);
    // The total alone needs no per-element values: only the tile that holds the last element walks and applies. The walk is
    // where the result is verified: every composite between two SEQ blocks is applied to a value that must lie in its binade
    // and must leave the result there, and the threads of such a run all use that one binade (each takes it from the value
    // its predecessor handed on), so with non-decreasing sums every add in between happened in that binade.
    const bool need_rest = CDF || t == nt - 1;
    if (!need_rest && sm_fail) atomicExch(ws.counters + 2, epoch);
    if (need_rest) {
    // ---- stage 3: the composite carried into this tile, and the SEQ blocks below it ---------------------------------------------------
    __shared__ int sm_src[XSF_WALK];                  // walk position -> tile * XSF_BSTRIDE + block
    __shared__ Par sm_cin[XSF_WALK];                  // walk position -> composite carried into that block's tile
    for (int c0 = 0; c0 < t; c0 += XS_THREADS) {
        const int k = c0 + tid;
        ScanState mine;
        mine.v = par_identity(); mine.reset = 0; mine.cnt = 0;
        if (k < t) {
            const unsigned long long* pk = ws.pub + (size_t)k * XSF_PSTRIDE + 2;
            unsigned long long a, b;
            ld_relaxed_u64x2(pk, a, b);
            while (!(a & b & XSF_VALID)) ld_relaxed_u64x2(pk, a, b);
            mine.v.e = a & XSF_VMASK; mine.v.o = b & XSF_VMASK;
            if (mine.v.e >= XSF_VCLAMP) mine.v.e = SAT;
            if (mine.v.o >= XSF_VCLAMP) mine.v.o = SAT;
            mine.cnt = (int)(((a >> 59) & 15ull) | (((b >> 59) & 15ull) << 4));
            mine.reset = mine.cnt ? 1 : 0;
        }
        XSF_STAMP(7);
        if (ws.trace != nullptr) {                // which lower tile's summary this block saw last, and when
            __shared__ unsigned long long sm_tmax;
            const unsigned long long tp = k < t ? xsf_now() : 0ull;       // when this thread's own poll succeeded
            if (tid == 0) sm_tmax = 0;
            __syncthreads();
            atomicMax(&sm_tmax, tp);
            __syncthreads();
            if (tp == sm_tmax) { ws.trace[(size_t)t * 16 + 14] = tp; ws.trace[(size_t)t * 16 + 15] = (unsigned long long)k; }
        }
        ScanState in2 = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            ScanState up = st_shfl_up(in2, o);
            if (lane >= o) in2 = st_combine(up, in2);
        }
        __syncthreads();                          // sm_st / sm_carry_in of the previous chunk consumed
        if (lane == 31) sm_st[warp] = in2;
        __syncthreads();
        ScanState p2 = sm_carry_in;
        for (int q = 0; q < warp; q++) p2 = st_combine(p2, sm_st[q]);
        {
            ScanState lane_excl = st_shfl_up(in2, 1);
            if (lane > 0) p2 = st_combine(p2, lane_excl);            // state entering tile k: p2.cnt = SEQ blocks below it
        }
        for (int q = 0; q < mine.cnt; q++) {      // where tile k's blocks go in walk order, and what is carried into tile k
            const int pos = p2.cnt + q;
            if (pos >= XSF_WALK) { sm_fail = 1; break; }
            sm_src[pos] = k * XSF_BSTRIDE + q;
            sm_cin[pos] = p2.v;
        }
        __syncthreads();
        if (tid == XS_THREADS - 1) sm_carry_in = st_combine(p2, mine);       // state leaving the chunk
        __syncthreads();
    }
    {
        // all threads fetch the blocks below this tile, 16 bytes each (their tiles' summary words were seen, and a fence
        // follows: the blocks written before those words are visible); then the first block of every tile takes in what is
        // carried into its tile
        const int n_below = sm_fail ? 0 : min(sm_carry_in.cnt, XSF_WALK);
        XSF_STAMP(8);
        if (!single) __threadfence();
        XSF_STAMP(9);
        for (int q = tid; q < n_below * BPIECES; q += XS_THREADS) {
            const int f = q / BPIECES, piece = q - f * BPIECES;
            reinterpret_cast<int4*>(sm_walk + f)[piece] = __ldcg(reinterpret_cast<const int4*>(reinterpret_cast<const SB*>(ws.blocks) + sm_src[f]) + piece);
        }
        __syncthreads();
        for (int f = tid; f < n_below; f += XS_THREADS)
            if (sm_walk[f].first_in_tile) { sm_walk[f].pre = par_compose(sm_cin[f], sm_walk[f].pre); sm_walk[f].first_in_tile = 0; }
        __syncthreads();
    }
    XSF_STAMP(4);
    const ScanState carry = sm_carry_in;             // state entering this tile
    const ScanState tstate = sm_tile_state;
    if (tid == 0) {
        bool ok = sm_fail == 0;
        double sv = 0.0;
        if (ok) {
            const int n_below = min(carry.cnt, XSF_WALK);
            // one dependent chain of hardware additions (exactly the reference's roundings); the next block's weights are fetched
            // while the current block's sixteen additions run. Items past a block's count are +0: adding them changes nothing.
            float4 nx[ITEMS / 4];
            if (n_below > 0) {
#pragma unroll
                for (int k = 0; k < ITEMS / 4; k++) nx[k] = reinterpret_cast<const float4*>(sm_walk[0].w)[k];
            }
            for (int q = 0; q < n_below; q++) {
                const SB* b = sm_walk + q;
                float4 cur[ITEMS / 4];
#pragma unroll
                for (int k = 0; k < ITEMS / 4; k++) cur[k] = nx[k];
                if (q + 1 < n_below) {
#pragma unroll
                    for (int k = 0; k < ITEMS / 4; k++) nx[k] = reinterpret_cast<const float4*>(sm_walk[q + 1].w)[k];
                }
                sv = par_apply(sv, b->pre, b->E_prev, ok);
#pragma unroll
                for (int k = 0; k < ITEMS / 4; k++) {
                    sv = dadd(sv, (double)cur[k].x); sv = dadd(sv, (double)cur[k].y); sv = dadd(sv, (double)cur[k].z); sv = dadd(sv, (double)cur[k].w);
                }
            }
            sm_start = sv; sm_start_valid = carry.cnt > 0;
            const int own = min(tstate.cnt, XSF_BLOCKS);
            if (own > 0) {
#pragma unroll
                for (int k = 0; k < ITEMS / 4; k++) nx[k] = reinterpret_cast<const float4*>(sm_own[0].w)[k];
            }
            for (int q = 0; q < own; q++) {                          // (items past a block's count are +0, as above)
                const SB* b = sm_own + q;
                float4 cur[ITEMS / 4];
#pragma unroll
                for (int k = 0; k < ITEMS / 4; k++) cur[k] = nx[k];
                if (q + 1 < own) {
#pragma unroll
                    for (int k = 0; k < ITEMS / 4; k++) nx[k] = reinterpret_cast<const float4*>(sm_own[q + 1].w)[k];
                }
                const Par comp = b->first_in_tile ? par_compose(carry.v, b->pre) : b->pre;
                sv = par_apply(sv, comp, b->E_prev, ok);
                sm_own_pre[q] = sv;
                double* const so = sm_own_s + q * ITEMS;
#pragma unroll
                for (int k = 0; k < ITEMS / 4; k++) {
                    sv = dadd(sv, (double)cur[k].x); so[4 * k] = sv;
                    sv = dadd(sv, (double)cur[k].y); so[4 * k + 1] = sv;
                    sv = dadd(sv, (double)cur[k].z); so[4 * k + 2] = sv;
                    sv = dadd(sv, (double)cur[k].w); so[4 * k + 3] = sv;
                }
            }
        }
        if (!ok) sm_fail = 1;
    }
    __syncthreads();
    XSF_STAMP(5);
    // ---- stage 4: apply ----------------------------------------------------------------------------------------------------------
    bool okflag = sm_fail == 0;
    const double start0 = sm_start_valid ? sm_start : 0.0;
    if (okflag && (fast || easy)) {
        // value after the last SEQ block before this thread, and the composite of the PAR items since then
        double thr_start = start0;
        Par thr_comp = carry.v;
        bool have_start = sm_start_valid != 0;
        if (!fast) {
            if (pre.reset) { thr_start = sm_own_s[min(pre.cnt - 1, XSF_BLOCKS - 1) * ITEMS + ITEMS - 1]; thr_comp = pre.v; have_start = true; }
            else thr_comp = par_compose(carry.v, pre.v);
        }
        // integer significand entering the thread; every item adds its increment
        const unsigned long long sb = (unsigned long long)__double_as_longlong(thr_start);
        const int be = (int)((sb >> 52) & 0x7ff);
        const unsigned long long A_start = (sb & ((1ull << 52) - 1)) | (1ull << 52);
        const unsigned long long csel = (A_start & 1) ? thr_comp.o : thr_comp.e;
        okflag = have_start && !(sb >> 63) && be - 1023 == thr_E && csel < SAT;
        unsigned long long A = A_start + csel + excl_d;
        const unsigned long long hi_bits = (unsigned long long)be << 52;
        // floor(v * buckets) of a value v = A 2^(be - 1075) of this thread's binade is a shift of its integer significand
        const int gsh = min(63, max(0, 1075 - be - guide.log2_buckets));
        int pf = -1;
        if (want_guide && base > 0) pf = (int)min(A >> gsh, (unsigned long long)guide.buckets);
#pragma unroll
        for (int j = 0; j < ITEMS; j++) {
            const int64_t i = base + j;
            if (i >= n) break;
            unsigned long long o; bool inv;
            const unsigned long long e = increment(x[j], o, inv);
            A += (A & 1) ? o : e;
            const double v = __longlong_as_double((long long)(hi_bits | (A & ((1ull << 52) - 1))));
            if (CDF) cdf_out[i] = v;
            if (want_guide) {
                const int hi = (int)min(A >> gsh, (unsigned long long)guide.buckets);
                if (hi != pf) { guide_emit(pf + 1, hi, (int)i); pf = hi; }
            }
            if (i == n - 1 && total_out) *total_out = v;
        }
        if (want_guide && n_mine > 0 && base + n_mine == n) guide_emit(pf + 1, guide.buckets, (int)n);      // edges beyond the last CDF value
        if (A >= (1ull << 53)) okflag = false;           // the thread's last value left the binade: the prediction was wrong
    } else if (okflag && is_block) {
        int pf = -1;
        if (want_guide && base > 0) pf = xsf_guide_floor(sm_own_pre[min(pre.cnt, XSF_BLOCKS - 1)], GB, guide.buckets);
#pragma unroll
        for (int j = 0; j < ITEMS; j++) {
            const int64_t i = base + j;
            if (i >= n) break;
            const double v = sm_own_s[min(pre.cnt, XSF_BLOCKS - 1) * ITEMS + j];
            if (CDF) cdf_out[i] = v;
            if (want_guide) guide_item(i, v, pf);
            if (i == n - 1 && total_out) *total_out = v;
        }
        if (want_guide && n_mine > 0 && base + n_mine == n) guide_emit(pf + 1, guide.buckets, (int)n);
    } else if (okflag) {
        // every weight up to and including this thread's is zero: so are the sums
        if (n_mine > 0 && (carry.cnt != 0 || pre.cnt != 0)) okflag = false;       // (cannot happen: a SEQ block means a non-zero P~ before here)
        int pf = base > 0 ? 0 : -1;
#pragma unroll
        for (int j = 0; j < ITEMS; j++) {
            const int64_t i = base + j;
            if (i >= n) break;
            if (CDF) cdf_out[i] = 0.0;
            if (want_guide) guide_item(i, 0.0, pf);
            if (i == n - 1 && total_out) *total_out = 0.0;
        }
        if (want_guide && n_mine > 0 && base + n_mine == n) guide_emit(pf + 1, guide.buckets, (int)n);
    }
    if (!okflag) atomicExch(ws.counters + 2, epoch);
    if (want_guide) {
        __syncthreads();
        const int nq = min(sm_gq_n, XSF_GQ);
        for (int q = 0; q < nq; q++) {
            const int hi = sm_gq[3 * q + 1], val = sm_gq[3 * q + 2];
            for (int b = sm_gq[3 * q] + tid; b <= hi; b += XS_THREADS) guide.table[b] = val;
        }
    }
    }   // need_rest
    XSF_STAMP(6);
    // ---- the last block to finish: clean-up, fallback if anybody asked for it, then the adaptive-injection state ---------------------
    __threadfence();
    __syncthreads();
    if (!single) {
        if (tid == 0) sm_is_last = atomicAdd(ws.counters + 1, 1u) == (unsigned)nt - 1u;
        __syncthreads();
        if (!sm_is_last) return;
        __threadfence();
    }
    for (int k = tid; k < nt; k += XS_THREADS) {
        unsigned long long* pk = ws.pub + (size_t)k * XSF_PSTRIDE;
        pk[0] = XSF_EMPTY1; pk[2] = 0ull; pk[3] = 0ull;
    }
    if (tid == 0 && !single) { ws.counters[0] = 0; ws.counters[1] = 0; }
    if (*(volatile unsigned*)(ws.counters + 2) == epoch) {
        __syncthreads();                          // every thread is done with what lived in sm_raw
        fused_sequential<CDF>(w, n, CDF ? (single ? __ldcg(divisor) : *divisor) : 1.0, cdf_out, total_out, reinterpret_cast<float(*)[XS_SEQ_TILE]>(sm_raw),
                              reinterpret_cast<double(*)[XS_SEQ_TILE]>(sm_raw + 2 * XS_SEQ_TILE * sizeof(float)));
        __syncthreads();
        if (want_guide) {                         // whatever the tiles scattered came from values that were not trusted: all of it again
            for (int64_t i = tid; i < n; i += XS_THREADS) {
                const int hi = xsf_guide_floor(cdf_out[i], GB, guide.buckets);
                const int lo = i == 0 ? 0 : xsf_guide_floor(cdf_out[i - 1], GB, guide.buckets) + 1;
                for (int b = lo; b <= hi; b++) guide.table[b] = (int)i;
                if (i == n - 1) for (int b = hi + 1; b <= guide.buckets; b++) guide.table[b] = (int)n;
            }
        }
    }
    if (ema.inj != nullptr && tid == 0) {
        // adaptive injection (MC:469-492): the same IEEE operations in the same order as the host form in Engine::ref_resample
        ema.counters[0] = 0; ema.counters[1] = 0; ema.counters[2] = 0; ema.counters[3] = 0;
        const double tv = *(volatile double*)total_out;
        const double avg = ddiv(tv, ema.n);
        const double slow = dadd(ema.inj[0], dmul(ema.a_slow, dsub(avg, ema.inj[0])));
        const double fast_w = dadd(ema.inj[1], dmul(ema.a_fast, dsub(avg, ema.inj[1])));
        const double p = dsub(1.0, ddiv(fast_w, slow));
        ema.inj[0] = slow; ema.inj[1] = fast_w;
        ema.inj[2] = (0.0 < p) ? p : 0.0;                                   // std::max(0.0, p): NaN gives 0.0 (MC:492)
        ema.inj[3] = (tv > 0.0 && tv < 1.0e300) ? 1.0 : 0.0;                // finite positive total: the CDF is non-decreasing
        ema.inj[4] = tv;
    }
}

// CDF == false: *total_out = w_0 + w_1 + ... left to right                                  (MC:675)
// CDF == true : x_i = (float)((double)w_i / *divisor), cdf[i] = x_0 + ... + x_i              (MC:496-505)
template <bool CDF>
__global__ void __launch_bounds__(XS_THREADS, 3) k_xs_fused(const float* __restrict__ w, int64_t n, int nt, unsigned epoch, FusedWs ws,
                                                            const double* __restrict__ divisor, double* __restrict__ cdf_out,
                                                            double* __restrict__ total_out, FusedEma ema, FusedGuide guide) {
    pdl_enter();
    if (ws.abort != nullptr && *ws.abort != 0) return;
    __shared__ __align__(16) unsigned char sm_raw[XSF_RAW_BYTES];
    xsf_run<CDF, XSF_ITEMS>(sm_raw, false, w, n, nt, epoch, ws, divisor, cdf_out, total_out, ema, guide);
}

}  // namespace xs
}  // namespace mcl
