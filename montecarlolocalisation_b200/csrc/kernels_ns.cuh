// kernels_ns.cuh — MCL_MODE_NS kernels (the north-star formulation; definitions in ns_core.cuh and DESIGN.md "NS").
//
//   k_ns_edt_cols / k_ns_edt_rows   exact capped squared Euclidean distance transform (integers) -> likelihood field
//   k_ns_init                       uniform particles from Philox (global index keyed)
//   k_ns_predict                    odometry motion model with per-particle Philox noise, float4 in/out
//   k_ns_update                     likelihood-field sensor model: warp per particle, beams across lanes, field staged in
//                                   shared memory by TMA (cp.async.bulk) when it fits, else read through L2
//   k_ns_weights_sum / k_ns_weights_scan   Q32 fixed-point weights + block-scan prefix sum (integers: order-independent)
//   k_ns_resample                   systematic resampling by 128-bit integer search; each output is stored straight into
//                                   the shard that owns its slot (own memory or a peer GPU's over NVLink)
#pragma once
#include "mcl_device.cuh"
#include "ns_core.cuh"
#include "ns_plan.hpp"

namespace mcl {

// ---- distance transform -------------------------------------------------------------------------------------------------
// pass 1, one thread per column: g[y][x] = rows to the nearest occupied cell of the same column, capped at R+1.
__global__ void k_ns_edt_cols(const uint8_t* __restrict__ occ, int W, int H, int R, uint16_t* __restrict__ g) {
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= W) return;
    int d = R + 1;
    for (int y = 0; y < H; y++) {
        d = occ[(size_t)y * W + x] ? 0 : min(d + 1, R + 1);
        g[(size_t)y * W + x] = (uint16_t)d;
    }
    d = R + 1;
    for (int y = H - 1; y >= 0; y--) {
        d = occ[(size_t)y * W + x] ? 0 : min(d + 1, R + 1);
        size_t i = (size_t)y * W + x;
        if (d < g[i]) g[i] = (uint16_t)d;
    }
}
// pass 2, one thread per cell: d2 = min over |dx| <= R of dx^2 + g^2, capped at R^2; field = table[d2].
// The field is stored with a border of `pad` cells on every side (row pitch Wp = W + 2 pad) pre-filled with the
// outside-the-grid value, so that the sensor-model kernel needs no bounds test for particles inside the map.
__global__ void k_ns_fill_f32(float* __restrict__ dst, size_t n, float v) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) dst[i] = v;
}
__global__ void k_ns_edt_rows(const uint16_t* __restrict__ g, int W, int H, int R, const float* __restrict__ lf_of_d2,
                              uint16_t* __restrict__ d2_out, float* __restrict__ lf_out, int pad) {
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= W) return;
    const int cap = R * R;
    int best = cap;
    const uint16_t* row = g + (size_t)y * W;
    int lo = max(0, x - R), hi = min(W - 1, x + R);
    for (int xx = lo; xx <= hi; xx++) {
        int gy = row[xx];
        if (gy > R) continue;
        int dx = xx - x;
        int v = dx * dx + gy * gy;
        best = min(best, v);
    }
    size_t i = (size_t)y * W + x;
    if (d2_out) d2_out[i] = (uint16_t)best;
    lf_out[(size_t)(y + pad) * (W + 2 * pad) + (x + pad)] = lf_of_d2[best];
}

// ---- init / predict -------------------------------------------------------------------------------------------------------
struct NsMotion {
    float rot1, trans, rot2;          // odometry increment (MC:699-702)
    float sd_rot1, sd_trans, sd_rot2; // sqrt of the reference's noise variances (MC:706-710)
};

// particle i (global index g0 + i): x,y uniform over the map extent, theta uniform in [-pi,pi), weight 1.
__global__ void __launch_bounds__(256) k_ns_init(float4* __restrict__ part, int64_t n, int64_t g0, double ox, double oy,
                                                 double ext_x, double ext_y, uint32_t k0, uint32_t k1) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t g = (uint64_t)(g0 + i);
    uint32_t r[4];
    Philox::gen((uint32_t)g, (uint32_t)(g >> 32), 0x60u, 0u, k0, k1, r);
    const double s = 2.3283064365386963e-10;
    double u1 = ns::mul(ns::add((double)r[0], 0.5), s), u2 = ns::mul(ns::add((double)r[1], 0.5), s), u3 = ns::mul(ns::add((double)r[2], 0.5), s);
    float4 p;
    p.x = (float)ns::add(ox, ns::mul(u1, ext_x));
    p.y = (float)ns::add(oy, ns::mul(u2, ext_y));
    p.z = (float)ns::add(-3.14159265358979323846, ns::mul(u3, 6.28318530717958647692));
    p.w = 1.0f;
    part[i] = p;
}

__global__ void __launch_bounds__(256) k_ns_predict(float4* __restrict__ part, int64_t n, int64_t g0, NsMotion m, uint32_t step,
                                                    uint32_t k0, uint32_t k1) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t g = (uint64_t)(g0 + i);
    uint32_t r[4];
    Philox::gen((uint32_t)g, (uint32_t)(g >> 32), 0x50u, step, k0, k1, r);
    float z0, z1, z2, z3;
    ns::det_normal_pair(r[0], r[1], z0, z1);
    ns::det_normal_pair(r[2], r[3], z2, z3);
    (void)z3;
    float4 p = part[i];
    const float r1 = ns::fmaf_(z0, m.sd_rot1, m.rot1);
    const float tr = ns::fmaf_(z1, m.sd_trans, m.trans);
    const float r2 = ns::fmaf_(z2, m.sd_rot2, m.rot2);
    float s, c;
    ns::det_sincosf(ns::addf(p.z, r1), s, c);
    p.x = ns::fmaf_(tr, c, p.x);
    p.y = ns::fmaf_(tr, s, p.y);
    p.z = ns::wrap_pi(ns::addf(p.z, ns::addf(r1, r2)));
    part[i] = p;
}

// ---- update: likelihood field -----------------------------------------------------------------------------------------------
struct NsField {
    const float* lf;        // [(H + 2 pad) * (W + 2 pad)] log-likelihood per cell, border = lf_out
    int W, H;               // the grid proper
    int pad, Wp;            // border width and row pitch (cells)
    float ox, oy, inv_res;
    float lf_out;           // value for endpoints outside the grid
    int bytes_padded;       // field bytes rounded up to 16 (TMA bulk copy granularity)
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// TMA 1-D bulk copy global -> shared, completion on an mbarrier (PTX cp.async.bulk; SASS UBLKCP).
__device__ __forceinline__ void tma_stage(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    const uint32_t b = smem_u32(bar);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
        const uint32_t CH = 32768;
        for (uint32_t off = 0; off < bytes; off += CH) {
            uint32_t sz = min(CH, bytes - off);
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                             smem_u32((char*)smem_dst + off)),
                         "l"((const char*)gsrc + off), "r"(sz), "r"(b)
                         : "memory");
        }
    }
    // every thread waits for phase 0 of the barrier
    uint32_t done = 0;
    while (!done) {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\nselp.u32 %0, 1, 0, p;\n}"
                     : "=r"(done)
                     : "r"(b)
                     : "memory");
    }
}

// One warp scores 32 particles at a time. Lane l owns particle l's pose: it computes sin/cos and the particle position
// in CELL units (origin and resolution folded in once per particle, and -0.5 so that round-to-nearest gives the cell).
// Particles are then taken four at a time: their poses are broadcast by shuffle and the 32 lanes take beams l, l+32, ...
// (beam points are pre-scaled to cell units on the host), so one shared-memory beam load feeds four evaluations.
// Cell index = round-to-nearest-even via the 1.5*2^23 magic add (FMA pipe) instead of F2I (XU pipe).
// When all 32 particles of a batch lie inside the map (the overwhelmingly common case) no endpoint can leave the
// bordered field (border >= longest beam), so the evaluation is 4 FFMA + 2 FADD + IMAD + LEA + load + FADD with no
// bounds test: the magic-add bit patterns are fed to the address arithmetic as they are, their constant parts folded
// into the base. Any batch with a particle outside the map takes the bounds-tested path; both give the same values.
// Per-lane partial sums of the 32 particles are combined by a transpose-reduction: 31 shuffles per 32 particles, and
// for every particle exactly the xor-butterfly (16,8,4,2,1) summation tree that DESIGN.md NS-3 specifies.
constexpr float NS_MAGIC = 12582912.0f;          // 1.5 * 2^23
constexpr int NS_MAGIC_BITS = 0x4B400000;

template <bool SMEM>
struct NsFieldView {
    const float* lf;        // generic pointer (global path, and the bounds-tested path)
    uint32_t base;          // SMEM fast path: shared byte address of the field + folded constant; else folded cell constant
    uint32_t Wp, Wp4;
    unsigned W, H;
    int pad;
    float lf_out;
};

// bounds-tested evaluation (any particle position)
template <bool SMEM>
__device__ __forceinline__ float ns_eval_checked(const NsFieldView<SMEM>& V, float gx0, float gy0, float c, float s, float2 bm) {
    const float tx = ns::addf(ns::fmaf_(c, bm.x, ns::fmaf_(-s, bm.y, gx0)), NS_MAGIC);
    const float ty = ns::addf(ns::fmaf_(s, bm.x, ns::fmaf_(c, bm.y, gy0)), NS_MAGIC);
    const unsigned ix = (unsigned)(__float_as_int(tx) - NS_MAGIC_BITS);
    const unsigned iy = (unsigned)(__float_as_int(ty) - NS_MAGIC_BITS);
    const bool in = ix < V.W && iy < V.H;
    const unsigned idx = in ? (iy + V.pad) * V.Wp + ix + V.pad : 0u;
    const float v = SMEM ? V.lf[idx] : __ldg(V.lf + idx);
    return in ? v : V.lf_out;
}
// evaluation for a particle inside the map: the endpoint is inside the bordered field by construction
template <bool SMEM>
__device__ __forceinline__ float ns_eval_fast(const NsFieldView<SMEM>& V, float gx0, float gy0, float c, float s, float2 bm) {
    const float tx = ns::addf(ns::fmaf_(c, bm.x, ns::fmaf_(-s, bm.y, gx0)), NS_MAGIC);
    const float ty = ns::addf(ns::fmaf_(s, bm.x, ns::fmaf_(c, bm.y, gy0)), NS_MAGIC);
    if (SMEM) {
        // byte address = field + 4 * ((iy + pad) * Wp + ix + pad), iy = bits(ty) - MAGIC_BITS (wrapping u32 arithmetic)
        // (IMAD, LEA, LDS: written as PTX so the compiler does not re-associate it into three integer operations)
        float v;
        asm("{\n.reg .u32 t, u;\nmad.lo.u32 t, %1, %2, %3;\nshl.b32 u, %4, 2;\nadd.u32 t, t, u;\nld.shared.f32 %0, [t];\n}"
            : "=f"(v)
            : "r"(__float_as_uint(ty)), "r"(V.Wp4), "r"(V.base), "r"(__float_as_uint(tx)));
        return v;
    } else {
        const uint32_t idx = __float_as_uint(ty) * V.Wp + V.base + __float_as_uint(tx);
        return __ldg(V.lf + idx);
    }
}

template <bool SMEM, bool FAST>
__device__ __forceinline__ float ns_score_batch(const NsFieldView<SMEM>& V, const float2* __restrict__ s_beams, int n_beams, int lane, float gx0,
                                                float gy0, float c, float s) {
    float acc[32];
#pragma unroll
    for (int k0 = 0; k0 < 32; k0 += 4) {
        float X[4], Y[4], C[4], S[4], a[4];
#pragma unroll
        for (int q = 0; q < 4; q++) {
            X[q] = __shfl_sync(0xffffffffu, gx0, k0 + q); Y[q] = __shfl_sync(0xffffffffu, gy0, k0 + q);
            C[q] = __shfl_sync(0xffffffffu, c, k0 + q); S[q] = __shfl_sync(0xffffffffu, s, k0 + q);
            a[q] = 0.f;
        }
#pragma unroll 2
        for (int b = lane; b < n_beams; b += 32) {
            const float2 bm = s_beams[b];
#pragma unroll
            for (int q = 0; q < 4; q++)
                a[q] = ns::addf(a[q], FAST ? ns_eval_fast<SMEM>(V, X[q], Y[q], C[q], S[q], bm) : ns_eval_checked<SMEM>(V, X[q], Y[q], C[q], S[q], bm));
        }
#pragma unroll
        for (int q = 0; q < 4; q++) acc[k0 + q] = a[q];
    }
    // transpose-reduction: after the stage with offset o, lanes with bit o set hold the upper half of the particles
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const bool up = (lane & o) != 0;
#pragma unroll
        for (int j = 0; j < o; j++) {
            const float keep = up ? acc[j + o] : acc[j];
            const float send = up ? acc[j] : acc[j + o];
            acc[j] = ns::addf(keep, __shfl_xor_sync(0xffffffffu, send, o));
        }
    }
    return acc[0];
}

template <bool SMEM_FIELD>
__global__ void __launch_bounds__(512, 1) k_ns_update(const float4* __restrict__ part, int64_t n, NsField F,
                                                   const float2* __restrict__ beams, int n_beams, float* __restrict__ ll_out,
                                                   int* __restrict__ max_bits /* ordered-int max of ll */) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ uint64_t bar;
    __shared__ float warp_max[16];
    float* s_lf = reinterpret_cast<float*>(smem_raw);
    float2* s_beams = reinterpret_cast<float2*>(smem_raw + (SMEM_FIELD ? F.bytes_padded : 0));
    if (SMEM_FIELD) tma_stage(s_lf, F.lf, (uint32_t)F.bytes_padded, &bar);
    for (int b = threadIdx.x; b < n_beams; b += blockDim.x) s_beams[b] = beams[b];
    __syncthreads();
    NsFieldView<SMEM_FIELD> V;
    V.lf = SMEM_FIELD ? s_lf : F.lf;
    V.Wp = (uint32_t)F.Wp; V.Wp4 = 4u * (uint32_t)F.Wp;
    V.W = (unsigned)F.W; V.H = (unsigned)F.H; V.pad = F.pad; V.lf_out = F.lf_out;
    // (pad - MAGIC_BITS) * (Wp + 1): turns the raw magic-add bit patterns into the bordered cell index (mod 2^32)
    const uint32_t fold = (uint32_t)(F.pad - NS_MAGIC_BITS) * ((uint32_t)F.Wp + 1u);
    V.base = SMEM_FIELD ? smem_u32(s_lf) + 4u * fold : fold;
    const bool fast_ok = F.pad > 0;
    const float ox = F.ox, oy = F.oy, inv_res = F.inv_res;
    const float x_hi = (float)F.W - 0.5f, y_hi = (float)F.H - 0.5f;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, warps_per_block = blockDim.x >> 5;
    const int64_t n_batches = (n + 31) / 32;
    float best = -3.0e38f;
    for (int64_t batch = (int64_t)blockIdx.x * warps_per_block + warp; batch < n_batches; batch += (int64_t)gridDim.x * warps_per_block) {
        const int64_t i = batch * 32 + lane;
        float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
        if (i < n) p = part[i];
        float s, c;
        ns::det_sincosf(p.z, s, c);
        const float gx0 = ns::fmaf_(ns::addf(p.x, -ox), inv_res, -0.5f);
        const float gy0 = ns::fmaf_(ns::addf(p.y, -oy), inv_res, -0.5f);
        const bool inside = gx0 >= -0.5f && gx0 <= x_hi && gy0 >= -0.5f && gy0 <= y_hi;        // false for NaN
        float ll;
        if (fast_ok && __all_sync(0xffffffffu, inside)) ll = ns_score_batch<SMEM_FIELD, true>(V, s_beams, n_beams, lane, gx0, gy0, c, s);
        else ll = ns_score_batch<SMEM_FIELD, false>(V, s_beams, n_beams, lane, gx0, gy0, c, s);
        if (i < n) { ll_out[i] = ll; best = fmaxf(best, ll); }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) best = fmaxf(best, __shfl_xor_sync(0xffffffffu, best, o));
    if (lane == 0) warp_max[warp] = best;
    __syncthreads();
    if (threadIdx.x == 0) {
        float m = warp_max[0];
        for (int w = 1; w < warps_per_block; w++) m = fmaxf(m, warp_max[w]);
        // order-preserving float -> int so atomicMax works for negative values too
        int bits = __float_as_int(m);
        bits = bits >= 0 ? bits : bits ^ 0x7fffffff;
        atomicMax(max_bits, bits);
    }
}

// ---- weights + prefix sum -------------------------------------------------------------------------------------------------
constexpr int NS_SCAN_THREADS = 256;
constexpr int NS_SCAN_ITEMS = 8;
constexpr int NS_SCAN_TILE = NS_SCAN_THREADS * NS_SCAN_ITEMS;

// the global maximum log-likelihood lives in device memory as an order-preserving int (so atomicMax / NCCL max work)
__device__ __forceinline__ float ns_decode_max(const int* __restrict__ max_bits) {
    int b = *max_bits;
    b = b >= 0 ? b : b ^ 0x7fffffff;
    return __int_as_float(b);
}
__device__ __forceinline__ uint64_t ns_weight(float ll, float max_ll, float temper) {
    return ns::det_exp_q32(ns::mulf(temper, ns::addf(ll, -max_ll)));
}
__device__ __forceinline__ uint64_t block_scan_u64(uint64_t v, uint64_t* sm8, uint64_t& block_total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint64_t t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += t;
    }
    if (lane == 31) sm8[warp] = v;
    __syncthreads();
    uint64_t pre = 0, tot = 0;
    for (int k = 0; k < NS_SCAN_THREADS / 32; k++) { if (k < warp) pre += sm8[k]; tot += sm8[k]; }
    __syncthreads();
    block_total = tot;
    return pre + v;      // inclusive
}
// pass 1: per-tile sums of W
__global__ void __launch_bounds__(NS_SCAN_THREADS) k_ns_weights_sum(const float* __restrict__ ll, int64_t n, const int* __restrict__ max_bits,
                                                                   float temper, uint64_t* __restrict__ tile_sums) {
    __shared__ uint64_t sm[8];
    const float max_ll = ns_decode_max(max_bits);
    const int64_t base = (int64_t)blockIdx.x * NS_SCAN_TILE + (int64_t)threadIdx.x * NS_SCAN_ITEMS;
    uint64_t s = 0;
#pragma unroll
    for (int j = 0; j < NS_SCAN_ITEMS; j++)
        if (base + j < n) s += ns_weight(ll[base + j], max_ll, temper);
    uint64_t tot;
    block_scan_u64(s, sm, tot);
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = tot;
}
// pass 2: exclusive scan of tile sums (one block), grand total
__global__ void __launch_bounds__(1024) k_ns_tile_offsets(uint64_t* __restrict__ tile_sums, int nt, uint64_t* __restrict__ total) {
    __shared__ uint64_t sm[1024];
    uint64_t carry = 0;
    for (int c0 = 0; c0 < nt; c0 += 1024) {
        const int i = c0 + threadIdx.x;
        uint64_t v = i < nt ? tile_sums[i] : 0;
        sm[threadIdx.x] = v;
        __syncthreads();
        for (int o = 1; o < 1024; o <<= 1) {
            uint64_t t = threadIdx.x >= o ? sm[threadIdx.x - o] : 0;
            __syncthreads();
            sm[threadIdx.x] += t;
            __syncthreads();
        }
        if (i < nt) tile_sums[i] = carry + sm[threadIdx.x] - v;      // exclusive
        carry += sm[1023];
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = carry;
}
// pass 3: inclusive prefix per particle (local to this shard), and the unnormalised weight into the particle record
__global__ void __launch_bounds__(NS_SCAN_THREADS) k_ns_weights_scan(const float* __restrict__ ll, int64_t n, const int* __restrict__ max_bits,
                                                                    float temper, const uint64_t* __restrict__ tile_offsets,
                                                                    uint64_t* __restrict__ prefix, float4* __restrict__ part) {
    __shared__ uint64_t sm[8];
    const float max_ll = ns_decode_max(max_bits);
    const int64_t base = (int64_t)blockIdx.x * NS_SCAN_TILE + (int64_t)threadIdx.x * NS_SCAN_ITEMS;
    uint64_t w[NS_SCAN_ITEMS];
    uint64_t s = 0;
#pragma unroll
    for (int j = 0; j < NS_SCAN_ITEMS; j++) {
        w[j] = (base + j < n) ? ns_weight(ll[base + j], max_ll, temper) : 0;
        s += w[j];
    }
    uint64_t tot;
    uint64_t incl = block_scan_u64(s, sm, tot);
    uint64_t run = tile_offsets[blockIdx.x] + incl - s;
#pragma unroll
    for (int j = 0; j < NS_SCAN_ITEMS; j++) {
        run += w[j];
        if (base + j < n) {
            prefix[base + j] = run;
            part[base + j].w = (float)((double)w[j] * 2.3283064365386963e-10);      // W * 2^-32, max particle = 1
        }
    }
}

// ---- systematic resampling ------------------------------------------------------------------------------------------------
struct NsDest {
    float4* part[8];        // next-step particle buffer of every shard (own memory or mapped peer memory)
    int* anc[8];            // ancestor (global index) per output slot, same sharding
    int64_t per_rank;       // slots [r*per_rank, (r+1)*per_rank) live on shard r
    int world;
};
// Output slot k (global) takes the first particle i of this shard with (offset + prefix[i]) selecting k. Slots
// [k_lo, k_hi) are exactly those whose ancestor lives here (host plan, ns_plan.hpp).
__global__ void __launch_bounds__(256) k_ns_resample(const float4* __restrict__ src, const uint64_t* __restrict__ prefix, int64_t n_local,
                                                     int64_t g0, uint64_t offset, uint64_t total, uint64_t n_global, uint32_t u0,
                                                     int64_t k_lo, int64_t k_hi, NsDest D, float new_weight) {
    const int64_t k = k_lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= k_hi) return;
    const ns::U128 rhs = ns::rhs_of((uint64_t)k, u0, total);
    int64_t lo = 0, len = n_local;
    while (len > 0) {                       // first i with selects(offset + prefix[i])
        const int64_t half = len >> 1;
        if (!ns::selects(offset + prefix[lo + half], n_global, rhs)) { lo += half + 1; len -= half + 1; } else len = half;
    }
    if (lo >= n_local) lo = n_local - 1;    // cannot happen when the plan is right; keeps the access in range
    float4 p = src[lo];
    p.w = new_weight;
    int r = (int)(k / D.per_rank);
    if (r >= D.world) r = D.world - 1;
    const int64_t slot = k - (int64_t)r * D.per_rank;
    D.part[r][slot] = p;
    D.anc[r][slot] = (int)(g0 + lo);
}

// ---- gather micro-benchmark: the denominator for the sensor-model kernel ------------------------------------------------
// Independent random 4-byte reads from a table held in shared memory (SMEM) or in global memory (L2- or HBM-resident,
// depending on its size), issued with the thread/block shape of k_ns_update. Reports nothing itself: the host times it.
template <bool SMEM>
__global__ void __launch_bounds__(512) k_gather_bench(const float* __restrict__ table, uint32_t n_words, int iters, float* __restrict__ out) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* s_t = reinterpret_cast<float*>(smem_raw);
    if (SMEM) {
        for (uint32_t i = threadIdx.x; i < n_words; i += blockDim.x) s_t[i] = table[i];
        __syncthreads();
    }
    const float* t = SMEM ? s_t : table;
    uint32_t state = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u;
    float acc = 0.f;
#pragma unroll 4
    for (int it = 0; it < iters; it++) {
        state = state * 1664525u + 1013904223u;
        const uint32_t idx = __umulhi(state, n_words);
        acc += SMEM ? t[idx] : __ldg(t + idx);
    }
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

// The resampling plan of this shard, computed on the device from the all-gathered Q32 totals (no host round trip).
struct NsPlan {
    uint64_t offset, total;
    int64_t k_lo, k_hi;
};
__global__ void k_ns_plan(const uint64_t* __restrict__ totals, int world, int rank, uint64_t n_global, uint32_t u0, NsPlan* __restrict__ plan) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    uint64_t off = 0, tot = 0;
    for (int r = 0; r < world; r++) { if (r < rank) off += totals[r]; tot += totals[r]; }
    NsPlan p;
    p.offset = off; p.total = tot;
    p.k_lo = tot ? ns::first_slot(off, tot, n_global, u0) : 0;
    p.k_hi = tot ? ns::first_slot(off + totals[rank], tot, n_global, u0) : 0;
    *plan = p;
}
// k_ns_resample with the plan read from device memory and a grid-stride loop (the slot count is not known on the host).
__global__ void __launch_bounds__(256) k_ns_resample_planned(const float4* __restrict__ src, const uint64_t* __restrict__ prefix, int64_t n_local,
                                                             int64_t g0, const NsPlan* __restrict__ plan, uint64_t n_global, uint32_t u0, NsDest D,
                                                             float new_weight) {
    const NsPlan P = *plan;
    for (int64_t k = P.k_lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < P.k_hi; k += (int64_t)gridDim.x * blockDim.x) {
        const ns::U128 rhs = ns::rhs_of((uint64_t)k, u0, P.total);
        int64_t lo = 0, len = n_local;
        while (len > 0) {
            const int64_t half = len >> 1;
            if (!ns::selects(P.offset + prefix[lo + half], n_global, rhs)) { lo += half + 1; len -= half + 1; } else len = half;
        }
        if (lo >= n_local) lo = n_local - 1;
        float4 p = src[lo];
        p.w = new_weight;
        int r = (int)(k / D.per_rank);
        if (r >= D.world) r = D.world - 1;
        const int64_t slot = k - (int64_t)r * D.per_rank;
        D.part[r][slot] = p;
        D.anc[r][slot] = (int)(g0 + lo);
    }
    __threadfence_system();        // peer stores must be visible before the step's closing collective lets anyone read them
}
// out[k] = sum over blocks of partials[b*5+k] (one warp per output, fixed order)
__global__ void k_ns_pose_reduce(const double* __restrict__ partials, int n_blocks, double* __restrict__ out5) {
    const int o = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (o >= 5) return;
    double s = 0;
    for (int b = lane; b < n_blocks; b += 32) s += partials[(size_t)b * 5 + o];
    s = warp_sum(s);
    if (lane == 0) out5[o] = s;
}

// Weighted pose sums with the particle weights as they stand: {sum w, sum w x, sum w y, sum w sin, sum w cos} per block.
__global__ void __launch_bounds__(256) k_ns_pose_partials(const float4* __restrict__ part, int64_t n, double* __restrict__ partials) {
    __shared__ double ws[8][5];
    double a[5] = {0, 0, 0, 0, 0};
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        float4 p = part[i];
        float s, c;
        ns::det_sincosf(p.z, s, c);
        const double w = (double)p.w;
        a[0] += w; a[1] += w * (double)p.x; a[2] += w * (double)p.y; a[3] += w * (double)s; a[4] += w * (double)c;
    }
#pragma unroll
    for (int k = 0; k < 5; k++) a[k] = warp_sum(a[k]);
    if ((threadIdx.x & 31) == 0)
        for (int k = 0; k < 5; k++) ws[threadIdx.x >> 5][k] = a[k];
    __syncthreads();
    if (threadIdx.x < 5) { double s = 0; for (int w = 0; w < 8; w++) s += ws[w][threadIdx.x]; partials[(size_t)blockIdx.x * 5 + threadIdx.x] = s; }
}

}  // namespace mcl
