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
__global__ void k_ns_fill_u8(uint8_t* __restrict__ dst, size_t n, uint8_t v) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) dst[i] = v;
}
// code_of_d2 / code_out (both or neither): the same field as one byte per cell, the rank of d2 among the attainable
// squared distances; k_ns_update turns the code back into table[d2] through a <= 256-entry shared-memory table
__global__ void k_ns_edt_rows(const uint16_t* __restrict__ g, int W, int H, int R, const float* __restrict__ lf_of_d2,
                              uint16_t* __restrict__ d2_out, float* __restrict__ lf_out, int pad,
                              const uint8_t* __restrict__ code_of_d2, uint8_t* __restrict__ code_out) {
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
    if (code_out) code_out[(size_t)(y + pad) * (W + 2 * pad) + (x + pad)] = code_of_d2[best];
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

// mcl_ns_step: the first kernel of a step also resets what the later kernels of that step accumulate into (the running
// max log-likelihood, the scan's tile states / group sums), so the step needs no memset / copy commands in between.
struct NsStepPrep { int* maxbits; unsigned long long* zero; int zero_words; };      // maxbits == null: nothing to prepare
__global__ void __launch_bounds__(256) k_ns_predict(float4* __restrict__ part, int64_t n, int64_t g0, NsMotion m, uint32_t step,
                                                    uint32_t k0, uint32_t k1, NsStepPrep prep) {
    pdl_enter();
    if (prep.maxbits != nullptr && blockIdx.x == 0) {
        if (threadIdx.x == 0) *prep.maxbits = INT32_MIN;
        for (int j = threadIdx.x; j < prep.zero_words; j += blockDim.x) prep.zero[j] = 0ull;
    }
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
    ns::det_sincosf32(ns::addf(p.z, r1), s, c);
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
    const uint8_t* lf8;     // NS_FIELD_U8: one code per cell, same bordered layout; value = codes[code]
    const float* codes;     // NS_FIELD_U8: code -> log-likelihood (n_codes <= 256 entries, global)
    int n_codes;
};
// where k_ns_update reads the field from
constexpr int NS_FIELD_SMEM = 0;     // fp32 field staged into shared memory by TMA
constexpr int NS_FIELD_GLOBAL = 1;   // fp32 field through L1/L2
constexpr int NS_FIELD_U8 = 2;       // one-byte codes through L1/L2 (a quarter of the footprint: fields too large for L2 as fp32)
                                     // + code table in shared memory

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

template <int KIND>
struct NsFieldView {
    const float* lf;        // generic pointer (global path, and the bounds-tested path)
    const uint8_t* lf8;     // NS_FIELD_U8: the code field
    const float* codes;     // NS_FIELD_U8: shared-memory code table
    uint32_t win_hi;           // global fast path: high address word when the field lies inside one 4 GiB window
    uint32_t base;          // SMEM fast path: shared byte address of the field + folded constant; else folded cell constant
    uint32_t Wp, Wp4;
    unsigned W, H;
    int pad;
    float lf_out;
};

// bounds-tested evaluation (any particle position)
template <int KIND>
__device__ __forceinline__ float ns_eval_checked(const NsFieldView<KIND>& V, float gx0, float gy0, float c, float s, float2 bm) {
    const float tx = ns::addf(ns::fmaf_(c, bm.x, ns::fmaf_(-s, bm.y, gx0)), NS_MAGIC);
    const float ty = ns::addf(ns::fmaf_(s, bm.x, ns::fmaf_(c, bm.y, gy0)), NS_MAGIC);
    const unsigned ix = (unsigned)(__float_as_int(tx) - NS_MAGIC_BITS);
    const unsigned iy = (unsigned)(__float_as_int(ty) - NS_MAGIC_BITS);
    const bool in = ix < V.W && iy < V.H;
    const unsigned idx = in ? (iy + V.pad) * V.Wp + ix + V.pad : 0u;
    const float v = KIND == NS_FIELD_U8 ? V.codes[__ldg(V.lf8 + idx)] : KIND == NS_FIELD_SMEM ? V.lf[idx] : __ldg(V.lf + idx);
    return in ? v : V.lf_out;
}
// The field value for a particle inside the map from the raw magic-add bit patterns of the endpoint's cell coordinates:
// the endpoint is inside the bordered field by construction, the constant parts of the bit patterns are folded into the
// base. SHIFT (global fields): the field lies inside one 4 GiB-aligned window, so only the low address word depends on
// the cell (IMAD + LEA/IADD like the shared-memory path, constant high word) and the 64-bit IMAD.WIDE address form is
// avoided - the FMA pipe is this kernel's busiest unit.
template <int KIND, bool SHIFT>
__device__ __forceinline__ float ns_load_fast(const NsFieldView<KIND>& V, uint32_t tx_bits, uint32_t ty_bits) {
    if (KIND == NS_FIELD_SMEM) {
        // byte address = field + 4 * ((iy + pad) * Wp + ix + pad), iy = bits(ty) - MAGIC_BITS (wrapping u32 arithmetic)
        // (IMAD, LEA, LDS: written as PTX so the compiler does not re-associate it into three integer operations)
        float v;
        asm("{\n.reg .u32 t, u;\nmad.lo.u32 t, %1, %2, %3;\nshl.b32 u, %4, 2;\nadd.u32 t, t, u;\nld.shared.f32 %0, [t];\n}"
            : "=f"(v)
            : "r"(ty_bits), "r"(V.Wp4), "r"(V.base), "r"(tx_bits));
        return v;
    } else if (KIND == NS_FIELD_GLOBAL) {
        if (SHIFT) {
            const uint32_t lo = ty_bits * V.Wp4 + V.base + (tx_bits << 2);
            return __ldg(reinterpret_cast<const float*>(((uint64_t)V.win_hi << 32) | (uint64_t)lo));
        }
        const uint32_t idx = ty_bits * V.Wp + V.base + tx_bits;
        return __ldg(V.lf + idx);
    } else {
        uint32_t code;
        if (SHIFT) {
            const uint32_t lo = ty_bits * V.Wp + V.base + tx_bits;
            code = __ldg(reinterpret_cast<const uint8_t*>(((uint64_t)V.win_hi << 32) | (uint64_t)lo));
        } else {
            const uint32_t idx = ty_bits * V.Wp + V.base + tx_bits;
            code = __ldg(V.lf8 + idx);
        }
        return V.codes[code];
    }
}
template <int KIND, bool SHIFT>
__device__ __forceinline__ float ns_eval_fast(const NsFieldView<KIND>& V, float gx0, float gy0, float c, float s, float2 bm) {
    const float tx = ns::addf(ns::fmaf_(c, bm.x, ns::fmaf_(-s, bm.y, gx0)), NS_MAGIC);
    const float ty = ns::addf(ns::fmaf_(s, bm.x, ns::fmaf_(c, bm.y, gy0)), NS_MAGIC);
    return ns_load_fast<KIND, SHIFT>(V, __float_as_uint(tx), __float_as_uint(ty));
}

// per-lane partial sums of P particles -> lane (g0 + j) holds particle j's total: the xor-butterfly (16,8,4,2,1) tree of
// DESIGN.md NS-3 for every particle, as a transpose-reduction (31 shuffles per 32 particles)
template <int P>
__device__ __forceinline__ float ns_transpose_reduce(float (&acc)[P], int lane) {
#pragma unroll
    for (int o = 16; o >= P; o >>= 1) {
#pragma unroll
        for (int j = 0; j < P; j++) acc[j] = ns::addf(acc[j], __shfl_xor_sync(0xffffffffu, acc[j], o));
    }
    // after the stage with offset o, lanes with bit o set hold the upper half of the particles
#pragma unroll
    for (int o = (P < 32 ? P / 2 : 16); o > 0; o >>= 1) {
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

template <int KIND, bool FAST, int P, bool SHIFT = false>
__device__ __forceinline__ float ns_score_batch(const NsFieldView<KIND>& V, const float2* __restrict__ s_beams, int n_beams, int lane, float gx0,
                                                float gy0, float c, float s) {
    float mine = 0.f;
#pragma unroll 1
    for (int g0 = 0; g0 < 32; g0 += P) {
        float acc[P];
#pragma unroll
        for (int k0 = 0; k0 < P; k0 += 4) {
            float X[4], Y[4], C[4], S[4], a[4];
#pragma unroll
            for (int q = 0; q < 4; q++) {
                X[q] = __shfl_sync(0xffffffffu, gx0, g0 + k0 + q); Y[q] = __shfl_sync(0xffffffffu, gy0, g0 + k0 + q);
                C[q] = __shfl_sync(0xffffffffu, c, g0 + k0 + q); S[q] = __shfl_sync(0xffffffffu, s, g0 + k0 + q);
                a[q] = 0.f;
            }
#pragma unroll (KIND == NS_FIELD_SMEM ? 2 : 4)
            for (int b = lane; b < n_beams; b += 32) {
                const float2 bm = s_beams[b];
#pragma unroll
                for (int q = 0; q < 4; q++)
                    a[q] = ns::addf(a[q], FAST ? ns_eval_fast<KIND, SHIFT>(V, X[q], Y[q], C[q], S[q], bm) : ns_eval_checked<KIND>(V, X[q], Y[q], C[q], S[q], bm));
            }
#pragma unroll
            for (int q = 0; q < 4; q++) acc[k0 + q] = a[q];
        }
        const float total = ns_transpose_reduce<P>(acc, lane);
        if ((lane & ~(P - 1)) == g0) mine = total;
    }
    return mine;
}

// ---- packed fp32 (sm_100 FFMA2 / FADD2) form of the fast path ---------------------------------------------------------------
// fma.rn.f32x2 / add.rn.f32x2 perform two independent IEEE fp32 operations on a 64-bit register pair per issue slot; the
// values are bit-identical to the scalar form. Two particles share a pair: (pose_q, pose_q+1) against the same beam point
// (a scalar operand broadcast to both halves). Measured (profiles/): FFMA2 occupies the FMA pipe for two cycles, so it
// saves issue slots, not pipe time: +4 % on the global-field path (issue bound), nothing on the shared-memory path (FMA
// pipe and LDS bound), which therefore keeps the scalar form.
// same summation order, same values as ns_score_batch<KIND, true, P, SHIFT>
template <int KIND, int P, bool SHIFT>
__device__ __forceinline__ float ns_score_batch_packed(const NsFieldView<KIND>& V, const float2* __restrict__ s_beams, int n_beams, int lane,
                                                       float gx0, float gy0, float c, float s) {
    float mine = 0.f;
    const f32x2 magic2 = pk2(NS_MAGIC, NS_MAGIC);
    const float ns_ = -s;
#pragma unroll 1
    for (int g0 = 0; g0 < 32; g0 += P) {
        float acc[P];
#pragma unroll
        for (int k0 = 0; k0 < P; k0 += 4) {
            f32x2 X[2], Y[2], C[2], S[2], NSn[2], a[2];
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const int q = g0 + k0 + 2 * h;
                X[h] = pk2(__shfl_sync(0xffffffffu, gx0, q), __shfl_sync(0xffffffffu, gx0, q + 1));
                Y[h] = pk2(__shfl_sync(0xffffffffu, gy0, q), __shfl_sync(0xffffffffu, gy0, q + 1));
                C[h] = pk2(__shfl_sync(0xffffffffu, c, q), __shfl_sync(0xffffffffu, c, q + 1));
                S[h] = pk2(__shfl_sync(0xffffffffu, s, q), __shfl_sync(0xffffffffu, s, q + 1));
                NSn[h] = pk2(__shfl_sync(0xffffffffu, ns_, q), __shfl_sync(0xffffffffu, ns_, q + 1));
                a[h] = pk2(0.f, 0.f);
            }
#pragma unroll (KIND == NS_FIELD_SMEM ? 2 : 4)
            for (int b = lane; b < n_beams; b += 32) {
                const float2 bm = s_beams[b];
                const f32x2 bx2 = pk2(bm.x, bm.x), by2 = pk2(bm.y, bm.y);
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    const f32x2 tx2 = add2(fma2(C[h], bx2, fma2(NSn[h], by2, X[h])), magic2);
                    const f32x2 ty2 = add2(fma2(S[h], bx2, fma2(C[h], by2, Y[h])), magic2);
                    uint32_t tx0, tx1, ty0, ty1;
                    upk2u(tx2, tx0, tx1); upk2u(ty2, ty0, ty1);
                    const float v0 = ns_load_fast<KIND, SHIFT>(V, tx0, ty0);
                    const float v1 = ns_load_fast<KIND, SHIFT>(V, tx1, ty1);
                    a[h] = add2(a[h], pk2(v0, v1));
                }
            }
#pragma unroll
            for (int h = 0; h < 2; h++) upk2(a[h], acc[k0 + 2 * h], acc[k0 + 2 * h + 1]);
        }
        const float total = ns_transpose_reduce<P>(acc, lane);
        if ((lane & ~(P - 1)) == g0) mine = total;
    }
    return mine;
}

// Mixed form: the four FFMAs per evaluation stay scalar (they can issue on either half of the FMA pipe), the two magic adds
// of an evaluation travel as one FADD2 on the (x, y) pair and the running sums of two particles as one FADD2. FFMA2/FADD2
// occupy the heavy half of the FMA pipe for two cycles (measured), so packing everything moves the bound from the issue
// slots to that half-pipe; packing only the adds balances the two halves and still saves 3 of 13 issue slots.
template <int KIND, int P, bool SHIFT>
__device__ __forceinline__ float ns_score_batch_mixed(const NsFieldView<KIND>& V, const float2* __restrict__ s_beams, int n_beams, int lane,
                                                      float gx0, float gy0, float c, float s) {
    float mine = 0.f;
    const f32x2 magic2 = pk2(NS_MAGIC, NS_MAGIC);
#pragma unroll 1
    for (int g0 = 0; g0 < 32; g0 += P) {
        float acc[P];
#pragma unroll
        for (int k0 = 0; k0 < P; k0 += 4) {
            float X[4], Y[4], C[4], S[4];
            f32x2 a[2];
#pragma unroll
            for (int q = 0; q < 4; q++) {
                X[q] = __shfl_sync(0xffffffffu, gx0, g0 + k0 + q); Y[q] = __shfl_sync(0xffffffffu, gy0, g0 + k0 + q);
                C[q] = __shfl_sync(0xffffffffu, c, g0 + k0 + q); S[q] = __shfl_sync(0xffffffffu, s, g0 + k0 + q);
            }
            a[0] = a[1] = pk2(0.f, 0.f);
#pragma unroll (KIND == NS_FIELD_SMEM ? 2 : 4)
            for (int b = lane; b < n_beams; b += 32) {
                const float2 bm = s_beams[b];
                float v[4];
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const float ex = ns::fmaf_(C[q], bm.x, ns::fmaf_(-S[q], bm.y, X[q]));
                    const float ey = ns::fmaf_(S[q], bm.x, ns::fmaf_(C[q], bm.y, Y[q]));
                    uint32_t tx, ty;
                    upk2u(add2(pk2(ex, ey), magic2), tx, ty);
                    v[q] = ns_load_fast<KIND, SHIFT>(V, tx, ty);
                }
                a[0] = add2(a[0], pk2(v[0], v[1]));
                a[1] = add2(a[1], pk2(v[2], v[3]));
            }
            upk2(a[0], acc[k0], acc[k0 + 1]);
            upk2(a[1], acc[k0 + 2], acc[k0 + 3]);
        }
        const float total = ns_transpose_reduce<P>(acc, lane);
        if ((lane & ~(P - 1)) == g0) mine = total;
    }
    return mine;
}

constexpr int NS_UPD_THREADS = 1024;
constexpr int NS_UPD_P = 8;
constexpr int NS_MAX_CODES = 256;

// dynamic shared memory: [field (NS_FIELD_SMEM) | code table, NS_MAX_CODES floats (NS_FIELD_U8)] [beam points]
// PACK: 0 scalar arithmetic, 1 everything packed (FFMA2/FADD2), 2 only the adds packed
template <int KIND, int PACK>
__global__ void __launch_bounds__(NS_UPD_THREADS, 1) k_ns_update(const float4* __restrict__ part, int64_t n, NsField F,
                                                                const float2* __restrict__ beams, int n_beams, float* __restrict__ ll_out,
                                                                int* __restrict__ max_bits /* ordered-int max of ll */) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ uint64_t bar;
    __shared__ float warp_max[NS_UPD_THREADS / 32];
    float* s_lf = reinterpret_cast<float*>(smem_raw);
    const size_t front = KIND == NS_FIELD_SMEM ? (size_t)F.bytes_padded : KIND == NS_FIELD_U8 ? NS_MAX_CODES * sizeof(float) : 0;
    float2* s_beams = reinterpret_cast<float2*>(smem_raw + front);
    if (KIND == NS_FIELD_SMEM) tma_stage(s_lf, F.lf, (uint32_t)F.bytes_padded, &bar);
    if (KIND == NS_FIELD_U8)
        for (int k = threadIdx.x; k < NS_MAX_CODES; k += blockDim.x) s_lf[k] = k < F.n_codes ? F.codes[k] : F.lf_out;
    for (int b = threadIdx.x; b < n_beams; b += blockDim.x) s_beams[b] = beams[b];
    __syncthreads();
    NsFieldView<KIND> V;
    V.lf = KIND == NS_FIELD_SMEM ? s_lf : F.lf;
    V.lf8 = F.lf8; V.codes = s_lf;
    V.Wp = (uint32_t)F.Wp; V.Wp4 = 4u * (uint32_t)F.Wp;
    V.W = (unsigned)F.W; V.H = (unsigned)F.H; V.pad = F.pad; V.lf_out = F.lf_out;
    // (pad - MAGIC_BITS) * (Wp + 1): turns the raw magic-add bit patterns into the bordered cell index (mod 2^32)
    const uint32_t fold = (uint32_t)(F.pad - NS_MAGIC_BITS) * ((uint32_t)F.Wp + 1u);
    V.base = KIND == NS_FIELD_SMEM ? smem_u32(s_lf) + 4u * fold : fold;
    // global field inside one 4 GiB-aligned window: 32-bit address arithmetic, constant high word
    const uint64_t cell_bytes = KIND == NS_FIELD_U8 ? 1 : 4;
    const uint64_t f_lo = KIND == NS_FIELD_U8 ? reinterpret_cast<uint64_t>(F.lf8) : reinterpret_cast<uint64_t>(F.lf);
    const uint64_t f_hi = f_lo + (uint64_t)F.Wp * (uint64_t)(F.H + 2 * F.pad) * cell_bytes - 1;
    const bool one_window = KIND != NS_FIELD_SMEM && (f_lo >> 32) == (f_hi >> 32);
    V.win_hi = (uint32_t)(f_lo >> 32);
    if (one_window) V.base = (uint32_t)f_lo + (uint32_t)cell_bytes * fold;
    const bool fast_ok = F.pad > 0;
    const float ox = F.ox, oy = F.oy, inv_res = F.inv_res;
    const float x_hi = (float)F.W - 0.5f, y_hi = (float)F.H - 0.5f;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, warps_per_block = blockDim.x >> 5;
    const int64_t n_batches = (n + 31) / 32;
    float best = -3.0e38f;
    // The staging above read only the field, the code table and the beams (written by host copies / map set-up, never by a
    // kernel of the step), so under a programmatic launch it overlaps the predecessor's tail; particles, ll_out and
    // max_bits belong to the step.
    pdl_enter();
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
        if (fast_ok && __all_sync(0xffffffffu, inside)) {
            if (PACK == 1) {
                if (one_window) ll = ns_score_batch_packed<KIND, NS_UPD_P, true>(V, s_beams, n_beams, lane, gx0, gy0, c, s);
                else ll = ns_score_batch_packed<KIND, NS_UPD_P, false>(V, s_beams, n_beams, lane, gx0, gy0, c, s);
            } else if (PACK == 2) {
                if (one_window) ll = ns_score_batch_mixed<KIND, NS_UPD_P, true>(V, s_beams, n_beams, lane, gx0, gy0, c, s);
                else ll = ns_score_batch_mixed<KIND, NS_UPD_P, false>(V, s_beams, n_beams, lane, gx0, gy0, c, s);
            } else {
                if (one_window) ll = ns_score_batch<KIND, true, NS_UPD_P, true>(V, s_beams, n_beams, lane, gx0, gy0, c, s);
                else ll = ns_score_batch<KIND, true, NS_UPD_P>(V, s_beams, n_beams, lane, gx0, gy0, c, s);
            }
        }
        else ll = ns_score_batch<KIND, false, NS_UPD_P>(V, s_beams, n_beams, lane, gx0, gy0, c, s);
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
// W_i = Q32 weight of particle i, prefix[i] = W_0 + ... + W_i over this shard, in two passes: tile sums (with group sums
// accumulated by integer atomics: exact, order-independent), then the prefix, every tile finding its own offset from the
// sums before it. The log-likelihoods were just written by the sensor-model kernel and are re-read from L2; HBM sees the
// 8-byte prefix writes. No spin-waits, no inter-block dependency inside a kernel.
constexpr int NS_SCAN_THREADS = 256;
constexpr int NS_SCAN_ITEMS = 16;                                    // per thread: 4 x float4
constexpr int NS_SCAN_WARP_ITEMS = 32 * NS_SCAN_ITEMS;               // a warp owns 512 consecutive particles
constexpr int NS_SCAN_TILE = NS_SCAN_THREADS * NS_SCAN_ITEMS;        // 4096
constexpr int NS_SCAN_GROUP = 64;                                    // tiles per group sum

// the global maximum log-likelihood lives in device memory as an order-preserving int (so atomicMax / NCCL max work)
__device__ __forceinline__ float ns_decode_max(const int* __restrict__ max_bits) {
    int b = *max_bits;
    b = b >= 0 ? b : b ^ 0x7fffffff;
    return __int_as_float(b);
}
// ns::det_exp_q32 (NS-4) for the device, same value for every input, without the three conversion (XU pipe) operations of
// the plain form: rintf + (int) become a magic-add whose bit pattern holds k, and the exact power-of-two scaling followed
// by the float -> uint64 truncation becomes a shift of the polynomial's significand.
__device__ __forceinline__ uint64_t ns_exp_q32_dev(float t) {
    const float tc = fminf(fmaxf(t, -30.f), 0.f);                     // keeps the bit tricks in range; NaN -> -30 (selected away below)
    const float km = ns::addf(ns::mulf(tc, 1.44269502f), NS_MAGIC);   // nearest-even integer k = rintf(tc * log2 e), in the low bits
    const float kf = ns::addf(km, -NS_MAGIC);
    float g = ns::fmaf_(-kf, 0.693145752f, tc);
    g = ns::fmaf_(-kf, 1.42860677e-06f, g);
    float p = 1.98412701e-04f;
    p = ns::fmaf_(p, g, 1.38888892e-03f);
    p = ns::fmaf_(p, g, 8.33333377e-03f);
    p = ns::fmaf_(p, g, 4.16666679e-02f);
    p = ns::fmaf_(p, g, 1.66666672e-01f);
    p = ns::fmaf_(p, g, 0.5f);
    p = ns::fmaf_(p, g, 1.0f);
    p = ns::fmaf_(p, g, 1.0f);                                        // exp(g) in [0.70, 1.42]
    // trunc(p * 2^(32+k)): p = m * 2^(e-23) with the 24-bit significand m, so the value is (m << 9) >> sh, sh = 32 - e - k.
    // m << 9 is the 33-bit number {1, m << 9 (32 bits)}; sh is 0 (p = 1, k = 0: the value 2^32) .. 33 (k = -32, p < 1: 0)
    const uint32_t pb = __float_as_uint(p);
    const int sh = 127 + NS_MAGIC_BITS - (int)(pb >> 23) - __float_as_int(km);
    uint32_t lo = __funnelshift_rc(pb << 9, 1u, (unsigned)sh);       // low word of (2^32 + (m << 9 mod 2^32)) >> min(sh, 32)
    lo = sh > 32 ? 0u : lo;
    const bool one = (sh == 0) | (t >= 0.f);                          // exactly 2^32
    lo = (one | !(t > -22.5f)) ? 0u : lo;                             // t <= -22.5 and NaN: below 2^-32
    return ((uint64_t)(one ? 1u : 0u) << 32) | lo;
}
__device__ __forceinline__ uint64_t ns_weight(float ll, float max_ll, float temper) {
    return ns_exp_q32_dev(ns::mulf(temper, ns::addf(ll, -max_ll)));
}
__device__ __forceinline__ uint64_t warp_scan_u64(uint64_t v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint64_t t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += t;
    }
    return v;
}
__device__ __forceinline__ uint64_t warp_sum_u64(uint64_t v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
// 32-byte store (sm_100: 256-bit global accesses)
__device__ __forceinline__ void st_v4_u64(uint64_t* p, uint64_t a, uint64_t b, uint64_t c, uint64_t d) {
    asm volatile("st.global.v4.b64 [%0], {%1, %2, %3, %4};" ::"l"(p), "l"(a), "l"(b), "l"(c), "l"(d) : "memory");
}
// Lane l of a warp holds the four float4 groups 4 (l + 32 j) .. + 3 (j = 0..3) of the warp's 512 particles: loads are
// 512-byte and stores 1-KiB contiguous per warp instruction, and the scan needs no shared-memory transpose.
__device__ __forceinline__ void ns_load_weights(const float* __restrict__ ll, int64_t n, int64_t base, float max_ll, float temper,
                                                uint64_t (&w)[4][4], uint64_t (&s)[4]) {
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const int64_t i = base + j * 128;
        float v[4] = {0.f, 0.f, 0.f, 0.f};
        if (i + 3 < n) { const float4 q = *reinterpret_cast<const float4*>(ll + i); v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w; }
        else { for (int e = 0; e < 4; e++) if (i + e < n) v[e] = ll[i + e]; }
        s[j] = 0;
#pragma unroll
        for (int e = 0; e < 4; e++) { w[j][e] = (i + e < n) ? ns_weight(v[e], max_ll, temper) : 0ull; s[j] += w[j][e]; }
    }
}
// pass 1: tile_sums[tile], group_sums[tile / 64] += (group_sums zeroed before the launch)
__global__ void __launch_bounds__(NS_SCAN_THREADS) k_ns_weights_sum(const float* __restrict__ ll, int64_t n, const int* __restrict__ max_bits,
                                                                   float temper, uint64_t* __restrict__ tile_sums, uint64_t* __restrict__ group_sums) {
    pdl_enter();
    __shared__ uint64_t s_warp[NS_SCAN_THREADS / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float max_ll = ns_decode_max(max_bits);
    const int64_t base = (int64_t)blockIdx.x * NS_SCAN_TILE + (int64_t)warp * NS_SCAN_WARP_ITEMS + lane * 4;
    uint64_t w[4][4], s[4];
    ns_load_weights(ll, n, base, max_ll, temper, w, s);
    const uint64_t tot = warp_sum_u64(s[0] + s[1] + s[2] + s[3]);
    if (lane == 0) s_warp[warp] = tot;
    __syncthreads();
    if (tid == 0) {
        uint64_t t = 0;
#pragma unroll
        for (int k = 0; k < NS_SCAN_THREADS / 32; k++) t += s_warp[k];
        tile_sums[blockIdx.x] = t;
        atomicAdd((unsigned long long*)(group_sums + blockIdx.x / NS_SCAN_GROUP), (unsigned long long)t);
    }
}
// pass 2: inclusive prefix per particle (local to this shard); block 0 also writes the shard total
__global__ void __launch_bounds__(NS_SCAN_THREADS) k_ns_weights_scan(const float* __restrict__ ll, int64_t n, const int* __restrict__ max_bits,
                                                                    float temper, const uint64_t* __restrict__ tile_sums,
                                                                    const uint64_t* __restrict__ group_sums, int n_tiles,
                                                                    uint64_t* __restrict__ prefix, uint64_t* __restrict__ total_out) {
    pdl_enter();
    __shared__ uint64_t s_warp[NS_SCAN_THREADS / 32], s_offp[NS_SCAN_THREADS / 32], s_tot[NS_SCAN_THREADS / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tile = blockIdx.x, group = tile / NS_SCAN_GROUP;
    // the tile's own weights first: the loads are in flight while the tile's offset is summed from the (L2-resident) sums
    const float max_ll = ns_decode_max(max_bits);
    const int64_t base = (int64_t)tile * NS_SCAN_TILE + (int64_t)warp * NS_SCAN_WARP_ITEMS + lane * 4;
    uint64_t w[4][4], s[4], incl[4];
    ns_load_weights(ll, n, base, max_ll, temper, w, s);
    // this tile's offset: whole groups before it + the tiles of its own group before it
    uint64_t part = 0;
    for (int g = tid; g < group; g += NS_SCAN_THREADS) part += group_sums[g];
    if (tid < tile - group * NS_SCAN_GROUP) part += tile_sums[group * NS_SCAN_GROUP + tid];
    part = warp_sum_u64(part);
    if (lane == 0) s_offp[warp] = part;
    uint64_t t = 0;
    if (tile == 0) {                                                 // shard total = all group sums
        const int n_groups = (n_tiles + NS_SCAN_GROUP - 1) / NS_SCAN_GROUP;
        for (int g = tid; g < n_groups; g += NS_SCAN_THREADS) t += group_sums[g];
        t = warp_sum_u64(t);
        if (lane == 0) s_tot[warp] = t;
    }
    uint64_t carry = 0;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        incl[j] = warp_scan_u64(s[j], lane) + carry;
        carry = __shfl_sync(0xffffffffu, incl[j], 31);
    }
    if (lane == 31) s_warp[warp] = carry;                            // warp total
    __syncthreads();
    if (tile == 0 && tid == 0) { uint64_t a = 0; for (int k = 0; k < NS_SCAN_THREADS / 32; k++) a += s_tot[k]; *total_out = a; }
    uint64_t off = 0;
#pragma unroll
    for (int k = 0; k < NS_SCAN_THREADS / 32; k++) { off += s_offp[k]; if (k < warp) off += s_warp[k]; }
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const int64_t i = base + j * 128;
        uint64_t run = off + incl[j] - s[j];
        uint64_t o[4];
#pragma unroll
        for (int e = 0; e < 4; e++) { run += w[j][e]; o[e] = run; }
        if (i + 3 < n) st_v4_u64(prefix + i, o[0], o[1], o[2], o[3]);
        else { for (int e = 0; e < 4; e++) if (i + e < n) prefix[i + e] = o[e]; }
    }
}
// Single pass (decoupled look-back): every tile computes its weights once, publishes its aggregate, finds its offset from
// the tiles before it and writes its prefix; tiles are taken in ticket order, so a tile only ever waits for tiles that
// are already running. State word per tile = status in bits 63:62 (0 none, 1 aggregate, 2 inclusive) | value (< 2^62):
// one 64-bit store/load, no fence needed. HBM sees the log-likelihoods once (4 B) and the prefix once (8 B).
constexpr uint64_t NS_LB_AGG = 1ull << 62, NS_LB_INC = 2ull << 62, NS_LB_VAL = (1ull << 62) - 1;
__device__ __forceinline__ uint64_t ld_volatile_u64(const uint64_t* p) {
    uint64_t v;
    asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_volatile_u64(uint64_t* p, uint64_t v) { asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory"); }
__global__ void __launch_bounds__(NS_SCAN_THREADS) k_ns_weights_scan1(const float* __restrict__ ll, int64_t n, const int* __restrict__ max_bits,
                                                                     float temper, uint64_t* __restrict__ tile_state /* n_tiles + 1, zeroed */,
                                                                     int n_tiles, uint64_t* __restrict__ prefix, uint64_t* __restrict__ total_out) {
    pdl_enter();
    __shared__ uint64_t s_warp[NS_SCAN_THREADS / 32];
    __shared__ uint64_t s_excl;
    __shared__ int s_tile;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_tile = (int)atomicAdd((unsigned long long*)(tile_state + n_tiles), 1ull);
    __syncthreads();
    const int tile = s_tile;
    const float max_ll = ns_decode_max(max_bits);
    const int64_t base = (int64_t)tile * NS_SCAN_TILE + (int64_t)warp * NS_SCAN_WARP_ITEMS + lane * 4;
    uint64_t w[4][4], s[4], incl[4];
    ns_load_weights(ll, n, base, max_ll, temper, w, s);
    uint64_t carry = 0;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        incl[j] = warp_scan_u64(s[j], lane) + carry;
        carry = __shfl_sync(0xffffffffu, incl[j], 31);
    }
    if (lane == 31) s_warp[warp] = carry;                            // warp total
    __syncthreads();
    if (warp == 0) {
        uint64_t agg = 0;
#pragma unroll
        for (int k = 0; k < NS_SCAN_THREADS / 32; k++) agg += s_warp[k];
        uint64_t excl = 0;
        if (tile > 0) {
            if (lane == 0) st_volatile_u64(tile_state + tile, NS_LB_AGG | agg);
            for (int j = tile - 1;; j -= 32) {                       // look back 32 tiles at a time
                const int idx = j - lane;
                uint64_t v;
                do { v = idx >= 0 ? ld_volatile_u64(tile_state + idx) : NS_LB_INC; } while (__any_sync(0xffffffffu, (v >> 62) == 0));
                const unsigned inc_mask = __ballot_sync(0xffffffffu, (v >> 62) == 2);
                const int stop = inc_mask ? __ffs((int)inc_mask) - 1 : 32;      // nearest tile with an inclusive prefix
                excl += warp_sum_u64(lane <= stop ? (v & NS_LB_VAL) : 0ull);
                if (inc_mask) break;
            }
        }
        if (lane == 0) {
            st_volatile_u64(tile_state + tile, NS_LB_INC | (excl + agg));
            s_excl = excl;
            if (tile == n_tiles - 1) *total_out = excl + agg;
        }
    }
    __syncthreads();
    uint64_t off = s_excl;
#pragma unroll
    for (int k = 0; k < NS_SCAN_THREADS / 32; k++) if (k < warp) off += s_warp[k];
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const int64_t i = base + j * 128;
        uint64_t run = off + incl[j] - s[j];
        uint64_t o[4];
#pragma unroll
        for (int e = 0; e < 4; e++) { run += w[j][e]; o[e] = run; }
        if (i + 3 < n) st_v4_u64(prefix + i, o[0], o[1], o[2], o[3]);
        else { for (int e = 0; e < 4; e++) if (i + e < n) prefix[i + e] = o[e]; }
    }
}
// The unnormalised weight W * 2^-32 (max particle = 1) into the particle records: only when somebody asks for the
// particles or the pose between update and resample (the filter loop itself never needs it).
__global__ void __launch_bounds__(256) k_ns_materialise_w(const float* __restrict__ ll, int64_t n, const int* __restrict__ max_bits, float temper,
                                                          float4* __restrict__ part) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    part[i].w = (float)((double)ns_weight(ll[i], ns_decode_max(max_bits), temper) * 2.3283064365386963e-10);
}

// ---- systematic resampling ------------------------------------------------------------------------------------------------
struct NsDest {
    float4* part[8];        // next-step particle buffer of every shard (own memory or mapped peer memory)
    int* anc[8];            // ancestor (global index) per output slot, same sharding
    int64_t per_rank;       // slots [r*per_rank, (r+1)*per_rank) live on shard r
    int world;
};
// The resampling plan of this shard: slots [k_lo, k_hi) are exactly those whose ancestor lives here (ns_plan.hpp).
// Slot k selects the first particle i with offset + prefix[i] > thr(k), thr(k) = floor(((k << 32) + u0) * total / (N << 32))
// (the integer form of u_k = (k + u0 / 2^32) / N against the normalised CDF). Moving one slot on adds total * 2^32 to the
// numerator, i.e. (dq1, dr1 << 32) in (quotient, remainder) form; thresholds are advanced with that, exactly.
struct NsPlan {
    uint64_t offset, total;
    int64_t k_lo, k_hi;
    uint64_t dq1, dr1;      // total = dq1 * N + dr1
    uint64_t dqs, drs;      // NS_RS_THREADS slots on: (NS_RS_THREADS * total) = dqs * N + drs
};
constexpr int NS_RS_THREADS = 256;
constexpr int NS_RS_ITEMS = 4;
constexpr int NS_RS_TILE = NS_RS_THREADS * NS_RS_ITEMS;             // output slots per tile
constexpr int NS_RS_CAP = 2560;                                      // prefix entries staged in shared memory per tile (8 CTAs/SM fit)

__device__ __forceinline__ void ns_plan_steps(NsPlan& p, uint64_t n_global) {
    p.dq1 = p.total / n_global; p.dr1 = p.total % n_global;
    // NS_RS_THREADS * total < 2^8 * 2^62: split so that nothing overflows
    const uint64_t x = p.dr1 * NS_RS_THREADS;                         // < 2^39
    p.dqs = p.dq1 * NS_RS_THREADS + x / n_global; p.drs = x % n_global;
}
// plan from the all-gathered Q32 totals, on the device (no host round trip)
__device__ __forceinline__ NsPlan ns_make_plan(const uint64_t* totals, int world, int rank, uint64_t n_global, uint32_t u0) {
    uint64_t off = 0, tot = 0;
    for (int r = 0; r < world; r++) { if (r < rank) off += totals[r]; tot += totals[r]; }
    NsPlan p;
    p.offset = off; p.total = tot;
    p.k_lo = tot ? ns::first_slot(off, tot, n_global, u0) : 0;
    p.k_hi = tot ? ns::first_slot(off + totals[rank], tot, n_global, u0) : 0;
    ns_plan_steps(p, n_global);
    return p;
}
__global__ void k_ns_plan(const uint64_t* __restrict__ totals, int world, int rank, uint64_t n_global, uint32_t u0, NsPlan* __restrict__ plan) {
    pdl_enter();
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    *plan = ns_make_plan(totals, world, rank, n_global, u0);
}

// ---- peer-memory exchange: the step's collectives without NCCL on the data path ---------------------------------------------
// Every shard owns a mailbox in its own HBM, mapped into the other shards (CUDA IPC, like the particle buffers). A shard
// POSTS by storing {payload, tag} into slot [its rank] of every peer's mailbox over NVLink (payload, then a system-scope
// release store of the tag) and COMPLETES by polling its own mailbox (system-scope acquire loads of local memory) until
// all `world` slots carry this step's tag. One 32-thread kernel per exchange: lane r talks to shard r. What an NCCL
// all-reduce(max) + all-gather + all-reduce(sum) + barrier cost in launches and protocol latency becomes three tiny
// kernels and three synchronisation points: the totals and the pose sums travel in one exchange, fused with the
// resampling plan the totals determine.
// Value slots are double-buffered by step parity and a shard cannot start step s+1's exchange before every shard posted
// step s's closing barrier, so a slot is never overwritten before it is read. Polls are bounded: a shard that never
// posts makes the others give up after 30 s (MCL_NS_EXCHANGE_TIMEOUT_S; 0 = wait for ever) and raise the mailbox status
// instead of hanging the GPU.
struct NsMailSlot {
    unsigned long long v[5];
    unsigned tag, pad[5];
};
static_assert(sizeof(NsMailSlot) == 64, "one slot per 64-byte line pair");
struct NsMailbox {
    NsMailSlot max_ll[2][8], total[2][8], pose[2][8];
    unsigned barrier[8];            // monotonic step tags
    int status;                     // != 0: an exchange timed out
    int pad[7];
};
struct NsPeers {
    NsMailbox* box[8];              // box[rank] = this shard's own mailbox
    int world, rank;
    unsigned long long timeout_ns;  // a poll gives up after this long (0: never)
};
__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) { asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_relaxed_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long ns_globaltimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
template <int WORDS>
__device__ __forceinline__ void ns_mail_post(NsMailSlot* slot_on_peer, const unsigned long long* payload, unsigned tag) {
#pragma unroll
    for (int w = 0; w < WORDS; w++) st_relaxed_sys(&slot_on_peer->v[w], payload[w]);
    st_release_sys(&slot_on_peer->tag, tag);
}
// wait until *flag reaches `tag` (AT_LEAST: monotonic barrier tags, compared modulo 2^32; else equality)
template <bool AT_LEAST>
__device__ __forceinline__ bool ns_mail_wait(const unsigned* flag, unsigned tag, int* status, unsigned long long timeout_ns) {
    const unsigned long long t0 = ns_globaltimer();
    for (unsigned spins = 0;; ++spins) {
        const unsigned seen = ld_acquire_sys(flag);
        if (AT_LEAST ? (int)(seen - tag) >= 0 : seen == tag) return true;
        if (timeout_ns && (spins & 1023u) == 1023u && ns_globaltimer() - t0 > timeout_ns) { atomicExch(status, 1); return false; }
        __nanosleep(32);
    }
}

// all-reduce(max) of the ordered-int maximum log-likelihood
__global__ void k_ns_xchg_max(int* __restrict__ max_bits, NsPeers P, unsigned tag, int parity) {
    const int r = threadIdx.x;
    int v = INT32_MIN;
    if (r < P.world) {
        NsMailbox* mine = P.box[P.rank];
        unsigned long long pay = (unsigned long long)(unsigned)(*max_bits);
        ns_mail_post<1>(&P.box[r]->max_ll[parity][P.rank], &pay, tag);
        if (ns_mail_wait<false>(&mine->max_ll[parity][r].tag, tag, &mine->status, P.timeout_ns)) v = (int)(unsigned)ld_relaxed_sys(&mine->max_ll[parity][r].v[0]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
    if (r == 0) *max_bits = v;
}
// all-gather of the shards' Q32 totals fused with the resampling plan they determine, and - in the same exchange - the
// all-reduce(sum) of the five weighted pose sums, added in rank order: the same bits on every shard.
// pose5[5] receives the mailbox status (0 = every exchange so far completed).
__global__ void k_ns_plan_xchg(const uint64_t* __restrict__ local_total, double* __restrict__ pose5, NsPeers P, unsigned tag, int parity,
                               uint64_t n_global, uint32_t u0, NsPlan* __restrict__ plan, uint64_t* __restrict__ totals_out) {
    __shared__ uint64_t s_tot[8];
    const int r = threadIdx.x;
    NsMailbox* mine = P.box[P.rank];
    bool ok = true;
    if (r < P.world) {
        unsigned long long pay[5];
#pragma unroll
        for (int k = 0; k < 5; k++) pay[k] = (unsigned long long)__double_as_longlong(pose5[k]);
        // the pose slot carries its own tag, so the totals' tag (release) is stored last and covers both payloads
        NsMailSlot* ps = &P.box[r]->pose[parity][P.rank];
#pragma unroll
        for (int k = 0; k < 5; k++) st_relaxed_sys(&ps->v[k], pay[k]);
        unsigned long long tot = *local_total;
        ns_mail_post<1>(&P.box[r]->total[parity][P.rank], &tot, tag);
        uint64_t t = 0;
        ok = ns_mail_wait<false>(&mine->total[parity][r].tag, tag, &mine->status, P.timeout_ns);
        if (ok) t = ld_relaxed_sys(&mine->total[parity][r].v[0]);
        s_tot[r] = t;
        totals_out[r] = t;
    }
    ok = __all_sync(0xffffffffu, ok);
    __syncwarp();                                                      // lane 0 reads what the other lanes' acquires made visible
    if (r == 0) {
        *plan = ns_make_plan(s_tot, P.world, P.rank, n_global, u0);
        for (int k = 0; k < 5; k++) {
            double s = 0.0;
            for (int q = 0; q < P.world; q++) s += __longlong_as_double((long long)ld_relaxed_sys(&mine->pose[parity][q].v[k]));
            pose5[k] = ok ? s : __longlong_as_double(0x7ff8000000000000ll);
        }
        pose5[5] = (double)(*(volatile int*)&mine->status);
    }
}
// closing barrier: every shard's resampled particles have landed in their owners' buffers before anybody swaps
__global__ void k_ns_xchg_barrier(NsPeers P, unsigned tag) {
    const int r = threadIdx.x;
    if (r < P.world) {
        NsMailbox* mine = P.box[P.rank];
        __threadfence_system();
        st_release_sys(&P.box[r]->barrier[P.rank], tag);
        ns_mail_wait<true>(&mine->barrier[r], tag, &mine->status, P.timeout_ns);
    }
}

// plan handed in by the host (phase-by-phase API)
__global__ void k_ns_plan_set(NsPlan p, uint64_t n_global, NsPlan* __restrict__ plan) { ns_plan_steps(p, n_global); *plan = p; }

struct NsThr {
    uint64_t q, r;          // ((k << 32) + u0) * total = q * (N << 32) + r
};
// by long division (once per output tile)
__device__ __forceinline__ NsThr ns_thr_of(uint64_t k, uint32_t u0, uint64_t total, uint64_t n_global) {
    const ns::U128 P = ns::mul64((k << 32) + u0, total);             // < 2^63 * 2^62
    // floor(P / (N << 32)) = floor((P >> 32) / N), in 32-bit limbs; the top two limbs are P.hi itself (< 2^61)
    const uint64_t lo = P.lo;
    const uint64_t q1 = P.hi / n_global;
    uint64_t rem = P.hi % n_global;
    const uint64_t t = (rem << 32) | (lo >> 32);
    const uint64_t q0 = t / n_global;
    rem = t % n_global;
    NsThr o;
    o.q = (q1 << 32) + q0;
    o.r = (rem << 32) | (lo & 0xffffffffull);
    return o;
}
// threshold `steps` (< 2^9) slots after `t`: exact, one small quotient (<= steps) found with an f64 estimate and fixed up
__device__ __forceinline__ NsThr ns_thr_ahead(const NsThr& t, uint32_t steps, uint64_t dq1, uint64_t dr1, uint64_t n_global, double inv_n) {
    const uint64_t x = (t.r >> 32) + (uint64_t)steps * dr1;           // < 2^31 + 2^9 * 2^31
    uint64_t c = (uint64_t)((double)x * inv_n);
    if (c * n_global > x) --c;
    else if ((c + 1) * n_global <= x) ++c;
    NsThr o;
    o.q = t.q + (uint64_t)steps * dq1 + c;
    o.r = ((x - c * n_global) << 32) | (t.r & 0xffffffffull);
    return o;
}
// first i in [lo, hi) with prefix[i] > v, else hi
__device__ __forceinline__ int64_t ns_search_global(const uint64_t* __restrict__ prefix, uint64_t v, int64_t lo, int64_t hi) {
    int64_t len = hi - lo;
    while (len > 0) {
        const int64_t half = len >> 1;
        if (!(prefix[lo + half] > v)) { lo += half + 1; len -= half + 1; } else len = half;
    }
    return lo;
}
struct NsTileHead {
    uint64_t q, r;          // threshold of the tile's first slot
    int i_lo;               // its ancestor (local index)
    int pad;
};

// first i in [lo, hi) with prefix[i] > v, else hi, by a whole warp: 32 evenly spaced probes per round (the prefix is
// non-decreasing, so their ballot is a prefix mask and its population count selects the sub-range): log32 instead of
// log2 dependent loads
__device__ __forceinline__ int64_t ns_search_warp(const uint64_t* __restrict__ prefix, uint64_t v, int64_t lo, int64_t hi, int lane) {
    int64_t len = hi - lo;
    while (len > 0) {
        if (len <= 32) {
            const bool below = lane < len && !(prefix[lo + lane] > v);
            lo += __popc(__ballot_sync(0xffffffffu, below));
            break;
        }
        const int64_t step = len >> 5;
        const int c = __popc(__ballot_sync(0xffffffffu, !(prefix[lo + (lane + 1) * step - 1] > v)));
        lo += c * step;
        len = c < 32 ? step - 1 : len - 32 * step;
    }
    return lo;
}
// pass 1 (merge-path partition): head[t] = threshold and ancestor of the first slot of output tile t (t = 0 .. tiles;
// the last entry describes the last slot instead). One warp per tile head.
__global__ void __launch_bounds__(256) k_ns_resample_bounds(const uint64_t* __restrict__ prefix, int64_t n_local, const NsPlan* __restrict__ plan,
                                                            uint64_t n_global, uint32_t u0, NsTileHead* __restrict__ head) {
    pdl_enter();
    const NsPlan P = *plan;
    const int64_t tiles = (P.k_hi - P.k_lo + NS_RS_TILE - 1) / NS_RS_TILE;
    if (tiles <= 0) return;
    const int lane = threadIdx.x & 31;
    const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t t = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; t <= tiles; t += warps) {
        int64_t k = P.k_lo + t * NS_RS_TILE;
        if (k >= P.k_hi) k = P.k_hi - 1;
        const NsThr th = ns_thr_of((uint64_t)k, u0, P.total, n_global);
        // by the plan thr >= offset for every slot of this shard; a smaller one would select the first particle
        int64_t i = th.q >= P.offset ? ns_search_warp(prefix, th.q - P.offset, 0, n_local, lane) : 0;
        if (i >= n_local) i = n_local - 1;
        if (lane == 0) {
            NsTileHead h;
            h.q = th.q; h.r = th.r; h.i_lo = (int)i; h.pad = 0;
            head[t] = h;
        }
    }
}
// pass 2: every tile of NS_RS_TILE consecutive output slots stages the prefix range its ancestors lie in (coalesced).
// Thread t resolves the four CONSECUTIVE slots 4t .. 4t+3: one branch-free binary search for the first, then short
// forward walks (ancestors are non-decreasing in the slot index: a merge); the ancestors are exchanged through shared
// memory so that gathers and stores run slot-strided (coalesced), each survivor going straight into the shard that owns
// its slot (own memory or a peer GPU's over NVLink). Ranges too long for shared memory are searched in global memory.
__global__ void __launch_bounds__(NS_RS_THREADS) k_ns_resample(const float4* __restrict__ src, const uint64_t* __restrict__ prefix, int64_t n_local,
                                                               int64_t g0, const NsPlan* __restrict__ plan, const NsTileHead* __restrict__ head,
                                                               uint64_t n_global, double inv_n, NsDest D, float new_weight) {
    pdl_enter();
    __shared__ uint64_t s_pre[NS_RS_CAP];
    __shared__ __align__(16) int s_anc[NS_RS_TILE];
    const NsPlan P = *plan;
    const int64_t tiles = (P.k_hi - P.k_lo + NS_RS_TILE - 1) / NS_RS_TILE;
    for (int64_t t = blockIdx.x; t < tiles; t += gridDim.x) {
        const int64_t k0 = P.k_lo + t * NS_RS_TILE;
        const NsTileHead H = head[t];
        const int i_lo = H.i_lo, i_hi = head[t + 1].i_lo;           // ancestors of this tile lie in [i_lo, i_hi]
        const int range = i_hi - i_lo + 1;
        const bool staged = range <= NS_RS_CAP;
        if (staged)
            for (int j = threadIdx.x; j < range; j += NS_RS_THREADS) s_pre[j] = prefix[i_lo + j];
        __syncthreads();
        NsThr th;
        th.q = H.q; th.r = H.r;
        th = ns_thr_ahead(th, NS_RS_ITEMS * threadIdx.x, P.dq1, P.dr1, n_global, inv_n);
        const uint64_t* pre = prefix + i_lo;
        int pos = 0;                                                  // entries of the range that do not select the current slot
        int anc[NS_RS_ITEMS];
#pragma unroll
        for (int it = 0; it < NS_RS_ITEMS; it++) {
            // by the plan thr >= offset for every slot of this shard; a smaller one would select the first particle
            const bool none = th.q < P.offset;
            const uint64_t v = th.q - P.offset;
            if (staged) {
                if (it == 0) {
#pragma unroll
                    for (int step = 2048; step > 0; step >>= 1) {    // NS_RS_CAP <= 4095
                        const int np = pos + step;
                        if (np <= range && !none && s_pre[np - 1] <= v) pos = np;
                    }
                } else {
                    int walked = 0;
                    while (pos < range && !none && s_pre[pos] <= v && walked < 8) { ++pos; ++walked; }
                    if (walked == 8) {
                        int len = range - pos;                        // long gap of (near-)weightless particles: search the rest
                        while (len > 0) {
                            const int half = len >> 1;
                            if (s_pre[pos + half] <= v) { pos += half + 1; len -= half + 1; } else len = half;
                        }
                    }
                }
            } else if (!none) {
                pos = (int)(ns_search_global(pre, v, pos, range));
            }
            anc[it] = i_lo + min(pos, range - 1);
            th.q += P.dq1;                                            // one slot on
            uint64_t rh = (th.r >> 32) + P.dr1;                       // < 2 N
            if (rh >= n_global) { rh -= n_global; th.q += 1; }
            th.r = (rh << 32) | (th.r & 0xffffffffull);
        }
        *reinterpret_cast<int4*>(s_anc + NS_RS_ITEMS * threadIdx.x) = make_int4(anc[0], anc[1], anc[2], anc[3]);
        __syncthreads();
        const int r0 = (int)min((int64_t)(D.world - 1), k0 / D.per_rank);       // shard of the tile's first slot
        const int64_t base0 = (int64_t)r0 * D.per_rank;
        int a[NS_RS_ITEMS];
        float4 p[NS_RS_ITEMS];
#pragma unroll
        for (int it = 0; it < NS_RS_ITEMS; it++) {
            const int j = threadIdx.x + it * NS_RS_THREADS;
            a[it] = (k0 + j < P.k_hi) ? s_anc[j] : -1;
        }
#pragma unroll
        for (int it = 0; it < NS_RS_ITEMS; it++)
            if (a[it] >= 0) p[it] = src[a[it]];
#pragma unroll
        for (int it = 0; it < NS_RS_ITEMS; it++) {
            if (a[it] >= 0) {
                const int64_t kk = k0 + threadIdx.x + it * NS_RS_THREADS;
                int r = r0;
                int64_t slot = kk - base0;
                while (slot >= D.per_rank && r < D.world - 1) { slot -= D.per_rank; ++r; }
                p[it].w = new_weight;
                D.part[r][slot] = p[it];
                D.anc[r][slot] = (int)(g0 + a[it]);
            }
        }
        __syncthreads();
    }
    if (D.world > 1) __threadfence_system();        // peer stores visible before the step's closing collective lets anyone read them
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

// out[k] = sum over blocks of partials[b*5+k] (one warp per output, fixed order)
// host5 (single-shard mcl_ns_step): the sums also go straight into the caller's pinned host block (zero-copy), so the
// step ends without a device-to-host copy command.
__global__ void k_ns_pose_reduce(const double* __restrict__ partials, int n_blocks, double* __restrict__ out5, double* __restrict__ host5) {
    pdl_enter();
    const int o = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (o >= 5) return;
    double s = 0;
    for (int b = lane; b < n_blocks; b += 32) s += partials[(size_t)b * 5 + o];
    s = warp_sum(s);
    if (lane == 0) { out5[o] = s; if (host5) host5[o] = s; }
}

// Weighted pose sums {sum w, sum w x, sum w y, sum w sin, sum w cos} per block, w = the unnormalised fp32 weight
// W * 2^-32 recomputed from the log-likelihood (ll != null: between update and resample) or read from the particle record.
__global__ void __launch_bounds__(256) k_ns_pose_partials(const float4* __restrict__ part, int64_t n, const float* __restrict__ ll,
                                                          const int* __restrict__ max_bits, float temper, double* __restrict__ partials) {
    pdl_enter();
    __shared__ double ws[8][5];
    double a[5] = {0, 0, 0, 0, 0};
    const float max_ll = ll ? ns_decode_max(max_bits) : 0.f;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        float4 p = part[i];
        float s, c;
        ns::det_sincosf32(p.z, s, c);          // fp32 evaluation (~1 ulp): NS-8 is tolerance-graded, the f64-evaluated form is not needed here
        const float wf = ll ? (float)((double)ns_weight(ll[i], max_ll, temper) * 2.3283064365386963e-10) : p.w;
        const double w = (double)wf;
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
