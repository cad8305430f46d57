// kernels_ref.cuh — MCL_MODE_REF kernels: results identical to the reference's CPU filter
// (pink_fundamentals/src/monte_carlo.cpp, "MC") for the same particles, scan, motion and injected draws.
//
// All parity-critical f64 arithmetic goes through __dmul_rn/__dadd_rn/__ddiv_rn so it can never be contracted
// into FMAs; the translation unit is additionally compiled with -fmad=false.
#pragma once
#include "mcl_device.cuh"

namespace mcl {

// One used beam (every beam_stride-th of the FOV-filtered list, MC:650-657), prepared on the host in f64.
struct RefBeam {
    double off_deg;    // -(angle) * 180.0 / M_PI                           (MC:653)
    double obs;        // observed_distance                                   (MC:657)
    double rand_term;  // w_rand * (|obs - max_range| < 0.01 ? 1.0 : 0.0)     (MC:669)
};

constexpr int RU_INLINE_BEAMS = 40;        // (the reference scores 12 beams of its 360-beam scan, 35 of its 683-beam one)

struct RefParams {
    // occupancy: 1 byte per cell, 1 = value > 50 (MC:327,377)
    const uint8_t* occ;
    int width, height;
    int map_in_smem;           // stage occ into shared memory (w*h small enough)
    // bordered ray-march table (k_ref_update_v2's fp32 path): (width + 2 pad) x (height + 2 pad) bytes, 0 free, 1 occupied,
    // 2 outside the grid; row/column -1 replicate row/column 0 (the reference's cast truncates (-1,0) to cell 0, Q7)
    const uint8_t* occ_pad;
    int pad, wp;
    double res, inv_res;       // (double)float32 resolution (Q10) and its rounded reciprocal
    double ox, oy;             // origin (MC:300-301)
    double max_x, max_y;       // isInsideMap upper bounds (MC:688-689)
    double laser_offset;       // MC:631
    double validity_offset;    // MC:333
    double max_range;          // MC:628
    double w_hit;              // MC:180
    // ray radii r = 0, step, 2*step ... accumulated in f64 exactly like `r += step` (MC:372)
    const double* radii;
    int n_radii;
    // Gaussian LUT (MC:139-177)
    const double* gauss;
    int gauss_size;
    double gauss_res, gauss_min, gauss_max;
    // ray direction LUT (MC:192, 355-366): entry k holds (cos,sin) for angle_key = key_min + k
    const double2* lut;
    const uint8_t* lut_filled;
    int key_min, n_keys;
    // beams
    const RefBeam* beams;
    int n_beams;
    int trig;                  // TRIG_*: how cosf/sinf of MC:644-645 are evaluated (mcl_device.cuh)
    // mcl_step: updateParticlePos (MC:740-755) applied to every particle as it is loaded, so the tick has no predict pass of
    // its own (k_ref_update_v2 only; the particle is written back whole, with its weight)
    int do_predict;
    float rot1, trans, dtheta;
    // beams == null: the scored beams ride in the launch parameters (k_ref_update_v2 only), so a tick whose scan arrives from the
    // host needs no copy command ahead of its first kernel
    RefBeam inline_beams[RU_INLINE_BEAMS];
    // optimistic tick (mcl_step): non-null and *abort != 0 = the first-touch pre-pass of this tick found ray directions that the
    // host has to evaluate first; every kernel of the tick returns at once and the host runs the tick again
    const int* abort;
};
__device__ __forceinline__ RefBeam ref_beam_of(const RefParams& P, int i) { return P.beams ? P.beams[i] : P.inline_beams[i]; }

// ---- map probes ---------------------------------------------------------------------------------------
struct MapView {
    const uint8_t* occ;
    int width, height;
    double res, inv_res, ox, oy;
    // worldToMap + getCell > 50 (MC:298-319). Returns 0 free, 1 occupied, -1 outside the grid.
    __device__ __forceinline__ int probe(double wx, double wy) const {
        int mx = cell_of(dsub(wx, ox), res, inv_res);
        int my = cell_of(dsub(wy, oy), res, inv_res);
        if (mx < 0 || my < 0 || mx >= width || my >= height) return -1;
        return occ[my * width + mx];
    }
};

// isValidPos (MC:331-349): inside the map and none of the 9 stencil points occupied.
__device__ __forceinline__ bool ref_is_valid(const MapView& m, const RefParams& P, double x, double y) {
    if (!((x >= P.ox && x < P.max_x) && (y >= P.oy && y < P.max_y))) return false;
    const double o = P.validity_offset;
    const double offx[9] = {0, o, 0, -o, 0, o, o, -o, -o};
    const double offy[9] = {0, 0, o, 0, -o, o, -o, o, -o};
#pragma unroll
    for (int k = 0; k < 9; k++)
        if (m.probe(dadd(x, offx[k]), dadd(y, offy[k])) == 1) return false;
    return true;
}

// tf::getYaw(tf::createQuaternionMsgFromYaw(theta)) in the library's operation order (Q8), then degrees.
__device__ __forceinline__ double ref_yaw(float theta) {
    double h = dmul((double)theta, 0.5);
    double sy, cy;
    sincos(h, &sy, &cy);
    double d = dadd(dmul(sy, sy), dmul(cy, cy));
    double s = ddiv(2.0, d);
    double zs = dmul(sy, s);
    double wz = dmul(cy, zs);
    double zz = dmul(sy, zs);
    double m00 = dsub(1.0, zz);
    return atan2(wz, m00);
}

// tf::getYaw(tf::createQuaternionMsgFromYaw(theta)) in degrees, the cheap way. Mathematically the round trip returns theta
// wrapped to [-pi, pi]; numerically it lands within ~1e-15 rad of it. The only consumer is (int)round(yaw_deg + off_deg), so
// z = theta - k 2 pi (two-term 2 pi, fused) is used wherever the wrap is unambiguous, and the caller redoes the round trip for
// the rays whose sum lies within 1e-9 of a rounding boundary. exact = true: the full round trip was evaluated here (near
// +-pi, where atan2 may return either sign, for huge |theta|, NaN and Inf).
__device__ __forceinline__ double ref_yaw_deg_fast(float theta, bool& exact) {
    const double PI = 3.14159265358979323846;
    const double t = (double)theta;
    const double k = rint(dmul(t, 0.15915494309189535));
    const double z = fma(-k, 2.4492935982947064e-16, fma(-k, 6.283185307179586, t));
    exact = !(fabs(t) < 1.0e5 && fabs(fabs(z) - PI) > 1e-9);
    if (exact) return ddiv(dmul(ref_yaw(theta), 180.0), PI);
    return dmul(z, 57.295779513082323);
}

// GaussianLookup::get (MC:154-168)
__device__ __forceinline__ double ref_gauss(const RefParams& P, double diff) {
    if (diff < P.gauss_min || diff > P.gauss_max) return 0.0;
    double index_f = ddiv(dsub(diff, P.gauss_min), P.gauss_res);
    int index = trunc_x86(index_f);
    if (index + 1 < P.gauss_size) {
        double w = dsub(index_f, (double)index);
        return dadd(dmul(dsub(1.0, w), __ldg(P.gauss + index)), dmul(w, __ldg(P.gauss + index + 1)));
    }
    return __ldg(P.gauss + index);
}

struct RefSmem {
    double2* lut;
    RefBeam* beams;
    double* radii;
    uint8_t* occ;
};

__device__ __forceinline__ RefSmem ref_stage_smem(const RefParams& P, unsigned char* raw, bool stage_map) {
    RefSmem s;
    s.lut = reinterpret_cast<double2*>(raw);
    s.beams = reinterpret_cast<RefBeam*>(s.lut + P.n_keys);
    s.radii = reinterpret_cast<double*>(s.beams + P.n_beams);
    s.occ = reinterpret_cast<uint8_t*>(s.radii + P.n_radii);
    for (int i = threadIdx.x; i < P.n_keys; i += blockDim.x) s.lut[i] = P.lut[i];
    for (int i = threadIdx.x; i < P.n_beams; i += blockDim.x) s.beams[i] = ref_beam_of(P, i);
    for (int i = threadIdx.x; i < P.n_radii; i += blockDim.x) s.radii[i] = P.radii[i];
    if (stage_map) {
        int cells = P.width * P.height;
        for (int i = threadIdx.x; i < cells; i += blockDim.x) s.occ[i] = P.occ[i];
    }
    __syncthreads();
    return s;
}

// angle_key = (int)round(yaw_deg + off_deg) (MC:352-355) as an index into the LUT arrays.
__device__ __forceinline__ int ref_key_index(const RefParams& P, double yaw_deg, double off_deg) {
    int key = trunc_x86(round(dadd(yaw_deg, off_deg)));
    int k = key - P.key_min;
    return (k < 0 || k >= P.n_keys) ? -1 : k;
}

// ---- first-touch pre-pass (Q9) ---------------------------------------------------------------------------
// The reference memoises the direction of a missing key from whichever (particle j ascending, beam i ascending)
// touches it first. For every still-unfilled key find that first toucher: touch[k] = min over (j<<32 | beam).
__global__ void __launch_bounds__(256) k_ref_first_touch(const float4* __restrict__ part, int64_t n, RefParams P,
                                                         unsigned long long* __restrict__ touch) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    RefSmem S = ref_stage_smem(P, smem_raw, P.map_in_smem != 0);
    MapView m{P.map_in_smem ? S.occ : P.occ, P.width, P.height, P.res, P.inv_res, P.ox, P.oy};
    int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    float4 p = part[j];
    if (P.do_predict) {                  // the motion the computeWeight kernel will apply (MC:746-753), here without writing it back
        const float h = __fadd_rn(p.z, P.rot1);
        p.x = __fadd_rn(p.x, __fmul_rn(P.trans, ref_cosf(h, P.trig)));
        p.y = __fadd_rn(p.y, __fmul_rn(P.trans, ref_sinf(h, P.trig)));
        p.z = __fadd_rn(p.z, P.dtheta);
    }
    if (!ref_is_valid(m, P, (double)p.x, (double)p.y)) return;
    double yaw_deg = ddiv(dmul(ref_yaw(p.z), 180.0), 3.14159265358979323846);
    for (int b = 0; b < P.n_beams; b++) {
        int k = ref_key_index(P, yaw_deg, S.beams[b].off_deg);
        if (k >= 0 && !P.lut_filled[k])
            atomicMin(&touch[k], ((unsigned long long)j << 32) | (unsigned)b);
    }
}

// theta of each first toucher, so the host can evaluate the direction with the same libm as the CPU filter.
__global__ void k_ref_touch_theta(const float4* __restrict__ part, const unsigned long long* __restrict__ touch, int n_keys,
                                  float* __restrict__ theta_out) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_keys) return;
    unsigned long long t = touch[k];
    theta_out[k] = (t == ~0ull) ? 0.f : part[t >> 32].z;
}
// The same for an optimistic tick (one block): first touchers and their theta (after the tick's motion) go straight into the
// caller's pinned host block, the touch table is left empty for the next pre-pass, and *abort tells the kernels that follow
// whether any key was found (then they return at once: the host evaluates those directions and runs the tick again).
struct RefTouchReport { unsigned long long seq; int n_new; int pad; };      // followed by touch[n_keys] (u64) and theta[n_keys] (f32)
__global__ void __launch_bounds__(1024) k_ref_touch_report(const float4* __restrict__ part, unsigned long long* __restrict__ touch, int n_keys,
                                                           int do_predict, float dtheta, RefTouchReport* __restrict__ report,
                                                           unsigned long long* __restrict__ h_touch, float* __restrict__ h_theta,
                                                           unsigned long long seq, int* __restrict__ abort) {
    int mine = 0;
    for (int k = threadIdx.x; k < n_keys; k += blockDim.x) {
        const unsigned long long t = touch[k];
        if (t != ~0ull) {
            float th = part[t >> 32].z;
            if (do_predict) th = __fadd_rn(th, dtheta);
            h_touch[k] = t; h_theta[k] = th;
            touch[k] = ~0ull;
            mine++;
        } else h_touch[k] = ~0ull;
    }
    const int total = __syncthreads_count(mine != 0);            // threads that found something: zero = nothing new
    if (threadIdx.x == 0) {
        *abort = total != 0 ? 1 : 0;
        report->n_new = total;
        __threadfence_system();
        *(volatile unsigned long long*)&report->seq = seq;
    }
}

// ---- computeWeight (MC:623-682): one thread per particle ---------------------------------------------------
__global__ void __launch_bounds__(256) k_ref_update(float4* __restrict__ part, float* __restrict__ w_dense, int64_t n, RefParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    RefSmem S = ref_stage_smem(P, smem_raw, P.map_in_smem != 0);
    MapView m{P.map_in_smem ? S.occ : P.occ, P.width, P.height, P.res, P.inv_res, P.ox, P.oy};
    int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    float4 p = part[j];
    double prob = 0.0;
    if (ref_is_valid(m, P, (double)p.x, (double)p.y)) {                                   // MC:648
        double posx = dadd((double)p.x, dmul(P.laser_offset, (double)ref_cosf(p.z, P.trig)));   // MC:644
        double posy = dadd((double)p.y, dmul(P.laser_offset, (double)ref_sinf(p.z, P.trig)));   // MC:645
        double yaw_deg = ddiv(dmul(ref_yaw(p.z), 180.0), 3.14159265358979323846);         // MC:351-352
        for (int b = 0; b < P.n_beams; b++) {
            RefBeam bm = S.beams[b];
            int k = ref_key_index(P, yaw_deg, bm.off_deg);
            double2 dir = (k >= 0) ? S.lut[k] : make_double2(0.0, 0.0);
            double expected = P.max_range;                                                // MC:389
            for (int s = 0; s < P.n_radii; s++) {                                         // MC:372
                double r = S.radii[s];
                int c = m.probe(dadd(posx, dmul(r, dir.x)), dadd(posy, dmul(r, dir.y)));
                if (c < 0) break;                                                         // MC:376
                if (c == 1) { expected = r; break; }                                      // MC:377-381
            }
            double diff = fabs(dsub(bm.obs, expected));                                   // MC:662
            prob = dadd(prob, dmul(P.w_hit, ref_gauss(P, diff)));                         // MC:665
            prob = dadd(prob, bm.rand_term);                                              // MC:669
        }
    }
    const float wf = __double2float_rn(prob);                                             // MC:673
    part[j].w = wf;
    w_dense[j] = wf;
}

// ---- computeWeight, restructured: a block takes 256 particles; rays, not particles, are the unit of parallel work -----------
// The one-thread-per-particle kernel above is issue-bound with only ~70 % of lanes active: invalid particles idle for the
// whole ray march and ray lengths vary between lanes. Here:
//   phase A (thread = particle)   validity stencil (3+3 cell lookups instead of 18), laser origin, yaw; valid particles are
//                                 compacted into shared memory
//   phase B (thread = ray)        the (valid particle, beam) pairs are spread evenly over the block; each marches its ray
//                                 and leaves w_hit*gauss in shared memory
//   phase C (thread = particle)   the per-beam terms are added in the reference's order (f64, left to right) -> weight
// Cell lookups avoid both the IEEE division and the XU-pipe conversions, and their fast path is branch-free: q = a*(1/res)
// is rounded with the 2^52+2^51 magic add, accepted when it is farther than 1e-6 from an integer
// (|q - a/res| <= 3.4e-16*|q|), and resolved by the exact division only otherwise. Results are bit-identical to the kernel
// above (and to the CPU).
constexpr int RU_TILE = 256;
constexpr double RU_MAGIC = 6755399441055744.0;       // 2^52 + 2^51

// static_cast<int>((w - o) / res) along one axis, branch-free fast path. q = (w-o)*(1/res) differs from the correctly
// rounded quotient by at most 3.4e-16*|q|; rounding q to the nearest integer with the magic add and looking at the
// remainder d tells floor(q) whenever |d| > 1e-6, and truncation toward zero (what the reference's cast does, Q7) is
// floor + 1 for negative non-integers. `ok` is false when the shortcut cannot be trusted (within 1e-6 of an integer,
// |q| >= 1.9e9, NaN): the caller then redoes the probe with the IEEE division.
// BOUNDED: the caller guarantees |q| < 2^31 (a point within max_range of a particle that is inside the map), so only the
// remainder test is needed; NaN fails it and takes the exact path. ZERO_ORIGIN: o == 0 (w - 0.0 == w except for -0.0,
// whose cell is 0 either way).
template <bool BOUNDED, bool ZERO_ORIGIN>
__device__ __forceinline__ int cell_fast(double w, double o, double inv_res, bool& ok) {
    const double q = dmul(ZERO_ORIGIN ? w : dsub(w, o), inv_res);
    const double t = dadd(q, RU_MAGIC);
    const double d = dsub(q, dsub(t, RU_MAGIC));
    ok = BOUNDED ? (fabs(d) > 1e-6) : ((fabs(d) > 1e-6) & (fabs(q) < 1.9e9));
    const int fl = __double2loint(t) - (d < 0.0 ? 1 : 0);      // floor(q)
    return fl + (int)((unsigned)fl >> 31);                      // trunc(q) for non-integers
}
// One map probe (worldToMap + bounds + getCell > 50, MC:298-319): 0 free, 1 occupied, -1 outside the grid.
template <bool ZERO_ORIGIN>
__device__ __forceinline__ int probe_fast(const uint8_t* __restrict__ occ, double wx, double wy, double ox, double oy, double res,
                                          double inv_res, int W, int H) {
    bool okx, oky;
    int mx = cell_fast<true, ZERO_ORIGIN>(wx, ox, inv_res, okx);
    int my = cell_fast<true, ZERO_ORIGIN>(wy, oy, inv_res, oky);
    if (__builtin_expect(!(okx & oky), 0)) {                    // rare: exact path
        mx = trunc_x86(ddiv(dsub(wx, ox), res));
        my = trunc_x86(ddiv(dsub(wy, oy), res));
    }
    if ((unsigned)mx >= (unsigned)W || (unsigned)my >= (unsigned)H) return -1;
    return occ[my * W + mx];
}

// fp32 pre-filter for the same lookup. The probe's cell coordinate q = ((p + r*dir) - o)/res is first evaluated in fp32
// from per-ray fp32 copies (q0 = (p-o)/res, dq = dir/res) with one FFMA; its error against the f64 quantity the reference
// truncates is below 2^-23*(W + 2*max_range/res + 2). When q sits farther than four times that from every integer, its
// truncation IS the reference's cell and no f64 instruction is spent; otherwise (a fraction ~1e-4..1e-2 of the probes)
// the f64 path above decides. Either way the result is the reference's, bit for bit.
constexpr float RU_MAGICF = 12582912.0f;     // 1.5 * 2^23
constexpr int RU_MAGICF_BITS = 0x4B400000;
__device__ __forceinline__ int cell_fast32(float q, float tol, bool& ok) {
    const float t = __fadd_rn(q, RU_MAGICF);
    const float d = __fadd_rn(q, -__fadd_rn(t, -RU_MAGICF));
    ok = fabsf(d) > tol;                                          // NaN fails
    const int fl = (__float_as_int(t) - RU_MAGICF_BITS) - (d < 0.f ? 1 : 0);
    return fl + (int)((unsigned)fl >> 31);
}

struct RuSmem {
    const double2* lut;   // f64 ray directions: shared memory, or (fp32 march) the global table itself, read by the ~1e-3 undecided rays only
    RefBeam* beams;
    double* radii;
    double* posx;
    double* posy;
    double* yawd;
    double* terms;
    int* vlist;
    float* radii_f;
    uint8_t* occ;
    uint8_t* occ_pad;
    float2* dq;           // per ray-table key: fp32 direction in cells per metre (dir / res)
    float2* q0;           // per valid particle: fp32 laser origin in cell units, minus 0.5
    double* gterm;        // [beam][hit index 0..n_radii]: w_hit * GaussianLookup(|obs - expected|), last = no hit (max range)
    float* yawf;          // per valid particle: (float)yaw_deg, for the fp32 pre-decision of the ray-table key
    float* offf;          // per beam: (float)off_deg
    float* wout;          // per particle of the tile: its weight, handed from the slot's thread back to the particle's thread
};
// map_bytes = plain + bordered table bytes when they are staged in shared memory, else 0
// lut_in_smem: false for the fp32 march, which leaves the f64 direction table (16 B per key, 38 KB for the reference's
// 2401 keys) in global memory: 60 -> 22 KB per block, four resident blocks per SM instead of three.
__host__ __device__ inline size_t ru_smem_bytes(int n_keys, int n_beams, int n_radii, size_t map_bytes, size_t pad_bytes, bool lut_in_smem) {
    return (lut_in_smem ? (size_t)n_keys * 16 : 0) + (size_t)n_beams * 24 + (size_t)n_radii * 8 + 3 * RU_TILE * 8 + (size_t)RU_TILE * (n_beams + 1) * 8 +
           RU_TILE * 4 + (((size_t)n_radii * 4 + 15) & ~(size_t)15) + ((map_bytes + 15) & ~(size_t)15) + ((pad_bytes + 15) & ~(size_t)15) +
           (size_t)n_keys * 8 + RU_TILE * 8 + (size_t)n_beams * (n_radii + 1) * 8 + RU_TILE * 4 + (((size_t)n_beams * 4 + 15) & ~(size_t)15) + RU_TILE * 4;
}

// NR: number of ray steps known at compile time (11 for the reference's 1.0 m / 0.1 m), 0 = run-time count (<= 16)
// MAP_SMEM: the bordered ray-march table is in shared memory (maps up to ~30 KB): probes are addressed with 32-bit shared
// addresses (IMAD + IADD + LDS.U8) instead of through a generic 64-bit pointer.
template <bool ZERO_ORIGIN, bool FAST32, int NR, bool MAP_SMEM>
__global__ void __launch_bounds__(RU_TILE) k_ref_update_v2(float4* __restrict__ part, float* __restrict__ w_dense, int64_t n, RefParams P,
                                                           uint32_t div_magic /* ceil(2^32 / n_beams) */, float tol32,
                                                           int ppb /* particles a block takes per pass, <= RU_TILE: fewer for small filters,
                                                                      so that their rays spread over more SMs */) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ int warp_cnt[RU_TILE / 32];
    RuSmem S;
    double2* s_lut = reinterpret_cast<double2*>(smem_raw);
    S.lut = FAST32 ? P.lut : s_lut;
    S.beams = reinterpret_cast<RefBeam*>(s_lut + (FAST32 ? 0 : P.n_keys));
    S.radii = reinterpret_cast<double*>(S.beams + P.n_beams);
    S.posx = S.radii + P.n_radii;
    S.posy = S.posx + RU_TILE;
    S.yawd = S.posy + RU_TILE;
    S.terms = S.yawd + RU_TILE;
    S.vlist = reinterpret_cast<int*>(S.terms + (size_t)RU_TILE * (P.n_beams + 1));
    S.radii_f = reinterpret_cast<float*>(S.vlist + RU_TILE);
    S.occ = reinterpret_cast<uint8_t*>(S.radii_f + ((P.n_radii + 3) & ~3));
    S.occ_pad = S.occ + ((P.map_in_smem ? (size_t)P.width * P.height + 15 : 0) & ~(size_t)15);
    S.dq = reinterpret_cast<float2*>(S.occ_pad + ((P.map_in_smem ? (size_t)P.wp * (P.height + 2 * P.pad) + 15 : 0) & ~(size_t)15));
    S.q0 = S.dq + P.n_keys;
    S.gterm = reinterpret_cast<double*>(S.q0 + RU_TILE);
    S.yawf = reinterpret_cast<float*>(S.gterm + (size_t)P.n_beams * (P.n_radii + 1));
    S.offf = S.yawf + RU_TILE;
    S.wout = S.offf + ((P.n_beams + 3) & ~3);
    for (int i = threadIdx.x; i < P.n_beams; i += RU_TILE) S.offf[i] = __double2float_rn(ref_beam_of(P, i).off_deg);
    for (int i = threadIdx.x; i < P.n_radii; i += RU_TILE) S.radii_f[i] = __double2float_rn(P.radii[i]);
    if (!FAST32) for (int i = threadIdx.x; i < P.n_keys; i += RU_TILE) s_lut[i] = P.lut[i];
    for (int i = threadIdx.x; i < P.n_beams; i += RU_TILE) S.beams[i] = ref_beam_of(P, i);
    for (int i = threadIdx.x; i < P.n_radii; i += RU_TILE) S.radii[i] = P.radii[i];
    if (P.map_in_smem) {
        const int cells = P.width * P.height;
        for (int i = threadIdx.x; i < cells; i += RU_TILE) S.occ[i] = P.occ[i];
        const int cells_p = P.wp * (P.height + 2 * P.pad);
        for (int i = threadIdx.x; i < cells_p; i += RU_TILE) S.occ_pad[i] = P.occ_pad[i];
    }
    __syncthreads();
    if (FAST32) {
        // per key: the direction in fp32 cell units; per (beam, step at which the ray ended): the beam's score term. A ray
        // ends at one of n_radii + 1 distances, so GaussianLookup::get (MC:154-168) is evaluated here once per pair.
        for (int i = threadIdx.x; i < P.n_keys; i += RU_TILE)
            { const double2 d = P.lut[i]; S.dq[i] = make_float2(__double2float_rn(dmul(d.x, P.inv_res)), __double2float_rn(dmul(d.y, P.inv_res))); }
        const int per = P.n_radii + 1;
        for (int i = threadIdx.x; i < P.n_beams * per; i += RU_TILE) {
            const int b = i / per, k = i - b * per;
            const double expected = k < P.n_radii ? S.radii[k] : P.max_range;            // MC:379 / MC:389
            S.gterm[i] = dmul(P.w_hit, ref_gauss(P, fabs(dsub(S.beams[b].obs, expected))));   // MC:662-665
        }
        __syncthreads();
    }
    const uint8_t* occ = P.map_in_smem ? S.occ : P.occ;
    const uint8_t* occ_pad = P.map_in_smem ? S.occ_pad : P.occ_pad;
    // raw magic-add bit patterns -> bordered index: (by - MB + pad) * wp + (bx - MB + pad), constants folded (mod 2^32)
    const uint32_t pad_fold = (uint32_t)(P.pad - RU_MAGICF_BITS) * ((uint32_t)P.wp + 1u);
    const uint32_t wp_u = (uint32_t)P.wp;
    const float lim32 = 0.5f - tol32;
    const uint32_t pad_last = (uint32_t)P.wp * (uint32_t)(P.height + 2 * P.pad) - 1u;
    const uint32_t pad_saddr = (uint32_t)__cvta_generic_to_shared(S.occ_pad) + pad_fold;      // MAP_SMEM: shared byte address of cell (0,0) + folded constants
    // constants in registers
    const double res = P.res, inv_res = P.inv_res, ox = P.ox, oy = P.oy;
    const int W = P.width, H = P.height;
    const int nb = P.n_beams, nr = P.n_radii, stride = nb + 1;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t n_tiles = (n + ppb - 1) / ppb;
    // Everything above read only tables that no kernel of a tick writes (map, ray directions, beams, Gaussian table: set by
    // host copies, which serialise the stream), so under a programmatic launch it overlaps the predecessor's tail; the
    // particles are the predecessor's output.
    pdl_enter();
    if (P.abort != nullptr && *P.abort != 0) return;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        // ---- phase A -------------------------------------------------------------------------------------------------
        const int64_t j = tile * ppb + threadIdx.x;
        const bool mine = (int)threadIdx.x < ppb && j < n;                                // this thread holds a particle of the tile
        bool valid = false, yaw_exact = false;
        double posx = 0, posy = 0, yawd = 0;
        float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
        if (mine) {
            p = part[j];
            if (P.do_predict) {                                                          // updateParticlePos, MC:746-753 (fp32 element math)
                const float h = __fadd_rn(p.z, P.rot1);
                p.x = __fadd_rn(p.x, __fmul_rn(P.trans, ref_cosf(h, P.trig)));           // MC:747
                p.y = __fadd_rn(p.y, __fmul_rn(P.trans, ref_sinf(h, P.trig)));           // MC:748
                p.z = __fadd_rn(p.z, P.dtheta);                                          // MC:753
            }
            const double x = (double)p.x, y = (double)p.y;
            if ((x >= ox && x < P.max_x) && (y >= oy && y < P.max_y)) {                  // isInsideMap, MC:685-692
                const double o = P.validity_offset;
                const double xs[3] = {dadd(x, -o), x, dadd(x, o)}, ys[3] = {dadd(y, -o), y, dadd(y, o)};
                int cx[3], cy[3];
                bool okall = true;
#pragma unroll
                for (int a = 0; a < 3; a++) {
                    bool k1, k2;
                    cx[a] = cell_fast<false, false>(xs[a], ox, inv_res, k1);
                    cy[a] = cell_fast<false, false>(ys[a], oy, inv_res, k2);
                    okall = okall & k1 & k2;
                }
                if (__builtin_expect(!okall, 0)) {
#pragma unroll
                    for (int a = 0; a < 3; a++) { cx[a] = trunc_x86(ddiv(dsub(xs[a], ox), res)); cy[a] = trunc_x86(ddiv(dsub(ys[a], oy), res)); }
                }
                bool hit = false;                                                        // any of the 9 stencil points occupied
#pragma unroll
                for (int a = 0; a < 3; a++)
#pragma unroll
                    for (int b = 0; b < 3; b++)
                        if ((unsigned)cx[a] < (unsigned)W && (unsigned)cy[b] < (unsigned)H && occ[cy[b] * W + cx[a]]) hit = true;
                valid = !hit;
            }
            if (valid) {
                float snf, csf;
                if (P.trig == TRIG_CR) {
                    double sn, cs;
                    sincos((double)p.z, &sn, &cs);
                    snf = __double2float_rn(sn); csf = __double2float_rn(cs);             // correctly rounded sinf / cosf
                } else { snf = ref_sinf(p.z, P.trig); csf = ref_cosf(p.z, P.trig); }        // the host libm's sinf / cosf
                posx = dadd(x, dmul(P.laser_offset, (double)csf));                        // MC:644
                posy = dadd(y, dmul(P.laser_offset, (double)snf));                        // MC:645
                yawd = ref_yaw_deg_fast(p.z, yaw_exact);                                   // MC:351-352 (see ref_yaw_deg_fast)
            }
        }
        const unsigned bal = __ballot_sync(0xffffffffu, valid);
        if (lane == 0) warp_cnt[warp] = __popc(bal);
        __syncthreads();
        int slot = __popc(bal & ((1u << lane) - 1u)), nv = 0;
        for (int w = 0; w < RU_TILE / 32; w++) { if (w < warp) slot += warp_cnt[w]; nv += warp_cnt[w]; }
        if (valid) {
            S.vlist[slot] = threadIdx.x | (yaw_exact ? 0x10000 : 0); S.posx[slot] = posx; S.posy[slot] = posy; S.yawd[slot] = yawd;
            S.yawf[slot] = __double2float_rn(yawd);
            if (FAST32) S.q0[slot] = make_float2(__double2float_rn(dsub(dmul(dsub(posx, ox), inv_res), 0.5)), __double2float_rn(dsub(dmul(dsub(posy, oy), inv_res), 0.5)));
        }
        __syncthreads();
        // ---- phase B -------------------------------------------------------------------------------------------------
        const int n_rays = nv * nb;
        const int nrr = NR ? NR : nr;
        for (int r = threadIdx.x; r < n_rays; r += RU_TILE) {
            // r / nb by multiplication with ceil(2^32 / nb) (exact for r < 2^32 / nb); ceil(2^32 / 1) does not fit 32 bits,
            // so one scored beam (a scan with 1..beam_stride filtered readings inside the FOV) is its own case
            const int v = nb == 1 ? r : (int)__umulhi((unsigned)r, div_magic);
            const int b = r - v * nb;
            // angle_key = (int)round(yaw_deg + off_deg) (MC:352-355): decided in fp32 unless the sum lies within 1e-3 of a
            // rounding boundary (the fp32 sum is within 4e-5 of the f64 one), then in f64 from the fast yaw unless within 1e-9
            // (the fast yaw is within 1e-12 of the tf round trip), then - once per ~1e9 rays - through the round trip itself
            int k;
            {
                const float a32 = __fadd_rn(S.yawf[v], S.offf[b]);
                const float tk = __fadd_rn(a32, RU_MAGICF);
                const float ek = __fadd_rn(a32, -__fadd_rn(tk, -RU_MAGICF));
                if (__builtin_expect(fabsf(ek) < 0.499f, 1)) {
                    const int kk = (__float_as_int(tk) - RU_MAGICF_BITS) - P.key_min;
                    k = (kk < 0 || kk >= P.n_keys) ? -1 : kk;
                } else {
                    const double off = S.beams[b].off_deg;
                    double a = dadd(S.yawd[v], off);
                    const double fr = fabs(dsub(a, trunc(a)));
                    if (!(S.vlist[v] & 0x10000) && !(fabs(fr - 0.5) > 1e-9)) {
                        const float th = part[tile * ppb + (S.vlist[v] & 0xffff)].z;
                        a = dadd(ddiv(dmul(ref_yaw(th), 180.0), 3.14159265358979323846), off);
                    }
                    const int kk = trunc_x86(round(a)) - P.key_min;
                    k = (kk < 0 || kk >= P.n_keys) ? -1 : kk;
                }
            }
            if (FAST32) {
                // All probes of the ray at once, branch-free, in fp32: q' = cell coordinate - 0.5, so that round-to-nearest
                // (magic add) gives floor(q) and e = q' - RN(q') = frac(q) - 0.5. A probe whose q lies within tol of a cell
                // edge in either axis (|e| >= 0.5 - tol) is marked undecided. Codes from the bordered table: 1 = occupied
                // (MC:377), 2 = left the grid (MC:376); the first non-zero code ends the march. Only if an undecided probe
                // comes at or before it does the f64 march decide (a fraction ~1e-3 of the rays).
                // The x and y halves travel as one packed fp32 pair (sm_100 FFMA2 / FADD2: two IEEE operations per issue slot,
                // same values as the scalar form; this kernel is issue bound).
                const float2 q0 = S.q0[v];
                const float2 dq = (k >= 0) ? S.dq[k] : make_float2(0.f, 0.f);
                const f32x2 q02 = pk2(q0.x, q0.y), dq2 = pk2(dq.x, dq.y);
                const f32x2 magic2 = pk2(RU_MAGICF, RU_MAGICF), nmagic2 = pk2(-RU_MAGICF, -RU_MAGICF), mone2 = pk2(-1.0f, -1.0f);
                uint32_t codes = 0;
#pragma unroll
                for (int s = 0; s < (NR ? NR : 16); s++) {                              // MC:372
                    if (!NR && s >= nr) break;
                    const float rf = S.radii_f[s];
                    const f32x2 q2 = fma2(pk2(rf, rf), dq2, q02);
                    const f32x2 t2 = add2(q2, magic2);
                    const f32x2 e2 = fma2(mone2, add2(t2, nmagic2), q2);                 // q - RN(q): one rounding, as the subtraction
                    float tx, ty, ex, ey;
                    upk2(t2, tx, ty); upk2(e2, ex, ey);
                    const bool decided = (fabsf(ex) < lim32) & (fabsf(ey) < lim32);      // NaN fails
                    // a valid particle is inside the map and its rays end inside the border, so the index is in range;
                    // an undecided probe's code is never used (the f64 march takes over if it matters)
                    uint32_t code = 3u;                                                  // 3 = undecided
                    if (MAP_SMEM) {
                        // a decided probe of a valid particle (inside the map, ray inside the border) is always in range
                        if (decided)
                            asm("{\n.reg .u32 a;\nmad.lo.u32 a, %1, %2, %3;\nadd.u32 a, a, %4;\nld.shared.u8 %0, [a];\n}"
                                : "=r"(code)
                                : "r"(__float_as_uint(ty)), "r"(wp_u), "r"(__float_as_uint(tx)), "r"(pad_saddr));
                    } else {
                        const uint32_t idx = __float_as_uint(ty) * wp_u + __float_as_uint(tx) + pad_fold;
                        if (decided) code = (uint32_t)occ_pad[min(idx, pad_last)];
                    }
                    codes += code << (2 * s);
                }
                const int first = codes ? (__ffs((int)codes) - 1) >> 1 : nrr;            // probe index of the first event
                double term;
                if (__builtin_expect(first < nrr && ((codes >> (2 * first)) & 3u) == 3u, 0)) {
                    const double px = S.posx[v], py = S.posy[v];
                    const double2 dir = (k >= 0) ? S.lut[k] : make_double2(0.0, 0.0);
                    double expected = P.max_range;                                       // MC:389
                    for (int s = 0; s < nrr; s++) {                                      // exact march
                        const double rr = S.radii[s];
                        const int c = probe_fast<ZERO_ORIGIN>(occ, dadd(px, dmul(rr, dir.x)), dadd(py, dmul(rr, dir.y)), ox, oy, res, inv_res, W, H);
                        if (c < 0) break;                                                // MC:376
                        if (c) { expected = rr; break; }                                 // MC:377-381
                    }
                    term = dmul(P.w_hit, ref_gauss(P, fabs(dsub(S.beams[b].obs, expected))));
                } else {
                    const bool hit = first < nrr && ((codes >> (2 * first)) & 3u) == 1u;
                    term = S.gterm[b * (nrr + 1) + (hit ? first : nrr)];
                }
                S.terms[v * stride + b] = term;                                          // MC:665
            } else {
                const RefBeam bm = S.beams[b];
                const double px = S.posx[v], py = S.posy[v];
                const double2 dir = (k >= 0) ? S.lut[k] : make_double2(0.0, 0.0);
                double expected = P.max_range;                                           // MC:389
                for (int s = 0; s < nr; s++) {                                           // MC:372
                    const double rr = S.radii[s];
                    const int c = probe_fast<ZERO_ORIGIN>(occ, dadd(px, dmul(rr, dir.x)), dadd(py, dmul(rr, dir.y)), ox, oy, res, inv_res, W, H);
                    if (c < 0) break;                                                    // MC:376
                    if (c) { expected = rr; break; }                                     // MC:377-381
                }
                const double diff = fabs(dsub(bm.obs, expected));                        // MC:662
                S.terms[v * stride + b] = dmul(P.w_hit, ref_gauss(P, diff));             // MC:665
            }
        }
        __syncthreads();
        // ---- phase C -------------------------------------------------------------------------------------------------
        if ((int)threadIdx.x < nv) {
            double prob = 0.0;
            const double* t = S.terms + threadIdx.x * stride;
            for (int b = 0; b < nb; b++) {
                prob = dadd(prob, t[b]);                                                 // MC:665
                prob = dadd(prob, S.beams[b].rand_term);                                 // MC:669
            }
            S.wout[S.vlist[threadIdx.x] & 0xffff] = __double2float_rn(prob);             // MC:673
        }
        __syncthreads();
        if (mine) {                                                                      // the particle's own thread writes it back whole
            p.w = valid ? S.wout[threadIdx.x] : 0.f;
            part[j] = p;
            w_dense[j] = p.w;
        }
        __syncthreads();
    }
}

// ---- sequential f64 accumulations (MC:675 and MC:496-505) --------------------------------------------------
// The reference sums fp32 weights into an f64 total one by one, and builds the CDF the same way. fp addition is
// not associative, so bit-identical results need the same roundings. The fast path is the parallel exact scan
// (exact_scan.cuh); these single-chain kernels are its always-correct fallback (they run only when *run_if != 0, or
// when run_if is null) and the cross-check the tests compare it with.
constexpr int SEQ_TILE = 1024;

__global__ void __launch_bounds__(256) k_ref_seq_total(const float* __restrict__ w, int64_t n, double* __restrict__ total_out,
                                                       const int* __restrict__ run_if) {
    if (run_if && *run_if == 0) return;
    __shared__ float tile[2][SEQ_TILE];
    double acc = 0.0;
    int64_t n_tiles = (n + SEQ_TILE - 1) / SEQ_TILE;
    for (int i = threadIdx.x; i < SEQ_TILE; i += blockDim.x) { int64_t g = i; tile[0][i] = g < n ? w[g] : 0.f; }
    __syncthreads();
    for (int64_t t = 0; t < n_tiles; t++) {
        int cur = t & 1;
        if (threadIdx.x == 0) {
            int64_t cnt = min((int64_t)SEQ_TILE, n - t * SEQ_TILE);
            for (int i = 0; i < cnt; i++) acc = dadd(acc, (double)tile[cur][i]);
        } else if (t + 1 < n_tiles) {
            int64_t base = (t + 1) * SEQ_TILE;
            for (int i = threadIdx.x - 1; i < SEQ_TILE; i += blockDim.x - 1) { int64_t g = base + i; tile[cur ^ 1][i] = g < n ? w[g] : 0.f; }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) *total_out = acc;
}

// cdf[i] = cdf[i-1] + (double)w_i over already-normalised weights (MC:498,504).
__global__ void __launch_bounds__(256) k_ref_seq_cdf(const float* __restrict__ wn_in, int64_t n, double* __restrict__ cdf,
                                                     const int* __restrict__ run_if) {
    pdl_enter();
    if (run_if && *run_if == 0) return;
    __shared__ float wn[2][SEQ_TILE];
    __shared__ double out[2][SEQ_TILE];
    double acc = 0.0;
    int64_t n_tiles = (n + SEQ_TILE - 1) / SEQ_TILE;
    auto stage = [&](int buf, int64_t t, int first, int stride) {
        int64_t base = t * SEQ_TILE;
        for (int i = first; i < SEQ_TILE; i += stride) { int64_t g = base + i; if (g < n) wn[buf][i] = wn_in[g]; }
    };
    auto flush = [&](int buf, int64_t t, int first, int stride) {
        int64_t base = t * SEQ_TILE;
        for (int i = first; i < SEQ_TILE; i += stride) { int64_t g = base + i; if (g < n) cdf[g] = out[buf][i]; }
    };
    stage(0, 0, threadIdx.x, blockDim.x);
    __syncthreads();
    for (int64_t t = 0; t < n_tiles; t++) {
        int cur = t & 1;
        if (threadIdx.x == 0) {
            int64_t cnt = min((int64_t)SEQ_TILE, n - t * SEQ_TILE);
            for (int i = 0; i < cnt; i++) { acc = dadd(acc, (double)wn[cur][i]); out[cur][i] = acc; }
        } else {
            if (t > 0) flush(cur ^ 1, t - 1, threadIdx.x - 1, blockDim.x - 1);
            if (t + 1 < n_tiles) stage(cur ^ 1, t + 1, threadIdx.x - 1, blockDim.x - 1);
        }
        __syncthreads();
    }
    flush((n_tiles - 1) & 1, n_tiles - 1, threadIdx.x, blockDim.x);
}

// ---- estimateWeightedPose (MC:782-800): the pieces k_pose_sums and (for small filters) k_ref_resample share --------------
// Sums of w, w x, w y, w sin(theta), w cos(theta) with w = weight / weight_sum: fp32 element math as the reference,
// f64 accumulation in a fixed order: thread -> warp tree -> the block's warps in order -> partials[block][4]; the last block
// to finish adds the per-block partials (lane-strided, then a warp tree) and leaves the four sums in out4. In a whole-tick call
// it also stores the tick's scalar results straight into the caller's pinned host block (zero-copy): the four pose sums,
// the adaptive-injection state and the resampling counters, the tick's sequence number last. (struct RefStepReport: mcl_engine.hpp)
struct PoseTail {
    double* partials;            // [grid][4]; null: no pose sums asked of this kernel
    unsigned* ticket;
    double* out4;
    RefStepReport* report;       // null: none
    const double* inj5;
    const int* counters4;
    unsigned long long seq;
};
__device__ __forceinline__ void pose_add(const float4& p, float weight_sum, double (&a)[4]) {
    float w = __fdiv_rn(p.w, weight_sum);
    a[0] += (double)__fmul_rn(w, p.x);
    a[1] += (double)__fmul_rn(w, p.y);
    // fp32 sin/cos within 2 ulp (the reference's are Eigen's fp32 psin/pcos, MC:790-791; the estimate is graded to 1e-5
    // and feeds nothing downstream, so the f64-evaluated correctly rounded form predict needs would be wasted here)
    float sn, cs;
    sincosf(p.z, &sn, &cs);
    a[2] += (double)__fmul_rn(w, sn);
    a[3] += (double)__fmul_rn(w, cs);
}
__device__ __forceinline__ void pose_report_aborted(const PoseTail& T) {          // optimistic tick that did not run: tell the watching host
    if (blockIdx.x == 0 && threadIdx.x == 0 && T.report) {
        T.report->aborted = 1;
        __threadfence_system();
        *(volatile unsigned long long*)&T.report->seq = T.seq;
    }
}
__device__ __forceinline__ void pose_block_finish(double (&a)[4], const PoseTail& T) {          // whole block, 256 threads
    __shared__ double ws[8][4];
    __shared__ bool last;
#pragma unroll
    for (int k = 0; k < 4; k++) a[k] = warp_sum(a[k]);
    if ((threadIdx.x & 31) == 0)
        for (int k = 0; k < 4; k++) ws[threadIdx.x >> 5][k] = a[k];
    __syncthreads();
    if (threadIdx.x < 4) { double s = 0; for (int w = 0; w < 8; w++) s += ws[w][threadIdx.x]; T.partials[(size_t)blockIdx.x * 4 + threadIdx.x] = s; }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) last = atomicAdd(T.ticket, 1u) == gridDim.x - 1;
    __syncthreads();
    if (last) {
        __threadfence();
        const int o = threadIdx.x >> 5, lane = threadIdx.x & 31;
        if (o < 4) {
            double s = 0;
            for (int b0 = lane; b0 < (int)gridDim.x; b0 += 32 * 8) {          // eight loads in flight, added in the same order as one by one
                double v[8];
#pragma unroll
                for (int k = 0; k < 8; k++) { const int b = b0 + 32 * k; v[k] = b < (int)gridDim.x ? __ldcg(T.partials + (size_t)b * 4 + o) : 0.0; }
#pragma unroll
                for (int k = 0; k < 8; k++) if (b0 + 32 * k < (int)gridDim.x) s += v[k];
            }
            s = warp_sum(s);
            if (lane == 0) { T.out4[o] = s; if (T.report) T.report->pose[o] = s; }
        } else if (T.report) {
            const int k = threadIdx.x - 128;
            if (k < 5) T.report->inj[k] = __ldcg(T.inj5 + k);
            else if (k < 9) T.report->counters[k - 5] = __ldcg(T.counters4 + (k - 5));
        }
        if (threadIdx.x == 0) *T.ticket = 0;
        if (T.report) {                             // the report is complete: the host may be spinning on its sequence number
            __syncthreads();
            if (threadIdx.x == 0) { T.report->aborted = 0; __threadfence_system(); *(volatile unsigned long long*)&T.report->seq = T.seq; }
        }
    }
}

// ---- resample (MC:508-555) ---------------------------------------------------------------------------------
struct RefResampleParams {
    double p_inject;       // MC:492
    int max_inject;        // MC:474/479
    int jitter_state;      // 1 = lost (x,y,theta jitter), 0 = confident (x,y only)
    double jit_xy_a, jit_xy_w;     // uniform(a,b): u*(b-a)+a with w=b-a  (libstdc++ formula)
    double jit_th_a, jit_th_w;
    float new_weight;      // (float)(1.0/N) (MC:524,551)
    // sampleParticles constants (MC:396-403, 431-432, 442-443)
    double cell_meters, half_cell, init_a, init_w, yaw_a, yaw_w, init_shift;
    uint32_t inj_rows, inj_cols;   // coarse-cell counts the injected particle's row / col draws are reduced to (MC:423-424)
};

// Production draws: the u_r / u_jitter streams come from Philox4x32-10, counter = (2i | 2i+1, stream 0x30, step):
//   u_r[i] = c53(A.0, A.1), u_jitter[i*per + 0] = c53(A.2, A.3), [+1] = c53(B.0, B.1), [+2] = c53(B.2, B.3),  A = philox(2i), B = philox(2i+1)
// Generated inside the kernels that consume them (GEN = true); k_fill_resample_draws materialises the same streams when a
// checker asks for them (mcl_debug_download_resample_draws).
struct RefDrawGen { uint32_t step, k0, k1; };
__device__ __forceinline__ void ref_philox_draws(uint64_t counter, const RefDrawGen& G, uint32_t (&o)[4]) {
    Philox::gen((uint32_t)counter, (uint32_t)(counter >> 32), 0x30u, G.step, G.k0, G.k1, o);
}
// The named draws of an injected particle (sampleParticles(1), MC:434-446): stream 0x31 of the same generator.
__device__ __forceinline__ void ref_philox_inject_draws(uint64_t counter, const RefDrawGen& G, uint32_t (&o)[4]) {
    Philox::gen((uint32_t)counter, (uint32_t)(counter >> 32), 0x31u, G.step, G.k0, G.k1, o);
}

// flags[i] = (u_r[i] < p_inject); block_counts[b] = number of flags in block b.
// Adaptive injection (MC:469-492) on the device, for mcl_step: the same IEEE operations in the same order as the host
// form in Engine::ref_resample, so both give the same bits. inj = {weight_slow, weight_fast, p_inject, cdf_is_monotone}.
__global__ void k_ref_ema(const double* __restrict__ total, double n, double a_slow, double a_fast, double* __restrict__ inj,
                          int* __restrict__ counters /* [0] injected, [1] clamped, [2] flagged, [3]: cleared for the resampling that follows */) {
    pdl_enter();
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    counters[0] = 0; counters[1] = 0; counters[2] = 0; counters[3] = 0;
    const double t = *total;
    const double avg = ddiv(t, n);
    const double slow = dadd(inj[0], dmul(a_slow, dsub(avg, inj[0])));
    const double fast = dadd(inj[1], dmul(a_fast, dsub(avg, inj[1])));
    const double p = dsub(1.0, ddiv(fast, slow));
    inj[0] = slow; inj[1] = fast;
    inj[2] = (0.0 < p) ? p : 0.0;                                   // std::max(0.0, p): NaN gives 0.0 (MC:492)
    inj[3] = (t > 0.0 && t < 1.0e300) ? 1.0 : 0.0;                  // finite positive total: the CDF is non-decreasing
    inj[4] = t;
}

// inj_dev (mcl_step): p_inject lives in device memory (inj_dev[2]); nothing to do when it is zero
// The last block to finish turns the per-block counts into exclusive offsets in place (the n-th flagged particle of the whole
// population takes the n-th injection draw, MC:508-523) and leaves their total in *total; `ticket` resets itself.
// n_counts = ceil(n / 256) counts (one per 256 slots: k_ref_resample's blocks); a block of this kernel produces RIC_SEGS of them,
// so that the whole population is one wave of blocks with one fence and one ticket each.
constexpr int RIC_SEGS = 4;
template <bool GEN>
__global__ void __launch_bounds__(256) k_ref_inject_count(const double* __restrict__ u_r, int64_t n, double p_inject,
                                                          int* __restrict__ block_counts, int n_counts, RefDrawGen G, const double* __restrict__ inj_dev,
                                                          int* __restrict__ total, unsigned* __restrict__ ticket, const int* __restrict__ abort) {
    pdl_enter();
    if (abort != nullptr && *abort != 0) return;
    if (inj_dev) { p_inject = inj_dev[2]; if (!(p_inject > 0.0)) return; }
    __shared__ bool last;
    __shared__ int warp_tot[8];
#pragma unroll
    for (int sgm = 0; sgm < RIC_SEGS; sgm++) {
        const int cb = (int)blockIdx.x * RIC_SEGS + sgm;                 // the count this segment produces
        if (cb >= n_counts) break;
        const int64_t i = (int64_t)cb * 256 + threadIdx.x;
        double r = 2.0;
        if (i < n) {
            if (GEN) { uint32_t a[4]; ref_philox_draws(2 * (uint64_t)i, G, a); r = canonical53(a[0], a[1]); }
            else r = u_r[i];
        }
        const int c = __syncthreads_count((r < p_inject) ? 1 : 0);
        if (threadIdx.x == 0) block_counts[cb] = c;
    }
    if (threadIdx.x == 0) {
        __threadfence();
        last = atomicAdd(ticket, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!last) return;
    __threadfence();
    const int nb = n_counts, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int carry = 0;
    for (int base = 0; base < nb; base += 256 * 16) {               // 16 consecutive counts per thread, all loads in flight at once
        int v[16], sum = 0;
        const int first = base + (int)threadIdx.x * 16;
#pragma unroll
        for (int k = 0; k < 16; k++) v[k] = first + k < nb ? __ldcg(block_counts + first + k) : 0;
#pragma unroll
        for (int k = 0; k < 16; k++) sum += v[k];
        int inc = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int up = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += up; }
        if (lane == 31) warp_tot[warp] = inc;
        __syncthreads();
        int woff = 0, chunk = 0;
#pragma unroll
        for (int k = 0; k < 8; k++) { if (k < warp) woff += warp_tot[k]; chunk += warp_tot[k]; }
        int run = carry + woff + inc - sum;
#pragma unroll
        for (int k = 0; k < 16; k++) { if (first + k < nb) block_counts[first + k] = run; run += v[k]; }
        carry += chunk;
        __syncthreads();
    }
    if (threadIdx.x == 0) { *total = carry; *ticket = 0u; }
}

// Up to one tile of weights (the reference's own few thousand particles), where a tick is bound by the number of dependent
// launches: the weight total (+ adaptive-injection state), the injection counts of k_ref_inject_count and the normalised CDF
// by ONE block in one launch (the exact-scan body of exact_scan_fused.cuh run twice; epochs `epoch` and `epoch + 1`). No guide
// table: below 4096 particles the CDF search is the plain lower_bound. Production draws only (mcl_step).
template <int ITEMS>          // weights per thread of the exact-scan passes: 4 up to 1024 particles, 8 up to 2048, else 16
__global__ void __launch_bounds__(xs::XS_THREADS, 3) k_ref_scans_one_tile(const float* __restrict__ w, int64_t n, unsigned epoch, xs::FusedWs ws,
                                                                         double* __restrict__ cdf_out, double* __restrict__ total_out, xs::FusedEma ema,
                                                                         int force_fallback, RefDrawGen G, int* __restrict__ block_counts,
                                                                         int* __restrict__ flagged_total) {
    pdl_enter();
    if (ws.abort != nullptr && *ws.abort != 0) return;
    __shared__ __align__(16) unsigned char sm_raw[xs::XSF_RAW_BYTES];
    __shared__ int seg_counts[xs::XSF_TILE / 256];
    xs::FusedGuide no_guide;
    no_guide.table = nullptr; no_guide.buckets = 0; no_guide.log2_buckets = 0; no_guide.force_fallback = force_fallback;
    xs::FusedEma no_ema;
    no_ema.inj = nullptr; no_ema.counters = nullptr; no_ema.n = 0; no_ema.a_slow = 0; no_ema.a_fast = 0;
    xs::xsf_run<false, ITEMS>(sm_raw, true, w, n, 1, epoch, ws, nullptr, nullptr, total_out, ema, no_guide);
    __threadfence();
    __syncthreads();
    // slots flagged for injection (u_r < p_inject, MC:518) per 256-slot segment, then their exclusive offsets: what
    // k_ref_inject_count leaves for k_ref_resample. Thread t draws for slots 16 t .. 16 t + 15, sixteen threads make a segment.
    const double p_inject = __ldcg(ema.inj + 2);
    if (block_counts != nullptr && p_inject > 0.0) {
        int c = 0;
#pragma unroll 2
        for (int j = 0; j < xs::XSF_ITEMS; j++) {
            const int64_t i = (int64_t)threadIdx.x * xs::XSF_ITEMS + j;
            if (i < n) { uint32_t a[4]; ref_philox_draws(2 * (uint64_t)i, G, a); c += canonical53(a[0], a[1]) < p_inject ? 1 : 0; }
        }
#pragma unroll
        for (int o = 1; o < 16; o <<= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
        if ((threadIdx.x & 15) == 0) seg_counts[threadIdx.x >> 4] = c;
        __syncthreads();
        if (threadIdx.x == 0) {
            const int n_counts = (int)((n + 255) / 256);
            int acc = 0;
            for (int k = 0; k < n_counts; k++) { const int v = seg_counts[k]; block_counts[k] = acc; acc += v; }
            *flagged_total = acc;
        }
    }
    __syncthreads();
    xs::xsf_run<true, ITEMS>(sm_raw, true, w, n, 1, epoch + 1u, ws, total_out, cdf_out, nullptr, no_ema, no_guide);
}

// (float)atan2(sin(t), cos(t)) (MC:550). For |t| < 3 pi the mathematical value is t, t - 2 pi or t + 2 pi; libm's composed
// result differs from it by a few 1e-16, so the float rounding agrees unless the value sits within 1e-14 of a float
// rounding boundary: only then (or near +-pi, or for larger |t|) is the libm chain evaluated.
__device__ __forceinline__ float ref_wrap_theta(double t) {
    const double PI = 3.14159265358979323846, TWO_PI_HI = 6.283185307179586, TWO_PI_LO = 2.4492935982947064e-16;
    if (fabs(t) < 9.42) {
        double z = t;
        if (t > PI) z = dsub(dsub(t, TWO_PI_HI), TWO_PI_LO);
        else if (t < -PI) z = dadd(dadd(t, TWO_PI_HI), TWO_PI_LO);
        const float f = __double2float_rn(z);
        if (__double2float_rn(z - 1e-14) == f && __double2float_rn(z + 1e-14) == f && fabs(fabs(t) - PI) > 1e-12) return f;
    }
    double sn, cs;
    sincos(t, &sn, &cs);
    return __double2float_rn(atan2(sn, cs));
}

// Guide table for the CDF search: guide[b] = std::lower_bound(cdf, b / buckets) for b = 0..buckets (buckets a power of two, so
// b / buckets and floor(r * buckets) are exact). lower_bound is monotone in its argument, hence for r in [b, b+1) / buckets
// the answer lies in [guide[b], guide[b+1]]: each particle's search shrinks from log2(N) dependent loads to one guide load
// plus log2(N / buckets) = 3 probes on neighbouring lines (the search is bound by L2 sector traffic: 32 B per 8-byte probe).
// Valid when the CDF is non-decreasing, i.e. the total weight is finite and positive; otherwise (NaN CDF, MC:492/530) the
// full-range search reproduces std::lower_bound's walk.
// Built by scattering, one coalesced pass over the CDF: element i is the answer for exactly the bucket edges b / buckets in
// (cdf[i-1], cdf[i]], i.e. b = floor(cdf[i-1] * buckets) + 1 .. floor(cdf[i] * buckets) (exact: power-of-two scaling);
// edges beyond cdf[n-1] get n. Elements owning more than four edges (a particle holding a large share of the weight) are
// written by their whole warp.
__device__ __forceinline__ int ref_guide_floor(double c, double B, int buckets) {
    const double x = c * B;
    return !(x >= 0.0) ? -1 : (x >= B ? buckets : (int)x);           // NaN owns nothing
}
__global__ void __launch_bounds__(256) k_ref_guide(const double* __restrict__ cdf, int64_t n, int buckets, int* __restrict__ guide) {
    pdl_enter();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const double B = (double)buckets;
    int lo_b = 0, hi_b = -1, tail_lo = 0, tail_hi = -1;               // inclusive bucket-edge ranges
    if (i < n) {
        hi_b = ref_guide_floor(cdf[i], B, buckets);
        lo_b = i == 0 ? 0 : ref_guide_floor(cdf[i - 1], B, buckets) + 1;
        if (i == n - 1) { tail_lo = hi_b + 1; tail_hi = buckets; }
    }
    const int cnt = hi_b - lo_b + 1;
    if (cnt > 0 && cnt <= 4)
        for (int b = lo_b; b <= hi_b; b++) guide[b] = (int)i;
    unsigned wide = __ballot_sync(0xffffffffu, cnt > 4);
    while (wide) {
        const int src = __ffs((int)wide) - 1;
        wide &= wide - 1;
        const int L = __shfl_sync(0xffffffffu, lo_b, src), H = __shfl_sync(0xffffffffu, hi_b, src);
        const int v = __shfl_sync(0xffffffffu, (int)i, src);
        for (int b = L + lane; b <= H; b += 32) guide[b] = v;
    }
    const unsigned has_tail = __ballot_sync(0xffffffffu, tail_hi >= tail_lo);
    if (has_tail) {
        const int src = __ffs((int)has_tail) - 1;
        const int L = __shfl_sync(0xffffffffu, tail_lo, src), H = __shfl_sync(0xffffffffu, tail_hi, src);
        for (int b = L + lane; b <= H; b += 32) guide[b] = (int)n;
    }
}

template <bool GEN>
__global__ void __launch_bounds__(256) k_ref_resample(const float4* __restrict__ src, float4* __restrict__ dst, int64_t n,
                                                      const double* __restrict__ cdf, const double* __restrict__ u_r,
                                                      const double* __restrict__ u_jit,
                                                      const double* __restrict__ inj_u_yaw, const int* __restrict__ inj_row,
                                                      const int* __restrict__ inj_col, const double* __restrict__ inj_u_dx,
                                                      const double* __restrict__ inj_u_dy,
                                                      const int* __restrict__ block_flag_offsets,   // null when p_inject == 0
                                                      RefResampleParams R, int* __restrict__ ancestors,
                                                      int* __restrict__ counters /* [0]=injected, [1]=clamped */, RefDrawGen G,
                                                      const int* __restrict__ guide /* null: full-range search */, int buckets,
                                                      const double* __restrict__ inj_dev /* mcl_step: {.., p_inject, cdf_is_monotone} on the device */,
                                                      const int* __restrict__ abort /* optimistic tick: see RefParams */,
                                                      PoseTail T /* small filters inside mcl_step: the pose sums of the new particles too */,
                                                      float pose_weight_sum,
                                                      int cdf_in_smem /* a whole small CDF (n doubles of dynamic shared memory) is staged first:
                                                                         the search's dependent loads then cost a shared-memory round trip each */) {
    pdl_enter();
    extern __shared__ __align__(16) unsigned char rs_smem[];
    if (abort != nullptr && *abort != 0) { if (T.partials != nullptr) pose_report_aborted(T); return; }
    __shared__ int warp_counts[8];
    const double* cs = cdf;                      // where the search reads the CDF (a generic pointer: global or shared)
    if (cdf_in_smem) {
        double* sc = reinterpret_cast<double*>(rs_smem);
        for (int64_t k = threadIdx.x; k < n; k += blockDim.x) sc[k] = cdf[k];
        __syncthreads();
        cs = sc;
    }
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool live = i < n;
    uint32_t A[4] = {0, 0, 0, 0};
    double r = 2.0;
    if (live) {
        if (GEN) { ref_philox_draws(2 * (uint64_t)i, G, A); r = canonical53(A[0], A[1]); }
        else r = u_r[i];
    }
    const double p_inject = inj_dev ? inj_dev[2] : R.p_inject;
    if (inj_dev) {
        if (!(p_inject > 0.0)) block_flag_offsets = nullptr;       // block-uniform: no slot can be injected
        if (inj_dev[3] == 0.0) guide = nullptr;                    // NaN CDF: std::lower_bound's full-range walk
    }
    // rank of this slot among the slots whose draw fell below p_inject (sequential injection counter, MC:518,525)
    int flag = (block_flag_offsets != nullptr && r < p_inject) ? 1 : 0;
    int rank = 0;
    if (block_flag_offsets != nullptr) {
        unsigned ballot = __ballot_sync(0xffffffffu, flag);
        int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        if (lane == 0) warp_counts[warp] = __popc(ballot);
        __syncthreads();
        int before = block_flag_offsets[blockIdx.x];
        for (int w = 0; w < warp; w++) before += warp_counts[w];
        rank = before + __popc(ballot & ((1u << lane) - 1u));
    }
    if (!live && T.partials == nullptr) return;
    float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
    if (live) {
        if (flag && rank < R.max_inject) {
            // sampleParticles(1) with named draws (MC:434-446). Production draws (GEN): injection slot `rank` takes Philox
            // counters 2*rank, 2*rank+1 of stream 0x31: u_yaw = c53(A.0,A.1), u_dx = c53(A.2,A.3), u_dy = c53(B.0,B.1),
            // row = B.2 mod rows, col = B.3 mod cols.
            double u_yaw, u_dx, u_dy;
            int row, col;
            if (GEN) {
                uint32_t a[4], b[4];
                ref_philox_inject_draws(2 * (uint64_t)rank, G, a); ref_philox_inject_draws(2 * (uint64_t)rank + 1, G, b);
                u_yaw = canonical53(a[0], a[1]); u_dx = canonical53(a[2], a[3]); u_dy = canonical53(b[0], b[1]);
                row = (int)(b[2] % R.inj_rows); col = (int)(b[3] % R.inj_cols);
            } else {
                u_yaw = inj_u_yaw[rank]; u_dx = inj_u_dx[rank]; u_dy = inj_u_dy[rank];
                row = inj_row[rank]; col = inj_col[rank];
            }
            double orientation = dadd(dmul(u_yaw, R.yaw_w), R.yaw_a);
            double x_move = dadd(dmul(u_dx, R.init_w), R.init_a);
            double y_move = dadd(dmul(u_dy, R.init_w), R.init_a);
            double base_x = dadd(dmul((double)col, R.cell_meters), R.half_cell);
            double base_y = dadd(dmul((double)row, R.cell_meters), R.half_cell);
            o.x = __double2float_rn(dadd(dadd(base_x, x_move), R.init_shift));
            o.y = __double2float_rn(dadd(dadd(base_y, y_move), R.init_shift));
            o.z = __double2float_rn(orientation);
            o.w = R.new_weight;
            ancestors[i] = -1;
            atomicAdd(&counters[0], 1);
        } else {
            // std::lower_bound(cdf, r): first idx with !(cdf[idx] < r) (MC:530); NaN entries compare false.
            int64_t lo = 0, len = n;
            if (guide && r >= 0.0 && r < 1.0) {                           // (injected draws outside [0, 1): full-range search)
                const int b = (int)(r * (double)buckets);                // exact: power-of-two scaling, r in [0, 1)
                lo = guide[b]; len = (int64_t)guide[b + 1] - lo;         // answer in [guide[b], guide[b+1]]
            }
            while (len > 0) {
                int64_t half = len >> 1;
                if (cs[lo + half] < r) { lo += half + 1; len -= half + 1; } else { len = half; }
            }
            if (lo >= n) { lo = n - 1; atomicAdd(&counters[1], 1); }
            float4 a = src[lo];
            int64_t inj_before = rank < R.max_inject ? rank : R.max_inject;
            // jitter draws are consumed in slot order by non-injected slots only: this slot takes stream entry i - inj_before
            double ux, uy, ut = 0.0;
            if (GEN) {
                const uint64_t j = (uint64_t)(i - inj_before);
                uint32_t B[4];
                if (inj_before != 0) ref_philox_draws(2 * j, G, A);
                ref_philox_draws(2 * j + 1, G, B);
                ux = canonical53(A[2], A[3]); uy = canonical53(B[0], B[1]); ut = canonical53(B[2], B[3]);
            } else {
                const int64_t jpos = (i - inj_before) * (R.jitter_state ? 3 : 2);
                ux = u_jit[jpos]; uy = u_jit[jpos + 1];
                if (R.jitter_state) ut = u_jit[jpos + 2];
            }
            double jx = dadd(dmul(ux, R.jit_xy_w), R.jit_xy_a);
            double jy = dadd(dmul(uy, R.jit_xy_w), R.jit_xy_a);
            double jt = (double)a.z;
            if (R.jitter_state) jt = dadd(jt, dadd(dmul(ut, R.jit_th_w), R.jit_th_a));
            o.x = __double2float_rn(dadd((double)a.x, jx));          // MC:548
            o.y = __double2float_rn(dadd((double)a.y, jy));          // MC:549
            o.z = ref_wrap_theta(jt);                                // MC:550
            o.w = R.new_weight;                                      // MC:551
            ancestors[i] = (int)lo;
        }
        dst[i] = o;
    }   // live
    if (T.partials != nullptr) {                    // same thread -> particle mapping and summation order as k_pose_sums: the same bits
        double a[4] = {0, 0, 0, 0};
        if (live) pose_add(o, pose_weight_sum, a);
        pose_block_finish(a, T);
    }
}

// ---- the production draw streams materialised (for checkers) ---------------------------------------------------------------
__global__ void __launch_bounds__(256) k_fill_resample_draws(double* __restrict__ u_r, double* __restrict__ u_jit, int64_t n, int per,
                                                             uint32_t step, uint32_t k0, uint32_t k1) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t a[4], b[4];
    uint64_t c = 2 * (uint64_t)i;
    Philox::gen((uint32_t)c, (uint32_t)(c >> 32), 0x30u, step, k0, k1, a);
    Philox::gen((uint32_t)(c + 1), (uint32_t)((c + 1) >> 32), 0x30u, step, k0, k1, b);
    u_r[i] = canonical53(a[0], a[1]);
    u_jit[i * per] = canonical53(a[2], a[3]);
    u_jit[i * per + 1] = canonical53(b[0], b[1]);
    if (per == 3) u_jit[i * per + 2] = canonical53(b[2], b[3]);
}

// ---- sampleParticles(N) (MC:415-450) with named draws ------------------------------------------------------
__global__ void __launch_bounds__(256) k_ref_init(float4* __restrict__ part, int64_t n, const double* __restrict__ u_yaw,
                                                  const int* __restrict__ row, const int* __restrict__ col,
                                                  const double* __restrict__ u_dx, const double* __restrict__ u_dy,
                                                  RefResampleParams R) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double orientation = dadd(dmul(u_yaw[i], R.yaw_w), R.yaw_a);
    double x_move = dadd(dmul(u_dx[i], R.init_w), R.init_a);
    double y_move = dadd(dmul(u_dy[i], R.init_w), R.init_a);
    double base_x = dadd(dmul((double)col[i], R.cell_meters), R.half_cell);
    double base_y = dadd(dmul((double)row[i], R.cell_meters), R.half_cell);
    float4 o;
    o.x = __double2float_rn(dadd(dadd(base_x, x_move), R.init_shift));
    o.y = __double2float_rn(dadd(dadd(base_y, y_move), R.init_shift));
    o.z = __double2float_rn(orientation);
    o.w = 1.0f;
    part[i] = o;
}

// ---- updateParticlePos (MC:740-755): fp32 element math -----------------------------------------------------
__global__ void __launch_bounds__(256) k_ref_predict(float4* __restrict__ part, int64_t n, float rot1, float trans, float dtheta, int trig) {
    pdl_enter();
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 p = part[i];
    float h = __fadd_rn(p.z, rot1);
    p.x = __fadd_rn(p.x, __fmul_rn(trans, ref_cosf(h, trig)));      // MC:747
    p.y = __fadd_rn(p.y, __fmul_rn(trans, ref_sinf(h, trig)));      // MC:748
    p.z = __fadd_rn(p.z, dtheta);
    part[i] = p;
}

// the kernels' float trig on an array (mcl_debug_trigf): what a checker compares with the host libm
__global__ void __launch_bounds__(256) k_debug_trigf(const float* __restrict__ x, int64_t n, int trig, float* __restrict__ s_out, float* __restrict__ c_out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float v = x[i];
    s_out[i] = ref_sinf(v, trig);
    c_out[i] = ref_cosf(v, trig);
}

// ---- estimateWeightedPose (MC:782-800) ---------------------------------------------------------------------
// pass 1: sum of weights (f64). pass 2: sums of (w/ws)*{x, y, sin, cos} with fp32 element math, f64 accumulation.
// Partials are written per block and reduced in a fixed order, so the result is deterministic.
__global__ void __launch_bounds__(256) k_pose_wsum(const float4* __restrict__ part, int64_t n, double* __restrict__ partials) {
    __shared__ double ws[8];
    double acc = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) acc += (double)part[i].w;
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) { double s = 0; for (int w = 0; w < 8; w++) s += ws[w]; partials[blockIdx.x] = s; }
}
// out[o] = sum over b of partials[b*stride + o]; one warp per output, fixed lane striding => deterministic.
__global__ void k_reduce_partials(const double* __restrict__ partials, int n_partials, int stride, int n_out, double* __restrict__ out) {
    const int o = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (o >= n_out) return;
    double s = 0;
    for (int b = lane; b < n_partials; b += 32) s += partials[(size_t)b * stride + o];
    s = warp_sum(s);
    if (lane == 0) out[o] = s;
}
// weight_sum by value (known on the host after update / resample / init, else from k_pose_wsum + k_reduce_partials). The last
// block to finish adds the per-block partials in block order (deterministic) and leaves the four sums in out4.
// report (mcl_step): the tick's scalar results, written by the last block straight into the caller's pinned host block
// (zero-copy; visible to the host once the stream has drained), so a tick ends without any device-to-host copy command:
// the four pose sums, the adaptive-injection state k_ref_ema left and the resampling counters.
// (struct RefStepReport: mcl_engine.hpp)
__global__ void __launch_bounds__(256) k_pose_sums(const float4* __restrict__ part, int64_t n, const double* __restrict__ wsum_dev, double wsum_host,
                                                   PoseTail T, const int* __restrict__ abort /* optimistic tick: see RefParams */) {
    pdl_enter();
    if (abort != nullptr && *abort != 0) { pose_report_aborted(T); return; }
    const float weight_sum = __double2float_rn(wsum_dev ? *wsum_dev : wsum_host);
    double a[4] = {0, 0, 0, 0};
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) pose_add(part[i], weight_sum, a);
    pose_block_finish(a, T);
}

}  // namespace mcl
