// kernels_next.cuh — the rows either side of the hot path: the k-means confidence estimate the node
// actually publishes (MC:802-949) and the particle pose array for rviz (MC:563-579).
//
//   k_km_assign        assignment step (MC:821-840), fused with per-block f64 partial sums of the centre update
//   k_km_update_seq    centre update with the reference's SEQUENTIAL fp32 accumulation (MC:851-856), one warp per cluster:
//                      bit-exact with the CPU; used for particle counts the reference itself can run (<= MCL_KMEANS_EXACT_MAX)
//   k_km_finalize      new centres = sum / count (MC:857-863) incl. re-initialisation of emptied clusters
//   k_km_cluster_weights, k_km_best_stats   cluster weights (MC:903-907), circular mean of the best cluster's theta
//                      (MC:924-932) and the density count around its centre (MC:869-884)
//   k_pose_array       (x, y, qz, qw) of every stride-th particle: createQuaternionMsgFromYaw on the device
#pragma once
#include "mcl_device.cuh"

namespace mcl {

constexpr int KM_K = 3;
constexpr int KM_BLOCK = 256;

struct KmState {
    float centers[2 * KM_K];
    int changed;            // set by k_km_assign when any assignment changed
    int reinit_used;
    long long counts[KM_K];
    double cluster_weight[KM_K];
    double sin_sum, cos_sum;
    long long near_count;
    float seq_sum[2 * KM_K];   // sequential fp32 sums (exact mode)
};

// squared distance exactly as MC:828-830 (fp32, no contraction), ties to the lowest k (strict <, MC:831)
__device__ __forceinline__ int km_best(float x, float y, const float* c) {
    float min_dist = 3.402823466e+38f;
    int best = -1;
#pragma unroll
    for (int k = 0; k < KM_K; k++) {
        const float dx = __fadd_rn(x, -c[2 * k]), dy = __fadd_rn(y, -c[2 * k + 1]);
        const float dist = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
        if (dist < min_dist) { min_dist = dist; best = k; }
    }
    return best < 0 ? 0 : best;      // all-NaN distances index out of bounds in the reference; defined as cluster 0 here
}

// partials[b][0..2] = sum x, [3..5] = sum y, [6..8] = count (as double) of block b's particles per NEW cluster
__global__ void __launch_bounds__(KM_BLOCK) k_km_assign(const float4* __restrict__ part, int64_t n, KmState* __restrict__ st, int* __restrict__ assign,
                                                        double* __restrict__ partials) {
    __shared__ double sm[KM_BLOCK / 32][9];
    __shared__ float c[2 * KM_K];
    if (threadIdx.x < 2 * KM_K) c[threadIdx.x] = st->centers[threadIdx.x];
    __syncthreads();
    double a[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    bool changed = false;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float4 p = part[i];
        const int b = km_best(p.x, p.y, c);
        if (assign[i] != b) { assign[i] = b; changed = true; }
#pragma unroll
        for (int k = 0; k < KM_K; k++) {
            const bool m = b == k;
            a[k] += m ? (double)p.x : 0.0; a[3 + k] += m ? (double)p.y : 0.0; a[6 + k] += m ? 1.0 : 0.0;
        }
    }
    if (__any_sync(0xffffffffu, changed) && (threadIdx.x & 31) == 0) atomicOr(&st->changed, 1);
#pragma unroll
    for (int k = 0; k < 9; k++) a[k] = warp_sum(a[k]);
    if ((threadIdx.x & 31) == 0)
        for (int k = 0; k < 9; k++) sm[threadIdx.x >> 5][k] = a[k];
    __syncthreads();
    if (threadIdx.x < 9) { double s = 0; for (int w = 0; w < KM_BLOCK / 32; w++) s += sm[w][threadIdx.x]; partials[(size_t)blockIdx.x * 9 + threadIdx.x] = s; }
}

// Exact mode: warp k walks the particles in index order and accumulates cluster k's x and y in fp32, one add after the
// other, as `new_centers[c].first += particles(0, i)` does (MC:851-856). Loads are coalesced (32 particles per step) and
// handed round by shuffle; every lane carries the same running sums.
__global__ void __launch_bounds__(32 * KM_K) k_km_update_seq(const float4* __restrict__ part, int64_t n, const int* __restrict__ assign,
                                                             KmState* __restrict__ st) {
    const int k = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float sx = 0.f, sy = 0.f;
    for (int64_t base = 0; base < n; base += 32) {
        const int64_t i = base + lane;
        float x = 0.f, y = 0.f;
        int a = -1;
        if (i < n) { const float4 p = part[i]; x = p.x; y = p.y; a = assign[i]; }
#pragma unroll
        for (int j = 0; j < 32; j++) {
            const float xj = __shfl_sync(0xffffffffu, x, j), yj = __shfl_sync(0xffffffffu, y, j);
            const int aj = __shfl_sync(0xffffffffu, a, j);
            if (aj == k) { sx = __fadd_rn(sx, xj); sy = __fadd_rn(sy, yj); }
        }
    }
    if (lane == 0) { st->seq_sum[2 * k] = sx; st->seq_sum[2 * k + 1] = sy; }
}

// One thread: counts from the partials (fixed order), then centre = sum / count in fp32 (float /= int, MC:857-858) or a
// re-initialisation draw for an emptied cluster (MC:859-861). exact: sums from k_km_update_seq, else fl32 of the f64 sums.
__global__ void k_km_finalize(const double* __restrict__ partials, int n_blocks, const float4* __restrict__ part, const int* __restrict__ reinit_idx,
                              int n_reinit, int exact, KmState* __restrict__ st) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    double s[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int b = 0; b < n_blocks; b++)
        for (int k = 0; k < 9; k++) s[k] += partials[(size_t)b * 9 + k];
    int used = st->reinit_used;
    for (int k = 0; k < KM_K; k++) {
        const long long cnt = (long long)s[6 + k];
        st->counts[k] = cnt;
        if (cnt > 0) {
            const float fx = exact ? st->seq_sum[2 * k] : (float)s[k], fy = exact ? st->seq_sum[2 * k + 1] : (float)s[3 + k];
            st->centers[2 * k] = __fdiv_rn(fx, (float)cnt);
            st->centers[2 * k + 1] = __fdiv_rn(fy, (float)cnt);
        } else {
            const int idx = used < n_reinit ? reinit_idx[used] : 0;
            ++used;
            const float4 p = part[idx];
            st->centers[2 * k] = p.x; st->centers[2 * k + 1] = p.y;
        }
    }
    st->reinit_used = used;
}

// cluster_weights[c] += particles(3, i) (MC:903-907): f64 sums of the fp32 weights, fixed reduction order; counts too
__global__ void __launch_bounds__(KM_BLOCK) k_km_cluster_weights(const float4* __restrict__ part, int64_t n, const int* __restrict__ assign,
                                                                 double* __restrict__ partials) {
    __shared__ double sm[KM_BLOCK / 32][6];
    double a[6] = {0, 0, 0, 0, 0, 0};
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int b = assign[i];
        const double w = (double)part[i].w;
#pragma unroll
        for (int k = 0; k < KM_K; k++) { a[k] += b == k ? w : 0.0; a[3 + k] += b == k ? 1.0 : 0.0; }
    }
#pragma unroll
    for (int k = 0; k < 6; k++) a[k] = warp_sum(a[k]);
    if ((threadIdx.x & 31) == 0)
        for (int k = 0; k < 6; k++) sm[threadIdx.x >> 5][k] = a[k];
    __syncthreads();
    if (threadIdx.x < 6) { double s = 0; for (int w = 0; w < KM_BLOCK / 32; w++) s += sm[w][threadIdx.x]; partials[(size_t)blockIdx.x * 6 + threadIdx.x] = s; }
}
// Exact mode: the reference's left-to-right f64 accumulation of the weights per cluster (warp k, as k_km_update_seq)
__global__ void __launch_bounds__(32 * KM_K) k_km_cluster_weights_seq(const float4* __restrict__ part, int64_t n, const int* __restrict__ assign,
                                                                      KmState* __restrict__ st) {
    const int k = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double s = 0.0;
    long long cnt = 0;
    for (int64_t base = 0; base < n; base += 32) {
        const int64_t i = base + lane;
        float w = 0.f;
        int a = -1;
        if (i < n) { w = part[i].w; a = assign[i]; }
#pragma unroll
        for (int j = 0; j < 32; j++) {
            const float wj = __shfl_sync(0xffffffffu, w, j);
            const int aj = __shfl_sync(0xffffffffu, a, j);
            if (aj == k) { s = dadd(s, (double)wj); ++cnt; }
        }
    }
    if (lane == 0) { st->cluster_weight[k] = s; st->counts[k] = cnt; }
}
// sin/cos sums of the best cluster's theta in f64 (MC:924-931) and the count within `radius` of (xc, yc) (MC:869-884)
__global__ void __launch_bounds__(KM_BLOCK) k_km_best_stats(const float4* __restrict__ part, int64_t n, const int* __restrict__ assign, int best,
                                                            float xc, float yc, float radius_sq, double* __restrict__ partials) {
    __shared__ double sm[KM_BLOCK / 32][3];
    double a[3] = {0, 0, 0};
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float4 p = part[i];
        if (assign[i] == best) {
            double sn, cs;
            sincos((double)p.z, &sn, &cs);
            a[0] += sn; a[1] += cs;
        }
        const float dx = __fadd_rn(p.x, -xc), dy = __fadd_rn(p.y, -yc);
        const float d = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
        a[2] += d <= radius_sq ? 1.0 : 0.0;
    }
#pragma unroll
    for (int k = 0; k < 3; k++) a[k] = warp_sum(a[k]);
    if ((threadIdx.x & 31) == 0)
        for (int k = 0; k < 3; k++) sm[threadIdx.x >> 5][k] = a[k];
    __syncthreads();
    if (threadIdx.x < 3) { double s = 0; for (int w = 0; w < KM_BLOCK / 32; w++) s += sm[w][threadIdx.x]; partials[(size_t)blockIdx.x * 3 + threadIdx.x] = s; }
}

// out[j] = (x, y, sin(theta/2), cos(theta/2)) of particle first + j*stride: geometry_msgs/Pose position.xy and
// orientation.zw as publishParticles builds them (MC:570-574; tf::createQuaternionMsgFromYaw, x = y = 0)
__global__ void __launch_bounds__(256) k_pose_array(const float4* __restrict__ part, int64_t first, int64_t stride, int64_t count,
                                                    double4* __restrict__ out) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= count) return;
    const float4 p = part[first + j * stride];
    double sn, cs;
    sincos(dmul((double)p.z, 0.5), &sn, &cs);
    out[j] = make_double4((double)p.x, (double)p.y, sn, cs);
}

}  // namespace mcl
