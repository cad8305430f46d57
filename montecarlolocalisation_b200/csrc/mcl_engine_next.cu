// mcl_engine_next.cu — host orchestration of the rows either side of the hot path: k-means confidence
// estimate (MC:886-949), pose-array download for rviz (MC:563-579). Host-only adapters live in host_models.hpp.
#include <algorithm>
#include <cmath>
#include <cstring>

#include "mcl_engine.hpp"
#include "engine_internal.hpp"
#include "kernels_next.cuh"

namespace mcl {

// isLocalizationLost_densitiy_cluster: K = 3, at most 20 iterations, radius 0.4 (MC:889-890, 933).
int Engine::kmeans_confidence(const int32_t* init_idx, const int32_t* reinit_idx, int n_reinit, double ratio_threshold, mcl_kmeans_result* out) {
    CK(cudaSetDevice(cfg.device));
    if (!out) return fail(MCL_ERR_ARG, "kmeans_confidence: null result");
    if (n == 0) return fail(MCL_ERR_ARG, "kmeans_confidence: no particles");
    if (cfg.mode == MCL_MODE_NS && shard_world != 1) return fail(MCL_ERR_STATE, "kmeans_confidence: single-shard filters only");
    if (n_reinit < 0 || (n_reinit > 0 && !reinit_idx)) return fail(MCL_ERR_ARG, "kmeans_confidence: bad re-initialisation draws");
    { int rc = ns_materialise_weights(); if (rc) return rc; }
    const int max_iters = 20;
    const bool exact = n <= MCL_KMEANS_EXACT_MAX;
    // the rand() % N draws (MC:813, 860): injected, or from the engine's Philox stream
    int32_t init[KM_K];
    std::vector<int32_t> reinit((size_t)max_iters * KM_K);
    for (int k = 0; k < KM_K; k++) {
        if (init_idx) { if (init_idx[k] < 0 || init_idx[k] >= n) return fail(MCL_ERR_ARG, "kmeans_confidence: initial index out of range"); init[k] = init_idx[k]; }
        else { uint32_t r[4]; philox_host(0x70, (uint64_t)k, r); init[k] = (int32_t)(r[0] % (uint32_t)n); }
    }
    for (size_t j = 0; j < reinit.size(); j++) {
        if ((int)j < n_reinit) { if (reinit_idx[j] < 0 || reinit_idx[j] >= n) return fail(MCL_ERR_ARG, "kmeans_confidence: re-initialisation index out of range"); reinit[j] = reinit_idx[j]; }
        else if (reinit_idx) reinit[j] = 0;      // injected draws exhausted: the oracle's convention
        else { uint32_t r[4]; philox_host(0x71, (uint64_t)j, r); reinit[j] = (int32_t)(r[0] % (uint32_t)n); }
    }
    const int blocks = (int)std::min<int64_t>(592, grid_for(n, KM_BLOCK));
    CK(d_assign.ensure((size_t)n)); CK(d_km.ensure(sizeof(KmState))); CK(d_km_reinit.ensure(reinit.size()));
    CK(d_partials.ensure(std::max<size_t>(4 * 1024, (size_t)blocks * 9)));
    KmState* st = (KmState*)d_km.p;
    // initial centres = the drawn particles; assignments start at 0 (fresh std::vector<int>, MC:805)
    KmState h;
    memset(&h, 0, sizeof(h));
    std::vector<float> three(4 * KM_K);
    for (int k = 0; k < KM_K; k++) CK(cudaMemcpyAsync(three.data() + 4 * k, part[cur].p + init[k], sizeof(float4), cudaMemcpyDeviceToHost, stream));
    CK(cudaMemsetAsync(d_assign.p, 0, (size_t)n * sizeof(int), stream));
    CK(cudaMemcpyAsync(d_km_reinit.p, reinit.data(), reinit.size() * sizeof(int32_t), cudaMemcpyHostToDevice, stream));
    CK(cudaStreamSynchronize(stream));
    for (int k = 0; k < KM_K; k++) { h.centers[2 * k] = three[4 * k]; h.centers[2 * k + 1] = three[4 * k + 1]; }
    CK(cudaMemcpyAsync(st, &h, sizeof(h), cudaMemcpyHostToDevice, stream));
    int passes = 0;
    for (int iter = 0; iter < max_iters; ++iter) {
        ++passes;
        CK(cudaMemsetAsync(&st->changed, 0, sizeof(int), stream));
        LAUNCH(K_KM_ASSIGN, k_km_assign, blocks, KM_BLOCK, 0, part[cur].p, n, st, d_assign.p, d_partials.p);
        CK(cudaGetLastError());
        int changed = 0;
        CK(cudaMemcpyAsync(&changed, &st->changed, sizeof(int), cudaMemcpyDeviceToHost, stream));
        CK(cudaStreamSynchronize(stream));
        if (!changed) break;                                                      // MC:842-845
        if (exact) LAUNCH(K_KM_UPDATE, k_km_update_seq, 1, 32 * KM_K, 0, part[cur].p, n, d_assign.p, st);
        LAUNCH(K_KM_UPDATE, k_km_finalize, 1, 32, 0, d_partials.p, blocks, part[cur].p, d_km_reinit.p, (int)reinit.size(), exact ? 1 : 0, st);
        CK(cudaGetLastError());
    }
    // cluster weights -> best cluster (first maximum, MC:910-917)
    if (exact) {
        LAUNCH(K_KM_STATS, k_km_cluster_weights_seq, 1, 32 * KM_K, 0, part[cur].p, n, d_assign.p, st);
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(&h, st, sizeof(h), cudaMemcpyDeviceToHost, stream));
        CK(cudaStreamSynchronize(stream));
    } else {
        LAUNCH(K_KM_STATS, k_km_cluster_weights, blocks, KM_BLOCK, 0, part[cur].p, n, d_assign.p, d_partials.p);
        CK(cudaGetLastError());
        std::vector<double> hp((size_t)blocks * 6);
        CK(cudaMemcpyAsync(hp.data(), d_partials.p, hp.size() * sizeof(double), cudaMemcpyDeviceToHost, stream));
        CK(cudaMemcpyAsync(&h, st, sizeof(h), cudaMemcpyDeviceToHost, stream));
        CK(cudaStreamSynchronize(stream));
        for (int k = 0; k < KM_K; k++) {
            double w = 0, c = 0;
            for (int b = 0; b < blocks; b++) { w += hp[(size_t)b * 6 + k]; c += hp[(size_t)b * 6 + 3 + k]; }
            h.cluster_weight[k] = w; h.counts[k] = (long long)c;
        }
    }
    int best = 0;
    double max_w = h.cluster_weight[0];
    for (int k = 1; k < KM_K; k++) if (h.cluster_weight[k] > max_w) { max_w = h.cluster_weight[k]; best = k; }
    const float xc = h.centers[2 * best], yc = h.centers[2 * best + 1];
    const float radius = (float)cfg.kmeans_radius;
    LAUNCH(K_KM_STATS, k_km_best_stats, blocks, KM_BLOCK, 0, part[cur].p, n, d_assign.p, best, xc, yc, radius * radius, d_partials.p);
    CK(cudaGetLastError());
    std::vector<double> hp((size_t)blocks * 3);
    CK(cudaMemcpyAsync(hp.data(), d_partials.p, hp.size() * sizeof(double), cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
    double ss = 0, cs = 0, near = 0;
    for (int b = 0; b < blocks; b++) { ss += hp[(size_t)b * 3]; cs += hp[(size_t)b * 3 + 1]; near += hp[(size_t)b * 3 + 2]; }
    const double ratio = near / (double)n;                                        // MC:933
    memset(out, 0, sizeof(*out));
    out->ratio = ratio;
    if (ratio > ratio_threshold) { out->x_best = xc; out->y_best = yc; out->theta_best = std::atan2(ss, cs); }       // MC:934-937
    else { out->x_best = -1; out->y_best = -1; out->theta_best = -1; }                                               // MC:938-940
    for (int k = 0; k < KM_K; k++) { out->cluster_weight[k] = h.cluster_weight[k]; out->counts[k] = h.counts[k]; out->centers[2 * k] = h.centers[2 * k]; out->centers[2 * k + 1] = h.centers[2 * k + 1]; }
    out->best_cluster = best; out->passes = passes; out->reinit_used = h.reinit_used; out->exact = exact ? 1 : 0;
    return MCL_OK;
}

int Engine::download_assignments(int32_t* a) {
    CK(cudaSetDevice(cfg.device));
    if (!a || n == 0 || d_assign.n < (size_t)n) return fail(MCL_ERR_ARG, "download_assignments: run mcl_kmeans_confidence first");
    CK(cudaMemcpyAsync(a, d_assign.p, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
    return MCL_OK;
}

// publishParticles (MC:563-579) for every stride-th particle: (x, y, qz, qw) in f64, built on the device
int Engine::download_pose_array(int64_t first, int64_t stride, int64_t count, double* out) {
    CK(cudaSetDevice(cfg.device));
    if (!out || count <= 0 || stride <= 0 || first < 0 || first + (count - 1) * stride >= n) return fail(MCL_ERR_ARG, "download_pose_array: range outside the particle set");
    CK(d_posearr.ensure((size_t)count * 4));
    LAUNCH(K_POSE_ARRAY, k_pose_array, grid_for(count, 256), 256, 0, part[cur].p, first, stride, count, (double4*)d_posearr.p);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(out, d_posearr.p, (size_t)count * 4 * sizeof(double), cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
    return MCL_OK;
}

}  // namespace mcl
