// mcl_engine_ns.cu — host orchestration of MCL_MODE_NS (likelihood field, Philox motion noise, fixed-point systematic
// resampling, sharding across GPUs). Definitions: ns_core.cuh / DESIGN.md "NS".
#include <algorithm>
#include <cmath>
#include <cstring>

#include "mcl_engine.hpp"
#include "engine_internal.hpp"
#include "kernels_ns.cuh"
#include "nccl_dyn.hpp"
#include "ns_plan.hpp"

namespace mcl {

int Engine::ns_set_shard(int rank, int world, int64_t ng) {
    if (cfg.mode != MCL_MODE_NS) return fail(MCL_ERR_STATE, "set_shard: NS mode only (REF resampling needs the global f64 CDF on one GPU)");
    if (world < 1 || world > 8 || rank < 0 || rank >= world || ng <= 0) return fail(MCL_ERR_ARG, "set_shard: bad rank/world/count");
    // the Q32 resampling arithmetic keeps totals below 2^60 (N * 2^32), thresholds' remainders below 2^31 + 2^9 * 2^31 and
    // stores ancestors as int: all of that assumes fewer than 2^28 particles in the whole filter
    if (ng >= (1ll << 28)) return fail(MCL_ERR_ARG, "set_shard: n_global must be below 2^28 (Q32 totals stay below 2^60, ancestors are 32-bit)");
    shard_rank = rank; shard_world = world; n_global = ng;
    int64_t cnt;
    ns::shard_range(ng, world, rank, &shard_begin, &cnt, &per_rank);
    if (cnt <= 0) return fail(MCL_ERR_ARG, "set_shard: empty shard");
    int rc = ensure_particles(cnt);
    if (rc) return rc;
    n = cnt;
    return world > 1 ? ensure_mailbox() : MCL_OK;
}

#define NS_UPD_FOR_ALL(X) X(NS_FIELD_SMEM, 0) X(NS_FIELD_SMEM, 1) X(NS_FIELD_SMEM, 2) X(NS_FIELD_GLOBAL, 0) X(NS_FIELD_GLOBAL, 1) X(NS_FIELD_GLOBAL, 2) \
                          X(NS_FIELD_U8, 0) X(NS_FIELD_U8, 1) X(NS_FIELD_U8, 2)
int Engine::ns_preload_update(int kind, int pack) {
    cudaFuncAttributes fa;
#define X(K, P) if (kind == K && pack == P) CK(cudaFuncGetAttributes(&fa, k_ns_update<K, P>));
    NS_UPD_FOR_ALL(X)
#undef X
    return MCL_OK;
}

int Engine::ensure_mailbox() {
    if (d_mbox.p) return MCL_OK;
    CK(d_mbox.ensure(sizeof(NsMailbox)));
    CK(cudaMemset(d_mbox.p, 0, sizeof(NsMailbox)));
    // everything mcl_ns_step needs, allocated now: cudaMalloc waits for the device, and a shard whose stream already holds
    // an exchange kernel waiting for its peers must not be waited on by a peer living in the same process
    CK(d_totals.ensure(8)); CK(d_plan.ensure(sizeof(NsPlan))); CK(d_pose.ensure(8)); CK(d_bar.ensure(1)); CK(d_partials.ensure(5 * 2048));
    const int64_t max_tiles = (n_global + NS_RS_TILE - 1) / NS_RS_TILE;
    CK(d_bounds.ensure(((size_t)max_tiles + 2) * sizeof(NsTileHead)));
    // ... and every kernel of the step loaded now: with lazy module loading the first launch of a kernel synchronises
    // the context, which would wait for exchange kernels already spinning on this device
    cudaFuncAttributes fa;
#define PRELOAD(k) CK(cudaFuncGetAttributes(&fa, k))
    PRELOAD(k_ns_predict);
    for (int kk = 0; kk < 3; ++kk) for (int pp = 0; pp < 3; ++pp) { int rc = ns_preload_update(kk, pp); if (rc) return rc; }
    PRELOAD(k_ns_weights_sum); PRELOAD(k_ns_weights_scan); PRELOAD(k_ns_weights_scan1); PRELOAD(k_ns_plan); PRELOAD(k_ns_plan_xchg); PRELOAD(k_ns_xchg_max);
    PRELOAD(k_ns_pose_partials); PRELOAD(k_ns_pose_reduce); PRELOAD(k_ns_xchg_barrier);
    PRELOAD(k_ns_resample_bounds); PRELOAD(k_ns_resample);
#undef PRELOAD
    return MCL_OK;
}
bool Engine::peers_have_mailboxes() const {
    if (!d_mbox.p) return false;
    for (int r = 0; r < shard_world; ++r)
        if (r != shard_rank && !peer_ptr[3][r]) return false;
    return true;
}

// Exact capped squared distance transform -> log-likelihood field.
int Engine::ns_build_field() {
    const double max_dist = 2.0;                                   // metres beyond which the field is flat
    ns_tune_cost[0] = ns_tune_cost[1] = ns_tune_cost[2] = -1.0; ns_tune_launches = 0;       // a new field: measure afresh
    ns_R = std::min(255, std::max(1, (int)std::ceil(max_dist / (double)res_f)));
    const int cap = ns_R * ns_R;
    std::vector<float> table(cap + 1);
    const double sigma = cfg.ns_sigma_hit;
    for (int d2 = 0; d2 <= cap; ++d2) {
        const double d = (double)res_f * std::sqrt((double)d2);
        const double p = cfg.ns_z_hit * std::exp(-(d * d) / (2.0 * sigma * sigma)) / (sigma * std::sqrt(2.0 * M_PI)) + cfg.ns_z_rand / cfg.ns_max_range;
        table[d2] = (float)std::log(p);
    }
    lf_out = table[cap];
    const size_t cells = (size_t)map_w * map_h;
    // border: no beam endpoint of a particle inside the map can leave the stored field (beams >= ns_max_range are dropped)
    const double reach = (std::fabs(cfg.laser_offset) + cfg.ns_max_range) / (double)res_f;
    lf_pad = reach < 4096.0 ? (int)std::ceil(reach) + 2 : 0;        // 0: absurd range, keep the bounds-tested path only
    lf_wp = map_w + 2 * lf_pad; lf_hp = map_h + 2 * lf_pad;
    const size_t cells_p = (size_t)lf_wp * lf_hp;
    if (cells_p >= (1ull << 31)) return fail(MCL_ERR_ARG, "set_map: grid too large");
    lf_bytes_padded = (cells_p * sizeof(float) + 15) & ~(size_t)15;
    CK(d_lf_table.ensure(table.size())); CK(d_lf.ensure(lf_bytes_padded / sizeof(float))); CK(d_d2.ensure(cells)); CK(d_g.ensure(cells));
    // the sensor-model kernel addresses the field with 32-bit arithmetic when it lies inside one 4 GiB-aligned window
    // (else it falls back to 64-bit address arithmetic): if this allocation straddles a boundary, try for another one
    for (int attempt = 0; attempt < 3; ++attempt) {
        const uint64_t a = (uint64_t)(uintptr_t)d_lf.p, b = a + lf_bytes_padded - 1;
        if ((a >> 32) == (b >> 32) || lf_bytes_padded > (1ull << 32)) break;
        DevBuf<float> other;
        if (other.ensure(lf_bytes_padded / sizeof(float)) != cudaSuccess) { cudaGetLastError(); break; }
        std::swap(other.p, d_lf.p); std::swap(other.n, d_lf.n);
        other.release();
    }
    // The same field as one byte per cell: d2 takes only the values a^2 + b^2 <= R^2 (and the cap), 125 of them for R = 20,
    // so its rank fits a byte whenever there are <= 256. A quarter of the footprint: an 8193 x 8193 field is 66 MiB instead
    // of 264 MiB and stays L2-resident. k_ns_update maps code -> table[d2] through shared memory: identical values.
    std::vector<uint8_t> code_of_d2(cap + 1, 0);
    std::vector<float> codes;
    {
        std::vector<char> attainable(cap + 1, 0);
        attainable[cap] = 1;
        for (int a = 0; a <= ns_R; ++a)
            for (int b = a; a * a + b * b <= cap; ++b) attainable[a * a + b * b] = 1;
        int count = 0;
        for (int d2 = 0; d2 <= cap; ++d2) count += attainable[d2];
        ns_n_codes = count <= NS_MAX_CODES ? count : 0;
        if (ns_n_codes) {
            for (int d2 = 0; d2 <= cap; ++d2)
                if (attainable[d2]) { code_of_d2[d2] = (uint8_t)codes.size(); codes.push_back(table[d2]); }
        }
    }
    CK(cudaMemcpyAsync(d_lf_table.p, table.data(), table.size() * sizeof(float), cudaMemcpyHostToDevice, stream));
    LAUNCH(K_NS_EDT_COLS, k_ns_fill_f32, 148 * 8, 256, 0, d_lf.p, lf_bytes_padded / sizeof(float), lf_out);
    if (ns_n_codes) {
        const size_t bytes8 = (cells_p + 15) & ~(size_t)15;
        CK(d_codes.ensure(codes.size())); CK(d_code_of_d2.ensure(code_of_d2.size())); CK(d_lf8.ensure(bytes8));
        for (int attempt = 0; attempt < 3; ++attempt) {            // inside one 4 GiB window, like the fp32 field
            const uint64_t a = (uint64_t)(uintptr_t)d_lf8.p, b = a + bytes8 - 1;
            if ((a >> 32) == (b >> 32)) break;
            DevBuf<uint8_t> other;
            if (other.ensure(bytes8) != cudaSuccess) { cudaGetLastError(); break; }
            std::swap(other.p, d_lf8.p); std::swap(other.n, d_lf8.n);
            other.release();
        }
        CK(cudaMemcpyAsync(d_codes.p, codes.data(), codes.size() * sizeof(float), cudaMemcpyHostToDevice, stream));
        CK(cudaMemcpyAsync(d_code_of_d2.p, code_of_d2.data(), code_of_d2.size(), cudaMemcpyHostToDevice, stream));
        LAUNCH(K_NS_EDT_COLS, k_ns_fill_u8, 148 * 8, 256, 0, d_lf8.p, bytes8, code_of_d2[cap]);
    }
    LAUNCH(K_NS_EDT_COLS, k_ns_edt_cols, grid_for(map_w, 128), 128, 0, d_occ.p, map_w, map_h, ns_R, d_g.p);
    LAUNCH(K_NS_EDT_ROWS, k_ns_edt_rows, dim3(grid_for(map_w, 128), map_h), 128, 0, d_g.p, map_w, map_h, ns_R, d_lf_table.p, d_d2.p, d_lf.p, lf_pad,
           ns_n_codes ? d_code_of_d2.p : nullptr, ns_n_codes ? d_lf8.p : nullptr);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(stream));
    return MCL_OK;
}

int Engine::ns_download_field(float* lf, uint16_t* d2) {
    CK(cudaSetDevice(cfg.device));
    if (!map_ready || cfg.mode != MCL_MODE_NS) return fail(MCL_ERR_STATE, "download_field: NS mode with a map only");
    const size_t cells = (size_t)map_w * map_h;
    if (lf) CK(cudaMemcpy2DAsync(lf, (size_t)map_w * sizeof(float), d_lf.p + (size_t)lf_pad * lf_wp + lf_pad, (size_t)lf_wp * sizeof(float),
                                 (size_t)map_w * sizeof(float), map_h, cudaMemcpyDeviceToHost, stream));
    if (d2) CK(cudaMemcpyAsync(d2, d_d2.p, cells * sizeof(uint16_t), cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
    return MCL_OK;
}

int Engine::ns_init(int64_t count) {
    if (shard_world == 1) {
        int rc = ensure_particles(count);
        if (rc) return rc;
        n = count; n_global = count; shard_begin = 0; per_rank = count;
    } else if (count != n_global) {
        return fail(MCL_ERR_ARG, "init: pass the GLOBAL particle count given to mcl_ns_set_shard");
    }
    const double ext_x = (double)map_w * (double)res_f, ext_y = (double)map_h * (double)res_f;
    LAUNCH(K_NS_INIT, k_ns_init, grid_for(n, 256), 256, 0, part[cur].p, n, shard_begin, origin_x, origin_y, ext_x, ext_y,
           (uint32_t)cfg.seed, (uint32_t)(cfg.seed >> 32));
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(stream));
    have_weights = false; ns_have_ll = false;
    return MCL_OK;
}

// The clean odometry increment plus, per particle, N(0, variance) noise with the reference's variances (MC:706-710).
int Engine::ns_predict(const Motion& m, bool prep_step) {
    NsMotion k;
    k.rot1 = (float)m.rot_1; k.trans = (float)m.trans; k.rot2 = (float)m.rot_2;
    k.sd_rot1 = (float)std::sqrt(cfg.alpha[0] * std::fabs(m.rot_1) + cfg.alpha[1] * std::fabs(m.trans));
    k.sd_trans = (float)std::sqrt(cfg.alpha[2] * std::fabs(m.trans) + cfg.alpha[3] * (std::fabs(m.rot_1) + std::fabs(m.rot_2)));
    k.sd_rot2 = (float)std::sqrt(cfg.alpha[0] * std::fabs(m.rot_2) + cfg.alpha[1] * std::fabs(m.trans));
    NsStepPrep prep{nullptr, nullptr, 0};
    if (prep_step) {                     // mcl_ns_step: reset the step's accumulators here instead of by memset / copy commands
        bool two_pass; int nt, ng;
        ns_scan_shape(two_pass, nt, ng);
        prep.maxbits = d_maxbits.p;
        prep.zero = (unsigned long long*)(two_pass ? d_tile_sums.p + nt : d_tile_sums.p);
        prep.zero_words = two_pass ? ng : nt + 1;
    }
    LAUNCH_PDL(K_NS_PREDICT, k_ns_predict, grid_for(n, 256), 256, 0, part[cur].p, n, shard_begin, k, (uint32_t)step_counter,
           (uint32_t)cfg.seed, (uint32_t)(cfg.seed >> 32), prep);
    CK(cudaGetLastError());
    ns_maxbits_prepped = ns_scan_prepped = prep_step;
    have_weights = false; ns_have_ll = false;
    return MCL_OK;
}

// Scan -> beam endpoints in the robot frame.
void Engine::ns_prepare_beams(const float* ranges, int n_beams, float angle_min, float angle_inc, float range_min, float range_max,
                              std::vector<float2>& pts) const {
    pts.clear();
    const int stride = std::max(1, cfg.ns_beam_stride);
    int kept = 0;
    for (int i = 0; i < n_beams; ++i) {
        const double r = ranges[i];
        if (std::isnan(r) || std::isinf(r)) continue;
        if (!(r >= range_min && r <= range_max) || r >= cfg.ns_max_range) continue;
        const double ang = (double)angle_min + ((size_t)i * (double)angle_inc);
        if (cfg.ns_use_fov) {
            const double deg = ang * 180.0 / M_PI;
            if (!(deg > cfg.fov_lower_deg && deg < cfg.fov_upper_deg)) continue;
        }
        if ((kept++ % stride) != 0) continue;
        const double phi = -ang;                                        // the reference mirrors beam angles (MC:653)
        const double inv_res = 1.0 / (double)res_f;                     // beam points in CELL units
        pts.push_back(make_float2((float)((cfg.laser_offset + r * std::cos(phi)) * inv_res), (float)((r * std::sin(phi)) * inv_res)));
    }
}

int Engine::ns_update_local(const float* ranges, int n_beams, float angle_min, float angle_inc, float range_min, float range_max,
                            float* local_max) {
    CK(cudaSetDevice(cfg.device));
    if (cfg.mode != MCL_MODE_NS) return fail(MCL_ERR_STATE, "ns_update_local: NS mode only");
    if (n_beams < 0 || (n_beams > 0 && !ranges)) return fail(MCL_ERR_ARG, "update: bad scan");
    std::vector<float2> pts;
    ns_prepare_beams(ranges, n_beams, angle_min, angle_inc, range_min, range_max, pts);
    CK(d_ns_beams.ensure(std::max<size_t>(1, pts.size())));
    if (!pts.empty()) {
        int rc = ensure_pinned(pts.size() * sizeof(float2));
        if (rc) return rc;
        memcpy(h_pinned, pts.data(), pts.size() * sizeof(float2));
        CK(cudaMemcpyAsync(d_ns_beams.p, h_pinned, pts.size() * sizeof(float2), cudaMemcpyHostToDevice, stream));
    }
    return ns_run_update(d_ns_beams.p, (int)pts.size(), local_max);
}

int Engine::ns_stage_scan(int slot, const float* ranges, int n_beams, float angle_min, float angle_inc, float range_min, float range_max) {
    if ((size_t)slot >= ns_staged.size()) ns_staged.resize(slot + 1);
    std::vector<float2> pts;
    ns_prepare_beams(ranges, n_beams, angle_min, angle_inc, range_min, range_max, pts);
    NsStagedScan& s = ns_staged[slot];
    // steps queued by mcl_ns_step_staged may still read this slot: drain the (non-blocking) stream, then copy on it
    CK(cudaStreamSynchronize(stream));
    CK(s.d_pts.ensure(std::max<size_t>(1, pts.size())));
    if (!pts.empty()) {
        CK(cudaMemcpyAsync(s.d_pts.p, pts.data(), pts.size() * sizeof(float2), cudaMemcpyHostToDevice, stream));
        CK(cudaStreamSynchronize(stream));
    }
    s.n = (int)pts.size();
    s.valid = true;
    return MCL_OK;
}

int Engine::ns_update_local_staged(int slot, float* local_max) {
    CK(cudaSetDevice(cfg.device));
    if (cfg.mode != MCL_MODE_NS) return fail(MCL_ERR_STATE, "ns_update_local_staged: NS mode only");
    if (slot < 0 || (size_t)slot >= ns_staged.size() || !ns_staged[slot].valid) return fail(MCL_ERR_ARG, "update_staged: empty slot");
    return ns_run_update(ns_staged[slot].d_pts.p, ns_staged[slot].n, local_max);
}

// The likelihood-field kernel over this shard. Returns the shard's maximum log-likelihood.
int Engine::ns_run_update(const float2* d_pts, int n_pts, float* local_max) {
    ns_maxbits_prepped = false;                  // phase by phase: only mcl_ns_step's own predict prepares the step
    int rc = ns_launch_update(d_pts, n_pts);
    if (rc) return rc;
    int bits = 0;
    CK(cudaMemcpyAsync(&bits, d_maxbits.p, sizeof(int), cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
    bits = bits >= 0 ? bits : bits ^ 0x7fffffff;
    float mx;
    memcpy(&mx, &bits, 4);
    ns_last_max = mx;
    if (local_max) *local_max = mx;
    return MCL_OK;
}

int Engine::ns_launch_update(const float2* d_pts, int n_pts) {
    if (!map_ready) return fail(MCL_ERR_ARG, "update: no map");
    if (n == 0) return fail(MCL_ERR_ARG, "update: no particles");
    ns_beams_n = n_pts;
    NsField F;
    F.lf = d_lf.p; F.W = map_w; F.H = map_h; F.pad = lf_pad; F.Wp = lf_wp; F.ox = (float)origin_x; F.oy = (float)origin_y;
    F.inv_res = 1.0f / res_f; F.lf_out = lf_out; F.bytes_padded = (int)lf_bytes_padded;
    if (!ns_maxbits_prepped) {
        const int init_bits = INT32_MIN;
        CK(cudaMemcpyAsync(d_maxbits.p, &init_bits, sizeof(int), cudaMemcpyHostToDevice, stream));
    }
    ns_maxbits_prepped = false;
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, cfg.device);
    F.lf8 = d_lf8.p; F.codes = d_codes.p; F.n_codes = ns_n_codes;
    const size_t beam_bytes = (size_t)ns_beams_n * sizeof(float2);
    // Where the field is read from. Shared memory (TMA-staged) when it fits; else fp32 through L1/L2 while that stays
    // L2-resident next to the particle stream; else the one-byte coded field (same values, a quarter of the footprint).
    const size_t l2_budget = 96ull << 20;
    int kind = lf_bytes_padded + beam_bytes <= 200 * 1024 ? NS_FIELD_SMEM : NS_FIELD_GLOBAL;
    bool tuning = false;
    if (kind == NS_FIELD_GLOBAL && lf_bytes_padded > l2_budget && ns_n_codes && ns_force_field < 0) {
        // both global forms are candidates: keep the faster one by measurement, re-trying the other every 32nd launch
        tuning = true;
        if (!ns_tune_ev[0]) { CK(cudaEventCreate(&ns_tune_ev[0])); CK(cudaEventCreate(&ns_tune_ev[1])); }
        if (ns_tune_pending && cudaEventQuery(ns_tune_ev[1]) == cudaSuccess) {
            float ms = 0.f;
            if (cudaEventElapsedTime(&ms, ns_tune_ev[0], ns_tune_ev[1]) == cudaSuccess && ns_beams_n > 0) ns_tune_cost[ns_tune_kind] = ms / ns_tune_beams;
            ns_tune_pending = false;
        }
        cudaGetLastError();
        const double cu8 = ns_tune_cost[NS_FIELD_U8], cf32 = ns_tune_cost[NS_FIELD_GLOBAL];
        if (cu8 < 0) kind = NS_FIELD_U8;                                  // uniform start: the coded field first
        else if (cf32 < 0) kind = NS_FIELD_GLOBAL;
        else {
            kind = cu8 <= cf32 ? NS_FIELD_U8 : NS_FIELD_GLOBAL;
            if (ns_tune_launches % 32 == 31) kind = kind == NS_FIELD_U8 ? NS_FIELD_GLOBAL : NS_FIELD_U8;
        }
        ++ns_tune_launches;
    }
    if (ns_force_field == NS_FIELD_U8 && ns_n_codes) kind = NS_FIELD_U8;
    if (ns_force_field == NS_FIELD_GLOBAL) kind = NS_FIELD_GLOBAL;
    // arithmetic form (identical values): 0 scalar, 1 everything packed (FFMA2/FADD2), 2 only the adds packed. Defaults by
    // measurement (profiles/); MCL_NS_PACK=<smem><global><u8> digits override (experiments), debug bit 5 forces scalar.
    static const char* env_pack = getenv("MCL_NS_PACK");
    int pack = kind == NS_FIELD_SMEM ? 0 : kind == NS_FIELD_GLOBAL ? 1 : 2;      // measured: 158/158/161 us, 3416/3341/3286 us, 4472/4294/4428 us
    if (env_pack && strlen(env_pack) == 3 && env_pack[kind] >= '0' && env_pack[kind] <= '2') pack = env_pack[kind] - '0';
    if (ns_force_scalar) pack = 0;
    if (!ns_attr_set) {
#define X(K, P) CK(cudaFuncSetAttribute(k_ns_update<K, P>, cudaFuncAttributeMaxDynamicSharedMemorySize, (K == NS_FIELD_SMEM ? 200 : 64) * 1024));
        NS_UPD_FOR_ALL(X)
#undef X
        ns_attr_set = true;
    }
    const int threads = NS_UPD_THREADS;
    const int64_t batches = (n + 31) / 32;
    const int64_t ctas_needed = (batches + threads / 32 - 1) / (threads / 32);
    const size_t smem = beam_bytes + (kind == NS_FIELD_SMEM ? lf_bytes_padded : kind == NS_FIELD_U8 ? NS_MAX_CODES * sizeof(float) : 0);
    if (kind != NS_FIELD_SMEM && smem > 64 * 1024) return fail(MCL_ERR_ARG, "update: too many beams");
    auto launch = [&](auto kernel) -> int {
        int occ = 1;
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, threads, smem));
        const int grid = (int)std::min<int64_t>((int64_t)sms * std::max(1, occ), ctas_needed);     // persistent: resident CTAs only
        const bool measure = tuning && !ns_tune_pending && ns_beams_n > 0;
        if (measure) CK(cudaEventRecord(ns_tune_ev[0], stream));
        LAUNCH_PDL(K_NS_UPDATE, kernel, std::max(1, grid), threads, smem, part[cur].p, n, F, d_pts, ns_beams_n, d_ll.p, d_maxbits.p);
        if (measure) { CK(cudaEventRecord(ns_tune_ev[1], stream)); ns_tune_pending = true; ns_tune_kind = kind; ns_tune_beams = ns_beams_n; }
        return MCL_OK;
    };
    int lrc = MCL_ERR_ARG;
#define X(K, P) if (kind == K && pack == P) lrc = launch(k_ns_update<K, P>);
    NS_UPD_FOR_ALL(X)
#undef X
    if (lrc != MCL_OK) return lrc;
    ns_field_kind = kind;
    CK(cudaGetLastError());
    ns_have_ll = true;
    have_weights = false;
    return MCL_OK;
}

int Engine::ns_weights_local(float global_max, uint64_t* local_total) {
    CK(cudaSetDevice(cfg.device));
    if (cfg.mode != MCL_MODE_NS) return fail(MCL_ERR_STATE, "ns_weights_local: NS mode only");
    if (!ns_have_ll) return fail(MCL_ERR_ARG, "weights: run the update first");
    int gbits;
    memcpy(&gbits, &global_max, 4);
    gbits = gbits >= 0 ? gbits : gbits ^ 0x7fffffff;
    CK(cudaMemcpyAsync(d_maxbits.p, &gbits, sizeof(int), cudaMemcpyHostToDevice, stream));
    ns_scan_prepped = false;
    int rc0 = ns_launch_weights();
    if (rc0) return rc0;
    uint64_t tot = 0;
    CK(cudaMemcpyAsync(&tot, d_u64.p, sizeof(uint64_t), cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
    have_weights = true;
    last_total = (double)tot * 2.3283064365386963e-10;
    if (local_total) *local_total = tot;
    return MCL_OK;
}

// One wave of tiles or less: single pass with decoupled look-back (one launch, weights computed once). More: the
// look-back frontier (32 tiles per L2 round trip) would cap the rate below HBM speed (measured 1.7 TB/s), so the
// dependency-free two-pass form takes over (tile sums, then the prefix with offsets read from the sums).
void Engine::ns_scan_shape(bool& two_pass, int& nt, int& ng) const {
    nt = (int)((n + NS_SCAN_TILE - 1) / NS_SCAN_TILE);
    ng = (nt + NS_SCAN_GROUP - 1) / NS_SCAN_GROUP;
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, cfg.device);
    two_pass = force_sequential || nt > sms * (2048 / NS_SCAN_THREADS);
}

int Engine::ns_launch_weights() {
    const float temper = (float)cfg.ns_temper;
    bool two_pass; int nt, ng;
    ns_scan_shape(two_pass, nt, ng);
    const bool prepped = ns_scan_prepped;              // mcl_ns_step: k_ns_predict already cleared the scan's scratch
    ns_scan_prepped = false;
    if (two_pass) {
        uint64_t* group_sums = d_tile_sums.p + nt;
        if (!prepped) CK(cudaMemsetAsync(group_sums, 0, (size_t)ng * sizeof(uint64_t), stream));
        LAUNCH_PDL(K_NS_WSUM, k_ns_weights_sum, nt, NS_SCAN_THREADS, 0, d_ll.p, n, d_maxbits.p, temper, d_tile_sums.p, group_sums);
        LAUNCH_PDL(K_NS_WSCAN, k_ns_weights_scan, nt, NS_SCAN_THREADS, 0, d_ll.p, n, d_maxbits.p, temper, d_tile_sums.p, group_sums, nt, d_prefix.p, d_u64.p);
    } else {
        if (!prepped) CK(cudaMemsetAsync(d_tile_sums.p, 0, (size_t)(nt + 1) * sizeof(uint64_t), stream));      // tile states + ticket
        LAUNCH_PDL(K_NS_WSCAN, k_ns_weights_scan1, nt, NS_SCAN_THREADS, 0, d_ll.p, n, d_maxbits.p, temper, d_tile_sums.p, nt, d_prefix.p, d_u64.p);
    }
    CK(cudaGetLastError());
    ns_w_in_records = false;
    return MCL_OK;
}

// The fp32 weights into the particle records, for callers that look at the particles between update and resample.
int Engine::ns_materialise_weights() {
    if (cfg.mode != MCL_MODE_NS || !have_weights || ns_w_in_records) return MCL_OK;
    LAUNCH(K_NS_WSCAN, k_ns_materialise_w, grid_for(n, 256), 256, 0, d_ll.p, n, d_maxbits.p, (float)cfg.ns_temper, part[cur].p);
    CK(cudaGetLastError());
    ns_w_in_records = true;
    return MCL_OK;
}

uint32_t Engine::ns_u0() const {
    uint32_t o[4];
    Philox::gen(0u, 0u, 0x40u, (uint32_t)step_counter, (uint32_t)cfg.seed, (uint32_t)(cfg.seed >> 32), o);
    return o[0];
}

int Engine::ns_resample_local(uint64_t offset, uint64_t total, uint32_t u0, int64_t* k_lo_out, int64_t* k_hi_out) {
    CK(cudaSetDevice(cfg.device));
    if (cfg.mode != MCL_MODE_NS) return fail(MCL_ERR_STATE, "ns_resample_local: NS mode only");
    if (!have_weights) return fail(MCL_ERR_ARG, "resample: run the update first");
    if (total == 0) return fail(MCL_ERR_ARG, "resample: total weight is zero");
    uint64_t mine = 0;
    CK(cudaMemcpyAsync(&mine, d_u64.p, sizeof(uint64_t), cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
    NsPlan P;
    P.offset = offset; P.total = total; P.dq1 = P.dr1 = P.dqs = P.drs = 0;
    P.k_lo = ns::first_slot(offset, total, (uint64_t)n_global, u0);
    P.k_hi = ns::first_slot(offset + mine, total, (uint64_t)n_global, u0);
    CK(d_plan.ensure(sizeof(NsPlan)));
    LAUNCH(K_NS_PLAN, k_ns_plan_set, 1, 1, 0, P, (uint64_t)n_global, (NsPlan*)d_plan.p);
    int rc = ns_launch_resample(u0);
    if (rc) return rc;
    CK(cudaStreamSynchronize(stream));       // peers may only swap once every shard's stores have landed
    if (k_lo_out) *k_lo_out = P.k_lo;
    if (k_hi_out) *k_hi_out = P.k_hi;
    return MCL_OK;
}

// Resampling proper, plan in d_plan: merge-path partition (ancestor of every output tile's first slot), then the tiles.
int Engine::ns_launch_resample(uint32_t u0) {
    NsDest D;
    D.per_rank = per_rank; D.world = shard_world;
    const int next = cur ^ 1;
    for (int r = 0; r < 8; ++r) { D.part[r] = nullptr; D.anc[r] = nullptr; }
    for (int r = 0; r < shard_world; ++r) {
        if (r == shard_rank) { D.part[r] = part[next].p; D.anc[r] = ancestors.p; }
        else {
            if (!peer_ptr[next][r] || !peer_ptr[2][r])
                return fail(MCL_ERR_COMM, "resample: peer buffers of a shard are not mapped (mcl_comm_init, or mcl_peer_import / mcl_peer_set)");
            D.part[r] = (float4*)peer_ptr[next][r]; D.anc[r] = (int*)peer_ptr[2][r];
        }
    }
    // a shard can own up to all n_global output slots
    const int64_t max_tiles = (n_global + NS_RS_TILE - 1) / NS_RS_TILE;
    CK(d_bounds.ensure(((size_t)max_tiles + 2) * sizeof(NsTileHead)));
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, cfg.device);
    const int g1 = (int)std::min<int64_t>((max_tiles + 1 + 7) / 8, (int64_t)sms * 8);            // one warp per tile head, 8 per CTA
    LAUNCH_PDL(K_NS_BOUNDS, k_ns_resample_bounds, g1, 256, 0, d_prefix.p, n, (const NsPlan*)d_plan.p, (uint64_t)n_global, u0, (NsTileHead*)d_bounds.p);
    int occ = 8;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_ns_resample, NS_RS_THREADS, 0));
    const int g2 = (int)std::min<int64_t>(max_tiles, (int64_t)sms * std::max(1, occ));      // resident CTAs only: tiles are taken grid-stride
    LAUNCH_PDL(K_NS_RESAMPLE, k_ns_resample, g2, NS_RS_THREADS, 0, part[cur].p, d_prefix.p, n, shard_begin, (const NsPlan*)d_plan.p,
           (const NsTileHead*)d_bounds.p, (uint64_t)n_global, 1.0 / (double)n_global, D, (float)(1.0 / (double)n_global));
    CK(cudaGetLastError());
    return MCL_OK;
}

int Engine::ns_end_step() {
    cur ^= 1;
    have_weights = false; ns_have_ll = false;
    ++step_counter;
    return MCL_OK;
}

int Engine::ns_pose_partials(double* out5) {
    CK(cudaSetDevice(cfg.device));
    if (n == 0 || !out5) return fail(MCL_ERR_ARG, "pose_partials: no particles");
    const int blocks = (int)std::min<int64_t>(148 * 8, grid_for(n, 256));      // one wave of resident CTAs, grid-stride
    CK(d_partials.ensure(5 * 2048));
    const bool from_ll = have_weights && !ns_w_in_records;
    LAUNCH_PDL(K_NS_POSE, k_ns_pose_partials, blocks, 256, 0, part[cur].p, n, from_ll ? (const float*)d_ll.p : (const float*)nullptr, (const int*)d_maxbits.p,
           (float)cfg.ns_temper, d_partials.p);
    CK(cudaGetLastError());
    std::vector<double> h((size_t)blocks * 5);
    CK(cudaMemcpyAsync(h.data(), d_partials.p, h.size() * sizeof(double), cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
    for (int k = 0; k < 5; ++k) { double s = 0; for (int b = 0; b < blocks; ++b) s += h[(size_t)b * 5 + k]; out5[k] = s; }
    return MCL_OK;
}

int Engine::ns_download_loglik(float* ll) {
    CK(cudaSetDevice(cfg.device));
    if (!ll || !ns_have_ll) return fail(MCL_ERR_ARG, "download_loglik: run the update first");
    CK(cudaMemcpyAsync(ll, d_ll.p, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
    return MCL_OK;
}
int Engine::ns_download_prefix(uint64_t* prefix) {
    CK(cudaSetDevice(cfg.device));
    if (!prefix || !have_weights) return fail(MCL_ERR_ARG, "download_prefix: run the update first");
    CK(cudaMemcpyAsync(prefix, d_prefix.p, (size_t)n * sizeof(uint64_t), cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
    return MCL_OK;
}

int Engine::ns_last_plan(int64_t* k_lo, int64_t* k_hi, int64_t* own_begin, int64_t* own_count) {
    CK(cudaSetDevice(cfg.device));
    if (cfg.mode != MCL_MODE_NS || !d_plan.p) return fail(MCL_ERR_STATE, "ns_last_plan: no NS step yet");
    NsPlan P;
    CK(cudaMemcpyAsync(&P, d_plan.p, sizeof(P), cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
    if (k_lo) *k_lo = P.k_lo;
    if (k_hi) *k_hi = P.k_hi;
    if (own_begin) *own_begin = shard_begin;
    if (own_count) *own_count = n;
    return MCL_OK;
}

// ---- engine-native collectives: NCCL on the engine's stream ------------------------------------------------------------------
#define NCK(call)                                                                                          \
    do {                                                                                                   \
        ncclResult_t r__ = (call);                                                                         \
        if (r__ != ncclSuccess) { err = std::string("NCCL error: ") + nccl_api().GetErrorString(r__) + " at " #call; return MCL_ERR_COMM; } \
    } while (0)

void Engine::ns_comm_destroy() {
    if (comm && nccl_api().lib) nccl_api().CommDestroy((ncclComm_t)comm);
    comm = nullptr;
}

int Engine::comm_unique_id(void* out128) {
    std::string e;
    if (!nccl_api().load(e)) return fail(MCL_ERR_COMM, e);
    ncclUniqueId id;
    NCK(nccl_api().GetUniqueId(&id));
    memcpy(out128, &id, sizeof(id));
    return MCL_OK;
}

// Joins the communicator (collective: every shard calls it), then maps every peer's particle / ancestor buffers through
// CUDA IPC; the 64-byte handles travel by ncclAllGather, so no other transport is needed.
int Engine::comm_init(const void* id128) {
    CK(cudaSetDevice(cfg.device));
    if (cfg.mode != MCL_MODE_NS) return fail(MCL_ERR_STATE, "comm_init: NS mode only");
    if (n == 0) return fail(MCL_ERR_ARG, "comm_init: call mcl_ns_set_shard first");
    std::string e;
    if (!nccl_api().load(e)) return fail(MCL_ERR_COMM, e);
    ncclUniqueId id;
    memcpy(&id, id128, sizeof(id));
    ncclComm_t c = nullptr;
    NCK(nccl_api().CommInitRank(&c, shard_world, id, shard_rank));
    comm = c;
    CK(d_totals.ensure(8)); CK(d_plan.ensure(sizeof(NsPlan))); CK(d_pose.ensure(8)); CK(d_bar.ensure(1));
    if (shard_world > 1) {
        DevBuf<unsigned char> mine, all;
        int mrc = ensure_mailbox();
        if (mrc) return mrc;
        CK(mine.ensure(256)); CK(all.ensure((size_t)256 * shard_world));
        unsigned char h[256];
        for (int w = 0; w < 4; w++) { int rc = peer_export(w, h + 64 * w); if (rc) return rc; }
        CK(cudaMemcpyAsync(mine.p, h, 256, cudaMemcpyHostToDevice, stream));
        NCK(nccl_api().AllGather(mine.p, all.p, 256, ncclUint8, (ncclComm_t)comm, stream));
        std::vector<unsigned char> hall((size_t)256 * shard_world);
        CK(cudaMemcpyAsync(hall.data(), all.p, hall.size(), cudaMemcpyDeviceToHost, stream));
        CK(cudaStreamSynchronize(stream));
        for (int r = 0; r < shard_world; r++) {
            if (r == shard_rank) continue;
            for (int w = 0; w < 4; w++) { int rc = peer_import(r, w, hall.data() + (size_t)256 * r + 64 * w); if (rc) return rc; }
        }
        mine.release(); all.release();
    }
    return MCL_OK;
}

// One whole filter step of a (possibly sharded) NS filter, enqueued on the engine's stream with no host round trip:
//   predict -> likelihood field -> all-reduce(max) -> Q32 weights + prefix -> all-gather(totals) -> device plan ->
//   resample straight into the owning shards -> closing all-reduce (barrier) -> swap.
// pose3 != null additionally reduces the weighted pose (before resampling) and waits for it.
int Engine::ns_step(double rot1, double trans, double rot2, int slot, const float* ranges, int n_beams, float angle_min, float angle_inc,
                    float range_min, float range_max, double* pose3) {
    CK(cudaSetDevice(cfg.device));
    if (cfg.mode != MCL_MODE_NS) return fail(MCL_ERR_STATE, "ns_step: NS mode only");
    // the step's collectives: peer-memory mailboxes (default once every peer's mailbox is mapped) or NCCL
    static const int env_exchange = [] { const char* e = getenv("MCL_NS_EXCHANGE"); return !e ? -1 : !strcmp(e, "nccl") ? 0 : !strcmp(e, "peer") ? 1 : -1; }();
    const int want = ns_exchange >= 0 ? ns_exchange : env_exchange;
    const bool mail = shard_world > 1 && want != 0 && peers_have_mailboxes();
    if (shard_world > 1 && want == 1 && !mail) return fail(MCL_ERR_COMM, "ns_step: peer-memory exchange asked for but a peer mailbox is not mapped");
    if (shard_world > 1 && !mail && !comm) return fail(MCL_ERR_COMM, "ns_step: sharded filter without a communicator (mcl_comm_init)");
    NsPeers PX;
    PX.world = shard_world; PX.rank = shard_rank;
    static const double env_timeout_s = [] { const char* e = getenv("MCL_NS_EXCHANGE_TIMEOUT_S"); return e ? atof(e) : 30.0; }();
    PX.timeout_ns = env_timeout_s > 0 ? (unsigned long long)(env_timeout_s * 1e9) : 0ull;
    for (int r = 0; r < 8; ++r) PX.box[r] = r >= shard_world ? nullptr : r == shard_rank ? (NsMailbox*)d_mbox.p : (NsMailbox*)peer_ptr[3][r];
    // every argument check comes before the exchange tag advances: a shard that returned an error with its tag already
    // incremented would never again match its peers' tags
    if (n == 0) return fail(MCL_ERR_ARG, "ns_step: no particles");
    if (!map_ready) return fail(MCL_ERR_ARG, "ns_step: no map");
    if (ranges ? (n_beams < 0) : (slot < 0 || (size_t)slot >= ns_staged.size() || !ns_staged[slot].valid))
        return fail(MCL_ERR_ARG, ranges ? "ns_step: bad scan" : "ns_step: empty scan slot");
    if (shard_world > 1) ns_exchange_used = mail ? 1 : 0;
    CK(d_totals.ensure(8)); CK(d_plan.ensure(sizeof(NsPlan))); CK(d_pose.ensure(8)); CK(d_bar.ensure(1));
    Motion m; m.rot_1 = rot1; m.trans = trans; m.rot_2 = rot2;
    int rc;
    // the scan goes to the device first, so that no copy command sits between the step's kernels (programmatic launches
    // overlap only kernel with kernel)
    const float2* d_pts; int n_pts;
    if (ranges) {
        std::vector<float2> pts;
        ns_prepare_beams(ranges, n_beams, angle_min, angle_inc, range_min, range_max, pts);
        CK(d_ns_beams.ensure(std::max<size_t>(1, pts.size())));
        if (!pts.empty()) {
            // a small ring of pinned slots so the host never overwrites a scan the GPU has not consumed yet
            rc = ensure_pinned_ring(pts.size() * sizeof(float2));
            if (rc) return rc;
            void* hp = pinned_ring_next();
            memcpy(hp, pts.data(), pts.size() * sizeof(float2));
            CK(cudaMemcpyAsync(d_ns_beams.p, hp, pts.size() * sizeof(float2), cudaMemcpyHostToDevice, stream));
            CK(cudaEventRecord(ring_events[ring_pos], stream));
        }
        d_pts = d_ns_beams.p; n_pts = (int)pts.size();
    } else {
        d_pts = ns_staged[slot].d_pts.p; n_pts = ns_staged[slot].n;
    }
    rc = ns_predict(m, true);
    if (rc) return rc;
    // sensor model (asynchronous variant of ns_run_update: the max stays on the device)
    rc = ns_launch_update(d_pts, n_pts);
    if (rc) return rc;
    auto& N = nccl_api();
    const unsigned tag = mail ? ++xchg_seq : 0u;       // advanced only now, immediately before the first exchange kernel is enqueued
    const int parity = (int)(tag & 1u);
    if (mail) LAUNCH(K_NS_PLAN, k_ns_xchg_max, 1, 32, 0, d_maxbits.p, PX, tag, parity);
    else if (shard_world > 1) NCK(N.AllReduce(d_maxbits.p, d_maxbits.p, 1, ncclInt32, ncclMax, (ncclComm_t)comm, stream));
    if (shard_world > 1) pdl_hold = true;        // the exchange polls other shards: its successor is launched in stream order
    rc = ns_launch_weights();
    if (rc) return rc;
    if (mail) {}                                                              // gathered inside k_ns_plan_xchg
    else if (shard_world > 1) { NCK(N.AllGather(d_u64.p, d_totals.p, 1, ncclUint64, (ncclComm_t)comm, stream)); pdl_hold = true; }
    // (one shard: k_ns_plan reads the local total where the scan left it)
    {   // the weighted-mean pose (before resampling) is part of every step; it crosses to the host only when pose3 asks
        const int blocks = (int)std::min<int64_t>(148 * 8, grid_for(n, 256));      // one wave of resident CTAs, grid-stride
        CK(d_partials.ensure(5 * 2048));
        LAUNCH_PDL(K_NS_POSE, k_ns_pose_partials, blocks, 256, 0, part[cur].p, n, (const float*)d_ll.p, (const int*)d_maxbits.p, (float)cfg.ns_temper,
               d_partials.p);
        if (shard_world == 1 && !h_ns_pose) CK(cudaMallocHost((void**)&h_ns_pose, 8 * sizeof(double)));
        LAUNCH_PDL(K_NS_POSE_REDUCE, k_ns_pose_reduce, 1, 160, 0, d_partials.p, blocks, d_pose.p, shard_world == 1 ? h_ns_pose : (double*)nullptr);
        if (mail) {}                                                          // summed over the shards inside k_ns_plan_xchg
        else if (shard_world > 1) { NCK(N.AllReduce(d_pose.p, d_pose.p, 5, ncclFloat64, ncclSum, (ncclComm_t)comm, stream)); pdl_hold = true; }
    }
    const uint32_t u0 = ns_u0();
    if (mail) LAUNCH(K_NS_PLAN, k_ns_plan_xchg, 1, 32, 0, d_u64.p, d_pose.p, PX, tag, parity, (uint64_t)n_global, u0, (NsPlan*)d_plan.p, d_totals.p);
    else LAUNCH_PDL(K_NS_PLAN, k_ns_plan, 1, 32, 0, shard_world > 1 ? d_totals.p : d_u64.p, shard_world, shard_rank, (uint64_t)n_global, u0, (NsPlan*)d_plan.p);
    have_weights = true;
    if (shard_world > 1) pdl_hold = true;
    rc = ns_launch_resample(u0);
    if (rc) return rc;
    if (mail) LAUNCH(K_NS_PLAN, k_ns_xchg_barrier, 1, 32, 0, PX, tag);                                            // closing barrier
    else if (shard_world > 1) NCK(N.AllReduce(d_bar.p, d_bar.p, 1, ncclInt32, ncclSum, (ncclComm_t)comm, stream));
    if (shard_world > 1) pdl_hold = true;        // next step's first kernel follows the closing barrier in stream order
    cur ^= 1;
    have_weights = false; ns_have_ll = false;
    ++step_counter;
    if (pose3) {
        double h[6] = {0, 0, 0, 0, 0, 0};
        if (shard_world == 1) {                      // the sums are already on their way to the pinned block
            CK(cudaStreamSynchronize(stream));
            memcpy(h, h_ns_pose, 5 * sizeof(double));
        } else {
            CK(cudaMemcpyAsync(h, d_pose.p, (mail ? 6 : 5) * sizeof(double), cudaMemcpyDeviceToHost, stream));
            CK(cudaStreamSynchronize(stream));
        }
        if (mail && h[5] != 0.0) return fail(MCL_ERR_COMM, "ns_step: a peer-memory exchange timed out (a shard never posted)");
        pose3[0] = h[1] / h[0]; pose3[1] = h[2] / h[0]; pose3[2] = std::atan2(h[3], h[4]);
    }
    return MCL_OK;
}

int Engine::ensure_pinned_ring(size_t bytes) {
    if (bytes <= ring_bytes && ring_base) return MCL_OK;
    if (ring_base) { cudaStreamSynchronize(stream); cudaFreeHost(ring_base); for (auto& e : ring_events) cudaEventDestroy(e); ring_events.clear(); }
    ring_bytes = std::max<size_t>(bytes, 16 * 1024);
    CK(cudaMallocHost(&ring_base, ring_bytes * RING));
    ring_events.resize(RING);
    for (auto& e : ring_events) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    ring_pos = 0;
    return MCL_OK;
}
void* Engine::pinned_ring_next() {
    ring_pos = (ring_pos + 1) % RING;
    cudaEventSynchronize(ring_events[ring_pos]);          // the copy that last used this slot has completed
    return (char*)ring_base + ring_bytes * ring_pos;
}

// Random-gather micro-benchmark: reads/s for a table of `table_bytes` in shared memory (tier 0) or global
// memory (tier 1: whatever level of the hierarchy a table of that size lives in).
int Engine::gather_bench(int tier, size_t table_bytes, int iters, double* reads_per_s) {
    CK(cudaSetDevice(cfg.device));
    if (!reads_per_s || table_bytes < 64 || iters < 1) return fail(MCL_ERR_ARG, "gather_bench: bad argument");
    if (tier == 0 && table_bytes > 200 * 1024) return fail(MCL_ERR_ARG, "gather_bench: shared-memory tier is limited to 200 KiB");
    const uint32_t n_words = (uint32_t)(table_bytes / 4);
    DevBuf<float> table, out;
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, cfg.device);
    const int threads = 512;
    const int ctas_per_sm = tier == 0 ? std::max(1, std::min(4, (int)((220 * 1024) / (table_bytes + 1024)))) : 4;
    const int grid = sms * ctas_per_sm;
    CK(table.ensure(n_words)); CK(out.ensure((size_t)grid * threads));
    CK(cudaMemsetAsync(table.p, 0, (size_t)n_words * 4, stream));
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    float best = 1e30f;
    for (int rep = 0; rep < 4; rep++) {
        CK(cudaEventRecord(a, stream));
        if (tier == 0) {
            CK(cudaFuncSetAttribute(k_gather_bench<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            k_gather_bench<true><<<grid, threads, table_bytes, stream>>>(table.p, n_words, iters, out.p);
        } else {
            k_gather_bench<false><<<grid, threads, 0, stream>>>(table.p, n_words, iters, out.p);
        }
        ++launches;
        CK(cudaEventRecord(b, stream));
        CK(cudaEventSynchronize(b));
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, a, b));
        if (rep > 0) best = std::min(best, ms);
    }
    cudaEventDestroy(a); cudaEventDestroy(b);
    CK(cudaGetLastError());
    *reads_per_s = (double)grid * threads * iters / (best * 1e-3);
    table.release(); out.release();
    return MCL_OK;
}

// ---- peer memory: every shard's two particle buffers and its ancestor buffer are visible to the others -------------------
void* Engine::device_buffer(int which) {
    if (which == 0 || which == 1) return part[which].p;
    if (which == 2) return ancestors.p;
    if (which == 3) return d_mbox.p;
    return nullptr;
}
int Engine::peer_export(int which, void* out64) {
    CK(cudaSetDevice(cfg.device));
    void* p = device_buffer(which);
    if (!p || !out64) return fail(MCL_ERR_ARG, "peer_export: allocate the shard first (mcl_ns_set_shard)");
    cudaIpcMemHandle_t hnd;
    CK(cudaIpcGetMemHandle(&hnd, p));
    static_assert(sizeof(hnd) == 64, "cudaIpcMemHandle_t is 64 bytes");
    memcpy(out64, &hnd, 64);
    return MCL_OK;
}
int Engine::peer_import(int rank, int which, const void* in64) {
    CK(cudaSetDevice(cfg.device));
    if (rank < 0 || rank >= 8 || which < 0 || which > 3 || !in64) return fail(MCL_ERR_ARG, "peer_import: bad argument");
    cudaIpcMemHandle_t hnd;
    memcpy(&hnd, in64, 64);
    void* p = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&p, hnd, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) { err = std::string("cudaIpcOpenMemHandle: ") + cudaGetErrorString(e); return MCL_ERR_COMM; }
    peer_ptr[which][rank] = p; peer_ipc[which][rank] = true;
    return MCL_OK;
}
int Engine::peer_set(int rank, int which, void* devptr) {
    if (rank < 0 || rank >= 8 || which < 0 || which > 3) return fail(MCL_ERR_ARG, "peer_set: bad argument");
    peer_ptr[which][rank] = devptr; peer_ipc[which][rank] = false;
    return MCL_OK;
}

}  // namespace mcl
