// mcl_engine.cu — host orchestration of the B200 particle filter (one handle = one GPU = one stream).
#include "mcl_engine.hpp"

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <sstream>

#include "engine_internal.hpp"
#include "exact_scan.cuh"
#include "exact_scan_fused.cuh"
#include "kernels_ref.cuh"

namespace mcl {

Engine::Engine(const mcl_config& c) : cfg(c) {}

Engine::~Engine() {
    if (!opened) return;
    cudaSetDevice(cfg.device);
    part[0].release(); part[1].release(); cdf.release(); d_wraw.release(); d_wn.release(); xs_pub.release(); xs_blocks.release(); xs_counters.release(); xs_trace.release(); xs_tsum.release(); xs_toff.release(); xs_seq_s.release(); xs_tiles.release(); xs_entries.release(); xs_carry.release(); xs_seq_base.release(); xs_flag.release(); for (auto& e : ns_tune_ev) if (e) { cudaEventDestroy(e); e = nullptr; }
    d_guide.release(); d_mbox.release(); d_lf.release(); d_lf8.release(); d_codes.release(); d_code_of_d2.release(); d_lf_table.release(); d_ll.release(); d_d2.release(); d_g.release(); d_ns_beams.release(); d_prefix.release(); d_tile_sums.release(); d_u64.release(); d_maxbits.release();
    for (int w = 0; w < 4; w++) for (int r = 0; r < 8; r++) if (peer_ipc[w][r] && peer_ptr[w][r]) cudaIpcCloseMemHandle(peer_ptr[w][r]);
    d_assign.release(); d_km_reinit.release(); d_km.release(); d_posearr.release(); d_bounds.release(); d_totals.release(); d_plan.release(); d_pose.release(); d_bar.release();
    if (h_step) cudaFreeHost(h_step);
    if (h_touch_report) cudaFreeHost(h_touch_report);
    if (h_ns_pose) cudaFreeHost(h_ns_pose);
    d_inj.release();
    if (ring_base) { cudaFreeHost(ring_base); for (auto& e : ring_events) cudaEventDestroy(e); }
    ns_comm_destroy(); ancestors.release(); d_occ.release(); d_occ_pad.release(); d_gauss.release();
    d_radii.release(); d_lut.release(); d_lut_filled.release(); d_touch.release(); d_touch_theta.release();
    d_beams.release(); for (auto& sc : staged) sc.d_used.release(); for (auto& sc : ns_staged) sc.d_pts.release(); d_u_r.release(); d_u_jit.release(); d_inj_f64.release(); d_inj_i32.release();
    d_block_counts.release(); d_counters.release(); d_scalars.release(); d_partials.release();
    if (h_pinned) cudaFreeHost(h_pinned);
    if (stream) cudaStreamDestroy(stream);
}

const char* Engine::kernel_name(int id) {
    static const char* names[K_COUNT] = {"k_ref_init", "k_ref_predict", "k_ref_first_touch", "k_ref_touch_theta", "k_ref_update", "k_ref_update_v2",
                                         "k_ref_seq_total", "k_fill_resample_draws", "k_ref_inject_count", "k_ref_ema",
                                         "k_ref_seq_cdf", "k_ref_guide", "k_ref_resample", "k_xs_tilesum", "k_xs_offsets", "k_xs_scan", "k_xs_chain", "k_xs_apply", "k_xs_total", "k_xs_cdf", "k_ns_edt_cols", "k_ns_edt_rows", "k_ns_init", "k_ns_predict", "k_ns_update", "k_ns_weights_sum", "k_ns_weights_scan", "k_ns_plan", "k_ns_resample_bounds", "k_ns_resample", "k_ns_pose_partials", "k_ns_pose_reduce", "k_km_assign", "k_km_update", "k_km_stats", "k_pose_array", "k_pose_wsum", "k_pose_sums", "k_reduce_partials", "k_ref_scans_one_tile"};
    return (id >= 0 && id < K_COUNT) ? names[id] : "?";
}
void Engine::profile_enable(bool on) {
    prof_collect();
    profiling = on;
    if (on) for (int i = 0; i < K_COUNT; i++) { prof_ms[i] = 0; prof_count[i] = 0; }
}
void Engine::prof_begin(int id) {
    if (!profiling) return;
    ProfEvent e; e.id = id;
    cudaEventCreate(&e.a); cudaEventCreate(&e.b);
    cudaEventRecord(e.a, stream);
    prof_events.push_back(e);
}
void Engine::prof_end() {
    if (!profiling) return;
    cudaEventRecord(prof_events.back().b, stream);
}
void Engine::prof_collect() {
    if (prof_events.empty()) return;
    cudaStreamSynchronize(stream);
    for (auto& e : prof_events) {
        float ms = 0;
        if (cudaEventElapsedTime(&ms, e.a, e.b) == cudaSuccess) { prof_ms[e.id] += ms; prof_count[e.id]++; }
        cudaEventDestroy(e.a); cudaEventDestroy(e.b);
    }
    prof_events.clear();
}
int Engine::profile_read(int id, double* total_ms, int64_t* count) {
    if (id < 0 || id >= K_COUNT) return fail(MCL_ERR_ARG, "profile_read: bad kernel id");
    CK(cudaSetDevice(cfg.device));
    prof_collect();
    if (total_ms) *total_ms = prof_ms[id];
    if (count) *count = prof_count[id];
    return MCL_OK;
}

int Engine::fail(int code, const std::string& what) { err = what; return code; }
int Engine::cuda_fail(cudaError_t e, const char* where) {
    err = std::string("CUDA error: ") + cudaGetErrorString(e) + " at " + where;
    return MCL_ERR_CUDA;
}

int Engine::ensure_pinned(size_t bytes) {
    if (bytes <= h_pinned_bytes) return MCL_OK;
    if (h_pinned) cudaFreeHost(h_pinned);
    h_pinned = nullptr; h_pinned_bytes = 0;
    CK(cudaMallocHost(&h_pinned, bytes));
    h_pinned_bytes = bytes;
    return MCL_OK;
}

int Engine::open() {
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(MCL_ERR_CUDA, std::string("no CUDA device: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0") +
                                      " (this engine has no CPU fallback)");
    if (cfg.device < 0 || cfg.device >= count) return fail(MCL_ERR_ARG, "device ordinal out of range");
    CK(cudaSetDevice(cfg.device));
    CK(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    opened = true;
    { const char* e_pdl = getenv("MCL_PDL"); use_pdl = !(e_pdl && e_pdl[0] == '0'); }
    {   // which float trig the kernels evaluate (include/mcl.h, MCL_TRIG_*)
        int rc = resolve_trig();
        if (rc) return rc;
    }
    // static tables
    gauss.build(cfg.sigma_hit);
    CK(d_gauss.ensure(gauss.v.size()));
    CK(cudaMemcpyAsync(d_gauss.p, gauss.v.data(), gauss.v.size() * sizeof(double), cudaMemcpyHostToDevice, stream));
    h_radii.clear();
    for (double r = 0.0; r < cfg.max_laser_range; r += cfg.ray_step) h_radii.push_back(r);       // MC:372
    CK(d_radii.ensure(h_radii.size()));
    CK(cudaMemcpyAsync(d_radii.p, h_radii.data(), h_radii.size() * sizeof(double), cudaMemcpyHostToDevice, stream));
    // ray-direction LUT: keys reachable by round(yaw_deg + off_deg), yaw in [-180,180], off in (-fov_hi, -fov_lo)
    const int span = 181 + (int)std::ceil(std::max(std::fabs(cfg.fov_lower_deg), std::fabs(cfg.fov_upper_deg)));
    key_min = -span;
    n_keys = 2 * span + 1;
    h_lut.assign(n_keys, make_double2(0.0, 0.0));
    h_lut_filled.assign(n_keys, 0);
    n_unfilled = n_keys;
    CK(d_lut.ensure(n_keys)); CK(d_lut_filled.ensure(n_keys)); CK(d_touch.ensure(n_keys)); CK(d_touch_theta.ensure(n_keys));
    CK(cudaMemcpyAsync(d_lut.p, h_lut.data(), n_keys * sizeof(double2), cudaMemcpyHostToDevice, stream));
    CK(cudaMemcpyAsync(d_lut_filled.p, h_lut_filled.data(), n_keys, cudaMemcpyHostToDevice, stream));
    CK(d_counters.ensure(8)); CK(d_scalars.ensure(8));
    { const double one = 1.0; CK(cudaMemcpyAsync(d_scalars.p + 7, &one, sizeof(double), cudaMemcpyHostToDevice, stream)); }      // divisor of an un-normalised CDF
    CK(cudaMemsetAsync(d_counters.p, 0, 8 * sizeof(int), stream));      // [4], [5] = tickets of k_pose_sums and k_ref_inject_count (they reset themselves); [6] = abort flag of an optimistic tick
    CK(d_partials.ensure(4 * 1024));
    CK(cudaStreamSynchronize(stream));
    return MCL_OK;
}

// MCL_TRIG_LIBM: find out which build of glibc's sinf/cosf this host's libm selected (ifunc: FMA + AVX2 CPUs get the fused
// build). The two builds differ at 17 magnitudes (34 floats) of the 2^32 arguments; the host libm is probed there and on a
// spread of ordinary arguments, and must agree with one restatement everywhere.
int Engine::resolve_trig() {
    if (cfg.trig_mode == MCL_TRIG_CORRECTLY_ROUNDED) { trig_kind = TRIG_CR; return MCL_OK; }
    static const uint32_t differ[] = {0x4255b0a9u, 0x42a35c07u, 0x42a35d44u, 0x42a97360u, 0x42cf5854u, 0x42e87a55u,            // sinf
                                      0x418a3adbu, 0x418a3adcu, 0x418a3addu, 0x418a3adeu, 0x41bc76d9u, 0x4202eb4bu, 0x42687a55u,
                                      0x4280ce28u, 0x42870e40u, 0x42c55faau, 0x42d8d23eu};                                    // cosf
    bool ok[2] = {true, true};
    auto probe = [&](uint32_t bits) {
        float y;
        memcpy(&y, &bits, 4);
        volatile float vy = y;
        const float hs = sinf(vy), hc = cosf(vy);
        const float s1 = glibc_trig::sinf_as_glibc<true>(y), c1 = glibc_trig::cosf_as_glibc<true>(y);
        const float s0 = glibc_trig::sinf_as_glibc<false>(y), c0 = glibc_trig::cosf_as_glibc<false>(y);
        if (memcmp(&hs, &s1, 4) || memcmp(&hc, &c1, 4)) ok[1] = false;
        if (memcmp(&hs, &s0, 4) || memcmp(&hc, &c0, 4)) ok[0] = false;
    };
    for (uint32_t b : differ) { probe(b); probe(b | 0x80000000u); }
    for (uint32_t k = 0; k < 4096; ++k) probe(0x30000000u + k * 0x0004F1A3u);      // 2^-31 .. 2^33: every path of the function
    if (ok[1] == ok[0])
        return fail(MCL_ERR_STATE, ok[1] ? "trig_mode MCL_TRIG_LIBM: the host libm's sinf/cosf could not be told apart (unexpected)"
                                         : "trig_mode MCL_TRIG_LIBM: the host libm's sinf/cosf is not glibc's (>= 2.28, x86-64) FMA or SSE2 build; "
                                           "use MCL_TRIG_CORRECTLY_ROUNDED");
    trig_kind = ok[1] ? TRIG_GLIBC_FMA : TRIG_GLIBC_SSE2;
    return MCL_OK;
}

int Engine::synchronize() {
    CK(cudaSetDevice(cfg.device));
    CK(cudaStreamSynchronize(stream));
    if (ns_exchange_used == 1 && d_mbox.p) {
        // last word group of the mailbox: {barrier[8], status, pad}
        int status = 0;
        CK(cudaMemcpy(&status, d_mbox.p + d_mbox.n - 8 * sizeof(int), sizeof(int), cudaMemcpyDeviceToHost));
        if (status) return fail(MCL_ERR_COMM, "a peer-memory exchange timed out (a shard never posted)");
    }
    return MCL_OK;
}

// ---- map ----------------------------------------------------------------------------------------------------
int Engine::set_map(const int8_t* occ, int w, int h, float res, double ox, double oy) {
    if (!occ || w <= 0 || h <= 0 || !(res > 0.f)) return fail(MCL_ERR_ARG, "set_map: bad grid");
    CK(cudaSetDevice(cfg.device));
    map_w = w; map_h = h; res_f = res; origin_x = ox; origin_y = oy;
    max_x = ox + w * res;      // int * float -> float, then + double (MC:688)
    max_y = oy + h * res;      // MC:689
    h_occ.assign(occ, occ + (size_t)w * h);
    std::vector<uint8_t> flags((size_t)w * h);
    for (size_t i = 0; i < flags.size(); ++i) flags[i] = occ[i] > 50 ? 1 : 0;    // MC:327,377
    CK(d_occ.ensure(flags.size()));
    CK(cudaMemcpyAsync(d_occ.p, flags.data(), flags.size(), cudaMemcpyHostToDevice, stream));
    // bordered ray-march table: 0 free, 1 occupied, 2 outside; row/column -1 replicate row/column 0 (truncation quirk, Q7).
    // Border = longest ray (max range + laser offset) in cells + 3, so a ray from a particle inside the map never leaves it.
    {
        const double reach = (cfg.max_laser_range + std::fabs(cfg.laser_offset)) / (double)res;
        occ_pad = reach < 1024.0 ? (int)std::ceil(reach) + 3 : 0;
        occ_wp = w + 2 * occ_pad;
        std::vector<uint8_t> padded((size_t)occ_wp * (h + 2 * occ_pad), 2);
        if (occ_pad > 0) {
            for (int y = -1; y < h; ++y)
                for (int x = -1; x < w; ++x)
                    padded[(size_t)(y + occ_pad) * occ_wp + (x + occ_pad)] = flags[(size_t)std::max(y, 0) * w + std::max(x, 0)];
        }
        CK(d_occ_pad.ensure(std::max<size_t>(1, padded.size())));
        CK(cudaMemcpyAsync(d_occ_pad.p, padded.data(), padded.size(), cudaMemcpyHostToDevice, stream));
    }
    CK(cudaStreamSynchronize(stream));
    map_ready = true;
    if (cfg.mode == MCL_MODE_NS) return ns_build_field();
    return MCL_OK;
}

int Engine::load_map_txt(const char* path) {
    std::ifstream f(path ? path : "");
    if (!f) return fail(MCL_ERR_IO, std::string("cannot open ") + (path ? path : "(null)"));
    std::stringstream ss; ss << f.rdbuf();
    WallGrid g; std::string perr;
    if (!parse_map_txt(ss.str(), g, perr)) return fail(MCL_ERR_IO, "map.txt: " + perr);
    std::vector<int8_t> occ; int w, h;
    rasterise_walls(g, cfg.cell_size_px, occ, w, h);
    return set_map(occ.data(), w, h, (float)(cfg.cell_meters / cfg.cell_size_px), 0.0, 0.0);   // RV:274,420-426
}

// precomputeRayDirections (MC:1017-1023): keys are (int)(a*100) although lookups use whole degrees (Q9).
int Engine::precompute_ray_directions(double lo, double hi, double step) {
    CK(cudaSetDevice(cfg.device));
    if (!(step > 0)) return fail(MCL_ERR_ARG, "precompute_ray_directions: step must be > 0");
    for (double a = lo; a <= hi; a += step) {
        const int key = static_cast<int>(a * 100);
        const double rad = a * M_PI / 180.0;
        const int k = key - key_min;
        if (k < 0 || k >= n_keys) continue;          // never looked up
        h_lut[k] = make_double2(std::cos(rad), std::sin(rad));
        if (!h_lut_filled[k]) { h_lut_filled[k] = 1; --n_unfilled; }
    }
    CK(cudaMemcpyAsync(d_lut.p, h_lut.data(), n_keys * sizeof(double2), cudaMemcpyHostToDevice, stream));
    CK(cudaMemcpyAsync(d_lut_filled.p, h_lut_filled.data(), n_keys, cudaMemcpyHostToDevice, stream));
    CK(cudaStreamSynchronize(stream));
    return MCL_OK;
}

int Engine::get_ray_lut(int32_t* keys, double* dx, double* dy, int32_t cap, int32_t* count) {
    int c = 0;
    for (int k = 0; k < n_keys; ++k) {
        if (!h_lut_filled[k]) continue;
        if (c < cap && keys && dx && dy) { keys[c] = key_min + k; dx[c] = h_lut[k].x; dy[c] = h_lut[k].y; }
        ++c;
    }
    if (count) *count = c;
    return MCL_OK;
}

// ---- particles ----------------------------------------------------------------------------------------------
int Engine::ensure_particles(int64_t count) {
    if (count <= 0) return fail(MCL_ERR_ARG, "particle count must be > 0");
    if (cfg.max_particles > 0 && count > cfg.max_particles) return fail(MCL_ERR_ARG, "particle count exceeds max_particles");
    if (count > (int64_t)INT32_MAX) return fail(MCL_ERR_ARG, "particle count exceeds 2^31-1 per GPU");
    CK(part[0].ensure(count)); CK(part[1].ensure(count));
    CK(ancestors.ensure(count));
    if (cfg.mode == MCL_MODE_NS) {
        if (shard_world == 1 && count >= (1ll << 28)) return fail(MCL_ERR_ARG, "NS filters hold fewer than 2^28 particles (Q32 totals stay below 2^60)");
        CK(d_ll.ensure(count)); CK(d_prefix.ensure(count));
        CK(d_tile_sums.ensure((size_t)(count + 4095) / 4096 + (size_t)(count + 4095) / 4096 / 64 + 4)); CK(d_u64.ensure(4)); CK(d_maxbits.ensure(1));
        if (shard_world == 1) { n_global = count; shard_begin = 0; per_rank = count; }
        return MCL_OK;
    }
    CK(cdf.ensure(count));
    CK(d_wraw.ensure(count)); CK(d_wn.ensure(count));
    return ensure_xs(count);
}

int Engine::ensure_xs(int64_t count) {
    const int nt = (int)((count + xs::XS_TILE - 1) / xs::XS_TILE);
    CK(xs_flag.ensure(2));
    CK(cudaMemsetAsync(xs_flag.p, 0, 2 * sizeof(int), stream));      // [0] fallback flag, [1] ticket of k_xs_tilesum
    if (nt <= xs_tiles_cap) return MCL_OK;
    CK(xs_tsum.ensure(nt)); CK(xs_toff.ensure(nt + 1)); CK(xs_seq_s.ensure((size_t)nt * xs::XS_SEQ_CAP));
    CK(xs_tiles.ensure((size_t)nt * sizeof(xs::TileSummary))); CK(xs_entries.ensure((size_t)nt * xs::XS_SEQ_CAP * sizeof(xs::SeqEntry)));
    CK(xs_carry.ensure((size_t)(nt + 1) * sizeof(xs::Par))); CK(xs_seq_base.ensure(nt + 1));
    // the one-kernel form's published words: EMPTY once; every launch's last block leaves them EMPTY again
    CK(xs_pub.ensure((size_t)nt * xs::XSF_PSTRIDE)); CK(xs_counters.ensure(4)); CK(xs_blocks.ensure((size_t)nt * xs::XSF_BSTRIDE * sizeof(xs::SeqBlock)));
    {   // word 0 of every line EMPTY (all ones), the summary words 0
        std::vector<unsigned long long> init((size_t)nt * xs::XSF_PSTRIDE, 0ull);
        for (int k = 0; k < nt; ++k) init[(size_t)k * xs::XSF_PSTRIDE] = ~0ull;
        CK(cudaMemcpyAsync(xs_pub.p, init.data(), init.size() * sizeof(unsigned long long), cudaMemcpyHostToDevice, stream));
        CK(cudaStreamSynchronize(stream));
    }
    CK(cudaMemsetAsync(xs_counters.p, 0, 4 * sizeof(unsigned), stream));
    xs_epoch = 0;
    xs_tiles_cap = nt;
    return MCL_OK;
}

// The reference's sequential f64 accumulations, reproduced bit-exactly in parallel (exact_scan.cuh):
//   normalise == false: *d_total_out = w_0 + w_1 + ... left to right over d_wraw            (MC:675)
//   normalise == true : w_i <- (float)((double)w_i / total) into d_wn, then cdf[i]            (MC:496-505)
int Engine::exact_accumulate(bool normalise, double* d_total_out, const EmaArgs* ema, int guide_buckets_wanted) {
    return exact_accumulate_on(d_wraw.p, normalise, normalise, d_total_out, ema, guide_buckets_wanted);
}

// w: the dense fp32 terms. normalise: they are divided by the total in d_scalars[0] first (the quotients are what is
// accumulated; the multi-launch form also leaves them in d_wn). ema (mcl_step, with the total): the accumulation's last block
// also advances the adaptive-injection state on the device; ema_in_total tells the caller whether that happened.
int Engine::exact_accumulate_on(const float* w, bool normalise, bool want_cdf, double* d_total_out, const EmaArgs* ema, int guide_buckets_wanted) {
    const int nt = (int)((n + xs::XS_TILE - 1) / xs::XS_TILE);
    const int ntf = (int)((n + xs::XSF_TILE - 1) / xs::XSF_TILE);          // the one-kernel form's tiles (never more than nt)
    ema_in_total = false;
    guide_in_cdf = false;
    if (!force_sequential && !force_multilaunch_scan && ntf <= xs::XSF_MAX_TILES) {
        xs::FusedWs fw;
        fw.pub = xs_pub.p; fw.blocks = (xs::SeqBlock*)xs_blocks.p; fw.counters = xs_counters.p;
        fw.trace = nullptr;
        fw.abort = tick_abort;
        if (xs_resident_tiles < 0) {           // how many tiles the device holds at once: up to there a tile can be its block index
            int per_sm_a = 0, per_sm_b = 0, sms = 0;
            CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_a, xs::k_xs_fused<true>, xs::XS_THREADS, 0));
            CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_b, xs::k_xs_fused<false>, xs::XS_THREADS, 0));
            CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, cfg.device));
            xs_resident_tiles = std::min(per_sm_a, per_sm_b) * sms;
        }
        fw.by_index = (ntf <= xs_resident_tiles && !force_scan_tickets) ? 1 : 0;
        if (xs_trace_on) { CK(xs_trace.ensure((size_t)nt * 16)); CK(cudaMemsetAsync(xs_trace.p, 0, (size_t)nt * 16 * sizeof(unsigned long long), stream)); fw.trace = xs_trace.p; }      // (nt >= ntf)
        xs::FusedEma fe;
        fe.inj = nullptr; fe.counters = nullptr; fe.n = (double)n; fe.a_slow = 0; fe.a_fast = 0;
        if (ema && !want_cdf) { fe.inj = d_inj.p; fe.counters = d_counters.p; fe.a_slow = ema->a_slow; fe.a_fast = ema->a_fast; ema_in_total = true; }
        xs::FusedGuide fg;
        fg.table = nullptr; fg.buckets = 0; fg.log2_buckets = 0; fg.force_fallback = force_scan_fallback ? 1 : 0;
        if (want_cdf && guide_buckets_wanted > 0 && !force_separate_guide) { fg.table = d_guide.p; fg.buckets = guide_buckets_wanted; fg.log2_buckets = __builtin_ctz((unsigned)guide_buckets_wanted); guide_in_cdf = true; }      // (d_guide sized by the caller)
        // One tile (the reference's own few thousand particles) inside mcl_step: total + adaptive-injection state + normalised
        // CDF in one launch; ref_resample_front then finds the CDF done.
        cdf_by_total = false;
        if (!want_cdf) inject_by_scans = false;
        if (!want_cdf && fuse_cdf_into_total && !force_two_scan_launches && ema && ntf == 1 && n < 4096 && !xs_trace_on && d_total_out == d_scalars.p) {
            xs_epoch += 2;
            CK(d_block_counts.ensure(xs::XSF_TILE / 256));
            const RefDrawGen G{(uint32_t)step_counter, (uint32_t)cfg.seed, (uint32_t)(cfg.seed >> 32)};       // (ref_resample's, which follows in this call)
            // (the fewer weights, the fewer per thread: the passes are chains of per-thread latencies)
            if (n <= 1024 && !force_scan_items16)
                LAUNCH_PDL(K_SCANS_ONE_TILE, k_ref_scans_one_tile<4>, 1, xs::XS_THREADS, 0, w, n, xs_epoch - 1, fw, cdf.p, d_total_out, fe, fg.force_fallback, G,
                           d_block_counts.p, d_counters.p + 2);
            else if (n <= 2048 && !force_scan_items16)
                LAUNCH_PDL(K_SCANS_ONE_TILE, k_ref_scans_one_tile<8>, 1, xs::XS_THREADS, 0, w, n, xs_epoch - 1, fw, cdf.p, d_total_out, fe, fg.force_fallback, G,
                           d_block_counts.p, d_counters.p + 2);
            else
                LAUNCH_PDL(K_SCANS_ONE_TILE, k_ref_scans_one_tile<16>, 1, xs::XS_THREADS, 0, w, n, xs_epoch - 1, fw, cdf.p, d_total_out, fe, fg.force_fallback, G,
                           d_block_counts.p, d_counters.p + 2);
            CK(cudaGetLastError());
            cdf_by_total = true; inject_by_scans = true;
            return MCL_OK;
        }
        ++xs_epoch;
        if (want_cdf)      // divisor: the total (normalise) or the constant 1.0 parked in d_scalars[7]
            LAUNCH_PDL(K_XS_CDF, xs::k_xs_fused<true>, ntf, xs::XS_THREADS, 0, w, n, ntf, xs_epoch, fw, (const double*)(normalise ? d_scalars.p : d_scalars.p + 7), cdf.p,
                       d_total_out, fe, fg);
        else
            LAUNCH_PDL(K_XS_TOTAL, xs::k_xs_fused<false>, ntf, xs::XS_THREADS, 0, w, n, ntf, xs_epoch, fw, (const double*)nullptr, (double*)nullptr, d_total_out, fe, fg);
        CK(cudaGetLastError());
        return MCL_OK;
    }
    xs::Workspace ws;
    ws.tsum = xs_tsum.p; ws.toff = xs_toff.p; ws.tiles = (xs::TileSummary*)xs_tiles.p; ws.entries = (xs::SeqEntry*)xs_entries.p;
    ws.carry = (xs::Par*)xs_carry.p; ws.seq_base = xs_seq_base.p; ws.seq_s = xs_seq_s.p; ws.flag = xs_flag.p;
    const float* wa = normalise ? d_wn.p : w;                  // what the multi-launch form accumulates
    if (!force_sequential) {
        if (normalise)
            LAUNCH_PDL(K_XS_TILESUM, xs::k_xs_tilesum<true>, nt, xs::XS_THREADS, 0, w, d_wn.p, part[cur].p, n, d_scalars.p, xs_tsum.p, xs_toff.p, xs_flag.p);
        else
            LAUNCH_PDL(K_XS_TILESUM, xs::k_xs_tilesum<false>, nt, xs::XS_THREADS, 0, w, (float*)nullptr, (float4*)nullptr, n, (const double*)nullptr,
                   xs_tsum.p, xs_toff.p, xs_flag.p);
        LAUNCH_PDL(K_XS_SCAN, xs::k_xs_scan<false>, nt, xs::XS_THREADS, 0, wa, n, nt, ws, (double*)nullptr);
        LAUNCH_PDL(K_XS_CHAIN, xs::k_xs_chain, 1, xs::XS_CHAIN_THREADS, 0, nt, ws, d_total_out, wa, n);     // total: falls back in-kernel
        if (want_cdf) {
            LAUNCH_PDL(K_XS_APPLY, xs::k_xs_scan<true>, nt, xs::XS_THREADS, 0, wa, n, nt, ws, cdf.p);
            LAUNCH_PDL(K_SEQ_CDF, k_ref_seq_cdf, 1, 256, 0, wa, n, cdf.p, (const int*)xs_flag.p);            // runs only if flagged
        }
    } else {
        if (normalise)
            LAUNCH_PDL(K_XS_TILESUM, xs::k_xs_tilesum<true>, nt, xs::XS_THREADS, 0, w, d_wn.p, part[cur].p, n, d_scalars.p, xs_tsum.p, xs_toff.p, xs_flag.p);
        if (want_cdf) LAUNCH_PDL(K_SEQ_CDF, k_ref_seq_cdf, 1, 256, 0, wa, n, cdf.p, (const int*)nullptr);
        if (d_total_out) LAUNCH(K_SEQ_TOTAL, k_ref_seq_total, 1, 256, 0, wa, n, d_total_out, (const int*)nullptr);
    }
    CK(cudaGetLastError());
    return MCL_OK;
}

// Debug/test entry: the exact accumulation of an arbitrary fp32 vector (no normalisation).
int Engine::debug_exact_scan(const float* w, int64_t count, double* cdf_out, double* total_out, int* fell_back) {
    CK(cudaSetDevice(cfg.device));
    if (!w || count <= 0) return fail(MCL_ERR_ARG, "debug_exact_scan: bad input");
    int rc = ensure_particles(count);
    if (rc) return rc;
    const int64_t saved_n = n;
    n = count;
    CK(cudaMemcpyAsync(d_wn.p, w, (size_t)count * sizeof(float), cudaMemcpyHostToDevice, stream));
    CK(cudaMemsetAsync(xs_flag.p, 0, sizeof(int), stream));
    rc = exact_accumulate_on(d_wn.p, false, true, d_scalars.p + 6);
    if (rc) { n = saved_n; return rc; }
    int flag = 0;
    unsigned fused_flag = 0;
    if (cdf_out) CK(cudaMemcpyAsync(cdf_out, cdf.p, (size_t)count * sizeof(double), cudaMemcpyDeviceToHost, stream));
    if (total_out) CK(cudaMemcpyAsync(total_out, d_scalars.p + 6, sizeof(double), cudaMemcpyDeviceToHost, stream));
    CK(cudaMemcpyAsync(&flag, xs_flag.p, sizeof(int), cudaMemcpyDeviceToHost, stream));
    CK(cudaMemcpyAsync(&fused_flag, xs_counters.p + 2, sizeof(unsigned), cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
    const bool fused = !force_sequential && !force_multilaunch_scan && (count + xs::XSF_TILE - 1) / xs::XSF_TILE <= xs::XSF_MAX_TILES;
    if (fused) flag = (xs_epoch != 0 && fused_flag == xs_epoch) ? 1 : 0;
    if (fell_back) *fell_back = force_sequential ? 1 : flag;
    n = saved_n;
    have_weights = false;
    return MCL_OK;
}

// Stage time stamps (%globaltimer, ns) of every tile of the one-kernel exact scan, for the NEXT mcl_debug_exact_scan calls
// (out == null: switch tracing on) or of the last one (out != null: [n_tiles][8], stamps 0..6 used).
int Engine::debug_exact_scan_trace(unsigned long long* out, int64_t cap_tiles, int* n_tiles) {
    CK(cudaSetDevice(cfg.device));
    if (!out) { xs_trace_on = cap_tiles != 0; return MCL_OK; }
    const int64_t have = (int64_t)(xs_trace.n / 16);
    const int64_t cnt = std::min(cap_tiles, have);
    if (n_tiles) *n_tiles = (int)have;
    if (cnt > 0) {
        CK(cudaMemcpyAsync(out, xs_trace.p, (size_t)cnt * 16 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, stream));
        CK(cudaStreamSynchronize(stream));
    }
    return MCL_OK;
}

int Engine::debug_last_scan_fell_back(int* fell_back) {
    CK(cudaSetDevice(cfg.device));
    if (!fell_back) return fail(MCL_ERR_ARG, "debug_last_scan_fell_back: null pointer");
    unsigned flag = 0;
    *fell_back = 0;
    if (!xs_counters.p || xs_epoch == 0) return MCL_OK;
    CK(cudaMemcpyAsync(&flag, xs_counters.p + 2, sizeof(unsigned), cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
    // a tick uses two consecutive epochs (total, CDF): either of them flagged counts
    *fell_back = (flag == xs_epoch || flag == xs_epoch - 1) ? 1 : 0;
    return MCL_OK;
}

int Engine::debug_trigf(const float* x, int64_t count, float* s_out, float* c_out, int* kind_out) {
    CK(cudaSetDevice(cfg.device));
    if (!x || !s_out || !c_out || count <= 0) return fail(MCL_ERR_ARG, "debug_trigf: bad argument");
    DevBuf<float> d;
    CK(d.ensure(3 * (size_t)count));
    CK(cudaMemcpyAsync(d.p, x, (size_t)count * sizeof(float), cudaMemcpyHostToDevice, stream));
    LAUNCH(K_PREDICT, k_debug_trigf, grid_for(count, 256), 256, 0, d.p, count, trig_kind, d.p + count, d.p + 2 * count);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(s_out, d.p + count, (size_t)count * sizeof(float), cudaMemcpyDeviceToHost, stream));
    CK(cudaMemcpyAsync(c_out, d.p + 2 * count, (size_t)count * sizeof(float), cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
    d.release();
    if (kind_out) *kind_out = trig_kind;
    return MCL_OK;
}

void Engine::philox_host(uint32_t stream_id, uint64_t index, uint32_t out[4]) const {
    uint32_t o[4];
    Philox::gen((uint32_t)index, (uint32_t)(index >> 32), stream_id, (uint32_t)step_counter, (uint32_t)cfg.seed, (uint32_t)(cfg.seed >> 32), o);
    for (int i = 0; i < 4; i++) out[i] = o[i];
}

static RefResampleParams make_resample_params(const mcl_config& c, int64_t n, int jitter_state, double p_inject) {
    RefResampleParams R;
    R.p_inject = p_inject;
    R.max_inject = (int)(jitter_state ? c.inject_max_lost : c.inject_max_conf);
    R.jitter_state = jitter_state;
    const double jxy = jitter_state ? c.jitter_xy_lost : c.jitter_xy_conf;
    R.jit_xy_a = -jxy; R.jit_xy_w = jxy - (-jxy);
    R.jit_th_a = -c.jitter_theta_lost; R.jit_th_w = c.jitter_theta_lost - (-c.jitter_theta_lost);
    R.new_weight = (float)(1.0 / (double)n);
    R.cell_meters = c.cell_meters;
    R.half_cell = 0.5 * c.cell_meters;
    R.init_a = -c.init_offset; R.init_w = c.init_offset - (-c.init_offset);
    R.yaw_a = -M_PI; R.yaw_w = M_PI - (-M_PI);
    R.init_shift = c.init_shift;
    R.inj_rows = 1; R.inj_cols = 1;       // set by the caller from the map (MC:423-424)
    return R;
}

int Engine::init(int64_t count, const mcl_init_draws* d) {
    CK(cudaSetDevice(cfg.device));
    if (!map_ready) return fail(MCL_ERR_ARG, "init: set a map first (sampleParticles reads the grid size, MC:423-424)");
    if (cfg.mode == MCL_MODE_NS) return ns_init(count);
    int rc = ensure_particles(count);
    if (rc) return rc;
    const int n_cols = (int)((unsigned)map_w / (unsigned)cfg.cell_size_px);    // MC:423
    const int n_rows = (int)((unsigned)map_h / (unsigned)cfg.cell_size_px);    // MC:424
    if (n_cols <= 0 || n_rows <= 0) return fail(MCL_ERR_ARG, "init: map smaller than one cell");
    // stage the five named draws: [u_yaw | u_dx | u_dy] f64 and [row | col] i32
    const size_t f64_bytes = 3 * (size_t)count * sizeof(double), i32_bytes = 2 * (size_t)count * sizeof(int);
    rc = ensure_pinned(f64_bytes + i32_bytes);
    if (rc) return rc;
    double* hf = (double*)h_pinned;
    int* hi = (int*)((char*)h_pinned + f64_bytes);
    if (d) {
        if (!d->u_yaw || !d->row || !d->col || !d->u_dx || !d->u_dy) return fail(MCL_ERR_ARG, "init: null draw array");
        memcpy(hf, d->u_yaw, count * sizeof(double));
        memcpy(hf + count, d->u_dx, count * sizeof(double));
        memcpy(hf + 2 * count, d->u_dy, count * sizeof(double));
        memcpy(hi, d->row, count * sizeof(int));
        memcpy(hi + count, d->col, count * sizeof(int));
    } else {
        for (int64_t i = 0; i < count; ++i) {
            uint32_t a[4], b[4];
            philox_host(0x10, 2 * (uint64_t)i, a);
            philox_host(0x10, 2 * (uint64_t)i + 1, b);
            hf[i] = canonical53(a[0], a[1]);
            hf[count + i] = canonical53(a[2], a[3]);
            hf[2 * count + i] = canonical53(b[0], b[1]);
            hi[i] = (int)(b[2] % (uint32_t)n_rows);
            hi[count + i] = (int)(b[3] % (uint32_t)n_cols);
        }
    }
    CK(d_u_jit.ensure(3 * (size_t)count));
    CK(d_u_r.ensure((size_t)count));
    // reuse d_u_jit as the f64 staging area and ancestors/cdf space for the ints
    CK(cudaMemcpyAsync(d_u_jit.p, hf, f64_bytes, cudaMemcpyHostToDevice, stream));
    int* d_rc = (int*)cdf.p;      // cdf holds count doubles = 2*count ints
    CK(cudaMemcpyAsync(d_rc, hi, i32_bytes, cudaMemcpyHostToDevice, stream));
    RefResampleParams R = make_resample_params(cfg, count, 1, 0.0);
    LAUNCH(K_INIT, k_ref_init, grid_for(count, 256), 256, 0, part[cur].p, count, d_u_jit.p, d_rc, d_rc + count, d_u_jit.p + count,
           d_u_jit.p + 2 * count, R);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(stream));
    n = count;
    have_weights = false;
    wsum_known = true; known_wsum = (double)count;                 // every weight is 1 (MC:445)
    return MCL_OK;
}

int Engine::upload(const float* p, int64_t count) {
    CK(cudaSetDevice(cfg.device));
    if (!p) return fail(MCL_ERR_ARG, "upload: null pointer");
    int rc = ensure_particles(count);
    if (rc) return rc;
    CK(cudaMemcpyAsync(part[cur].p, p, (size_t)count * sizeof(float4), cudaMemcpyHostToDevice, stream));
    CK(cudaStreamSynchronize(stream));
    n = count;
    have_weights = false;
    wsum_known = false;
    return MCL_OK;
}

int Engine::download(float* p) {
    CK(cudaSetDevice(cfg.device));
    if (!p) return fail(MCL_ERR_ARG, "download: null pointer");
    if (n == 0) return fail(MCL_ERR_ARG, "download: no particles");
    { int rc = ns_materialise_weights(); if (rc) return rc; }
    CK(cudaMemcpyAsync(p, part[cur].p, (size_t)n * sizeof(float4), cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
    return MCL_OK;
}

int Engine::download_ancestors(int32_t* idx) {
    CK(cudaSetDevice(cfg.device));
    if (!idx || n == 0) return fail(MCL_ERR_ARG, "download_ancestors: nothing to download");
    CK(cudaMemcpyAsync(idx, ancestors.p, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
    return MCL_OK;
}

int Engine::download_resample_draws(double* u_r, double* u_jit) {
    CK(cudaSetDevice(cfg.device));
    if (!u_r || !u_jit || n == 0) return fail(MCL_ERR_ARG, "download_resample_draws: nothing to download");
    if (draws_generated) {        // the last resample generated its draws inside the kernels: materialise the same streams
        CK(d_u_r.ensure((size_t)n)); CK(d_u_jit.ensure((size_t)n * 3));
        LAUNCH(K_FILL_DRAWS, k_fill_resample_draws, grid_for(n, 256), 256, 0, d_u_r.p, d_u_jit.p, n, last_per, draws_step, (uint32_t)cfg.seed,
               (uint32_t)(cfg.seed >> 32));
        CK(cudaGetLastError());
    }
    CK(cudaMemcpyAsync(u_r, d_u_r.p, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, stream));
    CK(cudaMemcpyAsync(u_jit, d_u_jit.p, (size_t)n * last_per * sizeof(double), cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
    return MCL_OK;
}

int Engine::download_cdf(double* out) {
    CK(cudaSetDevice(cfg.device));
    if (!out || n == 0) return fail(MCL_ERR_ARG, "download_cdf: nothing to download");
    CK(cudaMemcpyAsync(out, cdf.p, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
    return MCL_OK;
}

// ---- predict --------------------------------------------------------------------------------------------------
int Engine::predict_motion(double r1, double t, double r2) {
    CK(cudaSetDevice(cfg.device));
    if (n == 0) return fail(MCL_ERR_ARG, "predict: no particles");
    if (cfg.mode == MCL_MODE_NS) { Motion m; m.rot_1 = r1; m.trans = t; m.rot_2 = r2; return ns_predict(m); }
    // Eigen narrows the f64 scalars to the array's fp32 first (MC:746-753)
    if (pending_motion.valid) { int rc = flush_pending_motion(); if (rc) return rc; }
    if (defer_predict) {           // mcl_step: the computeWeight kernel that follows applies it while it loads the particles
        pending_motion.valid = true; pending_motion.rot1 = (float)r1; pending_motion.trans = (float)t; pending_motion.dtheta = (float)(r1 + r2);
        have_weights = false;
        return MCL_OK;
    }
    LAUNCH_PDL(K_PREDICT, k_ref_predict, grid_for(n, 256), 256, 0, part[cur].p, n, (float)r1, (float)t, (float)(r1 + r2), trig_kind);
    CK(cudaGetLastError());
    have_weights = false;
    return MCL_OK;
}

// A motion that no computeWeight kernel took along (first-touch pre-pass needed, per-particle kernel, ...): its own pass.
int Engine::flush_pending_motion() {
    if (!pending_motion.valid) return MCL_OK;
    pending_motion.valid = false;
    LAUNCH_PDL(K_PREDICT, k_ref_predict, grid_for(n, 256), 256, 0, part[cur].p, n, pending_motion.rot1, pending_motion.trans, pending_motion.dtheta, trig_kind);
    CK(cudaGetLastError());
    return MCL_OK;
}

int Engine::predict_encoders(double enc_l, double enc_r, const double* z3, double* motion_out) {
    double z[3];
    if (z3) { z[0] = z3[0]; z[1] = z3[1]; z[2] = z3[2]; }
    else {
        // Box-Muller on host Philox draws (production path; parity runs inject z3)
        uint32_t a[4], b[4];
        philox_host(0x20, 0, a); philox_host(0x20, 1, b);
        const double u1 = 1.0 - canonical53(a[0], a[1]), u2 = canonical53(a[2], a[3]);
        const double u3 = 1.0 - canonical53(b[0], b[1]), u4 = canonical53(b[2], b[3]);
        z[0] = std::sqrt(-2.0 * std::log(u1)) * std::cos(2 * M_PI * u2);
        z[1] = std::sqrt(-2.0 * std::log(u1)) * std::sin(2 * M_PI * u2);
        z[2] = std::sqrt(-2.0 * std::log(u3)) * std::cos(2 * M_PI * u4);
    }
    if (cfg.mode == MCL_MODE_NS) {
        // per-particle noise: the odometry increment itself stays clean (zero draws), the kernel adds Philox noise
        const double zero[3] = {0.0, 0.0, 0.0};
        Motion m = odometry_step(odo, cfg, enc_l, enc_r, zero);
        if (motion_out) { motion_out[0] = m.rot_1; motion_out[1] = m.trans; motion_out[2] = m.rot_2; }
        return ns_predict(m);
    }
    Motion m = odometry_step(odo, cfg, enc_l, enc_r, z);
    if (motion_out) { motion_out[0] = m.rot_1; motion_out[1] = m.trans; motion_out[2] = m.rot_2; }
    ++step_counter;
    return predict_motion(m.rot_1, m.trans, m.rot_2);
}

// ---- update -----------------------------------------------------------------------------------------------------
int Engine::update(const float* ranges, int n_beams, float angle_min, float angle_inc, float range_min, float range_max, double* total) {
    CK(cudaSetDevice(cfg.device));
    if (!map_ready) return fail(MCL_ERR_ARG, "update: no map (the reference warns 'NO MAP RECEVIED', MC:311)");
    if (n == 0) return fail(MCL_ERR_ARG, "update: no particles");
    if (n_beams < 0 || (n_beams > 0 && !ranges)) return fail(MCL_ERR_ARG, "update: bad scan");
    if (cfg.mode == MCL_MODE_NS) {
        if (shard_world != 1) return fail(MCL_ERR_STATE, "update: sharded NS filters are driven through the mcl_ns_*_local phases");
        float local_max = 0.f;
        int rc = ns_update_local(ranges, n_beams, angle_min, angle_inc, range_min, range_max, &local_max);
        if (rc) return rc;
        uint64_t tot = 0;
        rc = ns_weights_local(local_max, &tot);
        if (rc) return rc;
        if (total) *total = (double)tot * 2.3283064365386963e-10;
        return MCL_OK;
    }
    std::vector<RefBeam> used;
    int rc = ref_prepare_beams(ranges, n_beams, angle_min, angle_inc, range_min, range_max, beams_all, used);
    if (rc) return rc;
    return ref_run_update(nullptr, used.data(), (int)used.size(), beams_all, total);
}

int Engine::stage_scan(int slot, const float* ranges, int n_beams, float angle_min, float angle_inc, float range_min, float range_max) {
    CK(cudaSetDevice(cfg.device));
    if (slot < 0 || slot >= 4096) return fail(MCL_ERR_ARG, "stage_scan: slot out of range [0,4096)");
    if (n_beams < 0 || (n_beams > 0 && !ranges)) return fail(MCL_ERR_ARG, "stage_scan: bad scan");
    if (cfg.mode == MCL_MODE_NS) return ns_stage_scan(slot, ranges, n_beams, angle_min, angle_inc, range_min, range_max);
    if ((size_t)slot >= staged.size()) staged.resize(slot + 1);
    StagedScan& s = staged[slot];
    std::vector<RefBeam> used;
    int rc = ref_prepare_beams(ranges, n_beams, angle_min, angle_inc, range_min, range_max, s.all, used);
    if (rc) return rc;
    // Ticks queued by mcl_step_staged may still be reading this slot (or the buffer ensure() is about to free): the engine's
    // stream is non-blocking, so a plain cudaMemcpy would not be ordered behind them. Drain the stream, then copy on it.
    CK(cudaStreamSynchronize(stream));
    CK(s.d_used.ensure(std::max<size_t>(1, used.size())));
    if (!used.empty()) {
        CK(cudaMemcpyAsync(s.d_used.p, used.data(), used.size() * sizeof(RefBeam), cudaMemcpyHostToDevice, stream));
        CK(cudaStreamSynchronize(stream));
    }
    s.n_used = (int)used.size();
    s.valid = true;
    return MCL_OK;
}

int Engine::update_staged(int slot, double* total) {
    CK(cudaSetDevice(cfg.device));
    if (!map_ready) return fail(MCL_ERR_ARG, "update: no map");
    if (n == 0) return fail(MCL_ERR_ARG, "update: no particles");
    if (cfg.mode == MCL_MODE_NS) {
        if (shard_world != 1) return fail(MCL_ERR_STATE, "update_staged: sharded NS filters use mcl_ns_update_local_staged");
        float local_max = 0.f;
        int rc = ns_update_local_staged(slot, &local_max);
        if (rc) return rc;
        uint64_t tot = 0;
        rc = ns_weights_local(local_max, &tot);
        if (rc) return rc;
        if (total) *total = (double)tot * 2.3283064365386963e-10;
        return MCL_OK;
    }
    if (slot < 0 || (size_t)slot >= staged.size() || !staged[slot].valid) return fail(MCL_ERR_ARG, "update_staged: empty slot");
    return ref_run_update(staged[slot].d_used.p, nullptr, staged[slot].n_used, staged[slot].all, total);
}

static size_t ref_smem_bytes(int n_keys, int n_beams, int n_radii, size_t map_bytes) {
    return (size_t)n_keys * sizeof(double2) + (size_t)n_beams * sizeof(RefBeam) + (size_t)n_radii * sizeof(double) + map_bytes;
}

// filterLaserReadings + filterAngles (MC:635), then the stride-picked beams in f64 (MC:650-669).
int Engine::ref_prepare_beams(const float* ranges, int n_beams, float angle_min, float angle_inc, float range_min, float range_max,
                              std::vector<HostBeam>& all, std::vector<RefBeam>& used) {
    filter_scan(ranges, n_beams, angle_min, angle_inc, range_min, range_max, true, cfg.fov_lower_deg, cfg.fov_upper_deg, all);
    const int stride = std::max(1, cfg.beam_stride);
    used.clear();
    for (size_t i = 0; i < all.size(); i += stride) {                             // MC:650
        RefBeam b;
        b.off_deg = -(all[i].angle) * 180.0 / M_PI;                               // MC:653
        b.obs = all[i].radius;                                                    // MC:657
        b.rand_term = cfg.w_rand * ((std::abs(b.obs - cfg.max_laser_range) < 0.01) ? 1.0 : 0.0);   // MC:669
        used.push_back(b);
    }
    return MCL_OK;
}

// d_used: the scored beams in device memory (a staged scan), or null with h_used: the beams of a scan that has just arrived
// from the host. Those ride in k_ref_update_v2's launch parameters when they fit (RU_INLINE_BEAMS); the kernels that read
// them through a pointer (first-touch pre-pass, the per-particle kernel) and longer lists get a copy through the pinned ring.
int Engine::ref_run_update(const RefBeam* d_used, const RefBeam* h_used, int n_used, const std::vector<HostBeam>& all, double* total, bool defer_sync, const EmaArgs* ema) {
    RefParams P;
    auto beams_on_device = [&]() -> int {
        if (d_used || n_used == 0) return MCL_OK;
        CK(d_beams.ensure((size_t)n_used));
        // a ring of pinned slots: earlier ticks may still be in flight, their scans must not be overwritten
        int rc = ensure_pinned_ring((size_t)n_used * sizeof(RefBeam));
        if (rc) return rc;
        void* hp = pinned_ring_next();
        memcpy(hp, h_used, (size_t)n_used * sizeof(RefBeam));
        CK(cudaMemcpyAsync(d_beams.p, hp, (size_t)n_used * sizeof(RefBeam), cudaMemcpyHostToDevice, stream));
        CK(cudaEventRecord(ring_events[ring_pos], stream));
        d_used = d_beams.p;
        P.beams = d_used;
        return MCL_OK;
    };
    if (!d_used && n_used > 0 && !h_used) return fail(MCL_ERR_ARG, "update: no beams");
    if (!d_used && n_used > RU_INLINE_BEAMS) { int rc = beams_on_device(); if (rc) return rc; }
    if (!d_used)
        for (int i = 0; i < n_used; i++) P.inline_beams[i] = h_used[i];
    P.occ = d_occ.p; P.width = map_w; P.height = map_h;
    P.occ_pad = d_occ_pad.p; P.pad = occ_pad; P.wp = occ_wp;
    const size_t map_bytes = (size_t)map_w * map_h;
    const size_t pad_bytes = (size_t)occ_wp * (map_h + 2 * occ_pad);
    P.map_in_smem = map_bytes + pad_bytes <= 64 * 1024 ? 1 : 0;
    P.res = (double)res_f; P.inv_res = 1.0 / (double)res_f;
    P.ox = origin_x; P.oy = origin_y; P.max_x = max_x; P.max_y = max_y;
    P.laser_offset = cfg.laser_offset; P.validity_offset = cfg.validity_offset; P.max_range = cfg.max_laser_range;
    P.w_hit = cfg.w_hit;
    P.radii = d_radii.p; P.n_radii = (int)h_radii.size();
    P.gauss = d_gauss.p; P.gauss_size = (int)gauss.v.size(); P.gauss_res = gauss.step; P.gauss_min = gauss.lo; P.gauss_max = gauss.hi;
    P.lut = d_lut.p; P.lut_filled = d_lut_filled.p; P.key_min = key_min; P.n_keys = n_keys;
    P.beams = d_used; P.n_beams = n_used;
    P.trig = trig_kind;
    P.do_predict = 0; P.rot1 = P.trans = P.dtheta = 0.f;
    const size_t smem = ref_smem_bytes(n_keys, n_used, P.n_radii, P.map_in_smem ? map_bytes : 0);
    if (smem > 200 * 1024) return fail(MCL_ERR_ARG, "update: too many beams for the shared-memory staging area");
    if (!attr_set) {
        CK(cudaFuncSetAttribute(k_ref_update, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        CK(cudaFuncSetAttribute(k_ref_first_touch, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr_set = true;
    }
    // First-touch memoisation of still-missing ray directions (Q9). Only keys round(yaw_deg + off_deg) with yaw in
    // [-180,180] and off among this scan's used beams can be touched; once those are all filled the pre-pass is skipped.
    bool need_prepass = false;
    if (n_unfilled > 0 && n_used > 0) {
        const int stride = std::max(1, cfg.beam_stride);
        double off_lo = 1e300, off_hi = -1e300;
        for (size_t i = 0; i < all.size(); i += stride) {
            const double off = -(all[i].angle) * 180.0 / M_PI;
            off_lo = std::min(off_lo, off); off_hi = std::max(off_hi, off);
        }
        // yaw_deg lies in [-180, 180] up to an ulp; rounding is monotone, so these are the extreme reachable keys
        const int k_lo = std::max(0, (int)std::round(-180.000001 + off_lo) - key_min);
        const int k_hi = std::min(n_keys - 1, (int)std::round(180.000001 + off_hi) - key_min);
        for (int k = k_lo; k <= k_hi && !need_prepass; ++k) need_prepass = !h_lut_filled[k];
    }
    // fp32 pre-filter tolerance: 4x the bound 2^-23*(cells + 2*max_range/res + 2) on the fp32 cell coordinate's error
    const double span = (double)std::max(map_w, map_h) + 2.0 * cfg.max_laser_range / (double)res_f + 2.0;
    const float tol32 = (float)(span * 4.76837158203125e-07);                  // 2^-21
    const bool fast32 = !force_f64_probe && span < 2.0e6 && tol32 < 0.05f && occ_pad > 0 && P.n_radii <= 16;
    const size_t smem2 = ru_smem_bytes(n_keys, n_used, P.n_radii, P.map_in_smem ? map_bytes : 0, P.map_in_smem ? pad_bytes : 0, !fast32);
    const bool bounded_ok = ((double)std::max(map_w, map_h) + cfg.max_laser_range / (double)res_f + 16.0) < 1.0e9;
    const bool use_v2 = !force_v1_update && n_used > 0 && smem2 <= 200 * 1024 && bounded_ok;
    P.abort = tick_abort;                      // (null outside an optimistic mcl_step tick)
    tick_optimistic = false;
    if (need_prepass && tick_abort != nullptr && use_v2 && scans_are_fused()) {
        // Optimistic tick: with a few thousand particles the table never fills (only the keys some particle's heading reaches
        // are ever touched), so EVERY tick would pay the pre-pass and its round trip through the host, although after the
        // first ticks it almost never finds a key. Here the pre-pass only reports (zero-copy, into pinned memory) and raises
        // *abort if it found something; the tick's kernels are enqueued behind it at once and return immediately if the flag
        // is up, in which case ref_step evaluates the directions and runs the tick again. Nothing of the tick is written
        // before that decision: the pre-pass applies the pending motion to the particles it looks at without storing it.
        if (pending_motion.valid) { P.do_predict = 1; P.rot1 = pending_motion.rot1; P.trans = pending_motion.trans; P.dtheta = pending_motion.dtheta; }
        { int rc = ensure_touch_block(); if (rc) return rc; }
        if (!touch_clean) { CK(cudaMemsetAsync(d_touch.p, 0xFF, n_keys * sizeof(unsigned long long), stream)); touch_clean = true; }
        LAUNCH(K_FIRST_TOUCH, k_ref_first_touch, grid_for(n, 256), 256, smem, part[cur].p, n, P, d_touch.p);
        CK(cudaGetLastError());
        LAUNCH(K_TOUCH_THETA, k_ref_touch_report, 1, 1024, 0, part[cur].p, d_touch.p, n_keys, P.do_predict, P.dtheta, h_touch_report, h_touch_keys, h_touch_theta,
               ++touch_seq, d_counters.p + 6);
        CK(cudaGetLastError());
        tick_optimistic = true;
    } else if (need_prepass) {
        { int rc = beams_on_device(); if (rc) return rc; }
        { int rc = flush_pending_motion(); if (rc) return rc; }           // the pre-pass looks at the predicted particles
        CK(cudaMemsetAsync(d_touch.p, 0xFF, n_keys * sizeof(unsigned long long), stream));
        touch_clean = false;
        LAUNCH(K_FIRST_TOUCH, k_ref_first_touch, grid_for(n, 256), 256, smem, part[cur].p, n, P, d_touch.p);
        CK(cudaGetLastError());
        LAUNCH(K_TOUCH_THETA, k_ref_touch_theta, grid_for(n_keys, 256), 256, 0, part[cur].p, d_touch.p, n_keys, d_touch_theta.p);
        CK(cudaGetLastError());
        int rc = ref_fill_ray_lut(all);
        if (rc) return rc;
    }
    if (use_v2) {
        if (pending_motion.valid) {
            P.do_predict = 1; P.rot1 = pending_motion.rot1; P.trans = pending_motion.trans; P.dtheta = pending_motion.dtheta;
            pending_motion.valid = false;
        }
        const bool ms = P.map_in_smem != 0;
        // instantiation = (zero origin, fp32 march, compile-time ray steps, map in shared memory)
#define RU_FOR_ALL(X) X(true, true, 11, true) X(false, true, 11, true) X(true, true, 0, true) X(false, true, 0, true) X(true, false, 0, true) X(false, false, 0, true) \
                      X(true, true, 11, false) X(false, true, 11, false) X(true, true, 0, false) X(false, true, 0, false) X(true, false, 0, false) X(false, false, 0, false)
        if (!attr_set2) {
#define X(Z, F, N, M) CK(cudaFuncSetAttribute((k_ref_update_v2<Z, F, N, M>), cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            RU_FOR_ALL(X)
#undef X
            attr_set2 = true;
        }
        int occ_blocks = 1, sms = 148;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, cfg.device);
        // the bounded fast path needs every probe quotient below 2^31: particle inside the map, ray at most max_range long
        const bool zero_origin = origin_x == 0.0 && origin_y == 0.0;
        const int nr_ct = (fast32 && P.n_radii == 11) ? 11 : 0;
        // particles per block and pass: a whole tile normally; a small filter is cut finer so that its rays reach more SMs
        // (1500 particles: 47 blocks of 32 instead of 6 of 256)
        int ppb = RU_TILE;
        while (ppb > 32 && (n + ppb - 1) / ppb < (int64_t)sms) ppb >>= 1;
        const int64_t tiles = (n + ppb - 1) / ppb;
        // ceil(2^32 / n_used); n_used == 1 would need 2^32 itself: the kernel divides by one without it
        const uint32_t div_magic = n_used == 1 ? 0u : (uint32_t)((0x100000000ull + (uint64_t)n_used - 1) / (uint64_t)n_used);
        bool launched = false;
#define X(Z, F, N, M)                                                                                                                    \
        if (!launched && zero_origin == Z && fast32 == F && nr_ct == N && ms == M) {                                                     \
            CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_blocks, (k_ref_update_v2<Z, F, N, M>), RU_TILE, smem2));               \
            const int grid = (int)std::min<int64_t>(tiles, (int64_t)sms * std::max(1, occ_blocks));                                      \
            LAUNCH_PDL(K_UPDATE_V2, (k_ref_update_v2<Z, F, N, M>), grid, RU_TILE, smem2, part[cur].p, d_wraw.p, n, P, div_magic, tol32, ppb); \
            launched = true;                                                                                                             \
        }
        RU_FOR_ALL(X)
#undef X
#undef RU_FOR_ALL
    } else {
        { int rc = beams_on_device(); if (rc) return rc; }
        { int rc = flush_pending_motion(); if (rc) return rc; }
        LAUNCH(K_UPDATE, k_ref_update, grid_for(n, 256), 256, smem, part[cur].p, d_wraw.p, n, P);
    }
    CK(cudaGetLastError());
    {
        int rc = exact_accumulate(false, d_scalars.p, ema);
        if (rc) return rc;
    }
    if (defer_sync) {            // mcl_step: the total stays on the device (k_ref_ema reads it there)
        have_weights = true;
        wsum_known = false;
        return MCL_OK;
    }
    CK(cudaMemcpyAsync(&last_total, d_scalars.p, sizeof(double), cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
    have_weights = true;
    wsum_known = true; known_wsum = last_total;                    // the f64 sum of the fp32 weights just written (MC:675)
    if (total) *total = last_total;
    return MCL_OK;
}

// Evaluate the direction of each newly touched key from its first toucher's theta, with host libm (MC:361-362).
int Engine::ref_fill_ray_lut(const std::vector<HostBeam>& all) {
    std::vector<unsigned long long> touch(n_keys);
    std::vector<float> theta(n_keys);
    CK(cudaMemcpyAsync(touch.data(), d_touch.p, n_keys * sizeof(unsigned long long), cudaMemcpyDeviceToHost, stream));
    CK(cudaMemcpyAsync(theta.data(), d_touch_theta.p, n_keys * sizeof(float), cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
    return ref_fill_ray_lut_from(all, touch.data(), theta.data());
}

// pinned block of an optimistic tick's pre-pass: {RefTouchReport, touch[n_keys], theta[n_keys]}
int Engine::ensure_touch_block() {
    if (h_touch_report && h_touch_cap >= n_keys) return MCL_OK;
    if (h_touch_report) cudaFreeHost(h_touch_report);
    h_touch_report = nullptr; h_touch_cap = 0;
    const size_t bytes = sizeof(RefTouchReport) + (size_t)n_keys * (sizeof(unsigned long long) + sizeof(float));
    void* p = nullptr;
    CK(cudaMallocHost(&p, bytes));
    memset(p, 0, bytes);
    h_touch_report = (RefTouchReport*)p;
    h_touch_keys = (unsigned long long*)(h_touch_report + 1);
    h_touch_theta = (float*)(h_touch_keys + n_keys);
    h_touch_cap = n_keys;
    return MCL_OK;
}

bool Engine::scans_are_fused() const {
    return !force_sequential && !force_multilaunch_scan && (n + xs::XSF_TILE - 1) / xs::XSF_TILE <= xs::XSF_MAX_TILES;
}

// first touchers (touch[k] = particle << 32 | beam, ~0 = none) and their theta -> directions evaluated with the host's libm
int Engine::ref_fill_ray_lut_from(const std::vector<HostBeam>& all, const unsigned long long* touch, const float* theta) {
    bool changed = false;
    const int stride = std::max(1, cfg.beam_stride);
    for (int k = 0; k < n_keys; ++k) {
        if (h_lut_filled[k] || touch[k] == ~0ull) continue;
        const unsigned beam = (unsigned)(touch[k] & 0xffffffffu);
        const double off_deg = -(all[(size_t)beam * stride].angle) * 180.0 / M_PI;
        const double yaw = tf_yaw_roundtrip((double)theta[k]);
        const double angle_rad = yaw + off_deg * M_PI / 180.0;
        h_lut[k] = make_double2(std::cos(angle_rad), std::sin(angle_rad));
        h_lut_filled[k] = 1;
        --n_unfilled;
        changed = true;
    }
    if (changed) {
        CK(cudaMemcpyAsync(d_lut.p, h_lut.data(), n_keys * sizeof(double2), cudaMemcpyHostToDevice, stream));
        CK(cudaMemcpyAsync(d_lut_filled.p, h_lut_filled.data(), n_keys, cudaMemcpyHostToDevice, stream));
        CK(cudaStreamSynchronize(stream));     // h_lut must not change under an in-flight copy
    }
    return MCL_OK;
}

// ---- resample ---------------------------------------------------------------------------------------------------
int Engine::resample(int jitter_state, const mcl_resample_draws* d, mcl_resample_stats* st) {
    CK(cudaSetDevice(cfg.device));
    if (n == 0) return fail(MCL_ERR_ARG, "resample: no particles");
    if (!have_weights) return fail(MCL_ERR_ARG, "resample: call mcl_update first (resampleParticles weighs before it resamples, MC:468)");
    if (cfg.mode == MCL_MODE_REF) return ref_resample(jitter_state, d, st);
    if (shard_world != 1) return fail(MCL_ERR_STATE, "resample: sharded NS filters are driven through the mcl_ns_*_local phases");
    uint64_t tot = 0;
    CK(cudaMemcpyAsync(&tot, d_u64.p, sizeof(uint64_t), cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
    int64_t k_lo = 0, k_hi = 0;
    int rc = ns_resample_local(0, tot, ns_u0(), &k_lo, &k_hi);
    if (rc) return rc;
    rc = ns_end_step();
    if (rc) return rc;
    if (st) { st->injected = 0; st->clamped = 0; st->p_inject = 0; st->weight_slow = 0; st->weight_fast = 0; st->total_weight = (double)tot * 2.3283064365386963e-10; }
    return MCL_OK;
}

// normalise + sequential CDF (MC:496-505) and the guide table of the CDF search: everything resampling needs that does not
// depend on a host decision, so mcl_step can enqueue it before it waits for the weight total.
int Engine::ref_resample_front() {
    guide_built = false;
    int buckets = 0;
    if (n >= 4096 && !force_sequential) {
        buckets = 1024;
        while ((int64_t)buckets * 8 < n && buckets < (1 << 24)) buckets <<= 1;          // ~8 CDF entries per bucket: 3 probes
        CK(d_guide.ensure((size_t)buckets + 2));
    }
    if (cdf_by_total && buckets == 0) { cdf_by_total = false; return MCL_OK; }      // (k_ref_scans_one_tile: the total's launch wrote the CDF as well)
    cdf_by_total = false;
    int rc = exact_accumulate(true, nullptr, nullptr, buckets);          // the one-kernel form scatters the guide table as it writes the CDF
    if (rc) return rc;
    if (buckets) {
        if (!guide_in_cdf) LAUNCH_PDL(K_GUIDE, k_ref_guide, grid_for(n, 256), 256, 0, cdf.p, n, buckets, d_guide.p);
        guide_built = true; guide_buckets = buckets;
    }
    return MCL_OK;
}

// dev_ema (mcl_step): the adaptive-injection state lives in device memory (d_inj, advanced by k_ref_ema) and the kernels
// read p_inject from there, so nothing here waits for the weight total; draws are staged through the pinned ring because
// several steps may be in flight; counters travel to the pinned step block and the call does not synchronise.
int Engine::ref_resample(int jitter_state, const mcl_resample_draws* d, mcl_resample_stats* st, bool front_done, bool dev_ema) {
    jitter_state = jitter_state ? 1 : 0;
    const double a_slow = jitter_state ? cfg.inject_alpha_slow_lost : cfg.inject_alpha_slow_conf;
    const double a_fast = jitter_state ? cfg.inject_alpha_fast_lost : cfg.inject_alpha_fast_conf;
    double p_inject = 0.0;
    if (dev_ema) {
        // (the one-kernel accumulation of the total has already advanced the state and cleared the counters: ema_in_total)
        if (!ema_in_total) LAUNCH_PDL(K_INJECT_SCAN, k_ref_ema, 1, 1, 0, d_scalars.p, (double)n, a_slow, a_fast, d_inj.p, d_counters.p);    // also clears the counters
    } else {
        int rc0 = inj_sync_to_host();
        if (rc0) return rc0;
        // adaptive injection EMA (MC:469-492)
        const double weight_avg = last_total / (double)n;
        inj_slow = inj_slow + a_slow * (weight_avg - inj_slow);
        inj_fast = inj_fast + a_fast * (weight_avg - inj_fast);
        p_inject = std::max(0.0, 1.0 - (inj_fast / inj_slow));
    }
    const double* inj_dev = dev_ema ? (const double*)d_inj.p : (const double*)nullptr;
    RefResampleParams R = make_resample_params(cfg, n, jitter_state, p_inject);
    R.inj_cols = (uint32_t)std::max(1, (int)((unsigned)map_w / (unsigned)cfg.cell_size_px));
    R.inj_rows = (uint32_t)std::max(1, (int)((unsigned)map_h / (unsigned)cfg.cell_size_px));
    const int per = jitter_state ? 3 : 2;
    const int max_inj = R.max_inject;
    // stage draws. Without injected draws nothing crosses: the kernels generate u_r / u_jitter (stream 0x30) and the named
    // draws of injected particles (stream 0x31) from the engine's Philox generator themselves.
    CK(d_inj_f64.ensure(3 * (size_t)std::max(1, max_inj))); CK(d_inj_i32.ensure(2 * (size_t)std::max(1, max_inj)));
    int rc;
    if (d) {
        CK(d_u_r.ensure((size_t)n)); CK(d_u_jit.ensure((size_t)n * 3));
        const size_t inj_f64 = 3 * (size_t)max_inj, inj_i32 = 2 * (size_t)max_inj;
        const size_t stage_bytes = ((size_t)n * (1 + per) + inj_f64) * sizeof(double) + inj_i32 * sizeof(int);
        double* hr;
        if (dev_ema) {
            rc = ensure_pinned_ring(std::max<size_t>(stage_bytes, 64));
            if (rc) return rc;
            hr = (double*)pinned_ring_next();
        } else {
            rc = ensure_pinned(stage_bytes);
            if (rc) return rc;
            hr = (double*)h_pinned;
        }
        double* hj = hr + n;
        double* hif = hj + (size_t)n * per;
        int* hii = (int*)(hif + inj_f64);
        if (!d->u_r || !d->u_jitter) return fail(MCL_ERR_ARG, "resample: null draw array");
        if (d->n_jitter < (int64_t)n * per) return fail(MCL_ERR_ARG, "resample: u_jitter shorter than N*(2|3)");
        memcpy(hr, d->u_r, (size_t)n * sizeof(double));
        memcpy(hj, d->u_jitter, (size_t)n * per * sizeof(double));
        const int have = std::min(d->n_inject, max_inj);
        if (p_inject > 0.0 && have < max_inj)
            if (!d->inject.u_yaw || d->n_inject < max_inj) return fail(MCL_ERR_ARG, "resample: p_inject > 0 needs max_injection named draws");
        for (int i = 0; i < max_inj; ++i) {
            const bool ok = i < have && d->inject.u_yaw;
            hif[i] = ok ? d->inject.u_yaw[i] : 0.0;
            hif[max_inj + i] = ok ? d->inject.u_dx[i] : 0.0;
            hif[2 * max_inj + i] = ok ? d->inject.u_dy[i] : 0.0;
            hii[i] = ok ? d->inject.row[i] : 0;
            hii[max_inj + i] = ok ? d->inject.col[i] : 0;
        }
        CK(cudaMemcpyAsync(d_u_r.p, hr, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, stream));
        CK(cudaMemcpyAsync(d_u_jit.p, hj, (size_t)n * per * sizeof(double), cudaMemcpyHostToDevice, stream));
        if (max_inj > 0) {
            CK(cudaMemcpyAsync(d_inj_f64.p, hif, inj_f64 * sizeof(double), cudaMemcpyHostToDevice, stream));
            CK(cudaMemcpyAsync(d_inj_i32.p, hii, inj_i32 * sizeof(int), cudaMemcpyHostToDevice, stream));
        }
        if (dev_ema) CK(cudaEventRecord(ring_events[ring_pos], stream));
    }
    last_per = per;
    draws_generated = d == nullptr;
    draws_step = (uint32_t)step_counter;
    const RefDrawGen G{(uint32_t)step_counter, (uint32_t)cfg.seed, (uint32_t)(cfg.seed >> 32)};
    if (!dev_ema) CK(cudaMemsetAsync(d_counters.p, 0, 4 * sizeof(int), stream));
    const unsigned blocks = grid_for(n, 256);
    // NaN p_inject compares false (MC:492, std::max(0.0, NaN) = 0.0); dev_ema: the kernels decide from inj_dev[2]
    const bool inject_possible = max_inj > 0 && (dev_ema || p_inject > 0.0);
    if (inject_possible && inject_by_scans && dev_ema && !d) {
        // (k_ref_scans_one_tile counted the flagged slots behind the total)
    } else if (inject_possible) {
        CK(d_block_counts.ensure(blocks));
        // (counts per block, then - by the last block to finish - their exclusive offsets and total: [5] = that kernel's ticket)
        if (d) LAUNCH_PDL(K_INJECT_COUNT, k_ref_inject_count<false>, (blocks + RIC_SEGS - 1) / RIC_SEGS, 256, 0, d_u_r.p, n, p_inject, d_block_counts.p, (int)blocks, G, inj_dev, d_counters.p + 2, (unsigned*)(d_counters.p + 5), tick_abort);
        else LAUNCH_PDL(K_INJECT_COUNT, k_ref_inject_count<true>, (blocks + RIC_SEGS - 1) / RIC_SEGS, 256, 0, (const double*)nullptr, n, p_inject, d_block_counts.p, (int)blocks, G, inj_dev, d_counters.p + 2, (unsigned*)(d_counters.p + 5), tick_abort);
        CK(cudaGetLastError());
    }
    inject_by_scans = false;
    if (!front_done) { rc = ref_resample_front(); if (rc) return rc; }
    // the guide table is used whenever the CDF is known to be non-decreasing: finite positive total weight (dev_ema: inj_dev[3])
    const bool use_guide = guide_built && (dev_ema || (std::isfinite(last_total) && last_total > 0.0));
    const int* guide = use_guide ? d_guide.p : nullptr;
    const int buckets = use_guide ? guide_buckets : 0;
    // Small filters inside mcl_step: the resampling kernel also sums the pose of the particles it writes and its last block
    // stores the tick's report (k_pose_sums' work: same thread -> particle mapping, same summation order, same bits), so the
    // tick ends one dependent launch earlier. (Not at large N: a fence and a ticket per 256-particle block cost more than the
    // launch they save.)
    PoseTail PT;
    PT.partials = nullptr; PT.ticket = nullptr; PT.out4 = nullptr; PT.report = nullptr; PT.inj5 = nullptr; PT.counters4 = nullptr; PT.seq = 0;
    pose_by_resample = false;
    if (dev_ema && fuse_pose_into_resample && !force_two_scan_launches && blocks <= 64) {
        CK(d_partials.ensure(4 * 1024));
        PT.partials = d_partials.p; PT.ticket = (unsigned*)(d_counters.p + 4); PT.out4 = d_scalars.p + 2; PT.report = h_step;
        PT.inj5 = d_inj.p; PT.counters4 = d_counters.p; PT.seq = ++step_seq;
        pose_by_resample = true;
    }
    const size_t cdf_smem = (!guide && n <= 4096) ? (size_t)n * sizeof(double) : 0;      // (32 KB at most: no opt-in needed)
    if (d)
        LAUNCH_PDL(K_RESAMPLE, k_ref_resample<false>, blocks, 256, cdf_smem, part[cur].p, part[cur ^ 1].p, n, cdf.p, d_u_r.p, d_u_jit.p, d_inj_f64.p, d_inj_i32.p,
               d_inj_i32.p + max_inj, d_inj_f64.p + max_inj, d_inj_f64.p + 2 * max_inj,
               inject_possible ? (const int*)d_block_counts.p : (const int*)nullptr, R, ancestors.p, d_counters.p, G, guide, buckets, inj_dev, tick_abort, PT, (float)((double)n * (double)R.new_weight), cdf_smem ? 1 : 0);
    else
        LAUNCH_PDL(K_RESAMPLE, k_ref_resample<true>, blocks, 256, cdf_smem, part[cur].p, part[cur ^ 1].p, n, cdf.p, (const double*)nullptr, (const double*)nullptr,
               d_inj_f64.p, d_inj_i32.p, d_inj_i32.p + max_inj, d_inj_f64.p + max_inj, d_inj_f64.p + 2 * max_inj,
               inject_possible ? (const int*)d_block_counts.p : (const int*)nullptr, R, ancestors.p, d_counters.p, G, guide, buckets, inj_dev, tick_abort, PT, (float)((double)n * (double)R.new_weight), cdf_smem ? 1 : 0);
    CK(cudaGetLastError());
    int counters[4] = {0, 0, 0, 0};
    if (!dev_ema) {
        CK(cudaMemcpyAsync(counters, d_counters.p, sizeof(counters), cudaMemcpyDeviceToHost, stream));
        CK(cudaStreamSynchronize(stream));
    }
    cur ^= 1;
    have_weights = false;
    ++step_counter;
    wsum_known = true; known_wsum = (double)n * (double)R.new_weight;        // every new particle weighs (float)(1/N), MC:524,551
    if (st && !dev_ema) {
        st->injected = counters[0]; st->clamped = counters[1]; st->p_inject = p_inject;
        st->weight_slow = inj_slow; st->weight_fast = inj_fast; st->total_weight = last_total;
    }
    return MCL_OK;
}

// The adaptive-injection state has two homes: the host (per-function calls, mcl_get/set_injection_state) and device memory
// (mcl_step). Whoever advanced it last is the owner; the other side is refreshed on demand.
int Engine::inj_sync_to_host() {
    if (!inj_on_device) return MCL_OK;
    double h[5];
    CK(cudaMemcpyAsync(h, d_inj.p, sizeof(h), cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
    inj_slow = h[0]; inj_fast = h[1];
    inj_on_device = false;
    return MCL_OK;
}
int Engine::inj_sync_to_device() {
    if (inj_on_device) return MCL_OK;
    CK(d_inj.ensure(8));
    const double h[5] = {inj_slow, inj_fast, 0.0, 0.0, last_total};
    CK(cudaMemcpyAsync(d_inj.p, h, sizeof(h), cudaMemcpyHostToDevice, stream));      // pageable source: staged before the call returns
    inj_on_device = true;
    return MCL_OK;
}

// ---- estimate -----------------------------------------------------------------------------------------------------
// h_sums4: where the four sums are copied to; step_report (mcl_step, pinned host block): the sums, the injection state and
// the resampling counters are written there by the kernel itself instead (no copy command).
int Engine::estimate_enqueue(double* h_sums4, RefStepReport* step_report) {
    if (pose_by_resample && step_report && !h_sums4) { pose_by_resample = false; return MCL_OK; }       // (k_ref_resample did it)
    pose_by_resample = false;
    { int rc = ns_materialise_weights(); if (rc) return rc; }
    const int blocks = (int)std::min<int64_t>(1024, grid_for(n, 256));
    CK(d_partials.ensure(4 * 1024));
    const double* wsum_dev = nullptr;
    if (!(wsum_known && cfg.mode == MCL_MODE_REF)) {            // weights of unknown provenance (uploaded, NS records): sum them first
        LAUNCH(K_POSE_WSUM, k_pose_wsum, blocks, 256, 0, part[cur].p, n, d_partials.p);
        LAUNCH(K_REDUCE, k_reduce_partials, 1, 32, 0, d_partials.p, blocks, 1, 1, d_scalars.p + 1);
        wsum_dev = d_scalars.p + 1;
    }
    PoseTail PT;
    PT.partials = d_partials.p; PT.ticket = (unsigned*)(d_counters.p + 4); PT.out4 = d_scalars.p + 2; PT.report = step_report;
    PT.inj5 = d_inj.p; PT.counters4 = d_counters.p; PT.seq = step_report ? ++step_seq : 0ull;
    LAUNCH_PDL(K_POSE_SUMS, k_pose_sums, blocks, 256, 0, part[cur].p, n, wsum_dev, known_wsum, PT, step_report ? tick_abort : (const int*)nullptr);
    CK(cudaGetLastError());
    if (h_sums4) CK(cudaMemcpyAsync(h_sums4, d_scalars.p + 2, 4 * sizeof(double), cudaMemcpyDeviceToHost, stream));
    return MCL_OK;
}

int Engine::estimate(double* x, double* y, double* th) {
    CK(cudaSetDevice(cfg.device));
    if (n == 0) return fail(MCL_ERR_ARG, "estimate: no particles");
    double s[4];
    { int rc = estimate_enqueue(s); if (rc) return rc; }
    CK(cudaStreamSynchronize(stream));
    const float xm = (float)s[0], ym = (float)s[1];
    const float tm = std::atan2((float)s[2], (float)s[3]);      // MC:796 (fp32 atan2)
    if (x) *x = xm;
    if (y) *y = ym;
    if (th) *th = tm;
    return MCL_OK;
}

// ---- one whole step of the reference loop (MC:1084-1092) with a single wait at its end ---------------------------------------
// predict -> computeWeight -> [total starts travelling to the host] -> normalise + CDF + guide table -> (host: adaptive
// injection from the total, MC:469-492, while those kernels run) -> resample -> pose sums -> one synchronisation for the
// counters and the pose. Same kernels, same results as mcl_predict_encoders + mcl_update + mcl_resample + mcl_estimate,
// which wait for the GPU three times. Draws come from the engine's Philox streams (parity runs inject theirs per call).
int Engine::ref_step(double enc_l, double enc_r, int slot, const float* ranges, int n_beams, float angle_min, float angle_inc, float range_min,
                     float range_max, int jitter_state, double* pose3, mcl_resample_stats* st, bool allow_optimistic) {
    CK(cudaSetDevice(cfg.device));
    if (cfg.mode != MCL_MODE_REF) return fail(MCL_ERR_STATE, "step: MCL_MODE_REF only (NS filters use mcl_ns_step)");
    if (!map_ready) return fail(MCL_ERR_ARG, "step: no map");
    if (n == 0) return fail(MCL_ERR_ARG, "step: no particles");
    if (!h_step) { CK(cudaMallocHost((void**)&h_step, sizeof(StepScalars))); memset(h_step, 0, sizeof(StepScalars)); }
    int rc = inj_sync_to_device();
    if (rc) return rc;
    // An optimistic tick (ref_run_update) needs a caller that waits for the report, and everything the host side of a tick
    // changes is kept so that the tick can be run again
    tick_abort = (allow_optimistic && (pose3 || st)) ? d_counters.p + 6 : nullptr;
    struct Saved { decltype(odo) odo; decltype(step_counter) step_counter; int cur; PendingMotion pending; bool have_weights, wsum_known; double known_wsum;
                   int last_per; bool draws_generated; uint32_t draws_step; double last_total; } saved{odo, step_counter, cur, pending_motion, have_weights,
                   wsum_known, known_wsum, last_per, draws_generated, draws_step, last_total};
    struct Reset { const int*& p; ~Reset() { p = nullptr; } } reset_abort{tick_abort};
    // A scan from the host: its scored beams ride in the computeWeight kernel's launch parameters (ref_run_update), so the
    // tick is kernels only (programmatic launches overlap only kernel with kernel).
    const bool host_scan = ranges || slot < 0;
    int n_used = 0;
    if (host_scan) {
        if (n_beams < 0 || (n_beams > 0 && !ranges)) return fail(MCL_ERR_ARG, "step: bad scan");
        std::vector<RefBeam> used;
        rc = ref_prepare_beams(ranges, n_beams, angle_min, angle_inc, range_min, range_max, beams_all, used);
        if (rc) return rc;
        step_used.swap(used);
        n_used = (int)step_used.size();
    } else if ((size_t)slot >= staged.size() || !staged[slot].valid) return fail(MCL_ERR_ARG, "step: empty scan slot");
    defer_predict = true;
    rc = predict_encoders(enc_l, enc_r, nullptr, nullptr);
    defer_predict = false;
    if (rc) return rc;
    EmaArgs ema;
    ema.a_slow = jitter_state ? cfg.inject_alpha_slow_lost : cfg.inject_alpha_slow_conf;
    ema.a_fast = jitter_state ? cfg.inject_alpha_fast_lost : cfg.inject_alpha_fast_conf;
    fuse_cdf_into_total = true;              // (the resampling that needs the CDF follows in this very call)
    if (host_scan) rc = ref_run_update(nullptr, step_used.data(), n_used, beams_all, nullptr, true, &ema);
    else rc = ref_run_update(staged[slot].d_used.p, nullptr, staged[slot].n_used, staged[slot].all, nullptr, true, &ema);
    fuse_cdf_into_total = false;
    if (rc) { flush_pending_motion(); return rc; }        // (a tick that failed before its computeWeight kernel still moves the particles)
    fuse_pose_into_resample = true;
    rc = ref_resample(jitter_state, nullptr, nullptr, false, true);
    fuse_pose_into_resample = false;
    if (rc) return rc;
    // the estimate is part of every tick; its last block writes the tick's scalars straight into the pinned block
    rc = estimate_enqueue(nullptr, h_step);
    if (rc) return rc;
    if (!pose3 && !st) return MCL_OK;                           // nothing asked for: the tick is queued, the host moves on
    // The report lands in pinned memory, its sequence number last: the host watches for that instead of waiting for the
    // stream to be reported idle (a few microseconds later). Every few thousand looks it asks the runtime as well, so that a
    // failed launch ends the wait with its error.
    {
        const volatile unsigned long long* seq = &h_step->seq;
        const unsigned long long want = step_seq;
        for (unsigned spins = 1; *seq != want; ++spins) {
            if ((spins & 0xfff) == 0 && cudaStreamQuery(stream) != cudaErrorNotReady) { CK(cudaStreamSynchronize(stream)); break; }
#if defined(__x86_64__) || defined(__i386__)
            __builtin_ia32_pause();
#endif
        }
        if (*seq != want) return fail(MCL_ERR_CUDA, "step: the tick finished without writing its report");
    }
    if (tick_optimistic && h_step->aborted) {
        // the pre-pass found ray directions that must be evaluated first: no kernel of the tick ran. Evaluate them (host libm),
        // put the host side back and run the tick again, this time with the pre-pass waited for.
        const std::vector<HostBeam>& all = host_scan ? beams_all : staged[slot].all;
        if (h_touch_report->seq != touch_seq) return fail(MCL_ERR_CUDA, "step: aborted tick without a pre-pass report");
        rc = ref_fill_ray_lut_from(all, h_touch_keys, h_touch_theta);
        if (rc) return rc;
        CK(cudaMemsetAsync(d_counters.p + 6, 0, sizeof(int), stream));
        odo = saved.odo; step_counter = saved.step_counter; cur = saved.cur; pending_motion = saved.pending; have_weights = saved.have_weights;
        wsum_known = saved.wsum_known; known_wsum = saved.known_wsum; last_per = saved.last_per; draws_generated = saved.draws_generated;
        draws_step = saved.draws_step; last_total = saved.last_total;
        ++optimistic_redos;
        return ref_step(enc_l, enc_r, slot, ranges, n_beams, angle_min, angle_inc, range_min, range_max, jitter_state, pose3, st, false);
    }
    last_total = h_step->inj[4];
    if (st) {
        st->injected = h_step->counters[0]; st->clamped = h_step->counters[1]; st->p_inject = h_step->inj[2];
        st->weight_slow = h_step->inj[0]; st->weight_fast = h_step->inj[1]; st->total_weight = h_step->inj[4];
    }
    if (pose3) {
        pose3[0] = (double)(float)h_step->pose[0]; pose3[1] = (double)(float)h_step->pose[1];
        pose3[2] = (double)std::atan2((float)h_step->pose[2], (float)h_step->pose[3]);      // MC:796 (fp32 atan2)
    }
    return MCL_OK;
}

}  // namespace mcl
