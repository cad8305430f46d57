// mcl_capi.cu — the extern "C" boundary declared in include/mcl.h (+ the instrumentation of include/mcl_debug.h). Nothing
// throws across it.
#include <cmath>
#include <cstring>
#include <new>
#include <string>

#include "../../include/mcl_debug.h"
#include "mcl_engine.hpp"
#include "ns_plan.hpp"

struct mcl_handle {
    mcl::Engine engine;
    explicit mcl_handle(const mcl_config& c) : engine(c) {}
};

static thread_local std::string g_create_error;

#define GUARD(h) do { if (!(h)) return MCL_ERR_ARG; } while (0)
#define TRY(expr)                                                                       \
    try { return (expr); }                                                              \
    catch (const std::exception& ex) { h->engine.err = std::string("exception: ") + ex.what(); return MCL_ERR_ARG; } \
    catch (...) { h->engine.err = "unknown exception"; return MCL_ERR_ARG; }

extern "C" {

const char* mcl_version(void) { return "mcl_b200 0.1 (sm_100a)"; }

void mcl_config_default(mcl_config* c) {
    if (!c) return;
    memset(c, 0, sizeof(*c));
    c->device = 0; c->mode = MCL_MODE_REF; c->max_particles = 0;
    c->sigma_hit = 0.1; c->max_laser_range = 1.0; c->laser_offset = 0.1;            // MC:627-631
    c->w_hit = 0.8; c->w_rand = 0.2;                                                 // MC:180-181
    c->fov_lower_deg = -120.00; c->fov_upper_deg = 120.00; c->beam_stride = 20;      // MC:635,650
    c->ray_step = 0.1; c->validity_offset = 0.1;                                     // MC:370,333
    c->alpha[0] = 0.001; c->alpha[1] = 0.001; c->alpha[2] = 0.0001; c->alpha[3] = 0.0001;   // MC:1198
    c->wheel_size = 0.0620; c->wheel_space = 0.265;                                  // PID_lib.hpp:19-20
    c->cell_size_px = 8; c->cell_meters = 0.8; c->init_offset = 0.2; c->init_shift = 0.05;  // MC:422,396,431,442
    c->inject_max_lost = 200; c->inject_alpha_slow_lost = 0.05; c->inject_alpha_fast_lost = 0.5;   // MC:474-476
    c->inject_max_conf = 50; c->inject_alpha_slow_conf = 0.02; c->inject_alpha_fast_conf = 2;      // MC:479-481
    c->jitter_xy_lost = 0.05; c->jitter_theta_lost = M_PI / 12; c->jitter_xy_conf = 0.01;          // MC:537-545
    c->seed = 0x9E3779B97F4A7C15ull;
    c->ns_sigma_hit = 0.1; c->ns_z_hit = 0.8; c->ns_z_rand = 0.2; c->ns_max_range = 5.6;
    c->ns_beam_stride = 1; c->ns_use_fov = 0; c->ns_temper = 0.05;
    c->kmeans_radius = 0.4;                                                                         // MC:933
    c->trig_mode = MCL_TRIG_LIBM;                                                                   // what the reference binary executes
}

int mcl_create(const mcl_config* cfg, mcl_handle** out) {
    if (!cfg || !out) { g_create_error = "mcl_create: null argument"; return MCL_ERR_ARG; }
    *out = nullptr;
    if (cfg->mode != MCL_MODE_REF && cfg->mode != MCL_MODE_NS) { g_create_error = "mcl_create: unknown mode"; return MCL_ERR_ARG; }
    if (cfg->trig_mode != MCL_TRIG_LIBM && cfg->trig_mode != MCL_TRIG_CORRECTLY_ROUNDED) { g_create_error = "mcl_create: unknown trig_mode"; return MCL_ERR_ARG; }
    mcl_handle* h = new (std::nothrow) mcl_handle(*cfg);
    if (!h) { g_create_error = "mcl_create: out of host memory"; return MCL_ERR_ARG; }
    int rc;
    try { rc = h->engine.open(); } catch (...) { rc = MCL_ERR_CUDA; h->engine.err = "exception during open"; }
    if (rc) { g_create_error = h->engine.err; delete h; return rc; }
    *out = h;
    return MCL_OK;
}
void mcl_destroy(mcl_handle* h) { delete h; }
const char* mcl_last_error(mcl_handle* h) { return h ? h->engine.err.c_str() : g_create_error.c_str(); }

int mcl_rasterise_map_txt(const char* text, int8_t* out, int64_t cap, int32_t* width, int32_t* height) {
    if (!text || !width || !height) return MCL_ERR_ARG;
    mcl::WallGrid g; std::string perr;
    if (!mcl::parse_map_txt(text, g, perr)) { g_create_error = "map.txt: " + perr; return MCL_ERR_IO; }
    std::vector<int8_t> occ; int w, hh;
    mcl::rasterise_walls(g, 8, occ, w, hh);
    *width = w; *height = hh;
    if (!out) return MCL_OK;
    if ((int64_t)occ.size() > cap) return MCL_ERR_ARG;
    memcpy(out, occ.data(), occ.size());
    return MCL_OK;
}

int mcl_set_map(mcl_handle* h, const int8_t* occ, int32_t w, int32_t hh, float res, double ox, double oy) { GUARD(h); TRY(h->engine.set_map(occ, w, hh, res, ox, oy)) }
int mcl_load_map_txt(mcl_handle* h, const char* path) { GUARD(h); TRY(h->engine.load_map_txt(path)) }
int mcl_precompute_ray_directions(mcl_handle* h, double a, double b, double s) { GUARD(h); TRY(h->engine.precompute_ray_directions(a, b, s)) }
int mcl_init(mcl_handle* h, int64_t n, const mcl_init_draws* d) { GUARD(h); TRY(h->engine.init(n, d)) }
int mcl_upload(mcl_handle* h, const float* p, int64_t n) { GUARD(h); TRY(h->engine.upload(p, n)) }
int mcl_download(mcl_handle* h, float* p) { GUARD(h); TRY(h->engine.download(p)) }
int64_t mcl_num_particles(mcl_handle* h) { return h ? h->engine.n : 0; }
int mcl_predict_encoders(mcl_handle* h, double l, double r, const double* z3, double* m) { GUARD(h); TRY(h->engine.predict_encoders(l, r, z3, m)) }
int mcl_predict_motion(mcl_handle* h, double r1, double t, double r2) { GUARD(h); TRY(h->engine.predict_motion(r1, t, r2)) }
int mcl_update(mcl_handle* h, const float* ranges, int32_t nb, float amin, float ainc, float rmin, float rmax, double* total) {
    GUARD(h); TRY(h->engine.update(ranges, nb, amin, ainc, rmin, rmax, total))
}
int mcl_scan_stage(mcl_handle* h, int32_t slot, const float* ranges, int32_t nb, float amin, float ainc, float rmin, float rmax) {
    GUARD(h); TRY(h->engine.stage_scan(slot, ranges, nb, amin, ainc, rmin, rmax))
}
int mcl_update_staged(mcl_handle* h, int32_t slot, double* total) { GUARD(h); TRY(h->engine.update_staged(slot, total)) }
int mcl_resample(mcl_handle* h, int32_t js, const mcl_resample_draws* d, mcl_resample_stats* st) { GUARD(h); TRY(h->engine.resample(js, d, st)) }
int mcl_download_ancestors(mcl_handle* h, int32_t* idx) { GUARD(h); TRY(h->engine.download_ancestors(idx)) }
int mcl_download_cdf(mcl_handle* h, double* cdf) { GUARD(h); TRY(h->engine.download_cdf(cdf)) }
int mcl_estimate(mcl_handle* h, double* x, double* y, double* th) { GUARD(h); TRY(h->engine.estimate(x, y, th)) }
int mcl_step(mcl_handle* h, double el, double er, const float* ranges, int32_t nb, float amin, float ainc, float rmin, float rmax, int32_t js,
             double* pose3, mcl_resample_stats* st) {
    GUARD(h);
    TRY(h->engine.ref_step(el, er, -1, ranges, nb, amin, ainc, rmin, rmax, js, pose3, st))
}
int mcl_step_staged(mcl_handle* h, double el, double er, int32_t slot, int32_t js, double* pose3, mcl_resample_stats* st) {
    GUARD(h);
    if (slot < 0) return MCL_ERR_ARG;
    TRY(h->engine.ref_step(el, er, slot, nullptr, 0, 0.f, 0.f, 0.f, 0.f, js, pose3, st))
}
int mcl_get_injection_state(mcl_handle* h, double* s, double* f) {
    GUARD(h);
    int rc = h->engine.inj_sync_to_host();
    if (rc) return rc;
    if (s) *s = h->engine.inj_slow;
    if (f) *f = h->engine.inj_fast;
    return MCL_OK;
}
int mcl_set_injection_state(mcl_handle* h, double s, double f) {
    GUARD(h);
    int rc = h->engine.inj_sync_to_host();
    if (rc) return rc;
    h->engine.inj_slow = s; h->engine.inj_fast = f;
    return MCL_OK;
}
int mcl_get_ray_lut(mcl_handle* h, int32_t* keys, double* dx, double* dy, int32_t cap, int32_t* count) { GUARD(h); TRY(h->engine.get_ray_lut(keys, dx, dy, cap, count)) }
int mcl_debug_download_resample_draws(mcl_handle* h, double* u_r, double* u_jitter) { GUARD(h); TRY(h->engine.download_resample_draws(u_r, u_jitter)) }
int mcl_debug_exact_scan(mcl_handle* h, const float* w, int64_t n, double* cdf, double* total, int32_t* fell_back) { GUARD(h); TRY(h->engine.debug_exact_scan(w, n, cdf, total, fell_back)) }
int mcl_debug_force_sequential(mcl_handle* h, int32_t on) { GUARD(h); h->engine.force_sequential = (on & 1) != 0; h->engine.force_v1_update = (on & 2) != 0; h->engine.force_f64_probe = (on & 4) != 0;
    h->engine.ns_force_field = (on & 8) ? 2 : (on & 16) ? 1 : -1; h->engine.ns_force_scalar = (on & 32) != 0; h->engine.force_multilaunch_scan = (on & 64) != 0; h->engine.force_scan_fallback = (on & 128) != 0; h->engine.force_scan_tickets = (on & 256) != 0; h->engine.force_separate_guide = (on & 512) != 0; h->engine.force_two_scan_launches = (on & 1024) != 0; h->engine.force_scan_items16 = (on & 2048) != 0; return MCL_OK; }
int mcl_debug_exact_scan_trace(mcl_handle* h, unsigned long long* out, int64_t cap_tiles, int32_t* n_tiles) { GUARD(h); TRY(h->engine.debug_exact_scan_trace(out, cap_tiles, n_tiles)) }
int mcl_debug_ns_last_plan(mcl_handle* h, int64_t* k_lo, int64_t* k_hi, int64_t* own_begin, int64_t* own_count) { GUARD(h); TRY(h->engine.ns_last_plan(k_lo, k_hi, own_begin, own_count)) }
int mcl_debug_trigf(mcl_handle* h, const float* x, int64_t n, float* s, float* c, int32_t* kind) { GUARD(h); TRY(h->engine.debug_trigf(x, n, s, c, kind)) }
int mcl_profile_enable(mcl_handle* h, int32_t on) { GUARD(h); h->engine.profile_enable(on != 0); return MCL_OK; }
int mcl_profile_kernel_count(void) { return mcl::Engine::K_COUNT; }
const char* mcl_profile_kernel_name(int32_t id) { return mcl::Engine::kernel_name(id); }
int mcl_profile_read(mcl_handle* h, int32_t id, double* total_ms, int64_t* count) { GUARD(h); TRY(h->engine.profile_read(id, total_ms, count)) }
int mcl_ns_set_shard(mcl_handle* h, int32_t rank, int32_t world, int64_t ng) { GUARD(h); TRY(h->engine.ns_set_shard(rank, world, ng)) }
int mcl_ns_update_local(mcl_handle* h, const float* ranges, int32_t nb, float amin, float ainc, float rmin, float rmax, float* lm) {
    GUARD(h); TRY(h->engine.ns_update_local(ranges, nb, amin, ainc, rmin, rmax, lm))
}
int mcl_ns_update_local_staged(mcl_handle* h, int32_t slot, float* lm) { GUARD(h); TRY(h->engine.ns_update_local_staged(slot, lm)) }
int mcl_ns_weights_local(mcl_handle* h, float gm, uint64_t* lt) { GUARD(h); TRY(h->engine.ns_weights_local(gm, lt)) }
int mcl_ns_resample_local(mcl_handle* h, uint64_t off, uint64_t tot, uint32_t u0, int64_t* klo, int64_t* khi) { GUARD(h); TRY(h->engine.ns_resample_local(off, tot, u0, klo, khi)) }
int mcl_ns_end_step(mcl_handle* h) { GUARD(h); TRY(h->engine.ns_end_step()) }
uint32_t mcl_ns_u0(mcl_handle* h) { return h ? h->engine.ns_u0() : 0; }
int mcl_ns_pose_partials(mcl_handle* h, double* out5) { GUARD(h); TRY(h->engine.ns_pose_partials(out5)) }
int mcl_comm_unique_id(mcl_handle* h, void* out128) { GUARD(h); if (!out128) return MCL_ERR_ARG; TRY(h->engine.comm_unique_id(out128)) }
int mcl_comm_init(mcl_handle* h, const void* id128) { GUARD(h); if (!id128) return MCL_ERR_ARG; TRY(h->engine.comm_init(id128)) }
int mcl_ns_step(mcl_handle* h, double rot_1, double trans, double rot_2, const float* ranges, int32_t nb, float amin, float ainc, float rmin,
                float rmax, double* pose3) {
    GUARD(h);
    if (nb < 0 || (nb > 0 && !ranges)) return MCL_ERR_ARG;
    static const float none = 0.f;
    TRY(h->engine.ns_step(rot_1, trans, rot_2, -1, ranges ? ranges : &none, nb, amin, ainc, rmin, rmax, pose3))
}
int mcl_ns_step_staged(mcl_handle* h, double rot_1, double trans, double rot_2, int32_t slot, double* pose3) {
    GUARD(h); TRY(h->engine.ns_step(rot_1, trans, rot_2, slot, nullptr, 0, 0.f, 0.f, 0.f, 0.f, pose3))
}
int mcl_kmeans_confidence(mcl_handle* h, const int32_t* init_idx, const int32_t* reinit_idx, int32_t n_reinit, double thr, mcl_kmeans_result* out) {
    GUARD(h); TRY(h->engine.kmeans_confidence(init_idx, reinit_idx, n_reinit, thr, out))
}
int mcl_download_assignments(mcl_handle* h, int32_t* a) { GUARD(h); TRY(h->engine.download_assignments(a)) }
int mcl_pose_to_cell(double wx, double wy, double angle, double cell_meters, int32_t* row, int32_t* column, int32_t* orientation) {
    if (!row || !column || !orientation || !(cell_meters > 0)) return MCL_ERR_ARG;
    int r, c, o;
    mcl::pose_to_cell(wx, wy, angle, cell_meters, r, c, o);
    *row = r; *column = c; *orientation = o;
    return MCL_OK;
}
int mcl_exact_pose(double x, double y, double theta, float* out3) {
    if (!out3) return MCL_ERR_ARG;
    out3[0] = (float)x; out3[1] = (float)y; out3[2] = (float)theta;       // float32 message fields (msg/ExactPose.msg)
    return MCL_OK;
}
int mcl_download_pose_array(mcl_handle* h, int64_t first, int64_t stride, int64_t count, double* out) {
    GUARD(h); TRY(h->engine.download_pose_array(first, stride, count, out))
}
int mcl_config_preset(mcl_config* cfg, const char* name) {
    if (!cfg || !name) return MCL_ERR_ARG;
    const std::string n(name);
    const int32_t mode = cfg->mode, device = cfg->device;
    if (n == "reference") { mcl_config_default(cfg); cfg->mode = mode; cfg->device = device; return MCL_OK; }
    if (n == "playground") {
        mcl_config_default(cfg); cfg->mode = mode; cfg->device = device;
        cfg->ray_step = 0.05;            // playground.cpp:328
        cfg->beam_stride = 3;            // playground.cpp:604
        cfg->fov_lower_deg = -90.0; cfg->fov_upper_deg = 90.0;       // playground.cpp:601
        return MCL_OK;
    }
    return MCL_ERR_ARG;
}
int mcl_ns_first_slot(uint64_t off, uint64_t tot, uint64_t n, uint32_t u0, int64_t* slot) {
    if (!slot || tot == 0) return MCL_ERR_ARG;
    *slot = mcl::ns::first_slot(off, tot, n, u0);
    return MCL_OK;
}
int mcl_ns_shard_range(int64_t ng, int32_t world, int32_t rank, int64_t* b, int64_t* c, int64_t* per) {
    if (!b || !c || !per || world < 1 || rank < 0 || rank >= world) return MCL_ERR_ARG;
    mcl::ns::shard_range(ng, world, rank, b, c, per);
    return MCL_OK;
}
int mcl_bench_gather(mcl_handle* h, int32_t tier, int64_t table_bytes, int32_t iters, double* reads_per_s) { GUARD(h); TRY(h->engine.gather_bench(tier, (size_t)table_bytes, iters, reads_per_s)) }
int mcl_ns_set_exchange(mcl_handle* h, int32_t mode) { GUARD(h); if (mode < -1 || mode > 1) return MCL_ERR_ARG; h->engine.ns_exchange = mode; return MCL_OK; }
int mcl_ns_exchange_used(mcl_handle* h) { return h ? h->engine.ns_exchange_used : -1; }
int mcl_peer_export(mcl_handle* h, int32_t which, void* out64) { GUARD(h); TRY(h->engine.peer_export(which, out64)) }
int mcl_peer_import(mcl_handle* h, int32_t rank, int32_t which, const void* in64) { GUARD(h); TRY(h->engine.peer_import(rank, which, in64)) }
int mcl_peer_set(mcl_handle* h, int32_t rank, int32_t which, void* p) { GUARD(h); TRY(h->engine.peer_set(rank, which, p)) }
void* mcl_device_buffer(mcl_handle* h, int32_t which) { return h ? h->engine.device_buffer(which) : nullptr; }
int mcl_ns_download_field(mcl_handle* h, float* lf, uint16_t* d2) { GUARD(h); TRY(h->engine.ns_download_field(lf, d2)) }
int mcl_ns_download_loglik(mcl_handle* h, float* ll) { GUARD(h); TRY(h->engine.ns_download_loglik(ll)) }
int mcl_ns_download_prefix(mcl_handle* h, uint64_t* p) { GUARD(h); TRY(h->engine.ns_download_prefix(p)) }
int mcl_ns_field_form(mcl_handle* h) { return h ? h->engine.ns_field_kind : -1; }
void* mcl_stream(mcl_handle* h) { return h ? (void*)h->engine.stream : nullptr; }
int mcl_synchronize(mcl_handle* h) { GUARD(h); TRY(h->engine.synchronize()) }
int64_t mcl_kernel_launches(mcl_handle* h) { return h ? h->engine.launches : 0; }
int64_t mcl_debug_optimistic_redos(mcl_handle* h) { return h ? h->engine.optimistic_redos : 0; }
int mcl_debug_last_scan_fell_back(mcl_handle* h, int32_t* fell_back) { GUARD(h); TRY(h->engine.debug_last_scan_fell_back(fell_back)) }

}  // extern "C"
