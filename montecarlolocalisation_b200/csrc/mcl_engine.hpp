// mcl_engine.hpp — host side of the engine: owns the device state of one GPU shard and sequences the kernels.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <string>
#include <vector>

#include "../../include/mcl.h"
#include "host_models.hpp"

namespace mcl {

struct RefBeam;

template <class T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    cudaError_t ensure(size_t count) {
        if (count <= n) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; n = 0;
        cudaError_t e = cudaMalloc((void**)&p, count * sizeof(T));
        if (e == cudaSuccess) n = count;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
};

// The scalar results of one mcl_step tick, in pinned host memory; k_pose_sums' last block writes them there itself.
struct RefStepReport { double inj[5]; double pose[4]; int counters[4]; unsigned long long seq; int aborted; int pad; };      // seq: written last, the tick's number; aborted: optimistic tick that did not run

class Engine {
public:
    explicit Engine(const mcl_config& cfg);
    ~Engine();
    int open();                       // device, stream, static tables
    std::string err;

    // boundary (see include/mcl.h)
    int set_map(const int8_t* occ, int w, int h, float res, double ox, double oy);
    int load_map_txt(const char* path);
    int precompute_ray_directions(double lo, double hi, double step);
    int init(int64_t n, const mcl_init_draws* d);
    int upload(const float* p, int64_t n);
    int download(float* p);
    int predict_encoders(double enc_l, double enc_r, const double* z3, double* motion_out);
    int predict_motion(double r1, double t, double r2);
    // mcl_step: the motion of the tick being enqueued, applied by the computeWeight kernel as it loads the particles
    struct PendingMotion { bool valid = false; float rot1 = 0, trans = 0, dtheta = 0; } pending_motion;
    bool defer_predict = false;
    int flush_pending_motion();
    int update(const float* ranges, int n_beams, float angle_min, float angle_inc, float range_min, float range_max, double* total);
    // scans pre-staged in HBM (bench: inputs resident before the timed region)
    int stage_scan(int slot, const float* ranges, int n_beams, float angle_min, float angle_inc, float range_min, float range_max);
    int update_staged(int slot, double* total);
    int resample(int jitter_state, const mcl_resample_draws* d, mcl_resample_stats* st);
    int download_ancestors(int32_t* idx);
    int download_cdf(double* cdf);
    int estimate(double* x, double* y, double* th);
    int inj_sync_to_host();                 // the adaptive-injection state back on the host (after mcl_step)
    int ref_step(double enc_l, double enc_r, int slot, const float* ranges, int n_beams, float angle_min, float angle_inc, float range_min,
                 float range_max, int jitter_state, double* pose3, mcl_resample_stats* st, bool allow_optimistic = true);
    int get_ray_lut(int32_t* keys, double* dx, double* dy, int32_t cap, int32_t* count);
    int synchronize();
    // NS mode (north-star formulation); the *_local phases are what a multi-GPU driver sequences around its collectives
    int ns_set_shard(int rank, int world, int64_t n_global);
    int ns_update_local(const float* ranges, int n_beams, float angle_min, float angle_inc, float range_min, float range_max, float* local_max);
    int ns_stage_scan(int slot, const float* ranges, int n_beams, float angle_min, float angle_inc, float range_min, float range_max);
    int ns_update_local_staged(int slot, float* local_max);
    int ns_weights_local(float global_max, uint64_t* local_total);
    int ns_resample_local(uint64_t offset, uint64_t total, uint32_t u0, int64_t* k_lo, int64_t* k_hi);
    int ns_end_step();
    uint32_t ns_u0() const;
    int ns_pose_partials(double* out5);
    int comm_unique_id(void* out128);
    void ns_comm_destroy();
    int comm_init(const void* id128);
    int ns_step(double rot1, double trans, double rot2, int slot, const float* ranges, int n_beams, float angle_min, float angle_inc,
                float range_min, float range_max, double* pose3);
    int ns_download_field(float* lf, uint16_t* d2);
    int ns_download_loglik(float* ll);
    int ns_download_prefix(uint64_t* prefix);
    int ns_last_plan(int64_t* k_lo, int64_t* k_hi, int64_t* own_begin, int64_t* own_count);
    // rows either side of the hot path
    int kmeans_confidence(const int32_t* init_idx, const int32_t* reinit_idx, int n_reinit, double ratio_threshold, mcl_kmeans_result* out);
    int download_assignments(int32_t* a);
    int download_pose_array(int64_t first, int64_t stride, int64_t count, double* out);
    DevBuf<int> d_assign, d_km_reinit;
    DevBuf<unsigned char> d_km;
    DevBuf<double> d_posearr;
    int gather_bench(int tier, size_t table_bytes, int iters, double* reads_per_s);
    int peer_export(int which, void* out64);
    int peer_import(int rank, int which, const void* in64);
    int peer_set(int rank, int which, void* devptr);
    void* device_buffer(int which);
    int shard_rank = 0, shard_world = 1;
    int64_t n_global = 0, shard_begin = 0, per_rank = 0;
    int download_resample_draws(double* u_r, double* u_jit);
    int debug_exact_scan(const float* w, int64_t count, double* cdf_out, double* total_out, int* fell_back);
    int debug_trigf(const float* x, int64_t count, float* s_out, float* c_out, int* kind_out);
    int debug_exact_scan_trace(unsigned long long* out, int64_t cap_tiles, int* n_tiles);
    int debug_last_scan_fell_back(int* fell_back);
    // per-kernel CUDA-event timing (off by default; bench.py turns it on for its roofline pass)
    enum KernelId { K_INIT = 0, K_PREDICT, K_FIRST_TOUCH, K_TOUCH_THETA, K_UPDATE, K_UPDATE_V2, K_SEQ_TOTAL, K_FILL_DRAWS, K_INJECT_COUNT,
                    K_INJECT_SCAN, K_SEQ_CDF, K_GUIDE, K_RESAMPLE, K_XS_TILESUM, K_XS_OFFSETS, K_XS_SCAN, K_XS_CHAIN, K_XS_APPLY, K_XS_TOTAL, K_XS_CDF, K_NS_EDT_COLS, K_NS_EDT_ROWS, K_NS_INIT, K_NS_PREDICT, K_NS_UPDATE, K_NS_WSUM, K_NS_WSCAN, K_NS_PLAN, K_NS_BOUNDS, K_NS_RESAMPLE, K_NS_POSE, K_NS_POSE_REDUCE, K_KM_ASSIGN, K_KM_UPDATE, K_KM_STATS, K_POSE_ARRAY, K_POSE_WSUM, K_POSE_SUMS, K_REDUCE, K_SCANS_ONE_TILE, K_COUNT };
    static const char* kernel_name(int id);
    void profile_enable(bool on);
    int profile_read(int id, double* total_ms, int64_t* count);

    mcl_config cfg;
    bool force_f64_probe = false;       // tests: ray-parallel kernel without the fp32 pre-filter
    bool force_v1_update = false;       // tests: the one-thread-per-particle computeWeight kernel
    int ns_exchange = -1;               // sharded step's collectives: 0 = NCCL, 1 = peer-memory mailboxes, -1 = mailboxes when mapped
    int ns_exchange_used = -1;          // what the last sharded step used
    int ns_field_kind = -1;             // NS_FIELD_* the last sensor-model launch used
    int ns_force_field = -1;            // tests: NS_FIELD_* to use regardless of size (-1 = by size)
    bool ns_force_scalar = false;       // tests: scalar FFMA form of the sensor model on every path
    bool force_sequential = false;      // tests: use the single-chain kernels instead of the exact parallel scan
    int xs_resident_tiles = -1;          // blocks of the one-kernel exact scan the device holds at once (queried on first use)
    bool force_scan_tickets = false;     // tests: ticket order even when the grid is co-resident
    bool force_separate_guide = false;   // tests / A-B: the guide table by its own launch (k_ref_guide) behind the one-kernel CDF
    bool fuse_cdf_into_total = false;    // mcl_step: the accumulation of the total may also write the normalised CDF (one tile: k_xs_both)
    bool cdf_by_total = false;           // ... and did
    bool fuse_pose_into_resample = false; // mcl_step: the resampling kernel of a small filter may also produce the pose sums and the report
    bool pose_by_resample = false;       // ... and did
    bool inject_by_scans = false;        // ... and counted the slots flagged for injection (k_ref_inject_count's work) too
    bool force_scan_items16 = false;      // tests / A-B: k_ref_scans_one_tile with 16 weights per thread whatever the size
    bool force_two_scan_launches = false; // tests / A-B: never k_xs_both
    bool force_scan_fallback = false;    // tests: the one-kernel exact scan takes its in-kernel single-chain fallback every time
    bool guide_in_cdf = false;           // the last CDF accumulation also scattered the guide table
    bool force_multilaunch_scan = false; // tests / A-B: the multi-launch exact scan (exact_scan.cuh) instead of the one-kernel form
    cudaStream_t stream = nullptr;
    int64_t n = 0;
    int64_t launches = 0;
    int64_t optimistic_redos = 0;       // mcl_step ticks that were run again because their pre-pass found new ray directions
    double inj_slow = 0, inj_fast = 0;      // adaptiveInjection (MC:191)

private:
    struct EmaArgs { double a_slow, a_fast; };     // mcl_step: the total's accumulation also advances the injection state
    int fail(int code, const std::string& what);
    int resolve_trig();                     // cfg.trig_mode -> trig_kind (probes the host libm for MCL_TRIG_LIBM)
    int trig_kind = 2;                      // TRIG_GLIBC_FMA / TRIG_GLIBC_SSE2 / TRIG_CR (mcl_device.cuh)
    int cuda_fail(cudaError_t e, const char* where);
    int ensure_particles(int64_t count);
    int ref_prepare_beams(const float* ranges, int n_beams, float angle_min, float angle_inc, float range_min, float range_max,
                          std::vector<HostBeam>& all, std::vector<RefBeam>& used);
    int ref_run_update(const RefBeam* d_used, const RefBeam* h_used, int n_used, const std::vector<HostBeam>& all, double* total, bool defer_sync = false,
                       const EmaArgs* ema = nullptr);
    int ref_resample(int jitter_state, const mcl_resample_draws* d, mcl_resample_stats* st, bool front_done = false, bool dev_ema = false);
    int inj_sync_to_device();
    DevBuf<double> d_inj;                   // {weight_slow, weight_fast, p_inject, cdf_is_monotone, total}: adaptive injection on the device (mcl_step)
    bool inj_on_device = false;             // who advanced the injection state last
    int ref_resample_front();               // normalise + CDF + guide table: needs nothing from the host
    int estimate_enqueue(double* h_sums4, RefStepReport* step_report = nullptr);
    // whole-step entry (mcl_step / mcl_step_staged): scalars the host needs travel through this pinned block
    typedef RefStepReport StepScalars;      // {inj[5], pose[4], counters[4]}: pinned, written by k_pose_sums (zero-copy)
    StepScalars* h_step = nullptr;
    // optimistic tick (ref_run_update): the pre-pass reports into this pinned block, the tick's kernels watch d_counters[6]
    const int* tick_abort = nullptr;         // non-null while an mcl_step tick whose caller waits for the report is being enqueued
    bool tick_optimistic = false;            // the tick being enqueued has an unwaited pre-pass in front of it
    bool touch_clean = false;                // d_touch holds no first touchers (k_ref_touch_report leaves it that way)
    struct RefTouchReport* h_touch_report = nullptr;
    unsigned long long* h_touch_keys = nullptr;
    float* h_touch_theta = nullptr;
    int h_touch_cap = 0;
    unsigned long long touch_seq = 0;
    int ensure_touch_block();
    bool scans_are_fused() const;
    int ref_fill_ray_lut_from(const std::vector<HostBeam>& all, const unsigned long long* touch, const float* theta);
    std::vector<RefBeam> step_used;          // scored beams of the tick being enqueued (host scan)
    unsigned long long step_seq = 0;         // ticks enqueued with a report; the report carries the number of the tick that wrote it
    bool guide_built = false;
    int guide_buckets = 0;
    int ref_fill_ray_lut(const std::vector<HostBeam>& all);
    void philox_host(uint32_t stream_id, uint64_t index, uint32_t out[4]) const;

    bool opened = false;
    bool attr_set = false, attr_set2 = false;
    bool profiling = false;
    bool use_pdl = true;                    // programmatic dependent launch between the kernels of a tick (MCL_PDL=0: off)
    bool pdl_hold = false;                  // the next LAUNCH_PDL is an ordinary launch (set after a kernel that waits for other shards)
    struct ProfEvent { int id; cudaEvent_t a, b; };
    std::vector<ProfEvent> prof_events;
    double prof_ms[K_COUNT] = {0};
    int64_t prof_count[K_COUNT] = {0};
    void prof_begin(int id);
    void prof_end();
    void prof_collect();
    int last_per = 3;
    // particles: ping-pong float4 {x,y,theta,w} (= column-major 4xN of the reference)
    DevBuf<float4> part[2];
    int cur = 0;
    DevBuf<double> cdf;
    DevBuf<float> d_wraw, d_wn;       // dense raw / normalised weights (4 B/particle streams for the scans)
    // exact-scan workspace
    DevBuf<double> xs_tsum, xs_toff, xs_seq_s;
    DevBuf<unsigned char> xs_tiles, xs_entries, xs_carry;
    DevBuf<int> xs_seq_base, xs_flag;
    int xs_tiles_cap = 0;
    // one-kernel form (exact_scan_fused.cuh): per-tile flags stamped with the launch's epoch, ticket / finished / fall-back words
    DevBuf<unsigned long long> xs_pub;                           // one 128-byte line of published words per tile
    DevBuf<unsigned char> xs_blocks;                             // SEQ blocks of the one-kernel form: 64 x 80 bytes per tile
    DevBuf<unsigned> xs_counters;
    DevBuf<unsigned long long> xs_trace;                         // stage time stamps of the last debug_exact_scan (mcl_debug_exact_scan_trace)
    bool xs_trace_on = false;
    unsigned xs_epoch = 0;
    bool ema_in_total = false;                                   // ... and did so for the tick being enqueued
    int ensure_xs(int64_t count);
    int exact_accumulate(bool normalise, double* d_total_out, const EmaArgs* ema = nullptr, int guide_buckets_wanted = 0);   // total of d_wraw, or normalise + CDF
    int exact_accumulate_on(const float* w, bool normalise, bool want_cdf, double* d_total_out, const EmaArgs* ema = nullptr, int guide_buckets_wanted = 0);
    DevBuf<int> ancestors;
    // map
    bool map_ready = false;
    int map_w = 0, map_h = 0;
    float res_f = 0.f;
    double origin_x = 0, origin_y = 0, max_x = 0, max_y = 0;
    std::vector<int8_t> h_occ;
    DevBuf<uint8_t> d_occ;            // 1 = occupied (value > 50)
    DevBuf<uint8_t> d_occ_pad;        // bordered ray-march table (0 free, 1 occupied, 2 outside)
    int occ_pad = 0, occ_wp = 0;
    // tables
    GaussTable gauss;
    DevBuf<double> d_gauss;
    std::vector<double> h_radii;
    DevBuf<double> d_radii;
    int key_min = 0, n_keys = 0, n_unfilled = 0;
    std::vector<double2> h_lut;
    std::vector<uint8_t> h_lut_filled;
    DevBuf<double2> d_lut;
    DevBuf<uint8_t> d_lut_filled;
    DevBuf<unsigned long long> d_touch;
    DevBuf<float> d_touch_theta;
    // per-step scan
    std::vector<HostBeam> beams_all;
    DevBuf<RefBeam> d_beams;
    struct StagedScan { std::vector<HostBeam> all; DevBuf<RefBeam> d_used; int n_used = 0; bool valid = false; };
    std::vector<StagedScan> staged;
    // motion / odometry
    OdometryState odo;
    uint64_t step_counter = 0;
    // resample
    DevBuf<double> d_u_r, d_u_jit, d_inj_f64;   // inj_f64: [u_yaw | u_dx | u_dy] x max_inject
    DevBuf<int> d_inj_i32;                      // [row | col] x max_inject
    DevBuf<int> d_block_counts;
    DevBuf<int> d_counters;                     // [0] injected, [1] clamped, [2] flagged total
    DevBuf<double> d_scalars;                   // [0] total weight, [1] wsum, [2..5] pose sums
    DevBuf<double> d_partials;
    double last_total = 0;
    bool have_weights = false;
    bool wsum_known = false;                    // sum of the particle weights known on the host (skips a reduction pass in estimate)
    double known_wsum = 0;
    bool draws_generated = false;               // the last resample generated its draws in-kernel (Philox)
    uint32_t draws_step = 0;
    // NS state
    int ns_build_field();
    int ns_init(int64_t count);
    int ns_predict(const Motion& clean, bool prep_step = false);      // prep_step: also reset the step's accumulators (mcl_ns_step)
    void ns_scan_shape(bool& two_pass, int& nt, int& ng) const;
    bool ns_maxbits_prepped = false, ns_scan_prepped = false;
    double* h_ns_pose = nullptr;            // pinned: the single-shard step's five pose sums, written by k_ns_pose_reduce
    void ns_prepare_beams(const float* ranges, int n_beams, float angle_min, float angle_inc, float range_min, float range_max,
                          std::vector<float2>& pts) const;
    int ns_run_update(const float2* d_pts, int n_pts, float* local_max);
    int ns_launch_update(const float2* d_pts, int n_pts);
    int ns_launch_weights();
    int ns_materialise_weights();
    int ns_launch_resample(uint32_t u0);
    bool ns_w_in_records = false;
    DevBuf<unsigned char> d_bounds;
    void* comm = nullptr;                 // ncclComm_t
    DevBuf<uint64_t> d_totals;
    DevBuf<unsigned char> d_plan;
    DevBuf<double> d_pose;
    DevBuf<int> d_bar;
    static constexpr int RING = 8;
    void* ring_base = nullptr; size_t ring_bytes = 0; int ring_pos = 0;
    std::vector<cudaEvent_t> ring_events;
    int ensure_pinned_ring(size_t bytes);
    void* pinned_ring_next();
    struct NsStagedScan { DevBuf<float2> d_pts; int n = 0; bool valid = false; };
    std::vector<NsStagedScan> ns_staged;
    DevBuf<float> d_lf, d_lf_table, d_ll, d_codes;
    DevBuf<uint8_t> d_lf8, d_code_of_d2;      // the field as one-byte codes (NS_FIELD_U8), same bordered layout
    int ns_n_codes = 0;                        // 0: more than 256 attainable squared distances, no code field
    // fields too large for L2 as fp32: the sensor model alternates between the fp32 and the coded form on measured time
    // (spread-out particles favour the small coded field, converged ones the cheaper fp32 lookup; same values either way)
    cudaEvent_t ns_tune_ev[2] = {nullptr, nullptr};
    bool ns_tune_pending = false;
    int ns_tune_kind = -1, ns_tune_beams = 1;
    int64_t ns_tune_launches = 0;
    double ns_tune_cost[3] = {-1.0, -1.0, -1.0};   // last measured ms per beam, by NS_FIELD_*

    DevBuf<uint16_t> d_d2, d_g;
    DevBuf<float2> d_ns_beams;
    DevBuf<uint64_t> d_prefix, d_tile_sums, d_u64;      // d_u64[0] = local total
    DevBuf<int> d_maxbits;
    int ns_R = 0, ns_beams_n = 0;
    size_t lf_bytes_padded = 0;
    int lf_pad = 0, lf_wp = 0, lf_hp = 0;      // bordered field: border width, row pitch, rows
    float lf_out = 0.f, ns_last_max = 0.f;
    bool ns_attr_set = false;
    bool ns_have_ll = false;
    void* peer_ptr[4][8] = {{nullptr}};    // [0],[1]: particle ping-pong buffers of shard r; [2]: ancestors; [3]: mailbox
    bool peer_ipc[4][8] = {{false}};
    DevBuf<int> d_guide;                   // REF resampling: guide table of the CDF search
    DevBuf<unsigned char> d_mbox;          // this shard's NsMailbox (peer-memory exchange)
    uint32_t xchg_seq = 0;                 // exchange tag: sharded steps taken by this handle (never reset; same on every shard)
    int ensure_mailbox();
    int ns_preload_update(int kind, int pack);
    bool peers_have_mailboxes() const;
    // pinned staging
    void* h_pinned = nullptr;
    size_t h_pinned_bytes = 0;
    int ensure_pinned(size_t bytes);
};

}  // namespace mcl
