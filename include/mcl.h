/* mcl.h — C-ABI of the B200 Monte Carlo localisation engine (libmcl_b200.so).
 *
 * Drop-in boundary for the particle-filter hot path of Bright8787/MonteCarloLocalisation,
 * pink_fundamentals/src/monte_carlo.cpp ("MC"). The reference has no FFI/plugin layer: the boundary is the
 * set of call sites of its free functions in the ROS node shell. Each entry point below names
 * the reference function / call site it replaces. INTEGRATION.md shows the few lines a maintainer changes in
 * monte_carlo.cpp to bind them.
 *
 * Conventions
 *   - plain pointers and sizes only; the engine owns all device memory behind an opaque handle; the caller owns
 *     every host pointer, which is read/written only during the call.
 *   - particles cross the boundary as float[4*N], column-major 4xN = {x,y,theta,w} per particle: the exact
 *     layout of the reference's Eigen::MatrixXf particles(4,N).data() (MC:82,417), so
 *     Eigen::Map<Eigen::MatrixXf>(buf,4,N) works in the node.
 *   - every function returns 0 on success, <0 on error (mcl_last_error gives the text); nothing throws across
 *     the ABI; CUDA/NCCL errors are captured, not aborted on. There is no CPU fallback: without a usable
 *     CUDA device mcl_create fails with MCL_ERR_CUDA.
 *   - one handle = one GPU = one CUDA stream; calls on a handle must be sequential (the reference is a
 *     single roscpp spinner thread, MC:1212).
 *
 * Two sensor/resampling modes share the interface:
 *   MCL_MODE_REF  results identical to the reference given the same injected draws: step-scalar motion noise,
 *                 ray-march sensor model, multinomial resampling with adaptive injection and jitter.
 *   MCL_MODE_NS   the north-star formulation for scale: per-particle Philox motion noise, likelihood-field
 *                 sensor model over a Euclidean distance transform, fixed-point systematic resampling that is
 *                 bit-identical for 1/2/4/8 GPUs.
 */
#ifndef MCL_B200_H
#define MCL_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mcl_handle mcl_handle;

enum { MCL_MODE_REF = 0, MCL_MODE_NS = 1 };

/* How MCL_MODE_REF evaluates the reference's fp32 cos/sin.
 *   MCL_TRIG_LIBM (default)        bit for bit what the reference BINARY computes on this host: glibc's sinf/cosf (2.28 and
 *                                  later, x86-64), whose FMA and SSE2 builds differ at 34 of the 2^32 arguments; mcl_create
 *                                  probes the host's libm at those arguments and the kernels run that build's operation
 *                                  sequence in f64 (csrc/glibc_trigf.cuh). mcl_create fails if the host's libm is neither.
 *   MCL_TRIG_CORRECTLY_ROUNDED     (float)cos((double)x): the portable definition, independent of any libm; differs from
 *                                  glibc's result for 0.26 % / 0.55 % of arguments by one fp32 ulp. */
enum { MCL_TRIG_LIBM = 0, MCL_TRIG_CORRECTLY_ROUNDED = 1 };

enum {
    MCL_OK = 0,
    MCL_ERR_ARG = -1,      /* bad argument / call order (e.g. update before set_map) */
    MCL_ERR_CUDA = -2,     /* CUDA runtime error, or no CUDA device */
    MCL_ERR_IO = -3,       /* map.txt could not be read or parsed */
    MCL_ERR_COMM = -4,     /* NCCL / peer-memory error */
    MCL_ERR_STATE = -5     /* feature not available in this mode */
};

/* Every literal the reference hard-codes on the path, as one POD. mcl_config_default() fills in the
 * reference's values; the synthetic large-map configs override a few. */
typedef struct mcl_config {
    int32_t device;              /* CUDA device ordinal */
    int32_t mode;                /* MCL_MODE_REF | MCL_MODE_NS */
    int64_t max_particles;       /* capacity of the particle buffers on this GPU */
    /* sensor model, MC:627-631, 180-181, 635, 650, 370, 333 */
    double sigma_hit;            /* 0.1   Gaussian LUT sigma (MC:177) */
    double max_laser_range;      /* 1.0   (MC:628) */
    double laser_offset;         /* 0.1   (MC:631) */
    double w_hit, w_rand;        /* 0.8, 0.2 (MC:180-181) */
    double fov_lower_deg;        /* -120  (MC:635) */
    double fov_upper_deg;        /*  120 */
    int32_t beam_stride;         /* 20    (MC:650) */
    int32_t _pad0;
    double ray_step;             /* 0.1   (MC:370) */
    double validity_offset;      /* 0.1   (MC:333) */
    /* odometry noise, MC:1198, and robot geometry, PID_lib.hpp:19-20 */
    double alpha[4];             /* 0.001, 0.001, 0.0001, 0.0001 */
    double wheel_size;           /* 0.0620 */
    double wheel_space;          /* 0.265 */
    /* sampleParticles, MC:422, 396, 431, 442-443 */
    int32_t cell_size_px;        /* 8 */
    int32_t _pad1;
    double cell_meters;          /* 0.8 */
    double init_offset;          /* 0.2  uniform half-width inside a cell */
    double init_shift;           /* 0.05 global offset */
    /* resampleParticles, MC:473-482, 537-546 */
    double inject_max_lost, inject_alpha_slow_lost, inject_alpha_fast_lost;   /* 200, 0.05, 0.5 */
    double inject_max_conf, inject_alpha_slow_conf, inject_alpha_fast_conf;   /* 50, 0.02, 2 */
    double jitter_xy_lost, jitter_theta_lost;                                 /* 0.05, pi/12 */
    double jitter_xy_conf;                                                    /* 0.01 */
    /* production draws (used when a call is given no injected draws) and NS mode */
    uint64_t seed;               /* Philox4x32-10 key */
    double ns_sigma_hit;         /* likelihood-field Gaussian sigma (m) */
    double ns_z_hit, ns_z_rand;  /* mixture weights */
    double ns_max_range;         /* beams at or beyond this range are skipped */
    int32_t ns_beam_stride;      /* 1 = score every beam */
    int32_t ns_use_fov;          /* 0 = full circle, 1 = apply fov_lower/upper like the reference */
    double ns_temper;            /* weight = exp(ns_temper * (loglik - max loglik)); 1 = plain product of beam likelihoods */
    /* confidence estimate, MC:889-890, 933 */
    double kmeans_radius;        /* 0.4: density radius around the best cluster (countParticlesNearCluster) */
    /* float trig of the path: cosf/sinf at MC:644-645 (laser origin) and Eigen's fp32 cos/sin at MC:747-748 (predict) */
    int32_t trig_mode;           /* MCL_TRIG_LIBM | MCL_TRIG_CORRECTLY_ROUNDED */
    int32_t _pad2;
} mcl_config;

/* Named draws for sampleParticles (MC:427-440): per particle yaw~U[0,1) canonical, row, col, dx, dy canonical. */
typedef struct mcl_init_draws {
    const double* u_yaw;
    const int32_t* row;
    const int32_t* col;
    const double* u_dx;
    const double* u_dy;
} mcl_init_draws;

/* Draws consumed by resampleParticles (MC:508-555), as sequential streams like the reference's engines:
 *   u_r[N]      one canonical draw per output slot (MC:514)
 *   u_jitter[]  consumed in slot order by non-injected slots: x, y, and theta when jitterState (MC:537-546)
 *   inject      named sampleParticles(1) draws, consumed in order by injected slots (MC:520); n_inject entries */
typedef struct mcl_resample_draws {
    const double* u_r;
    const double* u_jitter;
    int64_t n_jitter;
    mcl_init_draws inject;
    int32_t n_inject;
} mcl_resample_draws;

typedef struct mcl_resample_stats {
    int32_t injected;            /* value logged at MC:559 */
    int32_t clamped;             /* slots whose lower_bound ran off the end (UB in the reference), clamped to N-1 */
    double p_inject;             /* MC:492 */
    double weight_slow;          /* MC:487 */
    double weight_fast;          /* MC:488 */
    double total_weight;         /* computeWeight's return, MC:681 */
} mcl_resample_stats;

/* ---- lifecycle --------------------------------------------------------------------------------------- */
void mcl_config_default(mcl_config* cfg);
int mcl_create(const mcl_config* cfg, mcl_handle** out);
void mcl_destroy(mcl_handle* h);
const char* mcl_last_error(mcl_handle* h);       /* h may be NULL: last mcl_create error */
const char* mcl_version(void);

/* ---- map: replaces mapCallback's stored grid (MC:291-295) and the map.txt pipeline
 *      (publish_map.py:8-33 -> publish_map_rviz.cpp:306-437) ---------------------------------------------- */
int mcl_rasterise_map_txt(const char* text, int8_t* out, int64_t cap, int32_t* width, int32_t* height);  /* host only */
int mcl_set_map(mcl_handle* h, const int8_t* occ, int32_t width, int32_t height, float resolution,
                double origin_x, double origin_y);
int mcl_load_map_txt(mcl_handle* h, const char* path);    /* resolution 0.1f, origin (0,0) as RV:420-426 */
int mcl_precompute_ray_directions(mcl_handle* h, double min_deg, double max_deg, double step_deg);  /* MC:1017-1023, 1199 */

/* ---- particles --------------------------------------------------------------------------------------- */
int mcl_init(mcl_handle* h, int64_t n, const mcl_init_draws* draws);   /* sampleParticles(N), MC:415-450, 1205 */
int mcl_upload(mcl_handle* h, const float* particles4xn, int64_t n);   /* checkpoint/resume; tests */
int mcl_download(mcl_handle* h, float* particles4xn);                  /* publishParticles' input, MC:563-579 */
int64_t mcl_num_particles(mcl_handle* h);

/* ---- predict: diffDriveModel + sampleMotionModelOdometry + updateParticlePos (MC:695-755, 1084-1086) -- */
/* z3 = three standard-normal draws in call order (rot1, trans, rot2) or NULL to draw from Philox.
 * motion_out3 (optional) receives the noised (rot_1, trans, rot_2) that was applied. */
int mcl_predict_encoders(mcl_handle* h, double encoder_left, double encoder_right, const double* z3, double* motion_out3);
int mcl_predict_motion(mcl_handle* h, double rot_1, double trans, double rot_2);   /* updateParticlePos with a given motionModel */

/* ---- update: computeWeight(particles, latest_scan) (MC:623-682) --------------------------------------- */
int mcl_update(mcl_handle* h, const float* ranges, int32_t n_beams, float angle_min, float angle_increment,
               float range_min, float range_max, double* total_weight);

/* The same with the scan pre-processed and parked in device memory (slot in [0,4096)): replaying recorded scans,
 * and benchmarks that want the inputs resident in HBM before the timed region. Slot lifetime: a slot may be re-staged at
 * any time; mcl_scan_stage first waits for every tick already queued on the handle (mcl_step_staged / mcl_ns_step_staged
 * return before their ticks ran), so a queued tick always reads the scan that was in its slot when it was queued. */
int mcl_scan_stage(mcl_handle* h, int32_t slot, const float* ranges, int32_t n_beams, float angle_min, float angle_increment,
                   float range_min, float range_max);
int mcl_update_staged(mcl_handle* h, int32_t slot, double* total_weight);

/* ---- resample: the rest of resampleParticles(particles, jitterState) (MC:469-561). Must follow mcl_update. */
int mcl_resample(mcl_handle* h, int32_t jitter_state, const mcl_resample_draws* draws, mcl_resample_stats* stats);
int mcl_download_ancestors(mcl_handle* h, int32_t* idx);   /* ancestor index per output slot, -1 = injected */
int mcl_download_cdf(mcl_handle* h, double* cdf);          /* REF: the f64 CDF of MC:496-505 */

/* ---- estimate: estimateWeightedPose(particles) (MC:782-800) -------------------------------------------
 * Tolerance-graded, not bit-exact: the reference reduces in fp32 through Eigen packets (vendored without Eigen/Core, so
 * its summation order cannot be pinned); the engine does the fp32 element math (w/weight_sum, w*x, w*sin, w*cos) and
 * accumulates in f64 in a fixed order, with weight_sum taken from the known f64 total. Agreement with the reference is
 * 1e-5 relative, the bound north_star states for poses; the value feeds nothing downstream on the path. */
int mcl_estimate(mcl_handle* h, double* x, double* y, double* theta);

/* ---- one tick of the node's loop: executeParticleFilter (MC:1084-1092) = diffDriveModel + updateParticlePos,
 *      resampleParticles (computeWeight inside, MC:468), estimateWeightedPose --------------------------------
 * The same kernels and results as mcl_predict_encoders + mcl_update + mcl_resample + mcl_estimate with the engine's own
 * draw streams, enqueued as one piece with no host decision in the middle: the adaptive-injection state (MC:469-492) is
 * advanced on the device from the weight total, with the same IEEE operations as the host form. With pose3 and stats both
 * NULL the call returns as soon as the tick is queued (the estimate is still computed; several ticks may be in flight;
 * any call that returns data synchronises); otherwise it waits for the GPU once, at the end, instead of three times. MCL_MODE_REF only (NS filters:
 * mcl_ns_step). pose3 = {x, y, theta} of the resampled particles; stats as mcl_resample. The per-function calls and
 * mcl_step can be mixed on one handle. mcl_step_staged takes a scan parked by mcl_scan_stage.
 * What crosses the bus per tick: the scored beams in (24 B each: up to 40 of them ride in the computeWeight kernel's launch
 * parameters, longer lists take one copy command at the head of the tick; nothing with a staged scan) and a 104-byte report
 * out (pose sums, injection state, counters, the tick's sequence number), which the tick's last kernel stores straight into
 * the engine's pinned host block; a call that returns data returns when that sequence number lands. No other copy or memset
 * command is enqueued, and the tick's kernels follow one another by programmatic dependent launch (MCL_PDL=0 in the
 * environment: ordinary stream order). While the ray-direction table is incomplete (first tick; clouds with a narrow spread
 * of headings) a call that returns data enqueues the tick behind the first-touch pre-pass without waiting for it, and runs
 * the tick a second time only when the pre-pass found a direction the host has to evaluate (same results either way).
 * The engine's own draw streams (draws == NULL / mcl_step), all Philox4x32-10 keyed by mcl_config.seed with counter
 * (index lo, index hi, stream, tick number): stream 0x30 = u_r and the jitter draws of slot i (indices 2i, 2i+1; readable
 * through mcl_debug_download_resample_draws), stream 0x31 = the named draws of the rank-th injected particle (indices
 * 2*rank, 2*rank+1: u_yaw, u_dx, u_dy as 53-bit fractions, row and col as words mod the coarse grid), generated inside the
 * resampling kernel; stream 0x20 = the three odometry normals (host); stream 0x10 = mcl_init without draws. */
int mcl_step(mcl_handle* h, double enc_left, double enc_right, const float* ranges, int32_t n_beams, float angle_min,
             float angle_increment, float range_min, float range_max, int32_t jitter_state, double* pose3, mcl_resample_stats* stats);
int mcl_step_staged(mcl_handle* h, double enc_left, double enc_right, int32_t slot, int32_t jitter_state, double* pose3,
                    mcl_resample_stats* stats);

/* ---- NS mode across GPUs: particles shard by contiguous global index; one handle per GPU. A driver sequences the
 *      *_local phases around three tiny collectives (max of the local maxima; all-gather of the local totals; a barrier),
 *      which keeps systematic resampling globally exact and bit-identical for any GPU count. Each shard stores its
 *      resampled particles straight into the shard that owns the output slot (peer memory over NVLink), so the
 *      rebalance is fused into the resampling kernel. With world == 1, mcl_update / mcl_resample do all of this. ---- */
int mcl_ns_set_shard(mcl_handle* h, int32_t rank, int32_t world, int64_t n_global);   /* allocates this shard */
int mcl_ns_update_local(mcl_handle* h, const float* ranges, int32_t n_beams, float angle_min, float angle_increment,
                        float range_min, float range_max, float* local_max_loglik);
int mcl_ns_update_local_staged(mcl_handle* h, int32_t slot, float* local_max_loglik);   /* scan parked by mcl_scan_stage */
int mcl_ns_weights_local(mcl_handle* h, float global_max_loglik, uint64_t* local_total_q32);
int mcl_ns_resample_local(mcl_handle* h, uint64_t offset_q32, uint64_t total_q32, uint32_t u0, int64_t* k_lo, int64_t* k_hi);
int mcl_ns_end_step(mcl_handle* h);                   /* after every shard finished resampling: swap buffers */
uint32_t mcl_ns_u0(mcl_handle* h);                    /* this step's systematic offset (same on every shard) */
int mcl_ns_pose_partials(mcl_handle* h, double* out5);/* {sum w, sum w x, sum w y, sum w sin, sum w cos} of this shard */
/* The same step driven entirely by the engine: NCCL (loaded at run time) on the handle's stream, resampling plan computed
 * on the device, no host round trip. mcl_comm_unique_id on one shard -> hand the 128 bytes to every shard (any transport)
 * -> mcl_comm_init on every shard (collective; also maps all peers' buffers through CUDA IPC) -> mcl_ns_step per tick on
 * every shard. world == 1 needs no communicator. pose3 (optional) = {x, y, theta} weighted mean before resampling. */
int mcl_comm_unique_id(mcl_handle* h, void* out128);
int mcl_comm_init(mcl_handle* h, const void* id128);
int mcl_ns_step(mcl_handle* h, double rot_1, double trans, double rot_2, const float* ranges, int32_t n_beams, float angle_min,
                float angle_increment, float range_min, float range_max, double* pose3);
int mcl_ns_step_staged(mcl_handle* h, double rot_1, double trans, double rot_2, int32_t slot, double* pose3);
/* host-only planning helpers (no GPU touched) */
int mcl_ns_first_slot(uint64_t offset_q32, uint64_t total_q32, uint64_t n_global, uint32_t u0, int64_t* slot);
int mcl_ns_shard_range(int64_t n_global, int32_t world, int32_t rank, int64_t* begin, int64_t* count, int64_t* per_rank);
/* How mcl_ns_step does the step's collectives (all-reduce of the maximum log-likelihood, all-gather of the Q32 totals,
 * pose all-reduce, closing barrier): 1 = peer-memory mailboxes (each shard stores {payload, tag} straight into the other
 * shards' mailboxes over NVLink and polls its own; three 32-thread kernels, no NCCL on the data path; the default once
 * every peer's mailbox is mapped, which mcl_comm_init does), 0 = NCCL collectives on the handle's stream, -1 = default.
 * Results are identical. Environment override for the default: MCL_NS_EXCHANGE=nccl|peer. Mode 1 expects the shards'
 * streams to make progress independently: one process per GPU (the intended deployment), or in one process at most as many
 * shards as the device has hardware queues (CUDA_DEVICE_MAX_CONNECTIONS, 8 by default) and scans parked with
 * mcl_scan_stage so that enqueueing a step never waits for the device. Polls give up after MCL_NS_EXCHANGE_TIMEOUT_S
 * seconds (30; 0 = never) with MCL_ERR_COMM.
 * mcl_ns_exchange_used: what the last sharded step used (-1: none yet). */
int mcl_ns_set_exchange(mcl_handle* h, int32_t mode);
int mcl_ns_exchange_used(mcl_handle* h);
/* peer memory. which: 0/1 = the two particle buffers, 2 = ancestors, 3 = the exchange mailbox. export/import move a
 * 64-byte CUDA IPC handle between processes; mcl_peer_set wires two handles of one process (tests, single-process
 * multi-GPU). */
int mcl_peer_export(mcl_handle* h, int32_t which, void* out64);
int mcl_peer_import(mcl_handle* h, int32_t rank, int32_t which, const void* in64);
int mcl_peer_set(mcl_handle* h, int32_t rank, int32_t which, void* device_ptr);
void* mcl_device_buffer(mcl_handle* h, int32_t which);
/* inspection */
int mcl_ns_download_field(mcl_handle* h, float* loglik_field, uint16_t* d2);
int mcl_ns_download_loglik(mcl_handle* h, float* loglik);
int mcl_ns_download_prefix(mcl_handle* h, uint64_t* prefix_q32);
/* Where the last sensor-model launch read the likelihood field from: 0 = fp32 field staged into shared memory by TMA,
 * 1 = fp32 field through L1/L2, 2 = one-byte coded field through L1/L2 + shared-memory code table (fields too large for
 * L2 as fp32; chosen by measured time against form 1). All three give identical values. -1 before the first update. */
int mcl_ns_field_form(mcl_handle* h);

/* ---- the rows either side of the hot path ------------------------------------------------- */
#define MCL_KMEANS_K 3
#define MCL_KMEANS_EXACT_MAX 65536   /* up to this many particles the centre sums are accumulated sequentially in fp32 like
                                        the reference's loop (MC:851-856): results bit-exact with the CPU */
typedef struct mcl_kmeans_result {
    double ratio;                /* return value of isLocalizationLost_densitiy_cluster (MC:933, 948) */
    double x_best, y_best, theta_best;        /* globals x_best.. (MC:934-940); all -1 when ratio <= threshold */
    double cluster_weight[MCL_KMEANS_K];      /* MC:903-907 */
    float centers[2 * MCL_KMEANS_K];          /* x0,y0,x1,y1,x2,y2 */
    int64_t counts[MCL_KMEANS_K];
    int32_t best_cluster, passes, reinit_used, exact;
} mcl_kmeans_result;
/* replaces isLocalizationLost_densitiy_cluster(particles, cluster_distance, cluster_ratio_threshold) (MC:886-949, call
 * site MC:1090) incl. kMeansClustering (MC:802-868, K = 3, <= 20 iterations) and countParticlesNearCluster (MC:869-884).
 * The reference seeds srand(time) and draws rand() % N for the 3 initial centres and for every emptied cluster: pass those
 * indices (init_idx[3], reinit_idx[n_reinit], consumption order) for reproducible/parity runs, or NULL to draw them from
 * the handle's Philox stream. cluster_distance is unused by the reference and therefore absent. */
int mcl_kmeans_confidence(mcl_handle* h, const int32_t* init_idx, const int32_t* reinit_idx, int32_t n_reinit,
                          double cluster_ratio_threshold, mcl_kmeans_result* out);
int mcl_download_assignments(mcl_handle* h, int32_t* cluster_of_particle);     /* assignments of the last call */
/* replaces publishPosMsg (MC:958-994): world pose -> msg/Pose.msg fields {row, column, orientation}; RIGHT=0 UP=1 LEFT=2
 * DOWN=3; all -1 for the "not localised" sentinel (wx < 0 or wy < 0). Host only. cell_meters = 0.8 in the reference. */
int mcl_pose_to_cell(double wx, double wy, double angle, double cell_meters, int32_t* row, int32_t* column, int32_t* orientation);
/* replaces publishExactPose (MC:995-1008): msg/ExactPose.msg float32 fields {x, y, theta}. Host only. */
int mcl_exact_pose(double x, double y, double theta, float* out3);
/* replaces the loop of publishParticles (MC:563-579) for particles first, first+stride, ...: out[4*j..] =
 * {position.x, position.y, orientation.z, orientation.w} (position.z = orientation.x = orientation.y = 0), quaternions
 * built on the device; stride > 1 streams a subsample instead of 16 B x N every tick. */
int mcl_download_pose_array(mcl_handle* h, int64_t first, int64_t stride, int64_t count, double* out);
/* config presets: "reference" (= mcl_config_default) or "playground" = the knobs of the earlier, unbuilt variant
 * src/playground.cpp that the config can express (ray step 0.05 :328, every 3rd beam :604, FOV +-90 :601). */
int mcl_config_preset(mcl_config* cfg, const char* name);

/* ---- state the reference keeps in globals (MC:189-192), for checkpoint/resume and tests ---------------- */
int mcl_get_injection_state(mcl_handle* h, double* weight_slow, double* weight_fast);
int mcl_set_injection_state(mcl_handle* h, double weight_slow, double weight_fast);
int mcl_get_ray_lut(mcl_handle* h, int32_t* keys, double* dx, double* dy, int32_t cap, int32_t* count);

/* ---- handle plumbing (cross-check switches, micro-benchmarks and per-kernel timers live in mcl_debug.h) ---- */
void* mcl_stream(mcl_handle* h);                  /* the cudaStream_t all kernels of this handle run on */
int mcl_synchronize(mcl_handle* h);
int64_t mcl_kernel_launches(mcl_handle* h);       /* kernels launched by this handle so far */

#ifdef __cplusplus
}
#endif
#endif /* MCL_B200_H */
