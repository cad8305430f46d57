// ref_harness.cpp — TEST INFRASTRUCTURE. Builds oracle/_ref/libmclref.so: the UNMODIFIED reference
// translation unit /root/reference/pink_fundamentals/src/monte_carlo.cpp (included where it lies, never
// copied) behind the stub headers in oracle/shim/, plus a small C API that drives the reference's own free
// functions and globals. Used to pin oracle/mcl_oracle.cpp and to generate tests/golden/ fixtures.
// Never loaded by the product.
#define main mcl_ref_main
#include MCL_REF_SOURCE
#undef main

#include <cstring>
#include <algorithm>

namespace {
void reseed_static_engines(unsigned s) { for (auto* e : mclshim::engines()) e->seed(s); }
}

extern "C" {

// ---- RNG control -------------------------------------------------------------------------------
void ref_push_seed(unsigned s) { mclshim::seeds().q.push_back(s); }
void ref_clear_seeds() { mclshim::seeds().q.clear(); mclshim::seeds().history.clear(); }
// The two function-local static engines (MC:411 `sample`, MC:452 `uniformJitter`) are constructed on
// first use from the seed queue; force both into existence, then reseed them individually.
void ref_touch_static_engines() { (void)sample(1.0); (void)uniformJitter(0.0, 1.0); }
int ref_num_static_engines() { return (int)mclshim::engines().size(); }
void ref_seed_static_engine(int which, unsigned s) { mclshim::engines().at(which)->seed(s); }

// Canonical draws exactly as the reference's distributions consume them, for building injected-draw
// arrays: mt19937 + uniform_real_distribution<>(0,1) (MC:509-514).
void ref_stream_mt_canonical(unsigned seed, int n, double* out) {
    std::mt19937 gen(seed);
    std::uniform_real_distribution<> dis(0.0, 1.0);
    for (int i = 0; i < n; i++) out[i] = dis(gen);
}
// minstd_rand0 canonical draws (uniformJitter builds a fresh distribution per call; a (0,1) distribution
// returns the canonical value itself).
void ref_stream_minstd_canonical(unsigned seed, int n, double* out) {
    std::minstd_rand0 gen(seed);
    for (int i = 0; i < n; i++) { std::uniform_real_distribution<double> d(0.0, 1.0); out[i] = d(gen); }
}
// minstd_rand0 standard normals, one fresh normal_distribution per draw like `sample` (MC:412).
void ref_stream_minstd_normal(unsigned seed, int n, double* out) {
    std::minstd_rand0 gen(seed);
    for (int i = 0; i < n; i++) { std::normal_distribution<double> d(0.0, 1.0); out[i] = d(gen); }
}
// The named draws of one sampleParticles(1) call (MC:427-440) for a given mt19937 seed, using the SAME call
// shape so this compiler picks the same argument evaluation order as in the reference TU (Q21).
static void named_draws_sink(int row, int col, double dx, double dy, int* r, int* c, double* x, double* y) { *r = row; *c = col; *x = dx; *y = dy; }
void ref_named_sample_draws(unsigned seed, int n_rows, int n_cols, int count, double* u_yaw, int* row, int* col, double* u_dx, double* u_dy) {
    std::mt19937 gen(seed);
    std::uniform_int_distribution<> cell_x(0, n_rows - 1);
    std::uniform_int_distribution<> cell_y(0, n_cols - 1);
    std::uniform_real_distribution<> offset(0.0, 1.0);
    std::uniform_real_distribution<> yaw_dist(0.0, 1.0);
    for (int i = 0; i < count; i++) {
        u_yaw[i] = yaw_dist(gen);
        named_draws_sink(cell_x(gen), cell_y(gen), offset(gen), offset(gen), &row[i], &col[i], &u_dx[i], &u_dy[i]);
    }
}

// ---- globals -------------------------------------------------------------------------------------
void ref_set_map(const int8_t* occ, int w, int h, float res, double ox, double oy) {
    auto g = std::make_shared<nav_msgs::OccupancyGrid>();
    g->info.resolution = res; g->info.width = w; g->info.height = h;
    g->info.origin.position.x = ox; g->info.origin.position.y = oy;
    g->data.assign(occ, occ + (size_t)w * h);
    map_msg = g; map_ready = true;
}
void ref_set_scan(const float* ranges, int B, float angle_min, float angle_inc, float range_min, float range_max) {
    latest_scan.ranges.assign(ranges, ranges + B);
    latest_scan.angle_min = angle_min; latest_scan.angle_increment = angle_inc;
    latest_scan.angle_max = angle_min + angle_inc * (B - 1);
    latest_scan.range_min = range_min; latest_scan.range_max = range_max;
}
void ref_reset_state() {
    ray_direction_lookup.clear();
    adaptiveInjection = AdaptiveInjection();
    encoderData = EncoderData{0, 0, 0, 0};
    previous_position = RobotPosition{0, 0, 0};
    current_position = RobotPosition{0, 0, 0};
    motionModel = OdometryModel{0, 0, 0};
    noiseParam = NoiseParameter{0.001, 0.001, 0.0001, 0.0001};   // MC:1198
}
void ref_precompute_ray_directions(double a, double b, double s) { precomputeRayDirections(a, b, s); }
int ref_ray_lut_dump(int lo, int hi, int* keys, double* dx, double* dy, int cap) {
    std::vector<int> ks;
    for (auto& kv : ray_direction_lookup) if (kv.first >= lo && kv.first <= hi) ks.push_back(kv.first);
    std::sort(ks.begin(), ks.end());
    int n = 0;
    for (int k : ks) { if (n < cap) { keys[n] = k; dx[n] = ray_direction_lookup[k].first; dy[n] = ray_direction_lookup[k].second; } n++; }
    return n;
}
void ref_get_injection_state(double* out) { out[0] = adaptiveInjection.weight_slow; out[1] = adaptiveInjection.weight_fast; }
void ref_set_injection_state(double slow, double fast) { adaptiveInjection.weight_slow = slow; adaptiveInjection.weight_fast = fast; }
void ref_set_motion(double r1, double t, double r2) { motionModel = OdometryModel{r1, t, r2}; }

static void load(Eigen::MatrixXf& m, const float* P, int N) { m = Eigen::MatrixXf(4, N); memcpy(m.data(), P, sizeof(float) * 4 * (size_t)N); }
static void store(const Eigen::MatrixXf& m, float* P) { memcpy(P, m.data(), sizeof(float) * 4 * (size_t)m.cols()); }

// ---- the reference's own functions -----------------------------------------------------------------
double ref_gauss_get(double d) { return exp_gauss.get(d); }
int ref_filter_scan(double lower, double upper, double* radius, double* angle, int cap) {
    auto v = filterAngles(filterLaserReadings(latest_scan), lower, upper);
    for (size_t i = 0; i < v.size() && (int)i < cap; i++) { radius[i] = v[i].radius; angle[i] = v[i].angle; }
    return (int)v.size();
}
int ref_is_valid_pos(double x, double y) { return isValidPos(x, y) ? 1 : 0; }
int ref_is_occupied(double x, double y) { return isOccupied(x, y) ? 1 : 0; }
double ref_yaw_roundtrip(double t) { return tf::getYaw(tf::createQuaternionMsgFromYaw(t)); }
double ref_raycast(double x, double y, double theta, double off_deg, double max_range) {
    geometry_msgs::Pose p; p.position.x = x; p.position.y = y; p.orientation = tf::createQuaternionMsgFromYaw(theta);
    geometry_msgs::Point hit;
    return raycast(p, off_deg, hit, max_range);
}
// sampleParticles(N) with the mt19937 seeded from the next queued seed.
void ref_sample_particles(int N, float* P) { Eigen::MatrixXf m = sampleParticles(N); store(m, P); }
// sensorCallback + diffDriveModel; out = noised (rot1, trans, rot2). Static engine 0 (`sample`) supplies noise.
void ref_diff_drive(double enc_left, double enc_right, double* out) {
    encoderData.current_encoderLeft = enc_left; encoderData.current_encoderRight = enc_right;
    diffDriveModel(encoderData, current_position, previous_position);
    out[0] = motionModel.rot_1; out[1] = motionModel.trans; out[2] = motionModel.rot_2;
}
void ref_update_particle_pos(float* P, int N) { Eigen::MatrixXf m; load(m, P, N); updateParticlePos(m); store(m, P); }
double ref_compute_weight(float* P, int N) { Eigen::MatrixXf m; load(m, P, N); double t = computeWeight(m, latest_scan); store(m, P); return t; }
// resampleParticles(particles, jitterState) on the GLOBAL `particles` (as at MC:1089). P is updated with the
// normalised weights; Pout receives the new set. Returns the injected count (captured from the ROS_INFO at MC:559).
int ref_resample(float* P, int N, int jitterState, float* Pout) {
    load(particles, P, N);
    mclshim::log().last_injected = -1;
    Eigen::MatrixXf out = resampleParticles(particles, jitterState != 0);
    store(particles, P);
    store(out, Pout);
    return mclshim::log().last_injected;
}
void ref_estimate_weighted_pose(const float* P, int N, double* out) {
    Eigen::MatrixXf m; load(m, P, N);
    RobotPosition r = estimateWeightedPose(m);
    out[0] = r.x; out[1] = r.y; out[2] = r.theta;
}
// ---- confidence estimate and output adapters ----------------------------------------
void ref_set_time(long t) { mclshim::fake_time() = (time_t)t; }
// isLocalizationLost_densitiy_cluster (MC:886-949) on P; out = {x_best, y_best, theta_best} (the globals it sets)
double ref_kmeans_confidence(const float* P, int N, double cluster_distance, double ratio_threshold, double* out) {
    Eigen::MatrixXf m; load(m, P, N);
    double ratio = isLocalizationLost_densitiy_cluster(m, cluster_distance, ratio_threshold);
    out[0] = x_best; out[1] = y_best; out[2] = theta_best;
    return ratio;
}
// kMeansClustering alone (MC:802-868): assignments[N], centers[2K]
void ref_kmeans(const float* P, int N, int K, int max_iters, int* assignments, float* centers) {
    Eigen::MatrixXf m; load(m, P, N);
    std::vector<int> a; std::vector<std::pair<float, float>> c;
    kMeansClustering(m, K, max_iters, a, c);
    for (int i = 0; i < N; i++) assignments[i] = a[i];
    for (int k = 0; k < K; k++) { centers[2 * k] = c[k].first; centers[2 * k + 1] = c[k].second; }
}
int ref_count_near(const float* P, int N, float x, float y, float radius) { Eigen::MatrixXf m; load(m, P, N); return countParticlesNearCluster(m, x, y, radius); }
// publishPosMsg (MC:958-994): out = {row, column, orientation} of the message it published
void ref_publish_pos_msg(double wx, double wy, double angle, int* out) {
    publishPosMsg(wx, wy, angle);
    const pink_fundamentals::Pose& p = mclshim::last_published<pink_fundamentals::Pose>();
    out[0] = p.row; out[1] = p.column; out[2] = p.orientation;
}
// publishExactPose (MC:995-1008): out = {x, y, theta} as float32 message fields
void ref_publish_exact_pose(double x, double y, double theta, float* out) {
    ros::Publisher pub;
    publishExactPose(x, y, theta, pub);
    const pink_fundamentals::ExactPose& p = mclshim::last_published<pink_fundamentals::ExactPose>();
    out[0] = p.x; out[1] = p.y; out[2] = p.theta;
}
// publishParticles (MC:563-579): out[i] = {x, y, qz, qw} of pose i (qx = qy = 0)
int ref_publish_particles(const float* P, int N, double* out) {
    Eigen::MatrixXf m; load(m, P, N);
    ros::Publisher pub;
    publishParticles(m, pub);
    const geometry_msgs::PoseArray& a = mclshim::last_published<geometry_msgs::PoseArray>();
    for (size_t i = 0; i < a.poses.size(); i++) {
        out[4 * i] = a.poses[i].position.x; out[4 * i + 1] = a.poses[i].position.y;
        out[4 * i + 2] = a.poses[i].orientation.z; out[4 * i + 3] = a.poses[i].orientation.w;
    }
    return (int)a.poses.size();
}
}  // extern "C"
