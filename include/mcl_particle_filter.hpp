// mcl_particle_filter.hpp — C++ host class over the C-ABI (include/mcl.h) for the ROS node in
// pink_fundamentals/src/monte_carlo.cpp ("MC").
//
// The reference has no particle-filter class: the filter is a set of free functions over globals. This header gives
// those functions one home, with the SAME names, argument meaning and call order, so that the swap inside
// executeParticleFilter (MC:1084-1092) and main (MC:1198-1206) is line for line:
//
//     reference global / function                     ->  mcl::ParticleFilter member
//     map_msg (mapCallback, MC:291-295)               ->  setMap(*msg)  / loadMapTxt(path)
//     precomputeRayDirections(-120,120,0.1) MC:1199   ->  precomputeRayDirections(-120,120,0.1)
//     particles = sampleParticles(N)        MC:1205   ->  sampleParticles(N)
//     diffDriveModel(...) + updateParticlePos(...)    ->  diffDriveModel(encoderLeft, encoderRight)
//     particles = resampleParticles(particles, lost)  ->  resampleParticles(latest_scan..., lost)   (computeWeight inside)
//     estimateWeightedPose(particles)       MC:782    ->  estimateWeightedPose()
//     particles (Eigen::MatrixXf 4xN)                 ->  downloadParticles(float*)  (same column-major 4xN layout)
//     isLocalizationLost_densitiy_cluster(...) MC:1090 ->  isLocalizationLost_densitiy_cluster(threshold)  (sets x_best.. below)
//     publishPosMsg / publishExactPose      MC:958-1008 ->  poseMsg(x, y, theta) / exactPoseMsg(x, y, theta)  (message fields)
//     publishParticles(particles, pub)      MC:563-579  ->  poseArray(stride)  (x, y, qz, qw per pose, quaternions built on the GPU)
//
// Header-only, no ROS or Eigen dependency; link with -lmcl_b200. Errors throw std::runtime_error on this side of the ABI.
#pragma once
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "mcl.h"

namespace mcl {

struct RobotPosition { double x, y, theta; };          // MC:106-110
struct OdometryModel { double rot_1, trans, rot_2; };  // MC:112-116

class ParticleFilter {
public:
    explicit ParticleFilter(int device = 0, int mode = MCL_MODE_REF) {
        mcl_config_default(&cfg_);
        cfg_.device = device;
        cfg_.mode = mode;
        open();
    }
    explicit ParticleFilter(const mcl_config& cfg) : cfg_(cfg) { open(); }
    ~ParticleFilter() { mcl_destroy(h_); }
    ParticleFilter(const ParticleFilter&) = delete;
    ParticleFilter& operator=(const ParticleFilter&) = delete;

    // ---- map ----
    // occupancy: int8 row-major data[my*width+mx], occupied iff > 50 (MC:316-327); resolution is the float32 on the wire.
    void setMap(const int8_t* data, int width, int height, float resolution, double origin_x = 0.0, double origin_y = 0.0) {
        check(mcl_set_map(h_, data, width, height, resolution, origin_x, origin_y));
    }
    // Works with nav_msgs::OccupancyGrid without including ROS headers here.
    template <class OccupancyGridMsg>
    void setMap(const OccupancyGridMsg& msg) {
        setMap(msg.data.data(), (int)msg.info.width, (int)msg.info.height, msg.info.resolution, msg.info.origin.position.x,
               msg.info.origin.position.y);
    }
    void loadMapTxt(const std::string& path) { check(mcl_load_map_txt(h_, path.c_str())); }
    void precomputeRayDirections(double min_angle_deg, double max_angle_deg, double step_deg) {
        check(mcl_precompute_ray_directions(h_, min_angle_deg, max_angle_deg, step_deg));
    }

    // ---- particles ----
    void sampleParticles(int num_particles) { check(mcl_init(h_, num_particles, nullptr)); }
    int64_t cols() const { return mcl_num_particles(h_); }
    // dst: 4*cols() floats, column-major 4xN {x,y,theta,w} = Eigen::Map<Eigen::MatrixXf>(dst, 4, cols())
    void downloadParticles(float* dst) { check(mcl_download(h_, dst)); }
    void uploadParticles(const float* src, int64_t n) { check(mcl_upload(h_, src, n)); }

    // ---- predict: diffDriveModel + sampleMotionModelOdometry + updateParticlePos (MC:1084-1086) ----
    OdometryModel diffDriveModel(double current_encoderLeft, double current_encoderRight) {
        double m[3];
        check(mcl_predict_encoders(h_, current_encoderLeft, current_encoderRight, nullptr, m));
        return OdometryModel{m[0], m[1], m[2]};
    }
    void updateParticlePos(const OdometryModel& motionModel) { check(mcl_predict_motion(h_, motionModel.rot_1, motionModel.trans, motionModel.rot_2)); }

    // ---- update + resample ----
    // computeWeight(particles, real_scan) (MC:623-682): returns totalWeight
    double computeWeight(const float* ranges, int n_beams, float angle_min, float angle_increment, float range_min, float range_max) {
        double total = 0;
        check(mcl_update(h_, ranges, n_beams, angle_min, angle_increment, range_min, range_max, &total));
        return total;
    }
    template <class LaserScanMsg>
    double computeWeight(const LaserScanMsg& s) {
        return computeWeight(s.ranges.data(), (int)s.ranges.size(), s.angle_min, s.angle_increment, s.range_min, s.range_max);
    }
    // resampleParticles(particles, jitterState) (MC:457-561), computeWeight included like the reference; returns the
    // injected-particle count the reference logs at MC:559.
    template <class LaserScanMsg>
    int resampleParticles(const LaserScanMsg& latest_scan, bool jitterState) {
        computeWeight(latest_scan);
        mcl_resample_stats st;
        check(mcl_resample(h_, jitterState ? 1 : 0, nullptr, &st));
        last_stats_ = st;
        return st.injected;
    }
    const mcl_resample_stats& lastResampleStats() const { return last_stats_; }

    // ---- estimate ----
    RobotPosition estimateWeightedPose() {
        RobotPosition p;
        check(mcl_estimate(h_, &p.x, &p.y, &p.theta));
        return p;
    }

    // ---- one tick: executeParticleFilter (MC:1084-1092) as a single engine call (mcl_step) ----
    // = diffDriveModel + updateParticlePos + resampleParticles (computeWeight inside) + estimateWeightedPose with one wait for
    // the GPU instead of three; same results as the separate calls. Returns the weighted pose of the resampled particles.
    template <class LaserScanMsg>
    RobotPosition executeParticleFilter(double current_encoderLeft, double current_encoderRight, const LaserScanMsg& latest_scan, bool jitterState) {
        double pose[3];
        mcl_resample_stats st;
        check(mcl_step(h_, current_encoderLeft, current_encoderRight, latest_scan.ranges.data(), (int)latest_scan.ranges.size(), latest_scan.angle_min,
                       latest_scan.angle_increment, latest_scan.range_min, latest_scan.range_max, jitterState ? 1 : 0, pose, &st));
        last_stats_ = st;
        RobotPosition p;
        p.x = pose[0]; p.y = pose[1]; p.theta = pose[2];
        return p;
    }

    // ---- confidence estimate (MC:886-949): returns the density ratio; x_best / y_best / theta_best carry the reference's
    // globals of the same names (MC:73-75), -1 when the ratio does not exceed the threshold (MC:938-940) ----
    double isLocalizationLost_densitiy_cluster(double cluster_ratio_threshold) {
        mcl_kmeans_result r;
        check(mcl_kmeans_confidence(h_, nullptr, nullptr, 0, cluster_ratio_threshold, &r));
        x_best = r.x_best; y_best = r.y_best; theta_best = r.theta_best;
        last_kmeans_ = r;
        return r.ratio;
    }
    double x_best = -1, y_best = -1, theta_best = -1;
    const mcl_kmeans_result& lastKmeans() const { return last_kmeans_; }

    // ---- output adapters ----
    struct PoseMsg { int32_t row, column, orientation; };          // msg/Pose.msg
    struct ExactPoseMsg { float x, y, theta; };                    // msg/ExactPose.msg (the fields publishExactPose fills)
    static PoseMsg poseMsg(double wx, double wy, double angle, double cell_meters = 0.8) {      // publishPosMsg, MC:958-994
        PoseMsg m;
        if (mcl_pose_to_cell(wx, wy, angle, cell_meters, &m.row, &m.column, &m.orientation)) throw std::runtime_error("mcl_pose_to_cell: bad argument");
        return m;
    }
    static ExactPoseMsg exactPoseMsg(double x, double y, double theta) {                        // publishExactPose, MC:995-1008
        float o[3];
        mcl_exact_pose(x, y, theta, o);
        return ExactPoseMsg{o[0], o[1], o[2]};
    }
    // publishParticles (MC:563-579): every stride-th particle as {position.x, position.y, orientation.z, orientation.w}
    std::vector<double> poseArray(int64_t stride = 1) {
        const int64_t count = (cols() + stride - 1) / stride;
        std::vector<double> out((size_t)count * 4);
        check(mcl_download_pose_array(h_, 0, stride, count, out.data()));
        return out;
    }

    mcl_handle* handle() { return h_; }
    const mcl_config& config() const { return cfg_; }

private:
    void open() {
        int rc = mcl_create(&cfg_, &h_);
        if (rc) throw std::runtime_error(std::string("mcl_create: ") + mcl_last_error(nullptr));
    }
    void check(int rc) {
        if (rc) throw std::runtime_error(std::string("mcl: ") + mcl_last_error(h_));
    }
    mcl_config cfg_;
    mcl_handle* h_ = nullptr;
    mcl_resample_stats last_stats_{};
    mcl_kmeans_result last_kmeans_{};
};

}  // namespace mcl
