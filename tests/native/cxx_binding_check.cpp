// Compile-and-link check of include/mcl_particle_filter.hpp against libmcl_b200.so, using message structs shaped like
// nav_msgs::OccupancyGrid / sensor_msgs::LaserScan. Without a GPU the constructor must throw (no CPU fallback).
#include <cstdio>
#include <vector>

#include "../../include/mcl_particle_filter.hpp"

struct FakeGrid { struct { float resolution; unsigned width, height; struct { struct { double x, y; } position; } origin; } info; std::vector<int8_t> data; };
struct FakeScan { float angle_min, angle_increment, range_min, range_max; std::vector<float> ranges; };

int main() {
    try {
        mcl::ParticleFilter pf(0, MCL_MODE_REF);
        FakeGrid g;
        g.info.resolution = 0.1f; g.info.width = 49; g.info.height = 49; g.info.origin.position.x = 0; g.info.origin.position.y = 0;
        g.data.assign(49 * 49, 0);
        pf.setMap(g);
        pf.precomputeRayDirections(-120.0, 120.0, 0.1);
        pf.sampleParticles(1500);
        FakeScan s{-3.14159274f, 0.0174532924f, 0.02f, 5.6f, std::vector<float>(360, 0.7f)};
        pf.diffDriveModel(0.5, 0.5);
        int injected = pf.resampleParticles(s, true);
        mcl::RobotPosition p = pf.estimateWeightedPose();
        std::vector<float> P(4 * pf.cols());
        pf.downloadParticles(P.data());
        double confident_level = pf.isLocalizationLost_densitiy_cluster(0.6);
        mcl::ParticleFilter::PoseMsg cell = mcl::ParticleFilter::poseMsg(p.x, p.y, p.theta);
        std::vector<double> poses = pf.poseArray(10);
        mcl::RobotPosition q = pf.executeParticleFilter(0.6, 0.6, s, true);          // the whole tick as one engine call
        if (!(q.x == q.x) || pf.lastResampleStats().total_weight < 0) return 5;
        if (poses.size() != 4 * 150 || cell.row < -1 || confident_level < 0 || confident_level > 1) return 4;
        printf("gpu ok: injected %d pose %.3f %.3f %.3f confidence %.3f best %.3f %.3f\n", injected, p.x, p.y, p.theta, confident_level, pf.x_best, pf.y_best);
        return 0;
    } catch (const std::exception& e) {
        printf("threw: %s\n", e.what());
        return 3;
    }
}
