// mcl_oracle.cpp — CPU ORACLE (test infrastructure, NOT product code).
//
// A plain, single-threaded C++ restatement of the particle-filter hot path of
// Bright8787/MonteCarloLocalisation, pink_fundamentals/src/monte_carlo.cpp ("MC" below),
// plus the map rasteriser pink_fundamentals/src/publish_map_rviz.cpp ("RV") and the
// map.txt semantics of src/publish_map.py + msg/Cell.msg.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
// legs may load this library; the product (libmcl_b200.so) never links or calls it.
//
// Differences from the reference, all forced by "make it runnable and reproducible":
//   * ROS message types -> plain arrays; Eigen::MatrixXf(4,N) -> float* col-major 4xN.
//   * every RNG engine (all random_device-seeded, MC:411,427,452,508) -> caller-supplied
//     draw arrays (canonical u in [0,1), standard normal z, integer cells).
//   * ROS `tf` yaw<->quaternion round trip restated from the tf formulas (not vendored).
//   * Eigen fp32 array math restated with scalar fp32 ops (vendored Eigen lacks Eigen/Core).
//   * lower_bound returning N (out-of-bounds read, UB at MC:548) is clamped to N-1 and counted.
//   * float trig at MC:644-645, MC:747-748: selectable. trig_mode 0 = libm cosf/sinf (what the
//     reference compiles to; used to pin this oracle against oracle/_ref), trig_mode 1 =
//     correctly-rounded float trig, (float)cos((double)x) (the portable definition the CUDA
//     engine is held to; see DESIGN.md "float trig").
//
// Parity status: pinned against oracle/_ref (the unmodified monte_carlo.cpp compiled behind
// stub ROS/tf/Eigen headers) by tests/test_oracle_vs_ref.py and the fixtures in tests/golden/.
// The tf and Eigen arithmetic is stubbed there by our own headers, so those two boundaries
// remain "parity unpinned" (SURVEY.md §8c).
//
// Build: see oracle/Makefile (g++ -O2 -ffp-contract=off, no -ffast-math, no -march=native).

#include <algorithm>
#include <cmath>
#include <limits>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <utility>
#include <vector>

namespace {

// ----------------------------------------------------------------------------------------------
// Types (MC:87-128)
// ----------------------------------------------------------------------------------------------
struct EncoderData { double cur_l = 0, cur_r = 0, prev_l = 0, prev_r = 0; };     // MC:87-93,184
struct RobotPosition { double x = 0, y = 0, theta = 0; };                          // MC:106-110
struct OdometryModel { double rot_1 = 0, trans = 0, rot_2 = 0; };                  // MC:112-116
struct NoiseParameter { double a1 = 0.001, a2 = 0.001, a3 = 0.0001, a4 = 0.0001; }; // MC:118-123,1198
struct AdaptiveInjection { double weight_slow = 0, weight_fast = 0; };             // MC:125-128,191
struct Beam { double radius, angle; };                                             // MC:101-104

// GaussianLookup (MC:139-177). Float literals held in doubles on purpose (Q12).
struct GaussLut {
    double sigma, min_diff, max_diff, resolution;
    int size;
    std::vector<double> table;
    explicit GaussLut(double s, double res = 0.0001f) : sigma(s), min_diff(0.0f), max_diff(1.1f), resolution(res) {
        size = static_cast<int>((max_diff - min_diff) / resolution) + 1;
        table.resize(size);
        double denom = sigma * std::sqrt(2.0f * M_PI);
        for (int i = 0; i < size; ++i) {
            double diff = min_diff + i * resolution;
            table[i] = std::exp(-(diff * diff) / (2 * sigma * sigma)) / denom;
        }
    }
    double get(double diff) const {
        if (diff < min_diff || diff > max_diff) return 0.0f;
        double index_f = (diff - min_diff) / resolution;
        int index = static_cast<int>(index_f);
        if (index + 1 < size) {
            double weight = index_f - index;
            return (1.0f - weight) * table[index] + weight * table[index + 1];
        }
        return table[index];
    }
};

struct Ctx {
    // map (nav_msgs::OccupancyGrid fields that the path reads; MC:291-319)
    std::vector<int8_t> occ;
    int width = 0, height = 0;
    float resolution_f32 = 0.f;      // float32 on the wire (Q10)
    double origin_x = 0, origin_y = 0;
    bool map_ready = false;
    // global state of the reference
    GaussLut gauss{0.1};                                 // MC:177
    double w_hit = 0.8, w_rand = 0.2;                    // MC:180-181
    EncoderData enc;                                     // MC:184
    RobotPosition previous_position, current_position;   // MC:186-187
    OdometryModel motion;                                // MC:189
    NoiseParameter noise;                                // MC:190,1198
    AdaptiveInjection inj;                               // MC:191
    std::map<int, std::pair<double, double>> ray_lut;    // MC:192 (container kind is irrelevant)
    int trig_mode = 0;
    // bookkeeping for tests
    long long clamp_count = 0;
};

inline float f32cos(const Ctx& c, float t) { return c.trig_mode ? (float)std::cos((double)t) : cosf(t); }
inline float f32sin(const Ctx& c, float t) { return c.trig_mode ? (float)std::sin((double)t) : sinf(t); }

// ----------------------------------------------------------------------------------------------
// tf round trip (Q8): tf::createQuaternionMsgFromYaw (MC:646) then tf::getYaw (MC:351).
// ----------------------------------------------------------------------------------------------
struct Quat { double x, y, z, w; };
inline Quat quat_from_yaw(double yaw) {
    // tf::Quaternion::setRPY(0,0,yaw)
    double halfYaw = yaw * 0.5, halfPitch = 0.0 * 0.5, halfRoll = 0.0 * 0.5;
    double cosYaw = std::cos(halfYaw), sinYaw = std::sin(halfYaw);
    double cosPitch = std::cos(halfPitch), sinPitch = std::sin(halfPitch);
    double cosRoll = std::cos(halfRoll), sinRoll = std::sin(halfRoll);
    Quat q;
    q.x = sinRoll * cosPitch * cosYaw - cosRoll * sinPitch * sinYaw;
    q.y = cosRoll * sinPitch * cosYaw + sinRoll * cosPitch * sinYaw;
    q.z = cosRoll * cosPitch * sinYaw - sinRoll * sinPitch * cosYaw;
    q.w = cosRoll * cosPitch * cosYaw + sinRoll * sinPitch * sinYaw;
    return q;
}
inline double yaw_from_quat(const Quat& q) {
    // quaternionMsgToTF (renormalise only if |len2-1| > 0.1: never for a yaw quaternion)
    // Matrix3x3::setRotation + getEulerYPR, first solution.
    double d = q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w;
    double s = 2.0 / d;
    double ys = q.y * s, zs = q.z * s;
    double wy = q.w * ys, wz = q.w * zs;
    double xy = q.x * ys, xz = q.x * zs;
    double yy = q.y * ys, zz = q.z * zs;
    double m00 = 1.0 - (yy + zz);
    double m10 = xy + wz;
    double m20 = xz - wy;
    if (std::fabs(m20) >= 1) {               // gimbal branch; unreachable for pure yaw
        return 0.0;
    }
    double pitch = -std::asin(m20);
    return std::atan2(m10 / std::cos(pitch), m00 / std::cos(pitch));
}

// ----------------------------------------------------------------------------------------------
// Map probes (MC:298-349, 685-692)
// ----------------------------------------------------------------------------------------------
inline bool worldToMap(const Ctx& c, double wx, double wy, int& mx, int& my) {
    double resolution = c.resolution_f32;                         // MC:302 (float32 -> double)
    mx = static_cast<int>((wx - c.origin_x) / resolution);       // MC:304
    my = static_cast<int>((wy - c.origin_y) / resolution);       // MC:305
    return mx >= 0 && my >= 0 && mx < c.width && my < c.height;  // MC:307-309
}
inline int getCell(const Ctx& c, int mx, int my) { return c.occ[(size_t)my * c.width + mx]; }   // MC:316-319
inline bool isOccupied(const Ctx& c, double wx, double wy) {      // MC:320-328
    int xi, yi = 0;
    if (!worldToMap(c, wx, wy, xi, yi)) return false;
    return getCell(c, xi, yi) > 50;
}
inline bool isInsideMap(const Ctx& c, double x, double y) {       // MC:685-692
    double map_min_x = c.origin_x, map_min_y = c.origin_y;
    double map_max_x = map_min_x + c.width * c.resolution_f32;    // unsigned*float -> float, then + double
    double map_max_y = map_min_y + c.height * c.resolution_f32;
    return (x >= map_min_x && x < map_max_x) && (y >= map_min_y && y < map_max_y);
}
inline bool isValidPos(const Ctx& c, double wx, double wy) {      // MC:331-349
    double offset = 0.1;
    if (!isInsideMap(c, wx, wy)) return false;
    const double offs[9][2] = {{0, 0}, {offset, 0}, {0, offset}, {-offset, 0}, {0, -offset},
                               {offset, offset}, {offset, -offset}, {-offset, offset}, {-offset, -offset}};
    for (auto& o : offs)
        if (isOccupied(c, wx + o[0], wy + o[1])) return false;
    return true;
}

// raycast (MC:350-390). The pose arrives as (x, y, quaternion-of-theta).
inline double raycast(Ctx& c, double px, double py, const Quat& q, double angle_offset_deg, double max_range) {
    double robot_yaw = yaw_from_quat(q);                                           // MC:351
    double absolute_angle_deg = robot_yaw * 180.0 / M_PI + angle_offset_deg;       // MC:352
    int angle_key = static_cast<int>(std::round(absolute_angle_deg));             // MC:355
    auto it = c.ray_lut.find(angle_key);                                           // MC:358
    if (it == c.ray_lut.end()) {
        double angle_rad = robot_yaw + angle_offset_deg * M_PI / 180.0;            // MC:361
        it = c.ray_lut.emplace(angle_key, std::make_pair(std::cos(angle_rad), std::sin(angle_rad))).first;
    }
    double dx = it->second.first, dy = it->second.second;
    double step = 0.1;
    for (double r = 0.0; r < max_range; r += step) {                               // MC:372
        double rx = px + r * dx;
        double ry = py + r * dy;
        int mx, my;
        if (!worldToMap(c, rx, ry, mx, my)) break;
        if (getCell(c, mx, my) > 50) return r;
    }
    return max_range;
}

// filterLaserReadings (MC:254-278) then filterAngles (MC:610-620).
std::vector<Beam> filter_scan(const float* ranges, int B, float angle_min, float angle_increment_f,
                              float range_min, float range_max, double lower, double upper) {
    std::vector<Beam> readings;
    double min_angle = angle_min;
    double angle_increment = angle_increment_f;
    for (size_t i = 0; i < (size_t)B; i++) {
        double r = ranges[i];
        if (std::isnan(r) || std::isinf(r)) {
            readings.push_back({1.05, min_angle + (i * angle_increment)});
        } else if (r >= range_min && r <= range_max) {
            readings.push_back({r, min_angle + (i * angle_increment)});
        }
    }
    std::vector<Beam> filtered;
    for (size_t i = 0; i < readings.size(); i++) {
        if (readings[i].angle * 180.0 / M_PI > lower && readings[i].angle * 180.0 / M_PI < upper)
            filtered.push_back(readings[i]);
    }
    return filtered;
}

// computeWeight (MC:623-682). P is 4xN col-major (x,y,theta,w per particle).
double compute_weight(Ctx& c, float* P, int N, const std::vector<Beam>& laser_data) {
    double max_laser_range = 1.0;
    double prob = 0;
    double totalWeight = 0;
    double laser_offset = 0.1;
    for (int j = 0; j < N; ++j) {
        float* p = P + 4 * (size_t)j;
        prob = 0;
        double posx = p[0] + laser_offset * f32cos(c, p[2]);      // MC:644 (float + double*float)
        double posy = p[1] + laser_offset * f32sin(c, p[2]);      // MC:645
        Quat q = quat_from_yaw(p[2]);                             // MC:646
        if (isValidPos(c, p[0], p[1])) {                          // MC:648
            for (size_t i = 0; i < laser_data.size(); i += 20) {  // MC:650
                double angle = -(laser_data[i].angle) * 180.0 / M_PI;                   // MC:653
                double expected_distance = raycast(c, posx, posy, q, angle, max_laser_range);
                double observed_distance = laser_data[i].radius;
                double diff = std::fabs(observed_distance - expected_distance);         // MC:662
                prob += c.w_hit * c.gauss.get(diff);                                    // MC:665
                prob += c.w_rand * ((std::abs(observed_distance - max_laser_range) < 0.01) ? 1.0 : 0.0);  // MC:669
            }
        }
        p[3] = prob;                      // MC:673 (narrow to float)
        totalWeight += p[3];              // MC:675
    }
    return totalWeight;
}

// sample(variance) (MC:410-414) with the engine's standard-normal output injected:
// libstdc++ normal_distribution returns ret * stddev + mean.
inline double sample_normal(double variance, double z) { return z * std::sqrt(variance) + 0.0; }
// uniform_real_distribution(a,b) with the canonical draw injected: libstdc++ returns u*(b-a)+a.
inline double uniform_ab(double a, double b, double u) { return (u * (b - a)) + a; }

// sampleParticles body for one particle (MC:434-446), draws named (Q21).
inline void sample_one(const Ctx&, double u_yaw, int row, int col, double u_dx, double u_dy, float* out) {
    double orientation = uniform_ab(-M_PI, M_PI, u_yaw);          // MC:432,436
    const double CELL_METERS = 0.8;                               // to_cell MC:394-407
    double x_move = uniform_ab(-0.2, 0.2, u_dx), y_move = uniform_ab(-0.2, 0.2, u_dy);
    double base_x = col * CELL_METERS + 0.5 * CELL_METERS;
    double base_y = row * CELL_METERS + 0.5 * CELL_METERS;
    double px = base_x + x_move, py = base_y + y_move;
    out[0] = px + 0.05;      // MC:442
    out[1] = py + 0.05;      // MC:443
    out[2] = orientation;    // MC:444
    out[3] = 1.0;            // MC:445
}

}  // namespace

// ================================================================================================
// C API (ctypes-friendly)
// ================================================================================================
extern "C" {

void* orc_create() { return new Ctx(); }
void orc_destroy(void* h) { delete (Ctx*)h; }
void orc_set_trig_mode(void* h, int m) { ((Ctx*)h)->trig_mode = m; }
// this process's libm sinf / cosf (what the reference binary calls at MC:644-645 and, through Eigen, MC:747-748) over an
// array: the checker of the engine's device restatement of them (csrc/glibc_trigf.cuh)
void orc_libm_trigf(const float* x, long long n, float* s_out, float* c_out) {
    for (long long i = 0; i < n; ++i) { volatile float v = x[i]; s_out[i] = sinf(v); c_out[i] = cosf(v); }
}
long long orc_clamp_count(void* h) { return ((Ctx*)h)->clamp_count; }

// ---- map.txt -> wall lists -> occupancy grid (publish_map.py:8-16, Cell.msg:2-5, RV:272-276,306-437)
// Parses the nested Python list literal with bare identifiers T,B,L,R. Returns 0 on success.
// walls_out: flattened; rows/cols via row_len; uses std::vector internally.
static int parse_map_txt(const char* text, std::vector<std::vector<std::vector<int>>>& walls) {
    // Cell.msg: RIGHT=0, TOP=1, LEFT=2, BOTTOM=3
    int depth = 0;
    walls.clear();
    for (const char* p = text; *p; ++p) {
        char ch = *p;
        if (ch == '[') {
            depth++;
            if (depth == 2) walls.emplace_back();
            else if (depth == 3) { if (walls.empty()) return -1; walls.back().emplace_back(); }
            else if (depth > 3) return -1;
        } else if (ch == ']') {
            depth--;
            if (depth < 0) return -1;
        } else if (ch == 'T' || ch == 'B' || ch == 'L' || ch == 'R') {
            if (depth != 3) return -1;
            int v = ch == 'R' ? 0 : ch == 'T' ? 1 : ch == 'L' ? 2 : 3;
            walls.back().back().push_back(v);
        } else if (ch == ',' || ch == ' ' || ch == '\n' || ch == '\r' || ch == '\t') {
        } else {
            return -1;
        }
    }
    return depth == 0 && !walls.empty() ? 0 : -1;
}

// createOccupancyGrid (RV:306-437). out must hold (maxcols*8+1)*(rows*8+1) bytes.
int orc_rasterise_map_txt(const char* text, int8_t* out, int cap, int* w_out, int* h_out) {
    std::vector<std::vector<std::vector<int>>> walls;
    if (parse_map_txt(text, walls)) return -1;
    const int CELL_SIZE = 8, WALL_OCCUPIED = 100, FREE_SPACE = 0;
    int grid_height = walls.size();
    int grid_width = 0, column_size = 0;
    for (const auto& row : walls)
        if ((int)row.size() > grid_width) { grid_width = row.size(); column_size = grid_width; }
    int map_width = grid_width * CELL_SIZE + 1;
    int map_height = grid_height * CELL_SIZE + 1;
    if (map_width * map_height > cap) return -2;
    std::vector<int8_t> data(map_width * map_height, FREE_SPACE);
    auto put = [&](int idx) { if (idx >= 0 && idx < (int)data.size()) data[idx] = WALL_OCCUPIED; };
    for (int cell_y = 0; cell_y < grid_height; ++cell_y) {
        const auto& row = walls[cell_y];
        int row_width = row.size();
        for (int cell_x = 0; cell_x < row_width; ++cell_x) {
            int x = cell_y * CELL_SIZE;       // note: "x" is the grid ROW offset (RV:340)
            int y = cell_x * CELL_SIZE;       //       "y" is the grid COLUMN offset
            for (int w : row[cell_x]) {
                if (w == 1) {                 // top
                    for (int dx = 0; dx <= CELL_SIZE; ++dx) put(x * map_width + (y + dx));
                } else if (w == 2) {          // left
                    for (int dy = 0; dy <= CELL_SIZE; ++dy) put((x + dy) * map_width + y);
                } else if (w == 0) {          // right: only on the last column of the row
                    if (cell_x == row_width - 1)
                        for (int dy = 0; dy < CELL_SIZE; ++dy) put((x + dy) * map_width + (y + CELL_SIZE));
                } else if (w == 3) {          // bottom: last row, or no cell below
                    if (cell_y == grid_height - 1)
                        for (int dx = 0; dx < CELL_SIZE; ++dx) put((x + CELL_SIZE) * map_width + (y + dx + 1));
                    if (cell_y + 1 < grid_height)
                        if (!((size_t)cell_x < walls[cell_y + 1].size()))
                            for (int dx = 0; dx < CELL_SIZE; ++dx) put((x + CELL_SIZE) * map_width + (y + dx + 1));
                }
            }
        }
        while (row_width < column_size) {     // ragged rows: fill the missing cells solid (RV:395-408)
            double x = cell_y * CELL_SIZE;
            double y = row_width * CELL_SIZE;
            for (int dy = 0; dy < CELL_SIZE; ++dy)
                for (int dx = 0; dx <= CELL_SIZE; ++dx) put((int)((x + dy) * map_width + (y + dx)));
            row_width++;
        }
    }
    memcpy(out, data.data(), data.size());
    *w_out = map_width;
    *h_out = map_height;
    return 0;
}

void orc_set_map(void* h, const int8_t* occ, int w, int hgt, float res, double ox, double oy) {
    Ctx& c = *(Ctx*)h;
    c.occ.assign(occ, occ + (size_t)w * hgt);
    c.width = w; c.height = hgt; c.resolution_f32 = res; c.origin_x = ox; c.origin_y = oy;
    c.map_ready = true;
}

// precomputeRayDirections (MC:1017-1023) — including its key-scale bug (Q9).
void orc_precompute_ray_directions(void* h, double min_angle_deg, double max_angle_deg, double step_deg) {
    Ctx& c = *(Ctx*)h;
    for (double a = min_angle_deg; a <= max_angle_deg; a += step_deg) {
        int angle_key = static_cast<int>(a * 100);
        double rad = a * M_PI / 180.0;
        c.ray_lut[angle_key] = std::make_pair(std::cos(rad), std::sin(rad));
    }
}
void orc_clear_ray_directions(void* h) { ((Ctx*)h)->ray_lut.clear(); }
int orc_ray_lut_size(void* h) { return (int)((Ctx*)h)->ray_lut.size(); }
// Dump entries with lo <= key <= hi; returns count.
int orc_ray_lut_dump(void* h, int lo, int hi, int* keys, double* dx, double* dy, int cap) {
    Ctx& c = *(Ctx*)h;
    int n = 0;
    for (auto& kv : c.ray_lut) {
        if (kv.first < lo || kv.first > hi) continue;
        if (n < cap) { keys[n] = kv.first; dx[n] = kv.second.first; dy[n] = kv.second.second; }
        n++;
    }
    return n;
}
void orc_ray_lut_set(void* h, int key, double dx, double dy) { ((Ctx*)h)->ray_lut[key] = std::make_pair(dx, dy); }

double orc_gauss_get(void* h, double diff) { return ((Ctx*)h)->gauss.get(diff); }
int orc_gauss_size(void* h) { return ((Ctx*)h)->gauss.size; }
void orc_gauss_table(void* h, double* out) { Ctx& c = *(Ctx*)h; memcpy(out, c.gauss.table.data(), sizeof(double) * c.gauss.size); }

int orc_filter_scan(const float* ranges, int B, float angle_min, float angle_inc, float range_min, float range_max,
                    double lower, double upper, double* radius, double* angle, int cap) {
    auto v = filter_scan(ranges, B, angle_min, angle_inc, range_min, range_max, lower, upper);
    for (size_t i = 0; i < v.size() && (int)i < cap; i++) { radius[i] = v[i].radius; angle[i] = v[i].angle; }
    return (int)v.size();
}

int orc_is_valid_pos(void* h, double x, double y) { return isValidPos(*(Ctx*)h, x, y) ? 1 : 0; }
int orc_is_occupied(void* h, double x, double y) { return isOccupied(*(Ctx*)h, x, y) ? 1 : 0; }
double orc_yaw_roundtrip(double theta) { return yaw_from_quat(quat_from_yaw(theta)); }
double orc_raycast(void* h, double x, double y, double theta, double angle_offset_deg, double max_range) {
    return raycast(*(Ctx*)h, x, y, quat_from_yaw(theta), angle_offset_deg, max_range);
}

// sampleParticles (MC:415-450) with named draws. cells: row ~ U{0..height/8-1}, col ~ U{0..width/8-1}.
void orc_sample_particles(void* h, int N, const double* u_yaw, const int* row, const int* col,
                          const double* u_dx, const double* u_dy, float* P) {
    Ctx& c = *(Ctx*)h;
    for (int i = 0; i < N; i++) sample_one(c, u_yaw[i], row[i], col[i], u_dx[i], u_dy[i], P + 4 * (size_t)i);
}
void orc_cell_ranges(void* h, int* n_rows, int* n_cols) {
    Ctx& c = *(Ctx*)h;
    int cell_size = 8;
    *n_cols = (unsigned)c.width / cell_size;      // maze_coordiante_width  (MC:423)
    *n_rows = (unsigned)c.height / cell_size;     // maze_coordiante_height (MC:424)
}

// diffDriveModel (MC:719-739) + sampleMotionModelOdometry (MC:695-717). z[3] = standard normal draws
// in call order (rot1, trans, rot2). out[3] = noised (rot_1, trans, rot_2).
void orc_set_encoders(void* h, double left, double right) { Ctx& c = *(Ctx*)h; c.enc.cur_l = left; c.enc.cur_r = right; }   // MC:756-763
void orc_diff_drive(void* h, const double* z, double* out) {
    Ctx& c = *(Ctx*)h;
    const double wheel_size = 0.0620, wheel_space = 0.265;       // PID_lib.hpp:19-20
    double d_left = (c.enc.cur_l - c.enc.prev_l) * wheel_size * 0.5;
    double d_right = (c.enc.cur_r - c.enc.prev_r) * wheel_size * 0.5;
    double d_center = 0.5 * (d_left + d_right);
    double delta_theta = (d_left - d_right) / wheel_space;
    RobotPosition& prev = c.previous_position;
    double current_theta = delta_theta + prev.theta;
    double current_pos_x = prev.x + d_center * std::cos(prev.theta + 0.5 * delta_theta);
    double current_pos_y = prev.y + d_center * std::sin(prev.theta + 0.5 * delta_theta);
    c.current_position = {current_pos_x, current_pos_y, std::atan2(std::sin(current_theta), std::cos(current_theta))};
    RobotPosition& cur = c.current_position;
    // sampleMotionModelOdometry
    OdometryModel& m = c.motion;
    const NoiseParameter& n = c.noise;
    m.rot_1 = std::atan2(cur.y - prev.y, cur.x - prev.x) - prev.theta;
    m.trans = std::sqrt((cur.y - prev.y) * (cur.y - prev.y) + (cur.x - prev.x) * (cur.x - prev.x));
    m.rot_2 = cur.theta - prev.theta - m.rot_1;
    double rot_1_temp = m.rot_1 + sample_normal(n.a1 * std::fabs(m.rot_1) + n.a2 * m.trans, z[0]);
    double translation_noise = sample_normal(n.a3 * m.trans + n.a4 * (std::fabs(m.rot_1) + std::fabs(m.rot_2)), z[1]);
    double trans_temp = m.trans + translation_noise;
    double rot_2_temp = m.rot_2 + sample_normal(n.a1 * std::fabs(m.rot_2) + n.a2 * m.trans, z[2]);
    m.rot_1 = rot_1_temp; m.trans = trans_temp; m.rot_2 = rot_2_temp;
    prev = cur;
    c.enc.prev_l = c.enc.cur_l;
    c.enc.prev_r = c.enc.cur_r;
    out[0] = m.rot_1; out[1] = m.trans; out[2] = m.rot_2;
}
void orc_set_motion(void* h, double rot1, double trans, double rot2) { Ctx& c = *(Ctx*)h; c.motion = {rot1, trans, rot2}; }
void orc_get_robot_position(void* h, double* out) { Ctx& c = *(Ctx*)h; out[0] = c.previous_position.x; out[1] = c.previous_position.y; out[2] = c.previous_position.theta; }

// updateParticlePos (MC:740-755): fp32 array ops; double scalars are narrowed to float first (Eigen
// converts a scalar operand to the array's Scalar type).
void orc_update_particle_pos(void* h, float* P, int N) {
    Ctx& c = *(Ctx*)h;
    float rot1 = (float)c.motion.rot_1, trans = (float)c.motion.trans;
    float dtheta = (float)(c.motion.rot_1 + c.motion.rot_2);
    for (int i = 0; i < N; i++) {
        float* p = P + 4 * (size_t)i;
        float moved_heading = p[2] + rot1;
        float dx = trans * f32cos(c, moved_heading);
        float dy = trans * f32sin(c, moved_heading);
        p[0] += dx; p[1] += dy; p[2] += dtheta;
    }
}

// computeWeight (MC:623-682) on a raw scan. Returns totalWeight.
double orc_compute_weight(void* h, float* P, int N, const float* ranges, int B, float angle_min, float angle_inc,
                          float range_min, float range_max) {
    Ctx& c = *(Ctx*)h;
    auto laser = filter_scan(ranges, B, angle_min, angle_inc, range_min, range_max, -120.00, 120.00);  // MC:635
    return compute_weight(c, P, N, laser);
}

// resampleParticles (MC:457-561). P (4xN) is both `particles_local` and the global `particles`
// (they alias at both call sites MC:1089,1206); its weight row is normalised in place like the reference.
//   u_r[N]        canonical draw per output slot (MC:514)
//   u_jit[]       canonical jitter draws, consumed in order by non-injected slots: x, y, (theta if jitterState)
//   inj_*[]       named draws of sampleParticles(1), consumed in order by injected slots
// Outputs: Pout (4xN), idx_out[N] = ancestor index or -1 for injected slots, cdf_out[N] (may be null),
//          stats[0]=injected count, [1]=p_inject, [2]=weight_slow, [3]=weight_fast, [4]=total_weight.
void orc_resample(void* h, float* P, int N, int jitterState,
                  const float* ranges, int B, float angle_min, float angle_inc, float range_min, float range_max,
                  const double* u_r, const double* u_jit,
                  const double* inj_u_yaw, const int* inj_row, const int* inj_col, const double* inj_u_dx, const double* inj_u_dy,
                  float* Pout, int* idx_out, double* cdf_out, double* stats) {
    Ctx& c = *(Ctx*)h;
    double max_injection, alpha_slow, alpha_fast;
    double total_weight = orc_compute_weight(h, P, N, ranges, B, angle_min, angle_inc, range_min, range_max);  // MC:468
    double weight_avg = total_weight / N;
    if (jitterState) { max_injection = 200; alpha_slow = 0.05; alpha_fast = 0.5; }
    else { max_injection = 50; alpha_slow = 0.02; alpha_fast = 2; }
    c.inj.weight_slow = c.inj.weight_slow + alpha_slow * (weight_avg - c.inj.weight_slow);   // MC:487
    c.inj.weight_fast = c.inj.weight_fast + alpha_fast * (weight_avg - c.inj.weight_fast);   // MC:488
    double p_inject = std::max(0.0, 1.0 - (c.inj.weight_fast / c.inj.weight_slow));          // MC:492
    int injected = 0;
    std::vector<double> cdf(N, 0.0);
    P[3] = P[3] / total_weight;                    // MC:497 (float/double -> double -> float)
    cdf[0] = P[3];
    for (int i = 1; i < N; ++i) {
        P[4 * (size_t)i + 3] = P[4 * (size_t)i + 3] / total_weight;   // MC:503
        cdf[i] = cdf[i - 1] + P[4 * (size_t)i + 3];                   // MC:504
    }
    size_t jit_pos = 0;
    for (int i = 0; i < N; ++i) {
        double r = u_r[i];                                            // MC:514
        float* o = Pout + 4 * (size_t)i;
        if (r < p_inject && injected < max_injection) {               // MC:518
            float np[4];
            sample_one(c, inj_u_yaw[injected], inj_row[injected], inj_col[injected], inj_u_dx[injected], inj_u_dy[injected], np);
            o[0] = np[0]; o[1] = np[1]; o[2] = np[2];
            o[3] = 1.0 / N;
            if (idx_out) idx_out[i] = -1;
            injected++;
        } else {
            auto it = std::lower_bound(cdf.begin(), cdf.end(), r);    // MC:530
            int idx = std::distance(cdf.begin(), it);
            if (idx >= N) { idx = N - 1; c.clamp_count++; }           // documented deviation from UB
            const float* a = P + 4 * (size_t)idx;
            double jitter_theta, jitter_x, jitter_y;
            if (jitterState) {
                jitter_x = uniform_ab(-0.05, 0.05, u_jit[jit_pos++]);
                jitter_y = uniform_ab(-0.05, 0.05, u_jit[jit_pos++]);
                jitter_theta = a[2] + uniform_ab(-M_PI / 12, M_PI / 12, u_jit[jit_pos++]);
            } else {
                jitter_x = uniform_ab(-0.01, 0.01, u_jit[jit_pos++]);
                jitter_y = uniform_ab(-0.01, 0.01, u_jit[jit_pos++]);
                jitter_theta = a[2];
            }
            o[0] = a[0] + jitter_x;                                   // MC:548
            o[1] = a[1] + jitter_y;                                   // MC:549
            o[2] = std::atan2(std::sin(jitter_theta), std::cos(jitter_theta));   // MC:550
            o[3] = 1.0 / N;                                           // MC:551
            if (idx_out) idx_out[i] = idx;
        }
    }
    if (cdf_out) memcpy(cdf_out, cdf.data(), sizeof(double) * N);
    if (stats) { stats[0] = injected; stats[1] = p_inject; stats[2] = c.inj.weight_slow; stats[3] = c.inj.weight_fast; stats[4] = total_weight; }
}
void orc_get_injection_state(void* h, double* out) { Ctx& c = *(Ctx*)h; out[0] = c.inj.weight_slow; out[1] = c.inj.weight_fast; }
void orc_set_injection_state(void* h, double slow, double fast) { Ctx& c = *(Ctx*)h; c.inj.weight_slow = slow; c.inj.weight_fast = fast; }

// estimateWeightedPose (MC:782-800): fp32 Eigen reductions; restated with fp32 element math and
// f64 accumulation (Eigen's SIMD reduction order is unknowable here; graded at 1e-5 relative).
void orc_estimate_weighted_pose(void* h, const float* P, int N, double* out) {
    Ctx& c = *(Ctx*)h;
    double ws = 0;
    for (int i = 0; i < N; i++) ws += P[4 * (size_t)i + 3];
    float weight_sum = (float)ws;
    double sx = 0, sy = 0, ss = 0, sc = 0;
    for (int i = 0; i < N; i++) {
        const float* p = P + 4 * (size_t)i;
        float w = p[3] / weight_sum;
        sx += (float)(w * p[0]);
        sy += (float)(w * p[1]);
        ss += (float)(w * f32sin(c, p[2]));
        sc += (float)(w * f32cos(c, p[2]));
    }
    float x_mean = (float)sx, y_mean = (float)sy;
    float theta_mean = std::atan2((float)ss, (float)sc);
    out[0] = x_mean; out[1] = y_mean; out[2] = theta_mean;
}

// ---- k-means confidence estimate and output adapters ---------------------------------------------

// kMeansClustering (MC:802-868). The reference seeds srand(time) (MC:808) and draws rand() % N for the K initial
// centres (MC:812-815) and for every emptied cluster (MC:859-861); here those indices are INJECTED in consumption
// order. assignments start at 0 (a fresh std::vector<int>::resize, MC:805 + MC:893). Centres accumulate in fp32,
// sequentially over the particles (MC:851-853), then divide by the int count (MC:857-858).
// Returns the number of assignment passes made; *reinit_used = how many re-initialisation draws were consumed.
int orc_kmeans(const float* P, int N, int K, int max_iters, const int* init_idx, const int* reinit_idx, int n_reinit, int* assignments,
               float* centers, int* reinit_used) {
    for (int i = 0; i < N; i++) assignments[i] = 0;
    for (int k = 0; k < K; k++) { centers[2 * k] = P[4 * (size_t)init_idx[k]]; centers[2 * k + 1] = P[4 * (size_t)init_idx[k] + 1]; }
    int used = 0, passes = 0;
    for (int iter = 0; iter < max_iters; ++iter) {
        bool changed = false;
        ++passes;
        for (int i = 0; i < N; i++) {                                              // MC:821-840
            float x = P[4 * (size_t)i], y = P[4 * (size_t)i + 1];
            float min_dist = std::numeric_limits<float>::max();
            int best = -1;
            for (int k = 0; k < K; k++) {
                float dx = x - centers[2 * k], dy = y - centers[2 * k + 1];
                float dist = dx * dx + dy * dy;
                if (dist < min_dist) { min_dist = dist; best = k; }
            }
            if (assignments[i] != best) { assignments[i] = best; changed = true; }
        }
        if (!changed) break;                                                       // MC:842-845
        std::vector<int> counts(K, 0);
        std::vector<float> nc(2 * (size_t)K, 0.0f);
        for (int i = 0; i < N; i++) {                                              // MC:851-856
            // best == -1 (all distances NaN) would index out of bounds in the reference; particles are finite here
            int c = assignments[i];
            nc[2 * c] += P[4 * (size_t)i];
            nc[2 * c + 1] += P[4 * (size_t)i + 1];
            counts[c]++;
        }
        for (int k = 0; k < K; k++) {                                              // MC:857-863
            if (counts[k] > 0) { nc[2 * k] /= counts[k]; nc[2 * k + 1] /= counts[k]; }
            else {
                int idx = used < n_reinit ? reinit_idx[used] : 0;
                ++used;
                nc[2 * k] = P[4 * (size_t)idx]; nc[2 * k + 1] = P[4 * (size_t)idx + 1];
            }
        }
        for (int k = 0; k < 2 * K; k++) centers[k] = nc[k];
    }
    if (reinit_used) *reinit_used = used;
    return passes;
}

// countParticlesNearCluster (MC:869-884)
int orc_count_near(const float* P, int N, float xc, float yc, float radius) {
    int count = 0;
    float radius_sq = radius * radius;
    for (int i = 0; i < N; i++) {
        float dx = P[4 * (size_t)i] - xc, dy = P[4 * (size_t)i + 1] - yc;
        float d = dx * dx + dy * dy;
        if (d <= radius_sq) ++count;
    }
    return count;
}

// isLocalizationLost_densitiy_cluster (MC:886-949), K = 3, 20 iterations, radius 0.4 (its cluster_distance argument is
// unused). out = {x_best, y_best, theta_best} with the -1 sentinel (MC:938-940); info = {best_cluster, passes, reinit_used};
// cluster_weights[K] optional. Returns the density ratio.
double orc_kmeans_confidence(const float* P, int N, const int* init_idx, const int* reinit_idx, int n_reinit, double ratio_threshold,
                             double* out, float* centers, int* assignments, int* info, double* cluster_weights) {
    const int K = 3, max_iters = 20;
    int used = 0;
    int passes = orc_kmeans(P, N, K, max_iters, init_idx, reinit_idx, n_reinit, assignments, centers, &used);
    double cw[3] = {0.0, 0.0, 0.0};
    for (int i = 0; i < N; i++) cw[assignments[i]] += P[4 * (size_t)i + 3];       // MC:903-907
    int best = 0;
    double max_w = cw[0];
    for (int k = 1; k < K; k++) if (cw[k] > max_w) { max_w = cw[k]; best = k; }    // MC:910-917
    double xb = centers[2 * best], yb = centers[2 * best + 1];                     // MC:920-921
    double ss = 0.0, cs = 0.0;
    for (int i = 0; i < N; i++) {                                                  // MC:924-931
        if (assignments[i] == best) { double th = P[4 * (size_t)i + 2]; ss += std::sin(th); cs += std::cos(th); }
    }
    double tb = std::atan2(ss, cs);
    double ratio = static_cast<double>(orc_count_near(P, N, (float)xb, (float)yb, 0.4f)) / N;      // MC:933 (float parameters)
    if (ratio > ratio_threshold) { out[0] = xb; out[1] = yb; out[2] = tb; } else { out[0] = -1; out[1] = -1; out[2] = -1; }
    if (info) { info[0] = best; info[1] = passes; info[2] = used; }
    if (cluster_weights) for (int k = 0; k < K; k++) cluster_weights[k] = cw[k];
    return ratio;
}

// publishPosMsg (MC:958-994): world pose -> maze cell (row, column) and 4-way direction; out = {row, column, orientation},
// all -1 for the "not localised" sentinel (negative coordinates). RIGHT=0, UP=1, LEFT=2, DOWN=3 (MC:130-135, msg/Pose.msg).
void orc_pose_to_cell(double wx, double wy, double angle, int* out) {
    const double CELL_METERS = 0.8;
    if (wx < 0 || wy < 0) { out[0] = out[1] = out[2] = -1; return; }
    double col_wx = (wx - 0.5 * CELL_METERS) / CELL_METERS;
    double row_wx = (wy - 0.5 * CELL_METERS) / CELL_METERS;
    int col = static_cast<int>(std::floor(col_wx + 0.5));
    int row = static_cast<int>(std::floor(row_wx + 0.5));
    const double TWO_PI = 2.0 * M_PI;                                              // wrapTo2Pi, MC:951-957
    double wrapped = std::fmod(angle, TWO_PI);
    if (wrapped < 0) wrapped += TWO_PI;
    double deg = wrapped * 180.0 / M_PI;
    int dir;
    if (deg >= 45 && deg < 135) dir = 3;
    else if (deg >= 135 && deg < 225) dir = 2;
    else if (deg >= 225 && deg < 315) dir = 1;
    else dir = 0;
    out[0] = row; out[1] = col; out[2] = dir;
}
// publishExactPose (MC:995-1008): float32 narrowing of the message fields (msg/ExactPose.msg)
void orc_exact_pose(double x, double y, double theta, float* out) { out[0] = (float)x; out[1] = (float)y; out[2] = (float)theta; }
// publishParticles (MC:563-579): pose i = (x, y, quaternion of yaw); tf::createQuaternionMsgFromYaw = setRPY(0,0,yaw):
// qz = sin(yaw/2), qw = cos(yaw/2), qx = qy = 0 (restated from the tf formulas: PARITY UNPINNED at this boundary)
void orc_particle_poses(const float* P, int N, double* out) {
    for (int i = 0; i < N; i++) {
        double yaw = P[4 * (size_t)i + 2], h = yaw * 0.5;
        out[4 * (size_t)i] = P[4 * (size_t)i]; out[4 * (size_t)i + 1] = P[4 * (size_t)i + 1];
        out[4 * (size_t)i + 2] = std::sin(h); out[4 * (size_t)i + 3] = std::cos(h);
    }
}

}  // extern "C"
