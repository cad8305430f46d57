// host_models.hpp — the scalar, per-step host work of the filter, in f64 with the same libm the reference's CPU
// build uses: map.txt rasterisation, scan filtering, the odometry motion model, the Gaussian and ray-direction
// tables. None of this is data-parallel; it feeds the kernels a few hundred bytes per step.
#pragma once
#include <cmath>
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/mcl.h"

namespace mcl {

// ---- map.txt -> occupancy grid ---------------------------------------------------------------------------
// map.txt is a nested Python list literal of per-cell wall lists with bare identifiers T,B,L,R
// (publish_map.py:8-16; Cell.msg:2-5). The grid rule is publish_map_rviz.cpp:306-437: 8 px per cell plus one
// closing line; a cell draws its top and left walls; a right wall only on the last cell of its row; a bottom wall
// only when no cell lies below; cells missing from a short row are filled solid.
struct WallGrid {
    std::vector<std::vector<uint8_t>> rows;   // bit0 = right, bit1 = top, bit2 = left, bit3 = bottom
};

inline bool parse_map_txt(const std::string& text, WallGrid& g, std::string& err) {
    g.rows.clear();
    int depth = 0;
    for (size_t pos = 0; pos < text.size(); ++pos) {
        const char ch = text[pos];
        switch (ch) {
            case '[':
                ++depth;
                if (depth == 2) g.rows.emplace_back();
                else if (depth == 3) {
                    if (g.rows.empty()) { err = "cell list outside a row"; return false; }
                    g.rows.back().push_back(0);
                } else if (depth > 3) { err = "nesting deeper than rows/cells/walls"; return false; }
                break;
            case ']':
                if (--depth < 0) { err = "unbalanced ']'"; return false; }
                break;
            case 'R': case 'T': case 'L': case 'B': {
                if (depth != 3) { err = "wall identifier outside a cell"; return false; }
                const int bit = ch == 'R' ? 0 : ch == 'T' ? 1 : ch == 'L' ? 2 : 3;
                g.rows.back().back() |= (uint8_t)(1u << bit);
                break;
            }
            case ',': case ' ': case '\t': case '\n': case '\r':
                break;
            default:
                err = std::string("unexpected character '") + ch + "'";
                return false;
        }
    }
    if (depth != 0 || g.rows.empty()) { err = "unbalanced or empty map"; return false; }
    return true;
}

inline void rasterise_walls(const WallGrid& g, int px, std::vector<int8_t>& occ, int& width, int& height) {
    const int n_rows = (int)g.rows.size();
    int n_cols = 0;
    for (auto& r : g.rows) n_cols = std::max(n_cols, (int)r.size());
    width = n_cols * px + 1;
    height = n_rows * px + 1;
    occ.assign((size_t)width * height, 0);
    auto hline = [&](int row, int c0, int c1) { for (int c = c0; c <= c1; ++c) occ[(size_t)row * width + c] = 100; };
    auto vline = [&](int col, int r0, int r1) { for (int r = r0; r <= r1; ++r) occ[(size_t)r * width + col] = 100; };
    for (int cy = 0; cy < n_rows; ++cy) {
        const auto& row = g.rows[cy];
        const int len = (int)row.size();
        const int top = cy * px;
        for (int cx = 0; cx < len; ++cx) {
            const int left = cx * px;
            const uint8_t w = row[cx];
            if (w & 2) hline(top, left, left + px);
            if (w & 4) vline(left, top, top + px);
            if ((w & 1) && cx == len - 1) vline(left + px, top, top + px - 1);
            const bool nothing_below = (cy == n_rows - 1) || cx >= (int)g.rows[cy + 1].size();
            if ((w & 8) && nothing_below) hline(top + px, left + 1, left + px);
        }
        for (int cx = len; cx < n_cols; ++cx)
            for (int r = top; r < top + px; ++r) hline(r, cx * px, cx * px + px);
    }
}

// ---- Gaussian lookup (MC:139-177) --------------------------------------------------------------------------
struct GaussTable {
    double sigma, lo, hi, step;
    std::vector<double> v;
    void build(double sigma_) {
        sigma = sigma_;
        lo = 0.0f; hi = 1.1f; step = 0.0001f;               // float literals widened, as the reference holds them
        const int n = static_cast<int>((hi - lo) / step) + 1;
        v.resize(n);
        const double denom = sigma * std::sqrt(2.0f * M_PI);
        for (int i = 0; i < n; ++i) {
            const double d = lo + i * step;
            v[i] = std::exp(-(d * d) / (2 * sigma * sigma)) / denom;
        }
    }
};

// ---- scan preprocessing (MC:254-278, 610-620) --------------------------------------------------------------
struct HostBeam { double radius, angle; };
inline void filter_scan(const float* ranges, int n, float angle_min, float angle_inc, float range_min, float range_max,
                        bool use_fov, double lower_deg, double upper_deg, std::vector<HostBeam>& out) {
    out.clear();
    const double a0 = angle_min, da = angle_inc;
    for (size_t i = 0; i < (size_t)n; ++i) {
        const double r = ranges[i];
        const double ang = a0 + (i * da);
        double keep;
        if (std::isnan(r) || std::isinf(r)) keep = 1.05;
        else if (r >= range_min && r <= range_max) keep = r;
        else continue;
        if (use_fov) {
            const double deg = ang * 180.0 / M_PI;
            if (!(deg > lower_deg && deg < upper_deg)) continue;
        }
        out.push_back({keep, ang});
    }
}

// ---- odometry (MC:695-739) -----------------------------------------------------------------------------------
struct Pose2 { double x = 0, y = 0, theta = 0; };
struct Motion { double rot_1 = 0, trans = 0, rot_2 = 0; };
struct OdometryState {
    double enc_l_prev = 0, enc_r_prev = 0;
    Pose2 prev;
};
// z[3] are the standard-normal draws of the three sample() calls, in call order.
inline Motion odometry_step(OdometryState& st, const mcl_config& cfg, double enc_l, double enc_r, const double z[3]) {
    const double d_left = (enc_l - st.enc_l_prev) * cfg.wheel_size * 0.5;
    const double d_right = (enc_r - st.enc_r_prev) * cfg.wheel_size * 0.5;
    const double d_center = 0.5 * (d_left + d_right);
    const double delta_theta = (d_left - d_right) / cfg.wheel_space;
    const double th_new = delta_theta + st.prev.theta;
    Pose2 cur;
    cur.x = st.prev.x + d_center * std::cos(st.prev.theta + 0.5 * delta_theta);
    cur.y = st.prev.y + d_center * std::sin(st.prev.theta + 0.5 * delta_theta);
    cur.theta = std::atan2(std::sin(th_new), std::cos(th_new));
    Motion m;
    const double ddx = cur.x - st.prev.x, ddy = cur.y - st.prev.y;
    m.rot_1 = std::atan2(ddy, ddx) - st.prev.theta;
    m.trans = std::sqrt(ddy * ddy + ddx * ddx);
    m.rot_2 = cur.theta - st.prev.theta - m.rot_1;
    auto noise = [](double variance, double zz) { return zz * std::sqrt(variance) + 0.0; };   // normal_distribution: z*stddev+mean
    const double r1 = m.rot_1 + noise(cfg.alpha[0] * std::fabs(m.rot_1) + cfg.alpha[1] * m.trans, z[0]);
    const double tn = noise(cfg.alpha[2] * m.trans + cfg.alpha[3] * (std::fabs(m.rot_1) + std::fabs(m.rot_2)), z[1]);
    const double tr = m.trans + tn;
    const double r2 = m.rot_2 + noise(cfg.alpha[0] * std::fabs(m.rot_2) + cfg.alpha[1] * m.trans, z[2]);
    m.rot_1 = r1; m.trans = tr; m.rot_2 = r2;
    st.prev = cur;
    st.enc_l_prev = enc_l;
    st.enc_r_prev = enc_r;
    return m;
}

// ---- tf yaw round trip on the host (Q8), for the first-touch ray directions ------------------------------------
inline double tf_yaw_roundtrip(double theta) {
    const double h = theta * 0.5;
    const double sy = std::sin(h), cy = std::cos(h);
    // setRPY(0,0,yaw): roll = pitch = 0 so x = y = 0, z = sy, w = cy (products with exact 0 and 1)
    const double d = sy * sy + cy * cy;
    const double s = 2.0 / d;
    const double zs = sy * s;
    const double wz = cy * zs, zz = sy * zs;
    const double m00 = 1.0 - zz, m10 = wz;
    return std::atan2(m10, m00);
}

// publishPosMsg (MC:958-994) with wrapTo2Pi (MC:951-957): maze cell and 4-way direction of a world pose.
inline void pose_to_cell(double wx, double wy, double angle, double cell_meters, int& row, int& col, int& dir) {
    if (wx < 0 || wy < 0) { row = col = dir = -1; return; }                       // "not localised" sentinel, MC:964-971
    const double col_wx = (wx - 0.5 * cell_meters) / cell_meters;                 // MC:973-974
    const double row_wx = (wy - 0.5 * cell_meters) / cell_meters;
    col = static_cast<int>(std::floor(col_wx + 0.5));                             // MC:976-977
    row = static_cast<int>(std::floor(row_wx + 0.5));
    const double two_pi = 2.0 * M_PI;
    double wrapped = std::fmod(angle, two_pi);
    if (wrapped < 0) wrapped += two_pi;
    const double deg = wrapped * 180.0 / M_PI;                                    // MC:978
    if (deg >= 45 && deg < 135) dir = 3;                                          // DOWN   (MC:980-987; msg/Pose.msg constants)
    else if (deg >= 135 && deg < 225) dir = 2;                                    // LEFT
    else if (deg >= 225 && deg < 315) dir = 1;                                    // UP
    else dir = 0;                                                                 // RIGHT
}

}  // namespace mcl
