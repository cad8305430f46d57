// mcl_oracle_ns.cpp — CPU ORACLE for MCL_MODE_NS (test infrastructure, NOT product code).
//
// NS mode (likelihood field over a distance transform, per-particle Philox motion noise, fixed-point systematic
// resampling) is the engine's own north-star formulation: the reference has none of it, so
// PARITY IS UNPINNED BY THE REFERENCE here. This file is an independent, single-threaded restatement of the NS
// definitions in DESIGN.md ("NS-1".."NS-8"), written without including any product header and, where possible, with a
// different algorithm (window-search distance transform instead of separable passes; division-based systematic
// thresholds instead of cross-multiplied comparisons), so agreement with the CUDA engine is evidence, not tautology.
// What it borrows from the reference: the odometry increment and noise variances (MC:695-739), the beam-angle
// mirroring and laser offset (MC:644-653), and the particle record layout.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

namespace {

typedef unsigned __int128 u128;

// ---- Philox4x32-10 (Salmon et al., SC'11), restated ------------------------------------------------------------------
void philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t out[4]) {
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// ---- deterministic elementary functions: IEEE basic ops + fma only, fixed order -----------------------
const double INV_FACT[] = {1.0, 1.0, 1.0 / 2.0, 1.0 / 6.0, 1.0 / 24.0, 1.0 / 120.0, 1.0 / 720.0, 1.0 / 5040.0, 1.0 / 40320.0, 1.0 / 362880.0,
                           1.0 / 3628800.0, 1.0 / 39916800.0, 1.0 / 479001600.0, 1.0 / 6227020800.0, 1.0 / 87178291200.0,
                           1.0 / 1307674368000.0, 1.0 / 20922789888000.0};

void sincos_small(double x, double& s, double& c) {           // |x| <= pi/4, Taylor via Horner in x^2
    double x2 = x * x;
    double ps = -INV_FACT[15];
    const int sodd[] = {13, 11, 9, 7, 5, 3};
    double sign = 1.0;
    for (int n : sodd) { ps = fma(ps, x2, sign * INV_FACT[n]); sign = -sign; }
    s = fma(ps * x2, x, x);
    double pc = INV_FACT[16];
    const int ceven[] = {14, 12, 10, 8, 6, 4, 2};
    sign = -1.0;
    for (int n : ceven) { pc = fma(pc, x2, sign * INV_FACT[n]); sign = -sign; }
    c = fma(pc, x2, 1.0);
}
void det_sincos(double t, double& s, double& c) {
    double kf = rint(t * 0.63661977236758134308);
    double r = fma(-kf, 1.57079632673412561417e+00, t);
    r = fma(-kf, 6.07710050630396597660e-11, r);
    r = fma(-kf, 2.02226624879595063154e-21, r);
    double sr, cr;
    sincos_small(r, sr, cr);
    long long q = ((long long)kf) & 3;
    if (q == 0) { s = sr; c = cr; } else if (q == 1) { s = cr; c = -sr; } else if (q == 2) { s = -sr; c = -cr; } else { s = -cr; c = sr; }
}
double det_log(double x) {
    int e;
    double m = frexp(x, &e);          // m in [0.5,1)
    m *= 2.0; e -= 1;                 // m in [1,2)
    if (m > 1.4142135623730951) { m *= 0.5; e += 1; }
    double t = (m - 1.0) / (m + 1.0), t2 = t * t;
    double p = 1.0 / 23.0;
    for (int d = 21; d >= 3; d -= 2) p = fma(p, t2, 1.0 / (double)d);
    p = fma(p * t2, t, t);
    return fma((double)e, 0.69314718055994530942, 2.0 * p);
}
// fp32 elementary functions of the motion noise: IEEE +,*,/,sqrt,fma in fp32 only, fixed order.
// Coefficient tables (Horner, highest power first); literals are the decimal strings of DESIGN.md.
const float SIN_C[] = {2.75573192e-06f, -1.98412701e-04f, 8.33333377e-03f, -1.66666672e-01f};                 // 1/9! -1/7! 1/5! -1/3!
const float COS_C[] = {-2.75573200e-07f, 2.48015876e-05f, -1.38888892e-03f, 4.16666679e-02f, -0.5f};           // -1/10! 1/8! -1/6! 1/4! -1/2!
const float ATANH_C[] = {0.111111112f, 0.142857149f, 0.200000003f, 0.333333343f};                              // 1/9 1/7 1/5 1/3

void sincos_small_f(float x, float& s, float& c) {
    const float x2 = x * x;
    float ps = SIN_C[0];
    for (int i = 1; i < 4; i++) ps = fmaf(ps, x2, SIN_C[i]);
    s = fmaf(ps * x2, x, x);
    float pc = COS_C[0];
    for (int i = 1; i < 5; i++) pc = fmaf(pc, x2, COS_C[i]);
    c = fmaf(pc, x2, 1.0f);
}
void rotate_quadrant_f(int q, float sr, float cr, float& s, float& c) {
    q &= 3;
    if (q == 0) { s = sr; c = cr; } else if (q == 1) { s = cr; c = -sr; } else if (q == 2) { s = -sr; c = -cr; } else { s = -cr; c = sr; }
}
void det_sincos_f32(float t, float& s, float& c) {                 // Cody-Waite by pi/2 = 1.5703125 + 4.8375129699707e-4 + 7.5497899548919e-8
    const float kf = rintf(t * 0.636619747f);
    float r = fmaf(-kf, 1.5703125f, t);
    r = fmaf(-kf, 4.837512969970703125e-4f, r);
    r = fmaf(-kf, 7.54978995489188216e-8f, r);
    float sr, cr;
    sincos_small_f(r, sr, cr);
    rotate_quadrant_f((int)kf, sr, cr, s, c);
}
void det_sincos_turns_f32(float u, float& s, float& c) {           // sin, cos of 2 pi u, u in [0,1)
    const float q = rintf(u * 4.0f);
    const float f = fmaf(q, -0.25f, u);
    const float x = fmaf(f, 6.28318548f, f * -1.74845553e-07f);
    float sr, cr;
    sincos_small_f(x, sr, cr);
    rotate_quadrant_f((int)q, sr, cr, s, c);
}
float det_log_f32(float x) {
    int e;
    float m = frexpf(x, &e) * 2.0f;   // m in [1,2)
    e -= 1;
    if (m > 1.41421354f) { m *= 0.5f; e += 1; }
    const float t = (m + -1.0f) / (m + 1.0f), t2 = t * t;
    float p = ATANH_C[0];
    for (int i = 1; i < 4; i++) p = fmaf(p, t2, ATANH_C[i]);
    p = fmaf(p * t2, t, t);
    return fmaf((float)e, 0.693147182f, 2.0f * p);
}
void det_normal_pair(uint32_t w1, uint32_t w2, float& z0, float& z1) {
    const float u1 = ((float)(w1 >> 9) + 0.5f) * 1.1920929e-07f, u2 = ((float)(w2 >> 9) + 0.5f) * 1.1920929e-07f;
    const float r = sqrtf(-2.0f * det_log_f32(u1));
    float s, c;
    det_sincos_turns_f32(u2, s, c);
    z0 = r * c;
    z1 = r * s;
}
const float EXP_C[] = {1.98412701e-04f, 1.38888892e-03f, 8.33333377e-03f, 4.16666679e-02f, 1.66666672e-01f, 0.5f, 1.0f, 1.0f};   // 1/7! .. 1/0!
uint64_t det_exp_q32(float t) {                                // NS-4: trunc(2^32 * exp(t)), exp in fp32 (IEEE ops only)
    if (!(t > -22.5f)) return 0;
    if (t >= 0.f) return 1ull << 32;
    const float kf = rintf(t * 1.44269502f);
    float g = fmaf(-kf, 0.693145752f, t);
    g = fmaf(-kf, 1.42860677e-06f, g);
    float p = EXP_C[0];
    for (int i = 1; i < 8; i++) p = fmaf(p, g, EXP_C[i]);
    const float scaled = ldexpf(p, 32 + (int)kf);
    uint64_t w = (uint64_t)scaled;
    return std::min<uint64_t>(w, 1ull << 32);
}
float wrap_pi(float t) {                                       // one conditional turn each way (NS-7)
    const float PI_F = 3.14159274f, TWO_PI_F = 6.28318548f;
    if (t > PI_F) t = t - TWO_PI_F;
    if (t < -PI_F) t = t + TWO_PI_F;
    return t;
}

struct NsCtx {
    int W = 0, H = 0, R = 0;
    float res = 0;
    double ox = 0, oy = 0;
    std::vector<uint16_t> d2;
    std::vector<float> lf;
    float lf_out = 0;
    // config
    double sigma = 0.1, z_hit = 0.8, z_rand = 0.2, max_range = 5.6, laser_offset = 0.1, temper = 0.05;
    double alpha[4] = {0.001, 0.001, 0.0001, 0.0001};
    uint64_t seed = 0x9E3779B97F4A7C15ull;
    int beam_stride = 1;
};

}  // namespace

extern "C" {

void* ons_create() { return new NsCtx(); }
void ons_destroy(void* h) { delete (NsCtx*)h; }
void ons_config(void* h, double sigma, double z_hit, double z_rand, double max_range, double laser_offset, double temper, uint64_t seed,
                int beam_stride) {
    NsCtx& c = *(NsCtx*)h;
    c.sigma = sigma; c.z_hit = z_hit; c.z_rand = z_rand; c.max_range = max_range; c.laser_offset = laser_offset; c.temper = temper;
    c.seed = seed; c.beam_stride = beam_stride;
}

// NS-1: capped squared distance (cells) to the nearest occupied cell by direct window search, then the field.
void ons_set_map(void* h, const int8_t* occ, int W, int H, float res, double ox, double oy) {
    NsCtx& c = *(NsCtx*)h;
    c.W = W; c.H = H; c.res = res; c.ox = ox; c.oy = oy;
    c.R = std::min(255, std::max(1, (int)std::ceil(2.0 / (double)res)));
    const int R = c.R, cap = R * R;
    c.d2.assign((size_t)W * H, (uint16_t)cap);
    // scatter from occupied cells: every occupied cell lowers the cells inside its disc
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) {
            if (!(occ[(size_t)y * W + x] > 50)) continue;
            for (int dy = -R; dy <= R; dy++) {
                int yy = y + dy;
                if (yy < 0 || yy >= H) continue;
                for (int dx = -R; dx <= R; dx++) {
                    int xx = x + dx;
                    if (xx < 0 || xx >= W) continue;
                    int v = dx * dx + dy * dy;
                    if (v < c.d2[(size_t)yy * W + xx]) c.d2[(size_t)yy * W + xx] = (uint16_t)v;
                }
            }
        }
    std::vector<float> table(cap + 1);
    for (int q = 0; q <= cap; q++) {
        double d = (double)res * std::sqrt((double)q);
        double p = c.z_hit * std::exp(-(d * d) / (2.0 * c.sigma * c.sigma)) / (c.sigma * std::sqrt(2.0 * M_PI)) + c.z_rand / c.max_range;
        table[q] = (float)std::log(p);
    }
    c.lf.resize((size_t)W * H);
    for (size_t i = 0; i < c.lf.size(); i++) c.lf[i] = table[c.d2[i]];
    c.lf_out = table[cap];
}
void ons_get_field(void* h, float* lf, uint16_t* d2) {
    NsCtx& c = *(NsCtx*)h;
    if (lf) memcpy(lf, c.lf.data(), c.lf.size() * 4);
    if (d2) memcpy(d2, c.d2.data(), c.d2.size() * 2);
}

// uniform initial particles keyed by global index (NS "init")
void ons_init(void* h, int64_t g0, int64_t n, float* P) {
    NsCtx& c = *(NsCtx*)h;
    double ext_x = (double)c.W * (double)c.res, ext_y = (double)c.H * (double)c.res;
    for (int64_t i = 0; i < n; i++) {
        uint64_t g = (uint64_t)(g0 + i);
        uint32_t r[4];
        philox((uint32_t)g, (uint32_t)(g >> 32), 0x60u, 0u, (uint32_t)c.seed, (uint32_t)(c.seed >> 32), r);
        double u1 = ((double)r[0] + 0.5) * 2.3283064365386963e-10, u2 = ((double)r[1] + 0.5) * 2.3283064365386963e-10,
               u3 = ((double)r[2] + 0.5) * 2.3283064365386963e-10;
        P[4 * i + 0] = (float)(c.ox + u1 * ext_x);
        P[4 * i + 1] = (float)(c.oy + u2 * ext_y);
        P[4 * i + 2] = (float)(-3.14159265358979323846 + u3 * 6.28318530717958647692);
        P[4 * i + 3] = 1.0f;
    }
}

// NS-7: odometry increment (rot1, trans, rot2) + per-particle noise with the reference's variances
void ons_predict(void* h, float* P, int64_t g0, int64_t n, double rot1, double trans, double rot2, uint32_t step) {
    NsCtx& c = *(NsCtx*)h;
    float r1m = (float)rot1, trm = (float)trans, r2m = (float)rot2;
    float sd1 = (float)std::sqrt(c.alpha[0] * std::fabs(rot1) + c.alpha[1] * std::fabs(trans));
    float sdt = (float)std::sqrt(c.alpha[2] * std::fabs(trans) + c.alpha[3] * (std::fabs(rot1) + std::fabs(rot2)));
    float sd2 = (float)std::sqrt(c.alpha[0] * std::fabs(rot2) + c.alpha[1] * std::fabs(trans));
    for (int64_t i = 0; i < n; i++) {
        uint64_t g = (uint64_t)(g0 + i);
        uint32_t r[4];
        philox((uint32_t)g, (uint32_t)(g >> 32), 0x50u, step, (uint32_t)c.seed, (uint32_t)(c.seed >> 32), r);
        float z0, z1, z2, z3;
        det_normal_pair(r[0], r[1], z0, z1);
        det_normal_pair(r[2], r[3], z2, z3);
        float* p = P + 4 * i;
        float r1 = fmaf(z0, sd1, r1m), tr = fmaf(z1, sdt, trm), r2 = fmaf(z2, sd2, r2m);
        float s, cs;
        det_sincos_f32(p[2] + r1, s, cs);
        p[0] = fmaf(tr, cs, p[0]);
        p[1] = fmaf(tr, s, p[1]);
        p[2] = wrap_pi(p[2] + (r1 + r2));
    }
}

// NS-2: scan -> beam points in the robot frame. Returns the number of beams kept.
int ons_beams(void* h, const float* ranges, int B, float angle_min, float angle_inc, float range_min, float range_max, float* pts, int cap) {
    NsCtx& c = *(NsCtx*)h;
    int kept = 0, out = 0;
    for (int i = 0; i < B; i++) {
        double r = ranges[i];
        if (std::isnan(r) || std::isinf(r)) continue;
        if (!(r >= range_min && r <= range_max) || r >= c.max_range) continue;
        double ang = (double)angle_min + ((size_t)i * (double)angle_inc);
        if ((kept++ % std::max(1, c.beam_stride)) != 0) continue;
        double phi = -ang;
        double inv_res = 1.0 / (double)c.res;           // beam points in CELL units
        if (out < cap) { pts[2 * out] = (float)((c.laser_offset + r * std::cos(phi)) * inv_res); pts[2 * out + 1] = (float)((r * std::sin(phi)) * inv_res); }
        out++;
    }
    return out;
}

// NS-3: particle position in cell units, g0 = fma(x - ox, 1/res, -0.5); endpoint g = fma(c,bx, fma(-s,by, g0)); cell =
// round-to-nearest-even(g) (so the -0.5 makes it the containing cell); per-particle log-likelihood = 32 lane-strided
// fp32 partial sums combined by an xor butterfly (16,8,4,2,1)
void ons_loglik(void* h, const float* P, int64_t n, const float* pts, int nb, float* ll) {
    NsCtx& c = *(NsCtx*)h;
    float oxf = (float)c.ox, oyf = (float)c.oy, inv_res = 1.0f / c.res;
    for (int64_t i = 0; i < n; i++) {
        const float* p = P + 4 * i;
        double sd, cd;
        det_sincos((double)p[2], sd, cd);
        float s = (float)sd, cs = (float)cd;
        float gx0 = fmaf(p[0] + -oxf, inv_res, -0.5f), gy0 = fmaf(p[1] + -oyf, inv_res, -0.5f);
        float lane[32];
        for (int l = 0; l < 32; l++) {
            float acc = 0.f;
            for (int b = l; b < nb; b += 32) {
                float bx = pts[2 * b], by = pts[2 * b + 1];
                float gx = fmaf(cs, bx, fmaf(-s, by, gx0));
                float gy = fmaf(s, bx, fmaf(cs, by, gy0));
                float v = c.lf_out;
                if (fabsf(gx) < 4194304.f && fabsf(gy) < 4194304.f) {         // |g| < 2^22: lrintf is the magic-add result
                    long ix = lrintf(gx), iy = lrintf(gy);                     // ties to even (default rounding mode)
                    if (ix >= 0 && iy >= 0 && ix < c.W && iy < c.H) v = c.lf[(size_t)iy * c.W + ix];
                }
                acc = acc + v;
            }
            lane[l] = acc;
        }
        for (int o = 16; o > 0; o >>= 1) {
            float nxt[32];
            for (int l = 0; l < 32; l++) nxt[l] = lane[l] + lane[l ^ o];
            memcpy(lane, nxt, sizeof(lane));
        }
        ll[i] = lane[0];
    }
}

// NS-4/5: Q32 weights and their inclusive prefix (local to the span given); returns the span total
uint64_t ons_weights(void* h, const float* ll, int64_t n, float max_ll, uint64_t* W, uint64_t* prefix, float* wfloat) {
    NsCtx& c = *(NsCtx*)h;
    float temper = (float)c.temper;
    uint64_t run = 0;
    for (int64_t i = 0; i < n; i++) {
        uint64_t w = det_exp_q32(temper * (ll[i] + -max_ll));
        run += w;
        if (W) W[i] = w;
        if (prefix) prefix[i] = run;
        if (wfloat) wfloat[i] = (float)((double)w * 2.3283064365386963e-10);
    }
    return run;
}

uint32_t ons_u0(void* h, uint32_t step) {
    NsCtx& c = *(NsCtx*)h;
    uint32_t o[4];
    philox(0u, 0u, 0x40u, step, (uint32_t)c.seed, (uint32_t)(c.seed >> 32), o);
    return o[0];
}

// NS-6: systematic resampling over the GLOBAL prefix C[0..N): slot k takes the first i with C_i > t_k,
// t_k = floor(((k<<32)+u0) * T / (N<<32)). Division-based on purpose (the engine cross-multiplies).
void ons_resample(const uint64_t* C, int64_t N, uint32_t u0, int64_t k_begin, int64_t k_end, int64_t* ancestors) {
    const uint64_t T = C[N - 1];
    const u128 D = (u128)N << 32;
    for (int64_t k = k_begin; k < k_end; k++) {
        u128 X = ((u128)(uint64_t)k << 32) + u0;
        uint64_t t = (uint64_t)((X * T) / D);
        ancestors[k - k_begin] = std::upper_bound(C, C + N, t) - C;
    }
}

// NS-8: weighted pose sums {sum w, sum w x, sum w y, sum w sin, sum w cos}
void ons_pose_partials(const float* P, int64_t n, double* out5) {
    double a[5] = {0, 0, 0, 0, 0};
    for (int64_t i = 0; i < n; i++) {
        const float* p = P + 4 * i;
        double sd, cd;
        det_sincos((double)p[2], sd, cd);
        double w = p[3];
        a[0] += w; a[1] += w * (double)p[0]; a[2] += w * (double)p[1]; a[3] += w * (double)(float)sd; a[4] += w * (double)(float)cd;
    }
    memcpy(out5, a, sizeof(a));
}

// exposed for unit tests of the deterministic math
void ons_det_sincos(double t, double* s, double* c) { det_sincos(t, *s, *c); }
double ons_det_log(double x) { return det_log(x); }
void ons_det_sincos_f32(float t, float* s, float* c) { det_sincos_f32(t, *s, *c); }
void ons_det_sincos_turns_f32(float u, float* s, float* c) { det_sincos_turns_f32(u, *s, *c); }
float ons_det_log_f32(float x) { return det_log_f32(x); }
void ons_normal_pairs(const uint32_t* w1, const uint32_t* w2, int64_t n, float* z0, float* z1) {
    for (int64_t i = 0; i < n; i++) det_normal_pair(w1[i], w2[i], z0[i], z1[i]);
}
uint64_t ons_det_exp_q32(float t) { return det_exp_q32(t); }
void ons_philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t* out) { philox(c0, c1, c2, c3, k0, k1, out); }

}  // extern "C"
