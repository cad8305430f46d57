"""ctypes bindings for the CPU oracle (oracle/_build/libmcl_oracle.so) and, when present, the compiled
reference (oracle/_ref/libmclref.so).

TEST INFRASTRUCTURE ONLY. Importable from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs. The product package montecarlolocalisation_b200 never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "_build", "libmcl_oracle.so")
REF_SO = os.path.join(HERE, "_ref", "libmclref.so")

c_dp = C.POINTER(C.c_double)
c_fp = C.POINTER(C.c_float)
c_ip = C.POINTER(C.c_int)
c_bp = C.POINTER(C.c_int8)


def build(force=False):
    """Compile the oracle (and oracle/_ref when /root/reference exists). Building the checker is not using it."""
    srcs = [os.path.join(HERE, f) for f in ("mcl_oracle.cpp", "mcl_oracle_ns.cpp", "Makefile")]
    stale = force or not os.path.exists(ORACLE_SO) or any(
        os.path.getmtime(s) > os.path.getmtime(ORACLE_SO) for s in srcs)
    ref_src = "/root/reference/pink_fundamentals/src/monte_carlo.cpp"
    ref_deps = [os.path.join(HERE, "ref_harness.cpp"), os.path.join(HERE, "shim", "shim_prelude.h"), os.path.join(HERE, "shim", "ros", "ros.h")]
    ref_stale = os.path.exists(ref_src) and (force or not os.path.exists(REF_SO) or
                                             any(os.path.getmtime(d) > os.path.getmtime(REF_SO) for d in ref_deps))
    if stale or ref_stale:
        subprocess.run(["make", "-s", "-C", HERE, "-B"], check=True)


def _p(a, t):
    return a.ctypes.data_as(t) if a is not None else None


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


class Scan:
    """sensor_msgs/LaserScan fields the path reads (float32 on the wire)."""

    def __init__(self, ranges, angle_min, angle_inc, range_min, range_max):
        self.ranges = np.ascontiguousarray(ranges, dtype=np.float32)
        self.angle_min = np.float32(angle_min)
        self.angle_inc = np.float32(angle_inc)
        self.range_min = np.float32(range_min)
        self.range_max = np.float32(range_max)

    def args(self):
        return (_p(self.ranges, c_fp), C.c_int(len(self.ranges)), C.c_float(self.angle_min), C.c_float(self.angle_inc),
                C.c_float(self.range_min), C.c_float(self.range_max))


_oracle_lib = None


def oracle_lib():
    global _oracle_lib
    if _oracle_lib is None:
        if not os.path.exists(ORACLE_SO):
            build()
        L = C.CDLL(ORACLE_SO)
        L.orc_create.restype = C.c_void_p
        L.orc_clamp_count.restype = C.c_longlong
        for name in ("orc_gauss_get", "orc_yaw_roundtrip", "orc_raycast", "orc_compute_weight"):
            getattr(L, name).restype = C.c_double
        _oracle_lib = L
    return _oracle_lib


def libm_trigf(x):
    """(sinf(x), cosf(x)) of this process's libm for a float32 array."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    s = np.empty_like(x); c = np.empty_like(x)
    oracle_lib().orc_libm_trigf(x.ctypes.data_as(C.POINTER(C.c_float)), C.c_longlong(x.size), s.ctypes.data_as(C.POINTER(C.c_float)),
                                c.ctypes.data_as(C.POINTER(C.c_float)))
    return s, c


class Oracle:
    """One reference-process worth of global state (map, LUTs, motion model, injection EMA)."""

    def __init__(self, trig_mode=1):
        self.L = oracle_lib()
        self.h = C.c_void_p(self.L.orc_create())
        self.L.orc_set_trig_mode(self.h, trig_mode)

    def __del__(self):
        try:
            self.L.orc_destroy(self.h)
        except Exception:
            pass

    # -- map ------------------------------------------------------------------------------------
    @staticmethod
    def rasterise_map_txt(text):
        L = oracle_lib()
        cap = 1 << 22
        buf = np.zeros(cap, dtype=np.int8)
        w, h = C.c_int(), C.c_int()
        rc = L.orc_rasterise_map_txt(text.encode(), _p(buf, c_bp), cap, C.byref(w), C.byref(h))
        if rc:
            raise ValueError("map.txt parse error %d" % rc)
        return buf[: w.value * h.value].reshape(h.value, w.value).copy()

    def set_map(self, occ, res=np.float32(0.1), ox=0.0, oy=0.0):
        occ = np.ascontiguousarray(occ, dtype=np.int8)
        self.map_h, self.map_w = occ.shape
        self.L.orc_set_map(self.h, _p(occ, c_bp), self.map_w, self.map_h, C.c_float(res), C.c_double(ox), C.c_double(oy))

    def cell_ranges(self):
        r, c = C.c_int(), C.c_int()
        self.L.orc_cell_ranges(self.h, C.byref(r), C.byref(c))
        return r.value, c.value

    def precompute_ray_directions(self, lo=-120.0, hi=120.0, step=0.1):
        self.L.orc_precompute_ray_directions(self.h, C.c_double(lo), C.c_double(hi), C.c_double(step))

    def ray_lut(self, lo=-100000, hi=100000):
        n = self.L.orc_ray_lut_dump(self.h, lo, hi, None, None, None, 0)
        k = np.zeros(n, np.int32); dx = np.zeros(n); dy = np.zeros(n)
        self.L.orc_ray_lut_dump(self.h, lo, hi, _p(k, c_ip), _p(dx, c_dp), _p(dy, c_dp), n)
        return k, dx, dy

    def ray_lut_set(self, key, dx, dy):
        self.L.orc_ray_lut_set(self.h, int(key), C.c_double(dx), C.c_double(dy))

    def gauss_get(self, d):
        return self.L.orc_gauss_get(self.h, C.c_double(d))

    def gauss_table(self):
        n = self.L.orc_gauss_size(self.h)
        t = np.zeros(n)
        self.L.orc_gauss_table(self.h, _p(t, c_dp))
        return t

    @staticmethod
    def filter_scan(scan, lower=-120.0, upper=120.0):
        L = oracle_lib()
        cap = len(scan.ranges)
        r = np.zeros(cap); a = np.zeros(cap)
        n = L.orc_filter_scan(*scan.args(), C.c_double(lower), C.c_double(upper), _p(r, c_dp), _p(a, c_dp), cap)
        return r[:n].copy(), a[:n].copy()

    def is_valid_pos(self, x, y):
        return bool(self.L.orc_is_valid_pos(self.h, C.c_double(x), C.c_double(y)))

    def is_occupied(self, x, y):
        return bool(self.L.orc_is_occupied(self.h, C.c_double(x), C.c_double(y)))

    def yaw_roundtrip(self, t):
        return self.L.orc_yaw_roundtrip(C.c_double(t))

    def raycast(self, x, y, theta, off_deg, max_range=1.0):
        return self.L.orc_raycast(self.h, C.c_double(x), C.c_double(y), C.c_double(theta), C.c_double(off_deg), C.c_double(max_range))

    # -- filter stages ----------------------------------------------------------------------------
    def sample_particles(self, u_yaw, row, col, u_dx, u_dy):
        n = len(u_yaw)
        P = np.zeros((n, 4), np.float32)
        u_yaw, u_dx, u_dy, row, col = _f64(u_yaw), _f64(u_dx), _f64(u_dy), _i32(row), _i32(col)
        self.L.orc_sample_particles(self.h, n, _p(u_yaw, c_dp), _p(row, c_ip), _p(col, c_ip), _p(u_dx, c_dp), _p(u_dy, c_dp), _p(P, c_fp))
        return P

    def diff_drive(self, enc_left, enc_right, z3):
        self.L.orc_set_encoders(self.h, C.c_double(enc_left), C.c_double(enc_right))
        z = _f64(z3); out = np.zeros(3)
        self.L.orc_diff_drive(self.h, _p(z, c_dp), _p(out, c_dp))
        return out

    def set_motion(self, rot1, trans, rot2):
        self.L.orc_set_motion(self.h, C.c_double(rot1), C.c_double(trans), C.c_double(rot2))

    def update_particle_pos(self, P):
        assert P.dtype == np.float32 and P.flags.c_contiguous
        self.L.orc_update_particle_pos(self.h, _p(P, c_fp), len(P))

    def compute_weight(self, P, scan):
        assert P.dtype == np.float32 and P.flags.c_contiguous
        return self.L.orc_compute_weight(self.h, _p(P, c_fp), len(P), *scan.args())

    def resample(self, P, jitter_state, scan, u_r, u_jit, inj=None):
        """Returns (Pout, idx, cdf, stats dict). P's weight row is normalised in place (as in the reference)."""
        n = len(P)
        Pout = np.zeros((n, 4), np.float32)
        idx = np.zeros(n, np.int32)
        cdf = np.zeros(n)
        stats = np.zeros(5)
        u_r, u_jit = _f64(u_r), _f64(u_jit)
        if inj is None:
            z = np.zeros(1)
            inj = dict(u_yaw=z, row=np.zeros(1, np.int32), col=np.zeros(1, np.int32), u_dx=z, u_dy=z)
        iy, ir, ic, ix, iyy = _f64(inj["u_yaw"]), _i32(inj["row"]), _i32(inj["col"]), _f64(inj["u_dx"]), _f64(inj["u_dy"])
        self.L.orc_resample(self.h, _p(P, c_fp), n, int(bool(jitter_state)), *scan.args(), _p(u_r, c_dp), _p(u_jit, c_dp),
                            _p(iy, c_dp), _p(ir, c_ip), _p(ic, c_ip), _p(ix, c_dp), _p(iyy, c_dp),
                            _p(Pout, c_fp), _p(idx, c_ip), _p(cdf, c_dp), _p(stats, c_dp))
        return Pout, idx, cdf, dict(injected=int(stats[0]), p_inject=stats[1], weight_slow=stats[2], weight_fast=stats[3], total_weight=stats[4])

    def injection_state(self):
        out = np.zeros(2)
        self.L.orc_get_injection_state(self.h, _p(out, c_dp))
        return out

    def set_injection_state(self, slow, fast):
        self.L.orc_set_injection_state(self.h, C.c_double(slow), C.c_double(fast))

    def clamp_count(self):
        return self.L.orc_clamp_count(self.h)

    def estimate_weighted_pose(self, P):
        out = np.zeros(3)
        self.L.orc_estimate_weighted_pose(self.h, _p(np.ascontiguousarray(P, np.float32), c_fp), len(P), _p(out, c_dp))
        return out


def _P32(P):
    P = np.ascontiguousarray(P, dtype=np.float32)
    assert P.ndim == 2 and P.shape[1] == 4
    return P


def kmeans(P, init_idx, reinit_idx=(), K=3, max_iters=20):
    """kMeansClustering (MC:802-868) with the rand() draws injected. Returns (assignments, centers[K,2], passes, reinit_used)."""
    L = oracle_lib(); P = _P32(P)
    a = np.zeros(len(P), np.int32); c = np.zeros((K, 2), np.float32); used = C.c_int()
    ii = _i32(init_idx); ri = _i32(reinit_idx if len(reinit_idx) else [0])
    passes = L.orc_kmeans(_p(P, c_fp), len(P), K, max_iters, _p(ii, c_ip), _p(ri, c_ip), len(reinit_idx), _p(a, c_ip), _p(c, c_fp), C.byref(used))
    return a, c, passes, used.value


def kmeans_confidence(P, init_idx, reinit_idx=(), ratio_threshold=0.6):
    """isLocalizationLost_densitiy_cluster (MC:886-949). Returns dict(ratio, best[3], centers, assignments, best_cluster, passes, ...)."""
    L = oracle_lib(); P = _P32(P)
    L.orc_kmeans_confidence.restype = C.c_double
    a = np.zeros(len(P), np.int32); c = np.zeros((3, 2), np.float32); out = np.zeros(3); info = np.zeros(3, np.int32); cw = np.zeros(3)
    ii = _i32(init_idx); ri = _i32(reinit_idx if len(reinit_idx) else [0])
    ratio = L.orc_kmeans_confidence(_p(P, c_fp), len(P), _p(ii, c_ip), _p(ri, c_ip), len(reinit_idx), C.c_double(ratio_threshold), _p(out, c_dp),
                                    _p(c, c_fp), _p(a, c_ip), _p(info, c_ip), _p(cw, c_dp))
    return dict(ratio=ratio, best=out, centers=c, assignments=a, best_cluster=int(info[0]), passes=int(info[1]), reinit_used=int(info[2]),
                cluster_weights=cw)


def count_near(P, x, y, radius=0.4):
    P = _P32(P)
    return oracle_lib().orc_count_near(_p(P, c_fp), len(P), C.c_float(x), C.c_float(y), C.c_float(radius))


def pose_to_cell(wx, wy, angle):
    """publishPosMsg (MC:958-994): (row, column, orientation)."""
    out = np.zeros(3, np.int32)
    oracle_lib().orc_pose_to_cell(C.c_double(wx), C.c_double(wy), C.c_double(angle), _p(out, c_ip))
    return tuple(int(v) for v in out)


def exact_pose(x, y, theta):
    out = np.zeros(3, np.float32)
    oracle_lib().orc_exact_pose(C.c_double(x), C.c_double(y), C.c_double(theta), _p(out, c_fp))
    return out


def particle_poses(P):
    """publishParticles (MC:563-579): [N,4] = x, y, qz, qw."""
    P = _P32(P); out = np.zeros((len(P), 4))
    oracle_lib().orc_particle_poses(_p(P, c_fp), len(P), _p(out, c_dp))
    return out


def libc_rand_sequence(seed, count):
    """What srand(seed); rand() ... yields in this process's libc: the draws kMeansClustering makes after srand(time) (MC:808)."""
    libc = C.CDLL(None)
    libc.srand(C.c_uint(int(seed)))
    return [int(libc.rand()) for _ in range(count)]


# ------------------------------------------------------------------------------------------------------
# Compiled reference (oracle/_ref). Process-global state, exactly like the reference node.
# ------------------------------------------------------------------------------------------------------
_ref_lib = None


def ref_available():
    return os.path.exists(REF_SO)


def ref_lib():
    global _ref_lib
    if _ref_lib is None:
        L = C.CDLL(REF_SO)
        for name in ("ref_gauss_get", "ref_yaw_roundtrip", "ref_raycast", "ref_compute_weight"):
            getattr(L, name).restype = C.c_double
        _ref_lib = L
    return _ref_lib


class Ref:
    def __init__(self):
        self.L = ref_lib()
        self.L.ref_reset_state()

    def push_seeds(self, *seeds):
        for s in seeds:
            self.L.ref_push_seed(C.c_uint(int(s)))

    def clear_seeds(self):
        self.L.ref_clear_seeds()

    def seed_static_engines(self, seed_sample, seed_jitter):
        """Engine 0 = `sample` (MC:411), engine 1 = `uniformJitter` (MC:452)."""
        if self.L.ref_num_static_engines() < 2:
            self.push_seeds(1, 1)
            self.L.ref_touch_static_engines()
        assert self.L.ref_num_static_engines() == 2
        self.L.ref_seed_static_engine(0, C.c_uint(int(seed_sample)))
        self.L.ref_seed_static_engine(1, C.c_uint(int(seed_jitter)))

    def stream_mt_canonical(self, seed, n):
        out = np.zeros(n); self.L.ref_stream_mt_canonical(C.c_uint(int(seed)), n, _p(out, c_dp)); return out

    def stream_minstd_canonical(self, seed, n):
        out = np.zeros(n); self.L.ref_stream_minstd_canonical(C.c_uint(int(seed)), n, _p(out, c_dp)); return out

    def stream_minstd_normal(self, seed, n):
        out = np.zeros(n); self.L.ref_stream_minstd_normal(C.c_uint(int(seed)), n, _p(out, c_dp)); return out

    # -- SURVEY §8f rows
    def set_time(self, t):
        """The value std::time(nullptr) returns inside the reference (it seeds srand with it, MC:808)."""
        self.L.ref_set_time(C.c_long(int(t)))

    def kmeans_confidence(self, P, ratio_threshold=0.6, cluster_distance=0.5):
        P = _P32(P); out = np.zeros(3)
        self.L.ref_kmeans_confidence.restype = C.c_double
        ratio = self.L.ref_kmeans_confidence(_p(P, c_fp), len(P), C.c_double(cluster_distance), C.c_double(ratio_threshold), _p(out, c_dp))
        return ratio, out

    def kmeans(self, P, K=3, max_iters=20):
        P = _P32(P); a = np.zeros(len(P), np.int32); c = np.zeros((K, 2), np.float32)
        self.L.ref_kmeans(_p(P, c_fp), len(P), K, max_iters, _p(a, c_ip), _p(c, c_fp))
        return a, c

    def count_near(self, P, x, y, radius=0.4):
        P = _P32(P)
        return self.L.ref_count_near(_p(P, c_fp), len(P), C.c_float(x), C.c_float(y), C.c_float(radius))

    def publish_pos_msg(self, wx, wy, angle):
        out = np.zeros(3, np.int32)
        self.L.ref_publish_pos_msg(C.c_double(wx), C.c_double(wy), C.c_double(angle), _p(out, c_ip))
        return tuple(int(v) for v in out)

    def publish_exact_pose(self, x, y, theta):
        out = np.zeros(3, np.float32)
        self.L.ref_publish_exact_pose(C.c_double(x), C.c_double(y), C.c_double(theta), _p(out, c_fp))
        return out

    def publish_particles(self, P):
        P = _P32(P); out = np.zeros((len(P), 4))
        n = self.L.ref_publish_particles(_p(P, c_fp), len(P), _p(out, c_dp))
        assert n == len(P)
        return out

    def named_sample_draws(self, seed, n_rows, n_cols, count):
        uy = np.zeros(count); r = np.zeros(count, np.int32); c = np.zeros(count, np.int32); ux = np.zeros(count); uyy = np.zeros(count)
        self.L.ref_named_sample_draws(C.c_uint(int(seed)), n_rows, n_cols, count, _p(uy, c_dp), _p(r, c_ip), _p(c, c_ip), _p(ux, c_dp), _p(uyy, c_dp))
        return dict(u_yaw=uy, row=r, col=c, u_dx=ux, u_dy=uyy)

    def reset_state(self):
        self.L.ref_reset_state()

    def set_map(self, occ, res=np.float32(0.1), ox=0.0, oy=0.0):
        occ = np.ascontiguousarray(occ, dtype=np.int8)
        h, w = occ.shape
        self.L.ref_set_map(_p(occ, c_bp), w, h, C.c_float(res), C.c_double(ox), C.c_double(oy))

    def set_scan(self, scan):
        self.L.ref_set_scan(*scan.args())

    def precompute_ray_directions(self, lo=-120.0, hi=120.0, step=0.1):
        self.L.ref_precompute_ray_directions(C.c_double(lo), C.c_double(hi), C.c_double(step))

    def ray_lut(self, lo=-100000, hi=100000):
        n = self.L.ref_ray_lut_dump(lo, hi, None, None, None, 0)
        k = np.zeros(n, np.int32); dx = np.zeros(n); dy = np.zeros(n)
        self.L.ref_ray_lut_dump(lo, hi, _p(k, c_ip), _p(dx, c_dp), _p(dy, c_dp), n)
        return k, dx, dy

    def injection_state(self):
        out = np.zeros(2); self.L.ref_get_injection_state(_p(out, c_dp)); return out

    def set_injection_state(self, slow, fast):
        self.L.ref_set_injection_state(C.c_double(slow), C.c_double(fast))

    def set_motion(self, r1, t, r2):
        self.L.ref_set_motion(C.c_double(r1), C.c_double(t), C.c_double(r2))

    def gauss_get(self, d):
        return self.L.ref_gauss_get(C.c_double(d))

    def filter_scan(self, lower=-120.0, upper=120.0, cap=4096):
        r = np.zeros(cap); a = np.zeros(cap)
        n = self.L.ref_filter_scan(C.c_double(lower), C.c_double(upper), _p(r, c_dp), _p(a, c_dp), cap)
        return r[:n].copy(), a[:n].copy()

    def is_valid_pos(self, x, y):
        return bool(self.L.ref_is_valid_pos(C.c_double(x), C.c_double(y)))

    def is_occupied(self, x, y):
        return bool(self.L.ref_is_occupied(C.c_double(x), C.c_double(y)))

    def yaw_roundtrip(self, t):
        return self.L.ref_yaw_roundtrip(C.c_double(t))

    def raycast(self, x, y, theta, off_deg, max_range=1.0):
        return self.L.ref_raycast(C.c_double(x), C.c_double(y), C.c_double(theta), C.c_double(off_deg), C.c_double(max_range))

    def sample_particles(self, n):
        P = np.zeros((n, 4), np.float32); self.L.ref_sample_particles(n, _p(P, c_fp)); return P

    def diff_drive(self, enc_left, enc_right):
        out = np.zeros(3); self.L.ref_diff_drive(C.c_double(enc_left), C.c_double(enc_right), _p(out, c_dp)); return out

    def update_particle_pos(self, P):
        assert P.dtype == np.float32 and P.flags.c_contiguous
        self.L.ref_update_particle_pos(_p(P, c_fp), len(P))

    def compute_weight(self, P):
        assert P.dtype == np.float32 and P.flags.c_contiguous
        return self.L.ref_compute_weight(_p(P, c_fp), len(P))

    def resample(self, P, jitter_state):
        Pout = np.zeros_like(P)
        inj = self.L.ref_resample(_p(P, c_fp), len(P), int(bool(jitter_state)), _p(Pout, c_fp))
        return Pout, inj

    def estimate_weighted_pose(self, P):
        out = np.zeros(3)
        self.L.ref_estimate_weighted_pose(_p(np.ascontiguousarray(P, np.float32), c_fp), len(P), _p(out, c_dp))
        return out


# ------------------------------------------------------------------------------------------------------
# NS-mode oracle (oracle/mcl_oracle_ns.cpp): the engine's own north-star formulation; parity unpinned by the reference.
# ------------------------------------------------------------------------------------------------------
c_u64p = C.POINTER(C.c_uint64)
c_i64p = C.POINTER(C.c_int64)


class NsOracle:
    def __init__(self, sigma=0.1, z_hit=0.8, z_rand=0.2, max_range=5.6, laser_offset=0.1, temper=0.05, seed=0x9E3779B97F4A7C15, beam_stride=1):
        L = self.L = oracle_lib()
        L.ons_create.restype = C.c_void_p
        L.ons_weights.restype = C.c_uint64
        L.ons_u0.restype = C.c_uint32
        L.ons_det_log.restype = C.c_double
        L.ons_det_exp_q32.restype = C.c_uint64
        self.h = C.c_void_p(L.ons_create())
        L.ons_config(self.h, C.c_double(sigma), C.c_double(z_hit), C.c_double(z_rand), C.c_double(max_range), C.c_double(laser_offset),
                     C.c_double(temper), C.c_uint64(seed), beam_stride)

    def __del__(self):
        try:
            self.L.ons_destroy(self.h)
        except Exception:
            pass

    def set_map(self, occ, res=np.float32(0.1), ox=0.0, oy=0.0):
        occ = np.ascontiguousarray(occ, dtype=np.int8)
        self.H, self.W = occ.shape
        self.L.ons_set_map(self.h, _p(occ, c_bp), self.W, self.H, C.c_float(res), C.c_double(ox), C.c_double(oy))

    def field(self):
        lf = np.zeros((self.H, self.W), np.float32); d2 = np.zeros((self.H, self.W), np.uint16)
        self.L.ons_get_field(self.h, _p(lf, c_fp), d2.ctypes.data_as(C.POINTER(C.c_uint16)))
        return lf, d2

    def init(self, g0, n):
        P = np.zeros((n, 4), np.float32)
        self.L.ons_init(self.h, C.c_int64(g0), C.c_int64(n), _p(P, c_fp))
        return P

    def predict(self, P, g0, rot1, trans, rot2, step):
        assert P.dtype == np.float32 and P.flags.c_contiguous
        self.L.ons_predict(self.h, _p(P, c_fp), C.c_int64(g0), C.c_int64(len(P)), C.c_double(rot1), C.c_double(trans), C.c_double(rot2), C.c_uint32(step))

    def beams(self, scan):
        cap = len(scan.ranges)
        pts = np.zeros((max(cap, 1), 2), np.float32)
        n = self.L.ons_beams(self.h, *scan.args(), _p(pts, c_fp), cap)
        return pts[:n].copy()

    def loglik(self, P, pts):
        ll = np.zeros(len(P), np.float32)
        pts = np.ascontiguousarray(pts, np.float32)
        self.L.ons_loglik(self.h, _p(np.ascontiguousarray(P, np.float32), c_fp), C.c_int64(len(P)), _p(pts, c_fp), len(pts), _p(ll, c_fp))
        return ll

    def weights(self, ll, max_ll):
        n = len(ll)
        W = np.zeros(n, np.uint64); pre = np.zeros(n, np.uint64); wf = np.zeros(n, np.float32)
        ll = np.ascontiguousarray(ll, np.float32)
        tot = self.L.ons_weights(self.h, _p(ll, c_fp), C.c_int64(n), C.c_float(max_ll), W.ctypes.data_as(c_u64p), pre.ctypes.data_as(c_u64p), _p(wf, c_fp))
        return W, pre, wf, int(tot)

    def pose(self, P, wf):
        """NS-8: weighted-mean pose {x, y, theta} of particles P (N x 4) under the fp32 weights wf (ons_pose_partials)."""
        Q = np.ascontiguousarray(P, np.float32).copy()
        Q[:, 3] = wf
        a = np.zeros(5)
        self.L.ons_pose_partials(_p(Q, c_fp), C.c_int64(len(Q)), _p(a, c_dp))
        return np.array([a[1] / a[0], a[2] / a[0], np.arctan2(a[3], a[4])])

    def u0(self, step):
        return int(self.L.ons_u0(self.h, C.c_uint32(step)))

    def resample(self, prefix_global, u0, k_begin=0, k_end=None):
        Cg = np.ascontiguousarray(prefix_global, np.uint64)
        n = len(Cg)
        k_end = n if k_end is None else k_end
        anc = np.zeros(k_end - k_begin, np.int64)
        self.L.ons_resample(Cg.ctypes.data_as(c_u64p), C.c_int64(n), C.c_uint32(u0), C.c_int64(k_begin), C.c_int64(k_end), anc.ctypes.data_as(c_i64p))
        return anc

    def pose_partials(self, P):
        out = np.zeros(5)
        self.L.ons_pose_partials(_p(np.ascontiguousarray(P, np.float32), c_fp), C.c_int64(len(P)), _p(out, c_dp))
        return out

    def step(self, P, g0, scan, motion, step):
        """One full NS filter step on a single span holding ALL particles: returns (new particles, ancestors, ll, prefix)."""
        self.predict(P, g0, motion[0], motion[1], motion[2], step)
        pts = self.beams(scan)
        ll = self.loglik(P, pts)
        W, pre, wf, tot = self.weights(ll, ll.max())
        anc = self.resample(pre, self.u0(step))
        newP = P[anc].copy()
        newP[:, 3] = np.float32(1.0 / len(P))
        return newP, anc, ll, pre
