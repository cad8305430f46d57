"""ctypes binding of the C-ABI in include/mcl.h (libmcl_b200.so, built in-tree by csrc/Makefile).

There is no fallback: if the shared library is missing or no CUDA device is usable, loading / mcl_create raises.
"""
import ctypes as C
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libmcl_b200.so")

MODE_REF, MODE_NS = 0, 1
TRIG_LIBM, TRIG_CORRECTLY_ROUNDED = 0, 1


class MclError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("mcl error %d: %s" % (code, msg))
        self.code = code


class Config(C.Structure):
    _fields_ = [
        ("device", C.c_int32), ("mode", C.c_int32), ("max_particles", C.c_int64),
        ("sigma_hit", C.c_double), ("max_laser_range", C.c_double), ("laser_offset", C.c_double),
        ("w_hit", C.c_double), ("w_rand", C.c_double), ("fov_lower_deg", C.c_double), ("fov_upper_deg", C.c_double),
        ("beam_stride", C.c_int32), ("_pad0", C.c_int32), ("ray_step", C.c_double), ("validity_offset", C.c_double),
        ("alpha", C.c_double * 4), ("wheel_size", C.c_double), ("wheel_space", C.c_double),
        ("cell_size_px", C.c_int32), ("_pad1", C.c_int32), ("cell_meters", C.c_double), ("init_offset", C.c_double),
        ("init_shift", C.c_double),
        ("inject_max_lost", C.c_double), ("inject_alpha_slow_lost", C.c_double), ("inject_alpha_fast_lost", C.c_double),
        ("inject_max_conf", C.c_double), ("inject_alpha_slow_conf", C.c_double), ("inject_alpha_fast_conf", C.c_double),
        ("jitter_xy_lost", C.c_double), ("jitter_theta_lost", C.c_double), ("jitter_xy_conf", C.c_double),
        ("seed", C.c_uint64), ("ns_sigma_hit", C.c_double), ("ns_z_hit", C.c_double), ("ns_z_rand", C.c_double),
        ("ns_max_range", C.c_double), ("ns_beam_stride", C.c_int32), ("ns_use_fov", C.c_int32), ("ns_temper", C.c_double),
        ("kmeans_radius", C.c_double), ("trig_mode", C.c_int32), ("_pad2", C.c_int32),
    ]


class InitDraws(C.Structure):
    _fields_ = [("u_yaw", C.POINTER(C.c_double)), ("row", C.POINTER(C.c_int32)), ("col", C.POINTER(C.c_int32)),
                ("u_dx", C.POINTER(C.c_double)), ("u_dy", C.POINTER(C.c_double))]


class ResampleDraws(C.Structure):
    _fields_ = [("u_r", C.POINTER(C.c_double)), ("u_jitter", C.POINTER(C.c_double)), ("n_jitter", C.c_int64),
                ("inject", InitDraws), ("n_inject", C.c_int32)]


class ResampleStats(C.Structure):
    _fields_ = [("injected", C.c_int32), ("clamped", C.c_int32), ("p_inject", C.c_double), ("weight_slow", C.c_double),
                ("weight_fast", C.c_double), ("total_weight", C.c_double)]


class KmeansResult(C.Structure):
    _fields_ = [("ratio", C.c_double), ("x_best", C.c_double), ("y_best", C.c_double), ("theta_best", C.c_double),
                ("cluster_weight", C.c_double * 3), ("centers", C.c_float * 6), ("counts", C.c_int64 * 3),
                ("best_cluster", C.c_int32), ("passes", C.c_int32), ("reinit_used", C.c_int32), ("exact", C.c_int32)]


# every symbol include/mcl.h and include/mcl_debug.h declare: (name, restype, argtypes)
_vp, _i32, _i64, _d, _f = C.c_void_p, C.c_int32, C.c_int64, C.c_double, C.c_float
_dp, _fp, _ip, _bp = C.POINTER(C.c_double), C.POINTER(C.c_float), C.POINTER(C.c_int32), C.POINTER(C.c_int8)
SYMBOLS = [
    ("mcl_config_default", None, [C.POINTER(Config)]),
    ("mcl_create", _i32, [C.POINTER(Config), C.POINTER(_vp)]),
    ("mcl_destroy", None, [_vp]),
    ("mcl_last_error", C.c_char_p, [_vp]),
    ("mcl_version", C.c_char_p, []),
    ("mcl_rasterise_map_txt", _i32, [C.c_char_p, _bp, _i64, _ip, _ip]),
    ("mcl_set_map", _i32, [_vp, _bp, _i32, _i32, _f, _d, _d]),
    ("mcl_load_map_txt", _i32, [_vp, C.c_char_p]),
    ("mcl_precompute_ray_directions", _i32, [_vp, _d, _d, _d]),
    ("mcl_init", _i32, [_vp, _i64, C.POINTER(InitDraws)]),
    ("mcl_upload", _i32, [_vp, _fp, _i64]),
    ("mcl_download", _i32, [_vp, _fp]),
    ("mcl_num_particles", _i64, [_vp]),
    ("mcl_predict_encoders", _i32, [_vp, _d, _d, _dp, _dp]),
    ("mcl_predict_motion", _i32, [_vp, _d, _d, _d]),
    ("mcl_update", _i32, [_vp, _fp, _i32, _f, _f, _f, _f, _dp]),
    ("mcl_scan_stage", _i32, [_vp, _i32, _fp, _i32, _f, _f, _f, _f]),
    ("mcl_update_staged", _i32, [_vp, _i32, _dp]),
    ("mcl_resample", _i32, [_vp, _i32, C.POINTER(ResampleDraws), C.POINTER(ResampleStats)]),
    ("mcl_download_ancestors", _i32, [_vp, _ip]),
    ("mcl_download_cdf", _i32, [_vp, _dp]),
    ("mcl_estimate", _i32, [_vp, _dp, _dp, _dp]),
    ("mcl_step", _i32, [_vp, _d, _d, _vp, _i32, _f, _f, _f, _f, _i32, _dp, _vp]),       # ranges as an address: the hot call skips the pointer object
    ("mcl_step_staged", _i32, [_vp, _d, _d, _i32, _i32, _dp, _vp]),
    ("mcl_kmeans_confidence", _i32, [_vp, _ip, _ip, _i32, _d, C.POINTER(KmeansResult)]),
    ("mcl_download_assignments", _i32, [_vp, _ip]),
    ("mcl_pose_to_cell", _i32, [_d, _d, _d, _d, _ip, _ip, _ip]),
    ("mcl_exact_pose", _i32, [_d, _d, _d, _fp]),
    ("mcl_download_pose_array", _i32, [_vp, _i64, _i64, _i64, _dp]),
    ("mcl_config_preset", _i32, [C.POINTER(Config), C.c_char_p]),
    ("mcl_get_injection_state", _i32, [_vp, _dp, _dp]),
    ("mcl_set_injection_state", _i32, [_vp, _d, _d]),
    ("mcl_get_ray_lut", _i32, [_vp, _ip, _dp, _dp, _i32, _ip]),
    ("mcl_debug_download_resample_draws", _i32, [_vp, _dp, _dp]),
    ("mcl_debug_exact_scan", _i32, [_vp, _fp, _i64, _dp, _dp, _ip]),
    ("mcl_debug_force_sequential", _i32, [_vp, _i32]),
    ("mcl_debug_exact_scan_trace", _i32, [_vp, C.POINTER(C.c_uint64), _i64, _ip]),
    ("mcl_debug_ns_last_plan", _i32, [_vp, C.POINTER(_i64), C.POINTER(_i64), C.POINTER(_i64), C.POINTER(_i64)]),
    ("mcl_debug_trigf", _i32, [_vp, _fp, _i64, _fp, _fp, _ip]),
    ("mcl_bench_gather", _i32, [_vp, _i32, _i64, _i32, _dp]),
    ("mcl_profile_enable", _i32, [_vp, _i32]),
    ("mcl_profile_kernel_count", _i32, []),
    ("mcl_profile_kernel_name", C.c_char_p, [_i32]),
    ("mcl_profile_read", _i32, [_vp, _i32, _dp, C.POINTER(C.c_int64)]),
    ("mcl_ns_set_shard", _i32, [_vp, _i32, _i32, _i64]),
    ("mcl_ns_update_local", _i32, [_vp, _fp, _i32, _f, _f, _f, _f, _fp]),
    ("mcl_ns_update_local_staged", _i32, [_vp, _i32, _fp]),
    ("mcl_ns_weights_local", _i32, [_vp, _f, C.POINTER(C.c_uint64)]),
    ("mcl_ns_resample_local", _i32, [_vp, C.c_uint64, C.c_uint64, C.c_uint32, C.POINTER(_i64), C.POINTER(_i64)]),
    ("mcl_ns_end_step", _i32, [_vp]),
    ("mcl_ns_u0", C.c_uint32, [_vp]),
    ("mcl_ns_pose_partials", _i32, [_vp, _dp]),
    ("mcl_comm_unique_id", _i32, [_vp, _vp]),
    ("mcl_comm_init", _i32, [_vp, _vp]),
    ("mcl_ns_step", _i32, [_vp, _d, _d, _d, _fp, _i32, _f, _f, _f, _f, _dp]),
    ("mcl_ns_step_staged", _i32, [_vp, _d, _d, _d, _i32, _dp]),
    ("mcl_ns_first_slot", _i32, [C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint32, C.POINTER(_i64)]),
    ("mcl_ns_shard_range", _i32, [_i64, _i32, _i32, C.POINTER(_i64), C.POINTER(_i64), C.POINTER(_i64)]),
    ("mcl_peer_export", _i32, [_vp, _i32, _vp]),
    ("mcl_peer_import", _i32, [_vp, _i32, _i32, _vp]),
    ("mcl_peer_set", _i32, [_vp, _i32, _i32, _vp]),
    ("mcl_device_buffer", _vp, [_vp, _i32]),
    ("mcl_ns_download_field", _i32, [_vp, _fp, C.POINTER(C.c_uint16)]),
    ("mcl_ns_download_loglik", _i32, [_vp, _fp]),
    ("mcl_ns_download_prefix", _i32, [_vp, C.POINTER(C.c_uint64)]),
    ("mcl_ns_field_form", _i32, [_vp]),
    ("mcl_ns_set_exchange", _i32, [_vp, _i32]),
    ("mcl_ns_exchange_used", _i32, [_vp]),
    ("mcl_stream", _vp, [_vp]),
    ("mcl_synchronize", _i32, [_vp]),
    ("mcl_kernel_launches", _i64, [_vp]),
    ("mcl_debug_optimistic_redos", _i64, [_vp]),
    ("mcl_debug_last_scan_fell_back", _i32, [_vp, _ip]),
]

_lib = None


def build(verbose=False):
    """Compile libmcl_b200.so for sm_100a (nvcc cross-compiles without a GPU)."""
    r = subprocess.run(["make", "-j4", "-C", os.path.join(HERE, "csrc")], capture_output=not verbose, text=True)
    if r.returncode != 0:
        raise RuntimeError("building libmcl_b200.so failed:\n%s\n%s" % (r.stdout, r.stderr))


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError("%s is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(there is no CPU fallback)" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        for name, res, args in SYMBOLS:
            fn = getattr(L, name)       # AttributeError here = a symbol of include/mcl.h is not exported
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib
