"""The C-ABI library loads and exports every symbol include/mcl.h and include/mcl_debug.h declare; host-only entry points work;
without a GPU the engine refuses to start instead of falling back."""
import ctypes as C
import hashlib
import os
import re

import numpy as np
import pytest

import montecarlolocalisation_b200 as m
from montecarlolocalisation_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def gpu_present():
    """A usable CUDA device in this process (driver API; no torch import needed)."""
    try:
        cu = C.CDLL("libcuda.so.1")
        n = C.c_int(0)
        return cu.cuInit(0) == 0 and cu.cuDeviceGetCount(C.byref(n)) == 0 and n.value > 0
    except OSError:
        return False


def declared_symbols(headers=("mcl.h", "mcl_debug.h")):
    names = set()
    for hname in headers:
        text = open(os.path.join(ROOT, "include", hname)).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        names |= set(re.findall(r"\b(mcl_[a-z0-9_]+)\s*\(", text))
    return sorted(names)


def test_debug_entry_points_are_not_in_the_boundary_header():
    """Cross-check switches, the gather micro-benchmark and the per-kernel timers live in include/mcl_debug.h."""
    boundary = declared_symbols(("mcl.h",))
    assert not [n for n in boundary if n.startswith(("mcl_debug_", "mcl_bench_", "mcl_profile_"))]
    assert "mcl_debug_exact_scan" in declared_symbols(("mcl_debug.h",))


def test_every_declared_symbol_is_exported_and_bound():
    L = C.CDLL(_lib.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(L, n), n
    assert sorted(s[0] for s in _lib.SYMBOLS) == names       # the Python binding covers the whole header


def test_config_struct_matches_header_defaults():
    cfg = m.default_config()
    assert C.sizeof(cfg) % 8 == 0
    assert (cfg.sigma_hit, cfg.max_laser_range, cfg.laser_offset, cfg.w_hit, cfg.w_rand) == (0.1, 1.0, 0.1, 0.8, 0.2)
    assert (cfg.fov_lower_deg, cfg.fov_upper_deg, cfg.beam_stride) == (-120.0, 120.0, 20)
    assert list(cfg.alpha) == [0.001, 0.001, 0.0001, 0.0001]
    assert (cfg.wheel_size, cfg.wheel_space, cfg.cell_size_px, cfg.cell_meters) == (0.062, 0.265, 8, 0.8)
    assert (cfg.inject_max_lost, cfg.inject_max_conf, cfg.jitter_xy_lost, cfg.jitter_xy_conf) == (200, 50, 0.05, 0.01)
    assert cfg.ns_beam_stride == 1 and cfg.kmeans_radius == 0.4      # last fields land where the C struct puts them
    assert b"sm_100a" in _lib.load().mcl_version()


def test_host_rasteriser_matches_kat_and_oracle(map_txt):
    from oracle.pyoracle import Oracle
    occ = m.rasterise_map_txt(map_txt)
    assert hashlib.sha256(occ.tobytes()).hexdigest() == "9d700e0d21c8b669621222f4c2514d9d80a7e502fc476f77120c348c4849c275"
    from scenario import load_map
    assert np.array_equal(load_map(), occ) and np.array_equal(Oracle.rasterise_map_txt(map_txt), occ)      # the committed fixture
    for txt in ("[[[T,L],[T,R]],[[L,B]]]", "[[[T,L,B,R]]]", "[[[],[B]],[[R],[L,T]],[[B],[B,R]]]"):
        assert np.array_equal(m.rasterise_map_txt(txt), Oracle.rasterise_map_txt(txt))
    with pytest.raises(m.MclError):
        m.rasterise_map_txt("[[[Q]]]")


@pytest.mark.skipif(gpu_present(), reason="GPU present")
def test_no_gpu_means_loud_failure_not_fallback():
    with pytest.raises(m.MclError) as e:
        m.ParticleFilter()
    assert e.value.code == -2 and "no CPU fallback" in str(e.value)


def _cxx_host_class(tmp_path):
    import subprocess
    exe = str(tmp_path / "cxx_binding_check")
    libdir = os.path.dirname(_lib.LIB_PATH)
    subprocess.run(["g++", "-std=c++17", "-o", exe, os.path.join(ROOT, "tests", "native", "cxx_binding_check.cpp"),
                    "-L" + libdir, "-lmcl_b200", "-Wl,-rpath," + libdir], check=True)
    return subprocess.run([exe], capture_output=True, text=True)


@pytest.mark.skipif(gpu_present(), reason="GPU present: test_cxx_host_class_runs_on_the_gpu covers it")
def test_cxx_host_class_compiles_and_links(tmp_path):
    """include/mcl_particle_filter.hpp (the class the ROS node would use) builds against the library; without a GPU its
    constructor throws instead of silently computing on the CPU."""
    r = _cxx_host_class(tmp_path)
    assert r.returncode == 3 and "no CPU fallback" in r.stdout, r.stdout + r.stderr


@pytest.mark.gpu
def test_cxx_host_class_runs_on_the_gpu(tmp_path):
    """The same program on a GPU box: map, particles, a tick through the per-function members and one through
    executeParticleFilter, the k-means confidence row and the pose-array download, all through the C++ class."""
    r = _cxx_host_class(tmp_path)
    assert r.returncode == 0 and "gpu ok" in r.stdout, r.stdout + r.stderr


def test_host_only_pose_adapters_match_oracle():
    """mcl_pose_to_cell / mcl_exact_pose (publishPosMsg / publishExactPose, MC:958-1008) need no GPU."""
    from oracle import pyoracle
    rng = np.random.default_rng(3)
    pts = [(-0.1, 1.0, 0.3), (1.0, -1e-9, 0.0), (-1, -1, -1), (0.0, 0.0, 0.0), (0.4, 0.4, 0.0), (0.79999, 0.8, np.pi / 4), (1.2, 3.6, -7.0),
           (1.2, 3.6, 100.0), (1.2, 3.6, np.deg2rad(135.0)), (1.2, 3.6, np.deg2rad(225.0)), (1.2, 3.6, np.deg2rad(315.0)), (1.2, 3.6, np.deg2rad(45.0))]
    pts += [(rng.uniform(-0.2, 4.8), rng.uniform(-0.2, 4.8), rng.uniform(-10, 10)) for _ in range(3000)]
    for wx, wy, th in pts:
        assert m.pose_to_cell(wx, wy, th) == pyoracle.pose_to_cell(wx, wy, th)
        assert np.array_equal(m.exact_pose(wx, wy, th), pyoracle.exact_pose(wx, wy, th))


def test_config_presets():
    cfg = m.default_config()
    L = _lib.load()
    assert L.mcl_config_preset(C.byref(cfg), b"playground") == 0
    assert (cfg.ray_step, cfg.beam_stride, cfg.fov_lower_deg, cfg.fov_upper_deg) == (0.05, 3, -90.0, 90.0)
    assert L.mcl_config_preset(C.byref(cfg), b"reference") == 0
    assert (cfg.ray_step, cfg.beam_stride, cfg.kmeans_radius) == (0.1, 20, 0.4)
    assert L.mcl_config_preset(C.byref(cfg), b"nope") != 0
