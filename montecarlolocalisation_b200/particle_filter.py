"""Host-side mirror of the reference's particle-filter functions over the C-ABI (include/mcl.h).

The reference (pink_fundamentals/src/monte_carlo.cpp, "MC") has no class: the filter is the free functions
sampleParticles / diffDriveModel + updateParticlePos / computeWeight / resampleParticles / estimateWeightedPose over
globals. ParticleFilter keeps those names and argument meanings so tests read like calls into the reference.
The C++ twin for the ROS node is include/mcl_particle_filter.hpp.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import MODE_NS, MODE_REF, Config, InitDraws, MclError, ResampleDraws, ResampleStats

_dp, _fp, _ip, _bp = C.POINTER(C.c_double), C.POINTER(C.c_float), C.POINTER(C.c_int32), C.POINTER(C.c_int8)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def default_config(**overrides):
    cfg = Config()
    _lib.load().mcl_config_default(C.byref(cfg))
    for k, v in overrides.items():
        if k == "alpha":
            for i in range(4):
                cfg.alpha[i] = v[i]
        else:
            if not hasattr(cfg, k):
                raise AttributeError("mcl_config has no field %r" % k)
            setattr(cfg, k, v)
    return cfg


def rasterise_map_txt(text):
    """map.txt text -> int8 occupancy grid (publish_map.py + publish_map_rviz.cpp:306-437). Host only."""
    L = _lib.load()
    w, h = C.c_int32(), C.c_int32()
    rc = L.mcl_rasterise_map_txt(text.encode(), None, 0, C.byref(w), C.byref(h))
    if rc:
        raise MclError(rc, L.mcl_last_error(None).decode())
    out = np.zeros((h.value, w.value), np.int8)
    rc = L.mcl_rasterise_map_txt(text.encode(), out.ctypes.data_as(_bp), out.size, C.byref(w), C.byref(h))
    if rc:
        raise MclError(rc, L.mcl_last_error(None).decode())
    return out


class ParticleFilter:
    def __init__(self, cfg=None, prefill_ray_directions=True, **overrides):
        self.L = _lib.load()
        self.cfg = cfg if cfg is not None else default_config(**overrides)
        h = C.c_void_p()
        rc = self.L.mcl_create(C.byref(self.cfg), C.byref(h))
        if rc:
            raise MclError(rc, self.L.mcl_last_error(None).decode())
        self.h = h
        self._step_pose = (C.c_double * 3)()
        self._step_stats = ResampleStats()
        self._step_stats_ref = C.byref(self._step_stats)
        if prefill_ray_directions and self.cfg.mode == MODE_REF:
            self.precomputeRayDirections(-120.0, 120.0, 0.1)       # MC:1199

    def close(self):
        if getattr(self, "h", None):
            self.L.mcl_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc:
            raise MclError(rc, self.L.mcl_last_error(self.h).decode())

    # -- map ------------------------------------------------------------------------------------------
    def setMap(self, occ, resolution=np.float32(0.1), origin_x=0.0, origin_y=0.0):
        occ = np.ascontiguousarray(occ, dtype=np.int8)
        h, w = occ.shape
        self._ck(self.L.mcl_set_map(self.h, occ.ctypes.data_as(_bp), w, h, C.c_float(resolution), origin_x, origin_y))

    def loadMapTxt(self, path):
        self._ck(self.L.mcl_load_map_txt(self.h, path.encode()))

    def precomputeRayDirections(self, lo, hi, step):
        self._ck(self.L.mcl_precompute_ray_directions(self.h, lo, hi, step))

    # -- particles ----------------------------------------------------------------------------------------
    def sampleParticles(self, n, draws=None):
        """draws: dict(u_yaw,row,col,u_dx,u_dy) named draws, or None for the engine's Philox stream."""
        if draws is None:
            self._ck(self.L.mcl_init(self.h, n, None))
            return
        keep = [_f64(draws["u_yaw"]), _i32(draws["row"]), _i32(draws["col"]), _f64(draws["u_dx"]), _f64(draws["u_dy"])]
        d = InitDraws(keep[0].ctypes.data_as(_dp), keep[1].ctypes.data_as(_ip), keep[2].ctypes.data_as(_ip),
                      keep[3].ctypes.data_as(_dp), keep[4].ctypes.data_as(_dp))
        self._ck(self.L.mcl_init(self.h, n, C.byref(d)))

    def uploadParticles(self, P):
        P = np.ascontiguousarray(P, dtype=np.float32)
        assert P.ndim == 2 and P.shape[1] == 4
        self._ck(self.L.mcl_upload(self.h, P.ctypes.data_as(_fp), len(P)))

    def downloadParticles(self):
        n = self.L.mcl_num_particles(self.h)
        P = np.empty((n, 4), np.float32)
        self._ck(self.L.mcl_download(self.h, P.ctypes.data_as(_fp)))
        return P

    @property
    def num_particles(self):
        return self.L.mcl_num_particles(self.h)

    # -- predict --------------------------------------------------------------------------------------------
    def diffDriveModel(self, encoder_left, encoder_right, z3=None):
        """diffDriveModel + updateParticlePos (MC:1084-1086). Returns the noised (rot_1, trans, rot_2)."""
        out = np.zeros(3)
        z = _f64(z3) if z3 is not None else None
        self._ck(self.L.mcl_predict_encoders(self.h, encoder_left, encoder_right,
                                             z.ctypes.data_as(_dp) if z is not None else None, out.ctypes.data_as(_dp)))
        return out

    def updateParticlePos(self, rot_1, trans, rot_2):
        self._ck(self.L.mcl_predict_motion(self.h, rot_1, trans, rot_2))

    # -- update ---------------------------------------------------------------------------------------------
    def computeWeight(self, ranges, angle_min, angle_inc, range_min, range_max):
        r = np.ascontiguousarray(ranges, dtype=np.float32)
        total = C.c_double()
        self._ck(self.L.mcl_update(self.h, r.ctypes.data_as(_fp), len(r), C.c_float(angle_min), C.c_float(angle_inc),
                                   C.c_float(range_min), C.c_float(range_max), C.byref(total)))
        return total.value

    def stageScan(self, slot, ranges, angle_min, angle_inc, range_min, range_max):
        r = np.ascontiguousarray(ranges, dtype=np.float32)
        self._ck(self.L.mcl_scan_stage(self.h, slot, r.ctypes.data_as(_fp), len(r), C.c_float(angle_min), C.c_float(angle_inc),
                                       C.c_float(range_min), C.c_float(range_max)))

    def computeWeightStaged(self, slot):
        total = C.c_double()
        self._ck(self.L.mcl_update_staged(self.h, slot, C.byref(total)))
        return total.value

    # -- resample ---------------------------------------------------------------------------------------------
    def resampleParticles(self, jitter_state, u_r=None, u_jitter=None, inject=None):
        """The part of resampleParticles after computeWeight (MC:469-561). Returns the stats dict."""
        st = ResampleStats()
        if u_r is None:
            self._ck(self.L.mcl_resample(self.h, int(bool(jitter_state)), None, C.byref(st)))
        else:
            keep = [_f64(u_r), _f64(u_jitter)]
            d = ResampleDraws()
            d.u_r = keep[0].ctypes.data_as(_dp)
            d.u_jitter = keep[1].ctypes.data_as(_dp)
            d.n_jitter = len(keep[1])
            d.n_inject = 0
            if inject is not None:
                ik = [_f64(inject["u_yaw"]), _i32(inject["row"]), _i32(inject["col"]), _f64(inject["u_dx"]), _f64(inject["u_dy"])]
                keep += ik
                d.inject = InitDraws(ik[0].ctypes.data_as(_dp), ik[1].ctypes.data_as(_ip), ik[2].ctypes.data_as(_ip),
                                     ik[3].ctypes.data_as(_dp), ik[4].ctypes.data_as(_dp))
                d.n_inject = len(ik[0])
            self._ck(self.L.mcl_resample(self.h, int(bool(jitter_state)), C.byref(d), C.byref(st)))
        return dict(injected=st.injected, clamped=st.clamped, p_inject=st.p_inject, weight_slow=st.weight_slow,
                    weight_fast=st.weight_fast, total_weight=st.total_weight)

    def ancestors(self):
        idx = np.empty(self.num_particles, np.int32)
        self._ck(self.L.mcl_download_ancestors(self.h, idx.ctypes.data_as(_ip)))
        return idx

    def cdf(self):
        out = np.empty(self.num_particles, np.float64)
        self._ck(self.L.mcl_download_cdf(self.h, out.ctypes.data_as(_dp)))
        return out

    # -- estimate ---------------------------------------------------------------------------------------------
    def estimateWeightedPose(self):
        x, y, t = C.c_double(), C.c_double(), C.c_double()
        self._ck(self.L.mcl_estimate(self.h, C.byref(x), C.byref(y), C.byref(t)))
        return np.array([x.value, y.value, t.value])

    # -- one tick (executeParticleFilter, MC:1084-1092) as a single engine call ----------------------------------
    def executeParticleFilter(self, enc_left, enc_right, jitter_state, scan=None, slot=None, want_result=True):
        """predict + computeWeight + resample + estimate enqueued as one piece (mcl_step / mcl_step_staged): one wait for
        the GPU instead of three, same results. Returns (pose, stats); with want_result=False nothing is read back and the
        call returns as soon as the tick is queued (None)."""
        # (the tick is ~0.25 ms at 1M particles and ~0.08 ms at 1500: the wrapper keeps its own share small - reused output
        # blocks, the scan handed over by address, no per-call ctypes objects)
        if want_result:
            pose, st, stp = self._step_pose, self._step_stats, self._step_stats_ref
        else:
            pose = st = stp = None
        if slot is not None:
            rc = self.L.mcl_step_staged(self.h, enc_left, enc_right, slot, 1 if jitter_state else 0, pose, stp)
        else:
            r = scan["ranges"]
            if not (type(r) is np.ndarray and r.dtype == np.float32 and r.flags.c_contiguous):
                r = np.ascontiguousarray(r, dtype=np.float32)
            rc = self.L.mcl_step(self.h, enc_left, enc_right, r.ctypes.data, len(r), scan["angle_min"], scan["angle_inc"],
                                 scan["range_min"], scan["range_max"], 1 if jitter_state else 0, pose, stp)
        if rc:
            self._ck(rc)
        if not want_result:
            return None
        return np.array(pose), dict(injected=st.injected, clamped=st.clamped, p_inject=st.p_inject, weight_slow=st.weight_slow,
                                    weight_fast=st.weight_fast, total_weight=st.total_weight)

    # -- state ------------------------------------------------------------------------------------------------
    def injectionState(self):
        s, f = C.c_double(), C.c_double()
        self._ck(self.L.mcl_get_injection_state(self.h, C.byref(s), C.byref(f)))
        return np.array([s.value, f.value])

    def setInjectionState(self, slow, fast):
        self._ck(self.L.mcl_set_injection_state(self.h, slow, fast))

    def rayLut(self):
        cnt = C.c_int32()
        self._ck(self.L.mcl_get_ray_lut(self.h, None, None, None, 0, C.byref(cnt)))
        k = np.zeros(cnt.value, np.int32); dx = np.zeros(cnt.value); dy = np.zeros(cnt.value)
        self._ck(self.L.mcl_get_ray_lut(self.h, k.ctypes.data_as(_ip), dx.ctypes.data_as(_dp), dy.ctypes.data_as(_dp), cnt.value, C.byref(cnt)))
        return k, dx, dy

    # -- rows either side of the hot path ------------------------------------------------------
    def isLocalizationLost_densitiy_cluster(self, cluster_ratio_threshold, init_idx=None, reinit_idx=None):
        """MC:886-949 (sic). Returns dict(ratio, best=(x,y,theta) with the -1 sentinel, centers, counts, ...). init_idx /
        reinit_idx: the rand() % N draws of kMeansClustering (MC:813, 860) for reproducible runs; None = Philox."""
        res = _lib.KmeansResult()
        ii = np.ascontiguousarray(init_idx, dtype=np.int32) if init_idx is not None else None
        ri = np.ascontiguousarray(reinit_idx, dtype=np.int32) if reinit_idx is not None and len(reinit_idx) else None
        self._ck(self.L.mcl_kmeans_confidence(self.h, ii.ctypes.data_as(_ip) if ii is not None else None,
                                              ri.ctypes.data_as(_ip) if ri is not None else None, len(ri) if ri is not None else 0,
                                              C.c_double(cluster_ratio_threshold), C.byref(res)))
        return dict(ratio=res.ratio, best=np.array([res.x_best, res.y_best, res.theta_best]), centers=np.array(res.centers, np.float32).reshape(3, 2),
                    counts=np.array(res.counts), cluster_weights=np.array(res.cluster_weight), best_cluster=res.best_cluster, passes=res.passes,
                    reinit_used=res.reinit_used, exact=bool(res.exact))

    def clusterAssignments(self):
        a = np.zeros(self.num_particles, np.int32)
        self._ck(self.L.mcl_download_assignments(self.h, a.ctypes.data_as(_ip)))
        return a

    def poseArray(self, first=0, stride=1, count=None):
        """publishParticles (MC:563-579): [count, 4] = position.x, position.y, orientation.z, orientation.w."""
        if count is None:
            count = (self.num_particles - first + stride - 1) // stride
        out = np.zeros((count, 4))
        self._ck(self.L.mcl_download_pose_array(self.h, first, stride, count, out.ctypes.data_as(_dp)))
        return out

    # -- instrumentation ----------------------------------------------------------------------------------------
    def lastResampleDraws(self, jitter_state):
        n = self.num_particles
        u_r = np.empty(n); u_j = np.empty(n * (3 if jitter_state else 2))
        self._ck(self.L.mcl_debug_download_resample_draws(self.h, u_r.ctypes.data_as(_dp), u_j.ctypes.data_as(_dp)))
        return u_r, u_j

    def exactScan(self, w):
        """(cdf, total, fell_back) of the engine's bit-exact sequential f64 accumulation over fp32 w."""
        w = np.ascontiguousarray(w, dtype=np.float32)
        cdf = np.empty(len(w)); total = C.c_double(); fb = C.c_int32()
        self._ck(self.L.mcl_debug_exact_scan(self.h, w.ctypes.data_as(_fp), len(w), cdf.ctypes.data_as(_dp), C.byref(total), C.byref(fb)))
        return cdf, total.value, fb.value

    def exactScanTrace(self, w):
        """Stage time stamps [tiles, 16] (ns) of the one-kernel exact scan over fp32 w (mcl_debug_exact_scan_trace)."""
        self._ck(self.L.mcl_debug_exact_scan_trace(self.h, None, 1, None))
        self.exactScan(w)
        nt = C.c_int32()
        tiles = (len(w) + 4095) // 4096          # XSF_TILE
        out = np.zeros((tiles, 16), np.uint64)
        self._ck(self.L.mcl_debug_exact_scan_trace(self.h, out.ctypes.data_as(C.POINTER(C.c_uint64)), tiles, C.byref(nt)))
        self._ck(self.L.mcl_debug_exact_scan_trace(self.h, None, 0, None))
        return out

    def trigf(self, x):
        """(sin, cos, kind) of the float trig the REF kernels evaluate, on the device (mcl_debug_trigf)."""
        x = np.ascontiguousarray(x, dtype=np.float32)
        s = np.empty_like(x); c = np.empty_like(x); kind = C.c_int32()
        self._ck(self.L.mcl_debug_trigf(self.h, x.ctypes.data_as(_fp), x.size, s.ctypes.data_as(_fp), c.ctypes.data_as(_fp), C.byref(kind)))
        return s, c, kind.value

    def forceSequential(self, on):
        """bit 0: single-chain accumulation kernels; bit 1: per-particle computeWeight kernel (cross-checks)."""
        self._ck(self.L.mcl_debug_force_sequential(self.h, int(on)))

    def benchGather(self, tier, table_bytes, iters=256):
        """Random 4-byte gathers per second from a table in shared memory (tier 0) or global memory (tier 1)."""
        out = C.c_double()
        self._ck(self.L.mcl_bench_gather(self.h, tier, table_bytes, iters, C.byref(out)))
        return out.value

    def profileEnable(self, on):
        self._ck(self.L.mcl_profile_enable(self.h, int(bool(on))))

    def profileRead(self):
        """{kernel name: (total ms, launches)} accumulated since profileEnable(True)."""
        out = {}
        for i in range(self.L.mcl_profile_kernel_count()):
            ms, cnt = C.c_double(), C.c_int64()
            self._ck(self.L.mcl_profile_read(self.h, i, C.byref(ms), C.byref(cnt)))
            if cnt.value:
                out[self.L.mcl_profile_kernel_name(i).decode()] = (ms.value, cnt.value)
        return out

    def stream(self):
        return self.L.mcl_stream(self.h)

    def synchronize(self):
        self._ck(self.L.mcl_synchronize(self.h))

    def kernelLaunches(self):
        return self.L.mcl_kernel_launches(self.h)

    def lastScanFellBack(self):
        """True if the last tick's exact accumulations took the single-chain fallback (mcl_debug_last_scan_fell_back)."""
        f = C.c_int32()
        self._ck(self.L.mcl_debug_last_scan_fell_back(self.h, C.byref(f)))
        return bool(f.value)

    def optimisticRedos(self):
        """Whole-tick calls that ran twice because their first-touch pre-pass found new ray directions (mcl_debug_optimistic_redos)."""
        return self.L.mcl_debug_optimistic_redos(self.h)


class NsShard:
    """One shard (= one GPU handle) of an MCL_MODE_NS filter, phase by phase (include/mcl.h "NS mode across GPUs")."""

    def __init__(self, rank=0, world=1, n_global=None, device=0, **overrides):
        self.pf = ParticleFilter(prefill_ray_directions=False, mode=MODE_NS, device=device, **overrides)
        self.L, self.h = self.pf.L, self.pf.h
        self.rank, self.world = rank, world
        if n_global is not None:
            self.pf._ck(self.L.mcl_ns_set_shard(self.h, rank, world, n_global))
            self.n_global = n_global

    def update_local(self, ranges, angle_min, angle_inc, range_min, range_max):
        r = np.ascontiguousarray(ranges, dtype=np.float32)
        mx = C.c_float()
        self.pf._ck(self.L.mcl_ns_update_local(self.h, r.ctypes.data_as(_fp), len(r), C.c_float(angle_min), C.c_float(angle_inc),
                                               C.c_float(range_min), C.c_float(range_max), C.byref(mx)))
        return mx.value

    def update_local_staged(self, slot):
        mx = C.c_float()
        self.pf._ck(self.L.mcl_ns_update_local_staged(self.h, slot, C.byref(mx)))
        return mx.value

    def weights_local(self, global_max):
        t = C.c_uint64()
        self.pf._ck(self.L.mcl_ns_weights_local(self.h, C.c_float(global_max), C.byref(t)))
        return t.value

    def resample_local(self, offset, total, u0):
        lo, hi = C.c_int64(), C.c_int64()
        self.pf._ck(self.L.mcl_ns_resample_local(self.h, C.c_uint64(offset), C.c_uint64(total), C.c_uint32(u0), C.byref(lo), C.byref(hi)))
        return lo.value, hi.value

    def end_step(self):
        self.pf._ck(self.L.mcl_ns_end_step(self.h))

    def u0(self):
        return self.L.mcl_ns_u0(self.h)

    def pose_partials(self):
        out = np.zeros(5)
        self.pf._ck(self.L.mcl_ns_pose_partials(self.h, out.ctypes.data_as(_dp)))
        return out

    def field(self, shape):
        lf = np.zeros(shape, np.float32); d2 = np.zeros(shape, np.uint16)
        self.pf._ck(self.L.mcl_ns_download_field(self.h, lf.ctypes.data_as(_fp), d2.ctypes.data_as(C.POINTER(C.c_uint16))))
        return lf, d2

    def loglik(self):
        ll = np.zeros(self.pf.num_particles, np.float32)
        self.pf._ck(self.L.mcl_ns_download_loglik(self.h, ll.ctypes.data_as(_fp)))
        return ll

    def set_exchange(self, mode):
        """Collectives of the sharded step: 'peer' (mailboxes in peer memory), 'nccl', or None for the default."""
        self.pf._ck(self.L.mcl_ns_set_exchange(self.h, {None: -1, "nccl": 0, "peer": 1}[mode]))

    def exchange_used(self):
        return {0: "nccl", 1: "peer"}.get(self.L.mcl_ns_exchange_used(self.h), "none")

    def field_form(self):
        """Where the last sensor-model launch read the field: 'smem-f32', 'global-f32' or 'global-u8' (mcl_ns_field_form)."""
        return {0: "smem-f32", 1: "global-f32", 2: "global-u8"}.get(self.L.mcl_ns_field_form(self.h), "none")

    def prefix(self):
        p = np.zeros(self.pf.num_particles, np.uint64)
        self.pf._ck(self.L.mcl_ns_download_prefix(self.h, p.ctypes.data_as(C.POINTER(C.c_uint64))))
        return p

    def comm_unique_id(self):
        buf = C.create_string_buffer(128)
        self.pf._ck(self.L.mcl_comm_unique_id(self.h, buf))
        return buf.raw

    def comm_init(self, raw128):
        self.pf._ck(self.L.mcl_comm_init(self.h, C.create_string_buffer(raw128, 128)))

    def step(self, motion, scan=None, slot=None, want_pose=False):
        """One whole filter step inside the engine (NCCL + device-side plan, no host round trip)."""
        pose = (C.c_double * 3)() if want_pose else None
        if slot is not None:
            self.pf._ck(self.L.mcl_ns_step_staged(self.h, motion[0], motion[1], motion[2], slot, pose))
        else:
            r = np.ascontiguousarray(scan["ranges"], dtype=np.float32)
            self.pf._ck(self.L.mcl_ns_step(self.h, motion[0], motion[1], motion[2], r.ctypes.data_as(_fp), len(r), C.c_float(scan["angle_min"]),
                                           C.c_float(scan["angle_inc"]), C.c_float(scan["range_min"]), C.c_float(scan["range_max"]), pose))
        return np.array(pose) if want_pose else None

    def last_plan(self):
        """(k_lo, k_hi, own_begin, own_count) of the last step: the output slots this shard resolved and the ones it holds."""
        v = [C.c_int64() for _ in range(4)]
        self.pf._ck(self.L.mcl_debug_ns_last_plan(self.h, *[C.byref(x) for x in v]))
        return tuple(x.value for x in v)

    def device_buffer(self, which):
        return self.L.mcl_device_buffer(self.h, which)

    def peer_set(self, rank, which, ptr):
        self.pf._ck(self.L.mcl_peer_set(self.h, rank, which, C.c_void_p(ptr)))

    def peer_export(self, which):
        buf = C.create_string_buffer(64)
        self.pf._ck(self.L.mcl_peer_export(self.h, which, buf))
        return buf.raw

    def peer_import(self, rank, which, raw64):
        self.pf._ck(self.L.mcl_peer_import(self.h, rank, which, C.create_string_buffer(raw64, 64)))


def pose_to_cell(wx, wy, angle, cell_meters=0.8):
    """publishPosMsg (MC:958-994): (row, column, orientation); RIGHT=0 UP=1 LEFT=2 DOWN=3, all -1 = not localised."""
    r, c, o = C.c_int32(), C.c_int32(), C.c_int32()
    rc = _lib.load().mcl_pose_to_cell(wx, wy, angle, cell_meters, C.byref(r), C.byref(c), C.byref(o))
    if rc:
        raise MclError(rc, "mcl_pose_to_cell")
    return r.value, c.value, o.value


def exact_pose(x, y, theta):
    """publishExactPose (MC:995-1008): the float32 message fields."""
    out = np.zeros(3, np.float32)
    rc = _lib.load().mcl_exact_pose(x, y, theta, out.ctypes.data_as(_fp))
    if rc:
        raise MclError(rc, "mcl_exact_pose")
    return out


def ns_first_slot(offset, total, n_global, u0):
    out = C.c_int64()
    rc = _lib.load().mcl_ns_first_slot(C.c_uint64(offset), C.c_uint64(total), C.c_uint64(n_global), C.c_uint32(u0), C.byref(out))
    if rc:
        raise MclError(rc, "mcl_ns_first_slot")
    return out.value


def ns_shard_range(n_global, world, rank):
    b, c, p = C.c_int64(), C.c_int64(), C.c_int64()
    rc = _lib.load().mcl_ns_shard_range(n_global, world, rank, C.byref(b), C.byref(c), C.byref(p))
    if rc:
        raise MclError(rc, "mcl_ns_shard_range")
    return b.value, c.value, p.value


def ns_step_in_process(shards, scan, motion):
    """Drive G shards living in ONE process through a filter step; the three collectives are plain Python here.
    (bench.py and multi-process users do the same with torch.distributed over NCCL.)"""
    for s in shards:
        s.pf.updateParticlePos(*motion)
    maxes = [s.update_local(scan["ranges"], scan["angle_min"], scan["angle_inc"], scan["range_min"], scan["range_max"]) for s in shards]
    gmax = max(maxes)                                        # all-reduce(max)
    totals = [s.weights_local(gmax) for s in shards]         # all-gather of the local totals
    total = sum(totals)
    u0 = shards[0].u0()
    off = 0
    ranges = []
    for s, t in zip(shards, totals):
        ranges.append(s.resample_local(off, total, u0))
        off += t
    for s in shards:                                         # barrier, then everyone swaps
        s.end_step()
    return dict(max=gmax, totals=totals, total=total, u0=u0, slots=ranges)
