"""CPU checks of the NS-mode definitions: the oracle's deterministic math against libm, its distance transform against
scipy's independent EDT, and the product's host-side resampling plan (mcl_ns_first_slot / mcl_ns_shard_range) against
the oracle's division-based systematic resampler, including a simulated multi-shard step."""
import ctypes as C

import numpy as np
import pytest
from scipy import ndimage

import montecarlolocalisation_b200 as m
from montecarlolocalisation_b200 import synth
from oracle.pyoracle import NsOracle, Scan, oracle_lib
from scenario import RES, Scenario


def test_deterministic_math_is_accurate():
    L = oracle_lib()
    L.ons_det_log.restype = C.c_double
    L.ons_det_exp_q32.restype = C.c_uint64
    rng = np.random.default_rng(0)
    s, c = C.c_double(), C.c_double()
    for t in rng.uniform(-2000, 2000, 20000):
        L.ons_det_sincos(C.c_double(t), C.byref(s), C.byref(c))
        assert abs(s.value - np.sin(t)) < 3e-16 and abs(c.value - np.cos(t)) < 3e-16
    for u in rng.uniform(1e-10, 1.0, 20000):
        assert abs(L.ons_det_log(C.c_double(u)) - np.log(u)) <= 4e-16 * max(1.0, abs(np.log(u)))
    for t in rng.uniform(-22.4, 0.0, 20000).astype(np.float32):
        want = np.exp(np.float64(t)) * 2.0**32
        assert abs(int(L.ons_det_exp_q32(C.c_float(t))) - want) <= 2.5e-7 * want + 1
    assert L.ons_det_exp_q32(C.c_float(0.0)) == 2**32 and L.ons_det_exp_q32(C.c_float(-23.0)) == 0
    assert L.ons_det_exp_q32(C.c_float(np.nan)) == 0


def test_fp32_noise_functions_are_accurate_and_normal():
    """NS-7's fp32 log / sincos (IEEE-only arithmetic) against libm, and the Box-Muller output against N(0,1) moments."""
    L = oracle_lib()
    L.ons_det_log_f32.restype = C.c_float
    rng = np.random.default_rng(1)
    s, c = C.c_float(), C.c_float()
    for t in rng.uniform(-8, 8, 20000).astype(np.float32):
        L.ons_det_sincos_f32(C.c_float(t), C.byref(s), C.byref(c))
        assert abs(s.value - np.sin(np.float64(t))) < 2.5e-7 and abs(c.value - np.cos(np.float64(t))) < 2.5e-7
    for u in rng.random(20000).astype(np.float32):
        L.ons_det_sincos_turns_f32(C.c_float(u), C.byref(s), C.byref(c))
        assert abs(s.value - np.sin(2 * np.pi * np.float64(u))) < 2.5e-7 and abs(c.value - np.cos(2 * np.pi * np.float64(u))) < 2.5e-7
    us = np.concatenate([rng.uniform(2.0**-24, 1.0, 20000), [2.0**-24, 1.0 - 2.0**-24, 0.5, 0.70710678, 0.70710679]]).astype(np.float32)
    for u in us:
        got = L.ons_det_log_f32(C.c_float(u))
        assert abs(got - np.log(np.float64(u))) <= 3e-7 * max(1.0, abs(np.log(np.float64(u)))), (u, got)
    n = 400000
    w1 = rng.integers(0, 2**32, n, dtype=np.uint64).astype(np.uint32); w2 = rng.integers(0, 2**32, n, dtype=np.uint64).astype(np.uint32)
    z0 = np.zeros(n, np.float32); z1 = np.zeros(n, np.float32)
    u32p, fp = C.POINTER(C.c_uint32), C.POINTER(C.c_float)
    L.ons_normal_pairs(w1.ctypes.data_as(u32p), w2.ctypes.data_as(u32p), C.c_int64(n), z0.ctypes.data_as(fp), z1.ctypes.data_as(fp))
    z = np.concatenate([z0, z1]).astype(np.float64)
    assert np.isfinite(z).all() and abs(z.mean()) < 0.005 and abs(z.std() - 1.0) < 0.005
    assert abs((z ** 4).mean() - 3.0) < 0.05 and abs(np.corrcoef(z0, z1)[0, 1]) < 0.005 and np.abs(z).max() < 5.8
    # against the textbook formula evaluated in f64
    u1 = ((w1 >> 9).astype(np.float64) + 0.5) * 2.0**-23; u2 = ((w2 >> 9).astype(np.float64) + 0.5) * 2.0**-23
    ref0 = np.sqrt(-2 * np.log(u1)) * np.cos(2 * np.pi * u2)
    assert np.abs(z0 - ref0).max() < 3e-6


def test_philox_known_answer():
    """Philox4x32-10 known-answer vectors from the Random123 distribution (kat_vectors)."""
    L = oracle_lib()
    out = (C.c_uint32 * 4)()
    L.ons_philox(0, 0, 0, 0, 0, 0, out)
    assert [hex(v) for v in out] == ["0x6627e8d5", "0xe169c58d", "0xbc57ac4c", "0x9b00dbd8"]
    L.ons_philox(0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff, out)
    assert [hex(v) for v in out] == ["0x408f276d", "0x41c83b0e", "0xa20bc7c6", "0x6d5451fd"]
    L.ons_philox(0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344, 0xa4093822, 0x299f31d0, out)
    assert [hex(v) for v in out] == ["0xd16cfe09", "0x94fdcceb", "0x5001e420", "0x24126ea1"]


@pytest.mark.parametrize("cells,seed", [(6, 0), (16, 3), (32, 4)])
def test_distance_transform_matches_scipy(cells, seed):
    occ = Scenario(1).occ if cells == 6 else synth.maze_occupancy(cells, seed)
    o = NsOracle()
    o.set_map(occ, RES)
    lf, d2 = o.field()
    edt = ndimage.distance_transform_edt(~(occ > 50))
    want = np.minimum(np.rint(edt ** 2).astype(np.int64), 400)        # capped at R^2, R = ceil(2 m / 0.1 m) = 20
    assert np.array_equal(d2.astype(np.int64), want)
    # the field is the log of a Gaussian-plus-uniform mixture of the metric distance
    d = float(RES) * np.sqrt(d2.astype(np.float64))
    p = 0.8 * np.exp(-d * d / 0.02) / (0.1 * np.sqrt(2 * np.pi)) + 0.2 / 5.6
    assert np.allclose(lf, np.log(p), rtol=1e-6, atol=1e-6)


def oracle_world(n, steps=3):
    sc = Scenario(steps)
    o = NsOracle()
    o.set_map(sc.occ, RES)
    return sc, o


def test_host_plan_matches_division_based_resampler():
    sc, o = oracle_world(0)
    rng = np.random.default_rng(5)
    for n in (1, 2, 17, 1000, 4097):
        for trial in range(6):
            W = rng.integers(0, 2**32 + 1, n, dtype=np.uint64)
            W[rng.random(n) < 0.6] = 0
            W[rng.integers(0, n)] = 2**32                     # the best particle always has full weight
            pre = np.cumsum(W, dtype=np.uint64)
            total = int(pre[-1])
            u0 = int(rng.integers(0, 2**32))
            anc = o.resample(pre, u0)
            assert (np.diff(anc) >= 0).all() and W[anc].min() > 0          # sorted, never a zero-weight ancestor
            counts = np.bincount(anc, minlength=n)
            expect = W.astype(np.float64) * n / total
            assert (np.abs(counts - expect) < 1.0 + 1e-9).all()             # systematic: |count - expectation| < 1
            for world in (1, 2, 3, 8):
                if world > n:
                    continue
                off = 0
                covered = 0
                for r in range(world):
                    b, c, per = m.ns_shard_range(n, world, r)
                    t = int(pre[b + c - 1]) - (int(pre[b - 1]) if b else 0) if c else 0
                    lo = m.ns_first_slot(off, total, n, u0)
                    hi = m.ns_first_slot(off + t, total, n, u0)
                    a = anc[lo:hi]
                    assert len(a) == 0 or (a.min() >= b and a.max() < b + c)
                    covered += hi - lo
                    off += t
                assert covered == n


def test_sharded_step_equals_single_span_step():
    """The multi-GPU protocol on CPU: shards compute local maxima / totals, exchange them, plan their slot ranges with
    the product's host code, resample locally and scatter by destination slot. Result == the one-span oracle."""
    sc, o = oracle_world(0)
    n = 3001
    P0 = o.init(0, n)
    motion = (0.02, 0.03, -0.01)
    scan = Scan(**sc.scans[0])
    want_P, want_anc, _, _ = o.step(P0.copy(), 0, scan, motion, step=0)
    for world in (2, 4, 8):
        new = np.zeros_like(P0)
        anc_all = np.full(n, -1, np.int64)
        shards = []
        for r in range(world):
            b, c, per = m.ns_shard_range(n, world, r)
            Pr = P0[b:b + c].copy()
            o.predict(Pr, b, *motion, 0)
            ll = o.loglik(Pr, o.beams(scan))
            shards.append((b, c, Pr, ll))
        gmax = max(float(s[3].max()) for s in shards)
        locs = [o.weights(s[3], gmax) for s in shards]
        totals = [l[3] for l in locs]
        total = sum(totals)
        u0 = o.u0(0)
        off = 0
        for (b, c, Pr, ll), (W, pre, wf, t) in zip(shards, locs):
            lo = m.ns_first_slot(off, total, n, u0)
            hi = m.ns_first_slot(off + t, total, n, u0)
            for k in range(lo, hi):
                thr = ((k << 32) + u0) * total // (n << 32)
                i = int(np.searchsorted(pre.astype(object) + off, thr, side="right"))
                new[k, :3] = Pr[i, :3]
                new[k, 3] = np.float32(1.0 / n)
                anc_all[k] = b + i
            off += t
        assert np.array_equal(anc_all, want_anc) and np.array_equal(new, want_P), world
