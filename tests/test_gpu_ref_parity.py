"""GPU parity, MCL_MODE_REF: the CUDA engine (through the C-ABI) against the CPU oracle and the reference's golden
vectors. Bar (BASELINE.json north_star): resampled indices and counts bit-exact; poses and weights within 1e-5
relative. In practice every stage below is required to be bit-identical to the oracle, except theta after
atan2(sin,cos) and the pose estimate, which go through device libm (<= 1 fp32 ulp / 1e-5).
Float trig (MC:644-645, 747-748) comes in two definitions on both sides: "libm" = what the reference binary computes
(engine MCL_TRIG_LIBM, the default, against oracle trig_mode 0 = this host's sinf/cosf) and "cr" = correctly rounded
(engine MCL_TRIG_CORRECTLY_ROUNDED against oracle trig_mode 1). With "libm" the engine is also bit-identical to the golden
vectors recorded from the compiled reference itself."""
import ctypes as C
import os

import numpy as np
import pytest

import montecarlolocalisation_b200 as m
from oracle.pyoracle import Oracle, Scan, oracle_lib
from scenario import RES, Scenario, load_map

pytestmark = pytest.mark.gpu

G = os.path.join(os.path.dirname(__file__), "golden", "ref_config1.npz")


def draws_for(rng, n, n_rows, n_cols, max_inj=200):
    inj = dict(u_yaw=rng.random(max_inj), row=rng.integers(0, n_rows, max_inj).astype(np.int32),
               col=rng.integers(0, n_cols, max_inj).astype(np.int32), u_dx=rng.random(max_inj), u_dy=rng.random(max_inj))
    return rng.random(n), rng.random(3 * n), inj


def ulp_diff(a, b):
    return np.abs(a.astype(np.float64) - b.astype(np.float64)) / np.spacing(np.maximum(np.abs(a), np.abs(b)).astype(np.float32))


def make_pair(occ, trig="libm"):
    o = Oracle(trig_mode=0 if trig == "libm" else 1)
    o.set_map(occ, RES)
    o.precompute_ray_directions(-120.0, 120.0, 0.1)
    pf = m.ParticleFilter(trig_mode=m.TRIG_LIBM if trig == "libm" else m.TRIG_CORRECTLY_ROUNDED)
    pf.setMap(occ, RES)
    return o, pf


def run_loop(n, steps, seed, kidnap_at=None, jitter=None, n_beams=360, settle_injection_at=None, trig="libm"):
    """Free-running predict/update/resample/estimate on both sides with identical injected draws."""
    sc = Scenario(steps, n_beams=n_beams, kidnap_at=kidnap_at)
    o, pf = make_pair(sc.occ, trig)
    rng = np.random.default_rng(seed)
    n_rows, n_cols = o.cell_ranges()
    init = dict(u_yaw=rng.random(n), row=rng.integers(0, n_rows, n).astype(np.int32), col=rng.integers(0, n_cols, n).astype(np.int32),
                u_dx=rng.random(n), u_dy=rng.random(n))
    P = o.sample_particles(init["u_yaw"], init["row"], init["col"], init["u_dx"], init["u_dy"])
    pf.sampleParticles(n, init)
    assert np.array_equal(pf.downloadParticles(), P)
    injected = 0
    for s in range(steps):
        js = 1 if jitter is None else jitter[s % len(jitter)]
        if settle_injection_at and s in settle_injection_at:
            # the slow EMA lags for dozens of steps after start-up; jump to a chosen state so this step injects
            o.set_injection_state(*settle_injection_at[s])
            pf.setInjectionState(*settle_injection_at[s])
        z = rng.standard_normal(3)
        mo = o.diff_drive(sc.enc_left[s], sc.enc_right[s], z)
        mg = pf.diffDriveModel(sc.enc_left[s], sc.enc_right[s], z)
        assert np.array_equal(mo, mg)
        o.update_particle_pos(P)
        assert np.array_equal(pf.downloadParticles(), P), "predict step %d" % s
        scan = sc.scans[s]
        u_r, u_jit, inj = draws_for(rng, n, n_rows, n_cols)
        total_g = pf.computeWeight(scan["ranges"], scan["angle_min"], scan["angle_inc"], scan["range_min"], scan["range_max"])
        Pw = P.copy()
        Pnew, idx, cdf, st = o.resample(Pw, js, Scan(**scan), u_r, u_jit, inj)
        assert total_g == st["total_weight"], "total weight step %d" % s
        sg = pf.resampleParticles(js, u_r, u_jit, inj)
        assert np.array_equal(pf.cdf(), cdf, equal_nan=True), "cdf step %d" % s
        assert np.array_equal(pf.ancestors(), idx), "ancestor indices step %d" % s
        assert sg["injected"] == st["injected"] and sg["clamped"] == o.clamp_count()
        assert sg["p_inject"] == st["p_inject"] and sg["weight_slow"] == st["weight_slow"] and sg["weight_fast"] == st["weight_fast"]
        Pg = pf.downloadParticles()
        assert np.array_equal(Pg[:, [0, 1, 3]], Pnew[:, [0, 1, 3]]), "resampled x,y,w step %d" % s
        assert ulp_diff(Pg[:, 2], Pnew[:, 2]).max() <= 1.0, "resampled theta step %d" % s
        pose_o = o.estimate_weighted_pose(Pnew)
        pose_g = pf.estimateWeightedPose()
        assert np.allclose(pose_g, pose_o, rtol=1e-5, atol=1e-5), "pose step %d" % s
        injected += st["injected"]
        # keep both sides on the same state even if theta differed in the last bit
        if not np.array_equal(Pg, Pnew):
            pf.uploadParticles(Pnew)
        P = Pnew
    ko, xo, yo = o.ray_lut(-301, 301)
    kg, xg, yg = pf.rayLut()
    assert np.array_equal(ko, kg) and np.array_equal(xo, xg) and np.array_equal(yo, yg), "ray-direction LUT state"
    return injected


@pytest.mark.parametrize("trig", ["libm", "cr"])
def test_config1_loop_1000_particles(trig):
    """BASELINE.json config 1: map.txt, 1000 particles, 360-beam scan + odometry trace."""
    run_loop(1000, 40, seed=1, trig=trig)


def test_config1_kidnap_triggers_injection():
    injected = run_loop(1500, 14, seed=2, kidnap_at=6, jitter=[1, 1, 0, 0, 0, 0, 0, 0, 1, 1, 0, 0, 0, 0],
                        settle_injection_at={5: (6.0, 6.0), 8: (30.0, 30.0)})
    assert injected >= 50 + 200          # the confident cap (MC:479) at step 6 and the lost cap (MC:474) at step 8 were hit


def test_guide_table_search_with_injection():
    """Above 4096 particles the CDF search goes through the guide table (k_ref_guide): same ancestors as std::lower_bound
    in the oracle, with injections consuming part of the slots, over a kidnap."""
    injected = run_loop(20_000, 12, seed=6, kidnap_at=5, jitter=[1, 1, 0, 0, 0, 0, 1, 1, 0, 0, 0, 0],
                        settle_injection_at={4: (6.0, 6.0), 7: (30.0, 30.0)})
    assert injected > 0


@pytest.mark.parametrize("layout", ["few-heavy", "one-tile-crowded", "spread"])
def test_guide_table_forms_agree_and_match_oracle(layout):
    """The guide table of the CDF search is scattered by the one-kernel CDF accumulation (exact_scan_fused.cuh); its three other
    builders - the separate k_ref_guide launch behind the multi-launch scan (force bit 6), the rebuild after the in-kernel
    single-chain fallback (bit 7) and no table at all (bit 0: full-range search) - must pick the same ancestors, and all of
    them the oracle's std::lower_bound. Layouts: a handful of particles own hundreds of bucket edges each; more wide ranges
    in one tile than the block-wide queue holds; ordinary."""
    occ = load_map()
    rng = np.random.default_rng(31)
    n = 20000
    P = np.zeros((n, 4), np.float32)
    P[:, 0] = -50.0; P[:, 1] = -50.0; P[:, 3] = 1                       # off the map: weight 0
    def valid(k):
        q = np.zeros((k, 4), np.float32)
        q[:, 0] = rng.uniform(2.2, 2.6, k); q[:, 1] = rng.uniform(2.2, 2.6, k); q[:, 2] = rng.uniform(-3.1, 3.1, k); q[:, 3] = 1
        return q
    if layout == "few-heavy":
        where = np.sort(rng.choice(n, 10, replace=False))
    elif layout == "one-tile-crowded":
        where = 8192 + np.sort(rng.choice(4096, 200, replace=False))
    else:
        where = np.arange(n)
    P[where] = valid(len(where))
    scan = Scenario(1).scans[0]
    args = (scan["ranges"], scan["angle_min"], scan["angle_inc"], scan["range_min"], scan["range_max"])
    n_rows, n_cols = 49, 49
    u_r, u_jit, inj = draws_for(rng, n, n_rows, n_cols)
    o = Oracle(trig_mode=0); o.set_map(occ, RES); o.precompute_ray_directions(-120.0, 120.0, 0.1)
    Pnew, idx, cdf, st = o.resample(P.copy(), 1, Scan(**scan), u_r, u_jit, inj)
    for bits in (0, 64, 128, 1):
        pf = m.ParticleFilter(); pf.setMap(occ, RES)
        pf.forceSequential(bits)
        pf.uploadParticles(P)
        pf.computeWeight(*args)
        pf.resampleParticles(1, u_r, u_jit, inj)
        assert np.array_equal(pf.cdf(), cdf), "cdf, force bits %d" % bits
        assert np.array_equal(pf.ancestors(), idx), "ancestors, force bits %d" % bits
    # the one-kernel form built the table itself: no k_ref_guide launch
    pf = m.ParticleFilter(); pf.setMap(occ, RES)
    pf.uploadParticles(P); pf.profileEnable(True)
    pf.computeWeight(*args); pf.resampleParticles(1, u_r, u_jit, inj)
    names = pf.profileRead()
    assert "k_xs_cdf" in names and "k_ref_guide" not in names, sorted(names)

@pytest.mark.parametrize("n_beams", [720, 1080])
def test_more_beams(n_beams):
    run_loop(2000, 6, seed=3, n_beams=n_beams)


def test_ragged_sizes():
    for n in (1, 2, 31, 257, 1025, 4097):
        run_loop(n, 3, seed=100 + n)


@pytest.mark.parametrize("trig", ["libm", "cr"])
def test_config2_one_million_particles(trig):
    """BASELINE.json config 2 at full size: 1M particles; indices bit-exact against the oracle."""
    run_loop(1_000_000, 2, seed=4, trig=trig)


def test_stagewise_against_reference_golden():
    """Stage by stage from the golden inputs recorded from the COMPILED REFERENCE (tests/golden/make_golden.py): with the
    default float trig (MCL_TRIG_LIBM = the reference binary's sinf/cosf) the engine reproduces it bit for bit: predicted
    particles, injected count, resampled particles and the adaptive-injection state."""
    g = np.load(G)
    pf = m.ParticleFilter()
    pf.setMap(g["occ"], RES)
    P = g["P0"].copy()
    for s in range(int(g["steps"])):
        pf.uploadParticles(P)
        mo = g["motion%d" % s]
        pf.updateParticlePos(mo[0], mo[1], mo[2])
        pred = pf.downloadParticles()
        assert np.array_equal(pred[:, :3], g["pred%d" % s][:, :3]), "predict step %d" % s
        pf.computeWeight(g["scan%d_ranges" % s], g["scan%d_angle_min" % s], g["scan%d_angle_inc" % s], g["scan%d_range_min" % s],
                         g["scan%d_range_max" % s])
        a = g["inj%d" % s]
        inj = dict(u_yaw=a[0], row=a[1].astype(np.int32), col=a[2].astype(np.int32), u_dx=a[3], u_dy=a[4])
        st = pf.resampleParticles(int(g["jitter"][s]), g["u_r%d" % s], g["u_jit%d" % s], inj)
        assert st["injected"] == int(g["injected%d" % s])
        new = pf.downloadParticles()
        gold_new = g["new%d" % s]
        mism = (new[:, :3] != gold_new[:, :3]).any(axis=1)
        assert mism.sum() == 0, "step %d: %d particles differ from the reference" % (s, mism.sum())
        assert np.array_equal(pf.injectionState(), g["inj_state%d" % s])
        P = gold_new


def test_golden_with_correctly_rounded_trig_differs_only_by_rounding():
    """The portable float trig against the same golden vectors: a predicted coordinate may move by one fp32 ulp (glibc's
    sinf/cosf are not correctly rounded for 0.26 % / 0.55 % of arguments, which moves x + trans * cos only when the sum sits at
    a rounding boundary); nothing else may change."""
    g = np.load(G)
    pf = m.ParticleFilter(trig_mode=m.TRIG_CORRECTLY_ROUNDED)
    pf.setMap(g["occ"], RES)
    P = g["P0"].copy()
    for s in range(int(g["steps"])):
        pf.uploadParticles(P)
        mo = g["motion%d" % s]
        pf.updateParticlePos(mo[0], mo[1], mo[2])
        pred = pf.downloadParticles()
        assert ulp_diff(pred[:, :3], g["pred%d" % s][:, :3]).max() <= 1.0
        P = g["new%d" % s]


def test_edge_total_weight_zero_and_empty_scan():
    occ = load_map()
    o, pf = make_pair(occ)
    n = 300
    P = np.zeros((n, 4), np.float32)
    P[:, 0] = -3.0
    P[:, 3] = 1.0
    pf.uploadParticles(P)
    empty = np.zeros(0, np.float32)
    assert pf.computeWeight(empty, 0.0, 0.0, 0.0, 0.0) == 0.0           # Q25: no scan yet
    rng = np.random.default_rng(0)
    u_r, u_jit, inj = draws_for(rng, n, 6, 6)
    Pw = P.copy()
    Pnew, idx, cdf, st = o.resample(Pw, 1, Scan(empty, 0.0, 0.0, 0.0, 0.0), u_r, u_jit, inj)
    sg = pf.resampleParticles(1, u_r, u_jit, inj)
    assert sg["p_inject"] == 0.0 and sg["injected"] == 0
    assert np.isnan(pf.cdf()).all() and (pf.ancestors() == 0).all() and (idx == 0).all()
    Pg = pf.downloadParticles()
    assert np.array_equal(Pg[:, [0, 1, 3]], Pnew[:, [0, 1, 3]])


def test_error_behaviour():
    pf = m.ParticleFilter()
    with pytest.raises(m.MclError):
        pf.sampleParticles(10)                       # no map yet
    pf.setMap(load_map(), RES)
    pf.sampleParticles(10)
    with pytest.raises(m.MclError):
        pf.resampleParticles(1)                      # update must come first
    with pytest.raises(m.MclError):
        pf.uploadParticles(np.zeros((0, 4), np.float32))
    assert pf.kernelLaunches() > 0


def test_philox_production_draws_run():
    """Without injected draws the engine draws from its own Philox stream: finite particles, valid ancestors."""
    occ = load_map()
    pf = m.ParticleFilter()
    pf.setMap(occ, RES)
    sc = Scenario(3)
    pf.sampleParticles(5000)
    for s in range(3):
        pf.diffDriveModel(sc.enc_left[s], sc.enc_right[s])
        scan = sc.scans[s]
        pf.computeWeight(scan["ranges"], scan["angle_min"], scan["angle_inc"], scan["range_min"], scan["range_max"])
        pf.resampleParticles(1)
        P = pf.downloadParticles()
        idx = pf.ancestors()
        assert np.isfinite(P).all() and idx.min() >= -1 and idx.max() < 5000


def _philox(c0, c2, c3, seed):
    out = (C.c_uint32 * 4)()
    oracle_lib().ons_philox(C.c_uint32(c0), C.c_uint32(0), C.c_uint32(c2), C.c_uint32(c3), C.c_uint32(seed & 0xffffffff), C.c_uint32(seed >> 32), out)
    return [int(v) for v in out]


def _c53(hi, lo):
    return float(((hi << 32) | lo) >> 11) * 2.0 ** -53


@pytest.mark.parametrize("whole_step", [False, True])
def test_philox_injected_particles_are_sample_particles_of_the_named_draws(whole_step):
    """Production draws: an injected particle is sampleParticles(1) (MC:434-446) of the named draws the resampling kernel
    takes from Philox stream 0x31, counters 2*rank / 2*rank+1 at the tick's step number (rank = how many slots were
    injected before this one). Checked bit for bit against the oracle's sampleParticles fed with the same words generated
    on the CPU, through the per-function calls and through the whole-tick call."""
    seed, n = 0x123456789, 20011
    sc = Scenario(1, n_beams=360, seed=9)
    o = Oracle(trig_mode=0)
    o.set_map(sc.occ, RES)
    n_rows, n_cols = o.cell_ranges()
    pf = m.ParticleFilter(max_particles=n, seed=seed)
    pf.setMap(sc.occ, RES)
    pf.sampleParticles(n)
    pf.setInjectionState(10.0, 0.0)               # p_inject = 1 - fast/slow well above zero on this tick
    scan = sc.scans[0]
    if whole_step:
        _, st = pf.executeParticleFilter(sc.enc_left[0], sc.enc_right[0], True, scan=scan)
    else:
        pf.diffDriveModel(sc.enc_left[0], sc.enc_right[0])
        pf.computeWeight(scan["ranges"], scan["angle_min"], scan["angle_inc"], scan["range_min"], scan["range_max"])
        st = pf.resampleParticles(True)
    step = 1                                      # the engine's step number: +1 per predict, +1 per resample
    idx = pf.ancestors()
    P = pf.downloadParticles()
    slots = np.flatnonzero(idx == -1)
    assert st["p_inject"] > 0.0 and st["injected"] == len(slots) == 200          # max_injection in lost mode (MC:474)
    u_yaw, u_dx, u_dy, row, col = [], [], [], [], []
    for rank in range(len(slots)):
        a, b = _philox(2 * rank, 0x31, step, seed), _philox(2 * rank + 1, 0x31, step, seed)
        u_yaw.append(_c53(a[0], a[1])); u_dx.append(_c53(a[2], a[3])); u_dy.append(_c53(b[0], b[1]))
        row.append(b[2] % n_rows); col.append(b[3] % n_cols)
    E = o.sample_particles(np.array(u_yaw), np.array(row, np.int32), np.array(col, np.int32), np.array(u_dx), np.array(u_dy))
    assert np.array_equal(P[slots, :3], E[:, :3])
    assert np.all(P[slots, 3] == np.float32(1.0 / n))


def test_ray_parallel_update_kernel_matches_per_particle_kernel_and_oracle():
    """The restructured computeWeight kernel (rays as the unit of work, division-free cell lookup) against the plain
    one-thread-per-particle kernel and the oracle, on particles that straddle walls, edges and the outside of the map."""
    occ = load_map()
    rng = np.random.default_rng(11)
    n = 300_000
    P = np.zeros((n, 4), np.float32)
    P[:, 0] = rng.uniform(-0.4, 5.3, n)
    P[:, 1] = rng.uniform(-0.4, 5.3, n)
    P[:, 2] = rng.uniform(-9, 9, n)
    # exact cell edges and exact integers stress the near-integer slow path
    P[:5000, 0] = np.float32(0.1) * rng.integers(0, 50, 5000)
    P[5000:10000, 1] = np.float32(0.1) * rng.integers(0, 50, 5000)
    P[:, 3] = 1
    sc = Scenario(1)
    scan = sc.scans[0]
    args = (scan["ranges"], scan["angle_min"], scan["angle_inc"], scan["range_min"], scan["range_max"])
    o, pf = make_pair(occ)
    Po = P.copy()
    total_o = o.compute_weight(Po, Scan(**scan))
    pf.uploadParticles(P)
    total_v2 = pf.computeWeight(*args)
    w_v2 = pf.downloadParticles()[:, 3].copy()
    pf2 = m.ParticleFilter()
    pf2.setMap(occ, RES)
    pf2.forceSequential(2)                  # bit 1: per-particle kernel
    pf2.uploadParticles(P)
    total_v1 = pf2.computeWeight(*args)
    w_v1 = pf2.downloadParticles()[:, 3]
    pf3 = m.ParticleFilter()
    pf3.setMap(occ, RES)
    pf3.forceSequential(4)                  # bit 2: ray-parallel kernel, f64 probes only (no fp32 pre-filter)
    pf3.uploadParticles(P)
    total_v3 = pf3.computeWeight(*args)
    w_v3 = pf3.downloadParticles()[:, 3]
    assert np.array_equal(w_v2, Po[:, 3]) and np.array_equal(w_v1, Po[:, 3]) and np.array_equal(w_v3, Po[:, 3])
    assert total_v2 == total_o == total_v1 == total_v3
    assert "k_ref_update_v2" not in pf2.profileRead()


def test_large_map_not_in_shared_memory():
    """REF mode on a 1025x1025 maze: the occupancy grid is read through L2 instead of shared memory."""
    from montecarlolocalisation_b200 import synth
    occ = synth.maze_occupancy(128, 3)
    rng = np.random.default_rng(12)
    n = 20000
    P = np.zeros((n, 4), np.float32)
    P[:, 0] = rng.uniform(0, 102.5, n); P[:, 1] = rng.uniform(0, 102.5, n); P[:, 2] = rng.uniform(-3.2, 3.2, n); P[:, 3] = 1
    scan = synth.make_scan(occ, 0.1, (30.45, 40.45, 0.3), 360, 5)
    o = Oracle(trig_mode=0); o.set_map(occ, RES); o.precompute_ray_directions()
    pf = m.ParticleFilter(); pf.setMap(occ, RES)
    Po = P.copy()
    total_o = o.compute_weight(Po, Scan(**scan))
    pf.uploadParticles(P)
    total_g = pf.computeWeight(scan["ranges"], scan["angle_min"], scan["angle_inc"], scan["range_min"], scan["range_max"])
    assert total_g == total_o and np.array_equal(pf.downloadParticles()[:, 3], Po[:, 3])


def test_nonzero_map_origin():
    """origin != (0,0): the reference subtracts it in worldToMap (MC:300-305) and isInsideMap (MC:686-689)."""
    occ = load_map()
    ox, oy = 1.5, -2.25
    rng = np.random.default_rng(13)
    n = 50_000
    P = np.zeros((n, 4), np.float32)
    P[:, 0] = rng.uniform(-0.4, 5.3, n) + ox
    P[:, 1] = rng.uniform(-0.4, 5.3, n) + oy
    P[:, 2] = rng.uniform(-4, 4, n)
    P[:, 3] = 1
    scan = Scenario(1).scans[0]
    o = Oracle(trig_mode=1); o.set_map(occ, RES, ox, oy); o.precompute_ray_directions()
    pf = m.ParticleFilter(trig_mode=m.TRIG_CORRECTLY_ROUNDED); pf.setMap(occ, RES, ox, oy)
    Po = P.copy()
    total_o = o.compute_weight(Po, Scan(**scan))
    pf.uploadParticles(P)
    total_g = pf.computeWeight(scan["ranges"], scan["angle_min"], scan["angle_inc"], scan["range_min"], scan["range_max"])
    assert (Po[:, 3] > 0).mean() > 0.3
    assert total_g == total_o and np.array_equal(pf.downloadParticles()[:, 3], Po[:, 3])


@pytest.mark.parametrize("n", [1000, 8000, 30011])          # one tile; several tiles with the pose summed by the resampling kernel; neither
def test_whole_step_call_equals_separate_calls(n):
    """mcl_step / mcl_step_staged (one tick enqueued as one piece, the host waiting once) against the four separate calls
    on a twin filter with the same seed: same particles, ancestors, injection state, stats and pose, over steps that
    include injections (lost mode after a weight collapse) and both scan paths."""
    sc = Scenario(6, n_beams=360, seed=3)          # n = 1000: below the guide-table threshold, one scan tile
    a = m.ParticleFilter(max_particles=n, seed=77)
    b = m.ParticleFilter(max_particles=n, seed=77)
    for pf in (a, b):
        pf.setMap(sc.occ, RES)
        pf.sampleParticles(n)
    for i, scan in enumerate(sc.scans):
        a.stageScan(i, scan["ranges"], scan["angle_min"], scan["angle_inc"], scan["range_min"], scan["range_max"])
    for step in range(6):
        scan = dict(sc.scans[step])
        if step == 3:                       # a scan that fits nothing: weights collapse, the injection EMA reacts on the next steps
            scan["ranges"] = np.full_like(scan["ranges"], 0.05)
            a.stageScan(step, scan["ranges"], scan["angle_min"], scan["angle_inc"], scan["range_min"], scan["range_max"])
        lost = step >= 2
        if step % 2 == 0:
            pose_a, st_a = a.executeParticleFilter(sc.enc_left[step], sc.enc_right[step], lost, slot=step)
        else:
            pose_a, st_a = a.executeParticleFilter(sc.enc_left[step], sc.enc_right[step], lost, scan=scan)
        b.diffDriveModel(sc.enc_left[step], sc.enc_right[step])
        total = b.computeWeight(scan["ranges"], scan["angle_min"], scan["angle_inc"], scan["range_min"], scan["range_max"])
        st_b = b.resampleParticles(lost)
        pose_b = b.estimateWeightedPose()
        assert st_a == st_b and st_a["total_weight"] == total, (step, st_a, st_b)
        assert np.array_equal(pose_a, pose_b), (step, pose_a, pose_b)
        assert np.array_equal(a.downloadParticles(), b.downloadParticles()), step
        assert np.array_equal(a.ancestors(), b.ancestors()), step
    assert np.array_equal(a.injectionState(), b.injectionState())


def test_optimistic_tick_equals_separate_calls_on_the_robot_scan():
    """A small cloud with a narrow spread of headings on the robot's 683-beam / 0.352-degree scan (35 scored beams): most
    ray-direction keys are never touched, so the table stays incomplete and EVERY tick needs the first-touch pre-pass.
    mcl_step enqueues the tick behind the pre-pass without waiting for it and runs it again only when the pre-pass found a
    new key (DESIGN.md section 3 "Tick plumbing"): same particles, ancestors, pose, injection state and ray table as the
    separate calls (which wait for the pre-pass every tick) over 40 ticks; some ticks ran twice, most ran once behind an
    unwaited pre-pass (5 launches: pre-pass 2 + tick 3)."""
    from montecarlolocalisation_b200 import synth
    n, ticks = 300, 40
    sc = Scenario(ticks, n_beams=360, seed=9)
    scans = [synth.make_scan(sc.occ, 0.1, sc.truth[s], 683, 700 + s, angle_min=np.float32(-120.0 * np.pi / 180.0),
                             angle_inc=np.float32(0.352 * np.pi / 180.0)) for s in range(ticks)]
    rng = np.random.default_rng(5)
    P = np.zeros((n, 4), np.float32)
    P[:, 0] = sc.truth[0][0] + rng.uniform(-0.1, 0.1, n); P[:, 1] = sc.truth[0][1] + rng.uniform(-0.1, 0.1, n)
    P[:, 2] = sc.truth[0][2] + rng.uniform(-0.05, 0.05, n); P[:, 3] = 1
    a = m.ParticleFilter(max_particles=n, seed=31)
    b = m.ParticleFilter(max_particles=n, seed=31)
    for pf in (a, b):
        pf.setMap(sc.occ, RES)
        pf.uploadParticles(P)
    for i in range(0, ticks, 3):
        a.stageScan(i, scans[i]["ranges"], scans[i]["angle_min"], scans[i]["angle_inc"], scans[i]["range_min"], scans[i]["range_max"])
    unwaited = 0
    for step in range(ticks):
        scan = scans[step]
        lost = step % 7 < 2
        l0, r0 = a.kernelLaunches(), a.optimisticRedos()
        if step % 3 == 0:
            pose_a, st_a = a.executeParticleFilter(sc.enc_left[step], sc.enc_right[step], lost, slot=step)
        else:
            pose_a, st_a = a.executeParticleFilter(sc.enc_left[step], sc.enc_right[step], lost, scan=scan)
        if a.optimisticRedos() == r0 and a.kernelLaunches() - l0 == 5:          # pre-pass 2 + tick 3 (computeWeight; total, injection counts and CDF; resampling with the pose sums)
            unwaited += 1
        b.diffDriveModel(sc.enc_left[step], sc.enc_right[step])
        total = b.computeWeight(scan["ranges"], scan["angle_min"], scan["angle_inc"], scan["range_min"], scan["range_max"])
        st_b = b.resampleParticles(lost)
        pose_b = b.estimateWeightedPose()
        assert st_a == st_b and st_a["total_weight"] == total, (step, st_a, st_b)
        assert np.array_equal(pose_a, pose_b), (step, pose_a, pose_b)
        assert np.array_equal(a.downloadParticles(), b.downloadParticles()), step
        assert np.array_equal(a.ancestors(), b.ancestors()), step
        for x, y in zip(a.rayLut(), b.rayLut()):
            assert np.array_equal(x, y), step
    assert np.array_equal(a.injectionState(), b.injectionState())
    redos = a.optimisticRedos()
    assert redos >= 1 and unwaited >= 10, (redos, unwaited)
    assert b.optimisticRedos() == 0


@pytest.mark.parametrize("n", [700, 1024, 1500, 2048, 3000])
def test_one_launch_total_and_cdf_below_one_tile(n):
    """Below 4096 particles mcl_step accumulates the weight total (+ adaptive-injection state) and the normalised CDF in one
    launch (k_ref_scans_one_tile), the injection counts included, and the resampling kernel also sums the pose and writes the
    tick's report: three launches per tick. The exact-scan passes take 4 weights per thread up to 1024 particles, 8 up to
    2048, else 16. Same particles, ancestors, CDF, injection state and pose as with a large filter's launches (force bit 10),
    with 16 weights per thread (bit 11) and with the in-kernel single-chain fallback forced in both passes (bit 7), over
    ticks that include a weight collapse; and the unforced filters never need the fallback."""
    ticks = 8
    sc = Scenario(ticks, n_beams=360, seed=4)
    scans = [dict(s_) for s_ in sc.scans]
    scans[3]["ranges"] = np.full_like(scans[3]["ranges"], 0.05)          # nothing fits: the total collapses, injections follow
    pfs = []
    for bits in (0, 1024, 2048, 128):
        pf = m.ParticleFilter(max_particles=n, seed=55)
        pf.setMap(sc.occ, RES)
        pf.sampleParticles(n)
        pf.forceSequential(bits)
        pfs.append(pf)
    pfs[0].profileEnable(True)
    for step in range(ticks):
        out = [pf.executeParticleFilter(sc.enc_left[step], sc.enc_right[step], step >= 2, scan=scans[step]) for pf in pfs]
        P0, A0, C0 = pfs[0].downloadParticles(), pfs[0].ancestors(), pfs[0].cdf()
        for pf, o in zip(pfs[1:], out[1:]):
            assert o[1] == out[0][1] and np.array_equal(o[0], out[0][0]), step
            assert np.array_equal(pf.downloadParticles(), P0) and np.array_equal(pf.ancestors(), A0), step
            assert np.array_equal(pf.cdf(), C0, equal_nan=True), step
        if np.isfinite(out[0][1]["total_weight"]) and out[0][1]["total_weight"] > 0:
            assert not pfs[0].lastScanFellBack() and not pfs[1].lastScanFellBack() and not pfs[2].lastScanFellBack(), step
        assert pfs[3].lastScanFellBack()
    names = pfs[0].profileRead()
    assert "k_ref_scans_one_tile" in names and not {"k_xs_cdf", "k_xs_total", "k_ref_inject_count", "k_pose_sums"} & set(names), sorted(names)
    for pf in pfs[1:]:
        assert np.array_equal(pfs[0].injectionState(), pf.injectionState())


def test_whole_step_calls_queued_without_waiting():
    """mcl_step with no outputs asked for returns as soon as the tick is queued (the adaptive-injection state advances on
    the device): several ticks in flight, scans from host memory and from staged slots, then the same state as a twin
    filter driven tick by tick through the separate calls; the two APIs can be mixed on one filter."""
    sc = Scenario(8, n_beams=360, seed=5)
    n = 25013
    a = m.ParticleFilter(max_particles=n, seed=99)
    b = m.ParticleFilter(max_particles=n, seed=99)
    for pf in (a, b):
        pf.setMap(sc.occ, RES)
        pf.sampleParticles(n)
    scans = [dict(s) for s in sc.scans]
    scans[4]["ranges"] = np.full_like(scans[4]["ranges"], 0.05)          # weights collapse at tick 4: injections follow
    for i, scan in enumerate(scans):
        a.stageScan(i, scan["ranges"], scan["angle_min"], scan["angle_inc"], scan["range_min"], scan["range_max"])

    def separate(pf, step, lost):
        scan = scans[step]
        pf.diffDriveModel(sc.enc_left[step], sc.enc_right[step])
        pf.computeWeight(scan["ranges"], scan["angle_min"], scan["angle_inc"], scan["range_min"], scan["range_max"])
        return pf.resampleParticles(lost)

    for step in range(7):
        lost = step >= 3
        if step == 2:                                   # one tick through the separate calls in the middle of the queued ones
            separate(a, step, lost)
        elif step % 2 == 0:
            a.executeParticleFilter(sc.enc_left[step], sc.enc_right[step], lost, slot=step, want_result=False)
        else:
            a.executeParticleFilter(sc.enc_left[step], sc.enc_right[step], lost, scan=scans[step], want_result=False)
        separate(b, step, lost)
    assert np.array_equal(a.injectionState(), b.injectionState())
    assert np.array_equal(a.downloadParticles(), b.downloadParticles())
    assert np.array_equal(a.ancestors(), b.ancestors())
    # and a last tick that reads its results back
    pose_a, st_a = a.executeParticleFilter(sc.enc_left[7], sc.enc_right[7], True, slot=7)
    st_b = separate(b, 7, True)
    assert st_a == st_b
    assert np.array_equal(pose_a, b.estimateWeightedPose())
    assert st_a["injected"] + st_b["injected"] >= 0
