"""GPU parity, MCL_MODE_REF, scan edge cases: scans that leave one, two or a handful of scored beams, finite readings outside
[range_min, range_max] scattered through the scan (they are dropped BEFORE the stride-20 pick, so they shift which beams are
scored: Q5, MC:261-276, MC:613-617, MC:650), and re-staging a scan slot while ticks that read it are still queued.
Everything is compared with the CPU oracle bit for bit (weights, total, CDF, ancestors), through mcl_update + mcl_resample
and through the whole-tick call mcl_step / mcl_step_staged."""
import numpy as np
import pytest

import montecarlolocalisation_b200 as m
from oracle.pyoracle import Oracle, Scan
from scenario import RES, Scenario, load_map

pytestmark = pytest.mark.gpu


def make_pair(occ):
    o = Oracle(trig_mode=0)          # the host libm's sinf/cosf, like the engine's default MCL_TRIG_LIBM
    o.set_map(occ, RES)
    o.precompute_ray_directions(-120.0, 120.0, 0.1)
    pf = m.ParticleFilter()
    pf.setMap(occ, RES)
    return o, pf


def scored_beams(scan, stride=20):
    """How many beams the reference scores for this scan (filterLaserReadings + filterAngles + the stride pick)."""
    r = scan["ranges"].astype(np.float64)
    ang = np.float64(scan["angle_min"]) + np.arange(len(r)) * np.float64(scan["angle_inc"])
    keep = (np.isnan(r) | np.isinf(r)) | ((r >= scan["range_min"]) & (r <= scan["range_max"]))
    deg = ang * 180.0 / np.pi
    keep &= (deg > -120.0) & (deg < 120.0)
    return len(range(0, int(keep.sum()), stride))


def particles(n, seed):
    rng = np.random.default_rng(seed)
    P = np.zeros((n, 4), np.float32)
    P[:, 0] = rng.uniform(-0.2, 5.1, n)
    P[:, 1] = rng.uniform(-0.2, 5.1, n)
    P[:, 2] = rng.uniform(-3.2, 3.2, n)
    P[:, 3] = 1
    return P, rng


def one_step_against_oracle(scan, n=6000, seed=0):
    """computeWeight + resampleParticles with injected draws on both sides."""
    occ = load_map()
    o, pf = make_pair(occ)
    P, rng = particles(n, seed)
    u_r, u_jit = rng.random(n), rng.random(3 * n)
    inj = dict(u_yaw=rng.random(200), row=rng.integers(0, 6, 200).astype(np.int32), col=rng.integers(0, 6, 200).astype(np.int32),
               u_dx=rng.random(200), u_dy=rng.random(200))
    pf.uploadParticles(P)
    total_g = pf.computeWeight(scan["ranges"], scan["angle_min"], scan["angle_inc"], scan["range_min"], scan["range_max"])
    w_g = pf.downloadParticles()[:, 3].copy()
    Pc = P.copy()
    assert o.compute_weight(Pc, Scan(**scan)) == total_g and np.array_equal(w_g, Pc[:, 3])
    Pw = P.copy()
    Pnew, idx, cdf, st = o.resample(Pw, 1, Scan(**scan), u_r, u_jit, inj)       # weighs again itself (MC:468), then normalises in place
    assert total_g == st["total_weight"]
    sg = pf.resampleParticles(1, u_r, u_jit, inj)
    assert np.array_equal(pf.cdf(), cdf, equal_nan=True)
    assert np.array_equal(pf.ancestors(), idx)
    assert sg["injected"] == st["injected"]
    Pg = pf.downloadParticles()
    assert np.array_equal(Pg[:, [0, 1, 3]], Pnew[:, [0, 1, 3]])
    return total_g


def sparse_scan(n_valid, seed=3, n_beams=360):
    """A 360-beam scan whose readings are all 0.0 (below range_min: dropped) except n_valid in-FOV beams."""
    base = Scenario(1, n_beams=n_beams, seed=seed).scans[0]
    scan = dict(base)
    r = np.zeros(n_beams, np.float32)
    ang = np.float64(scan["angle_min"]) + np.arange(n_beams) * np.float64(scan["angle_inc"])
    deg = ang * 180.0 / np.pi
    inside = np.flatnonzero((deg > -119.0) & (deg < 119.0))
    pick = inside[np.linspace(0, len(inside) - 1, n_valid).astype(int)] if n_valid else np.zeros(0, int)
    src = np.nan_to_num(base["ranges"], nan=0.6, posinf=0.9)
    r[pick] = np.clip(src[pick], 0.1, 1.5)
    scan["ranges"] = r
    return scan


@pytest.mark.parametrize("n_valid,expect_scored", [(1, 1), (7, 1), (20, 1), (21, 2), (40, 2), (41, 3), (100, 5)])
def test_scans_that_leave_few_scored_beams(n_valid, expect_scored):
    """One scored beam used to divide by a 32-bit ceil(2^32 / 1) = 0 in the ray-parallel kernel."""
    scan = sparse_scan(n_valid)
    assert scored_beams(scan) == expect_scored
    total = one_step_against_oracle(scan, seed=n_valid)
    assert total > 0.0


@pytest.mark.parametrize("n_valid", [1, 21])
def test_few_scored_beams_through_the_whole_tick_call(n_valid):
    """mcl_step / mcl_step_staged with 1 and 2 scored beams: same particles and stats as the separate calls on a twin."""
    scan = sparse_scan(n_valid)
    sc = Scenario(2)
    n = 9000
    a = m.ParticleFilter(max_particles=n, seed=5)
    b = m.ParticleFilter(max_particles=n, seed=5)
    for pf in (a, b):
        pf.setMap(sc.occ, RES)
        pf.sampleParticles(n)
    a.stageScan(0, scan["ranges"], scan["angle_min"], scan["angle_inc"], scan["range_min"], scan["range_max"])
    for step in range(2):
        if step == 0:
            pose_a, st_a = a.executeParticleFilter(sc.enc_left[step], sc.enc_right[step], True, slot=0)
        else:
            pose_a, st_a = a.executeParticleFilter(sc.enc_left[step], sc.enc_right[step], True, scan=scan)
        b.diffDriveModel(sc.enc_left[step], sc.enc_right[step])
        total = b.computeWeight(scan["ranges"], scan["angle_min"], scan["angle_inc"], scan["range_min"], scan["range_max"])
        st_b = b.resampleParticles(True)
        assert st_a == st_b and st_a["total_weight"] == total and total > 0
        assert np.array_equal(pose_a, b.estimateWeightedPose())
        assert np.array_equal(a.downloadParticles(), b.downloadParticles())
        assert np.array_equal(a.ancestors(), b.ancestors())


@pytest.mark.parametrize("n_beams,seed", [(360, 1), (720, 2), (683, 3)])
def test_finite_out_of_range_readings_shift_the_stride_picks(n_beams, seed):
    """Q5: finite readings below range_min / above range_max are removed from the beam list before every 20th is picked, so
    one dropped reading moves every later pick by one beam. Scans with such readings scattered through them, with NaN and
    Inf in between (those stay, as 1.05, MC:268-272), against the oracle; and the pick really moved."""
    base = Scenario(1, n_beams=n_beams, seed=seed).scans[0]
    rng = np.random.default_rng(100 + seed)
    scan = dict(base)
    r = base["ranges"].copy()
    k = rng.choice(n_beams, n_beams // 6, replace=False)
    r[k[: len(k) // 3]] = np.float32(0.01)                       # below range_min = 0.02: dropped
    r[k[len(k) // 3: 2 * len(k) // 3]] = np.float32(7.5)         # above range_max = 5.6: dropped
    r[k[2 * len(k) // 3:]] = np.inf                              # kept (as 1.05)
    scan["ranges"] = r
    if n_beams == 683:                                           # the robot's own LIDAR geometry (comment MC:638-640)
        scan["angle_min"] = np.float32(-120.0 * np.pi / 180.0)
        scan["angle_inc"] = np.float32(0.352 * np.pi / 180.0)
        base = dict(base, angle_min=scan["angle_min"], angle_inc=scan["angle_inc"])
    assert scored_beams(scan) < scored_beams(base) or n_beams == 683
    one_step_against_oracle(scan, seed=seed)
    # the same scan with the dropped readings replaced by valid ones scores different beams: the totals differ
    scan2 = dict(scan)
    r2 = r.copy()
    r2[(r2 == np.float32(0.01)) | (r2 == np.float32(7.5))] = np.float32(0.5)
    scan2["ranges"] = r2
    occ = load_map()
    o, _ = make_pair(occ)
    P, _ = particles(4000, 9)
    Pa, Pb = P.copy(), P.copy()
    assert o.compute_weight(Pa, Scan(**scan)) != o.compute_weight(Pb, Scan(**scan2))


def test_restaging_a_slot_waits_for_queued_ticks():
    """mcl_scan_stage on a slot that queued ticks still read: the queued ticks see the scan that was in
    the slot when they were queued. A ring of 2 slots is re-staged while 6 ticks are in flight; a twin filter fed the same
    scans through host memory tick by tick must end in the same state."""
    sc = Scenario(6, n_beams=360, seed=7)
    n = 200_000
    a = m.ParticleFilter(max_particles=n, seed=11)
    b = m.ParticleFilter(max_particles=n, seed=11)
    for pf in (a, b):
        pf.setMap(sc.occ, RES)
        pf.sampleParticles(n)
    for step in range(6):
        scan = sc.scans[step]
        a.stageScan(step % 2, scan["ranges"], scan["angle_min"], scan["angle_inc"], scan["range_min"], scan["range_max"])
        a.executeParticleFilter(sc.enc_left[step], sc.enc_right[step], True, slot=step % 2, want_result=False)
        b.executeParticleFilter(sc.enc_left[step], sc.enc_right[step], True, scan=scan, want_result=False)
    assert np.array_equal(a.downloadParticles(), b.downloadParticles())
    assert np.array_equal(a.ancestors(), b.ancestors())
    assert np.array_equal(a.injectionState(), b.injectionState())
