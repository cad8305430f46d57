"""GPU soak: hundreds of whole-tick calls (mcl_step: optimistic pre-pass, one-launch scans, pose summed by the resampling kernel,
report awaited by sequence number) against a twin filter driven through the separate calls, compared bit for bit along the way.
Guards the tick's host/device hand-shakes (stale abort flag, report sequence, restored host state after a re-run) over many
ticks rather than the handful of the parity tests."""
import numpy as np
import pytest

import montecarlolocalisation_b200 as m
from montecarlolocalisation_b200 import synth
from scenario import RES, Scenario

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,narrow", [(1500, False), (300, True), (5000, False)])
def test_many_ticks_match_the_separate_calls(n, narrow):
    ticks, n_scans = 600, 48
    sc = Scenario(n_scans, n_beams=360, seed=21)
    scans = [synth.make_scan(sc.occ, 0.1, sc.truth[s], 683, 900 + s, angle_min=np.float32(-120.0 * np.pi / 180.0),
                             angle_inc=np.float32(0.352 * np.pi / 180.0)) if s % 2 else sc.scans[s] for s in range(n_scans)]
    a = m.ParticleFilter(max_particles=n, seed=77)
    b = m.ParticleFilter(max_particles=n, seed=77)
    for pf in (a, b):
        pf.setMap(sc.occ, RES)
    if narrow:                                       # a cloud whose headings reach few ray-direction keys: pre-pass every tick
        rng = np.random.default_rng(8)
        P = np.zeros((n, 4), np.float32)
        P[:, 0] = sc.truth[0][0] + rng.uniform(-0.1, 0.1, n); P[:, 1] = sc.truth[0][1] + rng.uniform(-0.1, 0.1, n)
        P[:, 2] = sc.truth[0][2] + rng.uniform(-0.05, 0.05, n); P[:, 3] = 1
        a.uploadParticles(P); b.uploadParticles(P)
    else:
        a.sampleParticles(n); b.sampleParticles(n)
    for i in range(0, n_scans, 4):
        s_ = scans[i]
        a.stageScan(i, s_["ranges"], s_["angle_min"], s_["angle_inc"], s_["range_min"], s_["range_max"])
    for t in range(ticks):
        k = t % n_scans
        scan = scans[k]
        lost = (t // 37) % 3 == 0
        if t % 5 == 4:                               # some ticks only queued: the next waited tick must still be right
            a.executeParticleFilter(sc.enc_left[k], sc.enc_right[k], lost, scan=scan, want_result=False)
            pose_a = None
        elif k % 4 == 0:
            pose_a, st_a = a.executeParticleFilter(sc.enc_left[k], sc.enc_right[k], lost, slot=k)
        else:
            pose_a, st_a = a.executeParticleFilter(sc.enc_left[k], sc.enc_right[k], lost, scan=scan)
        b.diffDriveModel(sc.enc_left[k], sc.enc_right[k])
        b.computeWeight(scan["ranges"], scan["angle_min"], scan["angle_inc"], scan["range_min"], scan["range_max"])
        st_b = b.resampleParticles(lost)
        pose_b = b.estimateWeightedPose()
        if pose_a is not None:
            assert np.array_equal(pose_a, pose_b) and st_a == st_b, (t, pose_a, pose_b, st_a, st_b)
        if t % 50 == 49 or t == ticks - 1:
            assert np.array_equal(a.downloadParticles(), b.downloadParticles()), t
            assert np.array_equal(a.ancestors(), b.ancestors()), t
            assert np.array_equal(a.injectionState(), b.injectionState()), t
    if narrow:
        assert a.optimisticRedos() >= 1
