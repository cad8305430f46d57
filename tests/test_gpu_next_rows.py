"""GPU parity for the rows either side of the hot path: k-means confidence estimate and pose array through
the C-ABI against the oracle (itself pinned bit-exact to the compiled reference, tests/test_oracle_next_rows.py)."""
import numpy as np
import pytest

import montecarlolocalisation_b200 as m
from oracle import pyoracle
from scenario import RES, Scenario
from test_oracle_next_rows import CASES, clustered_particles, rand_draws

pytestmark = pytest.mark.gpu


def make_pf(P, **cfg):
    pf = m.ParticleFilter(**cfg)
    pf.setMap(Scenario(1).occ, RES)
    pf.uploadParticles(P)
    return pf


@pytest.mark.parametrize("name,centres,spread,n", CASES)
@pytest.mark.parametrize("seed", [1, 7, 1234567])
def test_kmeans_confidence_bit_exact_at_reference_sizes(name, centres, spread, n, seed):
    """N <= MCL_KMEANS_EXACT_MAX: sequential fp32 centre sums on the device -> everything but theta bit-exact."""
    rng = np.random.default_rng(hash(name) % 2**32 + seed)
    if centres is None:
        P = np.zeros((n, 4), np.float32)
        P[:, 0] = rng.uniform(0, 4.8, n); P[:, 1] = rng.uniform(0, 4.8, n); P[:, 2] = rng.uniform(-np.pi, np.pi, n); P[:, 3] = rng.random(n)
    else:
        P = clustered_particles(rng, n, centres, spread, weights="random" if seed % 2 else "uniform")
    pf = make_pf(P)
    draws = rand_draws(seed, n)
    for thr in (0.3, 0.95):
        o = pyoracle.kmeans_confidence(P, draws[:3], draws[3:], ratio_threshold=thr)
        g = pf.isLocalizationLost_densitiy_cluster(thr, draws[:3], draws[3:])
        assert g["exact"]
        assert g["ratio"] == o["ratio"] and g["passes"] == o["passes"] and g["best_cluster"] == o["best_cluster"]
        assert np.array_equal(g["centers"], o["centers"])
        assert np.array_equal(pf.clusterAssignments(), o["assignments"])
        assert np.array_equal(g["cluster_weights"], o["cluster_weights"])
        assert np.array_equal(g["best"][:2], o["best"][:2])
        assert (o["best"][2] == -1 and g["best"][2] == -1) or abs(g["best"][2] - o["best"][2]) < 1e-12
        assert g["counts"].sum() == n


def test_kmeans_reinitialises_emptied_cluster():
    rng = np.random.default_rng(5)
    n = 600
    P = clustered_particles(rng, n, [(1.0, 1.0, 0.0), (3.0, 3.0, 1.0)], 0.1)
    P[10] = P[20]
    pf = make_pf(P)
    init, reinit = [10, 20, 300], [5, 77, 123, 9]
    o = pyoracle.kmeans_confidence(P, init, reinit, ratio_threshold=0.2)
    g = pf.isLocalizationLost_densitiy_cluster(0.2, init, reinit)
    assert o["reinit_used"] >= 1 and g["reinit_used"] == o["reinit_used"]
    assert np.array_equal(g["centers"], o["centers"]) and g["ratio"] == o["ratio"]
    assert np.array_equal(pf.clusterAssignments(), o["assignments"])


def kmeans_f64_sums(P, init, max_iters=20):
    """The engine's definition above MCL_KMEANS_EXACT_MAX: fp32 distances, f64 centre sums rounded to fp32, fp32 divide."""
    x, y = P[:, 0], P[:, 1]
    c = np.array([[x[i], y[i]] for i in init], np.float32)
    a = np.zeros(len(P), np.int32)
    for _ in range(max_iters):
        d = np.stack([(x - c[k, 0]) * (x - c[k, 0]) + (y - c[k, 1]) * (y - c[k, 1]) for k in range(3)])
        b = np.argmin(d, axis=0).astype(np.int32)          # first minimum = strict <
        if np.array_equal(a, b):
            break
        a = b
        for k in range(3):
            cnt = int((a == k).sum())
            if cnt:
                c[k, 0] = np.float32(x[a == k].astype(np.float64).sum()) / np.float32(cnt)
                c[k, 1] = np.float32(y[a == k].astype(np.float64).sum()) / np.float32(cnt)
    return a, c


def test_kmeans_confidence_one_million_particles():
    """configs[1] size: parallel f64 centre sums (the reference's sequential fp32 sums carry ~1e-5 relative rounding noise
    at this size, so the oracle is matched to tolerance, the engine's own definition nearly exactly)."""
    rng = np.random.default_rng(12)
    n = 1_000_000
    P = clustered_particles(rng, n, [(0.8, 1.2, 0.0), (3.2, 3.6, 2.0), (4.0, 0.8, -2.0)], 0.12, weights="random")
    pf = make_pf(P, max_particles=n)
    init = [5, 400_000, 900_001]
    g = pf.isLocalizationLost_densitiy_cluster(0.2, init, [])
    assert not g["exact"] and g["counts"].sum() == n
    a, c = kmeans_f64_sums(P, init)
    assert np.abs(g["centers"] - c).max() <= 5e-7 * 5
    assert (pf.clusterAssignments() != a).mean() < 1e-5
    o = pyoracle.kmeans_confidence(P, init, [], ratio_threshold=0.2)
    assert np.abs(g["centers"] - o["centers"]).max() < 2e-3
    assert abs(g["ratio"] - o["ratio"]) < 2e-3
    assert np.allclose(g["best"], o["best"], atol=2e-3)
    assert np.allclose(g["cluster_weights"], o["cluster_weights"], rtol=1e-3)


def test_kmeans_with_engine_draws_and_after_a_filter_step():
    """No injected indices: the draws come from the handle's Philox stream; run on the particles a real filter step left."""
    sc = Scenario(2)
    pf = m.ParticleFilter()
    pf.setMap(sc.occ, RES)
    pf.sampleParticles(1500)
    scan = sc.scans[0]
    pf.updateParticlePos(0.0, 0.01, 0.0)
    pf.computeWeight(scan["ranges"], scan["angle_min"], scan["angle_inc"], scan["range_min"], scan["range_max"])
    pf.resampleParticles(1)
    g = pf.isLocalizationLost_densitiy_cluster(0.05)
    assert 0.0 < g["ratio"] <= 1.0 and g["counts"].sum() == 1500 and 1 <= g["passes"] <= 20
    assert (g["best"] >= 0).all() or (g["best"] == -1).all()
    # replaying the same initial centres through the oracle reproduces it
    P = pf.downloadParticles()
    a = pf.clusterAssignments()
    assert set(np.unique(a)) <= {0, 1, 2}


def test_kmeans_in_ns_mode_uses_materialised_weights():
    sc = Scenario(1)
    s = m.NsShard()
    s.pf.setMap(sc.occ, RES)
    s.pf.sampleParticles(4000)
    scan = sc.scans[0]
    s.pf.updateParticlePos(0.0, 0.01, 0.0)
    s.pf.computeWeight(scan["ranges"], scan["angle_min"], scan["angle_inc"], scan["range_min"], scan["range_max"])
    init = [1, 2000, 3999]
    g = s.pf.isLocalizationLost_densitiy_cluster(0.01, init, [])
    P = s.pf.downloadParticles()
    o = pyoracle.kmeans_confidence(P, init, [], ratio_threshold=0.01)
    assert g["ratio"] == o["ratio"] and np.array_equal(g["centers"], o["centers"]) and np.array_equal(g["cluster_weights"], o["cluster_weights"])


def test_pose_array_matches_publish_particles():
    rng = np.random.default_rng(4)
    n = 10_000
    P = np.zeros((n, 4), np.float32)
    P[:, 0] = rng.uniform(0, 4.8, n); P[:, 1] = rng.uniform(0, 4.8, n); P[:, 2] = rng.uniform(-7, 7, n); P[:, 3] = 1.0 / n
    pf = make_pf(P)
    want = pyoracle.particle_poses(P)
    got = pf.poseArray()
    assert np.array_equal(got[:, :2], want[:, :2]) and np.abs(got[:, 2:] - want[:, 2:]).max() < 3e-16
    sub = pf.poseArray(first=3, stride=17, count=500)
    assert np.array_equal(sub, got[3:3 + 17 * 500:17])
    with pytest.raises(m.MclError):
        pf.poseArray(first=0, stride=2, count=n)
