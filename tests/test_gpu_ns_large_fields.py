"""GPU parity, MCL_MODE_NS, at the field sizes of BASELINE.json configs[3] (4097x4097) and configs[4] (8193x8193): the shapes
whose likelihood field no longer fits shared memory or (as fp32, 8193^2 = 268 MB) L2, where the sensor model addresses the
field with 32-bit window arithmetic and may read it as one-byte codes. The engine's distance transform and field are compared
with the oracle's in full, and the log-likelihoods of a random subsample of particles with NsOracle, once per field form the
engine can pick at that size (fp32 through L1/L2, one-byte coded) and per arithmetic form (packed / scalar). Plus: at
configs[3]'s full 1e8 particles a 4-shard step equals the 1-shard step bit for bit."""
import numpy as np
import pytest

from montecarlolocalisation_b200 import NsShard, ns_step_in_process, synth
from oracle.pyoracle import NsOracle, Scan
from scenario import RES

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("cells,n_beams,seed", [(512, 720, 4), (1024, 1080, 5)])
def test_large_fields_every_form_against_oracle(cells, n_beams, seed):
    occ = synth.maze_occupancy(cells, seed)
    side = cells * 8 + 1
    assert occ.shape == (side, side)
    o = NsOracle()
    o.set_map(occ, RES)
    n = 400_000
    s = NsShard(0, 1, n)
    s.pf.setMap(occ, RES)
    lf_g, d2_g = s.field(occ.shape)
    lf_o, d2_o = o.field()
    assert np.array_equal(d2_g, d2_o), "squared distance transform"
    assert np.array_equal(lf_g, lf_o), "log-likelihood field"
    del lf_g, d2_g, lf_o, d2_o
    s.pf.sampleParticles(n)                                   # uniform over the whole map: gathers all over the field
    P = s.pf.downloadParticles()
    ext = side * 0.1
    P[::9, 0] = np.float32(ext) - P[::9, 0] * np.float32(1e-3)        # a ninth hugging the far edges (beams reach into the border)
    P[4::9, 1] = P[4::9, 1] * np.float32(1e-3)
    P[7::31, 0] += np.float32(ext)                            # and some off the map: the bounds-tested path
    s.pf.uploadParticles(P)
    cx = (cells // 2) * 8 * 0.1 + 0.45
    scan = synth.make_scan(occ, 0.1, (cx, cx, 0.3), n_beams, 11)
    pts = o.beams(Scan(**scan))
    rng = np.random.default_rng(seed)
    idx = np.concatenate([rng.integers(0, n, 6000), np.arange(0, 9 * 300, 9), np.arange(4, 9 * 300, 9), np.arange(7, 31 * 300, 31)])
    want = o.loglik(np.ascontiguousarray(P[idx]), pts)
    forms = set()
    for bits in (0, 16, 16 | 32, 8, 8 | 32, 0):               # engine's choice, fp32 packed / scalar, coded mixed / scalar, engine's choice again
        s.pf.forceSequential(bits)
        mx = s.update_local(scan["ranges"], scan["angle_min"], scan["angle_inc"], scan["range_min"], scan["range_max"])
        ll = s.loglik()
        assert np.array_equal(ll[idx], want), "debug bits %d on %dx%d (%s)" % (bits, side, side, s.field_form())
        assert mx == ll.max()
        forms.add(s.field_form())
    assert forms == {"global-f32", "global-u8"}


def test_configs3_full_size_four_shards_equal_one_shard():
    """BASELINE.json configs[3] as written: 4097x4097 grid, 720 beams, 100,000,000 particles. One step through the
    engine-enqueued path on one shard against the same step on four in-process shards (phase API, collectives in Python):
    ancestors and particles identical, ancestors sorted, and the resampled particle count is exact."""
    occ = synth.maze_occupancy(512, 4)
    n = 100_000_000
    cx = 256 * 8 * 0.1 + 0.45
    scan = synth.make_scan(occ, 0.1, (cx, cx, 0.3), 720, 4000)
    motion = (0.01, 0.02, -0.005)
    one = NsShard(0, 1, n)
    one.pf.setMap(occ, RES)
    one.pf.sampleParticles(n)
    one.step(motion, scan=scan)
    one.pf.synchronize()
    anc1 = one.pf.ancestors()
    assert (np.diff(anc1) >= 0).all() and anc1[0] >= 0 and anc1[-1] < n
    P1 = one.pf.downloadParticles()
    del one
    four = [NsShard(r, 4, n) for r in range(4)]
    for s in four:
        s.pf.setMap(occ, RES)
    for a in four:
        for b in four:
            if a is not b:
                for which in (0, 1, 2, 3):
                    a.peer_set(b.rank, which, b.device_buffer(which))
    for s in four:
        s.pf.sampleParticles(n)
    ns_step_in_process(four, scan, motion)
    off = 0
    for s in four:
        a = s.pf.ancestors()
        assert np.array_equal(a, anc1[off:off + len(a)]), "ancestors of shard %d" % s.rank
        assert np.array_equal(s.pf.downloadParticles(), P1[off:off + len(a)]), "particles of shard %d" % s.rank
        off += len(a)
    assert off == n
