"""GPU parity, MCL_MODE_NS: the CUDA engine against the NS oracle (its own CPU restatement; parity unpinned by the
reference, see oracle/mcl_oracle_ns.cpp). NS is defined with IEEE-only arithmetic, so EVERYTHING is required to be
bit-exact: distance transform, field, predicted poses, log-likelihoods, Q32 prefix sums, ancestors, resampled particles,
for one shard and for 2..8 shards."""
import numpy as np
import pytest

import montecarlolocalisation_b200 as m
from montecarlolocalisation_b200 import NsShard, ns_step_in_process, synth
from oracle.pyoracle import NsOracle, Scan
from scenario import RES, Scenario

pytestmark = pytest.mark.gpu


def make_shards(world, n, occ, **cfg):
    shards = [NsShard(r, world, n, **cfg) for r in range(world)]
    for s in shards:
        s.pf.setMap(occ, RES)
    for a in shards:
        for b in shards:
            if a is not b:
                for which in (0, 1, 2, 3):        # particle buffers, ancestors, exchange mailbox
                    a.peer_set(b.rank, which, b.device_buffer(which))
    for s in shards:
        s.pf.sampleParticles(n)
    return shards


def gather(shards):
    return np.concatenate([s.pf.downloadParticles() for s in shards])


def gather_anc(shards):
    return np.concatenate([s.pf.ancestors() for s in shards]).astype(np.int64)


def run(world, n, steps, occ=None, n_beams=360, scenario=None, **cfg):
    sc = scenario or Scenario(steps, n_beams=n_beams)
    occ = sc.occ if occ is None else occ
    o = NsOracle(**{k.replace("ns_", ""): v for k, v in cfg.items() if k in ("ns_temper",)})
    o.set_map(occ, RES)
    shards = make_shards(world, n, occ, **cfg)
    P = o.init(0, n)
    assert np.array_equal(gather(shards), P), "init"
    for step in range(steps):
        motion = (0.01 * (step + 1), 0.02 + 0.005 * step, -0.015)
        scan = sc.scans[step]
        Pp = P.copy()
        o.predict(Pp, 0, *motion, step)
        pts = o.beams(Scan(**scan))
        ll = o.loglik(Pp, pts)
        W, pre, wf, tot = o.weights(ll, float(ll.max()))
        anc = o.resample(pre, o.u0(step))
        # engine, phase by phase so intermediate products can be compared
        for s in shards:
            s.pf.updateParticlePos(*motion)
        assert np.array_equal(gather(shards)[:, :3], Pp[:, :3]), "predict step %d" % step
        maxes = [s.update_local(scan["ranges"], scan["angle_min"], scan["angle_inc"], scan["range_min"], scan["range_max"]) for s in shards]
        assert np.array_equal(np.concatenate([s.loglik() for s in shards]), ll), "loglik step %d" % step
        gmax = max(maxes)
        assert gmax == float(ll.max())
        totals = [s.weights_local(gmax) for s in shards]
        assert sum(totals) == tot
        off = 0
        for s, t in zip(shards, totals):
            b, c, per = m.ns_shard_range(n, world, s.rank)
            assert np.array_equal(s.prefix() + np.uint64(off), pre[b:b + c]), "prefix step %d" % step
            off += t
        assert np.array_equal(gather(shards)[:, 3], wf), "unnormalised weights step %d" % step
        u0 = shards[0].u0()
        assert u0 == o.u0(step)
        off = 0
        slots = []
        for s, t in zip(shards, totals):
            slots.append(s.resample_local(off, tot, u0))
            off += t
        assert slots[0][0] == 0 and slots[-1][1] == n and all(a[1] == b[0] for a, b in zip(slots, slots[1:]))
        for s in shards:
            s.end_step()
        assert np.array_equal(gather_anc(shards), anc), "ancestors step %d" % step
        P = Pp[anc].copy()
        P[:, 3] = np.float32(1.0 / n)
        assert np.array_equal(gather(shards), P), "resampled particles step %d" % step
    return shards, P


def test_field_matches_oracle():
    for occ in (Scenario(1).occ, synth.maze_occupancy(32, 3)):
        o = NsOracle()
        o.set_map(occ, RES)
        s = NsShard()
        s.pf.setMap(occ, RES)
        lf_g, d2_g = s.field(occ.shape)
        lf_o, d2_o = o.field()
        assert np.array_equal(d2_g, d2_o) and np.array_equal(lf_g, lf_o)


def test_single_shard_loop_config1_shape():
    run(1, 5000, 6)


@pytest.mark.parametrize("n_beams", [720, 1080])
def test_single_shard_more_beams(n_beams):
    run(1, 3000, 2, n_beams=n_beams)


@pytest.mark.parametrize("world", [2, 3, 4, 8])
def test_sharded_equals_oracle(world):
    run(world, 4001, 4)


def test_ragged_sizes():
    for n in (1, 33, 2049):
        run(1, n, 2)
    run(2, 7, 2)


def test_field_through_l2_1024_map():
    """1024^2-cell map (4 MiB field: not shared-memory resident, gathered through L2)."""
    occ = synth.maze_occupancy(128, 3)
    assert occ.shape == (1025, 1025)
    pose = (30.45, 40.45, 0.3)
    sc = Scenario(2)
    sc.scans = [synth.make_scan(occ, float(RES), pose, 720, 50 + i) for i in range(2)]
    run(1, 20000, 2, occ=occ, scenario=sc)


def test_field_forms_give_identical_loglik():
    """The sensor model has three homes for the field (fp32 in shared memory, fp32 through L1/L2, one-byte codes through
    L1/L2 + a shared-memory code table) and a scalar and a packed (FFMA2) arithmetic form: identical values from all of
    them and from the oracle, for particles inside the map, near its edge and outside (bounds-tested path)."""
    for occ, pose, n_beams in ((Scenario(1).occ, (2.45, 2.45, 0.3), 360), (synth.maze_occupancy(128, 3), (30.45, 40.45, 0.3), 1080)):
        scan = synth.make_scan(occ, float(RES), pose, n_beams, 7)
        n = 4096 + 37
        (s,) = make_shards(1, n, occ)
        P = s.pf.downloadParticles()
        P[::7, 0] -= 3.0                       # a seventh of the particles shifted, many of them off the map
        P[5::11, 1] += 2.5
        s.pf.uploadParticles(P)
        o = NsOracle(); o.set_map(occ, RES)
        want = o.loglik(P, o.beams(Scan(**scan)))
        for bits in (0, 32, 8, 8 | 32, 16, 16 | 32):
            s.pf.forceSequential(bits)
            s.update_local(scan["ranges"], scan["angle_min"], scan["angle_inc"], scan["range_min"], scan["range_max"])
            assert np.array_equal(s.loglik(), want), "debug bits %d, map %s" % (bits, occ.shape)
        s.pf.forceSequential(0)


def test_one_shard_api_matches_phases():
    """mcl_update + mcl_resample (world == 1 convenience path) give the same particles as the explicit phases."""
    sc = Scenario(3)
    n = 10000
    a = NsShard()
    a.pf.setMap(sc.occ, RES)
    a.pf.sampleParticles(n)
    (b,) = make_shards(1, n, sc.occ)
    for step in range(3):
        scan = sc.scans[step]
        a.pf.updateParticlePos(0.01, 0.02, 0.0)
        a.pf.computeWeight(scan["ranges"], scan["angle_min"], scan["angle_inc"], scan["range_min"], scan["range_max"])
        a.pf.resampleParticles(1)
        ns_step_in_process([b], scan, (0.01, 0.02, 0.0))
        assert np.array_equal(a.pf.downloadParticles(), b.pf.downloadParticles())
        assert np.array_equal(a.pf.ancestors(), b.pf.ancestors())
    pose = a.pf.estimateWeightedPose()
    assert np.isfinite(pose).all()


def test_200k_particles_bit_exact_and_systematic_properties():
    shards, P = run(1, 200_000, 1)
    anc = gather_anc(shards)
    assert (np.diff(anc) >= 0).all()


def test_config3_full_size_properties_and_sharding():
    """BASELINE.json configs[2] at full size: 1025x1025 grid, 10,000,000 particles, 720 beams. Too big for the oracle
    in full, so: (1) a 4-shard run must equal the 1-shard run bit for bit (particles and ancestors); (2) log-likelihoods
    of a random subsample are checked against the oracle; (3) systematic-resampling invariants hold: ancestors sorted,
    every particle's offspring count within 1 of N*W_i/total, zero-weight particles never selected."""
    occ = synth.maze_occupancy(128, 3)
    n = 10_000_000
    pose = (51.25, 51.25, 0.3)
    scans = [synth.make_scan(occ, float(RES), pose, 720, 70 + i) for i in range(2)]
    one = make_shards(1, n, occ)
    four = make_shards(4, n, occ)
    motion = (0.01, 0.02, -0.005)
    o = NsOracle()
    o.set_map(occ, RES)
    rng = np.random.default_rng(3)
    for step in range(2):
        scan = scans[step]
        # phase by phase on the single shard so intermediate products can be inspected
        s = one[0]
        s.pf.updateParticlePos(*motion)
        Pp = s.pf.downloadParticles()
        mx = s.update_local(scan["ranges"], scan["angle_min"], scan["angle_inc"], scan["range_min"], scan["range_max"])
        ll = s.loglik()
        idx = rng.integers(0, n, 3000)
        assert np.array_equal(o.loglik(Pp[idx], o.beams(Scan(**scan))), ll[idx]), "log-likelihood subsample vs oracle"
        assert mx == ll.max()
        tot = s.weights_local(mx)
        pre = s.prefix()
        assert int(pre[-1]) == tot and (np.diff(pre.astype(np.int64)) >= 0).all()
        s.resample_local(0, tot, s.u0())
        s.end_step()
        anc = s.pf.ancestors().astype(np.int64)
        assert (np.diff(anc) >= 0).all()
        W = np.diff(np.concatenate([[0], pre.astype(np.int64)]))
        counts = np.bincount(anc, minlength=n)
        assert (counts[W == 0] == 0).all()
        expect = W.astype(np.float64) * (n / float(tot))
        assert np.abs(counts - expect).max() < 1.0 + 1e-6
        ns_step_in_process(four, scan, motion)
        assert np.array_equal(gather_anc(four), anc), "4 shards vs 1 shard: ancestors"
        assert np.array_equal(gather(four), s.pf.downloadParticles()), "4 shards vs 1 shard: particles"


def test_no_valid_beams_and_particles_off_the_map():
    """Every reading invalid -> zero beams -> all log-likelihoods 0 -> uniform weights -> ancestor(k) = k."""
    sc = Scenario(1)
    n = 1000
    (s,) = make_shards(1, n, sc.occ)
    bad = np.full(360, np.nan, np.float32)
    mx = s.update_local(bad, -np.pi, 2 * np.pi / 360, 0.02, 5.6)
    assert mx == 0.0 and (s.loglik() == 0).all()
    tot = s.weights_local(mx)
    assert tot == n << 32
    s.resample_local(0, tot, 12345)
    s.end_step()
    assert np.array_equal(s.pf.ancestors(), np.arange(n))
    # particles far outside the grid score the field's floor on every beam, identically in engine and oracle
    o = NsOracle(); o.set_map(sc.occ, RES)
    P = s.pf.downloadParticles()
    P[:, 0] += 100.0
    s.pf.uploadParticles(P)
    scan = sc.scans[0]
    s.update_local(scan["ranges"], scan["angle_min"], scan["angle_inc"], scan["range_min"], scan["range_max"])
    assert np.array_equal(s.loglik(), o.loglik(P, o.beams(Scan(**scan))))


def test_engine_native_step_matches_phases():
    """mcl_ns_step / mcl_ns_step_staged (whole step enqueued by the engine, plan computed on the device, no host round
    trip) produce the same particles and ancestors as the phase-by-phase API, and the pose they return is the weighted
    mean of the particles BEFORE resampling."""
    sc = Scenario(4)
    n = 30011
    (a,) = make_shards(1, n, sc.occ)
    (b,) = make_shards(1, n, sc.occ)
    o = NsOracle(); o.set_map(sc.occ, RES)
    for i, scan in enumerate(sc.scans):
        a.pf.stageScan(i, scan["ranges"], scan["angle_min"], scan["angle_inc"], scan["range_min"], scan["range_max"])
    for step in range(4):
        motion = (0.01 * (step + 1), 0.02, -0.005)
        scan = sc.scans[step]
        # expected pose from the oracle (weights of the predicted particles under this scan) ...
        Pp = b.pf.downloadParticles()
        o.predict(Pp, 0, *motion, step)
        ll_o = o.loglik(Pp, o.beams(Scan(**scan)))
        oracle_pose = o.pose(Pp, o.weights(ll_o, float(ll_o.max()))[2])
        # ... and from the phases
        b.pf.updateParticlePos(*motion)
        mx = b.update_local(scan["ranges"], scan["angle_min"], scan["angle_inc"], scan["range_min"], scan["range_max"])
        tot = b.weights_local(mx)
        pp = b.pose_partials()
        expect_pose = np.array([pp[1] / pp[0], pp[2] / pp[0], np.arctan2(pp[3], pp[4])])
        b.resample_local(0, tot, b.u0())
        b.end_step()
        if step % 2 == 0:
            pose = a.step(motion, scan=scan, want_pose=True)
        else:
            pose = a.step(motion, slot=step, want_pose=True)
        assert np.allclose(pose, expect_pose, rtol=1e-12, atol=1e-12), (pose, expect_pose)
        assert np.allclose(pose, oracle_pose, rtol=1e-5, atol=1e-5), (pose, oracle_pose)      # NS-8: tolerance-graded (1e-5)
        assert np.array_equal(a.pf.downloadParticles(), b.pf.downloadParticles()), "particles step %d" % step
        assert np.array_equal(a.pf.ancestors(), b.pf.ancestors()), "ancestors step %d" % step
    # without a pose the call does not wait for the GPU; results are the same
    a.step((0.01, 0.02, 0.0), slot=0)
    ns_step_in_process([b], sc.scans[0], (0.01, 0.02, 0.0))
    assert np.array_equal(a.pf.downloadParticles(), b.pf.downloadParticles())


@pytest.mark.parametrize("world", [2, 4])
def test_engine_native_step_peer_memory_exchange_in_process(world):
    """mcl_ns_step of a sharded filter with the collectives done through peer-memory mailboxes (no NCCL): `world` shards
    of one process on one GPU, each on its own stream, every step enqueued without a host round trip. Same particles and
    ancestors as the phase-by-phase path whose collectives are plain Python."""
    sc = Scenario(3)
    n = 20011
    a = make_shards(world, n, sc.occ)
    b = make_shards(world, n, sc.occ)
    for s in a:
        s.set_exchange("peer")
        for i, scan in enumerate(sc.scans):     # scans parked on the device: enqueueing a step then never waits for the GPU,
            s.pf.stageScan(i, scan["ranges"], scan["angle_min"], scan["angle_inc"], scan["range_min"], scan["range_max"])
    for step in range(3):                       # which shards sharing one process (and one host thread) depend on
        motion = (0.01 * (step + 1), 0.02, -0.005)
        for s in a:
            s.step(motion, slot=step)
        ns_step_in_process(b, sc.scans[step], motion)
    for s in a:
        s.pf.synchronize()
        assert s.exchange_used() == "peer"
    assert np.array_equal(gather(a), gather(b))
    assert np.array_equal(gather_anc(a), gather_anc(b))


@pytest.mark.parametrize("exchange", ["peer", "nccl"])
def test_engine_native_step_multi_gpu(exchange):
    """2+ GPUs, one process per GPU: the engine-enqueued sharded step - collectives through peer-memory mailboxes, or
    through the engine's NCCL communicator - + peer stores (tests/dist_ns_step_nccl.py) equals the single-span oracle bit
    for bit. Skipped on a 1-GPU box."""
    import subprocess
    import sys
    import os
    import torch
    g = torch.cuda.device_count()
    if g < 2:
        pytest.skip("needs >= 2 GPUs")
    world = min(g, 4)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, MCL_NS_EXCHANGE=exchange)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
                        "--master-port", "29631", os.path.join(root, "tests", "dist_ns_step_nccl.py")], capture_output=True, text=True, timeout=600,
                       env=env)
    assert r.returncode == 0 and "dist_ns_step ok" in r.stdout and "exchange %s" % exchange in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
