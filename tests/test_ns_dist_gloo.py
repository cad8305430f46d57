"""world_size-2 (and 3) CPU test of the sharded NS step over torch.distributed/gloo: every rank runs the oracle on its
shard, the three collectives go through gloo, the slot plan comes from the product's host code (mcl_ns_first_slot /
mcl_ns_shard_range). The gathered result must equal the single-span oracle bit for bit."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n, out_path):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import montecarlolocalisation_b200 as m
    from oracle.pyoracle import NsOracle, Scan
    from scenario import RES, Scenario
    sc = Scenario(2)
    o = NsOracle()
    o.set_map(sc.occ, RES)
    b, c, per = m.ns_shard_range(n, world, rank)
    P = o.init(b, c)
    result = []
    for step in range(2):
        motion = (0.02 * (step + 1), 0.03, -0.01)
        scan = Scan(**sc.scans[step])
        o.predict(P, b, *motion, step)
        ll = o.loglik(P, o.beams(scan))
        mx = torch.tensor([float(ll.max())], dtype=torch.float32)
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)                       # collective 1
        W, pre, wf, t = o.weights(ll, float(mx))
        tot = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(tot, torch.tensor([t], dtype=torch.int64))      # collective 2
        totals = [int(x) for x in tot]
        total, off = sum(totals), sum(totals[:rank])
        u0 = o.u0(step)
        lo, hi = m.ns_first_slot(off, total, n, u0), m.ns_first_slot(off + t, total, n, u0)
        out = np.zeros((hi - lo, 5), np.float64)
        for j, k in enumerate(range(lo, hi)):
            thr = ((k << 32) + u0) * total // (n << 32)
            i = int(np.searchsorted(pre.astype(object) + off, thr, side="right"))
            out[j] = (k, P[i, 0], P[i, 1], P[i, 2], b + i)
        # the "P2P rebalance": every produced particle goes to the shard that owns its slot (all_gather on CPU)
        gathered = [None] * world
        dist.all_gather_object(gathered, out)
        allp = np.concatenate(gathered)
        allp = allp[np.argsort(allp[:, 0])]
        assert len(allp) == n and np.array_equal(allp[:, 0], np.arange(n))
        mine = allp[b:b + c]
        P = np.zeros((c, 4), np.float32)
        P[:, :3] = mine[:, 1:4].astype(np.float32)
        P[:, 3] = np.float32(1.0 / n)
        result.append(allp)
        dist.barrier()                                                  # collective 3
    if rank == 0:
        np.save(out_path, np.stack(result))
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_gloo_sharded_ns_step(world, tmp_path):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle.pyoracle import NsOracle, Scan
    from scenario import RES, Scenario
    n = 1501
    out_path = str(tmp_path / "gathered.npy")
    port = 29500 + (os.getpid() % 2000) + world
    mp.spawn(_worker, args=(world, port, n, out_path), nprocs=world, join=True)
    got = np.load(out_path)
    sc = Scenario(2)
    o = NsOracle()
    o.set_map(sc.occ, RES)
    P = o.init(0, n)
    for step in range(2):
        P, anc, ll, pre = o.step(P, 0, Scan(**sc.scans[step]), (0.02 * (step + 1), 0.03, -0.01), step)
        assert np.array_equal(got[step][:, 4].astype(np.int64), anc)
        assert np.array_equal(got[step][:, 1:4].astype(np.float32), P[:, :3])
