"""torchrun script (one process per GPU): the engine-native sharded NS step (mcl_comm_init + mcl_ns_step: collectives
through peer-memory mailboxes or, with MCL_NS_EXCHANGE=nccl, NCCL on the engine's stream; device-side plan; resampled
particles stored straight into the owning shard over NVLink) against the
single-span NS oracle, bit for bit. torch.distributed (gloo) only carries the NCCL unique id and the final gather.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29631 tests/dist_ns_step_nccl.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    dist.init_process_group("gloo")
    torch.cuda.set_device(local)
    from montecarlolocalisation_b200 import NsShard
    from oracle.pyoracle import NsOracle, Scan
    from scenario import RES, Scenario
    n, steps = 40_003, 3
    sc = Scenario(steps)
    shard = NsShard(rank, world, n, device=local)
    shard.pf.setMap(sc.occ, RES)
    ids = [shard.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    shard.comm_init(ids[0])
    shard.pf.sampleParticles(n)
    poses = []
    for step in range(steps):
        poses.append(shard.step((0.01 * (step + 1), 0.02 + 0.005 * step, -0.015), scan=sc.scans[step], want_pose=(step == steps - 1)))
    shard.pf.synchronize()
    dist.barrier()
    parts = [None] * world
    ancs = [None] * world
    dist.all_gather_object(parts, shard.pf.downloadParticles())
    dist.all_gather_object(ancs, shard.pf.ancestors())
    allpose = [None] * world
    dist.all_gather_object(allpose, poses[-1])
    if rank == 0:
        o = NsOracle()
        o.set_map(sc.occ, RES)
        P = o.init(0, n)
        for step in range(steps):
            Pprev = P
            P, anc, ll, pre = o.step(P, 0, Scan(**sc.scans[step]), (0.01 * (step + 1), 0.02 + 0.005 * step, -0.015), step)
        assert np.array_equal(np.concatenate(parts), P), "particles"
        assert np.array_equal(np.concatenate(ancs).astype(np.int64), anc), "ancestors"
        for p in allpose[1:]:
            assert np.array_equal(p, allpose[0]), "pose differs between ranks"
        assert np.isfinite(allpose[0]).all()
        print("dist_ns_step ok: world %d, %d particles, exchange %s, pose %s" % (world, n, shard.exchange_used(), allpose[0]))
    dist.barrier()
    del shard
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
