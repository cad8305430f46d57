"""A/B of the REF tick under mcl_debug_force_sequential switches: python tools/tick_ab.py [particles] [bits ...]
Each configuration: 300 staged ticks queued back to back (CUDA events around the batch), then 40 ticks with an L2 flush before each."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import torch  # noqa: E402

import bench  # noqa: E402
import montecarlolocalisation_b200 as m  # noqa: E402
from scenario import RES  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
configs = [int(a) for a in sys.argv[2:]] or [0, 256, 512, 768]
T = 64
sc = bench.workload(T)
junk = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for rep in range(2):
    for bits in configs:
        pf = m.ParticleFilter(device=0, max_particles=n, seed=0x1234)
        pf.setMap(sc.occ, RES)
        pf.sampleParticles(n)
        pf.forceSequential(bits)
        stream = torch.cuda.ExternalStream(pf.stream(), device=0)
        for s in range(T):
            sca = sc.scans[s]
            pf.stageScan(s, sca["ranges"], sca["angle_min"], sca["angle_inc"], sca["range_min"], sca["range_max"])

        def tick(i):
            s = i % T
            pf.executeParticleFilter(sc.enc_left[s], sc.enc_right[s], 1, slot=s, want_result=False)

        for i in range(30):
            tick(i)
        stream.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = pf.kernelLaunches()
        with torch.cuda.stream(stream):
            e0.record()
            for i in range(30, 330):
                tick(i)
            e1.record()
        stream.synchronize()
        hot = e0.elapsed_time(e1) * 1e3 / 300
        launches = (pf.kernelLaunches() - l0) / 300
        cold = []
        with torch.cuda.stream(stream):
            for i in range(330, 370):
                junk.fill_(i & 255)
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(); tick(i); b.record()
                cold.append((a, b))
        stream.synchronize()
        cold = np.median([a.elapsed_time(b) for a, b in cold]) * 1e3
        print("bits %4d: %.1f us/tick queued (%.1f launches), %.1f us/tick after an L2 flush (median)" % (bits, hot, launches, cold))
        pf.close()
