"""Where the REF tick's end-to-end time goes beyond the device time (configs[1], 1M particles by default):
    a  executeParticleFilter(scan=host scan): host wall per call (what bench.py's e2e reports, here without the L2 flush)
    b  executeParticleFilter(slot=staged) + result read back: no scan filtering / H2D in the call
    c  the same staged tick queued back to back, CUDA events around the batch (bench.py's `value`)
    python tools/e2e_breakdown.py [particles] [steps]
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import torch  # noqa: E402

import bench  # noqa: E402
import montecarlolocalisation_b200 as m  # noqa: E402
from scenario import RES  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
K = int(sys.argv[2]) if len(sys.argv) > 2 else 40
W = 5
sc = bench.workload(K + W)
pinned = [torch.from_numpy(np.ascontiguousarray(sc.scans[s]["ranges"])).pin_memory() for s in range(K + W)]


def fresh():
    """The same filter for every mode: same seed, same ticks."""
    global pf, stream
    pf = m.ParticleFilter(device=0, max_particles=n, seed=0x1234)
    pf.setMap(sc.occ, RES)
    pf.sampleParticles(n)
    stream = torch.cuda.ExternalStream(pf.stream(), device=0)
    for s in range(K + W):
        sca = sc.scans[s]
        pf.stageScan(s, sca["ranges"], sca["angle_min"], sca["angle_inc"], sca["range_min"], sca["range_max"])



def a(s):
    sca = dict(sc.scans[s]); sca["ranges"] = pinned[s].numpy()
    return pf.executeParticleFilter(sc.enc_left[s], sc.enc_right[s], 1, scan=sca)[0]


def b(s):
    return pf.executeParticleFilter(sc.enc_left[s], sc.enc_right[s], 1, slot=s)[0]


def cq(s):
    return pf.executeParticleFilter(sc.enc_left[s], sc.enc_right[s], 1, slot=s, want_result=False)


def wall(fn, first):
    for s in range(first, first + W):
        fn(s)
    out = []
    for s in range(first + W, first + W + K):
        stream.synchronize()
        t0 = time.perf_counter()
        fn(s)
        out.append(time.perf_counter() - t0)
    return np.array(out) * 1e6


fresh()
wa = wall(a, 0)
fresh()
wb = wall(b, 0)
fresh()
for s in range(W):
    cq(s)
stream.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
with torch.cuda.stream(stream):
    e0.record()
    for s in range(W, K + W):
        cq(s)
    e1.record()
stream.synchronize()
dev = e0.elapsed_time(e1) * 1e3 / K
print("particles %d  steps %d" % (n, K))
print("a  host scan, result back : median %.1f us  min %.1f" % (np.median(wa), wa.min()))
print("b  staged scan, result back: median %.1f us  min %.1f" % (np.median(wb), wb.min()))
print("c  staged, queued (device) : %.1f us per tick" % dev)
fresh()
for s in range(W):
    cq(s)
pf.profileEnable(True)
for s in range(W, K + W):
    cq(s)
prof = pf.profileRead()
pf.profileEnable(False)
print("d  per kernel (events around every launch, serialised, ~4 us high each): " + "  ".join("%s %.1f" % (k.replace("k_ref_", "").replace("k_", ""), 1e3 * v[0] / v[1]) for k, v in sorted(prof.items(), key=lambda kv: -kv[1][0])))
