"""Where the one-kernel exact scan spends its time: per-tile stage stamps. python tools/xs_trace.py [n] [real]
real: the weights of a configs[1] filter after 12 ticks (bench.py's workload) instead of random ones."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import montecarlolocalisation_b200 as m

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
rng = np.random.default_rng(1)
w = (40.0 * rng.random(n)).astype(np.float32)
pf = m.ParticleFilter(max_particles=n)
if len(sys.argv) > 2 and sys.argv[2] == "real":
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import bench
    from scenario import RES
    sc = bench.workload(16)
    pf.setMap(sc.occ, RES)
    pf.sampleParticles(n)
    for s_ in range(12):
        sca = sc.scans[s_]
        pf.executeParticleFilter(sc.enc_left[s_], sc.enc_right[s_], 1, scan=sca)
    sca = sc.scans[12]
    pf.diffDriveModel(sc.enc_left[12], sc.enc_right[12])
    pf.computeWeight(sca["ranges"], sca["angle_min"], sca["angle_inc"], sca["range_min"], sca["range_max"])
    w = np.ascontiguousarray(pf.downloadParticles()[:, 3])
    print("real weights: min %g max %g sum %g" % (w.min(), w.max(), w.astype(np.float64).sum()))
for rep in range(3):
    tr = pf.exactScanTrace(w).astype(np.int64)
t0 = tr[:, 0].min()
rel = (tr[:, :7] - t0) / 1e3
names = ["start", "sum published", "edges known", "summary published", "lower summaries scanned", "SEQ walked", "written"]
print("tiles %d, kernel span %.1f us" % (len(tr), rel[:, 6].max()))
for k, nm in enumerate(names):
    print("  %-26s min %7.1f  median %7.1f  max %7.1f us" % (nm, rel[:, k].min(), np.median(rel[:, k]), rel[:, k].max()))
d = np.diff(rel, axis=1)
for k in range(6):
    print("  stage %d->%d duration: median %6.2f max %6.2f us" % (k, k + 1, np.median(d[:, k]), d[:, k].max()))
order = np.argsort(tr[:, 0])
print("  start times of tiles by ticket (every 32nd):", np.round(rel[::32, 0], 1))
print("  end times of tiles by ticket (every 32nd):  ", np.round(rel[::32, 6], 1))
slow = np.argsort(-d[:, 2])[:12]
print("  slowest stage 2->3 tiles (tile: start, sum, edges, summary, scanned, walked, written):")
for k in sorted(slow):
    print("   tile %4d: " % k + " ".join("%6.1f" % v for v in rel[k]))
print("  tiles 0..7:")
for k in range(min(8, len(rel))):
    print("   tile %4d: " % k + " ".join("%6.1f" % v for v in rel[k]))
print("  inside stage 3 (summary published, own poll done, scan done, fence done, blocks fetched), tiles every 24th:")
for k in range(0, len(tr), 24):
    a = (tr[k, [3, 7, 8, 9, 4]] - t0) / 1e3
    print("   tile %4d: " % k + " ".join("%6.1f" % v for v in a))
print("  inside stage 2 of the slow tiles (edges known, non-fast branch entered, aggregates scanned, block written, fence done, summary published):")
for k in sorted(slow):
    a = (tr[k, [2, 10, 11, 12, 13, 3]] - t0) / 1e3
    print("   tile %4d: " % k + " ".join("%6.1f" % v for v in a))
print("  the summary each tile saw last (tile: its scan done at; last summary seen at, of tile, which published at):")
for k in range(8, len(tr), 16):
    src = int(tr[k, 15])
    print("   tile %4d: scan done %6.1f; last seen %6.1f of tile %4d published %6.1f" % (k, (tr[k, 8] - t0) / 1e3, (tr[k, 14] - t0) / 1e3, src, (tr[src, 3] - t0) / 1e3))
