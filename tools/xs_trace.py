"""Where the one-kernel exact scan spends its time: per-tile stage stamps. python tools/xs_trace.py [n]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import montecarlolocalisation_b200 as m

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
rng = np.random.default_rng(1)
w = (40.0 * rng.random(n)).astype(np.float32)
pf = m.ParticleFilter()
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
