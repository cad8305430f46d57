"""Generates tests/golden/ref_next_rows.npz from the compiled reference (oracle/_ref): run where /root/reference exists.
    python tests/golden/make_golden_next.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import pyoracle
from oracle.pyoracle import Ref
from test_oracle_next_rows import clustered_particles

pyoracle.build()
rng = np.random.default_rng(2024)
n = 1500                                   # the reference's own particle count (MC:84)
P = clustered_particles(rng, n, [(0.4, 0.4, 0.0), (2.8, 3.6, 2.0), (4.0, 1.2, -2.5)], 0.08, weights="random")
seed, thr = 4242, 0.3
r = Ref()
r.set_time(seed)
ratio, best = r.kmeans_confidence(P, ratio_threshold=thr)
r.set_time(seed)
a, c = r.kmeans(P)
draws = np.array([v % n for v in pyoracle.libc_rand_sequence(seed, 64)], np.int32)
poses = np.array([(rng.uniform(-0.2, 4.8), rng.uniform(-0.2, 4.8), rng.uniform(-10, 10)) for _ in range(200)])
cells = np.array([r.publish_pos_msg(*p) for p in poses], np.int32)
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "ref_next_rows.npz"), P=P, draws=draws, threshold=thr, ratio=ratio, best=best,
                    assignments=a, centers=c, poses=poses, cells=cells, pose_array=r.publish_particles(P[:256]))
print("ratio", ratio, "best", best)
