"""Shared test scenario = BASELINE.json config 1 in miniature: pink_fundamentals/map.txt, a 360-beam synthetic scan
per step from a ground-truth pose, a wheel-encoder trace. Inputs only: no oracle, no engine, and the product library is
never loaded from here (the reference arm of bench.py builds its workload with this module). The occupancy grid is the
committed rasterisation of tests/golden/map.txt (tests/golden/map_occ_49x49.npy, SHA-256 9d700e0d...; tests/test_abi.py
checks that the engine's and the oracle's rasterisers both reproduce it)."""
import os

import numpy as np

from montecarlolocalisation_b200 import synth      # numpy-only workload generators; importing it does not dlopen anything

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RES = np.float32(0.1)


def load_map():
    return np.load(os.path.join(ROOT, "tests", "golden", "map_occ_49x49.npy"))


class Scenario:
    def __init__(self, n_steps, n_beams=360, seed=1, kidnap_at=None, start=(0.45, 0.45, np.pi / 2)):
        self.occ = load_map()
        self.n_steps = n_steps
        self.enc_left, self.enc_right = synth.encoder_trace(n_steps)
        self.truth = synth.integrate_odometry(self.enc_left, self.enc_right, start)
        self.scans = []
        for s in range(n_steps):
            pose = self.truth[s]
            if kidnap_at is not None and s >= kidnap_at:
                pose = (3.65, 2.05, -2.0)          # the scan suddenly comes from somewhere else
            self.scans.append(synth.make_scan(self.occ, float(RES), pose, n_beams, seed * 1000 + s))
