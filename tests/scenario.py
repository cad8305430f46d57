"""Shared test scenario = BASELINE.json config 1 in miniature: pink_fundamentals/map.txt, a 360-beam synthetic scan
per step from a ground-truth pose, a wheel-encoder trace. Inputs only; no oracle, no engine."""
import os

import numpy as np

from montecarlolocalisation_b200 import rasterise_map_txt, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RES = np.float32(0.1)


def load_map():
    with open(os.path.join(ROOT, "tests", "golden", "map.txt")) as f:
        return rasterise_map_txt(f.read())


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
