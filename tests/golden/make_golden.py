"""Generates tests/golden/ref_config1.npz by running the UNMODIFIED reference (oracle/_ref/libmclref.so, built from
/root/reference/pink_fundamentals/src/monte_carlo.cpp behind stub ROS/tf/Eigen headers) on the config-1 scenario.
Only runs where /root/reference exists; the fixture it writes is what travels.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle.pyoracle import Ref, Scan, build  # noqa: E402
from scenario import RES, Scenario  # noqa: E402

N = 1000
STEPS = 8
JITTER = [1, 1, 1, 0, 0, 1, 0, 0]
KIDNAP_AT = 4


def main():
    build()
    sc = Scenario(STEPS, kidnap_at=KIDNAP_AT)
    r = Ref()
    r.set_map(sc.occ, RES)
    r.precompute_ray_directions(-120.0, 120.0, 0.1)
    n_rows, n_cols = sc.occ.shape[0] // 8, sc.occ.shape[1] // 8
    out = dict(occ=sc.occ, n=N, steps=STEPS, jitter=np.array(JITTER), enc_left=sc.enc_left, enc_right=sc.enc_right)
    # initial particle set: sampleParticles(N) with mt19937(4242)
    r.clear_seeds(); r.push_seeds(4242)
    P = r.sample_particles(N)
    out["init_draws"] = np.stack([v.astype(np.float64) for v in r.named_sample_draws(4242, n_rows, n_cols, N).values()])
    out["P0"] = P.copy()
    SEED_SAMPLE, MAXINJ = 99, 200
    r.seed_static_engines(SEED_SAMPLE, 1)
    z_all = r.stream_minstd_normal(SEED_SAMPLE, 3 * STEPS)
    out["z"] = z_all
    for s in range(STEPS):
        scan = Scan(**sc.scans[s])
        r.set_scan(scan)
        for k in ("ranges", "angle_min", "angle_inc", "range_min", "range_max"):
            out["scan%d_%s" % (s, k)] = np.asarray(sc.scans[s][k])
        motion = r.diff_drive(sc.enc_left[s], sc.enc_right[s])
        out["motion%d" % s] = motion
        r.update_particle_pos(P)
        out["pred%d" % s] = P.copy()
        seed_r, seed_j = 7000 + s, 8000 + s
        inj_seeds = [9000 + 300 * s + i for i in range(MAXINJ)]
        r.clear_seeds(); r.push_seeds(seed_r, *inj_seeds)
        # re-seed only uniformJitter's engine; `sample`'s engine keeps running across steps like in the node
        r.L.ref_seed_static_engine(1, seed_j)
        Pw = P.copy()
        Pnew, injected = r.resample(Pw, JITTER[s])
        out["weights%d" % s] = Pw[:, 3].copy()            # normalised weights (MC:497-503)
        out["new%d" % s] = Pnew.copy()
        out["injected%d" % s] = injected
        out["u_r%d" % s] = r.stream_mt_canonical(seed_r, N)
        out["u_jit%d" % s] = r.stream_minstd_canonical(seed_j, 3 * N)
        inj = [r.named_sample_draws(sd, n_rows, n_cols, 1) for sd in inj_seeds]
        out["inj%d" % s] = np.stack([np.array([d[k][0] for d in inj], np.float64) for k in ("u_yaw", "row", "col", "u_dx", "u_dy")])
        out["inj_state%d" % s] = r.injection_state()
        out["pose%d" % s] = r.estimate_weighted_pose(Pnew)
        P = Pnew
    k, dx, dy = r.ray_lut(-400, 400)
    out["lut_keys"], out["lut_dx"], out["lut_dy"] = k, dx, dy
    path = os.path.join(ROOT, "tests", "golden", "ref_config1.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes; injected per step:", [int(out["injected%d" % s]) for s in range(STEPS)])


if __name__ == "__main__":
    main()
