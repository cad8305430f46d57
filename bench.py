#!/usr/bin/env python
"""bench.py — particle-filter hot path on B200 (and the reference's CPU filter beside it).

    python bench.py --gpus 1 --steps 10 --warmup 3                     # our arm, one GPU
    torchrun --nproc-per-node N ... bench.py --gpus N ...              # our arm, N GPUs of one node (one rank per GPU)
    python bench.py --impl reference --gpus N --steps 3 --warmup 1     # reference arm: the reference's own CPU filter

The metric is BASELINE.json's: particle-beam likelihood evaluations per second over the full predict -> update -> resample ->
estimate loop. One "step" = one loop iteration; one "eval" = one (particle, scored beam) pair.

--gpus 1   top-level line = BASELINE.json configs[1]: pink_fundamentals/map.txt (49x49 grid), 1,000,000 particles, 360-beam
           synthetic scan + wheel-encoder trace, reference-parity mode (MCL_MODE_REF: results identical to the reference for the
           same draws; the reference scores every 20th beam inside (-120,120) degrees = 12 of 360). Sub-records: the reference's
           own operating point (1000 / 1500 particles, 360-beam and 683-beam scans) on GPU and CPU in the same run, and the
           north-star (MCL_MODE_NS) legs: configs[1] shape, configs[2] (1025^2 grid, 10 M particles, 720 beams), configs[3] /
           configs[4] per-GPU shapes, and configs[3] at its full 1e8 particles on this one GPU (the N = 1 point of the
           strong-scaling curve the --gpus N > 1 lines continue).
--gpus N>1 top-level line = BASELINE.json configs[3]: 4097x4097 grid, 720 beams, 1e8 particles IN TOTAL sharded across the N
           GPUs (strong scaling), global systematic resampling, the sharded step of mcl_ns_step. Sub-records: the weak-scaling
           legs (12.5 M particles per GPU on the configs[3] / configs[4] grids, 1 M per GPU on map.txt) and N independent
           REF replicas.
Every run starts with parity gates (a timing only counts behind a green gate): the sharded NS step
against the CPU oracle (oracle/mcl_oracle_ns.cpp) with a 64-bit hash of particles + ancestors that is the same for every N,
and at N = 1 the REF loop at 1000 particles against oracle/mcl_oracle.cpp.

`value` times the loop with the scans already parked in HBM (CUDA events on the engine's stream); `e2e` times the same loop
through the public C-ABI call with the scan in pinned host memory and the pose read back every step, on the HOST's wall clock.
"""
import argparse
import gc
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

N_PARTICLES = 1_000_000
N_BEAMS = 360
NS_STRONG_PARTICLES = 100_000_000
NS_STRONG_CELLS = 512                  # 512 maze cells of 8 px -> 4097 x 4097 grid
NS_STRONG_BEAMS = 720
METRIC = "particle-beam likelihood evals/sec (full predict/update/resample/estimate loop)"
UNIT = "evals/s"

# algorithmic bytes per particle per launch (SURVEY.md §8d / DESIGN.md "kernels"); update adds 1 B per map probe
ALGO_BYTES = {
    "k_ref_predict": 32, "k_ref_update": 20, "k_ref_update_v2": 20, "k_ref_first_touch": 16, "k_ref_seq_total": 4, "k_ref_seq_cdf": 16,
    "k_ref_resample": 44 + 32, "k_fill_resample_draws": 32, "k_pose_wsum": 16, "k_pose_sums": 16,
    "k_xs_total": 4, "k_xs_cdf": 12, "k_ref_guide": 8,
}


def counters_for(kernel, key=""):
    """What binds `kernel` according to the committed ncu capture of it: {"bound", "frac", "counter", "file", "traffic"}
    from profiles/r2_counters.json (written by tools/ncu_counters.py from `ncu --set full` captures), else None."""
    p = os.path.join(ROOT, "profiles", "r2_counters.json")
    if not os.path.exists(p):
        return None
    t = json.load(open(p))
    return t.get(kernel + "@" + key) or t.get(kernel)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p))["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def roofline_record(kernels, key, particles):
    """The dominant kernel (largest measured share of the step) against the resource that binds it. `achieved` / `peak` /
    `hbm_frac` are always the algorithmic HBM bytes per launch over the live-measured duration against the measured HBM
    peak; `bound` and `frac` name the binding resource: for a streaming kernel that is HBM and frac = hbm_frac; for an
    issue- or pipe-bound kernel frac is the ncu counter named in `frac_counter` from the committed capture `frac_file`."""
    peak, peak_src = peaks()
    top = next(iter(kernels))
    k = kernels[top]
    c = counters_for(top, key) or {}
    hbm_frac = (k["gbs"] / peak) if k.get("gbs") else None
    bound = c.get("bound", "hbm")
    return {"kernel": top, "bound": bound, "achieved": k.get("gbs"), "peak": peak, "unit": "GB/s", "hbm_frac": hbm_frac,
            "frac": hbm_frac if bound == "hbm" else c.get("frac"), "frac_counter": c.get("counter", "algorithmic bytes / live duration / HBM peak"),
            "frac_file": c.get("file"), "traffic": c.get("traffic") if c.get("particles") in (None, particles) else None,
            "peak_source": peak_src, "share_of_step": k["share"], "note": c.get("note")}


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, devices):
        """devices: the GPU indices of this job. ONE sampler per job (rank 0): eight nvidia-smi pollers would contend with
        the filters' own driver calls."""
        self.devices = list(devices)
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", ",".join(str(d) for d in self.devices), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        per_gpu = {}
        for r in self.rows:
            if len(r) >= 9 and r[1].replace(".", "").isdigit():
                per_gpu.setdefault(r[0], []).append(float(r[1]))
        sm = [statistics.median(v) for v in per_gpu.values()]           # the slowest GPU's median is what is reported
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) < 9:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": min(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": sum(len(v) for v in per_gpu.values()), "gpus_sampled": len(per_gpu)}


def workload(n_steps):
    from scenario import Scenario
    return Scenario(n_steps, n_beams=N_BEAMS, seed=1)


def used_beams(scan):
    """Beams the reference scores: every 20th of those strictly inside (-120, 120) degrees that pass the range filter."""
    r = scan["ranges"].astype(np.float64)
    ang = np.float64(scan["angle_min"]) + np.arange(len(r)) * np.float64(scan["angle_inc"])
    keep = (np.isnan(r) | np.isinf(r)) | ((r >= scan["range_min"]) & (r <= scan["range_max"]))
    deg = ang * 180.0 / np.pi
    keep &= (deg > -120.0) & (deg < 120.0)
    return len(range(0, int(keep.sum()), 20))


def ns_workload(cells, n_beams, n_scans, seed):
    from montecarlolocalisation_b200 import synth
    occ = synth.maze_occupancy(cells, seed)
    res = 0.1
    # a pose in the middle of a cell near the map centre
    cx = (cells // 2) * 8 * res + 0.45
    pose = (cx, cx, 0.3)
    scans = [synth.make_scan(occ, res, pose, n_beams, 1000 * seed + i) for i in range(n_scans)]
    return occ, scans


def robot_scan(occ, pose, seed):
    """The robot's own LIDAR geometry (comment MC:638-640): 683 beams, 0.352 degrees apart, from -120 degrees: 35 scored."""
    from montecarlolocalisation_b200 import synth
    return synth.make_scan(occ, 0.1, pose, 683, seed, angle_min=np.float32(-120.0 * np.pi / 180.0), angle_inc=np.float32(0.352 * np.pi / 180.0))


def config_ref(n):
    return {"workload": "BASELINE.json configs[1]: pink_fundamentals/map.txt 49x49 grid @0.1 m, %d particles, %d-beam synthetic scan "
                        "(12 beams scored per particle: every 20th inside +-120 deg), full predict/update/resample/estimate loop" % (n, N_BEAMS),
            "particles": n, "beams": N_BEAMS, "map": "map.txt 49x49",
            "l2": "L2 flushed between timed steps (256 MiB write, outside the timed region)"}


def config_ns_strong(n_global):
    return {"workload": "BASELINE.json configs[3]: 4097x4097 synthetic maze occupancy grid (seed 4) @0.1 m, %d particles IN TOTAL sharded "
                        "across the GPUs of the job (strong scaling), %d-beam synthetic scan, full predict/update/resample/estimate loop "
                        "with global systematic resampling" % (n_global, NS_STRONG_BEAMS),
            "particles": n_global, "beams": NS_STRONG_BEAMS, "map": "maze 4097x4097 seed 4",
            "l2": "inputs larger than L2: the per-GPU working set (48 B per particle) exceeds the 126 MB L2 at every N"}


# ------------------------------------------------------------------------------------------------------------------
# reference arm: the reference's own CPU filter (oracle/_ref = its unmodified translation unit; else the oracle port)
# ------------------------------------------------------------------------------------------------------------------
def run_cpu_steps(n_particles, steps, warmup, occ, scans, enc_left, enc_right):
    """Times full filter steps of the reference on 1 host thread (it is single-threaded, MC:1212). Returns
    (seconds per timed step list, evals per step list, kind)."""
    from oracle import pyoracle
    from oracle.pyoracle import Oracle, Scan
    RES = np.float32(0.1)
    pyoracle.build()
    kind = "reference" if pyoracle.ref_available() else "port"
    rng = np.random.default_rng(123)
    times, evals = [], []
    ns = len(scans)
    if kind == "reference":
        ref = pyoracle.Ref()
        ref.set_map(occ, RES)
        ref.precompute_ray_directions(-120.0, 120.0, 0.1)
        ref.clear_seeds()
        ref.push_seeds(*[int(x) for x in rng.integers(1, 2**31 - 1, 64)])
        ref.seed_static_engines(11, 12)
        P = ref.sample_particles(n_particles)
        for s in range(warmup + steps):
            ref.set_scan(Scan(**scans[s % ns]))
            ref.push_seeds(*[int(x) for x in rng.integers(1, 2**31 - 1, 256)])
            t0 = time.perf_counter()
            ref.diff_drive(enc_left[s], enc_right[s])                # diffDriveModel          (MC:1084)
            ref.update_particle_pos(P)                               # updateParticlePos       (MC:1086)
            P, _ = ref.resample(P, 1)                                # resampleParticles       (MC:1089) incl. computeWeight
            ref.estimate_weighted_pose(P)                            # estimateWeightedPose    (MC:782)
            dt = time.perf_counter() - t0
            if s >= warmup:
                times.append(dt)
                evals.append(n_particles * used_beams(scans[s % ns]))
    else:
        o = Oracle(trig_mode=0)
        o.set_map(occ, RES)
        o.precompute_ray_directions(-120.0, 120.0, 0.1)
        n = n_particles
        n_rows, n_cols = o.cell_ranges()
        P = o.sample_particles(rng.random(n), rng.integers(0, n_rows, n), rng.integers(0, n_cols, n), rng.random(n), rng.random(n))
        for s in range(warmup + steps):
            scan = Scan(**scans[s % ns])
            u_r, u_j = rng.random(n), rng.random(3 * n)
            inj = dict(u_yaw=rng.random(200), row=rng.integers(0, n_rows, 200), col=rng.integers(0, n_cols, 200), u_dx=rng.random(200), u_dy=rng.random(200))
            z = rng.standard_normal(3)
            t0 = time.perf_counter()
            o.diff_drive(enc_left[s], enc_right[s], z)
            o.update_particle_pos(P)
            P, _, _, _ = o.resample(P, 1, scan, u_r, u_j, inj)
            o.estimate_weighted_pose(P)
            dt = time.perf_counter() - t0
            if s >= warmup:
                times.append(dt)
                evals.append(n_particles * used_beams(scans[s % ns]))
    return times, evals, kind


def reference_arm(args):
    """The reference's own CPU implementation of the path on this box's host cores. It is single-threaded by construction
    (one roscpp spinner, MC:1212), so "all the host threads it can use" is one."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from montecarlolocalisation_b200 import synth
    K, W = args.steps, args.warmup
    if args.gpus == 1:
        n = args.particles
        sc = workload(W + K)
        times, evals, kind = run_cpu_steps(n, K, W, sc.occ, sc.scans, sc.enc_left, sc.enc_right)
        cfg = config_ref(n)
        sample = "full workload: %d particles x %d steps (+%d warm-up), every step the whole predict/update/resample/estimate loop" % (n, K, W)
    else:
        # the workload of our --gpus N > 1 line (configs[3]); the reference's filter (ray march, every 20th beam inside the
        # FOV: 24 of 720) runs on the same map and scans over a bounded sample of the particles
        n = min(args.ns_strong_particles, 200_000)
        occ, scans = ns_workload(NS_STRONG_CELLS, NS_STRONG_BEAMS, 4, seed=4)
        enc_l, enc_r = synth.encoder_trace(W + K)
        times, evals, kind = run_cpu_steps(n, K, W, occ, scans, enc_l, enc_r)
        cfg = config_ns_strong(args.ns_strong_particles)
        sample = ("bounded sample of the configs[3] workload: the same 4097x4097 map and 720-beam scans, %d of the %d particles per step, "
                  "%d steps (+%d warm-up) of the reference's own filter (ray-march sensor model, %d beams scored per particle)" % (
                      n, args.ns_strong_particles, K, W, used_beams(scans[0])))
    value = sum(evals) / sum(times)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": K,
        "warmup": W, "ms_per_step": 1e3 * sum(times) / len(times), "higher_is_better": True, "scaling": "weak" if args.gpus == 1 else "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": cfg,
        "arm": {"where": "host CPU, 1 thread", "api": "the reference's free functions (oracle/_ref: its unmodified translation unit)" if kind == "reference"
                else "oracle port (oracle/mcl_oracle.cpp)"},
        "steps_per_s": len(times) / sum(times),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": kind, "sample": sample, "host_cores_available": os.cpu_count()},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------------------
# parity gates
# ------------------------------------------------------------------------------------------------------------------
def _mix64(x):
    with np.errstate(over="ignore"):
        x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return x ^ (x >> np.uint64(31))


def state_hash(P, anc, g0):
    """Order-independent 64-bit hash of (global index, particle bits, ancestor) triples: shard hashes add (mod 2^64), so
    the value is the same however the particles are sharded."""
    with np.errstate(over="ignore"):
        w = np.ascontiguousarray(P, np.float32).view(np.uint32).astype(np.uint64)
        idx = np.arange(g0, g0 + len(P), dtype=np.uint64)
        x = _mix64(idx * np.uint64(0x9E3779B97F4A7C15) ^ (w[:, 0] | (w[:, 1] << np.uint64(32))))
        x = _mix64(x ^ (w[:, 2] | (w[:, 3] << np.uint64(32))))
        x = _mix64(x ^ np.asarray(anc).astype(np.int64).view(np.uint64))
        return int(np.add.reduce(x, dtype=np.uint64))


def ns_parity_gate(dist, rank, world, local, n=60_000, steps=3):
    """The sharded NS step (mcl_comm_init + mcl_ns_step: mailbox exchange, peer stores) over `world` GPUs on a 1025x1025 maze
    (fp32 field through L2, the multi-GPU legs' path) against the single-span CPU oracle, bit for bit, plus the 64-bit hash
    of particles + ancestors, which must be the same number at N = 1, 2, 4, 8."""
    from montecarlolocalisation_b200 import NsShard
    from oracle.pyoracle import NsOracle, Scan
    occ, scans = ns_workload(128, 720, steps, seed=3)
    RES = np.float32(0.1)
    shard = NsShard(rank, world, n, device=local)
    shard.pf.setMap(occ, RES)
    if world > 1:
        ids = [shard.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        shard.comm_init(ids[0])
    shard.pf.sampleParticles(n)
    motions = [(0.01 * (s + 1), 0.02 + 0.005 * s, -0.015) for s in range(steps)]
    pose = None
    for s in range(steps):
        pose = shard.step(motions[s], scan=scans[s], want_pose=(s == steps - 1))
    shard.pf.synchronize()
    P, A = shard.pf.downloadParticles(), shard.pf.ancestors()
    from montecarlolocalisation_b200 import ns_shard_range
    g0 = ns_shard_range(n, world, rank)[0]
    mine = (state_hash(P, A, g0), [float(v) for v in pose])
    allh = [mine]
    if world > 1:
        allh = [None] * world
        dist.all_gather_object(allh, mine)
    exchange = shard.exchange_used()
    del shard
    out = None
    if rank == 0:
        got = sum(h for h, _ in allh) & ((1 << 64) - 1)
        o = NsOracle()
        o.set_map(occ, RES)
        Po = o.init(0, n)
        for s in range(steps):
            Po, anc, ll, pre = o.step(Po, 0, Scan(**scans[s]), motions[s], s)
        want = state_hash(Po, anc, 0)
        same_pose = all(p == allh[0][1] for _, p in allh)
        ok = got == want and same_pose and bool(np.isfinite(allh[0][1]).all())
        out = {"status": "pass" if ok else "fail", "hash": "%016x" % got, "expected_hash": "%016x" % want, "pose_identical_on_every_rank": same_pose,
               "what": "sharded mcl_ns_step x %d on %d GPU(s), %d particles, 1025x1025 maze, 720-beam scans: particles + ancestors bit-identical "
                       "to the single-span CPU oracle (oracle/mcl_oracle_ns.cpp on rank 0); the hash is the same for every GPU count" % (steps, world, n),
               "exchange": exchange if world > 1 else "none (1 GPU)"}
    return out


def ref_parity_gate(local, n=1000, steps=3):
    """resampled indices + injected count bit-exact against the oracle, weights' total and CDF equal,
    pose within 1e-5, on configs[0]'s size (1000 particles, map.txt, 360 beams) with injected draws."""
    import montecarlolocalisation_b200 as m
    from oracle.pyoracle import Oracle, Scan
    from scenario import RES, Scenario
    sc = Scenario(steps)
    o = Oracle(trig_mode=0)
    o.set_map(sc.occ, RES)
    o.precompute_ray_directions(-120.0, 120.0, 0.1)
    pf = m.ParticleFilter(device=local)
    pf.setMap(sc.occ, RES)
    rng = np.random.default_rng(7)
    n_rows, n_cols = o.cell_ranges()
    init = dict(u_yaw=rng.random(n), row=rng.integers(0, n_rows, n).astype(np.int32), col=rng.integers(0, n_cols, n).astype(np.int32),
                u_dx=rng.random(n), u_dy=rng.random(n))
    P = o.sample_particles(init["u_yaw"], init["row"], init["col"], init["u_dx"], init["u_dy"])
    pf.sampleParticles(n, init)
    why = None
    for s in range(steps):
        z = rng.standard_normal(3)
        o.diff_drive(sc.enc_left[s], sc.enc_right[s], z)
        pf.diffDriveModel(sc.enc_left[s], sc.enc_right[s], z)
        o.update_particle_pos(P)
        scan = sc.scans[s]
        u_r, u_jit = rng.random(n), rng.random(3 * n)
        inj = dict(u_yaw=rng.random(200), row=rng.integers(0, n_rows, 200).astype(np.int32), col=rng.integers(0, n_cols, 200).astype(np.int32),
                   u_dx=rng.random(200), u_dy=rng.random(200))
        total = pf.computeWeight(scan["ranges"], scan["angle_min"], scan["angle_inc"], scan["range_min"], scan["range_max"])
        Pnew, idx, cdf, st = o.resample(P.copy(), 1, Scan(**scan), u_r, u_jit, inj)
        sg = pf.resampleParticles(1, u_r, u_jit, inj)
        Pg = pf.downloadParticles()
        pose_ok = np.allclose(pf.estimateWeightedPose(), o.estimate_weighted_pose(Pnew), rtol=1e-5, atol=1e-5)
        checks = {"total weight": total == st["total_weight"], "cdf": np.array_equal(pf.cdf(), cdf, equal_nan=True),
                  "ancestor indices": np.array_equal(pf.ancestors(), idx), "injected count": sg["injected"] == st["injected"],
                  "resampled x, y, w": np.array_equal(Pg[:, [0, 1, 3]], Pnew[:, [0, 1, 3]]),
                  "resampled theta (1e-5)": np.allclose(Pg[:, 2], Pnew[:, 2], rtol=1e-5, atol=1e-6), "pose (1e-5)": pose_ok}
        bad = [k for k, v in checks.items() if not v]
        if bad:
            why = "step %d: %s" % (s, ", ".join(bad))
            break
        pf.uploadParticles(Pnew)
        P = Pnew
    del pf
    return {"status": "pass" if why is None else "fail", "why": why,
            "what": "MCL_MODE_REF, %d particles, map.txt, 360-beam scans, %d steps with injected draws against oracle/mcl_oracle.cpp: total weight, CDF, "
                    "ancestor indices, injected count and resampled x/y/w bit-exact; theta and pose within 1e-5" % (n, steps)}


# ------------------------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------------------------
class Ctx:
    pass


def setup():
    import torch
    c = Ctx()
    c.torch = torch
    c.rank = int(os.environ.get("RANK", "0"))
    c.world = int(os.environ.get("WORLD_SIZE", "1"))
    c.local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the engine has no CPU path (use --impl reference for the CPU filter)")
    torch.cuda.set_device(c.local)
    c.dist = None
    if c.world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")        # keep stdout to the one JSON line
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", c.local))
        c.dist = dist

    def barrier():
        if c.world > 1:
            c.dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(*vals):
        if c.world == 1:
            return list(vals)
        t = torch.tensor(list(vals), dtype=torch.float64, device="cuda")
        c.dist.all_reduce(t, op=c.dist.ReduceOp.MAX)
        return t.tolist()

    c.barrier, c.max_over_ranks = barrier, max_over_ranks
    c.flush_buf = None
    return c


def l2_flusher(c, stream):
    torch = c.torch
    if c.flush_buf is None:
        c.flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def flush():
        with torch.cuda.stream(stream):
            c.flush_buf.zero_()
    return flush


def timed_resident(c, stream, fn, first, K, W, flush, after_warmup=None):
    """W warm-up + K timed steps; each timed step bracketed by CUDA events on the engine's stream (`flush` runs between
    steps, outside the events). Returns (per-step ms list, host wall seconds of the whole loop). after_warmup(): called
    once between the warm-up and the timed steps (the launch counter is read there)."""
    torch = c.torch
    for s in range(first, first + W):
        fn(s)
    if after_warmup:
        after_warmup()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    gc.collect()
    gc.disable()                 # a collection inside a 0.3 ms step would be charged to the step
    c.barrier()
    wall0 = time.perf_counter()
    for k in range(K):
        if flush:
            flush()
        ev[k][0].record(stream)
        fn(first + W + k)
        ev[k][1].record(stream)
    c.barrier()
    wall = time.perf_counter() - wall0
    gc.enable()
    return [a.elapsed_time(b) for a, b in ev], wall


def timed_e2e(c, stream, fn, first, K, W, flush):
    """The same loop through the public call with host buffers: fn(s) returns only when the step's result is on the host, and
    the HOST's wall clock brackets every call (scan H2D, every kernel, result D2H, the call's own overhead). `flush` + a
    device synchronisation run between steps, outside the timed intervals. Returns per-step seconds."""
    for s in range(first, first + W):
        fn(s)
    gc.collect()
    gc.disable()
    c.barrier()
    out = []
    for k in range(K):
        if flush:
            flush()
            stream.synchronize()
        if c.world > 1:
            c.dist.barrier()          # sharded steps start together (a step cannot finish before its slowest shard started)
        t0 = time.perf_counter()
        fn(first + W + k)
        out.append(time.perf_counter() - t0)
    gc.enable()
    c.barrier()
    return out


def kernel_table(prof, algo_of):
    """Per kernel: live CUDA-event duration per launch (events around every launch: serialised, tiny kernels read ~4 us
    high), share of the step, algorithmic bytes and what they make against the measured HBM peak."""
    total_ms = sum(v[0] for v in prof.values()) or 1.0
    peak, _ = peaks()
    kernels = {}
    for name, (ms, cnt) in sorted(prof.items(), key=lambda kv: -kv[1][0]):
        per = ms / cnt
        a = algo_of(name)
        gbs = (a / (per * 1e-3) / 1e9) if a else None
        kernels[name] = {"ms_per_launch": per, "launches": cnt, "share": ms / total_ms, "algo_bytes": a, "gbs": gbs,
                         "hbm_frac": (gbs / peak) if gbs else None}
    return kernels


def ref_leg(c, args, n, K, W, replicas_note=False, want_kernels=True):
    """MCL_MODE_REF on this rank's GPU (REF does not shard: its multinomial draw needs the global sequential f64 CDF).
    Returns the record of the configs[1] loop at n particles."""
    import montecarlolocalisation_b200 as m
    from scenario import RES
    torch = c.torch
    total_steps = 2 * (W + K) + K            # value pass, e2e pass, per-kernel profiling pass
    sc = workload(total_steps)
    pf = m.ParticleFilter(device=c.local, max_particles=n, seed=0x1234 + c.rank)
    pf.setMap(sc.occ, RES)
    pf.sampleParticles(n)
    stream = torch.cuda.ExternalStream(pf.stream(), device=c.local)
    flush = l2_flusher(c, stream)
    for s in range(total_steps):
        sca = sc.scans[s]
        pf.stageScan(s, sca["ranges"], sca["angle_min"], sca["angle_inc"], sca["range_min"], sca["range_max"])
    evals_per_step = [n * used_beams(sc.scans[s]) for s in range(total_steps)]

    def step_resident(s):
        """One tick through mcl_step_staged (predict, computeWeight, resample, estimate enqueued as one piece; the scan is
        already parked in HBM). --separate-calls: the four per-function calls, which wait for the GPU three times."""
        if args.separate_calls:
            pf.diffDriveModel(sc.enc_left[s], sc.enc_right[s])
            pf.computeWeightStaged(s)
            pf.resampleParticles(1)
            return pf.estimateWeightedPose()
        return pf.executeParticleFilter(sc.enc_left[s], sc.enc_right[s], 1, slot=s, want_result=False)      # queued: nothing read back

    pinned = [torch.from_numpy(np.ascontiguousarray(sc.scans[s]["ranges"])).pin_memory() for s in range(total_steps)]

    def step_e2e(s):
        """The same tick through mcl_step with the scan in (pinned) host memory: H2D of the scored beams, D2H of the tick's
        report (pose sums, injection state, counters); the call returns with the pose."""
        sca = sc.scans[s]
        if args.separate_calls:
            pf.diffDriveModel(sc.enc_left[s], sc.enc_right[s])
            pf.computeWeight(pinned[s].numpy(), sca["angle_min"], sca["angle_inc"], sca["range_min"], sca["range_max"])
            pf.resampleParticles(1)
            return pf.estimateWeightedPose()
        sca = dict(sca)
        sca["ranges"] = pinned[s].numpy()
        return pf.executeParticleFilter(sc.enc_left[s], sc.enc_right[s], 1, scan=sca)[0]

    mark = {}
    ms_res, wall_res = timed_resident(c, stream, step_resident, 0, K, W, flush, after_warmup=lambda: mark.update(l0=pf.kernelLaunches()))
    launches = pf.kernelLaunches() - mark["l0"]          # kernels of the K timed ticks only
    s_e2e = timed_e2e(c, stream, step_e2e, W + K, K, W, flush)
    kernels = None
    if want_kernels:
        # per-kernel durations (CUDA events around every launch on the engine's stream) over K more steps
        pf.profileEnable(True)
        for k in range(K):
            flush()
            step_resident(2 * (W + K) + k)
        prof = pf.profileRead()
        pf.profileEnable(False)

        def algo(name):
            b = ALGO_BYTES.get(name)
            if b is None:
                return None
            a = n * b
            if name in ("k_ref_update", "k_ref_update_v2"):
                a += n * 12 * 11        # <= 11 one-byte map probes per scored beam
            return a
        kernels = kernel_table(prof, algo)
    t_res, t_e2e = c.max_over_ranks(sum(ms_res) * 1e-3, sum(s_e2e))
    ev_res = sum(evals_per_step[W:W + K])
    ev_e2e = sum(evals_per_step[2 * W + K:2 * (W + K)])
    scan_bytes = int(sc.scans[0]["ranges"].nbytes) + 16 + 16         # ranges + 4 float32 scan fields + 2 encoder doubles
    rec = {
        "value": c.world * ev_res / t_res, "unit": UNIT, "ms_per_step": 1e3 * t_res / K,
        "ms_per_step_rank0": {"min": min(ms_res), "median": statistics.median(ms_res), "max": max(ms_res)},
        "steps_per_s": c.world * K / t_res,
        "e2e": {"value": c.world * ev_e2e / t_e2e, "unit": UNIT, "h2d_bytes_per_step": scan_bytes,
                "d2h_bytes_per_step": 24 + 8 + 48 if args.separate_calls else 104,    # mcl_step: the 104-byte tick report (pose sums, injection state,
                # counters, sequence number), stored by the last kernel straight into the engine's pinned block. h2d: the scan as
                # the caller hands it over (host buffer); of that, the scored beams (24 B each) reach the GPU in the launch parameters
                "ms_per_step": 1e3 * t_e2e / K, "clock": "host wall clock around each mcl_step call (returns with the pose); L2 flush + device "
                "synchronisation between calls, outside the timed intervals", "ms_per_step_rank0": {"min": 1e3 * min(s_e2e), "median": 1e3 * statistics.median(s_e2e), "max": 1e3 * max(s_e2e)}},
        "gpu_launches": launches, "launches_per_step": launches / K,
        "wall_ms_per_step_queued": 1e3 * wall_res / K,
    }
    if kernels is not None:
        rec["kernels"] = kernels
        rec["roofline"] = roofline_record(kernels, "ref1M" if n == N_PARTICLES else "", n)
    del pf
    return rec


def reference_sized(c, ticks=300):
    """The reference's own operating point (MC:84: 1500 particles at 10 Hz; BASELINE.json configs[0]: 1000) on the 360-beam
    synthetic scan and on the robot's 683-beam / 0.352-degree scan (35 scored beams), GPU and CPU in this same run."""
    import montecarlolocalisation_b200 as m
    from scenario import RES, Scenario
    out = []
    sc = Scenario(ticks + 20)
    for label, scans in (("360-beam full-circle scan, 12 scored beams", sc.scans),
                         ("683-beam 0.352-degree scan from -120 degrees (the robot's LIDAR, MC:638-640), 35 scored beams",
                          [robot_scan(sc.occ, sc.truth[s], 500 + s) for s in range(ticks + 20)])):
        scored = used_beams(scans[0])
        for n in (1000, 1500):
            pf = m.ParticleFilter(device=c.local, max_particles=n, seed=0x77)
            pf.setMap(sc.occ, RES)
            pf.sampleParticles(n)
            l0 = None
            ts = []
            for s in range(ticks + 20):
                if s == 20:
                    l0 = pf.kernelLaunches()
                t0 = time.perf_counter()
                pf.executeParticleFilter(sc.enc_left[s], sc.enc_right[s], 1, scan=scans[s])
                if s >= 20:
                    ts.append(time.perf_counter() - t0)
            launches = (pf.kernelLaunches() - l0) / ticks
            del pf
            t_cpu, ev_cpu, kind = run_cpu_steps(n, 40, 3, sc.occ, scans, sc.enc_left, sc.enc_right)
            gpu_us = 1e6 * statistics.median(ts)
            cpu_ms = 1e3 * statistics.median(t_cpu)
            out.append({"particles": n, "scan": label, "scored_beams": scored,
                        "gpu_us_per_tick": gpu_us, "gpu_us_per_tick_min": 1e6 * min(ts), "gpu_launches_per_tick": launches,
                        "gpu_evals_per_s": n * scored / (gpu_us * 1e-6),
                        "cpu_ms_per_tick": cpu_ms, "cpu_evals_per_s": n * scored / (cpu_ms * 1e-3), "cpu_kind": kind, "cpu_cores": 1,
                        "gpu_over_cpu": cpu_ms * 1e3 / gpu_us,
                        "clock": "host wall clock around mcl_step (scan from host memory in, pose out), median of %d ticks; CPU: median of 40 steps of the "
                                 "reference's filter on one thread" % ticks})
    return out


# ------------------------------------------------------------------------------------------------------------------
# NS leg: the north-star formulation (likelihood field, Philox motion noise, fixed-point systematic resampling), sharded
# across the GPUs of the job by mcl_ns_step: the step's collectives go through peer-memory mailboxes over NVLink (or NCCL
# with MCL_NS_EXCHANGE=nccl), resampled particles are stored by the resampling kernel straight into the owning shard's
# memory (CUDA IPC peer mappings), so no separate rebalance pass exists.
# ------------------------------------------------------------------------------------------------------------------
def ns_leg(c, args, cells, n_global, n_beams, label, K, W, key, map_seed=4, scaling="weak", uniform=True, gather=True):
    from montecarlolocalisation_b200 import NsShard
    torch = c.torch
    rank, world, local = c.rank, c.world, c.local
    per_gpu = (n_global + world - 1) // world
    n_scans = 4
    if cells == 6:
        from scenario import Scenario
        sc = Scenario(n_scans, n_beams=n_beams)
        occ, scans = sc.occ, sc.scans
    else:
        occ, scans = ns_workload(cells, n_beams, n_scans, seed=map_seed)
    shard = NsShard(rank, world, n_global, device=local, max_particles=0, seed=0xABCDEF)
    shard.pf.setMap(occ, np.float32(0.1))
    if world > 1:
        # the engine's own NCCL communicator (used for bootstrap and, with MCL_NS_EXCHANGE=nccl, the step's collectives) +
        # CUDA-IPC peer mappings
        ids = [shard.comm_unique_id() if rank == 0 else None]
        c.dist.broadcast_object_list(ids, src=0)
        shard.comm_init(ids[0])
    shard.pf.sampleParticles(n_global)
    for i, sca in enumerate(scans):
        shard.pf.stageScan(i, sca["ranges"], sca["angle_min"], sca["angle_inc"], sca["range_min"], sca["range_max"])
    stream = torch.cuda.ExternalStream(shard.pf.stream(), device=local)
    pinned = [torch.from_numpy(np.ascontiguousarray(sca["ranges"])).pin_memory() for sca in scans]
    valid_beams = [int((np.isfinite(sca["ranges"]) & (sca["ranges"] >= sca["range_min"]) & (sca["ranges"] <= sca["range_max"]) &
                        (sca["ranges"] < 5.6)).sum()) for sca in scans]
    motion = (0.01, 0.02, -0.005)
    small = per_gpu * 48 < (256 << 20)          # working set that L2 could hold between steps: flush
    flush = l2_flusher(c, stream) if small else None

    def step_resident(i):
        """One whole filter step enqueued by the engine (mcl_ns_step_staged): predict -> likelihood field -> all-reduce(max)
        -> Q32 weights + prefix + pose sums -> all-gather(totals) -> device-side plan -> resample into the owning shards ->
        barrier. Nothing is read back."""
        return shard.step(motion, slot=i % n_scans)

    def step_e2e(i):
        """The same step through mcl_ns_step with the scan in (pinned) host memory and the weighted-mean pose read back."""
        sca = dict(scans[i % n_scans])
        sca["ranges"] = pinned[i % n_scans].numpy()
        return shard.step(motion, scan=sca, want_pose=True)

    mark = {}
    ms_res, wall_res = timed_resident(c, stream, step_resident, 0, K, W, flush, after_warmup=lambda: mark.update(l0=shard.pf.kernelLaunches()))
    launches = shard.pf.kernelLaunches() - mark["l0"]    # kernels of the K timed steps only
    s_e2e = timed_e2e(c, stream, step_e2e, W + K, K, W, flush)
    shard.pf.profileEnable(True)
    for k in range(K):
        if flush:
            flush()
        step_resident(2 * (W + K) + k)
    prof = shard.pf.profileRead()
    shard.pf.profileEnable(False)
    steady_form = shard.field_form()
    uni = None
    if uniform:
        # kidnapped-robot case: freshly uniform particles (no spatial locality in the field gathers)
        for k in range(6):
            if k == 2:          # two untimed launches first: fields larger than L2 settle on their field form by measurement
                shard.pf.profileEnable(True)
            shard.pf.sampleParticles(n_global)
            shard.pf.updateParticlePos(*motion)
            shard.update_local_staged(0)
        u = shard.pf.profileRead().get("k_ns_update", (0.0, 1))
        shard.pf.profileEnable(False)
        uni = {"k_ns_update_ms": u[0] / u[1], "field_form": shard.field_form(), "evals_per_s_per_gpu": per_gpu * valid_beams[0] / (u[0] / u[1] * 1e-3),
               "note": "sensor-model kernel alone on freshly uniform particles (kidnapped robot): worst case for gather locality"}
    t_res, t_e2e = c.max_over_ranks(sum(ms_res) * 1e-3, sum(s_e2e))
    n_mine = shard.pf.num_particles
    evals_res = sum(n_global * valid_beams[i % n_scans] for i in range(W, W + K))
    evals_e2e = sum(n_global * valid_beams[i % n_scans] for i in range(2 * W + K, 2 * (W + K)))
    nb = valid_beams[0]
    algo = {"k_ns_update": n_mine * (20 + 4 * nb), "k_ns_predict": n_mine * 32, "k_ns_weights_sum": n_mine * 4,
            "k_ns_weights_scan": n_mine * 12, "k_ns_weights_pose": n_mine * 32, "k_ns_resample": n_mine * 44, "k_ns_pose_partials": n_mine * 20}
    kernels = kernel_table(prof, lambda name: algo.get(name))
    per_rank = None
    if world > 1:
        # every shard's own kernel times (the table above is rank 0's): the sensor-model kernel's spread over the shards is the
        # skew that the step's exchanges wait for
        mine_t = {name: 1e3 * v[0] / v[1] for name, v in prof.items() if name in ("k_ns_update", "k_ns_resample", "k_ns_predict", "k_ns_plan")}
        # the rebalance: which of the slots this shard resolved in its last step live on another shard (stored over NVLink)
        k_lo, k_hi, own_b, own_n = shard.last_plan()
        local = max(0, min(k_hi, own_b + own_n) - max(k_lo, own_b))
        mine_t["remote_fraction"] = (1.0 - local / (k_hi - k_lo)) if k_hi > k_lo else 0.0
        mine_t["remote_mb"] = (k_hi - k_lo - local) * 20 / 1e6                # float4 + ancestor index per survivor
        allt = [None] * world
        c.dist.all_gather_object(allt, mine_t)
        per_rank = {name: [round(t.get(name, 0.0), 3 if name == "remote_fraction" else 1) for t in allt] for name in mine_t}
    field_bytes = occ.size * 4
    out = {
        "label": label, "value": evals_res / t_res, "unit": UNIT, "ms_per_step": 1e3 * t_res / K, "steps_per_s": K / t_res,
        "ms_per_step_rank0": {"min": min(ms_res), "median": statistics.median(ms_res), "max": max(ms_res)},
        "e2e": {"value": evals_e2e / t_e2e, "unit": UNIT, "ms_per_step": 1e3 * t_e2e / K, "h2d_bytes_per_step": int(scans[0]["ranges"].nbytes) + 28,
                "d2h_bytes_per_step": 40, "clock": "host wall clock around each mcl_ns_step call (returns with the pose), max over ranks of the "
                "sum; ranks start every step together%s" % ("; L2 flush + synchronisation between calls, outside the timed intervals" if small else "")},
        "config": {"workload": "%s: %dx%d occupancy grid, %d particles on %d GPU(s) (%d per GPU), %d-beam scan (%d valid beams scored per particle), "
                               "NS mode: Philox motion noise -> likelihood field -> weighted-mean pose -> Q32 systematic resampling" % (
                                   label, occ.shape[1], occ.shape[0], n_global, world, per_gpu, n_beams, nb),
                   "field": "%d KiB log-likelihood field, %s" % (field_bytes // 1024, {"smem-f32": "TMA-staged into shared memory", "global-f32": "fp32, gathered through L1/L2",
                                                                       "global-u8": "as one-byte codes (%d KiB) gathered through L1/L2 + shared-memory code table" % (field_bytes // 4096)}.get(steady_form, steady_form)),
                   "collectives": "none (1 GPU)" if world == 1 else (
                       "peer-memory mailboxes, no NCCL on the data path: every shard stores {payload, tag} into the other shards' mailboxes over NVLink (CUDA IPC) and polls its own; 32-thread kernels on the filter's stream, no host round trip: all-reduce(max), all-gather(Q32 totals) + all-reduce(pose) in one exchange fused with the resampling plan, closing barrier; resampled particles stored into peer shards over NVLink"
                       if shard.exchange_used() == "peer" else
                       "engine-enqueued NCCL on the filter's stream, no host round trip: all-reduce(max), all-gather(Q32 totals), all-reduce(pose), closing all-reduce as barrier; resampled particles stored into peer shards over NVLink (CUDA IPC)"),
                   "l2": ("L2 flushed between timed steps (256 MiB write, outside the timed region)" if small else
                          "inputs larger than L2: per-GPU working set %.0f MB" % (per_gpu * 48 / 1e6))},
        "gpu_launches": launches, "launches_per_step": launches / K, "scaling": scaling,
        "roofline": roofline_record(kernels, key, per_gpu),
        "kernels": kernels,
    }
    if per_rank:
        out["kernel_us_per_rank"] = per_rank
        u = per_rank.get("k_ns_update")
        if u and min(u) > 0:
            out["update_skew"] = {"max_over_min": max(u) / min(u), "max_over_mean": max(u) / (sum(u) / len(u))}
    if uni:
        out["uniform_particles"] = uni
    if gather:
        out["gather_microbench_reads_per_s"] = {"shared_memory_table": shard.pf.benchGather(0, min(field_bytes, 190 * 1024)),
                                                "global_table_of_field_size": shard.pf.benchGather(1, field_bytes)}
    del shard
    torch.cuda.empty_cache()
    return out


def ours(args):
    c = setup()
    K, W = args.steps, args.warmup
    world, rank = c.world, c.rank
    sampler = ClockSampler(range(world)) if rank == 0 else None      # rank 0 samples every GPU of the job
    gates = {}
    if not args.no_gate:
        gates["ns_sharded_step"] = ns_parity_gate(c.dist, rank, world, c.local)
        if world == 1:
            gates["ref_loop"] = ref_parity_gate(c.local)
    if sampler:
        sampler.start()
    line = {"metric": METRIC, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W, "higher_is_better": True, "vs_baseline": None,
            "data": "synthetic"}
    if world == 1:
        rec = ref_leg(c, args, args.particles, K, W)
        line.update({"value": rec["value"], "ms_per_step": rec["ms_per_step"], "scaling": "weak", "dtype": "f64",
                     "config": config_ref(args.particles),
                     "arm": {"where": "B200 x1", "mode": "MCL_MODE_REF (results identical to the reference for the same draws)",
                             "api": "four calls per tick: mcl_predict_encoders, mcl_update[_staged], mcl_resample, mcl_estimate" if args.separate_calls
                             else "one call per tick: mcl_step_staged (value) / mcl_step (e2e)"}})
        for k in ("ms_per_step_rank0", "steps_per_s", "e2e", "gpu_launches", "launches_per_step", "roofline", "kernels", "wall_ms_per_step_queued"):
            line[k] = rec[k]
        if not args.no_reference_sized:
            line["reference_sized"] = reference_sized(c)
        if not args.no_ns:
            c.flush_buf = None
            c.torch.cuda.empty_cache()
            ns = {}
            ns["map_txt_1M"] = ns_leg(c, args, 6, 1_000_000, 360, "configs[1] shape", K, W, "map_txt_1M")
            ns["configs2"] = ns_leg(c, args, 128, 10_000_000, 720, "BASELINE.json configs[2]: 1025x1025 maze (seed 3), 10 M uniformly initialised particles, 1 GPU",
                                    K, W, "configs2", map_seed=3)
            ns["grid4096"] = ns_leg(c, args, args.ns_cells, args.ns_particles, 720, "configs[3] per-GPU shape (weak scaling)", K, W, "grid4096")
            if not args.no_ns_large:
                ns["grid8192"] = ns_leg(c, args, 1024, args.ns_particles, 1080, "configs[4] per-GPU shape (kidnapped robot; weak scaling)", K, W, "grid8192")
                ns["configs3_strong"] = ns_leg(c, args, NS_STRONG_CELLS, args.ns_strong_particles, NS_STRONG_BEAMS,
                                               "BASELINE.json configs[3] in full on one GPU: the N = 1 point of the strong-scaling curve", K, W,
                                               "grid4096_strong", scaling="strong", uniform=False, gather=False)
            line["ns"] = ns
    else:
        rec = ns_leg(c, args, NS_STRONG_CELLS, args.ns_strong_particles, NS_STRONG_BEAMS, "BASELINE.json configs[3]", K, W, "grid4096_strong",
                     scaling="strong", uniform=False, gather=False)
        line.update({"value": rec["value"], "ms_per_step": rec["ms_per_step"], "scaling": "strong", "dtype": "f32",
                     "config": config_ns_strong(args.ns_strong_particles),
                     "arm": {"where": "B200 x%d, one rank per GPU" % world, "mode": "MCL_MODE_NS, sharded (mcl_ns_step)",
                             "api": "one call per step and shard: mcl_ns_step_staged (value) / mcl_ns_step (e2e)", "detail": rec["config"]}})
        for k in ("ms_per_step_rank0", "steps_per_s", "e2e", "gpu_launches", "launches_per_step", "roofline", "kernels"):
            line[k] = rec[k]
        if not args.no_ns:
            weak = {}
            weak["map_txt_1M"] = ns_leg(c, args, 6, 1_000_000 * world, 360, "configs[1] shape, 1 M particles per GPU (weak scaling)", K, W, "map_txt_1M", uniform=False, gather=False)
            weak["grid4096"] = ns_leg(c, args, args.ns_cells, args.ns_particles * world, 720, "configs[3] grid, 12.5 M particles per GPU (weak scaling)", K, W, "grid4096", gather=False)
            if not args.no_ns_large:
                weak["grid8192"] = ns_leg(c, args, 1024, args.ns_particles * world, 1080,
                                          "BASELINE.json configs[4]: 8193x8193 grid, 1080 beams, 12.5 M particles per GPU (1e8 on 8 GPUs), kidnapped robot", K, W,
                                          "grid8192", gather=False)
            line["ns_weak"] = weak
            rr = ref_leg(c, args, args.particles, K, W, want_kernels=False)
            rr["note"] = "MCL_MODE_REF does not shard (its multinomial draw needs the global sequential f64 CDF): %d independent replicas, no communication" % world
            line["ref_replicas"] = rr
    clocks = sampler.stop() if sampler else None           # sampled over every timed region of this run
    if rank == 0:
        line["clocks"] = clocks
        line["parity_gate"] = ("skipped" if not gates else "pass" if all(g and g["status"] == "pass" for g in gates.values()) else "fail")
        line["parity_gates"] = gates
        if world == 1 and not args.no_cpu_baseline:
            n_cpu = min(args.particles, 200_000)
            csc = workload(3)
            times, evals, kind = run_cpu_steps(n_cpu, 2, 1, csc.occ, csc.scans, csc.enc_left, csc.enc_right)
            line["cpu_baseline"] = {"value": sum(evals) / sum(times), "unit": UNIT, "cores": 1, "kind": kind,
                                    "sample": "same workload at %d particles, 2 timed full steps after 1 warm-up, 1 host thread "
                                              "(the reference is single-threaded)" % n_cpu,
                                    "ms_per_step": 1e3 * sum(times) / len(times), "host_cores_available": os.cpu_count()}
        print(json.dumps(line), flush=True)
    if world > 1:
        c.dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--particles", type=int, default=N_PARTICLES)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gate", action="store_true", help="skip the parity gates")
    ap.add_argument("--no-reference-sized", action="store_true", help="skip the 1000 / 1500-particle GPU + CPU records")
    ap.add_argument("--separate-calls", action="store_true", help="REF loop through the four per-function calls instead of mcl_step")
    ap.add_argument("--no-ns", action="store_true", help="skip the secondary NS (north-star) legs")
    ap.add_argument("--no-ns-large", action="store_true", help="skip the 8193x8193 / 1080-beam and the 1e8-particle single-GPU NS cases")
    ap.add_argument("--ns-cells", type=int, default=512, help="NS weak leg: maze cells per side (512 -> 4097x4097 grid)")
    ap.add_argument("--ns-particles", type=int, default=12_500_000, help="NS weak legs: particles per GPU")
    ap.add_argument("--ns-strong-particles", type=int, default=NS_STRONG_PARTICLES, help="NS strong-scaling leg: particles in total")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3          # timing rule: at least 3 warm-up steps
    if args.impl == "reference":
        reference_arm(args)
    else:
        ours(args)


if __name__ == "__main__":
    main()
