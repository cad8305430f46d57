#!/usr/bin/env python
"""bench.py — particle-filter hot path on B200 (and the reference's CPU filter beside it).

    python bench.py --gpus 1 --steps 10 --warmup 3            # our arm: CUDA engine through the C-ABI
    python bench.py --impl reference --steps 3 --warmup 1     # reference arm: the reference's own CPU filter

Workload = BASELINE.json configs[1]: pink_fundamentals/map.txt (49x49 grid), 1,000,000 particles, 360-beam synthetic
LIDAR scan + wheel-encoder trace, the full predict -> update -> resample -> estimate loop in the reference-parity mode
(results identical to the reference for the same draws). One "step" = one such loop iteration; one "eval" = one
(particle, scored beam) pair: the reference scores every 20th beam inside (-120,120) degrees = 12 of 360.

Prints ONE JSON line (see the keys below). `value` times the loop with the scans already parked in HBM; `e2e` times
the same loop through the public per-call C-ABI with host buffers (scan in, pose out, every step).
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
METRIC = "particle-beam likelihood evals/sec (full predict/update/resample/estimate loop)"
UNIT = "evals/s"

# algorithmic bytes per particle per launch (SURVEY.md §8d / DESIGN.md "kernels"); update adds 1 B per map probe
ALGO_BYTES = {
    "k_ref_predict": 32, "k_ref_update": 20, "k_ref_update_v2": 20, "k_ref_first_touch": 16, "k_ref_seq_total": 4, "k_ref_seq_cdf": 16,
    "k_ref_resample": 44 + 32, "k_fill_resample_draws": 32, "k_pose_wsum": 16, "k_pose_sums": 16,
    "k_ref_exact_scan": 12, "k_ref_normalise": 8,
}


def traffic_for(key, particles):
    """DRAM bytes per launch from the committed ncu capture of this kernel at this workload size, else None."""
    p = os.path.join(ROOT, "profiles", "traffic_r1.json")
    if not os.path.exists(p):
        return None
    t = json.load(open(p)).get(key)
    return t["dram_bytes"] if t and t.get("particles") == particles else None


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p))["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


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


# ------------------------------------------------------------------------------------------------------------------
# reference arm: the reference's own CPU filter (oracle/_ref = its unmodified translation unit; else the oracle port)
# ------------------------------------------------------------------------------------------------------------------
def cpu_filter(n_particles):
    from oracle import pyoracle
    from scenario import RES
    pyoracle.build()
    if pyoracle.ref_available():
        r = pyoracle.Ref()
        kind = "reference"
    else:
        r = None
        kind = "port"
    return pyoracle, r, kind, RES


def run_cpu_steps(n_particles, steps, warmup, sc):
    """Times full filter steps of the reference on 1 host thread (it is single-threaded, MC:1212). Returns
    (seconds per timed step list, evals per step list, kind)."""
    pyoracle, ref, kind, RES = cpu_filter(n_particles)
    from oracle.pyoracle import Oracle, Scan
    rng = np.random.default_rng(123)
    times, evals = [], []
    if kind == "reference":
        ref.set_map(sc.occ, RES)
        ref.precompute_ray_directions(-120.0, 120.0, 0.1)
        ref.clear_seeds()
        ref.push_seeds(*[int(x) for x in rng.integers(1, 2**31 - 1, 64)])
        ref.seed_static_engines(11, 12)
        P = ref.sample_particles(n_particles)
        for s in range(warmup + steps):
            scan = Scan(**sc.scans[s])
            ref.set_scan(scan)
            ref.push_seeds(*[int(x) for x in rng.integers(1, 2**31 - 1, 256)])
            t0 = time.perf_counter()
            ref.diff_drive(sc.enc_left[s], sc.enc_right[s])          # diffDriveModel          (MC:1084)
            ref.update_particle_pos(P)                               # updateParticlePos       (MC:1086)
            P, _ = ref.resample(P, 1)                                # resampleParticles       (MC:1089) incl. computeWeight
            ref.estimate_weighted_pose(P)                            # estimateWeightedPose    (MC:782)
            dt = time.perf_counter() - t0
            if s >= warmup:
                times.append(dt)
                evals.append(n_particles * used_beams(sc.scans[s]))
    else:
        o = Oracle(trig_mode=0)
        o.set_map(sc.occ, RES)
        o.precompute_ray_directions(-120.0, 120.0, 0.1)
        n = n_particles
        P = o.sample_particles(rng.random(n), rng.integers(0, 6, n), rng.integers(0, 6, n), rng.random(n), rng.random(n))
        for s in range(warmup + steps):
            scan = Scan(**sc.scans[s])
            u_r, u_j = rng.random(n), rng.random(3 * n)
            inj = dict(u_yaw=rng.random(200), row=rng.integers(0, 6, 200), col=rng.integers(0, 6, 200), u_dx=rng.random(200), u_dy=rng.random(200))
            z = rng.standard_normal(3)
            t0 = time.perf_counter()
            o.diff_drive(sc.enc_left[s], sc.enc_right[s], z)
            o.update_particle_pos(P)
            P, _, _, _ = o.resample(P, 1, scan, u_r, u_j, inj)
            o.estimate_weighted_pose(P)
            dt = time.perf_counter() - t0
            if s >= warmup:
                times.append(dt)
                evals.append(n_particles * used_beams(sc.scans[s]))
    return times, evals, kind


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n = args.particles
    sc = workload(args.warmup + args.steps)
    times, evals, kind = run_cpu_steps(n, args.steps, args.warmup, sc)
    value = sum(evals) / sum(times)
    sample = "full workload: %d particles x %d steps (+%d warm-up), every step the whole predict/update/resample/estimate loop" % (
        n, args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * sum(times) / len(times), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_dict(n, "host CPU, 1 thread"),
        "steps_per_s": len(times) / sum(times),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": kind, "sample": sample,
                         "host_cores_available": os.cpu_count()},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def config_dict(n, where):
    return {"workload": "BASELINE.json configs[1]: pink_fundamentals/map.txt 49x49 grid @0.1 m, %d particles, %d-beam synthetic scan "
                        "(12 beams scored per particle: every 20th inside +-120 deg), full predict/update/resample/estimate loop, "
                        "reference-parity mode (MCL_MODE_REF)" % (n, N_BEAMS),
            "particles": n, "beams": N_BEAMS, "map": "map.txt 49x49", "mode": "ref", "where": where,
            "l2": "L2 flushed between timed steps (256 MiB write, outside the timed events)"}


# ------------------------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------------------------
def ours(args):
    import torch
    import montecarlolocalisation_b200 as m
    from scenario import RES

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the engine has no CPU path (use --impl reference for the CPU filter)")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")        # keep stdout to the one JSON line
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    n = args.particles
    K, W = args.steps, args.warmup
    total_steps = 2 * (W + K) + K            # value pass, e2e pass, per-kernel profiling pass
    sc = workload(total_steps)
    # REF mode does not shard (multinomial needs the global f64 CDF): every rank runs an independent replica
    pf = m.ParticleFilter(device=local, max_particles=n, seed=0x1234 + rank)
    pf.setMap(sc.occ, RES)
    pf.sampleParticles(n)
    stream = torch.cuda.ExternalStream(pf.stream(), device=local)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def flush_l2():
        with torch.cuda.stream(stream):
            flush.zero_()

    for s in range(total_steps):
        sca = sc.scans[s]
        pf.stageScan(s, sca["ranges"], sca["angle_min"], sca["angle_inc"], sca["range_min"], sca["range_max"])
    # motion triples precomputed on the host (3 scalars per step, passed as kernel arguments)
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
        """The same tick through mcl_step with the scan in (pinned) host memory: H2D of the scored beams, D2H of the weight
        total, the counters and the pose inside the timed region."""
        sca = sc.scans[s]
        if args.separate_calls:
            pf.diffDriveModel(sc.enc_left[s], sc.enc_right[s])
            pf.computeWeight(pinned[s].numpy(), sca["angle_min"], sca["angle_inc"], sca["range_min"], sca["range_max"])
            pf.resampleParticles(1)
            return pf.estimateWeightedPose()
        sca = dict(sca)
        sca["ranges"] = pinned[s].numpy()
        return pf.executeParticleFilter(sc.enc_left[s], sc.enc_right[s], 1, scan=sca)[0]

    def timed(fn, first):
        """W warm-up + K timed steps; each timed step bracketed by CUDA events on the engine's stream, L2 flushed
        between steps outside the events. Returns per-step ms list."""
        for s in range(first, first + W):
            fn(s)
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
        gc.collect()
        gc.disable()                 # a collection inside a 0.4 ms step would be charged to the step
        barrier()
        wall0 = time.perf_counter()
        for k in range(K):
            flush_l2()
            ev[k][0].record(stream)
            fn(first + W + k)
            ev[k][1].record(stream)
        barrier()
        wall = time.perf_counter() - wall0
        gc.enable()
        return [a.elapsed_time(b) for a, b in ev], wall

    sampler = ClockSampler(range(world)) if rank == 0 else None      # rank 0 samples every GPU of the job
    if sampler:
        sampler.start()
    launches0 = pf.kernelLaunches()
    ms_res, wall_res = timed(step_resident, 0)
    launches = pf.kernelLaunches() - launches0
    ms_e2e, wall_e2e = timed(step_e2e, W + K)

    # per-kernel durations (CUDA events around every launch on the engine's stream) over K more steps
    pf.profileEnable(True)
    for k in range(K):
        flush_l2()
        step_resident(2 * (W + K) + k)
    prof = pf.profileRead()
    pf.profileEnable(False)

    ev_res = sum(evals_per_step[W:W + K])
    ev_e2e = sum(evals_per_step[2 * W + K:2 * (W + K)])
    t_res = sum(ms_res) * 1e-3
    t_e2e = sum(ms_e2e) * 1e-3
    if world > 1:
        t = torch.tensor([t_res, t_e2e], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_res, t_e2e = t.tolist()
    value = world * ev_res / t_res
    e2e_value = world * ev_e2e / t_e2e

    # roofline of the dominant kernel
    peak, peak_src = peaks()
    kernels = {}
    total_kernel_ms = sum(v[0] for v in prof.values())
    for name, (ms, cnt) in sorted(prof.items(), key=lambda kv: -kv[1][0]):
        per_launch_ms = ms / cnt
        b = ALGO_BYTES.get(name)
        algo = None
        if b is not None:
            algo = n * b
            if name in ("k_ref_update", "k_ref_update_v2"):
                algo += n * 12 * 11        # <= 11 one-byte map probes per scored beam
        kernels[name] = {"ms_per_launch": per_launch_ms, "launches": cnt, "share": ms / total_kernel_ms,
                         "algo_bytes": algo, "gbs": (algo / (per_launch_ms * 1e-3) / 1e9) if algo else None}
    top = next(iter(kernels))
    roof = {"bound": "hbm", "kernel": top, "achieved": kernels[top]["gbs"], "peak": peak, "unit": "GB/s",
            "frac": (kernels[top]["gbs"] / peak) if kernels[top]["gbs"] else None, "traffic": traffic_for(top, n), "peak_source": peak_src,
            "share_of_step": kernels[top]["share"],
            "note": "k_ref_update_v2 is instruction-issue bound (69% of issue slots, 111M warp-instructions per launch, ncu: profiles/r1_ref_update_v2_ncu.txt), not HBM bound: algorithmic bytes = 20 B/particle + "
                    "one byte per map probe (<= 11 per scored beam); its DRAM traffic is the 16 B/particle particle stream"}

    ns = None
    if not args.no_ns:
        del flush
        torch.cuda.empty_cache()
        ns = {}
        ns["map_txt_1M"] = ns_leg(args, torch, dist if world > 1 else None, rank, world, local, 6, 1_000_000, 360, "configs[1] shape", K, W, "map_txt_1M")
        ns["grid4096"] = ns_leg(args, torch, dist if world > 1 else None, rank, world, local, args.ns_cells, args.ns_particles, 720,
                                "configs[3] per-GPU shape", K, W, "grid4096")
        if not args.no_ns_large:
            ns["grid8192"] = ns_leg(args, torch, dist if world > 1 else None, rank, world, local, 1024, args.ns_particles, 1080,
                                    "configs[4] per-GPU shape (kidnapped robot)", K, W, "grid8192")
    clocks = sampler.stop() if sampler else None           # sampled over every timed region of this run (REF loop and NS legs)
    if rank == 0:
        scan_bytes = int(sc.scans[0]["ranges"].nbytes) + 16 + 16         # ranges + 4 float32 scan fields + 2 encoder doubles
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": 1e3 * t_res / K, "ms_per_step_rank0": {"min": min(ms_res), "median": statistics.median(ms_res), "max": max(ms_res)},
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": dict(config_dict(n, "B200 x%d%s" % (world, ", independent replicas" if world > 1 else "")),
                                                api="four calls per tick: mcl_predict_encoders, mcl_update[_staged], mcl_resample, mcl_estimate" if args.separate_calls
                                                else "one call per tick: mcl_step_staged (value) / mcl_step (e2e)"),
            "steps_per_s": world * K / t_res,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": scan_bytes, "d2h_bytes_per_step": 24 + 8 + 48 if args.separate_calls else 88,      # mcl_step: the 88-byte tick report (pose sums, injection
                    # state, counters), stored by the last kernel straight into the caller's pinned block
                    "ms_per_step": 1e3 * t_e2e / K, "wall_ms_per_step": 1e3 * wall_e2e / K},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": roof,
            "kernels": kernels,
            "wall_ms_per_step": 1e3 * wall_res / K,
        }
        if ns is not None:
            line["ns"] = ns
        if world == 1 and not args.no_cpu_baseline:
            n_cpu = min(n, 200_000)
            csc = workload(3)
            times, evals, kind = run_cpu_steps(n_cpu, 2, 1, csc)
            line["cpu_baseline"] = {"value": sum(evals) / sum(times), "unit": UNIT, "cores": 1, "kind": kind,
                                    "sample": "same workload at %d particles, 2 timed full steps after 1 warm-up, 1 host thread "
                                              "(the reference is single-threaded)" % n_cpu,
                                    "ms_per_step": 1e3 * sum(times) / len(times), "host_cores_available": os.cpu_count()}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------------------------
# NS leg: the north-star formulation (likelihood field, Philox motion noise, fixed-point systematic resampling), sharded
# across the GPUs of the job. Three collectives per step go through torch.distributed/NCCL (max of local maxima,
# all-gather of local Q32 totals, barrier); resampled particles are stored by the resampling kernel straight into the
# owning shard's memory (CUDA IPC peer mappings over NVLink), so no separate rebalance pass exists.
# ------------------------------------------------------------------------------------------------------------------
def ns_workload(cells, n_beams, n_scans, seed):
    from montecarlolocalisation_b200 import synth
    occ = synth.maze_occupancy(cells, seed)
    res = 0.1
    # a pose in the middle of a cell near the map centre
    cx = (cells // 2) * 8 * res + 0.45
    pose = (cx, cx, 0.3)
    scans = [synth.make_scan(occ, res, pose, n_beams, 1000 * seed + i) for i in range(n_scans)]
    return occ, scans


def ns_leg(args, torch, dist, rank, world, local, cells, per_gpu, n_beams, label, K, W, key=""):
    import montecarlolocalisation_b200 as m
    from montecarlolocalisation_b200 import NsShard
    n_global = per_gpu * world
    n_scans = 4
    if cells == 6:
        from scenario import Scenario
        sc = Scenario(n_scans, n_beams=n_beams)
        occ, scans = sc.occ, sc.scans
    else:
        occ, scans = ns_workload(cells, n_beams, n_scans, seed=4)
    shard = NsShard(rank, world, n_global, device=local, max_particles=0, seed=0xABCDEF)
    shard.pf.setMap(occ, np.float32(0.1))
    if world > 1:
        # the engine's own NCCL communicator (collectives enqueued on the engine's stream) + CUDA-IPC peer mappings
        ids = [shard.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        shard.comm_init(ids[0])
    shard.pf.sampleParticles(n_global)
    for i, sca in enumerate(scans):
        shard.pf.stageScan(i, sca["ranges"], sca["angle_min"], sca["angle_inc"], sca["range_min"], sca["range_max"])
    stream = torch.cuda.ExternalStream(shard.pf.stream(), device=local)
    pinned = [torch.from_numpy(np.ascontiguousarray(sca["ranges"])).pin_memory() for sca in scans]
    valid_beams = [int((np.isfinite(sca["ranges"]) & (sca["ranges"] >= sca["range_min"]) & (sca["ranges"] <= sca["range_max"]) &
                        (sca["ranges"] < 5.6)).sum()) for sca in scans]
    motion = (0.01, 0.02, -0.005)

    def step(i, e2e):
        """One whole filter step enqueued by the engine (mcl_ns_step): predict -> likelihood field -> all-reduce(max) ->
        Q32 weights + prefix -> all-gather(totals) -> device-side plan -> resample into the owning shards -> barrier.
        e2e: the scan comes from (pinned) host memory and the weighted-mean pose is read back, every step."""
        slot = i % n_scans
        if e2e:
            sca = dict(scans[slot])
            sca["ranges"] = pinned[slot].numpy()
            return shard.step(motion, scan=sca, want_pose=True)
        return shard.step(motion, slot=slot)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(e2e, first):
        for i in range(first, first + W):
            step(i, e2e)
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
        barrier()
        w0 = time.perf_counter()
        for k in range(K):
            ev[k][0].record(stream)
            step(first + W + k, e2e)
            ev[k][1].record(stream)
        barrier()
        wall = time.perf_counter() - w0
        return [a.elapsed_time(b) for a, b in ev], wall

    l0 = shard.pf.kernelLaunches()
    ms_res, wall_res = timed(False, 0)
    launches = shard.pf.kernelLaunches() - l0
    ms_e2e, wall_e2e = timed(True, W + K)
    shard.pf.profileEnable(True)
    for k in range(K):
        step(2 * (W + K) + k, False)
    prof = shard.pf.profileRead()
    shard.pf.profileEnable(False)
    # kidnapped-robot case: freshly uniform particles (no spatial locality in the field gathers)
    steady_form = shard.field_form()
    for k in range(6):
        if k == 2:          # two untimed launches first: fields larger than L2 settle on their field form by measurement
            shard.pf.profileEnable(True)
        shard.pf.sampleParticles(n_global)
        shard.pf.updateParticlePos(*motion)
        shard.update_local_staged(0)
    uni = shard.pf.profileRead().get("k_ns_update", (0.0, 1))
    shard.pf.profileEnable(False)
    uniform_form = shard.field_form()
    uniform_ms = uni[0] / uni[1]
    t_res, t_e2e = sum(ms_res) * 1e-3, sum(ms_e2e) * 1e-3
    if world > 1:
        tt = torch.tensor([t_res, t_e2e], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_res, t_e2e = tt.tolist()
    evals_res = sum(n_global * valid_beams[i % n_scans] for i in range(W, W + K))
    evals_e2e = sum(n_global * valid_beams[i % n_scans] for i in range(2 * W + K, 2 * (W + K)))
    peak, peak_src = peaks()
    kernels = {}
    tot_ms = sum(v[0] for v in prof.values())
    nb = valid_beams[0]
    algo = {"k_ns_update": per_gpu * (20 + 4 * nb), "k_ns_predict": per_gpu * 32, "k_ns_weights_sum": per_gpu * 4,
            "k_ns_weights_scan": per_gpu * 12, "k_ns_resample": per_gpu * 44}
    for name, (msv, cnt) in sorted(prof.items(), key=lambda kv: -kv[1][0]):
        a = algo.get(name)
        kernels[name] = {"ms_per_launch": msv / cnt, "launches": cnt, "share": msv / tot_ms, "algo_bytes": a,
                         "gbs": (a / (msv / cnt * 1e-3) / 1e9) if a else None}
    top = next(iter(kernels))
    field_bytes = occ.size * 4
    out = {
        "label": label, "value": evals_res / t_res, "unit": UNIT, "ms_per_step": 1e3 * t_res / K, "steps_per_s": K / t_res,
        "e2e": {"value": evals_e2e / t_e2e, "unit": UNIT, "ms_per_step": 1e3 * t_e2e / K, "h2d_bytes_per_step": int(scans[0]["ranges"].nbytes) + 28,
                "d2h_bytes_per_step": 40, "wall_ms_per_step": 1e3 * wall_e2e / K},
        "config": {"workload": "%s: %dx%d occupancy grid, %d particles per GPU x %d GPU(s) = %d, %d-beam scan (%d valid beams scored per particle), "
                               "NS mode: Philox motion noise -> likelihood field -> weighted-mean pose -> Q32 systematic resampling" % (
                                   label, occ.shape[1], occ.shape[0], per_gpu, world, n_global, n_beams, nb),
                   "field": "%d KiB log-likelihood field, %s" % (field_bytes // 1024, {"smem-f32": "TMA-staged into shared memory", "global-f32": "fp32, gathered through L1/L2",
                                                                       "global-u8": "as one-byte codes (%d KiB) gathered through L1/L2 + shared-memory code table" % (field_bytes // 4096)}[steady_form]),
                   "collectives": "none (1 GPU)" if world == 1 else (
                       "peer-memory mailboxes, no NCCL on the data path: every shard stores {payload, tag} into the other shards' mailboxes over NVLink (CUDA IPC) and polls its own; three 32-thread kernels on the filter's stream, no host round trip: all-reduce(max), all-gather(Q32 totals) + all-reduce(pose) in one exchange fused with the resampling plan, closing barrier; resampled particles stored into peer shards over NVLink"
                       if shard.exchange_used() == "peer" else
                       "engine-enqueued NCCL on the filter's stream, no host round trip: all-reduce(max), all-gather(Q32 totals), all-reduce(pose), closing all-reduce as barrier; resampled particles stored into peer shards over NVLink (CUDA IPC)"),
                   "l2": "per-GPU working set %.0f MB exceeds or displaces L2 between steps" % (per_gpu * 48 / 1e6)},
        "gpu_launches": launches, "scaling": "weak",
        "roofline": {"bound": "hbm", "kernel": top, "achieved": kernels[top]["gbs"], "peak": peak, "unit": "GB/s",
                     "frac": (kernels[top]["gbs"] / peak) if kernels[top]["gbs"] else None,
                     "traffic": traffic_for(top + "@" + key, per_gpu), "peak_source": peak_src,
                     "share_of_step": kernels[top]["share"],
                     "note": "algorithmic bytes = 20 B/particle + 4 B per scored beam (table gather); the gathers are served by shared memory or L2, not HBM"},
        "kernels": kernels,
        "uniform_particles": {"k_ns_update_ms": uniform_ms, "field_form": uniform_form, "evals_per_s_per_gpu": per_gpu * valid_beams[0] / (uniform_ms * 1e-3),
                              "note": "sensor-model kernel alone on freshly uniform particles (kidnapped robot): worst case for gather locality"},
        "gather_microbench_reads_per_s": {"shared_memory_table": shard.pf.benchGather(0, min(field_bytes, 190 * 1024)),
                                          "global_table_of_field_size": shard.pf.benchGather(1, field_bytes)},
    }
    del shard
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--particles", type=int, default=N_PARTICLES)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--separate-calls", action="store_true", help="REF loop through the four per-function calls instead of mcl_step")
    ap.add_argument("--no-ns", action="store_true", help="skip the NS (north-star) leg")
    ap.add_argument("--no-ns-large", action="store_true", help="skip the 8193x8193 / 1080-beam NS case")
    ap.add_argument("--ns-cells", type=int, default=512, help="NS leg: maze cells per side (512 -> 4097x4097 grid)")
    ap.add_argument("--ns-particles", type=int, default=12_500_000, help="NS leg: particles per GPU")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3          # timing rule: at least 3 warm-up steps
    if args.impl == "reference":
        reference_arm(args)
    else:
        ours(args)


if __name__ == "__main__":
    main()
