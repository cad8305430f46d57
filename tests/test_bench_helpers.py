"""CPU: the quantities bench.py turns into `value` are what it says they are. The number of scored beams per scan (evals =
particles x scored beams) against the oracle's own scan filter; the cross-N parity hash adds over shards; the two arms print
the same `config` for the same workload; the reference arm runs (tiny size) and prints the contract's keys."""
import io
import json
import os
import sys
from contextlib import redirect_stdout

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from oracle.pyoracle import Oracle, Scan  # noqa: E402


def test_scored_beams_match_the_oracles_filter():
    sc = bench.workload(3)
    for scan in list(sc.scans) + [bench.robot_scan(sc.occ, sc.truth[0], 501)]:
        r, a = Oracle.filter_scan(Scan(**scan))
        assert bench.used_beams(scan) == len(range(0, len(r), 20))
    assert bench.used_beams(sc.scans[0]) == 12
    assert bench.used_beams(bench.robot_scan(sc.occ, sc.truth[0], 501)) == 35


def test_state_hash_adds_over_shards_and_sees_every_field():
    rng = np.random.default_rng(3)
    n = 10_000
    P = rng.standard_normal((n, 4)).astype(np.float32)
    anc = rng.integers(-1, n, n).astype(np.int32)
    whole = bench.state_hash(P, anc, 0)
    for cuts in ([0, n], [0, 1, n], [0, 2500, 5000, 7500, n], [0, 3333, 3334, 9999, n]):
        parts = sum(bench.state_hash(P[a:b], anc[a:b], a) for a, b in zip(cuts[:-1], cuts[1:])) % (1 << 64)
        assert parts == whole, cuts
    Q = P.copy(); Q[1234, 2] = np.nextafter(Q[1234, 2], np.float32(9))
    assert bench.state_hash(Q, anc, 0) != whole
    anc2 = anc.copy(); anc2[77] += 1
    assert bench.state_hash(P, anc2, 0) != whole
    perm = rng.permutation(n)                               # same particles in other slots: a different state
    assert bench.state_hash(P[perm], anc[perm], 0) != whole


def test_arms_name_the_same_config():
    a = bench.config_ref(1_000_000)
    assert a == bench.config_ref(1_000_000) and "configs[1]" in a["workload"] and a["particles"] == 1_000_000
    s = bench.config_ns_strong(100_000_000)
    assert "configs[3]" in s["workload"] and json.dumps(s) == json.dumps(bench.config_ns_strong(100_000_000))


def test_reference_arm_line_has_the_contract_keys():
    class A:
        gpus = 1; steps = 2; warmup = 1; particles = 2000
    buf = io.StringIO()
    with redirect_stdout(buf):
        bench.reference_arm(A)
    line = json.loads(buf.getvalue().strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["n_gpus"] == 1 and line["steps"] == 2 and line["warmup"] == 1
    assert line["config"] == bench.config_ref(2000)
    assert line["unit"] == bench.UNIT and line["metric"] == bench.METRIC and line["higher_is_better"] is True
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["cores"] == 1
    assert line["e2e"] == {"value": line["value"], "unit": bench.UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # 2000 particles x 12 scored beams per step over the measured time
    assert line["value"] > 1e5 and abs(line["value"] * line["ms_per_step"] * 1e-3 - 2000 * 12) < 1e-6 * 2000 * 12
