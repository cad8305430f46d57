"""Which resource binds a kernel, from an `ncu --set full` capture: writes/updates profiles/r2_counters.json, which bench.py
reads for its `roofline` record.

    python tools/ncu_counters.py <report.ncu-rep> <kernel-name substring> <key> [particles] [note] [index among the matching launches]

key: "<kernel>" or "<kernel>@<bench leg>" (e.g. k_ns_update@grid4096). The binding resource is the busiest of: instruction
issue (smsp__issue_active), the FMA-heavy pipe, the FP64 pipe, the LSU pipe (shared-memory wavefronts) and DRAM; `frac` is that
counter's fraction of its peak, `traffic` = dram__bytes_read.sum + dram__bytes_write.sum of the launch.
"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, kern, key = sys.argv[1:4]
particles = int(sys.argv[4]) if len(sys.argv) > 4 and sys.argv[4].isdigit() else None
note = sys.argv[5] if len(sys.argv) > 5 and sys.argv[5] else None
which = int(sys.argv[6]) if len(sys.argv) > 6 else -1          # which of the matching launches (default: the last)
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
cands = [r for r in rows[2:] if kern in r[hdr.index("Kernel Name")]]
if not cands:
    raise SystemExit("no kernel matching %r in %s" % (kern, rep))
d = cands[which]


def val(name):
    if name not in hdr:
        return None
    try:
        v = float(d[hdr.index(name)].replace(",", ""))
    except ValueError:
        return None
    u = units[hdr.index(name)]
    mult = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "us": 1.0, "ms": 1e3, "ns": 1e-3}.get(u, 1.0)
    return v * mult


resources = {
    "issue": "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "fmaheavy": "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "fp64": "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "alu": "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "lsu": "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "smem": "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l2": "lts__t_sectors.avg.pct_of_peak_sustained_elapsed",
    "hbm": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
}
seen = {k: val(v) for k, v in resources.items()}
seen = {k: v for k, v in seen.items() if v is not None}
bound = max(seen, key=seen.get)
rd, wr = val("dram__bytes_read.sum"), val("dram__bytes_write.sum")
rec = {"bound": bound, "frac": seen[bound] / 100.0, "counter": resources[bound], "file": "profiles/" + os.path.basename(rep).replace(".ncu-rep", "_ncu.txt"),
       "all_pct": {k: round(v, 1) for k, v in sorted(seen.items(), key=lambda kv: -kv[1])},
       "traffic": (rd + wr) if rd is not None and wr is not None else None, "duration_us_under_ncu": val("gpu__time_duration.sum"),
       "kernel": d[hdr.index("Kernel Name")][:120], "particles": particles, "note": note}
path = os.path.join(ROOT, "profiles", "r2_counters.json")
table = json.load(open(path)) if os.path.exists(path) else {}
table[key] = rec
json.dump(table, open(path, "w"), indent=1, sort_keys=True)
print(key, json.dumps(rec, indent=1))
