"""One line per captured launch with the metrics that matter: python tools/ncu_table.py <rep>"""
import csv, subprocess, sys
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[0]
def col(name): return hdr.index(name) if name in hdr else None
cols = [("Kernel Name", 26, "kernel"), ("gpu__time_duration.sum", 9, "time"), ("dram__bytes_read.sum", 9, "rdMB"), ("dram__bytes_write.sum", 9, "wrMB"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", 6, "dram%"), ("smsp__issue_active.avg.pct_of_peak_sustained_active", 6, "issue%"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", 6, "occ%"), ("smsp__inst_executed.sum", 11, "warp-inst"),
        ("launch__registers_per_thread", 4, "regs"), ("launch__grid_size", 6, "grid"), ("launch__block_size", 5, "blk"),
        ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", 6, "fp64%"), ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", 6, "fma%"),
        ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", 6, "alu%"), ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", 6, "lsu%"),
        ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", 6, "xu%"), ("l1tex__t_sector_hit_rate.pct", 6, "l1hit"), ("lts__t_sector_hit_rate.pct", 6, "l2hit"),
        ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", 6, "st_lg"),
        ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", 6, "st_sh"),
        ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", 6, "st_bar"),
        ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", 6, "st_mth"),
        ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", 6, "st_wt")]
print(" ".join(("%-" + str(w) + "s") % t for _, w, t in cols))
units = rows[1]
for r in rows[2:]:
    out = []
    for name, w, t in cols:
        i = col(name)
        v = r[i] if i is not None else ""
        if name == "Kernel Name":
            v = v.split("(")[0].replace("void ", "")[:w]
        elif name == "gpu__time_duration.sum":
            v = "%.1fus" % (float(v) * {"ms": 1e3, "us": 1.0, "ns": 1e-3, "s": 1e6}[units[i]])
        elif name.startswith("dram__bytes"):
            v = "%.1f" % (float(v) * {"Mbyte": 1.0, "Kbyte": 1e-3, "byte": 1e-6, "Gbyte": 1e3}[units[i]])
        else:
            try:
                f = float(v); v = ("%d" % f) if f == int(f) and abs(f) > 100 else ("%.1f" % f)
            except ValueError:
                pass
        out.append(("%-" + str(w) + "s") % v[:w])
    print(" ".join(out))
