"""Summarise an .ncu-rep into a small text file for profiles/: python tools/ncu_summary.py <rep> <kernel-substr> <out.txt> "<title>" """
import collections, csv, subprocess, sys
rep, kern, out, title = sys.argv[1:5]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
d = next(r for r in rows[2:] if kern in r[hdr.index("Kernel Name")])
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_pipe_lsu_mem_global_op_ld_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.max",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio"]
lines = ["# " + title, "# kernel: " + d[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "", "# ncu --set full --clock-control none (numbers under ncu are not bench values)"]
for k in want:
    if k in hdr:
        lines.append("%-88s %s %s" % (k, d[hdr.index(k)], units[hdr.index(k)]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", "regex:" + kern.split("<")[0], "-c", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hi = [i for i, r in enumerate(rows) if "Source" in r][0]
h = rows[hi]
iS, iE = h.index("Source"), h.index("Instructions Executed")
ops, tot = collections.Counter(), 0
for r in rows[hi + 1:]:
    if "Source" in r:
        break
    try:
        e = int(r[iE])
    except Exception:
        continue
    op = (r[iS].split()[1] if r[iS].startswith("@") else r[iS].split()[0]).split(".")[0]
    ops[op] += e; tot += e
lines.append("# executed warp-instructions by SASS opcode")
for op, c in ops.most_common(14):
    lines.append("%-10s %12d %5.1f%%" % (op, c, 100.0 * c / tot))
open(out, "w").write("\n".join(lines) + "\n")
print("\n".join(lines))
