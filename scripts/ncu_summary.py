"""ncu report -> the text summary kept under profiles/ (raw metrics of one launch + executed warp instructions by opcode).
usage: ncu_summary.py <report.ncu-rep> <n_tasks> <antidiagonals per task> <warps per task> > profiles/ncu_dpx_fill_<tag>_summary.txt"""
import collections, csv, io, subprocess, sys
rep, n_tasks, n_diag, nw = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__cycles_elapsed.max"]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
print("kernel:", vals[hdr.index("Kernel Name")])
for i, h in enumerate(hdr):
    if h in WANT or (h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio") and float(vals[i].replace(",", "") or 0) >= 0.02):
        print("%-90s %-16s %s" % (h, units[i], vals[i]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
steps = n_tasks * n_diag * nw
ops, tot, samples, bar = collections.Counter(), 0, 0, 0
for r in rows[2:]:
    ex = int(r[ix["Instructions Executed"]]); s = int(r[ix["# Samples"]])
    op = [o for o in r[ix["Source"]].split() if not o.startswith("@")][0]
    ops[op] += ex; tot += ex; samples += s; bar += int(r[ix["stall_barrier"]])
print("\n--- executed warp instructions by opcode (ncu source page), per warp-antidiagonal = / (%d tasks x %d antidiagonals x %d warps)" % (n_tasks, n_diag, nw))
print("total warp instr %d per warp-diag %.1f ; warp-state samples %d, of which at the CTA barrier %d (%.1f %%)" % (tot, tot / steps, samples, bar, 100.0 * bar / max(samples, 1)))
for op, v in ops.most_common(32):
    print("%6.2f%% %8.1f %s" % (100.0 * v / tot, v / steps, op))
