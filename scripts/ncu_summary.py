"""ncu report -> the text summary kept under profiles/ (raw metrics of one launch + executed warp instructions by opcode).
usage: ncu_summary.py <report.ncu-rep> <n_tasks> <antidiagonals per task> <warps per task> [cells of the launch] [out.json] > profiles/ncu_dpx_fill_<tag>_summary.txt
(cells = GCUPS x ms x 1e6 of the same command run without ncu; with it the JSON holds thread-instructions per DP cell, what bench.py's roofline.issue uses)"""
import collections, csv, io, json, subprocess, sys
rep, n_tasks, n_diag, nw = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
cells = float(sys.argv[5]) if len(sys.argv) > 5 else None
out_json = sys.argv[6] if len(sys.argv) > 6 else None
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
num = lambda name: float(vals[hdr.index(name)].replace(",", ""))
if cells and out_json:
    ti = num("smsp__inst_executed.sum") * num("smsp__thread_inst_executed_per_inst_executed.ratio")
    json.dump({"kernel": vals[hdr.index("Kernel Name")], "report": rep.split("/")[-1], "cells": cells, "warp_instructions": num("smsp__inst_executed.sum"),
               "thread_instructions_per_cell": ti / cells, "issue_active_pct": num("smsp__issue_active.avg.pct_of_peak_sustained_active"),
               "alu_pipe_pct": num("sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active"), "fma_pipe_pct": num("sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active"),
               "barrier_stall_per_issue": num("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"), "registers_per_thread": num("launch__registers_per_thread"),
               "note": "one launch under ncu --set full (kbench uniform probe); cells from the same command without ncu"}, open(out_json, "w"), indent=1)
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
