#!/usr/bin/env python
"""Summarise an .ncu-rep (one kernel launch) into the text kept under profiles/.
usage: tools/summarize_profile.py gpurun_out/prof.ncu-rep > profiles/xyz.txt"""
import csv, io, subprocess, sys, re
from collections import Counter

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
want = [
    "Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size",
    "launch__registers_per_thread", "launch__shared_mem_per_block", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum",
    "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second",
]
print(f"# {rep}")
for k in want:
    if k in d:
        print(f"{k:75s} {d[k][0]} {d[k][1]}")
print("# warp issue stall reasons (cycles per issued instruction)")
for h in hdr:
    m = re.match(r"smsp__average_warps_issue_stalled_(\w+)_per_issue_active.ratio", h)
    if m and h in d:
        try:
            v = float(d[h][0])
        except ValueError:
            continue
        if v >= 0.05:
            print(f"  {m.group(1):28s} {v:.3f}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
if len(rows) > 2:
    h2 = rows[1]
    iS, iE = h2.index("Source"), h2.index("Instructions Executed")
    ops, tot = Counter(), 0
    for r in rows[2:]:
        if len(r) <= iE or not r[iE]:
            continue
        m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)", r[iS])
        if m:
            ops[m.group(2)] += int(r[iE]); tot += int(r[iE])
    print(f"# dynamic SASS opcode mix (warp instructions executed: {tot})")
    for k, v in ops.most_common(16):
        print(f"  {k:10s} {v:14d} {100.0 * v / tot:5.1f}%")
