#!/usr/bin/env python
"""Opcode mix + headline counters of an .ncu-rep (needs ncu on PATH): python tools/ncu_mix.py file.ncu-rep"""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
keys = ["Kernel Name", "gpu__time_duration.sum", "smsp__inst_executed.sum", "sm__cycles_elapsed.max", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active"]
for r in rows[2:]:
    for k in keys:
        if k in hdr:
            i = hdr.index(k)
            print(f"{k} = {r[i]} {units[i]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
for i, r in enumerate(rows):
    if "Source" in r and "Instructions Executed" in r:
        hdr, start = r, i + 1
        break
isrc, iex = hdr.index("Source"), hdr.index("Instructions Executed")
data = [r for r in rows[start:] if len(r) > iex and r[iex].isdigit()]
tot = sum(int(r[iex]) for r in data)
by = collections.Counter()
for r in data:
    op = r[isrc].strip().split()
    if op and op[0].startswith("@"):
        op = op[1:]
    by[op[0].split(".")[0] if op else "?"] += int(r[iex])
print("total warp instr", tot, "static", len(data))
print("  ".join(f"{k}:{100 * v / tot:.1f}%" for k, v in by.most_common(22)))
