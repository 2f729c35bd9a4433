#!/usr/bin/env python
"""Headline counters + hottest SASS lines (by warp-stall samples) of every kernel in an .ncu-rep:  python tools/ncu_hot.py file.ncu-rep [top]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
keys = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "l1tex__m_xbar2l1tex_read_bytes.sum",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"]
names = []
for r in rows[2:]:
    names.append(r[hdr.index("Kernel Name")])
    print("====", r[hdr.index("ID")], r[hdr.index("Kernel Name")][:80])
    for k in keys:
        if k in hdr:
            print(f"  {k} = {r[hdr.index(k)]} {units[hdr.index(k)]}")
    st = [(float(r[i]), h) for i, h in enumerate(hdr) if "issue_stalled" in h and h.endswith("per_issue_active.ratio") and r[i]]
    print("  stalls/issue:", ", ".join(f"{h.split('issue_stalled_')[1].split('_per_')[0]}={v:.2f}" for v, h in sorted(st, reverse=True)[:7]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
blocks = [i for i, r in enumerate(rows) if "Source" in r and "Instructions Executed" in r]
for bi, i in enumerate(blocks):
    hdr = rows[i]
    end = blocks[bi + 1] if bi + 1 < len(blocks) else len(rows)
    isrc, iex = hdr.index("Source"), hdr.index("Instructions Executed")
    js = [j for j, h in enumerate(hdr) if "Samples" in h][0]
    data = [r for r in rows[i + 1:end] if len(r) > iex and r[iex].isdigit()]
    tot = sum(int(r[js]) for r in data) or 1
    print(f"---- kernel {bi} ({names[bi][:60] if bi < len(names) else ''}): {tot} stall samples, {len(data)} SASS instructions, {sum(int(r[iex]) for r in data)} warp instr")
    top = sorted(range(len(data)), key=lambda k: -int(data[k][js]))[:top_n]
    for k in sorted(top):
        r = data[k]
        print(f"  {k:5d} {100 * int(r[js]) / tot:5.1f}%  x{r[iex]:>9}  {r[isrc][:100]}")
