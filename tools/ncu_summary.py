#!/usr/bin/env python
"""Summarise an .ncu-rep (raw + source pages) into text: key metrics, opcode mix, hottest SASS."""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 12
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "smsp__inst_executed.sum",
        "sm__inst_executed.avg.per_cycle_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__grid_size", "launch__block_size",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "lts__t_bytes.sum", "sm__cycles_elapsed.max"]
for vals in rows[2:]:
    print("== kernel:", vals[hdr.index("Kernel Name")][:100])
    for i, h in enumerate(hdr):
        if h in want:
            print(f"  {h:70s} {units[i]:>12s} {vals[i]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]
idx = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
S, I, SRC = idx["# Samples"], idx["Instructions Executed"], idx["Source"]
tot = sum(int(r[I] or 0) for r in data)
samp = sum(int(r[S] or 0) for r in data)
print(f"total warp-instructions {tot}  stall samples {samp}")
byop, bys = collections.Counter(), collections.Counter()
for r in data:
    s = r[SRC].strip()
    parts = s.split()
    op = parts[1] if parts and parts[0].startswith("@") and len(parts) > 1 else (parts[0] if parts else "")
    byop[op.split(".")[0]] += int(r[I] or 0)
    bys[op.split(".")[0]] += int(r[S] or 0)
print("opcode mix (inst% / stall-sample%):")
for op, c in byop.most_common(22):
    print(f"  {op:12s} {c / tot * 100:5.1f}%  {bys[op] / max(samp, 1) * 100:5.1f}%")
print("hottest SASS by stall samples:")
top = sorted(range(len(data)), key=lambda i: -int(data[i][S] or 0))[:ntop]
for t in top:
    print(f" -- line {t}: samples {data[t][S]} inst {data[t][I]}")
    for j in range(max(0, t - 4), min(len(data), t + 2)):
        print(f"      {j:6d} smp={data[j][S]:>7s} inst={data[j][I]:>10s}  {data[j][SRC][:100]}")
