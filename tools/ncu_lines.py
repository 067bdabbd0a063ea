#!/usr/bin/env python
"""Per-source-line instruction / stall-sample totals from `ncu --page source --csv --print-source cuda,sass`."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur = None
out = []
tot_i = tot_s = 0
for r in rows:
    if r and r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if len(r) < 8 or r[0] in ("Line No", "Function Name", ""):
        continue
    try:
        line = int(r[0]); samples = int(r[4]); inst = int(r[7])
    except ValueError:
        continue
    out.append((inst, samples, cur, line, r[1].strip()[:110]))
    tot_i += inst; tot_s += samples
print(f"total inst {tot_i} samples {tot_s}")
print("== by instructions")
for inst, s, f, l, src in sorted(out, reverse=True)[:top]:
    print(f"{inst:>12} {100*inst/tot_i:5.1f}% smp {100*s/max(tot_s,1):5.1f}% {f}:{l}: {src}")
print("== by samples")
for inst, s, f, l, src in sorted(out, key=lambda t: -t[1])[:top]:
    print(f"{inst:>12} {100*inst/tot_i:5.1f}% smp {100*s/max(tot_s,1):5.1f}% {f}:{l}: {src}")
