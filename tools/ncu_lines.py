#!/usr/bin/env python
"""Per-CUDA-source-line totals from `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass`: executed warp instructions and
stall samples of every source line (the SASS rows that follow a source row belong to it).  Usage: ncu_lines.py file.csv [top]"""
import csv
import sys
from collections import defaultdict

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 45
cur_file, hdr, ix = None, None, None
inst, samp, src = defaultdict(float), defaultdict(float), {}
line = None
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]; continue
    if len(r) == 2:
        continue
    if r and r[0] == "Line No":
        hdr = r; ix = {n: i for i, n in enumerate(hdr)}; continue
    if hdr is None or len(r) != len(hdr):
        continue
    if r[0].strip():                                   # a source row
        line = (cur_file, int(r[0])); src[line] = r[1].strip(); continue
    def f(n):
        try:
            return float(r[ix[n]])
        except (ValueError, KeyError):
            return 0.0
    inst[line] += f("Instructions Executed"); samp[line] += f("# Samples")
ti, ts = sum(inst.values()), sum(samp.values())
print(f"{ti:.0f} warp instructions, {ts:.0f} samples")
byfile = defaultdict(float)
for k, v in inst.items():
    byfile[k[0]] += v
print("by file:", ", ".join(f"{k} {v / ti:.1%}" for k, v in sorted(byfile.items(), key=lambda kv: -kv[1])))
print("\n-- top source lines by executed warp instructions --")
for k, v in sorted(inst.items(), key=lambda kv: -kv[1])[:top]:
    print(f"{k[0]:18s}:{k[1]:5d} {v / ti:6.2%} inst {samp[k] / ts:6.2%} samples  {src[k][:110]}")
print("\n-- top source lines by stall samples --")
for k, v in sorted(samp.items(), key=lambda kv: -kv[1])[:25]:
    print(f"{k[0]:18s}:{k[1]:5d} {inst[k] / ti:6.2%} inst {v / ts:6.2%} samples  {src[k][:110]}")
