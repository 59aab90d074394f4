#!/usr/bin/env python
"""Hot spots of one kernel from `ncu -i X.ncu-rep --page source --csv`: the SASS instructions with the most stall samples /
executed instructions, with their dominant stall reason, plus totals per access kind.  Usage: ncu_hot.py file.csv [top]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = rows[1]
ix = {n: i for i, n in enumerate(hdr)}
body = [r for r in rows[2:] if len(r) == len(hdr)]
def f(r, n):
    try:
        return float(r[ix[n]])
    except ValueError:
        return 0.0
tot_s = sum(f(r, "# Samples") for r in body)
tot_i = sum(f(r, "Instructions Executed") for r in body)
print(f"{len(body)} SASS instructions, {tot_s:.0f} samples, {tot_i:.0f} warp-instructions executed")
stalls = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
agg = {n: sum(f(r, n) for r in body) for n in stalls}
print("stall totals:", ", ".join(f"{k[6:]} {v / tot_s:.1%}" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
print("L1 tag requests (global):", sum(f(r, "L1 Tag Requests Global") for r in body), " L2 theoretical sectors:", sum(f(r, "L2 Theoretical Sectors Global") for r in body))
print("\n-- top by samples --")
for i, r in sorted(enumerate(body), key=lambda ir: -f(ir[1], "# Samples"))[:top]:
    dom = max(stalls, key=lambda n: f(r, n))
    print(f"{i:5d} {f(r, '# Samples') / tot_s:6.2%} exec {f(r, 'Instructions Executed'):10.0f} thr {f(r, 'Avg. Threads Executed'):4.1f} tag {f(r, 'L1 Tag Requests Global'):9.0f}  {dom[6:]:10s} {r[ix['Source']].strip()[:90]}")
print("\n-- memory instructions by L1 tag requests --")
for i, r in sorted(enumerate(body), key=lambda ir: -f(ir[1], "L1 Tag Requests Global"))[:24]:
    if f(r, "L1 Tag Requests Global") == 0:
        break
    print(f"{i:5d} tag {f(r, 'L1 Tag Requests Global'):10.0f} sectors {f(r, 'L2 Theoretical Sectors Global'):10.0f} exec {f(r, 'Instructions Executed'):9.0f} thr {f(r, 'Avg. Threads Executed'):4.1f}  {r[ix['Source']].strip()[:90]}")
