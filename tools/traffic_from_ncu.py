#!/usr/bin/env python
"""profiles/<round>_traffic.json (what bench.py's roofline.traffic reads) from the per-launch ncu CSV of tools/r2_profile.sh step 2.

  python tools/traffic_from_ncu.py gpurun_out/r2_traffic_ncu.csv profiles/r2_traffic.json [edge_updates_of_the_run p_dirty fanout]
"""
import csv
import json
import sys

src, dst = sys.argv[1], sys.argv[2]
eu = float(sys.argv[3]) if len(sys.argv) > 3 else 620985084.0
p_dirty = float(sys.argv[4]) if len(sys.argv) > 4 else 0.0322
fanout = float(sys.argv[5]) if len(sys.argv) > 5 else 4.0
rows = [r for r in csv.reader(l for l in open(src) if l.startswith('"')) if len(r) > 14]
hdr, rows = rows[0], rows[1:]
ki, mi, vi, idi = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
tot = {"dram__bytes_read.sum": 0.0, "dram__bytes_write.sum": 0.0, "gpu__time_duration.sum": 0.0}
ids, kernel = set(), None
for r in rows:
    if "tick_kernel" not in r[ki] or r[mi] not in tot:
        continue
    tot[r[mi]] += float(r[vi].replace(",", ""))
    ids.add(r[idi]); kernel = r[ki]
b_edge = 4 + 32 / fanout + 32 + 32 * p_dirty
alg = eu * b_edge
out = {"source": f"ncu per-launch DRAM bytes of every tick_kernel launch of one run ({src})", "kernel": kernel, "launches": len(ids),
       "dram_bytes_read": tot["dram__bytes_read.sum"], "dram_bytes_write": tot["dram__bytes_write.sum"],
       "dram_bytes_per_launch": (tot["dram__bytes_read.sum"] + tot["dram__bytes_write.sum"]) / max(1, len(ids)),
       "ncu_time_ns_total": tot["gpu__time_duration.sum"], "algorithmic_bytes_total": alg,
       "traffic_over_algorithmic": (tot["dram__bytes_read.sum"] + tot["dram__bytes_write.sum"]) / alg}
json.dump(out, open(dst, "w"), indent=1)
print(json.dumps(out, indent=1))
