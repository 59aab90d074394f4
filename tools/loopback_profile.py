#!/usr/bin/env python
"""Per-GPU work of a W-rank run on ONE GPU (serfsim_comm_loopback): the sharded tick kernel with (W-1)/W of its sends staged and
stored into windows, the publish kernel and the drain kernel folding W-1 windows — per-tick device times, split at the end of the
tick kernel (SERFSIM_XTIMING).  The handle exchanges with itself, so the simulation results are meaningless; the load is real.

    python tools/loopback_profile.py --world 8                  # shard 0 of the 10 M-node bench workload (1.25 M nodes)
    ncu --set full -k regex:tick_kernel --launch-skip 13 --launch-count 1 ... python tools/loopback_profile.py --world 8 --runs 1
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from serf_b200 import GossipSim, scenarios  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--nodes", type=int, default=10_000_000)
ap.add_argument("--world", type=int, default=8)
ap.add_argument("--slots", type=int, default=1)
ap.add_argument("--runs", type=int, default=2)
ap.add_argument("--fail", action="store_true", help="the leave + fail workload (2 tracked subjects)")
ap.add_argument("--out", default=None)
a = ap.parse_args()
sc = scenarios.dissemination_storm(a.nodes, 16, 4, slots=max(2, a.slots) if a.fail else a.slots, seed=1, with_fail=a.fail)
g = sc.build(lambda n, s, **kw: GossipSim(n, s, **kw), rank=0, world_size=a.world)
g.connect_loopback()
for run in range(a.runs):
    g.reset(1); sc.schedule(g)
    g.set_tick_timing(run == a.runs - 1)
    ticks, ok = g.run_until_converged(min(sc.max_ticks, 400))
ms = g.tick_times_ms()
tr = g.tick_trace()
rows = []
for t in range(len(ms)):
    rows.append({"tick": t, "edge_updates_x_world": int(tr["edge_updates"][t]), "ms": float(ms[t])})
    if t < 60:
        print(f"tick {t:3d}  global-row eu {int(tr['edge_updates'][t]):10d}  {ms[t]*1e3:9.1f} us")
print(f"world {a.world} loopback, shard of {g.count} nodes: {len(ms)} ticks, {ms.sum():.3f} ms device time")
if a.out:
    json.dump({"world": a.world, "nodes_local": int(g.count), "rows": rows, "kernel_ms": float(ms.sum())}, open(a.out, "w"), indent=1)
g.close()
