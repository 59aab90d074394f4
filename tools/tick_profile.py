#!/usr/bin/env python
"""Per-tick device time of the bench workload: edge-updates, ms and algorithmic GB/s of every tick launch."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from serf_b200 import GossipSim, scenarios  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--nodes", type=int, default=10_000_000)
ap.add_argument("--fanout", type=int, default=4)
ap.add_argument("--slots", type=int, default=1)
ap.add_argument("--degree", type=int, default=16)
ap.add_argument("--waves", type=int, default=1)
ap.add_argument("--runs", type=int, default=2)
ap.add_argument("--out", default=None)
ap.add_argument("--scenario", default="storm", choices=["storm", "storm_fail", "churn"])
a = ap.parse_args()
if a.scenario == "churn":      # BASELINE configs[2]: small-world graph, 5 % of the nodes fail / rejoin, 8 tracked subjects
    sc = scenarios.small_world_churn(a.nodes, a.degree, 0.1, 0.05, slots=a.slots, window=200, seed=1, fanout=a.fanout)
    extra = dict(suspicion_mult=2, suspicion_max_timeout_mult=2, probe_interval_ticks=2)
elif a.scenario == "storm_fail":   # bench.py's default workload: one subject leaves, one crashes
    sc = scenarios.dissemination_storm(a.nodes, a.degree, a.fanout, slots=max(2, a.slots), seed=1, waves=a.waves, with_fail=True)
    extra = {}
else:
    sc = scenarios.dissemination_storm(a.nodes, a.degree, a.fanout, slots=a.slots, seed=1, waves=a.waves)
    extra = {}
g = sc.build(lambda n, s, **kw: GossipSim(n, s, **kw), **extra)
for run in range(a.runs):
    g.reset(1); sc.schedule(g)
    g.set_tick_timing(run == a.runs - 1)
    ticks, ok = g.run_until_converged(sc.max_ticks)
tr, ms = g.tick_trace(), g.tick_times_ms()
st = g.stats()
p_dirty = st["changed"] / max(1, st["edge_updates"])
rows = []
for t in range(len(ms)):
    eu, ch = int(tr["edge_updates"][t]), int(tr["changed"][t])
    be = 4 + 32 / a.fanout + 32 + 32 * (ch / eu if eu else 0)
    rows.append({"tick": t, "edge_updates": eu, "changed": ch, "pending": int(tr["pending"][t]), "ms": float(ms[t]),
                 "alg_GBps": float(eu * be / (ms[t] * 1e-3) / 1e9) if ms[t] > 0 else 0.0})
    print(f"tick {t:3d}  eu {eu:10d}  changed {ch:9d}  pending {int(tr['pending'][t]):9d}  {ms[t]*1e3:9.1f} us  {rows[-1]['alg_GBps']:8.1f} GB/s(alg)")
tot = float(ms.sum())
print(f"total {tot:.3f} ms kernel time, {st['edge_updates']} edge-updates, {st['edge_updates'] / tot / 1e6:.2f} G edge-updates/s (kernel time only), p_dirty {p_dirty:.4f}")
if a.out:
    json.dump({"scenario": sc.name, "rows": rows, "kernel_ms": tot, "edge_updates": st["edge_updates"], "p_dirty": p_dirty}, open(a.out, "w"), indent=1)
