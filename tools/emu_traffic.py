#!/usr/bin/env python
"""Bytes ISSUED by the tick kernel's memory accessors over one run of the bench workload, by class, counted in the
host build (tests/emu).  An accounting aid for layout experiments (e.g. the queue-word branch): it counts what the code
asks for, not what DRAM serves (no cache model).

  python tools/emu_traffic.py [--nodes 10000000]
"""
import argparse
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from emu_lib import emu_sim, lib  # noqa: E402
from serf_b200 import scenarios  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--nodes", type=int, default=10_000_000)
ap.add_argument("--fanout", type=int, default=4)
a = ap.parse_args()
sc = scenarios.dissemination_storm(a.nodes, 16, a.fanout, slots=1, seed=1)
g = sc.build(emu_sim, trace=0)
L = lib()
L.emu_probe.restype = C.c_ulong
L.emu_probe_reset()
ticks, ok = g.run_until_converged(sc.max_ticks)
names = {6: "queue-word loads (queue-word layout only)", 7: "queue-word stores (queue-word layout only)", 8: "record loads", 9: "record stores", 10: "node-word loads", 11: "inbox / stream u32 loads", 12: "inbox clears (u32 stores)",
         13: "node-word stores", 14: "RED.MAX (4 B each)", 15: "read-only path: row offsets + neighbour gathers"}
st = g.stats()
out = {"nodes": a.nodes, "ticks": ticks, "edge_updates": st["edge_updates"], "bytes": {names[k]: int(L.emu_probe(k)) for k in names}}
out["total_issued_bytes"] = sum(out["bytes"].values())
out["issued_bytes_per_edge_update"] = out["total_issued_bytes"] / max(1, st["edge_updates"])
print(json.dumps(out, indent=1))
