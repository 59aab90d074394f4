#!/usr/bin/env python
"""Generates tests/golden/traces.json from the CPU oracle (which is pinned by the reference's KATs).

The reference is Rust and cannot run here, so these are not reference outputs: they freeze the tick model
(convergence tick, totals, final state hash, a digest of every per-tick trace row) so that an accidental
simultaneous drift of oracle and kernel is caught.  Regenerate only when the tick semantics change on purpose:
    python tools/make_golden.py
"""
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle_lib import oracle_sim  # noqa: E402
from serf_b200 import scenarios  # noqa: E402

CASES = [("full_mesh_leave", dict(n=256, fanout=3, seed=s)) for s in (1, 2, 3)] + \
        [("random_graph_leave", dict(n=100_000, degree=16, fanout=3, seed=1)),
         ("random_graph_leave", dict(n=20_000, degree=16, fanout=4, seed=5, slots=4)),
         ("random_graph_fail", dict(n=5_000, degree=24, fanout=3, seed=4))] + \
        [("fuzz", dict(seed=s)) for s in range(10)]


# user events / byzantine injectors: kept in their own fixture (tests/golden/features.json) and their own test files, so that
# the device tests of the parts that have not run on hardware yet come after the ones that have
FEATURE_CASES = [("user_event_storm", dict(n=20_000, degree=16, fanout=3, seed=1, n_events=4, spacing=3)),
                 ("user_event_storm", dict(n=8_000, degree=12, fanout=4, seed=5, n_events=3, spacing=2, alias=True, churn=40, with_leave=True)),
                 ("byzantine_injectors", dict(n=20_000, degree=16, fanout=4, frac=0.01, seed=1)),
                 ("byzantine_injectors", dict(n=6_000, degree=16, fanout=4, frac=0.2, seed=3))]


def digest(sim, n):
    tr = sim.tick_trace(0, n)
    return hashlib.sha256(tr.tobytes()).hexdigest()


def run_case(name, kwargs, factory):
    sc = getattr(scenarios, name)(**kwargs)
    sc.max_ticks = min(sc.max_ticks, 1500)
    sim = sc.build(factory, trace=1)
    ticks, ok = sim.run_until_converged(sc.max_ticks)
    st = sim.stats()
    out = {"ticks": int(ticks), "converged": bool(ok), "n_ticks": st["tick"], "edge_updates": st["edge_updates"], "messages": st["messages"],
           "changed": st["changed"], "state_hash": "%016x" % sim.state_hash(), "trace_sha256": digest(sim, st["tick"])}
    if sc.user_events is not None:
        out["user_events"] = sim.user_event_stats()
        out["user_event_records_sha256"] = hashlib.sha256(sim.user_event_records().tobytes()).hexdigest()
    if sc.byzantine is not None:
        out["byzantine"] = sim.byzantine_stats()
        out["anomaly_sha256"] = hashlib.sha256(sim.anomaly_flags().tobytes()).hexdigest()
    return out


if __name__ == "__main__":
    out = []
    for name, kwargs in CASES:
        r = run_case(name, kwargs, oracle_sim)
        out.append({"scenario": name, "args": kwargs, **r})
        print(name, kwargs, r["ticks"], r["state_hash"])
    path = os.path.join(ROOT, "tests", "golden", "traces.json")
    json.dump(out, open(path, "w"), indent=1)
    print("wrote", path)
    out = []
    for name, kwargs in FEATURE_CASES:
        r = run_case(name, kwargs, oracle_sim)
        out.append({"scenario": name, "args": kwargs, **r})
        print(name, kwargs, r["ticks"], r["state_hash"])
    path = os.path.join(ROOT, "tests", "golden", "features.json")
    json.dump(out, open(path, "w"), indent=1)
    print("wrote", path)
