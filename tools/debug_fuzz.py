#!/usr/bin/env python
"""Print the first difference between the CUDA path and the oracle on a fuzz scenario (debug aid)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle_lib import oracle_sim  # noqa: E402
from serf_b200 import GossipSim, scenarios  # noqa: E402

seed = int(sys.argv[1])
sc = scenarios.fuzz(seed)
print(sc.n, sc.slots, sc.cfg, "subjects", sc.subjects.tolist())
print(sorted(sc.ops))
g = sc.build(lambda n, s, **kw: GossipSim(n, s, **kw), trace=1)
o = sc.build(oracle_sim, trace=1)
for t in range(sc.max_ticks):
    g.step(1); o.step(1)
    rg, ro = g.tick_trace(t, 1)[0], o.tick_trace(t, 1)[0]
    if rg != ro:
        print("tick", t, "rows differ\n gpu   ", rg, "\n oracle", ro)
        for s in range(sc.slots):
            a, b = g.records(s), o.records(s)
            bad = np.nonzero(a != b)[0]
            for v in bad[:4]:
                print(" slot", s, "node", v, "subject", int(sc.subjects[s]), "\n   gpu   ", a[v], "\n   oracle", b[v])
        ca, cb = g.lamport_time(), o.lamport_time()
        bad = np.nonzero(ca != cb)[0]
        print(" clocks differ at", bad[:8], ca[bad[:8]], cb[bad[:8]])
        break
else:
    print("no difference in", sc.max_ticks, "ticks")
