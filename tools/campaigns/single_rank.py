"""Randomized campaign of fuzz / fuzz_features scenarios beyond the committed seeds: host-compiled kernels (tests/emu) vs oracle.
Run from the repo root; prints the failing seeds (none expected).  Takes 10-20 minutes."""
import sys, time
import os; ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
from emu_lib import emu_sim
from oracle_lib import oracle_sim
from serf_b200 import scenarios
import test_emu_parity as P
import numpy as np
bad=[]
t0=time.time()
def one(sc, tag):
    sc.max_ticks=min(sc.max_ticks,1500)
    try:
        o = sc.build(oracle_sim, trace=1); to=o.run_until_converged(sc.max_ticks)
        for trace in (0,1):
            g = sc.build(emu_sim, trace=trace)
            assert g.run_until_converged(sc.max_ticks)==to
            P.assert_same(g,o,sc.slots,with_hash=bool(trace))
            P._feature_checks(g,o,sc)
    except Exception as e:
        bad.append((tag, repr(e)[:200])); print('FAIL', tag, repr(e)[:200], flush=True)
for s in range(40, 400):
    one(scenarios.fuzz(s), f'fuzz{s}')
    if time.time()-t0 > 900: break
print('fuzz done up to', s, round(time.time()-t0), flush=True)
t1=time.time()
for s in range(30, 330):
    one(scenarios.fuzz_features(s), f'feat{s}')
    if time.time()-t1 > 900: break
print('features done up to', s, round(time.time()-t1), 'bad:', bad, flush=True)
