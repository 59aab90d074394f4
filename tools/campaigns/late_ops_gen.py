"""Scenario generator of the late-operations campaign (tools/campaigns/late_ops.py, tests/test_emu_host.py)."""
import numpy as np
from serf_b200 import scenarios
from serf_b200.sim import Op


def late(seed):
    sc = scenarios.fuzz(seed)
    rng = np.random.Generator(np.random.Philox(seed + 777))
    used = {(t, node) for (t, _, node, _) in sc.ops}
    t0 = max([t for (t, *_) in sc.ops] + [0])
    for _ in range(int(rng.integers(1, 7))):
        t = t0 + int(rng.integers(1, 300))
        kind = rng.choice([Op.JOIN, Op.FORCE_LEAVE, Op.FORCE_LEAVE, Op.REJOIN, Op.FAIL, Op.LEAVE])
        s = int(rng.integers(0, sc.slots))
        node = int(rng.integers(0, sc.n)) if kind == Op.FORCE_LEAVE else int(sc.subjects[s])
        if (t, node) not in used:
            used.add((t, node)); sc.ops.append((t, int(kind), node, s))
    sc.name = f"late_{seed}"; sc.max_ticks = 3000
    return sc
