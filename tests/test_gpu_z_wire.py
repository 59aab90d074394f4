"""Wire codec on the device: SerfDelegate::local_state of every node of a shard, encoded by the batch kernels
of serf_b200/csrc/wire_codec.cu through the C ABI, against the oracle's per-node encoding (oracle/wire_oracle.cpp) byte for byte,
and back through the decode kernel.  The single-message host entry points are compared with the oracle in tests/test_wire.py."""
import numpy as np
import pytest

import wire_lib as W
from oracle_lib import oracle_sim
from serf_b200 import GossipSim, scenarios
from serf_b200.sim import load_library

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("scen", ["leave_fail", "prune", "churn"])
def test_device_local_state_batch_equals_oracle(scen):
    if scen == "leave_fail":
        sc, cfg = scenarios.dissemination_storm(100_000, 16, 4, slots=2, seed=3, with_fail=True), {}
    elif scen == "prune":
        sc, cfg = scenarios.fuzz_prune(5, n=5000, slots=6), {}
    else:
        sc, cfg = scenarios.small_world_churn(60_000, 16, 0.1, 0.05, slots=8, window=40, seed=3), dict(suspicion_mult=2, suspicion_max_timeout_mult=2, probe_interval_ticks=2)
    P = W.bind_product(load_library())
    o = sc.build(oracle_sim, trace=0, **cfg)
    g = sc.build(lambda n, s, **kw: GossipSim(n, s, **kw), trace=0, **cfg)
    rng = np.random.default_rng(1)
    for ticks in (0, 12, 70):
        o.step(ticks); g.step(ticks)
        buf, off = W.local_state_batch(P, g)
        view = dict(status=[o.member_status(s) for s in range(sc.slots)], ltime=[o.status_ltime(s) for s in range(sc.slots)], clock=o.lamport_time())
        for v in [int(x) for x in rng.integers(0, sc.n, 400)] + [int(x) for x in sc.subjects] + [0, sc.n - 1]:
            ltime, status, left = W.expected_local_state(view, v, sc.subjects)
            assert bytes(buf[int(off[v]):int(off[v + 1])]) == W.o_encode_push_pull(ltime, status, left, 1, 1), (scen, ticks, v)
        assert int(off[-1]) == buf.size and (np.diff(off.astype(np.int64)) > 0).all()
        lt, ids, sts, ns = W.decode_batch(P, g, buf, off, sc.slots)
        assert (lt == o.lamport_time()).all()
        known = np.stack([o.member_status(s) != 0 for s in range(sc.slots)], axis=1)
        assert (ns == known.sum(axis=1)).all()
        for v in [int(x) for x in rng.integers(0, sc.n, 300)]:
            _, status, _ = W.expected_local_state(view, v, sc.subjects)
            assert [(int(ids[v, i]), int(sts[v, i])) for i in range(ns[v])] == status


def test_device_batch_one_million_nodes_round_trip():
    """1 M nodes × 8 subjects: the batch is a few tens of MB; every message decodes to the state the getters report."""
    sc = scenarios.small_world_churn(1_000_000, 16, 0.1, 0.05, slots=8, window=200, seed=1, fanout=3)
    P = W.bind_product(load_library())
    g = sc.build(lambda n, s, **kw: GossipSim(n, s, **kw), trace=0)
    g.step(150)
    buf, off = W.local_state_batch(P, g)
    lt, ids, sts, ns = W.decode_batch(P, g, buf, off, sc.slots)
    assert (lt == g.lamport_time()).all()
    status = np.stack([g.member_status(s) for s in range(sc.slots)], axis=1)
    ltime = np.stack([g.status_ltime(s) for s in range(sc.slots)], axis=1)
    assert (ns == (status != 0).sum(axis=1)).all()
    full = ns == sc.slots                                          # nodes that know every subject: entries are in slot order
    assert full.sum() > 900_000
    assert (ids[full] == sc.subjects.astype(np.uint64)).all() and (sts[full] == ltime[full]).all()
