"""Production mode (trace = 0: tile skipping, lazy loads, compaction) across converge → inject → continue sequences on
the device — the mirror of tests/test_emu_host.py::test_multi_phase_production_mode_rewind (the convergence
loop's rewind used to drop the watcher tiles' flags after an odd number of speculative ticks)."""
import pytest

from oracle_lib import oracle_sim
from serf_b200 import GossipSim, scenarios
from serf_b200.sim import Op

pytestmark = pytest.mark.gpu


def same(g, o, slots):
    assert g.stats() == o.stats()
    n = o.stats()["tick"]
    tg, to = g.tick_trace(0, n), o.tick_trace(0, n)
    for f in tg.dtype.names:
        if f != "hash":
            assert (tg[f] == to[f]).all(), f
    assert g.state_hash() == o.state_hash()
    for s in range(slots):
        assert (g.records(s) == o.records(s)).all()
    assert (g.lamport_time() == o.lamport_time()).all()


@pytest.mark.parametrize("seed", range(1, 9))
@pytest.mark.parametrize("chunk", ["4", "3"])
def test_multi_phase_production_mode(monkeypatch, seed, chunk):
    monkeypatch.setenv("SERFSIM_CHUNK", chunk)
    sc = scenarios.random_graph_leave(30_000, 12, 3, seed=seed, slots=2, graph_seed=seed + 20)
    cfg = dict(suspicion_mult=2, suspicion_max_timeout_mult=2, probe_interval_ticks=2)
    sc.ops = [(0, Op.JOIN, int(sc.subjects[0]), 0)]
    g, o = sc.build(lambda n, s, **kw: GossipSim(n, s, **kw), trace=0, **cfg), sc.build(oracle_sim, trace=1, **cfg)
    assert g.run_until_converged(sc.max_ticks) == o.run_until_converged(sc.max_ticks)
    for sim in (g, o):
        sim.inject(sim.stats()["tick"], Op.FAIL, int(sc.subjects[1]), 0)
    assert g.run_until_converged(5000) == o.run_until_converged(5000)
    same(g, o, sc.slots)
    for sim in (g, o):
        t = sim.stats()["tick"]
        sim.inject(t, Op.REJOIN, int(sc.subjects[1]), 0)
        sim.inject(t + 3, Op.FORCE_LEAVE, 7, 0)
    assert g.run_until_converged(5000) == o.run_until_converged(5000)
    same(g, o, sc.slots)
