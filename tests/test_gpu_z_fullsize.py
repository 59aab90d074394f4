"""Full-size parity against the (threaded) CPU oracle — BASELINE.json configs[2] and configs[4] at their stated sizes.

* configs[2]: 1 M-node Watts–Strogatz small world (k = 16, β = 0.1), 50 000 nodes fail / rejoin at random
  ticks in [0, 200), R = 8 tracked subjects sampled from the churn set, memberlist LAN timers, one GPU.
* configs[4]: 10 M-node random graph, 100 000 stale-record injectors (Δ = 2), anomaly-flag output; sharded
  over every GPU of the box (8 on the scaling box; the same test runs unsharded on a one-GPU box).
Production mode (trace = 0: tile skipping, compaction, no per-tick hash).  Compared: convergence step count, every non-hash
column of every trace row, the run totals, the final state hash (any differing byte of any record or clock changes it), the
SURVEY outputs (member status, status time, Lamport clock of every node) and the anomaly flags of every node."""
import os
import sys

import numpy as np
import pytest
import torch

from oracle_lib import oracle_sim_threaded
from serf_b200 import GossipSim, scenarios

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def compare_traces(trg, tro):
    for name in trg.dtype.names:
        if name != "hash":
            bad = np.nonzero(trg[name] != tro[name])[0]
            assert bad.size == 0, f"trace field {name} first differs at tick {bad[0]}: gpu {trg[name][bad[0]]} oracle {tro[name][bad[0]]}"


def test_config2_small_world_1m_r8():
    sc = scenarios.small_world_churn(1_000_000, 16, 0.1, 0.05, slots=8, window=200, seed=1, fanout=3)
    assert sc.n == 1_000_000 and sc.slots == 8 and sum(1 for op in sc.ops if op[1] == 4) == 50_000
    o = sc.build(oracle_sim_threaded, trace=0)
    to = o.run_until_converged(sc.max_ticks)
    g = sc.build(lambda n, s, **kw: GossipSim(n, s, **kw), trace=0)
    tg = g.run_until_converged(sc.max_ticks)
    assert tg == to and to[1], (tg, to)
    n = o.stats()["tick"]
    compare_traces(g.tick_trace(0, n), o.tick_trace(0, n))
    assert g.stats() == o.stats()
    assert g.state_hash() == o.state_hash()
    assert (g.lamport_time() == o.lamport_time()).all()
    for s in range(sc.slots):
        assert (g.member_status(s) == o.member_status(s)).all()
        assert (g.status_ltime(s) == o.status_ltime(s)).all()
        assert (g.records(s) == o.records(s)).all()


CONFIG4 = dict(n=10_000_000, degree=16, fanout=4, frac=0.01, delta=2, seed=1)


def _worker(rank, world, port, outdir):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from serf_b200 import GossipSim, scenarios
    from serf_b200 import dist as sdist
    sc = scenarios.byzantine_injectors(**CONFIG4)
    g = sc.build(lambda n, s, **kw: GossipSim(n, s, **kw), device=rank, rank=rank, world_size=world, trace=0)
    sdist.connect(g, dist, torch.device("cuda", rank))
    ticks, ok = g.run_until_converged(sc.max_ticks)
    st, bz = g.stats(), g.byzantine_stats()
    np.savez(os.path.join(outdir, f"r{rank}.npz"), ticks=ticks, ok=ok, trace=g.tick_trace(), first=g.first, count=g.count, hash=np.uint64(g.state_hash()),
             clock=g.lamport_time_u32(), flags=g.anomaly_flags(), stats=np.array([st[k] for k in sorted(st)], dtype=np.uint64),
             byz_stats=np.array([bz[k] for k in sorted(bz)], dtype=np.uint64),
             **{f"status{s}": g.member_status(s) for s in range(sc.slots)}, **{f"ltime{s}": g.status_ltime_u32(s) for s in range(sc.slots)})
    dist.barrier()
    dist.destroy_process_group()


def test_config4_byzantine_10m_sharded(tmp_path):
    ngpu = torch.cuda.device_count()
    world = int(os.environ.get("SERFSIM_CONFIG4_WORLD", "0")) or (8 if ngpu >= 8 else 4 if ngpu >= 4 else 2 if ngpu >= 2 else 1)
    sc = scenarios.byzantine_injectors(**CONFIG4)
    assert sc.n == 10_000_000 and 99_000 <= len(sc.byzantine) <= 100_000
    o = sc.build(oracle_sim_threaded, trace=0)
    to = o.run_until_converged(sc.max_ticks)
    assert to[1]
    n = o.stats()["tick"]
    tro, fo, so, bo = o.tick_trace(0, n), o.anomaly_flags(), o.stats(), o.byzantine_stats()
    assert fo[sc.byzantine].mean() > 0.95 and fo.sum() == fo[sc.byzantine].sum()        # injectors are flagged, nobody else is
    if world == 1:
        g = sc.build(lambda n_, s, **kw: GossipSim(n_, s, **kw), trace=0)
        assert g.run_until_converged(sc.max_ticks) == to
        compare_traces(g.tick_trace(0, n), tro)
        assert g.stats() == so and g.byzantine_stats() == bo and g.state_hash() == o.state_hash()
        assert (g.anomaly_flags() == fo).all() and (g.lamport_time_u32() == o.lamport_time()).all()
        for s in range(sc.slots):
            assert (g.member_status(s) == o.member_status(s)).all() and (g.status_ltime_u32(s) == o.status_ltime(s)).all()
        return
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    port = 29500 + os.getpid() % 1000
    procs = [ctx.Process(target=_worker, args=(r, world, port, str(tmp_path))) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(900)
        assert p.exitcode == 0
    res = [np.load(os.path.join(str(tmp_path), f"r{r}.npz")) for r in range(world)]
    for r in res:
        assert (int(r["ticks"]), bool(r["ok"])) == to
        compare_traces(r["trace"][:n], tro)
        assert int(r["hash"]) == o.state_hash()
        got = dict(zip(sorted(so), (int(x) for x in r["stats"])))
        glob = ("tick", "packets", "edge_updates", "messages", "changed", "events", "pending", "last_active_tick", "members")   # run totals; the agreement summary is per shard
        assert {k: got[k] for k in glob} == {k: so[k] for k in glob}
        assert dict(zip(sorted(bo), (int(x) for x in r["byz_stats"]))) == bo
    assert (np.concatenate([r["flags"] for r in res]) == fo).all()
    assert (np.concatenate([r["clock"] for r in res]) == o.lamport_time()).all()
    for s in range(sc.slots):
        assert (np.concatenate([r[f"status{s}"] for r in res]) == o.member_status(s)).all()
        assert (np.concatenate([r[f"ltime{s}"] for r in res]) == o.status_ltime(s)).all()
