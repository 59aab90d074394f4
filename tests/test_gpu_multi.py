"""Multi-GPU parity (needs ≥ 2 GPUs; run with `gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu`).
The id range is sharded over 2 (or 4) ranks, cross-shard gossip goes through the NVLink windows, and the
concatenated records / summed trace must equal the CPU oracle's — i.e. the result is independent of the sharding."""
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, scen_args, outdir):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from serf_b200 import GossipSim, scenarios
    from serf_b200 import dist as sdist
    name, kwargs, cfg = scen_args
    sc = getattr(scenarios, name)(**kwargs)
    g = sc.build(lambda n, s, **kw: GossipSim(n, s, **kw), device=rank, rank=rank, world_size=world, trace=1, **cfg)
    sdist.connect(g, dist, torch.device("cuda", rank))
    ticks, ok = g.run_until_converged(sc.max_ticks)
    tr = g.tick_trace()
    extra = {}
    if sc.user_events is not None:                 # collective getters: every rank makes the same calls in the same order
        st = g.user_event_stats()
        extra.update(ue=g.user_event_records(), ue_stats=np.array([st[k] for k in sorted(st)], dtype=np.uint64),
                     ue_ltime=np.array([g.user_event_ltime(e) for e in range(len(sc.user_events))], dtype=np.uint64))
    if sc.byzantine is not None:
        bz = g.byzantine_stats()
        extra.update(flags=g.anomaly_flags(), byz_stats=np.array([bz[k] for k in sorted(bz)], dtype=np.uint64))
    np.savez(os.path.join(outdir, f"r{rank}.npz"), ticks=ticks, ok=ok, trace=tr, first=g.first, count=g.count, hash=np.uint64(g.state_hash()),
             clock=g.lamport_time(), **{f"rec{s}": g.records(s) for s in range(sc.slots)}, **extra)
    dist.barrier()
    dist.destroy_process_group()


def _run(world, scen_args, tmp_path):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    port = 29500 + os.getpid() % 1000
    procs = [ctx.Process(target=_worker, args=(r, world, port, scen_args, str(tmp_path))) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(600)
        assert p.exitcode == 0
    return [np.load(os.path.join(str(tmp_path), f"r{r}.npz")) for r in range(world)]


WORLDS = [int(x) for x in os.environ.get("SERFSIM_TEST_WORLDS", "2,4").split(",")]


@pytest.mark.parametrize("world", WORLDS)
@pytest.mark.parametrize("scen", [
    ("random_graph_leave", dict(n=50_000, degree=16, fanout=3, seed=2, slots=1), {}),
    ("random_graph_leave", dict(n=30_001, degree=12, fanout=4, seed=3, slots=3), {}),
    ("random_graph_fail", dict(n=20_000, degree=16, fanout=3, seed=2), dict(suspicion_mult=2, suspicion_max_timeout_mult=2, probe_interval_ticks=2)),
    ("fuzz", dict(seed=7, n=3000, slots=4), dict(push_pull_interval_ticks=0)),
    # cross-shard traffic added after the last multi-GPU run (checked with the host-compiled kernels, tests/test_emu_multi.py)
    ("user_event_storm", dict(n=40_000, degree=16, fanout=3, seed=3, n_events=5, spacing=2, churn=100, with_leave=True), {}),
    ("byzantine_injectors", dict(n=40_000, degree=16, fanout=4, frac=0.02, seed=1), {}),
    ("fuzz", dict(seed=11, n=3000, slots=3), {}),                                   # push-pull rounds across shards (fuzz 11 has them on)
    ("fuzz_features", dict(seed=6, n=3000, slots=3), {}),
    ("user_event_storm", dict(n=20_000, degree=8, fanout=2, seed=6, n_events=5, spacing=2, churn=100, with_leave=True), dict(push_pull_interval_ticks=5, retransmit_mult=1)),
    ("byzantine_injectors", dict(n=20_000, degree=12, fanout=3, frac=0.05, seed=5), dict(push_pull_interval_ticks=6)),
])
def test_sharded_equals_oracle(world, scen, tmp_path):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle_lib import oracle_sim
    from serf_b200 import scenarios
    name, kwargs, cfg = scen
    sc = getattr(scenarios, name)(**kwargs)
    o = sc.build(oracle_sim, trace=1, **cfg)
    to, oko = o.run_until_converged(sc.max_ticks)
    res = _run(world, scen, tmp_path)
    n = o.stats()["tick"]
    tro = o.tick_trace(0, n)
    for r in res:
        assert (int(r["ticks"]), bool(r["ok"])) == (to, oko)
        for f in tro.dtype.names:                       # every rank holds the all-reduced (global) trace
            assert (r["trace"][f] == tro[f]).all(), f
        assert int(r["hash"]) == o.state_hash()
    assert (np.concatenate([r["clock"] for r in res]) == o.lamport_time()).all()
    for s in range(sc.slots):
        assert (np.concatenate([r[f"rec{s}"] for r in res]) == o.records(s)).all()
    if sc.user_events is not None:
        assert (np.concatenate([r["ue"] for r in res]) == o.user_event_records()).all()
        so = o.user_event_stats()
        keys = sorted(so)
        for r in res:
            got = dict(zip(keys, (int(x) for x in r["ue_stats"])))
            assert {k: v for k, v in got.items() if k != "event_time"} == {k: v for k, v in so.items() if k != "event_time"}
            assert [int(x) for x in r["ue_ltime"]] == [o.user_event_ltime(e) for e in range(len(sc.user_events))]
        assert max(int(dict(zip(keys, r["ue_stats"]))["event_time"]) for r in res) == so["event_time"]
    if sc.byzantine is not None:
        assert (np.concatenate([r["flags"] for r in res]) == o.anomaly_flags()).all()
        bo = o.byzantine_stats()
        for r in res:
            assert dict(zip(sorted(bo), (int(x) for x in r["byz_stats"]))) == bo
