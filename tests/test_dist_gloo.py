"""CPU coverage of the N > 1 host path with a world_size-2 gloo group: the shard split, the collective hooks
the library calls between ticks (IPC-handle all-gather, barrier, u64 all-reduce with wrap-around), and the
fact that a sharded trace is the sum of per-shard traces (checked on the CPU oracle split by id range)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from serf_b200 import dist as sdist
    gather, barrier, allreduce = sdist.make_hooks(dist)
    blobs = gather(bytes([rank]) * 200)
    assert [b[0] for b in blobs] == list(range(world)) and all(len(b) == 200 for b in blobs)
    arr = np.array([rank + 1, 2**63 + 5, 2**64 - 1], dtype=np.uint64)
    allreduce(arr)
    exp = np.array([sum(range(1, world + 1)), (world * (2**63 + 5)) % 2**64, (world * (2**64 - 1)) % 2**64], dtype=np.uint64)
    assert (arr == exp).all()
    barrier()
    first, count = sdist.shard_range(1001, rank, world)
    q.put((rank, first, count))
    dist.destroy_process_group()


def test_gloo_hooks_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29000 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    got = sorted(q.get() for _ in range(2))
    assert got == [(0, 0, 501), (1, 501, 500)]


def test_shard_ranges_cover_everything():
    from serf_b200.dist import shard_range
    for n in (7, 256, 100_000, 10_000_000):
        for w in (1, 2, 4, 8):
            r = [shard_range(n, k, w) for k in range(w)]
            assert r[0][0] == 0 and sum(c for _, c in r) == n
            for (f0, c0), (f1, _) in zip(r, r[1:]):
                assert f0 + c0 == f1


# ---- the sharded tick on CPU: two gloo ranks, each owning half of the id range of the ORACLE ----------------
def _sharded_worker(rank, world, port, q):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import ctypes as C
    from oracle_lib import lib, oracle_sim
    from serf_b200 import scenarios
    from serf_b200 import dist as sdist
    L = lib()
    L.oracle_sim_set_ownership.restype, L.oracle_sim_set_ownership.argtypes = C.c_int, [C.c_void_p, C.c_uint32, C.c_uint32]
    L.oracle_sim_export.restype, L.oracle_sim_export.argtypes = C.c_uint32, [C.c_void_p, C.c_void_p, C.c_uint32]
    L.oracle_sim_import.restype, L.oracle_sim_import.argtypes = C.c_int, [C.c_void_p, C.c_void_p, C.c_uint32]
    L.oracle_msg_size.restype = C.c_uint32
    msg_dt = np.dtype([("dst", "<u4"), ("src", "<u4"), ("val", "<u4"), ("slot", "u1"), ("kind", "u1"), ("pad", "<u2")])
    assert msg_dt.itemsize == L.oracle_msg_size()
    sc = scenarios.random_graph_fail(3001, 12, 3, seed=5)
    cfg = dict(suspicion_mult=2, suspicion_max_timeout_mult=2, probe_interval_ticks=2, trace=1)
    o = oracle_sim(sc.n, sc.slots, **dict(sc.cfg, **cfg))
    first, count = sdist.shard_range(sc.n, rank, world)
    assert L.oracle_sim_set_ownership(o._h, first, count) == 0
    o.set_topology(sc.row_ptr, sc.col); o.set_subjects(sc.subjects); o.reset(sc.cfg["seed"]); sc.schedule(o)
    ticks = 160
    for _ in range(ticks):
        o.step(1)
        buf = np.zeros(200_000, dtype=msg_dt)
        n = L.oracle_sim_export(o._h, buf.ctypes.data, len(buf))
        assert n <= len(buf)
        parts = [None] * world
        dist.all_gather_object(parts, buf[:n].tobytes())
        for r, b in enumerate(parts):
            if r == rank:
                continue
            m = np.frombuffer(b, dtype=msg_dt)
            mine = np.ascontiguousarray(m[(m["dst"] >= first) & (m["dst"] < first + count)])
            if len(mine):
                assert L.oracle_sim_import(o._h, mine.ctypes.data, len(mine)) == 0
    tr = o.tick_trace(0, ticks)
    q.put((rank, first, count, tr.tobytes(), [o.records(s)[first:first + count].tobytes() for s in range(sc.slots)],
           o.lamport_time()[first:first + count].tobytes(), o.state_hash()))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_oracle_equals_single_instance():
    """Range-sharding + commutative inbox reduction: two instances that own half of the ids each and exchange only the
    cross-shard messages reproduce the single-instance records, clocks, and (summed) per-tick trace incl. the state hash."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle_lib import oracle_sim
    from serf_b200 import scenarios
    from serf_b200.sim import RECORD_DTYPE, TRACE_DTYPE
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31000 + os.getpid() % 2000
    procs = [ctx.Process(target=_sharded_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=300) for _ in range(2))
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    sc = scenarios.random_graph_fail(3001, 12, 3, seed=5)
    ref = sc.build(oracle_sim, suspicion_mult=2, suspicion_max_timeout_mult=2, probe_interval_ticks=2, trace=1)
    ref.step(160)
    tr_ref = ref.tick_trace(0, 160)
    tr_sum = np.zeros(160, dtype=TRACE_DTYPE)
    for _, _, _, trb, _, _, _ in got:
        tr = np.frombuffer(trb, dtype=TRACE_DTYPE)
        for f in TRACE_DTYPE.names:
            tr_sum[f] = tr_sum[f] + tr[f]              # u64 wrap-around == the all-reduce the library does
    for f in TRACE_DTYPE.names:
        assert (tr_sum[f] == tr_ref[f]).all(), f
    for s in range(sc.slots):
        rec = np.concatenate([np.frombuffer(g[4][s], dtype=RECORD_DTYPE) for g in got])
        assert (rec == ref.records(s)).all()
    clk = np.concatenate([np.frombuffer(g[5], dtype=np.uint64) for g in got])
    assert (clk == ref.lamport_time()).all()
    assert (sum(g[6] for g in got) % 2**64) == ref.state_hash()
