"""Sharded path logic on the CPU: W ranks as threads of this process, each driving its own handle of the
host-compiled kernels (tests/emu); the "NVLink windows" are plain shared memory with real acquire/release on the
flags, the host collectives (barrier, u64 all-reduce) are thread primitives.  Concatenated records and the
all-reduced trace must equal the single oracle's — the comparisons of tests/test_gpu_multi.py at small sizes."""
import threading

import numpy as np
import pytest

from emu_lib import emu_sim
from oracle_lib import oracle_sim
from serf_b200 import scenarios


class ThreadComm:
    def __init__(self, world):
        self.world = world
        self.bar = threading.Barrier(world)
        self.blobs = [None] * world
        self.acc = None
        self.lock = threading.Lock()

    def hooks(self, rank):
        def all_gather_bytes(b):
            self.blobs[rank] = b
            self.bar.wait()
            out = list(self.blobs)
            self.bar.wait()
            return out

        def barrier():
            self.bar.wait()

        def allreduce_u64(arr):
            with self.lock:
                if self.acc is None:
                    self.acc = arr.copy()
                else:
                    self.acc = self.acc + arr                       # u64 wrap-around sum
            self.bar.wait()
            arr[:] = self.acc
            self.bar.wait()
            if rank == 0:
                self.acc = None
            self.bar.wait()
        return all_gather_bytes, barrier, allreduce_u64


def run_sharded(sc, world, trace=1, **cfg):
    comm = ThreadComm(world)
    res, errs = [None] * world, []

    def worker(rank):
        try:
            g = sc.build(emu_sim, rank=rank, world_size=world, trace=trace, **cfg)
            g.connect(*comm.hooks(rank))
            ticks, ok = g.run_until_converged(sc.max_ticks)
            res[rank] = dict(ticks=ticks, ok=ok, trace=g.tick_trace(), hash=g.state_hash(), clock=g.lamport_time(),
                             rec=[g.records(s) for s in range(sc.slots)], first=g.first, count=g.count)
            comm.bar.wait()
        except BaseException as e:                                  # noqa: BLE001 — surface it in the main thread
            errs.append(e)
            comm.bar.abort()
    th = [threading.Thread(target=worker, args=(r,)) for r in range(world)]
    for t in th:
        t.start()
    for t in th:
        t.join(600)
    if errs:
        raise errs[0]
    return res


def check(sc, world, **cfg):
    o = sc.build(oracle_sim, trace=1, **cfg)
    to, oko = o.run_until_converged(sc.max_ticks)
    n = o.stats()["tick"]
    tro = o.tick_trace(0, n)
    for trace in (1, 0):
        res = run_sharded(sc, world, trace=trace, **cfg)
        for r in res:
            assert (r["ticks"], r["ok"]) == (to, oko)
            for f in tro.dtype.names:                               # every rank holds the all-reduced (global) trace
                if f == "hash" and not trace:
                    continue
                bad = np.nonzero(r["trace"][f] != tro[f])[0]
                assert bad.size == 0, f"world {world} trace={trace}: field {f} first differs at tick {bad[0]}"
            assert r["hash"] == o.state_hash()
        assert (np.concatenate([r["clock"] for r in res]) == o.lamport_time()).all()
        for s in range(sc.slots):
            got = np.concatenate([r["rec"][s] for r in res])
            bad = np.nonzero(got != o.records(s))[0]
            assert bad.size == 0, f"world {world} trace={trace} slot {s}: record of node {bad[0]} differs"


@pytest.mark.parametrize("world", [2, 3, 4])
def test_sharded_random_graph(world):
    check(scenarios.random_graph_leave(3001, 12, 3, seed=2, slots=1), world)


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_multi_slot_fanout4(world):
    check(scenarios.random_graph_leave(2500, 12, 4, seed=3, slots=3), world)


def test_sharded_failure_detection():
    check(scenarios.random_graph_fail(2000, 16, 3, seed=2), 2, suspicion_mult=2, suspicion_max_timeout_mult=2, probe_interval_ticks=2)


@pytest.mark.parametrize("seed", [7, 9, 12])
def test_sharded_fuzz(seed):
    check(scenarios.fuzz(seed, n=600, slots=4), 2, push_pull_interval_ticks=0)
