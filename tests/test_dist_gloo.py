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
