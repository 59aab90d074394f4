"""torch.distributed plumbing for multi-GPU runs (one process per GPU, ids sharded by contiguous range).

The data path never goes through these helpers: cross-shard gossip entries are written by the tick kernel
straight into the peer GPU's window over NVLink.  torch.distributed is used for (1) exchanging the CUDA-IPC
handles once, (2) summing the per-tick trace rows / state hash when the host looks at them, (3) barriers.
"""
import numpy as np


def shard_range(n_nodes, rank, world):
    """Same split as serfsim_create: contiguous ranges of ceil(n / world) ids."""
    size = (n_nodes + world - 1) // world
    first = min(size * rank, n_nodes)
    return first, min(size, n_nodes - first)


def make_hooks(dist, device=None):
    """(all_gather_bytes, barrier, allreduce_u64) over the default process group (nccl or gloo)."""
    import torch
    world = dist.get_world_size()

    def all_gather_bytes(b):
        out = [None] * world
        dist.all_gather_object(out, b)
        return out

    def barrier():
        dist.barrier()

    def allreduce_u64(arr):
        # u64 sums modulo 2^64 == int64 sums with wrap-around
        t = torch.from_numpy(arr.view(np.int64).copy())
        if device is not None:
            t = t.to(device)
        dist.all_reduce(t)
        arr.view(np.int64)[:] = t.cpu().numpy()

    return all_gather_bytes, barrier, allreduce_u64


def connect(sim, dist, device=None):
    sim.connect(*make_hooks(dist, device))
