#!/usr/bin/env python
"""Per-tick device time of the sharded bench workload (launch under torchrun, one rank per GPU)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from serf_b200 import GossipSim, scenarios  # noqa: E402
from serf_b200 import dist as sdist  # noqa: E402

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
nodes = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
sc = scenarios.dissemination_storm(nodes, 16, 4, slots=1, seed=1)
g = sc.build(lambda n, s, **kw: GossipSim(n, s, **kw), device=lr, rank=rank, world_size=world)
sdist.connect(g, dist, torch.device("cuda", lr))
for run in range(2):
    g.reset(1); sc.schedule(g); g.set_tick_timing(run == 1)
    ticks, ok = g.run_until_converged(sc.max_ticks)
tr, ms = g.tick_trace(), g.tick_times_ms()
if rank == 0:
    for t in range(len(ms)):
        print(f"tick {t:3d} eu {int(tr['edge_updates'][t]):10d} {ms[t]*1e3:9.1f} us")
    print(f"world {world}: total {ms.sum():.3f} ms (rank 0 kernel time), {int(tr['edge_updates'].sum())} edge-updates")
g.close()
dist.barrier()
dist.destroy_process_group()
