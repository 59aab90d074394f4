#!/usr/bin/env python
"""bench.py — gossip edge-updates/s of the hot path on BASELINE.json's 10 M-node workload.

One STEP = one complete dissemination study on the resident cluster: reset to the bootstrap
state, schedule the host operations, run gossip ticks until the cluster is quiescent
(serfsim_run_until_converged).  `value` = edge-updates of all ranks ÷ device time (CUDA events on
the launch stream, max over ranks), inputs resident in HBM.  `e2e` = the same study driven through
the C ABI with HOST buffers: the operation schedule goes host→device and the member-status,
status-time and Lamport-clock vectors come back device→host inside the timed region (the two Lamport vectors through the
compact u32 getters: the device keeps them in 32 bits and fails loudly rather than wrap).

    python bench.py --gpus 1 --steps 5 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
           --master-port 29500 bench.py --gpus 8 --steps 5 --warmup 3
    python bench.py --impl reference          # the CPU oracle (port of the reference path) on the host cores
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "gossip edge-updates/sec @10M nodes"
UNIT = "edge-updates/s"
HBM_FALLBACK_GBS = 6650.0          # B200_PROFILING.md fallback when MEASURED_PEAKS.json is absent


def workload(args):
    from serf_b200 import scenarios
    return scenarios.dissemination_storm(args.nodes, args.degree, args.fanout, slots=args.slots, seed=1, waves=args.waves)


def config_dict(args, sc):
    return {"workload": f"configs[3] shape: {args.nodes}-node random graph (out-degree {args.degree}), fanout={args.fanout}, "
                        f"{args.slots} tracked subject(s) leave at tick 0, run to quiescence",
            "scenario": sc.name, "nodes": args.nodes, "degree": args.degree, "fanout": args.fanout, "slots": args.slots,
            "retransmit_mult": 4, "cache": "member records (%d MB) + CSR (%d MB) exceed the 126 MB L2; no flush needed"
            % (args.nodes * 32 * args.slots // 2**20, args.nodes * args.degree * 4 // 2**20)}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons (B200_PROFILING.md recipe).  nvidia-smi needs ~1 s to start, so it is
    launched before the warm-up; samples are time-stamped and only those inside [mark_begin, mark_end] — the timed
    regions — are summarised (all samples under load if the window caught none)."""

    QUERY = "timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index
        self.t0 = self.t1 = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [x.strip() for x in line.split(",")]))

    def mark_begin(self):
        self.t0 = time.time()

    def mark_end(self):
        self.t1 = time.time()

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.1)
        self.proc.terminate()

        def summarise(rows):
            sm, mx, reasons = [], [], set()
            for _, r in rows:
                try:
                    sm.append(float(r[1])); mx.append(float(r[2]))
                except (ValueError, IndexError):
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            return sm, mx, reasons
        inside = [x for x in self.rows if self.t0 is not None and self.t0 <= x[0] <= (self.t1 or 1e30)]
        window = "timed regions"
        if not inside:
            inside, window = self.rows, "warm-up + timed regions (no sample fell inside the timed window)"
        sm, mx, reasons = summarise(inside)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "window": window, "reasons": sorted(reasons)}


def b_edge(fanout, p_dirty):
    """ALGORITHMIC bytes per edge-update: edge index + amortised source record +
    destination record read + destination write-back when the merge changed it."""
    return 4.0 + 32.0 / fanout + 32.0 + 32.0 * p_dirty


_ORACLE_CACHE = {}


def time_oracle(args, nodes, threads=None):
    """CPU baseline: the oracle (C++ port of the reference path) on a bounded sample of the workload,
    node ranges split over `threads` host threads (default: every core of the box).  The cluster (topology,
    oracle handle) is built once and reused; a call times one reset + schedule + run to quiescence."""
    import ctypes
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle_lib import lib, oracle_sim
    from serf_b200 import scenarios
    threads = threads or (os.cpu_count() or 1)
    key = (nodes, threads)
    if key not in _ORACLE_CACHE:
        sc = scenarios.dissemination_storm(nodes, args.degree, args.fanout, slots=args.slots, seed=1, waves=args.waves)
        o = oracle_sim(sc.n, sc.slots, **sc.cfg)
        L = lib()
        L.oracle_sim_set_threads.restype, L.oracle_sim_set_threads.argtypes = ctypes.c_int, [ctypes.c_void_p, ctypes.c_int]
        assert L.oracle_sim_set_threads(o._h, threads) == 0
        o.set_topology(sc.row_ptr, sc.col); o.set_subjects(sc.subjects)
        _ORACLE_CACHE[key] = (sc, o)
    sc, o = _ORACLE_CACHE[key]
    t0 = time.perf_counter()
    o.reset(sc.cfg.get("seed", 1)); sc.schedule(o)
    ticks, ok = o.run_until_converged(sc.max_ticks)
    dt = time.perf_counter() - t0
    st = o.stats()
    return {"value": st["edge_updates"] / dt, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"same scenario at {nodes} nodes ({nodes / args.nodes:.3g} of the workload), full run to quiescence "
                      f"({ticks} ticks, {st['edge_updates']} edge-updates, {dt:.1f} s)", "seconds": dt, "ticks": ticks}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    nodes = args.ref_nodes or (args.nodes if (os.cpu_count() or 1) >= 32 else 1_000_000)
    vals, last = [], None
    for i in range(args.warmup + args.steps):
        r = time_oracle(args, nodes)
        if i >= args.warmup:
            vals.append(r); last = r
    total_eu = sum(float(r["value"]) * r["seconds"] for r in vals)
    total_s = sum(r["seconds"] for r in vals)
    v = total_eu / total_s
    sc_cfg = config_dict(args, workload_stub(args))
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * total_s / max(1, len(vals)), "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "u32", "data": "synthetic", "config": sc_cfg,
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": last["cores"], "kind": "port", "sample": last["sample"]},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line))


class _Stub:
    name = "storm"


def workload_stub(args):
    s = _Stub()
    s.name = f"storm_{args.nodes}_d{args.degree}_f{args.fanout}_r{args.slots}_w{args.waves}"
    return s


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--nodes", type=int, default=10_000_000)
    ap.add_argument("--degree", type=int, default=16)
    ap.add_argument("--fanout", type=int, default=4)
    ap.add_argument("--slots", type=int, default=1)
    ap.add_argument("--waves", type=int, default=1)
    ap.add_argument("--ref-nodes", type=int, default=0, help="size of the bounded CPU sample (0: 1 M nodes inside the b200 arm; the reference arm uses the full workload on hosts with >= 32 cores)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3

    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from serf_b200 import GossipSim

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    sc = workload(args)
    g = sc.build(lambda n, s, **kw: GossipSim(n, s, **kw), device=local_rank, rank=rank, world_size=world, trace=0)
    if world > 1:
        from serf_b200 import dist as sdist
        sdist.connect(g, dist, torch.device("cuda", local_rank))

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # pinned host buffers for the results a caller reads back (the C ABI copies into caller-owned memory)
    pin_status = [torch.empty(g.count, dtype=torch.uint8).pin_memory() for _ in range(sc.slots)]
    pin_ltime = [torch.empty(g.count, dtype=torch.int32).pin_memory() for _ in range(sc.slots)]     # Lamport times cross PCIe as u32 (serfsim_*_u32)
    pin_clock = torch.empty(g.count, dtype=torch.int32).pin_memory()

    def one_step(read_back):
        g.reset(1)
        sc.schedule(g)                                 # host→device: the operation schedule
        ticks, ok = g.run_until_converged(sc.max_ticks)
        ms, launches = g.last_step_device_ms()
        out_bytes = 0
        if read_back:                                  # device→host: the step's result vectors
            for s in range(sc.slots):
                out_bytes += g.member_status(s, out=pin_status[s].numpy()).nbytes
                out_bytes += g.status_ltime_u32(s, out=pin_ltime[s].numpy().view(np.uint32)).nbytes
            out_bytes += g.lamport_time_u32(out=pin_clock.numpy().view(np.uint32)).nbytes
        return ticks, ok, ms, launches, out_bytes

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        one_step(False)
    time.sleep(1.2)                                      # let nvidia-smi come up before the timed regions (all ranks: steps are collective)
    for _ in range(2):
        one_step(False)

    # ---- device-timed region: K steps ----
    sync_all()
    sampler.mark_begin()
    dev_ms, launches, ticks_list = 0.0, 0, []
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ticks, ok, ms, nl, _ = one_step(False)
        dev_ms += ms; launches += nl; ticks_list.append(ticks)
    sync_all()
    wall_dev = time.perf_counter() - t0
    st = g.stats()                                       # global sums (all ranks) of the LAST step
    eu_per_step, changed = st["edge_updates"], st["changed"]

    # ---- end-to-end region: host buffers in, host buffers out ----
    sync_all()
    t0 = time.perf_counter()
    d2h = 0
    for _ in range(args.steps):
        _, _, _, _, ob = one_step(True)
        d2h = ob
    sync_all()
    wall_e2e = time.perf_counter() - t0
    sampler.mark_end()
    clocks = sampler.stop() if rank == 0 else None

    t = torch.tensor([dev_ms, wall_e2e, wall_dev], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, wall_e2e, wall_dev = [float(x) for x in t.cpu()]

    if rank == 0:
        total_eu = eu_per_step * args.steps
        value = total_eu / wall_dev                    # K whole steps (reset + schedule + ticks), sync to sync
        p_dirty = changed / max(1, eu_per_step)
        be = b_edge(args.fanout, p_dirty)
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        else:
            peak, peak_src = HBM_FALLBACK_GBS, "fallback (B200_PROFILING.md)"
        tick_launches = launches                       # every kernel this library launched in the timed region (tick kernels ≥ 98 % of them)
        traffic, traffic_src = None, None
        import glob
        tfiles = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_traffic.json")))     # newest round's ncu capture of this workload
        if world == 1 and tfiles:                      # DRAM bytes per tick launch from the committed ncu capture
            tj = json.load(open(tfiles[-1]))
            traffic = tj.get("dram_bytes_per_launch")
            traffic_src = f"{os.path.relpath(tfiles[-1], ROOT)} (ncu dram__bytes_read.sum + dram__bytes_write.sum, mean per tick launch; {tj.get('source', '')})"
            if os.path.basename(tfiles[-1]) == "r1_traffic.json":
                traffic_src += " — captured on the round-1 kernel BEFORE multi-tile compaction and the queue-word layout"
        # per-GPU: each GPU runs its own tick kernel over its shard; algorithmic bytes split evenly
        achieved = (total_eu / world) * be / (dev_ms * 1e-3) / 1e9
        h2d = len(sc.ops) * 12
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": 1e3 * wall_dev / args.steps, "kernel_ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "u32", "data": "synthetic", "config": config_dict(args, sc),
                "ticks_to_convergence": ticks_list[-1], "edge_updates_per_step": eu_per_step, "p_dirty": p_dirty,
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                             "traffic": traffic, "traffic_source": traffic_src,
                             "algorithmic_bytes_per_launch": total_eu * be / world / max(1, tick_launches), "peak_source": peak_src, "kernel": "tick_kernel", "bytes_per_edge_update": be,
                             "launches": tick_launches, "avg_launch_us": 1e3 * dev_ms / max(1, tick_launches)},
                "e2e": {"value": total_eu / wall_e2e, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "ms_per_step": 1e3 * wall_e2e / args.steps},
                "gpu_launches": launches, "clocks": clocks}
        if not args.no_cpu_baseline:
            line["cpu_baseline"] = {k: v for k, v in time_oracle(args, args.ref_nodes or 1_000_000).items() if k not in ("seconds", "ticks")}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
