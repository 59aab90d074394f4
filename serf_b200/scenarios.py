"""Synthetic workloads: the BASELINE.json configs (and a seeded fuzzer) as concrete inputs.

A scenario is plain data — topology (CSR), tracked subjects, config overrides and a schedule
of host operations — that can be fed to any GossipSim-shaped driver.  Nothing here computes
simulation results.
"""
import numpy as np

from .sim import Op, full_mesh_graph, random_regular_graph, small_world_graph


class Scenario:
    def __init__(self, name, n, slots, topology, subjects, ops, cfg=None, max_ticks=2000, user_events=None, byzantine=None, delta=2):
        self.name, self.n, self.slots = name, n, slots
        self.byzantine = None if byzantine is None else np.asarray(byzantine, dtype=np.uint32)   # stale-record injector ids
        self.delta = delta
        self.user_events = None if user_events is None else np.asarray(user_events, dtype=np.uint32)   # content id per tracked user event
        self.row_ptr, self.col = topology
        self.subjects = np.asarray(subjects, dtype=np.uint32)
        self.ops = list(ops)              # (tick, op, node, slot)
        self.cfg = dict(cfg or {})
        self.max_ticks = max_ticks

    def build(self, factory, **extra_cfg):
        """factory(n, slots, **cfg) → GossipSim-shaped driver, configured and scheduled."""
        cfg = dict(self.cfg)
        cfg.update(extra_cfg)
        sim = factory(self.n, self.slots, **cfg)
        sim.set_topology(self.row_ptr, self.col)
        sim.set_subjects(self.subjects)
        if self.user_events is not None:
            sim.set_user_events(self.user_events)
        if self.byzantine is not None:
            sim.set_byzantine(self.byzantine, self.delta)
        self.schedule(sim)
        return sim

    def schedule(self, sim):
        for (tick, op, node, slot) in self.ops:
            sim.inject(tick, op, node, slot)


def full_mesh_leave(n=256, fanout=3, seed=1):
    """configs[0]: 256-node full mesh, fanout 3; node 0 leaves gracefully, node 1 re-announces its join."""
    ops = [(0, Op.LEAVE, 0, 0), (0, Op.JOIN, 1, 0)]
    return Scenario(f"full_mesh_{n}", n, 2, full_mesh_graph(n), [0, 1], ops, dict(fanout=fanout, seed=seed))


def random_graph_leave(n=100_000, degree=16, fanout=3, seed=1, slots=1, graph_seed=7):
    """configs[1]: random graph (each node draws `degree` out-neighbours), one leave intent at tick 0."""
    subjects = np.arange(slots, dtype=np.uint32) * np.uint32(max(1, n // max(1, slots)))
    ops = [(0, Op.LEAVE, int(subjects[s]), 0) for s in range(slots)]
    return Scenario(f"random_{n}_d{degree}_f{fanout}_r{slots}", n, slots, random_regular_graph(n, degree, graph_seed), subjects, ops,
                    dict(fanout=fanout, seed=seed))


def random_graph_fail(n=100_000, degree=16, fanout=3, seed=1, graph_seed=7):
    """Failure detection: subject 0 crashes at tick 0, subject 1 leaves; SWIM probe → suspect → dead → Failed."""
    subjects = [5, n // 2]
    ops = [(0, Op.FAIL, 5, 0), (0, Op.LEAVE, n // 2, 0)]
    return Scenario(f"random_fail_{n}", n, 2, random_regular_graph(n, degree, graph_seed), subjects, ops,
                    dict(fanout=fanout, seed=seed), max_ticks=5000)


def small_world_churn(n=1_000_000, k=16, beta=0.1, churn_frac=0.05, slots=8, window=200, seed=1, graph_seed=11, fanout=3):
    """configs[2]: Watts–Strogatz small world, `churn_frac` of the nodes fail / rejoin at random ticks in
    [0, window); `slots` tracked subjects are sampled from the churn set."""
    rng = np.random.Generator(np.random.Philox(seed + 1000))
    n_churn = max(slots, int(n * churn_frac))
    churn = rng.choice(n, size=n_churn, replace=False).astype(np.uint32)
    subjects = churn[:slots]
    ops = []
    for node in churn:
        t_fail = int(rng.integers(0, window))
        t_back = t_fail + 1 + int(rng.integers(1, window))
        ops.append((t_fail, Op.FAIL, int(node), 0))
        if rng.random() < 0.5:
            ops.append((t_back, Op.REJOIN, int(node), 0))
    return Scenario(f"small_world_{n}_churn", n, slots, small_world_graph(n, k, beta, graph_seed), subjects, ops,
                    dict(fanout=fanout, seed=seed), max_ticks=20000)


def dissemination_storm(n=10_000_000, degree=16, fanout=4, slots=1, seed=1, graph_seed=7, waves=1, spacing=8, with_fail=False):
    """configs[3] on one GPU / sharded: 10 M-node random graph, fanout 4.  Every tracked subject leaves at tick
    0; with `waves` > 1 further force-leave / join operations follow every `spacing` ticks so the gossip front
    never drains (sustained activity for throughput measurements).
    with_fail=True is SURVEY §8d item 4 verbatim — "one leave-intent + one fail at tick 0": the last tracked subject crashes
    instead of leaving, so the run also carries the SWIM probe → suspect (Lifeguard confirmations) → suspicion timeout →
    dead → Failed cycle with the memberlist LAN timers (three dissemination waves and a timer wait instead of one wave)."""
    subjects = (np.arange(slots, dtype=np.uint64) * np.uint64(max(1, n // max(1, slots))) + np.uint64(3)).astype(np.uint32)
    ops = [(0, Op.LEAVE, int(subjects[s]), 0) for s in range(slots)]
    if with_fail:
        assert slots >= 2, "leave + fail needs two tracked subjects"
        ops[-1] = (0, Op.FAIL, int(subjects[-1]), 0)
    for w in range(1, waves):
        for s in range(slots):
            origin = int((int(subjects[s]) + 1 + 7919 * w) % n)
            if origin in set(int(x) for x in subjects):
                origin = (origin + 1) % n
            ops.append((w * spacing, Op.FORCE_LEAVE, origin, s))
    return Scenario(f"storm_{n}_d{degree}_f{fanout}_r{slots}_w{waves}" + ("_fail" if with_fail else ""), n, slots, random_regular_graph(n, degree, graph_seed), subjects, ops,
                    dict(fanout=fanout, seed=seed), max_ticks=4000)


def fuzz(seed, n=None, slots=None):
    """Seeded random scenario for parity fuzzing: small random graph, every operation kind, random timing,
    short suspicion timers so failures resolve within a few hundred ticks."""
    rng = np.random.Generator(np.random.Philox(seed))
    n = n or int(rng.integers(8, 400))
    slots = slots or int(rng.integers(1, 9))
    slots = min(slots, n)
    degree = int(rng.integers(2, min(n - 1, 12) + 1))
    topo = random_regular_graph(n, degree, seed + 17) if rng.random() < 0.7 else small_world_graph(n, max(2, degree // 2 * 2), 0.2, seed + 17)
    subjects = rng.choice(n, size=slots, replace=False).astype(np.uint32)
    cfg = dict(fanout=int(rng.integers(1, 6)), seed=int(rng.integers(1, 2**40)),
               retransmit_mult=int(rng.integers(1, 5)), suspicion_mult=int(rng.integers(1, 6)),
               suspicion_max_timeout_mult=int(rng.integers(1, 4)), probe_interval_ticks=int(rng.integers(0, 4)),
               gossip_interval_ms=200, init_status_ltime=int(rng.integers(0, 3)), init_clock=int(rng.integers(1, 5)),
               push_pull_interval_ticks=int(rng.choice([0, 0, 5, 11, 30])),
               reap_interval_ticks=int(rng.choice([0, 0, 7, 25])), tombstone_timeout_ticks=int(rng.integers(5, 80)),
               reconnect_timeout_ticks=int(rng.integers(5, 80)), recent_intent_timeout_ticks=int(rng.integers(5, 80)))
    ops, used = [], set()
    horizon = int(rng.integers(5, 120))
    for _ in range(int(rng.integers(1, 30))):
        t = int(rng.integers(0, horizon))
        kind = rng.choice([Op.JOIN, Op.LEAVE, Op.FORCE_LEAVE, Op.FAIL, Op.REJOIN], p=[0.15, 0.2, 0.3, 0.2, 0.15])
        s = int(rng.integers(0, slots))
        if kind == Op.FORCE_LEAVE:
            node = int(rng.integers(0, n)) if rng.random() < 0.8 else int(subjects[int(rng.integers(0, slots))])
        elif kind in (Op.FAIL, Op.REJOIN) and rng.random() < 0.3:
            node = int(rng.integers(0, n))
        else:
            node = int(subjects[s])
        if (t, node) in used:
            continue
        used.add((t, node))
        ops.append((t, int(kind), node, s))
    return Scenario(f"fuzz_{seed}", n, slots, topo, subjects, ops, cfg, max_ticks=6000)


def user_event_storm(n=100_000, degree=16, fanout=3, seed=1, graph_seed=7, n_events=4, spacing=3, alias=False, churn=0, with_leave=False):
    """`n_events` tracked user events fired by different origins `spacing` ticks apart
    (Serf::user_event, serf/api.rs:241-299).  alias=True gives events 0 and 1 the same (name, payload) and fires them
    in the same tick from origins with equal event clocks, so they land in the same ring slot with equal content and
    every node keeps exactly one of the two.  `churn` nodes crash during the storm (some return); with_leave adds a
    membership leave so both kinds of broadcast share the gossip packets."""
    rng = np.random.Generator(np.random.Philox(seed + 4242))
    content = np.arange(1, n_events + 1, dtype=np.uint32) * np.uint32(17)
    if alias and n_events >= 2:
        content[1] = content[0]
    origins = rng.choice(np.arange(2, n), size=n_events, replace=False)
    ops = []
    for e in range(n_events):
        t = 0 if (alias and e < 2) else e * spacing
        ops.append((t, Op.USER_EVENT, int(origins[e]), e))
    busy = {(t, node) for (t, _, node, _) in ops}
    pool = np.setdiff1d(np.arange(2, n), origins)
    for node in rng.choice(pool, size=min(churn, pool.size), replace=False):
        t_fail = int(rng.integers(0, max(1, n_events * spacing + 4)))
        if (t_fail, int(node)) in busy:
            continue
        ops.append((t_fail, Op.FAIL, int(node), 0))
        if rng.random() < 0.5:
            ops.append((t_fail + 2 + int(rng.integers(0, 6)), Op.REJOIN, int(node), 0))
    subjects = [0]
    if with_leave:
        ops.append((1, Op.LEAVE, 0, 0))
    return Scenario(f"user_events_{n}_e{n_events}", n, 1, random_regular_graph(n, degree, graph_seed), subjects, ops,
                    dict(fanout=fanout, seed=seed), max_ticks=3000, user_events=content)


def byzantine_injectors(n=100_000, degree=16, fanout=4, frac=0.01, delta=2, seed=1, graph_seed=7, slots=2, churn=True):
    """BASELINE configs[4] shape: random graph, `frac` of the nodes re-inject stale (status_time - delta,
    incarnation - delta) copies of their views every tick; the run carries a leave, a crash with probing (so
    incarnations and Lamport times actually move) and a rejoin."""
    rng = np.random.Generator(np.random.Philox(seed + 777))
    subjects = [3, n // 2][:slots]
    byz = rng.choice(np.arange(8, n), size=max(1, int(n * frac)), replace=False)
    byz = byz[~np.isin(byz, subjects)]
    ops = [(0, Op.LEAVE, subjects[0], 0)]
    if slots > 1 and churn:
        ops += [(2, Op.FAIL, subjects[1], 0), (30, Op.REJOIN, subjects[1], 0)]
    cfg = dict(fanout=fanout, seed=seed, init_clock=12)      # Lamport times well above delta, so stale copies beat the bootstrap views
    if churn:
        cfg.update(suspicion_mult=2, suspicion_max_timeout_mult=2, probe_interval_ticks=2)
    return Scenario(f"byzantine_{n}_f{frac}", n, slots, random_regular_graph(n, degree, graph_seed), subjects, ops, cfg,
                    max_ticks=4000, byzantine=byz, delta=delta)


def remove_failed_node_prune(n=3, seed=1, at=40):
    """The reference's `serf_remove_failed_node_prune` (serf/base/tests/serf/remove.rs:95-165) as a scenario: node 1 crashes, is
    detected Failed (short timers), then node 0 calls remove_failed_node_prune(node 1): the leave intent carries `prune` and every
    node that accepts it erases the member — the survivors end with n - 1 members."""
    return Scenario(f"prune_{n}", n, 1, full_mesh_graph(n) if n <= 512 else random_regular_graph(n, 16, 7), [1],
                    [(0, Op.FAIL, 1, 0), (at, Op.FORCE_LEAVE_PRUNE, 0, 0)],
                    dict(fanout=min(3, n - 1), seed=seed, suspicion_mult=2, suspicion_max_timeout_mult=2, probe_interval_ticks=1), max_ticks=2000)


def fuzz_prune(seed, n=None, slots=None):
    """fuzz() with most force-leave operations pruning (Serf::remove_failed_node_prune) and a few more of them."""
    sc = fuzz(seed, n=n, slots=slots)
    sc.name = f"fuzz_prune_{seed}"
    rng = np.random.Generator(np.random.Philox(seed + 424242))
    used = {(t, node) for (t, _, node, _) in sc.ops}
    sc.ops = [(t, int(Op.FORCE_LEAVE_PRUNE) if (op == Op.FORCE_LEAVE and rng.random() < 0.7) else op, node, s) for (t, op, node, s) in sc.ops]
    for _ in range(int(rng.integers(1, 6))):
        t, node, s = int(rng.integers(0, 60)), int(rng.integers(0, sc.n)), int(rng.integers(0, sc.slots))
        if (t, node) not in used:
            used.add((t, node))
            sc.ops.append((t, int(Op.FORCE_LEAVE_PRUNE), node, s))
    return sc


def fuzz_features(seed, n=None, slots=None):
    """fuzz() plus the optional subsystems on top: tracked user events fired at random ticks / origins (possibly with
    equal content), and a random set of byzantine injectors — on top of whatever fuzz() drew (push-pull rounds, reaper, probing)."""
    sc = fuzz(seed, n=n, slots=slots)
    sc.name = f"fuzz_features_{seed}"
    rng = np.random.Generator(np.random.Philox(seed + 90001))
    used = {(t, node) for (t, _, node, _) in sc.ops}
    if rng.random() < 0.8:
        E = int(rng.integers(1, 9))
        sc.user_events = rng.integers(1, 4, size=E).astype(np.uint32)          # few distinct contents → aliases
        for e in range(E):
            for _ in range(8):
                t, node = int(rng.integers(0, 40)), int(rng.integers(0, sc.n))
                if (t, node) not in used:
                    used.add((t, node))
                    sc.ops.append((t, int(Op.USER_EVENT), node, e))
                    break
    if rng.random() < 0.6:
        k = int(rng.integers(1, max(2, sc.n // 4)))
        sc.byzantine = rng.choice(sc.n, size=k, replace=False).astype(np.uint32)
        sc.delta = int(rng.integers(0, 4))
        sc.cfg["init_clock"] = int(rng.integers(1, 12))
    return sc
