"""CPU tests of the tick oracle (oracle/serf_oracle.cpp, Part B) and of the host logic.

  * the packed-record rules (Part B) are the SAME function as the literal serf node (Part A,
    pinned by the reference KATs in test_oracle_kat.py): random operation sequences on both;
  * the reduction lemma the CUDA kernel relies on (greatest leave + greatest join ≡ every
    message one at a time in canonical order), on random message multisets;
  * scenario-level behaviour of the tick model (the outcomes the reference's integration tests
    assert: Join→Leave, Join→Failed, Failed→Leave(forced), Failed→Join; clocks; refutation).
"""
import ctypes as C

import numpy as np
import pytest

from oracle_lib import RefNode, lib, oracle_sim
from serf_b200 import MemberStatus, MlState, Op, full_mesh_graph, random_regular_graph
from serf_b200 import scenarios
from serf_b200.sim import RECORD_DTYPE

LIMIT = 12


def _view(known, status, st):
    buf = (C.c_uint8 * 32)()
    lib().oracle_view_init(buf, int(known), int(status), int(st))
    return buf


def _rec(buf):
    return np.frombuffer(bytes(buf), dtype=RECORD_DTYPE)[0]


def _apply(buf, kind, lt, self_=0, sstate=0):
    L = lib()
    if kind == "join":
        return bool(L.oracle_view_join_intent(buf, lt, LIMIT)), False
    rf = C.c_int(0)
    acc = bool(L.oracle_view_leave_intent(buf, lt, self_, sstate, C.byref(rf), LIMIT))
    return acc, bool(rf.value)


# ---- Part A ≡ Part B ------------------------------------------------------------------------
@pytest.mark.parametrize("seed", range(40))
def test_view_rules_equal_literal_node(seed):
    rng = np.random.default_rng(seed)
    node = RefNode(self_id=1)
    known = bool(rng.integers(0, 2))
    status0 = int(rng.choice([0, 1, 2, 3, 4])) if known else 0
    st0 = int(rng.integers(0, 6)) if known else 0
    if known:
        node.insert_member("x", status0, st0)
    buf = _view(known, status0, st0)
    for _ in range(30):
        op = rng.choice(["join", "leave", "node_join", "node_leave"], p=[0.35, 0.35, 0.15, 0.15])
        lt = int(rng.integers(0, 12))
        if op == "join":
            assert node.join_intent(lt, "x") == _apply(buf, "join", lt)[0]
        elif op == "leave":
            assert node.leave_intent(lt, "x") == _apply(buf, "leave", lt)[0]
        elif op == "node_join":
            node.node_join("x"); lib().oracle_view_node_join(buf)
        else:
            node.node_leave("x", 5); lib().oracle_view_node_leave(buf, 5)
        r, m = _rec(buf), node.member("x")
        if m is None:
            assert not (r["flags"] & 1)
            it_j, it_l = node.recent_intent("x", RefNode.JOIN), node.recent_intent("x", RefNode.LEAVE)
            if r["status"] == 0:
                assert it_j is None and it_l is None
            elif r["status"] == RefNode.JOIN:
                assert it_j == r["status_ltime"]
            else:
                assert it_l == r["status_ltime"]
        else:
            assert r["flags"] & 1
            assert (int(r["status"]), int(r["status_ltime"])) == m


# ---- reduction lemma: literal canonical order ≡ greatest-leave then greatest-join -------------
@pytest.mark.parametrize("seed", range(60))
def test_reduction_lemma(seed):
    rng = np.random.default_rng(1000 + seed)
    known = bool(rng.integers(0, 4))
    status0 = int(rng.choice([0, 1, 2, 3, 4])) if known else int(rng.choice([0, 1, 2]))
    st0 = int(rng.integers(0, 8))
    self_, sstate = int(rng.integers(0, 2)), int(rng.integers(0, 3))
    leaves = sorted(int(x) for x in rng.integers(0, 12, size=rng.integers(0, 6)))
    joins = sorted(int(x) for x in rng.integers(0, 12, size=rng.integers(0, 6)))
    a, b = _view(known, status0, st0), _view(known, status0, st0)
    rf_a = rf_b = False
    for lt in leaves:
        rf_a |= _apply(a, "leave", lt, self_, sstate)[1]
    for lt in joins:
        _apply(a, "join", lt)
    if leaves:
        rf_b |= _apply(b, "leave", max(leaves), self_, sstate)[1]
    if joins:
        _apply(b, "join", max(joins))
    assert bytes(a) == bytes(b) and rf_a == rf_b


# ---- scenarios --------------------------------------------------------------------------------
def test_full_mesh_leave_converges_to_left():
    sc = scenarios.full_mesh_leave(256, seed=5)
    o = sc.build(oracle_sim, trace=1)
    ticks, ok = o.run_until_converged(500)
    assert ok and 5 < ticks < 60
    st = o.member_status(0)
    assert st[0] == MemberStatus.LEAVING                 # the leaver's own view (api.rs:443-449)
    assert (st[1:] == MemberStatus.LEFT).all()           # Alive → Leaving (intent) → Left (memberlist left)
    assert (o.status_ltime(0) == 2).all()                # Leave{ltime = init_clock}
    assert (o.lamport_time()[2:] == 3).all()             # witness(2) → 3
    assert (o.member_status(1) == MemberStatus.ALIVE).all()
    assert (o.status_ltime(1) == 2).all()                # node 1 re-announced its join at clock 2
    s = o.stats()
    assert s["disagree_slots"] == 0 and s["pending"] == 0 and s["intent_queue"] == 0
    tr = o.tick_trace()
    assert tr["edge_updates"].sum() == s["edge_updates"] and tr["hash"][-1] == o.state_hash()
    # every node sends each accepted entry exactly retransmit_limit = 12 times (3 entries: 2 intents + 1 left)
    assert s["messages"] == 256 * 12 * 3


def test_deterministic_and_seed_sensitive():
    def run(seed):
        o = scenarios.random_graph_leave(3000, 8, 3, seed=seed).build(oracle_sim, trace=1)
        o.run_until_converged(500)
        return o.tick_trace()["hash"].tolist()
    assert run(1) == run(1)
    assert run(1) != run(2)


def test_failure_detection_and_rejoin():
    n = 64
    o = oracle_sim(n, 1, seed=9, suspicion_mult=2, suspicion_max_timeout_mult=2, probe_interval_ticks=2, trace=1)
    o.set_topology(*full_mesh_graph(n))
    o.set_subjects([7])
    o.fail(7, tick=1)
    ticks, ok = o.run_until_converged(3000)
    assert ok
    st = o.member_status(0)
    assert (np.delete(st, 7) == MemberStatus.FAILED).all()            # Alive → Failed (base.rs:1394-1402)
    assert (np.delete(o.ml_state(0), 7) == MlState.DEAD).all()
    assert o.tick_trace()["suspects"].sum() >= 1
    # force-leave the failed node: Failed → Left everywhere (base.rs:1520-1559)
    o.remove_failed_node(3, 0, tick=ticks + 1)
    t2, ok = o.run_until_converged(3000)
    assert ok and (np.delete(o.member_status(0), 7) == MemberStatus.LEFT).all()
    # the node comes back: alive(inc+1) → notify_join → Alive again, status_time = the new join intent
    o.rejoin(7, tick=t2 + 1)
    t3, ok = o.run_until_converged(3000)
    assert ok
    assert (o.member_status(0) == MemberStatus.ALIVE).all()
    assert (o.incarnation(0) == 2).all() and (o.ml_state(0) == MlState.ALIVE).all()
    assert o.stats()["disagree_slots"] == 0


def test_refutation_of_forced_leave_about_alive_node():
    n = 50
    o = oracle_sim(n, 1, seed=4)
    o.set_topology(*random_regular_graph(n, 6, 3))
    o.set_subjects([10])
    o.remove_failed_node(20, 0, tick=0)                               # somebody force-leaves an ALIVE node
    ticks, ok = o.run_until_converged(500)
    assert ok
    st, lt = o.member_status(0), o.status_ltime(0)
    assert st[10] == MemberStatus.ALIVE                               # the subject refuted (base.rs:1470-1480)
    assert (st == MemberStatus.ALIVE).all()                           # …and its newer join intent wins everywhere
    assert (lt == lt[10]).all() and lt[10] > 2


def test_one_op_per_node_per_tick_and_validation():
    o = oracle_sim(16, 2)
    o.set_topology(*full_mesh_graph(16))
    o.set_subjects([3, 4])
    o.leave(3, tick=0)
    with pytest.raises(Exception):
        o.join(3, tick=0)
    with pytest.raises(Exception):
        o.leave(9, tick=0)                                            # not a tracked subject
    with pytest.raises(Exception):
        o.inject(0, Op.FORCE_LEAVE, 2, 5)                             # slot out of range
    o.fail(9, tick=0)                                                 # untracked nodes may crash (they just go silent)


@pytest.mark.parametrize("seed", range(12))
def test_fuzz_scenarios_run_and_are_reproducible(seed):
    sc = scenarios.fuzz(seed)
    a = sc.build(oracle_sim, trace=1)
    a.run_until_converged(sc.max_ticks)
    n = a.stats()["tick"]
    b = sc.build(oracle_sim, trace=1)
    b.step(n)
    assert a.state_hash() == b.state_hash()
    assert (a.tick_trace(0, n) == b.tick_trace(0, n)).all()


def test_push_pull_backstop_converges_what_gossip_strands():
    """serf/delegate.rs:386-554: with tiny retransmit budgets on a sparse graph plain gossip strands members in
    Alive / Leaving / Failed; periodic push-pull (anti-entropy) brings every view to Left."""
    from serf_b200 import small_world_graph
    n = 3000
    res = {}
    for pp in (0, 15):
        o = oracle_sim(n, 1, seed=3, fanout=2, retransmit_mult=1, push_pull_interval_ticks=pp, probe_interval_ticks=0)
        o.set_topology(*small_world_graph(n, 4, 0.05, 5))
        o.set_subjects([7])
        o.leave(7, tick=0)
        ticks, ok = o.run_until_converged(3000)
        assert ok
        res[pp] = (np.bincount(np.delete(o.member_status(0), 7), minlength=5), ticks)
    assert res[0][0][MemberStatus.LEFT] < n - 1                       # gossip alone leaves stragglers
    assert res[15][0][MemberStatus.LEFT] == n - 1                     # anti-entropy finishes the job
    assert (res[15][1] + 1) % 15 == 0                                 # convergence is declared on a push-pull round after a gossip-free interval


def test_reaper_erases_tombstones_and_allows_a_fresh_join():
    """serf/base.rs:483-610 on a tick clock: a Failed member is erased from every view once reconnect_timeout has
    passed (erase_node!), a later return of the node is a brand-new member (handle_node_join's absent branch)."""
    n = 40
    o = oracle_sim(n, 1, seed=2, suspicion_mult=2, suspicion_max_timeout_mult=2, probe_interval_ticks=2,
                   reap_interval_ticks=10, reconnect_timeout_ticks=30, tombstone_timeout_ticks=30, recent_intent_timeout_ticks=20)
    o.set_topology(*full_mesh_graph(n))
    o.set_subjects([5])
    o.fail(5, tick=0)
    o.run_until_converged(2000)
    assert (np.delete(o.member_status(0), 5) == MemberStatus.FAILED).all()
    o.step(60)                                                          # reconnect_timeout + a reap interval
    assert (np.delete(o.member_status(0), 5) == MemberStatus.NONE).all()   # Serf::members no longer lists it
    rec = np.delete(o.records(0), 5)
    assert (rec["flags"] & 1 == 0).all() and (rec["leave_tick"] == 0).all()
    t = o.stats()["tick"]
    o.rejoin(5, tick=t)
    o.run_until_converged(2000)
    st = o.member_status(0)
    assert (st == MemberStatus.ALIVE).all() and (o.incarnation(0) == 2).all()


def test_reaper_drops_stale_buffered_intents():
    n = 30
    o = oracle_sim(n, 1, seed=2, probe_interval_ticks=0, reap_interval_ticks=5, recent_intent_timeout_ticks=12,
                   tombstone_timeout_ticks=10, reconnect_timeout_ticks=10, suspicion_mult=2, suspicion_max_timeout_mult=1)
    o.set_topology(*full_mesh_graph(n))
    o.set_subjects([3])
    # make the member unknown everywhere first: it fails, is detected, and is reaped
    o2 = o
    o2.cfg.probe_interval_ticks = 0
    o.fail(3, tick=0)
    o.remove_failed_node(1, 0, tick=1)                                 # Leave intent while everybody still sees it Alive → Leaving, never reaped
    o.run_until_converged(500)
    assert (np.delete(o.member_status(0), 3) == MemberStatus.LEAVING).all()
