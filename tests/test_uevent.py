"""User-event dissemination — CPU side.

Three layers are pinned against each other without a GPU:
  RefNode.handle_user_event   the literal 512-entry ring of one serf node (oracle Part A; KATs in test_oracle_kat.py
                              transliterate tests/serf/event.rs:8-85)
  TickSim user events         N literal nodes (map-backed ring) driven tick by tick (oracle Part B)
  uevent.cuh                  the packed 16-byte / mask rules the CUDA kernel runs, compiled for the host by
                              tests/cpp/uevent_rules_check.cpp and wrapped in the kernel's data flow
The GPU parity proper is tests/test_gpu_z_uevent.py.
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle_lib import RefNode, lib as oracle_lib_handle, oracle_sim
from serf_b200 import scenarios
from serf_b200.sim import Op, random_regular_graph

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
u32p, u64p = C.POINTER(C.c_uint32), C.POINTER(C.c_uint64)


@pytest.fixture(scope="module")
def uecheck(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("uecheck") / "uecheck.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wall", "-Wno-unknown-pragmas", "-o", so,
                           os.path.join(ROOT, "tests", "cpp", "uevent_rules_check.cpp")])
    L = C.CDLL(so)
    L.uecheck_run.restype = C.c_int
    L.uecheck_run.argtypes = [C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p,
                              C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.uecheck_handle_seq.restype = None
    L.uecheck_handle_seq.argtypes = [C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p]
    return L


def _oracle_handle_seq(content, ltimes, limit, seq):
    L = oracle_lib_handle()
    L.oracle_ue_handle_seq.restype = None
    L.oracle_ue_handle_seq.argtypes = [C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p]
    content, ltimes, seq = (np.ascontiguousarray(a, dtype=np.uint32) for a in (content, ltimes, seq))
    out, rec = np.zeros(len(seq), dtype=np.int32), np.zeros(4, dtype=np.uint32)
    L.oracle_ue_handle_seq(len(content), content.ctypes.data, ltimes.ctypes.data, limit, len(seq), seq.ctypes.data, out.ctypes.data, rec.ctypes.data)
    return out, rec


def _mask_handle_seq(uecheck, content, ltimes, limit, seq):
    content, ltimes, seq = (np.ascontiguousarray(a, dtype=np.uint32) for a in (content, ltimes, seq))
    out, rec = np.zeros(len(seq), dtype=np.int32), np.zeros(4, dtype=np.uint32)
    uecheck.uecheck_handle_seq(len(content), content.ctypes.data, ltimes.ctypes.data, limit, len(seq), seq.ctypes.data, out.ctypes.data, rec.ctypes.data)
    return out, rec


# ring collisions on purpose: equal ltimes, ltimes 512 apart (same ring slot, quirk ii), far-apart ltimes (too-old window)
LTIME_POOL = [1, 1, 2, 5, 5, 517, 1029, 7, 519, 600, 1200, 2000, 2001, 3000]


@pytest.mark.parametrize("seed", range(40))
def test_handler_three_way_agreement(uecheck, seed):
    """RefNode (512-entry ring) ≡ TickSim literal node ≡ packed-mask rules, on adversarial arrival sequences."""
    rng = np.random.default_rng(seed)
    E = int(rng.integers(1, 9))
    content = rng.integers(1, 4, size=E).astype(np.uint32)            # few distinct contents → equal (name, payload) pairs
    ltimes = rng.choice(LTIME_POOL, size=E).astype(np.uint32)
    seq = rng.integers(0, E, size=int(rng.integers(1, 40))).astype(np.uint32)
    limit = 12
    o_out, o_rec = _oracle_handle_seq(content, ltimes, limit, seq)
    m_out, m_rec = _mask_handle_seq(uecheck, content, ltimes, limit, seq)
    assert (o_out == m_out).all(), (content, ltimes, seq, o_out, m_out)
    assert (o_rec == m_rec).all(), (content, ltimes, seq, o_rec, m_rec)
    ref = RefNode()
    for i, e in enumerate(seq):
        acc = ref.L.ref_handle_user_event(ref.p, int(ltimes[e]), str(int(content[e])).encode(), b"payload")
        assert bool(acc) == (o_out[i] == 0), (i, content, ltimes, seq)
    assert ref.clock(1) == o_rec[0]                                   # event clock


def test_quirk_ring_slot_reused_without_ltime_check():
    """ltimes 5 and 517 share ring slot 5; equal (name, payload) → the later one is dropped as a
    duplicate although its Lamport time differs (serf/base.rs:784-808)."""
    out, rec = _oracle_handle_seq([9, 9, 4], [5, 517, 517], 8, [0, 1, 2, 1])
    assert out.tolist() == [0, 1, 0, 1]
    assert rec[0] == 518 and (rec[1] & 0xff) == 0b101 and (rec[1] >> 8) == 0b001     # event 0 created the slot; event 2 was pushed into it


def test_too_old_window():
    """serf/base.rs:771-781: once the event clock is past 512, an event more than 512 behind it is dropped
    (clock 2001: ltime 1488 < 2001 - 512 is old, 1489 is not)."""
    out, _ = _oracle_handle_seq([1, 2, 3, 4], [2000, 1400, 1488, 1489], 8, [0, 1, 2, 3])
    assert out.tolist() == [0, 2, 2, 0]


# ---------------------------------------------------------------------------------------------------------------
# N-node runs: oracle TickSim (literal) vs the packed rules in the kernel's data flow (host-compiled)
# ---------------------------------------------------------------------------------------------------------------
def _run_mask_model(uecheck, sc, n_ticks, limit, slots=1):
    ops = [(t, int(op), node, slot) for (t, op, node, slot) in sc.ops]
    a = [np.ascontiguousarray([o[i] for o in ops], dtype=np.uint32) for i in range(4)]
    row_ptr = np.ascontiguousarray(sc.row_ptr, dtype=np.uint64)
    col = np.ascontiguousarray(sc.col, dtype=np.uint32)
    content = np.ascontiguousarray(sc.user_events, dtype=np.uint32)
    rec = np.zeros((sc.n, 4), dtype=np.uint32)
    rows = np.zeros((n_ticks, 5), dtype=np.uint64)
    tot = np.zeros(5, dtype=np.uint64)
    lt = np.zeros(8, dtype=np.uint32)
    rc = uecheck.uecheck_run(sc.n, row_ptr.ctypes.data, col.ctypes.data, sc.cfg["fanout"], sc.cfg["seed"], limit, slots,
                             len(content), content.ctypes.data, len(ops), a[0].ctypes.data, a[1].ctypes.data, a[2].ctypes.data, a[3].ctypes.data,
                             n_ticks, rec.ctypes.data, rows.ctypes.data, tot.ctypes.data, lt.ctypes.data)
    assert rc == 0
    return rec, rows, tot, lt


def _compare_with_mask_model(uecheck, sc):
    o = sc.build(oracle_sim, trace=1)
    ticks, ok = o.run_until_converged(sc.max_ticks)
    assert ok
    n_ticks = o.stats()["tick"]
    L = oracle_lib_handle()
    limit = L.oracle_retransmit_limit(4, sc.n)
    rec, rows, tot, lt = _run_mask_model(uecheck, sc, n_ticks, limit)
    orec = o.user_event_records().view(np.uint32).reshape(sc.n, 4)
    bad = np.nonzero((orec != rec).any(axis=1))[0]
    assert bad.size == 0, f"event record of node {bad[0]}: oracle {orec[bad[0]]} packed {rec[bad[0]]}"
    st = o.user_event_stats()
    assert [st[k] for k in ("messages", "edge_updates", "delivered", "duplicates", "too_old")] == tot.tolist()
    for e in range(len(sc.user_events)):
        assert o.user_event_ltime(e) == lt[e]
    return o, rows, n_ticks


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_packed_rules_equal_literal_model(uecheck, seed):
    """Only user events (and crashes of untracked nodes) are scheduled, so the membership half contributes nothing to
    the trace rows: the oracle's rows must equal the packed model's, tick by tick, and the per-tick hash must differ
    by exactly the membership hash of the same run without user events."""
    sc = scenarios.user_event_storm(3000, 12, 3, seed=seed, n_events=5, spacing=2, churn=40)
    o, rows, n_ticks = _compare_with_mask_model(uecheck, sc)
    tr = o.tick_trace(0, n_ticks)
    assert (tr["edge_updates"] == rows[:, 0]).all() and (tr["messages"] == rows[:, 1]).all()
    assert (tr["changed"] == rows[:, 2]).all() and (tr["pending"] == rows[:, 3]).all()
    # the per-tick hash is additive: membership part (same run without the user events) + event records
    plain = scenarios.Scenario(sc.name, sc.n, sc.slots, (sc.row_ptr, sc.col), sc.subjects, [op for op in sc.ops if op[1] != Op.USER_EVENT], sc.cfg)
    m = plain.build(oracle_sim, trace=1)
    m.step(n_ticks)
    assert (tr["hash"] - m.tick_trace(0, n_ticks)["hash"] == rows[:, 4]).all()          # u64 wrap-around arithmetic


@pytest.mark.parametrize("fanout,events", [(1, 2), (4, 8), (8, 3)])
def test_packed_rules_fanouts_and_event_counts(uecheck, fanout, events):
    sc = scenarios.user_event_storm(1500, 10, fanout, seed=7, n_events=events, spacing=1, churn=10)
    _compare_with_mask_model(uecheck, sc)


def test_aliased_events_share_a_ring_slot(uecheck):
    """Events 0 and 1: same (name, payload), same Lamport time (both origins fire at event clock 1) → same ring slot,
    equal content: every node delivers exactly ONE of them, whichever reached it first, and re-broadcasts only that one."""
    sc = scenarios.user_event_storm(2000, 12, 3, seed=5, n_events=3, spacing=2, alias=True)
    o, _, _ = _compare_with_mask_model(uecheck, sc)
    assert o.user_event_ltime(0) == o.user_event_ltime(1) == 1
    s0, s1 = o.user_event_seen(0), o.user_event_seen(1)
    assert ((s0 + s1) == 1).all()
    assert s0.sum() > 0 and s1.sum() > 0
    rec = o.user_event_records()
    # exactly one event created each occupied ring slot
    lts = np.array([o.user_event_ltime(e) for e in range(3)])
    for slot in np.unique(lts % 512):
        group = int(sum(1 << e for e in range(3) if lts[e] % 512 == slot))
        occupied = (rec["seen"] & group) != 0
        firsts = rec["first"] & group
        assert (np.bitwise_count(firsts[occupied]) == 1).all() and (firsts[~occupied] == 0).all()


def test_full_dissemination_and_counters():
    sc = scenarios.user_event_storm(4000, 16, 3, seed=2, n_events=4, spacing=3)
    o = sc.build(oracle_sim, trace=1)
    ticks, ok = o.run_until_converged(sc.max_ticks)
    assert ok
    st = o.user_event_stats()
    for e in range(4):
        assert o.user_event_seen(e).sum() >= sc.n - 2                  # a random digraph may strand a node or two
    assert st["delivered"] == sum(int(o.user_event_seen(e).sum()) for e in range(4))
    assert st["event_queue"] == 0 and st["too_old"] == 0
    # origins stamp their event clock, then increment it (serf/api.rs:264, 285); everyone witnesses ltime + 1
    lts = [o.user_event_ltime(e) for e in range(4)]
    assert st["event_time"] == max(lts) + 1
    assert (o.event_time()[o.user_event_seen(3) == 1] == max(lts) + 1).all()
    tot = o.stats()
    assert tot["edge_updates"] == st["edge_updates"] and tot["messages"] == st["messages"]      # no membership traffic in this run


def test_threads_do_not_change_user_event_results():
    L = oracle_lib_handle()
    sc = scenarios.user_event_storm(3000, 12, 3, seed=4, n_events=6, spacing=1, churn=30, with_leave=True)
    res = []
    for th in (1, 3):
        o = sc.build(oracle_sim, trace=1)
        L.oracle_sim_set_threads(o._h, th)
        o.reset(sc.cfg["seed"])
        o.set_user_events(sc.user_events)
        sc.schedule(o)
        t = o.run_until_converged(sc.max_ticks)
        res.append((t, o.state_hash(), o.user_event_stats(), o.tick_trace().tobytes()))
    assert res[0] == res[1]


def test_user_events_share_packets_with_intents():
    """A leave intent and user events in flight together: both halves converge, and the rows are the sum of both."""
    sc = scenarios.user_event_storm(3000, 12, 3, seed=9, n_events=3, spacing=2, with_leave=True)
    o = sc.build(oracle_sim, trace=1)
    ticks, ok = o.run_until_converged(sc.max_ticks)
    assert ok
    st, tot = o.user_event_stats(), o.stats()
    assert tot["messages"] > st["messages"] > 0 and tot["edge_updates"] > st["edge_updates"]
    assert (o.member_status(0)[1:] == 3).sum() >= sc.n - 3


def test_inject_validation():
    sc = scenarios.user_event_storm(200, 8, 3, seed=1, n_events=2)
    o = sc.build(oracle_sim)
    from serf_b200.sim import SerfsimError
    with pytest.raises(SerfsimError):
        o.user_event(5, 0, tick=9)            # event 0 is already scheduled: a tracked event fires once
    with pytest.raises(SerfsimError):
        o.user_event(5, 2, tick=9)            # only 2 tracked events
