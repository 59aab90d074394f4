"""Cross-model check (CPU): the packed-record tick model (oracle Part B — what the CUDA kernel reproduces) against a
LITERAL multi-node execution: N full serf nodes (oracle Part A, the code the reference KATs pin), each with its own
member table of all N members and a multi-entry TransmitLimitedQueue, exchanging real message lists tick by tick with
the same peer selection.  Serf-only scenarios (join / force-leave operations incl. refutation of a leave about a live
node); the memberlist layer is not part of the literal model."""
import ctypes as C

import numpy as np
import pytest

from oracle_lib import lib, oracle_sim
from serf_b200 import Op, random_regular_graph

u64p, u32p = C.POINTER(C.c_uint64), C.POINTER(C.c_uint32)


def _faithful():
    L = lib()
    L.faithful_new.restype, L.faithful_new.argtypes = C.c_void_p, [C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint64, C.c_uint64, C.c_uint64]
    L.faithful_free.argtypes = [C.c_void_p]
    L.faithful_set_topology.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.faithful_set_subjects.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32]
    L.faithful_inject.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint64]
    L.faithful_step.argtypes = [C.c_void_p, C.c_uint32]
    L.faithful_view.restype, L.faithful_view.argtypes = C.c_int, [C.c_void_p, C.c_uint32, C.c_uint64, C.POINTER(C.c_uint8), u64p]
    L.faithful_clock.restype, L.faithful_clock.argtypes = C.c_uint64, [C.c_void_p, C.c_uint32]
    L.faithful_queue_len.restype, L.faithful_queue_len.argtypes = C.c_uint32, [C.c_void_p, C.c_uint32]
    L.faithful_inflight.restype, L.faithful_inflight.argtypes = C.c_uint32, [C.c_void_p]
    return L


def _run(seed, overlap, prune=False):
    rng = np.random.default_rng(seed)
    n, deg, fan, rm = int(rng.integers(60, 200)), int(rng.integers(4, 10)), int(rng.integers(2, 5)), int(rng.integers(2, 5))
    row_ptr, col = random_regular_graph(n, deg, seed + 3)
    slots = 3
    subjects = rng.choice(n, size=slots, replace=False).astype(np.uint32)
    ops, used = [], set()
    spacing = 3 if overlap else 45                      # ≥ 45 ticks apart: the previous intent about a subject has drained
    for s in range(slots):
        t = int(rng.integers(0, 5))
        for _ in range(int(rng.integers(1, 4))):
            kind = Op.JOIN if rng.random() < 0.4 else Op.FORCE_LEAVE
            if prune and kind == Op.FORCE_LEAVE and rng.random() < 0.6:
                kind = Op.FORCE_LEAVE_PRUNE                # Serf::remove_failed_node_prune: receivers erase the member
            node = int(subjects[s]) if kind == Op.JOIN else int(rng.integers(0, n))
            if (t, node) not in used:
                used.add((t, node)); ops.append((t, int(kind), node, s))
            t += spacing + int(rng.integers(0, 3))
    horizon = max(t for t, *_ in ops) + 80
    seed64 = int(rng.integers(1, 2**40))
    # Part B
    o = oracle_sim(n, slots, seed=seed64, fanout=fan, retransmit_mult=rm, probe_interval_ticks=0)
    o.set_topology(row_ptr, col); o.set_subjects(subjects)
    for (t, k, node, s) in ops:
        o.inject(t, k, node, s)
    o.step(horizon)
    # Part C
    L = _faithful()
    f = L.faithful_new(n, fan, rm, seed64, 1, 2)
    L.faithful_set_topology(f, row_ptr.ctypes.data, col.ctypes.data)
    sub64 = subjects.astype(np.uint64)
    L.faithful_set_subjects(f, sub64.ctypes.data, slots)
    for (t, k, node, s) in ops:
        L.faithful_inject(f, t, k, node, int(subjects[s]))
    L.faithful_step(f, horizon)
    assert L.faithful_inflight(f) == 0 and all(L.faithful_queue_len(f, v) == 0 for v in range(n)), "literal model not quiescent"
    assert o.stats()["pending"] == 0
    diffs = 0
    clk = o.lamport_time()
    for s in range(slots):
        st_b, lt_b = o.member_status(s), o.status_ltime(s)
        for v in range(n):
            st, lt = C.c_uint8(), C.c_uint64()
            if not L.faithful_view(f, v, int(subjects[s]), C.byref(st), C.byref(lt)):
                assert prune                               # the member is not in this node's table: erased by a pruning leave intent
                st.value, lt.value = 0, 0                  # the tick model reports MemberStatus::None, status time 0 for an unknown member
            diffs += int((st.value, lt.value) != (int(st_b[v]), int(lt_b[v])))
    diffs += sum(int(L.faithful_clock(f, v) != int(clk[v])) for v in range(n))
    L.faithful_free(f)
    return diffs, n * slots + n


@pytest.mark.parametrize("seed", range(12))
def test_tick_model_equals_literal_nodes_when_intents_do_not_overlap(seed):
    diffs, total = _run(seed, overlap=False)
    assert diffs == 0, f"{diffs} of {total} (status, status_time, clock) values differ from the literal multi-node run"


def test_prune_against_literal_nodes():
    """remove_failed_node_prune (serf/api.rs:513): LeaveMessage.prune travels with the intent and every receiver that accepts it
    erases the member (handle_prune, serf/base.rs:1628-1653).  Operations far enough apart that no two different leave intents
    about one subject meet.  One modelling deviation remains and is measured here (DESIGN rules P-1 / P-2): in the reference an
    erased member is unknown again, so the NEXT copy of the same pruning intent is buffered as a fresh intent and re-queued with a
    fresh budget.  The tick model delivers one reduced inbox word per (node, subject, tick): when two copies reach a node in the
    tick of its first reception it sees the second one a tick later than the literal node does (or never, if no further copy
    comes), so a few re-queues start a tick later."""
    tot_d = tot = exact = 0
    for seed in range(12):
        d, t = _run(200 + seed, overlap=False, prune=True)
        tot_d += d; tot += t; exact += d == 0
    assert exact >= 6 and tot_d <= 0.002 * tot, (exact, tot_d, tot)       # measured: 8 of 12 scenarios exact, 4 of 5532 values differ (0.07 %)


def test_overlapping_same_kind_intents_deviation_is_small():
    """Modelling rule 2 (two-entry queue): a newer accepted intent supersedes the queued one of the same kind, the
    reference keeps gossiping both.  With operations 3 ticks apart the models may differ; measure, do not hide."""
    tot_d = tot = 0
    for seed in range(12):
        d, t = _run(100 + seed, overlap=True)
        tot_d += d; tot += t
    assert tot_d <= 0.005 * tot, (tot_d, tot)          # measured: 0 of 5844 values over these 12 scenarios
