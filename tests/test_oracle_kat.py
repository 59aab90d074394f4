"""Pins the CPU oracle (oracle/serf_oracle.cpp, Part A) against the reference's own
known-answer tests, replayed literally.  Every test names the reference test it
transliterates; paths are relative to /root/reference/serf-core/src/ (SURVEY.md §8c).

CPU only.  If one of these fails the oracle is wrong, and no GPU parity claim stands.
"""
import ctypes as C

import pytest

from oracle_lib import RefNode, lib, u32p

ALIVE, LEAVING, LEFT, FAILED, NONE = RefNode.ALIVE, RefNode.LEAVING, RefNode.LEFT, RefNode.FAILED, RefNode.NONE
JOIN, LEAVE = RefNode.JOIN, RefNode.LEAVE


# ---- types/clock.rs:175-191  test_lamport_clock ------------------------------------------
def test_lamport_clock():
    L = lib()
    assert L.lamport_new_time() == 0                       # LamportClock::new().time() == 0
    # RefNode's clock starts at 1 (Serf::new increments once, base.rs:198-200) == "increment → 1"
    n = RefNode()
    assert n.clock() == 1
    n.witness(41)
    assert n.clock() == 42
    n.witness(41)
    assert n.clock() == 42
    n.witness(30)
    assert n.clock() == 42
    assert n.increment() == 43


# ---- serf/base/tests/serf.rs:873-950  test_recent_intent ---------------------------------
def test_recent_intent():
    n = RefNode()
    now = 100_000
    expire, save = now - 2000, now
    assert n.recent_intent("foo", JOIN) is None
    assert n.upsert_intent("foo", JOIN, 1, expire)
    assert n.upsert_intent("bar", LEAVE, 2, expire)
    assert n.upsert_intent("baz", JOIN, 3, save)
    assert n.upsert_intent("bar", JOIN, 4, expire)
    assert not n.upsert_intent("bar", JOIN, 0, expire)
    assert n.upsert_intent("bar", JOIN, 5, expire)
    assert n.recent_intent("foo", JOIN) == 1
    assert n.recent_intent("bar", JOIN) == 5
    assert n.recent_intent("baz", JOIN) == 3
    assert n.recent_intent("tubez", JOIN) is None
    n.L.ref_reap_intents(n.p, now, 1000)
    assert n.recent_intent("foo", JOIN) is None
    assert n.recent_intent("bar", JOIN) is None
    assert n.recent_intent("baz", JOIN) == 3
    assert n.recent_intent("tubez", JOIN) is None
    n.L.ref_reap_intents(n.p, now + 2000, 1000)
    assert n.recent_intent("baz", JOIN) is None


# ---- serf/base/tests/serf/join.rs ---------------------------------------------------------
def test_join_intent_buffer_early():          # join.rs:8-35
    n = RefNode()
    assert n.join_intent(10, "test"), "should rebroadcast"
    assert not n.join_intent(10, "test"), "should not rebroadcast"
    assert n.recent_intent("test", JOIN) == 10


def test_join_intent_old_message():           # join.rs:38-85
    n = RefNode()
    n.insert_member("test", ALIVE, 12)
    assert not n.join_intent(10, "test")
    assert n.recent_intent("test", JOIN) is None


def test_join_intent_newer():                 # join.rs:88-134
    n = RefNode()
    n.insert_member("test", ALIVE, 12)
    assert n.join_intent(14, "test")
    assert n.recent_intent("test", JOIN) is None
    assert n.member("test") == (ALIVE, 14)
    assert n.clock() == 15


def test_join_intent_reset_leaving():         # join.rs:137-185
    n = RefNode()
    n.insert_member("test", LEAVING, 12)
    assert n.join_intent(14, "test")
    assert n.recent_intent("test", JOIN) is None
    assert n.member("test") == (ALIVE, 14)
    assert n.clock() == 15


def test_join_pending_intent():               # join.rs:267-302
    n = RefNode()
    n.upsert_intent("test", JOIN, 5)
    n.node_join("test")
    assert n.member("test") == (ALIVE, 5)


def test_join_pending_intents():              # join.rs:305-347
    n = RefNode()
    n.upsert_intent("test", JOIN, 5)
    n.upsert_intent("test", LEAVE, 6)
    n.node_join("test")
    assert n.member("test") == (LEAVING, 6)


def test_join_leave_ltime():                  # join.rs:188-264 and SURVEY Appendix A.8 worked check
    s1, s2 = RefNode(1), RefNode(2)
    # memberlist join: each learns the other (notify_join → handle_node_join)
    s1.node_join(2); s2.node_join(1)
    assert s1.clock() == 1
    s1.L.ref_api_join(s1.p)                                # api.rs:339-342
    q = s1.queue()
    assert q == [(JOIN, 1, 1, 0)]                          # Join{ltime 1, id s1}
    assert s1.clock() == 2 and s1.member(1) == (ALIVE, 1)
    assert s2.join_intent(1, 1)                            # gossip delivery
    assert s2.member(1)[1] == 1                            # status_time == 1
    assert s2.clock() > s2.member(1)[1]
    old = s2.clock()
    assert s1.L.ref_api_leave(s1.p) == 0                   # api.rs:422-499
    assert s1.queue()[-1][:3] == (LEAVE, 2, 1)
    assert s1.clock() == 3 and s1.member(1) == (LEAVING, 2)
    assert s2.leave_intent(2, 1)
    assert s2.clock() > old, "leave should increment"
    assert s2.member(1) == (LEAVING, 2)


# ---- serf/base/tests/serf/leave.rs --------------------------------------------------------
def test_leave_intent_buffer_early():         # leave.rs:4-32
    n = RefNode()
    assert n.leave_intent(10, "test")
    assert not n.leave_intent(10, "test")
    assert n.recent_intent("test", LEAVE) == 10


def test_leave_intent_old_message():          # leave.rs:35-82
    n = RefNode()
    n.insert_member("test", ALIVE, 12)
    assert not n.leave_intent(10, "test")
    assert n.recent_intent("test", LEAVE) is None


def test_leave_intent_newer():                # leave.rs:85-136
    n = RefNode()
    n.insert_member("test", ALIVE, 12)
    assert n.leave_intent(14, "test")
    assert n.recent_intent("test", LEAVE) is None
    assert n.member("test")[0] == LEAVING
    assert n.clock() == 15


# ---- serf/base/tests/serf/delegate.rs:117-180  delegate_merge_remote_state ----------------
def test_delegate_merge_remote_state():
    n = RefNode()
    ids = (C.c_uint64 * 2)(n.id("test"), n.id("foo"))
    lts = (C.c_uint64 * 2)(20, 15)
    left = (C.c_uint64 * 1)(n.id("foo"))
    n.L.ref_merge_remote_state(n.p, 42, ids, lts, 2, left, 1, 50, 100)
    assert n.clock(0) == 42, "bad lamport clock"
    assert n.recent_intent("test", JOIN) == 20, "bad join ltime"
    assert n.recent_intent("foo", LEAVE) == 16, "bad leave ltime"
    assert n.clock(1) == 50, "bad event clock"
    assert n.clock(2) == 100, "bad query clock"


# ---- serf/base/tests/serf/delegate.rs:40-114  delegate_local_state (membership part) ------------
def test_delegate_local_state_round_trip():
    """local_state carries clock.time(), every member's status_time and the ids of the left list; feeding it to another
    node's merge_remote_state reproduces the intents the reference KAT above pins (Left member → Leave at ltime + 1)."""
    a = RefNode(1)
    a.node_join(2); a.node_join(3)
    assert a.join_intent(7, 2)                                   # member 2 joined at ltime 7
    assert a.leave_intent(9, 3); a.node_leave(3, 0)             # member 3 left gracefully: Leaving → Left, on the left list
    cap = 8
    ids = (C.c_uint64 * cap)(); lts = (C.c_uint64 * cap)(); left = (C.c_uint64 * cap)()
    ppl, evl, ql, nleft = C.c_uint64(), C.c_uint64(), C.c_uint64(), C.c_uint32()
    n = a.L.ref_local_state(a.p, C.byref(ppl), ids, lts, cap, left, cap, C.byref(nleft), C.byref(evl), C.byref(ql))
    assert ppl.value == a.clock() == 10 and evl.value == a.clock(1) and ql.value == a.clock(2)
    assert n == 3 and dict(zip(list(ids)[:n], list(lts)[:n])) == {1: 0, 2: 7, 3: 9}
    assert nleft.value == 1 and left[0] == 3
    b = RefNode(50)
    b.L.ref_merge_remote_state(b.p, ppl.value, ids, lts, n, left, nleft.value, evl.value, ql.value)
    assert b.clock() == 11                                       # witness(10-1) → 10, then the leave intent 9+1 = 10 → 11
    assert b.recent_intent(2, JOIN) == 7 and b.recent_intent(3, LEAVE) == 10 and b.recent_intent(1, JOIN) == 0


# ---- serf/base/tests/serf.rs:772-788  serf_stats (fresh node) -----------------------------
def test_serf_stats_fresh_node():
    n = RefNode()
    assert n.clock(0) == 1 and n.clock(1) == 1 and n.clock(2) == 1     # member_time / event_time / query_time
    assert n.L.ref_num_members(n.p) == 1
    assert n.queue() == []


# ---- serf/base/tests/serf.rs:57-160  serf_get_queue_max -----------------------------------
def test_get_queue_max():
    L = lib()
    assert L.ref_get_queue_max(4096, 0, 1) == 4096            # default
    assert L.ref_get_queue_max(4096, 1024, 1) == 1024         # min 1024 wins with few members
    assert L.ref_get_queue_max(4096, 16, 100) == 200          # 2 · members
    assert L.ref_get_queue_max(4096, 16, 101) == 202


# ---- serf/base/tests/serf/reap.rs:41-129  serf_reap_handler -------------------------------
def test_reap_handler():
    n = RefNode()
    now = 1_000_000
    for age in (0, 5000, 10000):
        n.L.ref_push_left(n.p, n.id("foo"), NONE, 0, now - age)
    n.upsert_intent("alice", JOIN, 1, now)
    n.upsert_intent("bob", JOIN, 2, now - 10000)
    n.upsert_intent("carol", LEAVE, 1, now)
    n.upsert_intent("doug", LEAVE, 2, now - 10000)
    day = 24 * 3600 * 1000
    n.L.ref_reap(n.p, now, day, 6000, 7000)                    # tombstone 6 s, recent_intent_timeout 7 s
    assert n.L.ref_left_count(n.p) == 2
    assert n.recent_intent("alice", JOIN) is not None
    assert n.recent_intent("bob", JOIN) is None
    assert n.recent_intent("carol", LEAVE) is not None
    assert n.recent_intent("doug", LEAVE) is None


# ---- serf/base/tests/serf/remove.rs:187-222  test_remove_old_member ----------------------------
def test_remove_old_member():
    n = RefNode()
    now = 1_000_000
    n.L.ref_push_left(n.p, n.id("foo"), NONE, 0, now)
    n.L.ref_push_left(n.p, n.id("bar"), NONE, 0, now - 5000)
    n.L.ref_push_left(n.p, n.id("baz"), NONE, 0, now - 5000)
    n.L.ref_remove_old_member(n.p, 0, n.id("bar"))
    assert n.L.ref_left_count(n.p) == 2


# ---- handle_node_leave transitions (base.rs:1375-1440) + event order scenarios
#      serf/base/tests/serf/event.rs:88-232, 405-467; reconnect.rs:10-72 ---------------------
def test_event_order_join_failed_leave_forced():
    n = RefNode()
    n.node_join("x")                                            # Join
    n.node_leave("x", 10)                                       # memberlist dead → Failed
    assert n.member("x")[0] == FAILED and n.L.ref_failed_count(n.p) == 1
    assert n.leave_intent(5, "x")                               # force-leave intent → Left (base.rs:1520-1559)
    assert n.member("x")[0] == LEFT
    assert n.L.ref_failed_count(n.p) == 0 and n.L.ref_left_count(n.p) == 1
    x = n.id("x")
    assert n.events() == [(0, x), (2, x), (1, x)]               # Join, Failed, Leave


def test_event_order_join_failed_join():
    n = RefNode()
    n.node_join("x"); n.node_leave("x", 10); n.node_join("x")
    assert n.member("x")[0] == ALIVE and n.L.ref_failed_count(n.p) == 0
    x = n.id("x")
    assert n.events() == [(0, x), (2, x), (0, x)]               # Join, Failed, Join


def test_event_order_join_leave():
    n = RefNode()
    n.node_join("x")
    assert n.leave_intent(3, "x")                               # Alive → Leaving
    n.node_leave("x", 10)                                       # Leaving → Left
    assert n.member("x") == (LEFT, 3)
    x = n.id("x")
    assert n.events() == [(0, x), (1, x)]
    n.node_leave("x", 11)                                       # Left: bad state, nothing happens (base.rs:1403-1406)
    assert n.member("x") == (LEFT, 3) and len(n.events()) == 2


def test_node_join_existing_forces_alive_keeps_status_time():   # SURVEY §8c quirk (iv), base.rs:1251-1263
    n = RefNode()
    n.insert_member("x", LEAVING, 9)
    n.node_join("x")
    assert n.member("x") == (ALIVE, 9)


# ---- the always-update rule, base.rs:1482-1497; regression scenario of
#      serf/base/tests/serf/event.rs:257-402 (no infinite rebroadcast) -----------------------
def test_leave_avoid_infinite_rebroadcast():
    a, b = RefNode(1), RefNode(2)
    for n in (a, b):
        n.insert_member("s2", LEFT, 5)
    # two successive leave messages for an already-left member: the second must be rejected
    assert a.leave_intent(7, "s2") and a.member("s2") == (LEFT, 7)
    assert not a.leave_intent(7, "s2")
    assert b.leave_intent(7, "s2") and not b.leave_intent(7, "s2")
    assert not a.leave_intent(6, "s2")


def test_leave_intent_status_none_and_unknown():                # base.rs:1501, 1560-1570
    n = RefNode()
    n.insert_member("x", NONE, 1)
    assert not n.leave_intent(4, "x")
    assert n.member("x") == (NONE, 4)                           # status_time still moves (always-update)
    n.insert_member("y", 9, 1)                                  # MemberStatus::Unknown(9)
    assert n.leave_intent(4, "y") and n.member("y") == (LEAVING, 4)


def test_refute_stale_leave_about_self():                       # base.rs:1470-1480
    n = RefNode(7)
    n.witness(9)                                                # clock 10
    assert not n.leave_intent(12, 7)                            # about ourselves while Alive → refute, no rebroadcast
    assert n.member(7) == (ALIVE, 0)                            # status_time untouched by the refuted leave
    assert n.L.ref_refutes(n.p) == 1
    n.L.ref_run_detached(n.p)                                   # the detached broadcast_join(clock.time())
    assert n.queue() == [(JOIN, 13, 7, 0)]
    assert n.member(7) == (ALIVE, 13) and n.clock() == 14
    # a node that is itself leaving does not refute (api.rs:443-449 sets Leaving first)
    m = RefNode(8)
    assert m.L.ref_api_leave(m.p) == 0
    assert m.member(8) == (LEAVING, 1) and m.clock() == 2 and m.L.ref_refutes(m.p) == 0
    assert m.queue() == []                                      # no other alive member → no broadcast (api.rs:451-453)


def test_force_leave():                                         # base.rs:454-480
    n = RefNode(1)
    n.node_join(2); n.node_join(3)
    n.node_leave(3, 0)                                          # 3 failed
    n.L.ref_api_force_leave(n.p, 3, 0)
    assert n.member(3) == (LEFT, 1)                             # Leave{ltime = clock.time() = 1}
    assert n.clock() == 2                                       # witness(1) → 2, no increment
    assert n.queue() == [(LEAVE, 1, 3, 0)]


# ---- adjacent row (user events): serf/base/tests/serf/event.rs:8-85 ---------------------------
def test_user_event_old_message():           # event.rs:8-33
    n = RefNode()
    n.L.ref_event_clock_witness(n.p, 512 + 1000)                # event_buffer_size + 1000
    assert not n.L.ref_handle_user_event(n.p, 1, b"old", b""), "should not rebroadcast"


def test_user_event_same_clock():            # event.rs:36-85
    n = RefNode()
    assert n.L.ref_handle_user_event(n.p, 1, b"first", b"test"), "should rebroadcast"
    assert n.L.ref_handle_user_event(n.p, 1, b"first", b"newpayload"), "should rebroadcast"
    assert n.L.ref_handle_user_event(n.p, 1, b"second", b"other"), "should rebroadcast"
    got = [(n.L.ref_user_event_get(n.p, i, 0), n.L.ref_user_event_get(n.p, i, 1)) for i in range(n.L.ref_user_event_count(n.p))]
    assert got == [(b"first", b"test"), (b"first", b"newpayload"), (b"second", b"other")]
    assert not n.L.ref_handle_user_event(n.p, 1, b"first", b"test")        # exact duplicate: already seen (base.rs:803-808)
    assert n.L.ref_event_buffer_has(n.p, 1) and n.clock(1) == 2


# ---- external TransmitLimitedQueue restated (parity unpinned; documents the restatement) ---
def test_transmit_limited_queue_restated():
    L = lib()
    assert L.oracle_retransmit_limit(4, 256) == 12
    assert L.oracle_retransmit_limit(4, 100_000) == 24
    assert L.oracle_retransmit_limit(4, 1_000_000) == 28
    assert L.oracle_retransmit_limit(4, 10_000_000) == 32
    assert L.oracle_retransmit_limit(4, 9) == 4 and L.oracle_retransmit_limit(4, 10) == 8
    n = RefNode(1, retransmit_mult=1)
    for i in range(2, 11):
        n.node_join(i)                                          # 10 members → limit 1·ceil(log10 11) = 2
    n.L.ref_api_join(n.p)
    ty = (C.c_uint8 * 4)(); lt = (C.c_uint64 * 4)(); idv = (C.c_uint64 * 4)()
    assert n.L.ref_get_broadcasts(n.p, 1400, 2, ty, lt, idv, 4) == 1
    assert n.queue()[0][3] == 1
    assert n.L.ref_get_broadcasts(n.p, 1400, 2, ty, lt, idv, 4) == 1
    assert n.queue() == []                                      # dropped at transmits == limit
    assert n.L.ref_get_broadcasts(n.p, 1400, 2, ty, lt, idv, 4) == 0


def test_suspicion_table_restated():
    L = lib()
    out = (C.c_uint32 * 8)()
    # LAN profile at n = 100000: min = 4·5·1 s = 20 s = 100 ticks, max = 600 ticks, k = 2
    k1 = L.oracle_suspicion_table(4, 6, 5, 200, 100_000, out, 8)
    assert k1 == 3 and out[0] == 600 and out[2] == 100 and 100 < out[1] < 600
    # suspicion_mult 1 (the reference's test profile, tests.rs:25-39): k = 0 → one entry = min
    assert L.oracle_suspicion_table(1, 6, 10, 5, 4, out, 8) == 1


# ---- Philox4x32-10 known answers (Random123 kat_vectors) -----------------------------------
@pytest.mark.parametrize("ctr,key,exp", [
    ((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
    ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
    ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
     (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
])
def test_philox_kat(ctr, key, exp):
    L = lib()
    c = (C.c_uint32 * 4)(*ctr); k = (C.c_uint32 * 2)(*key); o = (C.c_uint32 * 4)()
    L.oracle_philox4x32_10(c, k, o)
    assert tuple(o) == exp
