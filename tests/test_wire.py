"""Wire codec on the CPU: the oracle's restatement against hand-derived bytes and the reference's round-trip
property (types/tests.rs:8-25: decode(encode(m)) == m and consumed == encoded length, here incl. the message envelope of
types/message.rs:397-428), the product's host codec against the oracle byte for byte, the decode error paths of
types/{join,leave,push_pull}.rs, and the batch kernels (host build of tests/emu) against per-node oracle encodings.
Byte-level interop with a real serf node stays UNPINNED: memberlist_core::proto is not in the reference tree (wire.cuh)."""
import ctypes as C

import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

import wire_lib as W
from emu_lib import emu_sim, lib as emu_lib
from oracle_lib import oracle_sim
from serf_b200 import scenarios

U64 = st.integers(min_value=0, max_value=2**64 - 1)
SMALL = st.one_of(st.integers(0, 300), U64)


@pytest.fixture(scope="module")
def P():
    return W.bind_product(emu_lib())             # the product's wire_codec.cu compiled for the host (no GPU here)


def test_hand_derived_bytes():
    """Layout restated in wire.cuh: tag byte = tag << 3 | wire type (Byte 0, Varint 1, LengthDelimited 2), LEB128 varints."""
    # Join{ltime 5, id 9}: message byte 2<<3|2 = 0x12, payload length 4, ltime field 1<<3|1 = 0x09, id field 2<<3|1 = 0x11
    assert W.o_encode_intent(W.JOIN, 5, 9) == bytes([0x12, 4, 0x09, 5, 0x11, 9])
    # Leave{ltime 300, id 1, prune}: 300 = 0xAC 0x02; prune field 2<<3|0 = 0x10 value 1; id field 3<<3|1 = 0x19
    assert W.o_encode_intent(W.LEAVE, 300, 1, True) == bytes([0x0A, 7, 0x09, 0xAC, 0x02, 0x10, 1, 0x19, 1])
    assert W.o_encode_intent(W.LEAVE, 300, 1, False) == bytes([0x0A, 5, 0x09, 0xAC, 0x02, 0x19, 1])      # prune written only when true (leave.rs:149-156)
    # PushPull{ltime 42, status {7: 20}, left {7}, event_ltime 50, query_ltime 100} (values of delegate_merge_remote_state's KAT)
    assert W.o_encode_push_pull(42, [(7, 20)], [7], 50, 100) == bytes([0x1A, 14, 0x09, 42, 0x12, 4, 0x09, 7, 0x11, 20, 0x19, 7, 0x21, 50, 0x31, 100])


@settings(max_examples=300, deadline=None)
@given(st.sampled_from([W.JOIN, W.LEAVE]), SMALL, SMALL, st.booleans())
def test_intent_round_trip_and_product_equals_oracle(P, type_, ltime, id_, prune):
    prune = prune and type_ == W.LEAVE
    b = W.o_encode_intent(type_, ltime, id_, prune)
    assert W.o_decode_intent(b) == (0, (type_, ltime, id_, prune))                   # data_round_trip, types/tests.rs:8-25
    assert W.p_encode_intent(P, type_, ltime, id_, prune) == b
    assert W.p_decode_intent(P, b) == (0, (type_, ltime, id_, prune))
    t = C.c_uint32()
    assert P.serfsim_wire_message_type(W._buf(b), len(b), C.byref(t)) == 0 and t.value == type_


@settings(max_examples=200, deadline=None)
@given(SMALL, st.lists(st.tuples(SMALL, SMALL), max_size=12, unique_by=lambda kv: kv[0]), st.lists(SMALL, max_size=6, unique=True), SMALL, SMALL)
def test_push_pull_round_trip_and_product_equals_oracle(P, ltime, status, left, ev, q):
    b = W.o_encode_push_pull(ltime, status, left, ev, q)
    assert W.o_decode_push_pull(b) == (0, (ltime, status, left, ev, q, 0))
    assert W.p_encode_push_pull(P, ltime, status, left, ev, q) == b
    assert W.p_decode_push_pull(P, b) == (0, (ltime, status, left, ev, q, 0))


def test_decode_errors_and_unknown_fields(P):
    join = W.o_encode_intent(W.JOIN, 5, 9)
    for dec in (W.o_decode_intent, lambda b: W.p_decode_intent(P, b)):
        assert dec(join[:-1])[0] != 0                                                # truncated
        dup = bytes([0x12, 6, 0x09, 5, 0x09, 6, 0x11, 9])                            # ltime twice: duplicate_field (join.rs:66-72)
        assert dec(dup)[0] != 0
        assert dec(bytes([0x12, 2, 0x09, 5]))[0] != 0                                # id missing (join.rs:106-109)
        # unknown fields are skipped (join.rs:99, utils::skip): tag 7 varint, tag 6 length-delimited
        ext = bytes([0x12, 11, 0x39, 0x7F, 0x09, 5, 0x32, 3, 1, 2, 3, 0x11, 9])
        assert dec(ext) == (0, (W.JOIN, 5, 9, False))
        # an unknown field before the message inside the envelope stream (message.rs:684) is skipped as well
        assert dec(bytes([0x39, 1]) + join) == (0, (W.JOIN, 5, 9, False))
        assert dec(join + join)[0] != 0                                              # two messages: duplicate_field (message.rs:522-528)
        # leave.rs:103-108 does not reject a second id (the last one wins); join.rs:80-82 does
        assert dec(bytes([0x0A, 6, 0x09, 5, 0x19, 1, 0x19, 2])) == (0, (W.LEAVE, 5, 2, False))
        assert dec(bytes([0x12, 6, 0x09, 5, 0x11, 1, 0x11, 2]))[0] != 0
    pp = W.o_encode_push_pull(42, [(7, 20)], [7], 50, 100)
    for dec in (W.o_decode_push_pull, lambda b: W.p_decode_push_pull(P, b)):
        assert dec(pp[:-2] )[0] != 0
        assert dec(bytes([0x1A, 4, 0x09, 42, 0x21, 50]))[0] != 0                     # query_ltime missing (push_pull.rs:311-312)
        ev = bytes([0x1A, 18, 0x09, 42, 0x12, 4, 0x09, 7, 0x11, 20, 0x2A, 2, 0x09, 3, 0x19, 7, 0x21, 50, 0x31, 100])   # one `events` entry: counted, skipped
        assert dec(ev) == (0, (42, [(7, 20)], [7], 50, 100, 1))
        assert dec(W.o_encode_intent(W.JOIN, 5, 9))[0] != 0                          # not a push-pull message


def test_encode_reports_the_needed_size(P):
    m = W.Intent(W.LEAVE, 1, 2**40, 77)
    n = C.c_size_t()
    out = (C.c_uint8 * 4)()
    assert P.serfsim_wire_encode_intent(C.byref(m), out, 4, C.byref(n)) != 0 and n.value == P.serfsim_wire_encoded_len_intent(C.byref(m))


@pytest.mark.parametrize("scen", ["leave_fail", "prune"])
def test_local_state_batch_equals_per_node_oracle_encoding(P, scen):
    """The device batch (length kernel → scan → emit kernel; here the host build) encodes the push-pull message of EVERY node;
    each must equal the oracle's encoding of that node's member table, taken from the ORACLE's run of the same scenario."""
    if scen == "leave_fail":
        sc = scenarios.dissemination_storm(3000, 12, 3, slots=2, seed=3, with_fail=True)
    else:
        sc = scenarios.fuzz_prune(3, n=700, slots=4)
    o = sc.build(oracle_sim, trace=0)
    g = sc.build(emu_sim, trace=0)
    for ticks in (0, 9, 60):
        o.step(ticks); g.step(ticks)
        buf, off = W.local_state_batch(P, g)
        view = dict(status=[o.member_status(s) for s in range(sc.slots)], ltime=[o.status_ltime(s) for s in range(sc.slots)], clock=o.lamport_time())
        for v in list(range(0, sc.n, 37)) + [int(x) for x in sc.subjects]:
            ltime, status, left = W.expected_local_state(view, v, sc.subjects)
            want = W.o_encode_push_pull(ltime, status, left, 1, 1)
            assert bytes(buf[int(off[v]):int(off[v + 1])]) == want, (scen, ticks, v)
        # and back: the decode kernel returns what went in
        lt, ids, sts, ns = W.decode_batch(P, g, buf, off, sc.slots)
        assert (lt == o.lamport_time()).all()
        for v in range(0, sc.n, 53):
            _, status, _ = W.expected_local_state(view, v, sc.subjects)
            assert [(int(ids[v, i]), int(sts[v, i])) for i in range(ns[v])] == status
