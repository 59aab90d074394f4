"""Byzantine stale-record injectors (BASELINE configs[4]) — CPU side.

No reference semantics exist for this configuration: serf ignores stale intents silently.  The model is
defined by the oracle (oracle/serf_oracle.cpp, "byzantine" block); these tests pin its properties and check that the
host/device rules the CUDA kernel runs (serf_b200/csrc/byz.cuh, compiled for the host) agree with the oracle's on
random records.  GPU parity: tests/test_gpu_z_byzantine.py.
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle_lib import lib as oracle_lib_handle, oracle_sim
from serf_b200 import scenarios
from serf_b200.sim import RECORD_DTYPE, SerfsimError

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def byzcheck(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("byzcheck") / "byzcheck.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wall", "-Wno-unknown-pragmas", "-o", so,
                           os.path.join(ROOT, "tests", "cpp", "byz_rules_check.cpp")])
    L = C.CDLL(so)
    L.byzcheck_entries.restype = None
    L.byzcheck_entries.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p]
    L.byzcheck_anomalous.restype = C.c_int
    L.byzcheck_anomalous.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32]
    return L


def _random_records(rng, n):
    r = np.zeros(n, dtype=RECORD_DTYPE)
    r["status_ltime"] = rng.integers(0, 12, n)
    r["incarnation"] = rng.integers(0, 9, n)
    r["status"] = rng.integers(0, 5, n)
    r["ml"] = rng.integers(0, 4, n) | (rng.integers(0, 16, n) << 2)
    r["flags"] = rng.integers(0, 2, n)
    r["qjoin_lt"] = rng.integers(0, 12, n)
    r["qleave_lt"] = rng.integers(0, 12, n)
    r["conf_mask"] = rng.integers(0, 1 << 16, n)
    return r


def test_device_rules_equal_oracle_rules(byzcheck):
    O = oracle_lib_handle()
    O.oracle_byz_stale.restype = None
    O.oracle_byz_stale.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p]
    O.oracle_byz_judge.restype = C.c_int
    O.oracle_byz_judge.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32]
    rng = np.random.default_rng(11)
    recs = _random_records(rng, 4000)
    for delta in (0, 1, 2, 5):
        for i in range(len(recs)):
            p = recs[i:i + 1].ctypes.data
            d = np.zeros(5, dtype=np.uint32)
            o = np.zeros(3, dtype=np.uint32)
            byzcheck.byzcheck_entries(p, delta, d.ctypes.data)
            O.oracle_byz_stale(p, delta, o.ctypes.data)
            assert d[0] == (recs["flags"][i] & 1)
            assert (d[1], d[2], d[3]) == (o[0], o[1], o[2]), (recs[i], delta)
            assert d[4] == o[2] >> 6
            # the receiver's verdict: the kernel ORs what the oracle decides per arriving entry
            q = recs[(i * 7 + 3) % len(recs):][:1]
            want = bool(O.oracle_byz_judge(q.ctypes.data, int(o[0]), int(o[1]), delta)) or bool(O.oracle_byz_judge(q.ctypes.data, 2, int(o[2]), delta))
            assert bool(byzcheck.byzcheck_anomalous(q.ctypes.data, int(d[2]), int(d[4]), delta)) == want


def _run(sc, **cfg):
    o = sc.build(oracle_sim, trace=1, **cfg)
    t, ok = o.run_until_converged(sc.max_ticks)
    assert ok
    return o, t


def test_injectors_get_flagged_and_honest_nodes_do_not():
    sc = scenarios.byzantine_injectors(6000, 16, 4, 0.02, seed=1)
    o, _ = _run(sc)
    flags = o.anomaly_flags()
    honest = np.ones(sc.n, dtype=bool)
    honest[sc.byzantine] = False
    assert flags[honest].sum() == 0                                   # only injected entries are judged
    assert flags[sc.byzantine].mean() > 0.95
    st = o.byzantine_stats()
    assert st["flagged"] == flags.sum() and st["messages"] == 2 * st["edge_updates"] > 0


def test_stale_entries_never_win_against_the_truth():
    """With injectors on, every up node still ends with the same views as the honest run: a stale copy can only be
    accepted by a node whose view is even older, and the real message overtakes it."""
    sc = scenarios.byzantine_injectors(5000, 16, 4, 0.2, seed=3)
    honest = scenarios.byzantine_injectors(5000, 16, 4, 0.2, seed=3)
    honest.byzantine = None
    a, _ = _run(sc)
    b, _ = _run(honest)
    for s in range(sc.slots):
        same = (a.member_status(s) == b.member_status(s)) & (a.status_ltime(s) == b.status_ltime(s))
        assert same.mean() > 0.999                                    # a stranded node may keep a stale-but-newer-than-bootstrap view
    assert a.stats()["changed"] >= b.stats()["changed"]


def test_delta_zero_flags_on_equal_views_and_large_delta_never():
    sc0 = scenarios.byzantine_injectors(3000, 12, 3, 0.05, delta=0, seed=2, churn=False, slots=1)
    o0, _ = _run(sc0)
    assert o0.anomaly_flags()[sc0.byzantine].all()                    # delta 0: an equal view already counts
    big = scenarios.byzantine_injectors(3000, 12, 3, 0.05, delta=1000, seed=2, churn=False, slots=1)
    ob, _ = _run(big)
    assert ob.anomaly_flags().sum() == 0 and ob.byzantine_stats()["messages"] > 0


def test_crashed_injector_stops_and_crashed_receiver_does_not_judge():
    from serf_b200.sim import Op
    sc = scenarios.byzantine_injectors(3000, 12, 3, 0.02, seed=5, churn=False, slots=1)
    victim = int(sc.byzantine[0])
    sc.ops.append((0, Op.FAIL, victim, 0))                            # down before it ever sends
    o, _ = _run(sc)
    assert o.anomaly_flags()[victim] == 0
    assert o.anomaly_flags()[sc.byzantine[1:]].mean() > 0.9


def test_threads_do_not_change_byzantine_results():
    L = oracle_lib_handle()
    sc = scenarios.byzantine_injectors(4000, 12, 4, 0.05, seed=4)
    res = []
    for th in (1, 3):
        o = sc.build(oracle_sim, trace=1)
        L.oracle_sim_set_threads(o._h, th)
        o.reset(sc.cfg["seed"])
        sc.schedule(o)
        t = o.run_until_converged(sc.max_ticks)
        res.append((t, o.state_hash(), o.byzantine_stats(), o.anomaly_flags().tobytes(), o.tick_trace().tobytes()))
    assert res[0] == res[1]


def test_set_byzantine_validation():
    o = oracle_sim(100, 1)
    with pytest.raises(SerfsimError):
        o.set_byzantine([5, 5])
    with pytest.raises(SerfsimError):
        o.set_byzantine([100])
