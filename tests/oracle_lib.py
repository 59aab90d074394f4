"""ctypes loader for the CPU oracle (oracle/liboracle.so) — test infrastructure only.

The oracle restates the reference's algorithm (see the header of oracle/serf_oracle.cpp);
nothing under serf_b200/ imports this module.
"""
import ctypes as C
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
_LIB = None

u8p, u32p, u64p = C.POINTER(C.c_uint8), C.POINTER(C.c_uint32), C.POINTER(C.c_uint64)


def build():
    so = os.path.join(ORACLE_DIR, "liboracle.so")
    srcs = [os.path.join(ORACLE_DIR, "serf_oracle.cpp"), os.path.join(ORACLE_DIR, "wire_oracle.cpp"), os.path.join(ROOT, "include", "serfsim.h")]
    if (not os.path.exists(so)) or os.path.getmtime(so) < max(os.path.getmtime(f) for f in srcs):
        subprocess.check_call(["make", "-s", "-C", ORACLE_DIR])
    return so


def lib():
    global _LIB
    if _LIB is not None:
        return _LIB
    L = C.CDLL(build())
    vp, i32, i64, u8, u32, u64 = C.c_void_p, C.c_int, C.c_int64, C.c_uint8, C.c_uint32, C.c_uint64
    sigs = {
        "ref_node_new": (vp, [u64, u32]), "ref_node_free": (None, [vp]),
        "ref_clock_time": (u64, [vp, i32]), "ref_clock_increment": (u64, [vp]), "ref_clock_witness": (None, [vp, u64]),
        "lamport_new_time": (u64, []),
        "ref_set_serf_state": (None, [vp, i32]), "ref_get_serf_state": (i32, [vp]),
        "ref_insert_member": (None, [vp, u64, i32, u64]), "ref_member_get": (i32, [vp, u64, u8p, u64p]),
        "ref_num_members": (u64, [vp]),
        "ref_handle_node_join_intent": (i32, [vp, u64, u64]), "ref_handle_node_leave_intent": (i32, [vp, u64, u64, i32]),
        "ref_run_detached": (None, [vp]), "ref_refutes": (u32, [vp]),
        "ref_handle_node_join": (None, [vp, u64]), "ref_handle_node_leave": (None, [vp, u64, i64]),
        "ref_upsert_intent": (i32, [vp, u64, i32, u64, i64]), "ref_recent_intent": (i32, [vp, u64, i32, u64p]),
        "ref_reap_intents": (None, [vp, i64, i64]),
        "ref_merge_remote_state": (None, [vp, u64, u64p, u64p, u32, u64p, u32, u64, u64]),
        "ref_local_state": (u32, [vp, u64p, u64p, u64p, u32, u64p, u32, u32p, u64p, u64p]),
        "ref_handle_user_event": (i32, [vp, u64, C.c_char_p, C.c_char_p]), "ref_event_clock_witness": (None, [vp, u64]),
        "ref_user_event_count": (u32, [vp]), "ref_user_event_get": (C.c_char_p, [vp, u32, i32]), "ref_event_buffer_has": (i32, [vp, u64]),
        "ref_api_join": (None, [vp]), "ref_api_leave": (i32, [vp]), "ref_api_force_leave": (None, [vp, u64, i32]),
        "ref_queue_len": (u32, [vp]), "ref_queue_get": (i32, [vp, u32, u8p, u64p, u64p, u32p]),
        "ref_get_broadcasts": (u32, [vp, u32, u32, u8p, u64p, u64p, u32]),
        "ref_remove_old_member": (None, [vp, i32, u64]),
        "ref_left_count": (u32, [vp]), "ref_failed_count": (u32, [vp]),
        "ref_push_left": (None, [vp, u64, i32, u64, i64]), "ref_push_failed": (None, [vp, u64, i32, u64, i64]),
        "ref_reap": (None, [vp, i64, i64, i64, i64]),
        "ref_event_count": (u32, [vp]), "ref_event_get": (i32, [vp, u32, u32p, u64p]),
        "ref_get_queue_max": (u64, [u64, u64, u64]),
        "oracle_philox4x32_10": (None, [u32p, u32p, u32p]),
        "oracle_retransmit_limit": (u32, [u32, u64]),
        "oracle_suspicion_table": (u32, [u32, u32, u32, u32, u64, u32p, u32]),
        "oracle_mix64": (u64, [u64]), "oracle_from_hash": (u32, [u32]),
        "oracle_view_init": (None, [vp, i32, i32, u32]),
        "oracle_view_join_intent": (i32, [vp, u32, u32]),
        "oracle_view_leave_intent": (i32, [vp, u32, i32, i32, C.POINTER(i32), u32]),
        "oracle_view_node_join": (None, [vp]), "oracle_view_node_leave": (None, [vp, u32]),
        "oracle_last_error": (C.c_char_p, []),
    }
    for name, (res, args) in sigs.items():
        f = getattr(L, name)
        f.restype, f.argtypes = res, args
    _LIB = L
    return L


class RefNode:
    """One serf node of the oracle (Part A): the object the reference KATs are replayed on."""
    LEAVE, JOIN = 1, 2          # MessageType tags, types/message.rs:17-18
    NONE, ALIVE, LEAVING, LEFT, FAILED = 0, 1, 2, 3, 4

    def __init__(self, self_id=0, retransmit_mult=4):
        self.L = lib()
        self.p = self.L.ref_node_new(self_id, retransmit_mult)
        self.ids = {}

    def __del__(self):
        if getattr(self, "p", None):
            self.L.ref_node_free(self.p)
            self.p = None

    def id(self, name):
        """Map the reference tests' string ids ("test", "foo") to dense integers."""
        if isinstance(name, int):
            return name
        return self.ids.setdefault(name, 1000 + len(self.ids))

    def clock(self, which=0): return self.L.ref_clock_time(self.p, which)
    def increment(self): return self.L.ref_clock_increment(self.p)
    def witness(self, t): self.L.ref_clock_witness(self.p, t)
    def insert_member(self, name, status, status_time): self.L.ref_insert_member(self.p, self.id(name), status, status_time)

    def member(self, name):
        st, lt = C.c_uint8(), C.c_uint64()
        if not self.L.ref_member_get(self.p, self.id(name), C.byref(st), C.byref(lt)):
            return None
        return st.value, lt.value

    def join_intent(self, ltime, name): return bool(self.L.ref_handle_node_join_intent(self.p, ltime, self.id(name)))
    def leave_intent(self, ltime, name, prune=False): return bool(self.L.ref_handle_node_leave_intent(self.p, ltime, self.id(name), int(prune)))
    def node_join(self, name): self.L.ref_handle_node_join(self.p, self.id(name))
    def node_leave(self, name, now_ms=0): self.L.ref_handle_node_leave(self.p, self.id(name), now_ms)
    def upsert_intent(self, name, ty, ltime, wall_ms=0): return bool(self.L.ref_upsert_intent(self.p, self.id(name), ty, ltime, wall_ms))

    def recent_intent(self, name, ty):
        lt = C.c_uint64()
        return lt.value if self.L.ref_recent_intent(self.p, self.id(name), ty, C.byref(lt)) else None

    def queue(self):
        out = []
        for i in range(self.L.ref_queue_len(self.p)):
            ty, lt, idv, tx = C.c_uint8(), C.c_uint64(), C.c_uint64(), C.c_uint32()
            self.L.ref_queue_get(self.p, i, C.byref(ty), C.byref(lt), C.byref(idv), C.byref(tx))
            out.append((ty.value, lt.value, idv.value, tx.value))
        return out

    def events(self):
        out = []
        for i in range(self.L.ref_event_count(self.p)):
            ty, idv = C.c_uint32(), C.c_uint64()
            self.L.ref_event_get(self.p, i, C.byref(ty), C.byref(idv))
            out.append((ty.value, idv.value))
        return out


# ---- Part B: the tick oracle driven through the same Python driver as the product --------
def oracle_sim(n_nodes, slots=1, **cfg_kw):
    """A GossipSim-shaped object backed by the CPU oracle (oracle_sim_* entry points)."""
    from serf_b200 import sim as _sim
    L = lib()
    for name, (res, args) in _sim.SIGNATURES.items():
        f = getattr(L, "oracle_sim_" + name)
        f.restype, f.argtypes = res, args
    return _sim.GossipSim(n_nodes, slots, _lib=L, _prefix="oracle_sim_", _errfn="oracle_last_error", **cfg_kw)


def physical_cpus():
    """One logical CPU per physical core this process may use, alternating between packages (NUMA nodes)."""
    allowed = sorted(os.sched_getaffinity(0))
    by_pkg = {}
    for c in allowed:
        try:
            core = int(open(f"/sys/devices/system/cpu/cpu{c}/topology/core_id").read())
            pkg = int(open(f"/sys/devices/system/cpu/cpu{c}/topology/physical_package_id").read())
        except (OSError, ValueError):
            core, pkg = c, 0
        by_pkg.setdefault(pkg, {}).setdefault(core, c)
    lists = [list(v.values()) for _, v in sorted(by_pkg.items())]
    out = []
    for i in range(max(len(x) for x in lists)):
        out += [x[i] for x in lists if i < len(x)]
    return out


def oracle_sim_threaded(n_nodes, slots=1, **cfg_kw):
    """oracle_sim with one pinned worker per physical core (results do not depend on the thread count): the checker of the
    full-size parity tests, where a single-threaded run would take minutes."""
    o = oracle_sim(n_nodes, slots, **cfg_kw)
    L = lib()
    L.oracle_sim_set_threads.restype, L.oracle_sim_set_threads.argtypes = C.c_int, [C.c_void_p, C.c_int]
    L.oracle_sim_set_affinity.restype, L.oracle_sim_set_affinity.argtypes = C.c_int, [C.c_void_p, C.POINTER(C.c_int), C.c_int]
    cpus = physical_cpus()
    assert L.oracle_sim_set_threads(o._h, len(cpus)) == 0
    assert L.oracle_sim_set_affinity(o._h, (C.c_int * len(cpus))(*cpus), len(cpus)) == 0
    return o
