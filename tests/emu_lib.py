"""Builds and loads tests/emu/_build/libserfsim_emu.so: the product's OWN kernel and host sources (serf_b200/csrc/*.cu)
compiled by g++ against the CUDA-on-CPU shim in tests/emu/ — test infrastructure only.

What it is for: running the logic of the CUDA kernels (and the C-ABI host code around them) against the oracle on a
machine without a GPU.  What it is not: a CPU backend (nothing under serf_b200/ knows about it; the product library is
the nvcc build and refuses to run without a GPU), nor evidence about performance or memory-model behaviour.
"""
import ctypes as C
import os
import subprocess

from serf_b200 import sim as _sim

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "serf_b200", "csrc")
EMU = os.path.join(ROOT, "tests", "emu")
OUT = os.path.join(EMU, "_build", "libserfsim_emu.so")
_LIB = None


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def build():
    if os.environ.get("SERFSIM_EMU_LIB"):          # e.g. a --coverage build of the same sources (tools/emu_coverage.sh)
        return os.environ["SERFSIM_EMU_LIB"]
    deps = sources() + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh")]
    deps += [os.path.join(EMU, "cuda_runtime.h"), os.path.join(EMU, "emu_engine.cpp"), os.path.join(ROOT, "include", "serfsim.h")]
    if os.path.exists(OUT) and os.path.getmtime(OUT) >= max(os.path.getmtime(d) for d in deps):
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    cmd = ["g++", "-std=c++17", "-O1", "-g", "-fPIC", "-shared", "-DSERFSIM_EMU", "-I" + EMU, "-Wall", "-Wno-unknown-pragmas",
           "-Wno-unused-function", "-o", OUT]
    for s in sources():
        cmd += ["-x", "c++", s]
    cmd += ["-x", "c++", os.path.join(EMU, "emu_engine.cpp")]
    subprocess.check_call(cmd)
    return OUT


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        for name, (res, args) in _sim.SIGNATURES.items():
            f = getattr(L, "serfsim_" + name)
            f.restype, f.argtypes = res, args
        for name, (res, args) in _sim.PRODUCT_ONLY.items():
            f = getattr(L, name)
            f.restype, f.argtypes = res, args
        _LIB = L
    return _LIB


def emu_sim(n_nodes, slots=1, **cfg_kw):
    """GossipSim-shaped driver over the host-compiled kernels."""
    return _sim.GossipSim(n_nodes, slots, _lib=lib(), _prefix="serfsim_", _errfn="serfsim_last_error", **cfg_kw)
