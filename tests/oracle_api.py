"""ctypes binding of oracle/_build/libtrk3_oracle.so -- test infrastructure only."""
import ctypes as C
import os
import subprocess

import numpy as np

import trekis3_b200 as tk
from trekis3_b200 import _abi

_lib = None
_native = False


def build(native=False):
    subprocess.run(["make", "-C", os.path.join(_abi.REPO, "oracle")] + (["native"] if native else []), check=True, capture_output=True)


def use_native_build():
    """The CPU arm of bench.py: compile the oracle ON THIS MACHINE with -O3 -march=native and use that
    library from now on.  Returns the compiler flags for the report."""
    global _lib, _native
    build(native=True)
    _lib, _native = None, True
    return "g++ -O3 -march=native -ffp-contract=off"


def lib():
    global _lib
    if _lib is None:
        path = _abi.lib_path("oracle")
        if _native:
            path = os.path.join(os.path.dirname(path), "libtrk3_oracle_native.so")
        elif not os.path.exists(path):
            build()
        l = C.CDLL(path)
        PD = C.POINTER(C.c_double)
        l.trk3_oracle_run.argtypes = [C.POINTER(_abi.Config), C.POINTER(_abi.Tables), C.c_int64, C.c_int64, C.c_int,
                                      C.c_int, PD, C.POINTER(_abi.Stats), PD, PD]
        l.trk3_oracle_philox.argtypes = [C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
        l.trk3_oracle_version.restype = C.c_char_p
        _lib = l
    return _lib


def run(case, it_begin, it_end, rng_mode=1, threads=0, seed=None):
    """Returns (tallies_sum, stats_dict, iter_totE[n,Nt], iter_totNel[n,Nt])."""
    lay = case.layout()
    cfg = _abi.Config.from_buffer_copy(case.config)
    if seed is not None:
        cfg.seed = int(seed)
    n = it_end - it_begin
    tallies = np.zeros(lay.total)
    totE = np.zeros((n, lay.Nt))
    totN = np.zeros((n, lay.Nt))
    st = _abi.Stats()
    PD = C.POINTER(C.c_double)
    rc = lib().trk3_oracle_run(C.byref(cfg), C.byref(case.tables), it_begin, it_end, rng_mode, threads,
                               tallies.ctypes.data_as(PD), C.byref(st), totE.ctypes.data_as(PD), totN.ctypes.data_as(PD))
    if rc != 0:
        raise RuntimeError(f"oracle failed: {rc}")
    return tallies, st.as_dict(), totE, totN


def philox(ctr, key):
    c = (C.c_uint32 * 4)(*ctr)
    k = (C.c_uint32 * 2)(*key)
    o = (C.c_uint32 * 4)()
    lib().trk3_oracle_philox(c, k, o)
    return list(o)
