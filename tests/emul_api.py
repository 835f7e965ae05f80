"""ctypes binding of tests/emul/_build/libtrk3_emul.so (CPU emulation of the wavefront engine; tests only)."""
import ctypes as C
import os
import subprocess

import numpy as np

from trekis3_b200 import _abi

_lib = None
_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "emul")


def lib():
    global _lib
    if _lib is None:
        subprocess.run(["make", "-C", _DIR], check=True, capture_output=True)
        l = C.CDLL(os.path.join(_DIR, "_build", "libtrk3_emul.so"))
        PD = C.POINTER(C.c_double)
        l.trk3_emul_run.argtypes = [C.POINTER(_abi.Config), C.POINTER(_abi.Tables), C.c_int64, C.c_int64, C.c_int, PD,
                                    C.POINTER(_abi.Stats), PD, PD]
        _lib = l
    return _lib


def run(case, it_begin, it_end, batch=16, seed=None):
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
    rc = lib().trk3_emul_run(C.byref(cfg), C.byref(case.tables), it_begin, it_end, batch, tallies.ctypes.data_as(PD),
                             C.byref(st), totE.ctypes.data_as(PD), totN.ctypes.data_as(PD))
    if rc != 0:
        raise RuntimeError(f"emulation failed: {rc}")
    return tallies, st.as_dict(), totE, totN
