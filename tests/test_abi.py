"""The three shared libraries load and export every symbol that include/*.h declares (no compute calls)."""
import ctypes as C
import os
import re
import subprocess

import pytest

import trekis3_b200 as tk
from trekis3_b200 import _abi

ROOT = _abi.REPO


def declared_functions(header):
    txt = open(os.path.join(ROOT, "include", header)).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(trk3h?_[a-z0-9_]+)\s*\(", txt)))


def test_gpu_library_exports_every_declared_symbol():
    lib = C.CDLL(_abi.lib_path("gpu"))          # loading needs libcudart only, not a GPU
    names = declared_functions("trekis3_gpu.h")
    assert "trk3_mc_run" in names and "trk3_mc_create" in names
    for n in names:
        assert hasattr(lib, n), f"libtrekis3_gpu.so does not export {n}"
    lib.trk3_gpu_version.restype = C.c_char_p
    assert b"sm_100a" in lib.trk3_gpu_version()


def test_host_library_exports_every_declared_symbol():
    lib = C.CDLL(_abi.lib_path("host"))
    for n in declared_functions("trekis3_host.h"):
        assert hasattr(lib, n), f"libtrekis3_host.so does not export {n}"


def test_gpu_library_is_sm100a_cuda_code():
    out = subprocess.run(["cuobjdump", "-lelf", _abi.lib_path("gpu")], capture_output=True, text=True).stdout
    assert "sm_100a" in out, out


def test_struct_sizes_match_the_c_headers(tmp_path):
    src = tmp_path / "sz.c"
    src.write_text('#include "trekis3_gpu.h"\n'
                   'unsigned long sz_config(void){return sizeof(trk3_config);}\n'
                   'unsigned long sz_tables(void){return sizeof(trk3_tables);}\n'
                   'unsigned long sz_layout(void){return sizeof(trk3_tally_layout);}\n'
                   'unsigned long sz_stats(void){return sizeof(trk3_stats);}\n'
                   'int n_tallies(void){return TRK3_N_TALLIES;} int n_ev(void){return TRK3_N_EVENT_CLASSES;} int n_err(void){return TRK3_N_ERRORS;}\n')
    so = tmp_path / "sz.so"
    subprocess.run(["gcc", "-shared", "-fPIC", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(so)], check=True)
    l = C.CDLL(str(so))
    for f in ("sz_config", "sz_tables", "sz_layout", "sz_stats"):
        getattr(l, f).restype = C.c_ulong
    assert l.sz_config() == C.sizeof(_abi.Config)
    assert l.sz_tables() == C.sizeof(_abi.Tables)
    assert l.sz_layout() == C.sizeof(_abi.TallyLayout)
    assert l.sz_stats() == C.sizeof(_abi.Stats)
    assert l.n_tallies() == len(_abi.TALLY_NAMES)
    assert l.n_ev() == len(_abi.EVENT_NAMES) and l.n_err() == len(_abi.ERROR_NAMES)


def test_engine_fails_loudly_without_gpu(case_c1):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(RuntimeError, match="no CPU fallback|CUDA"):
        tk.do_Monte_Carlo(case_c1, NMC=1)


def test_product_code_never_touches_the_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py's baseline legs may use oracle/ (or the emulation)."""
    pkg = os.path.join(ROOT, "trekis-3_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cpp", ".cu", ".cuh", ".h", ".hpp", ".c")) or f == "Makefile":
                txt = open(os.path.join(dp, f), errors="replace").read()
                assert "trk3_oracle_run" not in txt and "trk3_emul_run" not in txt, os.path.join(dp, f)
                assert "oracle_api" not in txt and "emul_api" not in txt, os.path.join(dp, f)


def test_nccl_entry_points_fail_cleanly_without_an_engine():
    """The collective half of the ABI (trk3_mc_comm_init / trk3_mc_set_comm / trk3_mc_reset) validates its arguments before it
    touches NCCL or CUDA: callable in the GPU-less container."""
    lib = C.CDLL(_abi.lib_path("gpu"))
    lib.trk3_mc_comm_init.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    lib.trk3_mc_set_comm.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    lib.trk3_mc_reset.argtypes = [C.c_void_p]
    lib.trk3_mc_comm_size.argtypes = [C.c_void_p]
    buf = C.create_string_buffer(128)
    assert lib.trk3_mc_comm_init(None, 2, 0, buf) == -1
    assert lib.trk3_mc_set_comm(None, None, 1) == -1
    assert lib.trk3_mc_reset(None) == -1
    assert lib.trk3_mc_comm_size(None) == 1
    assert lib.trk3_nccl_unique_id(None) == -1
