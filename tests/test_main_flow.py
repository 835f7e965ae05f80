"""Program flow of Universal_MC_for_SHI_MAIN.f90 around the engine (trekis-3_b200/main.py): run directory in, table cache and
output tree out.  The table half runs without a GPU; the whole run is a GPU test."""
import os
import subprocess
import sys

import numpy as np
import pytest

import trekis3_b200 as tk
from trekis3_b200 import main as flow

ROOT = tk._abi.REPO


def test_tables_are_built_once_and_then_taken_from_the_reference_cache(tmp_path):
    d = tk.make_run_dir(str(tmp_path / "run"), "C3")
    first = flow.run(d, tables_only=True, evaluator=None, shi_window_only=True)
    assert first["tables"] == "built:host-direct"
    # window-only tables (test shortcut) must never land where the reference looks for its cache
    assert not os.path.exists(os.path.join(d, "OUTPUT_Diamond"))
    out = os.path.join(d, "TEST_ONLY_window_tables", "OUTPUT_Diamond")
    assert os.path.isfile(os.path.join(out, "OUTPUT_Electron_IMFPs_Free_CDF_DOS_0.00_K.dat"))
    assert os.path.isdir(os.path.join(out, "diff_CS")) and os.path.isdir(os.path.join(out, "OUTPUT_Xe_in_Diamond"))
    second = flow.run(d, tables_only=True, evaluator=None, shi_window_only=True)
    assert second["tables"] == "cache" and second["t_tables_s"] < first["t_tables_s"]
    # a stale cache (grid mismatch) is rebuilt, as the reference decides (Analytical_IMFPs.f90:317-323)
    p = os.path.join(out, "OUTPUT_Hole_IMFPs_CDF_CDF_0.00_K.dat")
    lines = open(p).read().splitlines(True)
    open(p, "w").writelines(lines[:-2])
    third = flow.run(d, tables_only=True, evaluator=None, shi_window_only=True)
    assert third["tables"] == "built:host-direct" and len(open(p).read().splitlines()) == len(lines)
    # and --redo-tables ignores a valid cache
    assert flow.run(d, tables_only=True, evaluator=None, redo_tables=True, shi_window_only=True)["tables"] == "built:host-direct"
    assert flow.run(d, tables_only=True, evaluator=None, shi_window_only=True)["tables"] == "cache"
    # the reference's own conditions (Analytical_IMFPs.f90:307-327): a redo_* keyword in INPUT_PARAMETERS.txt ...
    ip = os.path.join(d, "INPUT_PARAMETERS.txt")
    text = open(ip).read()
    open(ip, "w").write(text + "redo_IMFP\n")
    assert flow.run(d, tables_only=True, evaluator=None, shi_window_only=True)["tables"] == "built:host-direct"
    open(ip, "w").write(text)
    assert flow.run(d, tables_only=True, evaluator=None, shi_window_only=True)["tables"] == "cache"
    # ... or a .cdf file that is newer than the cached tables
    import shutil
    cdf_dir = os.path.join(d, "INPUT_CDF")
    real = os.path.realpath(cdf_dir)
    os.unlink(cdf_dir); shutil.copytree(real, cdf_dir)
    os.utime(os.path.join(cdf_dir, "Diamond.cdf"), None)            # touched: now newer than the tables
    assert flow.run(d, tables_only=True, evaluator=None, shi_window_only=True)["tables"] == "built:host-direct"
    assert flow.run(d, tables_only=True, evaluator=None, shi_window_only=True)["tables"] == "cache"


def test_command_line_tables_only(tmp_path):
    d = tk.make_run_dir(str(tmp_path / "run"), "C3")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "trekis3_run.py"), d, "--tables-only", "--evaluator", "host", "--quiet", "--shi-window-only"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    assert r.stdout.startswith("tables: built:host-direct")
    assert os.path.isdir(os.path.join(d, "TEST_ONLY_window_tables", "OUTPUT_Diamond", "diff_CS"))


def test_two_ranks_share_one_table_cache(tmp_path):
    """Under torchrun rank 0 builds the tables and writes the cache, the other ranks read it (the reference broadcasts)."""
    d = tk.make_run_dir(str(tmp_path / "run"), "C3")
    prog = ("import sys; sys.path.insert(0, %r); import trekis3_b200; from trekis3_b200 import main as flow; import os; "
            "i = flow.run(%r, tables_only=True, evaluator=None, shi_window_only=True); "
            "print('RANK', os.environ['RANK'], i['tables'], flush=True)") % (ROOT, d)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29541", "--no-python", sys.executable, "-c", prog], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    out = r.stdout
    assert "RANK 0 built:host-direct" in out and "RANK 1 cache" in out


def test_monte_carlo_needs_the_gpu(tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    d = tk.make_run_dir(str(tmp_path / "run"), "C1", nmc=2)
    flow.run(d, tables_only=True, evaluator=None, shi_window_only=True)
    with pytest.raises(RuntimeError):
        flow.run(d)


@pytest.mark.gpu
def test_whole_run_writes_the_reference_output_tree(tmp_path):
    d = tk.make_run_dir(str(tmp_path / "run"), "C1", nmc=20)
    info = flow.run(d, verbose=False)            # full tables (the ion over its whole grid), q-integrals on the GPU
    assert info["tables"] == "built:gpu" and tk.gpu_library_loaded()
    assert not info["stats"]["errors"] and info["stats"]["max_energy_drift"] < 1e-9
    od = info["out_dir"]
    assert od.startswith(os.path.join(d, "OUTPUT_Al2O3", "OUTPUT_Xe_in_Al2O3")) and os.path.isfile(os.path.join(od, "Total_numbers.txt"))
    tot = np.loadtxt(os.path.join(od, "Total_numbers.txt"), skiprows=1)
    assert tot.shape[0] == 5 and 15e3 < tot[-1, 3] < 40e3                  # deposited energy per ion ~ S_e x layer
    assert np.allclose(tot[1:, 3], tot[-1, 3], rtol=1e-9)                  # conserved once the ion has left
    # second run: tables from the cache it wrote, same result (streams are keyed by the global iteration index)
    info2 = flow.run(d, verbose=False)
    assert info2["tables"] == "cache" and info2["out_dir"] != od            # the reference numbers repeated output directories
    assert np.allclose(info2["tallies"], info["tallies"], rtol=1e-6, atol=1e-12)


def test_bench_reference_arm_prints_one_json_line(tmp_path):
    """`bench.py --impl reference`: the CPU restatement timed on the host cores; exactly one JSON line on stdout with the keys
    the driver reads (the GPU arm shares the printing path and cannot run here)."""
    import json
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--config", "C1", "--nmc", "10"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = r.stdout.splitlines()
    assert len(lines) == 1
    b = json.loads(lines[0])
    assert b["impl"] == "reference" and b["unit"] == "iterations/s" and b["higher_is_better"] is True and b["value"] > 0
    assert b["cpu_baseline"]["kind"] == "port" and b["cpu_baseline"]["cores"] >= 1
    assert b["e2e"] == {"value": b["value"], "unit": "iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert b["config"]["workload"].startswith("C1: Xe 167 MeV in Al2O3")
