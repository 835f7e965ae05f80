"""Multi-GPU parity on hardware: the tallies reduced over N GPUs by the library's own NCCL all-reduce equal the one-GPU run
of the same GLOBAL iteration range (reference semantics: iteration split Monte_Carlo.f90:111-129 + 26 MPI_Reduce :131-389).
Needs >= 2 GPUs (`gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu`); skipped on a one-GPU box."""
import os
import subprocess
import sys

import numpy as np
import pytest

import trekis3_b200 as tk
from conftest import table_options

pytestmark = pytest.mark.gpu
ROOT = tk._abi.REPO


def _n_gpus():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.parametrize("world", [2, 4, 8])
def test_reduced_tallies_equal_the_one_gpu_run(tmp_path, world):
    if _n_gpus() < world:
        pytest.skip(f"needs {world} GPUs")
    n_total = 48
    out = str(tmp_path)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
                        "--master-port", str(29560 + world), os.path.join(ROOT, "tests", "multi_gpu_worker.py"), out, "C1", str(n_total)],
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-3000:]
    case = tk.Case.load(tk.make_run_dir(os.path.join(out, "run_single"), "C1"))
    case.build_tables(cache_dir=os.path.join(ROOT, ".table_cache"), **table_options())
    eng = tk.Engine(case, batch=64)
    whole, sw = eng.run(0, n_total)
    eng.close()
    lay = case.layout()
    i = tk.TALLY_NAMES.index("Out_diff_coeff")                # per-process recurrence over iterations: depends on the split, as in the reference
    m = np.ones(lay.total, bool); m[lay.off[i]: lay.off[i] + lay.len[i]] = False
    first, last = np.load(os.path.join(out, "tallies_rank0.npy")), np.load(os.path.join(out, f"tallies_rank{world - 1}.npy"))
    assert np.array_equal(first, last)                         # every rank holds the same reduced buffer
    den = np.maximum(np.abs(first[m]), np.abs(whole[m]))
    assert np.all(np.abs(first[m] - whole[m]) <= 1e-11 * den + 1e-300)
    assert int(np.load(os.path.join(out, "events_total.npy"))[0]) == sw["total_events"]
