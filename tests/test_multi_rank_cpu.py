"""N>1 host logic on CPU (gloo, world_size 2): iterations are sharded by global index, tallies are summed with ONE
all-reduce of the packed buffer (the role of MPI_subroutines.f90 / the 26 MPI_Reduce calls, Monte_Carlo.f90:131-389).
The per-rank engine is the CPU emulation of the wavefront engine (no GPU here)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, run_dir, n_total, out_dir):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import trekis3_b200 as tk
    import emul_api
    case = tk.Case.load(run_dir)
    case.build_tables(shi_window_only=True, cache_dir=os.path.join(ROOT, ".table_cache"))
    # contiguous split of the global iteration range
    per = (n_total + world - 1) // world
    lo, hi = rank * per, min(n_total, (rank + 1) * per)
    t, st, _, _ = emul_api.run(case, lo, hi, batch=4)
    lay = case.layout()
    i_dc = tk.TALLY_NAMES.index("Out_diff_coeff")
    buf = torch.from_numpy(t)
    dist.all_reduce(buf)                       # the single collective of the path
    ev = torch.tensor([st["total_events"]], dtype=torch.int64); dist.all_reduce(ev)
    if rank == 0:
        np.save(os.path.join(out_dir, "reduced.npy"), buf.numpy())
        np.save(os.path.join(out_dir, "events.npy"), ev.numpy())
        np.save(os.path.join(out_dir, "dc_slice.npy"), np.array([lay.off[i_dc], lay.len[i_dc]]))
    dist.destroy_process_group()


def test_two_ranks_equal_one_rank(tmp_path):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import trekis3_b200 as tk
    import emul_api
    run_dir = tk.make_run_dir(str(tmp_path / "run"), "C1")
    case = tk.Case.load(run_dir)
    case.build_tables(shi_window_only=True, cache_dir=os.path.join(ROOT, ".table_cache"))
    n_total = 6
    single, st, _, _ = emul_api.run(case, 0, n_total, batch=4)
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, run_dir, n_total, str(tmp_path)), nprocs=2, join=True)
    red = np.load(tmp_path / "reduced.npy")
    ev = np.load(tmp_path / "events.npy")
    o, l = np.load(tmp_path / "dc_slice.npy")
    assert int(ev[0]) == st["total_events"]
    mask = np.ones(red.size, bool); mask[o:o + l] = False
    # every tally is a plain sum over iterations -> independent of the rank count ...
    assert np.allclose(red[mask], single[mask], rtol=1e-11, atol=1e-300)
    # ... except Out_diff_coeff, whose reference definition is a per-process recurrence (Monte_Carlo.f90:1098)
    assert np.all(red[o:o + l] > 0)
