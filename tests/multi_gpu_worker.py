"""One rank of the multi-GPU parity test (tests/test_gpu_multi.py), launched under torch.distributed.run:
every rank runs its contiguous share of the global iteration range on its own GPU through the C ABI with an NCCL
communicator attached to the engine (trk3_mc_comm_init), so the library itself ends the run with the single all-reduce of
the tally buffer.  Rank 0 and rank world-1 write what they got back (every rank must hold the reduced tallies)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    out_dir, config, n_total = sys.argv[1], sys.argv[2], int(sys.argv[3])
    import torch
    import torch.distributed as dist
    import trekis3_b200 as tk
    from conftest import table_options
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")            # only carries the 128-byte NCCL id; the data path is the library's own NCCL call
    case = tk.Case.load(tk.make_run_dir(os.path.join(out_dir, f"run_{rank}"), config))
    case.build_tables(cache_dir=os.path.join(ROOT, ".table_cache"), **table_options())
    eng = tk.Engine(case, device=local, batch=64)
    eng.comm_init_torch()
    assert eng.comm_size() == world
    lo, hi = rank * n_total // world, (rank + 1) * n_total // world
    tallies, stats = eng.run(lo, hi)           # collective: ends with ncclAllReduce inside trk3_mc_run
    assert not stats["errors"], stats["errors"]
    if rank in (0, world - 1):
        np.save(os.path.join(out_dir, f"tallies_rank{rank}.npy"), tallies)
    ev = torch.tensor([stats["total_events"]], dtype=torch.int64)
    dist.all_reduce(ev)
    if rank == 0:
        np.save(os.path.join(out_dir, "events_total.npy"), ev.numpy())
    eng.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
