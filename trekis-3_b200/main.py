"""The program flow of Universal_MC_for_SHI_MAIN.f90 around the CUDA engine: one run directory in, the reference's output tree out.

    python scripts/trekis3_run.py RUN_DIR                       one GPU
    torchrun --nproc-per-node N scripts/trekis3_run.py RUN_DIR  N GPUs: iterations sharded, one NCCL all-reduce of the tallies

RUN_DIR holds what the Fortran program reads (INPUT_PARAMETERS.txt, INPUT_CDF/, INPUT_DOS/, INPUT_EADL/) and receives what it
writes (OUTPUT_<material>/...).  Steps, with the reference lines they replace:
    Read_input_file                       MAIN.f90:120-143   Case.load
    Analytical_ion_dEdx / _electron_dEdx  MAIN.f90:146-238   tables: the reference-format cache under RUN_DIR/OUTPUT_<material>/ is
                                                             used when it is valid (written by the Fortran program or by an earlier
                                                             run); otherwise the tables are built -- on the GPU when the CUDA library
                                                             is there -- and the cache is written for the next run, of either program
    do_Monte_Carlo                        MAIN.f90:267-276   Engine.run_device on this rank's share of the global iteration indices
    MPI_Reduce x 26                       Monte_Carlo.f90:131-389   one all_reduce of the packed tally buffer
    Save_output                           MAIN.f90:278-330   Case.save_output (rank 0)
The Monte-Carlo engine has no CPU fallback: without libtrekis3_gpu.so and a GPU, run() raises (tables_only=True stops before it).
"""
import os
import sys
import time

from .host import Case


def _stamp(verbose, text, t0):
    if verbose:
        print("%-52s %9.3f s" % (text, time.perf_counter() - t0), file=sys.stderr, flush=True)


def prepare_tables(case, run_dir, evaluator="auto", threads=0, redo=False, write_cache=True, verbose=False, shi_window_only=False):
    """Tables of a loaded case: from the reference-format cache in run_dir if it is valid, else built (and cached).
    Returns "cache" or "built:<evaluator>".
    Tables built with shi_window_only (a test shortcut: ion rows outside [0.5 E, E] are placeholders) never go into the
    reference's cache directory, where a later run with another ion energy -- or the Fortran program itself -- would accept
    them: they are cached under <run_dir>/TEST_ONLY_window_tables/ instead."""
    cache_root = os.path.join(run_dir, "TEST_ONLY_window_tables") if shi_window_only else run_dir
    if not redo:
        try:
            case.read_reference_cache(cache_root, threads=threads)
            return "cache"
        except RuntimeError as e:          # missing file / grid mismatch: the conditions under which the reference recomputes
            if verbose:
                print(f"[trekis3] table cache not used ({e}); building the tables", file=sys.stderr)
    ev = evaluator
    if ev == "auto":
        ev = None
        try:
            import torch
            from . import _abi
            if torch.cuda.is_available() and os.path.exists(_abi.lib_path("gpu")):
                ev = "gpu"
        except Exception:
            ev = None
    case.build_tables(threads=threads, verbose=verbose, evaluator=ev, shi_window_only=shi_window_only)
    if write_cache:
        os.makedirs(cache_root, exist_ok=True)
        try:
            case.write_reference_cache(cache_root)
        except RuntimeError as e:          # closed-form shells (BEB, delta-function CDF) have no differential tables to cache
            if "no differential tables" not in str(e):
                raise
            if verbose:
                print(f"[trekis3] {e}", file=sys.stderr)
    return "built:" + (ev or "host-direct")


def run(run_dir, nmc=None, evaluator="auto", threads=0, redo_tables=False, tables_only=False, out_root=None, verbose=False, seed=None,
        shi_window_only=False):
    """One TREKIS-3 run.  Returns a dict: tables ("cache" / "built:..."), out_dir (rank 0, after the Monte-Carlo), stats, timings.
    shi_window_only (tests): tabulate the ion only around its own energy instead of over the whole grid."""
    t0 = time.perf_counter()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    case = Case.load(run_dir)
    verbose = verbose and rank == 0
    _stamp(verbose, "Input files read:", t0)
    info = {"rank": rank, "world": world}
    # tables: rank 0 prepares (and writes) the cache, the other ranks read it afterwards -- the reference broadcasts instead
    dist, own_group = None, False
    if world > 1:
        import torch
        import torch.distributed as dist
        if not tables_only:
            torch.cuda.set_device(local_rank)
        own_group = not dist.is_initialized()
        if own_group:
            if tables_only:
                dist.init_process_group("gloo")
            else:
                dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    failure = None
    if rank == 0:
        try:
            info["tables"] = prepare_tables(case, run_dir, evaluator, threads, redo_tables, True, verbose, shi_window_only)
        except Exception as e:             # bad input, write failure: the other ranks must not sit in a collective until it times out
            failure = f"{type(e).__name__}: {e}"
    if dist is not None:
        box = [failure]
        dist.broadcast_object_list(box, src=0)      # doubles as the barrier behind which the cache files exist
        if box[0] is not None:
            if own_group:
                dist.destroy_process_group()
            raise RuntimeError("rank 0 could not prepare the tables: " + box[0])
        if rank != 0:                       # a rank that cannot read the cache builds with the SAME options as rank 0 did
            info["tables"] = prepare_tables(case, run_dir, evaluator, threads, False, False, False, shi_window_only)
    elif failure is not None:
        raise RuntimeError(failure)
    _stamp(verbose, "Mean free paths and differential tables ready:", t0)
    info["t_tables_s"] = time.perf_counter() - t0
    if tables_only:
        if own_group:
            dist.destroy_process_group()
        return info

    import torch
    from .engine import Engine
    n = int(nmc if nmc is not None else case.get("NMC"))
    lay = case.layout()
    eng = Engine(case, device=local_rank, seed=seed)
    tally = torch.zeros(lay.total, dtype=torch.float64, device=f"cuda:{local_rank}")
    eng.set_stream(torch.cuda.current_stream().cuda_stream)
    eng.set_device_tallies(tally.data_ptr())
    # contiguous shares of the GLOBAL iteration indices (Philox streams are keyed by them: the sum does not depend on the split)
    first = (n * rank) // world
    last = (n * (rank + 1)) // world
    _stamp(verbose, "Starting MC iterations:", t0)
    t_mc = time.perf_counter()
    stats = eng.run_device(first, last) if last > first else None
    if dist is not None:
        dist.all_reduce(tally)
    torch.cuda.synchronize()
    info["t_mc_s"] = time.perf_counter() - t_mc
    info["iterations"] = n
    info["stats"] = stats
    _stamp(verbose, "Preparing MC output data:", t0)
    if rank == 0:
        info["out_dir"] = case.save_output(tally.cpu().numpy(), n, out_root or run_dir)
        info["tallies"] = tally.cpu().numpy()
    eng.close()
    if dist is not None:
        dist.barrier()
        if own_group:
            dist.destroy_process_group()
    _stamp(verbose, "Done:", t0)
    info["t_total_s"] = time.perf_counter() - t0
    return info


def main(argv=None):
    import argparse
    ap = argparse.ArgumentParser(description="TREKIS-3 run on the B200 engine (program flow of Universal_MC_for_SHI_MAIN.f90)")
    ap.add_argument("run_dir")
    ap.add_argument("--nmc", type=int, default=None, help="number of iterations (default: INPUT_PARAMETERS.txt)")
    ap.add_argument("--redo-tables", action="store_true", help="ignore the table cache (the reference's redo_MFP keywords)")
    ap.add_argument("--tables-only", action="store_true", help="stop after the tables (and their cache files) are ready")
    ap.add_argument("--evaluator", default="auto", choices=["auto", "gpu", "host"], help="who integrates the loss function")
    ap.add_argument("--threads", type=int, default=0)
    ap.add_argument("--out-root", default=None, help="where OUTPUT_<material>/ is written (default: the run directory)")
    ap.add_argument("--quiet", action="store_true")
    ap.add_argument("--shi-window-only", action="store_true", help="(tests) tabulate the ion only around its own energy")
    a = ap.parse_args(argv)
    info = run(a.run_dir, nmc=a.nmc, evaluator=None if a.evaluator == "host" else a.evaluator, threads=a.threads,
               redo_tables=a.redo_tables, tables_only=a.tables_only, out_root=a.out_root, verbose=not a.quiet,
               shi_window_only=a.shi_window_only)
    if info["rank"] == 0:
        if "out_dir" in info:
            st = info["stats"] or {}
            print(f"{info['iterations']} iterations in {info['t_mc_s']:.3f} s on {info['world']} GPU(s) "
                  f"({info['iterations'] / info['t_mc_s']:.1f} iterations/s); tables: {info['tables']}; output: {info['out_dir']}")
            if st.get("errors"):
                print("run-time error counters:", st["errors"])
        else:
            print(f"tables: {info['tables']} ({info['t_tables_s']:.2f} s)")
    return 0


if __name__ == "__main__":
    sys.exit(main())
