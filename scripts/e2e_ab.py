"""End-to-end call time of do_Monte_Carlo (host buffers in, host buffers out) with the staged and the direct table upload, and the
time of the table re-binding alone.  Usage: e2e_ab.py CFG NIT [REPS]"""
import os
import statistics
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import trekis3_b200 as tk

cfg, nit = sys.argv[1], int(sys.argv[2])
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 8
case = tk.Case.load(tk.make_run_dir(f"/tmp/run_e2e_{cfg}", cfg, nmc=nit))
case.build_tables(shi_window_only=False, cache_dir=os.path.join(ROOT, ".table_cache"))
for stage in (1, 0, 1, 0):
    tk.release_handles()
    tk.do_Monte_Carlo(case, NMC=nit, stage_uploads=stage)      # creates the handle
    tk.do_Monte_Carlo(case, NMC=nit, stage_uploads=stage)      # first re-binding (allocates the pinned mirror)
    ms, dev = [], []
    for r in range(reps):
        t0 = time.perf_counter()
        _, st = tk.do_Monte_Carlo(case, NMC=nit, it_begin=r * nit, stage_uploads=stage)
        ms.append((time.perf_counter() - t0) * 1e3); dev.append(st["device_ms"])
    eng = next(iter(tk.engine._handles.values()))
    rb = []
    for r in range(reps):
        t0 = time.perf_counter()
        eng.reload_tables(case)
        t1 = time.perf_counter()
        eng.zero_device_tallies()                               # synchronises the engine's stream
        rb.append(((t1 - t0) * 1e3, (time.perf_counter() - t0) * 1e3))
    print("stage_uploads=%d: call min %.3f median %.3f ms (device section min %.3f ms); re-binding alone: host %.3f ms, until the device has it %.3f ms; %d bytes" % (
        stage, min(ms), statistics.median(ms), min(dev), min(a for a, _ in rb), min(b for _, b in rb), eng.table_bytes()), flush=True)
tk.release_handles()
