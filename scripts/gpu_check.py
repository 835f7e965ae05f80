"""Quick GPU bring-up check: engine vs oracle (same Philox streams) + timing. Usage: gpu_check.py CFG NIT_CMP NIT_TIME"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import trekis3_b200 as tk
import oracle_api as oa
from trekis3_b200.host import split_tallies

cfg = sys.argv[1]; n_cmp = int(sys.argv[2]); n_time = int(sys.argv[3])
opts = {}
for kv in sys.argv[4:]:
    k, v = kv.split("="); opts[k] = float(v)
case = tk.Case.load(tk.make_run_dir(f"/tmp/run_{cfg}", cfg))
case.build_tables(shi_window_only=True, cache_dir=os.path.join(ROOT, ".table_cache"))
eng = tk.Engine(case, **opts)
if n_cmp > 0:
    t = time.time(); tg, sg = eng.run(0, n_cmp); print("gpu %.3fs" % (time.time() - t))
    t = time.time(); to, so, eo, no = oa.run(case, 0, n_cmp, rng_mode=1); print("oracle %.3fs" % (time.time() - t))
    print("gpu   ", json.dumps(sg))
    print("oracle", json.dumps(so["events"]))
    eg = eng.iteration_energies(n_cmp)
    print("iter energies rel diff max:", np.max(np.abs(eg - eo) / np.maximum(eo, 1e-300)))
    lay = case.layout()
    Tg, To = split_tallies(lay, tg), split_tallies(lay, to)
    for k in Tg:
        a, b = To[k], Tg[k]
        if not np.any(a) and not np.any(b):
            continue
        den = np.maximum(np.abs(a), np.abs(b)); den[den == 0] = 1
        rel = np.abs(a - b) / den
        print("%-16s max rel %.3e  median rel %.3e  sum o=%.8e g=%.8e" % (k, rel.max(), np.median(rel[den > 0]), a.sum(), b.sum()))
if n_time > 0:
    eng.set_option("profile", 1)
    for rep in range(3):
        t = time.time(); st = eng.run_device(0, n_time); dt = time.time() - t
        print("time: %d its wall %.4fs dev %.3f ms -> %.1f it/s, %.3e events/s, waves %d launches %d drift %.2e errors %s" % (
            n_time, dt, st["device_ms"], n_time / (st["device_ms"] * 1e-3), st["total_events"] / (st["device_ms"] * 1e-3), st["n_waves"], st["kernel_launches"], st["max_energy_drift"], st["errors"]))
    print("kernel ms (3 reps):", {k: round(v["ms"], 2) for k, v in eng.kernel_times().items()})
