"""Table builder: host threads vs the GPU evaluator.  Usage: bench_tables.py [CFG] [full|window]
Prints one JSON line: wall seconds of both builds, device ms and number of q-integrals of the GPU one, worst relative
difference between the two sets of tables."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import trekis3_b200 as tk
from trekis3_b200.engine import dcs_stats

cfg = sys.argv[1] if len(sys.argv) > 1 else "C2"
window = (sys.argv[2] if len(sys.argv) > 2 else "full") == "window"
res = {}
tabs = {}
for ev in ("gpu", None, "gpu"):          # the first GPU build also pays the CUDA context creation: reported separately
    c = tk.Case.load(tk.make_run_dir(f"/tmp/bt_{cfg}_{ev}", cfg))
    if ev == "gpu":
        dcs_stats(reset=True)
    t = time.time()
    c.build_tables(shi_window_only=window, evaluator=ev)
    dt = time.time() - t
    key = "gpu_first" if (ev == "gpu" and "gpu_first" not in res) else ("gpu" if ev == "gpu" else "host")
    res[key] = dt
    if ev == "gpu":
        ms, n = dcs_stats()
        res["gpu_device_ms"], res["q_integrals"] = ms, n
    tabs[key] = c.table_arrays()
worst = 0.0
for k, a in tabs["host"].items():
    b = tabs["gpu"][k]
    if a.dtype.kind != "f" or a.size == 0:
        continue
    den = np.maximum(np.abs(a), np.abs(b)); den[den == 0] = 1.0
    worst = max(worst, float(np.max(np.abs(a - b) / den)))
print(json.dumps({"what": "CDF table build (Cross_sections.f90 TotIMFP / Tot_EMFP / SHI_TotIMFP)", "config": cfg, "shi_grid": "window" if window else "full",
                  "host_threads": os.cpu_count(), "host_wall_s": round(res["host"], 3), "gpu_wall_s": round(res["gpu"], 3),
                  "gpu_first_call_wall_s": round(res["gpu_first"], 3), "gpu_device_ms": round(res["gpu_device_ms"], 2),
                  "q_integrals": res["q_integrals"], "q_integrals_per_s_device": round(res["q_integrals"] / (res["gpu_device_ms"] * 1e-3)),
                  "speedup_wall": round(res["host"] / res["gpu"], 2), "worst_relative_difference": worst}))
