"""One batch through the engine (for `ncu --kernel-name ... --launch-skip ...`). Usage: ncu_target.py [CFG] [NIT] [opt=val ...]
Prints: device_ms total_events n_waves kernel_launches; with NCU_STATS_OUT=<file> the run statistics go there as JSON."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import trekis3_b200 as tk

cfg = sys.argv[1] if len(sys.argv) > 1 else "C2"
nit = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
opts = {}
for kv in sys.argv[3:]:
    k, v = kv.split("="); opts[k] = float(v)
case = tk.Case.load(tk.make_run_dir(f"/tmp/run_{cfg}", cfg))
case.build_tables(shi_window_only=True, cache_dir=os.path.join(ROOT, ".table_cache"))
eng = tk.Engine(case, **opts)
st = eng.run_device(0, nit)
print(st["device_ms"], st["total_events"], st["n_waves"], st["kernel_launches"])
if os.environ.get("NCU_STATS_OUT"):
    with open(os.environ["NCU_STATS_OUT"], "w") as f:
        json.dump({"config": cfg, "iterations": nit, "options": opts, "events": st["events"], "cold_events": st["cold_events"],
                   "warm_events": st["warm_events"], "n_waves": st["n_waves"], "kernel_launches": st["kernel_launches"]}, f)
