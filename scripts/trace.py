"""Per-launch trace of one call (option profile=2 prints to stderr). Usage: trace.py CFG NIT [opt=val ...]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import trekis3_b200 as tk

cfg, nit = sys.argv[1], int(sys.argv[2])
opts = {}
for kv in sys.argv[3:]:
    k, v = kv.split("="); opts[k] = float(v)
case = tk.Case.load(tk.make_run_dir(f"/tmp/run_{cfg}", cfg))
case.build_tables(shi_window_only=True, cache_dir=os.path.join(ROOT, ".table_cache"))
eng = tk.Engine(case, **opts)
eng.run_device(0, nit); eng.run_device(0, nit)
eng.set_option("profile", float(os.environ.get("TRK_PROFILE", "2")))
st = eng.run_device(0, nit)
print("device_ms", st["device_ms"], "events", st["events"] if "events" in st else st["total_events"])
