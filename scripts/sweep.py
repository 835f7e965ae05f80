"""Option sweep on one configuration: device ms per call (min of REPS after one warm-up) for each variant.
Usage: sweep.py CFG NIT "opt=val,opt=val" "opt=val" ...   (an empty string = defaults)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import trekis3_b200 as tk

if os.environ.get("TRK3_GPU_LIB"):          # A/B against another build of the engine library
    _orig = tk._abi.lib_path
    tk._abi.lib_path = lambda name: os.environ["TRK3_GPU_LIB"] if name == "gpu" else _orig(name)

cfg, nit = sys.argv[1], int(sys.argv[2])
case = tk.Case.load(tk.make_run_dir(f"/tmp/run_{cfg}", cfg))
case.build_tables(shi_window_only=True, cache_dir=os.path.join(ROOT, ".table_cache"))
ref_events = None
for variant in sys.argv[3:]:
    opts = {}
    for kv in filter(None, variant.split(",")):
        k, v = kv.split("="); opts[k] = float(v)
    eng = tk.Engine(case, **opts)
    eng.set_option("profile", 1)
    ms = []
    for rep in range(4):
        st = eng.run_device(0, nit)
        ms.append(st["device_ms"])
    if ref_events is None:
        ref_events = st["total_events"]
    kt = {k: round(v["ms"] / 4, 2) for k, v in eng.kernel_times().items()}
    print("%-60s min %.3f ms  (all %s)  events %d%s launches %d errors %s\n    %s" % (
        variant or "defaults", min(ms[1:]), " ".join("%.2f" % m for m in ms), st["total_events"], "" if st["total_events"] == ref_events else " MISMATCH",
        st["kernel_launches"], st["errors"], kt), flush=True)
    del eng
