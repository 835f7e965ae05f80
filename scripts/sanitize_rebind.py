"""The table re-binding path of a persistent handle for compute-sanitizer: bindings (k_companions, k_monotone_rows, staged DMAs) and one
short run per call, staged and direct, two inputs of the same shapes alternating.  Usage: sanitize_rebind.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import trekis3_b200 as tk

cache = os.path.join(ROOT, ".table_cache")
a = tk.Case.load(tk.make_run_dir("/tmp/sanr_a", "C1")); a.build_tables(shi_window_only=True, cache_dir=cache)
b = tk.Case.load(tk.make_run_dir("/tmp/sanr_b", "C1", edits={10: "1   23.5   ! kind of Zeff; fixed value"})); b.build_tables(shi_window_only=True, cache_dir=cache)
for stage in (1, 0):
    for name, c in (("a", a), ("b", b), ("a", a)):
        t, s = tk.do_Monte_Carlo(c, NMC=1, batch=1, stage_uploads=stage)
        print(stage, name, s["total_events"], s["errors"], flush=True)
tk.release_handles()
