"""GPU engine vs oracle on the same Philox streams: per tally array the number of bins that differ beyond 1e-9 and the largest
relative difference, for several option sets.  Usage: parity_diff.py CFG NIT "opt=val,..." ..."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import trekis3_b200 as tk
from trekis3_b200.host import split_tallies
import oracle_api

if os.environ.get("TRK3_GPU_LIB"):          # A/B against another build of the engine library
    _orig = tk._abi.lib_path
    tk._abi.lib_path = lambda name: os.environ["TRK3_GPU_LIB"] if name == "gpu" else _orig(name)
    tk.engine._abi.lib_path = tk._abi.lib_path

cfg, nit = sys.argv[1], int(sys.argv[2])
case = tk.Case.load(tk.make_run_dir(f"/tmp/run_{cfg}", cfg))
case.build_tables(shi_window_only=True, cache_dir=os.path.join(ROOT, ".table_cache"))
to, so, eo, _ = oracle_api.run(case, 0, nit, rng_mode=1)
lay = case.layout()
To = split_tallies(lay, to)
for variant in sys.argv[3:] or [""]:
    opts = {}
    for kv in filter(None, variant.split(",")):
        k, v = kv.split("="); opts[k] = float(v)
    eng = tk.Engine(case, **opts)
    tg, sg = eng.run(0, nit)
    eg = eng.iteration_energies(nit)
    Tg = split_tallies(lay, tg)
    print("==", variant or "defaults", "events equal:", sg["events"] == so["events"], "errors", sg["errors"])
    if sg["events"] != so["events"]:
        print("   ", {k: (sg["events"][k], so["events"][k]) for k in so["events"] if sg["events"][k] != so["events"][k]})
    print("   iteration energies: max rel diff %.3e" % np.max(np.abs(eg - eo) / np.maximum(np.abs(eo), 1e-300)))
    for k in To:
        den = np.maximum(np.maximum(np.abs(Tg[k]), np.abs(To[k])), 1e-300)
        rel = np.abs(Tg[k] - To[k]) / den
        bad = rel > 1e-9
        if bad.any():
            print("   %-18s bins differing %4d of %5d nonzero, worst rel %.3e, sum rel diff %.3e" % (k, bad.sum(), np.count_nonzero(To[k]), rel.max(), abs(Tg[k].sum() - To[k].sum()) / max(abs(To[k].sum()), 1e-300)))
    eng.close()
