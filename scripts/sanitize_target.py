"""Small runs of the engine for compute-sanitizer (memcheck / racecheck / initcheck): the default path (C1, C2 with photons) and the
non-lean kernels with every table family bound (DSF scattering on diamond, delta-function CDF).  Usage: sanitize_target.py [N_ITER]"""
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import pathlib

import trekis3_b200 as tk

n = int(sys.argv[1]) if len(sys.argv) > 1 else 3
cache = os.path.join(ROOT, ".table_cache")
cases = []
for cfg in ("C1", "C2"):
    c = tk.Case.load(tk.make_run_dir(f"/tmp/san_{cfg}", cfg)); c.build_tables(shi_window_only=True, cache_dir=cache); cases.append((cfg, c))
c = tk.Case.load(tk.make_run_dir("/tmp/san_delta", "C1", edits={13: "4   0   ! delta-function CDF"})); c.build_tables(shi_window_only=True, cache_dir=cache); cases.append(("C1-delta", c))
from test_dsf import dsf_run_dir
c = tk.Case.load(dsf_run_dir(pathlib.Path(tempfile.mkdtemp()), cfg="C3", material="Diamond")); c.build_tables(shi_window_only=True); cases.append(("C3-DSF", c))
for name, c in cases:
    eng = tk.Engine(c, batch=2)
    t, s = eng.run(0, n)
    print(name, s["total_events"], s["errors"], s["max_energy_drift"], flush=True)
    eng.close()
