"""Regression fixtures of this repository's own results (NOT of the reference: it ships no golden vectors and cannot be
built here -- SURVEY.md 8(c), parity unpinned).  They pin (a) the host table builder and (b) the CPU oracle against
accidental changes: a few table values per configuration, and the event counts, total energies and a few tally sums of a
short oracle run with the Philox streams.  Regenerate with `python tests/golden/make_golden.py` after a deliberate change
of the algorithm (e.g. of the stream convention) and say so in the commit."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import trekis3_b200 as tk
import oracle_api
from trekis3_b200.host import split_tallies

N_IT = {"C1": 3, "C2": 3, "C3": 3, "C4": 1}


def fingerprint(cfg, run_dir):
    case = tk.Case.load(tk.make_run_dir(run_dir, cfg))
    case.build_tables(shi_window_only=True)
    a = case.table_arrays()
    tab = {}
    for k, v in sorted(a.items()):
        v = np.asarray(v).ravel()
        ok = v[np.isfinite(v) & (np.abs(v) < 1e14)] if v.dtype.kind == "f" else v
        tab[k] = {"n": int(v.size), "sum": float(np.sum(ok)), "probe": [float(x) for x in v[:: max(1, v.size // 7)][:7]]}
    n = N_IT[cfg]
    tallies, st, totE, totN = oracle_api.run(case, 0, n, rng_mode=1, threads=1)
    T = split_tallies(case.layout(), tallies)
    return {"tables": tab, "iterations": n, "events": st["events"], "n_electrons": st["n_electrons"],
            "iter_totE": totE.tolist(), "iter_totNel": totN.tolist(),
            "tally_sums": {k: float(np.sum(v)) for k, v in sorted(T.items())}}


if __name__ == "__main__":
    out = {cfg: fingerprint(cfg, f"/tmp/golden_{cfg}") for cfg in N_IT}
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "fingerprints.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print("wrote fingerprints.json")
