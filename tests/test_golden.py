"""Regression fixtures (tests/golden/fingerprints.json, made by tests/golden/make_golden.py): the host table builder and
the CPU oracle still give what they gave when the fixtures were committed.  These are this repository's own results --
the reference ships no golden vectors and cannot be built here (parity unpinned)."""
import json
import os

import numpy as np
import pytest

import oracle_api
import trekis3_b200 as tk
from trekis3_b200.host import split_tallies

GOLD = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "fingerprints.json")))


@pytest.mark.parametrize("cfg", ["C1", "C2", "C3", "C4"])
def test_tables_match_the_committed_fingerprints(cfg, request):
    case = request.getfixturevalue("case_" + cfg.lower())
    a = case.table_arrays()
    g = GOLD[cfg]["tables"]
    assert sorted(a) == sorted(g)
    for k, v in a.items():
        v = np.asarray(v).ravel()
        assert v.size == g[k]["n"], k
        ok = v[np.isfinite(v) & (np.abs(v) < 1e14)] if v.dtype.kind == "f" else v
        assert float(np.sum(ok)) == pytest.approx(g[k]["sum"], rel=1e-12, abs=1e-300), k
        probe = [float(x) for x in v[:: max(1, v.size // 7)][:7]]
        assert probe == pytest.approx(g[k]["probe"], rel=1e-12, abs=1e-300), k


@pytest.mark.parametrize("cfg", ["C1", "C2", "C3"])       # (C4's oracle run takes half an hour: its fixture is kept for manual checks)
def test_oracle_matches_the_committed_fingerprints(cfg, request):
    case = request.getfixturevalue("case_" + cfg.lower())
    g = GOLD[cfg]
    tallies, st, totE, totN = oracle_api.run(case, 0, g["iterations"], rng_mode=1, threads=1)
    assert st["events"] == g["events"] and st["n_electrons"] == g["n_electrons"]
    assert np.allclose(totE, np.array(g["iter_totE"]), rtol=1e-10, atol=0)
    assert np.array_equal(totN, np.array(g["iter_totNel"]))
    T = split_tallies(case.layout(), tallies)
    for k, v in T.items():
        assert float(np.sum(v)) == pytest.approx(g["tally_sums"][k], rel=1e-9, abs=1e-300), k
