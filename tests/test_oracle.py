"""The CPU oracle (time-ordered restatement of Monte_Carlo.f90) and the wavefront re-formulation.

No golden Monte-Carlo outputs exist in the reference (unseeded RNG, no tests), so the oracle is pinned by the
reference's own run-time invariants, and the wavefront engine (same header as the CUDA kernels, compiled for the
CPU in tests/emul) is required to reproduce the oracle event by event when both use the same Philox streams."""
import numpy as np
import pytest

import trekis3_b200 as tk
from trekis3_b200.host import split_tallies
import emul_api
import oracle_api

CACHE = tk._abi.REPO + "/.table_cache"


def test_philox_known_answers():
    # Random123 kat_vectors, philox4x32-10
    assert oracle_api.philox([0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert oracle_api.philox([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert oracle_api.philox([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def assert_same(case, to, te, so, se, rtol=1e-9):
    assert so["events"] == se["events"]
    assert so["n_electrons"] == se["n_electrons"] and so["n_photons"] == se["n_photons"]
    lay = case.layout()
    To, Te = split_tallies(lay, to), split_tallies(lay, te)
    for k in To:
        assert np.allclose(To[k], Te[k], rtol=rtol, atol=1e-300), k


@pytest.mark.parametrize("rng_mode", [0, 1])
def test_energy_is_conserved_per_iteration(case_c1, rng_mode):
    # manual section VI: "total energy in Total_numbers.txt is conserved"; north_star: 1e-9 relative
    _, st, totE, totN = oracle_api.run(case_c1, 0, 6, rng_mode=rng_mode)
    assert not st["errors"]
    # the ion (v = 156 A/fs) has left the 10 A layer long before 0.1 fs: rows 1.. must be equal
    drift = np.abs(totE[:, 1:] - totE[:, -1:]) / totE[:, -1:]
    assert drift.max() < 1e-9
    assert np.all(np.diff(totN, axis=1) >= 0) and np.all(totN[:, -1] > 500)
    # deposited energy ~ S_e * layer = 2.6 keV/A * 10 A
    assert 15e3 < totE[:, -1].mean() < 40e3


def test_wavefront_equals_time_ordered_loop_al2o3(case_c1):
    to, so, eo, no = oracle_api.run(case_c1, 0, 5, rng_mode=1)
    te, se, ee, ne = emul_api.run(case_c1, 0, 5, batch=2)
    assert_same(case_c1, to, te, so, se)
    assert np.array_equal(no, ne) and np.allclose(eo, ee, rtol=1e-12)
    assert se["n_waves"] >= 3


def test_atom_without_shells_of_its_own_water(tmp_path):
    """H2O.cdf (shipped with the reference) gives hydrogen ZERO shells: its electron is counted in the valence band of the
    first atom (Reading_files_and_parameters.f90:1496-1503 allocates arrays of length 0 and loops over nothing)."""
    case = tk.Case.load(tk.make_run_dir(str(tmp_path / "w"), ("H2O", 54, 167.0, 0, 10)))
    case.build_tables(shi_window_only=True, cache_dir=CACHE)
    assert case.get("n_atoms") == 2 and case.tables.n_shells == 2
    to, so, eo, no = oracle_api.run(case, 0, 4, rng_mode=1)
    te, se, ee, ne = emul_api.run(case, 0, 4, batch=3)
    assert not so["errors"] and so["events"]["shi"] > 100 and so["events"]["el_elastic"] > 1000
    assert_same(case, to, te, so, se)
    assert np.array_equal(no, ne) and np.allclose(eo, ee, rtol=1e-12)
    drift = np.abs(eo[:, 1:] - eo[:, -1:]) / eo[:, -1:]
    assert drift.max() < 1e-9


@pytest.mark.parametrize("kind", [1, 2, 3, 4])
def test_charge_models_in_the_monte_carlo(tmp_path, kind):
    """The ion's equilibrium charge is recomputed at every ion collision (Monte_Carlo.f90:2196) with the model of input line 10:
    the device code (CUDA physics header, emulated on the CPU) and the oracle agree for every model, and the models differ."""
    case = tk.Case.load(tk.make_run_dir(str(tmp_path / "r"), "C1", edits={10: "%d   23.5   ! kind of Zeff; fixed value" % kind}))
    case.build_tables(shi_window_only=True, cache_dir=CACHE)
    to, so, eo, no = oracle_api.run(case, 0, 3, rng_mode=1)
    te, se, ee, ne = emul_api.run(case, 0, 3, batch=3)
    assert not so["errors"]
    assert_same(case, to, te, so, se)
    assert np.allclose(eo, ee, rtol=1e-12)


SWITCHES = {                                  # INPUT_PARAMETERS.txt line -> text (Reading_files_and_parameters.f90:300-400)
    "target_at_300K": {9: "300.0"},                   # phonon absorption / emission weights
    "elastic_scattering_off": {12: "-1   1"},
    "elastic_Zeff_Barkas_like": {12: "1   0"},
    "plasmon_pole_dispersion": {13: "2   0"},
    "ritchie_dispersion": {13: "3   0"},
    "free_electron_mass": {13: "1   -1"},
    "effective_mass_0.5": {13: "1   0.5"},
    "plasmon_integration_limit": {14: "1"},
    "hole_mass_1": {15: "1.0"},
    "heavy_holes": {15: "1.0d10"},
}


@pytest.mark.parametrize("name", sorted(SWITCHES))
def test_input_switches_through_oracle_and_device_code(tmp_path, name):
    """Every switch of INPUT_PARAMETERS.txt that the path supports, beyond the defaults: tables are rebuilt with it and the
    device code (emulated) gives the oracle's events and tallies with the same Philox streams."""
    case = tk.Case.load(tk.make_run_dir(str(tmp_path / "r"), "C1", edits=SWITCHES[name]))
    case.build_tables(shi_window_only=True)
    to, so, eo, no = oracle_api.run(case, 0, 2, rng_mode=1)
    te, se, ee, ne = emul_api.run(case, 0, 2, batch=2)
    assert not so["errors"] and so["events"]["shi"] > 100
    assert_same(case, to, te, so, se)
    assert np.array_equal(no, ne) and np.allclose(eo, ee, rtol=1e-12)
    drift = np.abs(eo[:, 1:] - eo[:, -1:]) / eo[:, -1:]
    assert drift.max() < 1e-9
    if name == "elastic_scattering_off":
        assert so["events"]["el_elastic"] == 0 and so["events"]["vbh_elastic"] == 0


def test_wavefront_equals_time_ordered_loop_photons_and_radiative_decay(tmp_path):
    d = tk.make_run_dir(str(tmp_path / "c2"), "C2")
    case = tk.Case.load(d)
    case.build_tables(shi_window_only=True, cache_dir=CACHE)
    assert not case.warnings                                       # side-car radiative widths were found
    # short radiative times (test hook) so that the photon channel is exercised by a handful of iterations
    case.set("radiat:0:0", 1.0); case.set("radiat:0:1", 8.0); case.set("radiat:1:0", 2.0)
    to, so, eo, no = oracle_api.run(case, 0, 4, rng_mode=1)
    te, se, ee, ne = emul_api.run(case, 0, 4, batch=3)
    assert so["events"]["radiative"] > 5 and so["events"]["photon"] > 5 and so["events"]["vbh_inelastic"] > 50
    assert_same(case, to, te, so, se)
    drift = np.abs(eo[:, 1:] - eo[:, -1:]) / eo[:, -1:]
    assert drift.max() < 1e-9 and not so["errors"]


def test_wavefront_equals_time_ordered_loop_diamond_hole_ionisation(case_c3):
    to, so, eo, no = oracle_api.run(case_c3, 0, 3, rng_mode=1)
    te, se, ee, ne = emul_api.run(case_c3, 0, 3, batch=8)
    assert so["events"]["vbh_inelastic"] > 100
    assert_same(case_c3, to, te, so, se)


def test_variants_cutoff_linear_grid_emission_mott(tmp_path):
    # cut-off 5 eV, linear time grid of 2.5 fs up to 10 fs, electron emission on
    d = tk.make_run_dir(str(tmp_path / "v1"), "C1", edits={5: "10.0", 6: "2.5 0", 7: "5.0", 17: "4.5 10.0 6.18"})
    case = tk.Case.load(d)
    case.build_tables(shi_window_only=True, cache_dir=CACHE)
    lay = case.layout()
    assert lay.Nt == 4 and case.config.work_function == 4.5
    to, so, eo, no = oracle_api.run(case, 0, 4, rng_mode=1)
    te, se, ee, ne = emul_api.run(case, 0, 4, batch=4)
    assert_same(case, to, te, so, se)
    T = split_tallies(lay, to)
    assert T["Out_Ne_Em"][-1] > 0 and T["Out_E_Em"][-1] > 0 and T["Out_Ee_vs_E_Em"].sum() > 0
    # Mott elastic scattering (kind_of_EMFP = 0)
    d = tk.make_run_dir(str(tmp_path / "v2"), "C1", edits={12: "0   1"})
    case = tk.Case.load(d)
    case.build_tables(shi_window_only=True, cache_dir=CACHE)
    assert case.config.kind_of_EMFP == 0
    to, so, eo, no = oracle_api.run(case, 0, 2, rng_mode=1)
    te, se, ee, ne = emul_api.run(case, 0, 2, batch=2)
    assert_same(case, to, te, so, se)
    drift = np.abs(eo[:, 1:] - eo[:, -1:]) / eo[:, -1:]
    assert drift.max() < 1e-9


def test_batching_does_not_change_results(case_c1):
    t1, s1, _, _ = emul_api.run(case_c1, 3, 9, batch=1)
    t2, s2, _, _ = emul_api.run(case_c1, 3, 9, batch=6)
    assert s1["events"] == s2["events"]
    assert np.allclose(t1, t2, rtol=1e-11, atol=1e-300)


def test_sequential_and_philox_streams_agree_statistically(case_c1):
    # the two RNG modes are different random sequences: means must agree within 3 sigma of the batch-mean error
    n = 24
    _, _, e0, n0 = oracle_api.run(case_c1, 0, n, rng_mode=0)
    _, _, e1, n1 = oracle_api.run(case_c1, 1000, 1000 + n, rng_mode=1)
    for a, b in ((e0[:, -1], e1[:, -1]), (n0[:, -1], n1[:, -1]), (n0[:, 2], n1[:, 2])):
        sig = np.sqrt(a.var(ddof=1) / n + b.var(ddof=1) / n)
        assert abs(a.mean() - b.mean()) < 3.0 * sig + 1e-12


def test_high_multiplicity_metal_runs_in_the_wavefront_engine(case_c4):
    # U 2600 MeV in Au: >1e5 carriers per iteration (the reference's O(N^2) loop would need hours)
    te, se, ee, ne = emul_api.run(case_c4, 0, 1, batch=1)
    assert ne[0, -1] > 5e4 and not se["errors"]
    assert abs(ee[0, 2] - ee[0, -1]) / ee[0, -1] < 1e-9


def test_edge_cases_of_the_iteration_range(case_c1):
    """An empty range gives zero tallies and no events; the iterations just below the 32-bit limit of the iteration index (one
    word of the Philox counter on both sides) run like any other; ranges compose (a ragged split adds up to the whole)."""
    to, so, eo, _ = oracle_api.run(case_c1, 7, 7, rng_mode=1)
    te, se, ee, _ = emul_api.run(case_c1, 7, 7, batch=4)
    assert not to.any() and not te.any() and so["total_events"] == 0 and se["total_events"] == 0 and eo.shape[0] == 0 and ee.shape[0] == 0
    top = 0xffffffff
    to, so, eo, no = oracle_api.run(case_c1, top - 3, top, rng_mode=1)
    te, se, ee, ne = emul_api.run(case_c1, top - 3, top, batch=2)
    assert so["total_events"] > 1000 and not so["errors"]
    assert_same(case_c1, to, te, so, se)
    assert np.array_equal(no, ne) and np.allclose(eo, ee, rtol=1e-12)
    # the histories at the top of the range are not those of the bottom (the iteration index keys the random streams)
    t0, s0, _, _ = oracle_api.run(case_c1, 0, 3, rng_mode=1)
    assert s0["events"] != so["events"]
    # ragged split: [top-3, top-2) + [top-2, top) == [top-3, top)
    ta, sa, _, _ = emul_api.run(case_c1, top - 3, top - 2, batch=2)
    tb, sb, _, _ = emul_api.run(case_c1, top - 2, top, batch=2)
    assert {k: sa["events"][k] + sb["events"][k] for k in sa["events"]} == se["events"]
    lay = case_c1.layout()
    i = tk.TALLY_NAMES.index("Out_diff_coeff")          # a running mean over the iterations, not a sum (Monte_Carlo.f90:1094-1098)
    keep = np.ones(lay.total, bool); keep[lay.off[i]: lay.off[i] + lay.len[i]] = False
    assert np.allclose((ta + tb)[keep], te[keep], rtol=1e-12, atol=1e-300)


_REF = "/root/reference"
# shipped materials beyond the five of the BASELINE configurations and H2O: compounds of three and four elements, an ionic crystal,
# a polymer with an atom without shells of its own, a semiconductor given by its chemical formula, metals
_MORE_MATERIALS = ["LiF", "yag", "Olivine", "Si_sp", "C2H4", "Si", "Cu", "Al"]


@pytest.mark.skipif(not __import__("os").path.isdir(_REF + "/INPUT_CDF"), reason="the reference tree is only mounted in the build container")
@pytest.mark.parametrize("material", _MORE_MATERIALS)
def test_wavefront_equals_time_ordered_loop_on_more_shipped_materials(tmp_path, material):
    """The input contract over more of the .cdf / .dos files the reference ships (read in place): tables built by the host half,
    then the wavefront engine (the CUDA kernels' physics header) against the time-ordered oracle on the same random streams."""
    import os
    import shutil
    if not os.path.exists(f"{_REF}/INPUT_CDF/{material}.cdf"):
        pytest.skip("not shipped")
    dd = tmp_path / "data"
    os.makedirs(dd)
    os.symlink(f"{_REF}/INPUT_CDF", dd / "INPUT_CDF")
    os.symlink(f"{_REF}/INPUT_DOS", dd / "INPUT_DOS")
    os.symlink(os.path.join(tk._abi.REPO, "data", "INPUT_EADL"), dd / "INPUT_EADL")
    shutil.copy(os.path.join(tk._abi.REPO, "data", "INPUT_PARAMETERS.default.txt"), dd)
    try:
        case = tk.Case.load(tk.make_run_dir(str(tmp_path / "run"), (material, 54, 167.0, 0, 3), data_dir=str(dd),
                                            edits={5: "10.0"}))          # 10 fs: the oracle's cost grows with the square of the cascade
    except RuntimeError as e:
        pytest.skip(f"refused by the reader: {e}")
    case.build_tables(shi_window_only=True, cache_dir=CACHE)
    n_it = 1 if material == "Cu" else 2           # copper's cascade is the largest of the set (the oracle's cost is quadratic in it)
    to, so, eo, no = oracle_api.run(case, 0, n_it, rng_mode=1)
    te, se, ee, ne = emul_api.run(case, 0, n_it, batch=2)
    assert not so["errors"], so["errors"]
    assert so["events"]["shi"] > 50 and so["events"]["el_inelastic"] > 100 and so["total_events"] > 1000, so["events"]
    if so["events"] == se["events"]:
        assert_same(case, to, te, so, se)
        assert np.array_equal(no, ne) and np.allclose(eo, ee, rtol=1e-12)
    else:
        # Cu: the d band's DOS has gaps, check_hole_parameters (Monte_Carlo.f90:682-721) snaps many holes exactly onto grid
        # energies, and there one unit in the last place decides the bin of the next lookup.  The engine interpolates across
        # energies with the LOGARITHMS of its two row samples where the oracle (as the reference) takes exp and log again, so
        # transferred energies differ in the last bits (traced: the first differing event of a run is always such a last-bit
        # difference); a handful of histories per iteration then take the other branch.  Held to the bound of the GPU tests.
        assert material in ("Cu",), (material, so["events"], se["events"])
        for k in so["events"]:
            assert abs(so["events"][k] - se["events"][k]) <= max(2, 2e-3 * so["events"][k]), (k, so["events"][k], se["events"][k])
        lay = case.layout()
        To, Te = split_tallies(lay, to), split_tallies(lay, te)
        for k in To:
            assert np.isclose(To[k].sum(), Te[k].sum(), rtol=5e-3), k
        assert np.allclose(eo, ee, rtol=1e-9)          # total energy per iteration and grid time
    drift = np.abs(eo[:, 1:] - eo[:, -1:]) / eo[:, -1:]
    assert drift.max() < 1e-9
