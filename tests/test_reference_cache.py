"""The reference's on-disk table cache (SURVEY.md 3.5, row N2 of 8(f)): names, row layout and number formats of
Analytical_IMFPs.f90:262-786, 917-1606, 2306-2707, and the round trip through them.

What pins the formats: the reference writes with the width-less descriptors '(f)', '(e)', '(es)' and reads the files back
with the same descriptors and advance='no', i.e. as fixed 25-character fields (real(8): F25.16 / E25.16 / ES25.16); files
shipped with the reference that were written this way (INPUT_DOS/Al2O3.dos, copied to data/) show the same fields."""
import os
import re

import numpy as np
import pytest

import trekis3_b200 as tk

E_FIELD = re.compile(r"^ +-?0\.\d{16}E[+-]\d{2}$")
ES_FIELD = re.compile(r"^ +-?\d\.\d{16}E[+-]\d{2}$")
F_FIELD = re.compile(r"^ +-?\d+\.\d{16}$")


def fields(line, w=25):
    assert len(line) % w == 0, (len(line), line)
    return [line[i:i + w] for i in range(0, len(line), w)]


@pytest.fixture(scope="module")
def cache_c2(case_c2, tmp_path_factory):
    root = str(tmp_path_factory.mktemp("refcache_c2"))
    n = case_c2.write_reference_cache(root)
    return root, n


def test_shipped_dos_file_shows_the_width_less_e_format():
    with open(os.path.join(tk._abi.REPO, "data", "INPUT_DOS", "Al2O3.dos")) as f:
        line = f.readline().rstrip("\n")
    assert all(E_FIELD.match(x) for x in fields(line))


def test_names_follow_the_reference(case_c2, case_c1, cache_c2):
    root, n_files = cache_c2
    name = case_c2.reference_cache_name
    assert name("dir_material") == "OUTPUT_SiO2_cryst" and name("dir_ion") == "OUTPUT_SiO2_cryst/OUTPUT_Au_in_SiO2_cryst"
    # default input: Ritchie-Howie CDF (kind_of_CDF 0), free-electron dispersion, electron mass from DOS, 0 K, Z = 1 phonons
    assert name("el_imfp") == "OUTPUT_Electron_IMFPs_Free_CDF_DOS_0.00_K.dat"
    assert name("hole_imfp") == "OUTPUT_Hole_IMFPs_CDF_CDF_0.00_K.dat"
    assert name("photon_imfp") == "OUTPUT_Photon_IMFPs_CDF_0.00_K.dat"
    assert name("el_emfp") == "OUTPUT_Electron_EMFPs_CDF_Z=1_0.00_K.dat"
    assert name("hole_emfp") == "OUTPUT_Hole_CDF_EMFPs_0.00_K.dat"
    assert name("shi_stem") == "OUTPUT_Au_CDF_Barkas_P"
    assert case_c1.reference_cache_name("shi_stem") == "OUTPUT_Xe_CDF_Barkas_P"
    dm = os.path.join(root, name("dir_material"))
    for k in ("el_imfp", "hole_imfp", "photon_imfp", "el_emfp", "hole_emfp"):
        assert os.path.isfile(os.path.join(dm, name(k))), k
    for sfx in ("_IMFP.dat", "_dEdx.dat", "_effective_charges.dat", "_Range.dat"):
        assert os.path.isfile(os.path.join(root, name("dir_ion"), name("shi_stem") + sfx))
    a = case_c2.table_arrays()
    ns, n_ei, n_ee, n_hi, n_he = a["ei_L"].shape[0], len(a["ei_E"]), len(a["ee_E"]), len(a["hi_E"]), len(a["he_E"])
    # one differential file per (shell, grid energy) for electrons, per grid energy for the valence hole and the elastic channels
    diff = os.listdir(os.path.join(root, name("dir_diff")))
    assert len(diff) == ns * n_ei + n_hi + n_ee + n_he
    assert n_files == len(diff) + 5 + 4
    # <table file without OUTPUT_ and .dat>_<atom>_<shell>_<E as f14.3>.dat
    assert "Electron_IMFPs_Free_CDF_DOS_0.00_K_Si_K-shell_%.3f.dat" % a["ei_E"][0] in diff or \
        any(d.startswith("Electron_IMFPs_Free_CDF_DOS_0.00_K_Si_") and d.endswith("_%.3f.dat" % a["ei_E"][0]) for d in diff)
    assert "Electron_EMFPs_CDF_Z=1_0.00_K_%.3f.dat" % a["ee_E"][0] in diff
    assert "Hole_CDF_EMFPs_0.00_K_%.3f.dat" % a["he_E"][-1] in diff


def test_rows_are_fixed_25_character_fields(case_c2, cache_c2):
    root, _ = cache_c2
    name = case_c2.reference_cache_name
    a = case_c2.table_arrays()
    ns = a["ei_L"].shape[0]
    dm = os.path.join(root, name("dir_material"))
    with open(os.path.join(dm, name("el_imfp"))) as f:
        lines = f.read().splitlines()
    assert len(lines) == len(a["ei_E"])                      # the reference's validity test: rows == grid size (:317-323)
    for i in (0, len(lines) // 2, len(lines) - 1):
        fl = fields(lines[i])
        assert len(fl) == 1 + ns + 1 and F_FIELD.match(fl[0]) and all(E_FIELD.match(x) for x in fl[1:])
        L = np.array([float(x) for x in fl[1:]])
        assert float(fl[0]) == a["ei_E"][i]
        assert np.allclose(L[:-1], a["ei_L"][:, i], rtol=1e-15)
        assert L[-1] == pytest.approx(1.0 / np.sum(1.0 / a["ei_L"][:, i]), rel=1e-14)       # total column (:626)
    with open(os.path.join(dm, name("el_emfp"))) as f:
        fl = fields(f.readline().rstrip("\n"))
    assert len(fl) == 2 and F_FIELD.match(fl[0]) and E_FIELD.match(fl[1])
    d0 = sorted(os.listdir(os.path.join(root, name("dir_diff"))))[0]
    with open(os.path.join(root, name("dir_diff"), d0)) as f:
        fl = fields(f.readline().rstrip("\n"))
    assert len(fl) == 2 and all(ES_FIELD.match(x) for x in fl)
    # ion files: energies in MeV (:2663-2665), Range with two header lines (:2498-2499)
    di = os.path.join(root, name("dir_ion"))
    with open(os.path.join(di, name("shi_stem") + "_IMFP.dat")) as f:
        lines = f.read().splitlines()
    assert len(lines) == len(a["shi_E"])
    fl = fields(lines[0])
    assert len(fl) == ns + 2 and all(E_FIELD.match(x) for x in fl) and float(fl[0]) == pytest.approx(a["shi_E"][0] / 1e6, rel=1e-15)
    with open(os.path.join(di, name("shi_stem") + "_Range.dat")) as f:
        lines = f.read().splitlines()
    assert lines[0] == "# Energy dEdx    Range" and lines[1] == "# [eV] [eV/A]    [A]" and len(lines) == 2 + len(a["shi_E"])
    with open(os.path.join(di, name("shi_stem") + "_effective_charges.dat")) as f:
        z = np.array([[float(x) for x in fields(l)] for l in f.read().splitlines()])
    assert np.all(np.diff(z[:, 1]) > 0) and 0 < z[0, 1] and z[-1, 1] <= 79.0        # Barkas charge grows with the energy


@pytest.mark.parametrize("cfg", ["C2", "C3"])
def test_round_trip_gives_the_tables_back(cfg, request, tmp_path):
    """write -> read into a fresh case: every table equal to 16 significant digits (what the text format keeps)."""
    case = request.getfixturevalue("case_" + cfg.lower())
    root = str(tmp_path / "cache")
    case.write_reference_cache(root)
    c2 = tk.Case.load(tk.make_run_dir(str(tmp_path / "run"), cfg))
    c2.read_reference_cache(root)
    a, b = case.table_arrays(), c2.table_arrays()
    assert set(a) == set(b)
    for k in a:
        assert a[k].shape == b[k].shape, k
        if a[k].dtype.kind in "iu":
            assert np.array_equal(a[k], b[k]), k
        elif k == "shi_dEdx" or k.endswith("dEdx"):
            assert np.allclose(a[k], b[k], rtol=2e-15, atol=0), k
        else:
            assert np.allclose(a[k], b[k], rtol=2e-15, atol=0), k
    # the scalars of the configuration that depend on the tables
    assert bytes(case.config) == bytes(c2.config)


def test_oracle_results_from_reread_tables_are_statistically_identical(case_c1, tmp_path):
    """Tables that went through the text cache differ in the 16th digit: the Monte-Carlo result stays the same event for
    event (identical Philox streams) except where a draw falls within 1e-15 of a channel boundary."""
    import oracle_api as oa
    root = str(tmp_path / "cache")
    case_c1.write_reference_cache(root)
    c2 = tk.Case.load(tk.make_run_dir(str(tmp_path / "run"), "C1"))
    c2.read_reference_cache(root)
    t1, s1, e1, _ = oa.run(case_c1, 0, 2, rng_mode=1)
    t2, s2, e2, _ = oa.run(c2, 0, 2, rng_mode=1)
    assert s1["events"] == s2["events"]
    assert np.allclose(e1, e2, rtol=1e-9)
    assert np.allclose(t1, t2, rtol=1e-6, atol=1e-12)


def test_missing_or_truncated_files_are_refused(case_c1, tmp_path):
    root = str(tmp_path / "cache")
    case_c1.write_reference_cache(root)
    name = case_c1.reference_cache_name
    c2 = tk.Case.load(tk.make_run_dir(str(tmp_path / "run"), "C1"))
    with pytest.raises(RuntimeError, match="missing|cannot read"):
        c2.read_reference_cache(str(tmp_path / "nowhere"))
    p = os.path.join(root, name("dir_material"), name("el_imfp"))
    with open(p) as f:
        lines = f.readlines()
    with open(p, "w") as f:
        f.writelines(lines[:-1])
    with pytest.raises(RuntimeError, match="grid mismatch"):
        c2.read_reference_cache(root)
    with pytest.raises(RuntimeError, match="not built"):
        _ = c2.config
    with open(p, "w") as f:
        f.writelines(lines)
    c2.read_reference_cache(root)
    assert c2.tables.n_ei == case_c1.tables.n_ei


def test_reader_accepts_fortran_spellings(case_c1, tmp_path):
    """List-directed files of other compilers: D exponents, three-digit exponents without a letter, free spacing."""
    root = str(tmp_path / "cache")
    case_c1.write_reference_cache(root)
    name = case_c1.reference_cache_name
    p = os.path.join(root, name("dir_material"), name("hole_emfp"))
    with open(p) as f:
        lines = f.read().splitlines()
    with open(p, "w") as f:
        for l in lines:
            e, L = l.split()
            f.write(" %s  %s\n" % (e, L.replace("E", "D")))
    c2 = tk.Case.load(tk.make_run_dir(str(tmp_path / "run"), "C1"))
    c2.read_reference_cache(root)
    assert np.allclose(c2.table_arrays()["he_L"], case_c1.table_arrays()["he_L"], rtol=2e-15)
