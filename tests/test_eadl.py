"""EADL2023.ALL (ENDL format) as the source of what a .cdf leaves out (SURVEY.md Appendix C, row N4 of 8(f);
check_atomic_parameters / READ_EADL_TYPE_FILE_int / _real, Dealing_with_EADL.f90:312-742).

The EPICS2023 file is not part of the reference tree, so the tests write ENDL blocks themselves, in the layout the
reference's reader walks: two header lines ('(I3,I3,I2,I2,E11.4,I6)' / '(I2,I3,I3,E11.4)'), data lines '(2E11.4)' read
list-directed, and a terminator line with '1' in column 72.  The NUMBERS below are made up for the tests (they are not
EADL values); what is checked is the access pattern, the unit conversions and the sub-shell rules."""
import os
import shutil

import numpy as np
import pytest

import trekis3_b200 as tk

G_H, G_E = 1.05457162853e-34, 1.602176487e-19            # Universal_Constants.f90
DATA = os.path.join(tk._abi.REPO, "data")


def endl_block(Z, A, C, I, rows, S=0):
    h1 = "%3d%3d%2d%2d%11.4E%6d" % (Z, A, 0, 0, float(A), 230101)
    h2 = "%2d%3d%3d%11.4E" % (C, I, S, 0.0)
    out = [h1, h2]
    for d, v in rows:
        out.append("%11.4E%11.4E" % (d, v))
    out.append(" " * 71 + "1")
    return out


def write_eadl(path, table):
    """table: {Z: (A, {I: [(designator, value_in_MeV_or_count), ...]})}"""
    lines = []
    for Z in sorted(table):
        A, blocks = table[Z]
        for I in sorted(blocks):
            lines += endl_block(Z, A, 91 if I < 920 else 92, I, blocks[I])
    with open(path, "w") as f:
        f.write("\n".join(lines) + "\n")


# designators as EADL lists them: sub-shells only (1 K, 3 L1, 5 L2, 6 L3, 8 M1, 10 M2, 11 M3); no whole-shell entries
SI = (28, {
    912: [(1, 2.0), (3, 2.0), (5, 2.0), (6, 4.0), (8, 2.0), (10, 2.0)],
    913: [(1, 1.8389e-3), (3, 1.5e-4), (5, 1.0e-4), (6, 0.99e-4), (8, 1.3e-5), (10, 6.5e-6)],
    914: [(1, 2.5e-3), (3, 3.0e-4), (5, 2.8e-4), (6, 2.7e-4), (8, 4.0e-5), (10, 2.5e-5)],
    921: [(1, 2.0e-8), (3, 3.0e-13), (5, 6.0e-12), (6, 9.0e-12)],
    922: [(1, 4.0e-7), (3, 9.0e-7), (5, 3.0e-9), (6, 3.0e-9)],
})
OX = (16, {
    912: [(1, 2.0), (3, 2.0), (5, 1.33), (6, 2.67)],
    913: [(1, 5.38e-4), (3, 2.8e-5), (5, 1.4e-5), (6, 1.4e-5)],
    914: [(1, 7.9e-4), (3, 7.0e-5), (5, 6.0e-5), (6, 6.0e-5)],
    921: [(1, 1.0e-9)],
    922: [(1, 1.5e-7)],
})


def run_dir(tmp_path, cdf_text=None, eadl=None, sidecar=True, config="C2"):
    """A data directory of its own: INPUT_EADL with (or without) a synthetic EADL2023.ALL, optionally an edited .cdf."""
    dd = tmp_path / "data"
    os.makedirs(dd / "INPUT_EADL")
    shutil.copy(os.path.join(DATA, "INPUT_PARAMETERS.default.txt"), dd)
    shutil.copy(os.path.join(DATA, "INPUT_EADL", "INPUT_atomic_data.dat"), dd / "INPUT_EADL")
    if sidecar:
        shutil.copy(os.path.join(DATA, "INPUT_EADL", "radiative_widths.dat"), dd / "INPUT_EADL")
    os.symlink(os.path.join(DATA, "INPUT_DOS"), dd / "INPUT_DOS")
    if cdf_text is None:
        os.symlink(os.path.join(DATA, "INPUT_CDF"), dd / "INPUT_CDF")
    else:
        os.makedirs(dd / "INPUT_CDF")
        with open(dd / "INPUT_CDF" / "SiO2_cryst.cdf", "w") as f:
            f.write(cdf_text)
    if eadl is not None:
        write_eadl(str(dd / "INPUT_EADL" / "EADL2023.ALL"), eadl)
    return tk.make_run_dir(str(tmp_path / "run"), config, data_dir=str(dd))


def t_fs(width_eV):
    return 1e15 * G_H / (G_E * width_eV)


def test_radiative_times_come_from_eadl_when_the_file_is_there(tmp_path):
    c = tk.Case.load(run_dir(tmp_path, eadl={14: SI, 8: OX}))
    # Si K (designator 1): the I=921 entry, MeV -> eV, t = 1e15 hbar / (e Gamma)   (Dealing_with_EADL.f90:347-354, 671)
    assert c.get("atom:0:0:Radiat") == pytest.approx(t_fs(2.0e-8 * 1e6), rel=1e-14)
    # Si L is given as the whole shell (designator 2), which EADL does not list: the reader averages the sub-shells 3..6 that
    # the first imax-imin+1 = 4 lines of the block hold -- lines 1, 3, 5, 6 -> mean of 3, 5, 6  (:638-654)
    gl = (3.0e-13 + 6.0e-12 + 9.0e-12) / 3 * 1e6
    assert gl >= 1e-6 and c.get("atom:0:1:Radiat") == pytest.approx(t_fs(gl), rel=1e-14)
    # valence band: no radiative decay, no Auger decay (:344-346, 358-360)
    assert c.get("atom:0:2:Radiat") == 1e23 and c.get("atom:0:2:Auger") == 1e23
    assert c.get("atom:1:0:Radiat") == pytest.approx(t_fs(1.0e-9 * 1e6), rel=1e-14)
    # Auger times given by the .cdf are kept (:361)
    assert c.get("atom:0:0:Auger") == 1.6 and c.get("atom:1:0:Auger") == 8.0
    # kinetic energies always come from EADL (the .cdf has no column for them; Ek starts at -1e-15, :333-343);
    # for the valence band the designator after the last core shell is used: after L (2) comes M = 7 -> sub-shells 8..14,
    # of which the 7-line window 1,3,5,6,8,10,<end> holds 8 and 10
    assert c.get("atom:0:0:Ek") == pytest.approx(2.5e-3 * 1e6, rel=1e-14)
    assert c.get("atom:0:1:Ek") == pytest.approx((3.0e-4 + 2.8e-4 + 2.7e-4) / 3 * 1e6, rel=1e-14)
    assert c.get("atom:0:2:Ek") == pytest.approx((4.0e-5 + 2.5e-5) / 2 * 1e6, rel=1e-14)
    assert list(c.warnings) == []


def test_widths_below_1e_6_eV_mean_no_decay_and_missing_reactions_close_the_channel(tmp_path):
    si = (28, dict(SI[1]))
    si[1][921] = [(1, 5.0e-13), (3, 1e-13), (5, 1e-13), (6, 1e-13)]          # K: 5e-7 eV < 1e-6 eV
    ox = (16, {k: v for k, v in OX[1].items() if k != 921})                   # no I=921 block for oxygen
    c = tk.Case.load(run_dir(tmp_path, eadl={14: si, 8: ox}))
    assert c.get("atom:0:0:Radiat") == 1.1e35                                 # :349-350
    assert c.get("atom:1:0:Radiat") == 1.1e35                                 # value 1e-30 for a missing reaction (:604-607) -> no decay


def test_without_photons_no_radiative_time_is_read(tmp_path):
    c = tk.Case.load(run_dir(tmp_path, eadl={14: SI, 8: OX}))
    d = tk.make_run_dir(str(tmp_path / "run_nophot"), ("SiO2_cryst", 79, 2187.0, 0, 10), data_dir=str(tmp_path / "data"))
    c0 = tk.Case.load(d)
    assert c0.get("atom:0:0:Radiat") == 1e23 and c.get("atom:0:0:Radiat") < 1e23


CDF = open(os.path.join(DATA, "INPUT_CDF", "SiO2_cryst.cdf")).read()


def test_missing_cdf_entries_are_filled_from_eadl(tmp_path):
    """Nel <= 0, Ip <= -1e-14 and an Auger time <= 0 on a shell line mean 'take it from EADL' (:327-332, 361-363)."""
    text = CDF.replace("1\t1\t1844.1e0\t2\t1.6e0", "1\t1\t-1.0e15\t0\t-1.0e0")          # Si K: Ip, Nel, Auger left out
    text = text.replace("7\t2\t100.0e0\t8\t16.0e0", "7\t2\t-1.0e15\t0\t0.0e0")          # Si L (whole shell)
    assert text != CDF
    c = tk.Case.load(run_dir(tmp_path, cdf_text=text, eadl={14: SI, 8: OX}))
    assert c.get("atom:0:0:Nel") == 2.0
    assert c.get("atom:0:0:Ip") == pytest.approx(1838.9, rel=1e-14)
    assert c.get("atom:0:0:Auger") == pytest.approx(t_fs(4.0e-7 * 1e6), rel=1e-14)
    # whole L shell: electrons are SUMMED over the sub-shells in the window (:495-505), energies and widths AVERAGED (:638-654)
    assert c.get("atom:0:1:Nel") == 8.0
    assert c.get("atom:0:1:Ip") == pytest.approx((1.5e-4 + 1.0e-4 + 0.99e-4) / 3 * 1e6, rel=1e-14)
    assert c.get("atom:0:1:Auger") == pytest.approx(t_fs((9.0e-7 + 3.0e-9 + 3.0e-9) / 3 * 1e6), rel=1e-14)
    # and the case is usable: tables build from the completed parameters
    c.build_tables(shi_window_only=True)
    assert c.tables.n_ei > 100


def test_missing_cdf_entries_without_eadl_are_reported(tmp_path):
    text = CDF.replace("1\t1\t1844.1e0\t2\t1.6e0", "1\t1\t-1.0e15\t0\t-1.0e0")
    with pytest.raises(RuntimeError, match="EADL2023.ALL"):
        tk.Case.load(run_dir(tmp_path, cdf_text=text, eadl=None))


def test_side_car_is_only_the_fallback(tmp_path):
    """Without EADL2023.ALL the approximate side-car widths are used (and the bench says so); with it they are ignored."""
    c_side = tk.Case.load(run_dir(tmp_path / "a", eadl=None))
    c_eadl = tk.Case.load(run_dir(tmp_path / "b", eadl={14: SI, 8: OX}))
    assert c_side.get("atom:0:0:Radiat") == pytest.approx(t_fs(2.2e-2), rel=1e-12)          # data/INPUT_EADL/radiative_widths.dat
    assert c_eadl.get("atom:0:0:Radiat") == pytest.approx(t_fs(2.0e-2), rel=1e-12)


def test_sub_shell_windows_and_next_designator():
    """select_imin_imax (:687-722) and next_designator (:725-742) through a block that lists every sub-shell up to N."""
    import ctypes as C
    lib = tk.host._host()
    lib.trk3h_eadl_lookup.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double)]
    import tempfile
    subs = [1, 3, 5, 6, 8, 10, 11, 13, 14, 16, 18, 19, 21, 22]
    with tempfile.TemporaryDirectory() as d:
        p = os.path.join(d, "EADL2023.ALL")
        write_eadl(p, {79: (197, {912: [(s, 2.0 * (i + 1)) for i, s in enumerate(subs)],
                                   913: [(s, 1e-3 * (i + 1)) for i, s in enumerate(subs)]})})
        v = C.c_double()

        def look(I, des):
            assert lib.trk3h_eadl_lookup(p.encode(), 79, I, des, C.byref(v)) == 0
            return v.value

        assert look(913, 5) == pytest.approx(3e-3 * 1e6)                      # listed designator: its own entry
        assert look(913, 2) == pytest.approx(np.mean([2e-3, 3e-3, 4e-3]) * 1e6)          # L = 3..6: lines 1,3,5,6
        assert look(913, 4) == 0.0                                             # L23 = 5..6: the 2-line window holds lines 1 and 3 only (:655-657)
        # M = 8..14: window of 7 lines = 1,3,5,6,8,10,11 -> M1..M3 only (the reference's window, kept as it is)
        assert look(913, 7) == pytest.approx(np.mean([5e-3, 6e-3, 7e-3]) * 1e6)
        assert look(912, 7) == 2.0 * (5 + 6 + 7)                               # electrons: summed over the same window
        assert look(912, 1) == 2.0 and look(913, 63) == 1e22                   # valence band: 'infinite' (:662)
        assert lib.trk3h_eadl_lookup(p.encode(), 14, 913, 1, C.byref(v)) != 0  # element not in the file


# ---- files that leave their shells to the atomic database: chemical formula, single-pole CDF, VALENCE / PHONON keywords --------
SP_CDF = """Quartz from the atomic database
SiO2						! chemical formula instead of an element list
2.65	5900.0	5.0	8.9	! density [g/cm^3], speed of sound [m/s], fermi energy [eV], band gap [eV]
"""


def test_chemical_formula_file_is_built_from_the_atomic_database(tmp_path):
    """Decompose_compound (Dealing_with_EADL.f90:28-198) + the all-shells branch of check_atomic_parameters (:374-392) +
    make_valence_band (Reading_files_and_parameters.f90:1989-2137) + the electronic part of get_single_pole
    (Cross_sections.f90:570-633): 'SiO2' -> Si (Z = 14) first, O second; every EADL sub-shell becomes a shell; the outermost ones
    holding the valence electrons (4 of Si, 6 of O: INPUT_atomic_data.dat) are merged into ONE valence band on the first atom;
    every shell gets one oscillator whose k-sum rule gives the electrons of the shell."""
    d = run_dir(tmp_path, cdf_text=SP_CDF, eadl={14: SI, 8: OX}, config="C2")
    case = tk.Case.load(d)
    assert case.get("n_atoms") == 2 and case.get("kind_of_CDF") == 1 and case.get("kind_of_CDF_ph") == 1
    # Si: K, L1, L2, L3 core (M1 + M2 = 2 + 2 valence electrons) + the valence band; O: K core (L1..L3 = 6 valence electrons)
    assert case.get("nshl:0") == 5 and case.get("nshl:1") == 1
    assert case.get("Zat:0") == 14 and case.get("Zat:1") == 8
    n_vb = 4 * 1 + 6 * 2
    assert case.get("atom:0:4:Nel") == pytest.approx(n_vb) and case.get("atom:0:4:Ip") == pytest.approx(8.9)
    assert case.get("atom:0:0:Ip") == pytest.approx(1838.9) and case.get("atom:1:0:Ip") == pytest.approx(538.0)
    assert case.get("atom:0:0:Auger") == pytest.approx(t_fs(0.4), rel=1e-12) and case.get("atom:0:4:Auger") == 1.0e26
    assert case.get("atom:0:0:Radiat") == pytest.approx(t_fs(0.02), rel=1e-12)          # photons on in C2
    # single-pole CDFs: E0 = Ip + 10 eV for the core shells, Gamma = E0, and A from the k-sum rule
    case.build_tables(shi_window_only=True)
    for at, sh, nel in ((0, 0, 2.0), (0, 1, 2.0), (0, 3, 4.0), (1, 0, 2.0 * 2)):   # k-sum per MOLECULE: electrons of the shell x atoms of the kind
        ks, _ = case.sumrules(at, sh)
        assert ks == pytest.approx(nel, rel=1e-9), (at, sh)
    ks, _ = case.sumrules(0, 4)
    assert ks == pytest.approx(n_vb, rel=1e-9)                                      # valence band, per molecule
    assert case.reference_cache_name("el_imfp").startswith("OUTPUT_Electron_IMFPs_Free_spCDF")
    a = case.table_arrays()
    assert a["ei_L"].shape[0] == 6 and np.all(a["ei_L"] > 0)


def test_valence_and_phonon_keywords(tmp_path):
    text = SP_CDF + "VALENCE\n2 63 8.9 16.0 1.0e23\n22.0 200.0 10.0\n35.0 100.0 20.0\nPHONON\n1\n0.12 0.002 0.01\n"
    case = tk.Case.load(run_dir(tmp_path, cdf_text=text, eadl={14: SI, 8: OX}, config="C2"))
    assert case.get("kind_of_CDF") == 1 and case.get("kind_of_CDF_ph") == 0
    assert case.get("phonon_E0") == pytest.approx(0.12)
    case.build_tables(shi_window_only=True)
    ks_user, _ = case.sumrules(0, 4)
    assert ks_user != pytest.approx(16.0, rel=1e-3)                                 # the user's valence CDF is taken as it is
    ks, _ = case.sumrules(0, 0)
    assert ks == pytest.approx(2.0, rel=1e-9)                                       # the core shells are still single-pole


def test_database_files_are_refused_without_the_database(tmp_path):
    with pytest.raises(RuntimeError, match="EADL2023.ALL"):
        tk.Case.load(run_dir(tmp_path, cdf_text=SP_CDF, eadl=None, config="C2"))
