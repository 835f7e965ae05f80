"""BEB shells: a NEGATIVE shell designator in the .cdf makes electrons and valence holes ionise that shell with
the binary-encounter-Bethe cross section (Target_atoms%KOCS = 2, Reading_files_and_parameters.f90:1557-1561) while the ion and the
photons keep its CDF.  Total cross section: Sigma_BEB (Cross_sections.f90:3891-3906, TotIMFP :1041-1047); transferred energy:
bisection on the closed-form cumulative cross section (Electron_NRG_transfer_BEB :2128-2165).  The mean kinetic energy of the
shell comes from EADL2023.ALL (I = 914), which is not part of the reference tree: the tests write ENDL blocks themselves
(tests/test_eadl.py; made-up numbers), so what is pinned is the formula and the sampling, not a material."""
import numpy as np
import pytest

import trekis3_b200 as tk
import emul_api
import oracle_api
from test_eadl import CDF, OX, SI, run_dir

g_Pi, g_a0, g_Ry = 3.1415926535897932384626433832795, 0.5291772085936, 13.6056981


def beb_case(tmp_path, which="core", nmc=10):
    """SiO2_cryst with BEB shells: 'core' = Si L and O K; 'all' = every shell, the valence band included (then the valence holes
    ionise by BEB too).  Au 2187 MeV, photons on."""
    txt = CDF.replace("7\t2\t100.0e0", "7\t-2\t100.0e0").replace("1\t1\t538.25e0", "1\t-1\t538.25e0")
    if which == "all":
        txt = txt.replace("8\t63\t8.9e0", "8\t-63\t8.9e0").replace("1\t1\t1844.1e0", "1\t-1\t1844.1e0")
    assert txt.count("\t-") == (4 if which == "all" else 2)
    return tk.Case.load(run_dir(tmp_path, cdf_text=txt, eadl={14: SI, 8: OX}, config=("SiO2_cryst", 79, 2187.0, 1, nmc)))


def kim_rudd_dsigma_dw(t, u, w, S):
    """Kim & Rudd, Phys. Rev. A 50 (1994) 3954, eq. (57) with dN/dw -> N (BEB): singly differential cross section in
    reduced units (w = W/B, t = T/B, u = U/B), written out independently of the reference's antiderivatives."""
    return S / (t + u + 1.0) * (-(1.0 / (w + 1.0) + 1.0 / (t - w)) / (t + 1.0) + 1.0 / (w + 1.0) ** 2 + 1.0 / (t - w) ** 2
                                + np.log(t) * (1.0 / (w + 1.0) ** 3 + 1.0 / (t - w) ** 3))


def test_a_beb_shell_needs_the_kinetic_energy_from_eadl(tmp_path):
    txt = CDF.replace("7\t2\t100.0e0", "7\t-2\t100.0e0")
    with pytest.raises(RuntimeError, match="EADL2023.ALL"):
        tk.Case.load(run_dir(tmp_path, cdf_text=txt, eadl=None))


def test_total_beb_cross_section_is_the_integral_of_kim_and_rudd(tmp_path):
    c = beb_case(tmp_path)
    assert c.get("atom:0:1:KOCS") == 2 and c.get("atom:1:0:KOCS") == 2 and c.get("atom:0:0:KOCS") == 1 and c.get("atom:0:2:KOCS") == 1
    at_dens = c.get("At_Dens")
    for (at, sh, pers) in ((0, 1, 1.0), (1, 0, 2.0)):
        B, U, N = c.get(f"atom:{at}:{sh}:Ip"), c.get(f"atom:{at}:{sh}:Ek"), c.get(f"atom:{at}:{sh}:Nel")
        assert U > 0
        S = 4.0 * g_Pi * g_a0 * g_a0 * N * (g_Ry / B) ** 2
        for T in (1.5 * B, 5.0 * B, 40.0 * B):
            t, u = T / B, U / B
            w = np.linspace(0.0, (t - 1.0) / 2.0, 200001)                      # ejected-electron energy up to (T - B)/2
            f = kim_rudd_dsigma_dw(t, u, w, S)
            h = w[1] - w[0]
            sigma = h / 3.0 * (f[0] + f[-1] + 4.0 * f[1:-1:2].sum() + 2.0 * f[2:-1:2].sum())          # Simpson
            L, dEdx = c.eval_TotIMFP(T, at, sh, 0)
            temp1 = at_dens * 1e-24 * pers / 3.0
            assert 1.0 / (temp1 * L) == pytest.approx(sigma, rel=1e-9), (at, sh, T)
            # stopping power: the reference integrates w dsigma/dw up to (T - 1 eV)/2 (:1045), in eV
            w2 = np.linspace(0.0, (T - 1.0) / 2.0 / B, 200001)
            f2 = w2 * kim_rudd_dsigma_dw(t, u, w2, S)
            h2 = w2[1] - w2[0]
            s2 = B * h2 / 3.0 * (f2[0] + f2[-1] + 4.0 * f2[1:-1:2].sum() + 2.0 * f2[2:-1:2].sum())
            assert dEdx == pytest.approx(temp1 * s2, rel=1e-8), (at, sh, T)
        # below the binding energy: Sigma_BEB = 0 -> the mean free path is 1/0 (:3898-3899, :1044)
        L0, d0 = c.eval_TotIMFP(0.5 * B, at, sh, 0)
        assert np.isinf(L0)


def test_beb_tables_have_no_differential_rows_and_the_ion_keeps_the_cdf(tmp_path):
    c = beb_case(tmp_path)
    c.build_tables(shi_window_only=True)
    a, t = c.table_arrays(), c.tables
    off = a["eid_off"]
    for flat, beb in ((0, False), (1, True), (2, False), (3, True)):
        n = off[(flat + 1) * t.n_ei] - off[flat * t.n_ei]
        assert (n == 0) == beb, flat
        assert t.shell_kocs[flat] == (2 if beb else 1)
    assert t.shell_Ek[1] > 0 and t.at_dens == c.get("At_Dens")
    # table entries = single-point evaluations; the ion's tables of a BEB shell are the CDF ones (KOCS_SHI = 1)
    i = 300
    assert a["ei_L"][1, i] == c.eval_TotIMFP(float(a["ei_E"][i]), 0, 1, 0)[0]
    ref = tk.Case.load(run_dir(_sibling(c), eadl={14: SI, 8: OX}, config=("SiO2_cryst", 79, 2187.0, 1, 10)))
    ref.build_tables(shi_window_only=True)
    b = ref.table_arrays()
    assert np.array_equal(a["shi_L"], b["shi_L"]) and np.array_equal(a["ph_L"], b["ph_L"]) and np.array_equal(a["dshi_L"], b["dshi_L"])
    assert not np.array_equal(a["ei_L"][1], b["ei_L"][1]) and np.array_equal(a["ei_L"][0], b["ei_L"][0])


def _sibling(case):
    import pathlib
    p = pathlib.Path(case.dir).parent / "plain"
    p.mkdir()
    return p


@pytest.mark.parametrize("which", ["core", "all"])
def test_beb_monte_carlo_oracle_equals_device_code(tmp_path, which):
    """Same Philox streams: the time-ordered oracle and the device code (CUDA physics header compiled for the CPU) give the same
    events and tallies.  With BEB on the valence band the reference's clamp of the transferred energy to the kinematic maximum of a
    hole (Cross_sections.f90:1849-1850) can put it below the ionisation potential; the reference prints its error 10/40 for these
    events (Monte_Carlo.f90:1705, 2632) and goes on -- so do both sides, with equal counts."""
    c = beb_case(tmp_path, which)
    c.build_tables(shi_window_only=True)
    to, so, eo, no = oracle_api.run(c, 0, 4, rng_mode=1)
    te, se, ee, ne = emul_api.run(c, 0, 4, batch=3)
    assert so["events"] == se["events"] and so["errors"] == se["errors"]
    assert so["events"]["el_inelastic"] > 3000 and so["events"]["auger"] > 50
    if which == "core":
        assert not so["errors"]
        drift = np.abs(eo[:, 1:] - eo[:, -1:]) / eo[:, -1:]
        assert drift.max() < 1e-9
    else:
        assert set(so["errors"]) <= {"err10", "err40"}
    assert np.allclose(to, te, rtol=1e-9, atol=1e-300)
    assert np.array_equal(no, ne) and np.allclose(eo, ee, rtol=1e-12)
    # the BEB shells are actually ionised: the plain-CDF material gives other histories from the same streams
    plain = tk.Case.load(run_dir(_sibling(c), eadl={14: SI, 8: OX}, config=("SiO2_cryst", 79, 2187.0, 1, 10)))
    plain.build_tables(shi_window_only=True, cache_dir=tk._abi.REPO + "/.table_cache")
    tp, sp, _, _ = oracle_api.run(plain, 0, 4, rng_mode=1)
    assert sp["events"] != so["events"]
