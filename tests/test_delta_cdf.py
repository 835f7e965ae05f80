"""Delta-function CDF (kind_of_DR = 4): the inelastic cross section of electrons and valence holes in closed
form per CDF oscillator (Integral_CDF_delta_CS, Cross_sections.f90:1449-1522; TotIMFP :952-956) and the transferred energy sampled
from it by bisection with a random number of its own (get_inelastic_energy_transfer :2051-2123, called from
Electron_NRG_transfer_CDF :1894-1895).  Elastic scattering, the ion and the photons keep the Ritchie CDF with the free-electron
dispersion (select case default, :377-392).

Pinned by: the weights alpha against a numerical k-sum integral of the Ritchie oscillator; a numpy restatement of the cross section
written from the Fortran (a third copy beside the shared host/device header and the oracle's); continuity at the point where the
reference switches from its linear extrapolation to the delta model; oracle == device code on the same Philox streams."""
import math

import numpy as np
import pytest

import trekis3_b200 as tk
import emul_api
import oracle_api

g_e, g_me, g_cvel, g_Pi, g_a0 = 1.602176487e-19, 9.1093821545e-31, 299792458.0, 3.1415926535897932384626433832795, 0.5291772085936
g_me_eV = 0.51099906e6
EDIT = {13: "4   0   ! delta-function CDF"}


def delta_case(tmp_path, cfg="C1"):
    c = tk.Case.load(tk.make_run_dir(str(tmp_path / ("d" + cfg)), cfg, edits=EDIT))
    c.build_tables(shi_window_only=True)
    return c


def oscillators(c, flat):
    t = c.tables
    lo, hi = t.osc_off[flat], t.osc_off[flat + 1]
    return [t.osc_E0[i] for i in range(lo, hi)], [t.osc_alpha[i] for i in range(lo, hi)]


def np_sigma(E, E0s, alphas, Ip, nat, Emax_in=None):
    """Integral_CDF_delta_CS for M = mt = m_e, identical particles (the only case the reference reaches), written from the Fortran."""
    Mc2 = mtc2 = g_me * g_cvel * g_cvel / g_e

    def integral_CS(alpha, E0, W):
        return alpha / (g_me_eV * (2 * Mc2 - E0)) * ((2 * Mc2 - mtc2) * math.log(2 * Mc2 + W - E0) + 2 * Mc2 * mtc2 / E0 * (math.log(W) - math.log(abs(W - E0))))

    def prefactor(Ek):
        fact = Ek / Mc2 + 1.0
        beta2 = 1.0 - 1.0 / (fact * fact)
        return 1.0e24 / (g_Pi * g_a0 * nat * g_me_eV * beta2)

    CS, P = 0.0, 0.0
    for E0, alpha in zip(E0s, alphas):
        Wmin = max(Ip, E0 * (1.0 - 0.25 * E0 / E))
        Eeq = 1.25 * E0 - Ip / 2.0 + 0.25 * math.sqrt(17.0 * E0 * E0 - 12.0 * Ip * E0 + 4.0 * Ip * Ip)
        if E <= Ip:
            CS, P = 0.0, 0.0
        elif E <= Eeq * 1.01:
            Ex = Eeq + Eeq / 100.0
            wl, wh = max(Ip, E0 * (1.0 - 0.25 * E0 / Ex)), (Ex + Ip) * 0.5
            cs = -prefactor(Ex) * (integral_CS(alpha, E0, wh) - integral_CS(alpha, E0, wl))
            CS, P = cs / (Ex - Ip) * E - cs * Ip / (Ex - Ip), 1.0
        else:
            Wmax = (E + Ip) * 0.5
            if Emax_in is not None:
                Wmax = Wmin if Emax_in < Wmin else min(Wmax, Emax_in)

            def part(W):
                return 0.0 if (W < Wmin or E <= Ip) else integral_CS(alpha, E0, W)
            CS, P = CS - (part(Wmax) - part(Wmin)), prefactor(E)
    return abs(CS) * P


def test_alpha_is_the_k_sum_of_the_oscillator_above_the_threshold(tmp_path):
    """define_alpha (Reading_files_and_parameters.f90:2199-2206) = integral of x Im(-1/eps) of ONE Ritchie oscillator from Ip to
    infinity, here by numerical quadrature of A Gamma x^2 / ((x^2 - E0^2)^2 + (Gamma x)^2)."""
    c = tk.Case.load(tk.make_run_dir(str(tmp_path / "a"), "C1", edits=EDIT))
    c.build_tables(shi_window_only=True)
    t = c.tables
    assert t.delta_cdf == 1 and t.osc_off[t.n_shells] >= t.n_shells
    txt = open(tk._abi.REPO + "/data/INPUT_CDF/Al2O3.cdf").read().splitlines()
    # Al K shell of Al2O3.cdf: one oscillator; find its (E0, A, Gamma) line right after the first shell line
    i0 = next(i for i, l in enumerate(txt) if "number of shells" in l) + 1
    ncdf = int(txt[i0].split()[0]); Ip = float(txt[i0].split()[2].replace("d", "e").replace("D", "e"))
    E0, A, G = (float(v.replace("d", "e").replace("D", "e")) for v in txt[i0 + 1].split()[:3])
    E0s, alphas = oscillators(c, 0)
    assert len(E0s) == ncdf and E0s[0] == E0
    x = np.exp(np.linspace(math.log(Ip), math.log(1.0e9), 2_000_001))          # log grid: the integrand falls like 1/x^2
    f = A * G * x * x / ((x * x - E0 * E0) ** 2 + (G * x) ** 2) * x              # dx = x dlnx
    h = math.log(x[1] / x[0])
    integral = h / 3.0 * (f[0] + f[-1] + 4.0 * f[1:-1:2].sum() + 2.0 * f[2:-1:2].sum())
    integral += A * G / 1.0e9                                                    # tail beyond 1e9 eV: A G / x
    assert alphas[0] == pytest.approx(integral, rel=1e-7)


def test_tables_equal_the_numpy_restatement_and_have_no_differential_rows(tmp_path):
    c = delta_case(tmp_path)
    a, t = c.table_arrays(), c.tables
    nat = c.get("At_Dens")
    assert t.at_dens == nat
    off = a["eid_off"]
    assert off[-1] == 0 and a["hid_off"][-1] == 0                      # no differential tables at all for electrons and holes
    for flat in range(t.n_shells):
        E0s, alphas = oscillators(c, flat)
        Ip = max(a["shell_Ip"][flat], 1.0e-3)
        for i in range(5, t.n_ei, 23):
            E = float(a["ei_E"][i])
            s = np_sigma(E, E0s, alphas, Ip, nat)
            want = 1.0 / (s * nat * 1.0e-24) if s > 1.0e-24 else 1.0e30    # MFP_from_sigma :1417-1428
            assert a["ei_L"][flat, i] == pytest.approx(want, rel=1e-12), (flat, E)
    # valence holes: the same closed form with the free-electron mass (TotIMFP :953, kind_of_particle = 'Hole')
    vb = t.vb_shell
    E0s, alphas = oscillators(c, vb)
    for i in range(5, t.n_hi, 31):
        E = float(a["hi_E"][i])
        s = np_sigma(E, E0s, alphas, max(a["shell_Ip"][vb], 1.0e-3), nat)
        want = 1.0 / (s * nat * 1.0e-24) if s > 1.0e-24 else 1.0e30
        assert a["hi_L"].reshape(-1, t.n_hi)[vb, i] == pytest.approx(want, rel=1e-12), E
    # everything else keeps the Ritchie CDF with the free-electron dispersion: identical to kind_of_DR = 1
    plain = tk.Case.load(tk.make_run_dir(str(tmp_path / "p"), "C1"))
    plain.build_tables(shi_window_only=True, cache_dir=tk._abi.REPO + "/.table_cache")
    b = plain.table_arrays()
    for k in ("shi_L", "dshi_L", "ee_L", "eed_L", "he_L", "hed_L"):
        assert np.array_equal(a[k], b[k]), k
    assert c.reference_cache_name("el_imfp").startswith("OUTPUT_Electron_IMFPs_Delta_")      # Analytical_IMFPs.f90:276-278


def test_cross_section_is_continuous_where_the_linear_extrapolation_hands_over(tmp_path):
    """One oscillator (Al K shell): below 1.01 Eeq the reference uses a straight line from (Ip, 0) to the delta model's value at
    1.01 Eeq (Find_linear_a_b :1671-1688), so the cross section must be continuous there and vanish at the threshold."""
    c = delta_case(tmp_path)
    a = c.table_arrays()
    E0s, alphas = oscillators(c, 0)
    assert len(E0s) == 1
    Ip, nat = a["shell_Ip"][0], c.get("At_Dens")
    Eeq = 1.25 * E0s[0] - Ip / 2.0 + 0.25 * math.sqrt(17.0 * E0s[0] ** 2 - 12.0 * Ip * E0s[0] + 4.0 * Ip * Ip)
    Ex = 1.01 * Eeq

    def sigma(E):
        L, _ = c.eval_TotIMFP(E, 0, 0, 0)
        return 0.0 if L >= 1.0e30 else 1.0 / (L * nat * 1.0e-24)
    assert sigma(Ex * (1 - 1e-9)) == pytest.approx(sigma(Ex * (1 + 1e-9)), rel=1e-6)
    assert sigma(Ip * (1 + 1e-6)) < 1e-5 * sigma(Ex) and sigma(0.9 * Ip) == 0.0
    assert sigma(0.5 * (Ip + Ex)) == pytest.approx(0.5 * (sigma(Ip * (1 + 1e-12)) + sigma(Ex * (1 - 1e-12))), rel=1e-6)   # a straight line


@pytest.mark.parametrize("cfg", ["C1", "C3"])
def test_delta_cdf_monte_carlo_oracle_equals_device_code(tmp_path, cfg):
    """Same Philox streams: the time-ordered oracle (its own restatement of the closed forms) and the device code (shared header)
    give the same events and tallies; energy is conserved; the histories differ from the Ritchie-CDF ones."""
    c = delta_case(tmp_path, cfg)
    to, so, eo, no = oracle_api.run(c, 0, 3, rng_mode=1)
    te, se, ee, ne = emul_api.run(c, 0, 3, batch=2)
    assert so["events"] == se["events"] and not so["errors"] and not se["errors"]
    assert so["events"]["el_inelastic"] > 1000
    if cfg == "C3":
        assert so["events"]["vbh_inelastic"] > 50                       # valence holes ionise through the same closed form
    assert np.allclose(to, te, rtol=1e-9, atol=1e-300)
    assert np.array_equal(no, ne) and np.allclose(eo, ee, rtol=1e-12)
    drift = np.abs(eo[:, 1:] - eo[:, -1:]) / eo[:, -1:]
    assert drift.max() < 1e-9
    plain = tk.Case.load(tk.make_run_dir(str(tmp_path / "p"), cfg))
    plain.build_tables(shi_window_only=True, cache_dir=tk._abi.REPO + "/.table_cache")
    _, sp, _, _ = oracle_api.run(plain, 0, 3, rng_mode=1)
    assert sp["events"] != so["events"]
