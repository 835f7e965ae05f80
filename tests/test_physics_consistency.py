"""The consistency checks the reference's manual asks for before production runs (!READ_ME_TREKIS_3, section VI "Consistency checks
and known problems": electron inelastic mean free paths against the NIST database, the ion's energy loss against SRIM / Bethe),
applied to the tables this repository builds from the reference's shipped INPUT_CDF / INPUT_DOS files.

This is NOT parity with the reference program (that stays unpinned: no Fortran compiler, oracle/README.md); it anchors the ABSOLUTE
scale of the tables (units, densities, constants, the k- and f-sum-rule normalisation of the loss function) on numbers that do not
come from this repository:
  * the ion's electronic stopping power against the relativistic Bethe formula with the ICRU-49 mean excitation energies and the
    Barkas effective charge (closed form, evaluated here);
  * the Bethe asymptote of the electron mean free path (lambda ~ E / ln E);
  * electron inelastic mean free paths at 1 keV against the optical-data values of Tanuma, Powell and Penn (NIST SRD 71).  The
    literature numbers below are quoted from the published tables from memory (this container has no network access), hence the
    wide tolerance; they are a gate against factor-of-two errors, not a precision claim."""
import math

import numpy as np
import pytest

import trekis3_b200 as tk

CACHE = tk._abi.REPO + "/.table_cache"

# mean number of electrons per atom, ICRU-49 mean excitation energy [eV], lambda_inelastic(1 keV) [A] from optical data (TPP / NIST SRD 71)
MATERIAL = {"C1": (10.0, 145.2, 20.0), "C2": (10.0, 139.2, 27.0), "C3": (6.0, 81.0, 18.0), "C4": (79.0, 790.0, 13.2)}
ION = {54: 131.293, 79: 196.967, 92: 238.029}           # atomic masses [u]


def _tables(case):
    t = case.tables
    NS = t.n_shells
    E = np.ctypeslib.as_array(t.ei_E, (t.n_ei,))
    L = np.ctypeslib.as_array(t.ei_L, (NS * t.n_ei,)).reshape(NS, t.n_ei)
    with np.errstate(divide="ignore"):
        lam = 1.0 / np.sum(np.where(L > 1e-10, 1.0 / L, 0.0), axis=0)      # How_many_electrons' total, Monte_Carlo.f90:1991-2011
    Es = np.ctypeslib.as_array(t.shi_E, (t.n_shi,))
    S = np.ctypeslib.as_array(t.shi_dEdx, (NS * t.n_shi,)).reshape(NS, t.n_shi).sum(0)
    return E, lam, Es, S


def _case(request, cfg):
    return request.getfixturevalue({"C1": "case_c1", "C2": "case_c2", "C3": "case_c3", "C4": "case_c4"}[cfg])


@pytest.mark.parametrize("cfg", ["C1", "C2", "C3", "C4"])
def test_ion_stopping_power_against_bethe(request, cfg):
    case = _case(request, cfg)
    _, _, Es, S = _tables(case)
    z_ion = tk.CONFIGS[cfg][1]
    e_ion = case.get("shi_E")                             # [eV]
    se = float(np.interp(e_ion, Es, S))                   # [eV/A], summed over the shells
    zpa, i_mean, _ = MATERIAL[cfg]
    n_e = case.get("At_Dens") * 1e-24 * zpa               # electrons per A^3
    gamma = 1.0 + e_ion / (ION[z_ion] * 931.494e6)
    beta2 = 1.0 - 1.0 / gamma ** 2
    zeff = z_ion * (1.0 - math.exp(-125.0 * math.sqrt(beta2) * z_ion ** (-2.0 / 3.0)))      # Barkas
    mc2, e2 = 510998.95, 14.399645                        # [eV], [eV A]
    bethe = 4.0 * math.pi * n_e * zeff ** 2 * e2 ** 2 / (mc2 * beta2) * (math.log(2.0 * mc2 * beta2 * gamma ** 2 / i_mean) - beta2)
    assert 0.85 < se / bethe < 1.15, (cfg, se, bethe)


@pytest.mark.parametrize("cfg", ["C1", "C2", "C3", "C4"])
def test_electron_inelastic_mean_free_path(request, cfg):
    case = _case(request, cfg)
    E, lam, _, _ = _tables(case)
    at = lambda e: float(np.exp(np.interp(math.log(e), np.log(E), np.log(np.minimum(lam, 1e30)))))
    assert abs(at(1000.0) / MATERIAL[cfg][2] - 1.0) < 0.25, (cfg, at(1000.0))
    # the minimum of the "universal curve" lies between 30 and 150 eV at a few Angstrom
    sel = (E > 15.0) & (E < 3000.0)
    e_min = E[sel][np.argmin(lam[sel])]
    assert 25.0 < e_min < 200.0 and 2.0 < lam[sel].min() < 9.0, (cfg, e_min, lam[sel].min())
    # Bethe asymptote lambda ~ E / ln(c E): a decade in energy costs a factor 10 ln(c 1 keV) / ln(c 10 keV), between 5.5 and 8
    assert 5.5 < at(1.0e4) / at(1.0e3) < 8.0, (cfg, at(1.0e4) / at(1.0e3))
    # monotone above the minimum
    hi = lam[(E > 2.0 * e_min) & (E < 3.0e4)]
    assert np.all(np.diff(hi) > 0.0)
