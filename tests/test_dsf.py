"""DSF elastic scattering (kind_of_EMFP = 2): the energy an electron or a valence hole exchanges with the lattice
is sampled from tabulated dynamic-structure-factor cross sections -- arguments DSF_DEMFP / DSF_DEMFP_H of do_Monte_Carlo
(Monte_Carlo.f90:44, 2387-2389, 2668-2675; NRG_transfer_elastic_DSF, Cross_sections.f90:3652-3780; reader
reading_DSF_cross_sections, Reading_files_and_parameters.f90:2516-2678).  A collision can GIVE energy to the particle (absorption,
negative transfer).  The reference ships no INPUT_DSF file, so the tests write one in the layout its reader walks; the numbers are
made up (two phonon-like peaks at +-30 meV with a detailed-balance-like asymmetry), what is pinned is the reader, the tables and
the sampling."""
import os
import shutil

import numpy as np
import pytest

import trekis3_b200 as tk
import emul_api
import oracle_api

DATA = os.path.join(tk._abi.REPO, "data")
E_GRID = [0.05, 0.2, 1.0, 5.0, 30.0, 200.0, 2000.0, 2.0e4, 2.0e5]
NW = 61


def dsf_value(E, w, hole):
    """differential inverse mean free path [1/(A eV)] of the synthetic file"""
    amp = (0.8 if hole else 0.5) / (1.0 + E / 300.0)
    emit = 1.0 if E > 0.2 else 0.0          # a particle below the phonon energy cannot emit one (it would end with negative energy)
    return amp * (emit * np.exp(-((w - 0.03) / 0.012) ** 2) + 0.4 * np.exp(-((w + 0.03) / 0.012) ** 2))


def write_dsf(path, hole):
    w = np.linspace(0.25, -0.25, NW)                    # descending, as the reader expects (it reverses them, :2605-2610)
    with open(path, "w") as f:
        f.write("%d\n" % NW)
        for E in E_GRID:
            for x in w:
                f.write("%.6e %.6e %.6e\n" % (E, x, dsf_value(E, x, hole)))


def dsf_run_dir(tmp_path, cfg="C1", material="Al2O3", holes_file=True):
    dd = tmp_path / "data"
    os.makedirs(dd / "INPUT_DSF" / material)
    shutil.copy(os.path.join(DATA, "INPUT_PARAMETERS.default.txt"), dd)
    for d in ("INPUT_CDF", "INPUT_DOS", "INPUT_EADL"):
        os.symlink(os.path.join(DATA, d), dd / d)
    write_dsf(dd / "INPUT_DSF" / material / f"{material}_Electron_DSF_Differential_EMFPs_0K.dat", False)
    if holes_file:
        write_dsf(dd / "INPUT_DSF" / material / f"{material}_Hole_DSF_Differential_EMFPs_0K.dat", True)
    return tk.make_run_dir(str(tmp_path / "run"), cfg, data_dir=str(dd), edits={12: "2   1   ! DSF elastic scattering"})


def dsf_case(tmp_path, **kw):
    c = tk.Case.load(dsf_run_dir(tmp_path, **kw))
    c.build_tables(shi_window_only=True)
    return c


def test_reader_resamples_and_integrates_the_rows(tmp_path):
    """reading_DSF_cross_sections: rows reversed to ascending transfer, resampled on NTEPo points of (-0.2, 0.2] eV by linear
    interpolation and integrated: total, emission-only (dE >= 0) and absorption-only inverse mean free paths -- restated in numpy."""
    c = dsf_case(tmp_path)
    t, a = c.tables, c.table_arrays()
    assert c.config.kind_of_EMFP == 2 and t.n_dsf_e == NW and t.n_dsf_h == NW and t.n_ee == len(E_GRID) and t.n_he == len(E_GRID)
    assert list(a["ee_E"]) == E_GRID and a["eed_off"][-1] == 0
    rd = np.vectorize(lambda v: float("%.6e" % v))                          # the file holds 7 significant digits
    w_file = rd(np.linspace(0.25, -0.25, NW))[::-1]
    # the last data line of the file is never read (`do i = 2, M` over N = M + 1 lines, :2563): its slot stays zero -- that is the
    # point (last energy, lowest transfer); the rows below are all but the last one
    for i, E in enumerate(E_GRID[:-1]):
        y = rd(dsf_value(E, np.linspace(0.25, -0.25, NW), False))[::-1]
        y[y < 1e-10] = 0.0
        dE = 0.4 / NW
        grid = -0.2 + dE * np.arange(1, NW + 1)
        loc = np.interp(grid, w_file, y)
        tot = np.cumsum(loc * dE)
        emit = np.cumsum(np.where(grid >= 0.0, loc * dE, 0.0))
        row_dE = np.array([t.dsf_e_dE[i * NW + j] for j in range(NW)])
        row_em = np.array([t.dsf_e_emit[i * NW + j] for j in range(NW)])
        row_ab = np.array([t.dsf_e_absorb[i * NW + j] for j in range(NW)])
        assert np.allclose(row_dE, grid, rtol=0, atol=1e-15)
        want_em = np.where(np.abs(emit) > 1e-10, 1.0 / np.where(emit == 0, 1, emit), 1e30)
        assert np.allclose(row_em, want_em, rtol=1e-12)
        neg = grid < 0.0
        want_ab = np.where(np.abs(tot) > 1e-10, 1.0 / np.where(tot == 0, 1, tot), 1e30)
        assert np.allclose(row_ab[neg], want_ab[neg], rtol=1e-12)
        assert np.all(row_ab[~neg] == row_ab[neg][-1])                    # "absorption does not change" once dE >= 0
        # Elastic_MFP%Total / %Emit / %Absorb = the integrals over all transfers (Analytical_IMFPs.f90:913-919)
        assert a["ee_L"][i] == pytest.approx(1.0 / tot[-1], rel=1e-12)
        assert t.ee_emit[i] == pytest.approx(want_em[-1], rel=1e-12) and t.ee_absorb[i] == row_ab[-1]
        assert 1.0 / a["ee_L"][i] == pytest.approx(1.0 / t.ee_emit[i] + 1.0 / t.ee_absorb[i], rel=1e-12)


def test_a_missing_dsf_file_switches_to_mott_as_the_reference_does(tmp_path):
    c = tk.Case.load(dsf_run_dir(tmp_path, holes_file=False))
    c.build_tables(shi_window_only=True)
    assert c.config.kind_of_EMFP == 0                                         # Reading_files_and_parameters.f90:2544-2552
    assert any("Mott" in w and "Hole_DSF" in w for w in c.warnings)


@pytest.mark.parametrize("cfg,material", [("C1", "Al2O3"), ("C3", "Diamond")])
def test_dsf_monte_carlo_oracle_equals_device_code(tmp_path, cfg, material):
    """Same Philox streams: the oracle (which interpolates the whole row, as the reference does) and the device code (which
    evaluates the interpolated row where the bisection probes it) give the same events and tallies; the lattice receives a NET
    positive energy although single collisions take energy from it; total energy is conserved."""
    c = dsf_case(tmp_path, cfg=cfg, material=material)
    to, so, eo, no = oracle_api.run(c, 0, 3, rng_mode=1)
    te, se, ee, ne = emul_api.run(c, 0, 3, batch=2)
    assert so["events"] == se["events"] and not so["errors"] and not se["errors"]
    assert so["events"]["el_elastic"] > 1000 and so["events"]["vbh_elastic"] > 1000
    assert np.allclose(to, te, rtol=1e-9, atol=1e-300)
    assert np.array_equal(no, ne) and np.allclose(eo, ee, rtol=1e-12)
    drift = np.abs(eo[:, 1:] - eo[:, -1:]) / eo[:, -1:]
    assert drift.max() < 1e-9
    from trekis3_b200.host import split_tallies
    T = split_tallies(c.layout(), to)
    assert T["Out_E_at"][-1] > 0.0
