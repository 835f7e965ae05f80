"""Save_output (Sorting_output_data.f90:340-1140): directory name, file set, headers and number formats of the
reference's output tree, written by the host library from the packed tally buffer."""
import os
import re

import numpy as np

import trekis3_b200 as tk
from trekis3_b200.host import split_tallies
import emul_api

E_FIELD = re.compile(r"^ +-?0\.\d{16}E[+-]\d{2}$")            # '(e)' of real(8): E25.16 (csrc/host/fortran_fmt.hpp)


def read_table(path, skip=1):
    rows = [l.split() for l in open(path).read().splitlines()[skip:]]
    return np.array([[float(x) for x in r] for r in rows])


def test_output_tree_matches_the_reference_layout(tmp_path):
    case = tk.Case.load(tk.make_run_dir(str(tmp_path / "c2"), "C2"))
    case.build_tables(shi_window_only=True, cache_dir=tk._abi.REPO + "/.table_cache")
    nmc = 2
    t, s, _, _ = emul_api.run(case, 0, nmc, batch=2)
    d = case.save_output(t, nmc, str(tmp_path / "out"))
    # <root>/OUTPUT_<material>/OUTPUT_<ion>_in_<material>/<ion>_E_<f8.2>_MeV_<f10.2>_fs  (:443-447)
    assert d.endswith("OUTPUT_SiO2_cryst/OUTPUT_Au_in_SiO2_cryst/Au_E_2187.00_MeV_100.00_fs")
    files = set(os.listdir(d))
    expected = {"INPUT_PARAMETERS.txt", "!Parameters.txt", "Total_numbers.txt", "Total_energies.txt",
                "Hole_mean_diffusion_coefficient.txt", "Electrons_theta_distribution.txt", "VB_holes_theta_distribution.txt",
                "Radial_electron_density[1_cm^-3].txt", "Radial_electron_energy[eV_A^-3].txt", "Radial_electron_temperature[K].txt",
                "Radial_photon_density[1_cm^-3].txt", "Radial_photon_energy[eV_A^-3].txt",
                "Electron_distribution_vs_E[1_eV].txt", "VB_holes_distribution_vs_E[1_eV].txt",
                "Radial_Lattice_energy[eV_A^-3].txt", "Radial_Track_energy[eV_A^-3].txt", "Radial_Lattice_temperature[K].txt",
                "Radial_Si_K-shell_holes_density[1_cm^-3].txt", "Radial_Si_K-shell_holes_energy[eV_A^-3].txt",
                "Radial_Si_L-shell_holes_density[1_cm^-3].txt", "Radial_Si_L-shell_holes_energy[eV_A^-3].txt",
                "Radial_O_K-shell_holes_density[1_cm^-3].txt", "Radial_O_K-shell_holes_energy[eV_A^-3].txt",
                "Radial_Valence_holes_density[1_cm^-3].txt", "Radial_Valence_holes_pot_energy[eV_A^-3].txt",
                "Radial_Valence_holes_kin_energy[eV_A^-3].txt", "Radial_Valence_holes_temperature[K].txt"}
    assert files == expected
    lay = case.layout()
    T = split_tallies(lay, t)
    # Total_numbers.txt: header + Nt rows of 6 fields of 25 characters
    lines = open(os.path.join(d, "Total_numbers.txt")).read().splitlines()
    assert lines[0] == "#Time[fs]    Ne    Ne_Emitted    Energy[eV]     Energy_Emitted[eV] N_photons"
    assert len(lines) == 1 + lay.Nt and all(len(l) == 6 * 25 for l in lines[1:])
    assert all(E_FIELD.match(lines[1][i:i + 25]) for i in range(0, 6 * 25, 25))
    tot = read_table(os.path.join(d, "Total_numbers.txt"))
    assert np.allclose(tot[:, 0], [0.01, 0.1, 1.0, 10.0, 100.0])
    assert np.allclose(tot[:, 1], T["Out_tot_Ne"] / nmc, rtol=1e-14) and np.allclose(tot[:, 3], T["Out_tot_E"] / nmc, rtol=1e-14)
    # radial files: header '#Radius[A] ' + f10.2 times, rows f9.1 + Nt numbers
    head, *rows = open(os.path.join(d, "Radial_electron_density[1_cm^-3].txt")).read().splitlines()
    assert head.startswith("#Radius[A]       0.01[fs]   ") and len(rows) == 50
    assert rows[0][:9] == "      1.0" and len(rows[0]) == 9 + 25 * lay.Nt + 1
    ne = read_table(os.path.join(d, "Radial_electron_density[1_cm^-3].txt"))
    assert np.allclose(ne[:, 1:], (T["Out_ne"] / nmc * 1e24).T, rtol=1e-14)
    # the lattice file shows the running sum over the time intervals (:931)
    lat = read_table(os.path.join(d, "Radial_Lattice_energy[eV_A^-3].txt"))
    assert np.allclose(lat[:, 1:], np.cumsum(T["Out_Elat"] / nmc, axis=0).T, rtol=1e-13, atol=1e-300)
    # track energy = lattice + potential and kinetic energy of the valence holes (:960); VB = last shell of atom 1
    trk = read_table(os.path.join(d, "Radial_Track_energy[eV_A^-3].txt"))
    vb = (T["Out_Eh"][:, :, 0, 2] + T["Out_Ehkin"][:, :, 0, 2]) / nmc
    assert np.allclose(trk[:, 1:], (np.cumsum(T["Out_Elat"] / nmc, axis=0) + vb).T, rtol=1e-13, atol=1e-300)
    # spectra: the electron spectrum is tabulated on the radius grid (sic, :1024), the hole spectrum on the DOS grid
    sp = read_table(os.path.join(d, "Electron_distribution_vs_E[1_eV].txt"), skip=2)
    assert sp.shape == (50, 1 + lay.Nt) and np.allclose(sp[:, 1:], (T["Out_Ee_vs_E"] / nmc).T, rtol=1e-14)
    sph = read_table(os.path.join(d, "VB_holes_distribution_vs_E[1_eV].txt"))
    assert sph.shape == (lay.n_dos, 1 + lay.Nt)
    th = read_table(os.path.join(d, "Electrons_theta_distribution.txt"))
    assert th.shape == (180, 1 + lay.Nt) and th[0, 0] == 1.0 and th[-1, 0] == 180.0
    # energy columns of Total_energies.txt add up to the conserved total
    en = read_table(os.path.join(d, "Total_energies.txt"))
    assert en.shape == (lay.Nt, 5 + 4)
    assert np.allclose(en[:, 1] + en[:, 2] + en[:, 4] + en[:, 5:].sum(axis=1), tot[:, 3], rtol=1e-12)
    assert open(os.path.join(d, "INPUT_PARAMETERS.txt")).read().split()[0] == "SiO2_cryst"
    # a second save does not overwrite: the reference appends _1, _2, ... (:451-465)
    d2 = case.save_output(t, nmc, str(tmp_path / "out"))
    assert d2 == d + "_1" and set(os.listdir(d2)) == expected


def test_electron_only_case_has_no_photon_files(tmp_path, case_c1):
    t, s, _, _ = emul_api.run(case_c1, 0, 1, batch=1)
    d = case_c1.save_output(t, 1, str(tmp_path / "o"))
    files = set(os.listdir(d))
    assert d.endswith("OUTPUT_Al2O3/OUTPUT_Xe_in_Al2O3/Xe_E_167.00_MeV_100.00_fs")
    assert "Radial_photon_density[1_cm^-3].txt" not in files and "Radial_Al_L-shell_holes_energy[eV_A^-3].txt" in files
    assert open(os.path.join(d, "Total_numbers.txt")).readline().strip() == "#Time[fs]    Ne    Ne_Emitted    Energy[eV]     Energy_Emitted[eV]"
