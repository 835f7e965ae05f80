// output.cpp -- writer of the reference's output directory (Save_output, Sorting_output_data.f90:340-1140).
//
// Same directory name, file names, headers, column layout and number formats as the reference, so that existing
// post-processing and gnuplot scripts keep working:
//   <out_root>/OUTPUT_<material>/OUTPUT_<ion>_in_<material>/<ion>_E_<f8.2>_MeV_<f10.2>_fs[_n]/
// Numbers written with the width-less descriptor '(e)' use the DEC/Intel default for real(8), E25.16 (fortran_fmt.hpp:
// pinned by files shipped with the reference; gfortran needs -fdec-format-defaults).
// '!Parameters.txt' carries the same information as print_parameters (:42-338) but is not byte-identical:
// the title banner, the sum-rule table and the wall-clock duration are the reference program's own.
#include <sys/stat.h>
#include <cstdio>
#include <cstring>
#include <fstream>
#include "fortran_fmt.hpp"
#include "trk3_host.hpp"

namespace trk3 {
namespace {

std::string trim(const std::string &s) {
    size_t a = s.find_first_not_of(' '), b = s.find_last_not_of(' ');
    return a == std::string::npos ? std::string() : s.substr(a, b - a + 1);
}
bool dir_exists(const std::string &p) { struct stat st; return stat(p.c_str(), &st) == 0; }
bool make_dirs(const std::string &p) {
    std::string cur;
    for (size_t i = 0; i <= p.size(); ++i) {
        if (i == p.size() || p[i] == '/') {
            if (!cur.empty() && !dir_exists(cur) && mkdir(cur.c_str(), 0777) != 0 && !dir_exists(cur)) return false;
        }
        if (i < p.size()) cur += p[i];
    }
    return true;
}

struct Out {                       // tallies divided by NMC (MAIN.f90:281-314), Fortran element order
    const trk3_tally_layout &lay;
    std::vector<double> v;
    int Nt, NR, Nat, Ns1, Nd;
    Out(const trk3_tally_layout &l, const double *sum, int NMC) : lay(l), v((size_t)l.total) {
        for (int64_t i = 0; i < l.total; ++i) v[(size_t)i] = sum[i] / (double)NMC;
        Nt = l.Nt; NR = l.n_r; Nat = l.n_atoms; Ns1 = l.nshl1; Nd = l.n_dos;
    }
    double a1(int id, int k) const { return v[(size_t)(lay.off[id] + k)]; }                                   // (k)
    double a2(int id, int k, int i, int ld) const { return v[(size_t)(lay.off[id] + k + (int64_t)ld * i)]; }   // (k,i)
    double a4(int id, int k, int i, int at, int sh) const { return v[(size_t)(lay.off[id] + k + (int64_t)Nt * (i + (int64_t)NR * (at + (int64_t)Nat * sh)))]; }
    double a3(int id, int k, int at, int sh) const { return v[(size_t)(lay.off[id] + k + (int64_t)Nt * (at + (int64_t)Nat * sh))]; }
};

// the common layout of every radial file: header '#Radius[A]' + times, then one row per radius
template <class F>
bool radial_file(const std::string &path, const char *first, const std::vector<double> &tg, int Nt, const std::vector<double> &x, int xd, double xscale, F value) {
    FILE *f = fopen(path.c_str(), "w");
    if (!f) return false;
    fputs(first, f);
    for (int i = 0; i < Nt; ++i) fprintf(f, "%s[fs]   ", fmt_f(tg[i], 10, 2).c_str());
    fputs(" \n", f);
    for (size_t i = 0; i < x.size(); ++i) {
        fputs(fmt_f(x[i] * xscale, 9, xd).c_str(), f);
        for (int k = 0; k < Nt; ++k) fputs(fmt_e(value(k, (int)i)).c_str(), f);
        fputs(" \n", f);
    }
    fclose(f);
    return true;
}

void write_parameters(const Case &c, const Out &o, const std::string &path, int NMC) {
    FILE *f = fopen(path.c_str(), "w");
    if (!f) return;
    const std::string dash(100, '-');
    fprintf(f, " TREKIS-3 Monte-Carlo cascade engine, B200-native implementation (trekis3_b200)\n");
    fprintf(f, " Performing calculations for %s in %s\n", c.SHI.Full_Name.c_str(), c.Material_name.c_str());
    fprintf(f, " Ion %s (Element #%d, Mass %s)\n", c.SHI.Name.c_str(), c.SHI.Zat, trim(fmt_f(c.SHI.Mass, 12, 3)).c_str());
    fprintf(f, " With energy %s [MeV]\n%s\n", trim(fmt_f(c.SHI.E / 1e6, 9, 2)).c_str(), dash.c_str());
    fprintf(f, " Material: %s (%s)\n", c.Material_name.c_str(), c.Matter.Target_name.c_str());
    fprintf(f, " Material density: %s [g/cm^3] or %.7E [1/cm^3]\n", trim(fmt_f(c.Matter.Dens, 12, 3)).c_str(), c.Matter.At_Dens);
    fprintf(f, " Thickness of the analysed layer: %s [A]\n", trim(fmt_f(c.Matter.Layer, 12, 3)).c_str());
    fprintf(f, " Temperature of the target: %s [K]\n", trim(fmt_f(c.Matter.temp, 12, 3)).c_str());
    if (c.Matter.El_eff_mass == 0) fprintf(f, " Effective mass of valence electrons is calculated from DOS.\n");
    else fprintf(f, " Effective mass of valence electrons: %s [me]\n", trim(fmt_f(c.Matter.El_eff_mass, 10, 2)).c_str());
    if (c.Matter.hole_mass > 0) fprintf(f, " Effective mass of valence holes: %s [me]\n", std::fabs(c.Matter.hole_mass) < 1e6 ? trim(fmt_f(c.Matter.hole_mass, 10, 2)).c_str() : "infinite");
    else fprintf(f, " Effective mass of valence holes is calculated from DOS\n");
    fprintf(f, " DOS file used: %s\n CDF file used: %s\n%s\n", c.numpar.DOS_file.c_str(), c.numpar.CDF_file.c_str(), dash.c_str());
    fprintf(f, " Total time to be analysed: %s [fs]\n", trim(fmt_f(c.Tim, 12, 3)).c_str());
    if (c.numpar.dt_flag <= 0) fprintf(f, " with the timestep of %s [fs], linear time-scale\n", trim(fmt_f(c.dt, 12, 3)).c_str());
    else fprintf(f, " starting with 0.01 [fs] increasing by %s in logarithmic time-scale\n", trim(fmt_f(c.dt, 12, 3)).c_str());
    fprintf(f, "%s\n", dash.c_str());
    static const char *zn[] = {"Barkas", "Bohr", "Nikolaev-Dmitriev", "Schiwietz-Grande", "fixed"};
    fprintf(f, " Ion equilibrium charge is used with %s formula, %s model for SHI.\n", zn[(c.SHI.Kind_Zeff >= 1 && c.SHI.Kind_Zeff <= 4) ? c.SHI.Kind_Zeff : 0],
            c.SHI.Kind_ion == 1 ? "Brandt-Kitagawa" : "point-like charge");
    const char *dr = c.numpar.kind_of_DR == 3 ? "dispersion relation from Ritchie and Howie" : c.numpar.kind_of_DR == 2 ? "plasmon-pole dispersion relation of scattering centers"
                   : c.numpar.kind_of_DR == 4 ? "Delta-function CDF" : "free electrons";
    fprintf(f, " Inelastic scattering calculated with %s\n (using tabulated files with integrated diff.CS)\n", dr);
    fprintf(f, " Elastic scattering %s\n", c.numpar.kind_of_EMFP == 2 ? "calculated with DSF cross-sections" : c.numpar.kind_of_EMFP == 1
            ? (c.numpar.kind_of_CDF_ph == 0 ? "calculated with phonon CDF (Ritchie-Howie)" : "calculated with single-pole phonon CDF")
            : c.numpar.kind_of_EMFP == 0 ? "calculated with Mott atomic cross-sections" : "is excluded");
    if (c.Matter.cut_off > 0) fprintf(f, " Energy cut-off used is %s\n", trim(fmt_f(c.Matter.cut_off, 12, 3)).c_str());
    else fprintf(f, " No energy cut-off is used\n");
    if (c.Matter.work_function > 0) fprintf(f, " Electron emission included. Work function = %s [eV]\n Potential barrier length = %s [A] Barrier height = %s [eV]\n",
            trim(fmt_f(c.Matter.work_function, 12, 3)).c_str(), trim(fmt_f(c.Matter.bar_length, 12, 3)).c_str(), trim(fmt_f(c.Matter.bar_height, 12, 3)).c_str());
    else fprintf(f, " Electron emission is excluded\n");
    fprintf(f, " Radiative decays of deep-shell holes and photon transport are %s\n", c.numpar.include_photons ? "included" : "excluded");
    fprintf(f, " Transient electric fields are excluded\n Number of MC iterations: %d\n", NMC);
    fprintf(f, " Monte-Carlo engine: CUDA (sm_100a), Philox4x32-10 streams keyed by the global iteration index\n%s\n", dash.c_str());
    for (const Atom &a : c.atoms) {
        fprintf(f, " Atom %s (Z = %d, mass %s, contribution %s)\n", a.Name.c_str(), a.Zat, trim(fmt_f(a.Mass, 12, 3)).c_str(), trim(fmt_f(a.Pers, 8, 3)).c_str());
        for (int j = 0; j < a.nshl(); ++j)
            fprintf(f, "   shell %-12s Ip = %s [eV]  Nel = %s  Auger time = %.4E [fs]  radiative time = %.4E [fs]\n", a.Shell_name[j].c_str(),
                    trim(fmt_f(a.Ip[j], 12, 3)).c_str(), trim(fmt_f(a.Nel[j], 8, 3)).c_str(), a.Auger[j], a.Radiat[j]);
    }
    fprintf(f, "%s\n", dash.c_str());
    fprintf(f, "Ion equilibrium charge:         %s [electron charge]\n", trim(fmt_f(c.SHI.Zeff, 6, 3)).c_str());
    fprintf(f, "MC calculated energy loss (Se): %s [eV/A]\n%s\n", trim(fmt_f(o.a1(TRK3_OUT_TOT_E, o.Nt - 1) / c.Matter.Layer, 9, 2)).c_str(), dash.c_str());
    fclose(f);
}

}  // namespace

bool save_output(const Case &c, const trk3_tally_layout &lay, const double *sum, int NMC, const std::string &out_root,
                 std::string &out_dir, std::string &err) {
    if (!sum || NMC < 1) { err = "save_output: no tallies / NMC < 1"; return false; }
    if (lay.n_atoms != (int)c.atoms.size() || lay.n_r != (int)c.Out_R.size() || lay.n_dos != (int)c.dos.E.size()) { err = "save_output: layout does not belong to this case"; return false; }
    const Out o(lay, sum, NMC);
    const int N = o.Nt, NR = o.NR;
    // the time grid as Save_output rebuilds it (:390-412): last point clamped to Tim
    std::vector<double> tg((size_t)N);
    if (c.numpar.dt_flag <= 0) { tg[0] = c.dt; for (int i = 1; i < N; ++i) tg[i] = std::min(tg[i - 1] + c.dt, c.Tim); }
    else { tg[0] = 0.01; for (int i = 1; i < N; ++i) tg[i] = tg[i - 1] * c.dt; tg[N - 1] = std::min(tg[N - 1], c.Tim); }

    // ---- directory (:437-471); Output_path_SHI = OUTPUT_<material>/OUTPUT_<ion>_in_<material> (MAIN.f90, Reading_files...:432-450)
    const std::string base = out_root + "/OUTPUT_" + c.Material_name + "/OUTPUT_" + c.SHI.Name + "_in_" + c.Material_name;
    std::string stem = base + "/" + c.SHI.Name + "_E_" + trim(fmt_f(c.SHI.E / 1e6, 8, 2)) + "_MeV_" + trim(fmt_f(c.Tim, 10, 2)) + "_fs";
    std::string d = stem;
    for (int i = 1; dir_exists(d); ++i) d = stem + "_" + std::to_string(i);
    if (!make_dirs(d)) { err = "save_output: cannot create " + d; return false; }
    out_dir = d;
    {   // copy of the input file for reproducibility (:476-481)
        std::ofstream f(d + "/INPUT_PARAMETERS.txt");
        for (const std::string &l : c.input_lines) f << l << "\n";
    }
    write_parameters(c, o, d + "/!Parameters.txt", NMC);

    const std::vector<double> &R = c.Out_R;
    auto rad = [&](const std::string &name, auto value) { return radial_file(d + "/" + name, "#Radius[A] ", tg, N, R, 1, 1.0, value); };
    bool ok = true;
    // ---- angular distributions (:538-588): first column written with '(e)'
    for (int hole = 0; hole < 2; ++hole) {
        FILE *f = fopen((d + (hole ? "/VB_holes_theta_distribution.txt" : "/Electrons_theta_distribution.txt")).c_str(), "w");
        if (!f) { ok = false; continue; }
        fputs(hole ? "Angle[deg] " : "#Angle[deg] ", f);
        for (int i = 0; i < N; ++i) fprintf(f, "%s[fs]   ", fmt_f(tg[i], 10, 2).c_str());
        fputs(" \n", f);
        for (int i = 0; i < TRK3_NTHETA; ++i) {
            fputs(fmt_e((double)(i + 1)).c_str(), f);
            for (int k = 0; k < N; ++k) fputs(fmt_e(o.a2(hole ? TRK3_OUT_THETA_H : TRK3_OUT_THETA, k, i, N + 1)).c_str(), f);
            fputs(" \n", f);
        }
        fclose(f);
    }
    if (c.Matter.work_function > 0.0)       // :591-616, energy grid Out_R/10
        ok &= radial_file(d + "/Emitted_electron_distribution_vs_E[1_eV].txt", "#Energy[eV] ", tg, N, R, 1, 0.1, [&](int k, int i) { return o.a2(TRK3_OUT_EE_VS_E_EM, k, i, N); });
    {   // Total_numbers.txt (:620-643)
        FILE *f = fopen((d + "/Total_numbers.txt").c_str(), "w");
        if (f) {
            fputs(c.numpar.include_photons ? "#Time[fs]    Ne    Ne_Emitted    Energy[eV]     Energy_Emitted[eV] N_photons\n" : "#Time[fs]    Ne    Ne_Emitted    Energy[eV]     Energy_Emitted[eV]\n", f);
            for (int i = 0; i < N; ++i) {
                fprintf(f, "%s%s%s%s%s", fmt_e(tg[i]).c_str(), fmt_e(o.a1(TRK3_OUT_TOT_NE, i)).c_str(), fmt_e(o.a1(TRK3_OUT_NE_EM, i)).c_str(),
                        fmt_e(o.a1(TRK3_OUT_TOT_E, i)).c_str(), fmt_e(o.a1(TRK3_OUT_E_EM, i)).c_str());
                if (c.numpar.include_photons) fputs(fmt_e(o.a1(TRK3_OUT_TOT_NPHOT, i)).c_str(), f);
                fputs("\n", f);
            }
            fclose(f);
        } else ok = false;
    }
    {   // Hole_mean_diffusion_coefficient.txt (:648-663); the file exists (empty) when the holes are immobile
        FILE *f = fopen((d + "/Hole_mean_diffusion_coefficient.txt").c_str(), "w");
        if (f) {
            if (c.Matter.hole_mass < 1.0e3) {
                fputs("#Time[fs]    Diffusion_coeff[cm^2/s]\n", f);
                for (int i = 0; i < N; ++i) fprintf(f, "%s%s\n", fmt_e(tg[i]).c_str(), fmt_e(o.a1(TRK3_OUT_DIFF_COEFF, i)).c_str());
            }
            fclose(f);
        } else ok = false;
    }
    {   // Total_energies.txt (:667-705); Out_E_h has the shells of atom 1 as its last extent
        FILE *f = fopen((d + "/Total_energies.txt").c_str(), "w");
        if (f) {
            fputs(c.numpar.include_photons ? "#Time[fs]    Electrons[eV]   Atoms[eV]   Field[eV]  Photons[eV] " : "#Time[fs]    Electrons[eV]   Atoms[eV]   Field[eV]  ", f);
            for (const Atom &a : c.atoms) for (int j = 0; j < a.nshl(); ++j) {
                if (a.Shl_num[j] < 63) fprintf(f, "%s_%s[eV]    ", a.Name.c_str(), trim(a.Shell_name[j]).c_str());
                else fprintf(f, "%s[eV]    ", trim(a.Shell_name[j]).c_str());
            }
            fputs(" \n", f);
            for (int k = 0; k < N; ++k) {
                fprintf(f, "%s%s%s%s", fmt_e(tg[k]).c_str(), fmt_e(o.a1(TRK3_OUT_E_E, k)).c_str(), fmt_e(o.a1(TRK3_OUT_E_AT, k)).c_str(), fmt_e(o.a1(TRK3_OUT_E_FIELD, k)).c_str());
                if (c.numpar.include_photons) fputs(fmt_e(o.a1(TRK3_OUT_E_PHOT, k)).c_str(), f);
                for (int i = 0; i < o.Nat; ++i) for (int j = 0; j < c.atoms[i].nshl(); ++j) fputs(fmt_e(j < o.Ns1 ? o.a3(TRK3_OUT_E_H, k, i, j) : 0.0).c_str(), f);
                fputs(" \n", f);
            }
            fclose(f);
        } else ok = false;
    }
    // ---- electrons (:709-795)
    ok &= rad("Radial_electron_density[1_cm^-3].txt", [&](int k, int i) { return o.a2(TRK3_OUT_NE, k, i, N) * 1.0e24; });
    ok &= rad("Radial_electron_energy[eV_A^-3].txt", [&](int k, int i) { return o.a2(TRK3_OUT_EE, k, i, N); });
    ok &= rad("Radial_electron_temperature[K].txt", [&](int k, int i) { const double n = o.a2(TRK3_OUT_NE, k, i, N); return n > 0.0 ? o.a2(TRK3_OUT_EE, k, i, N) / n * g_kb * 2.0 / 3.0 : 0.0; });
    if (c.numpar.include_photons) {         // :797-847
        ok &= rad("Radial_photon_density[1_cm^-3].txt", [&](int k, int i) { return o.a2(TRK3_OUT_NPHOT, k, i, N) * 1.0e24; });
        ok &= rad("Radial_photon_energy[eV_A^-3].txt", [&](int k, int i) { return o.a2(TRK3_OUT_EPHOT, k, i, N); });
    }
    {   // Electron_distribution_vs_E[1_eV].txt (:850-877): two header lines, the radius grid is the energy grid (sic)
        FILE *f = fopen((d + "/Electron_distribution_vs_E[1_eV].txt").c_str(), "w");
        if (f) {
            fputs("#Energy[eV]   Spectrum[arb.units]\n#time[fs]:  ", f);
            for (int i = 0; i < N; ++i) fputs(fmt_f(tg[i], 10, 2).c_str(), f);
            fputs(" \n", f);
            for (int i = 0; i < NR; ++i) {
                fputs(fmt_f(R[i], 9, 1).c_str(), f);
                for (int k = 0; k < N; ++k) fputs(fmt_e(o.a2(TRK3_OUT_EE_VS_E, k, i, N)).c_str(), f);
                fputs(" \n", f);
            }
            fclose(f);
        } else ok = false;
    }
    ok &= radial_file(d + "/VB_holes_distribution_vs_E[1_eV].txt", "#Energy[eV] ", tg, N, c.dos.E, 1, 1.0, [&](int k, int i) { return o.a2(TRK3_OUT_EH_VS_E, k, i, N); });
    // ---- lattice (:906-992): Out_Elat is binned by time interval, the files show the running sum
    auto elat_cum = [&](int k, int i) { double s = 0.0; for (int q = 0; q <= k; ++q) s += o.a2(TRK3_OUT_ELAT, q, i, N); return s; };
    const int vb_at = c.Lowest_Ip_At, vb_sh = c.Lowest_Ip_Shl;
    ok &= rad("Radial_Lattice_energy[eV_A^-3].txt", elat_cum);
    ok &= rad("Radial_Track_energy[eV_A^-3].txt", [&](int k, int i) { return elat_cum(k, i) + o.a4(TRK3_OUT_EH, k, i, vb_at, vb_sh) + o.a4(TRK3_OUT_EHKIN, k, i, vb_at, vb_sh); });
    ok &= rad("Radial_Lattice_temperature[K].txt", [&](int k, int i) { return elat_cum(k, i) / (c.Matter.At_Dens * 1e-24) * g_kb * 2.0 / 3.0; });
    // ---- holes per shell (:995-1138)
    for (int j = 0; j < o.Nat; ++j) for (int l = 0; l < c.atoms[j].nshl(); ++l) {
        const Atom &a = c.atoms[j];
        const bool vb = !(a.Shl_num[l] < 63);
        const std::string nm = vb ? trim(a.Shell_name[l]) : a.Name + "_" + trim(a.Shell_name[l]);
        auto h = [&](int id) { return [&, id](int k, int i) { return l < o.Ns1 ? o.a4(id, k, i, j, l) : 0.0; }; };
        ok &= rad("Radial_" + nm + "_holes_density[1_cm^-3].txt", [&](int k, int i) { return (l < o.Ns1 ? o.a4(TRK3_OUT_NH, k, i, j, l) : 0.0) * 1.0e24; });
        if (!vb) ok &= rad("Radial_" + nm + "_holes_energy[eV_A^-3].txt", h(TRK3_OUT_EH));
        else {
            ok &= rad("Radial_" + nm + "_holes_pot_energy[eV_A^-3].txt", h(TRK3_OUT_EH));
            ok &= rad("Radial_" + nm + "_holes_kin_energy[eV_A^-3].txt", h(TRK3_OUT_EHKIN));
        }
        if (a.Shl_num[l] == 63) ok &= rad("Radial_" + nm + "_holes_temperature[K].txt", [&](int k, int i) {
            const double n = l < o.Ns1 ? o.a4(TRK3_OUT_NH, k, i, j, l) : 0.0;
            return n > 0.0 ? o.a4(TRK3_OUT_EHKIN, k, i, j, l) / n * g_kb * 2.0 / 3.0 : 0.0; });
    }
    if (!ok) { err = "save_output: could not write every file in " + d; return false; }
    return true;
}

}  // namespace trk3
