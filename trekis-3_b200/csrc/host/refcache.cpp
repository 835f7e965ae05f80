// refcache.cpp -- the reference's on-disk table cache, written and read in the reference's own layout (SURVEY.md 3.5,
// row N2 of 8(f)).  Tables built here (on the GPU in 0.3 s) can be dropped into an existing TREKIS-3 working directory,
// where the Fortran program finds them instead of integrating for minutes to hours; and tables cached by the reference
// feed this engine directly (level-1 parity by file comparison once a reference build exists).
//
//   OUTPUT_<material>/
//     OUTPUT_Electron_IMFPs_<Free|Plasmon_pole|Ritchie|Delta>_<CDF|spCDF>_<DOS|Me|x.xx_me>_<T>_K.dat   Analytical_IMFPs.f90:262-291
//     OUTPUT_Hole_IMFPs_CDF_<CDF|spCDF>_<T>_K.dat                                                      :385-408
//     OUTPUT_Photon_IMFPs_<CDF|spCDF>_<T>_K.dat                                                        :506-526
//         one row per grid energy: '(f)' E, '(e)' L of every (atom, shell), '(e)' total                :596-634
//     OUTPUT_Electron_EMFPs_<CDF|spCDF>_<Zeff|Z=1|Z_CDFe_FF|Z_CDFe>_<T>_K.dat | ..._Mott.dat | OUTPUT_Electron_No_elas_EMFPs.dat   :917-1112
//     OUTPUT_Hole_<CDF|spCDF>_EMFPs_<T>_K.dat | OUTPUT_Hole_Mott_EMFPs_<T>_K.dat | OUTPUT_Hole_No_elas_EMFPs.dat                    :1190-1380
//         rows '(f,e)' E, L                                                                            :1391-1418
//     diff_CS/<table file without 'OUTPUT_' and '.dat'>[_<atom>_<shell>]_<E as f14.3>.dat              :667-786, 1449-1606
//         rows '(es,es)' hw, cumulative MFP
//     OUTPUT_<ion>_in_<material>/OUTPUT_<ion>_<CDF|spCDF>_<Barkas|Bohr|ND|SG|fixed>_<P|BK>_{IMFP,dEdx,effective_charges,Range}.dat
//         '(e)' E [MeV], per-shell values, total (:2657-2707); Range: two header lines + '(es,es,es)' E [eV], dE/dx, range (:2462-2506)
//
// Validity as in the reference (:307-327, 422-437, 540-555, 966-981, 2364-2372): a table file counts if it has as many rows
// as the energy grid, is not older than the .cdf file it was computed from (get_file_stat, Last_mod_time_CDF), and no
// redo_MFP / redo_IMFP / redo_EMFP / redo_MFP_SHI keyword asks for a recalculation.  The cache is read all-or-nothing here:
// where the reference would recompute one table family, read_reference_cache declines and the caller rebuilds them all
// (the tables are deterministic, so the untouched families come out as they were).
#include <sys/stat.h>
#include <algorithm>
#include <cstdio>
#include <fstream>
#include "fortran_fmt.hpp"
#include "trk3_host.hpp"

namespace trk3 {
namespace {

std::string trim_s(const std::string &s) {
    size_t a = s.find_first_not_of(' '), b = s.find_last_not_of(' ');
    return a == std::string::npos ? std::string() : s.substr(a, b - a + 1);
}
bool exists(const std::string &p) { struct stat st; return stat(p.c_str(), &st) == 0; }
// modification time [s]; -1 if the file cannot be examined
double mtime_of(const std::string &p) { struct stat st; if (stat(p.c_str(), &st) != 0) return -1.0; return (double)st.st_mtim.tv_sec + 1.0e-9 * (double)st.st_mtim.tv_nsec; }
bool make_dirs(const std::string &p) {
    std::string cur;
    for (size_t i = 0; i <= p.size(); ++i) {
        if (i == p.size() || p[i] == '/') {
            if (!cur.empty() && !exists(cur) && mkdir(cur.c_str(), 0777) != 0 && !exists(cur)) return false;
        }
        if (i < p.size()) cur += p[i];
    }
    return true;
}
std::string kcs(int kind_of_CDF) { return kind_of_CDF == 0 ? "_CDF" : kind_of_CDF == 1 ? "_spCDF" : ""; }   // :243-249

bool read_rows(const std::string &path, size_t ncol_min, std::vector<std::vector<double>> &rows, std::string &err, int skip = 0) {
    std::ifstream f(path);
    if (!f) { err = "cannot read " + path; return false; }
    rows.clear();
    std::string line;
    std::vector<double> v;
    while (std::getline(f, line)) {
        if (skip > 0) { --skip; continue; }
        if (trim_s(line).empty()) continue;
        if (!parse_fortran_line(line, v) || v.size() < ncol_min) { err = "bad record in " + path + ": '" + line + "'"; return false; }
        rows.push_back(v);
    }
    return true;
}

// ---- E, L of every (atom, shell), total (:596-634)
bool write_shell_table(const std::string &path, const std::vector<std::vector<MFP>> &T, std::string &err) {
    FILE *f = fopen(path.c_str(), "w");
    if (!f) { err = "cannot write " + path; return false; }
    const size_t N = T[0][0].E.size();
    for (size_t i = 0; i < N; ++i) {
        fputs(fmt_fd(T[0][0].E[i]).c_str(), f);
        double inv = 0.0;
        for (auto &at : T) for (auto &m : at) { fputs(fmt_e(m.L[i]).c_str(), f); inv += 1.0 / m.L[i]; }
        fputs(fmt_e(1.0 / inv).c_str(), f);
        fputc('\n', f);
    }
    fclose(f);
    return true;
}
bool read_shell_table(const std::string &path, const Case &c, size_t N, std::vector<std::vector<MFP>> &T, std::string &err) {
    std::vector<std::vector<double>> rows;
    const size_t ns = (size_t)c.n_shells();
    if (!read_rows(path, ns + 2, rows, err)) return false;
    if (rows.size() != N) { err = "energy grid mismatch in " + path + " (" + std::to_string(rows.size()) + " rows, grid has " + std::to_string(N) + ")"; return false; }
    T.assign(c.atoms.size(), {});
    size_t col = 1;
    for (size_t j = 0; j < c.atoms.size(); ++j) {
        T[j].assign((size_t)c.atoms[j].nshl(), MFP{});
        for (auto &m : T[j]) {
            m.E.resize(N); m.L.resize(N); m.dEdx.assign(N, 0.0);
            for (size_t i = 0; i < N; ++i) { m.E[i] = rows[i][0]; m.L[i] = rows[i][col]; }
            ++col;
        }
    }
    return true;
}

// ---- E, L (:1391-1418)
bool write_pair_table(const std::string &path, const MFP &m, std::string &err) {
    FILE *f = fopen(path.c_str(), "w");
    if (!f) { err = "cannot write " + path; return false; }
    for (size_t i = 0; i < m.E.size(); ++i) fprintf(f, "%s%s\n", fmt_fd(m.E[i]).c_str(), fmt_e(m.L[i]).c_str());
    fclose(f);
    return true;
}
bool read_pair_table(const std::string &path, size_t N, MFP &m, std::string &err) {
    std::vector<std::vector<double>> rows;
    if (!read_rows(path, 2, rows, err)) return false;
    if (rows.size() != N) { err = "energy grid mismatch in " + path; return false; }
    m.E.resize(N); m.L.resize(N); m.dEdx.assign(N, 0.0);
    for (size_t i = 0; i < N; ++i) { m.E[i] = rows[i][0]; m.L[i] = rows[i][1]; }
    return true;
}

// ---- one differential table: a file per grid energy (:667-716)
bool write_diff(const std::string &dir, const std::string &table_file, const std::string &atom, const std::string &shell,
                const DiffCS &d, int &n_files, std::string &err) {
    for (size_t i = 0; i < d.E.size(); ++i) {
        const std::string path = dir + "/" + diff_cs_file_name(table_file, atom, shell, d.E[i]);
        FILE *f = fopen(path.c_str(), "w");
        if (!f) { err = "cannot write " + path; return false; }
        const DiffRow &r = d.row[i];
        for (size_t k = 0; k < r.hw.size(); ++k) fprintf(f, "%s%s\n", fmt_es(r.hw[k]).c_str(), fmt_es(r.L[k]).c_str());
        fclose(f);
        ++n_files;
    }
    return true;
}
bool read_diff(const std::string &dir, const std::string &table_file, const std::string &atom, const std::string &shell,
               const std::vector<double> &E, DiffCS &d, std::string &err) {
    d.E = E;
    d.row.assign(E.size(), DiffRow{});
    std::vector<std::vector<double>> rows;
    for (size_t i = 0; i < E.size(); ++i) {
        if (!read_rows(dir + "/" + diff_cs_file_name(table_file, atom, shell, E[i]), 2, rows, err)) return false;
        DiffRow &r = d.row[i];
        r.hw.resize(rows.size()); r.L.resize(rows.size());
        for (size_t k = 0; k < rows.size(); ++k) { r.hw[k] = rows[k][0]; r.L[k] = rows[k][1]; }
    }
    return true;
}

// the energy grids of MAIN.f90 / Analytical_electron_dEdx (same calls as build_tables)
struct Grids { std::vector<double> shi, el, el_elast, hole, hole_elast, phot; };
Grids make_grids(const Case &c) {
    Grids g;
    const double M = c.SHI.Mass * g_Mp;
    g.shi = get_grid_4CS(c.atoms, std::ceil((M + g_me) * (M + g_me) / (M * g_me) * c.atoms[0].Ip.back() / 4.0), 175.6e6 / 2.0 * c.SHI.Mass);
    const double Emax_e = 175.6e6 * 2.0 / 1836.0;
    const double Emax_h = *std::max_element(c.dos.E.begin(), c.dos.E.end()) + 1.0;
    g.el = get_grid_4CS(c.atoms, 0.1, Emax_e);  g.el_elast = get_grid_4CS(c.atoms, 0.01, Emax_e);
    g.hole = get_grid_4CS(c.atoms, 0.1, Emax_h); g.hole_elast = get_grid_4CS(c.atoms, 0.01, Emax_h);
    g.phot = g.el;
    return g;
}

}  // namespace

std::string diff_cs_file_name(const std::string &table_file, const std::string &atom, const std::string &shell, double E) {
    // full_CS_file(8 : LEN-4): without 'OUTPUT_' and '.dat' (:685-692)
    std::string s = table_file.substr(7, table_file.size() - 7 - 4);
    if (!atom.empty()) s += "_" + trim_s(atom) + "_" + trim_s(shell);
    return s + "_" + trim_s(fmt_f(E, 14, 3)) + ".dat";
}

RefCacheNames reference_cache_names(const Case &c) {
    RefCacheNames n;
    n.dir_material = "OUTPUT_" + c.Material_name;                                    // Reading_files_and_parameters.f90:432-450
    n.dir_ion = n.dir_material + "/OUTPUT_" + c.SHI.Name + "_in_" + c.Material_name;
    n.dir_diff = n.dir_material + "/diff_CS";
    const std::string T9 = trim_s(fmt_f(c.Matter.temp, 9, 2)), T7 = trim_s(fmt_f(c.Matter.temp, 7, 2));
    // electrons (:262-291)
    std::string mass;
    if (c.Matter.El_eff_mass == 0) mass = "DOS_" + T9 + "_K";
    else if (c.Matter.El_eff_mass < 0 || c.Matter.El_eff_mass == 1.0) mass = "Me_" + T9 + "_K";
    else mass = trim_s(fmt_f(c.Matter.El_eff_mass, 5, 2)) + "_me_" + T9 + "_K";
    const std::string k = kcs(c.numpar.kind_of_CDF);
    const std::string tc = (k.empty() ? std::string() : k.substr(1)) + "_" + mass;
    const int dr = c.numpar.kind_of_DR;
    n.el_imfp = std::string("OUTPUT_Electron_IMFPs_") + (dr == 4 ? "Delta_" : dr == 3 ? "Ritchie_" : dr == 2 ? "Plasmon_pole_" : "Free_") + tc + ".dat";
    n.hole_imfp = "OUTPUT_Hole_IMFPs_CDF" + k + "_" + T7 + "_K.dat";                 // :385-408
    n.photon_imfp = "OUTPUT_Photon_IMFPs" + k + "_" + T7 + "_K.dat";                 // :506-526
    const std::string kp = kcs(c.numpar.kind_of_CDF_ph);
    switch (c.numpar.kind_of_EMFP) {
    case 1: {
        const int z = c.numpar.CDF_elast_Zeff;
        n.el_emfp = "OUTPUT_Electron_EMFPs" + kp + "_" + (z == 1 ? "Z=1_" : z == 2 ? "Z_CDFe_FF_" : z == 3 ? "Z_CDFe_" : "Zeff_") + T7 + "_K.dat";   // :929-948
        n.hole_emfp = "OUTPUT_Hole" + kp + "_EMFPs_" + T7 + "_K.dat";                // :1199-1201
        break;
    }
    case 0:
        n.el_emfp = "OUTPUT_Electron_EMFPs_Mott.dat";                                 // :1052
        n.hole_emfp = "OUTPUT_Hole_Mott_EMFPs_" + T7 + "_K.dat";                      // :1318
        break;
    default:
        n.el_emfp = "OUTPUT_Electron_No_elas_EMFPs.dat";                              // :1104
        n.hole_emfp = "OUTPUT_Hole_No_elas_EMFPs.dat";                                // :1370
    }
    static const char *zn[] = {"_Barkas", "_Bohr", "_ND", "_SG", "_fixed"};        // :2314-2333
    n.shi_stem = "OUTPUT_" + trim_s(c.SHI.Name) + k + zn[(c.SHI.Kind_Zeff >= 1 && c.SHI.Kind_Zeff <= 4) ? c.SHI.Kind_Zeff : 0] + (c.SHI.Kind_ion == 1 ? "_BK" : "_P");
    return n;
}

// Closed-form shells have no differential tables: BEB shells (KOCS = 2) and every shell under the delta-function CDF
// (kind_of_DR = 4).  The reference still writes one diff_CS file per shell and grid energy for them -- from arrays its TotIMFP
// never filled (Cross_sections.f90:950-956, 1041-1047) -- and reads them back: that part of the cache carries no information and is
// not reproduced, so the reference-format cache is refused for such inputs (the binary .trk3tab cache serves them).
static bool closed_form_shells(const Case &c) {
    if (c.numpar.kind_of_DR == 4) return true;
    if (c.numpar.kind_of_EMFP == 2) return true;        // DSF: the elastic tables are the DSF file's own, nothing to cache
    for (const auto &a : c.atoms) for (int k : a.KOCS) if (k == 2) return true;
    return false;
}
bool write_reference_cache(const Case &c, const std::string &out_root, int *n_files_out, std::string &err) {
    if (!c.tables_built) { err = "write_reference_cache: tables not built"; return false; }
    if (closed_form_shells(c)) { err = "write_reference_cache: BEB shells / delta-function CDF / DSF scattering have no differential tables; the reference-format cache is not written for them"; return false; }
    const RefCacheNames n = reference_cache_names(c);
    const std::string dm = out_root + "/" + n.dir_material, di = out_root + "/" + n.dir_ion, dd = out_root + "/" + n.dir_diff;
    if (!make_dirs(di) || !make_dirs(dd)) { err = "cannot create " + di; return false; }
    int nf = 0;
    // ---- ion (:2657-2707): energies in MeV
    {
        FILE *f1 = fopen((di + "/" + n.shi_stem + "_IMFP.dat").c_str(), "w"), *f2 = fopen((di + "/" + n.shi_stem + "_dEdx.dat").c_str(), "w"),
             *f3 = fopen((di + "/" + n.shi_stem + "_effective_charges.dat").c_str(), "w"), *f4 = fopen((di + "/" + n.shi_stem + "_Range.dat").c_str(), "w");
        if (!f1 || !f2 || !f3 || !f4) { err = "cannot write the ion tables in " + di; return false; }
        const std::vector<double> &E = c.SHI_MFP[0][0].E;
        const size_t N = E.size();
        std::vector<double> dEdx_tot(N, 0.0);
        for (size_t i = 0; i < N; ++i) {
            Ion s1 = c.SHI; s1.E = E[i];
            equilibrium_charge_SHI(s1, c.atoms);
            fputs(fmt_e(E[i] / 1.0e6).c_str(), f1); fputs(fmt_e(E[i] / 1.0e6).c_str(), f2);
            fprintf(f3, "%s%s\n", fmt_e(E[i] / 1.0e6).c_str(), fmt_e(s1.Zeff).c_str());
            double inv = 0.0;
            for (auto &at : c.SHI_MFP) for (auto &m : at) {
                fputs(fmt_e(m.L[i]).c_str(), f1); fputs(fmt_e(m.dEdx[i]).c_str(), f2);
                if (m.L[i] > 1.0e-10) inv += 1.0 / m.L[i]; else inv = 1.1e29;         // :2690-2694
                dEdx_tot[i] += m.dEdx[i];
            }
            fprintf(f1, "%s\n", fmt_e(1.0 / inv).c_str());
            fprintf(f2, "%s\n", fmt_e(dEdx_tot[i]).c_str());
        }
        // Get_ion_range (:2462-2506)
        fputs("# Energy dEdx    Range\n# [eV] [eV/A]    [A]\n", f4);
        double range = 0.0;
        for (size_t i = 0; i < N; ++i) {
            if (i > 0) range += (dEdx_tot[i - 1] == 0.0 ? 0.5 / dEdx_tot[i] : 0.5 * (1.0 / dEdx_tot[i] + 1.0 / dEdx_tot[i - 1])) * (E[i] - E[i - 1]);
            fprintf(f4, "%s%s%s\n", fmt_es(E[i]).c_str(), fmt_es(dEdx_tot[i]).c_str(), fmt_es(range).c_str());
        }
        fclose(f1); fclose(f2); fclose(f3); fclose(f4);
        nf += 4;
    }
    // ---- electrons, holes, photons
    if (!write_shell_table(dm + "/" + n.el_imfp, c.Total_el_MFPs, err)) return false;
    if (!write_shell_table(dm + "/" + n.hole_imfp, c.Total_Hole_MFPs, err)) return false;
    nf += 2;
    if (c.numpar.include_photons && !c.Total_Photon_MFPs.empty()) { if (!write_shell_table(dm + "/" + n.photon_imfp, c.Total_Photon_MFPs, err)) return false; ++nf; }
    for (size_t j = 0; j < c.atoms.size(); ++j)
        for (int s = 0; s < c.atoms[j].nshl(); ++s)
            if (!write_diff(dd, n.el_imfp, c.atoms[j].Name, c.atoms[j].Shell_name[(size_t)s], c.EIdCS[j][(size_t)s], nf, err)) return false;
    if (!write_diff(dd, n.hole_imfp, c.atoms[0].Name, c.atoms[0].Shell_name.back(), c.HIdCS, nf, err)) return false;    // VB only (:743-786)
    if (!write_pair_table(dm + "/" + n.el_emfp, c.Elastic_MFP, err)) return false;
    if (!write_pair_table(dm + "/" + n.hole_emfp, c.Elastic_Hole_MFP, err)) return false;
    nf += 2;
    if (c.numpar.kind_of_EMFP == 1) {
        if (!write_diff(dd, n.el_emfp, "", "", c.EEdCS, nf, err)) return false;
        if (!write_diff(dd, n.hole_emfp, "", "", c.HEdCS, nf, err)) return false;
    }
    if (n_files_out) *n_files_out = nf;
    return true;
}

bool read_reference_cache(Case &c, const std::string &out_root, const BuildOptions &opt, std::string &err) {
    if (closed_form_shells(c)) { err = "read_reference_cache: BEB shells / delta-function CDF / DSF scattering have no differential tables; build the tables instead"; return false; }
    get_single_pole(c);                                   // MAIN.f90:146 (before any table, cached or not)
    const RefCacheNames n = reference_cache_names(c);
    const std::string dm = out_root + "/" + n.dir_material, di = out_root + "/" + n.dir_ion, dd = out_root + "/" + n.dir_diff;
    const Grids g = make_grids(c);
    const size_t ns = (size_t)c.n_shells();
    // ---- validity beyond existence and row count: recalculation keywords and the age of the files
    if (c.numpar.redo_IMFP || c.numpar.redo_EMFP || c.numpar.redo_IMFP_SHI) {
        err = std::string("recalculation requested by keyword (") + (c.numpar.redo_IMFP ? "redo_IMFP " : "") + (c.numpar.redo_EMFP ? "redo_EMFP " : "") +
              (c.numpar.redo_IMFP_SHI ? "redo_MFP_SHI" : "") + ")";
        return false;
    }
    {
        const double t_cdf = mtime_of(c.dir + "/" + c.numpar.CDF_file);
        std::vector<std::string> files = {dm + "/" + n.el_imfp, dm + "/" + n.hole_imfp, di + "/" + n.shi_stem + "_IMFP.dat"};
        if (c.numpar.kind_of_EMFP == 0 || c.numpar.kind_of_EMFP == 1) { files.push_back(dm + "/" + n.el_emfp); files.push_back(dm + "/" + n.hole_emfp); }
        if (c.numpar.include_photons) files.push_back(dm + "/" + n.photon_imfp);
        for (const auto &f : files) {
            const double t = mtime_of(f);
            if (t >= 0.0 && t_cdf >= 0.0 && t < t_cdf) { err = "the CDF file was modified more recently than " + f; return false; }
        }
    }
    // ---- ion: read_SHI_MFP (Reading_files_and_parameters.f90:2855-2903); all four files must exist (:2357-2365)
    for (const char *sfx : {"_IMFP.dat", "_dEdx.dat", "_effective_charges.dat", "_Range.dat"})
        if (!exists(di + "/" + n.shi_stem + sfx)) { err = "missing " + di + "/" + n.shi_stem + sfx; return false; }
    {
        std::vector<std::vector<double>> r1, r2;
        if (!read_rows(di + "/" + n.shi_stem + "_IMFP.dat", ns + 2, r1, err) || !read_rows(di + "/" + n.shi_stem + "_dEdx.dat", ns + 2, r2, err)) return false;
        if (r1.size() != g.shi.size() || r2.size() != g.shi.size()) { err = "energy grid mismatch in the SHI MFP files of " + di; return false; }
        const size_t N = r1.size();
        c.SHI_MFP.assign(c.atoms.size(), {});
        size_t col = 1;
        for (size_t j = 0; j < c.atoms.size(); ++j) {
            c.SHI_MFP[j].assign((size_t)c.atoms[j].nshl(), MFP{});
            for (auto &m : c.SHI_MFP[j]) {
                m.E.resize(N); m.L.resize(N); m.dEdx.resize(N);
                for (size_t i = 0; i < N; ++i) { m.E[i] = r1[i][0] * 1.0e6; m.L[i] = r1[i][col]; m.dEdx[i] = r2[i][col]; }
                ++col;
            }
        }
    }
    equilibrium_charge_SHI(c.SHI, c.atoms);               // MAIN.f90:171
    // ---- electrons and holes
    if (!read_shell_table(dm + "/" + n.el_imfp, c, g.el.size(), c.Total_el_MFPs, err)) return false;
    if (!read_shell_table(dm + "/" + n.hole_imfp, c, g.hole.size(), c.Total_Hole_MFPs, err)) return false;
    c.EIdCS.assign(c.atoms.size(), {});
    for (size_t j = 0; j < c.atoms.size(); ++j) {
        c.EIdCS[j].assign((size_t)c.atoms[j].nshl(), DiffCS{});
        for (int s = 0; s < c.atoms[j].nshl(); ++s)
            if (!read_diff(dd, n.el_imfp, c.atoms[j].Name, c.atoms[j].Shell_name[(size_t)s], c.Total_el_MFPs[j][(size_t)s].E, c.EIdCS[j][(size_t)s], err)) return false;
    }
    if (!read_diff(dd, n.hole_imfp, c.atoms[0].Name, c.atoms[0].Shell_name.back(), c.Total_Hole_MFPs[0].back().E, c.HIdCS, err)) return false;
    if (c.numpar.kind_of_EMFP == 0 || c.numpar.kind_of_EMFP == 1) {
        if (!read_pair_table(dm + "/" + n.el_emfp, g.el_elast.size(), c.Elastic_MFP, err)) return false;
        if (!read_pair_table(dm + "/" + n.hole_emfp, g.hole_elast.size(), c.Elastic_Hole_MFP, err)) return false;
    } else {                                              // 'No_elas': the reference does not read these back (:1104-1112)
        c.Elastic_MFP.E = g.el_elast; c.Elastic_MFP.L.assign(g.el_elast.size(), 1.0e30); c.Elastic_MFP.dEdx.assign(g.el_elast.size(), 0.0);
        c.Elastic_Hole_MFP.E = g.hole_elast; c.Elastic_Hole_MFP.L.assign(g.hole_elast.size(), 1.0e30); c.Elastic_Hole_MFP.dEdx.assign(g.hole_elast.size(), 0.0);
    }
    if (c.numpar.kind_of_EMFP == 1) {
        if (!read_diff(dd, n.el_emfp, "", "", c.Elastic_MFP.E, c.EEdCS, err)) return false;
        if (!read_diff(dd, n.hole_emfp, "", "", c.Elastic_Hole_MFP.E, c.HEdCS, err)) return false;
    } else {
        c.EEdCS.E = c.Elastic_MFP.E; c.EEdCS.row.assign(c.Elastic_MFP.E.size(), DiffRow{});
        c.HEdCS.E = c.Elastic_Hole_MFP.E; c.HEdCS.row.assign(c.Elastic_Hole_MFP.E.size(), DiffRow{});
    }
    // ---- photons
    if (c.numpar.include_photons) { if (!read_shell_table(dm + "/" + n.photon_imfp, c, g.phot.size(), c.Total_Photon_MFPs, err)) return false; }
    else c.Total_Photon_MFPs.clear();
    return finish_tables(c, opt, err);
}

}  // namespace trk3
