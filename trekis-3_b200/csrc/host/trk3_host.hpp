// trk3_host.hpp -- host-side data model of the TREKIS-3 driver surface (C++17).
//
// This is the host half of the drop-in: it re-expresses what the Fortran host
// (Universal_MC_for_SHI_MAIN.f90, Reading_files_and_parameters.f90, Cross_sections.f90,
// Analytical_IMFPs.f90) prepares before calling do_Monte_Carlo: parsed inputs and the
// CDF-derived mean-free-path / differential cross-section tables.  Nothing here runs on
// the GPU; the tables are flattened into trk3_tables (include/trekis3_gpu.h) and uploaded once.
#pragma once
#include <cmath>
#include <cstdint>
#include <array>
#include <memory>
#include <string>
#include <vector>
#include "../../../include/trekis3_host.h"

namespace trk3 {

// ---- Universal_Constants.f90:24-107 (values verbatim: 1e-12 table parity depends on them)
constexpr double g_Pi   = 3.1415926535897932384626433832795;
constexpr double g_e    = 1.602176487e-19;
constexpr double g_me   = 9.1093821545e-31;
constexpr double g_cvel = 299792458.0;
constexpr double g_Mp   = 1836.1526724780 * g_me;
constexpr double g_h    = 1.05457162853e-34;   // "Plank constant" of the reference is hbar
constexpr double g_kb   = 11604.0;             // [K/eV]
constexpr double g_e0   = 8.854187817620e-12;
constexpr double g_Ry   = 13.6056981;
constexpr double g_a0   = 0.5291772085936;
constexpr double g_me_eV = 0.51099906 * 1.0e6;
inline double g_alpha() { return g_e * g_e / (g_h * g_cvel * 4.0 * g_Pi * g_e0); }
inline double g_v0() { return std::sqrt(2.0 * g_Ry * g_e / g_me); }

// ---- Objects.f90:211-217 (type CDF)
// one particle energy of a DSF table (Differential_MFP, Objects.f90:170-178): transferred-energy grid and the mean free paths
// integrated up to each of its points (total, emission only, absorption only)
struct DsfPoint { double E = 0; std::vector<double> dE, dL, dL_absorb, dL_emit; };
struct CDFosc { std::vector<double> E0, A, Gamma, alpha; };   // alpha: weights of the delta-function CDF (kind_of_DR = 4)

// ---- Objects.f90:257-275 (type Atom)
struct Atom {
    int Zat = 0;
    double Mass = 0.0;     // [proton masses]
    double Pers = 0.0;
    std::string Name, Full_Name;
    std::vector<std::string> Shell_name;
    std::vector<int> Shl_num, PQN, KOCS, KOCS_SHI;
    std::vector<double> Nel, Ip, Ek, Auger, Radiat;
    std::vector<CDFosc> Ritchi;
    int nshl() const { return (int)Ip.size(); }
};

// ---- Objects.f90:86-105 (type Solid)
struct Solid {
    std::string Target_name, Chem;
    double Dens = 0, Vsound = 0, At_Dens = 0, Layer = 0, N_VB_el = 0;
    double work_function = 0, bar_length = 0, bar_height = 0, hole_mass = 0, El_eff_mass = 0;
    double E_F = 0, v_f = 0, cut_off = 0, temp = 0, Egap = 0;
};

// ---- Objects.f90:111-123 (type Density_of_states)
struct DOS {
    std::vector<double> E, dos, int_DOS, k, Eff_m, DOS_inv, int_DOS_inv, k_inv, Eff_m_inv;
};

// ---- Objects.f90:129-167 (type Flag); only the fields the hot path and its tables need
struct NumPar {
    int kind_of_EMFP = 1, kind_of_CDF = 1, kind_of_CDF_ph = 1, kind_of_DR = 0, dt_flag = 1;
    bool include_photons = false, plasmon_Emax = false;
    int CDF_elast_Zeff = 0;
    bool do_gnuplot = true, verbose = false, very_verbose = false;
    bool redo_IMFP = false, redo_EMFP = false, redo_IMFP_SHI = false, print_CDF = false,
         print_CDF_optical = false, get_thermal = false;
    std::string plot_extension = "jpeg";
    int CS_method = 1;
    bool VB_CDF_defined = false;
    std::string CDF_file, DOS_file;
    int out_dim = 1;
};

// ---- Objects.f90:44-53 (type Ion)
struct Ion {
    int Zat = 0;
    double E = 0, Mass = 0, Zeff = 0, fixed_Zeff = 0;
    int Kind_Zeff = 0, Kind_ion = 0;
    std::string Name, Full_Name;
};

// ---- Objects.f90:184-205 (types MFP, All_MFP)
struct MFP { std::vector<double> E, L, dEdx; };
// ---- Objects.f90:234-238 (type diff_CS): per grid energy a row (hw, dsdhw)
struct DiffRow { std::vector<double> hw, L; };
struct DiffCS { std::vector<double> E; std::vector<DiffRow> row; };

struct Case {                       // everything Read_input_file + MAIN.f90:120-247 produce
    std::string dir;                // directory holding INPUT_PARAMETERS.txt, INPUT_CDF/ ...
    std::string Material_name;
    Ion SHI;
    double Tim = 0, dt = 0;
    int NMC = 0, Num_th = 1;
    Solid Matter;
    NumPar numpar;
    std::vector<Atom> atoms;
    CDFosc CDF_Phonon;
    DOS dos;
    std::vector<std::string> input_lines;   // verbatim copy of INPUT_PARAMETERS.txt (Save_output copies it)
    std::vector<std::array<double, 5>> form_factor;   // Matter%form_factor: row Z-1 = coefficients a1..a5 of element Z (CDF_elast_Zeff = 2)
    // tables (MAIN.f90:166-238)
    std::vector<std::vector<MFP>> SHI_MFP, diff_SHI_MFP, Total_el_MFPs, Total_Hole_MFPs, Total_Photon_MFPs;
    MFP Elastic_MFP, Elastic_Hole_MFP;
    std::vector<DsfPoint> DSF_DEMFP, DSF_DEMFP_H;      // kind_of_EMFP = 2 (dsf.cpp)
    std::vector<std::vector<DiffCS>> EIdCS;   // [atom][shell]
    DiffCS EEdCS, HIdCS, HEdCS;
    int Lowest_Ip_At = 0, Lowest_Ip_Shl = 0;  // 0-based
    std::vector<double> Out_R, Out_V;
    bool tables_built = false;
    bool phonon_renormalised = false;       // get_single_pole has rescaled the user's phonon CDF (CDF_elast_Zeff = 2 / 3)
    std::vector<std::string> warnings;
    int n_shells() const { int n = 0; for (auto &a : atoms) n += a.nshl(); return n; }
};

// ---- searches (Reading_files_and_parameters.f90:3348-3559); all return the 1-based Fortran index
int find_monoton_1d(const double *A, int N, double v);                 // Find_in_monotonous_1D_array
int find_monoton_2d(const double *A, int stride, int N, double v);     // Find_in_monotonous_2D_array (row `Indx` of a (2,N) array)
int find_monoton_decreasing(const double *A, int N, double v);         // Find_in_monoton_array_decreasing
inline int find_monoton_1d(const std::vector<double> &A, double v) { return find_monoton_1d(A.data(), (int)A.size(), v); }
double interpolate(int flag, double E1, double E2, double S1, double S2, double E);   // Cross_sections.f90:4051

// ---- DSF elastic cross sections (dsf.cpp)
std::string dsf_file_name(const Case &c, bool hole);
bool read_dsf(const std::string &path, std::vector<DsfPoint> &out, bool &found, std::string &err);
void dsf_elastic_tables(const std::vector<DsfPoint> &D, MFP &Total, std::vector<double> &Emit, std::vector<double> &Absorb);

// ---- input (input.cpp)
void set_default_numpar(NumPar &np);
bool read_case(const std::string &dir, Case &c, std::string &err);

// ---- EADL2023.ALL (ENDL format) for what a .cdf leaves out (eadl.cpp; Dealing_with_EADL.f90:312-742)
struct Eadl {
    struct Row { double des, val; };
    struct Block { int Z = 0, C = 0, I = 0, S = 0; std::vector<Row> rows; };
    std::vector<Block> blocks;
    bool load(const std::string &path, std::string &err);
    const Block *find(int Z, int I) const;
    bool has_element(int Z) const;
    bool real_value(int Z, int I, int designator, double &out) const;   // READ_EADL_TYPE_FILE_real, [eV]
    bool electrons(int Z, int designator, double &out) const;           // READ_EADL_TYPE_FILE_int (I = 912)
};
void eadl_select_imin_imax(int designator, int &imin, int &imax);
int eadl_next_designator(int last_designator);
void eadl_check_shell(const Eadl &db, Atom &a, int k, bool include_photons, std::vector<std::string> &warnings);

// ---- physics of the table builder (cdf.cpp)
struct CtxFlat {        // storage behind the trk3_dcs_ctx handed to the shared integrands / the GPU evaluator
    std::vector<double> E0, A, G, scr;
    std::vector<int32_t> off;
    trk3_dcs_ctx d{};
};
// Requests of ONE outer integration (TotIMFP / Tot_EMFP / SHI_TotIMFP call): see dcs_request in cdf.cpp
struct DcsBatch {
    enum { DIRECT = 0, RECORD = 1, REPLAY = 2 };
    int mode = DIRECT;
    trk3_dcs_task task{};
    std::vector<double> hw;        // RECORD: transferred energies asked for, in call order
    const double *val = nullptr;   // REPLAY: their values
    size_t cursor = 0;
};
struct Ctx {            // immutable view used by the integrators
    const Case *c;
    const std::vector<double> *k = nullptr, *effm = nullptr;   // DOS (or inverted DOS for metals)
    bool mass_from_dos = false;
    std::shared_ptr<CtxFlat> flat;          // oscillator sets flattened: set0[atom] + shell, phonon CDF = set_phonon
    std::vector<int> set0;
    int set_phonon = 0;
    DcsBatch *batch = nullptr;              // null: integrate directly
};
double dcs_request(const Ctx &x, const trk3_dcs_task &t, double hw);
Ctx make_ctx(const Case &c);
void get_single_pole(Case &c);                                   // Cross_sections.f90:554
void sumrules(const CDFosc &o, double &ksum, double &fsum, double x_min, double Omega);   // :728
double w_plasma(double At_dens, double Mass = -1.0);             // :712
double equilibrium_charge_target(double Ekin, double Mass, double ZSHI, double Zmean, int Kind_Zeff, double fixed_Zeff); // :2601
void equilibrium_charge_SHI(Ion &shi, const std::vector<Atom> &atoms);      // :2641
double define_dE(int CS_method, int n, double E, bool has_min, double E0_min, bool has_max, double E0_max, double dE_min_use); // :1374
// per-point integrators; `row` (may be null) receives the cumulative table of this grid point
void TotIMFP(const Ctx &x, double Ele, int Nat, int Nshl, int kind /*0 e,1 h*/, double &Sigma, double &dEdx, DiffRow *row);  // :881
void Tot_EMFP(const Ctx &x, double Ele, int kind, double Zeff, double &Sigma, double &dEdx, DiffRow *row);                   // :2966
void Elastic_cross_section(const Ctx &x, double Ee, int kind, double &EMFP, double &dEdx, DiffRow *row);                  // :2892
void Tot_Phot_IMFP(const Ctx &x, double Ele, int Nat, int Nshl, double &Sigma, double &dEdx);                            // :833
void SHI_TotIMFP(const Ctx &x, Ion &shi, int Nat, int Nshl, double &Sigma, double &dEdx, MFP *dSedE);                     // :2452
void SHI_TotIMFP_BK(const Ctx &x, Ion &shi, int Nat, int Nshl, double &Sigma, double &dEdx);                              // :2748 (Brandt-Kitagawa ion)
void SHI_Total_IMFP(const Ctx &x, Ion &shi, int Nat, int Nshl, double &Sigma, double &dEdx);                              // :2431 (dispatch on Kind_ion)

// ---- table drivers (tables.cpp)
std::vector<double> get_grid_4CS(const std::vector<Atom> &atoms, double Emin, double Emax);   // Analytical_IMFPs.f90:2815
struct BuildOptions {
    int threads = 0;              // 0 => all
    bool shi_window_only = false; // only the SHI grid points the MC can touch (test speed-up)
    bool verbose = false;
    trk3_dcs_eval_fn evaluator = nullptr;   // of the integrands (trk3h_set_dcs_evaluator); null: integrate directly on the host
};
bool build_tables(Case &c, const BuildOptions &opt, std::string &err);     // MAIN.f90:146-247
bool finish_tables(Case &c, const BuildOptions &opt, std::string &err);    // MAIN.f90:231-247 (never cached by the reference)
void find_VB_numbers(Case &c);                                             // Reading_files...:3252
void radius_for_distributions(Case &c);                                    // Sorting_output_data.f90:1363
std::vector<double> set_time_grid(double Tim, double dt, int dt_flag);     // Monte_Carlo.f90:2118
int  count_time_points(double Tim, double dt, int dt_flag);               // Sorting_output_data.f90:1375-1386

// ---- binary cache of built tables (ours)
bool save_tables_bin(const Case &c, const std::string &path, std::string &err);
bool load_tables_bin(Case &c, const std::string &path, std::string &err);

// ---- the reference's own on-disk table cache (refcache.cpp): OUTPUT_<material>/OUTPUT_*_IMFPs_*.dat, *_EMFPs_*.dat,
// diff_CS/*.dat, OUTPUT_<ion>_in_<material>/OUTPUT_<ion>_*_{IMFP,dEdx,effective_charges,Range}.dat
// (Analytical_IMFPs.f90:262-291, 385-408, 506-526, 590-786, 917-1606, 2306-2349, 2462-2506, 2657-2707)
struct RefCacheNames {              // file names without directories, as the reference composes them
    std::string dir_material, dir_ion, dir_diff;             // OUTPUT_<mat>, OUTPUT_<mat>/OUTPUT_<ion>_in_<mat>, OUTPUT_<mat>/diff_CS
    std::string el_imfp, hole_imfp, photon_imfp, el_emfp, hole_emfp, shi_stem;
};
RefCacheNames reference_cache_names(const Case &c);
std::string diff_cs_file_name(const std::string &table_file, const std::string &atom, const std::string &shell, double E);
bool write_reference_cache(const Case &c, const std::string &out_root, int *n_files, std::string &err);
bool read_reference_cache(Case &c, const std::string &out_root, const BuildOptions &opt, std::string &err);

// ---- flattening into the C ABI (pack.cpp)
struct Packed {
    trk3_config cfg{};
    trk3_tables tab{};
    std::vector<double> ei_E, ei_L, ee_E, ee_L, hi_E, hi_L, he_E, he_L, ph_E, ph_L, shi_E, shi_L, shi_dEdx;
    std::vector<int64_t> dshi_off, eid_off, eed_off, hid_off, hed_off;
    std::vector<double> dshi_E, dshi_L, eid_hw, eid_L, eed_hw, eed_L, hid_hw, hid_L, hed_hw, hed_L;
    std::vector<double> dos_E, dos_DOS, dos_int, dos_effm, out_R, out_V, osc_E0, osc_alpha;
    std::vector<double> dsf_e_dE, dsf_e_emit, dsf_e_absorb, ee_emit, ee_absorb, dsf_h_dE, dsf_h_emit, dsf_h_absorb, he_emit, he_absorb;
};
void pack_case(const Case &c, Packed &p);
double define_alpha(double Ai, double Gammai, double E0i, double x_min);       // cdf.cpp

// ---- output files (output.cpp): Save_output, Sorting_output_data.f90:340-1140
bool save_output(const Case &c, const trk3_tally_layout &lay, const double *tallies_sum, int NMC,
                 const std::string &out_root, std::string &out_dir, std::string &err);

}  // namespace trk3
