// tables.cpp -- energy grids and table drivers, restating Analytical_IMFPs.f90 (grids :2815-3101,
// All_shells_Electron_MFP :1976, All_elastic_scattering :1830, All_shells_Photon_MFP :2146,
// Analytical_SHI_dEdx :2510) and the table-related part of Universal_MC_for_SHI_MAIN.f90:146-247.
// Grid points are independent, so they are distributed over OpenMP threads exactly as the
// reference does (`!$omp do schedule(dynamic)`).
#include "trk3_host.hpp"
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <atomic>
#include <functional>
#include <thread>

namespace trk3 {

namespace {
// Dynamic-schedule parallel loop over independent grid points (the reference uses
// `!$omp do schedule(dynamic)`, Analytical_IMFPs.f90:1957, 2120, 2636).
void parallel_for(int n, int nth, const std::function<void(int)> &fn) {
    if (nth <= 1 || n <= 1) { for (int i = 0; i < n; ++i) fn(i); return; }
    std::atomic<int> next(0);
    std::vector<std::thread> th;
    for (int t = 0; t < std::min(nth, n); ++t)
        th.emplace_back([&]() { for (;;) { int i = next.fetch_add(1); if (i >= n) break; fn(i); } });
    for (auto &t : th) t.join();
}

// The outer integrations of a table (one per task: grid point x shell), each calling body(t, ctx).  Without an evaluator
// they run directly on the host threads.  With one (the GPU: trk3_dcs_eval) every task is first run in RECORD mode, which
// only walks its outer loop and notes the q-integrals it needs, then all of them are evaluated in one go, and the tasks
// are run again in REPLAY mode with the values (see dcs_request, cdf.cpp): same operations, same order, same result.
bool run_integrations(const Ctx &x0, int n, int nth, trk3_dcs_eval_fn ev, std::string &err, const std::function<void(int, const Ctx &)> &body) {
    if (!ev) { parallel_for(n, nth, [&](int t) { body(t, x0); }); return true; }
    std::vector<DcsBatch> b((size_t)n);
    parallel_for(n, nth, [&](int t) { Ctx x = x0; b[t].mode = DcsBatch::RECORD; x.batch = &b[t]; body(t, x); });
    const size_t chunk_max = (size_t)8 << 20;             // requests per evaluator call (bounds the staging memory)
    std::vector<std::vector<double>> val((size_t)n);
    std::vector<trk3_dcs_task> tasks; std::vector<double> hw, out; std::vector<int32_t> task_of;
    for (int t0 = 0; t0 < n;) {
        tasks.clear(); hw.clear(); task_of.clear();
        int t1 = t0;
        while (t1 < n && (hw.empty() || hw.size() + b[t1].hw.size() <= chunk_max)) {
            tasks.push_back(b[t1].task);
            hw.insert(hw.end(), b[t1].hw.begin(), b[t1].hw.end());
            task_of.insert(task_of.end(), b[t1].hw.size(), (int32_t)(t1 - t0));
            ++t1;
        }
        out.assign(hw.size(), 0.0);
        if (!hw.empty()) {
            const int rc = ev(&x0.flat->d, tasks.data(), (int64_t)tasks.size(), hw.data(), task_of.data(), (int64_t)hw.size(), out.data());
            if (rc != TRK3_OK) { err = "the evaluator of the table integrands failed (code " + std::to_string(rc) + ")"; return false; }
        }
        size_t o = 0;
        for (int t = t0; t < t1; ++t) { val[t].assign(out.begin() + o, out.begin() + o + b[t].hw.size()); o += b[t].hw.size(); }
        t0 = t1;
    }
    parallel_for(n, nth, [&](int t) {
        Ctx x = x0;
        b[t].mode = DcsBatch::REPLAY; b[t].val = val[t].data(); b[t].cursor = 0;
        std::vector<double>().swap(b[t].hw);
        x.batch = &b[t]; body(t, x);
    });
    return true;
}

// find_order_of_number_real, Analytical_IMFPs.f90:3077: number of decimal digits of ceiling(num)
int find_order_of_number(double num) {
    char buf[64];
    if (num > 1e9) {
        double r = num / 1.0e8;
        std::snprintf(buf, sizeof buf, "%lld", (long long)std::ceil(r));
        return 8 + (int)std::strlen(buf);
    }
    std::snprintf(buf, sizeof buf, "%lld", (long long)std::ceil(num));
    return (int)std::strlen(buf);
}

void bubble_sort(std::vector<double> &a) {        // sort_array_r :3035 (result == any stable sort)
    int N = (int)a.size();
    for (int j = N - 1; j >= 1; --j) {
        bool swapped = false;
        for (int i = 0; i < j; ++i) if (a[i] > a[i + 1]) { std::swap(a[i], a[i + 1]); swapped = true; }
        if (!swapped) break;
    }
}

std::vector<double> special_points(const std::vector<Atom> &atoms) {   // define_special_points :2953 + exclude_doubles :2990
    std::vector<double> sp;
    for (auto &a : atoms) for (double ip : a.Ip) sp.push_back(ip);
    bubble_sort(sp);
    int Nsiz = (int)sp.size();
    std::vector<int> ind(Nsiz + 1, 0);
    int coun_ind = 0;
    for (int i = 1; i <= Nsiz - 1; ++i)
        for (int k = i + 1; k <= Nsiz; ++k)
            if (sp[i - 1] == sp[k - 1]) { if (coun_ind < Nsiz) ind[coun_ind++] = std::min(i, k); }
    int Ndub = 0; for (int i = 0; i < Nsiz; ++i) if (ind[i] > 0) ++Ndub;
    std::vector<double> out; out.reserve(Nsiz - Ndub);
    int ci = 0;
    for (int i = 1; i <= Nsiz; ++i) {
        if (ci < Nsiz && i == ind[ci]) ++ci;
        else if ((int)out.size() < Nsiz - Ndub) out.push_back(sp[i - 1]);
    }
    out.resize(std::max(0, Nsiz - Ndub), 0.0);
    return out;
}

// go_thru_grid :2866-2949; when `arr` is null only counts
int go_thru_grid(double Emin, double Emax, double E_sp_eps, const std::vector<double> &sp, double scale_dE, std::vector<double> *arr) {
    int SP_count = 1, N = 0;
    double dE_min = 0.01 * scale_dE;
    double E_cur = Emin - dE_min, dE;
    int size = arr ? (int)arr->size() : 0;
    while (E_cur < Emax) {
        N = N + 1;
        if (E_cur < 0.1 - dE_min * 0.5) dE = dE_min;
        else if (E_cur < 10.0 - dE_min * 0.5) dE = dE_min * 10.0;
        else if (E_cur < 100.0 - dE_min * 0.5) dE = dE_min * 100.0;
        else dE = std::pow(10.0, find_order_of_number(E_cur) - 2);
        E_cur = E_cur + dE;
        if (arr) { if (N <= size) (*arr)[N - 1] = E_cur; else return N; }
        bool here = false;
        if (SP_count <= (int)sp.size()) {
            if (E_cur >= sp[SP_count - 1] && (E_cur - dE) < sp[SP_count - 1]) { SP_count++; here = true; }
        }
        if (here) {
            if (arr) {
                if (N + 2 <= size) {
                    N = N + 1; (*arr)[N - 1] = sp[SP_count - 2] - E_sp_eps;
                    N = N + 1; (*arr)[N - 1] = sp[SP_count - 2] + E_sp_eps;
                } else return N;
            } else N = N + 2;
        }
    }
    if (arr) bubble_sort(*arr);
    return N;
}
}  // namespace

std::vector<double> get_grid_4CS(const std::vector<Atom> &atoms, double Emin, double Emax) {
    auto sp = special_points(atoms);
    int N = go_thru_grid(Emin, Emax, 1.0e-3, sp, 1.0, nullptr);
    std::vector<double> g(N, 0.0);
    go_thru_grid(Emin, Emax, 1.0e-3, sp, 1.0, &g);
    return g;
}

void find_VB_numbers(Case &c) {     // Reading_files_and_parameters.f90:3252-3270
    double temp = 1e10;
    for (int i = 0; i < (int)c.atoms.size(); ++i) {
        const Atom &a = c.atoms[i];
        for (int j = 0; j < a.nshl(); ++j) {
            if (a.Ip[j] < temp || a.Shl_num[j] >= 63) { temp = a.Ip[j]; c.Lowest_Ip_At = i; c.Lowest_Ip_Shl = j; }
            if (a.Shl_num[j] >= 63) break;
        }
    }
}

int count_time_points(double Tim, double dt, int dt_flag) {   // Sorting_output_data.f90:1375-1386
    if (dt_flag <= 0) return (int)std::ceil(Tim / dt);
    int i = 0; double t = 0.01;
    while (t < Tim) { ++i; t = t * dt; }
    return i + 1;
}

std::vector<double> set_time_grid(double Tim, double dt, int dt_flag) {   // Monte_Carlo.f90:2118-2149
    std::vector<double> g;
    if (dt_flag <= 0) {
        int n = (int)std::ceil(Tim / dt) + 1;
        g.assign(n, 0.0);
        g[0] = dt;
        for (int i = 1; i <= n - 2; ++i) g[i] = g[i - 1] + dt;
        g[n - 1] = Tim + dt;
    } else {
        int i = 0; double t = 0.01;
        while (t <= Tim) { ++i; t = t * dt; }
        g.assign(i + 1, 0.0);
        g[0] = 0.01;
        for (int k = 1; k <= (int)g.size() - 2; ++k) g[k] = g[k - 1] * dt;
        g[g.size() - 1] = Tim + dt;
    }
    return g;
}

void radius_for_distributions(Case &c) {      // Sorting_output_data.f90:1403-1434
    const int N = 50;
    c.Out_R.assign(N, 0.0); c.Out_V.assign(N, 0.0);
    double R = 0.0, R0 = 0.0;
    for (int i = 1; i <= N; ++i) {
        if (i < 11) R = (double)i;
        else if (i < 21) R = (i == 11) ? (double)((i - 10) * 15) : (double)((i - 10) * 10);
        else if (i < 31) R = (i == 21) ? (double)((i - 20) * 150) : (double)((i - 20) * 100);
        else if (i < 41) R = (i == 31) ? (double)((i - 30) * 1500) : (double)((i - 30) * 1000);
        else if (i == 41) R = (double)((i - 40) * 15000);
        else R = (double)((i - 40) * 10000);
        c.Out_R[i - 1] = R;
        c.Out_V[i - 1] = 1.0 / ((R * R - R0 * R0) * g_Pi * c.Matter.Layer);
        R0 = R;
    }
}

bool build_tables(Case &c, const BuildOptions &opt, std::string &err) {
    const trk3_dcs_eval_fn ev = opt.evaluator;
    int nth = opt.threads > 0 ? opt.threads : (int)std::thread::hardware_concurrency();
    if (nth < 1) nth = 1;
    get_single_pole(c);                                   // MAIN.f90:146
    Ctx x = make_ctx(c);
    const int Nat = (int)c.atoms.size();

    // ---- SHI: Analytical_ion_dEdx (:2242) -> Analytical_SHI_dEdx (:2510)
    {
        double M = c.SHI.Mass * g_Mp;
        double Emin = std::ceil((M + g_me) * (M + g_me) / (M * g_me) * c.atoms[0].Ip.back() / 4.0);
        double Emax = 175.6e6 / 2.0 * c.SHI.Mass;
        std::vector<double> grid = get_grid_4CS(c.atoms, Emin, Emax);
        const int N = (int)grid.size();
        c.SHI_MFP.assign(Nat, {});
        for (int j = 0; j < Nat; ++j) {
            c.SHI_MFP[j].assign(c.atoms[j].nshl(), MFP{});
            for (auto &m : c.SHI_MFP[j]) { m.E = grid; m.L.assign(N, 1.0e24); m.dEdx.assign(N, 0.0); }
        }
        std::vector<int> todo;
        for (int j = 0; j < N; ++j) {
            if (opt.shi_window_only) {
                bool in = grid[j] >= 0.5 * c.SHI.E && (j == 0 || grid[j - 1] <= c.SHI.E);
                if (!in) continue;
            }
            todo.push_back(j);
        }
        if (opt.verbose) std::fprintf(stderr, "[trk3] SHI table: %d of %d grid points\n", (int)todo.size(), N);
        // one task per (grid point, shell) for load balance; results are independent
        struct Task { int j, a, s; };
        std::vector<Task> tasks;
        for (int j : todo) for (int a = 0; a < Nat; ++a) for (int s = 0; s < c.atoms[a].nshl(); ++s) tasks.push_back({j, a, s});
        if (!run_integrations(x, (int)tasks.size(), nth, ev, err, [&](int t, const Ctx &x) {
            Ion shi1 = c.SHI;
            shi1.E = grid[tasks[t].j];
            double IMFP, dEdx;
            SHI_Total_IMFP(x, shi1, tasks[t].a, tasks[t].s, IMFP, dEdx);
            MFP &m = c.SHI_MFP[tasks[t].a][tasks[t].s];
            double L = (IMFP > 1.0e-10) ? 1.0 / IMFP : 1.0e28;     // :2644-2648
            if (L > 1e30) L = 1e30;                                 // :2682
            m.L[tasks[t].j] = L; m.dEdx[tasks[t].j] = dEdx;
        })) return false;
    }
    equilibrium_charge_SHI(c.SHI, c.atoms);               // MAIN.f90:171

    // ---- electrons, holes: Analytical_electron_dEdx (:100)
    const double Emax_e = 175.6e6 * 2.0 / 1836.0;
    for (int kind = 0; kind < 2; ++kind) {
        double Emax = (kind == 0) ? Emax_e : (*std::max_element(c.dos.E.begin(), c.dos.E.end()) + 1.0);
        std::vector<double> grid = get_grid_4CS(c.atoms, 0.1, Emax);
        std::vector<double> gridE = get_grid_4CS(c.atoms, 0.01, Emax);
        const int N = (int)grid.size(), Ne = (int)gridE.size();
        auto &T = (kind == 0) ? c.Total_el_MFPs : c.Total_Hole_MFPs;
        T.assign(Nat, {});
        for (int j = 0; j < Nat; ++j) {
            T[j].assign(c.atoms[j].nshl(), MFP{});
            for (auto &m : T[j]) { m.E = grid; m.L.assign(N, 0.0); m.dEdx.assign(N, 0.0); }
        }
        if (kind == 0) {
            c.EIdCS.assign(Nat, {});
            for (int j = 0; j < Nat; ++j) {
                c.EIdCS[j].assign(c.atoms[j].nshl(), DiffCS{});
                for (auto &d : c.EIdCS[j]) { d.E = grid; d.row.assign(N, DiffRow{}); }
            }
        } else { c.HIdCS.E = grid; c.HIdCS.row.assign(N, DiffRow{}); }
        struct Task { int i, a, s; };
        std::vector<Task> tasks;
        for (int i = 0; i < N; ++i) for (int a = 0; a < Nat; ++a) for (int s = 0; s < c.atoms[a].nshl(); ++s) tasks.push_back({i, a, s});
        if (opt.verbose) std::fprintf(stderr, "[trk3] %s inelastic table: %d points x %d shells\n", kind ? "hole" : "electron", N, c.n_shells());
        if (!run_integrations(x, (int)tasks.size(), nth, ev, err, [&](int t, const Ctx &x) {
            const Task &k = tasks[t];
            DiffRow *row = nullptr;
            if (kind == 0) row = &c.EIdCS[k.a][k.s].row[k.i];
            else if (k.a == 0 && k.s == c.atoms[0].nshl() - 1) row = &c.HIdCS.row[k.i];
            double S, dEdx;
            TotIMFP(x, grid[k.i], k.a, k.s, kind, S, dEdx, row);
            T[k.a][k.s].L[k.i] = S; T[k.a][k.s].dEdx[k.i] = dEdx;
        })) return false;
        // elastic
        MFP &El = (kind == 0) ? c.Elastic_MFP : c.Elastic_Hole_MFP;
        DiffCS &Ed = (kind == 0) ? c.EEdCS : c.HEdCS;
        El.E = gridE; El.L.assign(Ne, 0.0); El.dEdx.assign(Ne, 0.0);
        Ed.E = gridE; Ed.row.assign(Ne, DiffRow{});
        if (opt.verbose) std::fprintf(stderr, "[trk3] %s elastic table: %d points\n", kind ? "hole" : "electron", Ne);
        if (c.numpar.kind_of_EMFP == 1 || c.numpar.kind_of_EMFP == 0) {
            if (!run_integrations(x, Ne, nth, ev, err, [&](int i, const Ctx &x) {
                double S, dEdx;
                Elastic_cross_section(x, gridE[i], kind, S, dEdx, (c.numpar.kind_of_EMFP == 1) ? &Ed.row[i] : nullptr);
                El.L[i] = S; El.dEdx[i] = dEdx;
            })) return false;
        } else if (c.numpar.kind_of_EMFP == 2) {
            // DSF: the elastic tables live on the DSF file's own energy grid (Analytical_IMFPs.f90:885-919); no differential rows --
            // the transferred energy is sampled from the DSF rows themselves (NRG_transfer_elastic_DSF)
            std::vector<double> emit, absorb;
            dsf_elastic_tables(kind == 0 ? c.DSF_DEMFP : c.DSF_DEMFP_H, El, emit, absorb);
            Ed.E = El.E; Ed.row.assign(El.E.size(), DiffRow{});
        } else {
            // kind_of_EMFP = -1: elastic scattering disabled (Analytical_IMFPs.f90 'No_elas'): infinite MFP
            for (int i = 0; i < Ne; ++i) { El.L[i] = 1.0e30; El.dEdx[i] = 0.0; }
        }
    }
    // ---- photons
    if (c.numpar.include_photons) {
        std::vector<double> grid = get_grid_4CS(c.atoms, 0.1, Emax_e);
        const int N = (int)grid.size();
        c.Total_Photon_MFPs.assign(Nat, {});
        for (int j = 0; j < Nat; ++j) {
            c.Total_Photon_MFPs[j].assign(c.atoms[j].nshl(), MFP{});
            for (int s = 0; s < c.atoms[j].nshl(); ++s) {
                MFP &m = c.Total_Photon_MFPs[j][s];
                m.E = grid; m.L.assign(N, 0.0); m.dEdx.assign(N, 0.0);
                for (int i = 0; i < N; ++i) Tot_Phot_IMFP(x, grid[i], j, s, m.L[i], m.dEdx[i]);
            }
        }
    } else c.Total_Photon_MFPs.clear();

    return finish_tables(c, opt, err);
}

// What MAIN.f90 computes after the (possibly cached) tables: the differential SHI MFP for the given ion energy
// (:231-238), the valence-band indices and the radial grid.  Shared by build_tables and read_reference_cache.
bool finish_tables(Case &c, const BuildOptions &opt, std::string &err) {
    const trk3_dcs_eval_fn ev = opt.evaluator;
    int nth = opt.threads > 0 ? opt.threads : (int)std::thread::hardware_concurrency();
    if (nth < 1) nth = 1;
    Ctx x = make_ctx(c);
    const int Nat = (int)c.atoms.size();
    c.diff_SHI_MFP.assign(Nat, {});
    for (int j = 0; j < Nat; ++j) c.diff_SHI_MFP[j].assign(c.atoms[j].nshl(), MFP{});
    {
        struct Task { int a, s; };
        std::vector<Task> tasks;
        for (int a = 0; a < Nat; ++a) for (int s = 0; s < c.atoms[a].nshl(); ++s) tasks.push_back({a, s});
        if (!run_integrations(x, (int)tasks.size(), nth, ev, err, [&](int t, const Ctx &x) {
            Ion shi1 = c.SHI;
            double S, dEdx;
            SHI_TotIMFP(x, shi1, tasks[t].a, tasks[t].s, S, dEdx, &c.diff_SHI_MFP[tasks[t].a][tasks[t].s]);
        })) return false;
    }
    find_VB_numbers(c);
    radius_for_distributions(c);
    c.tables_built = true;
    return true;
}

// ---------------------------------------------------------------------------------------------
// binary cache (ours): magic, then length-prefixed double vectors in a fixed traversal order
// ---------------------------------------------------------------------------------------------
namespace {
void wv(std::ofstream &f, const std::vector<double> &v) { uint64_t n = v.size(); f.write((const char *)&n, 8); if (n) f.write((const char *)v.data(), n * 8); }
bool rv(std::ifstream &f, std::vector<double> &v) { uint64_t n = 0; f.read((char *)&n, 8); if (!f || n > (1ull << 32)) return false; v.resize(n); if (n) f.read((char *)v.data(), n * 8); return (bool)f; }
void wmfp(std::ofstream &f, const MFP &m) { wv(f, m.E); wv(f, m.L); wv(f, m.dEdx); }
bool rmfp(std::ifstream &f, MFP &m) { return rv(f, m.E) && rv(f, m.L) && rv(f, m.dEdx); }
void wd(std::ofstream &f, const DiffCS &d) { wv(f, d.E); uint64_t n = d.row.size(); f.write((const char *)&n, 8); for (auto &r : d.row) { wv(f, r.hw); wv(f, r.L); } }
bool rd(std::ifstream &f, DiffCS &d) { if (!rv(f, d.E)) return false; uint64_t n = 0; f.read((char *)&n, 8); if (!f || n > (1u << 24)) return false; d.row.resize(n); for (auto &r : d.row) if (!rv(f, r.hw) || !rv(f, r.L)) return false; return true; }
const char MAGIC[8] = {'T', 'R', 'K', '3', 'T', 'B', '0', '2'};
}  // namespace

bool save_tables_bin(const Case &c, const std::string &path, std::string &err) {
    std::ofstream f(path, std::ios::binary);
    if (!f) { err = "cannot write " + path; return false; }
    f.write(MAGIC, 8);
    uint64_t nat = c.atoms.size(); f.write((const char *)&nat, 8);
    for (size_t a = 0; a < nat; ++a) {
        uint64_t ns = c.atoms[a].nshl(); f.write((const char *)&ns, 8);
        for (size_t s = 0; s < ns; ++s) {
            wmfp(f, c.SHI_MFP[a][s]); wmfp(f, c.diff_SHI_MFP[a][s]); wmfp(f, c.Total_el_MFPs[a][s]); wmfp(f, c.Total_Hole_MFPs[a][s]);
            uint64_t hp = c.Total_Photon_MFPs.empty() ? 0 : 1; f.write((const char *)&hp, 8);
            if (hp) wmfp(f, c.Total_Photon_MFPs[a][s]);
            wd(f, c.EIdCS[a][s]);
        }
    }
    wmfp(f, c.Elastic_MFP); wmfp(f, c.Elastic_Hole_MFP); wd(f, c.EEdCS); wd(f, c.HIdCS); wd(f, c.HEdCS);
    wv(f, c.CDF_Phonon.E0); wv(f, c.CDF_Phonon.A); wv(f, c.CDF_Phonon.Gamma);
    return (bool)f;
}

bool load_tables_bin(Case &c, const std::string &path, std::string &err) {
    std::ifstream f(path, std::ios::binary);
    if (!f) { err = "cannot read " + path; return false; }
    char mg[8]; f.read(mg, 8);
    if (!f || std::memcmp(mg, MAGIC, 8) != 0) { err = "bad table cache " + path; return false; }
    uint64_t nat = 0; f.read((char *)&nat, 8);
    if (nat != c.atoms.size()) { err = "table cache does not match the material"; return false; }
    c.SHI_MFP.assign(nat, {}); c.diff_SHI_MFP.assign(nat, {}); c.Total_el_MFPs.assign(nat, {}); c.Total_Hole_MFPs.assign(nat, {});
    c.Total_Photon_MFPs.clear(); c.EIdCS.assign(nat, {});
    bool any_ph = false;
    std::vector<std::vector<MFP>> ph(nat);
    for (size_t a = 0; a < nat; ++a) {
        uint64_t ns = 0; f.read((char *)&ns, 8);
        if (ns != (uint64_t)c.atoms[a].nshl()) { err = "table cache does not match the material"; return false; }
        c.SHI_MFP[a].resize(ns); c.diff_SHI_MFP[a].resize(ns); c.Total_el_MFPs[a].resize(ns); c.Total_Hole_MFPs[a].resize(ns);
        c.EIdCS[a].resize(ns); ph[a].resize(ns);
        for (size_t s = 0; s < ns; ++s) {
            if (!rmfp(f, c.SHI_MFP[a][s]) || !rmfp(f, c.diff_SHI_MFP[a][s]) || !rmfp(f, c.Total_el_MFPs[a][s]) || !rmfp(f, c.Total_Hole_MFPs[a][s])) { err = "truncated table cache"; return false; }
            uint64_t hp = 0; f.read((char *)&hp, 8);
            if (hp) { any_ph = true; if (!rmfp(f, ph[a][s])) { err = "truncated table cache"; return false; } }
            if (!rd(f, c.EIdCS[a][s])) { err = "truncated table cache"; return false; }
        }
    }
    if (any_ph) c.Total_Photon_MFPs = ph;
    if (!rmfp(f, c.Elastic_MFP) || !rmfp(f, c.Elastic_Hole_MFP) || !rd(f, c.EEdCS) || !rd(f, c.HIdCS) || !rd(f, c.HEdCS)) { err = "truncated table cache"; return false; }
    if (!rv(f, c.CDF_Phonon.E0) || !rv(f, c.CDF_Phonon.A) || !rv(f, c.CDF_Phonon.Gamma)) { err = "truncated table cache"; return false; }
    if (c.numpar.include_photons && c.Total_Photon_MFPs.empty()) { err = "table cache has no photon tables"; return false; }
    equilibrium_charge_SHI(c.SHI, c.atoms);
    find_VB_numbers(c);
    radius_for_distributions(c);
    c.tables_built = true;
    return true;
}

}  // namespace trk3
