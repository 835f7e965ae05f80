// host_api.cpp -- extern "C" surface of libtrekis3_host.so (include/trekis3_host.h).
#include "../../../include/trekis3_host.h"
#include "trk3_host.hpp"
#include "../common/trk3_dcs.h"
#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstring>
#include <memory>
#include <thread>
#include <vector>

using namespace trk3;

struct trk3h_case {
    Case c;
    Packed p;
    bool packed = false;
    uint64_t seed = 20260101ull;
};

namespace {
void set_err(char *err, int errlen, const std::string &s) {
    if (err && errlen > 0) { std::snprintf(err, (size_t)errlen, "%s", s.c_str()); }
}
void repack(trk3h_case *h) {
    if (!h->c.tables_built) { find_VB_numbers(h->c); radius_for_distributions(h->c); }
    if (h->c.tables_built) { pack_case(h->c, h->p); h->p.cfg.seed = h->seed; h->packed = true; }
}
}  // namespace

extern "C" {

const char *trk3_host_version(void) { return "trekis3_host 0.1 (C++ host: readers + CDF table builder + output writer)"; }

trk3h_case *trk3h_load(const char *dir, char *err, int errlen) {
    auto h = std::make_unique<trk3h_case>();
    std::string e;
    if (!read_case(dir ? dir : ".", h->c, e)) { set_err(err, errlen, e); return nullptr; }
    return h.release();
}

void trk3h_free(trk3h_case *c) { delete c; }

static trk3_dcs_eval_fn g_dcs_evaluator = nullptr;
void trk3h_set_dcs_evaluator(trk3_dcs_eval_fn fn) { g_dcs_evaluator = fn; }

// the evaluator interface on the host threads (tests of the record / replay machinery without a GPU)
int trk3h_dcs_eval_host(const trk3_dcs_ctx *ctx, const trk3_dcs_task *tasks, int64_t n_tasks, const double *hw, const int32_t *task_of, int64_t n, double *out) {
    if (!ctx || !tasks || !hw || !task_of || !out || n < 0) return TRK3_E_INVALID;
    int nth = (int)std::thread::hardware_concurrency(); if (nth < 1) nth = 1;
    std::atomic<int64_t> next(0);
    std::atomic<int> bad(0);
    std::vector<std::thread> th;
    for (int t = 0; t < nth; ++t) th.emplace_back([&]() {
        for (;;) {
            const int64_t i0 = next.fetch_add(256);
            if (i0 >= n) break;
            for (int64_t i = i0; i < std::min<int64_t>(n, i0 + 256); ++i) {
                if (task_of[i] < 0 || task_of[i] >= n_tasks) { bad = 1; continue; }
                out[i] = trk3dcs::eval_request(*ctx, tasks[task_of[i]], hw[i]);
            }
        }
    });
    for (auto &t : th) t.join();
    return bad ? TRK3_E_INVALID : TRK3_OK;
}

int trk3h_build_tables(trk3h_case *h, int threads, int shi_window_only, int verbose, char *err, int errlen) {
    if (!h) return TRK3_E_INVALID;
    BuildOptions o; o.threads = threads; o.shi_window_only = shi_window_only != 0; o.verbose = verbose != 0;
    o.evaluator = g_dcs_evaluator;
    std::string e;
    if (!build_tables(h->c, o, e)) { set_err(err, errlen, e); return TRK3_E_INVALID; }
    h->packed = false;
    return TRK3_OK;
}

int trk3h_save_tables(trk3h_case *h, const char *path, char *err, int errlen) {
    std::string e;
    if (!h || !h->c.tables_built) { set_err(err, errlen, "tables not built"); return TRK3_E_INVALID; }
    if (!save_tables_bin(h->c, path, e)) { set_err(err, errlen, e); return TRK3_E_INVALID; }
    return TRK3_OK;
}

int trk3h_load_tables(trk3h_case *h, const char *path, char *err, int errlen) {
    std::string e;
    if (!h) return TRK3_E_INVALID;
    if (!load_tables_bin(h->c, path, e)) { set_err(err, errlen, e); return TRK3_E_INVALID; }
    h->packed = false;
    return TRK3_OK;
}

int trk3h_write_reference_cache(trk3h_case *h, const char *out_root, int *n_files, char *err, int errlen) {
    std::string e;
    if (!h || !h->c.tables_built) { set_err(err, errlen, "tables not built"); return TRK3_E_INVALID; }
    if (!write_reference_cache(h->c, out_root ? out_root : ".", n_files, e)) { set_err(err, errlen, e); return TRK3_E_INVALID; }
    return TRK3_OK;
}

int trk3h_read_reference_cache(trk3h_case *h, const char *out_root, int threads, char *err, int errlen) {
    if (!h) return TRK3_E_INVALID;
    BuildOptions o; o.threads = threads; o.evaluator = g_dcs_evaluator;
    std::string e;
    h->c.tables_built = false;
    if (!read_reference_cache(h->c, out_root ? out_root : ".", o, e)) { set_err(err, errlen, e); return TRK3_E_INVALID; }
    h->packed = false;
    return TRK3_OK;
}

int trk3h_reference_cache_name(trk3h_case *h, const char *which, char *out, int outlen) {
    if (!h || !which || !out || outlen < 1) return TRK3_E_INVALID;
    const RefCacheNames n = reference_cache_names(h->c);
    const std::string w(which);
    const std::string *s = w == "dir_material" ? &n.dir_material : w == "dir_ion" ? &n.dir_ion : w == "dir_diff" ? &n.dir_diff
                         : w == "el_imfp" ? &n.el_imfp : w == "hole_imfp" ? &n.hole_imfp : w == "photon_imfp" ? &n.photon_imfp
                         : w == "el_emfp" ? &n.el_emfp : w == "hole_emfp" ? &n.hole_emfp : w == "shi_stem" ? &n.shi_stem : nullptr;
    if (!s) return TRK3_E_INVALID;
    std::snprintf(out, (size_t)outlen, "%s", s->c_str());
    return TRK3_OK;
}

const trk3_config *trk3h_config(trk3h_case *h) {
    if (!h || !h->c.tables_built) return nullptr;
    if (!h->packed) repack(h);
    return &h->p.cfg;
}
const trk3_tables *trk3h_tables(trk3h_case *h) {
    if (!h || !h->c.tables_built) return nullptr;
    if (!h->packed) repack(h);
    return &h->p.tab;
}

int trk3h_set(trk3h_case *h, const char *key, double v) {
    if (!h || !key) return TRK3_E_INVALID;
    std::string k(key);
    Case &c = h->c;
    if (k == "NMC") c.NMC = (int)v;
    else if (k == "seed") h->seed = (uint64_t)v;
    else if (k == "Tim") c.Tim = v;
    else if (k == "dt") c.dt = v;
    else if (k == "dt_flag") c.numpar.dt_flag = (int)v;
    else if (k == "cut_off") c.Matter.cut_off = v;
    else if (k == "layer") { c.Matter.Layer = v; if (c.tables_built) radius_for_distributions(c); }
    else if (k == "hole_mass") c.Matter.hole_mass = v;          // NB: tables depend on it; rebuild afterwards
    else if (k == "work_function") c.Matter.work_function = v;
    else if (k == "bar_length") c.Matter.bar_length = v;
    else if (k == "bar_height") c.Matter.bar_height = v;
    else if (k.rfind("radiat:", 0) == 0 || k.rfind("auger:", 0) == 0) {
        // "radiat:<atom>:<shell>" (0-based): test hook for the decay channels; value in fs
        int a = -1, s = -1;
        bool rad = k[0] == 'r';
        if (std::sscanf(k.c_str() + (rad ? 7 : 6), "%d:%d", &a, &s) != 2 || a < 0 || a >= (int)c.atoms.size() || s < 0 || s >= c.atoms[a].nshl()) return TRK3_E_INVALID;
        (rad ? c.atoms[a].Radiat[s] : c.atoms[a].Auger[s]) = v;
    }
    else return TRK3_E_INVALID;
    h->packed = false;
    return TRK3_OK;
}

double trk3h_get(trk3h_case *h, const char *key) {
    if (!h || !key) return 0.0;
    std::string k(key);
    const Case &c = h->c;
    if (k == "NMC") return c.NMC;
    if (k == "Tim") return c.Tim;
    if (k == "dt") return c.dt;
    if (k == "dt_flag") return c.numpar.dt_flag;
    if (k == "layer") return c.Matter.Layer;
    if (k == "At_Dens") return c.Matter.At_Dens;
    if (k == "Dens") return c.Matter.Dens;
    if (k == "N_VB_el") return c.Matter.N_VB_el;
    if (k == "shi_E") return c.SHI.E;
    if (k == "shi_mass") return c.SHI.Mass;
    if (k == "shi_Zeff") return c.SHI.Zeff;
    if (k == "include_photons") return c.numpar.include_photons ? 1 : 0;
    if (k == "kind_of_CDF_ph") return c.numpar.kind_of_CDF_ph;
    if (k == "kind_of_CDF") return c.numpar.kind_of_CDF;
    if (k.rfind("nshl:", 0) == 0) { int j = atoi(key + 5); return (j >= 0 && j < (int)c.atoms.size()) ? (double)c.atoms[(size_t)j].nshl() : std::nan(""); }
    if (k.rfind("Zat:", 0) == 0) { int j = atoi(key + 4); return (j >= 0 && j < (int)c.atoms.size()) ? (double)c.atoms[(size_t)j].Zat : std::nan(""); }
    if (k == "n_atoms") return (double)c.atoms.size();
    if (k == "n_dos") return (double)c.dos.E.size();
    if (k == "Num_th") return c.Num_th;
    if (k == "phonon_E0") return c.CDF_Phonon.E0.empty() ? 0.0 : c.CDF_Phonon.E0[0];
    if (k == "phonon_A") return c.CDF_Phonon.A.empty() ? 0.0 : c.CDF_Phonon.A[0];
    if (k == "phonon_Gamma") return c.CDF_Phonon.Gamma.empty() ? 0.0 : c.CDF_Phonon.Gamma[0];
    // "atom:<j>:<k>:<field>": a parameter of shell k of atom j (0-based), field = Nel | Ip | Ek | Auger | Radiat | Shl_num | PQN
    if (k.rfind("atom:", 0) == 0) {
        int j = -1, sh = -1; char fld[32] = "";
        if (std::sscanf(key, "atom:%d:%d:%31s", &j, &sh, fld) != 3 || j < 0 || j >= (int)c.atoms.size()) return std::nan("");
        const Atom &a = c.atoms[(size_t)j];
        if (sh < 0 || sh >= a.nshl()) return std::nan("");
        const std::string f(fld);
        if (f == "Nel") return a.Nel[(size_t)sh];
        if (f == "Ip") return a.Ip[(size_t)sh];
        if (f == "Ek") return a.Ek[(size_t)sh];
        if (f == "Auger") return a.Auger[(size_t)sh];
        if (f == "Radiat") return a.Radiat[(size_t)sh];
        if (f == "KOCS") return (double)a.KOCS[(size_t)sh];
        if (f == "Shl_num") return a.Shl_num[(size_t)sh];
        if (f == "PQN") return a.PQN[(size_t)sh];
        return std::nan("");
    }
    return 0.0;
}

int trk3h_get_string(trk3h_case *h, const char *key, char *out, int outlen) {
    if (!h || !key || !out || outlen <= 0) return TRK3_E_INVALID;
    std::string k(key), v;
    if (k == "material") v = h->c.Material_name;
    else if (k == "target_name") v = h->c.Matter.Target_name;
    else if (k == "chem") v = h->c.Matter.Chem;
    else if (k == "ion") v = h->c.SHI.Name;
    else if (k == "cdf_file") v = h->c.numpar.CDF_file;          // relative to the run directory, as opened (flags 'CDF <file>' / 'DOS <file>' included)
    else if (k == "dos_file") v = h->c.numpar.DOS_file;
    else return TRK3_E_INVALID;
    std::snprintf(out, (size_t)outlen, "%s", v.c_str());
    return TRK3_OK;
}

/* One lookup in an ENDL file (tests of the reader): I = 912 -> electrons of the designator, else the value in eV */
int trk3h_eadl_lookup(const char *path, int Z, int I, int designator, double *out) {
    if (!path || !out) return TRK3_E_INVALID;
    Eadl db; std::string e;
    if (!db.load(path, e)) return TRK3_E_INVALID;
    const bool ok = (I == 912) ? db.electrons(Z, designator, *out) : db.real_value(Z, I, designator, *out);
    return ok ? TRK3_OK : TRK3_E_INVALID;
}

int trk3h_num_warnings(trk3h_case *h) { return h ? (int)h->c.warnings.size() : 0; }
int trk3h_warning(trk3h_case *h, int i, char *out, int outlen) {
    if (!h || i < 0 || i >= (int)h->c.warnings.size() || !out || outlen <= 0) return TRK3_E_INVALID;
    std::snprintf(out, (size_t)outlen, "%s", h->c.warnings[i].c_str());
    return TRK3_OK;
}

static bool shell_ok(trk3h_case *h, int atom, int shell) {
    return h && atom >= 0 && atom < (int)h->c.atoms.size() && shell >= 0 && shell < h->c.atoms[atom].nshl();
}

int trk3h_eval_TotIMFP(trk3h_case *h, double E, int atom, int shell, int kind, double *L, double *dEdx) {
    if (!shell_ok(h, atom, shell)) return TRK3_E_INVALID;
    if ((h->c.numpar.kind_of_CDF_ph == 1 && h->c.CDF_Phonon.A[0] == 0.0) || (h->c.numpar.CDF_elast_Zeff >= 2 && !h->c.phonon_renormalised)) get_single_pole(h->c);
    Ctx x = make_ctx(h->c);
    double S, d; TotIMFP(x, E, atom, shell, kind, S, d, nullptr);
    if (L) *L = S;
    if (dEdx) *dEdx = d;
    return TRK3_OK;
}
int trk3h_eval_EMFP(trk3h_case *h, double E, int kind, double *L, double *dEdx) {
    if (!h) return TRK3_E_INVALID;
    if ((h->c.numpar.kind_of_CDF_ph == 1 && h->c.CDF_Phonon.A[0] == 0.0) || (h->c.numpar.CDF_elast_Zeff >= 2 && !h->c.phonon_renormalised)) get_single_pole(h->c);
    Ctx x = make_ctx(h->c);
    double S, d; Elastic_cross_section(x, E, kind, S, d, nullptr);
    if (L) *L = S;
    if (dEdx) *dEdx = d;
    return TRK3_OK;
}
// one value of Diff_cross_section_phonon (Cross_sections.f90:3142-3300) for an electron: the integrand of Tot_EMFP, with the
// screening the case's CDF_elast_Zeff asks for; *phonon_A0 receives the (possibly renormalised) amplitude of the first phonon oscillator
int trk3h_eval_dcs_phonon(trk3h_case *h, double Ee, double hw, double *value, double *phonon_A0) {
    if (!h || !value) return TRK3_E_INVALID;
    if ((h->c.numpar.kind_of_CDF_ph == 1 && h->c.CDF_Phonon.A[0] == 0.0) || (h->c.numpar.CDF_elast_Zeff >= 2 && !h->c.phonon_renormalised)) get_single_pole(h->c);
    Ctx x = make_ctx(h->c);
    double sm = 0, sp = 0; for (auto &a : h->c.atoms) { sm += a.Pers * a.Mass; sp += a.Pers; }
    trk3_dcs_task t{TRK3_DCS_PHONON, x.set_phonon, Ee, 1.0, sm * g_Mp / sp, h->c.Matter.temp, 1.0};
    *value = trk3dcs::eval_request(x.flat->d, t, hw);
    if (phonon_A0) *phonon_A0 = h->c.CDF_Phonon.A[0];
    return TRK3_OK;
}
int trk3h_eval_SHI(trk3h_case *h, double E, int atom, int shell, double *inv_L, double *dEdx, double *Zeff) {
    if (!shell_ok(h, atom, shell)) return TRK3_E_INVALID;
    Ctx x = make_ctx(h->c);
    Ion s = h->c.SHI; s.E = E;
    double S, d; SHI_Total_IMFP(x, s, atom, shell, S, d);
    if (inv_L) *inv_L = S;
    if (dEdx) *dEdx = d;
    if (Zeff) *Zeff = s.Zeff;
    return TRK3_OK;
}
int trk3h_eval_photon(trk3h_case *h, double E, int atom, int shell, double *L) {
    if (!shell_ok(h, atom, shell)) return TRK3_E_INVALID;
    Ctx x = make_ctx(h->c);
    double S, d; Tot_Phot_IMFP(x, E, atom, shell, S, d);
    if (L) *L = S;
    return TRK3_OK;
}
int trk3h_sumrules(trk3h_case *h, int atom, int shell, double *ksum, double *fsum) {
    if (!h) return TRK3_E_INVALID;
    const Case &c = h->c;
    double N_at_mol = 0; for (auto &a : c.atoms) N_at_mol += a.Pers;
    double k, f;
    if (atom < 0) {
        double sm = 0; for (auto &a : c.atoms) sm += a.Pers * a.Mass;
        double Omega = w_plasma(1e6 * c.Matter.At_Dens / N_at_mol, sm * g_Mp / N_at_mol);
        sumrules(c.CDF_Phonon, k, f, 1.0e-8, Omega);
    } else {
        if (!shell_ok(h, atom, shell)) return TRK3_E_INVALID;
        // Sorting_output_data.f90:286-331: molecular density for the printed sum rules
        double Omega = w_plasma(1e6 * c.Matter.At_Dens / N_at_mol);
        sumrules(c.atoms[atom].Ritchi[shell], k, f, c.atoms[atom].Ip[shell], Omega);
    }
    if (ksum) *ksum = k;
    if (fsum) *fsum = f;
    return TRK3_OK;
}
int trk3h_grid(trk3h_case *h, double Emin, double Emax, double *out, int cap) {
    if (!h) return TRK3_E_INVALID;
    auto g = get_grid_4CS(h->c.atoms, Emin, Emax);
    if (out) for (int i = 0; i < cap && i < (int)g.size(); ++i) out[i] = g[i];
    return (int)g.size();
}

int trk3h_save_output(trk3h_case *h, const trk3_tally_layout *lay, const double *tallies, int NMC,
                      const char *out_root, char *out_dir, int out_dir_len, char *err, int errlen) {
    if (!h || !lay || !tallies) return TRK3_E_INVALID;
    std::string od, e;
    if (!save_output(h->c, *lay, tallies, NMC, out_root ? out_root : ".", od, e)) { set_err(err, errlen, e); return TRK3_E_INVALID; }
    if (out_dir && out_dir_len > 0) std::snprintf(out_dir, (size_t)out_dir_len, "%s", od.c_str());
    return TRK3_OK;
}

}  // extern "C"
