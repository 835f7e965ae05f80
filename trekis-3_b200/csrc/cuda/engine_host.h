// engine_host.h -- host-side preparation shared by the CUDA engine (engine.cu) and the CPU emulation
// (tests/emul/emul.cpp): scalar set-up of DevP, the total mean-free-path tables that the reference
// rebuilds in every iteration (How_many_electrons, Monte_Carlo.f90:1902-2057; hoisted out of the
// iteration loop here), the Eckart barrier constants and the set of tallies privatised per block.
#pragma once
#include <cmath>
#include <cstring>
#include <vector>
#include "engine_types.h"

namespace trk3 {

struct HostTotals { std::vector<double> ei_tot, hi_tot, ph_tot, shi_tot; };

inline void compute_totals(const trk3_tables &T, HostTotals &h) {
    const int NS = T.n_shells;
    h.shi_tot.assign(T.n_shi, 0.0);                                        // SHI_path, :1941-1953
    for (int s = 0; s < NS; ++s) for (int i = 0; i < T.n_shi; ++i) h.shi_tot[i] += 1.0 / T.shi_L[(size_t)s * T.n_shi + i];
    for (auto &v : h.shi_tot) v = (v < 1.0e-10) ? 1.0e30 : 1.0 / v;
    h.ei_tot.assign(T.n_ei, 0.0);                                          // El_IMFP, :1991-2011 (only L > 1e-10 summed)
    for (int s = 0; s < NS; ++s) for (int i = 0; i < T.n_ei; ++i) { double L = T.ei_L[(size_t)s * T.n_ei + i]; if (L > 1.0e-10) h.ei_tot[i] += 1.0 / L; }
    for (auto &v : h.ei_tot) v = (v < 1.0e-10) ? 1.0e30 : 1.0 / v;
    h.hi_tot.assign(T.n_hi, 0.0);                                          // Hole_IMFP, :2019-2031
    for (int s = 0; s < NS; ++s) for (int i = 0; i < T.n_hi; ++i) h.hi_tot[i] += 1.0 / T.hi_L[(size_t)s * T.n_hi + i];
    for (auto &v : h.hi_tot) v = (v < 1.0e-10) ? 1.0e30 : 1.0 / v;
    h.ph_tot.assign(T.n_ph, 0.0);                                          // Phot_IMFP, :2041-2055
    for (int s = 0; s < NS; ++s) for (int i = 0; i < T.n_ph; ++i) h.ph_tot[i] += 1.0 / T.ph_L[(size_t)s * T.n_ph + i];
    for (auto &v : h.ph_tot) v = (v < 1.0e-10) ? 1.0e30 : 1.0 / v;
}

// The "cold" range of a total inelastic MFP table: the leading run of identical entries >= 1e16 (energies below the
// lowest ionisation threshold).  For E < E[m] (m = last index of the run, m >= 1) Next_free_path returns exactly that
// constant whatever the search path, so the kernels skip the lookup.  Returns false if there is no such run.
inline bool cold_range(const double *E, const double *L, int N, double &e_cold, double &l_cold) {
    e_cold = -1.0e300; l_cold = 0.0;
    if (N < 2 || !(L[0] >= 1.0e16) || std::fabs(E[1] - E[0]) < 1.0e-6) return false;
    int m = 0;
    while (m + 1 < N && L[m + 1] == L[0]) ++m;
    if (m < 1) return false;
    e_cold = E[m]; l_cold = L[0];
    return true;
}

// Search accelerator of one energy grid (GridLut): bin b covers log(E) in [l0 + b/scale, l0 + (b+1)/scale)
inline void build_lut(const double *E, int N, std::vector<uint16_t> &lut, double &l0, double &scale) {
    lut.assign(TRK_NLUT, 1);
    l0 = 0.0; scale = 0.0;
    if (N < 2 || !(E[0] > 0.0) || N > 65535) return;               // scale = 0: every query starts its scan at index 1
    l0 = std::log(E[0]);
    const double l1 = std::log(E[N - 1]);
    if (!(l1 > l0)) return;
    scale = (double)TRK_NLUT / (l1 - l0);
    int j = 1;
    double lj = std::log(E[1]);                                      // log(E[j]), evaluated once per grid point (per-call cost of a table reload)
    for (int b = 0; b < TRK_NLUT; ++b) {
        const double edge = l0 + (double)b / scale;                  // lower edge of the bin
        while (j < N - 1 && lj <= edge) { ++j; lj = std::log(E[j]); }   // j = number of grid points <= edge (clamped to [1, N-1])
        lut[b] = (uint16_t)j;
    }
}
// Accelerator of the search "first index with 1/L >= x" in a cumulative table of the ion (SHI_energy_transfer).  The bins
// are uniform in x = 1/L itself: that is the measure the collisions sample x with (x = x0 + RN (x1 - x0), :1748), so every
// bin is hit equally often and the local scan that follows the lookup is N / TRK_NLUT steps long on average.
inline void build_inverse_lut(const double *L, int N, int M_temp, std::vector<uint16_t> &lut, double &l0, double &scale) {
    lut.assign(TRK_NLUT, 1);
    l0 = 0.0; scale = 0.0;
    if (N < 2 || N > 65535) return;
    int first = (M_temp >= 2) ? M_temp - 2 : 0;
    while (first < N - 1 && !(L[first] > 0.0 && std::isfinite(1.0 / L[first]))) ++first;
    if (!(L[N - 1] > 0.0)) return;
    l0 = 1.0 / L[first];
    const double l1 = 1.0 / L[N - 1];
    if (!(l1 > l0) || !std::isfinite(l0) || !std::isfinite(l1)) { l0 = 0.0; return; }
    scale = (double)TRK_NLUT / (l1 - l0);
    int j = first + 1;                                                  // 1-based
    double ij = (L[j - 1] > 0.0) ? 1.0 / L[j - 1] : -1.0;              // 1 / L[j-1] (or "never >= edge"), evaluated once per table entry
    for (int b = 0; b < TRK_NLUT; ++b) {
        const double edge = l0 + (double)b / scale;
        while (j < N && !(L[j - 1] > 0.0 && ij >= edge)) { ++j; ij = (L[j - 1] > 0.0) ? 1.0 / L[j - 1] : -1.0; }
        lut[b] = (uint16_t)j;
    }
}
inline double uniform_inv_step(const double *E, int N) {
    if (N < 3) return 0.0;
    const double step = (E[N - 1] - E[0]) / (double)(N - 1);
    if (!(step > 0.0) || E[0] != 0.0) return 0.0;
    for (int i = 0; i < N; ++i) if (std::fabs(E[i] - step * i) > 1e-6 * step) return 0.0;
    return 1.0 / step;
}
#define TRK3_LUT_GRIDS(X, T) X(LUT_EI, T.ei_E, T.n_ei) X(LUT_EE, T.ee_E, T.n_ee) X(LUT_HI, T.hi_E, T.n_hi) X(LUT_HE, T.he_E, T.n_he) \
    X(LUT_PH, T.ph_E, T.n_ph) X(LUT_SHI, T.shi_E, T.n_shi)

// Companion arrays of the tables (natural logarithms, reciprocals): X(dst, src, n, op) with op 0 = log(src),
// 1 = 1/src, 2 = log(1/src).  The CUDA engine evaluates them on the device, the emulation on the host.
#define TRK3_COMPANIONS(X, p, T, NS)                                                                              \
    X(lei_E, p.ei_E, T.n_ei, 0) X(lei_L, p.ei_L, (NS) * T.n_ei, 0) X(lei_tot, p.ei_tot, T.n_ei, 0)                  \
    X(lee_E, p.ee_E, T.n_ee, 0) X(lee_L, p.ee_L, T.n_ee, 0)                                                        \
    X(lhi_E, p.hi_E, T.n_hi, 0) X(lhi_L, p.hi_L, (NS) * T.n_hi, 0) X(lhi_tot, p.hi_tot, T.n_hi, 0)                  \
    X(lhe_E, p.he_E, T.n_he, 0) X(lhe_L, p.he_L, T.n_he, 0)                                                        \
    X(lph_E, p.ph_E, T.n_ph, 0) X(lph_L, p.ph_L, (NS) * T.n_ph, 0) X(lph_tot, p.ph_tot, T.n_ph, 0)                  \
    X(lshi_E, p.shi_E, T.n_shi, 0) X(lshi_L, p.shi_L, (NS) * T.n_shi, 0) X(lshi_tot, p.shi_tot, T.n_shi, 0)         \
    X(ldshi_E, p.dshi_E, T.dshi_off[NS], 0) X(dshi_iL, p.dshi_L, T.dshi_off[NS], 1) X(ldshi_iL, p.dshi_L, T.dshi_off[NS], 2) \
    X(leid_hw, p.eid_hw, T.eid_off[(NS) * T.n_ei], 0) X(leid_L, p.eid_L, T.eid_off[(NS) * T.n_ei], 0)               \
    X(leed_hw, p.eed_hw, T.eed_off[T.n_ee], 0) X(leed_L, p.eed_L, T.eed_off[T.n_ee], 0)                            \
    X(lhid_hw, p.hid_hw, T.hid_off[T.n_hi], 0) X(lhid_L, p.hid_L, T.hid_off[T.n_hi], 0)                            \
    X(lhed_hw, p.hed_hw, T.hed_off[T.n_he], 0) X(lhed_L, p.hed_L, T.hed_off[T.n_he], 0)

TRK_HD double companion_value(double x, int op) { return op == 0 ? log(x) : (op == 1 ? 1.0 / x : log(1.0 / x)); }

// Equilibrium_charge_SHI for the incoming ion (MAIN.f90:171)
inline double host_shi_zeff(const trk3_config &c, const trk3_tables &T) {
    const double g_e = 1.602176487e-19, g_me = 9.1093821545e-31, g_Mp = 1836.1526724780 * g_me, g_cvel = 299792458.0, g_Ry = 13.6056981;
    double vp = (c.shi_E > 0.0) ? std::sqrt(2.0 * c.shi_E * g_e / (c.shi_mass * g_Mp)) : 0.0;
    double sz = 0, sp = 0; for (int a = 0; a < T.n_atoms; ++a) { sz += T.atom_Z[a] * T.atom_pers[a]; sp += T.atom_pers[a]; }
    double Zt = sz / sp, Zp = (double)c.shi_Z, g_v0 = std::sqrt(2.0 * g_Ry * g_e / g_me);
    switch (c.shi_kind_Zeff) {
    case 1: return Zp * (1.0 - std::exp(-(vp / g_v0 / std::pow(Zp, 0.66666666))));
    case 2: { double c1 = 0.6, c2 = 0.45; return Zp * std::pow(1.0 + std::pow(vp / (std::pow(Zp, c2) * g_v0 * 4.0 / 3.0), -1.0 / c1), -c1); }
    case 3: {
        double c1 = 1.0 - 0.26 * std::exp(-Zt / 11.0 - (Zt - Zp) * (Zt - Zp) / 9.0);
        double vpvo = std::pow(Zp, -0.543) * vp / g_v0;
        double c2 = 1.0 + 0.03 * vpvo * std::log(Zt);
        double x = c1 * std::pow(vpvo / c2 / 1.54, 1.0 + 1.83 / Zp), x2 = x * x, x4 = x2 * x2;
        return Zp * (8.29 * x + x4) / (0.06 / x + 4.0 + 7.4 * x + x4); }
    case 4: return c.shi_fixed_Zeff;
    default: return Zp * (1.0 - std::exp(-(vp * 125.0 / g_cvel / std::pow(Zp, 0.66666666))));
    }
}

// Everything of DevP that does not involve device pointers.
inline int fill_devp_scalars(const trk3_config &c, const trk3_tables &T, const trk3_tally_layout &lay, DevP &p) {
    std::memset(&p, 0, sizeof p);
    if (T.n_atoms < 1 || T.n_atoms > TRK3_MAX_ATOMS || T.n_shells < 1 || T.n_shells > TRK3_MAX_SHELLS) return TRK3_E_INVALID;
    if (c.kind_of_EMFP == 2 && (T.n_dsf_e < 2 || T.n_dsf_h < 2 || T.n_ee < 2 || T.n_he < 2)) return TRK3_E_UNSUPPORTED;   // DSF scattering without its tables
    if (c.include_photons && T.n_ph <= 0) return TRK3_E_INVALID;
    const double g_me = 9.1093821545e-31, g_Mp = 1836.1526724780 * g_me, g_e = 1.602176487e-19, g_h = 1.05457162853e-34, g_Pi = 3.1415926535897932384626433832795;
    p.ion_E = c.shi_E; p.ion_mass = c.shi_mass; p.ion_fixed_Zeff = c.shi_fixed_Zeff; p.ion_Z = c.shi_Z; p.ion_kind_Zeff = c.shi_kind_Zeff;
    p.ion_Zeff0 = host_shi_zeff(c, T);
    p.ion_pow23 = std::pow((double)c.shi_Z, 0.66666666);
    p.Tim = c.Tim; p.cut_off = c.cut_off; p.layer = c.layer; p.hole_mass = c.hole_mass;
    p.work_function = c.work_function; p.bar_height = c.bar_height;
    p.include_photons = c.include_photons; p.kind_of_EMFP = c.kind_of_EMFP;
    p.seed_lo = (uint32_t)c.seed; p.seed_hi = (uint32_t)(c.seed >> 32);
    p.n_atoms = T.n_atoms; p.n_shells = T.n_shells; p.vb_shell = T.vb_shell; p.nshl_atom1 = T.nshl_atom1;
    double sm = 0, sp = 0;
    for (int a = 0; a < T.n_atoms; ++a) {
        p.atom_Z[a] = T.atom_Z[a]; p.atom_first[a] = T.atom_first[a]; p.atom_nshl[a] = T.atom_nshl[a];
        p.atom_mass[a] = T.atom_mass[a]; p.atom_pers[a] = T.atom_pers[a];
        sm += T.atom_mass[a] * T.atom_pers[a]; sp += T.atom_pers[a];
    }
    p.Mtarget = g_Mp * sm / sp; p.sum_pers = sp;
    p.Erest_target = p.Mtarget * 299792458.0 * 299792458.0 / 1.602176487e-19;      // rest_energy (physics.cuh), the same operations
    for (int s = 0; s < T.n_shells; ++s) {
        p.shell_atom[s] = T.shell_atom[s]; p.shell_num[s] = T.shell_num[s]; p.shell_Ip[s] = T.shell_Ip[s];
        p.shell_Nel[s] = T.shell_Nel[s]; p.shell_auger[s] = T.shell_auger[s]; p.shell_radiat[s] = T.shell_radiat[s];
        p.shell_kocs[s] = (T.shell_kocs[s] == 2) ? 2 : 1; p.shell_Ek[s] = T.shell_Ek[s];   // 0 from a caller that predates the fields = CDF
    }
    p.at_dens = T.at_dens;
    p.n_dsf_e = (c.kind_of_EMFP == 2) ? T.n_dsf_e : 0; p.n_dsf_h = (c.kind_of_EMFP == 2) ? T.n_dsf_h : 0;
    p.delta_cdf = T.delta_cdf ? 1 : 0;
    for (int s = 0; s <= TRK3_MAX_SHELLS; ++s) p.osc_off[s] = T.delta_cdf ? T.osc_off[s] : 0;
    p.Egap = T.shell_Ip[T.atom_first[0] + T.atom_nshl[0] - 1];            // Target_atoms(1)%Ip(size(...))
    p.n_ei = T.n_ei; p.n_ee = T.n_ee; p.n_hi = T.n_hi; p.n_he = T.n_he; p.n_ph = T.n_ph; p.n_shi = T.n_shi; p.n_dos = T.n_dos; p.n_r = T.n_r;
    p.Nt = lay.Nt;
    for (int i = 0; i < lay.Nt; ++i) p.tg[i] = (lay.time_grid[i] < c.Tim) ? lay.time_grid[i] : c.Tim;   // tim_glob = min(time_grid(i),Tim), :580
    if (c.work_function > 0) {                                             // barrier_parameters, :2094-2115
        double wf = c.work_function, bh = c.bar_height;
        double Em_L = c.bar_length * 1.0e-10;
        double Em_B = 2.0 * bh - wf + 2.0 * std::sqrt(bh * bh - bh * wf);
        double Em_ksi = 0.5 * std::sqrt(8.0 * g_me * Em_L * Em_L * Em_B * g_e / ((2.0 * g_Pi * g_h) * (2.0 * g_Pi * g_h)) - 1.0);
        double Em_bb = std::cosh(2.0 * g_Pi * Em_ksi);
        double Em_delta = 2.0 * g_Pi * Em_L * std::sqrt(2.0 * g_me * g_e) / (2.0 * g_Pi * g_h);
        p.Em_E1 = bh + 2.0 * std::sqrt(bh * (bh - wf)) * (std::acosh(Em_bb) / (Em_delta * (std::sqrt(bh) + std::sqrt(bh - wf))) - 1.0);
        double g1 = Em_delta * (std::sqrt(p.Em_E1) + std::sqrt(p.Em_E1 - wf)), g2 = Em_delta * (std::sqrt(p.Em_E1) - std::sqrt(p.Em_E1 - wf));
        p.Em_gamma = (g1 * std::sinh(g1) + 2.0 * g2 * std::sinh(g2)) / (std::sqrt(p.Em_E1 * (p.Em_E1 - wf)) * (Em_bb + std::cosh(g1)));
    }
    // tallies written from inside the event loop are candidates for block-private (shared-memory) copies
    const int priv[] = {TRK3_OUT_NE, TRK3_OUT_EE, TRK3_OUT_ELAT, TRK3_OUT_NH, TRK3_OUT_EH, TRK3_OUT_EHKIN, TRK3_OUT_NPHOT,
                        TRK3_OUT_EPHOT, TRK3_OUT_E_E, TRK3_OUT_E_H, TRK3_OUT_E_PHOT};
    for (int i = 0; i < TRK3_N_TALLIES; ++i) { p.g_off[i] = lay.off[i]; p.s_off[i] = -1; p.s_len[i] = 0; }
    int o = 0;
    for (int id : priv) {
        if ((id == TRK3_OUT_NPHOT || id == TRK3_OUT_EPHOT || id == TRK3_OUT_E_PHOT) && !c.include_photons) continue;
        p.s_off[id] = o; p.s_len[id] = (int)lay.len[id]; o += (int)lay.len[id];
    }
    p.s_total = o;
    return TRK3_OK;
}

// sizes of the per-iteration scratch for a batch of nb iterations
struct ScratchLayout {
    size_t u32_total, f64_total;
    size_t created, nvb, nph, spec_e, spec_h, th_e, th_h, diffN, em_cnt, em_spec;   // offsets in the u32 slab
    size_t diffS, esnap, elat, em_E;                                                  // offsets in the f64 slab
};
inline ScratchLayout scratch_layout(const DevP &p, size_t nb) {
    ScratchLayout s{};
    const size_t Nt = p.Nt;
    size_t o = 0;
    s.created = o; o += nb * (Nt + 2);
    s.nvb = o; o += nb * Nt;
    s.nph = o; o += nb * Nt;
    s.spec_e = o; o += nb * Nt * p.n_r;
    s.spec_h = o; o += nb * Nt * p.n_dos;
    s.th_e = o; o += nb * Nt * TRK3_NTHETA;
    s.th_h = o; o += nb * Nt * TRK3_NTHETA;
    s.diffN = o; o += nb * Nt;
    s.em_cnt = o; o += nb * (Nt + 2);
    s.em_spec = o; o += (p.work_function > 0) ? nb * (Nt + 2) * p.n_r : 0;
    s.u32_total = o;
    o = 0;
    s.diffS = o; o += nb * Nt;
    s.esnap = o; o += nb * Nt;
    s.elat = o; o += nb * (Nt + 2);
    s.em_E = o; o += nb * (Nt + 2);
    s.f64_total = o;
    return s;
}
inline void bind_scratch(DevP &p, const ScratchLayout &s, uint32_t *u, double *d) {
    p.it.created = u + s.created; p.it.nvb = u + s.nvb; p.it.nph = u + s.nph; p.it.spec_e = u + s.spec_e; p.it.spec_h = u + s.spec_h;
    p.it.th_e = u + s.th_e; p.it.th_h = u + s.th_h; p.it.diffN = u + s.diffN; p.it.em_cnt = u + s.em_cnt; p.it.em_spec = u + s.em_spec;
    p.it.diffS = d + s.diffS; p.it.esnap = d + s.esnap; p.it.elat = d + s.elat; p.it.em_E = d + s.em_E;
}

// estimate of particles per iteration (the reference's own array-size rule, Monte_Carlo.f90:1955-1965)
inline double estimate_nel(const trk3_config &c, const trk3_tables &T) {
    // total stopping power at the ion energy (linear interpolation between the bracketing grid points is enough here)
    int n = 1; while (n < T.n_shi - 1 && T.shi_E[n] < c.shi_E) ++n;
    double se = 0.0;
    for (int s = 0; s < T.n_shells; ++s) { double a = T.shi_dEdx[(size_t)s * T.n_shi + n - 1], b = T.shi_dEdx[(size_t)s * T.n_shi + n]; se += (a > b ? a : b); }
    double nel = std::ceil(se * c.layer / T.shell_Ip[T.vb_shell]);
    if (nel > 5000.0 * c.layer) nel = 5000.0 * c.layer;
    if (nel < 1000.0) nel = 1000.0;
    return nel;
}

}  // namespace trk3
