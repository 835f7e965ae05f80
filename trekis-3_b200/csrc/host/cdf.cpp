// cdf.cpp -- CDF-derived cross sections built on the host (fp64), restating
// Source_files/Cross_sections.f90 of the reference: Ritchie-Howie loss function with finite-q
// extension (:346-486), q- and hw-integrals (:881-1050, :2217-2282, :2452-2597, :2683-2744,
// :2966-3300), photon IMFP (:833-878), single-pole fallback and sum rules (:554-826),
// effective charges (:2601-2680), searches and interpolation.
// Loop orders and expression shapes follow the Fortran so that a reference build and this
// builder agree to rounding (target 1e-12 relative, BASELINE.json north_star).
#include "../common/trk3_delta.h"
#include "trk3_host.hpp"
#include "../common/trk3_dcs.h"
#include <algorithm>
#include <complex>
#include <cstdio>

namespace trk3 {

// ---------------------------------------------------------------------------------------------
// searches and interpolation
// ---------------------------------------------------------------------------------------------
int find_monoton_1d(const double *A, int N, double v) {   // Reading_files_and_parameters.f90:3433-3494
    int i_1 = 1, i_2 = N;
    int i_cur = (int)std::floor((i_1 + i_2) / 2.0);
    double temp_val = A[i_cur - 1];
    if (v < A[0]) i_cur = 0;
    else if (v >= A[N - 1]) i_cur = N - 1;
    else {
        for (;;) {
            if (i_1 == i_2 - 1) break;
            if (temp_val <= v) i_1 = i_cur; else i_2 = i_cur;
            i_cur = (int)std::floor((i_1 + i_2) / 2.0);
            temp_val = A[i_cur - 1];
        }
    }
    return i_cur + 1;
}

int find_monoton_2d(const double *A, int stride, int N, double v) {   // :3496-3559
    auto a = [&](int i) { return A[(size_t)(i - 1) * stride]; };
    int i_1 = 1, i_2 = N;
    int i_cur = (int)std::floor((i_1 + i_2) / 2.0);
    double temp_val = a(i_cur);
    if (v < a(1)) i_cur = 0;
    else if (v >= a(N)) i_cur = N - 1;
    else {
        int coun = 0;
        for (;;) {
            if (v >= a(i_cur) && v <= a(i_cur + 1)) break;
            if (temp_val <= v) i_1 = i_cur; else i_2 = i_cur;
            i_cur = (int)std::floor((i_1 + i_2) / 2.0);
            temp_val = a(i_cur);
            if (++coun > 1000) break;
        }
    }
    return i_cur + 1;
}

int find_monoton_decreasing(const double *A, int N, double v) {   // :3376-3429
    int i_1 = 1, i_2 = N;
    int i_cur = (int)std::floor((i_1 + i_2) / 2.0);
    double temp_val = A[i_cur - 1];
    if (v < A[N - 1]) i_cur = N;
    else if (v > A[0]) i_cur = 1;
    else {
        int coun = 0;
        while (std::abs(i_1 - i_2) > 1) {
            if (temp_val > v) i_1 = i_cur; else i_2 = i_cur;
            i_cur = (int)std::floor((i_1 + i_2) / 2.0);
            temp_val = A[i_cur - 1];
            if (++coun > 1000) break;
        }
    }
    return i_cur;
}

double interpolate(int flag, double E1, double E2, double S1, double S2, double E) {   // Cross_sections.f90:4051-4086
    if (std::fabs(E2 - E1) < 1.0e-6) return std::max(S1, S2);
    if (E == E1) return S1;
    switch (flag) {
    case 3: { double a = std::log(E2), b = std::log(E1), x = std::log(E); return S1 + (S2 - S1) / (a - b) * (x - b); }
    case 4: { double a = std::log(S1), b = std::log(S2); return std::exp(a + (b - a) / (E2 - E1) * (E - E1)); }
    case 5: {
        double E2l = std::log(E2), E1l = std::log(E1), El = std::log(E), S1l = std::log(S1), S2l = std::log(S2);
        return std::exp(S1l + (S2l - S1l) / (E2l - E1l) * (El - E1l));
    }
    default: return S1 + (S2 - S1) / (E2 - E1) * (E - E1);
    }
}

// ---------------------------------------------------------------------------------------------
// loss function
// ---------------------------------------------------------------------------------------------
Ctx make_ctx(const Case &c) {
    Ctx x; x.c = &c;
    if (c.Matter.El_eff_mass == 0.0) {     // Imewq, Cross_sections.f90:428-434
        x.mass_from_dos = true;
        if (c.atoms[0].Ip.back() < 0.2) { x.k = &c.dos.k_inv; x.effm = &c.dos.Eff_m_inv; }
        else { x.k = &c.dos.k; x.effm = &c.dos.Eff_m; }
    }
    // flattened view for the shared integrands (csrc/common/trk3_dcs.h): oscillator sets in (atom, shell) order, phonon CDF last
    auto flat = std::make_shared<CtxFlat>();
    flat->off.push_back(0);
    auto add = [&](const CDFosc &o) {
        flat->E0.insert(flat->E0.end(), o.E0.begin(), o.E0.end());
        flat->A.insert(flat->A.end(), o.A.begin(), o.A.end());
        flat->G.insert(flat->G.end(), o.Gamma.begin(), o.Gamma.end());
        flat->off.push_back((int32_t)flat->E0.size());
    };
    x.set0.clear();
    int nset = 0;
    for (auto &a : c.atoms) { x.set0.push_back(nset); for (int sh = 0; sh < a.nshl(); ++sh) { add(sh < (int)a.Ritchi.size() ? a.Ritchi[sh] : CDFosc{}); ++nset; } }
    add(c.CDF_Phonon); x.set_phonon = nset++;
    trk3_dcs_ctx &d = flat->d;
    d.osc_E0 = flat->E0.data(); d.osc_A = flat->A.data(); d.osc_G = flat->G.data(); d.osc_off = flat->off.data(); d.n_sets = nset;
    d.k = x.k ? x.k->data() : nullptr; d.effm = x.effm ? x.effm->data() : nullptr; d.n_k = x.k ? (int32_t)x.k->size() : 0;
    d.mass_from_dos = x.mass_from_dos ? 1 : 0; d.El_eff_mass = c.Matter.El_eff_mass;
    d.kind_DR = c.numpar.kind_of_DR; d.v_f = c.Matter.v_f; d.temp = c.Matter.temp;
    // dynamical screening of the elastic cross section (CDF_elast_Zeff = 2 / 3): what get_screening_ff / get_screening_all read
    d.screening = (c.numpar.kind_of_EMFP == 1 && c.numpar.CDF_elast_Zeff >= 2) ? c.numpar.CDF_elast_Zeff : 0;
    d.vb_set = x.set0[0] + c.atoms[0].nshl() - 1;
    flat->scr.assign(1, (double)c.atoms.size());
    for (size_t i = 0; i < c.atoms.size(); ++i) {
        const Atom &a = c.atoms[i];
        double core = 0.0;
        for (int sh = 0; sh < a.nshl() - (i == 0 ? 1 : 0); ++sh) core += a.Nel[(size_t)sh];
        flat->scr.insert(flat->scr.end(), {(double)a.Zat, a.Pers, core, (double)x.set0[i], (double)a.nshl()});
        std::array<double, 5> ff{};
        if (a.Zat >= 1 && a.Zat <= (int)c.form_factor.size()) ff = c.form_factor[(size_t)a.Zat - 1];
        flat->scr.insert(flat->scr.end(), ff.begin(), ff.end());
    }
    for (auto &a : c.atoms) for (int sh = 0; sh < a.nshl(); ++sh) { flat->scr.push_back(a.Nel[(size_t)sh]); flat->scr.push_back(a.Ip[(size_t)sh]); }
    d.scr = flat->scr.data(); d.n_scr = (int32_t)flat->scr.size();
    x.flat = flat;
    return x;
}

// One q-integral.  Direct mode: evaluated here.  Record mode: the request is noted and 0 returned (the outer loops never
// branch on the value).  Replay mode: the value the evaluator produced for this request.
double dcs_request(const Ctx &x, const trk3_dcs_task &t, double hw) {
    DcsBatch *b = x.batch;
    if (!b || b->mode == DcsBatch::DIRECT) return trk3dcs::eval_request(x.flat->d, t, hw);
    if (b->mode == DcsBatch::RECORD) {
        if (b->hw.empty()) b->task = t;
        b->hw.push_back(hw);
        return 0.0;
    }
    return b->val[b->cursor++];
}

namespace {
inline double imewq_photon(const CDFosc &o, double hw) {   // Loss_func with photon=.true.: q = 0
    double dE2 = hw * hw, ImE = 0.0;
    for (size_t i = 0; i < o.A.size(); ++i) {
        double E0 = o.E0[i], G = o.Gamma[i], E02 = E0 * E0;
        ImE = ImE + o.A[i] * G * hw / ((dE2 - E02) * (dE2 - E02) + G * G * dE2);
    }
    return ImE;
}
// the three q-integrals (csrc/common/trk3_dcs.h) as requests of the current outer integration
inline double diff_cross_section(const Ctx &x, int set, double Ee, double dE, double Mass) {
    trk3_dcs_task t{TRK3_DCS_INELASTIC, set, Ee, Mass, 0.0, 0.0, 0.0};
    return dcs_request(x, t, dE);
}
inline double shi_diff_cross_section(const Ctx &x, int set, double Ee, double MSHI, double Emax, double hw) {
    trk3_dcs_task t{TRK3_DCS_SHI, set, Ee, 1.0, MSHI, Emax, 0.0};
    return dcs_request(x, t, hw);
}
inline double shi_diff_cross_section_bk(const Ctx &x, int set, double Ee, double MSHI, double Emax, double hw, double Z_SHI, double Zeff) {
    trk3_dcs_task t{TRK3_DCS_SHI_BK, set, Ee, Z_SHI, MSHI, Emax, Zeff};
    return dcs_request(x, t, hw);
}
inline double diff_cross_section_phonon(const Ctx &x, double Ee, double dE, double Mtarget, double Mass, double Ttarget, double pref) {
    trk3_dcs_task t{TRK3_DCS_PHONON, x.set_phonon, Ee, Mass, Mtarget, Ttarget, pref};
    return dcs_request(x, t, dE);
}

// get_diff_CS_grid_size (new grid, both bounds present), Cross_sections.f90:1278-1343
int diff_CS_grid_size(int n, double Emax, double E0_min, double E0_max, double dE_min_use) {
    double E = E0_min;      // define_integration_limits: E_start = E0_min
    int i = 0;
    while (E <= Emax) { ++i; E = E + define_dE(1, n, E, true, E0_min, true, E0_max, dE_min_use); }
    return std::max(i, 10);
}

// get_E_low, Cross_sections.f90:1251-1274
void get_E_low(double lo_b, double hi_b, double Emin, double Emax, double &E_low, double &E_high) {
    if (Emin > hi_b || Emax < lo_b) { E_low = Emin; E_high = Emax; return; }
    E_low = std::max(lo_b, Emin); E_high = std::min(hi_b, Emax);
    if (std::fabs(E_low - E_high) < 1.0e-3) { E_low = Emin; E_high = Emax; return; }
    double dE = hi_b - lo_b;
    E_low = std::min(E_low, E_high - dE);
    E_low = std::max(E_low, Emin);
}

double hole_mass_at(const Case &c, double Ele) {      // e.g. Cross_sections.f90:915-921
    if (c.Matter.hole_mass >= 0) return c.Matter.hole_mass;
    int m = find_monoton_1d(c.dos.E.data(), (int)c.dos.E.size(), Ele);
    return c.dos.Eff_m[m - 1];
}
double mean_target_mass(const Case &c) {              // g_Mp*SUM(Mass*Pers)/SUM(Pers)
    double sm = 0, sp = 0;
    for (auto &a : c.atoms) { sm += a.Mass * a.Pers; sp += a.Pers; }
    return g_Mp * sm / sp;
}
}  // namespace

// define_dE, Cross_sections.f90:1374-1412
double define_dE(int CS_method, int n, double E, bool has_min, double E0_min, bool has_max, double E0_max, double dE_min_use) {
    double dE;
    if (CS_method == -1) dE = (1.0 / (E + 1.0) + E) / (double)n;
    else if (has_min && has_max) {
        if (E > E0_min && E < E0_max) dE = (E0_max - E0_min) / (double)n;
        else dE = E / (double)n;
    } else dE = (1.0 / (E + 1.0) + E) / (double)n;
    return std::max(dE, dE_min_use);
}

// ---------------------------------------------------------------------------------------------
// TotIMFP, Cross_sections.f90:881-1050 (CS_method = 1, Ritchie CDF shells)
// ---------------------------------------------------------------------------------------------
// BEB cross sections, Cross_sections.f90:3891-3906 (Sigma_BEB), :4000-4040 (dSigma_w_int_BEB, dSigma_dw_w_int): closed forms of
// Kim & Rudd, Phys. Rev. A 50 (1994) 3954.  T = energy of the incident particle, B = binding energy, U = mean kinetic energy of
// the shell, N = its electrons; [A^2] and [A^2 eV].
double Sigma_BEB(double T, double B, double U, double N) {
    if (T <= B) return 0.0;
    const double S = 4.0 * g_Pi * g_a0 * g_a0 * N * (g_Ry / B) * (g_Ry / B);
    const double t0 = T / B, u0 = U / B;
    return S / (t0 + u0 + 1.0) * (std::log(t0) * 0.5 * (1.0 - 1.0 / (t0 * t0)) + (1.0 - 1.0 / t0) - std::log(t0) / (t0 + 1.0));
}
static double dSigma_dw_w_int(double S, double t0, double u0, double w0) {
    const double tw = 1.0 / (t0 + u0 + 1.0);
    const double logwt = std::log((w0 + 1.0) * std::fabs(w0 - t0));
    const double overwt = 1.0 / (w0 - t0), overw1 = 1.0 / (w0 + 1.0), overt1 = 1.0 / (t0 + 1.0);
    const double A = tw * overt1 * overt1 * (t0 * t0 * std::log(std::fabs(w0 - t0)) + t0 * logwt + std::log(w0 + 1.0));
    const double B = tw * (-t0 * overwt + overw1 + logwt);
    const double C = tw * std::log(t0) * (0.5 * (overw1 * overw1 + t0 * overwt * overwt) + overwt - overw1);
    return S * (A + B + C);
}
double dSigma_w_int_BEB(double T, double w, double B, double U, double N) {
    const double S = 4.0 * g_Pi * g_a0 * g_a0 * N * (g_Ry / B) * (g_Ry / B);
    const double t0 = T / B, u0 = U / B, w0 = w / B;
    return B * (dSigma_dw_w_int(S, t0, u0, w0) - dSigma_dw_w_int(S, t0, u0, 0.0));
}

void TotIMFP(const Ctx &x, double Ele, int Nat, int Nshl, int kind, double &Sigma, double &dEdx, DiffRow *row) {
    const Case &c = *x.c;
    const Atom &at = c.atoms[Nat];
    if (Nshl < (int)at.KOCS.size() && at.KOCS[Nshl] == 2) {          // BEB shell (:1041-1047): closed form, no differential table
        const double Mass = (kind == 0) ? 1.0 : hole_mass_at(c, Ele);
        double sp = 0; for (auto &a : c.atoms) sp += a.Pers;
        const double temp1 = c.Matter.At_Dens * 1e-24 * at.Pers / sp;
        const double sig = Sigma_BEB(Ele, at.Ip[Nshl], at.Ek[Nshl], at.Nel[Nshl]);
        Sigma = 1.0 / (Mass * temp1 * sig);
        dEdx = Mass * temp1 * dSigma_w_int_BEB(Ele, (Ele - 1.0) / 2.0, at.Ip[Nshl], at.Ek[Nshl], at.Nel[Nshl]);
        if (row) { row->hw.clear(); row->L.clear(); }
        return;
    }
    const CDFosc &o = at.Ritchi[Nshl];
    double Emin = at.Ip[Nshl];
    double Egap = c.atoms[0].Ip.back();
    if (Emin <= 1.0e-3) Emin = 1.0e-3;
    if (c.numpar.kind_of_DR == 4) {          // Delta-CDF (:952-956): closed form for electrons AND holes (M = mt = g_me), no differential table
        const double sig = trk3delta::Integral_CDF_delta_CS(g_me, g_me, Ele, o.E0.data(), o.alpha.data(), (int)o.E0.size(), Emin, c.Matter.At_Dens, true, -1.0);
        Sigma = trk3delta::MFP_from_sigma(sig, c.Matter.At_Dens);
        dEdx = trk3delta::energy_loss_delta(Ele, g_me, 1.0, Emin, c.Matter.At_Dens, g_me, o.E0.data(), o.alpha.data(), (int)o.E0.size(), true);
        if (row) { row->hw.clear(); row->L.clear(); }
        return;
    }
    double Emax, Mass;
    bool save = false;
    if (kind == 0) { Emax = (Ele + Emin) / 2.0; Mass = 1.0; save = true; }
    else {
        save = (Nat == 0 && Nshl == at.nshl() - 1);
        Mass = hole_mass_at(c, Ele);
        Emax = 4.0 * Ele * Mass / ((Mass + 1.0) * (Mass + 1.0));
    }
    if (c.numpar.plasmon_Emax && Emin == Egap) {
        double Epl = std::sqrt(c.Matter.N_VB_el * c.Matter.At_Dens * 1e6 * g_h * g_h / (g_me * g_e0) + Egap * Egap);
        if (Epl >= Emax) Emax = Epl;
        if (Ele < Emax) Emax = Ele;
    }
    const int n = 100;                               // m_N_grid_e_inelast
    double lo = 1e300, hi = -1e300;
    for (size_t i = 0; i < o.E0.size(); ++i) { lo = std::min(lo, o.E0[i] - 5.0 * o.Gamma[i]); hi = std::max(hi, o.E0[i] + 5.0 * o.Gamma[i]); }
    double E_low = std::max(lo, Emin), E_high = std::min(hi, Emax);

    std::vector<double> hw, cum;
    double E = Emin, Ltot1 = 0.0, ddEdx = 0.0;
    const int set = x.set0[Nat] + Nshl;
    double Ltot0 = diff_cross_section(x, set, Ele, E, Mass);
    while (E <= Emax) {
        double dE = define_dE(1, n, E, true, E_low, true, E_high, 0.001);
        double a = E + dE / 2.0;
        double temp1 = diff_cross_section(x, set, Ele, a, Mass);
        double b = E + dE;
        double dL = diff_cross_section(x, set, Ele, b, Mass);
        double temp2 = dE / 6.0 * (Ltot0 + 4.0 * temp1 + dL);
        Ltot1 = Ltot1 + temp2;
        ddEdx = ddEdx + E * temp2;
        Ltot0 = dL;
        E = E + dE;
        if (row && save) { hw.push_back(E); cum.push_back(Ltot1); }
    }
    if (Ltot1 < 1.0e-15) { Sigma = 1.0e20; dEdx = 0.0; }
    else { Sigma = 1.0 / (Mass * Ltot1); dEdx = Mass * ddEdx; }
    if (row && save) {
        // allocate_diff_CS_tables (:1055-1142): the row is sized by a count loop that starts at E_low, while the
        // fill loop above starts at Emin.  Where E_low > Emin the reference writes past the allocation (undefined
        // behaviour); the readable part is the first Nsiz entries, which is what is kept here.  Short rows are
        // padded with the allocation's initial (hw=0, cum=0) -> (0, 1e15).
        int Nsiz = diff_CS_grid_size(n, Emax, E_low, E_high, 0.001);
        row->hw.assign(Nsiz, 0.0); row->L.assign(Nsiz, 0.0);
        for (int i = 0; i < Nsiz && i < (int)hw.size(); ++i) { row->hw[i] = hw[i]; row->L[i] = cum[i]; }
        for (int i = 0; i < Nsiz; ++i) row->L[i] = (row->L[i] < 1.0e-15) ? 1.0e15 : 1.0 / (Mass * row->L[i]);
    }
}

// ---------------------------------------------------------------------------------------------
// Tot_EMFP / Elastic_cross_section, Cross_sections.f90:2892-3139
// ---------------------------------------------------------------------------------------------
void Tot_EMFP(const Ctx &x, double Ele, int kind, double Zeff, double &Sigma, double &dEdx, DiffRow *row) {
    const Case &c = *x.c;
    const double pref = 1.0;
    double qdebye = std::pow(6.0 * g_Pi * g_Pi * (c.Matter.At_Dens * 1e6), 0.33333333);   // Debye_energy :692
    double Edebay = g_h * c.Matter.Vsound * qdebye / g_e;
    Edebay = Edebay * 3.0;
    double Emin = pref * 0.1e-8;
    double Mtarget = mean_target_mass(c);
    double Mass = (kind == 0) ? 1.0 : hole_mass_at(c, Ele);
    double Emax = 4.0 * Ele * Mass * g_me * Mtarget / ((Mtarget + Mass * g_me) * (Mtarget + Mass * g_me));
    if (Edebay >= Emax) Emax = Edebay;
    if (Ele < Emax) Emax = Ele;
    Emax = pref * Emax;
    const int n = 20;                                // m_N_grid_e_elast
    const CDFosc &p = c.CDF_Phonon;
    double lo = 1e300, hi = -1e300;
    for (size_t i = 0; i < p.E0.size(); ++i) { lo = std::min(lo, p.E0[i] - 5.0 * p.Gamma[i]); hi = std::max(hi, p.E0[i] + 5.0 * p.Gamma[i]); }
    double E_low, E_high;
    get_E_low(lo, hi, Emin, Emax, E_low, E_high);

    std::vector<double> hw, cum;
    double E = E_low, Ltot1 = 0.0, ddEdx = 0.0;
    double Ltot0 = diff_cross_section_phonon(x, Ele, E, Mtarget, Mass, c.Matter.temp, 1.0);
    while (std::fabs(E) <= std::fabs(Emax)) {
        double dE = define_dE(1, n, E, true, E_low, true, E_high, 1.0e-5);
        dE = pref * dE;
        double a = E + dE / 2.0;
        double temp1 = diff_cross_section_phonon(x, Ele, a, Mtarget, Mass, c.Matter.temp, 1.0);
        double b = E + dE;
        double dL = diff_cross_section_phonon(x, Ele, b, Mtarget, Mass, c.Matter.temp, 1.0);
        double temp2 = std::fabs(dE) / 6.0 * (Ltot0 + 4.0 * temp1 + dL);
        Ltot1 = Ltot1 + temp2;
        ddEdx = ddEdx + E * temp2;
        Ltot0 = dL;
        E = E + dE;
        if (row) { hw.push_back(E); cum.push_back(Ltot1); }
    }
    if (Mass < 1.0e-20 || Ltot1 < 1.0e-15) Sigma = 1.0e29;
    else Sigma = 1.0 / (Zeff * Zeff * Mass * Ltot1);
    if (Sigma > 1e30) Sigma = 1e30;
    dEdx = (Zeff * Zeff) * Mass * ddEdx;
    if (row) {
        int Nsiz = diff_CS_grid_size(n, Emax, E_low, E_high, 1.0e-5);   // allocate_diff_CS_elastic_tables :1148
        row->hw.assign(Nsiz, 0.0); row->L.assign(Nsiz, 0.0);
        for (int i = 0; i < Nsiz && i < (int)hw.size(); ++i) { row->hw[i] = hw[i]; row->L[i] = cum[i]; }
        for (int i = 0; i < Nsiz; ++i) row->L[i] = (row->L[i] < 1.0e-15) ? 1.0e20 : 1.0 / (Zeff * Zeff * Mass * row->L[i]);
    }
}

namespace {
// Atomic_elastic_sigma (Mott), Cross_sections.f90:3495-3514
double atomic_elastic_sigma(const Atom &a, double Ee) {
    double Zat = (double)a.Zat;
    double mec2e = g_me * g_cvel * g_cvel / g_e;
    double Zat137 = Zat / 137.0;
    double RyEe = g_Ry / Ee;
    double beta2 = 2.0 * Ee / mec2e;
    double pc = 1.7e-5 * std::pow(Zat, 2.0 / 3.0) * (1.0 - beta2) / beta2;
    double nc = pc * (1.13 + 3.76 * (Zat137 * Zat137) / beta2 * std::sqrt(Ee / (Ee + mec2e)));
    return g_Pi * g_a0 * g_a0 * Zat * (Zat + 1.0) / (nc * (nc + 1.0)) * RyEe * RyEe * 1e-16;
}
}  // namespace

void Elastic_cross_section(const Ctx &x, double Ee, int kind, double &EMFP, double &dEdx, DiffRow *row) {
    const Case &c = *x.c;
    EMFP = 1.34e16; dEdx = 0.0;
    if (c.numpar.kind_of_EMFP == 1) {
        double sz = 0, sp = 0;
        for (auto &a : c.atoms) { sz += a.Zat * a.Pers; sp += a.Pers; }
        double Zt = sz / sp;
        double Zeff = 1.0;
        if (c.numpar.CDF_elast_Zeff == 0) Zeff = 1.0 + equilibrium_charge_target(Ee, g_me, Zt, (Zt - 1.0), 0, 1.0);
        Tot_EMFP(x, Ee, kind, Zeff, EMFP, dEdx, row);
    } else {
        double Mass = (kind == 0) ? 1.0 : hole_mass_at(c, Ee);
        double Sigma_Tot = 0.0, sp = 0;
        for (auto &a : c.atoms) { Sigma_Tot = Sigma_Tot + atomic_elastic_sigma(a, Ee) * a.Pers; sp += a.Pers; }
        Sigma_Tot = Sigma_Tot / sp * Mass;
        EMFP = 1.0e8 / (Sigma_Tot * c.Matter.At_Dens);
    }
}

// Tot_Phot_IMFP, Cross_sections.f90:833-878 (CDF branch; EPDL branch needs EPDL2023.ALL)
void Tot_Phot_IMFP(const Ctx &x, double Ele, int Nat, int Nshl, double &Sigma, double &dEdx) {
    const Atom &at = x.c->atoms[Nat];
    if (Ele < at.Ip[Nshl]) { Sigma = 1e30; dEdx = Ele / Sigma; return; }
    double ImE = imewq_photon(at.Ritchi[Nshl], Ele);
    Sigma = g_cvel * g_h / (ImE * Ele * g_e) * 1e10;
    dEdx = Ele / Sigma;
}

// ---------------------------------------------------------------------------------------------
// effective charges
// ---------------------------------------------------------------------------------------------
double equilibrium_charge_target(double Ekin, double Mass, double ZSHI, double Zmean, int Kind_Zeff, double fixed_Zeff) {
    double vp = std::sqrt(2.0 * Ekin * g_e / Mass);
    double Zp = ZSHI;
    switch (Kind_Zeff) {
    case 1: return Zp * (1.0 - std::exp(-(vp / g_v0() / std::pow(Zp, 0.66666666))));
    case 2: { double c1 = 0.6, c2 = 0.45; return Zp * std::pow(1.0 + std::pow(vp / (std::pow(Zp, c2) * g_v0() * 4.0 / 3.0), -1.0 / c1), -c1); }
    case 3: {
        double Zt = Zmean;
        double c1 = 1.0 - 0.26 * std::exp(-Zt / 11.0 - (Zt - Zp) * (Zt - Zp) / 9.0);
        double vpvo = std::pow(Zp, -0.543) * vp / g_v0();
        double c2 = 1.0 + 0.03 * vpvo * std::log(Zt);
        double xx = c1 * std::pow(vpvo / c2 / 1.54, 1.0 + 1.83 / Zp);
        double x2 = xx * xx, x4 = x2 * x2;
        return Zp * (8.29 * xx + x4) / (0.06 / xx + 4.0 + 7.4 * xx + x4);
    }
    case 4: return fixed_Zeff;
    default: return Zp * (1.0 - std::exp(-(vp * 125.0 / g_cvel / std::pow(Zp, 0.66666666))));
    }
}

void equilibrium_charge_SHI(Ion &shi, const std::vector<Atom> &atoms) {   // Cross_sections.f90:2641-2680
    double vp = (shi.E > 0.0) ? std::sqrt(2.0 * shi.E * g_e / (shi.Mass * g_Mp)) : 0.0;
    double sz = 0, sp = 0;
    for (auto &a : atoms) { sz += a.Zat * a.Pers; sp += a.Pers; }
    double Zt = sz / sp, Zp = (double)shi.Zat;
    switch (shi.Kind_Zeff) {
    case 1: shi.Zeff = Zp * (1.0 - std::exp(-(vp / g_v0() / std::pow(Zp, 0.66666666)))); break;
    case 2: { double c1 = 0.6, c2 = 0.45; shi.Zeff = Zp * std::pow(1.0 + std::pow(vp / (std::pow(Zp, c2) * g_v0() * 4.0 / 3.0), -1.0 / c1), -c1); break; }
    case 3: {
        double c1 = 1.0 - 0.26 * std::exp(-Zt / 11.0 - (Zt - Zp) * (Zt - Zp) / 9.0);
        double vpvo = std::pow(Zp, -0.543) * vp / g_v0();
        double c2 = 1.0 + 0.03 * vpvo * std::log(Zt);
        double xx = c1 * std::pow(vpvo / c2 / 1.54, 1.0 + 1.83 / Zp);
        double x2 = xx * xx, x4 = x2 * x2;
        shi.Zeff = Zp * (8.29 * xx + x4) / (0.06 / xx + 4.0 + 7.4 * xx + x4); break;
    }
    case 4: shi.Zeff = shi.fixed_Zeff; break;
    default: shi.Zeff = Zp * (1.0 - std::exp(-(vp * 125.0 / g_cvel / std::pow(Zp, 0.66666666)))); break;
    }
}

// ---------------------------------------------------------------------------------------------
// SHI_TotIMFP, Cross_sections.f90:2452-2597 (point-like charge, CDF shells)
// ---------------------------------------------------------------------------------------------
void SHI_TotIMFP(const Ctx &x, Ion &shi, int Nat, int Nshl, double &Sigma, double &dEdx, MFP *dSedE) {
    const Case &c = *x.c;
    const Atom &at = c.atoms[Nat];
    const CDFosc &o = at.Ritchi[Nshl];
    equilibrium_charge_SHI(shi, c.atoms);
    double Ele = shi.E, MSHI = g_Mp * shi.Mass, Zeff = shi.Zeff;
    double Egap = c.atoms[0].Ip.back();
    double Emin = at.Ip[Nshl];
    if (Emin <= 1.0e-3) Emin = 1.0e-3;
    double Emax = 4.0 * Ele * g_me * MSHI / ((MSHI + g_me) * (MSHI + g_me));
    if (c.numpar.plasmon_Emax && Emin == Egap) {
        double Epl = std::sqrt(c.Matter.N_VB_el * c.Matter.At_Dens * 1e6 * g_h * g_h / (g_me * g_e0) + Egap * Egap);
        if (Epl >= Emax) Emax = Epl;
    }
    const int n = 1000;                              // m_N_grid_SHI
    double lo = 1e300, hi = -1e300;
    for (size_t i = 0; i < o.E0.size(); ++i) { lo = std::min(lo, o.E0[i] - 5.0 * o.Gamma[i]); hi = std::max(hi, o.E0[i] + 5.0 * o.Gamma[i]); }
    double E_low = std::max(lo, Emin), E_high = std::min(hi, Emax);
    size_t cap = 0;
    if (dSedE) {   // size of the cumulative table: count loop starting at Ip (:2512-2540)
        double E0 = at.Ip[Nshl];
        double E_low0 = std::max(lo, E0), E_high0 = std::min(hi, Emax);
        size_t k0 = 0;
        while (E0 <= Emax) { ++k0; E0 = E0 + define_dE(1, n, E0, true, E_low0, true, E_high0, 0.001); }
        cap = k0;
        dSedE->E.assign(k0, 0.0); dSedE->L.assign(k0, 0.0); dSedE->dEdx.assign(k0, 0.0);
    }
    double E = Emin, Ltot1 = 0.0, ddEdx = 0.0;
    const int set = x.set0[Nat] + Nshl;
    double Ltot0 = shi_diff_cross_section(x, set, Ele, MSHI, Emax, E);
    size_t i = 0;
    while (E <= Emax) {
        ++i;
        if (dSedE && i > cap) break;
        double dE = define_dE(1, n, E, true, E_low, true, E_high, 0.001);
        double a = E + dE / 2.0;
        double temp1 = shi_diff_cross_section(x, set, Ele, MSHI, Emax, a);
        double b = E + dE;
        double dL = shi_diff_cross_section(x, set, Ele, MSHI, Emax, b);
        double temp2 = dE / 6.0 * (Ltot0 + 4.0 * temp1 + dL);
        Ltot1 = Ltot1 + temp2;
        ddEdx = ddEdx + dE / 6.0 * (E * Ltot0 + a * 4.0 * temp1 + b * dL);
        Ltot0 = dL;
        if (dSedE) {
            dSedE->E[i - 1] = E;
            dSedE->L[i - 1] = (g_Pi * g_a0 * Ele * g_me) / (MSHI * Zeff * Zeff * Ltot1);
            dSedE->dEdx[i - 1] = (g_Pi * g_a0 * Ele * g_me) / (MSHI * Zeff * Zeff * ddEdx);
        }
        E = E + dE;
    }
    Sigma = 1.0 / (g_Pi * g_a0 * Ele) * MSHI / g_me * Zeff * Zeff * Ltot1;
    if (Sigma > 1e30) Sigma = 1e30;
    dEdx = 1.0 / (g_Pi * g_a0 * Ele) * MSHI / g_me * Zeff * Zeff * ddEdx;
}

// SHI_TotIMFP_BK, Cross_sections.f90:2748-2823: Brandt-Kitagawa ion.  The charge enters through the form factor inside the
// q-integral (no Zeff^2 in front); the stopping power is summed as E*(Simpson weight of the interval), as the reference does
void SHI_TotIMFP_BK(const Ctx &x, Ion &shi, int Nat, int Nshl, double &Sigma, double &dEdx) {
    const Case &c = *x.c;
    const Atom &at = c.atoms[Nat];
    const CDFosc &o = at.Ritchi[Nshl];
    equilibrium_charge_SHI(shi, c.atoms);
    const double Ele = shi.E, MSHI = g_Mp * shi.Mass, Zeff = shi.Zeff, Z_SHI = (double)shi.Zat;
    const double Egap = c.atoms[0].Ip.back();
    double Emin = at.Ip[Nshl];
    if (Emin <= 1.0e-3) Emin = 1.0e-3;
    double Emax = 4.0 * Ele * g_me * MSHI / ((MSHI + g_me) * (MSHI + g_me));
    if (c.numpar.plasmon_Emax && Emin == Egap) {
        double Epl = std::sqrt(c.Matter.N_VB_el * c.Matter.At_Dens * 1e6 * g_h * g_h / (g_me * g_e0) + Egap * Egap);
        if (Epl >= Emax) Emax = Epl;
    }
    const int n = 1000;                              // m_N_grid_SHI
    double lo = 1e300, hi = -1e300;
    for (size_t i = 0; i < o.E0.size(); ++i) { lo = std::min(lo, o.E0[i] - 5.0 * o.Gamma[i]); hi = std::max(hi, o.E0[i] + 5.0 * o.Gamma[i]); }
    const double E_low = std::max(lo, Emin), E_high = std::min(hi, Emax);
    const int set = x.set0[Nat] + Nshl;
    double E = Emin, Ltot1 = 0.0, ddEdx = 0.0;
    double Ltot0 = shi_diff_cross_section_bk(x, set, Ele, MSHI, Emax, E, Z_SHI, Zeff);
    while (E <= Emax) {
        double dE = define_dE(1, n, E, true, E_low, true, E_high, 0.001);
        double a = E + dE / 2.0;
        double temp1 = shi_diff_cross_section_bk(x, set, Ele, MSHI, Emax, a, Z_SHI, Zeff);
        double b = E + dE;
        double dL = shi_diff_cross_section_bk(x, set, Ele, MSHI, Emax, b, Z_SHI, Zeff);
        double temp2 = dE / 6.0 * (Ltot0 + 4.0 * temp1 + dL);
        Ltot1 = Ltot1 + temp2;
        ddEdx = ddEdx + E * temp2;
        Ltot0 = dL;
        E = E + dE;
    }
    Sigma = 1.0 / (g_Pi * g_a0 * Ele) * MSHI / g_me * Ltot1;
    if (Sigma > 1e30) Sigma = 1e30;
    dEdx = 1.0 / (g_Pi * g_a0 * Ele) * MSHI / g_me * ddEdx;
}

// SHI_Total_IMFP, Cross_sections.f90:2431-2446: the tabulated ion MFPs follow Kind_ion; the differential table of the given
// ion energy is always the point-charge one (MAIN.f90:233 calls SHI_TotIMFP directly)
void SHI_Total_IMFP(const Ctx &x, Ion &shi, int Nat, int Nshl, double &Sigma, double &dEdx) {
    if (shi.Kind_ion == 1) SHI_TotIMFP_BK(x, shi, Nat, Nshl, Sigma, dEdx);
    else SHI_TotIMFP(x, shi, Nat, Nshl, Sigma, dEdx, nullptr);
}

// ---------------------------------------------------------------------------------------------
// sum rules and the single-pole phonon CDF, Cross_sections.f90:554-826
// ---------------------------------------------------------------------------------------------
double w_plasma(double At_dens, double Mass) {
    if (Mass > 0) return At_dens * g_e * g_e / (g_e0 * Mass);
    return At_dens * g_e * g_e / (g_e0 * g_me);
}

namespace {
typedef std::complex<double> cd;
double Int_Ritchi_x(double A, double E, double Gamma, double xx) {
    const double sq2 = std::sqrt(2.0);
    const cd oneI(0.0, 1.0);
    double disc = -(2.0 * E) * (2.0 * E) + Gamma * Gamma;
    cd Sc = (disc >= 0.0) ? cd(std::sqrt(disc), 0.0) : cd(0.0, std::sqrt(std::fabs(disc)));
    cd Gc(Gamma * Gamma - 2.0 * E * E, 0.0);
    cd s_plus = sq2 * std::sqrt(Gc + Gamma * Sc);
    cd s_minus = sq2 * std::sqrt(Gc - Gamma * Sc);
    cd Bc = ((Gamma * Sc) == cd(0.0, 0.0)) ? cd(0.0, 0.0) : Gc / (Gamma * Sc);
    cd arg = 2.0 * xx / s_minus, arg2 = 2.0 * xx / s_plus;
    cd Ic = (A * Gamma * (0.5 * oneI * (std::log(1.0 - oneI * arg) - std::log(1.0 + oneI * arg)) / s_minus * (1.0 - Bc) +
                          0.5 * oneI * (std::log(1.0 - oneI * arg2) - std::log(1.0 + oneI * arg2)) / s_plus * (1.0 + Bc)));
    return Ic.real();
}
double Int_Ritchi_p_x(double A, double E, double Gamma, double xx) {
    const double sq2 = std::sqrt(2.0);
    const cd oneI(0.0, 1.0);
    double disc = -(2.0 * E) * (2.0 * E) + Gamma * Gamma;
    cd Sc = (disc >= 0.0) ? cd(std::sqrt(disc), 0.0) : cd(0.0, std::sqrt(std::fabs(disc)));
    cd Gc(Gamma * Gamma - 2.0 * E * E, 0.0);
    cd s_plus = std::sqrt(Gc + Gamma * Sc);
    cd s_minus = std::sqrt(Gc - Gamma * Sc);
    cd arg = sq2 * xx / s_minus, arg2 = sq2 * xx / s_plus;
    cd term1 = (0.5 * oneI * (std::log(1.0 - oneI * arg) - std::log(1.0 + oneI * arg))) / s_minus;
    cd term2 = (0.5 * oneI * (std::log(1.0 - oneI * arg2) - std::log(1.0 + oneI * arg2))) / s_plus;
    cd In = ((term1 - term2) == cd(0.0, 0.0)) ? cd(sq2 * A, 0.0) : sq2 * A / Sc * (term1 - term2);
    return In.real();
}
}  // namespace

// define_alpha, Reading_files_and_parameters.f90:2199-2206: weight of the delta function that replaces a Ritchie oscillator
double define_alpha(double Ai, double Gammai, double E0i, double x_min) {
    return Int_Ritchi_x(Ai, E0i, Gammai, 1.0e30) - Int_Ritchi_x(Ai, E0i, Gammai, x_min);
}

void sumrules(const CDFosc &o, double &ksum, double &fsum, double x_min, double Omega) {
    double ne = 0.0, f = 0.0;
    for (size_t j = 0; j < o.A.size(); ++j) {
        ne = ne + Int_Ritchi_x(o.A[j], o.E0[j], o.Gamma[j], 1e10) - Int_Ritchi_x(o.A[j], o.E0[j], o.Gamma[j], x_min);
        f = f + Int_Ritchi_p_x(o.A[j], o.E0[j], o.Gamma[j], 1e10) - Int_Ritchi_p_x(o.A[j], o.E0[j], o.Gamma[j], x_min);
    }
    ksum = 2.0 * g_e * g_e / (g_Pi * Omega * g_h * g_h) * ne;
    fsum = f * 2.0 / g_Pi;
}

void get_single_pole(Case &c) {
    // Cross_sections.f90:554-686.  Part 1, electronic single-pole CDF: files that leave their shells to the atomic database
    // (kind_of_CDF = 1; they need EADL2023.ALL).  Part 2: phonon CDF.
    double N_at_mol = 0; for (auto &a : c.atoms) N_at_mol += a.Pers;
    if (c.numpar.kind_of_CDF == 1) {
        // Part 1 (:570-633): one oscillator per shell.  Valence band: E0 = plasmon energy of the valence electrons, Gamma = E0;
        // core shells: E0 = Ip + 10 eV, Gamma = E0; A from the k-sum rule (electrons of the shell).
        for (size_t i = 0; i < c.atoms.size(); ++i) {
            Atom &a = c.atoms[i];
            const int N_shl = a.nshl();
            for (int j = 0; j < N_shl; ++j) {
                CDFosc &o = a.Ritchi[(size_t)j];
                double ksum, fsum;
                if (i == 0 && j == N_shl - 1) {
                    if (c.numpar.VB_CDF_defined) continue;
                    const double NVB = a.Nel[(size_t)j] / N_at_mol;
                    double Omega = w_plasma(1e6 * c.Matter.At_Dens * NVB);
                    o.E0.assign(1, std::sqrt((g_h / g_e) * (g_h / g_e) * Omega));
                    o.Gamma.assign(1, o.E0[0]); o.A.assign(1, 1.0);
                    Omega = w_plasma(1e6 * c.Matter.At_Dens);
                    sumrules(o, ksum, fsum, a.Ip[(size_t)j], Omega);
                    o.A[0] = NVB / ksum;
                } else {
                    const double NVB = a.Nel[(size_t)j], contrib = a.Pers / N_at_mol;
                    o.E0.assign(1, a.Ip[(size_t)j] + 10.0);
                    o.Gamma.assign(1, o.E0[0]); o.A.assign(1, 1.0);
                    const double Omega = w_plasma(1e6 * c.Matter.At_Dens * contrib);
                    sumrules(o, ksum, fsum, a.Ip[(size_t)j], Omega);
                    o.A[0] = NVB / ksum;
                }
            }
        }
    }
    if (c.numpar.kind_of_CDF_ph == 1) {
        double qdebye = std::pow(6.0 * g_Pi * g_Pi * (c.Matter.At_Dens * 1e6), 0.33333333);
        double E_debye = g_h * c.Matter.Vsound * qdebye / g_e;
        double m_Pi_6 = std::pow(g_Pi / 6.0, 1.0 / 3.0);
        double E_einstein = E_debye * m_Pi_6;
        CDFosc &p = c.CDF_Phonon;
        p.E0.assign(1, 2.0 * E_einstein);
        p.Gamma.assign(1, p.E0[0] * 0.5);
        p.A.assign(1, 1.0);
        double sm = 0; for (auto &a : c.atoms) sm += a.Pers * a.Mass;
        double Mean_Mass = sm * g_Mp / N_at_mol;
        double Omega = w_plasma(1e6 * c.Matter.At_Dens / N_at_mol, Mean_Mass);
        double ksum, fsum;
        sumrules(p, ksum, fsum, 1.0e-8, Omega);
        p.A[0] = N_at_mol / ksum;
    } else if (c.numpar.CDF_elast_Zeff == 2 || c.numpar.CDF_elast_Zeff == 3) {
        // user-provided phonon CDF with dynamical screening (:641-655): renormalised to the number of atoms per molecule -- the
        // optical charge it was fitted with is replaced by the screened nuclear charge inside the cross section.  Done once:
        // the tables may be built (or read from the cache) several times for one case
        if (!c.phonon_renormalised) {
            CDFosc &p = c.CDF_Phonon;
            double sm = 0; for (auto &a : c.atoms) sm += a.Pers * a.Mass;
            double Mean_Mass = sm * g_Mp / N_at_mol;
            double Omega = w_plasma(1e6 * c.Matter.At_Dens / N_at_mol, Mean_Mass);
            double ksum, fsum;
            sumrules(p, ksum, fsum, 1.0e-8, Omega);
            for (auto &A : p.A) A = A * N_at_mol / ksum;
            c.phonon_renormalised = true;
        }
    }
}

}  // namespace trk3
