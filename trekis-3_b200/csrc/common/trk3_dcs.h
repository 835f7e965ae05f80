// trk3_dcs.h -- the integrands of the CDF table builder, shared by the host builder (csrc/host/cdf.cpp, g++) and the
// GPU evaluator (csrc/cuda/tables_gpu.cu, nvcc): Ritchie-Howie loss function with finite-q extension and the three
// q-integrals of Cross_sections.f90
//     Diff_cross_section          :2217-2282   electrons / holes on a CDF shell
//     SHI_Diff_cross_section      :2683-2744   swift heavy ion on a CDF shell
//     Diff_cross_section_phonon   :3142-3300   electrons / holes on the phonon CDF (CDF_elast_Zeff 0/1)
// with Imewq :397-486 (Extend_E0_to_finite_q :346-393) and the effective-mass lookup :428-434.
//
// ONE source for both sides on purpose: every value is produced by the same sequence of IEEE-754 double operations
// (+ - * / sqrt are correctly rounded on CPU and GPU; the host is built with -ffp-contract=off and the device file
// with -fmad=false), so the GPU-built tables are identical to the host-built ones, not merely close.  Only
// kind_of_DR = 3 (pow) and a target temperature > 0 (exp) go through libm functions that may differ in the last bit.
#ifndef TRK3_DCS_H
#define TRK3_DCS_H
#include <math.h>
#include <stdint.h>
#include "../../../include/trekis3_gpu.h"

#if defined(__CUDACC__)
#define DCS_HD __host__ __device__ inline
#else
#define DCS_HD inline
#endif

namespace trk3dcs {

// Universal_Constants.f90:24-107 (the values the host builder uses)
#define DCS_PI 3.1415926535897932384626433832795
#define DCS_GE 1.602176487e-19
#define DCS_ME 9.1093821545e-31
#define DCS_H 1.05457162853e-34
#define DCS_KB 11604.0
#define DCS_A0 0.5291772085936

// Find_in_monotonous_1D_array, Reading_files_and_parameters.f90:3433-3494 (1-based result)
DCS_HD int find_monoton_1d(const double *A, int N, double v) {
    int i_1 = 1, i_2 = N;
    int i_cur = (int)floor((i_1 + i_2) / 2.0);
    double temp_val = A[i_cur - 1];
    if (v < A[0]) i_cur = 0;
    else if (v >= A[N - 1]) i_cur = N - 1;
    else {
        for (;;) {
            if (i_1 == i_2 - 1) break;
            if (temp_val <= v) i_1 = i_cur; else i_2 = i_cur;
            i_cur = (int)floor((i_1 + i_2) / 2.0);
            temp_val = A[i_cur - 1];
        }
    }
    return i_cur + 1;
}

// Imewq for electrons / holes / SHI on oscillator set `s` (Cross_sections.f90:397-486).  The effective-mass lookup
// depends only on q and is hoisted out of the oscillator loop (same arithmetic).
DCS_HD double imewq(const trk3_dcs_ctx &x, int s, double hw, double dq) {
    double hq2 = DCS_H * DCS_H * dq * dq;
    double Mass;
    if (x.mass_from_dos) {
        double qlim = fabs(dq) * sqrt(DCS_GE);
        if (qlim <= x.k[x.n_k - 1]) { int j = find_monoton_1d(x.k, x.n_k, qlim); Mass = x.effm[j - 1]; }
        else Mass = 1.0;
    } else if (x.El_eff_mass > 0) Mass = x.El_eff_mass;
    else Mass = 1.0;
    double sqq = hq2 / (2.0 * Mass * DCS_ME);
    double dE2 = hw * hw;
    double ImE = 0.0;
    const int i0 = x.osc_off[s], i1 = x.osc_off[s + 1];
    for (int i = i0; i < i1; ++i) {
        double E = x.osc_E0[i], Gamma = x.osc_G[i], E0, Gamma1;
        switch (x.kind_DR) {
        case 2: E0 = sqrt(E * E + x.v_f * x.v_f * hq2 * 0.3333333333333 + sqq * sqq); Gamma1 = Gamma; break;
        case 3: E0 = pow(pow(E, 0.666666666666) + pow(sqq, 0.666666666666), 1.5); Gamma1 = sqrt(Gamma * Gamma + sqq * sqq); break;
        default: E0 = E + sqq; Gamma1 = Gamma; break;
        }
        double E02 = E0 * E0;
        ImE = ImE + x.osc_A[i] * Gamma1 * hw / ((dE2 - E02) * (dE2 - E02) + Gamma1 * Gamma1 * dE2);
    }
    return ImE;
}

// phonon loss function, :3409-3433 (Extend_E0_to_finite_q with the target's mean atomic mass)
DCS_HD double imewq_phonon(const trk3_dcs_ctx &x, int s, double hw, double hq, double Mtarget) {
    double hq2 = DCS_H * DCS_H * hq * hq;
    double dE2 = hw * hw, ImE = 0.0;
    const int i0 = x.osc_off[s], i1 = x.osc_off[s + 1];
    for (int i = i0; i < i1; ++i) {
        double E0 = x.osc_E0[i] + hq2 / (2.0 * Mtarget);
        double G = x.osc_G[i], E02 = E0 * E0;
        ImE = ImE + x.osc_A[i] * G * hw / ((dE2 - E02) * (dE2 - E02) + G * G * dE2);
    }
    return ImE;
}

// Diff_cross_section, Cross_sections.f90:2217-2282
DCS_HD double diff_cross_section(const trk3_dcs_ctx &x, int s, double Ee, double dE, double Mass) {
    double pre = sqrt(2.0 * Mass * DCS_ME) / DCS_H;
    double qmin, qmax;
    if (dE > Ee) { qmin = pre * sqrt(Ee); qmax = pre * sqrt(Ee); }
    else { qmin = pre * (sqrt(Ee) - sqrt(Ee - dE)); qmax = pre * (sqrt(Ee) + sqrt(Ee - dE)); }
    double dLs = 0.0, hq = qmin, dLs0 = 0.0;
    const double n = 100.0;                       // m_N_p_grid_SHI
    while (hq < qmax) {
        double dq = hq / n;
        double a = hq + dq / 2.0;
        double temp1 = imewq(x, s, dE, a);
        double b = hq + dq;
        double dL = imewq(x, s, dE, b);
        dLs = dLs + dq / 6.0 * (dLs0 + 4.0 * temp1 + dL) / hq;
        dLs0 = dL;
        hq = hq + dq;
    }
    double T_fact = 1.0;
    if (x.temp > 0.0) T_fact = 1.0 / (1.0 - exp(-dE / x.temp * DCS_KB));
    return 1.0 / (DCS_PI * DCS_A0 * Ee) * dLs * T_fact;
}

// SHI_Diff_cross_section, Cross_sections.f90:2683-2744
DCS_HD double shi_diff_cross_section(const trk3_dcs_ctx &x, int s, double Ee, double MSHI, double Emax, double hw) {
    double qmin = (Ee > 0.0) ? hw / DCS_H / sqrt(2.0 * Ee / MSHI) : 0.0;
    if (!(Emax > 0.0)) return 0.0;
    double qmax = sqrt(2.0 * DCS_ME) / DCS_H * sqrt(Emax);
    double dLs = 0.0, hq = qmin, dLs0 = 0.0;
    const double n = 100.0;
    while (hq < qmax) {
        double dq = hq / n;
        double a = hq + dq / 2.0;
        double temp1 = imewq(x, s, hw, a);
        double b = hq + dq;
        double dL = imewq(x, s, hw, b);
        dLs = dLs + dq / 6.0 * (dLs0 + 4.0 * temp1 + dL) / hq;
        dLs0 = dL;
        hq = hq + dq;
    }
    double T_fact = 1.0;
    if (x.temp > 0.0) T_fact = 1.0 / (1.0 - exp(-hw / x.temp * DCS_KB));
    return dLs * T_fact;
}

// Brand_Kitagawa, Cross_sections.f90:2877-2889: form factor of a partially stripped ion (Z = ionisation deficit / Z_SHI)
DCS_HD double brandt_kitagawa(double hq, double Z_SHI, double Zeff) {
    const double a = 0.2400519147;
    double Z = (Z_SHI - Zeff) / Z_SHI;
    double kl = hq * (DCS_A0 * 1e-10 * sqrt(DCS_GE)) * 2.0 * a * pow(Z, 2.0 / 3.0) / (pow(Z_SHI, 2.0 / 3.0) * (1.0 - Z / 7.0));
    double kl2 = kl * kl;
    return Z_SHI * (1.0 - Z + kl2) / (1.0 + kl2);
}

// SHI_Diff_cross_section_BK, Cross_sections.f90:2826-2874: the loss function weighted with the squared form factor
DCS_HD double shi_diff_cross_section_bk(const trk3_dcs_ctx &x, int s, double Ee, double MSHI, double Emax, double hw, double Z_SHI, double Zeff) {
    double qmin = hw / DCS_H / sqrt(2.0 * Ee / MSHI);
    double qmax = sqrt(2.0 * DCS_ME) / DCS_H * sqrt(Emax);
    double dLs = 0.0, hq = qmin, dLs0 = 0.0;
    const double n = 100.0;                      // m_N_p_grid_SHI
    while (hq < qmax) {
        double dq = hq / n;
        double a = hq + dq / 2.0;
        double rho = brandt_kitagawa(a, Z_SHI, Zeff);
        double temp1 = imewq(x, s, hw, a) * rho * rho;
        double b = hq + dq;
        rho = brandt_kitagawa(b, Z_SHI, Zeff);
        double dL = imewq(x, s, hw, b) * rho * rho;
        dLs = dLs + dq / 6.0 * (dLs0 + 4.0 * temp1 + dL) / hq;
        dLs0 = dL;
        hq = hq + dq;
    }
    return dLs / (1.0 - exp(-hw / x.temp * DCS_KB));      // :2873, at 0 K: exp(-inf) = 0
}

// ---- dynamical screening of the elastic cross section (CDF_elast_Zeff = 2 / 3), Cross_sections.f90:61-82, 184-344, 3216-3395
#define DCS_CVEL 299792458.0
#define DCS_E0 8.854187817620e-12
// mass and finite-q extension of one oscillator exactly as in imewq (Extend_E0_to_finite_q :346-393)
DCS_HD double dcs_mass_at_q(const trk3_dcs_ctx &x, double dq) {
    if (x.mass_from_dos) {
        double qlim = fabs(dq) * sqrt(DCS_GE);
        if (qlim <= x.k[x.n_k - 1]) { int j = find_monoton_1d(x.k, x.n_k, qlim); return x.effm[j - 1]; }
        return 1.0;
    }
    return (x.El_eff_mass > 0) ? x.El_eff_mass : 1.0;
}
// Reewq + One_Reewq, :253-344: Re(-1/eps) of oscillator set s corresponding to the fitted loss function
DCS_HD double reewq(const trk3_dcs_ctx &x, int s, double hw, double dq) {
    double hq2 = DCS_H * DCS_H * dq * dq;
    double Mass = dcs_mass_at_q(x, dq);
    double sqq = hq2 / (2.0 * Mass * DCS_ME);
    double dE2 = hw * hw, ReE = 0.0;
    const int i0 = x.osc_off[s], i1 = x.osc_off[s + 1];
    for (int i = i0; i < i1; ++i) {
        double E = x.osc_E0[i], Gamma = x.osc_G[i], E0, Gamma1;
        switch (x.kind_DR) {
        case 2: E0 = sqrt(E * E + x.v_f * x.v_f * hq2 * 0.3333333333333 + sqq * sqq); Gamma1 = Gamma; break;
        case 3: E0 = pow(pow(E, 0.666666666666) + pow(sqq, 0.666666666666), 1.5); Gamma1 = sqrt(Gamma * Gamma + sqq * sqq); break;
        default: E0 = E + sqq; Gamma1 = Gamma; break;
        }
        double E02 = E0 * E0;
        ReE = ReE + x.osc_A[i] * (E02 - dE2) / ((dE2 - E02) * (dE2 - E02) + (Gamma1 * Gamma1) * dE2);
    }
    return -(1.0 - ReE);
}
// |CDF| of one shell, construct_CDF :184-248: eps = (-Re, Im) / (Re^2 + Im^2) of the loss function; a core shell does not
// respond below its ionisation potential (hw + recoil <= Ip): eps = 1
DCS_HD double abs_shell_cdf(const trk3_dcs_ctx &x, int s, double hw, double hq) {
    if (s != x.vb_set) {
        const double Ip = x.scr[1 + 10 * (int)x.scr[0] + 2 * s + 1];
        const double Q = (DCS_H * hq) * (DCS_H * hq) / (2.0 * DCS_ME);
        if (hw + Q <= Ip) return 1.0;
    }
    const double ImE = imewq(x, s, hw, hq), ReE = reewq(x, s, hw, hq);
    const double den = ReE * ReE + ImE * ImE;
    if (!(fabs(den) > 1.0e-12)) return 1.0;
    const double re = -ReE / den, im = ImE / den;
    return sqrt(re * re + im * im);
}
// form_factor, :61-82 (PENELOPE analytical form factors; q in kg m/s)
DCS_HD double form_factor(double q, const double *a, double Z) {
    const double mc = DCS_ME * DCS_CVEL;
    const double x = q / mc * 20.6074224164, x2 = x * x, x4 = x2 * x2;
    const double demf = 1.0 + a[3] * x2 + a[4] * x4;
    double FF = Z * (1.0 + a[0] * x2 + a[1] * x * x2 + a[2] * x4) / (demf * demf);
    if (Z > 10.0 && FF < 2.0) {
        const double g_alpha = DCS_GE * DCS_GE / (DCS_H * DCS_CVEL * 4.0 * DCS_PI * DCS_E0);
        const double al = g_alpha * (Z - 5.0 / 16.0);
        const double b = sqrt(1.0 - al * al);
        const double CapQ = q / (2.0 * mc * al);
        const double Fk = sin(2.0 * b * atan(CapQ)) / (b * CapQ * pow(1.0 + CapQ * CapQ, b));
        if (FF < Fk) FF = Fk;
    }
    return FF;
}
// (Z/eps)^2 per atom at transferred momentum hq for a particle of energy Ee: get_screening_ff :3303-3367 (screening 2, with
// the reference's empirical momentum 0.5 (p_e + hq)) and get_screening_all :3372-3406 (screening 3)
DCS_HD double elastic_screening(const trk3_dcs_ctx &x, double Ee, double Mass, double hw, double hq) {
    const int nat = (int)x.scr[0];
    double Zmol = 0.0, pers = 0.0, contrib = 0.0;
    for (int i = 0; i < nat; ++i) { const double *A = x.scr + 1 + 10 * i; Zmol += A[0] * A[1]; pers += A[1]; }
    const double *sets = x.scr + 1 + 10 * nat;
    if (x.screening == 2) {
        const double p_e = 0.5 * (1.0 * sqrt(2.0 * Mass * DCS_ME * Ee * DCS_GE) + hq * DCS_H * sqrt(DCS_GE));
        const double p_e_prime = 0.5 * (1.0 * sqrt(2.0 * Mass * DCS_ME * Ee) / DCS_H + hq);
        const double acdf = abs_shell_cdf(x, x.vb_set, hw, p_e_prime);
        for (int i = 0; i < nat; ++i) {
            const double *A = x.scr + 1 + 10 * i;
            double FF = form_factor(p_e, A + 5, A[0]);
            if (FF > A[2]) FF = A[2];                          // exclude the valence part of the charge
            contrib = contrib - FF * A[1];
            if (i == 0) contrib = contrib + sets[2 * x.vb_set] * (1.0 / acdf - 1.0);
        }
    } else {
        for (int i = 0; i < nat; ++i) {
            const double *A = x.scr + 1 + 10 * i;
            const int set0 = (int)A[3], nsh = (int)A[4];
            // sic (:3253): the CDF of atom i's shells is built from shell number size(Ip of atom i) of the FIRST atom
            const int s_used = (int)x.scr[1 + 3] + nsh - 1;
            const double acdf = abs_shell_cdf(x, s_used, hw, hq);
            for (int j = 0; j < nsh; ++j) {
                const int s = set0 + j;
                const double N_el = (s != x.vb_set) ? sets[2 * s] * A[1] : sets[2 * s];
                contrib = contrib + N_el * (1.0 / acdf - 1.0);
            }
        }
    }
    const double scr = (Zmol + contrib) / pers;
    return scr * scr;
}

// Diff_cross_section_phonon, Cross_sections.f90:3142-3300 (CDF_elast_Zeff 0/1: screening = 1; 2/3: elastic_screening)
DCS_HD double diff_cross_section_phonon(const trk3_dcs_ctx &x, int s, double Ee, double dE, double Mtarget, double Mass, double Ttarget, double pref) {
    const double eps = 1.0e-12;
    double pre = sqrt(2.0 * Mass * DCS_ME) / DCS_H;
    double qmin;
    if (fabs(dE) < eps) return 1.31e30;
    else if (dE > (Ee - eps)) qmin = pre * sqrt(Ee);
    else qmin = pre * (sqrt(Ee) - sqrt(fabs(Ee - dE)));
    double qmax = pref * pre * (sqrt(Ee) + sqrt(fabs(Ee - dE)));
    double dLs = 0.0, hq = qmin, dLs0 = 0.0;
    const double n = 100.0;
    while (fabs(hq) < fabs(qmax)) {
        double dq = hq / n;
        double a = hq + dq / 2.0;
        double temp1 = imewq_phonon(x, s, dE, a, Mtarget);
        double b = hq + dq;
        double dL = imewq_phonon(x, s, dE, b, Mtarget);
        double Pot = ((x.screening >= 2) ? elastic_screening(x, Ee, Mass, dE, hq) : 1.0) / hq;
        dLs = dLs + dq / 6.0 * (dLs0 + 4.0 * temp1 + dL) * Pot;
        dLs0 = dL;
        hq = hq + dq;
    }
    if (Ttarget > 1.0e-6) return 1.0 / (DCS_PI * DCS_A0 * Ee) * dLs / (1 - exp(-dE / Ttarget * DCS_KB));
    return 1.0 / (DCS_PI * DCS_A0 * Ee) * dLs;
}

// one request of a task (see trk3_dcs_task in include/trekis3_gpu.h)
DCS_HD double eval_request(const trk3_dcs_ctx &x, const trk3_dcs_task &t, double hw) {
    switch (t.type) {
    case TRK3_DCS_INELASTIC: return diff_cross_section(x, t.set, t.Ee, hw, t.Mass);
    case TRK3_DCS_PHONON: return diff_cross_section_phonon(x, t.set, t.Ee, hw, t.p1, t.Mass, t.p2, t.p3);
    case TRK3_DCS_SHI_BK: return shi_diff_cross_section_bk(x, t.set, t.Ee, t.p1, t.p2, hw, t.Mass, t.p3);
    default: return shi_diff_cross_section(x, t.set, t.Ee, t.p1, t.p2, hw);
    }
}

}  // namespace trk3dcs
#endif
