// trk3_delta.h -- the delta-function CDF model of the inelastic electron / valence-hole cross section (kind_of_DR = 4), shared by
// the host table builder (csrc/host/cdf.cpp: TotIMFP, Cross_sections.f90:952-956) and the device code that samples the
// transferred energy (csrc/cuda/physics.cuh: delta_transfer, Cross_sections.f90:1894-1895, 2051-2123).
//
// Every oscillator (E0, A, Gamma) of a shell's Ritchie CDF is replaced by a delta function at E0 with the weight
//     alpha = Int_Ritchi_x(infinity) - Int_Ritchi_x(Ip)                (define_alpha, Reading_files_and_parameters.f90:2199-2206)
// for which the cross section integrated over the transferred energy up to W has a closed form (integral_CS :1706-1712).
// Near the threshold, where the delta model does not apply (E <= 1.01 Eeq, Eeq = the energy at which W_max = W_min), the reference
// replaces the cross section by a straight line through the threshold and the model's value at 1.01 Eeq (Find_linear_a_b).
// The reference calls these routines for electrons AND valence holes with M = mt = g_me and identical = .true.
// (TotIMFP :953-956, get_inelastic_energy_transfer :2098): the general (ion-electron) forms of W_min / W_max /
// minimal_sufficient_E are restated too, so that the functions read like their originals, but only that case is exercised.
#ifndef TRK3_DELTA_H
#define TRK3_DELTA_H
#include <math.h>

#if defined(__CUDACC__)
#define DLT_HD __host__ __device__ inline
#else
#define DLT_HD inline
#endif
#ifndef DLT_LOG            // the device code routes its logarithms through ONE out-of-line copy (physics.cuh: m_log)
#define DLT_LOG(x) log(x)
#endif

namespace trk3delta {

#define DLT_PI 3.1415926535897932384626433832795
#define DLT_GE 1.602176487e-19
#define DLT_ME 9.1093821545e-31
#define DLT_CVEL 299792458.0
#define DLT_A0 0.5291772085936
#define DLT_ME_EV (0.51099906 * 1.0e6)          // g_me_eV, Universal_Constants.f90:31-32

DLT_HD double rest_energy(double M0) { return M0 * DLT_CVEL * DLT_CVEL / DLT_GE; }                 // :1557-1561
// W_min :1564-1590 (the optional E0 is always present on this path)
DLT_HD double W_min(double Ip, double Mc2, double mtc2, double E, double E0) {
    double Wmin = Ip, E0min;
    if (fabs(Mc2 - mtc2) / Mc2 < 1.0e-6) E0min = E0 * (1.0 - 0.25 * E0 / E);
    else if (E < 1.0e9 * E0) {
        const double mtc2_2 = mtc2 * mtc2;
        E0min = (-2.0 * Mc2 * sqrt(E) * (Mc2 * sqrt(E) - sqrt(Mc2 * E0 * mtc2 - E0 * mtc2_2 + mtc2_2 * E)) / (Mc2 - mtc2) + 2.0 * Mc2 * E - E0 * mtc2) / (Mc2 - mtc2);
    } else E0min = E0 * (1.0 - 0.25 * E0 / E * Mc2 / mtc2);
    return (Wmin > E0min) ? Wmin : E0min;
}
// W_max :1593-1616 (no phonon argument on this path)
DLT_HD double W_max(double Mc2, double mtc2, bool identical, double E, double Ip) {
    if (identical) return (E + Ip) * 0.5;
    const double Emtc22 = 2.0 * mtc2 * E, Mm = Mc2 + mtc2;
    return Emtc22 * (E + Mc2 * 2.0) / (Emtc22 + Mm * Mm);
}
// find_Wmax_equal_Wmin :1619-1649
DLT_HD double find_Wmax_equal_Wmin(double Mc2, double mtc2, bool identical, double Ip, double E0) {
    if (!identical) {
        const double Mm = Mc2 + mtc2, Mm4 = Mm * Mm * Mm * Mm;
        return E0 * (Mc2 / mtc2) * (Mm4 / fabs(4.0 * Mc2 * Mc2 - Mm4));
    }
    return 5.0 / 4.0 * E0 - Ip / 2.0 + 0.25 * sqrt(17.0 * E0 * E0 - 12.0 * Ip * E0 + 4.0 * Ip * Ip);
}
// minimal_sufficient_E :1653-1668
DLT_HD double minimal_sufficient_E(double Ip, double Mc2, double mtc2) {
    if (fabs(Mc2 - mtc2) / Mc2 < 1.0e-6) return Ip;
    const double Mc2_2 = Mc2 * Mc2, mtc2_2 = mtc2 * mtc2;
    return 0.50 * (Ip * Mc2 - 2.0 * mtc2 * Mc2 + 2.0 * Ip * mtc2 + sqrt(-Ip * Ip * Mc2_2 + 4.0 * mtc2_2 * Mc2_2 - 6.0 * mtc2_2 * Mc2 * Ip + 2.0 * Ip * Ip * mtc2_2 + 2.0 * Ip * Mc2 * Mc2_2)) / (Mc2 - Ip);
}
// P_prefactor :1525-1535 with velosity_from_kinetic_energy / beta_factor :1538-1556
DLT_HD double P_prefactor(double M, double E, double nat) {
    double v;
    if (M < 1.0e-10 * DLT_ME) v = DLT_CVEL;
    else { const double fact = E / rest_energy(M) + 1.0; v = DLT_CVEL * sqrt(1.0 - 1.0 / (fact * fact)); }
    const double beta = v / DLT_CVEL;
    return 1.0e24 / (DLT_PI * DLT_A0 * nat * DLT_ME_EV * (beta * beta));
}
// integral_CS :1706-1712
DLT_HD double integral_CS(double alpha, double Mc2, double mtc2, double E0, double W) {
    const double Mc22 = 2.0 * Mc2;
    return alpha / (DLT_ME_EV * (Mc22 - E0)) * ((Mc22 - mtc2) * DLT_LOG(Mc22 + W - E0) + Mc22 * mtc2 / E0 * (DLT_LOG(W) - DLT_LOG(fabs(W - E0))));
}
// integrated_delta_CDF_CS :1690-1704
DLT_HD double integrated_delta_CDF_CS(double alpha, double Mc2, double E0, double mtc2, double W, double Ip, double E) {
    const double Wmin = W_min(Ip, Mc2, mtc2, E, E0);
    if (W < Wmin || E <= Ip) return 0.0;
    return integral_CS(alpha, Mc2, mtc2, E0, W);
}
// int_energy_loss :1715-1719
DLT_HD double int_energy_loss(double alpha, double Mc2, double mtc2, double E0, double Wmax, double Wmin) {
    return -alpha * (mtc2 * DLT_LOG(fabs((Wmax - E0) / (Wmin - E0))) + (2.0 * Mc2 - mtc2) * DLT_LOG(fabs((-Wmax + E0 - 2.0 * Mc2) / (-Wmin + E0 - 2.0 * Mc2)))) / DLT_ME_EV;
}
// Find_linear_a_b :1671-1688
DLT_HD void Find_linear_a_b(double alpha, double M, double Mc2, double mtc2, double E0, double Ip, bool identical, double nat, double Eeq, double &a, double &b) {
    const double Wmin_lim = W_min(Ip, Mc2, mtc2, Eeq, E0), Wmax_lim = W_max(Mc2, mtc2, identical, Eeq, Ip);
    const double P = P_prefactor(M, Eeq, nat);
    const double CS = -P * (integral_CS(alpha, Mc2, mtc2, E0, Wmax_lim) - integral_CS(alpha, Mc2, mtc2, E0, Wmin_lim));
    const double IpMm = minimal_sufficient_E(Ip, Mc2, mtc2);
    a = CS / (Eeq - IpMm);
    b = -CS * Ip / (Eeq - IpMm);
}
// Integral_CDF_delta_CS :1449-1522.  Emax_in < 0: the optional argument is absent.  As in the reference the loop over the
// oscillators does not only accumulate: an oscillator whose branch is "cannot ionise" or "linear extrapolation" OVERWRITES what
// the previous ones have summed, and the prefactor of the LAST oscillator multiplies the result.
DLT_HD double Integral_CDF_delta_CS(double M, double mt, double E, const double *E0, const double *alpha, int Nosc, double Ip, double nat, bool identical, double Emax_in) {
    const double Mc2 = rest_energy(M), mtc2 = rest_energy(mt);
    double CS = 0.0, P = 0.0;
    for (int i = 0; i < Nosc; ++i) {
        const double Emin = W_min(Ip, Mc2, mtc2, E, E0[i]);
        const double Estart = (fabs(M - mt) / M < 1.0e-6) ? Ip : minimal_sufficient_E(Ip, Mc2, mtc2);
        const double Eeq = find_Wmax_equal_Wmin(Mc2, mtc2, identical, Ip, E0[i]);
        if (E <= Estart) { CS = 0.0; P = 0.0; }
        else {
            const double dEed = Eeq / 100.0;
            if (E <= Eeq + dEed) {
                double a, b;
                Find_linear_a_b(alpha[i], M, Mc2, mtc2, E0[i], Ip, identical, nat, Eeq + dEed, a, b);
                CS = a * E + b;
                P = 1.0;
            } else {
                double Emax = W_max(Mc2, mtc2, identical, E, Ip);
                if (Emax_in >= 0.0) {
                    if (Emax_in < Emin) Emax = Emin;
                    else if (Emax_in < Emax) Emax = Emax_in;
                }
                CS = CS - (integrated_delta_CDF_CS(alpha[i], Mc2, E0[i], mtc2, Emax, Ip, E) - integrated_delta_CDF_CS(alpha[i], Mc2, E0[i], mtc2, Emin, Ip, E));
                P = P_prefactor(M, E, nat);
            }
        }
    }
    return fabs(CS) * P;
}
// energy_loss_delta :1722-1786 (same overwrite / last-prefactor behaviour)
DLT_HD double energy_loss_delta(double E, double M, double Zeff, double Ip, double nat, double Mt, const double *E0, const double *alpha, int Nosc, bool identical) {
    const double Mc2 = rest_energy(M), mtc2 = rest_energy(Mt);
    double S_cur = 0.0, P = 0.0;
    for (int i = 0; i < Nosc; ++i) {
        const double Emin = W_min(Ip, Mc2, mtc2, E, E0[i]);
        const double Estart = (fabs(M - Mt) / M < 1.0e-6) ? Ip : minimal_sufficient_E(Ip, Mc2, mtc2);
        const double Eeq = find_Wmax_equal_Wmin(Mc2, mtc2, identical, Ip, E0[i]);
        if (E <= Estart) { S_cur = 0.0; P = 0.0; }
        else {
            const double dEed = Eeq / 100.0;
            const double Emax = W_max(Mc2, mtc2, identical, E, Ip);
            if (E <= Eeq + dEed) {
                double a, b;
                Find_linear_a_b(alpha[i], M, Mc2, mtc2, E0[i], Ip, identical, nat, Eeq + dEed, a, b);
                S_cur = 0.5 * a * (Emax - Estart);
                P = 1.0;
            } else {
                S_cur = S_cur - int_energy_loss(alpha[i], Mc2, mtc2, E0[i], Emax, Emin);
                P = P_prefactor(M, E, nat);
            }
        }
    }
    return fabs(S_cur) * (Zeff * Zeff) * P * nat * 1.0e-24;
}
// MFP_from_sigma :1417-1428
DLT_HD double MFP_from_sigma(double sigma, double nat) { return (sigma > 1.0e-24) ? 1.0 / (sigma * (nat * 1.0e-24)) : 1.0e30; }

// get_inelastic_energy_transfer :2051-2123 without its random number: RN in, sampled transferred energy out.  Bisection on the
// cross section integrated up to E_cur against RN x (the cross section integrated up to (Ip + Ee)/2), relative tolerance 1e-3
// or an interval below 1e-3 eV.
DLT_HD double inelastic_energy_transfer(double Ee, const double *E0, const double *alpha, int Nosc, double Ip, double nat, double RN) {
    const double eps = 1.0e-3;
    double E_left = Ip, E_right = (Ip + Ee) * 0.5;
    const double CS_tot = Integral_CDF_delta_CS(DLT_ME, DLT_ME, Ee, E0, alpha, Nosc, Ip, nat, true, E_right);
    const double CS_sampled = RN * CS_tot;
    double E_cur = (E_left + E_right) * 0.5;
    double CS_cur = Integral_CDF_delta_CS(DLT_ME, DLT_ME, Ee, E0, alpha, Nosc, Ip, nat, true, E_cur);
    while (fabs(CS_cur - CS_sampled) / CS_sampled > eps) {
        if (CS_cur > CS_sampled) E_right = E_cur; else E_left = E_cur;
        E_cur = (E_left + E_right) / 2.0;
        if (fabs(E_left - E_right) < eps) break;
        CS_cur = Integral_CDF_delta_CS(DLT_ME, DLT_ME, Ee, E0, alpha, Nosc, Ip, nat, true, E_cur);
    }
    return E_cur;
}

}  // namespace trk3delta
#endif
