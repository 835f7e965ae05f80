// dsf.cpp -- elastic scattering from tabulated dynamic-structure-factor cross sections (kind_of_EMFP = 2).
//   reading_DSF_cross_sections   Reading_files_and_parameters.f90:2516-2678   (file names :316-321)
//   elastic tables of the run    Analytical_IMFPs.f90:808-919 (electrons), the same for valence holes
// The reference ships no INPUT_DSF file; the format below is what its reader walks: the first line holds NTEPo, the number of
// transferred-energy points per particle energy; every further line holds "particle energy, transferred energy, differential
// inverse mean free path [1/(A eV)]", NTEPo lines per particle energy, transferred energies DESCENDING (the reader reverses them).
#include "trk3_host.hpp"

#include <cmath>
#include <fstream>
#include <sstream>

namespace trk3 {

// Linear_approx_2d(Array, In_val, Value1, El1, El2), Reading_files_and_parameters.f90:3272-3307 (as in the .dos reader, input.cpp)
static double linear_approx_2d(const std::vector<double> &X, const std::vector<double> &Y, double v, double El1, double El2) {
    const int N = (int)X.size();
    const int num = find_monoton_2d(X.data(), 1, N, v);
    if (num == 1) return El2 + (Y[0] - El2) / (X[0] - El1) * (v - El1);
    if (Y[(size_t)num - 2] > 1e20) return Y[(size_t)num - 2];
    return Y[(size_t)num - 2] + (Y[(size_t)num - 1] - Y[(size_t)num - 2]) / (X[(size_t)num - 1] - X[(size_t)num - 2]) * (v - X[(size_t)num - 2]);
}

std::string dsf_file_name(const Case &c, bool hole) {                                             // :312-321
    const int temper = (int)c.Matter.temp;
    return "INPUT_DSF/" + c.Material_name + "/" + c.Material_name + (hole ? "_Hole" : "_Electron") + "_DSF_Differential_EMFPs_" + std::to_string(temper) + "K.dat";
}

// returns false with err set on a malformed file; `found` = false if the file does not exist (the reference then falls back to
// Mott cross sections, :2544-2552)
bool read_dsf(const std::string &path, std::vector<DsfPoint> &out, bool &found, std::string &err) {
    out.clear();
    std::ifstream f(path);
    found = (bool)f;
    if (!found) return true;
    std::vector<std::string> lines;
    for (std::string l; std::getline(f, l);) lines.push_back(l);
    while (!lines.empty() && lines.back().find_first_not_of(" \t\r") == std::string::npos) lines.pop_back();
    const int N = (int)lines.size();                              // Count_lines_in_file
    const int M = N - 1;
    int NTEPo = 0;
    { std::istringstream is(N > 0 ? lines[0] : std::string()); if (!(is >> NTEPo) || NTEPo < 2) { err = path + ": first line must hold the number of transferred-energy points"; return false; } }
    // `do i = 2, M` reads lines 2..M of the N = M + 1 lines: the LAST data line of the file is never read (:2563-2569); its slot
    // keeps whatever the allocation held -- zero here
    std::vector<double> Temp_E((size_t)std::max(M, 0), 0.0), T1((size_t)std::max(M, 0), 0.0), T2((size_t)std::max(M, 0), 0.0);
    for (int i = 2; i <= M; ++i) {
        std::string l = lines[(size_t)i - 1];
        for (auto &ch : l) if (ch == 'd' || ch == 'D') ch = 'e';
        std::istringstream is(l);
        if (!(is >> Temp_E[(size_t)i - 2] >> T1[(size_t)i - 2] >> T2[(size_t)i - 2])) { err = path + ": problem reading line " + std::to_string(i); return false; }
        if (T2[(size_t)i - 2] < 1.0e-10) T2[(size_t)i - 2] = 0.0;
    }
    const int NEPo = M / NTEPo;
    if (NEPo < 2) { err = path + ": fewer than two particle energies"; return false; }
    out.assign((size_t)NEPo, DsfPoint{});
    std::vector<std::vector<double>> tdE((size_t)NEPo), tdL((size_t)NEPo);
    int k = 0;
    for (int i = 0; i < NEPo; ++i) {
        DsfPoint &d = out[(size_t)i];
        d.dE.assign((size_t)NTEPo, 0.0); d.dL.assign((size_t)NTEPo, 0.0); d.dL_absorb.assign((size_t)NTEPo, 0.0); d.dL_emit.assign((size_t)NTEPo, 0.0);
        tdE[(size_t)i].assign((size_t)NTEPo, 0.0); tdL[(size_t)i].assign((size_t)NTEPo, 0.0);
        d.E = Temp_E[(size_t)(NTEPo * i)];
        for (int j = 0; j < NTEPo; ++j, ++k) { tdE[(size_t)i][(size_t)j] = T1[(size_t)k]; tdL[(size_t)i][(size_t)j] = T2[(size_t)k]; }
    }
    for (int i = 0; i < NEPo; ++i) for (int j = 0; j < NTEPo; ++j) {        // reversed: ascending transferred energy (:2605-2610)
        out[(size_t)i].dE[(size_t)j] = tdE[(size_t)i][(size_t)(NTEPo - 1 - j)];
        out[(size_t)i].dL[(size_t)j] = tdL[(size_t)i][(size_t)(NTEPo - 1 - j)];
    }
    for (int i = 0; i < NEPo; ++i) {                                         // resampled to [-0.2, 0.2] eV and integrated (:2612-2659)
        DsfPoint &d = out[(size_t)i];
        const std::vector<double> X = d.dE, Y = d.dL;
        double Sum_MFP = 0.0, Sum_MFP_emit = 0.0;
        const double Emin = -0.2;
        double E = Emin;
        double Emax = X[(size_t)NTEPo - 1];
        if (Emax > 0.2) Emax = 0.2;
        const double dE = (Emax - Emin) / (double)NTEPo;
        for (int j = 0; j < NTEPo; ++j) {
            E = E + dE;
            const double loc = linear_approx_2d(X, Y, E, Emin - dE, 0.0);
            Sum_MFP = Sum_MFP + loc * dE;
            d.dE[(size_t)j] = E;
            d.dL[(size_t)j] = (std::fabs(Sum_MFP) > 1.0e-10) ? 1.0 / Sum_MFP : 1.0e30;
            if (d.dE[(size_t)j] >= 0.0) {                                    // emission; absorption does not change
                Sum_MFP_emit = Sum_MFP_emit + loc * dE;
                d.dL_emit[(size_t)j] = (std::fabs(Sum_MFP_emit) > 1.0e-10) ? 1.0 / Sum_MFP_emit : 1.0e30;
                d.dL_absorb[(size_t)j] = (j == 0) ? 1.0e30 : d.dL_absorb[(size_t)j - 1];
            } else {                                                         // absorption; emission does not change
                d.dL_absorb[(size_t)j] = (std::fabs(Sum_MFP) > 1.0e-10) ? 1.0 / Sum_MFP : 1.0e30;
                d.dL_emit[(size_t)j] = (j == 0) ? 1.0e30 : d.dL_emit[(size_t)j - 1];
            }
        }
    }
    return true;
}

// Elastic_MFP%Total / %Emit / %Absorb from the DSF rows (Analytical_IMFPs.f90:913-919): the integrals over ALL transferred energies
void dsf_elastic_tables(const std::vector<DsfPoint> &D, MFP &Total, std::vector<double> &Emit, std::vector<double> &Absorb) {
    const size_t n = D.size();
    Total.E.assign(n, 0.0); Total.L.assign(n, 0.0); Total.dEdx.assign(n, 0.0); Emit.assign(n, 0.0); Absorb.assign(n, 0.0);
    for (size_t i = 0; i < n; ++i) {
        Total.E[i] = D[i].E;
        Total.L[i] = D[i].dL.back(); Emit[i] = D[i].dL_emit.back(); Absorb[i] = D[i].dL_absorb.back();
    }
}

}  // namespace trk3
