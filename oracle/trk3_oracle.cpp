// trk3_oracle.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// CPU restatement of the reference's Monte-Carlo hot path (N-Medvedev/TREKIS-3 v3.3.0,
// Source_files/Monte_Carlo.f90 + the sampling routines of Cross_sections.f90), kept in the
// reference's own shape: array-of-structs particle arrays, ONE global time-ordered event loop
// driven by Find_min_time_particle, Calculated_statistics at every grid time.  It is the checker
// for the CUDA engine and the "restated reference" CPU baseline of bench.py; only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline/--impl reference legs may use it.
//
// PARITY UNPINNED: the reference ships no tests, golden vectors or example outputs, has no
// Fortran compiler in this image to be built with, and seeds its RNG from the clock
// (MPI_subroutines.f90:387-400).  The oracle is therefore pinned only by (i) line-by-line
// correspondence with the cited Fortran, (ii) the reference's own run-time invariants
// (energy conservation, Auger balance, monotone event time), see tests/.
//
// Every routine cites the Fortran it follows (file:line of /root/reference/Source_files).
// Indices are 1-based where that keeps the correspondence with the Fortran obvious.
#include "../include/trekis3_gpu.h"
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <thread>
#include <vector>

namespace {

// ---- Universal_Constants.f90:24-107
const double g_Pi = 3.1415926535897932384626433832795;
const double g_e = 1.602176487e-19;
const double g_me = 9.1093821545e-31;
const double g_cvel = 299792458.0;
const double g_Mp = 1836.1526724780 * g_me;
const double g_h = 1.05457162853e-34;
const double g_Ry = 13.6056981;
const double g_a0 = 0.5291772085936;
const double g_e0 = 8.854187817620e-12;

// ------------------------------------------------------------------------------------------
// Random numbers.  The reference calls the compiler's random_number (U[0,1), unseeded).
// mode 0: one sequential generator per iteration (reference-like);
// mode 1: counter-based Philox4x32-10 streams per particle, identical to the CUDA engine:
//         draw k of particle `id` in iteration `it` = philox(key = seed, ctr = (id_lo,id_hi,k,it)).
// Draws are mapped to (0,1] so that log(RN) and L/RN stay finite.
// ------------------------------------------------------------------------------------------
struct Stream { uint64_t id; uint32_t ctr; };

inline void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t out[4]) {
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

struct Rng {
    int mode; uint64_t seed; uint32_t it;
    uint64_t s[4];     // xoshiro256** state (mode 0)
    static uint64_t splitmix(uint64_t &x) { uint64_t z = (x += 0x9E3779B97F4A7C15ull); z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; return z ^ (z >> 31); }
    void init(int m, uint64_t sd, uint64_t iter) { mode = m; seed = sd; it = (uint32_t)iter; uint64_t x = sd ^ (0xA0761D6478BD642Full * (iter + 1)); for (auto &v : s) v = splitmix(x); }
    static uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
    uint64_t next64() { uint64_t r = rotl(s[1] * 5, 7) * 9, t = s[1] << 17; s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl(s[3], 45); return r; }
    double rn(Stream &st) {
        uint64_t bits;
        if (mode == 0) bits = next64();
        else {      // draw k of a stream = half (k & 1) of Philox block k >> 1 (one block yields two 64-bit draws)
            uint32_t o[4]; const uint32_t k = st.ctr++;
            philox4x32_10((uint32_t)st.id, (uint32_t)(st.id >> 32), k >> 1, it, (uint32_t)seed, (uint32_t)(seed >> 32), o);
            bits = (k & 1u) ? ((uint64_t)o[2] | ((uint64_t)o[3] << 32)) : ((uint64_t)o[0] | ((uint64_t)o[1] << 32));
        }
        return (double)((bits >> 11) + 1) * (1.0 / 9007199254740992.0);
    }
    // every collision starts at an even draw index, so that its draws pair up in Philox blocks (engine: event_begin)
    void align(Stream &st) { if (mode != 0) st.ctr = (st.ctr + 1u) & ~1u; }
    Stream child(Stream &parent, uint32_t tag) {      // id of a newly created particle
        Stream c; c.ctr = 0;
        if (mode == 0) { c.id = 0; return c; }
        uint32_t o[4];
        philox4x32_10((uint32_t)parent.id, (uint32_t)(parent.id >> 32), parent.ctr++, it, (uint32_t)seed, (uint32_t)(seed >> 32) ^ (0x80000000u | tag), o);
        c.id = (uint64_t)o[0] | ((uint64_t)o[1] << 32);
        return c;
    }
};

// ---- particle types, Objects.f90:31-68 (velocities/accelerations are never used by the MC)
struct Particle { double E, t0, tn, X, Y, Z, L, theta, phi; Stream rng; };
struct IonT : Particle { double Zeff, Mass; int Zat, Kind_Zeff; double fixed_Zeff; };
struct Electron : Particle {};
struct Hole : Particle { int KOA, Shl; double Mass, Ehkin; };    // KOA/Shl 1-based, 0 = empty slot
struct Photon : Particle {};

// ---- searches, Reading_files_and_parameters.f90:3348-3559 (1-based results)
int Find_in_monotonous_1D_array(const double *A, int N, double v) {
    int i_1 = 1, i_2 = N, i_cur = (int)std::floor((i_1 + i_2) / 2.0);
    double temp_val = A[i_cur - 1];
    if (v < A[0]) i_cur = 0;
    else if (v >= A[N - 1]) i_cur = N - 1;
    else for (;;) {
        if (i_1 == i_2 - 1) break;
        if (temp_val <= v) i_1 = i_cur; else i_2 = i_cur;
        i_cur = (int)std::floor((i_1 + i_2) / 2.0); temp_val = A[i_cur - 1];
    }
    return i_cur + 1;
}
int Find_in_monotonous_2D_array(const double *A, int N, double v) {   // A = row Indx of the (2,N) array
    int i_1 = 1, i_2 = N, i_cur = (int)std::floor((i_1 + i_2) / 2.0);
    double temp_val = A[i_cur - 1];
    if (v < A[0]) i_cur = 0;
    else if (v >= A[N - 1]) i_cur = N - 1;
    else { int coun = 0; for (;;) {
        if (v >= A[i_cur - 1] && v <= A[i_cur]) break;
        if (temp_val <= v) i_1 = i_cur; else i_2 = i_cur;
        i_cur = (int)std::floor((i_1 + i_2) / 2.0); temp_val = A[i_cur - 1];
        if (++coun > 1000) break; } }
    return i_cur + 1;
}
int Find_in_monoton_array_decreasing(const double *A, int N, double v) {
    int i_1 = 1, i_2 = N, i_cur = (int)std::floor((i_1 + i_2) / 2.0);
    double temp_val = A[i_cur - 1];
    if (v < A[N - 1]) i_cur = N;
    else if (v > A[0]) i_cur = 1;
    else { int coun = 0; while (std::abs(i_1 - i_2) > 1) {
        if (temp_val > v) i_1 = i_cur; else i_2 = i_cur;
        i_cur = (int)std::floor((i_1 + i_2) / 2.0); temp_val = A[i_cur - 1];
        if (++coun > 1000) break; } }
    return i_cur;
}
// Interpolate, Cross_sections.f90:4051-4086
double Interpolate(int Iflag, double E1, double E2, double S1, double S2, double En) {
    if (std::fabs(E2 - E1) < 1.0e-6) return std::max(S1, S2);
    if (En == E1) return S1;
    if (Iflag == 5) {
        double E2l = std::log(E2), E1l = std::log(E1), El = std::log(En), S1l = std::log(S1), S2l = std::log(S2);
        return std::exp(S1l + (S2l - S1l) / (E2l - E1l) * (El - E1l));
    }
    return S1 + (S2 - S1) / (E2 - E1) * (En - E1);
}

struct Tally {          // thin views on the packed buffer, Fortran element order
    double *b; const trk3_tally_layout *l;
    double &a1(int id, int i) { return b[l->off[id] + (i - 1)]; }
    double &a2(int id, int i, int j, int n1) { return b[l->off[id] + (i - 1) + (int64_t)n1 * (j - 1)]; }
    double &a3(int id, int i, int j, int k, int n1, int n2) { return b[l->off[id] + (i - 1) + (int64_t)n1 * ((j - 1) + (int64_t)n2 * (k - 1))]; }
    double &a4(int id, int i, int j, int k, int m, int n1, int n2, int n3) { return b[l->off[id] + (i - 1) + (int64_t)n1 * ((j - 1) + (int64_t)n2 * ((k - 1) + (int64_t)n3 * (m - 1)))]; }
};

struct IterOut { std::vector<double> totE, diffS; std::vector<int> totNel, diffN; };

// ------------------------------------------------------------------------------------------
struct MC {
    const trk3_config &cfg; const trk3_tables &T; const trk3_tally_layout &lay;
    Rng rng;
    uint64_t ev[TRK3_N_EVENT_CLASSES] = {0}, er[TRK3_N_ERRORS] = {0};
    uint64_t n_el = 0, n_ph = 0;
    int Nat, NS, Nt, Lowest_Ip_At, Lowest_Ip_Shl;
    double Egap_;
    double Out_theta1[TRK3_NTHETA];                 // Sorting_output_data.f90:1193-1196
    MC(const trk3_config &c, const trk3_tables &t, const trk3_tally_layout &l) : cfg(c), T(t), lay(l) {
        Nat = T.n_atoms; NS = T.n_shells; Nt = l.Nt;
        Lowest_Ip_At = T.shell_atom[T.vb_shell] + 1; Lowest_Ip_Shl = T.shell_num[T.vb_shell] + 1;
        Egap_ = T.shell_Ip[T.atom_first[0] + T.atom_nshl[0] - 1];     // Target_atoms(1)%Ip(size)
        for (int q = 0; q < TRK3_NTHETA; ++q) Out_theta1[q] = (double)(q + 1);
    }
    int flat(int KOA, int Shl) const { return T.atom_first[KOA - 1] + (Shl - 1); }
    double Ip(int KOA, int Shl) const { return T.shell_Ip[flat(KOA, Shl)]; }
    bool isVB(int KOA, int Shl) const { return KOA == Lowest_Ip_At && Shl == Lowest_Ip_Shl; }

    // per-iteration state (Monte_Carlo_modelling locals, Monte_Carlo.f90:485-521)
    std::vector<Electron> All_electrons; std::vector<Hole> All_holes; std::vector<Photon> All_photons;
    std::vector<double> Em_electrons;
    std::vector<double> SHI_path, El_IMFP, Hole_IMFP, Phot_IMFP;     // row 2 of the (2,N) totals; row 1 is the shared grid
    int Tot_Nel, Tot_Nphot, Em_Nel;
    double At_NRG, Em_gamma, Em_E1;
    IonT SHI_loc;

    // ---- Next_free_path_1d / _2d, Monte_Carlo.f90:1835-1899
    double Next_free_path_1d(double E, const double *Ea, const double *La, int N) {
        int n = Find_in_monotonous_1D_array(Ea, N, E);
        if (n == 1) {
            double MFP = Interpolate(1, Ea[0], Ea[1], La[0], La[1], E);
            if (MFP < Ea[0]) MFP = Ea[0];          // sic: compared with the ENERGY array (:1854)
            return MFP;
        }
        if (La[n - 2] >= 1.0e16) return La[n - 2];
        return Interpolate(5, Ea[n - 2], Ea[n - 1], La[n - 2], La[n - 1], E);
    }
    double Next_free_path_2d(double E, const double *Ea, const double *La, int N) {
        int n = Find_in_monotonous_2D_array(Ea, N, E);
        if (n == 1) {
            double MFP = Interpolate(1, Ea[0], Ea[1], La[0], La[1], E);
            if (MFP < La[0]) MFP = La[0];
            return MFP;
        }
        if (La[n - 2] >= 1.0e16) return La[n - 2];
        return Interpolate(5, Ea[n - 2], Ea[n - 1], La[n - 2], La[n - 1], E);
    }

    // ---- Which_shell, Monte_Carlo.f90:1786-1832
    void Which_shell(const double *Ea, const double *Lmat, int N, double E, Stream &st, int &Nat_o, int &Nshl_o) {
        double Temp[TRK3_MAX_SHELLS]; double MFP_tot = 0.0;
        for (int s = 0; s < NS; ++s) {
            const double *La = Lmat + (size_t)s * N;
            int n = Find_in_monotonous_1D_array(Ea, N, E);
            double MFP;
            if (n == 1) MFP = 1.0e20;
            else if (La[n - 2] == La[n - 1] || La[n - 2] > 1e20) MFP = La[n - 2];
            else MFP = Interpolate(5, Ea[n - 2], Ea[n - 1], La[n - 2], La[n - 1], E);
            Temp[s] = 1.0 / MFP; MFP_tot = MFP_tot + Temp[s];
        }
        double RN = rng.rn(st);
        MFP_tot = RN * MFP_tot;
        double MFP_sum = 0.0; int s_sel = NS - 1;
        for (int s = 0; s < NS; ++s) { MFP_sum = MFP_sum + Temp[s]; if (MFP_sum >= MFP_tot) { s_sel = s; break; } }
        Nat_o = T.shell_atom[s_sel] + 1; Nshl_o = T.shell_num[s_sel] + 1;
    }

    // ---- Get_velosity / Get_time_of_next_event, Monte_Carlo.f90:795-878
    double vel_ion(const IonT &p) { return std::sqrt(2.0 * p.E * g_e / (p.Mass * g_Mp)); }
    double vel_el(const Particle &p) { return std::sqrt(2.0 * p.E * g_e / g_me); }
    double vel_hole(const Hole &p) {
        if (p.Mass < 1.0e6) {
            if (p.Ehkin < -1.0e-6 || p.Mass < 1.0e-10) return 0.0;
            if (std::fabs(p.Ehkin) < 1.0e-6) return 0.0;
            return std::sqrt(2.0 * p.Ehkin * g_e / (p.Mass * g_me));
        }
        return 0.0;
    }
    static void tn_from(Particle &p, double V, double MFP) { if (V > 1.0e-10) p.tn = p.t0 + MFP / V * 1e5; else p.tn = 1.0e25; }
    void cut_off_e(Electron &e) { if (e.E < cfg.cut_off) e.tn = 1.0e20; }                                  // :3000-3014
    void cut_off_h(Hole &h) { if (h.Mass < 1e15 && h.Ehkin < cfg.cut_off) h.tn = 1.0e20; }

    // ---- Assign_holes_mass + Hole_parameters, Monte_Carlo.f90:724-792
    void Hole_parameters(Hole &h, double Eh, Stream &st) {
        if (isVB(h.KOA, h.Shl)) {
            double Etemp = h.Ehkin;
            h.Ehkin = Eh - Egap_;
            h.E = Egap_;
            if (cfg.hole_mass > 0) h.Mass = cfg.hole_mass;
            else { int m = Find_in_monotonous_1D_array(T.dos_E, T.n_dos, h.Ehkin); h.Mass = T.dos_effm[m - 1]; }
            if (h.Mass < 1.0e3) {
                double HIMFP = Next_free_path_2d(h.Ehkin, T.hi_E, Hole_IMFP.data(), T.n_hi), HEMFP;
                if (Etemp == (Eh - Egap_)) HEMFP = 1.0e30;
                else HEMFP = Next_free_path_2d(h.Ehkin, T.he_E, T.he_L, T.n_he);
                double RN = rng.rn(st);
                double MFP_tot = -std::log(RN) / (1.0 / HIMFP + 1.0 / HEMFP);
                tn_from(h, vel_hole(h), MFP_tot);
                h.L = MFP_tot;
            } else { h.L = 1.0e30; h.tn = 1.0e30; }
        } else {
            h.Mass = 1.0e29;
            double RN = rng.rn(st);
            int f = flat(h.KOA, h.Shl);
            double nu = 1.0 / T.shell_auger[f] + 1.0 / T.shell_radiat[f];
            h.tn = h.t0 - std::log(RN) / nu;
            h.L = 1.0e30; h.E = Eh; h.Ehkin = 0.0;
        }
    }

    // ---- check_hole_parameters, Monte_Carlo.f90:682-721
    void check_hole_parameters(double Eel, double &dE, double &Ehole, Electron *el) {
        Ehole = Eel - dE;
        int mhole = Find_in_monotonous_1D_array(T.dos_E, T.n_dos, Ehole);
        if (T.dos_DOS[mhole - 1] < 1.0e-4) {
            if (el) {
                while (mhole > 1 && T.dos_DOS[mhole - 1] < 1.0e-4) mhole = mhole - 1;
                double Eloc = T.dos_E[mhole - 1];
                el->E = el->E + (Ehole - Eloc);
                Ehole = Eloc;
            } else {
                while (mhole < T.n_dos && T.dos_DOS[mhole - 1] < 1.0e-4) mhole = mhole + 1;
                double Eloc = T.dos_E[mhole - 1];
                dE = dE + (Ehole - Eloc);
                Ehole = Eloc;
            }
        }
    }

    // ---- angles, Monte_Carlo.f90:1132-1360
    void Update_holes_angles_SHI(double &theta, double &phi, Stream &st) { double RN = rng.rn(st); theta = g_Pi * RN; double RN2 = rng.rn(st); phi = 2.0 * g_Pi * RN2; }
    void Update_holes_angles_el(const Hole &h, double Eel, double dE, double &theta, double &phi, double &theta1, double &phi1, Stream &st) {
        double E11 = Eel - dE, Mh = h.Mass * g_me;
        theta = std::acos(std::sqrt((Mh + g_me) * (Mh + g_me) / (4.0 * Mh * g_me) * dE / Eel));
        double RN2 = rng.rn(st); phi = 2.0 * g_Pi * RN2;
        if (std::isnan(theta)) { double RN = rng.rn(st); theta = g_Pi * RN; }
        theta1 = std::acos((Eel * (Mh - g_me) + E11 * (Mh + g_me)) / (2 * Mh * std::sqrt(Eel * E11)));
        phi1 = phi + g_Pi;
        if (std::isnan(theta1)) { double RN = rng.rn(st); theta1 = g_Pi * RN; }
    }
    void Update_electron_angles_SHI(const IonT &s, double dE, double &theta, double &phi, Stream &st) {
        double MSHI = s.Mass * g_Mp;
        if (s.E <= 0.0) theta = g_Pi / 2.0;
        else theta = std::acos(std::sqrt((MSHI + g_me) * (MSHI + g_me) / (4.0 * MSHI * g_me) * dE / s.E));
        double RN = rng.rn(st); phi = 2.0 * g_Pi * RN;
    }
    void Update_electron_angles_El(double E, double dE, double &theta, double &phi, Stream &st) {
        theta = std::acos((E - dE) / std::sqrt(E * (E - dE)));
        if (std::isnan(theta)) { double RN = rng.rn(st); theta = RN * g_Pi; }
        double RN = rng.rn(st); phi = 2.0 * g_Pi * RN;
    }
    static double rest_energy(double M0) { return M0 * g_cvel * g_cvel / g_e; }      // Cross_sections.f90:1557
    double cos_theta_from_W(double E, double W, double M_in, double mt, Stream &st) {   // Monte_Carlo.f90:1275-1299
        double Erest_in = rest_energy(M_in), Erest_t = rest_energy(mt);
        double E2mc = E + 2.0 * Erest_in, EmW = E - W;
        double W1 = E * E2mc - W * (E + Erest_in + Erest_t);
        double W2 = E * E2mc * EmW * (E2mc - W);
        double mu = (W2 > 0.0) ? W1 / std::sqrt(W2) : 0.0;
        if (std::fabs(mu) > 1.0) { double RN = rng.rn(st); mu = std::cos(g_Pi * RN); }
        return mu;
    }
    double Mtarget() const { double sm = 0, sp = 0; for (int a = 0; a < Nat; ++a) { sm += T.atom_mass[a] * T.atom_pers[a]; sp += T.atom_pers[a]; } return g_Mp * sm / sp; }
    void Update_particle_angles_lat(double E, double dE, double &theta, double &phi, double M_eff, Stream &st) {   // :1252-1272
        double arg = cos_theta_from_W(E, dE, M_eff * g_me, Mtarget(), st);
        theta = std::acos(arg);
        double RN2 = rng.rn(st); phi = 2.0 * g_Pi * RN2;
    }
    static void New_Angles_both(double phi0, double theta0, double theta, double psi, double &phi1, double &theta1) {   // :1328-1360
        phi1 = phi0 + theta * std::cos(theta0) * std::sin(psi);
        theta1 = theta0 + theta * std::cos(psi);
        while (theta1 < 0.0) { theta1 = std::fabs(theta1); phi1 = phi1 + g_Pi; }
        while (theta1 > g_Pi) { theta1 = 2.0 * g_Pi - theta1; phi1 = phi1 - g_Pi; }
        if (phi1 > 2.0 * g_Pi) phi1 = phi1 - std::floor(phi1 / (2.0 * g_Pi)) * 2.0 * g_Pi;
        if (phi1 < 0.0) phi1 = phi1 + std::ceil(std::fabs(phi1) / (2.0 * g_Pi)) * 2.0 * g_Pi;
    }

    // ---- From_where_in_VB, Monte_Carlo.f90:1522-1573
    double From_where_in_VB(bool haveE, double E, Stream &st) {
        int N = T.n_dos;
        if (!haveE) {
            double Sum_DOS = T.dos_int[N - 1];
            double RN = rng.rn(st); double Tot_N = RN * Sum_DOS;
            int n = Find_in_monotonous_1D_array(T.dos_int, N, Tot_N);
            if (n > 1) return T.dos_E[n - 2] + (T.dos_E[n - 1] - T.dos_E[n - 2]) * (Tot_N - T.dos_int[n - 2]) / (T.dos_int[n - 1] - T.dos_int[n - 2]);
            return T.dos_E[n - 1];
        }
        int M_temp = (E < T.dos_E[N - 1]) ? Find_in_monotonous_1D_array(T.dos_E, N, E) : N + 1;
        if (M_temp > 1) {
            double Sum_DOS = T.dos_int[M_temp - 2];
            double RN = rng.rn(st); double Tot_N = RN * Sum_DOS;
            int n = Find_in_monotonous_1D_array(T.dos_int, N, Tot_N);
            if (n > 1) return T.dos_E[n - 2] + (T.dos_E[n - 1] - T.dos_E[n - 2]) * (Tot_N - T.dos_int[n - 2]) / (T.dos_int[n - 1] - T.dos_int[n - 2]);
            return T.dos_E[n - 1];
        }
        return 0.0;
    }
    // ---- Electron_recieves_E, Monte_Carlo.f90:1653-1716
    double Electron_recieves_E(double dE, int Nat_cur, int Nshl_cur, Stream &st) {
        double E = dE - Ip(Nat_cur, Nshl_cur), dE_cur;
        if (isVB(Nat_cur, Nshl_cur)) {
            if (dE <= Ip(Nat_cur, Nshl_cur)) dE_cur = E;
            else {
                int N = T.n_dos;
                int M_temp = (E < T.dos_E[N - 1]) ? Find_in_monotonous_1D_array(T.dos_E, N, E) : N + 1;
                if (M_temp > 1) {
                    double Sum_DOS = T.dos_int[M_temp - 2];
                    double RN = rng.rn(st); double Tot_N = RN * Sum_DOS;
                    int n = Find_in_monotonous_1D_array(T.dos_int, N, Tot_N);
                    double E_DOS;
                    if (n > 1) E_DOS = T.dos_E[n - 2] + (T.dos_E[n - 1] - T.dos_E[n - 2]) * (Tot_N - T.dos_int[n - 2]) / (T.dos_int[n - 1] - T.dos_int[n - 2]);
                    else E_DOS = T.dos_E[n - 1];
                    dE_cur = E - E_DOS;
                } else dE_cur = E;
            }
        } else dE_cur = E;
        if (dE_cur < 0.0) er[TRK3_ERR_10]++;
        return dE_cur;
    }
    // ---- count_for_Auger_shells / Choose_for_Auger_shell, Monte_Carlo.f90:1576-1647
    double count_for_Auger_shells(double NRG, bool second_e) {
        double coun = 0.0;
        for (int s = 0; s < NS; ++s) {
            double E_delta = second_e ? 1.0e10 : NRG - T.shell_Ip[s];
            if (NRG > T.shell_Ip[s] + 1.0e-3 && E_delta >= Egap_) coun = coun + T.shell_Nel[s];
        }
        return coun;
    }
    void Choose_for_Auger_shell(double NRG, double Shel, int &Sh1, int &KOA1, bool second_e) {
        double coun_sh = 0.0;
        for (int s = 0; s < NS; ++s) {
            double E_delta = second_e ? 1.0e10 : NRG - T.shell_Ip[s];
            if (NRG > T.shell_Ip[s] + 1.0e-3 && E_delta >= Egap_) {
                coun_sh = coun_sh + T.shell_Nel[s];
                if (coun_sh >= Shel) { Sh1 = T.shell_num[s] + 1; KOA1 = T.shell_atom[s] + 1; }
            }
            if (coun_sh >= Shel) break;
        }
    }
    // ---- Auger_decay, Monte_Carlo.f90:1366-1443
    void Auger_decay(int KOA, int SHL, int &Sh1, int &KOA1, int &Sh2, int &KOA2, double &Ee, double &E_new1, double &E_new2, Stream &st) {
        Sh1 = SHL; Sh2 = 0; KOA1 = KOA; KOA2 = 0; Ee = -1.0e-10; E_new2 = 0.0;
        double coun = count_for_Auger_shells(Ip(KOA, SHL), false);
        double RN = rng.rn(st); double Shel = RN * coun;
        Choose_for_Auger_shell(Ip(KOA, SHL), Shel, Sh1, KOA1, false);
        double dE_cur = 0.0;
        if (isVB(KOA1, Sh1)) dE_cur = From_where_in_VB(false, 0.0, st);
        E_new1 = dE_cur + Ip(KOA1, Sh1);
        double Energy_diff = Ip(KOA, SHL) - E_new1;
        coun = count_for_Auger_shells(Energy_diff, true);
        if (coun > 0.0) {
            RN = rng.rn(st); Shel = RN * coun;
            Choose_for_Auger_shell(Energy_diff, Shel, Sh2, KOA2, true);
            dE_cur = 0.0;
            if (isVB(KOA2, Sh2)) dE_cur = From_where_in_VB(true, Energy_diff - Ip(KOA2, Sh2), st);
            E_new2 = dE_cur + Ip(KOA2, Sh2);
            Ee = Energy_diff - E_new2;
        }
        if (Ee < 0.0 || E_new1 < 0.0 || E_new2 < 0.0) er[TRK3_ERR_25]++;
    }
    // ---- Radiative_decay, Monte_Carlo.f90:2969-2997
    void Radiative_decay(int KOA, int SHL, int &Sh1, int &KOA1, double &dE, double &E_new1, Stream &st) {
        Sh1 = SHL; KOA1 = KOA;
        double coun = count_for_Auger_shells(Ip(KOA, SHL), true);
        double RN = rng.rn(st); double Shel = RN * coun;
        Choose_for_Auger_shell(Ip(KOA, SHL), Shel, Sh1, KOA1, true);
        double dE_cur = 0.0;
        if (isVB(KOA1, Sh1)) dE_cur = From_where_in_VB(false, 0.0, st);
        E_new1 = dE_cur + Ip(KOA1, Sh1);
        dE = Ip(KOA, SHL) - E_new1;
    }

    // ---- interpolate_transferred_energy, Cross_sections.f90:1968-2045
    double interpolate_transferred_energy(double Ele, const double *Eg, int NE, const int64_t *off, const double *hwA, const double *LA, double L_need) {
        int i_E = Find_in_monotonous_1D_array(Eg, NE, Ele);
        if (i_E > 1) { if (std::fabs(Eg[i_E - 2] - Ele) < 1.0e-6) i_E = i_E - 1; }
        auto row = [&](int iE, double &hw_o) {
            const double *L = LA + off[iE - 1]; const double *hw = hwA + off[iE - 1]; int n = (int)(off[iE] - off[iE - 1]);
            int i_hw = Find_in_monoton_array_decreasing(L, n, L_need);
            if (i_hw == 1 || i_hw == n) hw_o = hw[i_hw - 1];
            else hw_o = Interpolate(5, L[i_hw - 1], L[i_hw], hw[i_hw - 1], hw[i_hw], L_need);
        };
        double hw_1, hw_2, hw_out;
        row(i_E, hw_1);
        hw_out = hw_1;
        if (i_E > 1) {
            i_E = i_E - 1;
            row(i_E, hw_2);
            if (hw_1 < 1.0e-10 || hw_2 < 1.0e-10) hw_out = Interpolate(1, Eg[i_E - 1], Eg[i_E], hw_1, hw_2, Ele);
            else hw_out = Interpolate(5, Eg[i_E - 1], Eg[i_E], hw_1, hw_2, Ele);
        }
        return hw_out;
    }
    // ---- BEB shells (KOCS = 2).  dSigma_dw_int :3981-3984, dSigma_int_BEB :3958-3979, Electron_NRG_transfer_BEB :2128-2165
    static double dSigma_dw_int(double S, double t0, double u0, double w0) {
        return S / (t0 + u0 + 1.0) * (-(std::log(w0 + 1.0) - std::log(std::fabs(t0 - w0))) / (t0 + 1.0) + (1.0 / (t0 - w0) - 1.0 / (w0 + 1.0))
                                      + std::log(t0) * 0.5 * (1.0 / ((t0 - w0) * (t0 - w0)) - 1.0 / ((w0 + 1.0) * (w0 + 1.0))));
    }
    static double dSigma_int_BEB(double Tk, double w, double B, double U, double N) {
        const double S = 4.0 * g_Pi * g_a0 * g_a0 * N * (g_Ry / B) * (g_Ry / B);
        const double t0 = Tk / B, u0 = U / B, w0 = w / B;
        const double dSigma0 = dSigma_dw_int(S, t0, u0, 0.0);
        return dSigma_dw_int(S, t0, u0, w0) - dSigma0;
    }
    double Electron_NRG_transfer_BEB(double Ele, int Nat_cur, int Nshl_cur, double L_need, double Mass, double Emin) {
        const int f = flat(Nat_cur, Nshl_cur);
        const double B = T.shell_Ip[f], U = T.shell_Ek[f], N = T.shell_Nel[f];
        double sum_pers = 0.0;
        for (int a = 0; a < T.n_atoms; ++a) sum_pers += T.atom_pers[a];
        double Emin1 = Emin, Emax1 = (Ele - B) / 2.0, E = 0.0;
        int coun = 0;
        double Sigma_cur = dSigma_int_BEB(Ele, E, B, U, N);
        double temp1 = T.at_dens * 1e-24 * T.atom_pers[Nat_cur - 1] / sum_pers;
        double L_cur = 1.0 / (Mass * temp1 * Sigma_cur);
        while (std::fabs(L_cur - L_need) / L_need > 0.001) {
            coun = coun + 1;
            Sigma_cur = dSigma_int_BEB(Ele, E, B, U, N);
            L_cur = 1.0 / (Mass * temp1 * Sigma_cur);
            if (L_cur > L_need) Emin1 = E; else Emax1 = E;
            E = (Emax1 + Emin1) / 2.0;
            if (coun >= 1000) break;
        }
        return E + B;
    }
    // ---- Delta-function CDF (kind_of_DR = 4), Cross_sections.f90:1449-1720 and get_inelastic_energy_transfer :2051-2123.
    // The reference calls these with M = mt = g_me and identical = .true. for electrons and holes alike; the ion-electron
    // branches of W_min / W_max / minimal_sufficient_E are therefore not restated (they cannot be reached on this path).
    static double dl_rest_energy(double M0) { return M0 * g_cvel * g_cvel / g_e; }
    static double dl_W_min(double Ip_, double E, double E0) {                                  // :1564-1590, |Mc2 - mtc2|/Mc2 < 1e-6
        double Wmin = Ip_;
        const double E0min = E0 * (1.0 - 0.25 * E0 / E);
        return std::max(Wmin, E0min);
    }
    static double dl_W_max(double E, double Ip_) { return (E + Ip_) * 0.50; }                   // :1593-1616, identical
    static double dl_Eeq(double Ip_, double E0) {                                             // find_Wmax_equal_Wmin :1619-1649, identical
        return 5.0 / 4.0 * E0 - Ip_ / 2.0 + 0.25 * std::sqrt(17.0 * E0 * E0 - 12.0 * Ip_ * E0 + 4.0 * Ip_ * Ip_);
    }
    static double dl_P_prefactor(double M, double E, double nat) {                            // :1525-1556
        const double Erest = dl_rest_energy(M), fact = E / Erest + 1.0;
        const double v = g_cvel * std::sqrt(1.0 - 1.0 / (fact * fact));
        const double beta = v / g_cvel;
        const double g_me_eV = 0.51099906 * 1.0e6;
        return 1.0e24 / (g_Pi * g_a0 * nat * g_me_eV * (beta * beta));
    }
    static double dl_integral_CS(double alpha, double Mc2, double mtc2, double E0, double W) { // :1706-1712
        const double g_me_eV = 0.51099906 * 1.0e6;
        const double Mc22 = 2.0 * Mc2;
        return alpha / (g_me_eV * (Mc22 - E0)) * ((Mc22 - mtc2) * std::log(Mc22 + W - E0) + Mc22 * mtc2 / E0 * (std::log(W) - std::log(std::fabs(W - E0))));
    }
    static double dl_integrated_delta_CDF_CS(double alpha, double Mc2, double E0, double mtc2, double W, double Ip_, double E) {   // :1690-1704
        const double Wmin = dl_W_min(Ip_, E, E0);
        if (W < Wmin || E <= Ip_) return 0.0;
        return dl_integral_CS(alpha, Mc2, mtc2, E0, W);
    }
    double Integral_CDF_delta_CS(double E, int f, double Ip_, bool have_Emax_in, double Emax_in) const {   // :1449-1522
        const double M = g_me, Mc2 = dl_rest_energy(g_me), mtc2 = dl_rest_energy(g_me), nat = T.at_dens;
        double CS = 0.0, P = 0.0;
        for (int i = T.osc_off[f]; i < T.osc_off[f + 1]; ++i) {
            const double E0 = T.osc_E0[i], alpha = T.osc_alpha[i];
            const double Emin = dl_W_min(Ip_, E, E0);
            const double Estart = Ip_;
            const double Eeq = dl_Eeq(Ip_, E0);
            if (E <= Estart) { CS = 0.0; P = 0.0; }
            else {
                const double dEed = Eeq / 100.0;
                if (E <= Eeq + dEed) {                       // linear extrapolation, Find_linear_a_b :1671-1688
                    const double Ex = Eeq + dEed;
                    const double Wmin_lim = dl_W_min(Ip_, Ex, E0), Wmax_lim = dl_W_max(Ex, Ip_);
                    const double Pl = dl_P_prefactor(M, Ex, nat);
                    const double CSl = -Pl * (dl_integral_CS(alpha, Mc2, mtc2, E0, Wmax_lim) - dl_integral_CS(alpha, Mc2, mtc2, E0, Wmin_lim));
                    const double IpMm = Ip_;
                    const double a = CSl / (Ex - IpMm), b = -CSl * Ip_ / (Ex - IpMm);
                    CS = a * E + b;
                    P = 1.0;
                } else {
                    double Emax = dl_W_max(E, Ip_);
                    if (have_Emax_in) {
                        if (Emax_in < Emin) Emax = Emin;
                        else if (Emax_in < Emax) Emax = Emax_in;
                    }
                    CS = CS - (dl_integrated_delta_CDF_CS(alpha, Mc2, E0, mtc2, Emax, Ip_, E) - dl_integrated_delta_CDF_CS(alpha, Mc2, E0, mtc2, Emin, Ip_, E));
                    P = dl_P_prefactor(M, E, nat);
                }
            }
        }
        return std::fabs(CS) * P;
    }
    double get_inelastic_energy_transfer(double Ee, int Nat_cur, int Nshl_cur, double Ip_, Stream &st) {   // :2051-2123
        const int f = flat(Nat_cur, Nshl_cur);
        const double eps = 1.0e-3;
        double E_left = Ip_, E_right = (Ip_ + Ee) * 0.5;
        double RN = rng.rn(st);
        const double CS_tot = Integral_CDF_delta_CS(Ee, f, Ip_, true, E_right);
        const double CS_sampled = RN * CS_tot;
        double E_cur = (E_left + E_right) * 0.5;
        double CS_cur = Integral_CDF_delta_CS(Ee, f, Ip_, true, E_cur);
        while (std::fabs(CS_cur - CS_sampled) / CS_sampled > eps) {
            if (CS_cur > CS_sampled) E_right = E_cur; else E_left = E_cur;
            E_cur = (E_left + E_right) / 2.0;
            if (std::fabs(E_left - E_right) < eps) break;
            CS_cur = Integral_CDF_delta_CS(Ee, f, Ip_, true, E_cur);
        }
        return E_cur;
    }
    // ---- Electron_energy_transfer_inelastic (CS_method=1), Cross_sections.f90:1793-1871
    double Electron_energy_transfer_inelastic(double Ele, int Nat_cur, int Nshl_cur, double L_tot, bool hole, Stream &st) {
        double RN = rng.rn(st);
        double L_need = L_tot / RN;
        double Emin = Ip(Nat_cur, Nshl_cur);
        if (Emin <= 1.0e-3) Emin = 1.0e-3;
        double Emax, E;
        if (!hole) {
            Emax = (Ele + Emin) / 2.0;
            int f = flat(Nat_cur, Nshl_cur);
            if (T.shell_kocs[f] == 2) E = Electron_NRG_transfer_BEB(Ele, Nat_cur, Nshl_cur, L_need, 1.0, Emin);
            else if (T.delta_cdf) E = get_inelastic_energy_transfer(Ele, Nat_cur, Nshl_cur, Emin, st);      // :1894-1895
            else E = interpolate_transferred_energy(Ele, T.ei_E, T.n_ei, T.eid_off + (size_t)f * T.n_ei, T.eid_hw, T.eid_L, L_need);
        } else {
            double Mass;
            if (cfg.hole_mass >= 0) Mass = cfg.hole_mass;
            else { int m = Find_in_monotonous_1D_array(T.dos_E, T.n_dos, Ele); Mass = T.dos_effm[m - 1]; }
            Emax = 4.0 * Ele * Mass / ((Mass + 1.0) * (Mass + 1.0));
            if (T.shell_kocs[flat(Nat_cur, Nshl_cur)] == 2) E = Electron_NRG_transfer_BEB(Ele, Nat_cur, Nshl_cur, L_need, Mass, Emin);
            else if (T.delta_cdf) E = get_inelastic_energy_transfer(Ele, Nat_cur, Nshl_cur, Emin, st);
            else E = interpolate_transferred_energy(Ele, T.hi_E, T.n_hi, T.hid_off, T.hid_hw, T.hid_L, L_need);
        }
        if (E < Emin) E = Emin;
        if (E > Emax) E = Emax;
        if (std::isnan(E)) E = Emin;
        return E;
    }
    // ---- Electron_energy_transfer_elastic (CS_method=1), Cross_sections.f90:2285-2428
    double Electron_energy_transfer_elastic(double Ele, double L_tot, bool hole, Stream &st) {
        double RN = rng.rn(st);
        double L_need = L_tot / RN;
        double hw = hole ? interpolate_transferred_energy(Ele, T.he_E, T.n_he, T.hed_off, T.hed_hw, T.hed_L, L_need)
                         : interpolate_transferred_energy(Ele, T.ee_E, T.n_ee, T.eed_off, T.eed_hw, T.eed_L, L_need);
        if (hw >= Ele) hw = Ele;
        return hw;
    }
    // ---- NRG_transfer_elastic_atomic (Mott), Cross_sections.f90:3517-3613
    double NRG_transfer_elastic_atomic(double Mat, double Zat, double Ee, double M_eff, Stream &st) {
        double RN = rng.rn(st);
        double theta;
        {   // Mott_sample_mu (called without `mass`: electron rest mass)
            double me = g_me;
            double Erest = rest_energy(me);
            double fact = Ee / Erest + 1.0;
            double v = g_cvel * std::sqrt(1.0 - 1.0 / (fact * fact));
            if (v < 1.0e-6) theta = 0.0;
            else {
                double beta = v / g_cvel, beta2 = beta * beta;
                double tau = Ee / Erest;
                double alpha = g_e * g_e / (g_h * g_cvel * 4.0 * g_Pi * g_e0);
                double nu = 1.7e-5 * std::pow(Zat, 2.0 / 3.0) * (1.0 - beta2) / beta2 * (1.13 + 3.76 * alpha * alpha / beta2 * Zat * Zat * std::sqrt(tau / (1.0 + tau)));
                double mu = (RN * (2.0 * nu + 1.0) - nu) / (RN + nu);
                theta = std::acos(mu);
            }
        }
        // transfered_E_from_theta
        double mc2 = rest_energy(g_me * M_eff), Mct2 = rest_energy(Mat);
        double ct = std::cos(theta), ct2 = ct * ct, st2 = 1.0 - ct2;
        double Emc = Ee + mc2, E2mc = Ee + 2.0 * mc2, EmcMc = Emc + Mct2;
        double W1 = Emc * st2 + Mct2 - ct * std::sqrt(Mct2 * Mct2 - mc2 * mc2 * st2);
        double W2 = Ee * E2mc / (EmcMc * EmcMc - Ee * E2mc * ct2);
        return W1 * W2;
    }
    // ---- NRG_transfer_elastic_DSF, Cross_sections.f90:3652-3780, with Linear_approx_2x1d_DSF (Reading_files_and_parameters.f90:3718-3745)
    // and linear_interpolation (:3798-3802).  The row of integrated mean free paths is interpolated as a whole between the two
    // tabulated particle energies, as the reference does.
    double NRG_transfer_elastic_DSF(double Eel, bool hole, Stream &st) {
        const double *gE = hole ? T.he_E : T.ee_E, *gL = hole ? T.he_L : T.ee_L, *gEm = hole ? T.he_emit : T.ee_emit, *gAb = hole ? T.he_absorb : T.ee_absorb;
        const int NE = hole ? T.n_he : T.n_ee, NW = hole ? T.n_dsf_h : T.n_dsf_e;
        const double *dEt = hole ? T.dsf_h_dE : T.dsf_e_dE, *dEm = hole ? T.dsf_h_emit : T.dsf_e_emit, *dAb = hole ? T.dsf_h_absorb : T.dsf_e_absorb;
        int NumE = Find_in_monotonous_1D_array(gE, NE, Eel);
        if (NumE > 1) NumE = NumE - 1;
        std::vector<double> dL(NW, 0.0), dW(NW, 0.0);
        int i_MFP = Find_in_monotonous_1D_array(gE, NE, Eel);
        if (i_MFP > 1) i_MFP = i_MFP - 1;
        auto lin = [](double y1, double y2, double x1, double x2, double x) { return y1 + (y2 - y1) / (x2 - x1) * (x - x1); };
        const double EMFP_tot = lin(gL[i_MFP - 1], gL[i_MFP], gE[i_MFP - 1], gE[i_MFP], Eel);
        const double EMFP_emit = lin(gEm[i_MFP - 1], gEm[i_MFP], gE[i_MFP - 1], gE[i_MFP], Eel);
        const double EMFP_absorb = lin(gAb[i_MFP - 1], gAb[i_MFP], gE[i_MFP - 1], gE[i_MFP], Eel);
        double RN = rng.rn(st);
        const bool it_is_emission = RN < EMFP_tot / EMFP_emit;
        const double *src = it_is_emission ? dEm : dAb;
        if (NumE == NE) {
            for (int j = 0; j < NW; ++j) { dL[j] = src[(size_t)(NumE - 1) * NW + j]; dW[j] = dEt[(size_t)(NumE - 1) * NW + j]; }
        } else {
            const double Value1 = (Eel - gE[NumE - 1]) / (gE[NumE] - gE[NumE - 1]);
            for (int j = 0; j < NW; ++j) {
                dL[j] = src[(size_t)(NumE - 1) * NW + j] + (src[(size_t)NumE * NW + j] - src[(size_t)(NumE - 1) * NW + j]) * Value1;
                dW[j] = dEt[(size_t)(NumE - 1) * NW + j] + (dEt[(size_t)NumE * NW + j] - dEt[(size_t)(NumE - 1) * NW + j]) * Value1;
            }
            for (int j = 0; j < NW; ++j) if (dL[j] < 0.0) dL[j] = 0.0;
        }
        RN = rng.rn(st);
        const double L_need = it_is_emission ? EMFP_emit / RN : EMFP_absorb / RN;
        std::vector<double> neg(NW);
        for (int j = 0; j < NW; ++j) neg[j] = -dL[j];
        const int Number = Find_in_monotonous_1D_array(neg.data(), NW, -L_need);
        double dE;
        if (Number == 1) dE = dW[0] + (dW[1] - dW[0]) / (dL[1] - dL[0]) * (L_need - dL[0]);
        else if (std::fabs(dL[Number - 1] - dL[Number - 2]) < 1.0e-9) dE = dW[Number - 2];
        else if (dL[Number - 2] > 1e20) dE = dW[Number - 2];
        else dE = dW[Number - 2] + (dW[Number - 1] - dW[Number - 2]) / (dL[Number - 1] - dL[Number - 2]) * (L_need - dL[Number - 2]);
        if (std::fabs(dE) > 1.0 && it_is_emission && dE > Eel) dE = Eel;
        return dE;
    }
    double elastic_dE(double Eel, double EMFP, bool hole, double M_eff, Stream &st) {    // the kind_of_EMFP switch, Monte_Carlo.f90:2387-2407 / :2668-2692
        if (cfg.kind_of_EMFP == 2) return NRG_transfer_elastic_DSF(Eel, hole, st);
        if (cfg.kind_of_EMFP == 1) return Electron_energy_transfer_elastic(Eel, EMFP, hole, st);
        double dE = 0.0, sp = 0.0;
        for (int ii = 0; ii < Nat; ++ii) {
            double dE_loc = NRG_transfer_elastic_atomic(T.atom_mass[ii] * g_Mp, (double)T.atom_Z[ii], Eel, hole ? M_eff : 1.0, st);
            dE = dE + dE_loc * T.atom_pers[ii]; sp += T.atom_pers[ii];
        }
        return dE / sp;
    }

    // ---- Equilibrium_charge_SHI, Cross_sections.f90:2641-2680
    void Equilibrium_charge_SHI(IonT &s) {
        double vp = (s.E > 0.0) ? std::sqrt(2.0 * s.E * g_e / (s.Mass * g_Mp)) : 0.0;
        double sz = 0, sp = 0; for (int a = 0; a < Nat; ++a) { sz += T.atom_Z[a] * T.atom_pers[a]; sp += T.atom_pers[a]; }
        double Zt = sz / sp, Zp = (double)s.Zat, g_v0 = std::sqrt(2.0 * g_Ry * g_e / g_me);
        switch (s.Kind_Zeff) {
        case 1: s.Zeff = Zp * (1.0 - std::exp(-(vp / g_v0 / std::pow(Zp, 0.66666666)))); break;
        case 2: { double c1 = 0.6, c2 = 0.45; s.Zeff = Zp * std::pow(1.0 + std::pow(vp / (std::pow(Zp, c2) * g_v0 * 4.0 / 3.0), -1.0 / c1), -c1); break; }
        case 3: {
            double c1 = 1.0 - 0.26 * std::exp(-Zt / 11.0 - (Zt - Zp) * (Zt - Zp) / 9.0);
            double vpvo = std::pow(Zp, -0.543) * vp / g_v0;
            double c2 = 1.0 + 0.03 * vpvo * std::log(Zt);
            double x = c1 * std::pow(vpvo / c2 / 1.54, 1.0 + 1.83 / Zp), x2 = x * x, x4 = x2 * x2;
            s.Zeff = Zp * (8.29 * x + x4) / (0.06 / x + 4.0 + 7.4 * x + x4); break; }
        case 4: s.Zeff = s.fixed_Zeff; break;
        default: s.Zeff = Zp * (1.0 - std::exp(-(vp * 125.0 / g_cvel / std::pow(Zp, 0.66666666)))); break;
        }
    }
    // ---- SHI_energy_transfer (CDF), Monte_Carlo.f90:1719-1780
    double SHI_energy_transfer(int Nat_cur, int Nshl_cur, Stream &st) {
        int f = flat(Nat_cur, Nshl_cur);
        const double *Ea = T.dshi_E + T.dshi_off[f]; const double *La = T.dshi_L + T.dshi_off[f];
        int N = (int)(T.dshi_off[f + 1] - T.dshi_off[f]);
        double RN = rng.rn(st);
        double E_cur = T.shell_Ip[f], dL;
        int M_temp = Find_in_monotonous_1D_array(Ea, N, E_cur);
        if (M_temp > 1) {
            if (La[M_temp - 2] > 1.0e-10) dL = Interpolate(5, Ea[M_temp - 2], Ea[M_temp - 1], La[M_temp - 2], La[M_temp - 1], E_cur);
            else dL = Interpolate(1, Ea[M_temp - 2], Ea[M_temp - 1], La[M_temp - 2], La[M_temp - 1], E_cur);
        } else dL = La[0];
        double Tot_N;
        if (dL > 0.0 && La[N - 1] > 0.0) Tot_N = 1.0 / dL + RN * (1.0 / La[N - 1] - 1.0 / dL);
        else Tot_N = 1.5e21;
        int N_temmp;
        if (Tot_N < 1e20) { int i = 1; while (i < N && 1.0 / La[i - 1] < Tot_N) i = i + 1; N_temmp = i; }   // Find_in_1D_array (linear)
        else N_temmp = M_temp;
        if (N_temmp > M_temp) return Interpolate(5, 1.0 / La[N_temmp - 2], 1.0 / La[N_temmp - 1], Ea[N_temmp - 2], Ea[N_temmp - 1], Tot_N);
        return T.shell_Ip[f];
    }
    double Impact_parameter(const IonT &s, double dE) {     // Monte_Carlo.f90:1113-1126
        double MSHI = g_Mp * s.Mass, A = 1.0 + MSHI / g_me;
        return g_a0 * s.Zeff * g_Ry / s.E * std::sqrt(4.0 * s.E / dE * MSHI / g_me - A * A);
    }

    // ---- Check_size / resize_array, Objects.f90:420-486
    void Check_size(int N) {
        int i = (int)All_electrons.size();
        if (N > i) {
            int M = i + 1000;
            Electron e{}; e.E = 0; e.t0 = All_electrons[0].t0; e.tn = 1e20; e.X = e.Y = e.Z = 0; e.L = 1e20; e.theta = e.phi = 0; e.rng = Stream{0, 0};
            Hole h{}; h.E = 0; h.Ehkin = 0; h.t0 = All_electrons[0].t0; h.tn = 1e21; h.X = h.Y = h.Z = 0; h.L = 1e30; h.KOA = 0; h.Shl = 0; h.Mass = 1e30; h.theta = h.phi = 0; h.rng = Stream{0, 0};
            All_electrons.resize(M, e); All_holes.resize(M, h); Em_electrons.resize(M, 0.0);
        }
    }
    void Check_size_ph(int N) {
        int i = (int)All_photons.size();
        if (N > i) { Photon p{}; p.E = 0; p.t0 = All_photons.empty() ? 0.0 : All_photons[0].t0; p.tn = 1e20; p.X = p.Y = p.Z = 0; p.L = 1e20; p.theta = p.phi = 0; p.rng = Stream{0, 0}; All_photons.resize(i + 1000, p); }
    }

    // ---- How_many_electrons, Monte_Carlo.f90:1902-2057
    void How_many_electrons() {
        int N = T.n_shi;
        SHI_path.assign(N, 0.0); std::vector<double> SHI_loss(N, 0.0);
        for (int s = 0; s < NS; ++s) for (int i = 0; i < N; ++i) { SHI_loss[i] += T.shi_dEdx[(size_t)s * N + i]; SHI_path[i] += 1.0 / T.shi_L[(size_t)s * N + i]; }
        for (int i = 0; i < N; ++i) SHI_path[i] = (SHI_path[i] < 1.0e-10) ? 1.0e30 : 1.0 / SHI_path[i];
        int n = Find_in_monotonous_2D_array(T.shi_E, N, cfg.shi_E);
        if (n < 2) n = 2;
        double SHI_dEdx = Interpolate(5, T.shi_E[n - 2], T.shi_E[n - 1], SHI_loss[n - 2], SHI_loss[n - 1], cfg.shi_E);
        double dEdx = SHI_dEdx * cfg.layer;
        double Nel_d = std::ceil(dEdx / T.shell_Ip[T.vb_shell]);
        int Nel = (Nel_d > 5000 * cfg.layer) ? (int)(5000 * cfg.layer) : (int)Nel_d;
        if (Nel < 1000) Nel = 1000;
        Electron e{}; e.E = 0; e.t0 = 0.0; e.tn = 1e20; e.X = e.Y = e.Z = 0; e.L = 0; e.theta = e.phi = 0; e.rng = Stream{0, 0};
        Hole h{}; h.E = 0; h.Ehkin = 0; h.t0 = 0.0; h.tn = 1e21; h.X = h.Y = h.Z = 0; h.L = 1e30; h.KOA = 0; h.Shl = 0; h.Mass = 1e30; h.theta = h.phi = 0; h.rng = Stream{0, 0};
        All_electrons.assign(Nel, e); All_holes.assign(Nel, h); Em_electrons.assign(Nel, 0.0);
        N = T.n_ei; El_IMFP.assign(N, 0.0);
        for (int s = 0; s < NS; ++s) for (int i = 0; i < N; ++i) { double L = T.ei_L[(size_t)s * N + i]; if (L > 1.0e-10) El_IMFP[i] += 1.0 / L; }
        for (int i = 0; i < N; ++i) El_IMFP[i] = (El_IMFP[i] < 1.0e-10) ? 1.0e30 : 1.0 / El_IMFP[i];
        N = T.n_hi; Hole_IMFP.assign(N, 0.0);
        for (int s = 0; s < NS; ++s) for (int i = 0; i < N; ++i) Hole_IMFP[i] += 1.0 / T.hi_L[(size_t)s * N + i];
        for (int i = 0; i < N; ++i) Hole_IMFP[i] = (Hole_IMFP[i] < 1.0e-10) ? 1.0e30 : 1.0 / Hole_IMFP[i];
        All_photons.clear(); Phot_IMFP.clear();
        if (cfg.include_photons) {
            Photon p{}; p.E = 0; p.t0 = 0.0; p.tn = 1e20; p.X = p.Y = p.Z = 0; p.L = 0; p.theta = p.phi = 0; p.rng = Stream{0, 0};
            All_photons.assign(Nel, p);
            N = T.n_ph; Phot_IMFP.assign(N, 0.0);
            for (int s = 0; s < NS; ++s) for (int i = 0; i < N; ++i) Phot_IMFP[i] += 1.0 / T.ph_L[(size_t)s * N + i];
            for (int i = 0; i < N; ++i) Phot_IMFP[i] = (Phot_IMFP[i] < 1.0e-10) ? 1.0e30 : 1.0 / Phot_IMFP[i];
        }
    }
    // ---- barrier_parameters, Monte_Carlo.f90:2094-2115
    void barrier_parameters() {
        double wf = cfg.work_function, bh = cfg.bar_height;
        double Em_L = cfg.bar_length * 1.0e-10;
        double Em_B = 2.0 * bh - wf + 2.0 * std::sqrt(bh * bh - bh * wf);
        double Em_ksi = 0.5 * std::sqrt(8.0 * g_me * Em_L * Em_L * Em_B * g_e / ((2.0 * g_Pi * g_h) * (2.0 * g_Pi * g_h)) - 1.0);
        double Em_bb = std::cosh(2.0 * g_Pi * Em_ksi);
        double Em_delta = 2.0 * g_Pi * Em_L * std::sqrt(2.0 * g_me * g_e) / (2.0 * g_Pi * g_h);
        Em_E1 = bh + 2.0 * std::sqrt(bh * (bh - wf)) * (std::acosh(Em_bb) / (Em_delta * (std::sqrt(bh) + std::sqrt(bh - wf))) - 1.0);
        double g1 = Em_delta * (std::sqrt(Em_E1) + std::sqrt(Em_E1 - wf)), g2 = Em_delta * (std::sqrt(Em_E1) - std::sqrt(Em_E1 - wf));
        Em_gamma = (g1 * std::sinh(g1) + 2.0 * g2 * std::sinh(g2)) / (std::sqrt(Em_E1 * (Em_E1 - wf)) * (Em_bb + std::cosh(g1)));
    }
    // ---- calculate_emission, Monte_Carlo.f90:2477-2513
    void calculate_emission(int NOP, Stream &st) {
        Electron &e = All_electrons[NOP - 1];
        if (e.Z < 0.0) {
            if (e.E >= 1.5 * cfg.bar_height) { Em_Nel++; e.tn = 1e30; e.L = 1e30; Em_electrons[Em_Nel - 1] = e.E - cfg.work_function; }
            else {
                double RN = rng.rn(st);
                double Em_Penetr = 1.0 / (1.0 + std::exp(Em_gamma * (Em_E1 - e.E)));
                double Ekin = e.E - cfg.work_function;
                if (Ekin > 0.0 && RN < Em_Penetr) { Em_Nel++; e.tn = 1.0e30; e.L = 1.0e30; Em_electrons[Em_Nel - 1] = Ekin; }
                else if (std::cos(e.theta) < 0) e.theta = g_Pi - e.theta;
            }
        }
    }

    // ---- SHI_Monte_Carlo, Monte_Carlo.f90:2153-2249
    void SHI_Monte_Carlo() {
        Stream &st = SHI_loc.rng;
        rng.align(st);
        ev[TRK3_EV_SHI]++;
        int Nat_cur, Nshl_cur;
        Which_shell(T.shi_E, T.shi_L, T.n_shi, SHI_loc.E, st, Nat_cur, Nshl_cur);
        double dE = SHI_energy_transfer(Nat_cur, Nshl_cur, st);
        double SHI_IMFP = Next_free_path_2d(SHI_loc.E, T.shi_E, SHI_path.data(), T.n_shi);
        double RN = rng.rn(st);
        SHI_IMFP = -SHI_IMFP * std::log(RN);
        double Z = SHI_loc.Z + SHI_loc.L;
        SHI_loc.E = SHI_loc.E - dE; SHI_loc.t0 = SHI_loc.tn; SHI_loc.Z = Z; SHI_loc.L = SHI_IMFP;
        tn_from(SHI_loc, vel_ion(SHI_loc), SHI_IMFP);
        Equilibrium_charge_SHI(SHI_loc);
        Tot_Nel = Tot_Nel + 1; n_el++;
        Check_size(Tot_Nel);
        Electron &el = All_electrons[Tot_Nel - 1]; Hole &ho = All_holes[Tot_Nel - 1];
        el.rng = rng.child(st, 1); ho.rng = rng.child(st, 2);
        // stream convention of the engine (physics.cuh, shi_emit): the draws that create the electron come from the
        // electron's own stream, those that create the hole from the hole's (the ion's chain does not depend on them)
        Stream &se = el.rng, &sh = ho.rng;
        double dE_cur = Electron_recieves_E(dE, Nat_cur, Nshl_cur, se);
        double theta, phi;
        Update_electron_angles_SHI(SHI_loc, dE, theta, phi, se);
        double IMFP = Next_free_path_2d(dE_cur, T.ei_E, El_IMFP.data(), T.n_ei);
        double EMFP = Next_free_path_2d(dE_cur, T.ee_E, T.ee_L, T.n_ee);
        RN = rng.rn(se);
        double MFP_tot = -std::log(RN) / (1.0 / IMFP + 1.0 / EMFP);
        double L = Impact_parameter(SHI_loc, dE);
        double X = SHI_loc.X + L * std::sin(phi), Y = SHI_loc.Y + L * std::cos(phi);
        el.E = dE_cur; el.t0 = SHI_loc.t0; el.X = X; el.Y = Y; el.Z = Z; el.L = MFP_tot; el.theta = theta; el.phi = phi;
        tn_from(el, vel_el(el), MFP_tot);
        cut_off_e(el);
        if (el.E < -1.0e-9 || std::isnan(el.E)) er[TRK3_ERR_20]++;
        double htheta, hphi;
        Update_holes_angles_SHI(htheta, hphi, sh);
        ho.t0 = SHI_loc.t0; ho.X = X; ho.Y = Y; ho.Z = Z; ho.KOA = Nat_cur; ho.Shl = Nshl_cur; ho.theta = htheta; ho.phi = hphi;
        Hole_parameters(ho, dE - dE_cur, sh);
        cut_off_h(ho);
        if (ho.Ehkin < -1.0e-9 || std::isnan(ho.Ehkin)) er[TRK3_ERR_20]++;
    }

    // ---- Electron_Monte_Carlo, Monte_Carlo.f90:2253-2474
    void Electron_Monte_Carlo(int NOP, int i, Tally &out) {
        Stream st = All_electrons[NOP - 1].rng;      // copy: All_electrons may be reallocated by Check_size
        rng.align(st);
        double Eel = All_electrons[NOP - 1].E;
        double IMFP = Next_free_path_2d(Eel, T.ei_E, El_IMFP.data(), T.n_ei);
        double EMFP = Next_free_path_2d(Eel, T.ee_E, T.ee_L, T.n_ee);
        double RN = rng.rn(st);
        double L = All_electrons[NOP - 1].L, theta0 = All_electrons[NOP - 1].theta, phi0 = All_electrons[NOP - 1].phi;
        double X = All_electrons[NOP - 1].X + L * std::sin(theta0) * std::sin(phi0);
        double Y = All_electrons[NOP - 1].Y + L * std::sin(theta0) * std::cos(phi0);
        double Z = All_electrons[NOP - 1].Z + L * std::cos(theta0);
        double dE, theta, phi;
        if (RN * (1.0 / IMFP + 1.0 / EMFP) < 1.0 / IMFP) {
            ev[TRK3_EV_EL_INEL]++;
            int Nat_cur, Nshl_cur;
            Which_shell(T.ei_E, T.ei_L, T.n_ei, Eel, st, Nat_cur, Nshl_cur);
            Tot_Nel = Tot_Nel + 1; n_el++;
            Check_size(Tot_Nel);
            Electron &en = All_electrons[Tot_Nel - 1]; Hole &hn = All_holes[Tot_Nel - 1];
            en.rng = rng.child(st, 1); hn.rng = rng.child(st, 2);
            int f = flat(Nat_cur, Nshl_cur);
            IMFP = Next_free_path_1d(Eel, T.ei_E, T.ei_L + (size_t)f * T.n_ei, T.n_ei);
            dE = Electron_energy_transfer_inelastic(Eel, Nat_cur, Nshl_cur, IMFP, false, st);
            Update_electron_angles_El(Eel, dE, theta, phi, st);
            // stream convention of the engine (physics.cuh, electron_ion_emit): the draws that create the pair come from
            // the new electron's / the new hole's own streams (the primary's history does not depend on them)
            Stream &se = en.rng, &sh = hn.rng;
            double dE_cur = Electron_recieves_E(dE, Nat_cur, Nshl_cur, se);
            IMFP = Next_free_path_2d(dE_cur, T.ei_E, El_IMFP.data(), T.n_ei);
            EMFP = Next_free_path_2d(dE_cur, T.ee_E, T.ee_L, T.n_ee);
            RN = rng.rn(se);
            double MFP_tot = -std::log(RN) / (1.0 / IMFP + 1.0 / EMFP);
            double theta2 = g_Pi / 2.0 - theta, phi2 = phi + g_Pi, phi1, theta1;
            New_Angles_both(phi0, theta0, theta2, phi2, phi1, theta1);
            double t_ev = All_electrons[NOP - 1].tn;
            en.E = dE_cur; en.t0 = t_ev; en.X = X; en.Y = Y; en.Z = Z; en.L = MFP_tot; en.theta = theta1; en.phi = phi1;
            tn_from(en, vel_el(en), MFP_tot);
            cut_off_e(en);
            if (en.E < -1.0e-9 || std::isnan(en.E)) er[TRK3_ERR_21]++;
            double htheta, hphi;
            Update_holes_angles_SHI(htheta, hphi, sh);
            hn.t0 = t_ev; hn.X = X; hn.Y = Y; hn.Z = Z; hn.KOA = Nat_cur; hn.Shl = Nshl_cur; hn.theta = htheta; hn.phi = hphi;
            Hole_parameters(hn, dE - dE_cur, sh);
            cut_off_h(hn);
            if (hn.Ehkin < -1.0e-9 || std::isnan(hn.Ehkin)) er[TRK3_ERR_20]++;
        } else {
            ev[TRK3_EV_EL_ELAST]++;
            EMFP = Next_free_path_1d(Eel, T.ee_E, T.ee_L, T.n_ee);
            dE = elastic_dE(Eel, EMFP, false, 1.0, st);
            Update_particle_angles_lat(Eel, dE, theta, phi, 1.0, st);
            At_NRG = At_NRG + dE;
            double R = std::sqrt(X * X + Y * Y);
            if (std::isnan(R) || std::isnan(theta) || std::isnan(phi)) er[TRK3_ERR_NAN]++;
            int j = Find_in_monotonous_1D_array(T.out_R, T.n_r, R);
            out.a2(TRK3_OUT_ELAT, i, j, Nt) += dE * T.out_V[j - 1];
        }
        Electron &e = All_electrons[NOP - 1];
        IMFP = Next_free_path_2d(Eel - dE, T.ei_E, El_IMFP.data(), T.n_ei);
        EMFP = Next_free_path_2d(Eel - dE, T.ee_E, T.ee_L, T.n_ee);
        RN = rng.rn(st);
        double MFP_tot = -std::log(RN) / (1.0 / IMFP + 1.0 / EMFP);
        double phi1, theta1;
        New_Angles_both(phi0, theta0, theta, phi, phi1, theta1);
        e.E = Eel - dE; e.t0 = e.tn; e.X = X; e.Y = Y; e.Z = Z; e.L = MFP_tot; e.theta = theta1; e.phi = phi1;
        tn_from(e, vel_el(e), MFP_tot);
        cut_off_e(e);
        if (cfg.work_function > 0) calculate_emission(NOP, st);
        if (e.E < -1.0e-9 || std::isnan(e.E)) er[TRK3_ERR_22]++;
        All_electrons[NOP - 1].rng = st;
    }

    // ---- Hole_Monte_Carlo, Monte_Carlo.f90:2517-2869
    void Hole_Monte_Carlo(int NOP, int i, double t_cur, Tally &out) {
        Stream st = All_holes[NOP - 1].rng;
        rng.align(st);
        double Egap = Egap_;
        if (isVB(All_holes[NOP - 1].KOA, All_holes[NOP - 1].Shl)) {
            double Eel = All_holes[NOP - 1].Ehkin;
            double HIMFP = Next_free_path_2d(Eel, T.hi_E, Hole_IMFP.data(), T.n_hi);
            double HEMFP = Next_free_path_2d(Eel, T.he_E, T.he_L, T.n_he);
            double RN = rng.rn(st);
            double L = All_holes[NOP - 1].L, theta0 = All_holes[NOP - 1].theta, phi0 = All_holes[NOP - 1].phi;
            double X = All_holes[NOP - 1].X + L * std::sin(theta0) * std::sin(phi0);
            double Y = All_holes[NOP - 1].Y + L * std::sin(theta0) * std::cos(phi0);
            double Z = All_holes[NOP - 1].Z + L * std::cos(theta0);
            double dE, Ehole, htheta1, hphi1;
            if (RN * (1.0 / HIMFP + 1.0 / HEMFP) < 1.0 / HIMFP && HIMFP < 1e15) {
                ev[TRK3_EV_VBH_INEL]++;
                int Nat_cur, Nshl_cur;
                Which_shell(T.hi_E, T.hi_L, T.n_hi, Eel, st, Nat_cur, Nshl_cur);
                Tot_Nel = Tot_Nel + 1; n_el++;
                Check_size(Tot_Nel);
                Electron &en = All_electrons[Tot_Nel - 1]; Hole &hn = All_holes[Tot_Nel - 1];
                en.rng = rng.child(st, 1); hn.rng = rng.child(st, 2);
                int f = flat(Nat_cur, Nshl_cur);
                HIMFP = Next_free_path_1d(Eel, T.hi_E, T.hi_L + (size_t)f * T.n_hi, T.n_hi);
                dE = Electron_energy_transfer_inelastic(Eel, Nat_cur, Nshl_cur, HIMFP, true, st);
                double htheta, hphi;
                Update_holes_angles_el(All_holes[NOP - 1], Eel, dE, htheta, hphi, htheta1, hphi1, st);
                double dE_cur = Electron_recieves_E(dE, Nat_cur, Nshl_cur, st);
                double IMFP = Next_free_path_2d(dE_cur, T.ei_E, El_IMFP.data(), T.n_ei);
                double EMFP = Next_free_path_2d(dE_cur, T.ee_E, T.ee_L, T.n_ee);
                RN = rng.rn(st);
                double MFP_tot = -std::log(RN) / (1.0 / IMFP + 1.0 / EMFP);
                double theta2 = htheta;
                RN = rng.rn(st);                       // sic: drawn and discarded (:2611)
                double phi2 = hphi, phi1, theta1;
                New_Angles_both(phi0, theta0, theta2, phi2, phi1, theta1);
                double t_ev = All_holes[NOP - 1].tn;
                en.E = dE_cur; en.t0 = t_ev; en.X = X; en.Y = Y; en.Z = Z; en.L = MFP_tot; en.theta = theta1; en.phi = phi1;
                tn_from(en, vel_el(en), MFP_tot);
                cut_off_e(en);
                if (en.E < -1.0e-9 || std::isnan(en.E)) er[TRK3_ERR_40]++;
                Update_holes_angles_SHI(htheta, hphi, st);
                hn.t0 = t_ev; hn.X = X; hn.Y = Y; hn.Z = Z; hn.KOA = Nat_cur; hn.Shl = Nshl_cur; hn.theta = htheta; hn.phi = hphi;
                Hole_parameters(hn, dE - dE_cur, st);
                cut_off_h(hn);
                if (hn.Ehkin < -1.0e-9 || std::isnan(hn.Ehkin)) er[TRK3_ERR_41]++;
                if ((hn.E + hn.Ehkin) < Egap - 1.0e-12) er[TRK3_ERR_41]++;
                check_hole_parameters(Eel, dE, Ehole, &en);
            } else {
                ev[TRK3_EV_VBH_ELAST]++;
                HEMFP = Next_free_path_1d(Eel, T.he_E, T.he_L, T.n_he);
                dE = elastic_dE(Eel, HEMFP, true, All_holes[NOP - 1].Mass, st);
                Update_particle_angles_lat(Eel, dE, htheta1, hphi1, All_holes[NOP - 1].Mass, st);
                check_hole_parameters(Eel, dE, Ehole, nullptr);
                At_NRG = At_NRG + dE;
                double R = std::sqrt(X * X + Y * Y);
                if (std::isnan(R)) er[TRK3_ERR_NAN]++;
                int j = Find_in_monotonous_1D_array(T.out_R, T.n_r, R);
                out.a2(TRK3_OUT_ELAT, i, j, Nt) += dE * T.out_V[j - 1];
            }
            Hole &h = All_holes[NOP - 1];
            double hphi2, htheta2;
            New_Angles_both(phi0, theta0, htheta1, hphi1, hphi2, htheta2);
            h.t0 = h.tn; h.X = X; h.Y = Y; h.Z = Z; h.theta = htheta2; h.phi = hphi2;
            Hole_parameters(h, Ehole + Egap, st);
            cut_off_h(h);
            if (h.Ehkin < -1.0e-9 || std::isnan(h.Ehkin)) er[TRK3_ERR_20]++;
        } else {
            double RN = rng.rn(st);
            int f = flat(All_holes[NOP - 1].KOA, All_holes[NOP - 1].Shl);
            double t_Auger = T.shell_auger[f], t_Radiat = T.shell_radiat[f];
            if (RN * (1.0 / t_Auger + 1.0 / t_Radiat) < 1.0 / t_Auger) {
                int Sh1, KOA1, Sh2, KOA2; double dE, E_new1, E_new2;
                Auger_decay(All_holes[NOP - 1].KOA, All_holes[NOP - 1].Shl, Sh1, KOA1, Sh2, KOA2, dE, E_new1, E_new2, st);
                if (Sh2 > 0) {
                    ev[TRK3_EV_AUGER]++;
                    if (std::fabs(All_holes[NOP - 1].E - (dE + E_new1 + E_new2)) > 1e-10) er[TRK3_ERR_AUGER_BALANCE]++;
                    double htheta, hphi;
                    Update_holes_angles_SHI(htheta, hphi, st);
                    { Hole &h = All_holes[NOP - 1]; h.t0 = h.tn; h.KOA = KOA1; h.Shl = Sh1; h.theta = htheta; h.phi = hphi; Hole_parameters(h, E_new1, st); cut_off_h(h); }
                    Tot_Nel = Tot_Nel + 1; n_el++;
                    Check_size(Tot_Nel);
                    Hole &h = All_holes[NOP - 1];
                    Electron &en = All_electrons[Tot_Nel - 1]; Hole &hn = All_holes[Tot_Nel - 1];
                    en.rng = rng.child(st, 1); hn.rng = rng.child(st, 2);
                    Update_holes_angles_SHI(htheta, hphi, st);
                    hn.t0 = h.t0; hn.X = h.X; hn.Y = h.Y; hn.Z = h.Z; hn.KOA = KOA2; hn.Shl = Sh2; hn.theta = htheta; hn.phi = hphi;
                    Hole_parameters(hn, E_new2, st);
                    cut_off_h(hn);
                    double IMFP = Next_free_path_2d(dE, T.ei_E, El_IMFP.data(), T.n_ei);
                    double EMFP = Next_free_path_2d(dE, T.ee_E, T.ee_L, T.n_ee);
                    RN = rng.rn(st);
                    double MFP_tot = -std::log(RN) / (1.0 / IMFP + 1.0 / EMFP);
                    RN = rng.rn(st); double phi1 = 2.0 * g_Pi * RN;
                    RN = rng.rn(st); double theta1 = g_Pi * RN;
                    en.E = dE; en.t0 = h.t0; en.X = h.X; en.Y = h.Y; en.Z = h.Z; en.L = MFP_tot; en.theta = theta1; en.phi = phi1;
                    tn_from(en, vel_el(en), MFP_tot);
                    cut_off_e(en);
                    if (en.E < -1.0e-9 || std::isnan(en.E)) er[TRK3_ERR_23]++;
                } else {
                    ev[TRK3_EV_AUGER_FROZEN]++;
                    All_holes[NOP - 1].t0 = t_cur; All_holes[NOP - 1].tn = 1e21;
                }
            } else {
                ev[TRK3_EV_RADIATIVE]++;
                int Sh1, KOA1; double dE, E_new1;
                Hole &h = All_holes[NOP - 1];
                Radiative_decay(h.KOA, h.Shl, Sh1, KOA1, dE, E_new1, st);
                double htheta, hphi;
                Update_holes_angles_SHI(htheta, hphi, st);
                h.t0 = h.tn; h.KOA = KOA1; h.Shl = Sh1; h.theta = htheta; h.phi = hphi;
                Hole_parameters(h, E_new1, st);
                cut_off_h(h);
                // Photon creation.  The reference stores the new photon at All_photons(Tot_Nphot+1) while Tot_Nphot is
                // DEcremented on absorption (:2963), so a live photon can be overwritten (energy is then lost).  The
                // slot bug is NOT reproduced: the photon goes to the first free slot (documented in DESIGN.md).
                Tot_Nphot = Tot_Nphot + 1; n_ph++;
                int slot = -1;
                for (size_t k = 0; k < All_photons.size(); ++k) if (All_photons[k].tn >= 1e20 && All_photons[k].E == 0.0) { slot = (int)k; break; }
                if (slot < 0) { Check_size_ph((int)All_photons.size() + 1); for (size_t k = 0; k < All_photons.size(); ++k) if (All_photons[k].tn >= 1e20 && All_photons[k].E == 0.0) { slot = (int)k; break; } }
                Photon &p = All_photons[slot];
                p.rng = rng.child(st, 3);
                double IMFP = Next_free_path_2d(dE, T.ph_E, Phot_IMFP.data(), T.n_ph);
                RN = rng.rn(st);
                double MFP_tot = -std::log(RN) * IMFP;
                RN = rng.rn(st); double phi1 = 2.0 * g_Pi * RN;
                RN = rng.rn(st); double theta1 = g_Pi * RN;
                p.E = dE; p.t0 = h.t0; p.X = h.X; p.Y = h.Y; p.Z = h.Z; p.L = MFP_tot; p.theta = theta1; p.phi = phi1;
                tn_from(p, g_cvel, MFP_tot);
                if (p.E < -1.0e-9 || std::isnan(p.E)) er[TRK3_ERR_30]++;
            }
        }
        All_holes[NOP - 1].rng = st;
    }

    // ---- Photon_Monte_Carlo, Monte_Carlo.f90:2873-2965
    void Photon_Monte_Carlo(int NOP) {
        ev[TRK3_EV_PHOTON]++;
        Photon &p = All_photons[NOP - 1];
        Stream st = p.rng;
        rng.align(st);
        double Eel = p.E, L = p.L, theta0 = p.theta, phi0 = p.phi;
        double X = p.X + L * std::sin(theta0) * std::sin(phi0), Y = p.Y + L * std::sin(theta0) * std::cos(phi0), Z = p.Z + L * std::cos(theta0);
        int Nat_cur, Nshl_cur;
        Which_shell(T.ph_E, T.ph_L, T.n_ph, Eel, st, Nat_cur, Nshl_cur);
        Tot_Nel = Tot_Nel + 1; n_el++;
        Check_size(Tot_Nel);
        Electron &en = All_electrons[Tot_Nel - 1]; Hole &hn = All_holes[Tot_Nel - 1];
        en.rng = rng.child(st, 1); hn.rng = rng.child(st, 2);
        double dE_cur = Electron_recieves_E(Eel, Nat_cur, Nshl_cur, st);
        double IMFP = Next_free_path_2d(dE_cur, T.ei_E, El_IMFP.data(), T.n_ei);
        double EMFP = Next_free_path_2d(dE_cur, T.ee_E, T.ee_L, T.n_ee);
        double RN = rng.rn(st);
        double MFP_tot = -std::log(RN) / (1.0 / IMFP + 1.0 / EMFP);
        double phi1, theta1;
        New_Angles_both(p.phi, p.theta, g_Pi / 2.0, 0.0, phi1, theta1);
        en.E = dE_cur; en.t0 = p.tn; en.X = X; en.Y = Y; en.Z = Z; en.L = MFP_tot; en.theta = theta1; en.phi = phi1;
        tn_from(en, vel_el(en), MFP_tot);
        cut_off_e(en);
        if (en.E < -1.0e-9 || std::isnan(en.E)) er[TRK3_ERR_50]++;
        double htheta, hphi;
        Update_holes_angles_SHI(htheta, hphi, st);
        hn.t0 = p.tn; hn.X = X; hn.Y = Y; hn.Z = Z; hn.KOA = Nat_cur; hn.Shl = Nshl_cur; hn.theta = htheta; hn.phi = hphi;
        Hole_parameters(hn, Eel - dE_cur, st);
        cut_off_h(hn);
        if (hn.Ehkin < -1.0e-9 || std::isnan(hn.Ehkin)) er[TRK3_ERR_51]++;
        if ((hn.E + hn.Ehkin) < Egap_) er[TRK3_ERR_52]++;
        Tot_Nphot = Tot_Nphot - 1;
        p.E = 0.0; p.t0 = 1.0e27; p.tn = 1.0e27; p.X = p.Y = p.Z = 0.0; p.L = 1.0e26; p.theta = p.phi = 0.0;
    }

    // ---- Find_min_time_particle, Monte_Carlo.f90:2060-2092 (scans the whole allocated arrays)
    void Find_min_time_particle(int &KOP, int &NOP, double &t_cur) {
        double te = 1e300, th = 1e300, tp = 1e300; int ie = 0, ih = 0, ip = 0;
        for (size_t k = 0; k < All_electrons.size(); ++k) if (All_electrons[k].tn < te) { te = All_electrons[k].tn; ie = (int)k + 1; }
        for (size_t k = 0; k < All_holes.size(); ++k) if (All_holes[k].tn < th) { th = All_holes[k].tn; ih = (int)k + 1; }
        if (cfg.include_photons) for (size_t k = 0; k < All_photons.size(); ++k) if (All_photons[k].tn < tp) { tp = All_photons[k].tn; ip = (int)k + 1; }
        KOP = 1; t_cur = SHI_loc.tn; NOP = 1;
        if (te < t_cur) { KOP = 2; t_cur = te; NOP = ie; }
        if (th < t_cur) { KOP = 3; t_cur = th; NOP = ih; }
        if (cfg.include_photons && tp < t_cur) { KOP = 4; t_cur = tp; NOP = ip; }
    }

    // ---- Calculated_statistics, Monte_Carlo.f90:881-1110
    void Calculated_statistics(int i, double tim, Tally &out, IterOut &io) {
        const int NR = T.n_r, NSH1 = T.nshl_atom1;
        out.a1(TRK3_OUT_TOT_NE, i) += (double)Tot_Nel;
        double sumEe = 0, sumEh = 0, sumEhk = 0, sumEph = 0;
        for (auto &e : All_electrons) sumEe += e.E;
        for (auto &h : All_holes) { sumEh += h.E; sumEhk += h.Ehkin; }
        double totE;
        if (cfg.include_photons) {
            int nph = 0;
            for (auto &p : All_photons) { sumEph += p.E; if (p.E > 0.0) nph++; }
            out.a1(TRK3_OUT_TOT_NPHOT, i) += (double)nph;      // == Tot_Nphot without the slot bug
            out.a1(TRK3_OUT_E_PHOT, i) += sumEph;
            totE = sumEe + sumEh + sumEhk + sumEph + At_NRG;
        } else totE = sumEe + sumEh + sumEhk + At_NRG;
        out.a1(TRK3_OUT_TOT_E, i) += totE;
        io.totE[i - 1] = totE; io.totNel[i - 1] = Tot_Nel;
        out.a1(TRK3_OUT_E_E, i) += sumEe;
        out.a1(TRK3_OUT_E_AT, i) += At_NRG;
        out.a1(TRK3_OUT_NE_EM, i) += (double)Em_Nel;
        { double s = 0; for (auto v : Em_electrons) s += v; out.a1(TRK3_OUT_E_EM, i) += s; }
        for (auto &h : All_holes) if (h.KOA > 0 && h.Shl <= NSH1) out.a3(TRK3_OUT_E_H, i, h.KOA, h.Shl, Nt, Nat) += h.E + h.Ehkin;
        if (cfg.include_photons) {
            for (auto &p : All_photons) if (p.E > 0.0) {
                double L0 = g_cvel * (tim - p.t0) * 1.0e-5; if (L0 < 0.0) L0 = 0.0;
                double X = p.X + L0 * std::sin(p.theta) * std::sin(p.phi), Y = p.Y + L0 * std::sin(p.theta) * std::cos(p.phi);
                double R = std::sqrt(X * X + Y * Y);
                int j = Find_in_monotonous_1D_array(T.out_R, NR, R);
                out.a2(TRK3_OUT_NPHOT, i, j, Nt) += T.out_V[j - 1];
                out.a2(TRK3_OUT_EPHOT, i, j, Nt) += p.E * T.out_V[j - 1];
            }
        }
        int N_VB_h_tot = 0;
        for (auto &h : All_holes) if (isVB(h.KOA, h.Shl)) N_VB_h_tot++;
        int N_VB_h = 0; double dsum = 0.0;
        const double cut = std::max(0.0, cfg.cut_off);
        for (int k = 1; k <= Tot_Nel; ++k) {
            const Electron &e = All_electrons[k - 1]; const Hole &h = All_holes[k - 1];
            double L0, theta0, phi0;
            if (e.E > cut) { double V = vel_el(e); L0 = V * (tim - e.t0) * 1.0e-5; if (L0 < 0.0) L0 = 0.0; theta0 = e.theta; phi0 = e.phi; }
            else { L0 = 0.0; theta0 = 0.0; phi0 = 0.0; }
            double X = e.X + L0 * std::sin(theta0) * std::sin(phi0), Y = e.Y + L0 * std::sin(theta0) * std::cos(phi0);
            double R = std::sqrt(X * X + Y * Y);
            int j = Find_in_monotonous_1D_array(T.out_R, NR, R);
            out.a2(TRK3_OUT_NE, i, j, Nt) += T.out_V[j - 1];
            out.a2(TRK3_OUT_EE, i, j, Nt) += e.E * T.out_V[j - 1];
            j = Find_in_monotonous_1D_array(T.out_R, NR, e.E);      // sic: radius grid used as energy grid (:1024)
            if (j > 1) out.a2(TRK3_OUT_EE_VS_E, i, j, Nt) += 1.0 / (T.out_R[j - 1] - T.out_R[j - 2]) / (double)Tot_Nel;
            else out.a2(TRK3_OUT_EE_VS_E, i, j, Nt) += 1.0 / T.out_R[j - 1] / (double)Tot_Nel;
            if (isVB(h.KOA, h.Shl)) {
                j = Find_in_monotonous_1D_array(T.dos_E, T.n_dos, h.Ehkin);
                if (j > 1) out.a2(TRK3_OUT_EH_VS_E, i, j, Nt) += 1.0 / (T.dos_E[j - 1] - T.dos_E[j - 2]) / (double)N_VB_h_tot;
                else out.a2(TRK3_OUT_EH_VS_E, i, j, Nt) += 1.0 / (T.dos_E[j] - T.dos_E[j - 1]) / (double)N_VB_h_tot;
            }
            double xx = theta0 / g_Pi * 180.0;
            if (e.E > 0.0) { j = Find_in_monotonous_1D_array(Out_theta1, TRK3_NTHETA, xx); out.a2(TRK3_OUT_THETA, i, j, Nt + 1) += 1.0 / (double)Tot_Nel; }
            double Xh, Yh;
            if (h.Mass < 1.0e3 && h.Ehkin > cut) {
                double V = vel_hole(h); L0 = V * (tim - h.t0) * 1.0e-5; if (L0 < 0.0) L0 = 0.0;
                theta0 = h.theta; phi0 = h.phi;
                Xh = h.X + L0 * std::sin(theta0) * std::sin(phi0); Yh = h.Y + L0 * std::sin(theta0) * std::cos(phi0);
                xx = theta0 / g_Pi * 180.0;
                j = Find_in_monotonous_1D_array(Out_theta1, TRK3_NTHETA, xx);
                out.a2(TRK3_OUT_THETA_H, i, j, Nt + 1) += 1.0 / (double)N_VB_h_tot;
            } else { Xh = h.X; Yh = h.Y; }
            R = std::sqrt(Xh * Xh + Yh * Yh);
            j = Find_in_monotonous_1D_array(T.out_R, NR, R);
            int l = h.KOA, m = h.Shl;
            if (l > 0 && m <= NSH1) {
                out.a4(TRK3_OUT_NH, i, j, l, m, Nt, NR, Nat) += T.out_V[j - 1];
                out.a4(TRK3_OUT_EH, i, j, l, m, Nt, NR, Nat) += h.E * T.out_V[j - 1];
                out.a4(TRK3_OUT_EHKIN, i, j, l, m, Nt, NR, Nat) += h.Ehkin * T.out_V[j - 1];
            }
            if (h.Mass < 1.0e3 && h.L < 1.0e3) { N_VB_h++; dsum += 1.0 / 3.0 * vel_hole(h) * h.L * 1.0e-6; }
        }
        io.diffS[i - 1] = dsum; io.diffN[i - 1] = N_VB_h;
        // Emitted electrons (:1100-1109): Out_E = Out_R/10
        for (int k = 1; k <= Em_Nel; ++k) {
            std::vector<double> OutE(NR); for (int q = 0; q < NR; ++q) OutE[q] = T.out_R[q] / 10.0;
            int j = Find_in_monotonous_1D_array(OutE.data(), NR, Em_electrons[k - 1]);
            if (j > 1) out.a2(TRK3_OUT_EE_VS_E_EM, i, j, Nt) += 1.0 / (OutE[j - 1] - OutE[j - 2]) / (double)Em_Nel;
            else out.a2(TRK3_OUT_EE_VS_E_EM, i, j, Nt) += 1.0 / OutE[j - 1] / (double)Em_Nel;
        }
    }

    // ---- Monte_Carlo_modelling, Monte_Carlo.f90:427-679
    void Monte_Carlo_modelling(uint64_t iter, int rng_mode, Tally &out, IterOut &io) {
        rng.init(rng_mode, cfg.seed, iter);
        if (cfg.work_function > 0) barrier_parameters(); else { Em_E1 = 0; Em_gamma = 0; }
        SHI_loc = IonT{};
        SHI_loc.E = cfg.shi_E; SHI_loc.t0 = 0.0; SHI_loc.tn = 0.0; SHI_loc.X = SHI_loc.Y = SHI_loc.Z = 0.0; SHI_loc.L = 0; SHI_loc.theta = SHI_loc.phi = 0;
        SHI_loc.Mass = cfg.shi_mass; SHI_loc.Zat = cfg.shi_Z; SHI_loc.Kind_Zeff = cfg.shi_kind_Zeff; SHI_loc.fixed_Zeff = cfg.shi_fixed_Zeff;
        SHI_loc.rng = Stream{0, 0};
        Equilibrium_charge_SHI(SHI_loc);               // MAIN.f90:171 (SHI%Zeff of the incoming ion)
        How_many_electrons();
        io.totE.assign(Nt, 0.0); io.diffS.assign(Nt, 0.0); io.totNel.assign(Nt, 0); io.diffN.assign(Nt, 0);
        Tot_Nel = 0; Tot_Nphot = 0; Em_Nel = 0;
        int i = 0; double t_cur = 0.0, tim_glob = 0.0;
        double SHI_IMFP = Next_free_path_2d(SHI_loc.E, T.shi_E, SHI_path.data(), T.n_shi);
        double RN = rng.rn(SHI_loc.rng);
        SHI_IMFP = -SHI_IMFP * std::log(RN);
        tn_from(SHI_loc, vel_ion(SHI_loc), SHI_IMFP);
        SHI_loc.L = SHI_IMFP;
        int KOP, NOP;
        Find_min_time_particle(KOP, NOP, t_cur);
        At_NRG = 0.0;
        while (tim_glob <= cfg.Tim - 1e-6 && i < Nt) {
            i = i + 1;
            tim_glob = std::min(lay.time_grid[i - 1], cfg.Tim);
            while (t_cur < std::min(tim_glob, cfg.Tim)) {
                switch (KOP) {
                case 1: SHI_Monte_Carlo(); if (SHI_loc.Z >= cfg.layer) SHI_loc.tn = 1e16; break;
                case 2: Electron_Monte_Carlo(NOP, i, out); break;
                case 3: Hole_Monte_Carlo(NOP, i, t_cur, out); break;
                case 4: Photon_Monte_Carlo(NOP); break;
                }
                Find_min_time_particle(KOP, NOP, t_cur);
            }
            Calculated_statistics(i, tim_glob, out, io);
            Find_min_time_particle(KOP, NOP, t_cur);
        }
    }
};

}  // namespace

extern "C" {

const char *trk3_oracle_version(void) { return "trekis3 oracle 0.1 (CPU restatement of Monte_Carlo.f90; test infrastructure)"; }

// Philox known-answer hook (Random123 kat_vectors) for tests.
void trk3_oracle_philox(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) { philox4x32_10(ctr[0], ctr[1], ctr[2], ctr[3], key[0], key[1], out); }

// Runs iterations [it_begin,it_end) and ADDS into tallies (lay.total doubles).
// rng_mode: 0 sequential stream per iteration, 1 per-particle Philox (same streams as the CUDA engine).
// iter_totE / iter_totNel: optional [n_iter][Nt] outputs.
int trk3_oracle_run(const trk3_config *cfg, const trk3_tables *tab, int64_t it_begin, int64_t it_end, int rng_mode, int threads,
                    double *tallies, trk3_stats *stats, double *iter_totE, double *iter_totNel) {
    if (!cfg || !tab || !tallies || it_end < it_begin) return TRK3_E_INVALID;
    trk3_tally_layout lay;
    int rc = trk3_tally_layout_init(cfg, tab, &lay);
    if (rc != TRK3_OK) return rc;
    if (cfg->kind_of_EMFP == 2 && (tab->n_dsf_e < 2 || tab->n_dsf_h < 2)) return TRK3_E_UNSUPPORTED;
    const int64_t n_it = it_end - it_begin;
    if (threads < 1) threads = (int)std::thread::hardware_concurrency();
    if (threads < 1) threads = 1;
    if (threads > n_it) threads = (int)std::max<int64_t>(1, n_it);
    std::vector<std::vector<double>> priv(threads, std::vector<double>(lay.total, 0.0));
    std::vector<IterOut> ios(n_it);
    std::vector<MC *> mcs(threads, nullptr);
    std::atomic<int64_t> next(0);
    auto work = [&](int t) {
        MC *mc = new MC(*cfg, *tab, lay); mcs[t] = mc;
        Tally out{priv[t].data(), &lay};
        for (;;) { int64_t k = next.fetch_add(1); if (k >= n_it) break; mc->Monte_Carlo_modelling((uint64_t)(it_begin + k), rng_mode, out, ios[k]); }
    };
    if (threads == 1) work(0);
    else { std::vector<std::thread> th; for (int t = 0; t < threads; ++t) th.emplace_back(work, t); for (auto &x : th) x.join(); }
    for (int t = 0; t < threads; ++t) for (int64_t q = 0; q < lay.total; ++q) tallies[q] += priv[t][q];
    // Out_diff_coeff (Monte_Carlo.f90:1094-1098): D(i) = (D(i) + sum_k v L/3)/N_VB_h applied once per iteration to the
    // RUNNING array, i.e. a recurrence over iterations; evaluated here in iteration order (single-thread semantics).
    for (int i = 0; i < lay.Nt; ++i) {
        double D = tallies[lay.off[TRK3_OUT_DIFF_COEFF] + i];
        for (int64_t k = 0; k < n_it; ++k) { D = D + ios[k].diffS[i]; if (ios[k].diffN[i] > 0) D = D / ios[k].diffN[i]; }
        tallies[lay.off[TRK3_OUT_DIFF_COEFF] + i] = D;
    }
    if (iter_totE) for (int64_t k = 0; k < n_it; ++k) for (int i = 0; i < lay.Nt; ++i) iter_totE[k * lay.Nt + i] = ios[k].totE[i];
    if (iter_totNel) for (int64_t k = 0; k < n_it; ++k) for (int i = 0; i < lay.Nt; ++i) iter_totNel[k * lay.Nt + i] = ios[k].totNel[i];
    if (stats) {
        std::memset(stats, 0, sizeof *stats);
        for (auto mc : mcs) if (mc) {
            for (int q = 0; q < TRK3_N_EVENT_CLASSES; ++q) stats->events[q] += mc->ev[q];
            for (int q = 0; q < TRK3_N_ERRORS; ++q) stats->errors[q] += mc->er[q];
            stats->n_electrons += mc->n_el; stats->n_photons += mc->n_ph;
        }
    }
    for (auto mc : mcs) delete mc;
    return TRK3_OK;
}

}  // extern "C"
