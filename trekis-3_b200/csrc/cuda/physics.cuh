// physics.cuh -- per-history physics of the TREKIS-3 cascade, written for the wavefront engine.
//
// The reference advances ALL particles of an iteration in one global time-ordered loop
// (Monte_Carlo.f90:572-665, Find_min_time_particle :2060).  Particles never interact (the field code
// :3018-3244 is dead), so the same tallies are obtained by following every particle independently
// and depositing a SNAPSHOT of its state at each output grid time its current free flight spans
// (Calculated_statistics, :881-1110).  Each routine below names the Fortran it re-expresses.
//
// All functions are templates over a context `C` that provides the side effects:
//   c.p                      const DevP&  (tables, switches, time grid)
//   c.push(species, rec)     append a particle to a queue (the context routes it: hot/cold, see electron_is_cold)
//   c.push_hot(species, rec) append to the queue of the full handlers regardless of the routing rule
//   c.push_ion(ev)           hand an impact ionisation over: its electron-hole pair is created by electron_ion_emit
//   c.snap(species, rec, i)  the particle's contribution to the tallies of grid time i: snapshot_electron/hole/photon,
//                            now or (CUDA engine) from a queue of snapshot records by a kernel of its own
//   c.tally(id, idx, v)      add v to element idx of Out_* array `id` (block-private or global)
//   c.add_u32(base, idx)     atomic ++ on a per-iteration integer histogram
//   c.add_f64(base, idx, v)  atomic += on a per-iteration double array
//   c.event(cls) / c.error(code) / c.count_electron() / c.count_photon()
//   C::kLean                 compile-time promise of the default switches (see elastic_dE); false = everything compiled in
// so that the identical code runs in the CUDA kernels (engine.cu) and in the CPU emulation used by
// the no-GPU tests (tests/emul).  fp64 throughout (the reference is built with -real-size 64).
#pragma once
#include <math.h>
#include "engine_types.h"

namespace trk3 {

// ---- Universal_Constants.f90:24-107
#define TRK_PI 3.1415926535897932384626433832795
#define TRK_GE 1.602176487e-19
#define TRK_ME 9.1093821545e-31
#define TRK_CVEL 299792458.0
#define TRK_MP (1836.1526724780 * TRK_ME)
#define TRK_HBAR 1.05457162853e-34
#define TRK_RY 13.6056981
#define TRK_A0 0.5291772085936
#define TRK_E0 8.854187817620e-12

TRK_HD bool trk_isnan(double x) { return x != x; }

// Elementary functions.  With TRK_MATH_OUTLINE the device code calls ONE out-of-line copy of each instead of expanding
// the 40-80 instruction sequence at every use (55 uses): the kernels are bound by instruction fetch, see philox4x32_10.
#if defined(TRK_MATH_OUTLINE) && defined(__CUDA_ARCH__)
#define TRK_MATH __device__ __noinline__
#else
#define TRK_MATH TRK_HD
#endif
TRK_MATH double m_exp(double x) { return exp(x); }
TRK_MATH double m_log(double x) { return log(x); }
TRK_MATH double m_sin(double x) { return sin(x); }
TRK_MATH double m_cos(double x) { return cos(x); }
TRK_MATH double m_acos(double x) { return acos(x); }
// sine and cosine of one angle share the argument reduction (same values as m_sin / m_cos)
struct SinCos { double s, c; };
TRK_MATH SinCos m_sincos(double x) { SinCos r; sincos(x, &r.s, &r.c); return r; }
// (divisions and square roots out of line were measured too: 21.4 ms instead of 20.9 ms per step -- they stay inline)
TRK_HD double m_div(double a, double b) { return a / b; }
TRK_HD double m_sqrt(double a) { return sqrt(a); }
}  // namespace trk3
// the closed forms of the delta-function CDF (shared with the host table builder); their logarithms through the one copy above
#define DLT_LOG(x) trk3::m_log(x)
#include "../common/trk3_delta.h"
namespace trk3 {

// ------------------------------------------------------------------------------------------------
// Philox4x32-10 counter-based RNG (Salmon et al., SC'11).  Stream = (particle id, draw index,
// iteration); key = user seed.  Replaces the unseeded compiler random_number of the reference.
// ------------------------------------------------------------------------------------------------
TRK_HD uint32_t trk_mulhi(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
    return __umulhi(a, b);
#else
    return (uint32_t)(((uint64_t)a * b) >> 32);
#endif
}
// The wave kernels are bound by instruction issue AND fetch (a 66 KB loop body against a 32 KB instruction cache): the ten
// rounds stay a loop, but of FIVE rounds per trip (measured, profiles/r2u_sweep_philox_builds.txt, C2 step / 4096-iteration
// batch in ms: one round per trip 12.81 / 53.89, two 12.77 / 53.05, five 12.77 / 52.61, straight-line 12.87 / 53.04, one
// out-of-line straight-line copy 12.83 / 53.46).  Per cold collision the two Philox blocks were 17 % of the thread instructions.
// Build switches (A/B builds: scripts/ab_build.sh, profiles/r2u_*): TRK_PHILOX_UNROLL = rounds per trip of the loop (1 = rolled,
// 10 = straight-line code), TRK_PHILOX_OUTLINE = one out-of-line copy per kernel that returns its four words in registers.
// The products are formed as 64-bit mul.wide (one IMAD.WIDE per pair of halves instead of IMAD.HI + IMAD).
struct Phx4 { uint32_t x, y, z, w; };
#if defined(__CUDA_ARCH__) && defined(TRK_PHILOX_OUTLINE)
#define TRK_PHX __device__ __noinline__
#else
#define TRK_PHX TRK_HD
#endif
#if !defined(TRK_PHILOX_UNROLL)
#if defined(TRK_PHILOX_ROLLED)
#define TRK_PHILOX_UNROLL 1
#else
#define TRK_PHILOX_UNROLL 10
#endif
#endif
#define TRK_PRAGMA(x) _Pragma(#x)
#define TRK_UNROLL(n) TRK_PRAGMA(unroll n)
TRK_PHX Phx4 philox4x32_10v(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
#if defined(__CUDA_ARCH__)
    TRK_UNROLL(TRK_PHILOX_UNROLL)
#endif
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        c0 = n0; c1 = (uint32_t)p1; c2 = n2; c3 = (uint32_t)p0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    Phx4 o; o.x = c0; o.y = c1; o.z = c2; o.w = c3;
    return o;
}
TRK_HD void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t &o0, uint32_t &o1, uint32_t &o2, uint32_t &o3) {
    const Phx4 o = philox4x32_10v(c0, c1, c2, c3, k0, k1);
    o0 = o.x; o1 = o.y; o2 = o.z; o3 = o.w;
}
// uniform in (0,1]: the reference can draw RN = 0 (m_log(RN), L/RN -> inf); we exclude it
// Draw k of a stream is half (k & 1) of Philox block k >> 1: one Philox evaluation serves two consecutive draws (the
// second half waits in the record's registers).  Collisions start at even draw indices (event_begin), so that the usual
// sequence of a collision -- channel roulette, transferred energy | azimuth, next free path -- costs two evaluations.
TRK_HD double rn(const DevP &p, Rec &r) {
    const uint32_t k = r.ctr++, blk = k >> 1;
    uint32_t a, b;
    if ((k & 1u) && r.rc_blk == blk) { a = r.rc_a; b = r.rc_b; }
    else {
        uint32_t o0, o1, o2, o3;
        philox4x32_10((uint32_t)r.id, (uint32_t)(r.id >> 32), blk, r.iter, p.seed_lo, p.seed_hi, o0, o1, o2, o3);
        if (k & 1u) { a = o2; b = o3; }
        else { a = o0; b = o1; r.rc_a = o2; r.rc_b = o3; r.rc_blk = blk; }
    }
    uint64_t bits = (uint64_t)a | ((uint64_t)b << 32);
    return (double)((bits >> 11) + 1) * (1.0 / 9007199254740992.0);
}
TRK_HD void event_begin(Rec &r) { r.ctr = (r.ctr + 1u) & ~1u; }
// id of a particle created by `parent` (tag 1 electron, 2 hole, 3 photon)
TRK_HD uint64_t child_id(const DevP &p, Rec &parent, uint32_t tag) {
    uint32_t a, b, c, d;
    philox4x32_10((uint32_t)parent.id, (uint32_t)(parent.id >> 32), parent.ctr++, parent.iter, p.seed_lo, p.seed_hi ^ (0x80000000u | tag), a, b, c, d);
    return (uint64_t)a | ((uint64_t)b << 32);
}

// ------------------------------------------------------------------------------------------------
// Searches (Reading_files_and_parameters.f90:3376-3559), same bisection paths as the Fortran
// (results depend on the path at exact grid hits and in padded, non-monotone table tails).
// All return the 1-based Fortran index.
// ------------------------------------------------------------------------------------------------
TRK_HD int find_1d(const double *A, int N, double v) {                 // Find_in_monotonous_1D_array
    if (v < A[0]) return 1;
    if (v >= A[N - 1]) return N;
    int i_1 = 1, i_2 = N, i_cur = (1 + N) >> 1;
    double temp_val = A[i_cur - 1];
    while (i_1 != i_2 - 1) {
        if (temp_val <= v) i_1 = i_cur; else i_2 = i_cur;
        i_cur = (i_1 + i_2) >> 1;
        temp_val = A[i_cur - 1];
    }
    return i_cur + 1;
}
TRK_HD int find_2d(const double *A, int N, double v) {                 // Find_in_monotonous_2D_array
    if (v < A[0]) return 1;
    if (v >= A[N - 1]) return N;
    int i_1 = 1, i_2 = N, i_cur = (1 + N) >> 1;
    double temp_val = A[i_cur - 1];
    for (int coun = 0; coun <= 1000; ++coun) {
        if (v >= temp_val && v <= A[i_cur]) break;
        if (temp_val <= v) i_1 = i_cur; else i_2 = i_cur;
        i_cur = (i_1 + i_2) >> 1;
        temp_val = A[i_cur - 1];
    }
    return i_cur + 1;
}
TRK_HD int find_dec(const double *A, int N, double v) {                // Find_in_monoton_array_decreasing
    if (v < A[N - 1]) return N;
    if (v > A[0]) return 1;
    int i_1 = 1, i_2 = N, i_cur = (1 + N) >> 1;
    double temp_val = A[i_cur - 1];
    int coun = 0;
    while ((i_2 - i_1 > 1 || i_1 - i_2 > 1) && coun <= 1000) {
        if (temp_val > v) i_1 = i_cur; else i_2 = i_cur;
        i_cur = (i_1 + i_2) >> 1;
        temp_val = A[i_cur - 1];
        ++coun;
    }
    return i_cur;
}
// Interpolate, Cross_sections.f90:4051-4086 (modes 1 and 5 are the only ones the MC uses)
TRK_HD double interp5(double E1, double E2, double S1, double S2, double En) {
    if (fabs(E2 - E1) < 1.0e-6) return (S1 > S2) ? S1 : S2;
    if (En == E1) return S1;
    double E2l = m_log(E2), E1l = m_log(E1), El = m_log(En), S1l = m_log(S1), S2l = m_log(S2);
    return m_exp(S1l + (S2l - S1l) / (E2l - E1l) * (El - E1l));
}
// same with the logarithms of the table entries precomputed (identical arithmetic, identical result)
TRK_HD double interp5t(double E1, double E2, double S1, double S2, double lE1, double lE2, double lS1, double lS2, double En, double lEn) {
    if (fabs(E2 - E1) < 1.0e-6) return (S1 > S2) ? S1 : S2;
    if (En == E1) return S1;
    return m_exp(lS1 + m_div(lS2 - lS1, lE2 - lE1) * (lEn - lE1));
}
// interp5t that also returns the logarithm of its result: the exponent it evaluated, or the tabulated logarithm of the
// table entry it returns.  (log(exp(x)) = x up to rounding: callers that interpolate the result again in log space
// use it instead of a second logarithm -- interpolate_transferred_energy does so twice per sampled energy.)
TRK_HD double interp5t_l(double E1, double E2, double S1, double S2, double lE1, double lE2, double lS1, double lS2, double En, double lEn, double &lres) {
    if (fabs(E2 - E1) < 1.0e-6) { if (S1 > S2) { lres = lS1; return S1; } lres = lS2; return S2; }
    if (En == E1) { lres = lS1; return S1; }
    lres = lS1 + m_div(lS2 - lS1, lE2 - lE1) * (lEn - lE1);
    return m_exp(lres);
}
TRK_HD double interp1(double E1, double E2, double S1, double S2, double En) {
    if (fabs(E2 - E1) < 1.0e-6) return (S1 > S2) ? S1 : S2;
    if (En == E1) return S1;
    return S1 + m_div(S2 - S1, E2 - E1) * (En - E1);
}

// Find_in_monotonous_1D_array through the direct-index accelerator: lv = m_log(v).  A strictly increasing array has
// exactly one index n with A[n-2] <= v < A[n-1]; the scan from the looked-up start finds it, as the bisection does.
TRK_HD int find_lut(const double *A, int N, const GridLut &g, double v, double lv) {
    if (v < A[0]) return 1;
    if (v >= A[N - 1]) return N;
    int b = (int)((lv - g.l0) * g.scale);
    b = (b < 0) ? 0 : ((b >= TRK_NLUT) ? TRK_NLUT - 1 : b);
    int j = g.lut[b];                          // in [1, N-1]
    while (j < N - 1 && A[j] <= v) ++j;
    while (j > 1 && A[j - 1] > v) --j;
    return j + 1;
}

// A 1-D mean-free-path table with its log companion and the accelerator of its energy grid
struct Tab { const double *E, *L, *lE, *lL; int N; const GridLut *g; };
TRK_HD int tab_find(const Tab &t, double E, double lE) { return find_lut(t.E, t.N, *t.g, E, lE); }

// Find_in_monotonous_2D_array gives the same index as the 1D variant unless the value sits exactly on the lower
// bracketing grid point, where the result depends on the bisection path: only then is the 2D search really run.
TRK_HD int find_2d_from_1d(const double *A, int N, double v, int n1) {
    if (n1 > 1 && v == A[n1 - 2]) return find_2d(A, N, v);
    return n1;
}
// Next_free_path_1d / _2d, Monte_Carlo.f90:1835-1899, for a known search index n
TRK_HD double nfp_at(const Tab &t, int n, bool two_d, double E, double lE) {
    if (n == 1) {
        double MFP = interp1(t.E[0], t.E[1], t.L[0], t.L[1], E);
        const double lo = two_d ? t.L[0] : t.E[0];      // sic (:1854 vs :1888): the 1d variant clamps against the energy array
        if (MFP < lo) MFP = lo;
        return MFP;
    }
    const double Ll = t.L[n - 2];
    if (Ll >= 1.0e16) return Ll;
    return interp5t(t.E[n - 2], t.E[n - 1], Ll, t.L[n - 1], t.lE[n - 2], t.lE[n - 1], t.lL[n - 2], t.lL[n - 1], E, lE);
}
TRK_HD double nfp_1d(const Tab &t, double E, double lE) { return nfp_at(t, tab_find(t, E, lE), false, E, lE); }
TRK_HD double nfp_2d(const Tab &t, double E, double lE) { return nfp_at(t, find_2d_from_1d(t.E, t.N, E, tab_find(t, E, lE)), true, E, lE); }

TRK_HD Tab tab_ei_tot(const DevP &p) { return Tab{p.ei_E, p.ei_tot, p.lei_E, p.lei_tot, p.n_ei, &p.lut[LUT_EI]}; }
TRK_HD Tab tab_ee(const DevP &p) { return Tab{p.ee_E, p.ee_L, p.lee_E, p.lee_L, p.n_ee, &p.lut[LUT_EE]}; }
TRK_HD Tab tab_hi_tot(const DevP &p) { return Tab{p.hi_E, p.hi_tot, p.lhi_E, p.lhi_tot, p.n_hi, &p.lut[LUT_HI]}; }
TRK_HD Tab tab_he(const DevP &p) { return Tab{p.he_E, p.he_L, p.lhe_E, p.lhe_L, p.n_he, &p.lut[LUT_HE]}; }
TRK_HD Tab tab_ph_tot(const DevP &p) { return Tab{p.ph_E, p.ph_tot, p.lph_E, p.lph_tot, p.n_ph, &p.lut[LUT_PH]}; }
TRK_HD Tab tab_shi_tot(const DevP &p) { return Tab{p.shi_E, p.shi_tot, p.lshi_E, p.lshi_tot, p.n_shi, &p.lut[LUT_SHI]}; }
// per-shell matrices of a family (rows [shell][energy] on the family's grid); tab_shell = one row
TRK_HD Tab tab_ei_L(const DevP &p) { return Tab{p.ei_E, p.ei_L, p.lei_E, p.lei_L, p.n_ei, &p.lut[LUT_EI]}; }
TRK_HD Tab tab_hi_L(const DevP &p) { return Tab{p.hi_E, p.hi_L, p.lhi_E, p.lhi_L, p.n_hi, &p.lut[LUT_HI]}; }
TRK_HD Tab tab_ph_L(const DevP &p) { return Tab{p.ph_E, p.ph_L, p.lph_E, p.lph_L, p.n_ph, &p.lut[LUT_PH]}; }
TRK_HD Tab tab_shi_L(const DevP &p) { return Tab{p.shi_E, p.shi_L, p.lshi_E, p.lshi_L, p.n_shi, &p.lut[LUT_SHI]}; }
TRK_HD Tab tab_shell(const Tab &m, int shell) { return Tab{m.E, m.L + (size_t)shell * m.N, m.lE, m.lL + (size_t)shell * m.N, m.N, m.g}; }

// total inelastic / elastic MFP of an electron (El_IMFP, El_EMFP of How_many_electrons, both searched with the 2D routine)
TRK_HD double electron_imfp(const DevP &p, double E, double lE) {
    if (E < p.e_cold) return p.e_imfp_cold;            // below the lowest threshold the table is one constant >= 1e16
    return nfp_2d(tab_ei_tot(p), E, lE);
}
TRK_HD double hole_imfp(const DevP &p, double E, double lE) {
    if (E < p.h_cold) return p.h_imfp_cold;
    return nfp_2d(tab_hi_tot(p), E, lE);
}

// Lookups of a particle's current energy that the next collision needs again: carried in registers between events
// (the reference recomputes them at the start of every event: Monte_Carlo.f90:2298-2299 == :2449-2450 of the previous one)
struct Cache {
    double lE;      // m_log(E)
    double emfp;    // elastic MFP from the 2D-searched table (El_EMFP / Hole_EMFP)
    double imfp;    // total inelastic MFP (El_IMFP / Hole_IMFP)
    double iemfp, iimfp;   // their reciprocals: the channel roulette of the next collision (:2301) and the free path sampled
                           // at the end of this one (:2451) divide 1 by the same two numbers
    int n1, n2;     // 1D / 2D search indices of E in the elastic grid
};
TRK_HDN void cache_fill(const Tab &el, double E, Cache &k) {
    k.lE = m_log(E);
    k.n1 = tab_find(el, E, k.lE);
    k.n2 = find_2d_from_1d(el.E, el.N, E, k.n1);
    k.emfp = nfp_at(el, k.n2, true, E, k.lE);
    k.iemfp = m_div(1.0, k.emfp);
    k.imfp = 0.0; k.iimfp = 0.0;
}
// Elastic_MFP%Total searched with the 1D routine (:2385, :2664): same table as El_EMFP, so the value is shared
// whenever both searches agree and the energy is inside the grid
TRK_HD double elastic_total(const Tab &el, double E, const Cache &k) {
    if (k.n1 > 1 && k.n1 == k.n2) return k.emfp;
    return nfp_at(el, k.n1, false, E, k.lE);
}

// all lookups of an electron / valence hole of kinetic energy E (what :2298-2299 / :2579-2580 evaluate at the start of an event)
TRK_HD void cache_electron(const DevP &p, double E, Cache &k) {
    cache_fill(tab_ee(p), E, k);
    if (E < p.e_cold) { k.imfp = p.e_imfp_cold; k.iimfp = p.e_iimfp_cold; }
    else { k.imfp = electron_imfp(p, E, k.lE); k.iimfp = m_div(1.0, k.imfp); }
}
TRK_HD void cache_vbhole(const DevP &p, double E, Cache &k) {
    cache_fill(tab_he(p), E, k);
    if (E < p.h_cold) { k.imfp = p.h_imfp_cold; k.iimfp = p.h_iimfp_cold; }
    else { k.imfp = hole_imfp(p, E, k.lE); k.iimfp = m_div(1.0, k.imfp); }
}

// Which_shell, Monte_Carlo.f90:1786-1832: shell roulette on 1/lambda_shell(E); returns the flat shell
// `m` = the family's per-shell matrix; n_out receives the search index of E in the family's grid (reused by the caller)
TRK_HDN int which_shell(const DevP &p, Rec &r, const Tab &m, double E, double lE, int &n_out) {
    double Temp[TRK3_MAX_SHELLS];
    double MFP_tot = 0.0;
    const double *Ea = m.E, *lEa = m.lE, *Lmat = m.L, *lLmat = m.lL;
    const int N = m.N;
    const int n = tab_find(m, E, lE);           // all shells share the grid (MAIN.f90:234-236)
    n_out = n;
    for (int s = 0; s < p.n_shells; ++s) {
        const double *La = Lmat + (size_t)s * N;
        double MFP;
        if (n == 1) MFP = 1.0e20;
        else {
            double a = La[n - 2], b = La[n - 1];
            if (a == b || a > 1e20) MFP = a;
            else { const double *lLa = lLmat + (size_t)s * N; MFP = interp5t(Ea[n - 2], Ea[n - 1], a, b, lEa[n - 2], lEa[n - 1], lLa[n - 2], lLa[n - 1], E, lE); }
        }
        Temp[s] = 1.0 / MFP;
        MFP_tot = MFP_tot + Temp[s];
    }
    double RN = rn(p, r);
    MFP_tot = RN * MFP_tot;
    double MFP_sum = 0.0;
    int sel = p.n_shells - 1;
    for (int s = 0; s < p.n_shells; ++s) { MFP_sum = MFP_sum + Temp[s]; if (MFP_sum >= MFP_tot) { sel = s; break; } }
    return sel;
}

// Get_velosity, Monte_Carlo.f90:850-878 (non-relativistic for massive particles)
TRK_HD double vel_electron(double E) { return m_sqrt(2.0 * E * TRK_GE / TRK_ME); }
TRK_HD double vel_hole(const Rec &h) {
    if (h.Mass < 1.0e6) {
        if (h.Ehkin < -1.0e-6 || h.Mass < 1.0e-10) return 0.0;
        if (fabs(h.Ehkin) < 1.0e-6) return 0.0;
        return m_sqrt(m_div(2.0 * h.Ehkin * TRK_GE, h.Mass * TRK_ME));
    }
    return 0.0;
}
// Get_time_of_next_event, :795-848
TRK_HD double next_time(double t0, double V, double MFP) { return (V > 1.0e-10) ? t0 + m_div(MFP, V) * 1e5 : 1.0e25; }

// time-interval index of an event/creation time: smallest i (1-based) with t < tg(i); Nt+1 if t >= Tim
TRK_HD int interval_of(const DevP &p, double t) {
    int i = 1;
    while (i <= p.Nt && !(t < p.tg[i - 1])) ++i;
    return i;
}

// interpolate_transferred_energy, Cross_sections.f90:1968-2045, on a CSR differential table with log companions.
// i_E is the 1D search index of Ele in the table's energy grid (shared with the MFP lookup: same grid).
struct Csr { const double *Eg, *lEg; int NE; const int64_t *off; const double *hw, *L, *lhw, *lL; };
// Find_in_monoton_array_decreasing in two rows at once: the two bisections are independent chains of dependent loads,
// interleaving them hides half of the latency.  Each follows exactly the path of find_dec.
TRK_HD void find_dec2(const double *A, int NA, const double *B, int NB, double v, int &ia, int &ib) {
    int a1 = 1, a2 = NA, ac = (1 + NA) >> 1, b1 = 1, b2 = NB, bc = (1 + NB) >> 1;
    bool adone = false, bdone = false;
    if (v < A[NA - 1]) { ac = NA; adone = true; } else if (v > A[0]) { ac = 1; adone = true; }
    if (v < B[NB - 1]) { bc = NB; bdone = true; } else if (v > B[0]) { bc = 1; bdone = true; }
    double ta = A[ac - 1], tb = B[bc - 1];
    for (int coun = 0; coun <= 1000; ++coun) {
        const bool ago = !adone && (a2 - a1 > 1 || a1 - a2 > 1), bgo = !bdone && (b2 - b1 > 1 || b1 - b2 > 1);
        if (!ago && !bgo) break;
        if (ago) { if (ta > v) a1 = ac; else a2 = ac; ac = (a1 + a2) >> 1; }
        if (bgo) { if (tb > v) b1 = bc; else b2 = bc; bc = (b1 + b2) >> 1; }
        if (ago) ta = A[ac - 1];
        if (bgo) tb = B[bc - 1];
    }
    ia = ac; ib = bc;
}
TRK_HD double sample_row_at(const Csr &t, int64_t o, int n, int i_hw, double L_need, double lLn, double &lres) {
    const double *L = t.L + o, *hw = t.hw + o, *lhw = t.lhw + o;
    if (i_hw == 1 || i_hw == n) { lres = lhw[i_hw - 1]; return hw[i_hw - 1]; }
    const double *lL = t.lL + o;
    return interp5t_l(L[i_hw - 1], L[i_hw], hw[i_hw - 1], hw[i_hw], lL[i_hw - 1], lL[i_hw], lhw[i_hw - 1], lhw[i_hw], L_need, lLn, lres);
}
// The same sample without its final exponential: the across-energy interpolation of interpolate_transferred_energy only needs the
// LOGARITHM of the two row samples (interp5t works in log space), so the two exponentials per sampled energy are not evaluated
// unless a rare branch asks for the values themselves.  `direct` = the sample is a table entry (hw then holds it), else hw is
// not set and lres is the exponent interp5t_l would have passed to m_exp.
TRK_HD void sample_row_log(const Csr &t, int64_t o, int n, int i_hw, double L_need, double lLn, double &lres, double &hw, bool &direct) {
    const double *L = t.L + o, *hwv = t.hw + o, *lhw = t.lhw + o;
    if (i_hw == 1 || i_hw == n) { lres = lhw[i_hw - 1]; hw = hwv[i_hw - 1]; direct = true; return; }
    const double E1 = L[i_hw - 1], E2 = L[i_hw];
    if (fabs(E2 - E1) < 1.0e-6) { const double S1 = hwv[i_hw - 1], S2 = hwv[i_hw]; direct = true; if (S1 > S2) { lres = lhw[i_hw - 1]; hw = S1; } else { lres = lhw[i_hw]; hw = S2; } return; }
    if (L_need == E1) { lres = lhw[i_hw - 1]; hw = hwv[i_hw - 1]; direct = true; return; }
    const double *lL = t.lL + o;
    lres = lhw[i_hw - 1] + m_div(lhw[i_hw] - lhw[i_hw - 1], lL[i_hw] - lL[i_hw - 1]) * (lLn - lL[i_hw - 1]);
    direct = false;
}
#define TRK_LN_1E_10 (-23.025850929940457)        // log(1e-10): exp(x) < 1e-10 <=> x < log(1e-10)
// the tail of interpolate_transferred_energy (Cross_sections.f90:2020-2045) from the two row samples in log form
TRK_HD double combine_rows_log(const Csr &t, int i_E, double lhw_1, double hw_1, bool d1, double lhw_2, double hw_2, bool d2, double Ele, double lE) {
    const double E1 = t.Eg[i_E - 1], E2 = t.Eg[i_E];
    const bool tiny = (d1 ? hw_1 < 1.0e-10 : lhw_1 < TRK_LN_1E_10) || (d2 ? hw_2 < 1.0e-10 : lhw_2 < TRK_LN_1E_10);
    if (tiny || fabs(E2 - E1) < 1.0e-6 || Ele == E1) {           // the branches that need the samples themselves: rare
        if (!d1) hw_1 = m_exp(lhw_1);
        if (!d2) hw_2 = m_exp(lhw_2);
        if (hw_1 < 1.0e-10 || hw_2 < 1.0e-10) return interp1(E1, E2, hw_1, hw_2, Ele);
        return interp5t(E1, E2, hw_1, hw_2, t.lEg[i_E - 1], t.lEg[i_E], lhw_1, lhw_2, Ele, lE);
    }
    return m_exp(lhw_1 + m_div(lhw_2 - lhw_1, t.lEg[i_E] - t.lEg[i_E - 1]) * (lE - t.lEg[i_E - 1]));      // interp5t, same arithmetic
}
TRK_HD double sample_row(const Csr &t, int64_t o, int n, double L_need, double lLn) {
    const double *L = t.L + o, *hw = t.hw + o;
    int i_hw = find_dec(L, n, L_need);
    if (i_hw == 1 || i_hw == n) return hw[i_hw - 1];
    const double *lL = t.lL + o, *lhw = t.lhw + o;
    return interp5t(L[i_hw - 1], L[i_hw], hw[i_hw - 1], hw[i_hw], lL[i_hw - 1], lL[i_hw], lhw[i_hw - 1], lhw[i_hw], L_need, lLn);
}
TRK_HDN double transferred_energy(const Csr &t, double Ele, double lE, int i_E, double L_need) {
    if (i_E > 1) { if (fabs(t.Eg[i_E - 2] - Ele) < 1.0e-6) i_E = i_E - 1; }
    const double lLn = m_log(L_need);
    int64_t o = t.off[i_E - 1];
    if (i_E <= 1) return sample_row(t, o, (int)(t.off[i_E] - o), L_need, lLn);
    const int64_t o2 = t.off[i_E - 2];
    const int n1 = (int)(t.off[i_E] - o), n2 = (int)(o - o2);
    int i1, i2;
    find_dec2(t.L + o, n1, t.L + o2, n2, L_need, i1, i2);
    double lhw_1, lhw_2, hw_1 = 0.0, hw_2 = 0.0;
    bool d1, d2;
    sample_row_log(t, o, n1, i1, L_need, lLn, lhw_1, hw_1, d1);
    i_E = i_E - 1;
    sample_row_log(t, o2, n2, i2, L_need, lLn, lhw_2, hw_2, d2);
    return combine_rows_log(t, i_E, lhw_1, hw_1, d1, lhw_2, hw_2, d2, Ele, lE);
}
TRK_HD Csr csr_eid(const DevP &p, int shell) { return Csr{p.ei_E, p.lei_E, p.n_ei, p.eid_off + (size_t)shell * p.n_ei, p.eid_hw, p.eid_L, p.leid_hw, p.leid_L}; }
TRK_HD Csr csr_eed(const DevP &p) { return Csr{p.ee_E, p.lee_E, p.n_ee, p.eed_off, p.eed_hw, p.eed_L, p.leed_hw, p.leed_L}; }
TRK_HD Csr csr_hid(const DevP &p) { return Csr{p.hi_E, p.lhi_E, p.n_hi, p.hid_off, p.hid_hw, p.hid_L, p.lhid_hw, p.lhid_L}; }
TRK_HD Csr csr_hed(const DevP &p) { return Csr{p.he_E, p.lhe_E, p.n_he, p.hed_off, p.hed_hw, p.hed_L, p.lhed_hw, p.lhed_L}; }

// effective hole mass from the DOS, e.g. Cross_sections.f90:1820-1825
// search in the DOS energy grid: direct index when the grid is uniform (reading_material_DOS resamples to 0.1 eV)
TRK_HD int find_dos(const DevP &p, double E) {
    const double *A = p.dos_E; const int N = p.n_dos;
    if (!(p.dos_inv_step > 0.0)) return find_1d(A, N, E);
    if (E < A[0]) return 1;
    if (E >= A[N - 1]) return N;
    int j = (int)(E * p.dos_inv_step) + 1;
    j = (j < 1) ? 1 : ((j > N - 1) ? N - 1 : j);
    while (j < N - 1 && A[j] <= E) ++j;
    while (j > 1 && A[j - 1] > E) --j;
    return j + 1;
}
TRK_HD double hole_mass_dos(const DevP &p, double E) { int m = find_dos(p, E); return p.dos_effm[m - 1]; }

// BEB shells (Target_atoms%KOCS = 2: negative shell designator in the .cdf).  dSigma_dw_int, Cross_sections.f90:3981-3984:
// the binary-encounter-Bethe cross section integrated over the transferred energy up to w0 (in units of the binding energy).
TRK_HD double beb_dsigma_dw_int(double S, double t0, double u0, double w0) {
    return S / (t0 + u0 + 1.0) * (-(m_log(w0 + 1.0) - m_log(fabs(t0 - w0))) / (t0 + 1.0) + (1.0 / (t0 - w0) - 1.0 / (w0 + 1.0))
                                  + m_log(t0) * 0.5 * (1.0 / ((t0 - w0) * (t0 - w0)) - 1.0 / ((w0 + 1.0) * (w0 + 1.0))));
}
// Electron_NRG_transfer_BEB, Cross_sections.f90:2128-2165: the transferred energy whose cumulative mean free path
// (dSigma_int_BEB :3958-3979) is L_need, by the reference's bisection (0.1 % tolerance, first probe at E = 0, where the
// cumulative cross section is exactly zero and the path infinite); the binding energy is added at the end.
TRK_HD_OUTLINE double beb_transfer(const DevP &p, double Ele, int shell, double L_need, double Mass, double Emin) {
    const double B = p.shell_Ip[shell], U = p.shell_Ek[shell], N = p.shell_Nel[shell];
    const double S = 4.0 * TRK_PI * TRK_A0 * TRK_A0 * N * (TRK_RY / B) * (TRK_RY / B);
    const double t0 = Ele / B, u0 = U / B;
    const double dSigma0 = beb_dsigma_dw_int(S, t0, u0, 0.0);
    const double temp1 = p.at_dens * 1e-24 * p.atom_pers[p.shell_atom[shell]] / p.sum_pers;
    double Emin1 = Emin, Emax1 = (Ele - B) / 2.0, E = 0.0;
    double Sigma_cur = beb_dsigma_dw_int(S, t0, u0, E / B) - dSigma0;
    double L_cur = 1.0 / (Mass * temp1 * Sigma_cur);
    int coun = 0;
    while (fabs(L_cur - L_need) / L_need > 0.001) {
        ++coun;
        Sigma_cur = beb_dsigma_dw_int(S, t0, u0, E / B) - dSigma0;
        L_cur = 1.0 / (Mass * temp1 * Sigma_cur);
        if (L_cur > L_need) Emin1 = E; else Emax1 = E;
        E = (Emax1 + Emin1) / 2.0;
        if (coun >= 1000) break;
    }
    return E + B;
}

// Delta-function CDF (kind_of_DR = 4): Electron_NRG_transfer_CDF hands over to get_inelastic_energy_transfer
// (Cross_sections.f90:1894-1895, 2051-2123), which draws a random number of its own and finds the transferred energy by bisection
// on the closed-form cross section (csrc/common/trk3_delta.h, shared with the host table builder) -- for electrons and for
// valence holes alike with the free-electron mass.
TRK_HD_OUTLINE double delta_transfer(const DevP &p, double Ele, int shell, double Emin, double RN) {
    const int o = p.osc_off[shell];
    return trk3delta::inelastic_energy_transfer(Ele, p.osc_E0 + o, p.osc_alpha + o, p.osc_off[shell + 1] - o, Emin, p.at_dens, RN);
}

// Electron_energy_transfer_inelastic (CS_method = 1), Cross_sections.f90:1793-1871
TRK_HD double inelastic_dE(const DevP &p, Rec &r, double Ele, double lE, int n_E, int shell, double L_tot, bool hole) {
    double RN = rn(p, r);
    double L_need = m_div(L_tot, RN);
    double Emin = p.shell_Ip[shell];
    if (Emin <= 1.0e-3) Emin = 1.0e-3;
    double Emax, E;
    if (!hole) {
        Emax = (Ele + Emin) / 2.0;
        if (p.shell_kocs[shell] == 2) E = beb_transfer(p, Ele, shell, L_need, 1.0, Emin);     // :1845-1846
        else if (p.delta_cdf) { const double RNd = rn(p, r); E = delta_transfer(p, Ele, shell, Emin, RNd); }
        else E = transferred_energy(csr_eid(p, shell), Ele, lE, n_E, L_need);
    } else {
        double Mass = (p.hole_mass >= 0) ? p.hole_mass : hole_mass_dos(p, Ele);
        Emax = 4.0 * Ele * Mass / ((Mass + 1.0) * (Mass + 1.0));
        if (p.shell_kocs[shell] == 2) E = beb_transfer(p, Ele, shell, L_need, Mass, Emin);
        else if (p.delta_cdf) { const double RNd = rn(p, r); E = delta_transfer(p, Ele, shell, Emin, RNd); }
        else E = transferred_energy(csr_hid(p), Ele, lE, n_E, L_need);
    }
    if (E < Emin) E = Emin;
    if (E > Emax) E = Emax;
    if (trk_isnan(E)) E = Emin;
    return E;
}

// rest_energy, Cross_sections.f90:1557
TRK_HD double rest_energy(double M0) { return M0 * TRK_CVEL * TRK_CVEL / TRK_GE; }

// NRG_transfer_elastic_atomic (Mott), Cross_sections.f90:3517-3613
TRK_HDN double mott_dE(const DevP &p, Rec &r, double Mat, double Zat, double Ee, double M_eff) {
    double RN = rn(p, r);
    double theta;
    double Erest = rest_energy(TRK_ME);
    double fact = Ee / Erest + 1.0;
    double v = TRK_CVEL * sqrt(1.0 - 1.0 / (fact * fact));
    if (v < 1.0e-6) theta = 0.0;
    else {
        double beta = v / TRK_CVEL, beta2 = beta * beta, tau = Ee / Erest;
        double alpha = TRK_GE * TRK_GE / (TRK_HBAR * TRK_CVEL * 4.0 * TRK_PI * TRK_E0);
        double nu = 1.7e-5 * pow(Zat, 2.0 / 3.0) * (1.0 - beta2) / beta2 * (1.13 + 3.76 * alpha * alpha / beta2 * Zat * Zat * sqrt(tau / (1.0 + tau)));
        double mu = (RN * (2.0 * nu + 1.0) - nu) / (RN + nu);
        theta = m_acos(mu);
    }
    double mc2 = rest_energy(TRK_ME * M_eff), Mct2 = rest_energy(Mat);
    double ct = m_cos(theta), ct2 = ct * ct, st2 = 1.0 - ct2;
    double Emc = Ee + mc2, E2mc = Ee + 2.0 * mc2, EmcMc = Emc + Mct2;
    double W1 = Emc * st2 + Mct2 - ct * sqrt(Mct2 * Mct2 - mc2 * mc2 * st2);
    double W2 = Ee * E2mc / (EmcMc * EmcMc - Ee * E2mc * ct2);
    return W1 * W2;
}
// kind_of_EMFP = 0: stoichiometric average of the atomic (Mott) energy transfers, Monte_Carlo.f90:2393-2401
TRK_HD_RARE double mott_elastic_dE(const DevP &p, Rec &r, double Eel, double M_eff) {
    double dE = 0.0;
    for (int ii = 0; ii < p.n_atoms; ++ii) {
        double dE_loc = mott_dE(p, r, p.atom_mass[ii] * TRK_MP, (double)p.atom_Z[ii], Eel, M_eff);
        dE = dE + dE_loc * p.atom_pers[ii];
    }
    return dE / p.sum_pers;
}
// elastic energy transfer: the kind_of_EMFP switch of Monte_Carlo.f90:2387-2407 / :2668-2692.
// LEAN (compile time, C::kLean of the context): the caller guarantees the default switches -- CDF elastic scattering
// (kind_of_EMFP = 1) and no electron emission (work_function <= 0) --, and the code of the other settings is not compiled
// in.  The wave kernels are bound by instruction fetch: code that never runs still costs (measured: -6 % step time).
// NRG_transfer_elastic_DSF, Cross_sections.f90:3652-3780 (kind_of_EMFP = 2): the energy an electron or a valence hole exchanges
// with the lattice, sampled from tabulated dynamic-structure-factor cross sections.  Positive = emission (the particle loses
// energy), NEGATIVE = absorption (it gains: "allowed in DSF formalism", Monte_Carlo.f90:713).  The reference interpolates the whole
// row of integrated mean free paths between the two tabulated particle energies around Eel (clipping negative entries to zero) and
// then searches it with Find_in_array_monoton on the NEGATED row; here the elements of the interpolated row are evaluated where the
// bisection probes them -- the same expression per element, the same index.
struct DsfRow {
    const double *La, *Lb, *Wa, *Wb; double v; bool two;      // rows of the two particle energies, interpolation weight Value1
    TRK_HD double L(int j) const { double x = two ? La[j] + (Lb[j] - La[j]) * v : La[j]; return (two && x < 0.0) ? 0.0 : x; }
    TRK_HD double W(int j) const { return two ? Wa[j] + (Wb[j] - Wa[j]) * v : Wa[j]; }
};
TRK_HD_RARE double dsf_elastic_dE(const DevP &p, Rec &r, double Eel, bool hole) {
    const double *gE = hole ? p.he_E : p.ee_E, *gL = hole ? p.he_L : p.ee_L, *gEm = hole ? p.he_emit : p.ee_emit, *gAb = hole ? p.he_absorb : p.ee_absorb;
    const int NE = hole ? p.n_he : p.n_ee, NW = hole ? p.n_dsf_h : p.n_dsf_e;
    const double *dE = hole ? p.dsf_h_dE : p.dsf_e_dE, *dEm = hole ? p.dsf_h_emit : p.dsf_e_emit, *dAb = hole ? p.dsf_h_absorb : p.dsf_e_absorb;
    // 0) the DSF energy grid is the grid of the elastic tables (Analytical_IMFPs.f90:913-915): one search serves :3665 and :3682
    int NumE = find_1d(gE, NE, Eel);
    if (NumE > 1) NumE = NumE - 1;
    const int i_MFP = NumE;
    // 1) emission or absorption (:3690-3705); linear_interpolation :3798-3802
    const double x1 = gE[i_MFP - 1], x2 = gE[i_MFP];
    const double EMFP_tot = gL[i_MFP - 1] + (gL[i_MFP] - gL[i_MFP - 1]) / (x2 - x1) * (Eel - x1);
    const double EMFP_emit = gEm[i_MFP - 1] + (gEm[i_MFP] - gEm[i_MFP - 1]) / (x2 - x1) * (Eel - x1);
    const double EMFP_absorb = gAb[i_MFP - 1] + (gAb[i_MFP] - gAb[i_MFP - 1]) / (x2 - x1) * (Eel - x1);
    double RN = rn(p, r);
    const bool it_is_emission = RN < EMFP_tot / EMFP_emit;
    // 2) the row between the two particle energies (:3712-3746)
    DsfRow row;
    row.two = (NumE != NE);
    row.v = row.two ? (Eel - gE[NumE - 1]) / (gE[NumE] - gE[NumE - 1]) : 0.0;
    const double *src = it_is_emission ? dEm : dAb;
    row.La = src + (size_t)(NumE - 1) * NW; row.Lb = row.two ? src + (size_t)NumE * NW : row.La;
    row.Wa = dE + (size_t)(NumE - 1) * NW; row.Wb = row.two ? dE + (size_t)NumE * NW : row.Wa;
    RN = rn(p, r);
    const double L_need = it_is_emission ? EMFP_emit / RN : EMFP_absorb / RN;
    // Linear_approx_2x1d_DSF, Reading_files_and_parameters.f90:3718-3745: Find_in_monotonous_1D_array on (-row, -L_need)
    int Number;
    {
        const double v = -L_need;
        if (v < -row.L(0)) Number = 1;
        else if (v >= -row.L(NW - 1)) Number = NW;
        else {
            int i_1 = 1, i_2 = NW, i_cur = (1 + NW) >> 1;
            double temp_val = -row.L(i_cur - 1);
            while (i_1 != i_2 - 1) {
                if (temp_val <= v) i_1 = i_cur; else i_2 = i_cur;
                i_cur = (i_1 + i_2) >> 1;
                temp_val = -row.L(i_cur - 1);
            }
            Number = i_cur + 1;
        }
    }
    double W;
    if (Number == 1) W = row.W(0) + (row.W(1) - row.W(0)) / (row.L(1) - row.L(0)) * (L_need - row.L(0));
    else if (fabs(row.L(Number - 1) - row.L(Number - 2)) < 1.0e-9) W = row.W(Number - 2);
    else if (row.L(Number - 2) > 1e20) W = row.W(Number - 2);
    else W = row.W(Number - 2) + (row.W(Number - 1) - row.W(Number - 2)) / (row.L(Number - 1) - row.L(Number - 2)) * (L_need - row.L(Number - 2));
    if (fabs(W) > 1.0 && it_is_emission && W > Eel) W = Eel;        // :3758-3771 (the messages of the reference are not printed)
    return W;
}

template <bool LEAN>
TRK_HD double elastic_dE(const DevP &p, Rec &r, double Eel, const Cache &k, double EMFP, bool hole, double M_eff) {
    if (!LEAN && p.kind_of_EMFP == 2) return dsf_elastic_dE(p, r, Eel, hole);
    if (LEAN || p.kind_of_EMFP == 1) {      // Electron_energy_transfer_elastic, Cross_sections.f90:2403-2413
        double RN = rn(p, r);
        double L_need = m_div(EMFP, RN);
        double hw = transferred_energy(hole ? csr_hed(p) : csr_eed(p), Eel, k.lE, k.n1, L_need);
        if (hw >= Eel) hw = Eel;
        return hw;
    }
    return mott_elastic_dE(p, r, Eel, hole ? M_eff : 1.0);
}

// cos_theta_from_W + Update_particle_angles_lat, Monte_Carlo.f90:1252-1299
TRK_HD void angles_lattice(const DevP &p, Rec &r, double E, double W, double M_eff, double &theta, double &phi) {
    double Erest_in = rest_energy(M_eff * TRK_ME), Erest_t = p.Erest_target;
    double E2mc = E + 2.0 * Erest_in, EmW = E - W;
    double W1 = E * E2mc - W * (E + Erest_in + Erest_t);
    double W2 = E * E2mc * EmW * (E2mc - W);
    double mu = (W2 > 0.0) ? m_div(W1, m_sqrt(W2)) : 0.0;
    if (fabs(mu) > 1.0) { double RN = rn(p, r); mu = m_cos(TRK_PI * RN); }
    theta = m_acos(mu);
    double RN2 = rn(p, r);
    phi = 2.0 * TRK_PI * RN2;
}
// New_Angles_both, Monte_Carlo.f90:1328-1360 (not a rotation; kept as is)
// (cos_theta0 = m_cos(theta0): the event handlers have it already from the position update)
TRK_HDN void new_angles_c(double phi0, double theta0, double cos_theta0, double theta, double psi, double &phi1, double &theta1) {
    const SinCos ps = m_sincos(psi);
    phi1 = phi0 + theta * cos_theta0 * ps.s;
    theta1 = theta0 + theta * ps.c;
    while (theta1 < 0.0) { theta1 = fabs(theta1); phi1 = phi1 + TRK_PI; }
    while (theta1 > TRK_PI) { theta1 = 2.0 * TRK_PI - theta1; phi1 = phi1 - TRK_PI; }
    // (:1353-1358) phi0 is in [0, 2 pi] and the steps above move it by at most 2 pi, so the number of turns to take off is ONE
    // almost always: floor / ceil of the quotient is then 1 and the division need not be evaluated (same value: 1 * 2 * pi)
    if (phi1 > 2.0 * TRK_PI) phi1 = (phi1 < 4.0 * TRK_PI) ? phi1 - 2.0 * TRK_PI : phi1 - floor(phi1 / (2.0 * TRK_PI)) * 2.0 * TRK_PI;
    if (phi1 < 0.0) phi1 = (phi1 >= -2.0 * TRK_PI) ? phi1 + 2.0 * TRK_PI : phi1 + ceil(fabs(phi1) / (2.0 * TRK_PI)) * 2.0 * TRK_PI;
}
TRK_HD void new_angles(double phi0, double theta0, double theta, double psi, double &phi1, double &theta1) { new_angles_c(phi0, theta0, m_cos(theta0), theta, psi, phi1, theta1); }
// Update_holes_angles_SHI, :1132-1140: isotropic in ANGLE (theta uniform), as the reference
TRK_HD void random_angles(const DevP &p, Rec &r, double &theta, double &phi) {
    double RN = rn(p, r); theta = TRK_PI * RN;
    double RN2 = rn(p, r); phi = 2.0 * TRK_PI * RN2;
}

// inverse-CDF level in the valence band (shared tail of From_where_in_VB :1522 and Electron_recieves_E :1653)
TRK_HD double vb_level(const DevP &p, Rec &r, int M_temp) {
    double Sum_DOS = p.dos_int[M_temp - 2];
    double RN = rn(p, r);
    double Tot_N = RN * Sum_DOS;
    int n = find_1d(p.dos_int, p.n_dos, Tot_N);
    if (n > 1) return p.dos_E[n - 2] + (p.dos_E[n - 1] - p.dos_E[n - 2]) * (Tot_N - p.dos_int[n - 2]) / (p.dos_int[n - 1] - p.dos_int[n - 2]);
    return p.dos_E[n - 1];
}
// Electron_recieves_E, Monte_Carlo.f90:1653-1716
template <class C>
TRK_HD double electron_receives_E(C &c, Rec &r, double dE, int shell) {
    const DevP &p = c.p;
    double E = dE - p.shell_Ip[shell], dE_cur = E;
    if (shell == p.vb_shell && !(dE <= p.shell_Ip[shell])) {
        int N = p.n_dos;
        int M_temp = (E < p.dos_E[N - 1]) ? find_dos(p, E) : N + 1;
        if (M_temp > 1) dE_cur = E - vb_level(p, r, M_temp);
    }
    if (dE_cur < 0.0) c.error(TRK3_ERR_10);
    return dE_cur;
}
// From_where_in_VB, :1522-1573
TRK_HD double from_where_in_VB(const DevP &p, Rec &r, bool haveE, double E) {
    int N = p.n_dos;
    if (!haveE) return vb_level(p, r, N + 1);
    int M_temp = (E < p.dos_E[N - 1]) ? find_dos(p, E) : N + 1;
    if (M_temp > 1) return vb_level(p, r, M_temp);
    return 0.0;
}

// Hole_parameters (+Assign_holes_mass), Monte_Carlo.f90:724-792.  `Ehkin_prev` is the hole's kinetic energy
// before the update (0 for a newly created hole: How_many_electrons initialises Ehkin = 0).
// `kout` (optional) receives the lookups of the new kinetic energy for the hole's next collision.
TRK_HDN void hole_parameters(const DevP &p, Rec &st, Rec &h, double Eh, double Ehkin_prev, Cache *kout = nullptr) {
    if (h.shell == p.vb_shell) {
        h.Ehkin = Eh - p.Egap;
        h.E = p.Egap;
        h.Mass = (p.hole_mass > 0) ? p.hole_mass : hole_mass_dos(p, h.Ehkin);
        if (h.Mass < 1.0e3) {
            Cache k;
            cache_vbhole(p, h.Ehkin, k);
            if (kout) *kout = k;
            // 1/HIMFP + 1/HEMFP with HEMFP = 1e30 if the hole's energy did not change (:746-747)
            const double iHEMFP = (Ehkin_prev == (Eh - p.Egap)) ? (1.0 / 1.0e30) : k.iemfp;
            double RN = rn(p, st);
            double MFP_tot = m_div(-m_log(RN), k.iimfp + iHEMFP);
            h.tn = next_time(h.t0, vel_hole(h), MFP_tot);
            h.L = MFP_tot;
        } else { h.L = 1.0e30; h.tn = 1.0e30; }
    } else {
        h.Mass = 1.0e29;
        double RN = rn(p, st);
        double nu = 1.0 / p.shell_auger[h.shell] + 1.0 / p.shell_radiat[h.shell];
        h.tn = h.t0 - m_log(RN) / nu;
        h.L = 1.0e30; h.E = Eh; h.Ehkin = 0.0;
    }
    // cut_off, :3008-3012
    if (h.Mass < 1e15 && h.Ehkin < p.cut_off) h.tn = 1.0e20;
}

// hand a newly created electron to the engine: Tot_Nel bookkeeping (the electron exists from the time interval of its
// creation on, Monte_Carlo.f90:1026) + append to the queue
template <class C>
TRK_HD void push_new_electron(C &c, const Rec &e) {
    const DevP &p = c.p;
    c.count_electron();
    c.add_u32(p.it.created, (size_t)(e.iter - p.batch_begin) * (p.Nt + 2) + interval_of(p, e.t0));
    c.push(SP_ELECTRON, e);
}

// a new electron at (X,Y,Z,t) with energy Ee and direction (theta,phi): the block repeated in every handler
// (:2212-2225, :2331-2342, :2606-2621, :2800-2813, :2921-2931); draws use the stream `st`: the event particle's, or
// (pairs created by the ion, see shi_emit) the new particle's own stream, which it then continues
template <class C>
TRK_HD void emit_electron(C &c, Rec &st, uint64_t id, double Ee, double t, double X, double Y, double Z, double theta, double phi, int err_code) {
    const DevP &p = c.p;
    Rec e;
    const double lE = m_log(Ee);
    double IMFP = electron_imfp(p, Ee, lE);
    double EMFP = nfp_2d(tab_ee(p), Ee, lE);
    double RN = rn(p, st);
    double MFP_tot = m_div(-m_log(RN), m_div(1.0, IMFP) + m_div(1.0, EMFP));
    e.E = Ee; e.Ehkin = 0.0; e.Mass = 1.0; e.t0 = t; e.X = X; e.Y = Y; e.Z = Z; e.L = MFP_tot; e.theta = theta; e.phi = phi;
    e.tn = next_time(t, vel_electron(Ee), MFP_tot);
    if (e.E < p.cut_off) e.tn = 1.0e20;
    if (e.E < -1.0e-9 || trk_isnan(e.E)) c.error(err_code);
    e.id = id; e.ctr = (st.id == id) ? st.ctr : 0; e.iter = st.iter; e.shell = -1;
    push_new_electron(c, e);
}
// a new hole in `shell` with total energy Eh and random direction (:2236-2240 and the identical blocks)
template <class C>
TRK_HD void emit_hole(C &c, Rec &st, uint64_t id, int shell, double Eh, double t, double X, double Y, double Z, int err_code) {
    const DevP &p = c.p;
    Rec h;
    double htheta, hphi;
    random_angles(p, st, htheta, hphi);
    h.E = 0.0; h.Ehkin = 0.0; h.Mass = 1.0e30; h.t0 = t; h.tn = 1.0e21; h.X = X; h.Y = Y; h.Z = Z; h.L = 1.0e30; h.theta = htheta; h.phi = hphi;
    h.shell = shell; h.id = id; h.iter = st.iter;
    hole_parameters(p, st, h, Eh, 0.0);
    h.ctr = (st.id == id) ? st.ctr : 0;
    if (h.Ehkin < -1.0e-9 || trk_isnan(h.Ehkin)) c.error(err_code);
    c.push(shell == p.vb_shell ? SP_VBHOLE : SP_COREHOLE, h);
}

// Equilibrium_charge_SHI, Cross_sections.f90:2641-2680
TRK_HD double shi_zeff(const DevP &p, double E) {
    double vp = (E > 0.0) ? sqrt(2.0 * E * TRK_GE / (p.ion_mass * TRK_MP)) : 0.0;
    double Zp = (double)p.ion_Z;
    const double g_v0 = sqrt(2.0 * TRK_RY * TRK_GE / TRK_ME);
    switch (p.ion_kind_Zeff) {
    case 1: return Zp * (1.0 - m_exp(-(vp / g_v0 / p.ion_pow23)));
    case 2: { double c1 = 0.6, c2 = 0.45; return Zp * pow(1.0 + pow(vp / (pow(Zp, c2) * g_v0 * 4.0 / 3.0), -1.0 / c1), -c1); }
    case 3: {
        double sz = 0; for (int a = 0; a < p.n_atoms; ++a) sz += p.atom_Z[a] * p.atom_pers[a];
        double Zt = sz / p.sum_pers;
        double c1 = 1.0 - 0.26 * m_exp(-Zt / 11.0 - (Zt - Zp) * (Zt - Zp) / 9.0);
        double vpvo = pow(Zp, -0.543) * vp / g_v0;
        double c2 = 1.0 + 0.03 * vpvo * m_log(Zt);
        double x = c1 * pow(vpvo / c2 / 1.54, 1.0 + 1.83 / Zp), x2 = x * x, x4 = x2 * x2;
        return Zp * (8.29 * x + x4) / (0.06 / x + 4.0 + 7.4 * x + x4); }
    case 4: return p.ion_fixed_Zeff;
    default: return Zp * (1.0 - m_exp(-(vp * 125.0 / TRK_CVEL / p.ion_pow23)));
    }
}
// SHI_energy_transfer (CDF shells), Monte_Carlo.f90:1719-1780.  The reference's linear search over 1/L
// (Find_in_1D_array) is kept as a forward scan from the threshold row: same first index with 1/L >= Tot_N.
// the part of SHI_energy_transfer that only depends on the shell (:1735-1747): the cumulative MFP at the ionisation
// potential.  Evaluated once per shell when the tables are bound (DevP::shi_Mtemp, shi_dL).
TRK_HD void shi_threshold(const double *Ea, const double *La, int N, double E_cur, int &M_temp, double &dL) {
    if (N < 1) { M_temp = 1; dL = 0.0; return; }
    M_temp = find_1d(Ea, N, E_cur);
    if (M_temp > 1) {
        if (La[M_temp - 2] > 1.0e-10) dL = interp5(Ea[M_temp - 2], Ea[M_temp - 1], La[M_temp - 2], La[M_temp - 1], E_cur);
        else dL = interp1(Ea[M_temp - 2], Ea[M_temp - 1], La[M_temp - 2], La[M_temp - 1], E_cur);
    } else dL = La[0];
}
// Two halves, so that a caller can evaluate the logarithm between them together with others (k_shi, engine.cu):
// shi_transfer_target = the sampled cumulative inverse MFP, shi_transfer_energy = its inverse lookup.
TRK_HD double shi_transfer_target(const DevP &p, int shell, double RN) {
    const int64_t o = p.dshi_off[shell];
    const double *La = p.dshi_L + o;
    const int N = (int)(p.dshi_off[shell + 1] - o);
    const double dL = p.shi_dL[shell];
    return (dL > 0.0 && La[N - 1] > 0.0) ? 1.0 / dL + RN * (1.0 / La[N - 1] - 1.0 / dL) : 1.5e21;
}
TRK_HD double shi_transfer_energy(const DevP &p, int shell, double Tot_N, double lTot) {
    const int64_t o = p.dshi_off[shell];
    const double *Ea = p.dshi_E + o, *lEa = p.ldshi_E + o, *iLa = p.dshi_iL + o, *liLa = p.ldshi_iL + o;
    int N = (int)(p.dshi_off[shell + 1] - o);
    const int M_temp = p.shi_Mtemp[shell];
    int N_temmp;
    if (Tot_N < 1e20) {
        // 1/L is non-decreasing (cumulative cross section): the first index with 1/L >= Tot_N, clamped to N.  The tables have
        // thousands of rows: a direct index (bins uniform in 1/L, the measure Tot_N is sampled with) + local scan instead of
        // 13 dependent loads of a bisection.
        const GridLut &g = p.dshi_lut[shell];
        int j;
        if (g.scale > 0.0) {
            int b = (int)((Tot_N - g.l0) * g.scale);
            b = (b < 0) ? 0 : ((b >= TRK_NLUT) ? TRK_NLUT - 1 : b);
            j = g.lut[b];
            while (j < N && iLa[j - 1] < Tot_N) ++j;
            while (j > 1 && !(iLa[j - 2] < Tot_N)) --j;
        } else {
            int lo = 1, hi = N;
            while (lo < hi) { int mid = (lo + hi) >> 1; if (iLa[mid - 1] < Tot_N) lo = mid + 1; else hi = mid; }
            j = lo;
        }
        N_temmp = j;
    } else N_temmp = M_temp;
    if (N_temmp > M_temp) return interp5t(iLa[N_temmp - 2], iLa[N_temmp - 1], Ea[N_temmp - 2], Ea[N_temmp - 1], liLa[N_temmp - 2], liLa[N_temmp - 1], lEa[N_temmp - 2], lEa[N_temmp - 1], Tot_N, lTot);
    return p.shell_Ip[shell];
}
TRK_HD double shi_energy_transfer(const DevP &p, Rec &r, int shell) {
    const double Tot_N = shi_transfer_target(p, shell, rn(p, r));
    return shi_transfer_energy(p, shell, Tot_N, m_log(Tot_N));
}

// ------------------------------------------------------------------------------------------------
// Snapshots: the per-particle share of Calculated_statistics (Monte_Carlo.f90:881-1110) at grid time i
// ------------------------------------------------------------------------------------------------
template <class C>
TRK_HD void snapshot_electron(C &c, const Rec &e, int i) {
    const DevP &p = c.p;
    const double tim = p.tg[i - 1];
    const uint32_t il = e.iter - p.batch_begin;
    const double cut = (p.cut_off > 0.0) ? p.cut_off : 0.0;
    double L0 = 0.0, theta0 = 0.0, phi0 = 0.0;
    if (e.E > cut) { L0 = vel_electron(e.E) * (tim - e.t0) * 1.0e-5; if (L0 < 0.0) L0 = 0.0; theta0 = e.theta; phi0 = e.phi; }
    double st = m_sin(theta0);
    const SinCos sp = m_sincos(phi0);
    double X = e.X + L0 * st * sp.s, Y = e.Y + L0 * st * sp.c;
    double R = sqrt(X * X + Y * Y);
    int j = find_1d(p.out_R, p.n_r, R);
    c.tally(TRK3_OUT_NE, (i - 1) + (int64_t)p.Nt * (j - 1), p.out_V[j - 1]);
    c.tally(TRK3_OUT_EE, (i - 1) + (int64_t)p.Nt * (j - 1), e.E * p.out_V[j - 1]);
    c.tally(TRK3_OUT_E_E, i - 1, e.E);
    c.add_f64(p.it.esnap, (size_t)il * p.Nt + (i - 1), e.E);
    j = find_1d(p.out_R, p.n_r, e.E);                 // sic (:1024): the radius grid doubles as the energy grid
    c.add_u32(p.it.spec_e, ((size_t)il * p.Nt + (i - 1)) * p.n_r + (j - 1));
    if (e.E > 0.0) {
        double xx = theta0 / TRK_PI * 180.0;
        int jt = (xx < 1.0) ? 1 : ((xx >= 180.0) ? 180 : (int)floor(xx) + 1);   // Find(Out_theta1 = 1..180, xx)
        c.add_u32(p.it.th_e, ((size_t)il * p.Nt + (i - 1)) * TRK3_NTHETA + (jt - 1));
    }
}
template <class C>
TRK_HD void snapshot_hole(C &c, const Rec &h, int i) {
    const DevP &p = c.p;
    const double tim = p.tg[i - 1];
    const uint32_t il = h.iter - p.batch_begin;
    const size_t base = (size_t)il * p.Nt + (i - 1);
    const double cut = (p.cut_off > 0.0) ? p.cut_off : 0.0;
    const bool vb = (h.shell == p.vb_shell);
    if (vb) {
        c.add_u32(p.it.nvb, base);
        int j = find_dos(p, h.Ehkin);
        c.add_u32(p.it.spec_h, base * p.n_dos + (j - 1));
    }
    double Xh = h.X, Yh = h.Y;
    if (h.Mass < 1.0e3 && h.Ehkin > cut) {
        double L0 = vel_hole(h) * (tim - h.t0) * 1.0e-5; if (L0 < 0.0) L0 = 0.0;
        double st = m_sin(h.theta);
        const SinCos sp = m_sincos(h.phi);
        Xh = h.X + L0 * st * sp.s; Yh = h.Y + L0 * st * sp.c;
        double xx = h.theta / TRK_PI * 180.0;
        int jt = (xx < 1.0) ? 1 : ((xx >= 180.0) ? 180 : (int)floor(xx) + 1);
        c.add_u32(p.it.th_h, base * TRK3_NTHETA + (jt - 1));
    }
    double R = sqrt(Xh * Xh + Yh * Yh);
    int j = find_1d(p.out_R, p.n_r, R);
    int l = p.shell_atom[h.shell], m = p.shell_num[h.shell];     // 0-based (KOA-1, Shl-1)
    if (m < p.nshl_atom1) {
        int64_t idx = (i - 1) + (int64_t)p.Nt * ((j - 1) + (int64_t)p.n_r * (l + (int64_t)p.n_atoms * m));
        c.tally(TRK3_OUT_NH, idx, p.out_V[j - 1]);
        c.tally(TRK3_OUT_EH, idx, h.E * p.out_V[j - 1]);
        c.tally(TRK3_OUT_EHKIN, idx, h.Ehkin * p.out_V[j - 1]);
        c.tally(TRK3_OUT_E_H, (i - 1) + (int64_t)p.Nt * (l + (int64_t)p.n_atoms * m), h.E + h.Ehkin);
    }
    c.add_f64(p.it.esnap, base, h.E + h.Ehkin);
    if (h.Mass < 1.0e3 && h.L < 1.0e3) {      // Out_diff_coeff, :1089-1096
        c.add_u32(p.it.diffN, base);
        c.add_f64(p.it.diffS, base, 1.0 / 3.0 * vel_hole(h) * h.L * 1.0e-6);
    }
}
template <class C>
TRK_HD void snapshot_photon(C &c, const Rec &ph, int i) {
    const DevP &p = c.p;
    if (!(ph.E > 0.0)) return;
    const double tim = p.tg[i - 1];
    const uint32_t il = ph.iter - p.batch_begin;
    const size_t base = (size_t)il * p.Nt + (i - 1);
    double L0 = TRK_CVEL * (tim - ph.t0) * 1.0e-5; if (L0 < 0.0) L0 = 0.0;
    double st = m_sin(ph.theta);
    const SinCos sp = m_sincos(ph.phi);
    double X = ph.X + L0 * st * sp.s, Y = ph.Y + L0 * st * sp.c;
    double R = sqrt(X * X + Y * Y);
    int j = find_1d(p.out_R, p.n_r, R);
    c.tally(TRK3_OUT_NPHOT, (i - 1) + (int64_t)p.Nt * (j - 1), p.out_V[j - 1]);
    c.tally(TRK3_OUT_EPHOT, (i - 1) + (int64_t)p.Nt * (j - 1), ph.E * p.out_V[j - 1]);
    c.tally(TRK3_OUT_E_PHOT, i - 1, ph.E);
    c.add_u32(p.it.nph, base);
    c.add_f64(p.it.esnap, base, ph.E);
}

// the snapshot of a particle of species `sp` (valence and core holes share snapshot_hole)
template <class C>
TRK_HD void snapshot_any(C &c, int sp, const Rec &r, int i) {
    if (sp == SP_ELECTRON) snapshot_electron(c, r, i);
    else if (sp == SP_PHOTON) snapshot_photon(c, r, i);
    else snapshot_hole(c, r, i);
}

// lattice energy of an elastic event in time interval iv at radius R (Monte_Carlo.f90:2414-2436, :2702-2719)
template <class C>
TRK_HD void deposit_lattice(C &c, const Rec &r, int iv, double X, double Y, double dE) {
    const DevP &p = c.p;
    double R = sqrt(X * X + Y * Y);
    if (trk_isnan(R)) { c.error(TRK3_ERR_NAN); return; }
    int j = find_1d(p.out_R, p.n_r, R);
    c.tally(TRK3_OUT_ELAT, (iv - 1) + (int64_t)p.Nt * (j - 1), dE * p.out_V[j - 1]);
    c.add_f64(p.it.elat, (size_t)(r.iter - p.batch_begin) * (p.Nt + 2) + iv, dE);
}

// ------------------------------------------------------------------------------------------------
// Event handlers.  Each processes the collision of particle `r` at time r.tn (< Tim) and leaves `r`
// in its post-collision state with a new tn.  `iv` is the time interval of the event.
// ------------------------------------------------------------------------------------------------

// calculate_emission, Monte_Carlo.f90:2477-2513: an electron that has crossed the surface (Z < 0) is emitted or reflected
template <class C>
TRK_HD_RARE void electron_emission(C &c, Rec &e, int iv) {
    const DevP &p = c.p;
    bool emitted = false; double Ekin = 0.0;
    if (e.E >= 1.5 * p.bar_height) { emitted = true; Ekin = e.E - p.work_function; }
    else {
        double r2 = rn(p, e);
        double Em_Penetr = 1.0 / (1.0 + m_exp(p.Em_gamma * (p.Em_E1 - e.E)));
        Ekin = e.E - p.work_function;
        if (Ekin > 0.0 && r2 < Em_Penetr) emitted = true;
        else if (m_cos(e.theta) < 0) e.theta = TRK_PI - e.theta;
    }
    if (emitted) {
        e.tn = 1.0e30; e.L = 1.0e30;
        const size_t b = (size_t)(e.iter - p.batch_begin) * (p.Nt + 2) + iv;
        c.add_u32(p.it.em_cnt, b);
        c.add_f64(p.it.em_E, b, Ekin);
        const int j = find_1d(p.out_R, p.n_r, Ekin * 10.0);      // Out_E = Out_R/10 (:940-941)
        c.add_u32(p.it.em_spec, b * p.n_r + (j - 1));
    }
}

// An impact ionisation by an electron, as far as the creation of its electron-hole pair needs it (:2321-2371).
// The pair does not feed back into the history of the primary, so its creation is split off (electron_ion_emit) and
// runs as parallel work instead of lengthening the serial chain of a delta-electron, which is the critical path of a
// batch.  Stream convention as for the ion (shi_emit): the primary's stream gives the primary's draws and the
// children's ids; the draws that create the electron / the hole come from the new particle's own stream.
struct IonEvent {
    double dE, t, X, Y, Z;          // transferred energy, time and place of the collision
    double theta0, phi0;            // direction of the primary before the collision
    double theta, phi;              // its scattering angles (Update_electron_angles_El)
    int32_t shell;
    uint64_t pid; uint32_t ctr0;    // primary's stream at the creation of the pair (children's ids)
    uint32_t iter;
};
template <class C>
TRK_HD void electron_ion_emit(C &c, const IonEvent &ev) {
    const DevP &p = c.p;
    Rec par; par.id = ev.pid; par.ctr = ev.ctr0; par.iter = ev.iter;
    const uint64_t id_e = child_id(p, par, 1), id_h = child_id(p, par, 2);
    Rec se; se.id = id_e; se.ctr = 0; se.iter = ev.iter;
    Rec sh; sh.id = id_h; sh.ctr = 0; sh.iter = ev.iter;
    double dE_cur = electron_receives_E(c, se, ev.dE, ev.shell);
    double phi1, theta1;
    new_angles(ev.phi0, ev.theta0, TRK_PI / 2.0 - ev.theta, ev.phi + TRK_PI, phi1, theta1);
    emit_electron(c, se, id_e, dE_cur, ev.t, ev.X, ev.Y, ev.Z, theta1, phi1, TRK3_ERR_21);
    emit_hole(c, sh, id_h, ev.shell, ev.dE - dE_cur, ev.t, ev.X, ev.Y, ev.Z, TRK3_ERR_20);
}

// Electron_Monte_Carlo, Monte_Carlo.f90:2253-2474.
// RN is the first draw of the collision (the channel roulette, :2301), made by the caller.
// MODE selects what is compiled in: EV_ANY = both channels (run-time roulette), EV_ELASTIC / EV_INELASTIC = one
// channel only, for kernels whose warps are uniform in the event type (the caller has evaluated the roulette with
// electron_roulette_inelastic).
enum EventMode { EV_ANY = 0, EV_ELASTIC = 1, EV_INELASTIC = 2 };
TRK_HD bool electron_roulette_inelastic(const Cache &k, double RN) { const double ii = k.iimfp; return RN * (ii + k.iemfp) < ii; }
// The collision in three parts, so that a warp whose lanes disagree on the channel runs the channel-specific middle part
// once per channel but the common head and tail once for all of them (k_hot, engine.cu):
//   electron_event_head   the place of the collision (:2290-2296)
//   electron_event_inel / electron_event_elast   transferred energy and scattering angles of the channel
//   electron_event_tail   lookups of the new energy, next free flight, new direction (:2449-2465)
struct ElEvent { double X, Y, Z, cos_t0, dE, theta, phi; };
TRK_HD void electron_event_head(const Rec &e, ElEvent &s) {
    const double L = e.L;
    const SinCos sc_t = m_sincos(e.theta), sc_p = m_sincos(e.phi);
    const double st0 = sc_t.s;
    s.X = e.X + L * st0 * sc_p.s; s.Y = e.Y + L * st0 * sc_p.c; s.Z = e.Z + L * sc_t.c;
    s.cos_t0 = sc_t.c;
}
template <class C>
TRK_HD void electron_event_inel(C &c, Rec &e, const Cache &k, ElEvent &s) {     // impact ionisation
    const DevP &p = c.p;
    const double Eel = e.E;
    c.event(TRK3_EV_EL_INEL);
    int n_E;
    int shell = which_shell(p, e, tab_ei_L(p), Eel, k.lE, n_E);
    IonEvent ev;
    ev.pid = e.id; ev.ctr0 = e.ctr; ev.iter = e.iter;
    e.ctr += 2;                                              // the two child ids
    const double IMFP = nfp_at(tab_shell(tab_ei_L(p), shell), n_E, false, Eel, k.lE);      // Next_free_path_1d, same grid => same index
    const double dE = inelastic_dE(p, e, Eel, k.lE, n_E, shell, IMFP, false);
    double theta = m_acos((Eel - dE) / sqrt(Eel * (Eel - dE)));       // Update_electron_angles_El :1189
    if (trk_isnan(theta)) { double r2 = rn(p, e); theta = r2 * TRK_PI; }
    double phi; { double r2 = rn(p, e); phi = 2.0 * TRK_PI * r2; }
    ev.dE = dE; ev.t = e.tn; ev.X = s.X; ev.Y = s.Y; ev.Z = s.Z; ev.theta0 = e.theta; ev.phi0 = e.phi; ev.theta = theta; ev.phi = phi; ev.shell = shell;
    c.push_ion(ev);                                          // the pair: electron_ion_emit
    s.dE = dE; s.theta = theta; s.phi = phi;
}
template <class C>
TRK_HD void electron_event_elast(C &c, Rec &e, int iv, const Cache &k, ElEvent &s) {     // elastic: energy to the lattice
    const DevP &p = c.p;
    const double Eel = e.E;
    c.event(TRK3_EV_EL_ELAST);
    const double EMFP = elastic_total(tab_ee(p), Eel, k);
    s.dE = elastic_dE<C::kLean>(p, e, Eel, k, EMFP, false, 1.0);
    angles_lattice(p, e, Eel, s.dE, 1.0, s.theta, s.phi);
    if (trk_isnan(s.theta) || trk_isnan(s.phi)) c.error(TRK3_ERR_NAN);
    deposit_lattice(c, e, iv, s.X, s.Y, s.dE);
}
template <class C>
TRK_HD void electron_event_tail(C &c, Rec &e, int iv, Cache &k, const ElEvent &s) {
    const DevP &p = c.p;
    const double Eel = e.E, t_ev = e.tn;
    cache_electron(p, Eel - s.dE, k);                               // :2449-2450, kept for the next collision
    const double RN = rn(p, e);
    double MFP_tot = m_div(-m_log(RN), k.iimfp + k.iemfp);
    double phi1, theta1;
    new_angles_c(e.phi, e.theta, s.cos_t0, s.theta, s.phi, phi1, theta1);
    e.E = Eel - s.dE; e.t0 = t_ev; e.X = s.X; e.Y = s.Y; e.Z = s.Z; e.L = MFP_tot; e.theta = theta1; e.phi = phi1;
    e.tn = next_time(e.t0, vel_electron(e.E), MFP_tot);
    if (e.E < p.cut_off) e.tn = 1.0e20;
    if (!C::kLean && p.work_function > 0 && e.Z < 0.0) electron_emission(c, e, iv);
    if (e.E < -1.0e-9 || trk_isnan(e.E)) c.error(TRK3_ERR_22);
}
template <int MODE, class C>
TRK_HD void electron_event_t(C &c, Rec &e, int iv, Cache &k, double RN) {
    ElEvent s;
    electron_event_head(e, s);
    if (MODE == EV_INELASTIC || (MODE == EV_ANY && electron_roulette_inelastic(k, RN))) electron_event_inel(c, e, k, s);
    else electron_event_elast(c, e, iv, k, s);
    electron_event_tail(c, e, iv, k, s);
}

// check_hole_parameters, Monte_Carlo.f90:682-721: snap the scattered hole to a populated DOS level
TRK_HD void check_hole_level(const DevP &p, double Eel, double &dE, double &Ehole, double *E_new_electron) {
    Ehole = Eel - dE;
    int mhole = find_dos(p, Ehole);
    if (p.dos_DOS[mhole - 1] < 1.0e-4) {
        if (E_new_electron) {
            while (mhole > 1 && p.dos_DOS[mhole - 1] < 1.0e-4) mhole = mhole - 1;
            double Eloc = p.dos_E[mhole - 1];
            *E_new_electron = *E_new_electron + (Ehole - Eloc);
            Ehole = Eloc;
        } else {
            while (mhole < p.n_dos && p.dos_DOS[mhole - 1] < 1.0e-4) mhole = mhole + 1;
            double Eloc = p.dos_E[mhole - 1];
            dE = dE + (Ehole - Eloc);
            Ehole = Eloc;
        }
    }
}

// Hole_Monte_Carlo, valence-band branch, Monte_Carlo.f90:2560-2738.  RN = first draw (channel roulette, :2582); MODE as
// for electrons.  Holes below DevP::h_cold have a total inelastic MFP >= 1e16 and can never ionise (needs HIMFP < 1e15).
TRK_HD bool vbhole_roulette_inelastic(const Cache &k, double RN) { const double ii = k.iimfp; return RN * (ii + k.iemfp) < ii && k.imfp < 1e15; }
template <int MODE, class C>
TRK_HD void vbhole_event_t(C &c, Rec &h, int iv, Cache &k, double RN) {
    const DevP &p = c.p;
    const double Eel = h.Ehkin;
    double HIMFP = k.imfp, HEMFP = k.emfp;                        // :2579-2580
    const double L = h.L, theta0 = h.theta, phi0 = h.phi;
    const SinCos sc_t = m_sincos(theta0), sc_p = m_sincos(phi0);
    const double st0 = sc_t.s;
    const double X = h.X + L * st0 * sc_p.s, Y = h.Y + L * st0 * sc_p.c, Z = h.Z + L * sc_t.c;
    const double t_ev = h.tn;
    double dE, Ehole, htheta1, hphi1;
    if (MODE == EV_INELASTIC || (MODE == EV_ANY && vbhole_roulette_inelastic(k, RN))) {
        c.event(TRK3_EV_VBH_INEL);
        int n_E;
        int shell = which_shell(p, h, tab_hi_L(p), Eel, k.lE, n_E);
        uint64_t id_e = child_id(p, h, 1), id_h = child_id(p, h, 2);
        HIMFP = nfp_at(tab_shell(tab_hi_L(p), shell), n_E, false, Eel, k.lE);
        dE = inelastic_dE(p, h, Eel, k.lE, n_E, shell, HIMFP, true);
        // Update_holes_angles_el, :1142-1168
        double E11 = Eel - dE, Mh = h.Mass * TRK_ME;
        double htheta = m_acos(sqrt((Mh + TRK_ME) * (Mh + TRK_ME) / (4.0 * Mh * TRK_ME) * dE / Eel));
        double hphi; { double r2 = rn(p, h); hphi = 2.0 * TRK_PI * r2; }
        if (trk_isnan(htheta)) { double r2 = rn(p, h); htheta = TRK_PI * r2; }
        htheta1 = m_acos((Eel * (Mh - TRK_ME) + E11 * (Mh + TRK_ME)) / (2 * Mh * sqrt(Eel * E11)));
        hphi1 = hphi + TRK_PI;
        if (trk_isnan(htheta1)) { double r2 = rn(p, h); htheta1 = TRK_PI * r2; }
        double dE_cur = electron_receives_E(c, h, dE, shell);
        // the new electron is fully sampled first (:2606-2621), then the level check may add the surplus to it (:2660)
        const double lEn = m_log(dE_cur);
        double IMFP = electron_imfp(p, dE_cur, lEn);
        double EMFP = nfp_2d(tab_ee(p), dE_cur, lEn);
        RN = rn(p, h);
        double MFP_tot = m_div(-m_log(RN), m_div(1.0, IMFP) + m_div(1.0, EMFP));
        RN = rn(p, h);                                           // sic (:2611): drawn and discarded
        double phi1, theta1;
        new_angles_c(phi0, theta0, sc_t.c, htheta, hphi, phi1, theta1);
        Rec e;
        e.E = dE_cur; e.Ehkin = 0.0; e.Mass = 1.0; e.t0 = t_ev; e.X = X; e.Y = Y; e.Z = Z; e.L = MFP_tot; e.theta = theta1; e.phi = phi1;
        e.tn = next_time(t_ev, vel_electron(dE_cur), MFP_tot);
        if (e.E < p.cut_off) e.tn = 1.0e20;
        if (e.E < -1.0e-9 || trk_isnan(e.E)) c.error(TRK3_ERR_40);
        e.id = id_e; e.ctr = 0; e.iter = h.iter; e.shell = -1;
        emit_hole(c, h, id_h, shell, dE - dE_cur, t_ev, X, Y, Z, TRK3_ERR_41);
        check_hole_level(p, Eel, dE, Ehole, &e.E);               // NB: tn/L of the new electron keep the pre-shift energy, as in the reference
        push_new_electron(c, e);
    } else {
        c.event(TRK3_EV_VBH_ELAST);
        HEMFP = elastic_total(tab_he(p), Eel, k);
        dE = elastic_dE<C::kLean>(p, h, Eel, k, HEMFP, true, h.Mass);
        angles_lattice(p, h, Eel, dE, h.Mass, htheta1, hphi1);
        check_hole_level(p, Eel, dE, Ehole, nullptr);
        deposit_lattice(c, h, iv, X, Y, dE);
    }
    double hphi2, htheta2;
    new_angles_c(phi0, theta0, sc_t.c, htheta1, hphi1, hphi2, htheta2);
    h.t0 = t_ev; h.X = X; h.Y = Y; h.Z = Z; h.theta = htheta2; h.phi = hphi2;
    hole_parameters(p, h, h, Ehole + p.Egap, Eel, &k);
    if (h.Ehkin < -1.0e-9 || trk_isnan(h.Ehkin)) c.error(TRK3_ERR_20);
}


// count_for_Auger_shells / Choose_for_Auger_shell, Monte_Carlo.f90:1576-1647
TRK_HD double auger_count(const DevP &p, double NRG, bool second_e) {
    double coun = 0.0;
    for (int s = 0; s < p.n_shells; ++s) {
        double E_delta = second_e ? 1.0e10 : NRG - p.shell_Ip[s];
        if (NRG > p.shell_Ip[s] + 1.0e-3 && E_delta >= p.Egap) coun = coun + p.shell_Nel[s];
    }
    return coun;
}
TRK_HD int auger_choose(const DevP &p, double NRG, double Shel, bool second_e, int dflt) {
    double coun_sh = 0.0;
    int sel = dflt;
    for (int s = 0; s < p.n_shells; ++s) {
        double E_delta = second_e ? 1.0e10 : NRG - p.shell_Ip[s];
        if (NRG > p.shell_Ip[s] + 1.0e-3 && E_delta >= p.Egap) {
            coun_sh = coun_sh + p.shell_Nel[s];
            if (coun_sh >= Shel) sel = s;
        }
        if (coun_sh >= Shel) break;
    }
    return sel;
}

// Hole_Monte_Carlo, deep-shell branch (Auger / radiative decay), Monte_Carlo.f90:2739-2868 with
// Auger_decay :1366-1443 and Radiative_decay :2969-2997
template <class C>
TRK_HD void corehole_event(C &c, Rec &h) {
    const DevP &p = c.p;
    const double t_ev = h.tn;
    const int sh0 = h.shell;
    event_begin(h);
    double RN = rn(p, h);
    const double t_Auger = p.shell_auger[sh0], t_Radiat = p.shell_radiat[sh0];
    const double Ip0 = p.shell_Ip[sh0];
    if (RN * (1.0 / t_Auger + 1.0 / t_Radiat) < 1.0 / t_Auger) {
        double coun = auger_count(p, Ip0, false);
        RN = rn(p, h);
        int s1 = auger_choose(p, Ip0, RN * coun, false, sh0);
        double dE_cur = (s1 == p.vb_shell) ? from_where_in_VB(p, h, false, 0.0) : 0.0;
        double E_new1 = dE_cur + p.shell_Ip[s1];
        double Energy_diff = Ip0 - E_new1;
        coun = auger_count(p, Energy_diff, true);
        int s2 = -1; double Ee = -1.0e-10, E_new2 = 0.0;
        if (coun > 0.0) {
            RN = rn(p, h);
            s2 = auger_choose(p, Energy_diff, RN * coun, true, -1);
        }
        if (s2 >= 0) {
            dE_cur = (s2 == p.vb_shell) ? from_where_in_VB(p, h, true, Energy_diff - p.shell_Ip[s2]) : 0.0;
            E_new2 = dE_cur + p.shell_Ip[s2];
            Ee = Energy_diff - E_new2;
        }
        if (Ee < 0.0 || E_new1 < 0.0 || E_new2 < 0.0) c.error(TRK3_ERR_25);
        if (s2 >= 0) {
            c.event(TRK3_EV_AUGER);
            if (fabs(h.E - (Ee + E_new1 + E_new2)) > 1e-10) c.error(TRK3_ERR_AUGER_BALANCE);
            double htheta, hphi;
            random_angles(p, h, htheta, hphi);
            const double Ehk_prev = h.Ehkin;
            h.t0 = t_ev; h.shell = s1; h.theta = htheta; h.phi = hphi;
            hole_parameters(p, h, h, E_new1, Ehk_prev);
            uint64_t id_e = child_id(p, h, 1), id_h = child_id(p, h, 2);
            emit_hole(c, h, id_h, s2, E_new2, t_ev, h.X, h.Y, h.Z, TRK3_ERR_20);
            // Auger electron: isotropic in angle (:2800-2813); phi is drawn before theta
            const double lEn = m_log(Ee);
            double IMFP = electron_imfp(p, Ee, lEn);
            double EMFP = nfp_2d(tab_ee(p), Ee, lEn);
            RN = rn(p, h);
            double MFP_tot = m_div(-m_log(RN), m_div(1.0, IMFP) + m_div(1.0, EMFP));
            RN = rn(p, h); double phi1 = 2.0 * TRK_PI * RN;
            RN = rn(p, h); double theta1 = TRK_PI * RN;
            Rec e;
            e.E = Ee; e.Ehkin = 0.0; e.Mass = 1.0; e.t0 = t_ev; e.X = h.X; e.Y = h.Y; e.Z = h.Z; e.L = MFP_tot; e.theta = theta1; e.phi = phi1;
            e.tn = next_time(t_ev, vel_electron(Ee), MFP_tot);
            if (e.E < p.cut_off) e.tn = 1.0e20;
            if (e.E < -1.0e-9 || trk_isnan(e.E)) c.error(TRK3_ERR_23);
            e.id = id_e; e.ctr = 0; e.iter = h.iter; e.shell = -1;
            push_new_electron(c, e);
        } else {
            c.event(TRK3_EV_AUGER_FROZEN);
            h.t0 = t_ev; h.tn = 1e21;                             // :2828
        }
    } else {
        c.event(TRK3_EV_RADIATIVE);
        double coun = auger_count(p, Ip0, true);
        RN = rn(p, h);
        int s1 = auger_choose(p, Ip0, RN * coun, true, sh0);
        double dE_cur = (s1 == p.vb_shell) ? from_where_in_VB(p, h, false, 0.0) : 0.0;
        double E_new1 = dE_cur + p.shell_Ip[s1];
        double dE = Ip0 - E_new1;
        double htheta, hphi;
        random_angles(p, h, htheta, hphi);
        const double Ehk_prev = h.Ehkin;
        h.t0 = t_ev; h.shell = s1; h.theta = htheta; h.phi = hphi;
        hole_parameters(p, h, h, E_new1, Ehk_prev);
        // photon (:2843-2866); the reference's slot-reuse bug (:2844 vs :2963) does not exist here
        uint64_t id_p = child_id(p, h, 3);
        double IMFP = nfp_2d(tab_ph_tot(p), dE, m_log(dE));
        RN = rn(p, h);
        double MFP_tot = -m_log(RN) * IMFP;
        RN = rn(p, h); double phi1 = 2.0 * TRK_PI * RN;
        RN = rn(p, h); double theta1 = TRK_PI * RN;
        Rec ph;
        ph.E = dE; ph.Ehkin = 0.0; ph.Mass = 0.0; ph.t0 = t_ev; ph.X = h.X; ph.Y = h.Y; ph.Z = h.Z; ph.L = MFP_tot; ph.theta = theta1; ph.phi = phi1;
        ph.tn = next_time(t_ev, TRK_CVEL, MFP_tot);
        if (ph.E < -1.0e-9 || trk_isnan(ph.E)) c.error(TRK3_ERR_30);
        ph.id = id_p; ph.ctr = 0; ph.iter = h.iter; ph.shell = -1;
        c.count_photon();
        c.push(SP_PHOTON, ph);
    }
}

// Photon_Monte_Carlo, Monte_Carlo.f90:2873-2965: photoabsorption; the photon disappears
template <class C>
TRK_HD void photon_event(C &c, Rec &ph) {
    const DevP &p = c.p;
    c.event(TRK3_EV_PHOTON);
    event_begin(ph);
    const double Eel = ph.E, L = ph.L, theta0 = ph.theta, phi0 = ph.phi, t_ev = ph.tn;
    const SinCos sc_t = m_sincos(theta0), sc_p = m_sincos(phi0);
    const double st0 = sc_t.s;
    const double X = ph.X + L * st0 * sc_p.s, Y = ph.Y + L * st0 * sc_p.c, Z = ph.Z + L * sc_t.c;
    int n_E;
    int shell = which_shell(p, ph, tab_ph_L(p), Eel, m_log(Eel), n_E);
    uint64_t id_e = child_id(p, ph, 1), id_h = child_id(p, ph, 2);
    double dE_cur = electron_receives_E(c, ph, Eel, shell);
    double phi1, theta1;
    new_angles_c(phi0, theta0, sc_t.c, TRK_PI / 2.0, 0.0, phi1, theta1);
    emit_electron(c, ph, id_e, dE_cur, t_ev, X, Y, Z, theta1, phi1, TRK3_ERR_50);
    emit_hole(c, ph, id_h, shell, Eel - dE_cur, t_ev, X, Y, Z, TRK3_ERR_51);
    ph.E = 0.0; ph.t0 = 1.0e27; ph.tn = 1.0e27;
}

// SHI_Monte_Carlo + the ion part of Monte_Carlo_modelling (:548-570, :593-597, :2153-2249).
// The trajectory of an ion is one serial chain of collisions; everything that does not feed back into the chain (the
// creation of the electron-hole pair of a collision) is split off so that it can run in parallel over all collisions
// of all ions:  shi_begin / shi_step = the ion's own part, producing one ShiEvent per collision;  shi_emit = the pair.
// Stream convention: the ion's stream gives the ion's draws and the children's ids; the draws that create the
// electron (level in the band, azimuth, first free path) come from the new electron's own stream and those that
// create the hole (direction, first free path / decay time) from the new hole's own stream.
struct ShiEvent {
    double dE;          // transferred energy
    double E_after;     // ion energy after the collision (Update_electron_angles_SHI uses the updated ion)
    double t0, Z;       // time and depth of the collision
    int32_t shell;
    uint32_t ctr0;      // position of the ion's stream at the creation of the pair (children's ids)
    uint32_t iter;
};
TRK_HD void shi_begin(const DevP &p, Rec &s, uint32_t iter) {
    s.E = p.ion_E; s.t0 = 0.0; s.tn = 0.0; s.X = 0.0; s.Y = 0.0; s.Z = 0.0; s.L = 0.0; s.theta = 0.0; s.phi = 0.0; s.Ehkin = 0.0; s.Mass = p.ion_mass;
    s.id = 0; s.ctr = 0; s.iter = iter; s.shell = -1;
    const double MSHI = p.ion_mass * TRK_MP;
    double lam = nfp_2d(tab_shi_tot(p), s.E, m_log(s.E));
    double RN = rn(p, s);
    s.L = -lam * m_log(RN);
    s.tn = next_time(s.t0, sqrt(2.0 * s.E * TRK_GE / MSHI), s.L);
}
// one collision of the ion at s.tn (< Tim); leaves the ion in flight towards its next collision
template <class C>
TRK_HD void shi_step(C &c, Rec &s, ShiEvent &ev) {
    const DevP &p = c.p;
    const double MSHI = p.ion_mass * TRK_MP;
    c.event(TRK3_EV_SHI);
    event_begin(s);
    const double lEs = m_log(s.E);
    int n_E;
    int shell = which_shell(p, s, tab_shi_L(p), s.E, lEs, n_E);
    double dE = shi_energy_transfer(p, s, shell);
    const Tab tt = tab_shi_tot(p);
    double lam = nfp_at(tt, find_2d_from_1d(tt.E, tt.N, s.E, n_E), true, s.E, lEs);
    double RN = rn(p, s);
    double SHI_IMFP = -lam * m_log(RN);
    double Z = s.Z + s.L;
    s.E = s.E - dE; s.t0 = s.tn; s.Z = Z; s.L = SHI_IMFP;
    s.tn = next_time(s.t0, sqrt(2.0 * s.E * TRK_GE / MSHI), SHI_IMFP);
    ev.dE = dE; ev.E_after = s.E; ev.t0 = s.t0; ev.Z = Z; ev.shell = shell; ev.ctr0 = s.ctr; ev.iter = s.iter;
    s.ctr += 2;                                                   // the two child ids
    if (s.Z >= p.layer) s.tn = 1e16;                              // :597 the ion has left the layer
}
// the (electron, hole) pair of one ion collision, :2198-2247
template <class C>
TRK_HD void shi_emit(C &c, const ShiEvent &ev) {
    const DevP &p = c.p;
    const double MSHI = p.ion_mass * TRK_MP;
    Rec ion; ion.id = 0; ion.ctr = ev.ctr0; ion.iter = ev.iter;
    const uint64_t id_e = child_id(p, ion, 1), id_h = child_id(p, ion, 2);
    Rec se; se.id = id_e; se.ctr = 0; se.iter = ev.iter;          // the new electron's stream
    Rec sh; sh.id = id_h; sh.ctr = 0; sh.iter = ev.iter;          // the new hole's stream
    const double dE = ev.dE;
    double dE_cur = electron_receives_E(c, se, dE, ev.shell);
    // Update_electron_angles_SHI (:1170-1187) with the UPDATED ion energy
    double theta = (ev.E_after <= 0.0) ? TRK_PI / 2.0 : m_acos(sqrt((MSHI + TRK_ME) * (MSHI + TRK_ME) / (4.0 * MSHI * TRK_ME) * dE / ev.E_after));
    double phi; { double r2 = rn(p, se); phi = 2.0 * TRK_PI * r2; }
    // Impact_parameter (:1113-1126); the ion moves along the Z axis (X = Y = 0)
    double A = 1.0 + MSHI / TRK_ME;
    const double Zeff = shi_zeff(p, ev.E_after);                  // the equilibrium charge after the collision (:2196)
    double b = TRK_A0 * Zeff * TRK_RY / ev.E_after * sqrt(4.0 * ev.E_after / dE * MSHI / TRK_ME - A * A);
    double X = b * m_sin(phi), Y = b * m_cos(phi);
    emit_electron(c, se, id_e, dE_cur, ev.t0, X, Y, ev.Z, theta, phi, TRK3_ERR_20);
    emit_hole(c, sh, id_h, ev.shell, dE - dE_cur, ev.t0, X, Y, ev.Z, TRK3_ERR_20);
}
// the whole trajectory of the ion of one iteration (serial form, used by the CPU emulation)
template <class C>
TRK_HD void shi_history(C &c, uint32_t iter) {
    Rec s;
    shi_begin(c.p, s, iter);
    while (s.tn < c.p.Tim) { ShiEvent ev; shi_step(c, s, ev); shi_emit(c, ev); }
}

// ------------------------------------------------------------------------------------------------
// Histories: follow one particle from its record towards Tim, depositing snapshots on the way.
// `ig` = next grid index to snapshot (first i with t0 < tg(i)).
// ------------------------------------------------------------------------------------------------
// Queue routing.  Particles that can no longer ionise ("cold": below the lowest threshold of their total inelastic
// MFP table, or without any further collision before Tim) are handled by the cold kernels, whose code contains only
// snapshots + elastic scattering; everything else goes through the full ("hot") handlers.
TRK_HD bool electron_is_cold(const DevP &p, const Rec &e) { return e.E < p.e_cold || !(e.tn < p.Tim); }
TRK_HD bool vbhole_is_cold(const DevP &p, const Rec &h) { return h.Ehkin < p.h_cold || !(h.tn < p.Tim); }

TRK_HD bool electron_leaves_hot(const DevP &p, const Rec &e) { return e.E < p.e_warm && e.tn < p.Tim; }      // e_warm >= e_cold
TRK_HD bool vbhole_leaves_hot(const DevP &p, const Rec &h) { return h.Ehkin < p.h_warm && h.tn < p.Tim; }      // h_warm >= h_cold

enum StepStatus { ST_DONE = 0, ST_CONT = 1, ST_MOVE = 2, ST_MOVE_HOT = 3 };
// ST_DONE: history finished.  ST_CONT: call again.  ST_MOVE: hand the record to push() (it now belongs to the other
// temperature class or another species).  ST_MOVE_HOT: hand it to push_hot() (full handler required, see electron_event_t).

TRK_HD void begin_electron(const DevP &p, const Rec &e, int &ig, Cache &k) {
    cache_electron(p, e.E, k);
    ig = interval_of(p, e.t0);
}
// one step = snapshots spanned by the current free flight, then the collision at tn.  A record may only change
// queue in a state where the snapshots of (t0, tn] are still all to be taken: the consumer restarts from t0.
// `warm` (COLD only): the elastic-only handler also serves the "warm" electrons, which can still ionise but rarely do
// (E < DevP::e_warm: the channel roulette selects the ionisation with a small probability).  A warm electron whose
// roulette does select it is handed to the full handler with its stream rewound by that draw (ST_MOVE_HOT), so the
// result does not depend on which kernel followed it; one that has fallen below e_cold moves on to the cold queue.
template <bool COLD, class C>
TRK_HD int step_electron(C &c, Rec &e, int &ig, Cache &k, bool warm = false) {
    const DevP &p = c.p;
    double RN = 0.0;
    if (COLD && e.tn < p.tg[p.Nt - 1]) {                          // a collision is pending: is it really an elastic one?
        if (!(e.E < (warm ? p.e_warm : p.e_cold))) return ST_MOVE;
        event_begin(e);
        RN = rn(p, e);
        if (electron_roulette_inelastic(k, RN)) { e.ctr--; return ST_MOVE_HOT; }      // cold: probability ~1e-16 (IMFP >= 1e16)
    }
    while (ig <= p.Nt && p.tg[ig - 1] <= e.tn) { c.snap(SP_ELECTRON, e, ig); ++ig; }
    if (ig > p.Nt) return ST_DONE;
    if (COLD) {
        electron_event_t<EV_ELASTIC>(c, e, ig, k, RN);
        return (warm && (e.E < p.e_cold || !(e.tn < p.Tim))) ? ST_MOVE : ST_CONT;
    }
    event_begin(e);
    RN = rn(p, e);
    electron_event_t<EV_ANY>(c, e, ig, k, RN);
    return (e.E < p.e_cold && e.tn < p.Tim) ? ST_MOVE : ST_CONT;
}
// a valence hole taken from a queue: the lookups of its kinetic energy (only mobile holes ever collide)
TRK_HD void begin_vbhole(const DevP &p, const Rec &h, int &ig, Cache &k) {
    ig = interval_of(p, h.t0);
    if (h.tn < p.Tim) cache_vbhole(p, h.Ehkin, k);
}
// `warm` (COLD only): as for electrons (step_electron) -- valence holes with h_cold <= Ehkin < h_warm can ionise but rarely
// do; the elastic-only handler follows them and hands one over (stream rewound by the draw) when its roulette selects the
// ionisation.  (A cold hole never ionises: vbhole_roulette_inelastic needs HIMFP < 1e15.)
template <bool COLD, class C>
TRK_HD int step_vbhole(C &c, Rec &h, int &ig, Cache &k, bool warm = false) {
    const DevP &p = c.p;
    double RN = 0.0;
    if (COLD && h.tn < p.tg[p.Nt - 1]) {
        if (!(h.Ehkin < (warm ? p.h_warm : p.h_cold))) return ST_MOVE;
        if (warm) {
            event_begin(h);
            RN = rn(p, h);
            if (vbhole_roulette_inelastic(k, RN)) { h.ctr--; return ST_MOVE_HOT; }
        }
    }
    while (ig <= p.Nt && p.tg[ig - 1] <= h.tn) { c.snap(SP_VBHOLE, h, ig); ++ig; }
    if (ig > p.Nt) return ST_DONE;
    if (COLD) {
        if (!warm) { event_begin(h); RN = rn(p, h); }
        vbhole_event_t<EV_ELASTIC>(c, h, ig, k, RN);
        if (warm) return (h.Ehkin < p.h_cold || !(h.tn < p.Tim) || !(h.Ehkin < p.h_warm)) ? ST_MOVE : ST_CONT;
        // A hole that a collision has lifted out of the cold range (level snapping; absorption of lattice energy with DSF
        // scattering) goes back to the full handlers -- unless it has no collision left before Tim: then it only has snapshots to
        // deposit and stays here.  (Handed to push() it would be routed to the cold queue again, behind the records the running
        // cold launch covers, and its last snapshots would be lost.)
        return (h.Ehkin < p.h_cold || !(h.tn < p.Tim)) ? ST_CONT : ST_MOVE;
    }
    event_begin(h);
    vbhole_event_t<EV_ANY>(c, h, ig, k, rn(p, h));
    return (h.Ehkin < p.h_cold && h.tn < p.Tim) ? ST_MOVE : ST_CONT;
}
// core hole: after a decay the hole may have hopped into the valence band -> continue as a VB hole (other queue)
template <class C>
TRK_HD int step_corehole(C &c, Rec &h, int &ig) {
    const DevP &p = c.p;
    while (ig <= p.Nt && p.tg[ig - 1] <= h.tn) { c.snap(SP_VBHOLE, h, ig); ++ig; }
    if (ig > p.Nt) return ST_DONE;
    corehole_event(c, h);
    if (h.shell == p.vb_shell) { c.push(SP_VBHOLE, h); return ST_DONE; }
    return ST_CONT;
}
template <class C>
TRK_HD int step_photon(C &c, Rec &ph, int &ig) {
    const DevP &p = c.p;
    while (ig <= p.Nt && p.tg[ig - 1] <= ph.tn) { c.snap(SP_PHOTON, ph, ig); ++ig; }
    if (ig > p.Nt) return ST_DONE;
    photon_event(c, ph);
    return ST_DONE;
}

}  // namespace trk3
