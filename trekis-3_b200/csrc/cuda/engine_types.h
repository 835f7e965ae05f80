// engine_types.h -- device-side data model of the wavefront engine (shared by the CUDA build and the
// host emulation build used by the CPU tests).
//
// Layout decisions (DESIGN.md "Data layout in HBM"):
//  * tables: one flat fp64 image per family on a shared energy grid, rows [shell][energy]; totals
//    (How_many_electrons, Monte_Carlo.f90:1902) are precomputed once at upload; differential tables are CSR.
//  * particles: struct-of-arrays queues, one per species (electron, valence hole, core hole, photon), double
//    buffered per wavefront generation; a record is consumed exactly once and run to its end of history.
//  * tallies: the packed Out_* buffer of include/trekis3_gpu.h in global memory; the radial x time arrays
//    that are written from inside the event loop are privatised per block in shared memory and flushed once.
//  * per-iteration scratch: integer histograms for the three spectra that the reference normalises per
//    iteration (1/Tot_Nel, 1/N_VB_h_tot, Monte_Carlo.f90:1026-1069) and per-iteration energy sums.
#pragma once
#include <stdint.h>
#include "../../../include/trekis3_gpu.h"

#if defined(__CUDACC__)
#define TRK_HD __host__ __device__ inline
#define TRK_D __device__ inline
// (Keeping the large helpers -- Philox rounds, table sampling -- out of line was measured on B200: 36.3 ms instead of
// 31.0 ms per 1000 iterations of config C2; the call/stack traffic costs more than the smaller code gains.)
#define TRK_HDN __host__ __device__ inline
// (the same holds for the code of switches that are off in a run -- Mott scattering, electron emission: out of line it made
// the cold kernels 7-10 % slower)
#define TRK_HD_RARE __host__ __device__ inline
// one out-of-line copy: code that only some inputs ever reach and that only the hot kernels (latency-, not fetch-bound) contain
#define TRK_HD_OUTLINE __host__ __device__ __noinline__ inline
#else
#define TRK_HD inline
#define TRK_D inline
#define TRK_HDN inline
#define TRK_HD_RARE inline
#define TRK_HD_OUTLINE inline
#endif

namespace trk3 {

enum Species { SP_ELECTRON = 0, SP_VBHOLE = 1, SP_COREHOLE = 2, SP_PHOTON = 3, N_SPECIES = 4 };

// One particle (registers while in flight; SoA columns in a queue).  Objects.f90:31-68 minus the
// never-used velocities/accelerations, plus the Philox stream (id, ctr) and the iteration index.
struct Rec {
    double E, Ehkin, Mass, t0, tn, X, Y, Z, L, theta, phi;
    uint64_t id;      // Philox stream id of this particle
    uint32_t ctr;     // next draw index of the stream
    uint32_t iter;    // global iteration index (4th Philox counter word)
    int32_t shell;    // flat shell of a hole (unused for electrons/photons)
    // second half of the Philox block of the last even draw (registers only, never stored in a queue)
    uint32_t rc_a, rc_b, rc_blk = 0xffffffffu;
};

#define TRK_NCOL 11   // double columns of a queue

struct Queue {        // SoA particle queue in global memory
    double *col[TRK_NCOL];
    uint64_t *id; uint32_t *ctr; uint32_t *iter; int32_t *shell;
    uint32_t *count;  // number of records pushed (may exceed capacity on overflow; clamp when reading)
    uint32_t cap;
};

// Per-iteration scratch arrays (index it_local = iter - batch_begin)
struct IterArrays {
    uint32_t *created;   // [nb][Nt+2]  electrons created in time interval iv (1..Nt+1)
    uint32_t *nvb;       // [nb][Nt]    valence holes alive at grid time i
    uint32_t *nph;       // [nb][Nt]    photons alive at grid time i
    uint32_t *spec_e;    // [nb][Nt][n_r]
    uint32_t *spec_h;    // [nb][Nt][n_dos]
    uint32_t *th_e;      // [nb][Nt][180]
    uint32_t *th_h;      // [nb][Nt][180]
    uint32_t *diffN;     // [nb][Nt]
    double *diffS;       // [nb][Nt]
    double *esnap;       // [nb][Nt]    sum of particle energies at grid time i
    double *elat;        // [nb][Nt+2]  lattice energy deposited in interval iv
    uint32_t *em_cnt;    // [nb][Nt+2]  emitted electrons per interval
    double *em_E;        // [nb][Nt+2]
    uint32_t *em_spec;   // [nb][Nt+2][n_r]
};

// Direct-index accelerator of the searches in the (strictly increasing) energy grids: TRK_NLUT uniform bins in log(E),
// lut[b] = a grid index near the lower edge of bin b, corrected by a local scan (physics.cuh, find_lut).  The result
// is the index the reference's bisection finds; it replaces ~9 dependent loads by ~3.
#define TRK_NLUT 2048
struct GridLut { const uint16_t *lut; double l0, scale; };
enum { LUT_EI = 0, LUT_EE, LUT_HI, LUT_HE, LUT_PH, LUT_SHI, N_LUT };

struct DevP {
    // ---- scalars (trk3_config)
    double ion_E, ion_mass, ion_fixed_Zeff, ion_Zeff0;
    double Tim, cut_off, layer, hole_mass, work_function, bar_height, Em_E1, Em_gamma;
    double Egap, Mtarget, sum_pers;
    double Erest_target;                 // rest_energy(Mtarget) [eV], evaluated once (angles_lattice needs it at every elastic collision)
    int32_t ion_Z, ion_kind_Zeff, include_photons, kind_of_EMFP;
    uint32_t seed_lo, seed_hi;
    // ---- target (trk3_tables header)
    int32_t n_atoms, n_shells, vb_shell, nshl_atom1;
    int32_t atom_Z[TRK3_MAX_ATOMS], atom_first[TRK3_MAX_ATOMS], atom_nshl[TRK3_MAX_ATOMS];
    double atom_mass[TRK3_MAX_ATOMS], atom_pers[TRK3_MAX_ATOMS];
    int32_t shell_atom[TRK3_MAX_SHELLS], shell_num[TRK3_MAX_SHELLS];
    double shell_Ip[TRK3_MAX_SHELLS], shell_Nel[TRK3_MAX_SHELLS], shell_auger[TRK3_MAX_SHELLS], shell_radiat[TRK3_MAX_SHELLS];
    int32_t shell_kocs[TRK3_MAX_SHELLS];                 // 1: CDF shell, 2: BEB shell (electrons and holes; beb_transfer, physics.cuh)
    double shell_Ek[TRK3_MAX_SHELLS], at_dens;           // mean kinetic energy of the shell [eV], atomic density [1/cm^3] (BEB, delta-CDF)
    int32_t delta_cdf, osc_off[TRK3_MAX_SHELLS + 1];     // delta-function CDF (kind_of_DR = 4): oscillators [osc_off[s], osc_off[s+1]) of shell s
    const double *osc_E0, *osc_alpha;                    // (delta_transfer, physics.cuh)
    // DSF elastic scattering (kind_of_EMFP = 2): rows [i * n_dsf][0..n_dsf) per energy of the elastic grid (dsf_elastic_dE, physics.cuh)
    int32_t n_dsf_e, n_dsf_h;
    const double *dsf_e_dE, *dsf_e_emit, *dsf_e_absorb, *ee_emit, *ee_absorb, *dsf_h_dE, *dsf_h_emit, *dsf_h_absorb, *he_emit, *he_absorb;
    // ---- tables (device pointers).  Every table has a companion of natural logarithms (prefix l) computed once
    //      at upload with the same log() the kernels use, so that the log-log interpolation of the reference
    //      (Interpolate(5,...), Cross_sections.f90:4074-4081) costs one exp() instead of five log() + exp().
    int32_t n_ei, n_ee, n_hi, n_he, n_ph, n_shi, n_dos, n_r;
    const double *ei_E, *ei_L, *ei_tot, *ee_E, *ee_L, *hi_E, *hi_L, *hi_tot, *he_E, *he_L, *ph_E, *ph_L, *ph_tot;
    const double *shi_E, *shi_L, *shi_tot;
    const double *lei_E, *lei_L, *lei_tot, *lee_E, *lee_L, *lhi_E, *lhi_L, *lhi_tot, *lhe_E, *lhe_L, *lph_E, *lph_L, *lph_tot;
    const double *lshi_E, *lshi_L, *lshi_tot;
    const int64_t *dshi_off, *eid_off, *eed_off, *hid_off, *hed_off;
    const double *dshi_E, *dshi_L, *eid_hw, *eid_L, *eed_hw, *eed_L, *hid_hw, *hid_L, *hed_hw, *hed_L;
    const double *ldshi_E, *dshi_iL, *ldshi_iL;          // log(E), 1/L and log(1/L) of the SHI cumulative tables
    const double *leid_hw, *leid_L, *leed_hw, *leed_L, *lhid_hw, *lhid_L, *lhed_hw, *lhed_L;
    // per row of the electron differential tables: 1 = the cumulative MFPs of the row are non-increasing, so that the index
    // Find_in_monoton_array_decreasing ends on is clamp(#entries above the value, 1, n) (engine.cu, transferred_energy_warp)
    const uint8_t *eid_mono, *eed_mono;
    const double *dos_E, *dos_DOS, *dos_int, *dos_effm, *out_R, *out_V;
    int32_t shi_Mtemp[TRK3_MAX_SHELLS]; double shi_dL[TRK3_MAX_SHELLS];      // per-shell constants of SHI_energy_transfer
    GridLut dshi_lut[TRK3_MAX_SHELLS];                   // accelerators of the inverse-CDF search of SHI_energy_transfer (in log 1/L)
    double ion_pow23;                                    // Zion^0.66666666 (Equilibrium_charge_SHI, evaluated once)
    GridLut lut[N_LUT];                                  // search accelerators of the six energy grids
    double dos_inv_step;                                 // 1/step of the DOS energy grid if it is uniform, else 0
    // below these energies the total inelastic MFP is one constant >= 1e16 (no ionisation possible): lookup skipped
    double e_cold, e_imfp_cold, h_cold, h_imfp_cold;
    double e_iimfp_cold, h_iimfp_cold;                   // 1 / the two constants
    // hot electrons are queued by energy class (class c: e_class[c-1] <= E < e_class[c]; +inf = class not used): the number
    // of collisions an electron still has before it turns cold grows with its energy, and the engine gives the long
    // histories warps of their own (engine.cu, k_hot)
    double e_class[3];
    // "warm" electrons, e_cold <= E < e_warm, can still ionise but rarely do; the engine follows them with the elastic-only
    // kernel, generation by generation (e_warm = e_cold: no such class)
    double e_warm;
    double h_warm;                                       // the same for valence holes (kinetic energy)
    // ---- time grid: tg[i-1] = min(time_grid(i), Tim), i = 1..Nt
    int32_t Nt;
    double tg[TRK3_MAX_NT];
    // ---- tallies
    double *tally;                       // packed Out_* buffer (global)
    int64_t g_off[TRK3_N_TALLIES];       // offsets in `tally`
    int32_t s_off[TRK3_N_TALLIES];       // offsets in the block-private copy, -1 = not privatised
    int32_t s_len[TRK3_N_TALLIES];       // lengths of the privatised arrays
    int32_t s_total;                     // doubles in the block-private copy (0 = privatisation off)
    // ---- per-iteration scratch
    IterArrays it;
    uint32_t batch_begin, batch_n;       // global index of the first iteration of the batch, iterations in it
    int32_t defer_snap;                  // 1: the wave kernels queue their snapshots for k_snapshot instead of tallying them
    // ---- counters
    unsigned long long *events;          // [TRK3_N_EVENT_CLASSES]
    unsigned long long *errors;          // [TRK3_N_ERRORS]
    unsigned long long *cnt_el, *cnt_ph;
};

}  // namespace trk3
