/*
 * trekis3_gpu.h -- C ABI of the B200-native TREKIS-3 Monte-Carlo cascade engine.
 *
 * Drop-in boundary: the reference's only public symbol of MODULE Monte_Carlo is
 *     subroutine do_Monte_Carlo(NMC, SHI, SHI_MFP, diff_SHI_MFP, Target_atoms, ...)
 * (Source_files/Monte_Carlo.f90:39-44, sole caller Universal_MC_for_SHI_MAIN.f90:271-276).
 * Its 52 dummy arguments are Fortran derived types with allocatable components and are
 * therefore not C-interoperable; the structs below are the flattened (SoA, fp64) image of
 * those arguments, and trk3_mc_run() is the replacement of the call.  A Fortran
 * ISO_C_BINDING shim that flattens the derived types and calls these entry points is shown
 * in INTEGRATION.md.
 *
 * Plain C: pointers and sizes only, no C++/torch types.  All real data are double (the
 * reference is compiled with -real-size 64), all indices are 0-based in this header.
 */
#ifndef TREKIS3_GPU_H
#define TREKIS3_GPU_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TRK3_MAX_ATOMS   8
#define TRK3_MAX_SHELLS  32
#define TRK3_NR          50    /* radial bins, Sorting_output_data.f90:1373 */
#define TRK3_NTHETA      180   /* Out_theta1(180), Sorting_output_data.f90:1193 */
#define TRK3_MAX_NT      256   /* max number of output time-grid points */

/* ---------------------------------------------------------------------------------
 * Scalars and switches: image of SHI (Objects.f90:44-53), Matter (Objects.f90:86-105),
 * NumPar (Objects.f90:129-167), Tim, dt, NMC as passed to do_Monte_Carlo.
 * ------------------------------------------------------------------------------- */
typedef struct trk3_config {
    /* ion (type Ion) */
    double shi_E;          /* [eV]  SHI%E */
    double shi_mass;       /* [proton masses] SHI%Mass */
    double shi_fixed_Zeff; /* SHI%fixed_Zeff */
    int32_t shi_Z;         /* SHI%Zat */
    int32_t shi_kind_Zeff; /* 0 Barkas,1 Bohr,2 ND,3 SG,4 fixed (Cross_sections.f90:2657) */
    /* time grid (Monte_Carlo.f90:2118-2149) */
    double Tim;            /* [fs] */
    double dt;             /* [fs] step (linear) or factor (log) */
    int32_t dt_flag;       /* <=0 linear, >=1 logarithmic */
    int32_t include_photons; /* NumPar%include_photons */
    /* material scalars (type Solid) */
    double cut_off;        /* [eV] */
    double layer;          /* [A]  */
    double hole_mass;      /* [me]; <0 => effective mass from DOS (Monte_Carlo.f90:783-788) */
    double work_function;  /* [eV]; <=0 => no emission */
    double bar_length;     /* [A] */
    double bar_height;     /* [eV] */
    /* numerics */
    int32_t kind_of_EMFP;  /* 0 Mott, 1 CDF phonons (2 DSF unsupported) */
    int32_t reserved0;
    /* engine (not in the reference) */
    uint64_t seed;         /* Philox key; histories are keyed (seed, iteration, particle id) */
} trk3_config;

/* ---------------------------------------------------------------------------------
 * Tables: image of Target_atoms, SHI_MFP, diff_SHI_MFP, Total_el_MFPs, Elastic_MFP,
 * Total_Hole_MFPs, Elastic_Hole_MFP, Total_Photon_MFPs, aidCS, Mat_DOS, Out_R, Out_V.
 * Shells are flattened in (atom, shell) order; matrices are row-major [shell][energy].
 * Differential tables are CSR: row r occupies [off[r], off[r+1]).
 * All pointers are host pointers owned by the caller; trk3_mc_create copies them.
 * ------------------------------------------------------------------------------- */
typedef struct trk3_tables {
    int32_t n_atoms;
    int32_t n_shells;                 /* total over atoms */
    int32_t vb_shell;                 /* flat index of (Lowest_Ip_At, Lowest_Ip_Shl) */
    int32_t nshl_atom1;               /* size(Target_atoms(1)%Ip): 4th dim of Out_nh etc. */
    int32_t atom_Z[TRK3_MAX_ATOMS];
    int32_t atom_nshl[TRK3_MAX_ATOMS];
    int32_t atom_first[TRK3_MAX_ATOMS];   /* flat index of first shell of atom */
    double  atom_mass[TRK3_MAX_ATOMS];    /* [proton masses] */
    double  atom_pers[TRK3_MAX_ATOMS];    /* stoichiometry */
    int32_t shell_atom[TRK3_MAX_SHELLS];  /* 0-based atom of flat shell */
    int32_t shell_num[TRK3_MAX_SHELLS];   /* 0-based shell index inside its atom */
    double  shell_Ip[TRK3_MAX_SHELLS];
    double  shell_Nel[TRK3_MAX_SHELLS];
    double  shell_auger[TRK3_MAX_SHELLS];   /* [fs] */
    double  shell_radiat[TRK3_MAX_SHELLS];  /* [fs] */

    /* electron inelastic: Total_el_MFPs(:)%ELMFP(:)%{E,L} */
    int32_t n_ei;  const double *ei_E;  const double *ei_L;   /* [n_shells][n_ei] */
    /* electron elastic: Elastic_MFP%Total%{E,L} */
    int32_t n_ee;  const double *ee_E;  const double *ee_L;   /* [n_ee] */
    /* VB-hole inelastic / elastic */
    int32_t n_hi;  const double *hi_E;  const double *hi_L;   /* [n_shells][n_hi] */
    int32_t n_he;  const double *he_E;  const double *he_L;   /* [n_he] */
    /* photon (n_ph = 0 when photons are off) */
    int32_t n_ph;  const double *ph_E;  const double *ph_L;   /* [n_shells][n_ph] */
    /* SHI: SHI_MFP(:)%ELMFP(:)%{E,L,dEdx} */
    int32_t n_shi; const double *shi_E; const double *shi_L; const double *shi_dEdx;
    /* diff_SHI_MFP: per shell cumulative (E = transferred energy, L = MFP) */
    const int64_t *dshi_off;  /* [n_shells+1] */
    const double  *dshi_E;  const double *dshi_L;
    /* aidCS%EIdCS: rows (shell, iE); EEdCS rows iE; HIdCS; HEdCS */
    const int64_t *eid_off;   /* [n_shells*n_ei + 1] */
    const double  *eid_hw;  const double *eid_L;
    const int64_t *eed_off;   /* [n_ee + 1] */
    const double  *eed_hw;  const double *eed_L;
    const int64_t *hid_off;   /* [n_hi + 1] */
    const double  *hid_hw;  const double *hid_L;
    const int64_t *hed_off;   /* [n_he + 1] */
    const double  *hed_hw;  const double *hed_L;
    /* Mat_DOS (hole-energy grid, increasing) */
    int32_t n_dos; const double *dos_E; const double *dos_DOS; const double *dos_int; const double *dos_effm;
    /* Out_R / Out_V (Sorting_output_data.f90:1403-1434) */
    int32_t n_r;   const double *out_R; const double *out_V;
    /* BEB shells (negative shell designator in the .cdf: Target_atoms%KOCS = 2, Reading_files_and_parameters.f90:1557-1561):
     * electrons and valence holes ionise them with the binary-encounter-Bethe cross section (Cross_sections.f90:3891-3906);
     * the transferred energy is then sampled from its closed form by bisection (Electron_NRG_transfer_BEB, :2128-2165)
     * instead of a differential table.  shell_kocs: 1 CDF, 2 BEB; shell_Ek = mean kinetic energy of the shell (EADL I = 914);
     * at_dens = Matter%At_Dens [1/cm^3]. */
    int32_t shell_kocs[TRK3_MAX_SHELLS];
    double  shell_Ek[TRK3_MAX_SHELLS];
    double  at_dens;
    /* Delta-function CDF (kind_of_DR = 4, Cross_sections.f90:952-956, 1449-1786): the inelastic cross section of electrons and
     * valence holes has a closed form per CDF oscillator (position E0, weight alpha = define_alpha,
     * Reading_files_and_parameters.f90:2199-2206) and the transferred energy is sampled from it by bisection
     * (get_inelastic_energy_transfer, :2051-2123) instead of a differential table.  delta_cdf: 0 off, 1 on; the oscillators of flat
     * shell s are entries [osc_off[s], osc_off[s+1]) of osc_E0 / osc_alpha. */
    int32_t delta_cdf;
    int32_t osc_off[TRK3_MAX_SHELLS + 1];
    const double *osc_E0; const double *osc_alpha;
    /* DSF elastic scattering (kind_of_EMFP = 2; DSF_DEMFP / DSF_DEMFP_H of do_Monte_Carlo, NRG_transfer_elastic_DSF,
     * Cross_sections.f90:3652-3780).  The elastic tables ee_E / ee_L (he_E / he_L) then live on the DSF file's own particle-energy
     * grid; per grid energy i the row [i * n_dsf_e, (i + 1) * n_dsf_e) of dsf_e_dE holds the transferred energies (ascending, from
     * -0.2 eV: the particle absorbs energy from the lattice, to +0.2 eV: it emits) and the same row of dsf_e_emit / dsf_e_absorb the
     * mean free paths integrated up to them; ee_emit / ee_absorb = Elastic_MFP%Emit%L / %Absorb%L (the rows' last entries,
     * Analytical_IMFPs.f90:913-919).  n_dsf_e = 0: no DSF tables. */
    int32_t n_dsf_e; const double *dsf_e_dE; const double *dsf_e_emit; const double *dsf_e_absorb; const double *ee_emit; const double *ee_absorb;
    int32_t n_dsf_h; const double *dsf_h_dE; const double *dsf_h_emit; const double *dsf_h_absorb; const double *he_emit; const double *he_absorb;
} trk3_tables;

/* ---------------------------------------------------------------------------------
 * Tallies: one contiguous fp64 buffer; every Out_* array of do_Monte_Carlo lives at a
 * fixed offset in Fortran (column-major) element order, so a Fortran caller can
 * c_f_pointer straight onto it.  The engine ADDS per-iteration contributions (the
 * caller pre-zeroes and later divides by NMC, MAIN.f90:281-314).
 * ------------------------------------------------------------------------------- */
enum trk3_tally_id {
    TRK3_OUT_NE = 0,       /* (Nt,NR)            */
    TRK3_OUT_EE,           /* (Nt,NR)            */
    TRK3_OUT_NPHOT,        /* (Nt,NR)            */
    TRK3_OUT_EPHOT,        /* (Nt,NR)            */
    TRK3_OUT_EE_VS_E,      /* (Nt,NR)            */
    TRK3_OUT_EH_VS_E,      /* (Nt,Ndos)          */
    TRK3_OUT_ELAT,         /* (Nt,NR)            */
    TRK3_OUT_NH,           /* (Nt,NR,Nat,Nsh1)   */
    TRK3_OUT_EH,           /* (Nt,NR,Nat,Nsh1)   */
    TRK3_OUT_EHKIN,        /* (Nt,NR,Nat,Nsh1)   */
    TRK3_OUT_TOT_NE,       /* (Nt)               */
    TRK3_OUT_TOT_NPHOT,    /* (Nt)               */
    TRK3_OUT_TOT_E,        /* (Nt)               */
    TRK3_OUT_E_E,          /* (Nt)               */
    TRK3_OUT_E_PHOT,       /* (Nt)               */
    TRK3_OUT_E_AT,         /* (Nt)               */
    TRK3_OUT_E_H,          /* (Nt,Nat,Nsh1)      */
    TRK3_OUT_EAT_DENS,     /* (Nt,NR) never written by the reference */
    TRK3_OUT_THETA,        /* (Nt+1,180)         */
    TRK3_OUT_THETA_H,      /* (Nt+1,180)         */
    TRK3_OUT_NE_EM,        /* (Nt)               */
    TRK3_OUT_E_EM,         /* (Nt)               */
    TRK3_OUT_EE_VS_E_EM,   /* (Nt,NR)            */
    TRK3_OUT_FIELD_ALL,    /* (Nt,NR) dead code in the reference, stays 0 */
    TRK3_OUT_E_FIELD,      /* (Nt)    dead code in the reference, stays 0 */
    TRK3_OUT_DIFF_COEFF,   /* (Nt)               */
    TRK3_N_TALLIES
};

typedef struct trk3_tally_layout {
    int32_t Nt, n_r, n_atoms, nshl1, n_dos, reserved;
    int64_t off[TRK3_N_TALLIES];   /* element offset of each array */
    int64_t len[TRK3_N_TALLIES];   /* element count of each array  */
    int64_t total;                 /* doubles in the whole buffer  */
    double  time_grid[TRK3_MAX_NT + 1]; /* set_time_grid, Monte_Carlo.f90:2118 (Nt+1 entries) */
} trk3_tally_layout;

/* Event classes: one pass through `select case (KOP)`,
 * Monte_Carlo.f90:587-627, split by channel. */
enum trk3_event_class {
    TRK3_EV_SHI = 0, TRK3_EV_EL_INEL, TRK3_EV_EL_ELAST, TRK3_EV_VBH_INEL, TRK3_EV_VBH_ELAST,
    TRK3_EV_AUGER, TRK3_EV_RADIATIVE, TRK3_EV_AUGER_FROZEN, TRK3_EV_PHOTON, TRK3_N_EVENT_CLASSES
};

/* Error counters carry the reference's error numbers (Monte_Carlo.f90:2227-2960). */
enum trk3_error_code {
    TRK3_ERR_10 = 0, TRK3_ERR_20, TRK3_ERR_21, TRK3_ERR_22, TRK3_ERR_23, TRK3_ERR_25, TRK3_ERR_30,
    TRK3_ERR_40, TRK3_ERR_41, TRK3_ERR_50, TRK3_ERR_51, TRK3_ERR_52,
    TRK3_ERR_QUEUE_OVERFLOW, TRK3_ERR_NAN, TRK3_ERR_AUGER_BALANCE, TRK3_N_ERRORS
};

typedef struct trk3_stats {
    uint64_t events[TRK3_N_EVENT_CLASSES];
    uint64_t errors[TRK3_N_ERRORS];
    uint64_t n_electrons;       /* total electrons created */
    uint64_t n_photons;         /* total photons created   */
    uint64_t n_waves;           /* kernel generations launched */
    uint64_t kernel_launches;   /* CUDA kernels launched by the engine */
    double   device_ms;         /* CUDA-event time of the MC section */
    double   algorithmic_bytes; /* sum_class events*bytes */
    double   max_energy_drift;  /* max_it max_i |tot_E(it,i)-tot_E(it,Nt)|/tot_E(it,Nt), i >= first grid time after the ion left */
    uint64_t cold_events[2];    /* elastic collisions of electrons / valence holes handled by the cold (elastic-only) kernels */
    uint64_t warm_events[2];    /* elastic collisions of "warm" electrons / valence holes (can still ionise, rarely do) handled by the same kernels, generation by generation */
} trk3_stats;

/* Return codes */
#define TRK3_OK                 0
#define TRK3_E_INVALID         -1
#define TRK3_E_CUDA            -2
#define TRK3_E_UNSUPPORTED     -3
#define TRK3_E_NOMEM           -4
#define TRK3_E_OVERFLOW        -5

typedef struct trk3_engine trk3_engine;   /* opaque; owns all device memory */

/* Fill `lay` for a given configuration/tables (host only, no GPU needed).
 * Mirrors Radius_for_distributions / Allocate_out_arrays / set_time_grid. */
int trk3_tally_layout_init(const trk3_config *cfg, const trk3_tables *tab, trk3_tally_layout *lay);

/* Validate + upload tables once (replaces the per-iteration How_many_electrons
 * table flattening, Monte_Carlo.f90:1902-2057).  device < 0 => current device. */
int trk3_mc_create(const trk3_config *cfg, const trk3_tables *tab, int device, trk3_engine **out);

/* Copy configuration + tables into an existing engine again (same shapes as at creation): what a persistent plugin
 * handle does at every do_Monte_Carlo call -- the inputs go host->device, queues and scratch stay allocated.
 * The arrays are copied into a pinned mirror owned by the handle and leave in two DMAs on the engine's stream; the call
 * returns with them in flight: `cfg` and `tab` (and the arrays they point to) may be changed or freed as soon as it has
 * returned, and the next trk3_mc_run / trk3_mc_run_device is ordered behind the copies (option "stage_uploads" = 0:
 * array-by-array copies, synchronised before the call returns).
 * trk3_mc_table_bytes = bytes copied host->device by the last binding. */
int trk3_mc_reload_tables(trk3_engine *eng, const trk3_config *cfg, const trk3_tables *tab);
uint64_t trk3_mc_table_bytes(const trk3_engine *eng);

/* Run iterations [it_begin, it_end) (global 0-based iteration indices; RNG streams are keyed
 * by the global index so the union over ranks is independent of the split) and ADD their
 * contributions to `tallies` (host buffer of lay.total doubles, caller pre-zeroed).
 * Replaces do_Monte_Carlo (Monte_Carlo.f90:39).  `stats` may be NULL. */
int trk3_mc_run(trk3_engine *eng, int64_t it_begin, int64_t it_end, double *tallies, trk3_stats *stats);

/* Same, but leaves the tally sums in the engine's device buffer (for an NCCL all-reduce
 * issued by the caller on trk3_mc_device_tallies()).  */
int trk3_mc_run_device(trk3_engine *eng, int64_t it_begin, int64_t it_end, trk3_stats *stats);
double *trk3_mc_device_tallies(trk3_engine *eng);      /* device pointer, lay.total doubles */
int trk3_mc_zero_device_tallies(trk3_engine *eng);
int trk3_mc_download_tallies(trk3_engine *eng, double *tallies_add_into);

/* Per-iteration total energy at every grid time, [n_iter][Nt] row-major, for the
 * conservation check (Total_numbers.txt, manual section VI).  Valid after a run. */
int trk3_mc_iteration_energies(trk3_engine *eng, double *out, int64_t capacity_doubles, int64_t *n_iter);

/* Plumbing for frameworks that own the stream / the tally buffer (torch.distributed + NCCL all-reduce). */
int trk3_mc_set_stream(trk3_engine *eng, void *cuda_stream);
int trk3_mc_set_device_tallies(trk3_engine *eng, double *device_buffer);
/* Device time per kernel class since option "profile"=1 (0 hot electron wave, 1 hot valence-hole wave, 2 core-hole
 * wave, 3 photon wave, 4 ion tracks, 5 pair creation + snapshots + folding, 6 cold electrons, 7 cold valence holes, 8 warm electrons, 9 warm valence holes); returns the number
 * of classes.  Classes 0-3 and 8 of one generation run on concurrent streams: their times overlap. */
int trk3_mc_kernel_times(trk3_engine *eng, double *ms, uint64_t *launches, int n);

/* ---------------------------------------------------------------------------------------------------------------
 * Multi-GPU: the role of MPI_subroutines.f90 on this path.  The reference splits the iterations over the MPI ranks
 * (Monte_Carlo.f90:111-129) and then sums the 26 Out_* arrays with 26 MPI_Reduce calls (:131-389).  Here every rank
 * (one process per GPU) runs its own range of GLOBAL iteration indices, and -- once a communicator is attached -- every
 * trk3_mc_run / trk3_mc_run_device call is COLLECTIVE: it ends with ONE ncclAllReduce(ncclDouble, ncclSum) of the packed
 * tally buffer, issued on the engine's own stream right behind the folding kernels, and every rank returns the reduced
 * tallies (the reference leaves them on rank 0 only).  The per-rank statistics (trk3_stats) stay local.
 *   trk3_nccl_unique_id   rank 0 creates the 128-byte NCCL id; the caller distributes it (MPI_Bcast, a file, torch ...)
 *   trk3_mc_comm_init     ncclCommInitRank on the engine's device; the engine owns (and destroys) the communicator
 *   trk3_mc_set_comm      attach a caller-owned ncclComm_t instead (NULL detaches)
 * NCCL is loaded at run time (dlopen "libnccl.so.2": the copy already mapped into the process, e.g. PyTorch's, or the
 * system's), so the library itself has no link-time dependency on it.  If a rank fails before its all-reduce the other
 * ranks block in theirs: the caller aborts the job, as the reference's MPI_Abort does (MPI_subroutines.f90:177-188). */
#define TRK3_NCCL_UNIQUE_ID_BYTES 128
int trk3_nccl_unique_id(void *id_out_128_bytes);
int trk3_mc_comm_init(trk3_engine *eng, int nranks, int rank, const void *id_128_bytes);
int trk3_mc_set_comm(trk3_engine *eng, void *nccl_comm, int nranks);
int trk3_mc_comm_size(const trk3_engine *eng);          /* ranks of the attached communicator, 1 if none */

/* Bring an engine back to a defined state after a failed run (CUDA error, persistent overflow): waits for its streams,
 * clears the device queues, counters and tallies and the error text.  Tables and options stay.  Returns TRK3_E_CUDA if the
 * device itself is lost (sticky CUDA error): then only trk3_mc_destroy is left. */
int trk3_mc_reset(trk3_engine *eng);

/* Tunables: "batch" iterations in flight, "cap_factor" queue capacity, "use_smem", "refill_min", "profile" ... */
int trk3_mc_set_option(trk3_engine *eng, const char *name, double value);

const trk3_tally_layout *trk3_mc_layout(const trk3_engine *eng);
const char *trk3_mc_last_error(const trk3_engine *eng);
void trk3_mc_destroy(trk3_engine *eng);

/* ---------------------------------------------------------------------------------------------------------------
 * GPU evaluator of the table builder's integrands.  The reference builds its mean-free-path and
 * differential cross-section tables by nested Simpson integrations (TotIMFP Cross_sections.f90:881-1050, Tot_EMFP
 * :2966-3139, SHI_TotIMFP :2452-2597): an outer loop over the transferred energy hw whose every step needs the
 * q-integral of the loss function at two new points (Diff_cross_section :2217, SHI_Diff_cross_section :2683,
 * Diff_cross_section_phonon :3142).  The outer loops' control flow does not depend on the integrand, so the host builder
 * (libtrekis3_host.so) first records every (task, hw) it will ask for, has them all evaluated at once -- by this entry
 * point, one GPU thread per request -- and then replays the outer loops with the values: same operations in the same
 * order, identical tables.  A "task" is one outer integration: all its requests share these parameters. */
enum { TRK3_DCS_INELASTIC = 0, TRK3_DCS_PHONON = 1, TRK3_DCS_SHI = 2, TRK3_DCS_SHI_BK = 3 };
typedef struct trk3_dcs_task {
    int32_t type;          /* TRK3_DCS_* */
    int32_t set;           /* oscillator set (CDF shell; the phonon CDF is the last set) */
    double Ee;             /* energy of the incident particle [eV] */
    double Mass;           /* its mass [m_e] (electrons 1, holes from the DOS); SHI_BK: atomic number of the ion */
    double p1, p2, p3;     /* PHONON: mean target atom mass [kg], target temperature [K], pref;  SHI: ion mass [kg], Emax [eV] (SHI_BK: and the equilibrium charge) */
} trk3_dcs_task;
typedef struct trk3_dcs_ctx {
    const double *osc_E0, *osc_A, *osc_G;   /* oscillators of all sets, concatenated (type CDF, Objects.f90:211-217) */
    const int32_t *osc_off;                 /* set s = [osc_off[s], osc_off[s+1]) */
    int32_t n_sets;
    const double *k, *effm; int32_t n_k;    /* k-vector and effective mass on the DOS grid (Objects.f90:111-123) */
    int32_t mass_from_dos;                  /* El_eff_mass == 0: mass from the DOS (Imewq :428-434) */
    double El_eff_mass;
    int32_t kind_DR;                        /* dispersion relation of the oscillators */
    double v_f, temp;
    /* Dynamical screening of the nucleus in the elastic (phonon-CDF) cross section, CDF_elast_Zeff = 2 / 3
     * (Diff_cross_section_phonon, Cross_sections.f90:3216-3276): 0 = none (CDF_elast_Zeff 0 / 1), 2 = atomic form factors for
     * the core + CDF of the valence band, 3 = CDF of every shell.  `scr` packs what get_screening_ff / get_screening_all
     * read: [0] n_atoms, then per atom 10 doubles {Zat, Pers, electrons outside the valence band, first oscillator set,
     * number of shells, form-factor coefficients a1..a5 (INPUT_EADL/Atomic_form_factors.dat)}, then per oscillator set
     * (shell) 2 doubles {Nel, Ip}.  vb_set = oscillator set of the valence band (last shell of the first atom). */
    int32_t screening, vb_set;
    const double *scr; int32_t n_scr;
} trk3_dcs_ctx;
/* out[i] = the q-integral of request i = (tasks[task_of[i]], hw[i]).  All pointers are host memory.  Returns TRK3_OK or
 * a negative error (no CUDA device: there is no CPU fallback in this library). */
int trk3_dcs_eval(const trk3_dcs_ctx *ctx, const trk3_dcs_task *tasks, int64_t n_tasks,
                  const double *hw, const int32_t *task_of, int64_t n, double *out);
/* device time [ms] and number of requests of the calls since the last reset (reset != 0 clears them) */
int trk3_dcs_stats(double *device_ms, int64_t *requests, int reset);

/* Library identity, used by tests to prove the CUDA library (not a fallback) is loaded. */
const char *trk3_gpu_version(void);

#ifdef __cplusplus
}
#endif
#endif /* TREKIS3_GPU_H */
