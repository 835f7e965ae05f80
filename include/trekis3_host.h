/*
 * trekis3_host.h -- C ABI of the host half of the drop-in (libtrekis3_host.so, no CUDA).
 *
 * It replaces what the Fortran host does around do_Monte_Carlo:
 *   Read_input_file            Reading_files_and_parameters.f90:162   -> trk3h_load
 *   Analytical_ion_dEdx / Analytical_electron_dEdx / SHI_TotIMFP
 *                              Universal_MC_for_SHI_MAIN.f90:146-247  -> trk3h_build_tables
 *   Save_output                Sorting_output_data.f90:340            -> trk3h_save_output
 * and hands the flattened tables (trk3_config / trk3_tables of trekis3_gpu.h) to the engine.
 */
#ifndef TREKIS3_HOST_H
#define TREKIS3_HOST_H
#include "trekis3_gpu.h"
#ifdef __cplusplus
extern "C" {
#endif

typedef struct trk3h_case trk3h_case;

/* Parse <dir>/INPUT_PARAMETERS.txt, <dir>/INPUT_CDF/<material>.cdf, <dir>/INPUT_DOS/<material>.dos,
 * <dir>/INPUT_EADL/INPUT_atomic_data.dat [, <dir>/INPUT_EADL/radiative_widths.dat]. NULL on error. */
trk3h_case *trk3h_load(const char *dir, char *err, int errlen);
void trk3h_free(trk3h_case *c);

/* Build all MFP / differential cross-section tables on the host (OpenMP over grid points).
 * threads<=0: all cores.  shi_window_only!=0: only the SHI-grid points the ion can visit. */
int trk3h_build_tables(trk3h_case *c, int threads, int shi_window_only, int verbose, char *err, int errlen);
/* Evaluator of the integrands for the NEXT trk3h_build_tables calls (see trk3_dcs_eval in trekis3_gpu.h, whose address
 * is what a caller passes here; trk3h_dcs_eval_host is the same interface evaluated by the host threads, used by the CPU
 * tests of the record / replay machinery).  NULL: the builder integrates directly, point by point, on the host. */
typedef int (*trk3_dcs_eval_fn)(const trk3_dcs_ctx *ctx, const trk3_dcs_task *tasks, int64_t n_tasks,
                                const double *hw, const int32_t *task_of, int64_t n, double *out);
void trk3h_set_dcs_evaluator(trk3_dcs_eval_fn fn);
int trk3h_dcs_eval_host(const trk3_dcs_ctx *ctx, const trk3_dcs_task *tasks, int64_t n_tasks,
                        const double *hw, const int32_t *task_of, int64_t n, double *out);
int trk3h_save_tables(trk3h_case *c, const char *path, char *err, int errlen);
int trk3h_load_tables(trk3h_case *c, const char *path, char *err, int errlen);

/* The reference's own on-disk table cache (Analytical_IMFPs.f90:262-786, 917-1606, 2306-2349, 2462-2506, 2657-2707;
 * reader of the ion files Reading_files_and_parameters.f90:2855-2903): <out_root>/OUTPUT_<material>/OUTPUT_*_IMFPs_*.dat,
 * *_EMFPs_*.dat, diff_CS/<one file per shell and grid energy>, OUTPUT_<ion>_in_<material>/OUTPUT_<ion>_*_{IMFP,dEdx,
 * effective_charges,Range}.dat, same names, row layout and number formats ('(f)', '(e)', '(es)' of real(8) = F25.16,
 * E25.16, ES25.16, which the reference reads back as fixed 25-character fields).
 * write: the built tables of the case (e.g. built on the GPU) -> files the Fortran program accepts as its cache.
 * read:  tables cached by the reference (or by write) -> the case, instead of trk3h_build_tables; fails (and leaves the case
 *        without tables) if a file is missing or its row count differs from the energy grid, as the reference decides to redo.
 * name:  one component of the layout: dir_material, dir_ion, dir_diff, el_imfp, hole_imfp, photon_imfp, el_emfp, hole_emfp, shi_stem */
int trk3h_write_reference_cache(trk3h_case *c, const char *out_root, int *n_files, char *err, int errlen);
int trk3h_read_reference_cache(trk3h_case *c, const char *out_root, int threads, char *err, int errlen);
int trk3h_reference_cache_name(trk3h_case *c, const char *which, char *out, int outlen);

/* Flattened views; valid until the next trk3h_* call that modifies the case. */
const trk3_config *trk3h_config(trk3h_case *c);
const trk3_tables *trk3h_tables(trk3h_case *c);

/* Scalar overrides applied before packing (benchmarks, tests). */
int trk3h_set(trk3h_case *c, const char *key, double value);
double trk3h_get(trk3h_case *c, const char *key);
int trk3h_get_string(trk3h_case *c, const char *key, char *out, int outlen);
/* "atom:<j>:<k>:<field>" in trk3h_get: parameter of shell k of atom j (0-based): Nel, Ip, Ek, Auger, Radiat, Shl_num, PQN.
 * When <dir>/INPUT_EADL/EADL2023.ALL exists, trk3h_load takes from it what the .cdf leaves out, as check_atomic_parameters
 * does (Dealing_with_EADL.f90:312-413): Nel (I=912), Ip (913), Ek (914), radiative (921) and Auger (922) widths -> times.
 * trk3h_eadl_lookup: one lookup with the reader's sub-shell rules (READ_EADL_TYPE_FILE_int/_real :417-684); value in eV. */
int trk3h_eadl_lookup(const char *path, int Z, int I, int designator, double *out);
int trk3h_num_warnings(trk3h_case *c);
int trk3h_warning(trk3h_case *c, int i, char *out, int outlen);

/* Single-point evaluations of the table builder (parity tests of Cross_sections.f90 routines). */
int trk3h_eval_TotIMFP(trk3h_case *c, double E, int atom, int shell, int kind, double *L, double *dEdx);
int trk3h_eval_EMFP(trk3h_case *c, double E, int kind, double *L, double *dEdx);
/* one value of Diff_cross_section_phonon (Cross_sections.f90:3142) for an electron, with the screening of the case's CDF_elast_Zeff */
int trk3h_eval_dcs_phonon(trk3h_case *c, double Ee, double hw, double *value, double *phonon_A0);
int trk3h_eval_SHI(trk3h_case *c, double E, int atom, int shell, double *inv_L, double *dEdx, double *Zeff);
int trk3h_eval_photon(trk3h_case *c, double E, int atom, int shell, double *L);
int trk3h_sumrules(trk3h_case *c, int atom, int shell, double *ksum, double *fsum);   /* atom<0: phonon CDF */
int trk3h_grid(trk3h_case *c, double Emin, double Emax, double *out, int cap);        /* get_grid_4CS */

/* Write the reference's output directory tree; tallies are the SUMS returned by trk3_mc_run
 * (division by NMC happens here, MAIN.f90:281-314).  out_dir receives the created directory. */
int trk3h_save_output(trk3h_case *c, const trk3_tally_layout *lay, const double *tallies, int NMC,
                      const char *out_root, char *out_dir, int out_dir_len, char *err, int errlen);

const char *trk3_host_version(void);

#ifdef __cplusplus
}
#endif
#endif
