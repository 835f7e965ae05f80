/* trk3_layout.c -- shapes and offsets of the Out_* tally arrays of do_Monte_Carlo
 * (Allocate_out_arrays, Sorting_output_data.f90:1144-1250; Radius_for_distributions :1363-1401;
 * set_time_grid, Monte_Carlo.f90:2118-2149).  Plain C, no GPU; compiled into the CUDA library,
 * the host library and the oracle so that all three agree on the buffer layout. */
#include "../../../include/trekis3_gpu.h"
#include <math.h>
#include <string.h>

static int count_time_points(double Tim, double dt, int dt_flag)
{
    if (dt_flag <= 0) return (int)ceil(Tim / dt);
    int i = 0; double t = 0.01;
    while (t < Tim) { ++i; t = t * dt; }
    return i + 1;
}

int trk3_tally_layout_init(const trk3_config *cfg, const trk3_tables *tab, trk3_tally_layout *lay)
{
    if (!cfg || !tab || !lay) return TRK3_E_INVALID;
    memset(lay, 0, sizeof *lay);
    if (!(cfg->Tim > 0.0) || !(cfg->dt > 0.0)) return TRK3_E_INVALID;
    int Nt = count_time_points(cfg->Tim, cfg->dt, cfg->dt_flag);
    if (Nt < 1 || Nt > TRK3_MAX_NT) return TRK3_E_INVALID;
    /* set_time_grid: Nt+1 entries, the last one is Tim+dt */
    int ng;
    if (cfg->dt_flag <= 0) {
        ng = (int)ceil(cfg->Tim / cfg->dt) + 1;
        if (ng > TRK3_MAX_NT + 1) return TRK3_E_INVALID;
        lay->time_grid[0] = cfg->dt;
        for (int i = 1; i <= ng - 2; ++i) lay->time_grid[i] = lay->time_grid[i - 1] + cfg->dt;
        lay->time_grid[ng - 1] = cfg->Tim + cfg->dt;
    } else {
        int i = 0; double t = 0.01;
        while (t <= cfg->Tim) { ++i; t = t * cfg->dt; }
        ng = i + 1;
        if (ng > TRK3_MAX_NT + 1) return TRK3_E_INVALID;
        lay->time_grid[0] = 0.01;
        for (int k = 1; k <= ng - 2; ++k) lay->time_grid[k] = lay->time_grid[k - 1] * cfg->dt;
        lay->time_grid[ng - 1] = cfg->Tim + cfg->dt;
    }
    /* The MC loop (Monte_Carlo.f90:572-665) visits tim_glob = min(time_grid(i), Tim) for i = 1..Nt. */
    lay->Nt = Nt; lay->n_r = tab->n_r; lay->n_atoms = tab->n_atoms; lay->nshl1 = tab->nshl_atom1; lay->n_dos = tab->n_dos;
    int64_t NR = tab->n_r, NA = tab->n_atoms, NS = tab->nshl_atom1, ND = tab->n_dos;
    int64_t len[TRK3_N_TALLIES];
    len[TRK3_OUT_NE] = Nt * NR;  len[TRK3_OUT_EE] = Nt * NR;  len[TRK3_OUT_NPHOT] = Nt * NR;  len[TRK3_OUT_EPHOT] = Nt * NR;
    len[TRK3_OUT_EE_VS_E] = Nt * NR;  len[TRK3_OUT_EH_VS_E] = Nt * ND;  len[TRK3_OUT_ELAT] = Nt * NR;
    len[TRK3_OUT_NH] = Nt * NR * NA * NS;  len[TRK3_OUT_EH] = Nt * NR * NA * NS;  len[TRK3_OUT_EHKIN] = Nt * NR * NA * NS;
    len[TRK3_OUT_TOT_NE] = Nt;  len[TRK3_OUT_TOT_NPHOT] = Nt;  len[TRK3_OUT_TOT_E] = Nt;  len[TRK3_OUT_E_E] = Nt;
    len[TRK3_OUT_E_PHOT] = Nt;  len[TRK3_OUT_E_AT] = Nt;  len[TRK3_OUT_E_H] = Nt * NA * NS;  len[TRK3_OUT_EAT_DENS] = Nt * NR;
    len[TRK3_OUT_THETA] = (Nt + 1) * (int64_t)TRK3_NTHETA;  len[TRK3_OUT_THETA_H] = (Nt + 1) * (int64_t)TRK3_NTHETA;
    len[TRK3_OUT_NE_EM] = Nt;  len[TRK3_OUT_E_EM] = Nt;  len[TRK3_OUT_EE_VS_E_EM] = Nt * NR;
    len[TRK3_OUT_FIELD_ALL] = Nt * NR;  len[TRK3_OUT_E_FIELD] = Nt;  len[TRK3_OUT_DIFF_COEFF] = Nt;
    int64_t o = 0;
    for (int i = 0; i < TRK3_N_TALLIES; ++i) { lay->off[i] = o; lay->len[i] = len[i]; o += len[i]; }
    lay->total = o;
    return TRK3_OK;
}
