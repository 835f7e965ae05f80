// finalize.cuh -- end-of-batch reduction of the per-iteration scratch into the Out_* tallies.
//
// The reference normalises three spectra INSIDE each iteration (1/Tot_Nel and 1/N_VB_h_tot at every grid
// time, Monte_Carlo.f90:1026-1069, and 1/Em_Nel :1105) and accumulates totals that are cumulative in time
// (Tot_Nel, At_NRG, Em_Nel).  The wavefront engine records raw per-iteration counts while many iterations
// are in flight; these two passes turn them into exactly the reference's sums:
//   iter_prefix : one thread per iteration   -> cumulative counts/energies per grid time
//   fold_job    : one thread per output bin  -> sum over the iterations of the batch (fixed order, no atomics)
#pragma once
#include "engine_types.h"

namespace trk3 {

struct FoldAux {            // [nb][Nt] each, produced by iter_prefix
    double *totnel;         // Tot_Nel(it, i)
    double *totE;           // Out_tot_E contribution of the iteration
    double *latcum;         // At_NRG(it, i)
    double *emcnt;          // Em_Nel(it, i)
    double *emE;            // SUM(Em_electrons)(it, i)
};

TRK_HD void iter_prefix(const DevP &p, const FoldAux &a, uint32_t il) {
    const int Nt = p.Nt;
    double nel = 0.0, lat = 0.0, emc = 0.0, eme = 0.0;
    for (int i = 1; i <= Nt; ++i) {
        const size_t q = (size_t)il * (Nt + 2) + i;
        nel += (double)p.it.created[q];
        lat += p.it.elat[q];
        emc += (double)p.it.em_cnt[q];
        eme += p.it.em_E[q];
        const size_t o = (size_t)il * Nt + (i - 1);
        a.totnel[o] = nel; a.latcum[o] = lat; a.emcnt[o] = emc; a.emE[o] = eme;
        a.totE[o] = p.it.esnap[o] + lat;         // SUM(E_e)+SUM(E_h)+SUM(Ehkin)[+SUM(E_ph)]+At_NRG, :947-951
    }
}

TRK_HD int64_t fold_num_jobs(const DevP &p) { return (int64_t)p.Nt * (6 + 2 * p.n_r + p.n_dos + 2 * TRK3_NTHETA); }

// The batch's contribution to ONE element of the tally buffer: returns the partial sum over the iterations
// first, first+stride, ... and the element it belongs to (job-private: no atomics needed; the CUDA kernel gives a job
// to a warp, lanes stride over the iterations and the partial sums are combined in a fixed order).
TRK_HD double fold_partial(const DevP &p, const FoldAux &a, int64_t job, uint32_t first, uint32_t stride, double **dst) {
    const int Nt = p.Nt, NR = p.n_r, ND = p.n_dos;
    const uint32_t nb = p.batch_n;
    double sum = 0.0;
    *dst = nullptr;
    if (job < 6 * (int64_t)Nt) {
        const int which = (int)(job / Nt), i = (int)(job % Nt);      // i is 0-based
        for (uint32_t il = first; il < nb; il += stride) {
            const size_t o = (size_t)il * Nt + i;
            switch (which) {
            case 0: sum += a.totnel[o]; break;
            case 1: sum += a.totE[o]; break;
            case 2: sum += a.latcum[o]; break;
            case 3: sum += (double)p.it.nph[o]; break;
            case 4: sum += a.emcnt[o]; break;
            default: sum += a.emE[o]; break;
            }
        }
        const int ids[6] = {TRK3_OUT_TOT_NE, TRK3_OUT_TOT_E, TRK3_OUT_E_AT, TRK3_OUT_TOT_NPHOT, TRK3_OUT_NE_EM, TRK3_OUT_E_EM};
        *dst = &p.tally[p.g_off[ids[which]] + i];
        return sum;
    }
    job -= 6 * (int64_t)Nt;
    if (job < (int64_t)Nt * NR) {                      // Out_Ee_vs_E(i,j), :1024-1029
        const int i = (int)(job % Nt), j = (int)(job / Nt);
        const double w = (j > 0) ? 1.0 / (p.out_R[j] - p.out_R[j - 1]) : 1.0 / p.out_R[0];
        for (uint32_t il = first; il < nb; il += stride) {
            uint32_t c = p.it.spec_e[((size_t)il * Nt + i) * NR + j];
            if (c) sum += (double)c * (w / a.totnel[(size_t)il * Nt + i]);
        }
        *dst = &p.tally[p.g_off[TRK3_OUT_EE_VS_E] + i + (int64_t)Nt * j];
        return sum;
    }
    job -= (int64_t)Nt * NR;
    if (job < (int64_t)Nt * ND) {                      // Out_Eh_vs_E(i,j), :1032-1040
        const int i = (int)(job % Nt), j = (int)(job / Nt);
        const double w = (j > 0) ? 1.0 / (p.dos_E[j] - p.dos_E[j - 1]) : 1.0 / (p.dos_E[1] - p.dos_E[0]);
        for (uint32_t il = first; il < nb; il += stride) {
            uint32_t c = p.it.spec_h[((size_t)il * Nt + i) * ND + j];
            if (c) sum += (double)c * (w / (double)p.it.nvb[(size_t)il * Nt + i]);
        }
        *dst = &p.tally[p.g_off[TRK3_OUT_EH_VS_E] + i + (int64_t)Nt * j];
        return sum;
    }
    job -= (int64_t)Nt * ND;
    if (job < (int64_t)Nt * TRK3_NTHETA) {             // Out_theta(i,j), :1043-1050
        const int i = (int)(job % Nt), j = (int)(job / Nt);
        for (uint32_t il = first; il < nb; il += stride) {
            uint32_t c = p.it.th_e[((size_t)il * Nt + i) * TRK3_NTHETA + j];
            if (c) sum += (double)c * (1.0 / a.totnel[(size_t)il * Nt + i]);
        }
        *dst = &p.tally[p.g_off[TRK3_OUT_THETA] + i + (int64_t)(Nt + 1) * j];
        return sum;
    }
    job -= (int64_t)Nt * TRK3_NTHETA;
    if (job < (int64_t)Nt * TRK3_NTHETA) {             // Out_theta_h(i,j), :1063-1069
        const int i = (int)(job % Nt), j = (int)(job / Nt);
        for (uint32_t il = first; il < nb; il += stride) {
            uint32_t c = p.it.th_h[((size_t)il * Nt + i) * TRK3_NTHETA + j];
            if (c) sum += (double)c * (1.0 / (double)p.it.nvb[(size_t)il * Nt + i]);
        }
        *dst = &p.tally[p.g_off[TRK3_OUT_THETA_H] + i + (int64_t)(Nt + 1) * j];
        return sum;
    }
    job -= (int64_t)Nt * TRK3_NTHETA;
    if (p.work_function > 0) {                          // Out_Ee_vs_E_Em(i,j), :1100-1109 (Out_E = Out_R/10)
        const int i = (int)(job % Nt), j = (int)(job / Nt);
        const double w = (j > 0) ? 1.0 / (p.out_R[j] / 10.0 - p.out_R[j - 1] / 10.0) : 1.0 / (p.out_R[0] / 10.0);
        for (uint32_t il = first; il < nb; il += stride) {
            uint32_t c = 0;
            for (int iv = 1; iv <= i + 1; ++iv) c += p.it.em_spec[((size_t)il * (Nt + 2) + iv) * NR + j];
            if (c) sum += (double)c * (w / a.emcnt[(size_t)il * Nt + i]);
        }
        *dst = &p.tally[p.g_off[TRK3_OUT_EE_VS_E_EM] + i + (int64_t)Nt * j];
    }
    return sum;
}
TRK_HD void fold_job(const DevP &p, const FoldAux &a, int64_t job) {
    double *dst; const double sum = fold_partial(p, a, job, 0, 1, &dst);
    if (dst) *dst += sum;
}

}  // namespace trk3
