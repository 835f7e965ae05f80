// emul.cpp -- TEST INFRASTRUCTURE: CPU emulation of the wavefront engine.
//
// Compiles the very same physics.cuh / finalize.cuh that the CUDA kernels are built from, with a
// context whose side effects are plain host operations, and drives it generation by generation like
// engine.cu does.  It exists because the development container has no GPU: the no-GPU test-suite uses
// it to check the wavefront re-formulation (independent histories + snapshots + per-iteration
// normalisation) against the time-ordered oracle.  It is never reachable from the product API.
#include <cstdio>
#include <cstring>
#include <vector>
#include "../../trekis-3_b200/csrc/cuda/physics.cuh"
#include "../../trekis-3_b200/csrc/cuda/finalize.cuh"
#include "../../trekis-3_b200/csrc/cuda/engine_host.h"

using namespace trk3;

namespace {
struct EmuCtx {
    static constexpr bool kLean = false;      // every switch compiled in (see elastic_dE, physics.cuh)
    const DevP &p;
    std::vector<Rec> *next;     // [N_SPECIES] next hot generation
    std::vector<Rec> *cold;     // [2] cold electrons / valence holes (consumed after the hot generations)
    unsigned long long *ev, *er, *nel, *nph;
    void push_hot(int sp, const Rec &r) { next[sp].push_back(r); }
    void push(int sp, const Rec &r) {       // same routing as DevCtx::push in engine.cu
        if (sp == SP_ELECTRON && electron_is_cold(p, r)) cold[0].push_back(r);
        else if (sp == SP_VBHOLE && vbhole_is_cold(p, r)) cold[1].push_back(r);
        else next[sp].push_back(r);
    }
    void snap(int sp, const Rec &r, int i) { snapshot_any(*this, sp, r, i); }                 // the engine may defer this to k_snapshot
    void push_ion(const IonEvent &ev) { electron_ion_emit(*this, ev); }     // the engine defers this to k_ion_emit
    void tally(int id, int64_t idx, double v) { p.tally[p.g_off[id] + idx] += v; }
    void add_u32(uint32_t *b, size_t i) { b[i] += 1u; }
    void add_f64(double *b, size_t i, double v) { b[i] += v; }
    void event(int c) { ev[c]++; }
    void error(int c) { er[c]++; }
    void count_electron() { (*nel)++; }
    void count_photon() { (*nph)++; }
};
// run one record through step function `step` until it leaves the kernel
template <class F>
void follow(EmuCtx &c, int sp, Rec r, int ig, Cache k, F step) {
    for (;;) {
        const int st = step(c, r, ig, k);
        if (st == ST_CONT) continue;
        if (st == ST_MOVE) c.push(sp, r);
        else if (st == ST_MOVE_HOT) c.push_hot(sp, r);
        return;
    }
}
}  // namespace

extern "C" int trk3_emul_run(const trk3_config *cfg, const trk3_tables *tab, int64_t it_begin, int64_t it_end, int batch,
                             double *tallies, trk3_stats *stats, double *iter_totE, double *iter_totNel) {
    trk3_tally_layout lay;
    int rc = trk3_tally_layout_init(cfg, tab, &lay);
    if (rc != TRK3_OK) return rc;
    DevP p;
    rc = fill_devp_scalars(*cfg, *tab, lay, p);
    if (rc != TRK3_OK) return rc;
    HostTotals tot; compute_totals(*tab, tot);
    p.ei_E = tab->ei_E; p.ei_L = tab->ei_L; p.ei_tot = tot.ei_tot.data(); p.ee_E = tab->ee_E; p.ee_L = tab->ee_L;
    p.hi_E = tab->hi_E; p.hi_L = tab->hi_L; p.hi_tot = tot.hi_tot.data(); p.he_E = tab->he_E; p.he_L = tab->he_L;
    p.ph_E = tab->ph_E; p.ph_L = tab->ph_L; p.ph_tot = tot.ph_tot.data();
    p.shi_E = tab->shi_E; p.shi_L = tab->shi_L; p.shi_tot = tot.shi_tot.data();
    p.dshi_off = tab->dshi_off; p.dshi_E = tab->dshi_E; p.dshi_L = tab->dshi_L;
    p.eid_off = tab->eid_off; p.eid_hw = tab->eid_hw; p.eid_L = tab->eid_L; p.eed_off = tab->eed_off; p.eed_hw = tab->eed_hw; p.eed_L = tab->eed_L;
    p.hid_off = tab->hid_off; p.hid_hw = tab->hid_hw; p.hid_L = tab->hid_L; p.hed_off = tab->hed_off; p.hed_hw = tab->hed_hw; p.hed_L = tab->hed_L;
    p.dos_E = tab->dos_E; p.dos_DOS = tab->dos_DOS; p.dos_int = tab->dos_int; p.dos_effm = tab->dos_effm; p.out_R = tab->out_R; p.out_V = tab->out_V;
    p.osc_E0 = tab->osc_E0; p.osc_alpha = tab->osc_alpha;
    p.dsf_e_dE = tab->dsf_e_dE; p.dsf_e_emit = tab->dsf_e_emit; p.dsf_e_absorb = tab->dsf_e_absorb; p.ee_emit = tab->ee_emit; p.ee_absorb = tab->ee_absorb;
    p.dsf_h_dE = tab->dsf_h_dE; p.dsf_h_emit = tab->dsf_h_emit; p.dsf_h_absorb = tab->dsf_h_absorb; p.he_emit = tab->he_emit; p.he_absorb = tab->he_absorb;
    // companions of the tables (logs, reciprocals) and the cold ranges, as engine.cu prepares them on the device
    std::vector<std::vector<double>> comp;
    const size_t NS = tab->n_shells;
    const trk3_tables &T = *tab;
#define X(dst, src, n, op) { comp.emplace_back((size_t)(n)); std::vector<double> &v = comp.back(); for (size_t i = 0; i < v.size(); ++i) v[i] = companion_value((src)[i], op); p.dst = v.data(); }
    TRK3_COMPANIONS(X, p, T, NS)
#undef X
    std::vector<uint16_t> luts[N_LUT];
#define X(id, E, n) { build_lut(E, n, luts[id], p.lut[id].l0, p.lut[id].scale); p.lut[id].lut = luts[id].data(); }
    TRK3_LUT_GRIDS(X, T)
#undef X
    p.dos_inv_step = uniform_inv_step(T.dos_E, T.n_dos);
    std::vector<uint16_t> dluts[TRK3_MAX_SHELLS];
    for (int sh = 0; sh < T.n_shells; ++sh) {
        int Mt; double dl; const int Nsh = (int)(T.dshi_off[sh + 1] - T.dshi_off[sh]);
        shi_threshold(T.dshi_E + T.dshi_off[sh], T.dshi_L + T.dshi_off[sh], Nsh, T.shell_Ip[sh], Mt, dl);
        build_inverse_lut(T.dshi_L + T.dshi_off[sh], Nsh, Mt, dluts[sh], p.dshi_lut[sh].l0, p.dshi_lut[sh].scale);
        p.dshi_lut[sh].lut = dluts[sh].data();
    }
    for (int sh = 0; sh < T.n_shells; ++sh) shi_threshold(T.dshi_E + T.dshi_off[sh], T.dshi_L + T.dshi_off[sh], (int)(T.dshi_off[sh + 1] - T.dshi_off[sh]), T.shell_Ip[sh], p.shi_Mtemp[sh], p.shi_dL[sh]);
    cold_range(p.ei_E, p.ei_tot, p.n_ei, p.e_cold, p.e_imfp_cold);
    p.e_warm = p.e_cold; p.e_class[0] = p.e_class[1] = p.e_class[2] = 1.0e300;      // scheduling classes of the CUDA engine: not used here
    cold_range(p.hi_E, p.hi_tot, p.n_hi, p.h_cold, p.h_imfp_cold);
    p.e_iimfp_cold = (p.e_imfp_cold > 0.0) ? 1.0 / p.e_imfp_cold : 0.0; p.h_iimfp_cold = (p.h_imfp_cold > 0.0) ? 1.0 / p.h_imfp_cold : 0.0;
    p.h_warm = p.h_cold;
    p.tally = tallies;
    unsigned long long ev[TRK3_N_EVENT_CLASSES] = {0}, er[TRK3_N_ERRORS] = {0}, nel = 0, nph = 0;
    uint64_t waves = 0;
    if (batch < 1) batch = 64;
    const int Nt = lay.Nt;
    std::vector<double> Dcoef(Nt, 0.0);
    for (int i = 0; i < Nt; ++i) Dcoef[i] = 0.0;
    for (int64_t b0 = it_begin; b0 < it_end; b0 += batch) {
        const uint32_t nb = (uint32_t)std::min<int64_t>(batch, it_end - b0);
        p.batch_begin = (uint32_t)b0; p.batch_n = nb;
        ScratchLayout sl = scratch_layout(p, nb);
        std::vector<uint32_t> U(sl.u32_total, 0u); std::vector<double> D(sl.f64_total, 0.0);
        bind_scratch(p, sl, U.data(), D.data());
        std::vector<Rec> cur[N_SPECIES], nxt[N_SPECIES], cold[2], cold_cur[2];
        EmuCtx c{p, nxt, cold, ev, er, &nel, &nph};
        for (uint32_t k = 0; k < nb; ++k) shi_history(c, (uint32_t)(b0 + k));
        for (;;) {
            size_t n = 0;
            for (int s = 0; s < N_SPECIES; ++s) { cur[s].swap(nxt[s]); nxt[s].clear(); n += cur[s].size(); }
            if (n) {        // one hot generation
                ++waves;
                for (const Rec &r : cur[SP_ELECTRON]) { int ig; Cache k{}; begin_electron(p, r, ig, k); follow(c, SP_ELECTRON, r, ig, k, [](EmuCtx &c, Rec &r, int &ig, Cache &k) { return step_electron<false>(c, r, ig, k); }); }
                for (const Rec &r : cur[SP_VBHOLE]) { int ig; Cache k{}; begin_vbhole(p, r, ig, k); follow(c, SP_VBHOLE, r, ig, k, [](EmuCtx &c, Rec &r, int &ig, Cache &k) { return step_vbhole<false>(c, r, ig, k); }); }
                for (const Rec &r : cur[SP_COREHOLE]) { follow(c, SP_COREHOLE, r, interval_of(p, r.t0), Cache{}, [](EmuCtx &c, Rec &r, int &ig, Cache &) { return step_corehole(c, r, ig); }); }
                for (const Rec &r : cur[SP_PHOTON]) { follow(c, SP_PHOTON, r, interval_of(p, r.t0), Cache{}, [](EmuCtx &c, Rec &r, int &ig, Cache &) { return step_photon(c, r, ig); }); }
                continue;
            }
            if (cold[0].empty() && cold[1].empty()) break;
            // the hot cascade has died out: drain the cold queues (elastic scattering + snapshots only)
            ++waves;
            cold_cur[0].swap(cold[0]); cold[0].clear(); cold_cur[1].swap(cold[1]); cold[1].clear();
            for (const Rec &r : cold_cur[0]) { int ig; Cache k{}; begin_electron(p, r, ig, k); follow(c, SP_ELECTRON, r, ig, k, [](EmuCtx &c, Rec &r, int &ig, Cache &k) { return step_electron<true>(c, r, ig, k); }); }
            for (const Rec &r : cold_cur[1]) { int ig; Cache k{}; begin_vbhole(p, r, ig, k); follow(c, SP_VBHOLE, r, ig, k, [](EmuCtx &c, Rec &r, int &ig, Cache &k) { return step_vbhole<true>(c, r, ig, k); }); }
        }
        std::vector<double> a0((size_t)nb * Nt), a1((size_t)nb * Nt), a2((size_t)nb * Nt), a3((size_t)nb * Nt), a4((size_t)nb * Nt);
        FoldAux fa{a0.data(), a1.data(), a2.data(), a3.data(), a4.data()};
        for (uint32_t il = 0; il < nb; ++il) iter_prefix(p, fa, il);
        for (int64_t j = 0, nj = fold_num_jobs(p); j < nj; ++j) fold_job(p, fa, j);
        for (uint32_t il = 0; il < nb; ++il) for (int i = 0; i < Nt; ++i) {
            size_t o = (size_t)il * Nt + i;
            Dcoef[i] = Dcoef[i] + p.it.diffS[o];
            if (p.it.diffN[o] > 0) Dcoef[i] = Dcoef[i] / (double)p.it.diffN[o];
            if (iter_totE) iter_totE[(size_t)(b0 - it_begin + il) * Nt + i] = a1[o];
            if (iter_totNel) iter_totNel[(size_t)(b0 - it_begin + il) * Nt + i] = a0[o];
        }
    }
    for (int i = 0; i < Nt; ++i) tallies[lay.off[TRK3_OUT_DIFF_COEFF] + i] += Dcoef[i];
    if (stats) {
        std::memset(stats, 0, sizeof *stats);
        for (int q = 0; q < TRK3_N_EVENT_CLASSES; ++q) stats->events[q] = ev[q];
        for (int q = 0; q < TRK3_N_ERRORS; ++q) stats->errors[q] = er[q];
        stats->n_electrons = nel; stats->n_photons = nph; stats->n_waves = waves;
    }
    return TRK3_OK;
}
