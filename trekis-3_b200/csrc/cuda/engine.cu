// engine.cu -- B200 (sm_100a) wavefront Monte-Carlo engine behind the C ABI of include/trekis3_gpu.h.
//
// Replaces do_Monte_Carlo / Monte_Carlo_modelling (Monte_Carlo.f90:39-679).  Structure:
//   * a batch of iterations is in flight at once; the ion tracks of the batch are walked by k_shi,
//     which fills the first generation of the species queues (electrons, valence holes, core holes);
//   * k_wave<species, HOT> consumes one generation of particles that can still ionise: a warp pulls records
//     from the queue (one atomic per refill, lanes that finish are refilled so warps stay full), every lane
//     follows its particle with the state in registers, secondaries are appended to the next-generation
//     queues with warp-aggregated atomics (one atomicAdd per queue per warp);
//   * particles that can no longer ionise (below the lowest threshold: >90 % of all collisions) are routed to
//     the COLD queues and drained by k_wave<species, COLD> once the hot cascade has died out: those kernels
//     contain only snapshots + elastic scattering, so warps do not diverge into the ionisation code;
//   * radial x time tallies are accumulated in a shared-memory private copy per block and flushed once
//     per block; spectra that the reference normalises per iteration go to per-iteration integer
//     histograms and are folded by k_fold (no atomics, fixed summation order);
//   * the tally sums stay in one packed device buffer so that multi-GPU runs need a single all-reduce.
// No tensor cores: nothing here is a dense contraction.  No CPU fallback: every entry point fails if CUDA does.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <algorithm>
#include <cstdio>
#include <chrono>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>
#include "physics.cuh"
#include "finalize.cuh"
#include "engine_host.h"

using namespace trk3;

// The run's constants (DevP: switches, table pointers, time grid ...): ONE __constant__ image per device, copied at the start of
// every batch; engines that share a device take turns (trk3_mc_run_device holds a per-device lock for the length of a run).
//
// -DTRK_PARAM_CONSTANTS passes DevP as a __grid_constant__ kernel parameter instead (8 KB of parameters per launch; the constants
// then belong to the launch and engines on one device need no lock).  Measured in round 2 and NOT the default: with the
// parameter the pair-creation kernels (k_shi_emit / k_ion_emit) give a handful of particles per iteration another first
// free flight than the oracle (same event counts and energies, 36 of 86 Out_Elat bins off by up to 3 % on C1, 8 iterations),
// while the symbol build agrees with the oracle in every bin; kernel group by kernel group (-DTRK_HYBRID=mask keeps both and lets
// group g read the symbol if bit g is set) the difference follows those two kernels alone.  Not understood, so not shipped.
#ifdef TRK_PARAM_CONSTANTS
#define TRK_P const __grid_constant__ DevP c_p
#define TRK_PA(eng) (eng)->hp
#else
#define TRK_CONST_SYMBOL 1
__constant__ DevP c_p;
#define TRK_P int
#define TRK_PA(eng) 0
#endif
#ifdef TRK_HYBRID
#undef TRK_P
#undef TRK_PA
#ifndef TRK_CONST_SYMBOL
__constant__ DevP c_p;
#endif
#define TRK_PA(eng) (eng)->hp
#define TRK_PSEL(g) const __grid_constant__ DevP TRK_CAT(c_, TRK_NAME(g))
#define TRK_CAT(a, b) TRK_CAT2(a, b)
#define TRK_CAT2(a, b) a##b
#if TRK_HYBRID & 1
#define TRK_NAME_0 unused
#else
#define TRK_NAME_0 p
#endif
#if TRK_HYBRID & 2
#define TRK_NAME_1 unused
#else
#define TRK_NAME_1 p
#endif
#if TRK_HYBRID & 4
#define TRK_NAME_2 unused
#else
#define TRK_NAME_2 p
#endif
#define TRK_NAME(g) TRK_NAME_##g
#define TRK_P0 TRK_PSEL(0)
#define TRK_P1 TRK_PSEL(1)
#define TRK_P2 TRK_PSEL(2)
#else
#define TRK_P0 TRK_P
#define TRK_P1 TRK_P
#define TRK_P2 TRK_P
#endif

// queue ids: 0..3 = next hot generation per species, 4 / 5 = cold electrons / cold valence holes
#define Q_EL_COLD N_SPECIES
#define Q_VB_COLD (N_SPECIES + 1)
#define Q_ION (N_SPECIES + 2)         // impact ionisations of the current generation (IonEvent records, see k_ion_emit)
#define Q_SNAP (N_SPECIES + 3)        // snapshot records of the whole batch (see k_snapshot)
#define N_ECLASS 4                    // energy classes of the hot electrons (class 0 lives in queue SP_ELECTRON)
#define Q_ELC (N_SPECIES + 4)         // classes 1..N_ECLASS-1: Q_ELC + (class - 1)
#define Q_ELW (N_SPECIES + 4 + N_ECLASS - 1)      // warm electrons of the next generation (see DevP::e_warm)
#define Q_VBW (N_SPECIES + 5 + N_ECLASS - 1)      // warm valence holes of the next generation (DevP::h_warm)
#define N_QUEUES (N_SPECIES + 6 + N_ECLASS - 1)
struct QueueSet { Queue q[N_QUEUES]; };
#define N_CLASSES (N_SPECIES + 6)     // timing classes: hot waves per species, k_shi, finalize, cold electrons, cold holes, warm electrons, warm holes

#define TRK_BLOCK_MAX 256          // compile-time upper bound of the wave-kernel block size (launch bounds)
#ifndef TRK_MIN_BLOCKS
#define TRK_MIN_BLOCKS 3
#endif
#ifndef TRK_HOT_MIN_BLOCKS
#define TRK_HOT_MIN_BLOCKS 2
#endif
#define S_EV 0                      // s_cnt layout: events[TRK3_N_EVENT_CLASSES], n_el, n_ph
#define S_NEL TRK3_N_EVENT_CLASSES
#define S_NPH (TRK3_N_EVENT_CLASSES + 1)
#define S_DONE (TRK3_N_EVENT_CLASSES + 2)     // warps of the block that have finished (block_epilogue_last_warp)
#define S_NCNT (TRK3_N_EVENT_CLASSES + 3)

// ------------------------------------------------------------------------------------------------
// device context: the side effects of physics.cuh
// ------------------------------------------------------------------------------------------------
// LEAN: the kernel was chosen for the default switches (CDF elastic scattering, no electron emission, queued snapshots):
// the code of the other settings is not compiled into it (physics.cuh, elastic_dE).
template <bool LEAN>
struct DevCtxT {
    static constexpr bool kLean = LEAN;
    const DevP &p;
    const QueueSet &out;
    double *s_tally;        // block-private tallies (nullptr: straight to global)
    unsigned int *s_cnt;

    // hot electrons are queued by energy class (DevP::e_class); a queue set without class queues takes them all in class 0
    // An electron that used up the time slice of its class without turning cold has proved to be a long history: it is
    // promoted (the tag travels in the otherwise unused `shell` of an electron record: -1 - class).
    __device__ int hot_queue(int sp, const Rec &r) const {
        if (sp != SP_ELECTRON) return sp;
        const int c = max((r.E >= p.e_class[0]) + (r.E >= p.e_class[1]) + (r.E >= p.e_class[2]), min(-1 - r.shell, N_ECLASS - 1));
        if (c == 0 || out.q[Q_ELC + c - 1].cap == 0u) return SP_ELECTRON;
        return Q_ELC + c - 1;
    }
    __device__ void push(int sp, const Rec &r) {
        int qi;
        if (sp == SP_ELECTRON && electron_is_cold(p, r)) qi = Q_EL_COLD;
        else if (sp == SP_ELECTRON && r.E < p.e_warm && out.q[Q_ELW].cap != 0u) qi = Q_ELW;
        else if (sp == SP_VBHOLE && vbhole_is_cold(p, r)) qi = Q_VB_COLD;
        else if (sp == SP_VBHOLE && r.Ehkin < p.h_warm && out.q[Q_VBW].cap != 0u) qi = Q_VBW;
        else qi = hot_queue(sp, r);
        push_q(qi, r);
    }
    __device__ void push_hot(int sp, const Rec &r) { push_q(hot_queue(sp, r), r); }
    // A snapshot is ~450 instructions that only a few lanes of a warp need in any given round, and the kernels are bound by
    // instruction fetch: the lanes just append the particle's state (what snapshot_* reads: 9 numbers) to a queue, and
    // k_snapshot turns the records into tallies with full warps.  `defer` = 0: tally right here (k_snapshot itself).
    int defer;
    __device__ void snap(int sp, const Rec &r, int i) {
        if (!LEAN && !defer) { snapshot_any(*this, sp, r, i); return; }
        const unsigned am = __activemask();
        const int lane = threadIdx.x & 31;
        const int leader = __ffs(am) - 1;
        const Queue &q = out.q[Q_SNAP];
        unsigned base = 0;
        if (lane == leader) base = atomicAdd(q.count, (unsigned)__popc(am));
        base = __shfl_sync(am, base, leader);
        const unsigned slot = base + __popc(am & ((1u << lane) - 1u));
        if (slot >= q.cap) { atomicAdd(&p.errors[TRK3_ERR_QUEUE_OVERFLOW], 1ull); return; }
        q.col[0][slot] = r.E; q.col[3][slot] = r.t0; q.col[5][slot] = r.X; q.col[6][slot] = r.Y; q.col[9][slot] = r.theta; q.col[10][slot] = r.phi;
        if (sp != SP_ELECTRON && sp != SP_PHOTON) { q.col[1][slot] = r.Ehkin; q.col[2][slot] = r.Mass; q.col[8][slot] = r.L; }
        q.iter[slot] = r.iter; q.shell[slot] = r.shell; q.ctr[slot] = (uint32_t)i | ((uint32_t)sp << 16);
    }
    __device__ void push_ion(const IonEvent &ev) {       // an IonEvent travels in the columns of an ordinary record
        Rec r;
        r.E = ev.dE; r.Ehkin = ev.t; r.Mass = ev.X; r.t0 = ev.Y; r.tn = ev.Z; r.X = ev.theta0; r.Y = ev.phi0; r.Z = ev.theta; r.L = ev.phi;
        r.theta = 0.0; r.phi = 0.0; r.id = ev.pid; r.ctr = ev.ctr0; r.iter = ev.iter; r.shell = ev.shell;
        push_q(Q_ION, r);
    }
    __device__ void push_q(int qi, const Rec &r) {
        // warp-aggregated append: lanes pushing to the same queue share one atomicAdd
        const unsigned am = __activemask();
        const unsigned m = __match_any_sync(am, qi);
        const int lane = threadIdx.x & 31;
        const int leader = __ffs(m) - 1;
        const Queue &q = out.q[qi];
        unsigned base = 0;
        if (lane == leader) base = atomicAdd(q.count, (unsigned)__popc(m));
        base = __shfl_sync(m, base, leader);
        const unsigned slot = base + __popc(m & ((1u << lane) - 1u));
        if (slot >= q.cap) { atomicAdd(&p.errors[TRK3_ERR_QUEUE_OVERFLOW], 1ull); return; }
        q.col[0][slot] = r.E; q.col[1][slot] = r.Ehkin; q.col[2][slot] = r.Mass; q.col[3][slot] = r.t0; q.col[4][slot] = r.tn;
        q.col[5][slot] = r.X; q.col[6][slot] = r.Y; q.col[7][slot] = r.Z; q.col[8][slot] = r.L; q.col[9][slot] = r.theta; q.col[10][slot] = r.phi;
        q.id[slot] = r.id; q.ctr[slot] = r.ctr; q.iter[slot] = r.iter; q.shell[slot] = r.shell;
    }
    // The lean kernels queue their snapshots, so the only tally they write is the lattice energy (deposit_lattice): their
    // block-private copy holds Out_Elat alone (2 KB instead of ~46 KB per block: the rest of the SM's memory stays L1 cache).
    __device__ void tally(int id, int64_t idx, double v) {
        if (LEAN) {
            if (s_tally && id == TRK3_OUT_ELAT) atomicAdd(&s_tally[idx], v);
            else atomicAdd(&p.tally[p.g_off[id] + idx], v);
            return;
        }
        const int so = p.s_off[id];
        if (s_tally && so >= 0) atomicAdd(&s_tally[so + idx], v);
        else atomicAdd(&p.tally[p.g_off[id] + idx], v);
    }
    __device__ void add_u32(uint32_t *b, size_t i) { atomicAdd(b + i, 1u); }
    __device__ void add_f64(double *b, size_t i, double v) { atomicAdd(b + i, v); }
    __device__ void event(int cls) { atomicAdd(&s_cnt[S_EV + cls], 1u); }
    __device__ void error(int code) { atomicAdd(&p.errors[code], 1ull); }
    __device__ void count_electron() { atomicAdd(&s_cnt[S_NEL], 1u); }
    __device__ void count_photon() { atomicAdd(&s_cnt[S_NPH], 1u); }
};

__device__ inline void load_rec(const Queue &q, uint32_t i, Rec &r) {
    r.E = q.col[0][i]; r.Ehkin = q.col[1][i]; r.Mass = q.col[2][i]; r.t0 = q.col[3][i]; r.tn = q.col[4][i];
    r.X = q.col[5][i]; r.Y = q.col[6][i]; r.Z = q.col[7][i]; r.L = q.col[8][i]; r.theta = q.col[9][i]; r.phi = q.col[10][i];
    r.id = q.id[i]; r.ctr = q.ctr[i]; r.iter = q.iter[i]; r.shell = q.shell[i];
    r.rc_blk = 0xffffffffu;
}

__device__ inline void block_prologue(double *s_tally, unsigned int *s_cnt, int s_total) {
    for (int i = threadIdx.x; i < s_total; i += blockDim.x) s_tally[i] = 0.0;
    if (threadIdx.x < S_NCNT) s_cnt[threadIdx.x] = 0u;
    __syncthreads();
}
__device__ inline void block_epilogue(const DevP &p, double *s_tally, unsigned int *s_cnt, int cold_species = -1, int warm = 0, bool elat_only = false) {
    __syncthreads();
    if (s_tally && elat_only) {
        double *g = p.tally + p.g_off[TRK3_OUT_ELAT];
        for (int i = threadIdx.x; i < p.s_len[TRK3_OUT_ELAT]; i += blockDim.x) { const double v = s_tally[i]; if (v != 0.0) atomicAdd(g + i, v); }
    } else if (s_tally) {
        // flush the private copy once per block (non-zero bins only)
        for (int id = 0; id < TRK3_N_TALLIES; ++id) {
            const int so = p.s_off[id];
            if (so < 0) continue;
            double *g = p.tally + p.g_off[id];
            for (int i = threadIdx.x; i < p.s_len[id]; i += blockDim.x) {
                const double v = s_tally[so + i];
                if (v != 0.0) atomicAdd(g + i, v);
            }
        }
    }
    if (threadIdx.x < TRK3_N_EVENT_CLASSES) { unsigned v = s_cnt[S_EV + threadIdx.x]; if (v) atomicAdd(&p.events[threadIdx.x], (unsigned long long)v); }
    if (threadIdx.x == 0) {
        // collisions handled by the cold kernels (per species), for the per-kernel roofline
        if (cold_species == SP_ELECTRON && s_cnt[S_EV + TRK3_EV_EL_ELAST]) atomicAdd(p.cnt_el + (warm ? 4 : 2), (unsigned long long)s_cnt[S_EV + TRK3_EV_EL_ELAST]);
        if (cold_species == SP_VBHOLE && s_cnt[S_EV + TRK3_EV_VBH_ELAST]) atomicAdd(p.cnt_el + (warm ? 5 : 3), (unsigned long long)s_cnt[S_EV + TRK3_EV_VBH_ELAST]);
        if (s_cnt[S_NEL]) atomicAdd(p.cnt_el, (unsigned long long)s_cnt[S_NEL]);
        if (s_cnt[S_NPH]) atomicAdd(p.cnt_ph, (unsigned long long)s_cnt[S_NPH]);
    }
}

// The same closing WITHOUT a block-wide barrier, for the hot kernels: there a block lives as long as its longest history, and
// with a barrier at the end every warp that has finished sits in it until then (ncu: stall `barrier` 2-6 cycles per issued
// instruction in k_hot).  Instead a warp that is done counts itself out and leaves; the LAST warp to arrive flushes the block's
// private tallies and counters with its 32 lanes.  (The warps' shared-memory atomics are ordered before their count-out by the
// block-level fence; the last arriver reads after its own atomic on the same counter.)
__device__ inline void block_epilogue_last_warp(const DevP &p, double *s_tally, unsigned int *s_cnt, bool elat_only) {
    const int lane = threadIdx.x & 31;
    __syncwarp();
    __threadfence_block();
    unsigned last = 0u;
    if (lane == 0) last = (atomicAdd(&s_cnt[S_DONE], 1u) == (blockDim.x >> 5) - 1u) ? 1u : 0u;
    last = __shfl_sync(0xffffffffu, last, 0);
    if (!last) return;
    __threadfence_block();
    const volatile double *st = s_tally;
    const volatile unsigned int *sc = s_cnt;
    if (s_tally && elat_only) {
        double *g = p.tally + p.g_off[TRK3_OUT_ELAT];
        for (int i = lane; i < p.s_len[TRK3_OUT_ELAT]; i += 32) { const double v = st[i]; if (v != 0.0) atomicAdd(g + i, v); }
    } else if (s_tally) {
        for (int id = 0; id < TRK3_N_TALLIES; ++id) {
            const int so = p.s_off[id];
            if (so < 0) continue;
            double *g = p.tally + p.g_off[id];
            for (int i = lane; i < p.s_len[id]; i += 32) { const double v = st[so + i]; if (v != 0.0) atomicAdd(g + i, v); }
        }
    }
    if (lane < TRK3_N_EVENT_CLASSES) { const unsigned v = sc[S_EV + lane]; if (v) atomicAdd(&p.events[lane], (unsigned long long)v); }
    if (lane == 0) {
        if (sc[S_NEL]) atomicAdd(p.cnt_el, (unsigned long long)sc[S_NEL]);
        if (sc[S_NPH]) atomicAdd(p.cnt_ph, (unsigned long long)sc[S_NPH]);
    }
}

// ------------------------------------------------------------------------------------------------
// kernels
// ------------------------------------------------------------------------------------------------
// k_shi: the ion track of every iteration of the batch (SHI_Monte_Carlo, Monte_Carlo.f90:2153-2249).
// k_shi: the ion tracks of the batch (SHI_Monte_Carlo, Monte_Carlo.f90:2153-2249).  An ion track is one serial chain
// of a few hundred collisions (latency bound): one warp per ion with a single working lane, so that the chains of
// different ions never serialise each other through divergence.  Only the ion's own part of a collision runs here; the
// collision is written to a staging queue and its electron-hole pair is created by k_shi_emit, one thread per collision.
#define SHI_WARPS 4
// An ion track is one serial chain of a few hundred collisions, so k_shi is bound by the latency of ONE collision.  A warp
// follows one ion and its 32 lanes share the work of a collision instead of 31 of them idling:
//   * lane j keeps Philox block (base + j) of the ion's stream: one evaluation serves ~10 collisions;
//   * the mean free paths of all shells and the total one (which_shell + Next_free_path_2d: one log-log interpolation each)
//     are evaluated by lanes 0..n_shells in one pass; log(E) and log(RN) share one call of the logarithm.
// Every value is computed by the same expressions as in shi_step (physics.cuh): the result is bit-identical.
struct PhiloxWarp {
    uint32_t o0, o1, o2, o3, base;
    bool valid;
    __device__ double draw(const DevP &p, uint32_t iter, uint32_t k, uint64_t id = 0ull) {       // draw k of stream `id` (the ion: 0), see rn()
        const uint32_t blk = k >> 1;
        if (!valid || blk - base >= 32u) {
            base = blk; valid = true;
            philox4x32_10((uint32_t)id, (uint32_t)(id >> 32), base + (threadIdx.x & 31), iter, p.seed_lo, p.seed_hi, o0, o1, o2, o3);
        }
        const int src = (int)(blk - base);
        const uint32_t a = __shfl_sync(0xffffffffu, (k & 1u) ? o2 : o0, src), b = __shfl_sync(0xffffffffu, (k & 1u) ? o3 : o1, src);
        const uint64_t bits = (uint64_t)a | ((uint64_t)b << 32);
        return (double)((bits >> 11) + 1) * (1.0 / 9007199254740992.0);
    }
};
// one collision of the ion (shi_step), all lanes hold the same `s`
__device__ inline void shi_step_warp(DevCtxT<false> &c, Rec &s, ShiEvent &ev, PhiloxWarp &pw) {
    const DevP &p = c.p;
    const int lane = threadIdx.x & 31, NS = p.n_shells;
    const double MSHI = p.ion_mass * TRK_MP;
    if (lane == 0) c.event(TRK3_EV_SHI);
    event_begin(s);
    const uint32_t c0 = s.ctr;
    const double RN1 = pw.draw(p, s.iter, c0), RN2 = pw.draw(p, s.iter, c0 + 1u), RN3 = pw.draw(p, s.iter, c0 + 2u);
    s.ctr = c0 + 3u;
    // log(E) [lane 0] and log(RN3) [lane 1] in one call
    const double lg = m_log(lane == 1 ? RN3 : s.E);
    const double lEs = __shfl_sync(0xffffffffu, lg, 0), lRN3 = __shfl_sync(0xffffffffu, lg, 1);
    // lanes 0..NS-1: MFP of shell `lane` (Which_shell :1786-1832); lane NS: total MFP (Next_free_path_2d)
    const Tab m = tab_shi_L(p), tt = tab_shi_tot(p);
    const int n = tab_find(m, s.E, lEs);
    double val = 1.0e20;
    bool need = false;
    double E1 = 1.0, E2 = 2.0, S1 = 1.0, S2 = 1.0, lE1 = 0.0, lE2 = 1.0, lS1 = 0.0, lS2 = 0.0;
    if (lane < NS) {
        if (n > 1) {
            const double *La = m.L + (size_t)lane * m.N;
            const double a = La[n - 2], b = La[n - 1];
            if (a == b || a > 1e20) val = a;
            else { const double *lLa = m.lL + (size_t)lane * m.N; need = true; E1 = m.E[n - 2]; E2 = m.E[n - 1]; S1 = a; S2 = b; lE1 = m.lE[n - 2]; lE2 = m.lE[n - 1]; lS1 = lLa[n - 2]; lS2 = lLa[n - 1]; }
        }
    } else if (lane == NS) {
        const int n2 = find_2d_from_1d(tt.E, tt.N, s.E, n);
        if (n2 == 1) val = nfp_at(tt, n2, true, s.E, lEs);
        else {
            const double Ll = tt.L[n2 - 2];
            if (Ll >= 1.0e16) val = Ll;
            else { need = true; E1 = tt.E[n2 - 2]; E2 = tt.E[n2 - 1]; S1 = Ll; S2 = tt.L[n2 - 1]; lE1 = tt.lE[n2 - 2]; lE2 = tt.lE[n2 - 1]; lS1 = tt.lL[n2 - 2]; lS2 = tt.lL[n2 - 1]; }
        }
    }
    if (need) val = interp5t(E1, E2, S1, S2, lE1, lE2, lS1, lS2, s.E, lEs);
    const double inv = 1.0 / val;
    double MFP_tot = 0.0;
    for (int q = 0; q < NS; ++q) MFP_tot = MFP_tot + __shfl_sync(0xffffffffu, inv, q);
    MFP_tot = RN1 * MFP_tot;
    double MFP_sum = 0.0;
    int shell = NS - 1;
    for (int q = 0; q < NS; ++q) { MFP_sum = MFP_sum + __shfl_sync(0xffffffffu, inv, q); if (MFP_sum >= MFP_tot) { shell = q; break; } }
    const double lam = __shfl_sync(0xffffffffu, val, NS);
    // SHI_energy_transfer :1719-1780
    const double Tot_N = shi_transfer_target(p, shell, RN2);
    const double dE = shi_transfer_energy(p, shell, Tot_N, m_log(Tot_N));
    const double SHI_IMFP = -lam * lRN3;
    const double Z = s.Z + s.L;
    s.E = s.E - dE; s.t0 = s.tn; s.Z = Z; s.L = SHI_IMFP;
    s.tn = next_time(s.t0, sqrt(2.0 * s.E * TRK_GE / MSHI), SHI_IMFP);
    ev.dE = dE; ev.E_after = s.E; ev.t0 = s.t0; ev.Z = Z; ev.shell = shell; ev.ctr0 = s.ctr; ev.iter = s.iter;
    s.ctr += 2;
    if (s.Z >= p.layer) s.tn = 1e16;
}
// ------------------------------------------------------------------------------------------------
// One hot electron per warp (the classes whose quota is one history per warp: the long delta-electron lineages that are
// the critical path of a batch, and every class of a late, small generation).  A lone history advances one collision per
// ~12 us when a single lane follows it (3 500 dependent instructions, ~25 dependent table loads, 4-5 Philox evaluations);
// here the 32 lanes of its warp share the collision, as in k_shi:
//   * lane j keeps Philox block (base + j) of the electron's stream: one evaluation serves ~8 collisions;
//   * evaluations of the SAME elementary function on independent arguments go through ONE call, one argument per lane:
//     sincos(theta) | sincos(phi);  the shell MFPs of Which_shell and the Next_free_path_1d of every shell;
//     log(E') | log(RN);  elastic MFP | total inelastic MFP of the new energy;  the two row interpolations of
//     interpolate_transferred_energy;
//   * the two inverse-CDF row searches of interpolate_transferred_energy (Find_in_monoton_array_decreasing: ~2 x 7
//     dependent loads) become ONE pass: the lanes count the entries above the sampled value (the rows are stored back to
//     back), which for a non-increasing row is exactly the index the reference's bisection ends on; rows that are not
//     monotone (a handful per table, flagged at upload) take the bisection.
// All lanes hold the same record and the same lookups; every value is computed by the same expressions as in physics.cuh.
// ------------------------------------------------------------------------------------------------
__device__ inline double transferred_energy_warp(const Csr &t, const uint8_t *mono, double Ele, double lE, int i_E, double L_need) {
    const int lane = threadIdx.x & 31;
    if (i_E > 1) { if (fabs(t.Eg[i_E - 2] - Ele) < 1.0e-6) i_E = i_E - 1; }
    const double lLn = m_log(L_need);
    const int64_t o = t.off[i_E - 1];
    if (i_E <= 1) return sample_row(t, o, (int)(t.off[i_E] - o), L_need, lLn);
    const int64_t o2 = t.off[i_E - 2];
    const int n1 = (int)(t.off[i_E] - o), n2 = (int)(o - o2);
    int i1, i2;
    if (mono[i_E - 1] && mono[i_E - 2]) {
        // entries above L_need in rows [o2, o) and [o, o + n1): clamp(count, 1, n) is where Find_in_monoton_array_decreasing ends
        unsigned cnt = 0u;
        const double *La = t.L + o2;
        const int ntot = n1 + n2;
        for (int j = lane; j < ntot; j += 32) { if (La[j] > L_need) cnt += (j < n2) ? 0x10000u : 1u; }
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, s);
        i1 = (int)(cnt & 0xffffu); i2 = (int)(cnt >> 16);
        i1 = min(max(i1, 1), n1); i2 = min(max(i2, 1), n2);
    } else find_dec2(t.L + o, n1, t.L + o2, n2, L_need, i1, i2);
    // the two row interpolations in one call: lane 0 row i_E, lane 1 row i_E - 1
    double lres = 0.0;
    const double hw = (lane & 1) ? sample_row_at(t, o2, n2, i2, L_need, lLn, lres) : sample_row_at(t, o, n1, i1, L_need, lLn, lres);
    const double hw_1 = __shfl_sync(0xffffffffu, hw, 0), hw_2 = __shfl_sync(0xffffffffu, hw, 1);
    const double lhw_1 = __shfl_sync(0xffffffffu, lres, 0), lhw_2 = __shfl_sync(0xffffffffu, lres, 1);
    i_E = i_E - 1;
    if (hw_1 < 1.0e-10 || hw_2 < 1.0e-10) return interp1(t.Eg[i_E - 1], t.Eg[i_E], hw_1, hw_2, Ele);
    return interp5t(t.Eg[i_E - 1], t.Eg[i_E], hw_1, hw_2, t.lEg[i_E - 1], t.lEg[i_E], lhw_1, lhw_2, Ele, lE);
}
// One collision of the electron `e` (Electron_Monte_Carlo, Monte_Carlo.f90:2253-2474; electron_event_head / _inel / _elast /
// _tail of physics.cuh) shared by the lanes of a warp.  RN of the channel roulette has been drawn: `inel` is its outcome.
template <class C>
__device__ inline void electron_collision_warp(C &c, Rec &e, int iv, Cache &k, PhiloxWarp &pw, bool inel) {
    const DevP &p = c.p;
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const double Eel = e.E, t_ev = e.tn, L = e.L;
    // head: the place of the collision (:2290-2296)
    const SinCos sc = m_sincos(lane == 1 ? e.phi : e.theta);
    const double st0 = __shfl_sync(FULL, sc.s, 0), ct0 = __shfl_sync(FULL, sc.c, 0), sp0 = __shfl_sync(FULL, sc.s, 1), cp0 = __shfl_sync(FULL, sc.c, 1);
    const double X = e.X + L * st0 * sp0, Y = e.Y + L * st0 * cp0, Z = e.Z + L * ct0;
    double dE, theta, phi;
    if (inel) {                                                  // impact ionisation (electron_event_inel)
        if (lane == 0) c.event(TRK3_EV_EL_INEL);
        const Tab m = tab_ei_L(p);
        const int NS = p.n_shells;
        const int n = tab_find(m, Eel, k.lE);
        // lanes 0..NS-1: shell MFP of Which_shell (:1786-1832); lanes NS..2NS-1: Next_free_path_1d of the same shells (:2312)
        double val = 1.0e20, E1 = 1.0, E2 = 2.0, S1 = 1.0, S2 = 1.0, lE1 = 0.0, lE2 = 1.0, lS1 = 0.0, lS2 = 0.0;
        bool need = false;
        if (lane < 2 * NS) {
            const int sh = (lane < NS) ? lane : lane - NS;
            const double *La = m.L + (size_t)sh * m.N, *lLa = m.lL + (size_t)sh * m.N;
            if (n > 1) {
                const double a = La[n - 2], b = La[n - 1];
                const bool flat = (lane < NS) ? (a == b || a > 1e20) : (a >= 1.0e16);
                if (flat) val = a;
                else { need = true; E1 = m.E[n - 2]; E2 = m.E[n - 1]; S1 = a; S2 = b; lE1 = m.lE[n - 2]; lE2 = m.lE[n - 1]; lS1 = lLa[n - 2]; lS2 = lLa[n - 1]; }
            } else if (lane >= NS) val = nfp_at(tab_shell(m, sh), 1, false, Eel, k.lE);
        }
        if (need) val = interp5t(E1, E2, S1, S2, lE1, lE2, lS1, lS2, Eel, k.lE);
        const double inv = 1.0 / val;
        double MFP_tot = 0.0;
        for (int q = 0; q < NS; ++q) MFP_tot = MFP_tot + __shfl_sync(FULL, inv, q);
        const double RN2 = pw.draw(p, e.iter, e.ctr++, e.id);
        MFP_tot = RN2 * MFP_tot;
        double MFP_sum = 0.0;
        int shell = NS - 1;
        for (int q = 0; q < NS; ++q) { MFP_sum = MFP_sum + __shfl_sync(FULL, inv, q); if (MFP_sum >= MFP_tot) { shell = q; break; } }
        IonEvent ev;
        ev.pid = e.id; ev.ctr0 = e.ctr; ev.iter = e.iter;
        e.ctr += 2;                                              // the two child ids
        const double IMFP = __shfl_sync(FULL, val, NS + shell);
        // Electron_energy_transfer_inelastic (inelastic_dE)
        const double RN3 = pw.draw(p, e.iter, e.ctr++, e.id);
        const double L_need = m_div(IMFP, RN3);
        double Emin = p.shell_Ip[shell];
        if (Emin <= 1.0e-3) Emin = 1.0e-3;
        const double Emax = (Eel + Emin) / 2.0;
        double E;
        if (p.shell_kocs[shell] == 2) E = beb_transfer(p, Eel, shell, L_need, 1.0, Emin);          // BEB shell: every lane the same bisection
        else if (p.delta_cdf) { const double RNd = pw.draw(p, e.iter, e.ctr++, e.id); E = delta_transfer(p, Eel, shell, Emin, RNd); }   // delta-function CDF: likewise
        else E = transferred_energy_warp(csr_eid(p, shell), p.eid_mono + (size_t)shell * p.n_ei, Eel, k.lE, n, L_need);
        if (E < Emin) E = Emin;
        if (E > Emax) E = Emax;
        if (trk_isnan(E)) E = Emin;
        dE = E;
        theta = m_acos((Eel - dE) / sqrt(Eel * (Eel - dE)));       // Update_electron_angles_El :1189
        if (trk_isnan(theta)) { const double r2 = pw.draw(p, e.iter, e.ctr++, e.id); theta = r2 * TRK_PI; }
        { const double r2 = pw.draw(p, e.iter, e.ctr++, e.id); phi = 2.0 * TRK_PI * r2; }
        if (lane == 0) {
            ev.dE = dE; ev.t = t_ev; ev.X = X; ev.Y = Y; ev.Z = Z; ev.theta0 = e.theta; ev.phi0 = e.phi; ev.theta = theta; ev.phi = phi; ev.shell = shell;
            c.push_ion(ev);                                      // the pair: electron_ion_emit
        }
    } else {                                                     // elastic: energy to the lattice (electron_event_elast, LEAN: CDF scattering)
        if (lane == 0) c.event(TRK3_EV_EL_ELAST);
        const double EMFP = elastic_total(tab_ee(p), Eel, k);
        const double RN = pw.draw(p, e.iter, e.ctr++, e.id);
        const double L_need = m_div(EMFP, RN);
        double hw = transferred_energy_warp(csr_eed(p), p.eed_mono, Eel, k.lE, k.n1, L_need);
        if (hw >= Eel) hw = Eel;
        dE = hw;
        // cos_theta_from_W + Update_particle_angles_lat (angles_lattice, M_eff = 1)
        const double Erest_in = rest_energy(1.0 * TRK_ME), Erest_t = p.Erest_target;
        const double E2mc = Eel + 2.0 * Erest_in, EmW = Eel - dE;
        const double W1 = Eel * E2mc - dE * (Eel + Erest_in + Erest_t);
        const double W2 = Eel * E2mc * EmW * (E2mc - dE);
        double mu = (W2 > 0.0) ? m_div(W1, m_sqrt(W2)) : 0.0;
        if (fabs(mu) > 1.0) { const double RNm = pw.draw(p, e.iter, e.ctr++, e.id); mu = m_cos(TRK_PI * RNm); }
        theta = m_acos(mu);
        const double RN2 = pw.draw(p, e.iter, e.ctr++, e.id);
        phi = 2.0 * TRK_PI * RN2;
        if (lane == 0) {
            if (trk_isnan(theta) || trk_isnan(phi)) c.error(TRK3_ERR_NAN);
            deposit_lattice(c, e, iv, X, Y, dE);
        }
    }
    // tail: lookups of the new energy, next free flight, new direction (:2449-2465)
    const double En = Eel - dE;
    const double RN4 = pw.draw(p, e.iter, e.ctr++, e.id);
    const double lg = m_log(lane == 1 ? RN4 : En);
    const double lEn = __shfl_sync(FULL, lg, 0), lRN4 = __shfl_sync(FULL, lg, 1);
    {   // cache_electron(p, En, k): lane 0 the elastic MFP, lane 1 the total inelastic MFP, one call
        const Tab el = tab_ee(p);
        k.lE = lEn;
        k.n1 = tab_find(el, En, lEn);
        k.n2 = find_2d_from_1d(el.E, el.N, En, k.n1);
        const bool cold = En < p.e_cold;
        const Tab ti = tab_ei_tot(p);
        int ni = 1;
        if (!cold) ni = find_2d_from_1d(ti.E, ti.N, En, tab_find(ti, En, lEn));
        double v = 0.0;
        if (lane == 0 || (lane == 1 && !cold)) v = nfp_at(lane == 0 ? el : ti, lane == 0 ? k.n2 : ni, true, En, lEn);
        k.emfp = __shfl_sync(FULL, v, 0);
        k.iemfp = m_div(1.0, k.emfp);
        if (cold) { k.imfp = p.e_imfp_cold; k.iimfp = p.e_iimfp_cold; }
        else { k.imfp = __shfl_sync(FULL, v, 1); k.iimfp = m_div(1.0, k.imfp); }
    }
    const double MFP_tot = m_div(-lRN4, k.iimfp + k.iemfp);
    double phi1, theta1;
    new_angles_c(e.phi, e.theta, ct0, theta, phi, phi1, theta1);
    e.E = En; e.t0 = t_ev; e.X = X; e.Y = Y; e.Z = Z; e.L = MFP_tot; e.theta = theta1; e.phi = phi1;
    e.tn = next_time(e.t0, vel_electron(e.E), MFP_tot);
    if (e.E < p.cut_off) e.tn = 1.0e20;
    if (lane == 0 && (e.E < -1.0e-9 || trk_isnan(e.E))) c.error(TRK3_ERR_22);
}

__global__ void __launch_bounds__(32 * SHI_WARPS) k_shi(TRK_P0, Queue stage, QueueSet qout, int lanes) {
    __shared__ unsigned int s_cnt[S_NCNT];
    block_prologue(nullptr, s_cnt, 0);
    DevCtxT<false> c{c_p, qout, nullptr, s_cnt, c_p.defer_snap};
    if (lanes <= 0) {       // one ion per warp, the lanes cooperate (n_shells + 1 <= 32 lanes are needed)
        const uint32_t k = blockIdx.x * SHI_WARPS + (threadIdx.x >> 5);
        if (k < c_p.batch_n) {
            Rec s;
            shi_begin(c_p, s, c_p.batch_begin + k);
            PhiloxWarp pw; pw.valid = false; pw.base = 0u;
            while (s.tn < c_p.Tim) {
                ShiEvent ev;
                shi_step_warp(c, s, ev, pw);
                if ((threadIdx.x & 31) == 0) {
                    const unsigned slot = atomicAdd(stage.count, 1u);
                    if (slot < stage.cap) {
                        stage.col[0][slot] = ev.dE; stage.col[1][slot] = ev.E_after; stage.col[3][slot] = ev.t0; stage.col[4][slot] = ev.Z;
                        stage.shell[slot] = ev.shell; stage.ctr[slot] = ev.ctr0; stage.iter[slot] = ev.iter;
                    }
                }
            }
        }
        block_epilogue(c_p, nullptr, s_cnt);
        return;
    }
    // `lanes` ions per warp, each lane on its own: 1 = no divergence between ions at all, more = fewer instruction streams
    const uint32_t k = (blockIdx.x * SHI_WARPS + (threadIdx.x >> 5)) * (uint32_t)lanes + (threadIdx.x & 31);
    if ((int)(threadIdx.x & 31) < lanes && k < c_p.batch_n) {
        Rec s;
        shi_begin(c_p, s, c_p.batch_begin + k);
        while (s.tn < c_p.Tim) {
            ShiEvent ev;
            shi_step(c, s, ev);
            const unsigned slot = atomicAdd(stage.count, 1u);
            if (slot < stage.cap) {
                stage.col[0][slot] = ev.dE; stage.col[1][slot] = ev.E_after; stage.col[3][slot] = ev.t0; stage.col[4][slot] = ev.Z;
                stage.shell[slot] = ev.shell; stage.ctr[slot] = ev.ctr0; stage.iter[slot] = ev.iter;
            }
        }
    }
    block_epilogue(c_p, nullptr, s_cnt);
}
__global__ void __launch_bounds__(256) k_shi_emit(TRK_P0, Queue stage, QueueSet qout) {
    __shared__ unsigned int s_cnt[S_NCNT];
    block_prologue(nullptr, s_cnt, 0);
    DevCtxT<false> c{c_p, qout, nullptr, s_cnt, c_p.defer_snap};
    const uint32_t n = min(*stage.count, stage.cap);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        ShiEvent ev;
        ev.dE = stage.col[0][i]; ev.E_after = stage.col[1][i]; ev.t0 = stage.col[3][i]; ev.Z = stage.col[4][i];
        ev.shell = stage.shell[i]; ev.ctr0 = stage.ctr[i]; ev.iter = stage.iter[i];
        shi_emit(c, ev);
    }
    block_epilogue(c_p, nullptr, s_cnt);
}

// k_ion_emit: the electron-hole pairs of the impact ionisations of a generation (electron_ion_emit), one thread per
// ionisation; launched right after the hot kernels of the generation, it fills the same next-generation queues.
__global__ void __launch_bounds__(256) k_ion_emit(TRK_P0, Queue ionq, QueueSet qout) {
    __shared__ unsigned int s_cnt[S_NCNT];
    block_prologue(nullptr, s_cnt, 0);
    DevCtxT<false> c{c_p, qout, nullptr, s_cnt, c_p.defer_snap};
    const uint32_t n = min(*ionq.count, ionq.cap);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        Rec r;
        load_rec(ionq, i, r);
        IonEvent ev;
        ev.dE = r.E; ev.t = r.Ehkin; ev.X = r.Mass; ev.Y = r.t0; ev.Z = r.tn; ev.theta0 = r.X; ev.phi0 = r.Y; ev.theta = r.Z; ev.phi = r.L;
        ev.pid = r.id; ev.ctr0 = r.ctr; ev.iter = r.iter; ev.shell = r.shell;
        electron_ion_emit(c, ev);
    }
    block_epilogue(c_p, nullptr, s_cnt);
}

// k_snapshot: the snapshot records of a batch -> tallies (Calculated_statistics, Monte_Carlo.f90:881-1110), one thread per record
__global__ void __launch_bounds__(256) k_snapshot(TRK_P2, Queue sq, QueueSet qout, int use_smem) {
    extern __shared__ double s_dyn[];
    __shared__ unsigned int s_cnt[S_NCNT];
    double *s_tally = (use_smem && c_p.s_total > 0) ? s_dyn : nullptr;
    block_prologue(s_tally ? s_tally : s_dyn, s_cnt, s_tally ? c_p.s_total : 0);
    DevCtxT<false> c{c_p, qout, s_tally, s_cnt, 0};
    const uint32_t n = min(*sq.count, sq.cap);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        Rec r;
        const uint32_t tag = sq.ctr[i];
        const int sp = (int)(tag >> 16), ig = (int)(tag & 0xffffu);
        r.E = sq.col[0][i]; r.t0 = sq.col[3][i]; r.X = sq.col[5][i]; r.Y = sq.col[6][i]; r.theta = sq.col[9][i]; r.phi = sq.col[10][i];
        r.Ehkin = 0.0; r.Mass = 1.0; r.L = 0.0;
        if (sp != SP_ELECTRON && sp != SP_PHOTON) { r.Ehkin = sq.col[1][i]; r.Mass = sq.col[2][i]; r.L = sq.col[8][i]; }
        r.iter = sq.iter[i]; r.shell = sq.shell[i];
        snapshot_any(c, sp, r, ig);
    }
    block_epilogue(c_p, s_tally, s_cnt);
}

// k_wave<SP, COLD>: records [first, n_in) of queue qin, histories followed with lane refill until they end or
// have to change queue (hot -> cold when the particle can no longer ionise, core hole -> valence hole, ...).
// The number of records is read from the queue's device counter: the host enqueues the launches of a generation without
// knowing how many records the previous one produced (no host round trip per generation); blocks without work leave at once.
template <int SP, bool COLD, bool LEAN>
__global__ void __launch_bounds__(TRK_BLOCK_MAX, TRK_MIN_BLOCKS) k_wave(TRK_P1, Queue qin, uint32_t first, uint32_t *head, QueueSet qout, int use_smem, int refill_min, int slice, int warm) {
    extern __shared__ double s_dyn[];
    __shared__ unsigned int s_cnt[S_NCNT];
    const uint32_t n_in = min(*qin.count, qin.cap);       // records [first, n_in)
    if (first + blockIdx.x * blockDim.x >= n_in) return;          // nothing left for this block (uniform per block, before any barrier)
    double *s_tally = (use_smem && c_p.s_total > 0) ? s_dyn : nullptr;
    block_prologue(s_tally ? s_tally : s_dyn, s_cnt, s_tally ? (LEAN ? c_p.s_len[TRK3_OUT_ELAT] : c_p.s_total) : 0);
    DevCtxT<LEAN> c{c_p, qout, s_tally, s_cnt, c_p.defer_snap};
    const int lane = threadIdx.x & 31;
    bool active = false, exhausted = false;
    Rec r;
    Cache k{};
    int ig = 0, nev = 0;
    for (;;) {
        const unsigned idle = __ballot_sync(0xffffffffu, !active);
        if (idle && !exhausted && (__popc(idle) >= refill_min || idle == 0xffffffffu)) {
            const int nidle = __popc(idle);
            uint32_t base = 0;
            if (lane == 0) base = atomicAdd(head, (uint32_t)nidle);
            base = __shfl_sync(0xffffffffu, base, 0) + first;
            if (base + (uint32_t)nidle >= n_in) exhausted = true;
            if (!active) {
                const uint32_t rank = __popc(idle & ((1u << lane) - 1u));
                const uint32_t my = base + rank;
                if (rank < (uint32_t)nidle && my < n_in) {
                    load_rec(qin, my, r);
                    active = true; nev = 0;
                    if (SP == SP_ELECTRON) begin_electron(c_p, r, ig, k);
                    else if (SP == SP_VBHOLE) begin_vbhole(c_p, r, ig, k);
                    else ig = interval_of(c_p, r.t0);
                }
            }
        }
        if (__ballot_sync(0xffffffffu, active) == 0u) { if (exhausted) break; continue; }
        if (active) {
            int st;
            if (SP == SP_ELECTRON) st = step_electron<COLD>(c, r, ig, k, COLD && warm);
            else if (SP == SP_VBHOLE) st = step_vbhole<COLD>(c, r, ig, k, COLD && warm);
            else if (SP == SP_COREHOLE) st = step_corehole(c, r, ig);
            else st = step_photon(c, r, ig);
            // time slicing of the hot cascade: after `slice` collisions the record goes back to the queue, so that the
            // duration of a generation is bounded and the work of long histories spreads over many lanes
            if (!COLD && st == ST_CONT && ++nev >= slice) st = ST_MOVE_HOT;
            if (COLD && warm && st == ST_CONT && ++nev >= slice) st = ST_MOVE;      // a warm electron: on to the next generation
            if (st != ST_CONT) {
                if (st == ST_MOVE) c.push(SP, r);
                else if (st == ST_MOVE_HOT) c.push_hot(SP, r);
                active = false;
            }
        }
    }
    block_epilogue(c_p, s_tally, s_cnt, COLD ? SP : -1, warm, LEAN);
}

// k_hot<SP>: one generation of carriers that can still ionise (SP = electron or valence hole).  Every round a lane draws
// the channel roulette of its collision first, so that the kernel (not the event handler) branches on the channel: the
// handlers are instantiated for one channel each, and for electrons the head and tail of the collision are shared.
// (Making the ionising lanes wait until several have gathered was measured: no gain once the pairs were deferred.)
template <int SP> __device__ inline bool hot_roulette(const Cache &k, double RN) { return SP == SP_ELECTRON ? electron_roulette_inelastic(k, RN) : vbhole_roulette_inelastic(k, RN); }
template <int SP, int MODE, class C> __device__ inline void hot_event(C &c, Rec &r, int ig, Cache &k, double RN) {
    if (SP == SP_ELECTRON) electron_event_t<MODE>(c, r, ig, k, RN); else vbhole_event_t<MODE>(c, r, ig, k, RN);
}
template <int SP> __device__ inline bool hot_leaves(const DevP &p, const Rec &r) { return SP == SP_ELECTRON ? electron_leaves_hot(p, r) : vbhole_leaves_hot(p, r); }

// The input of k_hot: the records of one generation by energy class (class 0 = all of them for the valence holes).
// A history's remaining number of collisions grows with its energy and a warp advances at the pace of its slowest
// lane: one collision of a lone history takes ~4 us, one round of a warp whose 32 lanes disagree on the channel and are
// refilled all the time ~25 us (measured).  So the long histories get warps of their own: `quota[c]` = how
// many histories of class c a warp follows at once, `wend` = the warps that start on class c (highest class first, i.e.
// in the blocks that are scheduled first).  A warp whose class is used up goes on with the highest class that has
// records left.
struct HotIn {
    Queue q[N_ECLASS];
    uint32_t *head[N_ECLASS];
    int qmax[N_ECLASS];            // most histories of class c a warp follows at once
    int slice[N_ECLASS];           // collisions after which a history of class c goes back to the queue (promoted by one class)
    int ncls, spread, quota_min, coop, weighted;
};
// What the launch works with, derived ON THE DEVICE from the queue counters (the host does not know them when it enqueues
// the launch): records per class, histories per warp, and the warps that start on each class.
struct HotPlan {
    uint32_t n[N_ECLASS];
    int quota[N_ECLASS];
    uint32_t wend[N_ECLASS];       // warps [wend[c+1], wend[c]) start on class c (wend[ncls] = 0); warps >= wend[0] have nothing to do
};
// warps per class.  Least: n / (the class's largest quota).  If that fills the GPU, the warps are shared out in proportion
// (quota = largest).  Otherwise the spare warps go to the classes from the top down, until every history of a class has a
// warp of its own: a small generation spreads over all warps, the long histories first.  W = warps of the grid.
__device__ inline void hot_plan(const HotIn &in, uint32_t W, HotPlan &pl) {
    uint32_t need[N_ECLASS], w[N_ECLASS], total_need = 0;
    for (int c = 0; c < N_ECLASS; ++c) { pl.n[c] = 0; pl.quota[c] = 1; pl.wend[c] = 0; need[c] = 0; w[c] = 0; }
    for (int c = 0; c < in.ncls; ++c) {
        pl.n[c] = min(*in.q[c].count, in.q[c].cap);
        const uint32_t qmax = (uint32_t)in.qmax[c];
        need[c] = (pl.n[c] + qmax - 1) / qmax; total_need += need[c];
        pl.quota[c] = (int)qmax;
    }
    if (total_need >= W || !in.spread) {
        // more warp-loads than warps: the generation is bound by throughput, and a warp-load of class c keeps its warp busy
        // for slice[c] rounds of a duration that grows with the histories per warp (measured: ~12 us for a lone lane, ~6 us
        // when the lanes share it, ~25 us for 32 lanes that disagree on the channel).  The warps are shared out in
        // proportion to that WORK, so that all classes end together (in proportion to the warp-loads alone, the long
        // time slices of the upper classes were the tail of the early generations).
        float cost[N_ECLASS], total = 0.f;
        for (int c = 0; c < in.ncls; ++c) {
            const float tau = (in.qmax[c] == 1 && in.coop ? 6.f : 12.f) + 0.42f * (float)(in.qmax[c] - 1);
            cost[c] = in.weighted ? (float)need[c] * (float)in.slice[c] * tau : (float)need[c];
            total += cost[c];
        }
        for (int c = 0; c < in.ncls; ++c) w[c] = need[c] ? max(1u, (uint32_t)(cost[c] * (float)W / fmaxf(total, 1.f))) : 0u;
        if (total_need < W) for (int c = 0; c < in.ncls; ++c) w[c] = need[c];
    } else {
        uint32_t spare = W - total_need;
        for (int c = in.ncls - 1; c >= 0; --c) {
            const uint32_t extra = min(spare, pl.n[c] - need[c]);
            w[c] = need[c] + extra; spare -= extra;
            if (w[c]) pl.quota[c] = (int)max((pl.n[c] + w[c] - 1) / w[c], (uint32_t)in.quota_min);
        }
    }
    uint32_t wsum = 0;
    for (int c = in.ncls - 1; c >= 0; --c) { wsum += w[c]; pl.wend[c] = wsum; }
}
template <int SP, bool LEAN>
__global__ void __launch_bounds__(TRK_BLOCK_MAX, TRK_HOT_MIN_BLOCKS) k_hot(TRK_P1, HotIn in, QueueSet qout, int use_smem, int refill_min, int coop) {
    extern __shared__ double s_dyn[];
    __shared__ unsigned int s_cnt[S_NCNT];
    __shared__ HotPlan pl;
    if (threadIdx.x == 0) hot_plan(in, gridDim.x * (blockDim.x >> 5), pl);
    double *s_tally = (use_smem && c_p.s_total > 0) ? s_dyn : nullptr;
    block_prologue(s_tally ? s_tally : s_dyn, s_cnt, s_tally ? (LEAN ? c_p.s_len[TRK3_OUT_ELAT] : c_p.s_total) : 0);
    if (blockIdx.x * (blockDim.x >> 5) >= pl.wend[0]) return;          // no warp of this block has work (uniform per block)
    DevCtxT<LEAN> c{c_p, qout, s_tally, s_cnt, c_p.defer_snap};
    const int lane = threadIdx.x & 31;
    bool active = false, have_rn = false;
    Rec r;
    Cache k{};
    double RN = 0.0;
    int ig = 0, nev = 0, my_slice = 0, my_next = 0;
    const uint32_t gw = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    bool exhausted = gw >= pl.wend[0];       // a warp beyond the planned ones: nothing to do (it still joins the block's epilogue)
    int cls = 0;
    unsigned exh = 0u;              // classes whose queue is used up (warp-uniform)
    for (int q = in.ncls - 1; q >= 0; --q) { if (pl.n[q] == 0u) exh |= 1u << q; }
    for (int q = in.ncls - 1; q > 0; --q) { if (gw < pl.wend[q]) { cls = q; break; } }
    // classes whose quota is ONE history per warp are followed by the whole warp (electron_collision_warp)
    const bool coop_ok = LEAN && SP == SP_ELECTRON && coop && 2 * c_p.n_shells <= 32;
    PhiloxWarp pw; pw.valid = false; pw.base = 0u;
    for (;;) {
        const unsigned idle = __ballot_sync(0xffffffffu, !active);
        if (!exhausted && ((exh >> cls) & 1u)) {
            int q = in.ncls - 1;
            while (q >= 0 && ((exh >> q) & 1u)) --q;
            if (q < 0) exhausted = true; else cls = q;
        }
        if (coop_ok && !exhausted && pl.quota[cls] == 1 && idle == 0xffffffffu) {
            // one record of class `cls`, all lanes hold it
            uint32_t my = 0;
            if (lane == 0) my = atomicAdd(in.head[cls], 1u);
            my = __shfl_sync(0xffffffffu, my, 0);
            if (my >= pl.n[cls]) { exh |= 1u << cls; continue; }
            load_rec(in.q[cls], my, r);
            begin_electron(c_p, r, ig, k);
            pw.valid = false;
            const int slice_c = in.slice[cls], next_c = min(cls + 1, in.ncls - 1);
            for (int n_coll = 0;;) {
                while (ig <= c_p.Nt && c_p.tg[ig - 1] <= r.tn) { if (lane == 0) c.snap(SP, r, ig); ++ig; }
                if (ig > c_p.Nt) break;
                event_begin(r);
                const double RN1 = pw.draw(c_p, r.iter, r.ctr++, r.id);
                electron_collision_warp(c, r, ig, k, pw, electron_roulette_inelastic(k, RN1));
                ++n_coll;
                if (electron_leaves_hot(c_p, r)) { if (lane == 0) c.push(SP, r); break; }
                if (n_coll >= slice_c && r.tn < c_p.Tim) { r.shell = -1 - next_c; if (lane == 0) c.push_hot(SP, r); break; }
            }
            continue;
        }
        const int quota = pl.quota[cls];
        const int room = quota - (32 - __popc(idle));
        if (room > 0 && !exhausted && (room >= min(refill_min, quota) || idle == 0xffffffffu)) {
            const uint32_t n_in = pl.n[cls];
            uint32_t base = 0;
            if (lane == 0) base = atomicAdd(in.head[cls], (uint32_t)room);
            base = __shfl_sync(0xffffffffu, base, 0);
            if (base + (uint32_t)room >= n_in) exh |= 1u << cls;
            if (!active) {
                const uint32_t rank = __popc(idle & ((1u << lane) - 1u));
                const uint32_t my = base + rank;
                if (rank < (uint32_t)room && my < n_in) {
                    load_rec(in.q[cls], my, r);
                    active = true; nev = 0; have_rn = false;
                    my_slice = in.slice[cls]; my_next = min(cls + 1, in.ncls - 1);
                    if (SP == SP_ELECTRON) begin_electron(c_p, r, ig, k); else begin_vbhole(c_p, r, ig, k);
                }
            }
        }
        if (__ballot_sync(0xffffffffu, active) == 0u) { if (exhausted) break; continue; }
        // snapshots of the current free flight, then the channel roulette of the collision that ends it
        int want = 0;       // 1 elastic, 2 inelastic
        if (active) {
            while (ig <= c_p.Nt && c_p.tg[ig - 1] <= r.tn) { c.snap(SP, r, ig); ++ig; }
            if (ig > c_p.Nt) active = false;
            else {
                if (!have_rn) { event_begin(r); RN = rn(c_p, r); have_rn = true; }
                want = hot_roulette<SP>(k, RN) ? 2 : 1;
            }
        }
        bool done_event = false;
        const bool go_in = (want == 2);
        if (SP == SP_ELECTRON) {
            // head and tail of a collision are common to both channels: lanes that disagree on the channel only take the
            // channel-specific middle part one after the other (electron_event_head / _inel / _elast / _tail, physics.cuh)
            ElEvent s;
            const bool ev_now = go_in || want == 1;
            if (ev_now) electron_event_head(r, s);
            if (go_in) electron_event_inel(c, r, k, s);
            if (want == 1) electron_event_elast(c, r, ig, k, s);
            if (ev_now) { electron_event_tail(c, r, ig, k, s); done_event = true; }
        } else {
            if (go_in) { hot_event<SP, EV_INELASTIC>(c, r, ig, k, RN); done_event = true; }
            if (want == 1) { hot_event<SP, EV_ELASTIC>(c, r, ig, k, RN); done_event = true; }
        }
        if (done_event) {
            have_rn = false;
            // leave when the carrier can no longer ionise (cold queue) or, after `slice` collisions, back to the hot queue
            ++nev;
            if (hot_leaves<SP>(c_p, r)) { c.push(SP, r); active = false; }
            else if (nev >= my_slice && r.tn < c_p.Tim) {
                if (SP == SP_ELECTRON) r.shell = -1 - my_next;
                c.push_hot(SP, r); active = false;
            }
        }
    }
    block_epilogue_last_warp(c_p, s_tally, s_cnt, LEAN);
}

__global__ void k_iter_prefix(TRK_P2, FoldAux a) {
    const uint32_t il = blockIdx.x * blockDim.x + threadIdx.x;
    if (il < c_p.batch_n) iter_prefix(c_p, a, il);
}
__global__ void k_fold(TRK_P2, FoldAux a, int64_t njobs) {      // one warp per output element
    const int64_t j = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (j >= njobs) return;
    double *dst;
    double sum = fold_partial(c_p, a, j, threadIdx.x & 31, 32, &dst);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if ((threadIdx.x & 31) == 0 && dst) *dst += sum;
}
// companion arrays of the tables (TRK3_COMPANIONS): evaluated on the device so that they carry the device's log()
// DevP::eid_mono / eed_mono: 1 for the rows of a differential table (CSR offsets `off`) that are non-increasing; one warp per row
__global__ void k_monotone_rows(const int64_t *off, const double *L, size_t nrows, uint8_t *flag) {
    const size_t r = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (r >= nrows) return;                                         // whole warps leave together
    const int lane = threadIdx.x & 31;
    const int64_t a = off[r], b = off[r + 1];
    int bad = 0;
    for (int64_t j = a + lane; j + 1 < b; j += 32) bad |= !(L[j + 1] <= L[j]);
    bad = __any_sync(0xffffffffu, bad);
    if (lane == 0) flag[r] = bad ? 0 : 1;
}
struct CompanionJob { double *dst; const double *src; unsigned long long n; int op; };
#define TRK_MAX_COMPANIONS 32
struct CompanionJobs { CompanionJob j[TRK_MAX_COMPANIONS]; int nj; };
// all companion arrays in one launch: blockIdx.y = array, the blocks of a row stride over its elements
__global__ void k_companions(const CompanionJobs jobs) {
    const CompanionJob jb = jobs.j[blockIdx.y];
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < jb.n; i += (size_t)gridDim.x * blockDim.x)
        jb.dst[i] = companion_value(jb.src[i], jb.op);
}
// end of a generation: remember the fill of the ionisation queue (overflow check on the host) and empty it
__global__ void k_ion_reset(uint32_t *cnt) { if (threadIdx.x == 0) { if (cnt[0] > cnt[1]) cnt[1] = cnt[0]; cnt[0] = 0; } }
__global__ void k_axpy(double *dst, const double *src, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] += src[i];
}

// ------------------------------------------------------------------------------------------------
// NCCL, bound at run time (see include/trekis3_gpu.h "Multi-GPU"): the handful of entry points the path needs
// ------------------------------------------------------------------------------------------------
namespace nccl_rt {
typedef struct { char internal[TRK3_NCCL_UNIQUE_ID_BYTES]; } UniqueId;      // ncclUniqueId (nccl.h: 128 bytes)
typedef void *Comm;
enum { kSuccess = 0, kFloat64 = 8, kSum = 0 };                               // ncclSuccess, ncclDouble, ncclSum
struct Api {
    void *handle = nullptr;
    int (*GetUniqueId)(UniqueId *) = nullptr;
    int (*CommInitRank)(Comm *, int, UniqueId, int) = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, Comm, cudaStream_t) = nullptr;
    int (*CommDestroy)(Comm) = nullptr;
    int (*CommAbort)(Comm) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    std::string err;
};
static Api &api() {
    static Api a;
    static std::once_flag once;
    std::call_once(once, [] {
        for (const char *name : {"libnccl.so.2", "libnccl.so"}) { a.handle = dlopen(name, RTLD_NOW | RTLD_GLOBAL); if (a.handle) break; }
        if (!a.handle) { a.err = std::string("NCCL not found: ") + dlerror(); return; }
#define SYM(field, name) a.field = (decltype(a.field))dlsym(a.handle, name); if (!a.field) { a.err = std::string("NCCL symbol missing: ") + name; return; }
        SYM(GetUniqueId, "ncclGetUniqueId") SYM(CommInitRank, "ncclCommInitRank") SYM(AllReduce, "ncclAllReduce")
        SYM(CommDestroy, "ncclCommDestroy") SYM(CommAbort, "ncclCommAbort") SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
    });
    return a;
}
}  // namespace nccl_rt

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
struct trk3_engine {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t stream_c = nullptr;       // second stream: the cold kernels run beside the hot cascade
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    cudaStream_t stream_w = nullptr; cudaEvent_t ev_w = nullptr;   // warm electrons of a generation
    cudaStream_t stream_wh = nullptr; cudaEvent_t ev_wh = nullptr; // warm valence holes
    cudaStream_t stream_sp[3] = {nullptr, nullptr, nullptr};      // the rarer species of a generation run beside the electrons
    cudaEvent_t ev_gen = nullptr, ev_sp[3] = {nullptr, nullptr, nullptr};
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    trk3_config cfg{};
    trk3_tally_layout lay{};
    DevP hp{};                         // the run's constants (device pointers inside): passed to every launch by value
    std::vector<void *> allocs;
    std::vector<std::pair<void *, size_t>> tab_allocs;   // table arrays in binding order (re-used by trk3_mc_reload_tables)
    size_t tab_cursor = 0;
    // the table arrays live back to back in ONE arena, so that one access-policy window can pin them in L2 (option l2_persist)
    char *tab_arena = nullptr; size_t tab_arena_cap = 0, tab_arena_used = 0;
    int opt_l2_persist = 1, l2_persist_applied = 0;
    uint64_t h2d_bytes = 0;                 // bytes of the last table binding
    // A table RE-binding (the per-call input copy of a persistent handle) goes through a pinned host mirror of the arena: every
    // array is copied into the mirror at its arena offset and a contiguous run of arrays leaves in ONE DMA (stage_flush),
    // instead of ~60 pageable copies with a synchronisation behind every temporary (option "stage_uploads", default on).
    char *stage = nullptr; size_t stage_cap = 0, stage_lo = (size_t)-1, stage_hi = 0;
    bool staging = false, direct_pending = false, stage_in_flight = false;
    cudaEvent_t ev_stage = nullptr;
    int opt_stage_uploads = 1;
    double nel_est = 1000.0;
    // options
    uint32_t *h_qcount = nullptr;       // pinned ring of counter snapshots (run-ahead generation loop)
    std::vector<cudaEvent_t> ring_ev;   // one event per ring slot
    int opt_weighted = 0;               // warps per energy class in proportion to the class's work, not its warp-loads
    int opt_coop = 1;                   // hot electrons that have a warp of their own: the lanes share the collision
    int opt_run_ahead = 1;              // generations the host may enqueue beyond the last counter snapshot it has seen
    int opt_batch = 4096, opt_use_smem = 1, opt_refill_min = 8, opt_blocks_per_sm = 0, opt_max_generations = 1 << 20, opt_block = 256;
    int opt_hot_slice = 64;
    int opt_shi_lanes = 0;                 // 0: one ion per warp, the lanes share the work of a collision (k_shi)
    int opt_hot_block = 0;
    // energy classes of the hot electrons: lower edges [eV] of classes 1..3 and the most histories a warp follows at once
    double opt_warm_pinel = 0.5;           // electrons are "warm" below the energy where the ionisation probability per collision reaches this (0: off)
    int opt_warm_slice = 64;
    int opt_warm_holes = 1;                // the warm class for valence holes too
    double e_warm_auto = -1.0, h_warm_auto = -1.0;      // from warm_P / warm_Ph and opt_warm_pinel (< 0: to be evaluated)
    std::vector<double> warm_E, warm_P, warm_Eh, warm_Ph;    // ionisation probability per collision on the inelastic energy grids
    int opt_lean = 1;                      // 0: always the kernels with every switch compiled in
    int opt_hot_classes = N_ECLASS;
    double opt_class_E[N_ECLASS - 1] = {200.0, 500.0, 1300.0};
    int opt_class_quota[N_ECLASS] = {32, 8, 3, 1};
    int opt_class_slice[N_ECLASS] = {8, 16, 32, 64};
    int opt_spread = 1, opt_quota_min = 1, opt_species_streams = 1, opt_defer_snap = 1;

    double opt_cap_factor = 2.0;
    size_t opt_queue_bytes_max = (size_t)64 << 30;       // of the 180 GB of a B200
    // per-batch resources
    uint32_t nb_alloc = 0;
    std::vector<double> batch_sig;     // what the per-batch resources were sized for (batch_signature)
    QueueSet qs[2]{};
    QueueSet qs_x{};                   // output set of the cold kernels (small hot queues + the shared cold queues)
    uint32_t *d_qcount = nullptr;      // QC_* layout below: hot counts of both generations, cold counts, heads
    uint32_t *d_u32 = nullptr; double *d_f64 = nullptr; ScratchLayout sl{};
    FoldAux fa{};
    double *d_tally = nullptr, *d_small = nullptr, *d_tally_bak = nullptr;
    unsigned long long *d_counters = nullptr, *d_counters_bak = nullptr;      // events, errors, n_el, n_ph
    int n_sm = 0, smem_optin = 0;
    bool own_stream = true, own_tally = true;
    int opt_profile = 0;               // 1: time every kernel class with CUDA events (bench.py roofline)
    double class_ms[N_CLASSES] = {0};          // hot k_wave<species>, k_shi, finalize, cold electrons, cold holes
    uint64_t class_launches[N_CLASSES] = {0};
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev_pool;
    std::vector<std::pair<int, int>> ev_pending;   // (class, pool index)
    std::vector<std::pair<int, uint32_t>> ev_info; // (generation, records) of the same launches: option "profile" = 2 prints them
    int cur_gen = -1;
    // multi-GPU: with a communicator attached every run ends with one all-reduce of the tally buffer
    nccl_rt::Comm comm = nullptr; int comm_size = 1; bool own_comm = false;
    double allreduce_ms = 0.0;
    bool failed = false;               // a run ended with an error: device state undefined until trk3_mc_reset
    // results
    std::vector<double> iter_totE;
    std::vector<double> Dcoef;
    std::string err;
    uint64_t launches = 0;
};

// d_qcount layout (uint32): [0..3] hot counts generation A, [4..7] generation B, [8..9] cold counts,
// [10..13] hot counts of set X (particles handed back by the cold kernels), [14] ionisations of the generation,
// [15] their high-water mark, [16] snapshot records of the batch, [17..20] hot heads, [21..22] cold heads
#define QC_HOT(b) ((b) * N_SPECIES)
#define QC_COLD (2 * N_SPECIES)
#define QC_X (2 * N_SPECIES + 2)
#define QC_ION (3 * N_SPECIES + 2)
#define QC_SNAP (3 * N_SPECIES + 4)
#define QC_HEAD (3 * N_SPECIES + 5)
#define QC_ELC(b) (4 * N_SPECIES + 7 + (b) * (N_ECLASS - 1))      // counts of the electron class queues 1.. of generation set b
#define QC_HEADC (4 * N_SPECIES + 7 + 2 * (N_ECLASS - 1))          // their heads
#define QC_ELW(b) (4 * N_SPECIES + 7 + 3 * (N_ECLASS - 1) + (b))   // warm electrons of generation set b
#define QC_HEADW (4 * N_SPECIES + 9 + 3 * (N_ECLASS - 1))
#define QC_VBW(b) (4 * N_SPECIES + 10 + 3 * (N_ECLASS - 1) + (b))  // warm valence holes of generation set b
#define QC_HEADWH (4 * N_SPECIES + 12 + 3 * (N_ECLASS - 1))
#define QC_TOTAL (4 * N_SPECIES + 13 + 3 * (N_ECLASS - 1))
#define QC_OVERFLOW QC_TOTAL                                         // sticky flag: some queue of a generation ran over its capacity
#define QC_WORDS (QC_TOTAL + 1)
// capacities of the per-generation queues, in the order the overflow check of k_gen_reset reads them
struct GenCaps { uint32_t hot[N_SPECIES], cls[N_ECLASS - 1], warm_e, warm_h; };
// Start of a generation (input set `cur`, output set `nxt`): the fill of the input set is final now -> a counter beyond its
// queue's capacity raises the sticky overflow flag (the host looks at it when it next reads the counters; records beyond
// the capacity were dropped, the batch is then rolled back and re-run with larger queues).  Then the counters of the set
// that is about to be filled and all queue heads are cleared: one launch instead of eight memsets.
__global__ void k_gen_reset(uint32_t *qc, int cur, int nxt, GenCaps caps) {
    const int t = threadIdx.x;
    bool over = false;
    if (t < N_SPECIES) over |= qc[QC_HOT(cur) + t] > caps.hot[t];
    if (t < N_ECLASS - 1) over |= qc[QC_ELC(cur) + t] > caps.cls[t];
    if (t == 0) over |= qc[QC_ELW(cur)] > caps.warm_e || qc[QC_VBW(cur)] > caps.warm_h;
    if (over) qc[QC_OVERFLOW] = 1u;
    if (t < N_SPECIES) { qc[QC_HOT(nxt) + t] = 0; qc[QC_HEAD + t] = 0; }
    if (t < N_ECLASS - 1) { qc[QC_ELC(nxt) + t] = 0; qc[QC_HEADC + t] = 0; }
    if (t == 0) { qc[QC_ELW(nxt)] = 0; qc[QC_HEADW] = 0; qc[QC_VBW(nxt)] = 0; qc[QC_HEADWH] = 0; }
}

namespace {
#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { eng->err = std::string(#call) + ": " + cudaGetErrorString(e_); return TRK3_E_CUDA; } } while (0)

// optional per-kernel-class timing: events are recorded on the launching stream around every launch
int prof_begin(trk3_engine *eng, int cls, cudaStream_t st = nullptr, uint32_t n = 0) {
    if (!st) st = eng->stream;
    if (!eng->opt_profile) return -1;
    size_t used = eng->ev_pending.size();
    if (used >= eng->ev_pool.size()) {
        cudaEvent_t a, b;
        if (cudaEventCreate(&a) != cudaSuccess || cudaEventCreate(&b) != cudaSuccess) return -1;
        eng->ev_pool.push_back({a, b});
    }
    cudaEventRecord(eng->ev_pool[used].first, st);
    eng->ev_pending.push_back({cls, (int)used});
    eng->ev_info.push_back({eng->cur_gen, n});
    return (int)used;
}
void prof_end(trk3_engine *eng, int idx, cudaStream_t st = nullptr) { if (idx >= 0) cudaEventRecord(eng->ev_pool[idx].second, st ? st : eng->stream); }
void prof_collect(trk3_engine *eng) {      // call after a stream synchronize
    for (size_t i = 0; i < eng->ev_pending.size(); ++i) {
        const auto &pe = eng->ev_pending[i];
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, eng->ev_pool[pe.second].first, eng->ev_pool[pe.second].second) == cudaSuccess) {
            eng->class_ms[pe.first] += ms; eng->class_launches[pe.first]++;
            if (eng->opt_profile >= 2) {
                float t0 = 0.f; cudaEventElapsedTime(&t0, eng->ev0, eng->ev_pool[pe.second].first);
                fprintf(stderr, "trace gen %3d class %d records %9u start %8.3f ms dur %8.3f ms\n", eng->ev_info[i].first, pe.first, eng->ev_info[i].second, t0, ms);
            }
        }
    }
    eng->ev_pending.clear(); eng->ev_info.clear();
}

template <class T>
int dev_alloc(trk3_engine *eng, T **p, size_t n) {
    *p = nullptr;
    if (n == 0) n = 1;
    CK(cudaMalloc((void **)p, n * sizeof(T)));
    eng->allocs.push_back((void *)*p);
    return TRK3_OK;
}
// table arrays: allocated by the first bind_tables(), re-used (same order, same sizes) by trk3_mc_reload_tables()
template <class T>
int tab_alloc(trk3_engine *eng, T **p, size_t n) {
    const size_t bytes = (n ? n : 1) * sizeof(T);
    if (eng->tab_cursor < eng->tab_allocs.size()) {
        auto &a = eng->tab_allocs[eng->tab_cursor++];
        if (a.second != bytes) { eng->err = "table shapes differ from those the engine was created with"; return TRK3_E_INVALID; }
        *p = (T *)a.first;
        return TRK3_OK;
    }
    if (!eng->tab_arena) {
        eng->tab_arena_cap = (size_t)192 << 20;        // tables + companions of the largest shipped material are ~80 MB; 180 GB of HBM
        int rc = dev_alloc(eng, &eng->tab_arena, eng->tab_arena_cap);
        if (rc) return rc;
    }
    const size_t at = (eng->tab_arena_used + 255) & ~(size_t)255;
    if (at + bytes <= eng->tab_arena_cap) { *p = (T *)(eng->tab_arena + at); eng->tab_arena_used = at + bytes; }
    else { int rc = dev_alloc(eng, p, n); if (rc) return rc; }      // does not fit: outside the arena (and outside the L2 window)
    eng->tab_allocs.push_back({(void *)*p, bytes}); eng->tab_cursor++;
    return TRK3_OK;
}
// Option l2_persist: the table arena as a persisting access-policy window on every stream of the engine.  The queues stream
// through L2 (108 B per record, written once, read once) and evict the tables that every collision looks up; with the window
// the table lines are kept (hitProp persisting) and everything else is treated as streaming.
int apply_l2_policy(trk3_engine *eng) {
    if (eng->l2_persist_applied == eng->opt_l2_persist) return TRK3_OK;
    eng->l2_persist_applied = eng->opt_l2_persist;
    int max_win = 0, max_persist = 0;
    CK(cudaDeviceGetAttribute(&max_win, cudaDevAttrMaxAccessPolicyWindowSize, eng->device));
    CK(cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, eng->device));
    cudaStreamAttrValue v; std::memset(&v, 0, sizeof v);
    if (eng->opt_l2_persist && eng->tab_arena && max_win > 0 && max_persist > 0) {
        const size_t bytes = std::min<size_t>(eng->tab_arena_used, (size_t)max_win);
        CK(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, std::min<size_t>(bytes, (size_t)max_persist)));
        v.accessPolicyWindow.base_ptr = eng->tab_arena; v.accessPolicyWindow.num_bytes = bytes;
        v.accessPolicyWindow.hitRatio = (float)std::min(1.0, (double)max_persist / (double)bytes);
        v.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting; v.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    } else { v.accessPolicyWindow.num_bytes = 0; v.accessPolicyWindow.hitProp = cudaAccessPropertyNormal; v.accessPolicyWindow.missProp = cudaAccessPropertyNormal; }
    cudaStream_t all[] = {eng->stream, eng->stream_c, eng->stream_w, eng->stream_wh, eng->stream_sp[0], eng->stream_sp[1], eng->stream_sp[2]};
    for (cudaStream_t st : all) if (st) CK(cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &v));
    if (!eng->opt_l2_persist) CK(cudaCtxResetPersistingL2Cache());
    return TRK3_OK;
}
template <class T>
int dev_upload(trk3_engine *eng, const T **dst, const T *src, size_t n) {
    T *d = nullptr;
    int rc = tab_alloc(eng, &d, n);
    if (rc) return rc;
    if (n && src) {
        const size_t bytes = n * sizeof(T);
        const char *dc = (const char *)d;
        if (eng->staging && dc >= eng->tab_arena && dc + bytes <= eng->tab_arena + eng->stage_cap) {
            const size_t at = (size_t)(dc - eng->tab_arena);
            std::memcpy(eng->stage + at, src, bytes);
            eng->stage_lo = std::min(eng->stage_lo, at); eng->stage_hi = std::max(eng->stage_hi, at + bytes);
        } else {
            CK(cudaMemcpyAsync(d, src, bytes, cudaMemcpyHostToDevice, eng->stream));
            eng->direct_pending = true;
        }
    }
    *dst = d;
    eng->h2d_bytes += n * sizeof(T);
    return TRK3_OK;
}
// the staged arrays collected since the last flush: one DMA out of the pinned mirror (the alignment gaps travel with them)
int stage_flush(trk3_engine *eng) {
    if (eng->stage_hi > eng->stage_lo) {
        CK(cudaMemcpyAsync(eng->tab_arena + eng->stage_lo, eng->stage + eng->stage_lo, eng->stage_hi - eng->stage_lo, cudaMemcpyHostToDevice, eng->stream));
        CK(cudaEventRecord(eng->ev_stage, eng->stream));
        eng->stage_in_flight = true;
    }
    eng->stage_lo = (size_t)-1; eng->stage_hi = 0;
    return TRK3_OK;
}
// a host temporary that was handed to dev_upload is about to go out of scope (or to be re-used)
int release_sources(trk3_engine *eng) {
    if (eng->direct_pending) { CK(cudaStreamSynchronize(eng->stream)); eng->direct_pending = false; }
    return TRK3_OK;
}
void dev_free(trk3_engine *eng, void *p) {
    if (!p) return;
    cudaFree(p);
    eng->allocs.erase(std::remove(eng->allocs.begin(), eng->allocs.end(), p), eng->allocs.end());
}

int alloc_queue(trk3_engine *eng, Queue &q, uint32_t cap, uint32_t *count) {
    q.cap = cap; q.count = count;
    for (int k = 0; k < TRK_NCOL; ++k) { int rc = dev_alloc(eng, &q.col[k], cap); if (rc) return rc; }
    int rc;
    if ((rc = dev_alloc(eng, &q.id, cap))) return rc;
    if ((rc = dev_alloc(eng, &q.ctr, cap))) return rc;
    if ((rc = dev_alloc(eng, &q.iter, cap))) return rc;
    if ((rc = dev_alloc(eng, &q.shell, cap))) return rc;
    return TRK3_OK;
}
void free_queue(trk3_engine *eng, Queue &q) {
    for (int k = 0; k < TRK_NCOL; ++k) { dev_free(eng, q.col[k]); q.col[k] = nullptr; }
    dev_free(eng, q.id); dev_free(eng, q.ctr); dev_free(eng, q.iter); dev_free(eng, q.shell);
    q.id = nullptr; q.ctr = nullptr; q.iter = nullptr; q.shell = nullptr; q.cap = 0;
}

// per-iteration capacities of the species queues (records of ONE generation)
void queue_caps(const trk3_engine *eng, double cap[N_QUEUES]) {
    const double n = eng->nel_est * eng->opt_cap_factor + 512.0;
    cap[SP_ELECTRON] = n; cap[SP_VBHOLE] = n; cap[SP_COREHOLE] = 0.5 * n + 256.0;
    cap[SP_PHOTON] = eng->cfg.include_photons ? 0.25 * n + 64.0 : 1.0;
    cap[Q_EL_COLD] = n; cap[Q_VB_COLD] = n;       // every carrier of an iteration ends up here once
    cap[Q_ION] = n;
    cap[Q_SNAP] = eng->opt_defer_snap ? 1.25 * n * (double)eng->lay.Nt : 1.0;     // every carrier at every grid time it lives to see
    // the higher energy classes of the hot electrons hold a few per cent of them (the spectrum falls like 1/E^2)
    for (int c = 1; c < N_ECLASS; ++c) cap[Q_ELC + c - 1] = (c < eng->opt_hot_classes) ? 0.25 * n + 64.0 : 0.0;
    cap[Q_ELW] = (eng->opt_warm_pinel > 0.0) ? n : 0.0;
    cap[Q_VBW] = (eng->opt_warm_pinel > 0.0) ? n : 0.0;
}
double queue_bytes_per_iteration(const trk3_engine *eng) {
    double cap[N_QUEUES]; queue_caps(eng, cap);
    double b = 0;
    for (int s = 0; s < N_SPECIES; ++s) b += 2.0 * cap[s] * (TRK_NCOL * 8 + 20);
    for (int s = N_SPECIES; s < Q_ELC; ++s) b += cap[s] * (TRK_NCOL * 8 + 20);
    for (int s = Q_ELC; s < N_QUEUES; ++s) b += 2.0 * cap[s] * (TRK_NCOL * 8 + 20);
    for (int s = 0; s < N_SPECIES; ++s) b += cap[s] / 16.0 * (TRK_NCOL * 8 + 20);
    return b;
}

// Everything the per-batch resources (queues, per-iteration scratch) are sized from, besides the batch size itself: a
// reload of configuration + tables that changes any of it (electron emission switched on -> em_spec, photons switched on
// -> photon queue, a longer time grid, larger cascades ...) must re-allocate them.
std::vector<double> batch_signature(const trk3_engine *eng) {
    double cap[N_QUEUES]; queue_caps(eng, cap);
    std::vector<double> sig(cap, cap + N_QUEUES);
    const ScratchLayout sl = scratch_layout(eng->hp, 1);
    sig.push_back((double)sl.u32_total); sig.push_back((double)sl.f64_total); sig.push_back((double)sl.em_spec);
    sig.push_back((double)eng->lay.Nt); sig.push_back((double)eng->hp.n_r); sig.push_back((double)eng->hp.n_dos);
    return sig;
}

int ensure_batch(trk3_engine *eng, uint32_t nb) {
    if (nb <= eng->nb_alloc) return TRK3_OK;
    // release the previous batch resources
    for (int b = 0; b < 2; ++b) for (int s = 0; s < N_SPECIES; ++s) free_queue(eng, eng->qs[b].q[s]);
    for (int b = 0; b < 2; ++b) for (int s = Q_ELC; s < N_QUEUES; ++s) free_queue(eng, eng->qs[b].q[s]);
    for (int s = N_SPECIES; s < Q_ELC; ++s) { free_queue(eng, eng->qs[0].q[s]); eng->qs[1].q[s] = Queue{}; eng->qs_x.q[s] = Queue{}; }
    for (int s = 0; s < N_SPECIES; ++s) free_queue(eng, eng->qs_x.q[s]);
    dev_free(eng, eng->d_u32); dev_free(eng, eng->d_f64);
    dev_free(eng, eng->fa.totnel); dev_free(eng, eng->fa.totE); dev_free(eng, eng->fa.latcum); dev_free(eng, eng->fa.emcnt); dev_free(eng, eng->fa.emE);
    eng->d_u32 = nullptr; eng->d_f64 = nullptr; eng->fa = FoldAux{};
    double cap[N_QUEUES]; queue_caps(eng, cap);
    for (int s = 0; s < N_QUEUES; ++s) if (cap[s] * (double)nb > 4.0e9) { eng->err = "queue capacity exceeds 2^32 records; lower the batch"; return TRK3_E_NOMEM; }
    for (int b = 0; b < 2; ++b) for (int s = 0; s < N_SPECIES; ++s) {
        int rc = alloc_queue(eng, eng->qs[b].q[s], (uint32_t)(cap[s] * (double)nb), eng->d_qcount + QC_HOT(b) + s);
        if (rc) return rc;
    }
    for (int s = N_SPECIES; s < Q_ELC; ++s) {      // the cold queues and the ionisation queue are shared by both generations
        int rc = alloc_queue(eng, eng->qs[0].q[s], (uint32_t)(cap[s] * (double)nb), eng->d_qcount + (s == Q_ION ? QC_ION : (s == Q_SNAP ? QC_SNAP : QC_COLD + (s - N_SPECIES))));
        if (rc) return rc;
        eng->qs[1].q[s] = eng->qs[0].q[s]; eng->qs_x.q[s] = eng->qs[0].q[s];
    }
    for (int b = 0; b < 2; ++b) for (int s = Q_ELC; s < N_QUEUES; ++s) {   // energy classes of the hot electrons, warm electrons (set X has none: cap 0)
        Queue &q = eng->qs[b].q[s];
        q = Queue{};
        if (cap[s] <= 0.0) continue;
        int rc = alloc_queue(eng, q, (uint32_t)(cap[s] * (double)nb), eng->d_qcount + (s == Q_VBW ? QC_VBW(b) : (s == Q_ELW ? QC_ELW(b) : QC_ELC(b) + (s - Q_ELC))));
        if (rc) return rc;
    }
    for (int s = 0; s < N_SPECIES; ++s) {             // handed back by the cold kernels: a rarity
        int rc = alloc_queue(eng, eng->qs_x.q[s], (uint32_t)(cap[s] * (double)nb / 16.0) + 4096u, eng->d_qcount + QC_X + s);
        if (rc) return rc;
    }
    eng->sl = scratch_layout(eng->hp, nb);
    int rc;
    if ((rc = dev_alloc(eng, &eng->d_u32, eng->sl.u32_total))) return rc;
    if ((rc = dev_alloc(eng, &eng->d_f64, eng->sl.f64_total))) return rc;
    const size_t n = (size_t)nb * eng->lay.Nt;
    if ((rc = dev_alloc(eng, &eng->fa.totnel, n))) return rc;
    if ((rc = dev_alloc(eng, &eng->fa.totE, n))) return rc;
    if ((rc = dev_alloc(eng, &eng->fa.latcum, n))) return rc;
    if ((rc = dev_alloc(eng, &eng->fa.emcnt, n))) return rc;
    if ((rc = dev_alloc(eng, &eng->fa.emE, n))) return rc;
    bind_scratch(eng->hp, eng->sl, eng->d_u32, eng->d_f64);
    eng->nb_alloc = nb;
    eng->batch_sig = batch_signature(eng);
    return TRK3_OK;
}

// The lean kernels serve the default switches (see DevCtxT); any other setting runs the kernels with everything compiled in.
inline bool engine_is_lean(const trk3_engine *eng) {
    return eng->opt_lean && eng->cfg.kind_of_EMFP == 1 && !(eng->cfg.work_function > 0.0) && eng->opt_defer_snap;
}
// Grids are persistent-style and sized for the device, not for the work: the number of records is only known on the device
// (queue counters) when the host enqueues the launch; blocks without work leave at once.
template <int SP, bool COLD>
int launch_wave(trk3_engine *eng, const Queue &qin, uint32_t *head, const QueueSet &qout, cudaStream_t st = nullptr, int warm = 0, uint32_t n_hint = 0, uint32_t first = 0) {
    if (!st) st = eng->stream;
    const bool lean = engine_is_lean(eng);
    size_t smem = (eng->opt_use_smem && eng->hp.s_total > 0) ? (size_t)(lean ? eng->hp.s_len[TRK3_OUT_ELAT] : eng->hp.s_total) * sizeof(double) : 8;
    int use_smem = eng->opt_use_smem;
    const size_t smem_max = (size_t)eng->smem_optin - 1024;             // static shared memory + driver reserve
    if (smem > smem_max) { smem = 8; use_smem = 0; }                    // too many output times for shared memory: global atomics
    auto kern = lean ? k_wave<SP, COLD, true> : k_wave<SP, COLD, false>;
    if (smem > 48 * 1024) CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int bps = eng->opt_blocks_per_sm;
    const int block = eng->opt_block;
    if (bps <= 0) { CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, kern, block, smem)); if (bps < 1) bps = 1; }
    // one resident wave of blocks, never more than the queue could hold records for
    const uint32_t want = (qin.cap + block - 1) / block;
    uint32_t grid = std::min<uint32_t>(want, (uint32_t)(eng->n_sm * bps));
    if (grid < 1) grid = 1;
    const int pi = prof_begin(eng, warm ? N_SPECIES + 4 + SP : (COLD ? N_SPECIES + 2 + SP : SP), st, n_hint);
    kern<<<grid, block, smem, st>>>(TRK_PA(eng), qin, first, head, qout, use_smem, eng->opt_refill_min, warm ? eng->opt_warm_slice : eng->opt_hot_slice, warm);
    prof_end(eng, pi, st);
    CK(cudaGetLastError());
    eng->launches++;
    return TRK3_OK;
}
// One generation of hot carriers of species SP: `ncls` class queues (valence holes: 1), see HotIn / hot_plan.
template <int SP>
int launch_hot(trk3_engine *eng, const Queue *const *qin, uint32_t *const *head, int ncls, const QueueSet &qout, cudaStream_t st = nullptr, uint32_t n_hint = 0) {
    if (!st) st = eng->stream;
    const bool lean = engine_is_lean(eng);
    size_t smem = (eng->opt_use_smem && eng->hp.s_total > 0) ? (size_t)(lean ? eng->hp.s_len[TRK3_OUT_ELAT] : eng->hp.s_total) * sizeof(double) : 8;
    int use_smem = eng->opt_use_smem;
    const size_t smem_max = (size_t)eng->smem_optin - 1024;
    if (smem > smem_max) { smem = 8; use_smem = 0; }
    auto kern = lean ? k_hot<SP, true> : k_hot<SP, false>;
    if (smem > 48 * 1024) CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int bps = eng->opt_blocks_per_sm;
    const int block = eng->opt_hot_block ? eng->opt_hot_block : eng->opt_block;
    if (bps <= 0) { CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, kern, block, smem)); if (bps < 1) bps = 1; }
    HotIn in{};
    in.ncls = ncls; in.spread = eng->opt_spread; in.quota_min = eng->opt_quota_min; in.coop = eng->opt_coop; in.weighted = eng->opt_weighted;
    for (int c = 0; c < ncls; ++c) {
        in.q[c] = *qin[c]; in.head[c] = head[c];
        in.qmax[c] = (ncls > 1) ? eng->opt_class_quota[c] : 32;
        in.slice[c] = (ncls > 1) ? eng->opt_class_slice[c] : eng->opt_hot_slice;
    }
    const uint32_t grid = (uint32_t)(eng->n_sm * bps);               // the warps the GPU holds at once: hot_plan shares them out
    const int pi = prof_begin(eng, SP, st, n_hint);
    kern<<<grid, block, smem, st>>>(TRK_PA(eng), in, qout, use_smem, eng->opt_refill_min, eng->opt_coop);
    prof_end(eng, pi, st);
    CK(cudaGetLastError());
    eng->launches++;
    return TRK3_OK;
}
// the hot electrons of generation set `qs`: all energy classes in one launch
int launch_hot_electrons(trk3_engine *eng, const QueueSet &qs, const QueueSet &qout, uint32_t n_hint = 0) {
    const Queue *q[N_ECLASS]; uint32_t *head[N_ECLASS];
    uint32_t *heads = eng->d_qcount + QC_HEAD, *headc = eng->d_qcount + QC_HEADC;
    q[0] = &qs.q[SP_ELECTRON]; head[0] = heads + SP_ELECTRON;
    int ncls = 1;
    for (int c = 1; c < N_ECLASS; ++c) {
        q[c] = &qs.q[Q_ELC + c - 1]; head[c] = headc + (c - 1);
        if (q[c]->cap) ncls = c + 1; else break;
    }
    return launch_hot<SP_ELECTRON>(eng, q, head, ncls, qout, nullptr, n_hint);
}
int launch_hot_vbholes(trk3_engine *eng, const Queue &qin, const QueueSet &qout, cudaStream_t st = nullptr, uint32_t n_hint = 0) {
    const Queue *q[1] = {&qin}; uint32_t *head[1] = {eng->d_qcount + QC_HEAD + SP_VBHOLE};
    return launch_hot<SP_VBHOLE>(eng, q, head, 1, qout, st, n_hint);
}
// Flattened tables -> device (first call allocates, later calls re-use the arrays): the inputs of do_Monte_Carlo.
int bind_tables(trk3_engine *eng, const trk3_config *cfg, const trk3_tables *tab) {
    // everything of the previous DevP that does not come from the tables survives a reload
    const DevP old = eng->hp;
    eng->tab_cursor = 0; eng->h2d_bytes = 0;
    // re-binding into the arrays of an earlier binding: stage the uploads (see trk3_engine::stage)
    eng->staging = false; eng->direct_pending = false; eng->stage_lo = (size_t)-1; eng->stage_hi = 0;
    if (eng->opt_stage_uploads && !eng->tab_allocs.empty() && eng->tab_arena && eng->tab_arena_used) {
        if (eng->stage_cap < eng->tab_arena_used) {
            if (eng->stage) { CK(cudaStreamSynchronize(eng->stream)); cudaFreeHost(eng->stage); eng->stage = nullptr; eng->stage_cap = 0; eng->stage_in_flight = false; }
            if (cudaHostAlloc((void **)&eng->stage, eng->tab_arena_used, cudaHostAllocDefault) == cudaSuccess) eng->stage_cap = eng->tab_arena_used;
            else { eng->stage = nullptr; (void)cudaGetLastError(); }          // no pinned memory to be had: direct copies
        }
        if (!eng->ev_stage) CK(cudaEventCreateWithFlags(&eng->ev_stage, cudaEventDisableTiming));
        if (eng->stage) {
            if (eng->stage_in_flight) { CK(cudaEventSynchronize(eng->ev_stage)); eng->stage_in_flight = false; }   // the mirror is read by the previous binding's DMA
            eng->staging = true;
        }
    }
    eng->cfg = *cfg;
    int rc = trk3_tally_layout_init(cfg, tab, &eng->lay);
    if (rc != TRK3_OK) { eng->err = "invalid time grid / layout"; return rc; }
    rc = fill_devp_scalars(*cfg, *tab, eng->lay, eng->hp);
    if (rc != TRK3_OK) { eng->err = (rc == TRK3_E_UNSUPPORTED) ? "DSF elastic scattering (kind_of_EMFP = 2) without DSF tables" : "invalid tables"; return rc; }
    DevP &p = eng->hp;
    HostTotals tot; compute_totals(*tab, tot);
    const size_t NS = tab->n_shells;
#define UP(dst, src, n) do { rc = dev_upload(eng, &p.dst, src, (size_t)(n)); if (rc) return rc; } while (0)
    UP(ei_E, tab->ei_E, tab->n_ei); UP(ei_L, tab->ei_L, NS * tab->n_ei); UP(ei_tot, tot.ei_tot.data(), tab->n_ei);
    UP(ee_E, tab->ee_E, tab->n_ee); UP(ee_L, tab->ee_L, tab->n_ee);
    UP(hi_E, tab->hi_E, tab->n_hi); UP(hi_L, tab->hi_L, NS * tab->n_hi); UP(hi_tot, tot.hi_tot.data(), tab->n_hi);
    UP(he_E, tab->he_E, tab->n_he); UP(he_L, tab->he_L, tab->n_he);
    UP(ph_E, tab->ph_E, tab->n_ph); UP(ph_L, tab->ph_L, NS * tab->n_ph); UP(ph_tot, tot.ph_tot.data(), tab->n_ph);
    UP(shi_E, tab->shi_E, tab->n_shi); UP(shi_L, tab->shi_L, NS * tab->n_shi); UP(shi_tot, tot.shi_tot.data(), tab->n_shi);
    UP(dshi_off, tab->dshi_off, NS + 1); UP(dshi_E, tab->dshi_E, tab->dshi_off[NS]); UP(dshi_L, tab->dshi_L, tab->dshi_off[NS]);
    const size_t n_eid = NS * tab->n_ei;
    UP(eid_off, tab->eid_off, n_eid + 1); UP(eid_hw, tab->eid_hw, tab->eid_off[n_eid]); UP(eid_L, tab->eid_L, tab->eid_off[n_eid]);
    UP(eed_off, tab->eed_off, tab->n_ee + 1); UP(eed_hw, tab->eed_hw, tab->eed_off[tab->n_ee]); UP(eed_L, tab->eed_L, tab->eed_off[tab->n_ee]);
    UP(hid_off, tab->hid_off, tab->n_hi + 1); UP(hid_hw, tab->hid_hw, tab->hid_off[tab->n_hi]); UP(hid_L, tab->hid_L, tab->hid_off[tab->n_hi]);
    UP(hed_off, tab->hed_off, tab->n_he + 1); UP(hed_hw, tab->hed_hw, tab->hed_off[tab->n_he]); UP(hed_L, tab->hed_L, tab->hed_off[tab->n_he]);
    {   // rows of the electron differential tables that are non-increasing (see DevP::eid_mono)
        // evaluated on the device from the uploaded rows (k_monotone_rows, launched with the companions below)
        uint8_t *fe = nullptr, *fl = nullptr;
        if ((rc = tab_alloc(eng, &fe, n_eid))) return rc;
        if ((rc = tab_alloc(eng, &fl, (size_t)tab->n_ee))) return rc;
        p.eid_mono = fe; p.eed_mono = fl;
    }
    UP(dos_E, tab->dos_E, tab->n_dos); UP(dos_DOS, tab->dos_DOS, tab->n_dos); UP(dos_int, tab->dos_int, tab->n_dos); UP(dos_effm, tab->dos_effm, tab->n_dos);
    UP(out_R, tab->out_R, tab->n_r); UP(out_V, tab->out_V, tab->n_r);
    { const int n_osc = tab->delta_cdf ? tab->osc_off[tab->n_shells] : 0;        // delta-function CDF (kind_of_DR = 4)
      UP(osc_E0, tab->osc_E0, n_osc); UP(osc_alpha, tab->osc_alpha, n_osc); }
    { const bool dsf = cfg->kind_of_EMFP == 2;                                     // DSF elastic scattering (kind_of_EMFP = 2)
      const size_t ne = dsf ? (size_t)tab->n_ee : 0, nh = dsf ? (size_t)tab->n_he : 0;
      const size_t re = ne * (size_t)tab->n_dsf_e, rh = nh * (size_t)tab->n_dsf_h;
      UP(dsf_e_dE, tab->dsf_e_dE, re); UP(dsf_e_emit, tab->dsf_e_emit, re); UP(dsf_e_absorb, tab->dsf_e_absorb, re); UP(ee_emit, tab->ee_emit, ne); UP(ee_absorb, tab->ee_absorb, ne);
      UP(dsf_h_dE, tab->dsf_h_dE, rh); UP(dsf_h_emit, tab->dsf_h_emit, rh); UP(dsf_h_absorb, tab->dsf_h_absorb, rh); UP(he_emit, tab->he_emit, nh); UP(he_absorb, tab->he_absorb, nh); }
#undef UP
    {   // log / reciprocal companions (one exp() per log-log interpolation instead of five log() + exp())
        const trk3_tables &T = *tab;
        if ((rc = stage_flush(eng))) return rc;          // the companions are computed from the arrays uploaded so far
        {   const size_t n_eid = NS * T.n_ei;
            if (n_eid) k_monotone_rows<<<(unsigned)((n_eid * 32 + 255) / 256), 256, 0, eng->stream>>>(p.eid_off, p.eid_L, n_eid, const_cast<uint8_t *>(p.eid_mono));
            if (T.n_ee) k_monotone_rows<<<(unsigned)(((size_t)T.n_ee * 32 + 255) / 256), 256, 0, eng->stream>>>(p.eed_off, p.eed_L, (size_t)T.n_ee, const_cast<uint8_t *>(p.eed_mono));
            CK(cudaGetLastError()); }
        CompanionJobs jobs; jobs.nj = 0;
        size_t n_max = 0;
        auto launch_companions = [&]() -> int {
            if (jobs.nj) k_companions<<<dim3((unsigned)std::min<size_t>((n_max + 255) / 256, (size_t)eng->n_sm * 2), (unsigned)jobs.nj), 256, 0, eng->stream>>>(jobs);
            jobs.nj = 0; n_max = 0;
            CK(cudaGetLastError());
            return TRK3_OK;
        };
#define X(dst, src, n, op) { double *d_ = nullptr; const size_t n_ = (size_t)(n); if ((rc = tab_alloc(eng, &d_, n_))) return rc; \
        if (n_) { if (jobs.nj == TRK_MAX_COMPANIONS && (rc = launch_companions())) return rc; \
                  jobs.j[jobs.nj++] = CompanionJob{d_, src, (unsigned long long)n_, op}; n_max = std::max(n_max, n_); } p.dst = d_; }
        TRK3_COMPANIONS(X, p, T, NS)
#undef X
        if ((rc = launch_companions())) return rc;
        std::vector<uint16_t> lut;
#define X(id, E, n) { build_lut(E, n, lut, p.lut[id].l0, p.lut[id].scale); rc = dev_upload(eng, &p.lut[id].lut, lut.data(), lut.size()); if (rc) return rc; if ((rc = release_sources(eng))) return rc; }
        TRK3_LUT_GRIDS(X, T)
#undef X
        p.dos_inv_step = uniform_inv_step(T.dos_E, T.n_dos);
        for (int sh = 0; sh < T.n_shells; ++sh) {
            int Mt; double dl; const int Nsh = (int)(T.dshi_off[sh + 1] - T.dshi_off[sh]);
            shi_threshold(T.dshi_E + T.dshi_off[sh], T.dshi_L + T.dshi_off[sh], Nsh, T.shell_Ip[sh], Mt, dl);
            build_inverse_lut(T.dshi_L + T.dshi_off[sh], Nsh, Mt, lut, p.dshi_lut[sh].l0, p.dshi_lut[sh].scale);
            rc = dev_upload(eng, &p.dshi_lut[sh].lut, lut.data(), lut.size()); if (rc) return rc;
            if ((rc = release_sources(eng))) return rc;
        }
        if ((rc = stage_flush(eng))) return rc;
    for (int sh = 0; sh < T.n_shells; ++sh) shi_threshold(T.dshi_E + T.dshi_off[sh], T.dshi_L + T.dshi_off[sh], (int)(T.dshi_off[sh + 1] - T.dshi_off[sh]), T.shell_Ip[sh], p.shi_Mtemp[sh], p.shi_dL[sh]);
        cold_range(tab->ei_E, tot.ei_tot.data(), tab->n_ei, p.e_cold, p.e_imfp_cold);
        cold_range(tab->hi_E, tot.hi_tot.data(), tab->n_hi, p.h_cold, p.h_imfp_cold);
        p.e_iimfp_cold = (p.e_imfp_cold > 0.0) ? 1.0 / p.e_imfp_cold : 0.0; p.h_iimfp_cold = (p.h_imfp_cold > 0.0) ? 1.0 / p.h_imfp_cold : 0.0;
    }
    eng->warm_E.assign(tab->ei_E, tab->ei_E + tab->n_ei); eng->warm_P.assign(tab->n_ei, 1.0);
    for (int i = 0; i < tab->n_ei; ++i) {       // ionisation probability per collision on the inelastic grid (scheduling only)
        const double E = tab->ei_E[i], li = tot.ei_tot[i];
        int j = (int)(std::upper_bound(tab->ee_E, tab->ee_E + tab->n_ee, E) - tab->ee_E);
        j = std::min(std::max(j, 1), tab->n_ee - 1);
        const double f = (E - tab->ee_E[j - 1]) / (tab->ee_E[j] - tab->ee_E[j - 1]);
        const double le = tab->ee_L[j - 1] + std::min(1.0, std::max(0.0, f)) * (tab->ee_L[j] - tab->ee_L[j - 1]);
        const double ii = (li > 0.0 && li < 1e15) ? 1.0 / li : 0.0, ie = (le > 0.0 && le < 1e15) ? 1.0 / le : 0.0;
        eng->warm_P[i] = (ii + ie > 0.0) ? ii / (ii + ie) : 0.0;
    }
    eng->e_warm_auto = -1.0;
    eng->warm_Eh.assign(tab->hi_E, tab->hi_E + tab->n_hi); eng->warm_Ph.assign(tab->n_hi, 1.0);
    for (int i = 0; i < tab->n_hi; ++i) {       // the same for valence holes
        const double E = tab->hi_E[i], li = tot.hi_tot[i];
        int j = (int)(std::upper_bound(tab->he_E, tab->he_E + tab->n_he, E) - tab->he_E);
        j = std::min(std::max(j, 1), tab->n_he - 1);
        const double f = (E - tab->he_E[j - 1]) / (tab->he_E[j] - tab->he_E[j - 1]);
        const double le = tab->he_L[j - 1] + std::min(1.0, std::max(0.0, f)) * (tab->he_L[j] - tab->he_L[j - 1]);
        const double ii = (li > 0.0 && li < 1e15) ? 1.0 / li : 0.0, ie = (le > 0.0 && le < 1e15) ? 1.0 / le : 0.0;
        eng->warm_Ph[i] = (ii + ie > 0.0) ? ii / (ii + ie) : 0.0;
    }
    eng->h_warm_auto = -1.0;
    // a first binding ends synchronised; a staged re-binding leaves its DMAs and the companion kernel in flight on the engine's
    // stream, in front of the run that follows (nothing on the host refers to the caller's arrays any more)
    if (!eng->staging) CK(cudaStreamSynchronize(eng->stream));
    p.tally = old.tally; p.events = old.events; p.errors = old.errors; p.cnt_el = old.cnt_el; p.cnt_ph = old.cnt_ph; p.it = old.it;

    eng->nel_est = estimate_nel(*cfg, *tab);
    return TRK3_OK;
}
}  // namespace

extern "C" {

const char *trk3_gpu_version(void) { return "trekis3_gpu 0.1 (sm_100a wavefront Monte-Carlo engine)"; }

int trk3_mc_create(const trk3_config *cfg, const trk3_tables *tab, int device, trk3_engine **out) {
    if (!cfg || !tab || !out) return TRK3_E_INVALID;
    *out = nullptr;
    trk3_engine *eng = new trk3_engine();
    *out = eng;                                   // returned even on failure so that the caller can read last_error
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) { eng->err = "no CUDA device: the engine has no CPU fallback"; return TRK3_E_CUDA; }
    if (device < 0) { if (cudaGetDevice(&device) != cudaSuccess) device = 0; }
    eng->device = device;
    CK(cudaSetDevice(device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    eng->n_sm = prop.multiProcessorCount;
    eng->smem_optin = (int)prop.sharedMemPerBlockOptin;
    {   // the hot cascade is the critical path: its stream gets the higher priority
        int lo = 0, hi = 0;
        CK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        CK(cudaStreamCreateWithPriority(&eng->stream, cudaStreamNonBlocking, hi));
        CK(cudaStreamCreateWithPriority(&eng->stream_c, cudaStreamNonBlocking, lo));
        for (int i = 0; i < 3; ++i) { CK(cudaStreamCreateWithPriority(&eng->stream_sp[i], cudaStreamNonBlocking, hi)); CK(cudaEventCreateWithFlags(&eng->ev_sp[i], cudaEventDisableTiming)); }
        CK(cudaEventCreateWithFlags(&eng->ev_gen, cudaEventDisableTiming));
        CK(cudaStreamCreateWithPriority(&eng->stream_w, cudaStreamNonBlocking, hi)); CK(cudaEventCreateWithFlags(&eng->ev_w, cudaEventDisableTiming));
        CK(cudaStreamCreateWithPriority(&eng->stream_wh, cudaStreamNonBlocking, hi)); CK(cudaEventCreateWithFlags(&eng->ev_wh, cudaEventDisableTiming));
    }
    CK(cudaEventCreate(&eng->ev0)); CK(cudaEventCreate(&eng->ev1));
    CK(cudaEventCreateWithFlags(&eng->ev_fork, cudaEventDisableTiming)); CK(cudaEventCreateWithFlags(&eng->ev_join, cudaEventDisableTiming));
    int rc = bind_tables(eng, cfg, tab);
    if (rc) return rc;
    // the pinned mirror of the table arena for the re-bindings of this handle (see trk3_engine::stage): allocated here, not in the
    // first trk3_mc_reload_tables call, where it would cost that call milliseconds; without it re-bindings copy array by array
    if (eng->opt_stage_uploads && eng->tab_arena_used && cudaHostAlloc((void **)&eng->stage, eng->tab_arena_used, cudaHostAllocDefault) == cudaSuccess) eng->stage_cap = eng->tab_arena_used;
    else { eng->stage = nullptr; (void)cudaGetLastError(); }
    DevP &p = eng->hp;
    if ((rc = dev_alloc(eng, &eng->d_tally, (size_t)eng->lay.total))) return rc;
    CK(cudaMemset(eng->d_tally, 0, (size_t)eng->lay.total * sizeof(double)));
    if ((rc = dev_alloc(eng, &eng->d_tally_bak, (size_t)eng->lay.total))) return rc;
    if ((rc = dev_alloc(eng, &eng->d_small, (size_t)TRK3_MAX_NT))) return rc;
    if ((rc = dev_alloc(eng, &eng->d_counters, (size_t)(TRK3_N_EVENT_CLASSES + TRK3_N_ERRORS + 6)))) return rc;
    if ((rc = dev_alloc(eng, &eng->d_counters_bak, (size_t)(TRK3_N_EVENT_CLASSES + TRK3_N_ERRORS + 6)))) return rc;
    if ((rc = dev_alloc(eng, &eng->d_qcount, (size_t)QC_WORDS))) return rc;

    p.tally = eng->d_tally;
    p.events = eng->d_counters; p.errors = eng->d_counters + TRK3_N_EVENT_CLASSES;
    p.cnt_el = eng->d_counters + TRK3_N_EVENT_CLASSES + TRK3_N_ERRORS; p.cnt_ph = p.cnt_el + 1;
    return TRK3_OK;
}

// Re-upload configuration + tables into an existing engine (same shapes): the per-call input copy of a persistent
// plugin handle; queues and scratch stay allocated.
int trk3_mc_reload_tables(trk3_engine *eng, const trk3_config *cfg, const trk3_tables *tab) {
    if (!eng || !cfg || !tab) return TRK3_E_INVALID;
    CK(cudaSetDevice(eng->device));
    const trk3_tally_layout lay0 = eng->lay;
    const double nel0 = eng->nel_est;
    int rc = bind_tables(eng, cfg, tab);
    if (rc) return rc;
    if (eng->lay.total != lay0.total || eng->lay.Nt != lay0.Nt) { eng->err = "tally layout differs from the one the engine was created with"; return TRK3_E_INVALID; }
    if (eng->nel_est > nel0) eng->nb_alloc = 0;          // larger cascades expected: re-size the queues at the next run
    // the scratch layout and the queue capacities follow the configuration (work_function > 0 -> emission spectrum,
    // include_photons -> photon queue): if the new inputs need other sizes, the next run re-allocates
    if (eng->nb_alloc && batch_signature(eng) != eng->batch_sig) eng->nb_alloc = 0;
    if (eng->nb_alloc) bind_scratch(eng->hp, eng->sl, eng->d_u32, eng->d_f64);
    return TRK3_OK;
}
uint64_t trk3_mc_table_bytes(const trk3_engine *eng) { return eng ? eng->h2d_bytes : 0; }

int trk3_mc_set_option(trk3_engine *eng, const char *name, double v) {
    if (!eng || !name) return TRK3_E_INVALID;
    std::string k(name);
    if (k == "batch") eng->opt_batch = std::max(1, (int)v);
    else if (k == "use_smem") eng->opt_use_smem = (v != 0.0);
    else if (k == "refill_min") eng->opt_refill_min = std::min(32, std::max(1, (int)v));
    else if (k == "blocks_per_sm") eng->opt_blocks_per_sm = (int)v;
    else if (k == "cap_factor") { eng->opt_cap_factor = std::max(0.1, v); eng->nb_alloc = 0; }
    else if (k == "queue_gib") eng->opt_queue_bytes_max = (size_t)(v * (double)(1ull << 30));
    else if (k == "block") eng->opt_block = std::min(TRK_BLOCK_MAX, std::max(32, ((int)v / 32) * 32));
    else if (k == "hot_slice") eng->opt_hot_slice = std::max(1, (int)v);
    else if (k == "shi_lanes") eng->opt_shi_lanes = std::min(32, std::max(0, (int)v));
    else if (k == "defer_snap") { eng->opt_defer_snap = (v != 0.0); eng->nb_alloc = 0; }
    else if (k == "species_streams") eng->opt_species_streams = (v != 0.0);
    else if (k == "spread") eng->opt_spread = (v != 0.0);
    else if (k == "quota_min") eng->opt_quota_min = std::min(32, std::max(1, (int)v));
    else if (k == "coop") eng->opt_coop = (v != 0.0);
    else if (k == "weighted") eng->opt_weighted = (v != 0.0);
    else if (k == "l2_persist") eng->opt_l2_persist = (v != 0.0);
    else if (k == "stage_uploads") eng->opt_stage_uploads = (v != 0.0);
    else if (k == "run_ahead") eng->opt_run_ahead = std::min(6, std::max(0, (int)v));
    else if (k == "warm_pinel") { eng->opt_warm_pinel = std::min(0.99, std::max(0.0, v)); eng->nb_alloc = 0; eng->e_warm_auto = -1.0; eng->h_warm_auto = -1.0; }
    else if (k == "warm_holes") eng->opt_warm_holes = (v != 0.0);
    else if (k == "warm_slice") eng->opt_warm_slice = std::max(1, (int)v);
    else if (k == "lean") eng->opt_lean = (v != 0.0);
    else if (k == "hot_classes") { eng->opt_hot_classes = std::min(N_ECLASS, std::max(1, (int)v)); eng->nb_alloc = 0; }
    else if (k == "class_E1") eng->opt_class_E[0] = v;
    else if (k == "class_E2") eng->opt_class_E[1] = v;
    else if (k == "class_E3") eng->opt_class_E[2] = v;
    else if (k == "class_s0") eng->opt_class_slice[0] = std::max(1, (int)v);
    else if (k == "class_s1") eng->opt_class_slice[1] = std::max(1, (int)v);
    else if (k == "class_s2") eng->opt_class_slice[2] = std::max(1, (int)v);
    else if (k == "class_s3") eng->opt_class_slice[3] = std::max(1, (int)v);
    else if (k == "class_q0") eng->opt_class_quota[0] = std::min(32, std::max(1, (int)v));
    else if (k == "class_q1") eng->opt_class_quota[1] = std::min(32, std::max(1, (int)v));
    else if (k == "class_q2") eng->opt_class_quota[2] = std::min(32, std::max(1, (int)v));
    else if (k == "class_q3") eng->opt_class_quota[3] = std::min(32, std::max(1, (int)v));
    else if (k == "hot_block") eng->opt_hot_block = std::min(TRK_BLOCK_MAX, std::max(0, ((int)v / 32) * 32));
    else if (k == "max_generations") eng->opt_max_generations = std::max(1, (int)v);
    else if (k == "profile") { eng->opt_profile = (int)v; for (auto &x : eng->class_ms) x = 0; for (auto &x : eng->class_launches) x = 0; }
    else return TRK3_E_INVALID;
    return TRK3_OK;
}

const trk3_tally_layout *trk3_mc_layout(const trk3_engine *eng) { return eng ? &eng->lay : nullptr; }
const char *trk3_mc_last_error(const trk3_engine *eng) { return eng ? eng->err.c_str() : "no engine"; }
double *trk3_mc_device_tallies(trk3_engine *eng) { return eng ? eng->d_tally : nullptr; }

int trk3_mc_zero_device_tallies(trk3_engine *eng) {
    if (!eng) return TRK3_E_INVALID;
    CK(cudaSetDevice(eng->device));
    CK(cudaMemsetAsync(eng->d_tally, 0, (size_t)eng->lay.total * sizeof(double), eng->stream));
    CK(cudaStreamSynchronize(eng->stream));
    return TRK3_OK;
}

int trk3_mc_download_tallies(trk3_engine *eng, double *dst) {
    if (!eng || !dst) return TRK3_E_INVALID;
    CK(cudaSetDevice(eng->device));
    std::vector<double> tmp((size_t)eng->lay.total);
    CK(cudaMemcpyAsync(tmp.data(), eng->d_tally, tmp.size() * sizeof(double), cudaMemcpyDeviceToHost, eng->stream));
    CK(cudaStreamSynchronize(eng->stream));
    for (size_t i = 0; i < tmp.size(); ++i) dst[i] += tmp[i];
    return TRK3_OK;
}


static int run_device_impl(trk3_engine *eng, int64_t it_begin, int64_t it_end, trk3_stats *stats);
int trk3_mc_run_device(trk3_engine *eng, int64_t it_begin, int64_t it_end, trk3_stats *stats) {
    if (!eng || it_end < it_begin || it_begin < 0 || it_end > 0xffffffffll) return TRK3_E_INVALID;
    if (eng->failed) { eng->err = "the previous run failed (" + eng->err + "): call trk3_mc_reset first"; return TRK3_E_INVALID; }
    const int rc = run_device_impl(eng, it_begin, it_end, stats);
    if (rc != TRK3_OK) eng->failed = true;          // queues / counters / tallies are in an undefined state now
    return rc;
}
static int run_device_impl(trk3_engine *eng, int64_t it_begin, int64_t it_end, trk3_stats *stats) {
#ifdef TRK_CONST_SYMBOL
    static std::mutex g_device_mutex[64];
    std::lock_guard<std::mutex> device_turn(g_device_mutex[eng->device % 64]);
#endif
    CK(cudaSetDevice(eng->device));
    { int rcp = apply_l2_policy(eng); if (rcp) return rcp; }
    const int Nt = eng->lay.Nt;
    const int64_t n_it = it_end - it_begin;
    eng->iter_totE.assign((size_t)n_it * Nt, 0.0);
    eng->Dcoef.assign(Nt, 0.0);
    // batch size: bounded by the option and by the queue-memory budget
    int64_t nb_max = std::min<int64_t>(eng->opt_batch, std::max<int64_t>(1, (int64_t)((double)eng->opt_queue_bytes_max / queue_bytes_per_iteration(eng))));
    nb_max = std::min<int64_t>(nb_max, std::max<int64_t>(n_it, 1));
    int rc = ensure_batch(eng, (uint32_t)nb_max);
    if (rc) return rc;
    const size_t n_counters = TRK3_N_EVENT_CLASSES + TRK3_N_ERRORS + 6;
    CK(cudaMemsetAsync(eng->d_counters, 0, n_counters * sizeof(unsigned long long), eng->stream));
    uint64_t waves = 0; const uint64_t launches0 = eng->launches;
    std::vector<double> h_diffS, h_totE; std::vector<uint32_t> h_diffN;
    std::vector<unsigned long long> h_c(n_counters, 0ull);
    uint32_t *heads = eng->d_qcount + QC_HEAD;
    CK(cudaEventRecord(eng->ev0, eng->stream));
    int retries = 0;
    for (int64_t b0 = it_begin; b0 < it_end;) {
        const uint32_t nb = (uint32_t)std::min<int64_t>(nb_max, it_end - b0);
        // snapshot of the accumulators: a batch whose queues overflow is rolled back and re-run with larger queues
        CK(cudaMemcpyAsync(eng->d_tally_bak, eng->d_tally, (size_t)eng->lay.total * sizeof(double), cudaMemcpyDeviceToDevice, eng->stream));
        CK(cudaMemcpyAsync(eng->d_counters_bak, eng->d_counters, n_counters * sizeof(unsigned long long), cudaMemcpyDeviceToDevice, eng->stream));
        eng->hp.batch_begin = (uint32_t)b0; eng->hp.batch_n = nb;
        eng->hp.defer_snap = eng->opt_defer_snap;
        if (eng->e_warm_auto < 0.0) {        // the lowest grid energy at which an ionisation is at least as likely as `warm_pinel`
            eng->e_warm_auto = eng->warm_E.empty() ? 0.0 : eng->warm_E.back();
            for (size_t i = 0; i < eng->warm_E.size(); ++i) if (eng->warm_E[i] >= eng->hp.e_cold && eng->warm_P[i] >= eng->opt_warm_pinel) { eng->e_warm_auto = eng->warm_E[i]; break; }
        }
        eng->hp.e_warm = (eng->opt_warm_pinel > 0.0) ? std::max(eng->e_warm_auto, eng->hp.e_cold) : eng->hp.e_cold;
        if (eng->h_warm_auto < 0.0) {
            eng->h_warm_auto = eng->warm_Eh.empty() ? 0.0 : eng->warm_Eh.back();
            for (size_t i = 0; i < eng->warm_Eh.size(); ++i) if (eng->warm_Eh[i] >= eng->hp.h_cold && eng->warm_Ph[i] >= eng->opt_warm_pinel) { eng->h_warm_auto = eng->warm_Eh[i]; break; }
        }
        eng->hp.h_warm = (eng->opt_warm_pinel > 0.0 && eng->opt_warm_holes) ? std::max(eng->h_warm_auto, eng->hp.h_cold) : eng->hp.h_cold;
        if (eng->opt_profile >= 2) fprintf(stderr, "e_cold %.3f e_warm %.3f eV, h_cold %.3f h_warm %.3f eV\n", eng->hp.e_cold, eng->hp.e_warm, eng->hp.h_cold, eng->hp.h_warm);
        for (int c = 1; c < N_ECLASS; ++c) eng->hp.e_class[c - 1] = (c < eng->opt_hot_classes) ? eng->opt_class_E[c - 1] : 1.0e300;

#if defined(TRK_CONST_SYMBOL) || defined(TRK_HYBRID)
        CK(cudaMemcpyToSymbolAsync(c_p, &eng->hp, sizeof(DevP), 0, cudaMemcpyHostToDevice, eng->stream));
#endif
        CK(cudaMemsetAsync(eng->d_u32, 0, eng->sl.u32_total * sizeof(uint32_t), eng->stream));
        CK(cudaMemsetAsync(eng->d_f64, 0, eng->sl.f64_total * sizeof(double), eng->stream));
        CK(cudaMemsetAsync(eng->d_qcount, 0, QC_WORDS * sizeof(uint32_t), eng->stream));
        { const int pi = prof_begin(eng, N_SPECIES);
          // collisions are staged in the (still unused) electron queue of the other generation
          const Queue &stage = eng->qs[1].q[SP_ELECTRON];
          const int shi_lanes = (eng->opt_shi_lanes > 0 || eng->hp.n_shells + 1 > 32) ? std::max(1, eng->opt_shi_lanes) : 0;
          const uint32_t shi_warps = shi_lanes ? (nb + shi_lanes - 1) / shi_lanes : nb;
          k_shi<<<(shi_warps + SHI_WARPS - 1) / SHI_WARPS, 32 * SHI_WARPS, 0, eng->stream>>>(TRK_PA(eng), stage, eng->qs[0], shi_lanes);
          k_shi_emit<<<eng->n_sm * 4, 256, 0, eng->stream>>>(TRK_PA(eng), stage, eng->qs[0]);
          prof_end(eng, pi); }
        CK(cudaGetLastError());
        eng->launches += 2;
        bool overflow = false;
        // ---- the hot cascade, generation by generation, WITHOUT the host in the loop.  Every kernel of a generation reads the
        // number of its records from the device counters, so the host enqueues generation after generation and only LOOKS at
        // the counters: after each generation an asynchronous copy lands in a pinned ring slot, the host examines the slots
        // that have arrived (never blocking unless it is `run_ahead` generations ahead of what it has seen) and stops
        // enqueuing when a slot shows an empty input set.  At most `run_ahead` + 1 empty generations are launched (a few tens
        // of microseconds of empty kernels) in exchange for ~12 host round trips per batch.
        const int depth = (eng->opt_profile >= 2) ? 0 : std::max(0, eng->opt_run_ahead);     // profile 2 prints the records per launch: synchronous
        const int R = depth + 2;
        if ((int)eng->ring_ev.size() < R || !eng->h_qcount) {
            if (eng->h_qcount) cudaFreeHost(eng->h_qcount);
            eng->h_qcount = nullptr;
            CK(cudaHostAlloc((void **)&eng->h_qcount, (size_t)R * QC_WORDS * sizeof(uint32_t), cudaHostAllocDefault));
            while ((int)eng->ring_ev.size() < R) { cudaEvent_t e; CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); eng->ring_ev.push_back(e); }
        }
        auto slot_of = [&](int g) { return eng->h_qcount + (size_t)(g % R) * QC_WORDS; };
        auto post_counts = [&](int g) -> int {      // the counters as they are before generation g runs
            CK(cudaMemcpyAsync(slot_of(g), eng->d_qcount, QC_WORDS * sizeof(uint32_t), cudaMemcpyDeviceToHost, eng->stream));
            CK(cudaEventRecord(eng->ring_ev[g % R], eng->stream));
            return TRK3_OK;
        };
        GenCaps caps{};
        auto set_caps = [&]() {
            for (int sp = 0; sp < N_SPECIES; ++sp) caps.hot[sp] = eng->qs[0].q[sp].cap;
            for (int c = 1; c < N_ECLASS; ++c) caps.cls[c - 1] = eng->qs[0].q[Q_ELC + c - 1].cap ? eng->qs[0].q[Q_ELC + c - 1].cap : 0xffffffffu;
            caps.warm_e = eng->qs[0].q[Q_ELW].cap ? eng->qs[0].q[Q_ELW].cap : 0xffffffffu;
            caps.warm_h = eng->qs[0].q[Q_VBW].cap ? eng->qs[0].q[Q_VBW].cap : 0xffffffffu;
        };
        set_caps();
        // total records of input set `cur` in a counter snapshot (0: the cascade has died out); sets the overflow flag
        auto hot_total = [&](const uint32_t *h, int cur, uint32_t n[6]) -> uint64_t {
            uint64_t total = 0;
            n[0] = h[QC_HOT(cur) + SP_ELECTRON];
            for (int c = 1; c < N_ECLASS; ++c) n[0] += h[QC_ELC(cur) + c - 1];
            n[1] = h[QC_HOT(cur) + SP_VBHOLE]; n[2] = h[QC_HOT(cur) + SP_COREHOLE]; n[3] = h[QC_HOT(cur) + SP_PHOTON];
            n[4] = h[QC_ELW(cur)]; n[5] = h[QC_VBW(cur)];
            for (int q = 0; q < 6; ++q) total += n[q];
            return total;
        };
        auto counters_overflow = [&](const uint32_t *h, int cur) {
            bool o = h[QC_OVERFLOW] != 0u;
            for (int sp = 0; sp < N_SPECIES; ++sp) o |= h[QC_HOT(cur) + sp] > eng->qs[cur].q[sp].cap;
            for (int c = 1; c < N_ECLASS; ++c) o |= eng->qs[cur].q[Q_ELC + c - 1].cap && h[QC_ELC(cur) + c - 1] > eng->qs[cur].q[Q_ELC + c - 1].cap;
            o |= eng->qs[cur].q[Q_ELW].cap && h[QC_ELW(cur)] > eng->qs[cur].q[Q_ELW].cap;
            o |= eng->qs[cur].q[Q_VBW].cap && h[QC_VBW(cur)] > eng->qs[cur].q[Q_VBW].cap;
            for (int q = 0; q < 2; ++q) o |= h[QC_COLD + q] > eng->qs[0].q[N_SPECIES + q].cap;
            o |= std::max(h[QC_ION], h[QC_ION + 1]) > eng->qs[0].q[Q_ION].cap;
            o |= h[QC_SNAP] > eng->qs[0].q[Q_SNAP].cap;
            for (int sp = 0; sp < N_SPECIES; ++sp) o |= h[QC_X + sp] > eng->qs_x.q[sp].cap;
            return o;
        };
        // one generation of the hot cascade: input set `cur`, output set `cur ^ 1` (time-sliced, see k_hot).  `n` = records per
        // kernel if the host happens to know them (profile trace only).
        auto enqueue_generation = [&](int cur, const uint32_t *n) -> int {
            const int nxt = cur ^ 1;
            static_assert(N_SPECIES <= 32 && N_ECLASS <= 32, "k_gen_reset uses one warp");
            k_gen_reset<<<1, 32, 0, eng->stream>>>(eng->d_qcount, cur, nxt, caps);
            eng->launches++;
            // the species of a generation are independent of each other: the (few) valence holes, core holes and photons run
            // on their own streams beside the electrons instead of lengthening the generation one after the other
            const bool par = eng->opt_species_streams != 0;
            cudaStream_t s1 = par ? eng->stream_sp[0] : eng->stream, s2 = par ? eng->stream_sp[1] : eng->stream, s3 = par ? eng->stream_sp[2] : eng->stream;
            cudaStream_t s4 = par ? eng->stream_w : eng->stream, s5 = par ? eng->stream_wh : eng->stream;
            const bool warm_e = eng->qs[cur].q[Q_ELW].cap != 0u, warm_h = eng->qs[cur].q[Q_VBW].cap != 0u && eng->hp.h_warm > eng->hp.h_cold;
            const bool photons = eng->cfg.include_photons != 0;
            if (par) { CK(cudaEventRecord(eng->ev_gen, eng->stream)); }
            int rc;
            if ((rc = launch_hot_electrons(eng, eng->qs[cur], eng->qs[nxt], n ? n[0] : 0))) return rc;
            if (par) CK(cudaStreamWaitEvent(s1, eng->ev_gen, 0));
            if ((rc = launch_hot_vbholes(eng, eng->qs[cur].q[SP_VBHOLE], eng->qs[nxt], s1, n ? n[1] : 0))) return rc;
            if (par) CK(cudaEventRecord(eng->ev_sp[0], s1));
            if (par) CK(cudaStreamWaitEvent(s2, eng->ev_gen, 0));
            if ((rc = launch_wave<SP_COREHOLE, false>(eng, eng->qs[cur].q[SP_COREHOLE], heads + SP_COREHOLE, eng->qs[nxt], s2, 0, n ? n[2] : 0))) return rc;
            if (par) CK(cudaEventRecord(eng->ev_sp[1], s2));
            if (warm_e) {       // warm electrons: the elastic-only kernel, one time slice per generation
                if (par) CK(cudaStreamWaitEvent(s4, eng->ev_gen, 0));
                if ((rc = launch_wave<SP_ELECTRON, true>(eng, eng->qs[cur].q[Q_ELW], eng->d_qcount + QC_HEADW, eng->qs[nxt], s4, 1, n ? n[4] : 0))) return rc;
                if (par) CK(cudaEventRecord(eng->ev_w, s4));
            }
            if (warm_h) {       // warm valence holes
                if (par) CK(cudaStreamWaitEvent(s5, eng->ev_gen, 0));
                if ((rc = launch_wave<SP_VBHOLE, true>(eng, eng->qs[cur].q[Q_VBW], eng->d_qcount + QC_HEADWH, eng->qs[nxt], s5, 1, n ? n[5] : 0))) return rc;
                if (par) CK(cudaEventRecord(eng->ev_wh, s5));
            }
            if (photons) {
                if (par) CK(cudaStreamWaitEvent(s3, eng->ev_gen, 0));
                if ((rc = launch_wave<SP_PHOTON, false>(eng, eng->qs[cur].q[SP_PHOTON], heads + SP_PHOTON, eng->qs[nxt], s3, 0, n ? n[3] : 0))) return rc;
                if (par) CK(cudaEventRecord(eng->ev_sp[2], s3));
            }
            if (par) {
                CK(cudaStreamWaitEvent(eng->stream, eng->ev_sp[0], 0));
                CK(cudaStreamWaitEvent(eng->stream, eng->ev_sp[1], 0));
                if (photons) CK(cudaStreamWaitEvent(eng->stream, eng->ev_sp[2], 0));
                if (warm_e) CK(cudaStreamWaitEvent(eng->stream, eng->ev_w, 0));
                if (warm_h) CK(cudaStreamWaitEvent(eng->stream, eng->ev_wh, 0));
            }
            {   // the pairs of this generation's impact ionisations join the next generation
                const int pi = prof_begin(eng, N_SPECIES + 1);       // timed with the finalisation kernels
                k_ion_emit<<<eng->n_sm * 4, 256, 0, eng->stream>>>(TRK_PA(eng), eng->qs[0].q[Q_ION], eng->qs[nxt]);
                k_ion_reset<<<1, 32, 0, eng->stream>>>(eng->d_qcount + QC_ION);
                prof_end(eng, pi);
                CK(cudaGetLastError());
                eng->launches += 2;
            }
            return TRK3_OK;
        };
        // run the cascade from input set `cur0` until it has died out; returns the set that is empty at the end
        auto run_cascade = [&](int cur0, int &cur_out) -> int {
            int enq = 0, seen = 0;          // generations enqueued / counter snapshots examined
            bool dead = false;
            int rc = post_counts(0);
            if (rc) return rc;
            while (!dead && !overflow && enq < eng->opt_max_generations) {
                uint32_t n[6]; bool have_n = false;
                while (seen <= enq) {
                    if (seen <= enq - depth) CK(cudaEventSynchronize(eng->ring_ev[seen % R]));
                    else if (cudaEventQuery(eng->ring_ev[seen % R]) != cudaSuccess) break;
                    const uint32_t *h = slot_of(seen);
                    const int cur = cur0 ^ (seen & 1);
                    if (counters_overflow(h, cur)) { overflow = true; break; }
                    const uint64_t total = hot_total(h, cur, n);
                    have_n = (seen == enq);
                    if (seen == 0 && h[QC_HOT(1) + SP_ELECTRON] > eng->qs[1].q[SP_ELECTRON].cap && cur0 == 0) { overflow = true; break; }     // staged ion collisions
                    if (total == 0) { dead = true; break; }
                    ++waves; ++seen;
                }
                if (dead || overflow) break;
                eng->cur_gen = enq;
                rc = enqueue_generation(cur0 ^ (enq & 1), have_n ? n : nullptr);
                if (rc) return rc;
                ++enq;
                if ((rc = post_counts(enq))) return rc;
            }
            cur_out = cur0 ^ (enq & 1);
            return TRK3_OK;
        };
        int cur = 0;
        rc = run_cascade(0, cur);
        if (rc) return rc;
        // ---- the cold kernels: launched once, after the hot cascade has died out, over ALL cold records of the batch
        uint32_t cold_done[2] = {0, 0};
        uint32_t *h_fin = slot_of(0);
        for (int round = 0; !overflow; ++round) {
            eng->cur_gen = 1000 + round;
            CK(cudaMemsetAsync(heads + N_SPECIES, 0, 2 * sizeof(uint32_t), eng->stream));
            rc = launch_wave<SP_ELECTRON, true>(eng, eng->qs[0].q[Q_EL_COLD], heads + Q_EL_COLD, eng->qs_x, nullptr, 0, 0, cold_done[0]); if (rc) return rc;
            rc = launch_wave<SP_VBHOLE, true>(eng, eng->qs[0].q[Q_VB_COLD], heads + Q_VB_COLD, eng->qs_x, nullptr, 0, 0, cold_done[1]); if (rc) return rc;
            // the one host round trip of a batch: did anything overflow, and did the cold kernels hand anything back?
            CK(cudaMemcpyAsync(h_fin, eng->d_qcount, QC_WORDS * sizeof(uint32_t), cudaMemcpyDeviceToHost, eng->stream));
            CK(cudaStreamSynchronize(eng->stream));
            if (counters_overflow(h_fin, cur)) { overflow = true; break; }
            uint64_t nx = 0;
            for (int sp = 0; sp < N_SPECIES; ++sp) nx += h_fin[QC_X + sp];
            if (!nx) break;                                        // cold kernels create no cold records: everything is done
            // a carrier that a cold kernel handed back (a valence hole lifted above the cold range by the level snapping):
            // one generation from set X into `nxt`, then the cascade goes on from there; the next cold round starts behind
            // the records that are done
            cold_done[0] = std::min(h_fin[QC_COLD], eng->qs[0].q[Q_EL_COLD].cap); cold_done[1] = std::min(h_fin[QC_COLD + 1], eng->qs[0].q[Q_VB_COLD].cap);
            const int nxt = cur ^ 1;
            k_gen_reset<<<1, 32, 0, eng->stream>>>(eng->d_qcount, cur, nxt, caps);
            eng->launches++;
            rc = launch_hot_electrons(eng, eng->qs_x, eng->qs[nxt]); if (rc) return rc;
            rc = launch_hot_vbholes(eng, eng->qs_x.q[SP_VBHOLE], eng->qs[nxt]); if (rc) return rc;
            k_ion_emit<<<eng->n_sm * 4, 256, 0, eng->stream>>>(TRK_PA(eng), eng->qs[0].q[Q_ION], eng->qs[nxt]);
            k_ion_reset<<<1, 32, 0, eng->stream>>>(eng->d_qcount + QC_ION);
            CK(cudaGetLastError());
            eng->launches += 2;
            CK(cudaMemsetAsync(eng->d_qcount + QC_X, 0, N_SPECIES * sizeof(uint32_t), eng->stream));
            rc = run_cascade(nxt, cur);
            if (rc) return rc;
        }
        if (eng->opt_defer_snap && !overflow) {
            // all histories of the batch have ended: turn the queued snapshot records into tallies (the kernel reads their number)
            size_t smem = (eng->opt_use_smem && eng->hp.s_total > 0) ? (size_t)eng->hp.s_total * sizeof(double) : 8;
            int use_smem = eng->opt_use_smem;
            if (smem > (size_t)eng->smem_optin - 1024) { smem = 8; use_smem = 0; }
            if (smem > 48 * 1024) CK(cudaFuncSetAttribute(k_snapshot, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            int bps = 0;
            CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, k_snapshot, 256, smem)); if (bps < 1) bps = 1;
            const uint32_t n_snap = h_fin[QC_SNAP];
            if (n_snap) {
                const uint32_t grid = std::min<uint32_t>((n_snap + 255u) / 256u, (uint32_t)(eng->n_sm * bps));
                const int pi = prof_begin(eng, N_SPECIES + 1);
                k_snapshot<<<grid, 256, smem, eng->stream>>>(TRK_PA(eng), eng->qs[0].q[Q_SNAP], eng->qs[0], use_smem);
                prof_end(eng, pi);
                CK(cudaGetLastError());
                eng->launches++;
            }
        }
        if (overflow) {
            if (++retries > 6) { eng->err = "particle queue overflow persists after 6 capacity doublings"; return TRK3_E_OVERFLOW; }
            CK(cudaMemcpyAsync(eng->d_tally, eng->d_tally_bak, (size_t)eng->lay.total * sizeof(double), cudaMemcpyDeviceToDevice, eng->stream));
            CK(cudaMemcpyAsync(eng->d_counters, eng->d_counters_bak, n_counters * sizeof(unsigned long long), cudaMemcpyDeviceToDevice, eng->stream));
            CK(cudaStreamSynchronize(eng->stream));
            eng->opt_cap_factor *= 2.0; eng->nb_alloc = 0;
            int64_t nb_new = std::min<int64_t>(nb_max, std::max<int64_t>(1, (int64_t)((double)eng->opt_queue_bytes_max / queue_bytes_per_iteration(eng))));
            rc = ensure_batch(eng, (uint32_t)nb_new);
            if (rc) return rc;
            nb_max = nb_new;
            continue;                // re-run from the same b0 (histories are keyed by the global iteration index: same result)
        }
        const int pf = prof_begin(eng, N_SPECIES + 1);
        k_iter_prefix<<<(nb + 127) / 128, 128, 0, eng->stream>>>(TRK_PA(eng), eng->fa);
        CK(cudaGetLastError());
        const int64_t njobs = fold_num_jobs(eng->hp);
        k_fold<<<(unsigned)((njobs * 32 + 255) / 256), 256, 0, eng->stream>>>(TRK_PA(eng), eng->fa, njobs);
        prof_end(eng, pf);
        CK(cudaGetLastError());
        eng->launches += 2;
        // per-iteration results back to the host: total energies (conservation check) and the
        // Out_diff_coeff recurrence (Monte_Carlo.f90:1094-1098), which is sequential over iterations
        const size_t n = (size_t)nb * Nt;
        h_diffS.resize(n); h_diffN.resize(n); h_totE.resize(n);
        CK(cudaMemcpyAsync(h_diffS.data(), eng->hp.it.diffS, n * sizeof(double), cudaMemcpyDeviceToHost, eng->stream));
        CK(cudaMemcpyAsync(h_diffN.data(), eng->hp.it.diffN, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, eng->stream));
        CK(cudaMemcpyAsync(h_totE.data(), eng->fa.totE, n * sizeof(double), cudaMemcpyDeviceToHost, eng->stream));
        // the run's counters travel with the last batch's results: no further host round trip after this one
        if (b0 + nb >= it_end) {
            CK(cudaMemcpyAsync(h_c.data(), eng->d_counters, n_counters * sizeof(unsigned long long), cudaMemcpyDeviceToHost, eng->stream));
            CK(cudaEventRecord(eng->ev1, eng->stream));
        }
        CK(cudaStreamSynchronize(eng->stream));
        for (uint32_t il = 0; il < nb; ++il) for (int i = 0; i < Nt; ++i) {
            const size_t o = (size_t)il * Nt + i;
            eng->Dcoef[i] = eng->Dcoef[i] + h_diffS[o];
            if (h_diffN[o] > 0) eng->Dcoef[i] = eng->Dcoef[i] / (double)h_diffN[o];
            eng->iter_totE[(size_t)(b0 - it_begin + il) * Nt + i] = h_totE[o];
        }
        b0 += nb;
    }
    CK(cudaMemcpyAsync(eng->d_small, eng->Dcoef.data(), Nt * sizeof(double), cudaMemcpyHostToDevice, eng->stream));
    k_axpy<<<1, 256, 0, eng->stream>>>(eng->d_tally + eng->lay.off[TRK3_OUT_DIFF_COEFF], eng->d_small, Nt);
    CK(cudaGetLastError());
    eng->launches++;
    if (eng->comm && eng->comm_size > 1) {
        // the single collective of the path (26 x MPI_Reduce in the reference, Monte_Carlo.f90:131-389): in place, on the
        // engine's stream, right behind the folding kernels -- no host synchronisation in between
        // TRK3_NCCL_PROBE=1 (diagnostic, synchronises): time the collective -- the first all-reduce includes the wait for the
        // slowest rank, a second one on a scratch buffer right behind it is the collective's own latency
        static const bool probe = std::getenv("TRK3_NCCL_PROBE") != nullptr;
        cudaEvent_t pe[3] = {nullptr, nullptr, nullptr};
        if (probe) { for (auto &e : pe) CK(cudaEventCreate(&e)); CK(cudaEventRecord(pe[0], eng->stream)); }
        const auto h0 = std::chrono::steady_clock::now();
        const int nrc = nccl_rt::api().AllReduce(eng->d_tally, eng->d_tally, (size_t)eng->lay.total, nccl_rt::kFloat64, nccl_rt::kSum, eng->comm, eng->stream);
        if (nrc != nccl_rt::kSuccess) { eng->err = std::string("ncclAllReduce: ") + nccl_rt::api().GetErrorString(nrc); return TRK3_E_CUDA; }
        if (probe) {
            const double host_us = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - h0).count();
            CK(cudaEventRecord(pe[1], eng->stream));
            nccl_rt::api().AllReduce(eng->d_tally_bak, eng->d_tally_bak, (size_t)eng->lay.total, nccl_rt::kFloat64, nccl_rt::kSum, eng->comm, eng->stream);
            CK(cudaEventRecord(pe[2], eng->stream));
            CK(cudaStreamSynchronize(eng->stream));
            float a = 0.f, b = 0.f, c = 0.f;
            cudaEventElapsedTime(&a, pe[0], pe[1]); cudaEventElapsedTime(&b, pe[1], pe[2]); cudaEventElapsedTime(&c, eng->ev1, pe[0]);
            fprintf(stderr, "[trk3 nccl probe] %d ranks: all-reduce of %lld doubles %.3f ms (with the wait for the slowest rank), repeated at once %.3f ms, host call %.1f us, folded tallies -> all-reduce %.3f ms\n",
                    eng->comm_size, (long long)eng->lay.total, a, b, host_us, c);
            for (auto &e : pe) cudaEventDestroy(e);
        }
    }
    // No synchronisation here: the last two operations (the Out_diff_coeff update and the all-reduce) stay in flight on the
    // engine's stream when the call returns, so that a rank does not sit in cudaStreamSynchronize until the slowest rank has
    // reached its all-reduce and the caller's next call is already queued behind them.  Whoever reads the tally buffer does so
    // in stream order (trk3_mc_download_tallies synchronises; a framework that owns the stream is ordered by it).
    // device_ms = the Monte-Carlo section up to the folded tallies.
    float ms = 0.f;
    if (it_end > it_begin) CK(cudaEventElapsedTime(&ms, eng->ev0, eng->ev1));
    prof_collect(eng);
    trk3_stats st; std::memset(&st, 0, sizeof st);
    for (int q = 0; q < TRK3_N_EVENT_CLASSES; ++q) st.events[q] = h_c[q];
    for (int q = 0; q < TRK3_N_ERRORS; ++q) st.errors[q] = h_c[TRK3_N_EVENT_CLASSES + q];
    st.n_electrons = h_c[TRK3_N_EVENT_CLASSES + TRK3_N_ERRORS]; st.n_photons = h_c[TRK3_N_EVENT_CLASSES + TRK3_N_ERRORS + 1];
    st.cold_events[0] = h_c[TRK3_N_EVENT_CLASSES + TRK3_N_ERRORS + 2]; st.cold_events[1] = h_c[TRK3_N_EVENT_CLASSES + TRK3_N_ERRORS + 3];
    st.warm_events[0] = h_c[TRK3_N_EVENT_CLASSES + TRK3_N_ERRORS + 4]; st.warm_events[1] = h_c[TRK3_N_EVENT_CLASSES + TRK3_N_ERRORS + 5];
    st.n_waves = waves; st.kernel_launches = eng->launches - launches0; st.device_ms = ms;
    static const double ev_bytes[TRK3_N_EVENT_CLASSES] = {176, 320, 144, 384, 208, 384, 280, 208, 248};
    for (int q = 0; q < TRK3_N_EVENT_CLASSES; ++q) st.algorithmic_bytes += ev_bytes[q] * (double)st.events[q];
    // energy conservation: after the ion has left, the total energy of an iteration must not change
    double drift = 0.0;
    for (int64_t k = 0; k < n_it; ++k) {
        const double *e = &eng->iter_totE[(size_t)k * Nt];
        const double ref = e[Nt - 1];
        if (!(ref > 0.0)) continue;
        for (int i = 0; i < Nt - 1; ++i) {
            // only grid times after the ion left the layer are comparable; the energy is non-decreasing before
            if (e[i] >= ref * (1.0 - 1e-6)) drift = std::max(drift, std::fabs(e[i] - ref) / ref);
        }
    }
    st.max_energy_drift = drift;
    if (stats) *stats = st;
    if (st.errors[TRK3_ERR_QUEUE_OVERFLOW]) { eng->err = "particle queue overflow: raise the 'cap_factor' option"; return TRK3_E_OVERFLOW; }
    return TRK3_OK;
}

int trk3_mc_run(trk3_engine *eng, int64_t it_begin, int64_t it_end, double *tallies, trk3_stats *stats) {
    if (!eng || !tallies) return TRK3_E_INVALID;
    CK(cudaSetDevice(eng->device));
    CK(cudaMemsetAsync(eng->d_tally, 0, (size_t)eng->lay.total * sizeof(double), eng->stream));     // stream order is all the run needs
    int rc = trk3_mc_run_device(eng, it_begin, it_end, stats);
    if (rc) return rc;
    return trk3_mc_download_tallies(eng, tallies);
}

int trk3_mc_iteration_energies(trk3_engine *eng, double *out, int64_t capacity, int64_t *n_iter) {
    if (!eng || !out || !n_iter) return TRK3_E_INVALID;
    const int64_t n = (int64_t)eng->iter_totE.size();
    const int64_t m = std::min<int64_t>(n, capacity);
    std::memcpy(out, eng->iter_totE.data(), (size_t)m * sizeof(double));
    *n_iter = m / eng->lay.Nt;
    return TRK3_OK;
}

// Use a caller-owned CUDA stream (e.g. torch.cuda.current_stream().cuda_stream) for all engine work.
int trk3_mc_set_stream(trk3_engine *eng, void *stream) {
    if (!eng) return TRK3_E_INVALID;
    CK(cudaSetDevice(eng->device));
    CK(cudaStreamSynchronize(eng->stream));
    if (eng->own_stream && eng->stream) cudaStreamDestroy(eng->stream);
    eng->stream = (cudaStream_t)stream; eng->own_stream = false;
    eng->l2_persist_applied = 0;            // the new stream has no access-policy window yet
    return TRK3_OK;
}
// Accumulate into a caller-owned DEVICE buffer of lay.total doubles (e.g. a torch tensor that is then all-reduced).
int trk3_mc_set_device_tallies(trk3_engine *eng, double *dptr) {
    if (!eng || !dptr) return TRK3_E_INVALID;
    eng->d_tally = dptr; eng->own_tally = false; eng->hp.tally = dptr;
    return TRK3_OK;
}
// Per-kernel-class device time of the runs since option "profile" was set: classes are
// 0 electron wave, 1 valence-hole wave, 2 core-hole wave, 3 photon wave, 4 ion tracks, 5 finalize.
int trk3_mc_kernel_times(trk3_engine *eng, double *ms, uint64_t *launches, int n) {
    if (!eng || !ms || !launches) return TRK3_E_INVALID;
    for (int i = 0; i < n && i < N_CLASSES; ++i) { ms[i] = eng->class_ms[i]; launches[i] = eng->class_launches[i]; }
    return N_CLASSES;
}

int trk3_nccl_unique_id(void *id) {
    if (!id) return TRK3_E_INVALID;
    nccl_rt::Api &a = nccl_rt::api();
    if (!a.err.empty()) return TRK3_E_UNSUPPORTED;
    return a.GetUniqueId((nccl_rt::UniqueId *)id) == nccl_rt::kSuccess ? TRK3_OK : TRK3_E_CUDA;
}
int trk3_mc_comm_init(trk3_engine *eng, int nranks, int rank, const void *id) {
    if (!eng || !id || nranks < 1 || rank < 0 || rank >= nranks) return TRK3_E_INVALID;
    nccl_rt::Api &a = nccl_rt::api();
    if (!a.err.empty()) { eng->err = a.err; return TRK3_E_UNSUPPORTED; }
    CK(cudaSetDevice(eng->device));
    if (eng->comm && eng->own_comm) a.CommDestroy(eng->comm);
    eng->comm = nullptr; eng->own_comm = false; eng->comm_size = 1;
    nccl_rt::UniqueId uid; std::memcpy(&uid, id, sizeof uid);
    const int nrc = a.CommInitRank(&eng->comm, nranks, uid, rank);
    if (nrc != nccl_rt::kSuccess) { eng->err = std::string("ncclCommInitRank: ") + a.GetErrorString(nrc); eng->comm = nullptr; return TRK3_E_CUDA; }
    eng->own_comm = true; eng->comm_size = nranks;
    return TRK3_OK;
}
int trk3_mc_set_comm(trk3_engine *eng, void *comm, int nranks) {
    if (!eng || (comm && nranks < 1)) return TRK3_E_INVALID;
    nccl_rt::Api &a = nccl_rt::api();
    if (comm && !a.err.empty()) { eng->err = a.err; return TRK3_E_UNSUPPORTED; }
    if (eng->comm && eng->own_comm) a.CommDestroy(eng->comm);
    eng->comm = comm; eng->own_comm = false; eng->comm_size = comm ? nranks : 1;
    return TRK3_OK;
}
int trk3_mc_comm_size(const trk3_engine *eng) { return eng ? eng->comm_size : 1; }

int trk3_mc_reset(trk3_engine *eng) {
    if (!eng) return TRK3_E_INVALID;
    if (cudaSetDevice(eng->device) != cudaSuccess) { eng->err = "device lost"; return TRK3_E_CUDA; }
    // wait for everything the failed run left in flight; a sticky error (illegal address ...) cannot be cleared
    cudaError_t e = cudaDeviceSynchronize();
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess && cudaDeviceSynchronize() != cudaSuccess) { eng->err = std::string("device lost: ") + cudaGetErrorString(e); return TRK3_E_CUDA; }
    const size_t n_counters = TRK3_N_EVENT_CLASSES + TRK3_N_ERRORS + 6;
    if (eng->d_counters) CK(cudaMemsetAsync(eng->d_counters, 0, n_counters * sizeof(unsigned long long), eng->stream));
    if (eng->d_qcount) CK(cudaMemsetAsync(eng->d_qcount, 0, QC_WORDS * sizeof(uint32_t), eng->stream));
    if (eng->d_tally) CK(cudaMemsetAsync(eng->d_tally, 0, (size_t)eng->lay.total * sizeof(double), eng->stream));
    if (eng->nb_alloc) { CK(cudaMemsetAsync(eng->d_u32, 0, eng->sl.u32_total * sizeof(uint32_t), eng->stream)); CK(cudaMemsetAsync(eng->d_f64, 0, eng->sl.f64_total * sizeof(double), eng->stream)); }
    CK(cudaStreamSynchronize(eng->stream));
    eng->ev_pending.clear(); eng->ev_info.clear();
    eng->iter_totE.clear(); eng->Dcoef.clear();
    eng->failed = false; eng->err.clear();
    return TRK3_OK;
}

void trk3_mc_destroy(trk3_engine *eng) {
    if (!eng) return;
    cudaSetDevice(eng->device);
    if (eng->comm && eng->own_comm) { if (eng->failed) nccl_rt::api().CommAbort(eng->comm); else nccl_rt::api().CommDestroy(eng->comm); }
    for (void *p : eng->allocs) cudaFree(p);
    if (eng->ev0) cudaEventDestroy(eng->ev0);
    if (eng->ev1) cudaEventDestroy(eng->ev1);
    for (auto &e : eng->ev_pool) { cudaEventDestroy(e.first); cudaEventDestroy(e.second); }
    if (eng->own_stream && eng->stream) cudaStreamDestroy(eng->stream);
    if (eng->stream_c) cudaStreamDestroy(eng->stream_c);
    for (int i = 0; i < 3; ++i) { if (eng->stream_sp[i]) cudaStreamDestroy(eng->stream_sp[i]); if (eng->ev_sp[i]) cudaEventDestroy(eng->ev_sp[i]); }
    if (eng->ev_gen) cudaEventDestroy(eng->ev_gen);
    if (eng->stream_w) cudaStreamDestroy(eng->stream_w);
    if (eng->ev_w) cudaEventDestroy(eng->ev_w);
    if (eng->stream_wh) cudaStreamDestroy(eng->stream_wh);
    if (eng->ev_wh) cudaEventDestroy(eng->ev_wh);
    if (eng->ev_fork) cudaEventDestroy(eng->ev_fork);
    if (eng->ev_join) cudaEventDestroy(eng->ev_join);
    if (eng->h_qcount) cudaFreeHost(eng->h_qcount);
    if (eng->stage) cudaFreeHost(eng->stage);            // the device is idle: cudaFree above synchronised it
    if (eng->ev_stage) cudaEventDestroy(eng->ev_stage);
    for (auto e : eng->ring_ev) cudaEventDestroy(e);
    delete eng;
}

}  // extern "C"
