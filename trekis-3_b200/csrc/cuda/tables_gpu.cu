// tables_gpu.cu -- GPU evaluator of the table builder's integrands (trk3_dcs_eval, include/trekis3_gpu.h): SURVEY.md 8(f) N1.
//
// The reference spends "minutes to hours" of a first run in TotIMFP / Tot_EMFP / SHI_TotIMFP (Cross_sections.f90:881-1050,
// :2966-3139, :2452-2597): nested Simpson integrations whose inner level, the q-integral of the loss function, is
// independent for every (grid energy, shell, transferred energy).  The host builder records those requests (millions per
// table) and this file evaluates them, one thread per request; the integrands are the very functions the host uses
// (csrc/common/trk3_dcs.h) and this translation unit is compiled with -fmad=false, so that every value is the result of
// the same IEEE-754 operations as on the host: the tables come out identical.
//
// Mapping: a request is ~10^2..10^4 sequential Simpson steps in q (dq = q/100: geometric), each two loss-function
// evaluations (an 8-step bisection in the DOS k-grid for the effective mass + a sum over <~10 oscillators): pure fp64
// arithmetic on ~2 KB of constants (oscillators + DOS, staged in shared memory), no HBM traffic to speak of -> bound by
// the FP64 pipe and the longest request of a warp.  Requests of a task are consecutive in hw, so the lanes of a warp run
// loops of similar length; blocks take 128 requests at a time from a global counter (the lengths differ between tasks).
// No tensor cores: nothing here is a contraction.
#include <cuda_runtime.h>
#include <algorithm>
#include <cstdio>
#include <mutex>
#include <vector>
#include "../common/trk3_dcs.h"

namespace {

#define DCS_BLOCK 128
#define DCS_SMEM_DOUBLES 5600          // most oscillators + DOS entries copied to shared memory (44.8 KB); larger inputs stay in global memory

__global__ void __launch_bounds__(DCS_BLOCK) k_dcs(trk3_dcs_ctx x, const trk3_dcs_task *tasks, const double *hw, const int32_t *task_of,
                                                   long long n, double *out, unsigned long long *next, int n_osc, int use_smem) {
    extern __shared__ double s_tab[];      // 3 n_osc + 2 n_k doubles (typically ~4 KB: does not limit the occupancy)
    __shared__ long long s_base;
    if (use_smem) {     // constants of the integrands -> shared memory (every loss-function evaluation reads them)
        double *p = s_tab;
        for (int i = threadIdx.x; i < n_osc; i += blockDim.x) { p[i] = x.osc_E0[i]; p[n_osc + i] = x.osc_A[i]; p[2 * n_osc + i] = x.osc_G[i]; }
        for (int i = threadIdx.x; i < x.n_k; i += blockDim.x) { p[3 * n_osc + i] = x.k[i]; p[3 * n_osc + x.n_k + i] = x.effm[i]; }
        __syncthreads();
        x.osc_E0 = p; x.osc_A = p + n_osc; x.osc_G = p + 2 * n_osc;
        if (x.n_k > 0) { x.k = p + 3 * n_osc; x.effm = p + 3 * n_osc + x.n_k; }
    }
    for (;;) {
        if (threadIdx.x == 0) s_base = (long long)atomicAdd(next, (unsigned long long)blockDim.x);
        __syncthreads();
        const long long base = s_base, i = base + threadIdx.x;
        __syncthreads();
        if (base >= n) break;
        if (i < n) out[i] = trk3dcs::eval_request(x, tasks[task_of[i]], hw[i]);
    }
}

struct DcsState {
    std::mutex mu;
    double device_ms = 0.0;
    long long requests = 0;
} g_state;

#define CKD(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { std::fprintf(stderr, "trk3_dcs_eval: %s: %s\n", #call, cudaGetErrorString(e_)); rc = TRK3_E_CUDA; goto done; } } while (0)

}  // namespace

extern "C" int trk3_dcs_eval(const trk3_dcs_ctx *ctx, const trk3_dcs_task *tasks, int64_t n_tasks,
                             const double *hw, const int32_t *task_of, int64_t n, double *out) {
    if (!ctx || !tasks || !hw || !task_of || !out || n < 0 || n_tasks < 0 || ctx->n_sets < 1) return TRK3_E_INVALID;
    if (n == 0) return TRK3_OK;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) { std::fprintf(stderr, "trk3_dcs_eval: no CUDA device (this library has no CPU fallback)\n"); return TRK3_E_CUDA; }
    int rc = TRK3_OK;
    const int n_osc = ctx->osc_off[ctx->n_sets];
    trk3_dcs_ctx d = *ctx;
    double *d_osc = nullptr, *d_dos = nullptr, *d_hw = nullptr, *d_out = nullptr, *d_scr = nullptr;
    int32_t *d_off = nullptr, *d_task_of = nullptr;
    trk3_dcs_task *d_tasks = nullptr;
    unsigned long long *d_next = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    cudaStream_t st = nullptr;
    float ms = 0.f;
    {
        CKD(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
        CKD(cudaEventCreate(&e0)); CKD(cudaEventCreate(&e1));
        CKD(cudaMalloc(&d_osc, sizeof(double) * 3 * std::max(n_osc, 1)));
        CKD(cudaMalloc(&d_dos, sizeof(double) * 2 * std::max(ctx->n_k, 1)));
        CKD(cudaMalloc(&d_off, sizeof(int32_t) * (ctx->n_sets + 1)));
        CKD(cudaMalloc(&d_tasks, sizeof(trk3_dcs_task) * std::max<int64_t>(n_tasks, 1)));
        CKD(cudaMalloc(&d_hw, sizeof(double) * n)); CKD(cudaMalloc(&d_out, sizeof(double) * n));
        CKD(cudaMalloc(&d_task_of, sizeof(int32_t) * n));
        CKD(cudaMalloc(&d_next, sizeof(unsigned long long)));
        if (n_osc) {
            CKD(cudaMemcpyAsync(d_osc, ctx->osc_E0, sizeof(double) * n_osc, cudaMemcpyHostToDevice, st));
            CKD(cudaMemcpyAsync(d_osc + n_osc, ctx->osc_A, sizeof(double) * n_osc, cudaMemcpyHostToDevice, st));
            CKD(cudaMemcpyAsync(d_osc + 2 * n_osc, ctx->osc_G, sizeof(double) * n_osc, cudaMemcpyHostToDevice, st));
        }
        if (ctx->n_k > 0) {
            CKD(cudaMemcpyAsync(d_dos, ctx->k, sizeof(double) * ctx->n_k, cudaMemcpyHostToDevice, st));
            CKD(cudaMemcpyAsync(d_dos + ctx->n_k, ctx->effm, sizeof(double) * ctx->n_k, cudaMemcpyHostToDevice, st));
        }
        if (ctx->screening >= 2) {       // what the screening of the elastic cross section reads (trk3_dcs_ctx::scr)
            if (!ctx->scr || ctx->n_scr < 1) { rc = TRK3_E_INVALID; goto done; }
            CKD(cudaMalloc(&d_scr, sizeof(double) * ctx->n_scr));
            CKD(cudaMemcpyAsync(d_scr, ctx->scr, sizeof(double) * ctx->n_scr, cudaMemcpyHostToDevice, st));
            d.scr = d_scr;
        }
        CKD(cudaMemcpyAsync(d_off, ctx->osc_off, sizeof(int32_t) * (ctx->n_sets + 1), cudaMemcpyHostToDevice, st));
        CKD(cudaMemcpyAsync(d_tasks, tasks, sizeof(trk3_dcs_task) * n_tasks, cudaMemcpyHostToDevice, st));
        CKD(cudaMemcpyAsync(d_hw, hw, sizeof(double) * n, cudaMemcpyHostToDevice, st));
        CKD(cudaMemcpyAsync(d_task_of, task_of, sizeof(int32_t) * n, cudaMemcpyHostToDevice, st));
        CKD(cudaMemsetAsync(d_next, 0, sizeof(unsigned long long), st));
        d.osc_E0 = d_osc; d.osc_A = d_osc + n_osc; d.osc_G = d_osc + 2 * n_osc; d.osc_off = d_off;
        d.k = d_dos; d.effm = d_dos + ctx->n_k;
        int dev = 0, n_sm = 0;
        CKD(cudaGetDevice(&dev));
        CKD(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
        int bps = 1;
        const int use_smem = (3 * n_osc + 2 * ctx->n_k <= DCS_SMEM_DOUBLES) ? 1 : 0;
        const size_t smem = use_smem ? sizeof(double) * (size_t)(3 * n_osc + 2 * ctx->n_k) : 0;
        CKD(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, k_dcs, DCS_BLOCK, smem));
        // persistent grid: a multiple of the SM count, never more blocks than there are chunks of requests
        const long long chunks = (n + DCS_BLOCK - 1) / DCS_BLOCK;
        const int grid = (int)std::min<long long>(chunks, (long long)n_sm * std::max(bps, 1));
        CKD(cudaEventRecord(e0, st));
        k_dcs<<<grid, DCS_BLOCK, smem, st>>>(d, d_tasks, d_hw, d_task_of, (long long)n, d_out, d_next, n_osc, use_smem);
        CKD(cudaGetLastError());
        CKD(cudaEventRecord(e1, st));
        CKD(cudaMemcpyAsync(out, d_out, sizeof(double) * n, cudaMemcpyDeviceToHost, st));
        CKD(cudaStreamSynchronize(st));
        CKD(cudaEventElapsedTime(&ms, e0, e1));
        std::lock_guard<std::mutex> lk(g_state.mu);
        g_state.device_ms += ms; g_state.requests += n;
    }
done:
    cudaFree(d_scr); cudaFree(d_osc); cudaFree(d_dos); cudaFree(d_off); cudaFree(d_tasks); cudaFree(d_hw); cudaFree(d_out); cudaFree(d_task_of); cudaFree(d_next);
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    if (st) cudaStreamDestroy(st);
    return rc;
}

extern "C" int trk3_dcs_stats(double *device_ms, int64_t *requests, int reset) {
    std::lock_guard<std::mutex> lk(g_state.mu);
    if (device_ms) *device_ms = g_state.device_ms;
    if (requests) *requests = g_state.requests;
    if (reset) { g_state.device_ms = 0.0; g_state.requests = 0; }
    return TRK3_OK;
}
