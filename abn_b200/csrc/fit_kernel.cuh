// fit_kernel.cuh - the streaming fit kernel: one warp owns one (child, parent set) fit.
//
// Per pass the 32 lanes stride over the observation rows of the column-major FP64 matrix
// (lane l reads rows l, l+32, ... of each design column: 256-byte coalesced segments per column),
// evaluate the family's row op in FP64, and accumulate the k(k+1)/2 Gram entries, the k score
// entries and the log-likelihood term in FP64 registers. A fixed-order butterfly of warp
// shuffles reduces them (deterministic: the same bits whatever the partition of candidates over
// GPUs), lane 0 runs the solver state machine (solver.cuh) on the warp's shared-memory state and
// broadcasts the next evaluation point. Tensor cores are not used: the work per row is a few
// dozen FP64 flops against 8(k+1) streamed bytes.
//
// Work is handed out by a global atomic counter over a candidate list that keeps the cache
// order (same child, shared leading parents adjacent), so warps that run concurrently read the
// same columns and hit in L1/L2.
#pragma once

#include <cuda_runtime.h>

#include "plan.cuh"
#include "rowops.cuh"
#include "solver.cuh"

namespace abn {

struct ResultPtrs {
    double *mlik, *aic, *bic, *mdl;
    int *status, *iters;
};

constexpr int FIT_WARPS = 8; // warps per CTA

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    return v;
}

// one pass of one fit by one warp. MODE is PASS_FULL / PASS_HG / PASS_GRAD / PASS_HESS / PASS_LL.
template <int KIND, int K, bool ICPT, int MODE>
__device__ __forceinline__ void warp_pass(const double *__restrict__ y, const double *const (&cp)[MAXK],
                                          const double *beta_sm, long long n, const RowConst rc, PassAcc *acc_sm)
{
    constexpr int KD = K - (ICPT ? 1 : 0); // streamed design columns
    constexpr bool NEED_H = (MODE == PASS_FULL || MODE == PASS_HG || MODE == PASS_HESS);
    constexpr bool NEED_G = (MODE == PASS_FULL || MODE == PASS_HG || MODE == PASS_GRAD);
    constexpr bool NEED_L = (MODE == PASS_FULL || MODE == PASS_LL);
    const int lane = threadIdx.x & 31;

    double beta[K];
#pragma unroll
    for (int j = 0; j < K; j++) beta[j] = beta_sm[j];

    double H[K * (K + 1) / 2];
    double g[K];
    double ll = 0.0;
    int bad = 0;
#pragma unroll
    for (int j = 0; j < K * (K + 1) / 2; j++) H[j] = 0.0;
#pragma unroll
    for (int j = 0; j < K; j++) g[j] = 0.0;

#pragma unroll 2
    for (long long i = lane; i < n; i += 32) {
        double x[K];
        if (ICPT) x[0] = 1.0;
#pragma unroll
        for (int j = 0; j < KD; j++) x[j + (ICPT ? 1 : 0)] = __ldg(cp[j] + i);
        const double yi = __ldg(y + i);
        double eta = ICPT ? beta[0] : 0.0;
#pragma unroll
        for (int j = (ICPT ? 1 : 0); j < K; j++) eta = fma(x[j], beta[j], eta);
        double w, r, l = 0.0;
        RowOp<KIND>::template apply<NEED_H, NEED_G, NEED_L>(eta, yi, rc, w, r, l, bad);
        if (NEED_H) {
            int t = 0;
#pragma unroll
            for (int p = 0; p < K; p++) {
                const double wx = (ICPT && p == 0) ? w : w * x[p];
#pragma unroll
                for (int q = p; q < K; q++) {
                    H[t] = fma(wx, x[q], H[t]);
                    t++;
                }
            }
        }
        if (NEED_G) {
#pragma unroll
            for (int p = 0; p < K; p++) g[p] = fma(r, x[p], g[p]);
        }
        if (NEED_L) ll += l;
    }

    // fixed-order butterfly: every lane ends with the same totals
    if (NEED_H) {
#pragma unroll
        for (int j = 0; j < K * (K + 1) / 2; j++) H[j] = warp_sum(H[j]);
    }
    if (NEED_G) {
#pragma unroll
        for (int j = 0; j < K; j++) g[j] = warp_sum(g[j]);
    }
    if (NEED_L) ll = warp_sum(ll);
    bad = __any_sync(0xffffffffu, bad);
    if (lane == 0) {
        if (NEED_H) {
#pragma unroll
            for (int j = 0; j < K * (K + 1) / 2; j++) acc_sm->H[j] = H[j];
        }
        if (NEED_G) {
#pragma unroll
            for (int j = 0; j < K; j++) acc_sm->g[j] = g[j];
        }
        acc_sm->ll = ll;
        acc_sm->bad = bad;
    }
}

// KIND: RK_BIN_MLE, RK_BIN_BAYES or RK_POIS (both methods; the mle start-value pass uses RK_PINIT)
template <int KIND, int K, bool ICPT>
__global__ void __launch_bounds__(FIT_WARPS * 32, 2)
fit_kernel(const DataMeta m, const Control c, const FitDesc *__restrict__ descs, const int *__restrict__ child,
           const int *__restrict__ list, int count, int *counter, ResultPtrs out)
{
    __shared__ FitState st_all[FIT_WARPS];
    __shared__ PassAcc acc_all[FIT_WARPS];
    __shared__ double exp_tab[32];
    if (threadIdx.x < 32) exp_tab[threadIdx.x] = c_exp2_table[threadIdx.x];
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    FitState &st = st_all[warp];
    PassAcc *acc = &acc_all[warp];

    for (;;) {
        int idx = 0;
        if (lane == 0) idx = atomicAdd(counter, 1);
        idx = __shfl_sync(0xffffffffu, idx, 0);
        if (idx >= count) break;
        const int row = list[idx];
        const FitDesc d = descs[row];
        const int ch = child[row];

        constexpr int KD = K - (ICPT ? 1 : 0);
        const double *cp[MAXK];
#pragma unroll
        for (int j = 0; j < KD; j++) {
            const int col = d.cols[j + (ICPT ? 1 : 0)];
            cp[j] = m.X + (long long)(col < 0 ? 0 : col) * m.ld;
        }
        const double *y = m.X + (long long)d.ycol * m.ld;
        const RowConst rc = make_row_const(K, (double)m.n, exp_tab);

        int mode = 0;
        if (lane == 0) {
            state_setup(m, d, ch, c, st);
            if (c.method == BAYES) bayes_inits(m, d, ch, st);
            mode = fit_begin(st, c);
        }
        mode = __shfl_sync(0xffffffffu, mode, 0);
        __syncwarp();
        while (mode != PASS_DONE) {
            // only the modes a row kind's state machines can ask for are instantiated
            if (mode == PASS_HG) warp_pass<KIND, K, ICPT, PASS_HG>(y, cp, st.x, m.n, rc, acc);
            else if (mode == PASS_FULL) warp_pass<KIND, K, ICPT, PASS_FULL>(y, cp, st.x, m.n, rc, acc);
            else if (mode == PASS_LL) warp_pass<KIND, K, ICPT, PASS_LL>(y, cp, st.x, m.n, rc, acc);
            else if (KIND != RK_BIN_MLE && mode == PASS_GRAD) warp_pass<KIND, K, ICPT, PASS_GRAD>(y, cp, st.x, m.n, rc, acc);
            else if (KIND != RK_BIN_MLE && mode == PASS_HESS) warp_pass<KIND, K, ICPT, PASS_HESS>(y, cp, st.x, m.n, rc, acc);
            else if (KIND == RK_POIS && mode == PASS_PINIT) warp_pass<RK_PINIT, K, ICPT, PASS_HG>(y, cp, st.x, m.n, rc, acc);
            __syncwarp();
            if (lane == 0) mode = fit_consume(st, c, *acc);
            mode = __shfl_sync(0xffffffffu, mode, 0);
            __syncwarp();
        }
        if (lane == 0) {
            if (out.mlik) out.mlik[row] = st.loglik;
            if (c.method == MLE) {
                if (out.aic) out.aic[row] = st.aic;
                if (out.bic) out.bic[row] = st.bic;
                if (out.mdl) out.mdl[row] = mdl_of(st.bic, d, c.nvars);
            }
            if (out.status) out.status[row] = st.status;
            if (out.iters) out.iters[row] = st.passes;
        }
        __syncwarp();
    }
}

// host-side launcher table entry
typedef void (*FitLaunchFn)(const DataMeta &, const Control &, const FitDesc *, const int *, const int *, int,
                            int *, ResultPtrs, int grid, cudaStream_t);

template <int KIND, int K, bool ICPT>
void launch_fit(const DataMeta &m, const Control &c, const FitDesc *descs, const int *child, const int *list,
                int count, int *counter, ResultPtrs out, int grid, cudaStream_t stream)
{
    fit_kernel<KIND, K, ICPT><<<grid, FIT_WARPS * 32, 0, stream>>>(m, c, descs, child, list, count, counter, out);
}

// defined in fit_bin_mle.cu / fit_bin_bayes.cu / fit_pois.cu: [k][icpt]
FitLaunchFn get_launch_bin_mle(int k, int icpt);
FitLaunchFn get_launch_bin_bayes(int k, int icpt);
FitLaunchFn get_launch_pois(int k, int icpt);

} // namespace abn
