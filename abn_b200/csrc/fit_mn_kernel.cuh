// fit_mn_kernel.cuh - streaming kernel for multinomial children (method = "mle"): one warp per
// fit, same row partition and fixed-order reduction as fit_kernel.cuh. A pass evaluates the
// softmax objective and its gradient at the current weights (solver_mn.cuh drives vmmin).
#pragma once

#include <cuda_runtime.h>

#include "fit_kernel.cuh"
#include "rowops_mn.cuh"
#include "solver_mn.cuh"

namespace ABN_NS {

constexpr int MN_UNROLL = 2;
constexpr int MN_WARPS = 8; // warps per CTA (its per-warp BFGS state is 5.8 KB of static shared memory)

template <int L, int K, bool ICPT>
constexpr size_t mn_stage_bytes()
{
    return (size_t)MN_WARPS * FIT_STAGES * (K - (ICPT ? 1 : 0) + L - 1) * 32 * MN_UNROLL * sizeof(double);
}

// same cp.async staging as warp_pass_staged (fit_kernel.cuh): each lane copies and reads back only
// its own rows; staged columns = the design columns plus the L-1 indicator columns of the child
template <int L, int K, bool ICPT>
__device__ __forceinline__ void mn_warp_pass(const double *__restrict__ y0, long long ld, const double *const (&cp)[MAXK],
                                             const double *w_sm, long long n, const double *tab, const double *logtab,
                                             MnAcc *acc_sm, double *tile)
{
    constexpr int KD = K - (ICPT ? 1 : 0);
    constexpr int I0 = ICPT ? 1 : 0;
    constexpr int NP = (L - 1) * K;
    constexpr int U = MN_UNROLL;
    constexpr int NC = KD + L - 1;
    constexpr int TROWS = 32 * U;
    const int lane = threadIdx.x & 31;
    double W[NP], g[NP], f = 0.0;
#pragma unroll
    for (int i = 0; i < NP; i++) { W[i] = w_sm[i]; g[i] = 0.0; }
    int bad = 0;
    const long long n_main = (n / TROWS) * TROWS;
    auto issue = [&](long long base, int stage) {
        double *t = tile + (size_t)stage * NC * TROWS;
#pragma unroll
        for (int u = 0; u < U; u++) {
#pragma unroll
            for (int j = 0; j < KD; j++) cp_async8(t + j * TROWS + u * 32 + lane, cp[j] + base + u * 32 + lane);
#pragma unroll
            for (int l = 1; l < L; l++) cp_async8(t + (KD + l - 1) * TROWS + u * 32 + lane, y0 + (long long)l * ld + base + u * 32 + lane);
        }
        cp_async_commit();
    };
    if (n_main > 0) issue(0, 0);
    int stage = 0;
    for (long long base = 0; base < n_main; base += TROWS) {
        const long long nb = base + TROWS;
        if (nb < n_main) { issue(nb, stage ^ 1); cp_async_wait<1>(); }
        else cp_async_wait<0>();
        const double *t = tile + (size_t)stage * NC * TROWS;
#pragma unroll
        for (int u = 0; u < U; u++) {
            double x[K], yl[L - 1];
            if (ICPT) x[0] = 1.0;
#pragma unroll
            for (int j = 0; j < KD; j++) x[j + I0] = t[j * TROWS + u * 32 + lane];
#pragma unroll
            for (int l = 1; l < L; l++) yl[l - 1] = t[(KD + l - 1) * TROWS + u * 32 + lane];
            mn_row<L, K, ICPT>(x, yl, W, tab, logtab, f, g);
        }
        stage ^= 1;
    }
    for (long long i = n_main + lane; i < n; i += 32) {
        double x[K], yl[L - 1];
        if (ICPT) x[0] = 1.0;
#pragma unroll
        for (int j = 0; j < KD; j++) x[j + I0] = __ldg(cp[j] + i);
#pragma unroll
        for (int l = 1; l < L; l++) yl[l - 1] = __ldg(y0 + (long long)l * ld + i);
        mn_row<L, K, ICPT>(x, yl, W, tab, logtab, f, g);
    }
    f = warp_sum(f);
#pragma unroll
    for (int i = 0; i < NP; i++) g[i] = warp_sum(g[i]);
    if (lane == 0) {
        acc_sm->f = f;
#pragma unroll
        for (int i = 0; i < NP; i++) acc_sm->g[i] = g[i];
        acc_sm->bad = bad;
    }
}

template <int L, int K, bool ICPT>
__global__ void __launch_bounds__(MN_WARPS * 32, (L - 1) * K <= 16 ? 2 : 1)
fit_mn_kernel(const DataMeta m, const Control c, const FitDesc *__restrict__ descs, const int *__restrict__ list,
              int count, int *counter, ResultPtrs out)
{
    __shared__ MnState st_all[MN_WARPS];
    __shared__ MnAcc acc_all[MN_WARPS];
    __shared__ double exp_tab[ABN_EXP2_N];
    __shared__ double log_tab[128];
    if (threadIdx.x < ABN_EXP2_N) exp_tab[threadIdx.x] = c_exp2_table[threadIdx.x];
    if (threadIdx.x < 64) fill_log_table(log_tab, threadIdx.x);
    __syncthreads();
    extern __shared__ __align__(16) double stage_tiles[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    MnState &st = st_all[warp];
    MnAcc *acc = &acc_all[warp];
    double *tile = stage_tiles + (size_t)warp * FIT_STAGES * (K - (ICPT ? 1 : 0) + L - 1) * 32 * MN_UNROLL;
    for (;;) {
        int idx = 0;
        if (lane == 0) idx = atomicAdd(counter, 1);
        idx = __shfl_sync(0xffffffffu, idx, 0);
        if (idx >= count) break;
        const int row = list[idx];
        const FitDesc d = descs[row];
        constexpr int KD = K - (ICPT ? 1 : 0);
        const double *cp[MAXK];
#pragma unroll
        for (int j = 0; j < KD; j++) {
            const int col = d.cols[j + (ICPT ? 1 : 0)];
            cp[j] = m.X + (long long)(col < 0 ? 0 : col) * m.ld;
        }
        const double *y0 = m.X + (long long)d.ycol * m.ld;
        int more = 0;
        if (lane == 0) {
            st.k = K; st.L = L; st.n = (double)m.n; st.rank = d.rank;
            more = mn_begin(st) ? 1 : 0;
        }
        more = __shfl_sync(0xffffffffu, more, 0);
        __syncwarp();
        while (more) {
            mn_warp_pass<L, K, ICPT>(y0, m.ld, cp, st.b, m.n, exp_tab, log_tab, acc, tile);
            __syncwarp();
            if (lane == 0) more = mn_consume(st, *acc) ? 1 : 0;
            more = __shfl_sync(0xffffffffu, more, 0);
            __syncwarp();
        }
        if (lane == 0) {
            if (out.mlik) out.mlik[row] = st.loglik;
            if (out.aic) out.aic[row] = st.aic;
            if (out.bic) out.bic[row] = st.bic;
            if (out.mdl) out.mdl[row] = mdl_of(st.bic, d, c.nvars);
            if (out.status) out.status[row] = st.status;
            if (out.iters) out.iters[row] = st.passes;
        }
        __syncwarp();
    }
}

typedef cudaError_t (*MnLaunchFn)(const DataMeta &, const Control &, const FitDesc *, const int *, int, int *, ResultPtrs,
                                  int grid, cudaStream_t);

template <int L, int K, bool ICPT>
cudaError_t launch_mn(const DataMeta &m, const Control &c, const FitDesc *descs, const int *list, int count, int *counter,
                      ResultPtrs out, int grid, cudaStream_t stream)
{
    constexpr size_t smem = mn_stage_bytes<L, K, ICPT>();
    static bool configured[64] = {false}; // per device, see launch_fit
    if (first_launch_on_device(configured)) {
        cudaError_t e = cudaFuncSetAttribute(fit_mn_kernel<L, K, ICPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    fit_mn_kernel<L, K, ICPT><<<grid, MN_WARPS * 32, smem, stream>>>(m, c, descs, list, count, counter, out);
    return cudaGetLastError();
}

MnLaunchFn get_launch_mn(int L, int k, int icpt); // fit_table.cu

} // namespace ABN_NS
