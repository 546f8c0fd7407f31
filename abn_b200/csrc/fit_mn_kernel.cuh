// fit_mn_kernel.cuh - streaming kernel for multinomial children (method = "mle"): one warp per
// fit, same row partition and fixed-order reduction as fit_kernel.cuh. A pass evaluates the
// softmax objective and its gradient at the current weights (solver_mn.cuh drives vmmin).
#pragma once

#include <cuda_runtime.h>

#include "fit_kernel.cuh"
#include "rowops_mn.cuh"
#include "solver_mn.cuh"

namespace ABN_NS {

// CTAs per SM the multinomial kernels are compiled for when the fit has at most MN_SMALL_NP weights (2: 128 registers
// per thread, 16 warps per SM; 1: 255 registers, 8 warps - room to interleave the four rows of a tile step).
// Measured on the config-5 slice of tests/run_mn_slice.py (one box): 12.17 s with 2, 11.26 s with 1.
#ifndef ABN_MN_BLOCKS
#define ABN_MN_BLOCKS 1
#endif
constexpr int MN_SMALL_NP = 16;
constexpr int MN_UNROLL = 2;
constexpr int MN_WARPS = 8; // warps per CTA (its per-warp BFGS state is 5.8 KB of static shared memory)

// per warp: the tiles of the careful pass (design + indicator columns, 64 rows, FIT_STAGES) or of the lean pass
// (design columns only, 128 rows, LEAN_STAGES), whichever is larger
template <int L, int K, bool ICPT>
constexpr size_t mn_stage_doubles_per_warp()
{
    const size_t careful = (size_t)FIT_STAGES * (K - (ICPT ? 1 : 0) + L - 1) * 32 * MN_UNROLL;
    const size_t lean = (size_t)LEAN_STAGES * (K - (ICPT ? 1 : 0)) * LEAN_TROWS;
    return careful > lean ? careful : lean;
}
// dynamic shared memory of a CTA: the tiles, then the pre-scaled exp table of the lean pass, then per warp
// (design column)'(level indicator) of the fit in flight (the static 48 KB are taken by the BFGS states)
template <int L, int K, bool ICPT>
constexpr size_t mn_stage_bytes()
{
    return ((size_t)MN_WARPS * mn_stage_doubles_per_warp<L, K, ICPT>() + ABN_EXP256_N + (size_t)MN_WARPS * (L - 1) * K) * sizeof(double);
}

// The lean pass (rowops_mn.cuh mn_row_lean): the tiles, copies and row order of fit_kernel.cuh's lean passes, design
// columns only. false: some linear predictor left (-700, 700) or is NaN - nothing written, the careful pass redoes it.
template <int L, int K, bool ICPT>
__device__ __forceinline__ bool mn_lean_pass(const double *const (&cp)[MAXK], const double *w_sm, const double *xty, long long n,
                                             const double *tab256, const double *logtab, MnAcc *acc_sm, double *tile)
{
    constexpr int KD = K - (ICPT ? 1 : 0);
    constexpr int KDA = KD > 0 ? KD : 1;
    constexpr int NP = (L - 1) * K;
    const int lane = threadIdx.x & 31;
    double W[NP], g[NP], f = 0.0;
#pragma unroll
    for (int i = 0; i < NP; i++) { W[i] = w_sm[i]; g[i] = 0.0; }
    RowTrack tk = row_track_init();
    auto row = [&](const double (&x)[KDA]) -> int {
        double o[L - 1];
        mn_row_lean<L, K, ICPT>(x, W, tab256, logtab, f, g, o);
#pragma unroll
        for (int l = 1; l < L - 1; l++) {
            const int h = hi_word(o[l]);
            tk.smax = max(tk.smax, h);
            tk.umax = max(tk.umax, (unsigned)h);
        }
        return hi_word(o[0]);
    };
    lean_stream<KD>(cp, n, tile, tk, row);
    if (__any_sync(0xffffffffu, (tk.smax >= 0x4085E000) | (tk.umax >= 0xC085E000u))) return false;
    f = warp_sum(f);
#pragma unroll
    for (int i = 0; i < NP; i++) g[i] = warp_sum(g[i]);
    if (lane == 0) {
        if (xty) {
            mn_lean_finish(NP, f, g, W, xty, &acc_sm->f, acc_sm->g);
        } else { // partial sums of a row range (fit_mn_cta_kernel): the Gram terms are added once, after the ranges are summed
            acc_sm->f = f;
#pragma unroll
            for (int i = 0; i < NP; i++) acc_sm->g[i] = g[i];
        }
        acc_sm->bad = 0;
    }
    return true;
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
__global__ void __launch_bounds__(MN_WARPS * 32, (L - 1) * K <= MN_SMALL_NP ? ABN_MN_BLOCKS : 1)
fit_mn_kernel(const DataMeta m, const Control c, const FitDesc *__restrict__ descs, const int *__restrict__ list,
              int count, int *counter, ResultPtrs out)
{
    __shared__ MnState st_all[MN_WARPS];
    __shared__ MnAcc acc_all[MN_WARPS];
    __shared__ double exp_tab[ABN_EXP2_N];
    __shared__ double log_tab[128];
    extern __shared__ __align__(16) double stage_tiles[];
    double *const exp256_tab = stage_tiles + (size_t)MN_WARPS * mn_stage_doubles_per_warp<L, K, ICPT>();
    double *const xty_all = exp256_tab + ABN_EXP256_N;
    if (threadIdx.x < ABN_EXP2_N) exp_tab[threadIdx.x] = c_exp2_table[threadIdx.x];
    if (threadIdx.x < 64) fill_log_table(log_tab, threadIdx.x);
    for (int j = threadIdx.x; j < ABN_EXP256_N; j += MN_WARPS * 32) exp256_tab[j] = exp256_entry(c_exp2_256_table[j], j);
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    MnState &st = st_all[warp];
    MnAcc *acc = &acc_all[warp];
    double *tile = stage_tiles + (size_t)warp * mn_stage_doubles_per_warp<L, K, ICPT>();
    double *xty = xty_all + (size_t)warp * (L - 1) * K;
    const bool lean_ok = !(c.flags & 1); // ABN_FLAG_EXACT_PATH: the row arithmetic of round 1 (response streamed, shifted softmax)
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
            for (int l = 1; l < L; l++)
                for (int j = 0; j < K; j++) xty[(l - 1) * K + j] = m.G[(long long)gidx(d, j) * m.D + gram_col(d.ycol + l)];
        }
        more = __shfl_sync(0xffffffffu, more, 0);
        __syncwarp();
        bool careful = !lean_ok;
        while (more) {
            if (!careful && !mn_lean_pass<L, K, ICPT>(cp, st.b, xty, m.n, exp256_tab, log_tab, acc, tile)) careful = true;
            if (careful) mn_warp_pass<L, K, ICPT>(y0, m.ld, cp, st.b, m.n, exp_tab, log_tab, acc, tile);
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

// ---- one fit per CTA ------------------------------------------------------------------------------------------
// A multinomial fit is hundreds of passes (vmmin runs to its iteration cap on most of BASELINE's config 5: ~900
// function evaluations of 500 000 rows), so with one warp per fit a GPU works on 2 368 fits of very different
// lengths at a time and idles through the tail of every bucket - and a node's 92 171 parent sets are only ~5 waves
// on 8 GPUs. Here the 8 warps of a CTA share ONE fit: warp w streams the w-th eighth of the row tiles, the per-warp
// sums are added in warp order (fixed: results do not depend on scheduling), thread 0 runs vmmin and the next
// weights are broadcast through shared memory. Fits are 8x shorter, the queue 8x finer.
template <int L, int K, bool ICPT>
__global__ void __launch_bounds__(MN_WARPS * 32, (L - 1) * K <= MN_SMALL_NP ? ABN_MN_BLOCKS : 1)
fit_mn_cta_kernel(const DataMeta m, const Control c, const FitDesc *__restrict__ descs, const int *__restrict__ list,
                  int count, int *counter, ResultPtrs out)
{
    constexpr int NP = (L - 1) * K;
    constexpr int KD = K - (ICPT ? 1 : 0);
    __shared__ MnState st;
    __shared__ MnAcc part[MN_WARPS];
    __shared__ MnAcc acc;
    __shared__ double xty[NP];
    __shared__ double exp_tab[ABN_EXP2_N];
    __shared__ double log_tab[128];
    __shared__ double exp256_tab[ABN_EXP256_N];
    __shared__ int sh_idx, sh_more, sh_careful, sh_lean_failed[MN_WARPS];
    extern __shared__ __align__(16) double stage_tiles[];
    if (threadIdx.x < ABN_EXP2_N) exp_tab[threadIdx.x] = c_exp2_table[threadIdx.x];
    if (threadIdx.x < 64) fill_log_table(log_tab, threadIdx.x);
    for (int j = threadIdx.x; j < ABN_EXP256_N; j += MN_WARPS * 32) exp256_tab[j] = exp256_entry(c_exp2_256_table[j], j);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double *tile = stage_tiles + (size_t)warp * mn_stage_doubles_per_warp<L, K, ICPT>();
    // rows of this warp: whole 128-row tiles, the last warp also takes what is left over
    const long long ntiles = m.n / LEAN_TROWS;
    const long long r0 = (ntiles * warp / MN_WARPS) * LEAN_TROWS;
    const long long r1 = (warp == MN_WARPS - 1) ? m.n : (ntiles * (warp + 1) / MN_WARPS) * LEAN_TROWS;
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) sh_idx = atomicAdd(counter, 1);
        __syncthreads();
        const int idx = sh_idx;
        if (idx >= count) break;
        const int row = list[idx];
        const FitDesc d = descs[row];
        const double *cp[MAXK];
#pragma unroll
        for (int j = 0; j < KD; j++) {
            const int col = d.cols[j + (ICPT ? 1 : 0)];
            cp[j] = m.X + (long long)(col < 0 ? 0 : col) * m.ld + r0;
        }
        const double *y0 = m.X + (long long)d.ycol * m.ld + r0;
        if (threadIdx.x == 0) {
            st.k = K; st.L = L; st.n = (double)m.n; st.rank = d.rank;
            sh_more = mn_begin(st) ? 1 : 0;
            sh_careful = (c.flags & 1) ? 1 : 0;
            for (int l = 1; l < L; l++)
                for (int j = 0; j < K; j++) xty[(l - 1) * K + j] = m.G[(long long)gidx(d, j) * m.D + gram_col(d.ycol + l)];
        }
        __syncthreads();
        while (sh_more) {
            bool careful = sh_careful != 0;
            if (!careful) {
                const bool ok = mn_lean_pass<L, K, ICPT>(cp, st.b, nullptr, r1 - r0, exp256_tab, log_tab, &part[warp], tile);
                if (lane == 0) sh_lean_failed[warp] = ok ? 0 : 1;
                __syncthreads();
                bool any = false;
#pragma unroll
                for (int w = 0; w < MN_WARPS; w++) any |= sh_lean_failed[w] != 0;
                careful = any; // some range left (-700, 700): every warp redoes its range with the careful pass
            }
            if (careful) {
                mn_warp_pass<L, K, ICPT>(y0, m.ld, cp, st.b, r1 - r0, exp_tab, log_tab, &part[warp], tile);
                __syncthreads();
            }
            if (threadIdx.x == 0) {
                double f = 0.0, g[NP];
                for (int i = 0; i < NP; i++) g[i] = 0.0;
                for (int w = 0; w < MN_WARPS; w++) { // warp order: fixed
                    f += part[w].f;
                    for (int i = 0; i < NP; i++) g[i] += part[w].g[i];
                }
                if (careful) {
                    acc.f = f;
                    for (int i = 0; i < NP; i++) acc.g[i] = g[i];
                } else {
                    mn_lean_finish(NP, f, g, st.b, xty, &acc.f, acc.g);
                }
                acc.bad = 0;
                sh_careful = careful ? 1 : 0;
                sh_more = mn_consume(st, acc) ? 1 : 0;
            }
            __syncthreads();
        }
        if (threadIdx.x == 0) {
            if (out.mlik) out.mlik[row] = st.loglik;
            if (out.aic) out.aic[row] = st.aic;
            if (out.bic) out.bic[row] = st.bic;
            if (out.mdl) out.mdl[row] = mdl_of(st.bic, d, c.nvars);
            if (out.status) out.status[row] = st.status;
            if (out.iters) out.iters[row] = st.passes;
        }
    }
}

// rows from which a fit is spread over a CTA: every warp gets at least 8 tiles
constexpr long long MN_CTA_MIN_ROWS = (long long)MN_WARPS * 8 * LEAN_TROWS;

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
    if (m.n >= MN_CTA_MIN_ROWS) {
        static bool configured_cta[64] = {false};
        if (first_launch_on_device(configured_cta)) {
            cudaError_t e = cudaFuncSetAttribute(fit_mn_cta_kernel<L, K, ICPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
        }
        int dev = 0, sms = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        const int per_sm = (L - 1) * K <= MN_SMALL_NP ? ABN_MN_BLOCKS : 1;
        const int grid_cta = count < sms * per_sm ? count : sms * per_sm;
        fit_mn_cta_kernel<L, K, ICPT><<<grid_cta, MN_WARPS * 32, smem, stream>>>(m, c, descs, list, count, counter, out);
        return cudaGetLastError();
    }
    fit_mn_kernel<L, K, ICPT><<<grid, MN_WARPS * 32, smem, stream>>>(m, c, descs, list, count, counter, out);
    return cudaGetLastError();
}

MnLaunchFn get_launch_mn(int L, int k, int icpt); // fit_table.cu

} // namespace ABN_NS
