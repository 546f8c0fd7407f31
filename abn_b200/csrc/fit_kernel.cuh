// fit_kernel.cuh - the streaming fit kernel: one warp owns one (child, parent set) fit.
//
// Per pass the 32 lanes stride over the observation rows of the column-major FP64 matrix
// (lane l reads rows l, l+32, ... of each design column: 256-byte coalesced segments per column),
// evaluate the family's row op in FP64, and accumulate the k(k+1)/2 Gram entries, the k score
// entries and the log-likelihood term in FP64 registers. A fixed-order butterfly of warp
// shuffles reduces them (deterministic: the same bits whatever the partition of candidates over
// GPUs), lane 0 runs the solver state machine (solver.cuh) on the warp's shared-memory state and
// broadcasts the next evaluation point. Tensor cores are not used: the work per row is a few
// dozen FP64 flops against 8 k_data streamed bytes.
//
// Designs with an intercept (all of them unless the rank rule drops it) never stream the response:
// X'y and X'1 come from the global Gram matrix, so the Poisson pass (warp_pass_pois) needs only
// H = sum E x x' - its first row gives the score, H_00 the log-likelihood - and the binomial pass
// (warp_pass_bin) one extra vector sum. The generic passes (warp_pass, warp_pass_staged) remain
// for designs without an intercept and for the reference's Poisson start-value pass
// (ABN_FLAG_EXACT_PATH).
//
// Work is handed out by a global atomic counter over a candidate list that keeps the cache
// order (same child, shared leading parents adjacent), so warps that run concurrently read the
// same columns and hit in L1/L2.
#pragma once

#include <cuda_runtime.h>

#include "plan.cuh"
#include "rowops.cuh"
#include "solver.cuh"

namespace ABN_NS {

// tunables (overridable at build time for sweeps: make EXTRA="-DABN_FIT_WARPS=.. -DABN_MIN_BLOCKS=.. -DABN_UNROLL=..")
#ifndef ABN_FIT_WARPS
#define ABN_FIT_WARPS 8
#endif
#ifndef ABN_MIN_BLOCKS
#define ABN_MIN_BLOCKS 2
#endif
#ifndef ABN_UNROLL
#define ABN_UNROLL 4
#endif
constexpr int FIT_WARPS = ABN_FIT_WARPS; // warps per CTA
constexpr int FIT_UNROLL = ABN_UNROLL;   // rows per lane in flight

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    return v;
}

// contribution of one row to the register accumulators
template <int KIND, int K, bool ICPT, bool NEED_H, bool NEED_G, bool NEED_L, bool VOTE = false>
__device__ __forceinline__ void row_update(const double (&xd)[K], double yi, const double (&beta)[K], const RowConst &rc,
                                           double (&H)[K * (K + 1) / 2], double (&g)[K], double &ll, RowTrack &bad)
{
    // xd[j] holds design column j; with ICPT column 0 is the constant 1 and xd[0] is unused
    double eta = ICPT ? beta[0] : 0.0;
#pragma unroll
    for (int j = (ICPT ? 1 : 0); j < K; j++) eta = fma(xd[j], beta[j], eta);
    double w, r, l = 0.0;
    // binomial mle: the rho correction is evaluated only when some row of this warp step needs it;
    // VOTE is set where all 32 lanes are known to be converged (full tiles), elsewhere rho is always on
    bool need_rho = true;
    if (KIND == RK_BIN_MLE && VOTE && (NEED_H || NEED_G)) need_rho = __any_sync(0xffffffffu, eta >= BIN_RHO_ETA);
    RowOp<KIND>::template apply<NEED_H, NEED_G, NEED_L>(eta, yi, rc, w, r, l, bad, need_rho);
    if (NEED_H) {
        int t = 0;
#pragma unroll
        for (int p = 0; p < K; p++) {
            const double wx = (ICPT && p == 0) ? w : w * xd[p];
#pragma unroll
            for (int q = p; q < K; q++) {
                H[t] = (ICPT && q == 0) ? (H[t] + wx) : fma(wx, xd[q], H[t]);
                t++;
            }
        }
    }
    if (NEED_G) {
#pragma unroll
        for (int p = 0; p < K; p++) g[p] = (ICPT && p == 0) ? (g[p] + r) : fma(r, xd[p], g[p]);
    }
    if (NEED_L) ll += l;
}

// one pass of one fit by one warp. MODE is PASS_FULL / PASS_HG / PASS_GRAD / PASS_HESS / PASS_LL.
// Lane l owns rows l, l+32, ...; FIT_UNROLL rows per lane form one iteration and the loads of the
// next iteration are issued before the current one is computed (register double buffering), so the
// L2 / HBM latency of the column reads is covered by a whole iteration of FP64 work. Rows are added
// to the accumulators in increasing order, whatever FIT_UNROLL is.
template <int KIND, int K, bool ICPT, int MODE>
__device__ __forceinline__ void warp_pass(const double *__restrict__ y, const double *const (&cp)[MAXK],
                                          const double *beta_sm, long long n, const RowConst rc, PassAcc *acc_sm)
{
    constexpr int KD = K - (ICPT ? 1 : 0); // streamed design columns
    constexpr int I0 = ICPT ? 1 : 0;
    constexpr int U = FIT_UNROLL;
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
    RowTrack tk = row_track_init();
#pragma unroll
    for (int j = 0; j < K * (K + 1) / 2; j++) H[j] = 0.0;
#pragma unroll
    for (int j = 0; j < K; j++) g[j] = 0.0;

    const long long it_rows = 32LL * U;
    const long long n_main = (n / it_rows) * it_rows;
    double xn[U][K], yn[U]; // next iteration's rows
    if (n_main > 0) {
#pragma unroll
        for (int u = 0; u < U; u++) {
#pragma unroll
            for (int j = 0; j < KD; j++) xn[u][j + I0] = __ldg(cp[j] + lane + 32 * u);
            yn[u] = __ldg(y + lane + 32 * u);
        }
    }
    for (long long base = lane; base < n_main; base += it_rows) {
        double xc[U][K], yc[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
#pragma unroll
            for (int j = I0; j < K; j++) xc[u][j] = xn[u][j];
            yc[u] = yn[u];
        }
        const long long nb = base + it_rows;
        if (nb < n_main) {
#pragma unroll
            for (int u = 0; u < U; u++) {
#pragma unroll
                for (int j = 0; j < KD; j++) xn[u][j + I0] = __ldg(cp[j] + nb + 32 * u);
                yn[u] = __ldg(y + nb + 32 * u);
            }
        }
#pragma unroll
        for (int u = 0; u < U; u++)
            row_update<KIND, K, ICPT, NEED_H, NEED_G, NEED_L, true>(xc[u], yc[u], beta, rc, H, g, ll, tk);
    }
    for (long long i = n_main + lane; i < n; i += 32) { // tail: fewer than 32 * U rows
        double xt[K];
#pragma unroll
        for (int j = 0; j < KD; j++) xt[j + I0] = __ldg(cp[j] + i);
        row_update<KIND, K, ICPT, NEED_H, NEED_G, NEED_L>(xt, __ldg(y + i), beta, rc, H, g, ll, tk);
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
    const int bad = __any_sync(0xffffffffu, RowOp<KIND>::bad_of(tk));
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
        acc_sm->has_gl = 0;
    }
}


#ifndef ABN_BIN_NOY
#define ABN_BIN_NOY 1 // binomial passes that do not stream the response (warp_pass_bin)
#endif
#ifndef ABN_STAGE
#define ABN_STAGE 1   // 1: stage the column tiles through shared memory with cp.async; 0: register double buffering
#endif
constexpr int FIT_STAGES = 2;

#ifndef ABN_LEAN_STAGES
#define ABN_LEAN_STAGES 2
#endif
// doubles of tile storage one warp needs: the staged pass (FIT_STAGES x (KD + 1) columns x 32 FIT_UNROLL rows)
// or the lean pass (ABN_LEAN_STAGES x KD columns x 128 rows), whichever is larger
template <int K, bool ICPT>
constexpr size_t fit_stage_doubles_per_warp()
{
    const size_t staged = ABN_STAGE ? (size_t)FIT_STAGES * (K - (ICPT ? 1 : 0) + 1) * 32 * FIT_UNROLL : 0;
    const size_t lean = ICPT ? (size_t)ABN_LEAN_STAGES * (K - 1) * 128 : 0;
    return staged > lean ? staged : lean;
}
// bytes of dynamic shared memory the passes need per CTA
template <int K, bool ICPT>
constexpr size_t fit_stage_bytes()
{
    return (size_t)FIT_WARPS * fit_stage_doubles_per_warp<K, ICPT>() * sizeof(double);
}

__device__ __forceinline__ void cp_async8(double *smem_dst, const double *gsrc)
{
    const unsigned dst = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Staged variant of warp_pass: every lane copies the rows it owns (FIT_UNROLL rows x (KD+1)
// columns per iteration) into its warp's shared-memory tile with cp.async, one iteration ahead,
// and reads back only what it copied itself - so the only synchronisation is the lane's own
// cp.async.wait_group, no registers are held by loads in flight, and the FP64 work of a whole
// iteration covers the L2 / HBM latency. Same row order and reduction as the register variant.
template <int KIND, int K, bool ICPT, int MODE>
__device__ __forceinline__ void warp_pass_staged(const double *__restrict__ y, const double *const (&cp)[MAXK],
                                                 const double *beta_sm, long long n, const RowConst rc,
                                                 PassAcc *acc_sm, double *tile)
{
    constexpr int KD = K - (ICPT ? 1 : 0);
    constexpr int I0 = ICPT ? 1 : 0;
    constexpr int U = FIT_UNROLL;
    constexpr int NC = KD + 1;          // staged columns: design columns + response
    constexpr int TROWS = 32 * U;       // rows per iteration
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
    RowTrack tk = row_track_init();
#pragma unroll
    for (int j = 0; j < K * (K + 1) / 2; j++) H[j] = 0.0;
#pragma unroll
    for (int j = 0; j < K; j++) g[j] = 0.0;

    const int ntiles = (int)(n / TROWS);
    const long long n_main = (long long)ntiles * TROWS;
    // tile layout: [stage][column][row in tile]; lane touches rows lane + 32 u only. The source
    // pointers advance by one tile per issue (no per-tile address arithmetic beyond K adds).
    const double *pc[KD > 0 ? KD : 1];
#pragma unroll
    for (int j = 0; j < KD; j++) pc[j] = cp[j] + lane;
    const double *py = y + lane;
    double *const tlane = tile + lane;
    auto issue = [&](int stage) {
        double *t = tlane + stage * (NC * TROWS);
#pragma unroll
        for (int u = 0; u < U; u++) {
#pragma unroll
            for (int j = 0; j < KD; j++) cp_async8(t + j * TROWS + u * 32, pc[j] + u * 32);
            cp_async8(t + KD * TROWS + u * 32, py + u * 32);
        }
        cp_async_commit();
#pragma unroll
        for (int j = 0; j < KD; j++) pc[j] += TROWS;
        py += TROWS;
    };
    if (ntiles > 0) issue(0);
    int stage = 0;
    for (int it = 0; it < ntiles; it++) {
        if (it + 1 < ntiles) { issue(stage ^ 1); cp_async_wait<1>(); }
        else cp_async_wait<0>();
        const double *t = tlane + stage * (NC * TROWS);
#pragma unroll
        for (int u = 0; u < U; u++) {
            double xr[K];
#pragma unroll
            for (int j = 0; j < KD; j++) xr[j + I0] = t[j * TROWS + u * 32];
            const double yi = t[KD * TROWS + u * 32];
            row_update<KIND, K, ICPT, NEED_H, NEED_G, NEED_L, true>(xr, yi, beta, rc, H, g, ll, tk);
        }
        stage ^= 1;
    }
    for (long long i = n_main + lane; i < n; i += 32) { // tail: fewer than 32 * U rows
        double xt[K];
#pragma unroll
        for (int j = 0; j < KD; j++) xt[j + I0] = __ldg(cp[j] + i);
        row_update<KIND, K, ICPT, NEED_H, NEED_G, NEED_L>(xt, __ldg(y + i), beta, rc, H, g, ll, tk);
    }

    if (NEED_H) {
#pragma unroll
        for (int j = 0; j < K * (K + 1) / 2; j++) H[j] = warp_sum(H[j]);
    }
    if (NEED_G) {
#pragma unroll
        for (int j = 0; j < K; j++) g[j] = warp_sum(g[j]);
    }
    if (NEED_L) ll = warp_sum(ll);
    const int bad = __any_sync(0xffffffffu, RowOp<KIND>::bad_of(tk));
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
        acc_sm->has_gl = 0;
    }
}


// Poisson with an intercept: the score and the log-likelihood need no sums of their own.
//   w = E = exp(eta), r = y - E  =>  g_j = (X'y)_j - sum E x_j = (X'y)_j - H_0j   (column 0 is the intercept)
//   ll = sum (y eta - E) = beta'(X'y) - H_00                                       (lfact is subtracted later)
// and X'y is a column of the global Gram matrix. So the pass streams only the design columns (not
// the response), evaluates one exp per row, and accumulates H - all of it (FULL / HG / HESS), its
// first row (GRAD) or H_00 alone (LL). `xty` holds (X'y)_j for this fit.
template <int K, int MODE>
__device__ __forceinline__ void warp_pass_pois(const double *const (&cp)[MAXK], const double *beta_sm, const double *xty,
                                               long long n, const RowConst rc, PassAcc *acc_sm, double *tile)
{
    constexpr int KD = K - 1;
    constexpr int U = FIT_UNROLL;
    constexpr int TROWS = 32 * U;
    constexpr bool ALL_H = (MODE == PASS_FULL || MODE == PASS_HG || MODE == PASS_HESS);
    constexpr int NH = ALL_H ? K * (K + 1) / 2 : (MODE == PASS_GRAD ? K : 1);
    const int lane = threadIdx.x & 31;

    double beta[K];
#pragma unroll
    for (int j = 0; j < K; j++) beta[j] = beta_sm[j];
    double H[NH];
#pragma unroll
    for (int j = 0; j < NH; j++) H[j] = 0.0;
    RowTrack tk = row_track_init();

    auto row = [&](const double (&x)[K]) { // x[0] unused (intercept)
        double eta = beta[0];
#pragma unroll
        for (int j = 1; j < K; j++) eta = fma(x[j], beta[j], eta);
        row_track(tk, eta);
        const double E = fast_exp(clamp_lo700(eta), rc.tab);
        H[0] += E;
        if (ALL_H) {
            int t = 1;
#pragma unroll
            for (int q = 1; q < K; q++) { H[t] = fma(E, x[q], H[t]); t++; }
#pragma unroll
            for (int p = 1; p < K; p++) {
                const double wx = E * x[p];
#pragma unroll
                for (int q = p; q < K; q++) { H[t] = fma(wx, x[q], H[t]); t++; }
            }
        } else if (MODE == PASS_GRAD) {
#pragma unroll
            for (int q = 1; q < K; q++) H[q] = fma(E, x[q], H[q]);
        }
    };

    constexpr int NCP = KD > 0 ? KD : 1;
    const int ntiles = (KD > 0) ? (int)(n / TROWS) : 0;
    const long long n_main = (long long)ntiles * TROWS;
    const double *pc[NCP];
#pragma unroll
    for (int j = 0; j < KD; j++) pc[j] = cp[j] + lane;
    double *const tlane = tile + lane;
    auto issue = [&](int stage) {
        double *t = tlane + stage * (NCP * TROWS);
#pragma unroll
        for (int u = 0; u < U; u++)
#pragma unroll
            for (int j = 0; j < KD; j++) cp_async8(t + j * TROWS + u * 32, pc[j] + u * 32);
        cp_async_commit();
#pragma unroll
        for (int j = 0; j < KD; j++) pc[j] += TROWS;
    };
    if (ntiles > 0) issue(0);
    int stage = 0;
    for (int it = 0; it < ntiles; it++) {
        if (it + 1 < ntiles) { issue(stage ^ 1); cp_async_wait<1>(); }
        else cp_async_wait<0>();
        const double *t = tlane + stage * (NCP * TROWS);
#pragma unroll
        for (int u = 0; u < U; u++) {
            double xr[K];
#pragma unroll
            for (int j = 0; j < KD; j++) xr[j + 1] = t[j * TROWS + u * 32];
            row(xr);
        }
        stage ^= 1;
    }
    for (long long i = n_main + lane; i < n; i += 32) {
        double xt[K];
#pragma unroll
        for (int j = 0; j < KD; j++) xt[j + 1] = __ldg(cp[j] + i);
        row(xt);
    }
#pragma unroll
    for (int j = 0; j < NH; j++) H[j] = warp_sum(H[j]);
    const int bad = __any_sync(0xffffffffu, RowOp<RK_POIS>::bad_of(tk));
    if (lane == 0) {
        if (ALL_H) {
#pragma unroll
            for (int j = 0; j < K * (K + 1) / 2; j++) acc_sm->H[j] = H[j];
        }
        if (MODE == PASS_FULL || MODE == PASS_HG || MODE == PASS_GRAD) {
            // H row 0 is H[0..K-1] in both layouts (packed upper triangle starts with row 0)
#pragma unroll
            for (int j = 0; j < K; j++) acc_sm->g[j] = xty[j] - H[j];
        }
        if (MODE == PASS_FULL || MODE == PASS_LL) {
            double s = -H[0];
#pragma unroll
            for (int j = 0; j < K; j++) s = fma(beta[j], xty[j], s);
            acc_sm->ll = s;
        } else {
            acc_sm->ll = 0.0;
        }
        acc_sm->bad = bad;
        acc_sm->has_gl = 0;
    }
}


// Binomial child with an intercept: the response column is not streamed either.
//   mle   r_i = rho_i (b'_i - g_i), b'_i = bp0 + y_i (bp1 - bp0), l_i = -log(1 + e_i) - (1 - y_i) eta_i
//         X'r = bp0 X'1 + (bp1 - bp0) X'y - S,   S_j = sum_i gv_i x_ij,   gv = g - (rho - 1)(b' - g)
//         ll  = beta'(X'y) - beta'(X'1) - sum_i log(1 + e_i)
//   bayes r_i = y_i - sigma_i, l_i = y_i eta_i - log(1 + E_i)   (eta_i for eta_i >= 700)
//         X'r = X'y - S,  S_j = sum_i sigma_i x_ij;   ll = beta'(X'y) - sum_i log(1 + E_i)
// with X'y and X'1 columns of the global Gram matrix (`xty`, `xt1`). The mle rho correction (see
// RowOp<RK_BIN_MLE>) is taken only in warp steps where some row has eta >= BIN_RHO_ETA; those
// steps fetch y straight from global memory, form g with the correctly rounded reciprocal and
// W = rho g' exactly as the row op does. Elsewhere rho = 1 to 2e-11, g comes from the cubic
// reciprocal step and g' = g - g^2 (equal to e g^2 to 1e-16 / (1 - g) <= 2e-11 relative; g'
// only weights the Hessian of the IRLS step, never the fixed point).
template <int KIND, int K, int MODE>
__device__ __forceinline__ void warp_pass_bin(const double *__restrict__ y, const double *const (&cp)[MAXK],
                                              const double *beta_sm, const double *xty, const double *xt1, long long n,
                                              const RowConst rc, PassAcc *acc_sm, double *tile)
{
    constexpr int KD = K - 1;
    constexpr int U = FIT_UNROLL;
    constexpr int TROWS = 32 * U;
    constexpr bool NEED_H = (MODE == PASS_FULL || MODE == PASS_HG || MODE == PASS_HESS);
    constexpr bool NEED_G = (MODE == PASS_FULL || MODE == PASS_HG || MODE == PASS_GRAD);
    constexpr bool NEED_L = (MODE == PASS_FULL || MODE == PASS_LL);
    constexpr bool MLE_ = (KIND == RK_BIN_MLE);
    const int lane = threadIdx.x & 31;

    double beta[K];
#pragma unroll
    for (int j = 0; j < K; j++) beta[j] = beta_sm[j];
    double H[K * (K + 1) / 2];
    double S[K];
    double L = 0.0;
#pragma unroll
    for (int j = 0; j < K * (K + 1) / 2; j++) H[j] = 0.0;
#pragma unroll
    for (int j = 0; j < K; j++) S[j] = 0.0;
    RowTrack tk = row_track_init();
    const double dbp = rc.bp1 - rc.bp0;
    int any_rho = 0; // some row of this pass had eta >= BIN_RHO_ETA: the reference's rho noise is in play

    // x[0] unused (intercept); yp points at this row's response, read only on the rho path
    auto row = [&](const double (&x)[K], const double *yp, bool vote) {
        double eta = beta[0];
#pragma unroll
        for (int j = 1; j < K; j++) eta = fma(x[j], beta[j], eta);
        double w = 0.0, gv = 0.0;
        if (MLE_) {
            row_track(tk, eta);
            const double e = fast_exp(-eta, rc.tab);
            const double d = 1.0 + e;
            if (NEED_H || NEED_G) {
                const bool need_rho = vote ? __any_sync(0xffffffffu, eta >= BIN_RHO_ETA) : true;
                any_rho |= vote ? (int)need_rho : (int)(eta >= BIN_RHO_ETA);
                if (need_rho) {
                    const double g = fast_rcp(d);
                    const double gp = e * g * g;
                    const double rho = gp * fast_rcp_approx(g * (1.0 - g));
                    const double bpg = fma(__ldg(yp), dbp, rc.bp0) - g;
                    w = rho * gp;
                    gv = fma(-(rho - 1.0), bpg, g);
                } else {
                    const double g = fast_rcp_approx(d);
                    w = fma(-g, g, g);
                    gv = g;
                }
            }
            if (NEED_L) L += fast_log_ge1(d, rc.logtab);
        } else {
            const double E = fast_exp(clamp700(eta), rc.tab);
            const double d = 1.0 + E;
            if (NEED_H || NEED_G) {
                const double inv = fast_rcp_approx(d);
                gv = E * inv;
                w = gv * inv;
            }
            if (NEED_L) L += (hi_word(eta) >= 0x4085E000) ? eta : fast_log_ge1(d, rc.logtab);
        }
        if (NEED_H) {
            H[0] += w;
            int t = 1;
#pragma unroll
            for (int q = 1; q < K; q++) { H[t] = fma(w, x[q], H[t]); t++; }
#pragma unroll
            for (int p = 1; p < K; p++) {
                const double wx = w * x[p];
#pragma unroll
                for (int q = p; q < K; q++) { H[t] = fma(wx, x[q], H[t]); t++; }
            }
        }
        if (NEED_G) {
            S[0] += gv;
#pragma unroll
            for (int q = 1; q < K; q++) S[q] = fma(gv, x[q], S[q]);
        }
    };

    constexpr int NCP = KD > 0 ? KD : 1;
    const int ntiles = (KD > 0) ? (int)(n / TROWS) : 0;
    const long long n_main = (long long)ntiles * TROWS;
    const double *pc[NCP];
#pragma unroll
    for (int j = 0; j < KD; j++) pc[j] = cp[j] + lane;
    double *const tlane = tile + lane;
    auto issue = [&](int stage) {
        double *t = tlane + stage * (NCP * TROWS);
#pragma unroll
        for (int u = 0; u < U; u++)
#pragma unroll
            for (int j = 0; j < KD; j++) cp_async8(t + j * TROWS + u * 32, pc[j] + u * 32);
        cp_async_commit();
#pragma unroll
        for (int j = 0; j < KD; j++) pc[j] += TROWS;
    };
    if (ntiles > 0) issue(0);
    int stage = 0;
    const double *yrow = y + lane;
    for (int it = 0; it < ntiles; it++) {
        if (it + 1 < ntiles) { issue(stage ^ 1); cp_async_wait<1>(); }
        else cp_async_wait<0>();
        const double *t = tlane + stage * (NCP * TROWS);
#pragma unroll
        for (int u = 0; u < U; u++) {
            double xr[K];
#pragma unroll
            for (int j = 0; j < KD; j++) xr[j + 1] = t[j * TROWS + u * 32];
            row(xr, yrow + u * 32, true);
        }
        yrow += TROWS;
        stage ^= 1;
    }
    for (long long i = n_main + lane; i < n; i += 32) {
        double xt[K];
#pragma unroll
        for (int j = 0; j < KD; j++) xt[j + 1] = __ldg(cp[j] + i);
        row(xt, y + i, false);
    }
    if (NEED_H) {
#pragma unroll
        for (int j = 0; j < K * (K + 1) / 2; j++) H[j] = warp_sum(H[j]);
    }
    if (NEED_G) {
#pragma unroll
        for (int j = 0; j < K; j++) S[j] = warp_sum(S[j]);
    }
    if (NEED_L) L = warp_sum(L);
    const int bad = __any_sync(0xffffffffu, RowOp<KIND>::bad_of(tk));
    any_rho = __any_sync(0xffffffffu, any_rho);
    if (lane == 0) {
        if (NEED_H) {
#pragma unroll
            for (int j = 0; j < K * (K + 1) / 2; j++) acc_sm->H[j] = H[j];
        }
        if (NEED_G) {
#pragma unroll
            for (int j = 0; j < K; j++) {
                acc_sm->g[j] = MLE_ ? fma(dbp, xty[j], rc.bp0 * xt1[j]) - S[j] : xty[j] - S[j];
                if (MLE_) acc_sm->gl[j] = xty[j] - S[j]; // X'(y - g) up to the rho correction (~1e-11 of it)
            }
        }
        double ll = 0.0;
        if (NEED_L) {
            ll = -L;
#pragma unroll
            for (int j = 0; j < K; j++) ll = fma(beta[j], MLE_ ? xty[j] - xt1[j] : xty[j], ll);
        }
        acc_sm->ll = ll;
        acc_sm->bad = bad;
        acc_sm->has_gl = (MLE_ && NEED_G && !any_rho) ? 1 : 0; // solver.cuh: shortcuts only for fits free of rho noise
    }
}

// ---- lean passes ---------------------------------------------------------------------------------
// The passes above are bound by instruction issue as much as by the FP64 pipe: per row they spend about as
// many integer / load-store / branch instructions as FP64 ones (profiles/sass_r1_*), and a warp-wide FP64
// instruction already occupies the pipe for two issue cycles. The lean passes do the same sums with the
// non-FP64 work cut to less than half, so that the FP64 pipe alone sets the pace:
//   * 16-byte cp.async and 16-byte shared reads: lane l owns rows 64 c + 2 l + {0, 1} of each 128-row tile
//     (half the copy / load instructions), one 32-bit row offset instead of a pointer per column;
//   * fast_exp256: 9 FP64 + 4 integer / LDS instructions instead of 10 + 7;
//   * no clamps, no per-row branches, no votes: the extreme high words of eta are tracked with 3-input
//     integer max (one VIMNMX3 per two rows and bound) and checked once after the pass. A pass that met an
//     eta outside the range the fast row ops are exact for - eta >= 12 for the bias-reduced binomial fit
//     (the rows where the reference's rho noise is visible, see warp_pass_bin), |eta| >= 700 or NaN
//     elsewhere - returns false, and the careful pass above redoes it (and all later passes of that fit).
//     On BASELINE-like data no fit ever does.
//   (Measured and dropped, profiles/ab_r2_variants.txt: 16-byte read-only loads straight into registers instead of
//   the shared-memory tiles. It halves the shared-memory traffic, but only one tile of registers fits under the
//   128-register budget, so the loads are no longer covered: 147-152 k fits/s against 212 k, SM clocks at their
//   maximum - the kernel waits on L2 instead of computing.)
//   * binomial, both methods: sigma = g = 1/(1+e), e = exp(-eta); w = g - g^2; sum log(1+E) = sum log(1+e)
//     + beta'(X'1), so the Bayes passes share the mle row op and the only per-row difference is gone.
// Rows are summed in a fixed order that depends on nothing but n, so results stay bit-identical under any
// partition of the candidate list.
#ifndef ABN_LEAN
#define ABN_LEAN 1
#endif
constexpr int LEAN_STAGES = ABN_LEAN_STAGES;
// exp of the lean row ops: 0 = 256-entry table (fast_exp256, 9 FP64 instructions, a conflict-prone table read),
// 1 = 64-entry table in 16 bank-interleaved replicas (fast_exp64r, conflict-free read, 10 FP64 instructions). The
// replicas take the shared-memory pipe of the Poisson pass from 68 % to 58 % busy (ncu), but the extra FP64
// instruction costs more than that buys
#ifndef ABN_EXP_REPLICATED
#define ABN_EXP_REPLICATED 0 // measured on one box (profiles/ab_r2_variants.txt): 206.6 k fits/s replicated, 216.8 k with 256 entries
#endif
#if ABN_EXP_REPLICATED
#define ABN_LEAN_EXP fast_exp64r
constexpr int LEAN_EXP_DOUBLES = ABN_EXP2_N * ABN_EXP64R_COPIES;
#else
// ABN_LEAN_EXP1 = 1: one reduction step (fast_exp256s, 8 FP64 instructions; see fastmath.cuh for why that is harmless)
#ifndef ABN_LEAN_EXP1
#define ABN_LEAN_EXP1 1
#endif
#if ABN_LEAN_EXP1
#define ABN_LEAN_EXP fast_exp256s
#else
#define ABN_LEAN_EXP fast_exp256
#endif
constexpr int LEAN_EXP_DOUBLES = ABN_EXP256_N;
#endif
// reciprocal of the lean binomial row op: 1 = one Newton step on the MUFU seed (2 FP64), 0 = the cubic step (3 FP64)
#ifndef ABN_LEAN_RCP1
#define ABN_LEAN_RCP1 1
#endif
#if ABN_LEAN_RCP1
#define ABN_LEAN_RCP fast_rcp_newton1
#else
#define ABN_LEAN_RCP fast_rcp_approx
#endif
constexpr int LEAN_TROWS = 128;

__device__ __forceinline__ void cp_async16(double *smem_dst, const double *gsrc)
{
    const unsigned dst = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(gsrc) : "memory");
}

// Streams the KD design columns of one fit through the warp's tiles and calls row(x) for every observation,
// x[j] = design column j + 1; row returns the high word of its eta, whose extremes end up in tk.
template <int KD, typename RowFn>
__device__ __forceinline__ void lean_stream(const double *const (&cp)[MAXK], long long n, double *tile, RowTrack &tk,
                                            RowFn &&row)
{
    constexpr int TR = LEAN_TROWS, ST = LEAN_STAGES;
    constexpr int KDA = KD > 0 ? KD : 1;
    const int lane = threadIdx.x & 31;
    const int ntiles = (KD > 0) ? (int)(n / TR) : 0;
    const double *base[KDA];
#pragma unroll
    for (int j = 0; j < KD; j++) base[j] = cp[j] + 2 * lane;
    double *const tl = tile + 2 * lane;
    unsigned off = 0; // first row of the next tile to issue
    auto issue = [&](int stage) {
        double *t = tl + stage * (KD * TR);
#pragma unroll
        for (int j = 0; j < KD; j++) {
            const double *p = base[j] + off;
            cp_async16(t + j * TR, p);
            cp_async16(t + j * TR + 64, p + 64);
        }
        off += TR;
    };
    // one commit group per tile slot, empty once the tiles run out, so that wait_group<ST - 1> always
    // means "the tile of this iteration has landed"
#pragma unroll
    for (int s = 0; s < ST - 1; s++) {
        if (s < ntiles) issue(s);
        cp_async_commit();
    }
    int stage = 0, fill = ST - 1;
    for (int it = 0; it < ntiles; it++) {
        if (it + ST - 1 < ntiles) issue(fill);
        cp_async_commit();
        cp_async_wait<ST - 1>();
        const double *t = tl + stage * (KD * TR);
        double x[4][KDA];
#pragma unroll
        for (int j = 0; j < KD; j++) {
            const double2 a = *reinterpret_cast<const double2 *>(t + j * TR);
            const double2 b = *reinterpret_cast<const double2 *>(t + j * TR + 64);
            x[0][j] = a.x; x[1][j] = a.y; x[2][j] = b.x; x[3][j] = b.y;
        }
        const int h0 = row(x[0]), h1 = row(x[1]), h2 = row(x[2]), h3 = row(x[3]);
        tk.smax = max(max(tk.smax, h0), h1);
        tk.umax = max(max(tk.umax, (unsigned)h0), (unsigned)h1);
        tk.smax = max(max(tk.smax, h2), h3);
        tk.umax = max(max(tk.umax, (unsigned)h2), (unsigned)h3);
        fill = stage;
        stage = (stage + 1 == ST) ? 0 : stage + 1;
    }
    cp_async_wait<0>();
    for (long long i = (long long)ntiles * TR + lane; i < n; i += 32) { // fewer than 128 rows (all of them for KD = 0)
        double xt[KDA];
#pragma unroll
        for (int j = 0; j < KD; j++) xt[j] = __ldg(cp[j] + i);
        const int h = row(xt);
        tk.smax = max(tk.smax, h);
        tk.umax = max(tk.umax, (unsigned)h);
    }
}

// ---- the same tiles moved by the TMA unit ------------------------------------------------------------------
// ABN_LEAN_TMA = 1: instead of 2 KD cp.async per lane and tile, lane 0 of the warp issues KD bulk copies
// (cp.async.bulk, 1 KB each: the 128 rows of one column) that complete on the warp's own mbarrier, and every lane
// waits on that barrier's phase. The copies do not pass through the load-store unit at all: no address
// arithmetic in the other 31 lanes and - what matters for the narrow Poisson designs - no shared-memory store
// wavefronts from the LSU, which together with the loads saturate the SM's 128 B/cycle shared-memory pipe before
// the FP64 pipe (ncu, round 2). Same tile layout, same row order, same bits.
#ifndef ABN_LEAN_TMA
#define ABN_LEAN_TMA 0
#endif

// Queue discipline of the streaming kernel. 0: every warp pulls its next fit on its own. 1: the CTA pulls FIT_WARPS
// CONSECUTIVE fits at a time and its warps start them together: neighbours in cache order share the child and all
// but the last parent, so the warps of a group read the same column tiles at about the same time and all but the
// first find them in L1 (a warp that runs ahead misses and is caught up by the ones that hit). The price is the
// barrier at the end of a group: warps whose fit needed fewer passes wait for the slowest of the eight.
#ifndef ABN_GROUP_FETCH
#define ABN_GROUP_FETCH 0
#endif

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(double *smem_dst, const double *gsrc, unsigned bytes, unsigned long long *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity)
{
    unsigned ok;
    do {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!ok);
}

// per-warp pipeline state that outlives a pass: the barriers (shared memory) and their phase bits
struct TmaPipe {
    unsigned long long *bar; // LEAN_STAGES barriers of this warp
    unsigned phase;          // bit s = parity the next wait on stage s expects
};

template <int KD, typename RowFn>
__device__ __forceinline__ void lean_stream_tma(const double *const (&cp)[MAXK], long long n, double *tile, TmaPipe &pipe,
                                                RowTrack &tk, RowFn &&row)
{
    constexpr int TR = LEAN_TROWS, ST = LEAN_STAGES;
    constexpr int KDA = KD > 0 ? KD : 1;
    const int lane = threadIdx.x & 31;
    const int ntiles = (KD > 0) ? (int)(n / TR) : 0;
    auto issue = [&](int stage, int t) { // lane 0 only
        double *dst = tile + stage * (KD * TR);
        mbar_expect_tx(pipe.bar + stage, KD * TR * (unsigned)sizeof(double));
#pragma unroll
        for (int j = 0; j < KD; j++) bulk_g2s(dst + j * TR, cp[j] + (long long)t * TR, TR * (unsigned)sizeof(double), pipe.bar + stage);
    };
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < ST - 1; s++)
            if (s < ntiles) issue(s, s);
    }
    int stage = 0, fill = ST - 1;
    for (int it = 0; it < ntiles; it++) {
        // the stage refilled here was read one iteration ago and its values have been consumed since
        __syncwarp();
        if (lane == 0 && it + ST - 1 < ntiles) issue(fill, it + ST - 1);
        mbar_wait(pipe.bar + stage, (pipe.phase >> stage) & 1u);
        pipe.phase ^= 1u << stage;
        const double *t = tile + stage * (KD * TR) + 2 * lane;
        double x[4][KDA];
#pragma unroll
        for (int j = 0; j < KD; j++) {
            const double2 a = *reinterpret_cast<const double2 *>(t + j * TR);
            const double2 b = *reinterpret_cast<const double2 *>(t + j * TR + 64);
            x[0][j] = a.x; x[1][j] = a.y; x[2][j] = b.x; x[3][j] = b.y;
        }
        const int h0 = row(x[0]), h1 = row(x[1]), h2 = row(x[2]), h3 = row(x[3]);
        tk.smax = max(max(tk.smax, h0), h1);
        tk.umax = max(max(tk.umax, (unsigned)h0), (unsigned)h1);
        tk.smax = max(max(tk.smax, h2), h3);
        tk.umax = max(max(tk.umax, (unsigned)h2), (unsigned)h3);
        fill = stage;
        stage = (stage + 1 == ST) ? 0 : stage + 1;
    }
    __syncwarp();
    for (long long i = (long long)ntiles * TR + lane; i < n; i += 32) {
        double xt[KDA];
#pragma unroll
        for (int j = 0; j < KD; j++) xt[j] = __ldg(cp[j] + i);
        const int h = row(xt);
        tk.smax = max(tk.smax, h);
        tk.umax = max(tk.umax, (unsigned)h);
    }
}

template <int KD, typename RowFn>
__device__ __forceinline__ void lean_rows(const double *const (&cp)[MAXK], long long n, double *tile, TmaPipe &pipe,
                                          RowTrack &tk, RowFn &&row)
{
#if ABN_LEAN_TMA
    lean_stream_tma<KD>(cp, n, tile, pipe, tk, row);
#else
    (void)pipe;
    lean_stream<KD>(cp, n, tile, tk, row);
#endif
}

// Poisson with an intercept (see warp_pass_pois for the identities). false: eta left (-700, 700), nothing written.
template <int K, int MODE>
__device__ __forceinline__ bool lean_pass_pois(const double *const (&cp)[MAXK], const double *beta_sm, const double *xty,
                                               long long n, const double *tab256, PassAcc *acc_sm, double *tile, TmaPipe &pipe)
{
    constexpr int KD = K - 1;
    constexpr int KDA = KD > 0 ? KD : 1;
    constexpr bool ALL_H = (MODE == PASS_FULL || MODE == PASS_HG || MODE == PASS_HESS);
    constexpr int NH = ALL_H ? K * (K + 1) / 2 : (MODE == PASS_GRAD ? K : 1);
    const int lane = threadIdx.x & 31;
    double beta[K];
#pragma unroll
    for (int j = 0; j < K; j++) beta[j] = beta_sm[j];
    double H[NH];
#pragma unroll
    for (int j = 0; j < NH; j++) H[j] = 0.0;
    RowTrack tk = row_track_init();
    auto row = [&](const double (&x)[KDA]) -> int {
        double eta = beta[0];
#pragma unroll
        for (int j = 0; j < KD; j++) eta = fma(x[j], beta[j + 1], eta);
        const double E = ABN_LEAN_EXP(eta, tab256);
        H[0] += E;
        if (ALL_H) {
            int t = 1;
#pragma unroll
            for (int q = 0; q < KD; q++) { H[t] = fma(E, x[q], H[t]); t++; }
#pragma unroll
            for (int p = 0; p < KD; p++) {
                const double wx = E * x[p];
#pragma unroll
                for (int q = p; q < KD; q++) { H[t] = fma(wx, x[q], H[t]); t++; }
            }
        } else if (MODE == PASS_GRAD) {
#pragma unroll
            for (int q = 0; q < KD; q++) H[q + 1] = fma(E, x[q], H[q + 1]);
        }
        return hi_word(eta);
    };
    lean_rows<KD>(cp, n, tile, pipe, tk, row);
    if (__any_sync(0xffffffffu, (tk.smax >= 0x4085E000) | (tk.umax >= 0xC085E000u))) return false; // |eta| >= 700, NaN
#pragma unroll
    for (int j = 0; j < NH; j++) H[j] = warp_sum(H[j]);
    if (lane == 0) {
        if (ALL_H) {
#pragma unroll
            for (int j = 0; j < K * (K + 1) / 2; j++) acc_sm->H[j] = H[j];
        }
        if (MODE == PASS_FULL || MODE == PASS_HG || MODE == PASS_GRAD) {
#pragma unroll
            for (int j = 0; j < K; j++) acc_sm->g[j] = xty[j] - H[j];
        }
        if (MODE == PASS_FULL || MODE == PASS_LL) {
            double s = -H[0];
#pragma unroll
            for (int j = 0; j < K; j++) s = fma(beta[j], xty[j], s);
            acc_sm->ll = s;
        } else {
            acc_sm->ll = 0.0;
        }
        acc_sm->bad = 0;
        acc_sm->has_gl = 0;
    }
    return true;
}

// Binomial child with an intercept, either method (see warp_pass_bin for the identities; the Bayes log term
// uses sum log(1+E) = sum log(1+e) + beta'(X'1)). false: some eta outside the range the fast row op is exact
// for - mle: eta >= 12 (rho rows) or <= -354.89; bayes: |eta| >= 700; NaN either way.
template <int KIND, int K, int MODE>
__device__ __forceinline__ bool lean_pass_bin(const double *const (&cp)[MAXK], const double *beta_sm, const double *xty,
                                              const double *xt1, long long n, const RowConst rc, const double *tab256,
                                              PassAcc *acc_sm, double *tile, TmaPipe &pipe)
{
    constexpr int KD = K - 1;
    constexpr int KDA = KD > 0 ? KD : 1;
    constexpr bool NEED_H = (MODE == PASS_FULL || MODE == PASS_HG || MODE == PASS_HESS);
    constexpr bool NEED_G = (MODE == PASS_FULL || MODE == PASS_HG || MODE == PASS_GRAD);
    constexpr bool NEED_L = (MODE == PASS_FULL || MODE == PASS_LL);
    constexpr bool MLE_ = (KIND == RK_BIN_MLE);
    constexpr int NHA = NEED_H ? K * (K + 1) / 2 : 1;
    const int lane = threadIdx.x & 31;
    double beta[K];
#pragma unroll
    for (int j = 0; j < K; j++) beta[j] = beta_sm[j];
    double H[NHA], S[K], L = 0.0;
#pragma unroll
    for (int j = 0; j < NHA; j++) H[j] = 0.0;
#pragma unroll
    for (int j = 0; j < K; j++) S[j] = 0.0;
    RowTrack tk = row_track_init();
    auto row = [&](const double (&x)[KDA]) -> int {
        double eta = beta[0];
#pragma unroll
        for (int j = 0; j < KD; j++) eta = fma(x[j], beta[j + 1], eta);
        const double e = ABN_LEAN_EXP(-eta, tab256);
        const double d = 1.0 + e;
        if (NEED_L) L += fast_log_ge1(d, rc.logtab);
        if (NEED_H || NEED_G) {
            const double g = ABN_LEAN_RCP(d);
            if (NEED_H) {
                const double w = fma(-g, g, g);
                H[0] += w;
                int t = 1;
#pragma unroll
                for (int q = 0; q < KD; q++) { H[t] = fma(w, x[q], H[t]); t++; }
#pragma unroll
                for (int p = 0; p < KD; p++) {
                    const double wx = w * x[p];
#pragma unroll
                    for (int q = p; q < KD; q++) { H[t] = fma(wx, x[q], H[t]); t++; }
                }
            }
            if (NEED_G) {
                S[0] += g;
#pragma unroll
                for (int q = 0; q < KD; q++) S[q + 1] = fma(g, x[q], S[q + 1]);
            }
        }
        return hi_word(eta);
    };
    lean_rows<KD>(cp, n, tile, pipe, tk, row);
    const int out_of_range = MLE_ ? ((tk.smax >= 0x40280000) | (tk.umax >= 0xC0762E42u))
                                  : ((tk.smax >= 0x4085E000) | (tk.umax >= 0xC085E000u));
    if (__any_sync(0xffffffffu, out_of_range)) return false;
    if (NEED_H) {
#pragma unroll
        for (int j = 0; j < NHA; j++) H[j] = warp_sum(H[j]);
    }
    if (NEED_G) {
#pragma unroll
        for (int j = 0; j < K; j++) S[j] = warp_sum(S[j]);
    }
    if (NEED_L) L = warp_sum(L);
    if (lane == 0) {
        if (NEED_H) {
#pragma unroll
            for (int j = 0; j < NHA; j++) acc_sm->H[j] = H[j];
        }
        if (NEED_G) {
            const double dbp = rc.bp1 - rc.bp0;
#pragma unroll
            for (int j = 0; j < K; j++) {
                acc_sm->g[j] = MLE_ ? fma(dbp, xty[j], rc.bp0 * xt1[j]) - S[j] : xty[j] - S[j];
                if (MLE_) acc_sm->gl[j] = xty[j] - S[j];
            }
        }
        double ll = 0.0;
        if (NEED_L) {
            ll = -L;
#pragma unroll
            for (int j = 0; j < K; j++) ll = fma(beta[j], xty[j] - xt1[j], ll);
        }
        acc_sm->ll = ll;
        acc_sm->bad = 0;
        acc_sm->has_gl = (MLE_ && NEED_G) ? 1 : 0;
    }
    return true;
}

template <int KIND, int K, bool ICPT, int MODE>
__device__ __forceinline__ void run_pass(const double *__restrict__ y, const double *const (&cp)[MAXK],
                                         const double *beta_sm, long long n, const RowConst rc, PassAcc *acc_sm,
                                         double *tile, const double *xty, const double *tab256, int &careful, TmaPipe &pipe)
{
#if ABN_LEAN
    if (ICPT && !careful && KIND != RK_PINIT) {
        bool ok;
        if (KIND == RK_POIS) ok = lean_pass_pois<K, MODE>(cp, beta_sm, xty, n, tab256, acc_sm, tile, pipe);
        else ok = lean_pass_bin<KIND == RK_BIN_MLE ? RK_BIN_MLE : RK_BIN_BAYES, K, MODE>(cp, beta_sm, xty, xty + MAXK, n, rc, tab256, acc_sm, tile, pipe);
        if (ok) return;
        careful = 1; // this fit leaves the range of the lean row ops: careful passes from here on
    }
#endif
#if ABN_STAGE
    if (KIND == RK_POIS && ICPT) {
        warp_pass_pois<K, MODE>(cp, beta_sm, xty, n, rc, acc_sm, tile);
        return;
    }
#if ABN_BIN_NOY
    if ((KIND == RK_BIN_MLE || KIND == RK_BIN_BAYES) && ICPT) {
        warp_pass_bin<KIND == RK_BIN_MLE ? RK_BIN_MLE : RK_BIN_BAYES, K, MODE>(y, cp, beta_sm, xty, xty + MAXK, n, rc, acc_sm, tile);
        return;
    }
#endif
    warp_pass_staged<KIND, K, ICPT, MODE>(y, cp, beta_sm, n, rc, acc_sm, tile);
#else
    warp_pass<KIND, K, ICPT, MODE>(y, cp, beta_sm, n, rc, acc_sm);
#endif
}

// `counter` points at the bucket's 16-int slot of abn_data::dCounters: [0] the work queue head, [2..15] seven
// 64-bit tallies of the data passes by PassMode (BUCKET_SLOT_INTS in abn_b200.cu).
// KIND: RK_BIN_MLE, RK_BIN_BAYES or RK_POIS (both methods; the mle start-value pass uses RK_PINIT)
template <int KIND, int K, bool ICPT>
__global__ void __launch_bounds__(FIT_WARPS * 32, ABN_MIN_BLOCKS)
fit_kernel(const DataMeta m, const Control c, const FitDesc *__restrict__ descs, const int *__restrict__ child,
           const int *__restrict__ list, int count, int *counter, ResultPtrs out)
{
    __shared__ FitState st_all[FIT_WARPS];
    __shared__ PassAcc acc_all[FIT_WARPS];
    __shared__ double xty_all[FIT_WARPS][2 * MAXK]; // (X'y)_j and (X'1)_j of the fit in flight (passes without y)
    __shared__ double exp_tab[ABN_EXP2_N];
    __shared__ double lnmu_tab[64];
    __shared__ double log_tab[128];
    __shared__ double exp256_tab[LEAN_EXP_DOUBLES]; // pre-scaled exp table of the lean passes (ABN_LEAN_EXP)
    if (threadIdx.x < 64) fill_log_table(log_tab, threadIdx.x);
    extern __shared__ __align__(16) double stage_tiles[]; // FIT_WARPS x fit_stage_doubles_per_warp
    if (threadIdx.x < ABN_EXP2_N) exp_tab[threadIdx.x] = c_exp2_table[threadIdx.x];
#if ABN_EXP_REPLICATED
    for (int i = threadIdx.x; i < LEAN_EXP_DOUBLES; i += FIT_WARPS * 32) {
        const int j = i / ABN_EXP64R_COPIES;
        exp256_tab[i] = exp64r_entry(c_exp2_table[j], j);
    }
#else
    for (int j = threadIdx.x; j < ABN_EXP256_N; j += FIT_WARPS * 32) exp256_tab[j] = exp256_entry(c_exp2_256_table[j], j);
#endif
    if (KIND == RK_POIS && threadIdx.x < 64) lnmu_tab[threadIdx.x] = log((double)threadIdx.x + 0.1);
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    FitState &st = st_all[warp];
    PassAcc *acc = &acc_all[warp];
    double *tile = stage_tiles + (size_t)warp * fit_stage_doubles_per_warp<K, ICPT>();
    double *xty = xty_all[warp];
    __shared__ unsigned long long tma_bars[FIT_WARPS][LEAN_STAGES];
    TmaPipe pipe{tma_bars[warp], 0u};
#if ABN_LEAN_TMA
    if (lane == 0) {
        for (int s = 0; s < LEAN_STAGES; s++) mbar_init(pipe.bar + s, 1u);
        mbar_fence_init();
    }
    __syncwarp();
#endif
    const double *lean_tab = exp256_tab + (ABN_EXP_REPLICATED ? (lane & (ABN_EXP64R_COPIES - 1)) : 0);

#if ABN_GROUP_FETCH
    __shared__ int group_base;
#endif
    for (;;) {
#if ABN_GROUP_FETCH
        __syncthreads(); // the previous group is finished (and its base has been read by everybody)
        if (threadIdx.x == 0) group_base = atomicAdd(counter, FIT_WARPS);
        __syncthreads();
        if (group_base >= count) break;
        const int idx = group_base + warp;
        if (idx >= count) continue;
#else
        int idx = 0;
        if (lane == 0) idx = atomicAdd(counter, 1);
        idx = __shfl_sync(0xffffffffu, idx, 0);
        if (idx >= count) break;
#endif
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
        const RowConst rc = make_row_const((c.flags & 2) ? 0 : K, (double)m.n, exp_tab, log_tab, KIND == RK_POIS ? lnmu_tab : nullptr);

        int mode = 0;
        if (lane == 0) {
            if (KIND != RK_PINIT) {
                const long long gy = gram_col(d.ycol);
                for (int j = 0; j < K; j++) {
                    xty[j] = m.G[(long long)gidx(d, j) * m.D + gy];
                    xty[MAXK + j] = m.G[(long long)gidx(d, j) * m.D];
                }
            }
            state_setup(m, d, ch, c, st);
            if (c.method == BAYES) bayes_inits(m, d, ch, st);
            mode = fit_begin(st, c);
            if (KIND == RK_BIN_MLE) { // iteration 1 from the Gram matrix, no data pass
                binomial_first_acc(m, d, *acc, !(c.flags & 2));
                st.have_ll = 1;
                mode = fit_consume(st, c, *acc);
                st.passes--;
            }
            if (KIND == RK_POIS && mode == PASS_PINIT && !(c.flags & 1)) { // Gram-only start; exact path: the reference's data pass
                poisson_init_acc(m, d, ch, *acc);
                mode = fit_consume(st, c, *acc);
                st.passes--;
            }
        }
        mode = __shfl_sync(0xffffffffu, mode, 0);
        __syncwarp();
        int careful = 0; // set by the first lean pass that meets an eta outside its range (run_pass)
        while (mode != PASS_DONE) {
            // passes by mode, per bucket: the numerator of the FP64 roofline (abn_timing.kernel_fp64)
            if (lane == 0) atomicAdd(reinterpret_cast<unsigned long long *>(counter + 2) + mode, 1ull);
            // only the modes a row kind's state machines can ask for are instantiated
            if (mode == PASS_HG) run_pass<KIND, K, ICPT, PASS_HG>(y, cp, st.x, m.n, rc, acc, tile, xty, lean_tab, careful, pipe);
            else if (mode == PASS_FULL) run_pass<KIND, K, ICPT, PASS_FULL>(y, cp, st.x, m.n, rc, acc, tile, xty, lean_tab, careful, pipe);
            else if (mode == PASS_LL) run_pass<KIND, K, ICPT, PASS_LL>(y, cp, st.x, m.n, rc, acc, tile, xty, lean_tab, careful, pipe);
            else if (KIND != RK_BIN_MLE && mode == PASS_GRAD) run_pass<KIND, K, ICPT, PASS_GRAD>(y, cp, st.x, m.n, rc, acc, tile, xty, lean_tab, careful, pipe);
            else if (KIND != RK_BIN_MLE && mode == PASS_HESS) run_pass<KIND, K, ICPT, PASS_HESS>(y, cp, st.x, m.n, rc, acc, tile, xty, lean_tab, careful, pipe);
            else if (KIND == RK_POIS && mode == PASS_PINIT) run_pass<RK_PINIT, K, ICPT, PASS_HG>(y, cp, st.x, m.n, rc, acc, tile, xty, lean_tab, careful, pipe);
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
            write_coef(out, row, d, st.x, st.M);
            write_se(out, row, d, st, c.method);
            if (out.sse && c.method == MLE) out.sse[row] = sse_of(m, d, st.x);
        }
        __syncwarp();
    }
}

// host-side launcher table entry
typedef cudaError_t (*FitLaunchFn)(const DataMeta &, const Control &, const FitDesc *, const int *, const int *, int,
                                   int *, ResultPtrs, int grid, cudaStream_t);

// the opt-in to more than 48 KB of dynamic shared memory is a per-device attribute of the kernel: it is
// set once per (instantiation, device) - a process may score on several devices (abn_control.devices)
inline bool first_launch_on_device(bool (&done)[64])
{
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return true;
    if (done[dev]) return false;
    done[dev] = true;
    return true;
}

template <int KIND, int K, bool ICPT>
cudaError_t launch_fit(const DataMeta &m, const Control &c, const FitDesc *descs, const int *child, const int *list,
                       int count, int *counter, ResultPtrs out, int grid, cudaStream_t stream)
{
    constexpr size_t smem = fit_stage_bytes<K, ICPT>();
    static bool configured[64] = {false};
    if (smem > 0 && first_launch_on_device(configured)) {
        cudaError_t e = cudaFuncSetAttribute(fit_kernel<KIND, K, ICPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    fit_kernel<KIND, K, ICPT><<<grid, FIT_WARPS * 32, smem, stream>>>(m, c, descs, child, list, count, counter, out);
    return cudaGetLastError();
}

// defined in fit_bin_mle.cu / fit_bin_bayes.cu / fit_pois.cu: [k][icpt]
FitLaunchFn get_launch_bin_mle(int k, int icpt);
FitLaunchFn get_launch_bin_bayes(int k, int icpt);
FitLaunchFn get_launch_pois(int k, int icpt);

} // namespace ABN_NS
