// wide.cu - designs wider than the 8 columns the register-resident kernels (fit_kernel.cuh) take.
//
// Built from the SAME solver headers with MAXK = 32 in namespace abnw (make passes
// -DABN_MAXK=32 -DABN_NS=abnw), so the state machines, row ops, rank handling and Gram shortcuts
// are identical; only the data pass differs. One warp per fit, one CTA per warp:
//   phase A  lane = row of a 32-row tile: eta = x'beta, the family's row op -> w, r, l; x, w, r go
//            to a shared-memory tile
//   phase B  lane = owner of k(k+1)/2 / 32 entries of X'WX (and lanes < k of one entry of X'r):
//            each owner sweeps the 32 rows of the tile - a small SYRK, no cross-lane reduction
// Multinomial children with more than 8 design columns are not handled (status 2).
#include <cuda_runtime.h>

#include <algorithm>

#ifndef ABN_MAXK
#error "compile wide.cu with -DABN_MAXK=32 -DABN_NS=abnw"
#endif

#include "plan.cuh"
#include "rowops.cuh"
#include "solver.cuh"

namespace ABN_NS {

static_assert(MAXK == 32, "the pair ownership below assumes MAXK = 32");
constexpr int PPL = (MAXK * (MAXK + 1) / 2 + 31) / 32; // pairs per lane: 17
constexpr int TSTR = 33;                               // tile row stride (bank-conflict padding)

struct WideShared {
    FitState st;
    PassAcc acc;
    double xt[MAXK * TSTR]; // [column][row]
    double wt[32], rt[32];
    double exp_tab[ABN_EXP2_N];
    double log_tab[128];
    unsigned short pair_pq[MAXK * (MAXK + 1) / 2]; // (p << 8) | q for pair index
    FitDesc d;
};

__device__ __forceinline__ double warp_sum_w(double v)
{
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    return v;
}

// one data pass at S.st.x: fills S.acc (H, g, ll, bad) - always the full set of sums
__device__ void wide_pass(const DataMeta &m, const Control &c, WideShared &S, int mode)
{
    const int lane = threadIdx.x;
    const FitDesc &d = S.d;
    const int k = d.k;
    const int npairs = k * (k + 1) / 2;
    const bool pinit = (mode == PASS_PINIT);
    const int kind = pinit ? RK_PINIT : (d.family == POISSON ? RK_POIS : (c.method == MLE ? RK_BIN_MLE : RK_BIN_BAYES));
    const RowConst rc = make_row_const((c.flags & 2) ? 0 : k, (double)m.n, S.exp_tab, S.log_tab, nullptr);
    const double *y = m.X + (long long)d.ycol * m.ld;

    double hacc[PPL];
#pragma unroll
    for (int t = 0; t < PPL; t++) hacc[t] = 0.0;
    double gacc = 0.0, ll = 0.0;
    RowTrack tk = row_track_init();

    for (long long base = 0; base < m.n; base += 32) {
        const long long i = base + lane;
        const bool live = i < m.n;
        // phase A
        double eta = 0.0;
        for (int j = 0; j < k; j++) {
            const int col = d.cols[j];
            const double xv = !live ? 0.0 : (col < 0 ? 1.0 : __ldg(m.X + (long long)col * m.ld + i));
            S.xt[j * TSTR + lane] = xv;
            eta = fma(xv, S.st.x[j], eta);
        }
        double w = 0.0, r = 0.0, l = 0.0;
        if (live) {
            const double yi = __ldg(y + i);
            switch (kind) {
            case RK_BIN_MLE: RowOp<RK_BIN_MLE>::apply<true, true, true>(eta, yi, rc, w, r, l, tk); break;
            case RK_POIS: RowOp<RK_POIS>::apply<true, true, true>(eta, yi, rc, w, r, l, tk); break;
            case RK_BIN_BAYES: RowOp<RK_BIN_BAYES>::apply<true, true, true>(eta, yi, rc, w, r, l, tk); break;
            default: RowOp<RK_PINIT>::apply<true, true, true>(eta, yi, rc, w, r, l, tk); break;
            }
        }
        S.wt[lane] = w;
        S.rt[lane] = r;
        ll += l;
        __syncwarp();
        // phase B
#pragma unroll
        for (int t = 0; t < PPL; t++) {
            const int pi = lane + 32 * t;
            if (pi < npairs) {
                const int pq = S.pair_pq[pi], p = pq >> 8, q = pq & 255;
                double s = hacc[t];
                for (int row = 0; row < 32; row++) s = fma(S.wt[row] * S.xt[p * TSTR + row], S.xt[q * TSTR + row], s);
                hacc[t] = s;
            }
        }
        if (lane < k) {
            double s = gacc;
            for (int row = 0; row < 32; row++) s = fma(S.rt[row], S.xt[lane * TSTR + row], s);
            gacc = s;
        }
        __syncwarp();
    }
    ll = warp_sum_w(ll);
    int bad_lane;
    switch (kind) {
    case RK_BIN_MLE: bad_lane = RowOp<RK_BIN_MLE>::bad_of(tk); break;
    case RK_POIS: bad_lane = RowOp<RK_POIS>::bad_of(tk); break;
    case RK_BIN_BAYES: bad_lane = RowOp<RK_BIN_BAYES>::bad_of(tk); break;
    default: bad_lane = RowOp<RK_PINIT>::bad_of(tk); break;
    }
    const int bad = __any_sync(0xffffffffu, bad_lane);
#pragma unroll
    for (int t = 0; t < PPL; t++) {
        const int pi = lane + 32 * t;
        if (pi < npairs) {
            const int pq = S.pair_pq[pi];
            S.acc.H[tri(pq >> 8, pq & 255, k)] = hacc[t];
        }
    }
    if (lane < k) S.acc.g[lane] = gacc;
    if (lane == 0) { S.acc.ll = ll; S.acc.bad = bad; S.acc.has_gl = 0; }
    __syncwarp();
}

__global__ void __launch_bounds__(32)
fit_wide_kernel(const DataMeta m, const Control c, const int *__restrict__ child, const int *__restrict__ parents,
                int width, const int *__restrict__ list, int count, int *counter, ResultPtrs out)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    WideShared &S = *reinterpret_cast<WideShared *>(smem_raw);
    const int lane = threadIdx.x;
    for (int t = lane; t < ABN_EXP2_N; t += 32) S.exp_tab[t] = c_exp2_table[t];
    for (int t = lane; t < 64; t += 32) fill_log_table(S.log_tab, t);
    __syncwarp();
    for (;;) {
        int idx = 0;
        if (lane == 0) idx = atomicAdd(counter, 1);
        idx = __shfl_sync(0xffffffffu, idx, 0);
        if (idx >= count) break;
        const int row = list[idx];
        const int ch = child[row];
        int mode = PASS_DONE, skip = 0;
        if (lane == 0) {
            resolve_design(m, c.method, ch, parents + (long long)row * width, width, S.d);
            const FitDesc &d = S.d;
            if (d.err || d.k == 0 || d.family == MULTINOMIAL) skip = 1;
            else {
                int pi = 0;
                for (int p = 0; p < d.k; p++)
                    for (int q = p; q < d.k; q++) S.pair_pq[pi++] = (unsigned short)((p << 8) | q);
            }
        }
        skip = __shfl_sync(0xffffffffu, skip, 0);
        __syncwarp();
        if (skip) {
            if (lane == 0) {
                if (out.mlik) out.mlik[row] = NAN;
                if (out.aic) out.aic[row] = NAN;
                if (out.bic) out.bic[row] = NAN;
                if (out.mdl) out.mdl[row] = NAN;
                if (out.status) out.status[row] = FIT_NON_FINITE;
                if (out.iters) out.iters[row] = 0;
            }
            continue;
        }
        const FitDesc &d = S.d;
        if (d.family == GAUSSIAN) { // Gram only, lane 0
            if (lane == 0) {
                if (c.method == MLE) {
                    double ll, aic, bic;
                    int status;
                    double sse;
                    gauss_mle_fit(m, d, &ll, &aic, &bic, &status, S.st.x, out.se ? S.st.dir : nullptr, &sse);
                    write_coef(out, row, d, S.st.x, d.k);
                    if (out.se) { ResultPtrs o = out; o.coef = out.se; write_coef(o, row, d, S.st.dir, d.k); }
                    if (out.sse) out.sse[row] = sse;
                    if (out.mlik) out.mlik[row] = ll;
                    if (out.aic) out.aic[row] = aic;
                    if (out.bic) out.bic[row] = bic;
                    if (out.mdl) out.mdl[row] = mdl_of(bic, d, c.nvars);
                    if (out.status) out.status[row] = status;
                } else {
                    state_setup(m, d, ch, c, S.st);
                    bayes_inits(m, d, ch, S.st);
                    gauss_gram_acc(m, d, S.acc);
                    int md = fit_begin(S.st, c);
                    while (md != PASS_DONE) md = fit_consume(S.st, c, S.acc);
                    if (out.mlik) out.mlik[row] = S.st.loglik;
                    if (out.status) out.status[row] = S.st.status;
                    write_coef(out, row, d, S.st.x, S.st.M);
                }
                if (out.iters) out.iters[row] = 1;
            }
            __syncwarp();
            continue;
        }
        if (lane == 0) {
            state_setup(m, d, ch, c, S.st);
            if (c.method == BAYES) bayes_inits(m, d, ch, S.st);
            mode = fit_begin(S.st, c);
        }
        mode = __shfl_sync(0xffffffffu, mode, 0);
        __syncwarp();
        while (mode != PASS_DONE) {
            wide_pass(m, c, S, mode);
            if (lane == 0) mode = fit_consume(S.st, c, S.acc);
            mode = __shfl_sync(0xffffffffu, mode, 0);
            __syncwarp();
        }
        if (lane == 0) {
            if (out.mlik) out.mlik[row] = S.st.loglik;
            if (c.method == MLE) {
                if (out.aic) out.aic[row] = S.st.aic;
                if (out.bic) out.bic[row] = S.st.bic;
                if (out.mdl) out.mdl[row] = mdl_of(S.st.bic, d, c.nvars);
            }
            if (out.status) out.status[row] = S.st.status;
            if (out.iters) out.iters[row] = S.st.passes;
            write_coef(out, row, d, S.st.x, S.st.M);
            write_se(out, row, d, S.st, c.method);
            if (out.sse && c.method == MLE) out.sse[row] = sse_of(m, d, S.st.x);
        }
        __syncwarp();
    }
}

} // namespace ABN_NS

// host entry used by abn_b200.cu for the rows its resolve step flagged as too wide
cudaError_t abn_wide_launch(const abnc::DataMeta &m, const abnc::Control &c, const int *child, const int *parents,
                            int width, const int *list, int count, int *counter, abnc::ResultPtrs out, int sm_count,
                            cudaStream_t stream)
{
    using namespace ABN_NS;
    const size_t smem = sizeof(WideShared);
    static bool configured[64] = {false}; // the attribute is per device
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || !configured[dev]) {
        cudaError_t e = cudaFuncSetAttribute(fit_wide_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        if (dev >= 0 && dev < 64) configured[dev] = true;
    }
    const int grid = std::min(count, sm_count * 4);
    fit_wide_kernel<<<grid, 32, smem, stream>>>(m, c, child, parents, width, list, count, counter, out);
    return cudaGetLastError();
}

int abn_wide_max_k() { return ABN_NS::MAXK; }
