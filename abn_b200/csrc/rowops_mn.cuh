// rowops_mn.cuh - row op of the multinomial-child (softmax) objective, __host__ __device__.
#pragma once

#include "fastmath.cuh"
#include "solver.cuh"

namespace ABN_NS {

// value and gradient contribution of one row. x: design row (x[0] = 1 with ICPT), yl[l-1] = one-hot
// indicator of level l (l = 1..L-1; level 0 is the reference), W: (L-1) x K weights.
template <int L, int K, bool ICPT>
ABN_HD void mn_row(const double (&x)[K], const double (&yl)[L - 1], const double (&W)[(L - 1) * K],
                   const double *tab, const double *logtab, double &f, double (&g)[(L - 1) * K])
{
    double o[L];
    o[0] = 0.0;
#pragma unroll
    for (int l = 1; l < L; l++) {
        double s = ICPT ? W[(l - 1) * K] : 0.0;
#pragma unroll
        for (int j = (ICPT ? 1 : 0); j < K; j++) s = fma(W[(l - 1) * K + j], x[j], s);
        o[l] = s;
    }
    double mx = o[0];
#pragma unroll
    for (int l = 1; l < L; l++) mx = fmax(mx, o[l]);
    double p[L], sum = 0.0;
#pragma unroll
    for (int l = 0; l < L; l++) { p[l] = fast_exp(fmax(o[l] - mx, -700.0), tab); sum += p[l]; }
    const double inv = fast_rcp_approx(sum);
    double oy = 0.0; // o of the observed level
#pragma unroll
    for (int l = 1; l < L; l++) oy = fma(yl[l - 1], o[l], oy);
    f -= (oy - mx) - fast_log_ge1(sum, logtab); // - log p_y (sum >= 1: the largest term is exp(0))
#pragma unroll
    for (int l = 1; l < L; l++) {
        const double e = fma(p[l], inv, -yl[l - 1]);
#pragma unroll
        for (int j = 0; j < K; j++) g[(l - 1) * K + j] = (ICPT && j == 0) ? (g[(l - 1) * K + j] + e) : fma(e, x[j], g[(l - 1) * K + j]);
    }
}

// The lean form of the same row (designs of any width, with or without intercept). Nothing of the response is read:
//     value    = sum_i log(1 + sum_l exp(o_il))  -  sum_l W_l'(X'y_l)
//     gradient = sum_i p_il x_i                  -  X'y_l,          p_il = exp(o_il) / (1 + sum_m exp(o_im)),
// and X'y_l - the one-hot columns of the child against the design columns - are entries of the global Gram matrix
// (mn_lean_finish). No shift by the largest o: the caller checks after the pass that every |o| stayed below 700
// (o[] returns them) and falls back to mn_row otherwise. xd: the KD data columns of the row.
template <int L, int K, bool ICPT>
ABN_HD void mn_row_lean(const double *xd, const double (&W)[(L - 1) * K], const double *tab256, const double *logtab,
                        double &f, double (&g)[(L - 1) * K], double (&o)[L - 1])
{
    constexpr int KD = K - (ICPT ? 1 : 0), I0 = ICPT ? 1 : 0;
#pragma unroll
    for (int l = 0; l < L - 1; l++) {
        double s = ICPT ? W[l * K] : 0.0;
#pragma unroll
        for (int j = 0; j < KD; j++) s = fma(W[l * K + I0 + j], xd[j], s);
        o[l] = s;
    }
    double p[L - 1], sum = 1.0;
#pragma unroll
    for (int l = 0; l < L - 1; l++) { p[l] = fast_exp256s(o[l], tab256); sum += p[l]; }
    const double inv = fast_rcp_approx(sum);
    f += fast_log_ge1(sum, logtab);
#pragma unroll
    for (int l = 0; l < L - 1; l++) {
        const double e = p[l] * inv;
        if (ICPT) g[l * K] += e;
#pragma unroll
        for (int j = 0; j < KD; j++) g[l * K + I0 + j] = fma(e, xd[j], g[l * K + I0 + j]);
    }
}

// xty[l * K + j] = (design column j)'(indicator of level l + 1)
ABN_HD void mn_lean_finish(int np, double fsum, const double *gsum, const double *W, const double *xty, double *f_out, double *g_out)
{
    double f = fsum;
    for (int i = 0; i < np; i++) f = fma(-W[i], xty[i], f);
    *f_out = f;
    for (int i = 0; i < np; i++) g_out[i] = gsum[i] - xty[i];
}

} // namespace ABN_NS
