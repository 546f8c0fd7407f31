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

} // namespace ABN_NS
