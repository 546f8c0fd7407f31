// rowops.cuh - what one observation row contributes to a pass: (w, r, l) from (eta, y).
//
//   H += w x x^T,   g += r x,   ll += l
//
// One struct per (family, method); all FP64. Reference formulas (paths relative to the
// reference tree) are cited per op. __host__ __device__ so tests/host_sim.cpp reuses them.
#pragma once

#include "solver.cuh"

namespace abn {

enum RowKind : int {
    RK_BIN_MLE = 0,   // src/irls_binomial_fast_br.cpp:53-66,97
    RK_POIS = 1,      // src/irls_poisson_fast.cpp:62-74 and src/node_poisson.c:545-667
    RK_BIN_BAYES = 2, // src/node_binomial.c:626-640,690-700,759-770
    RK_PINIT = 3      // src/irls_poisson_fast.cpp:37-40
};

struct RowConst {
    double bp0, bp1; // binomial MLE bias-reduced pseudo-responses for y = 0 / y = 1
};

// b' = (y + 0.5 k/n) / (1 + k/n), irls_binomial_fast_br.cpp:58
ABN_HD RowConst make_row_const(int k, double n)
{
    RowConst c;
    double dfn = (double)k / n;
    c.bp0 = (0.0 + dfn * 0.5) / (1 + dfn);
    c.bp1 = (1.0 + dfn * 0.5) / (1 + dfn);
    return c;
}

template <int KIND>
struct RowOp;

template <>
struct RowOp<RK_BIN_MLE> {
    // e = exp(-eta); g = 1/(1+e); g' = e/(1+e)^2; W = g'^2/(g(1-g)); z = eta + (b' - g)/g'.
    // In exact arithmetic W == g', but the reference forms 1 - g in floating point, and for
    // saturated rows (|eta| > ~20) the rounding of g makes rho = W/g' differ from 1 by up to
    // 2^-53/(1-g). Because the bias-reduced pseudo-response keeps (b' - g) ~ k/2n on those rows,
    // that noise moves the fixed point of the reference's IRLS at the 1e-8 level; parity needs
    // the same rho. So: W = rho g', W z = W eta + rho (b' - g), i.e. the Newton form
    //   x_new = x + (X'WX)^-1 X'[rho (b' - g)]  ==  solve(X'WX, X'Wz)  (irls_binomial_fast_br.cpp:72-73)
    // loglik term: -y log(1+e^-eta) - (1-y) log(1+e^eta) = -log(1+e) - (1-y) eta   (:97)
    template <bool NEED_L>
    static ABN_HD void apply(double eta, double y, const RowConst &c, double &w, double &r, double &l, int &bad)
    {
        double e = exp(-eta);
        double d = 1.0 + e;
        double g = 1.0 / d;          // IEEE division: the same g as the reference's
        double gp = e * g * g;
        double rho = gp / (g * (1.0 - g));
        w = rho * gp;
        r = rho * (((y != 0.0) ? c.bp1 : c.bp0) - g);
        // the reference's W or z is non-finite when g rounds to 1 (g(1-g) = 0) or (1+e)^2 overflows
        bad |= (d == 1.0) | !(d < 1.3407807929942596e154);
        if (NEED_L) l = -log(d) - (1.0 - y) * eta;
    }
};

template <>
struct RowOp<RK_POIS> {
    template <bool NEED_L>
    static ABN_HD void apply(double eta, double y, const RowConst &, double &w, double &r, double &l, int &bad)
    {
        double E = exp(eta);
        w = E;
        r = y - E;
        bad |= !isfinite(E);
        if (NEED_L) l = y * eta - E;
    }
};

template <>
struct RowOp<RK_BIN_BAYES> {
    template <bool NEED_L>
    static ABN_HD void apply(double eta, double y, const RowConst &, double &w, double &r, double &l, int &bad)
    {
        double E = exp(eta);
        if (E != INFINITY) {
            double d = 1.0 + E;
            double inv = 1.0 / d;
            double sig = E * inv;
            w = sig * inv;          // exp(eta - 2 log(1+e^eta)), node_binomial.c:761-764
            r = y - sig;            // :697-700
            if (NEED_L) l = y * eta - log(d); // :633-639
        } else {
            w = 0.0;
            r = y - 1.0;
            if (NEED_L) l = y * eta - eta;
        }
        (void)bad;
    }
};

template <>
struct RowOp<RK_PINIT> {
    // mu = y + 0.1; eta0 = log(mu); W0 = exp(eta0); z0 = eta0 + (y - W0)/W0; H = X'W0X, g = X'W0z0
    template <bool NEED_L>
    static ABN_HD void apply(double, double y, const RowConst &, double &w, double &r, double &l, int &bad)
    {
        double mu = y + 0.1;
        double eta0 = log(mu);
        double W0 = exp(eta0);
        double z0 = eta0 + (y - W0) / W0;
        w = W0;
        r = W0 * z0;
        bad |= !isfinite(W0) | !isfinite(z0);
        if (NEED_L) l = 0.0;
    }
};

} // namespace abn
