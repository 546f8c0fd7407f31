// solver.cuh - per-fit solver state machines of the B200 score-cache path.
//
// A fit is driven as:   mode = fit_begin(state);  while (mode) { acc = data_pass(state.x, mode);
//                                                                 mode = fit_consume(state, acc); }
// where data_pass streams the observation rows once and returns the raw reductions
//     H = sum_i w_i x_i x_i^T,  g = sum_i r_i x_i,  ll = sum_i l_i        (row ops in rowops.cuh)
// and everything else (tiny k x k algebra, step control, stop rules, scores) happens here, in
// registers / shared memory of the warp that owns the fit.
//
// What each machine follows in the reference (paths relative to the reference tree):
//   MLE binomial  src/irls_binomial_fast_br.cpp:50-101 (bias-reduced IRLS, loglik at the pre-update eta)
//   MLE Poisson   src/irls_poisson_fast.cpp:33-139     (ridge 1e-8, step halving, >= acceptance)
//   MLE Gaussian  src/irls_gaussian_fast.cpp:34-49     (closed form from Gram blocks)
//   Bayes         src/node_binomial.c:88-193, src/node_poisson.c:96-194, src/node_gaussian.c:93-185
//                 Laplace score at the root of grad g found by Powell's hybrid method exactly as
//                 GSL's hybridsj / hybridj (MINPACK HYBRJ) sequence the reference calls, or by
//                 plain Newton (ABN_SOLVER_NEWTON).
//
// The header is __host__ __device__ clean so tests can drive the very same state machines on the
// CPU (tests/host_sim.cpp) - the product path only ever runs them inside the CUDA kernels.
#pragma once

#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define ABN_HD __host__ __device__ __forceinline__
#define ABN_HDN inline __host__ __device__ __noinline__
#else
#define ABN_HD inline
#define ABN_HDN inline
#endif

#include "types.cuh"

// The same headers build twice: namespace abn with MAXK = 8 (register-resident accumulators, the
// fast path) and namespace abnw with MAXK = 32 (wide.cu: generic kernel for wider designs).
#ifndef ABN_MAXK
#define ABN_MAXK 8
#endif
#ifndef ABN_NS
#define ABN_NS abn
#endif

namespace ABN_NS {

using abnc::Control;
using abnc::DataMeta;
using abnc::ResultPtrs;

constexpr int MAXK = ABN_MAXK; // widest design (intercept + data columns)
constexpr int MAXM = MAXK + 1; // unknowns of the root problem (Gaussian Bayes adds the precision)
constexpr int NTRI = MAXK * (MAXK + 1) / 2;

enum Family : int { BINOMIAL = 1, GAUSSIAN = 2, POISSON = 3, MULTINOMIAL = 4 };
enum Method : int { MLE = 0, BAYES = 1 };
enum SolverKind : int { NEWTON = 0, HYBRID = 1 };
enum FitStatus : int { FIT_OK = 0, FIT_NOT_CONVERGED = 1, FIT_NON_FINITE = 2 };

// what the next data pass has to produce
enum PassMode : int {
    PASS_DONE = 0,
    PASS_FULL = 1,   // H, g, ll at state.x
    PASS_GRAD = 2,   // g only (hybrid trial points)
    PASS_HESS = 3,   // H only (hybrid Jacobian refresh)
    PASS_PINIT = 4,  // Poisson MLE start values: H = X'W0X, g = X'W0z0 (irls_poisson_fast.cpp:37-52)
    PASS_HG = 5,     // H and g, no log-likelihood term (saves the log per row)
    PASS_LL = 6      // log-likelihood term only
};

// raw reductions of one pass; H is the packed upper triangle, row-major: (a,b), a <= b
struct PassAcc {
    double H[NTRI];
    double g[MAXK];
    double ll;
    int bad; // a non-finite weight / mean was met in this pass
    int has_gl;      // gl is filled (binomial mle passes that keep X'g apart from the adjusted score)
    double gl[MAXK]; // X'(y - g): gradient of the reported, unpenalised binomial loglik
};

ABN_HD int tri(int a, int b, int k) { return a * k - a * (a - 1) / 2 + (b - a); } // a <= b

// ------------------------------------------------------------------------------------------------
// small dense algebra, row-major m x m with leading dimension MAXM
// ------------------------------------------------------------------------------------------------
ABN_HD double &A_(double *A, int i, int j) { return A[i * MAXM + j]; }
ABN_HD double A_(const double *A, int i, int j) { return A[i * MAXM + j]; }

// Cholesky in place (lower). false if a pivot is not positive / finite.
ABN_HDN bool chol(double *A, int m)
{
    for (int j = 0; j < m; j++) {
        double s = A_(A, j, j);
        for (int k = 0; k < j; k++) s -= A_(A, j, k) * A_(A, j, k);
        if (!(s > 0.0) || !isfinite(s)) return false;
        double l = sqrt(s);
        A_(A, j, j) = l;
        for (int i = j + 1; i < m; i++) {
            double t = A_(A, i, j);
            for (int k = 0; k < j; k++) t -= A_(A, i, k) * A_(A, j, k);
            A_(A, i, j) = t / l;
        }
    }
    return true;
}

ABN_HDN void chol_solve(const double *L, int m, const double *b, double *x)
{
    for (int i = 0; i < m; i++) {
        double s = b[i];
        for (int j = 0; j < i; j++) s -= A_(L, i, j) * x[j];
        x[i] = s / A_(L, i, i);
    }
    for (int i = m - 1; i >= 0; i--) {
        double s = x[i];
        for (int j = i + 1; j < m; j++) s -= A_(L, j, i) * x[j];
        x[i] = s / A_(L, i, i);
    }
}

// LU with partial pivoting (gsl_linalg_LU_decomp), in place
ABN_HDN void lu_decomp(double *A, int m, int *perm)
{
    for (int i = 0; i < m; i++) perm[i] = i;
    for (int j = 0; j + 1 < m; j++) {
        double mx = fabs(A_(A, j, j));
        int ip = j;
        for (int i = j + 1; i < m; i++) {
            double a = fabs(A_(A, i, j));
            if (a > mx) { mx = a; ip = i; }
        }
        if (ip != j) {
            for (int c = 0; c < m; c++) { double t = A_(A, j, c); A_(A, j, c) = A_(A, ip, c); A_(A, ip, c) = t; }
            int t = perm[j]; perm[j] = perm[ip]; perm[ip] = t;
        }
        double ajj = A_(A, j, j);
        if (ajj != 0.0) {
            for (int i = j + 1; i < m; i++) {
                double f = A_(A, i, j) / ajj;
                A_(A, i, j) = f;
                for (int c = j + 1; c < m; c++) A_(A, i, c) -= f * A_(A, j, c);
            }
        }
    }
}

ABN_HD bool lu_singular(const double *LU, int m)
{
    for (int i = 0; i < m; i++) if (A_(LU, i, i) == 0.0) return true;
    return false;
}

ABN_HD double lu_lndet(const double *LU, int m)
{
    double s = 0.0;
    for (int i = 0; i < m; i++) s += log(fabs(A_(LU, i, i)));
    return s;
}

ABN_HDN void lu_solve(const double *LU, const int *perm, int m, const double *b, double *x)
{
    for (int i = 0; i < m; i++) {
        double s = b[perm[i]];
        for (int j = 0; j < i; j++) s -= A_(LU, i, j) * x[j];
        x[i] = s;
    }
    for (int i = m - 1; i >= 0; i--) {
        double s = x[i];
        for (int j = i + 1; j < m; j++) s -= A_(LU, i, j) * x[j];
        x[i] = s / A_(LU, i, i);
    }
}

// cyclic Jacobi eigen-decomposition of a symmetric matrix (destroys A); V columns = vectors
ABN_HDN void jacobi_eig(double *A, int m, double *w, double *V)
{
    for (int i = 0; i < m; i++)
        for (int j = 0; j < m; j++) A_(V, i, j) = (i == j) ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 40; sweep++) {
        double off = 0.0;
        for (int i = 0; i < m; i++)
            for (int j = i + 1; j < m; j++) off += A_(A, i, j) * A_(A, i, j);
        if (off == 0.0) break;
        for (int p = 0; p < m; p++) {
            for (int q = p + 1; q < m; q++) {
                double apq = A_(A, p, q);
                if (apq == 0.0) continue;
                double theta = (A_(A, q, q) - A_(A, p, p)) / (2.0 * apq);
                double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
                for (int k = 0; k < m; k++) {
                    double akp = A_(A, k, p), akq = A_(A, k, q);
                    A_(A, k, p) = c * akp - s * akq;
                    A_(A, k, q) = s * akp + c * akq;
                }
                for (int k = 0; k < m; k++) {
                    double apk = A_(A, p, k), aqk = A_(A, q, k);
                    A_(A, p, k) = c * apk - s * aqk;
                    A_(A, q, k) = s * apk + c * aqk;
                }
                for (int k = 0; k < m; k++) {
                    double vkp = A_(V, k, p), vkq = A_(V, k, q);
                    A_(V, k, p) = c * vkp - s * vkq;
                    A_(V, k, q) = s * vkp + c * vkq;
                }
            }
        }
    }
    for (int i = 0; i < m; i++) w[i] = A_(A, i, i);
}

// arma::solve(A, b, likely_sympd [+ no_approx]) on a symmetric k x k system.
// Cholesky, then LU; if singular to working precision: minimum-norm least squares through the
// eigen-decomposition when allow_approx, otherwise failure. Returns 0 ok, 1 approximate, -1 fail.
// S (scratch) and A are MAXM-strided; A is preserved.
ABN_HDN int sym_solve(const double *A, const double *b, int m, bool allow_approx, double *x, double *S)
{
    bool finite = true;
    for (int i = 0; i < m; i++) {
        if (!isfinite(b[i])) finite = false;
        for (int j = 0; j < m; j++) if (!isfinite(A_(A, i, j))) finite = false;
    }
    if (finite) {
        for (int i = 0; i < m; i++) for (int j = 0; j < m; j++) A_(S, i, j) = A_(A, i, j);
        if (chol(S, m)) { chol_solve(S, m, b, x); return 0; }
        int perm[MAXM];
        for (int i = 0; i < m; i++) for (int j = 0; j < m; j++) A_(S, i, j) = A_(A, i, j);
        lu_decomp(S, m, perm);
        double umax = 0.0, umin = 1.79769313486231570815e308;
        for (int i = 0; i < m; i++) {
            double u = fabs(A_(S, i, i));
            if (u > umax) umax = u;
            if (u < umin) umin = u;
        }
        bool badlu = (umin == 0.0) || (allow_approx && umin < umax * 2.220446049250313e-16);
        if (!badlu) { lu_solve(S, perm, m, b, x); return 0; }
        if (allow_approx) {
            double w[MAXM], V[MAXM * MAXM];
            for (int i = 0; i < m; i++) for (int j = 0; j < m; j++) A_(S, i, j) = 0.5 * (A_(A, i, j) + A_(A, j, i));
            jacobi_eig(S, m, w, V);
            double wmax = 0.0;
            for (int i = 0; i < m; i++) if (fabs(w[i]) > wmax) wmax = fabs(w[i]);
            double tol = m * wmax * 2.220446049250313e-16;
            for (int i = 0; i < m; i++) x[i] = 0.0;
            for (int e = 0; e < m; e++) {
                if (fabs(w[e]) <= tol) continue;
                double s = 0.0;
                for (int i = 0; i < m; i++) s += A_(V, i, e) * b[i];
                s /= w[e];
                for (int i = 0; i < m; i++) x[i] += s * A_(V, i, e);
            }
            return 1;
        }
    }
    for (int i = 0; i < m; i++) x[i] = NAN;
    return -1;
}

// Householder QR: A = Q R (construction of gsl_linalg_QR_decomp + QR_unpack)
ABN_HDN void qr_decomp(const double *A, int m, double *Q, double *R)
{
    for (int i = 0; i < m; i++)
        for (int j = 0; j < m; j++) { A_(R, i, j) = A_(A, i, j); A_(Q, i, j) = (i == j) ? 1.0 : 0.0; }
    double v[MAXM];
    for (int j = 0; j < m; j++) {
        double xnorm = 0.0;
        for (int i = j + 1; i < m; i++) xnorm += A_(R, i, j) * A_(R, i, j);
        xnorm = sqrt(xnorm);
        if (xnorm == 0.0) continue;
        double alpha = A_(R, j, j);
        double beta = -(alpha >= 0.0 ? +1.0 : -1.0) * hypot(alpha, xnorm);
        double tau = (beta - alpha) / beta;
        double s = alpha - beta;
        v[j] = 1.0;
        for (int i = j + 1; i < m; i++) v[i] = A_(R, i, j) / s;
        for (int c = j; c < m; c++) {
            double w = 0.0;
            for (int i = j; i < m; i++) w += v[i] * A_(R, i, c);
            w *= tau;
            for (int i = j; i < m; i++) A_(R, i, c) -= w * v[i];
        }
        for (int r = 0; r < m; r++) {
            double w = 0.0;
            for (int i = j; i < m; i++) w += A_(Q, r, i) * v[i];
            w *= tau;
            for (int i = j; i < m; i++) A_(Q, r, i) -= w * v[i];
        }
        for (int i = j + 1; i < m; i++) A_(R, i, j) = 0.0;
    }
}

ABN_HD void givens(double a, double b, double *c, double *s)
{
    if (b == 0) { *c = 1; *s = 0; }
    else if (fabs(b) > fabs(a)) { double t = -a / b; double s1 = 1.0 / sqrt(1 + t * t); *s = s1; *c = s1 * t; }
    else { double t = -b / a; double c1 = 1.0 / sqrt(1 + t * t); *c = c1; *s = c1 * t; }
}

ABN_HD void apply_givens_qr(int m, double *Q, double *R, int i, int j, double c, double s)
{
    for (int k = 0; k < m; k++) {
        double qki = A_(Q, k, i), qkj = A_(Q, k, j);
        A_(Q, k, i) = qki * c - qkj * s;
        A_(Q, k, j) = qki * s + qkj * c;
    }
    for (int k = (i < j ? i : j); k < m; k++) {
        double rik = A_(R, i, k), rjk = A_(R, j, k);
        A_(R, i, k) = c * rik - s * rjk;
        A_(R, j, k) = s * rik + c * rjk;
    }
}

// gsl_linalg_QR_update: Q' R' = Q (R + w v^T)
ABN_HDN void qr_update(double *Q, double *R, int m, double *w, const double *v)
{
    for (int k = m - 1; k > 0; k--) {
        double c, s;
        givens(w[k - 1], w[k], &c, &s);
        double wi = w[k - 1], wj = w[k];
        w[k - 1] = c * wi - s * wj;
        w[k] = s * wi + c * wj;
        apply_givens_qr(m, Q, R, k - 1, k, c, s);
    }
    double w0 = w[0];
    for (int j = 0; j < m; j++) A_(R, 0, j) += w0 * v[j];
    for (int k = 1; k < m; k++) {
        double c, s;
        givens(A_(R, k - 1, k - 1), A_(R, k, k - 1), &c, &s);
        apply_givens_qr(m, Q, R, k - 1, k, c, s);
        A_(R, k, k - 1) = 0.0;
    }
}

ABN_HD double enorm(const double *f, int n) { double s = 0; for (int i = 0; i < n; i++) s += f[i] * f[i]; return sqrt(s); }
ABN_HD double scaled_enorm(const double *d, const double *f, int n)
{
    double s = 0;
    for (int i = 0; i < n; i++) { double u = d[i] * f[i]; s += u * u; }
    return sqrt(s);
}

// ------------------------------------------------------------------------------------------------
// fit state
// ------------------------------------------------------------------------------------------------
enum Phase : int {
    PH_MLE_BIN = 1, PH_MLE_BIN_LL, PH_FINAL_LL,
    PH_POIS_INIT, PH_POIS_FIRST, PH_POIS_TRIAL,
    PH_NEWTON,
    PH_HYB_SET, PH_HYB_TRIAL, PH_HYB_JAC, PH_FINAL
};

// Centred moments of one Gaussian child's design [1 | x_1 .. x_k1] (plan.cuh gauss_centered_block): with them
// ||y - X b||^2 = [Cyy - 2 b_s'Cxy + b_s'Cxx b_s] + n (ybar - b_0 - b_s'xbar)^2, free of the cancellation the raw
// form y'y - 2 b'X'y + b'X'X b suffers when |mean(y)| >> sd(y) (the reference sums the residuals themselves,
// src/irls_gaussian_fast.cpp:38-39).
struct GaussAux {
    int k1;
    double Cxx[MAXM * MAXM], Cxy[MAXM], xbar[MAXM], Cyy, ybar;
};

ABN_HD double gauss_aux_rss(const GaussAux &g, const double *x, double n)
{
    double q = g.Cyy, c = g.ybar - x[0];
    for (int i = 0; i < g.k1; i++) {
        double t = 0.0;
        for (int j = 0; j < g.k1; j++) t += A_(g.Cxx, i, j) * x[1 + j];
        q += x[1 + i] * (t - 2.0 * g.Cxy[i]);
        c -= x[1 + i] * g.xbar[i];
    }
    return q + n * c * c;
}

struct FitState {
    // problem
    int family, k, M; // k: design columns (H is k x k); M: unknowns
    double n;         // observations
    double lfact;     // Poisson: sum lgamma(y+1) of the child
    // Gram-based start values (filled by the caller before fit_begin): least-squares coefficients
    double x0[MAXM];
    // evaluation point of the next pass
    double x[MAXM];
    // bookkeeping
    int phase, iters, passes, status;
    int have_ll;      // the pass in flight also produces the log-likelihood term
    // results
    double loglik, aic, bic;
    // MLE Poisson
    double xold[MAXM], dir[MAXM], cur_ll, step;
    double q_cur, q_prev, gd, dHd; // IRLS shortcuts: eta-scale length of the step in flight and of the one before; g'd, d'Hd
    int icpt;               // design column 0 is the constant 1 (H_00 = sum of the weights)
    int shortcut;           // IRLS shortcuts enabled for this fit (off with ABN_FLAG_EXACT_PATH)
    int halvings;
    // Bayes: last accepted point, its residual f = grad g, Jacobian pieces
    double xc[MAXM], f[MAXM], dx[MAXM], diag[MAXM];
    double J[MAXM * MAXM], Q[MAXM * MAXM], R[MAXM * MAXM];
    double fnorm, delta;
    int hy_iter, ncfail, ncsuc, nslow1, nslow2, scale, attempt, loop_iter;
    const GaussAux *gaux; // Gaussian child with centred moments available (else null: raw Gram entries)
};

// ---- Bayes: raw sums -> grad g / hess g / g --------------------------------------------------
// binomial & Poisson: node_binomial.c:655-817, node_poisson.c:576-700
ABN_HD void bayes_f_from_acc(const FitState &s, const Control &c, const double *x, const PassAcc &a, double *f)
{
    const double sd2 = c.prior_sd * c.prior_sd;
    for (int i = 0; i < s.k; i++) f[i] = (-1.0 / s.n) * (-(x[i] - c.prior_mean) / sd2 + a.g[i]);
}

ABN_HD void bayes_J_from_acc(const FitState &s, const Control &c, const PassAcc &a, double *J)
{
    const double sd2 = c.prior_sd * c.prior_sd;
    for (int i = 0; i < s.k; i++)
        for (int j = 0; j < s.k; j++) {
            double h = a.H[i <= j ? tri(i, j, s.k) : tri(j, i, s.k)] / s.n;
            if (i == j) h += 1.0 / (s.n * sd2);
            A_(J, i, j) = h;
        }
}

ABN_HD double bayes_g_from_acc(const FitState &s, const Control &c, const double *x, const PassAcc &a)
{
    double t2 = 0.0, t3 = 0.0;
    for (int j = 0; j < s.k; j++) {
        t2 += -log(sqrt(2.0 * M_PI) * c.prior_sd);
        double d = x[j] - c.prior_mean;
        t3 += -0.5 * (d * d) / (c.prior_sd * c.prior_sd);
    }
    double lik = a.ll;
    if (s.family == POISSON) lik -= s.lfact;
    return (-1.0 / s.n) * (lik + t2 + t3);
}

// Gaussian (node_gaussian.c:503-738) from Gram blocks: a.H = X'X, a.g = X'y, a.ll = y'y (constant
// per fit, supplied by the Gram evaluator), x = (beta, tau)
ABN_HD void gaus_eval(const FitState &s, const Control &c, const double *x, const PassAcc &a,
                      double *f, double *J, double *gval)
{
    const int m = s.k;
    const double tau = x[m], n = s.n, sd2 = c.prior_sd * c.prior_sd;
    double xtxb[MAXK];
    double yxb = 0.0, xbxb = 0.0;
    for (int i = 0; i < m; i++) {
        double t = 0.0;
        for (int j = 0; j < m; j++) t += a.H[i <= j ? tri(i, j, m) : tri(j, i, m)] * x[j];
        xtxb[i] = t;
        yxb += a.g[i] * x[i];
        xbxb += t * x[i];
    }
    const double rss = s.gaux ? gauss_aux_rss(*s.gaux, x, n) : -2.0 * yxb + xbxb + a.ll;
    if (f) {
        for (int i = 0; i < m; i++)
            f[i] = (-1.0 / n) * (-(x[i] - c.prior_mean) / sd2 + tau * (a.g[i] - xtxb[i]));
        f[m] = (-1.0 / n) * ((n / (2.0 * tau)) + rss * (-0.5) + (c.gam_shape - 1.0) / tau - (1.0 / c.gam_scale));
    }
    if (J) {
        for (int i = 0; i < m; i++)
            for (int j = 0; j < m; j++) {
                double xx = a.H[i <= j ? tri(i, j, m) : tri(j, i, m)];
                A_(J, i, j) = (i != j) ? (-1.0 / n) * (-tau * xx) : (-1.0 / n) * (-tau * xx - 1.0 / sd2);
            }
        A_(J, m, m) = (-1.0 / n) * ((-n / (2.0 * tau * tau)) - ((c.gam_shape - 1) / (tau * tau)));
        for (int j = 0; j < m; j++) A_(J, m, j) = A_(J, j, m) = (-1.0 / n) * (a.g[j] - xtxb[j]);
    }
    if (gval) {
        double t2 = 0.0, t3 = 0.0;
        for (int j = 0; j < m; j++) {
            t2 += -log(sqrt(2.0 * M_PI) * c.prior_sd);
            double d = x[j] - c.prior_mean;
            t3 += -0.5 * (d * d) / sd2;
        }
        double term4 = -(tau / 2.0) * rss;
        double term5 = (n / 2.0) * log(tau / (2.0 * M_PI));
        double term6 = -c.gam_shape * log(c.gam_scale) - lgamma(c.gam_shape) + (c.gam_shape - 1) * log(tau) - (tau / c.gam_scale);
        *gval = (-1.0 / n) * (t2 + t3 + term4 + term5 + term6);
    }
}

// unified view used by the Bayes machines
ABN_HD void bayes_eval(const FitState &s, const Control &c, const double *x, const PassAcc &a,
                       double *f, double *J, double *gval)
{
    if (s.family == GAUSSIAN) { gaus_eval(s, c, x, a, f, J, gval); return; }
    if (f) bayes_f_from_acc(s, c, x, a, f);
    if (J) bayes_J_from_acc(s, c, a, J);
    if (gval) *gval = bayes_g_from_acc(s, c, x, a);
}

// node_binomial.c:172-193: Laplace score from g(x) and hess g(x) (in s.J, destroyed)
ABN_HDN void bayes_score(FitState &s, double gval)
{
    int perm[MAXM];
    lu_decomp(s.J, s.M, perm);
    double mydet = lu_lndet(s.J, s.M);
    double dm = (double)s.M;
    double score = -s.n * gval - 0.5 * mydet + (dm / 2.0) * log((2.0 * M_PI) / s.n);
    if (isnan(score)) { score = NAN; s.status = FIT_NON_FINITE; }
    s.loglik = score;
    s.aic = s.bic = NAN;
}

// ... from a pass that produced H, g and ll at s.x
ABN_HDN void bayes_finish(FitState &s, const Control &c, const PassAcc &a)
{
    double gval;
    bayes_eval(s, c, s.x, a, nullptr, s.J, &gval);
    bayes_score(s, gval);
}

// ---- Powell hybrid (GSL hybridj.c / dogleg.c) ---------------------------------------------------
ABN_HDN void dogleg(const FitState &s, const double *qtf, double *p)
{
    const int n = s.M;
    double newton[MAXM], grad[MAXM];
    for (int i = n - 1; i >= 0; i--) {
        double t = qtf[i];
        for (int j = i + 1; j < n; j++) t -= A_(s.R, i, j) * newton[j];
        newton[i] = t / A_(s.R, i, i);
    }
    for (int i = 0; i < n; i++) newton[i] = -newton[i];
    double qnorm = scaled_enorm(s.diag, newton, n);
    if (qnorm <= s.delta) { for (int i = 0; i < n; i++) p[i] = newton[i]; return; }
    for (int j = 0; j < n; j++) {
        double sum = 0;
        for (int i = 0; i < n; i++) sum += A_(s.R, i, j) * qtf[i];
        grad[j] = -sum / s.diag[j];
    }
    double gnorm = enorm(grad, n);
    if (gnorm == 0) {
        double alpha = s.delta / qnorm;
        for (int i = 0; i < n; i++) p[i] = alpha * newton[i];
        return;
    }
    for (int i = 0; i < n; i++) grad[i] = (grad[i] / gnorm) / s.diag[i];
    for (int i = 0; i < n; i++) {
        double sum = 0;
        for (int j = i; j < n; j++) sum += A_(s.R, i, j) * grad[j];
        p[i] = sum;
    }
    double temp = enorm(p, n);
    double sgnorm = (gnorm / temp) / temp;
    if (sgnorm > s.delta) { for (int i = 0; i < n; i++) p[i] = s.delta * grad[i]; return; }
    double bnorm = enorm(qtf, n);
    double bg = bnorm / gnorm, bq = bnorm / qnorm, dq = s.delta / qnorm, dq2 = dq * dq;
    double sd = sgnorm / s.delta, sd2 = sd * sd;
    double t1 = bg * bq * sd, u = t1 - dq;
    double t2 = t1 - dq * sd2 + sqrt(u * u + (1 - dq2) * (1 - sd2));
    double alpha = dq * (1 - sd2) / t2, beta = (1 - alpha) * sgnorm;
    for (int i = 0; i < n; i++) p[i] = alpha * newton[i] + beta * grad[i];
}

ABN_HD void hy_diag(const double *J, int n, double *diag, bool update)
{
    for (int j = 0; j < n; j++) {
        double sum = 0;
        for (int i = 0; i < n; i++) sum += A_(J, i, j) * A_(J, i, j);
        if (sum == 0) sum = 1.0;
        double c = sqrt(sum);
        if (!update || c > diag[j]) diag[j] = c;
    }
}

ABN_HD double hy_delta(const double *diag, const double *x, int n)
{
    double Dx = scaled_enorm(diag, x, n);
    return (Dx > 0) ? 100.0 * Dx : 100.0;
}

// which scaling the current attempt uses: binomial / Gaussian try hybridsj then hybridj
// (node_binomial.c:90,127); Poisson tries hybridj twice then hybridsj twice (node_poisson.c:96-135)
ABN_HD bool hy_attempt_scale(const FitState &s, int attempt, bool *valid)
{
    if (s.family == POISSON) { *valid = attempt < 4; return attempt >= 2; }
    *valid = attempt < 2;
    return attempt == 0;
}

// start attempt s.attempt from the start values: next pass = FULL at x0
ABN_HD int hy_start_attempt(FitState &s)
{
    bool valid;
    s.scale = hy_attempt_scale(s, s.attempt, &valid) ? 1 : 0;
    for (int i = 0; i < s.M; i++) { s.xc[i] = s.x0[i]; s.x[i] = s.x0[i]; }
    s.phase = PH_HYB_SET;
    return PASS_HG;
}

// the trial point of the next hybrid iteration: dogleg step from (xc, f, Q, R)
ABN_HDN int hy_next_trial(FitState &s)
{
    double qtf[MAXM];
    const int n = s.M;
    for (int j = 0; j < n; j++) {
        double sum = 0;
        for (int i = 0; i < n; i++) sum += A_(s.Q, i, j) * s.f[i];
        qtf[j] = sum;
    }
    dogleg(s, qtf, s.dx);
    for (int i = 0; i < n; i++) s.x[i] = s.xc[i] + s.dx[i];
    s.phase = PH_HYB_TRIAL;
    return PASS_GRAD;
}

// after an iterate: residual test + iteration cap of node_binomial.c:100-115; decides what is next
ABN_HDN int hy_after_iterate(FitState &s, const Control &c, int gsl_status)
{
    int status = gsl_status; // 0 = success so far
    if (!status) {
        double res = 0;
        for (int i = 0; i < s.M; i++) res += fabs(s.f[i]);
        status = (res < c.epsabs) ? 0 : -2; // GSL_CONTINUE
        if (status == -2 && s.loop_iter < c.max_iters) return hy_next_trial(s);
    }
    // this attempt is over
    if (status == 0) {
        s.status = FIT_OK;
    } else {
        s.status = FIT_NOT_CONVERGED;
        bool valid;
        s.attempt++;
        hy_attempt_scale(s, s.attempt, &valid);
        if (valid) { s.loop_iter = 0; return hy_start_attempt(s); }
    }
    for (int i = 0; i < s.M; i++) s.x[i] = s.xc[i];
    s.phase = PH_FINAL;
    return PASS_FULL;
}

// ---- IRLS shortcuts -----------------------------------------------------------------------------
// The reference iterates until a step is shorter than tol (1e-8). Newton's quadratic convergence
// makes the steps run ... 1e-2, 1e-4, 1e-8: the last one or two data passes only confirm what
// the earlier ones imply, and they change the reported score far below the 1e-8 parity bar. The
// state machines below stop paying for them, under bounds evaluated per fit:
//
//   q = sqrt(d'Hd / H_00)   length of a Newton step d in the scale of the linear predictor (the
//                           weight-averaged RMS change of eta; H_00 = sum of the IRLS weights),
//                           so the bounds do not depend on how the columns are scaled
//   rho = q_t / q_(t-1)     observed contraction; the step after d is predicted as q rho^2
//                           (quadratic convergence), times a safety factor of 10
//
//   Taylor finish  the loglik at x + d is taken as ll(x) + grad'd - d'Hd / 2 when the cubic
//                  remainder (sc_cubic) and the effect of the error still left after the step are
//                  each under SC_REL |ll|. Poisson: grad = g (plain ML: the fixed point is
//                  stationary, the remaining error e enters as H_00 e^2 / 2, e <= q rho). Binomial:
//                  grad = X'(y - g) (PassAcc::gl); the bias-reduced fixed point is NOT stationary for
//                  the reported loglik - an error of eta-length e moves it by <= (k/2) e - so the
//                  bound is (k/2) 2 q rho, with the linear-rate estimate of e.
//   predicted last trial (Poisson)  when H_00 (3 q rho^2)^2 / 2 < SC_REL |ll| the trial point only
//                  gets a loglik pass; on acceptance the quadratic model is checked against the gain
//                  it predicted (|gain - model| < 5 % of d'Hd/2), otherwise the point is redone in full.
//
// All of it needs an intercept column (H_00) and a contracting sequence (rho < 1); ABN_FLAG_EXACT_PATH in
// abn_control.flags (or -DABN_SHORTCUT=0)
// restores the reference's pass-for-pass iteration. tests/test_host_sim.py and the GPU parity
// tests hold both against the oracle, which iterates exactly like the reference.
#ifndef ABN_SHORTCUT
#define ABN_SHORTCUT 1
#endif
constexpr double SC_REL = 1e-11; // share of |loglik| each neglected term may reach
constexpr double SC_FULL_Q = 3e-4; // binomial: predicted eta-length under which the next pass also sums the loglik

// bound on the cubic remainder of the Taylor step: the third derivative of either family's loglik
// term is at most its weight w, so the remainder is at most (1/6) sum w |x'd|^3 ~ (1/6) H_00 q^3
// times a tail factor (10) for the spread of |x'd| around its RMS
ABN_HD double sc_cubic(double h00, double q) { return (10.0 / 6.0) * h00 * q * q * q; }

ABN_HD double quad_form(const double *A, const double *d, int k)
{
    double q = 0.0;
    for (int i = 0; i < k; i++) {
        double r = 0.0;
        for (int j = 0; j < k; j++) r += A_(const_cast<double *>(A), i, j) * d[j];
        q += d[i] * r;
    }
    return q;
}

// eta-scale length of d, or -1 when it is not defined (no intercept, non-finite or non-positive sums)
ABN_HD double eta_len(const FitState &s, double dHd)
{
    const double h00 = s.J[0];
    if (!s.shortcut || !s.icpt || !(h00 > 0.0) || !(dHd >= 0.0) || !isfinite(dHd) || !isfinite(h00)) return -1.0;
    return sqrt(dHd / h00);
}

// ---- entry points -------------------------------------------------------------------------------
// Caller fills: family, k, M, n, lfact, x0 (Bayes start values / MLE Gaussian unused).
ABN_HDN int fit_begin(FitState &s, const Control &c)
{
    s.iters = 0; s.passes = 0; s.status = FIT_OK;
    s.loglik = s.aic = s.bic = NAN;
    s.loop_iter = 0; s.attempt = 0;
    s.q_cur = s.q_prev = 0.0; s.gd = s.dHd = 0.0;
    s.shortcut = (ABN_SHORTCUT && !(c.flags & 1)) ? 1 : 0;
    if (c.method == MLE) {
        if (s.family == BINOMIAL) {
            for (int i = 0; i < s.k; i++) s.x[i] = 0.0; // x.zeros, irls_binomial_fast_br.cpp:19
            s.phase = PH_MLE_BIN;
            s.have_ll = (c.max_iters <= 1);
            return s.have_ll ? PASS_FULL : PASS_HG;
        }
        s.phase = PH_POIS_INIT;
        return PASS_PINIT;
    }
    if (c.solver == HYBRID) return hy_start_attempt(s);
    for (int i = 0; i < s.M; i++) s.x[i] = s.x0[i];
    s.phase = PH_NEWTON;
    return PASS_HG;
}

// Pass a Poisson-mle trial point needs: when the step to it is already shorter than the stop
// tolerance (or the iteration cap is reached) the trial can only end the fit - accepted it
// converges (irls_poisson_fast.cpp:129), rejected the next half step is shorter still - so only
// its log-likelihood is needed, not the Gram sums.
ABN_HD int pois_trial_mode(FitState &s, const Control &c)
{
    double nrm = 0.0;
    for (int i = 0; i < s.k; i++) { double d = s.x[i] - s.xold[i]; nrm += d * d; }
    nrm = sqrt(nrm);
#if defined(ABN_TRACE_STEPS) && !defined(__CUDA_ARCH__)
    fprintf(stderr, "STEP %p %d %d %.3e\n", (void *)&s, s.iters, s.halvings, nrm);
#endif
    bool pred = false;
    if (s.halvings == 0) { // full Newton step: is the point after it predicted to be converged?
        const double q = s.q_cur;
        if (q >= 0.0 && s.q_prev > 0.0 && q < s.q_prev) {
            const double rho = q / s.q_prev;
            const double rem = 3.0 * q * rho * rho; // checked against the observed gain on acceptance
            pred = 0.5 * s.J[0] * rem * rem < SC_REL * fabs(s.cur_ll);
        }
    }
    const bool last = (nrm < c.tol) || (s.iters >= c.max_iters) || pred;
    s.have_ll = last ? 2 : 1; // 2: the pass in flight has no H / g
    return last ? PASS_LL : PASS_FULL;
}

ABN_HD void mle_scores(FitState &s, double ll, double df)
{
    s.loglik = ll;
    s.aic = -2 * ll + 2 * df;
    s.bic = -2 * ll + log(s.n) * df;
    if (!isfinite(ll)) s.status = FIT_NON_FINITE;
}

ABN_HD void unpack_sym(const PassAcc &a, int k, double *A)
{
    for (int i = 0; i < k; i++)
        for (int j = i; j < k; j++) A_(A, i, j) = A_(A, j, i) = a.H[tri(i, j, k)];
}

ABN_HDN int fit_consume(FitState &s, const Control &c, const PassAcc &a)
{
    s.passes++;
    const int k = s.k, M = s.M;
    switch (s.phase) {
    case PH_MLE_BIN: {
        // a.g = X'[rho (b' - g)], a.H = X'WX at s.x (irls_binomial_fast_br.cpp:51-73); a.ll = loglik at
        // s.x when this pass was asked to produce it. The reported loglik is the one at the
        // coefficients BEFORE the last update (:51,:97): it is computed in the pass that detects
        // convergence when the previous step predicted it (quadratic convergence), otherwise in one
        // extra loglik-only pass at the same point - identical value either way.
        s.iters++;
        double step[MAXM];
        unpack_sym(a, k, s.J);
        int st = a.bad ? -1 : sym_solve(s.J, a.g, k, false, step, s.Q);
        if (st < 0) { // non-finite weights: Armadillo's solve() on such input is unpinned
            s.status = FIT_NON_FINITE;
            mle_scores(s, NAN, (double)k);
            return PASS_DONE;
        }
        double nrm = 0.0;
        for (int i = 0; i < k; i++) { s.xold[i] = s.x[i]; s.x[i] += step[i]; nrm += step[i] * step[i]; }
        nrm = sqrt(nrm);
#if defined(ABN_TRACE_STEPS) && !defined(__CUDA_ARCH__)
        fprintf(stderr, "BSTEP %d %.3e ll %.10e\n", s.iters, nrm, a.ll);
#endif
        const bool converged = nrm < c.tol;                       // :81
        if (converged || s.iters >= c.max_iters) {
            if (!converged) s.status = FIT_NOT_CONVERGED;
            if (s.have_ll) { mle_scores(s, a.ll, (double)k); return PASS_DONE; }
            for (int i = 0; i < k; i++) { s.dir[i] = s.x[i]; s.x[i] = s.xold[i]; } // dir keeps the final coefficients
            s.phase = PH_MLE_BIN_LL;
            return PASS_LL;
        }
        // shortcuts (see "IRLS shortcuts" above)
        const double dHd = quad_form(s.J, step, k);
        const double q = eta_len(s, dHd);
        double pred_next = nrm;
#if defined(ABN_TRACE_STEPS) && !defined(__CUDA_ARCH__)
        fprintf(stderr, "  BQ q %.3e prev %.3e have_ll %d has_gl %d ll %.6e h00 %.3e\n", q, s.q_cur, s.have_ll, a.has_gl, a.ll, s.J[0]);
#endif
        if (q >= 0.0 && s.q_cur > 0.0 && q < s.q_cur) {
            const double rho = q / s.q_cur;
            if (s.have_ll && a.has_gl && isfinite(a.ll)) {
                const double budget = SC_REL * fabs(a.ll);
                const double cubic = sc_cubic(s.J[0], q);
                const double rem = 2.0 * q * rho; // eta-length still missing after this step, linear-rate bound
                if (cubic < budget && 0.5 * k * rem < budget) {
                    double gd = 0.0;
                    for (int i = 0; i < k; i++) gd += a.gl[i] * step[i];
                    mle_scores(s, a.ll + gd - 0.5 * dHd, (double)k); // loglik at the updated point
                    return PASS_DONE;
                }
            }
            pred_next = 3.0 * q * rho * rho;
        } else if (q >= 0.0) pred_next = q;
        s.q_cur = q;
        s.have_ll = (pred_next < (s.shortcut ? SC_FULL_Q : 1e-4)) || (nrm < 1e-4) || (s.iters + 1 >= c.max_iters);
        return s.have_ll ? PASS_FULL : PASS_HG;
    }
    case PH_MLE_BIN_LL: {
        const int keep = s.status;
        for (int i = 0; i < k; i++) s.x[i] = s.dir[i]; // the coefficients after the last update (returned ones)
        mle_scores(s, a.ll, (double)k);
        if (s.status == FIT_OK) s.status = keep;
        return PASS_DONE;
    }
    case PH_POIS_INIT: {
        // irls_poisson_fast.cpp:37-58
        unpack_sym(a, k, s.J);
        bool ok = !a.bad;
        for (int i = 0; i < NTRI && ok; i++) if (i < k * (k + 1) / 2 && !isfinite(a.H[i])) ok = false;
        for (int i = 0; i < k; i++) s.x[i] = 0.0;
        if (ok) {
            double x[MAXM];
            sym_solve(s.J, a.g, k, true, x, s.Q);
            bool fin = true;
            for (int i = 0; i < k; i++) if (!isfinite(x[i])) fin = false;
            if (fin) for (int i = 0; i < k; i++) s.x[i] = x[i];
        }
        s.phase = PH_POIS_FIRST;
        return PASS_FULL;
    }
    case PH_POIS_FIRST:
    case PH_POIS_TRIAL: {
        // a.H = X'WX, a.g = X'(y - W), a.ll = sum(y eta - exp(eta)) at s.x
        double ll_here = a.ll - s.lfact;
        if (s.phase == PH_POIS_TRIAL) {
            // step-halving test at the trial point (irls_poisson_fast.cpp:98-120)
            bool acc = !a.bad && isfinite(ll_here) && ll_here >= s.cur_ll;
            if (!acc) {
                s.step *= 0.5;
                s.halvings++;
                if (s.halvings < 10) {
                    for (int i = 0; i < k; i++) s.x[i] = s.xold[i] + s.step * s.dir[i];
                    bool fin = true;
                    for (int i = 0; i < k; i++) if (!isfinite(s.x[i])) fin = false;
                    // a non-finite x_new costs a halving but no data pass (:102-105)
                    while (!fin && s.halvings < 10) {
                        s.step *= 0.5; s.halvings++;
                        fin = true;
                        for (int i = 0; i < k; i++) { s.x[i] = s.xold[i] + s.step * s.dir[i]; if (!isfinite(s.x[i])) fin = false; }
                    }
                    if (fin && s.halvings < 10) return pois_trial_mode(s, c);
                }
                // no valid step: revert, norm(x - xold) = 0 < tol -> stop (:122-131); ll stays cur_ll
                for (int i = 0; i < k; i++) s.x[i] = s.xold[i];
                mle_scores(s, s.cur_ll, (double)k);
                return PASS_DONE;
            }
            // accepted: convergence test (:129)
            double nrm = 0.0;
            for (int i = 0; i < k; i++) { double d = s.x[i] - s.xold[i]; nrm += d * d; }
            if (sqrt(nrm) < c.tol) { mle_scores(s, ll_here, (double)k); return PASS_DONE; } // final ll :137 at the same x
            if (s.iters >= c.max_iters) { s.status = FIT_NOT_CONVERGED; mle_scores(s, ll_here, (double)k); return PASS_DONE; }
            if (s.have_ll == 2) {
                // predicted last trial: accept it as final when the quadratic model that predicted
                // it also predicted its gain; otherwise evaluate the point in full and go on
                const double model = s.gd - 0.5 * s.dHd;
                if (fabs((ll_here - s.cur_ll) - model) < 0.05 * 0.5 * s.dHd) { mle_scores(s, ll_here, (double)k); return PASS_DONE; }
                s.phase = PH_POIS_FIRST;
                s.iters--;
                s.have_ll = 1;
                s.q_cur = 0.0;
                return PASS_FULL;
            }
        }
        // next IRLS iteration from s.x with this pass's sums (:62-91)
        if (s.phase == PH_POIS_FIRST && c.max_iters <= 0) { mle_scores(s, ll_here, (double)k); return PASS_DONE; }
        if (a.bad) { // non-finite weights -> break (:66-69); final ll at x is non-finite too
            s.status = FIT_NOT_CONVERGED;
            mle_scores(s, ll_here, (double)k);
            return PASS_DONE;
        }
        s.iters++;
        s.cur_ll = isfinite(ll_here) ? ll_here : -INFINITY;
        unpack_sym(a, k, s.J);
        bool finH = true;
        for (int i = 0; i < k; i++) {
            A_(s.J, i, i) += 1e-8; // :81-82
            for (int j = 0; j < k; j++) if (!isfinite(A_(s.J, i, j))) finH = false;
        }
        if (!finH) { s.status = FIT_NOT_CONVERGED; mle_scores(s, ll_here, (double)k); return PASS_DONE; }
        // solve(G, X'Wz) - x with X'Wz = (X'WX) x + X'(y - W)  ==  solve(G, g - lambda x)
        double rhs[MAXM];
        for (int i = 0; i < k; i++) rhs[i] = a.g[i] - 1e-8 * s.x[i];
        sym_solve(s.J, rhs, k, true, s.dir, s.Q);
        {
            // shortcuts (see "IRLS shortcuts" above): Taylor finish at x + dir
            s.dHd = quad_form(s.J, s.dir, k);
            s.gd = 0.0;
            for (int i = 0; i < k; i++) s.gd += a.g[i] * s.dir[i];
            const double q = eta_len(s, s.dHd);
            const double pq = s.q_cur;
            s.q_prev = pq;
            s.q_cur = q;
            const double budget = SC_REL * fabs(ll_here);
            if (q >= 0.0 && pq > 0.0 && q < pq && isfinite(s.gd) && isfinite(ll_here) && sc_cubic(s.J[0], q) < budget &&
                0.5 * s.J[0] * (q * q / pq) * (q * q / pq) < budget) {
                bool fin = true;
                for (int i = 0; i < k; i++) { s.x[i] += s.dir[i]; if (!isfinite(s.x[i])) fin = false; }
                if (fin) { mle_scores(s, ll_here + s.gd - 0.5 * s.dHd, (double)k); return PASS_DONE; }
                for (int i = 0; i < k; i++) s.x[i] -= s.dir[i];
            }
        }
        for (int i = 0; i < k; i++) s.xold[i] = s.x[i];
        s.step = 1.0;
        s.halvings = 0;
        bool fin = false;
        while (!fin && s.halvings < 10) {
            fin = true;
            for (int i = 0; i < k; i++) { s.x[i] = s.xold[i] + s.step * s.dir[i]; if (!isfinite(s.x[i])) fin = false; }
            if (!fin) { s.step *= 0.5; s.halvings++; }
        }
        if (!fin) {
            for (int i = 0; i < k; i++) s.x[i] = s.xold[i];
            mle_scores(s, s.cur_ll, (double)k);
            return PASS_DONE;
        }
        s.phase = PH_POIS_TRIAL;
        return pois_trial_mode(s, c);
    }
    case PH_NEWTON: {
        double f[MAXM], step[MAXM];
        bayes_eval(s, c, s.x, a, f, s.J, nullptr);
        double res = 0;
        bool fin = true;
        for (int i = 0; i < M; i++) { res += fabs(f[i]); if (!isfinite(f[i])) fin = false; }
        if (res < c.epsabs) { s.status = FIT_OK; }
        else if ((s.iters >= c.max_iters) || !fin) { s.status = FIT_NOT_CONVERGED; }
        else {
            int perm[MAXM];
            for (int i = 0; i < M * MAXM; i++) s.Q[i] = s.J[i];
            lu_decomp(s.Q, M, perm);
            if (lu_singular(s.Q, M)) { s.status = FIT_NOT_CONVERGED; }
            else {
                lu_solve(s.Q, perm, M, f, step);
                for (int i = 0; i < M; i++) s.x[i] -= step[i];
                s.iters++;
                return PASS_HG;
            }
        }
        // g(x) is still missing: one loglik-only pass at the same point, hess g(x) stays in s.J
        s.phase = PH_FINAL_LL;
        return PASS_LL;
    }
    case PH_FINAL_LL: {
        double gval;
        bayes_eval(s, c, s.x, a, nullptr, nullptr, &gval);
        bayes_score(s, gval);
        return PASS_DONE;
    }
    case PH_HYB_SET: {
        // gsl hybridj.c set(): f and J at the start point
        bayes_eval(s, c, s.x, a, s.f, s.J, nullptr);
        s.hy_iter = 1;
        s.fnorm = enorm(s.f, M);
        s.ncfail = s.ncsuc = s.nslow1 = s.nslow2 = 0;
        for (int i = 0; i < M; i++) { s.dx[i] = 0.0; s.diag[i] = 1.0; }
        if (s.scale) hy_diag(s.J, M, s.diag, false);
        s.delta = hy_delta(s.diag, s.xc, M);
        qr_decomp(s.J, M, s.Q, s.R);
        s.loop_iter = 1; // the do-loop's iter++ before the first iterate
        return hy_next_trial(s);
    }
    case PH_HYB_TRIAL: {
        // gsl hybridj.c iterate() from "evaluate function at x + p" on
        const double p1 = 0.1, p5 = 0.5, p001 = 0.001, p0001 = 0.0001;
        double f_trial[MAXM], df[MAXM], rdx[MAXM], qtf[MAXM], qtdf[MAXM], w[MAXM], v[MAXM];
        s.iters++;
        bayes_eval(s, c, s.x, a, f_trial, nullptr, nullptr);
        for (int j = 0; j < M; j++) {
            double sum = 0;
            for (int i = 0; i < M; i++) sum += A_(s.Q, i, j) * s.f[i];
            qtf[j] = sum;
        }
        const double fnorm = s.fnorm;
        double pnorm = scaled_enorm(s.diag, s.dx, M);
        if (s.hy_iter == 1 && pnorm < s.delta) s.delta = pnorm;
        for (int i = 0; i < M; i++) df[i] = f_trial[i] - s.f[i];
        double fnorm1 = enorm(f_trial, M);
        double actred = (fnorm1 < fnorm) ? 1 - (fnorm1 / fnorm) * (fnorm1 / fnorm) : -1;
        for (int i = 0; i < M; i++) {
            double sum = 0;
            for (int j = i; j < M; j++) sum += A_(s.R, i, j) * s.dx[j];
            rdx[i] = sum;
        }
        double fnorm1p = 0;
        for (int i = 0; i < M; i++) { double u = qtf[i] + rdx[i]; fnorm1p += u * u; }
        fnorm1p = sqrt(fnorm1p);
        double prered = (fnorm1p < fnorm) ? 1 - (fnorm1p / fnorm) * (fnorm1p / fnorm) : 0;
        double ratio = (prered > 0) ? actred / prered : 0;
        if (ratio < p1) { s.ncsuc = 0; s.ncfail++; s.delta *= p5; }
        else {
            s.ncfail = 0; s.ncsuc++;
            if (ratio >= p5 || s.ncsuc > 1) s.delta = fmax(s.delta, pnorm / p5);
            if (fabs(ratio - 1) <= p1) s.delta = pnorm / p5;
        }
        if (ratio >= p0001) {
            for (int i = 0; i < M; i++) { s.xc[i] = s.x[i]; s.f[i] = f_trial[i]; }
            s.fnorm = fnorm1;
            s.hy_iter++;
        }
        s.nslow1++;
        if (actred >= p001) s.nslow1 = 0;
        if (actred >= p1) s.nslow2 = 0;
        if (s.ncfail == 2) { // re-evaluate the Jacobian at the current point
            for (int i = 0; i < M; i++) s.x[i] = s.xc[i];
            s.phase = PH_HYB_JAC;
            return PASS_HESS;
        }
        for (int j = 0; j < M; j++) {
            double sum = 0;
            for (int i = 0; i < M; i++) sum += A_(s.Q, i, j) * df[i];
            qtdf[j] = sum;
        }
        for (int i = 0; i < M; i++) {
            w[i] = (qtdf[i] - rdx[i]) / pnorm;
            v[i] = s.diag[i] * s.diag[i] * s.dx[i] / pnorm;
        }
        qr_update(s.Q, s.R, M, w, v);
        int gst = 0;
        if (s.nslow2 == 5) gst = 28;      // GSL_ENOPROGJ
        else if (s.nslow1 == 10) gst = 27; // GSL_ENOPROG
#if defined(ABN_TRACE_STEPS) && !defined(__CUDA_ARCH__)
        { double res = 0; for (int i = 0; i < M; i++) res += fabs(s.f[i]); fprintf(stderr, "HY fam %d iter %d res %.3e ratio %.3f\n", s.family, s.loop_iter, res, ratio); }
#endif
        int nxt = hy_after_iterate(s, c, gst);
        if (nxt == PASS_GRAD) s.loop_iter++;
        return nxt;
    }
    case PH_HYB_JAC: {
        bayes_eval(s, c, s.x, a, nullptr, s.J, nullptr);
        s.nslow2++;
        if (s.hy_iter == 1) {
            if (s.scale) hy_diag(s.J, M, s.diag, false);
            s.delta = hy_delta(s.diag, s.xc, M);
        } else if (s.scale) {
            hy_diag(s.J, M, s.diag, true);
        }
        qr_decomp(s.J, M, s.Q, s.R);
        int nxt = hy_after_iterate(s, c, 0);
        if (nxt == PASS_GRAD) s.loop_iter++;
        return nxt;
    }
    case PH_FINAL: {
        bayes_finish(s, c, a);
        return PASS_DONE;
    }
    }
    return PASS_DONE;
}

} // namespace ABN_NS
