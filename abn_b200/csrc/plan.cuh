// plan.cuh - per-candidate design resolution and everything that needs only the global Gram.
//
// The global Gram G = Z^T Z is computed once per data set over
//     Z = [ 1 | every stored column | log(y + 0.1) for every Poisson node ]
// so that (a) Gaussian children (either method) are scored with no data pass at all
// (src/irls_gaussian_fast.cpp:34-49, src/node_gaussian.c), (b) the least-squares start values of
// the Bayes solvers (generate_inits_n*, src/node_binomial.c:832-883, src/node_poisson.c:705-763,
// src/node_gaussian.c:743-808) and (c) the rank test of rank_cpp (src/rank.cpp:15, called at
// R/build_score_cache_mle.R:282) come from k x k sub-blocks instead of n x k passes.
#pragma once

#include "solver.cuh"

namespace ABN_NS {

ABN_HD int gram_icpt() { return 0; }
ABN_HD int gram_col(int c) { return 1 + c; }
// virtual columns exist for Poisson nodes only, in the order of node_gw
ABN_HD int gram_logy(const DataMeta &m, int node) { return 1 + m.n_cols + m.node_gw[node]; }

// one candidate's resolved design
struct FitDesc {
    int cols[MAXK];       // stored column of design column j (-1 = intercept)
    int ycol;             // first stored column of the child
    short k;              // design columns after rank handling
    short kdata;          // selected stored columns before rank handling (for MDL)
    short num_na;         // leading columns dropped (binomial, rank deficient)
    short family;         // child family
    short rank;           // rank of the full design as rank_cpp sees it
    short err;            // 1 = design wider than MAXK, 2 = child not supported
    short nlev;           // levels of a multinomial child (else 1)
    short icpt0;          // the design had an intercept before rank handling
};

// kernel bucket of a resolved fit: the streaming kernels are instantiated per (row kind, k,
// intercept); Gaussian fits go to the Gram-only kernel. Family slots: 0 binomial, 1 Poisson,
// 2.. multinomial child with 2, 3, ... levels.
constexpr int MN_MAXL_ = 5;
constexpr int NFAMSLOT = 2 + (MN_MAXL_ - 1);
ABN_HD int bucket_of(const FitDesc &d)
{
    if (d.err) return 0;
    if (d.family == GAUSSIAN) return 1;
    int icpt = (d.k > 0 && d.cols[0] < 0) ? 1 : 0;
    int fam = d.family == BINOMIAL ? 0 : (d.family == POISSON ? 1 : d.nlev);
    return 2 + (fam * (MAXK + 1) + d.k) * 2 + icpt;
}
constexpr int NBUCKETS = 2 + NFAMSLOT * (MAXK + 1) * 2;
// all-discrete designs (fit_cell.cuh) get four buckets of their own after the streaming ones:
// CELL_BUCKET0 + 3 * (Poisson child) + class (0: small table, 1: large table, 2: all-binary design on bit planes)
constexpr int CELL_BUCKET0 = NBUCKETS;
constexpr int NBUCKETS_ALL = NBUCKETS + 6;
static_assert(MAXK > 8 || NBUCKETS_ALL <= 256, "bucket keys are bytes (the wide build does not bucket)");

ABN_HD int gidx(const FitDesc &d, int j) { return d.cols[j] < 0 ? 0 : 1 + d.cols[j]; }

// k x k Gram block of the design, MAXM-strided
ABN_HD void gram_block(const DataMeta &m, const FitDesc &d, int j0, int k, double *A)
{
    for (int i = 0; i < k; i++)
        for (int j = 0; j < k; j++)
            A_(A, i, j) = m.G[(long long)gidx(d, j0 + i) * m.D + gidx(d, j0 + j)];
}

// Cheap sufficient test for full rank, taken before the eigenvalue test: remove the intercept (rank [1 | X] =
// 1 + rank of the centred X), scale to unit diagonal and run a Cholesky factorisation; if every pivot of that
// correlation matrix stays above 1e-4 the design is far from any rank decision (the reference's SVD rule only
// drops a singular value below max(n, k) eps sigma_max) and the answer is k. This settles all but a handful of the
// 1e7 candidates of a cache with ~k^3 / 3 flops instead of Jacobi sweeps, and - being invariant to the location
// and scale of the columns - it also gives the reference's answer for an uncentred column with |mean| >> sd, which
// the eigenvalue test on raw moments below would call collinear with the intercept.
ABN_HDN bool gram_full_rank_fast(const double *A, int k, bool icpt_first)
{
    double C[MAXM * MAXM];
    const int o = icpt_first ? 1 : 0, k1 = k - o;
    if (icpt_first && !(A_(A, 0, 0) > 0.0)) return false;
    for (int i = 0; i < k1; i++)
        for (int j = 0; j <= i; j++) {
            double c = A_(A, o + i, o + j);
            if (icpt_first) c -= A_(A, 0, o + i) * A_(A, 0, o + j) / A_(A, 0, 0);
            A_(C, i, j) = c;
        }
    double sc[MAXM];
    for (int i = 0; i < k1; i++) {
        const double dgi = A_(C, i, i);
        if (!(dgi > 0.0) || !isfinite(dgi)) return false;
        // a centred second moment that is mostly cancellation (|mean| > ~1e5 sd) is not trusted here
        if (icpt_first && dgi < 1e-10 * A_(A, o + i, o + i)) return false;
        sc[i] = 1.0 / sqrt(dgi);
    }
    for (int j = 0; j < k1; j++) {
        double s = 1.0; // unit diagonal after scaling
        for (int t = 0; t < j; t++) s -= A_(C, j, t) * A_(C, j, t);
        if (!(s > 1e-4)) return false;
        const double l = sqrt(s);
        A_(C, j, j) = l;
        for (int i = j + 1; i < k1; i++) {
            double v = A_(C, i, j) * sc[i] * sc[j];
            for (int t = 0; t < j; t++) v -= A_(C, i, t) * A_(C, j, t);
            A_(C, i, j) = v / l;
        }
    }
    return true;
}

// arma::rank of the n x k design from the eigenvalues of its Gram block: singular values
// sqrt(lambda_i) against max(n,k) * sigma_max * eps, with the floor FP64 accumulation of the
// Gram imposes on what can be told from zero.
ABN_HDN int gram_rank(const DataMeta &m, const FitDesc &d, int j0, int k)
{
    if (k <= 0) return 0;
    double A[MAXM * MAXM], V[MAXM * MAXM], w[MAXM];
    gram_block(m, d, j0, k, A);
    if (gram_full_rank_fast(A, k, gidx(d, j0) == 0)) return k;
    jacobi_eig(A, k, w, V);
    double lmax = 0.0;
    for (int i = 0; i < k; i++) if (w[i] > lmax) lmax = w[i];
    const double eps = 2.220446049250313e-16;
    double nn = (double)(m.n > k ? m.n : k);
    double tol_sv = nn * sqrt(lmax) * eps;
    double noise = fmax(64.0, 8.0 * sqrt((double)m.n)) * k * eps * lmax;
    double thr = fmax(tol_sv * tol_sv, noise);
    int r = 0;
    for (int i = 0; i < k; i++) if (w[i] > thr) r++;
    return r;
}

// forLoopContent's design assembly, R/build_score_cache_mle.R:254-306 (mle) and
// build_designmatrix, src/node_binomial.c:497-572 (bayes)
ABN_HDN void resolve_design(const DataMeta &m, int method, int child, const int *parents, int width, FitDesc &d)
{
    d.family = (short)m.node_family[child];
    d.ycol = m.node_col0[child];
    d.num_na = 0;
    d.err = 0;
    d.nlev = (short)(m.node_col0[child + 1] - m.node_col0[child]);
    d.icpt0 = 0;
    bool any_multi = false;
    int kdata = 0;
    for (int j = 0; j < width; j++) {
        int p = parents[j];
        if (p < 0) continue;
        if (m.node_family[p] == MULTINOMIAL) any_multi = true;
        kdata += m.node_col0[p + 1] - m.node_col0[p];
    }
    bool icpt = (method == BAYES) || !any_multi; // _mle.R:266-270
    d.icpt0 = icpt ? 1 : 0;
    int k = kdata + (icpt ? 1 : 0);
    d.kdata = (short)kdata;
    for (int j = 0; j < MAXK; j++) d.cols[j] = -1;
    if (k > MAXK) { d.err = 1; d.k = 0; d.rank = 0; return; }
    int c = 0;
    if (icpt) d.cols[c++] = -1;
    for (int j = 0; j < width; j++) {
        int p = parents[j];
        if (p < 0) continue;
        for (int col = m.node_col0[p]; col < m.node_col0[p + 1]; col++) d.cols[c++] = col;
    }
    d.k = (short)k;
    d.rank = (short)k;
    if (method == MLE) {
        int R = gram_rank(m, d, 0, k);                         // _mle.R:282
        d.rank = (short)R;
        if (R != k && d.family == BINOMIAL) {                  // _mle.R:286-294: drop leading columns
            int j0 = 0;
            while (k - j0 > 0 && gram_rank(m, d, j0, k - j0) != k - j0) j0++;
            for (int j = 0; j + j0 < k; j++) d.cols[j] = d.cols[j + j0];
            // a dropped intercept leaves only data columns; mark slot 0 accordingly
            d.k = (short)(k - j0);
            d.num_na = (short)j0;
            for (int j = d.k; j < MAXK; j++) d.cols[j] = -1;
        }
    }
}

// ---- centred moments (DataMeta::Gc) ---------------------------------------------------------------
ABN_HD double gc_at(const DataMeta &m, int gi, int gj) { return m.Gc[(long long)gi * (1 + m.n_cols) + gj]; }
// sum (z_a - mean_a)(z_b - mean_b) for Gram indices a, b >= 1 (Gc is centred at col_mean, which is the true mean
// up to its rounding; the last term removes what is left)
ABN_HD double cmom(const DataMeta &m, int ga, int gb)
{
    return gc_at(m, ga, gb) - gc_at(m, 0, ga) * gc_at(m, 0, gb) / (double)m.n;
}
ABN_HD double cmean(const DataMeta &m, int ga) { return m.col_mean[ga - 1] + gc_at(m, 0, ga) / (double)m.n; }

// Centred blocks of a Gaussian child's design, when they apply: centred moments present, intercept in column 0.
ABN_HD bool gauss_centered_block(const DataMeta &m, const FitDesc &d, GaussAux &g)
{
    if (m.Gc == nullptr || d.k < 1 || d.cols[0] >= 0) return false;
    const int k1 = d.k - 1, gy = gram_col(d.ycol);
    g.k1 = k1;
    for (int i = 0; i < k1; i++) {
        const int gi = gidx(d, 1 + i);
        for (int j = 0; j < k1; j++) A_(g.Cxx, i, j) = cmom(m, gi, gidx(d, 1 + j));
        g.Cxy[i] = cmom(m, gi, gy);
        g.xbar[i] = cmean(m, gi);
    }
    g.Cyy = cmom(m, gy, gy);
    g.ybar = cmean(m, gy);
    return true;
}

// Standard errors as fitAbn(method = "mle") derives them from the `varcov` its IRLS functions return
// (R/fitabn_mle.R:715-745): se_i = sqrt(scale * inv(V)_ii) with V = X'WX of the last iteration (binomial,
// Poisson; scale = 1) or X'X (Gaussian; scale = mse = sse / (n - k)). V is MAXM-strided; NaN if singular.
ABN_HDN void se_from_information(const double *V, int k, double scale, double *se)
{
    double LU[MAXM * MAXM], e[MAXM], col[MAXM];
    int perm[MAXM];
    for (int i = 0; i < k; i++)
        for (int j = 0; j < k; j++) A_(LU, i, j) = A_(const_cast<double *>(V), i, j);
    lu_decomp(LU, k, perm);
    const bool bad = lu_singular(LU, k);
    for (int i = 0; i < k; i++) {
        if (bad) { se[i] = NAN; continue; }
        for (int j = 0; j < k; j++) e[j] = (j == i) ? 1.0 : 0.0;
        lu_solve(LU, perm, k, e, col);
        se[i] = sqrt(scale * col[i]);
    }
}

// ---- Gaussian child, method = "mle": src/irls_gaussian_fast.cpp:34-49 from Gram blocks ---------
ABN_HDN void gauss_mle_fit(const DataMeta &m, const FitDesc &d, double *loglik, double *aic, double *bic, int *status,
                           double *coef = nullptr, double *se = nullptr, double *sse = nullptr)
{
    const int k = d.k;
    double A[MAXM * MAXM], S[MAXM * MAXM], b[MAXM], x[MAXM];
    gram_block(m, d, 0, k, A);
    const long long gy = gram_col(d.ycol);
    for (int i = 0; i < k; i++) b[i] = m.G[(long long)gidx(d, i) * m.D + gy];
    int st;
    double ssr;
    GaussAux g;
    if (d.rank == k && gauss_centered_block(m, d, g)) {
        // full-rank design with an intercept: slopes from the centred normal equations, intercept from the means;
        // the residual sum of squares is formed from centred moments (no cancellation against n * mean(y)^2)
        st = sym_solve(g.Cxx, g.Cxy, g.k1, true, x + 1, S);
        x[0] = g.ybar;
        for (int i = 0; i < g.k1; i++) x[0] -= x[1 + i] * g.xbar[i];
        ssr = gauss_aux_rss(g, x, (double)m.n);
    } else {
        st = sym_solve(A, b, k, true, x, S);
        double yy = m.G[gy * m.D + gy];
        double xb = 0.0, xax = 0.0;
        for (int i = 0; i < k; i++) {
            double t = 0.0;
            for (int j = 0; j < k; j++) t += A_(A, i, j) * x[j];
            xax += t * x[i];
            xb += b[i] * x[i];
        }
        ssr = yy - 2.0 * xb + xax; // ||y - X x||^2
    }
    double n = (double)m.n, df = k + 1;
    double ll = 0.5 * (-n * (log(2 * M_PI) + 1 - log(n) + log(ssr)));
    *loglik = ll;
    *aic = -2 * ll + 2 * df;
    *bic = -2 * ll + log(n) * df;
    *status = (st < 0 || !isfinite(ll)) ? FIT_NON_FINITE : FIT_OK;
    if (coef)
        for (int i = 0; i < k; i++) coef[i] = x[i];
    if (se) se_from_information(A, k, ssr / (n - (double)k), se); // R/fitabn_mle.R:712-722
    if (sse) *sse = ssr;
}

// ---- Bayes start values: beta0 = inv(X'X) X'y~ by LU, as generate_inits_n* do -------------------
// y~ = y (binomial, Gaussian) or log(y + 0.1) (Poisson). Gaussian also gets tau0 = (n - m)/RSS.
ABN_HDN void bayes_inits(const DataMeta &m, const FitDesc &d, int child, FitState &s)
{
    const int k = d.k;
    double A[MAXM * MAXM], inv[MAXM * MAXM], b[MAXM], e[MAXM], col[MAXM];
    int perm[MAXM];
    gram_block(m, d, 0, k, A);
    const long long gy = (d.family == POISSON) ? gram_logy(m, child) : gram_col(d.ycol);
    for (int i = 0; i < k; i++) b[i] = m.G[(long long)gidx(d, i) * m.D + gy];
    lu_decomp(A, k, perm);
    bool singular = lu_singular(A, k);
    double beta[MAXM];
    if (!singular) {
        for (int j = 0; j < k; j++) {
            for (int i = 0; i < k; i++) e[i] = (i == j) ? 1.0 : 0.0;
            lu_solve(A, perm, k, e, col);
            for (int i = 0; i < k; i++) A_(inv, i, j) = col[i];
        }
        for (int i = 0; i < k; i++) {
            double t = 0.0;
            for (int j = 0; j < k; j++) t += A_(inv, i, j) * b[j];
            beta[i] = t;
        }
    } else {
        for (int i = 0; i < k; i++) beta[i] = (d.family == GAUSSIAN) ? 0.01 : 0.0;
    }
    for (int i = 0; i < k; i++) s.x0[i] = beta[i];
    if (d.family == GAUSSIAN) {
        gram_block(m, d, 0, k, A);
        double yy = m.G[gy * m.D + gy];
        double xb = 0.0, xax = 0.0;
        for (int i = 0; i < k; i++) {
            double t = 0.0;
            for (int j = 0; j < k; j++) t += A_(A, i, j) * beta[j];
            xax += t * beta[i];
            xb += b[i] * beta[i];
        }
        double rss = s.gaux ? gauss_aux_rss(*s.gaux, beta, (double)m.n) : yy - 2.0 * xb + xax;
        s.x0[k] = 1.0 / (rss / ((double)m.n - (double)k));
    }
}

// the "pass" of a Gaussian Bayes fit is a Gram lookup: H = X'X, g = X'y, ll = y'y
ABN_HD void gauss_gram_acc(const DataMeta &m, const FitDesc &d, PassAcc &a)
{
    const int k = d.k;
    const long long gy = gram_col(d.ycol);
    for (int i = 0; i < k; i++) {
        for (int j = i; j < k; j++) a.H[tri(i, j, k)] = m.G[(long long)gidx(d, i) * m.D + gidx(d, j)];
        a.g[i] = m.G[(long long)gidx(d, i) * m.D + gy];
    }
    a.ll = m.G[gy * m.D + gy];
    a.bad = 0;
    a.has_gl = 0;
}

// First IRLS iteration of the binomial mle from the Gram matrix: the reference starts at x = 0
// (irls_binomial_fast_br.cpp:19), where eta = 0 for every row, hence e = 1, g = 1/2, g' = 1/4,
// rho = g'/(g (1 - g)) = 1 exactly, W = 1/4 and W z = b' - 1/2 with b' = bp0 + y (bp1 - bp0):
//   X'WX = X'X / 4,   X'[rho (b' - g)] = (bp0 - 1/2) X'1 + (bp1 - bp0) X'y
// - all entries of the global Gram, so the first of the ~4-8 data passes of a fit is free.
ABN_HD void binomial_first_acc(const DataMeta &m, const FitDesc &d, PassAcc &a, bool bias_reduction = true)
{
    const int k = d.k;
    const double dfn = bias_reduction ? (double)k / (double)m.n : 0.0;
    const double bp0 = (0.0 + dfn * 0.5) / (1 + dfn), bp1 = (1.0 + dfn * 0.5) / (1 + dfn);
    const long long gy = gram_col(d.ycol);
    for (int i = 0; i < k; i++) {
        const long long gi = gidx(d, i);
        for (int j = i; j < k; j++) a.H[tri(i, j, k)] = 0.25 * m.G[gi * m.D + gidx(d, j)];
        a.g[i] = fma(bp1 - bp0, m.G[gi * m.D + gy], (bp0 - 0.5) * m.G[gi * m.D + 0]);
    }
    a.ll = -(double)m.n * 0.693147180559945309417; // loglik at eta = 0 (only used if max.iters = 1)
    a.bad = 0;
    a.has_gl = 0;
}

// Start values of the Poisson mle, no data pass. The reference starts from a weighted least-squares fit
// of log(y + 0.1) (irls_poisson_fast.cpp:37-52), about 0.26 away from the optimum in the scale of
// the linear predictor on BASELINE's data, and needs two Newton iterations to get within 1e-2. One
// Newton step from the intercept-only model is a better start and is made of Gram entries alone:
// at beta = (log ybar, 0, ..., 0) every mean is ybar, so H = ybar X'X and the score is X'(y - ybar):
//   beta = beta_null + (ybar X'X)^-1 X'(y - ybar)  <=>  solve(H, H beta_null + X'y - ybar X'1)
// with H beta_null = ybar log(ybar) X'1. It lands 2e-3..2e-2 from the optimum. The iteration
// converges to the same maximum either way (solver.cuh "IRLS shortcuts" on what that means for
// parity); an all-zero child (ybar = 0) starts at zero like a failed solve of the reference.
ABN_HD void poisson_init_acc(const DataMeta &m, const FitDesc &d, int child, PassAcc &a)
{
    (void)child;
    const int k = d.k;
    const long long gy = gram_col(d.ycol);
    const double ybar = m.G[gy] / (double)m.n; // G[0][y] = sum y
    const double c = ybar * (log(ybar) - 1.0);
    for (int i = 0; i < k; i++) {
        const long long gi = gidx(d, i);
        for (int j = i; j < k; j++) a.H[tri(i, j, k)] = ybar * m.G[gi * m.D + gidx(d, j)];
        a.g[i] = fma(c, m.G[gi * m.D], m.G[gi * m.D + gy]);
    }
    a.ll = 0.0;
    a.bad = !(ybar > 0.0) || !isfinite(c);
    a.has_gl = 0;
}

// common state setup
ABN_HD void state_setup(const DataMeta &m, const FitDesc &d, int child, const Control &c, FitState &s)
{
    s.family = d.family;
    s.k = d.k;
    s.M = d.k + ((c.method == BAYES && d.family == GAUSSIAN) ? 1 : 0);
    s.n = (double)m.n;
    s.lfact = m.node_lfact[child];
    s.icpt = (d.k > 0 && d.cols[0] < 0) ? 1 : 0;
    s.gaux = nullptr;
}

// Coefficient slots of a row of ResultPtrs::coef (abn_result.coef): slot 0 = intercept, slot 1 + t = the t-th
// stored column of the parents in candidate order (a multinomial parent owns one slot per level column); for a
// Gaussian child under method "bayes" the slot after the last column holds the precision. Columns the rank rule
// dropped, absent intercepts and rows without coefficients (multinomial children, failed designs) stay NaN.
ABN_HD void write_coef(const ResultPtrs &out, long long row, const FitDesc &d, const double *x, int M)
{
    if (!out.coef) return;
    double *o = out.coef + row * out.coef_stride;
    const int off = d.num_na + (d.icpt0 ? 0 : 1);
    for (int j = 0; j < d.k; j++)
        if (off + j < out.coef_stride) o[off + j] = x[j];
    if (M > d.k && 1 + d.kdata < out.coef_stride) o[1 + d.kdata] = x[d.k];
}

// `sse` of the reference's IRLS functions: e = b - A x on the scale of the linear predictor, whatever the family
// (irls_binomial_fast.cpp:128-130, irls_poisson_fast.cpp, irls_gaussian_fast.cpp:38-39) = y'y - 2 x'X'y + x'X'X x,
// Gram entries only
ABN_HD double sse_of(const DataMeta &m, const FitDesc &d, const double *x)
{
    const long long gy = gram_col(d.ycol);
    double s = m.G[gy * m.D + gy];
    for (int i = 0; i < d.k; i++) {
        const long long gi = gidx(d, i);
        double t = -2.0 * m.G[gi * m.D + gy];
        for (int j = 0; j < d.k; j++) t = fma(m.G[gi * m.D + gidx(d, j)], x[j], t);
        s = fma(t, x[i], s);
    }
    return s;
}

// standard errors of a finished mle fit whose state holds V = X'WX of its last iteration in s.J
ABN_HD void write_se(const ResultPtrs &out, long long row, const FitDesc &d, const FitState &s, int method)
{
    if (!out.se || method != MLE || s.status == FIT_NON_FINITE) return;
    double se[MAXM];
    se_from_information(s.J, d.k, 1.0, se);
    ResultPtrs o = out;
    o.coef = out.se;
    write_coef(o, row, d, se, d.k);
}

// MDL of a scored mle row: bic + (1 + #selected stored columns - num.na) * log(nvars), _mle.R:343
ABN_HD double mdl_of(double bic, const FitDesc &d, int nvars)
{
    return bic + (double)(1 + d.kdata - d.num_na) * log((double)nvars);
}

} // namespace ABN_NS
