/* TEST INFRASTRUCTURE - part of the CPU oracle (see abn_oracle.h). MLE fits. */
#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "abn_oracle.h"
#include "linalg.h"

/* Summation order of the Gram products. The reference leaves it to the BLAS behind Armadillo (blocked,
 * vectorised, platform dependent); 0 = ascending rows (default), 1 = descending rows. The tests use 1 to tell
 * fits whose result depends on that order - rounding-noise-driven near-separation fits, which the reference
 * itself does not reproduce across BLAS builds - from genuine parity failures. */
int or_sum_order_ = 0;
void or_set_sum_order(int order) { or_sum_order_ = order; }

/* A.t() * (W % A.each_col()) and A.t() * (W % z), X column-major n x k; w == NULL means 1 */
static void weighted_gram(const double *X, long n, int k, const double *w, const double *z,
                          double *G, double *rhs)
{
    const long i0 = or_sum_order_ ? n - 1 : 0, di = or_sum_order_ ? -1 : 1;
    for (int a = 0; a < k; a++) {
        const double *xa = X + (size_t)a * n;
        for (int b = a; b < k; b++) {
            const double *xb = X + (size_t)b * n;
            double s = 0.0;
            if (w) for (long c = 0, i = i0; c < n; c++, i += di) s += xa[i] * (w[i] * xb[i]);
            else   for (long c = 0, i = i0; c < n; c++, i += di) s += xa[i] * xb[i];
            G[a * k + b] = G[b * k + a] = s;
        }
        if (rhs) {
            double s = 0.0;
            if (w) for (long c = 0, i = i0; c < n; c++, i += di) s += xa[i] * (w[i] * z[i]);
            else   for (long c = 0, i = i0; c < n; c++, i += di) s += xa[i] * z[i];
            rhs[a] = s;
        }
    }
}

static void matvec(const double *X, long n, int k, const double *x, double *eta)
{
    for (long i = 0; i < n; i++) eta[i] = 0.0;
    for (int a = 0; a < k; a++) {
        const double *xa = X + (size_t)a * n;
        double c = x[a];
        for (long i = 0; i < n; i++) eta[i] += xa[i] * c;
    }
}

static int all_finite(const double *v, long n)
{
    for (long i = 0; i < n; i++) if (!isfinite(v[i])) return 0;
    return 1;
}

int or_rank_cpp(const double *X, long n, int k) { return or_rank(X, n, k); }

/* se[i] = sqrt(scale * inv(V)[i][i]), V k x k row-major (R: sqrt(diag(solve(varcov))), fitabn_mle.R:730,741;
 * Gaussian: sqrt(diag(mse * solve(t(X) %*% X))), :722); NaN where V is singular */
static void se_from_information(const double *V, int k, double scale, double *se)
{
    double LU[OR_MAXM * OR_MAXM], inv[OR_MAXM * OR_MAXM];
    int perm[OR_MAXM], signum;
    memcpy(LU, V, sizeof(double) * k * k);
    or_lu_decomp(LU, k, perm, &signum);
    if (or_lu_singular(LU, k) || or_lu_invert(LU, perm, k, inv) != 0) {
        for (int i = 0; i < k; i++) se[i] = NAN;
        return;
    }
    for (int i = 0; i < k; i++) se[i] = sqrt(scale * inv[i * k + i]);
}

/* ------------------------------------------------------------------------------------------ */
void or_mle_gaussian(const double *X, const double *y, long n, int k, or_fit *out)
{
    double G[OR_MAXM * OR_MAXM], rhs[OR_MAXM], x[OR_MAXM];
    memset(out, 0, sizeof(*out));
    weighted_gram(X, n, k, NULL, y, G, rhs);                 /* irls_gaussian_fast.cpp:34 */
    int st = or_arma_solve(G, rhs, k, 1, x);                 /* :35 solve(likely_sympd)   */
    double ssr = 0.0;                                        /* :38-39 */
    for (long i = 0; i < n; i++) {
        double e = y[i];
        for (int a = 0; a < k; a++) e -= X[(size_t)a * n + i] * x[a];
        ssr += e * e;
    }
    double dn = (double)n, df = k + 1;                       /* :40-41 */
    out->loglik = 0.5 * (-dn * (log(2 * M_PI) + 1 - log(dn) + log(ssr))); /* :45 */
    out->aic = -2 * out->loglik + 2 * df;                    /* :47 */
    out->bic = -2 * out->loglik + log(dn) * df;              /* :49 */
    out->iters = 1;
    out->passes = 1;
    out->status = (st < 0 || !isfinite(out->loglik)) ? ABN_FIT_NON_FINITE : ABN_FIT_OK;
    memcpy(out->coef, x, sizeof(double) * k);
    se_from_information(G, k, ssr / (dn - (double)k), out->se); /* mse = sse / df, df = n - (parents + 1), fitabn_mle.R:712-722 */
}

/* ------------------------------------------------------------------------------------------ */
/* both binomial IRLS variants of the reference: src/irls_binomial_fast_br.cpp (br = 1, line numbers below)
 * and src/irls_binomial_fast.cpp (br = 0: the same loop with b' = b, :47-76, loglik :89) */
static void mle_binomial_any(const double *X, const double *y, long n, int k, int maxit, double tol, int br,
                             or_fit *out)
{
    memset(out, 0, sizeof(*out));
    double x[OR_MAXM] = {0}, xold[OR_MAXM], G[OR_MAXM * OR_MAXM], rhs[OR_MAXM];
    double *eta = malloc(sizeof(double) * n), *W = malloc(sizeof(double) * n), *z = malloc(sizeof(double) * n);
    double nobs = (double)n, df = (double)k;
    int failed = 0, converged = 0, it = 0;
    for (long i = 0; i < n; i++) eta[i] = 0.0;
    for (it = 0; it < maxit; ++it) {                          /* irls_binomial_fast_br.cpp:50 */
        matvec(X, n, k, x, eta);                              /* :51 */
        for (long j = 0; j < n; ++j) {
            double e = exp(-1.0 * eta[j]);                    /* :54 */
            double g = 1.0 / (1.0 + e);                       /* :55 */
            double gprime = e / ((1.0 + e) * (1.0 + e));      /* :56 */
            double bprime = br ? (y[j] + (df / nobs) * (0.5)) / (1 + (df / nobs)) : y[j]; /* :58 */
            z[j] = eta[j] + (bprime - g) / gprime;            /* :63 */
            W[j] = (gprime * gprime) / (g * (1.0 - g));       /* :65-66 */
        }
        memcpy(xold, x, sizeof(double) * k);                  /* :68 */
        weighted_gram(X, n, k, W, z, G, rhs);                 /* :72 */
        int st = or_arma_solve(G, rhs, k, 0, x);              /* :73 no_approx + likely_sympd */
        out->iters = it + 1;
        if (st < 0) { failed = 1; break; }                    /* Armadillo: throw or NaN - unpinned */
        double s = 0.0;
        for (int a = 0; a < k; a++) s += (x[a] - xold[a]) * (x[a] - xold[a]);
        if (sqrt(s) < tol) { converged = 1; break; }          /* :81 */
    }
    /* :97 - eta is the one from the last loop pass (coefficients before the last update) */
    double t1 = 0.0, t2 = 0.0;
    for (long j = 0; j < n; j++) {
        t1 += y[j] * log(1.0 + exp(-eta[j]));
        t2 += (1.0 - y[j]) * log(1.0 + exp(eta[j]));
    }
    out->loglik = -t1 - t2;
    if (failed) out->loglik = NAN;
    out->aic = -2 * out->loglik + 2 * df;                     /* :99 */
    out->bic = -2 * out->loglik + log(nobs) * df;             /* :101 */
    out->passes = out->iters;
    out->status = (failed || !isfinite(out->loglik)) ? ABN_FIT_NON_FINITE
                  : (converged ? ABN_FIT_OK : ABN_FIT_NOT_CONVERGED);
    memcpy(out->coef, x, sizeof(double) * k);
    se_from_information(G, k, 1.0, out->se);                  /* varcov = varmatrix of the last loop pass */
    free(eta); free(W); free(z);
}

void or_mle_binomial_br(const double *X, const double *y, long n, int k, int maxit, double tol, or_fit *out)
{
    mle_binomial_any(X, y, n, k, maxit, tol, 1, out);
}

/* src/irls_binomial_fast.cpp:14-150 - what fitAbn(method = "mle") tries first (R/fitabn_mle.R:585) */
void or_mle_binomial(const double *X, const double *y, long n, int k, int maxit, double tol, or_fit *out)
{
    mle_binomial_any(X, y, n, k, maxit, tol, 0, out);
}

/* ------------------------------------------------------------------------------------------ */
static double pois_ll(const double *y, const double *eta, const double *f, long n)
{
    double s = 0.0; /* accu(b % eta - exp(eta) - f) */
    for (long i = 0; i < n; i++) s += y[i] * eta[i] - exp(eta[i]) - f[i];
    return s;
}

void or_mle_poisson(const double *X, const double *y, long n, int k, int maxit, double tol,
                    or_fit *out)
{
    memset(out, 0, sizeof(*out));
    double x[OR_MAXM] = {0}, xold[OR_MAXM], G[OR_MAXM * OR_MAXM], rhs[OR_MAXM], dir[OR_MAXM], xnew[OR_MAXM];
    double *eta = malloc(sizeof(double) * n), *W = malloc(sizeof(double) * n), *z = malloc(sizeof(double) * n),
           *f = malloc(sizeof(double) * n), *eta2 = malloc(sizeof(double) * n);
    double nobs = (double)n;
    int passes = 0, converged = 0, stopped_bad = 0;

    for (long j = 0; j < n; j++) f[j] = lgamma(y[j] + 1.0);   /* irls_poisson_fast.cpp:32-34 */

    /* :37-58 initial values */
    for (long j = 0; j < n; j++) {
        double mu = y[j] + 0.1;
        double eta_init = log(mu);
        W[j] = exp(eta_init);
        z[j] = eta_init + (y[j] - W[j]) / W[j];
    }
    passes++;
    if (all_finite(W, n) && all_finite(z, n)) {
        weighted_gram(X, n, k, W, z, G, rhs);
        if (all_finite(G, (long)k * k)) {
            or_arma_solve(G, rhs, k, 1, x);                   /* :52 */
            if (!all_finite(x, k)) memset(x, 0, sizeof(double) * k);
        }
    }

    int it;
    for (it = 0; it < maxit; ++it) {                          /* :61 */
        matvec(X, n, k, x, eta);
        for (long j = 0; j < n; j++) W[j] = exp(eta[j]);
        if (!all_finite(W, n)) { stopped_bad = 1; break; }    /* :66-69 */
        for (long j = 0; j < n; j++) z[j] = eta[j] + (y[j] - W[j]) / W[j];
        memcpy(xold, x, sizeof(double) * k);
        double current_ll = pois_ll(y, eta, f, n);            /* :74 */
        if (!isfinite(current_ll)) current_ll = -INFINITY;
        weighted_gram(X, n, k, W, z, G, rhs);                 /* :80 */
        for (int a = 0; a < k; a++) G[a * k + a] += 1e-8;     /* :81-82 */
        passes++;
        out->iters = it + 1;
        if (!all_finite(G, (long)k * k)) { stopped_bad = 1; break; } /* :85-88 */
        or_arma_solve(G, rhs, k, 1, dir);                     /* :91 */
        for (int a = 0; a < k; a++) dir[a] -= xold[a];
        double step = 1.0;
        int accepted = 0;
        for (int j = 0; j < 10; ++j) {                        /* :98 */
            for (int a = 0; a < k; a++) xnew[a] = xold[a] + step * dir[a];
            if (!all_finite(xnew, k)) { step /= 2.0; continue; }
            matvec(X, n, k, xnew, eta2);
            double new_ll = pois_ll(y, eta2, f, n);
            if (isfinite(new_ll) && new_ll >= current_ll) {   /* :111 */
                memcpy(x, xnew, sizeof(double) * k);
                accepted = 1;
                break;
            }
            passes++; /* a rejected half step costs a pass */
            step /= 2.0;
        }
        if (!accepted) memcpy(x, xold, sizeof(double) * k);   /* :122-125 */
        double s = 0.0;
        for (int a = 0; a < k; a++) s += (x[a] - xold[a]) * (x[a] - xold[a]);
        if (sqrt(s) < tol) { converged = 1; break; }          /* :129 */
    }
    matvec(X, n, k, x, eta);                                  /* :135 */
    double df = (double)k;
    out->loglik = pois_ll(y, eta, f, n);                      /* :137 */
    out->aic = -2 * out->loglik + 2 * df;
    out->bic = -2 * out->loglik + log(nobs) * df;
    passes++;
    out->passes = passes;
    out->status = !isfinite(out->loglik) ? ABN_FIT_NON_FINITE
                  : ((converged && !stopped_bad) ? ABN_FIT_OK : ABN_FIT_NOT_CONVERGED);
    memcpy(out->coef, x, sizeof(double) * k);
    for (long j = 0; j < n; j++) { W[j] = exp(eta[j]); z[j] = 0.0; }  /* :142 varmatrix = A' diag(exp(eta)) A */
    weighted_gram(X, n, k, W, z, G, rhs);
    se_from_information(G, k, 1.0, out->se);
    free(eta); free(W); free(z); free(f); free(eta2);
}

/* ------------------------------------------------------------------------------------------ */
/* nnet 7.3-20 (renv.lock) is not in the reference tree. This restates its published algorithm
 * for the configuration R/build_score_cache_mle.R:319-331 selects: a skip-layer-only softmax
 * network (size = 0) whose first output unit and all biases are masked, i.e. multinomial
 * logistic regression with reference level 1 on the columns of X; objective "value" =
 * - sum_i log p_{i, y_i}; optimiser = R's vmmin (BFGS, Nash's Algorithm 21) from zero
 * weights with maxit = 100, abstol = 1e-4, reltol = 1e-8. PARITY UNPINNED: the only reference
 * check is equality with nnet::multinom itself (tests/testthat/test-build_score_cache_mle.R:248). */
typedef struct {
    const double *X; const int *y; long n; int k, L;
    double *eta; /* scratch L */
} mn_ctx;

static double mn_value(const mn_ctx *c, const double *w, double *grad)
{
    /* w: (L-1) x k, row l-1 holds the weights of output l (l = 1..L-1), output 0 fixed at 0 */
    const int k = c->k, L = c->L;
    double total = 0.0;
    double out[64], p[64];
    if (grad) memset(grad, 0, sizeof(double) * (L - 1) * k);
    for (long i = 0; i < c->n; i++) {
        out[0] = 0.0;
        for (int l = 1; l < L; l++) {
            double s = 0.0;
            for (int a = 0; a < k; a++) s += w[(l - 1) * k + a] * c->X[(size_t)a * c->n + i];
            out[l] = s;
        }
        double mx = out[0];
        for (int l = 1; l < L; l++) if (out[l] > mx) mx = out[l];
        double sum = 0.0;
        for (int l = 0; l < L; l++) { p[l] = exp(out[l] - mx); sum += p[l]; }
        for (int l = 0; l < L; l++) p[l] /= sum;
        int yi = c->y[i];
        if (p[yi] > 0) total -= log(p[yi]); else total += 1000;
        if (grad) {
            for (int l = 1; l < L; l++) {
                double e = p[l] - (l == yi ? 1.0 : 0.0);
                for (int a = 0; a < k; a++) grad[(l - 1) * k + a] += e * c->X[(size_t)a * c->n + i];
            }
        }
    }
    return total;
}

void or_mle_multinomial(const double *X, const int *ycode, long n, int k, int L, int R, or_fit *out)
{
    memset(out, 0, sizeof(*out));
    const int np = (L - 1) * k;
    const int maxit = 100;
    const double abstol = 1.0e-4, reltol = 1.0e-8;
    const double stepredn = 0.2, acctol = 0.0001, reltest = 10.0;
    mn_ctx ctx = {X, ycode, n, k, L, NULL};
    double *b = calloc(np, sizeof(double)), *g = calloc(np, sizeof(double)), *t = calloc(np, sizeof(double)),
           *Xv = calloc(np, sizeof(double)), *c = calloc(np, sizeof(double)), *B = calloc((size_t)np * np, sizeof(double));
    int count = 0, funcount, gradcount, ilast, iter = 0, passes = 0;
    double f, Fmin, gradproj, s, steplength, D1, D2;
    int accpoint, enough;

    f = mn_value(&ctx, b, NULL); passes++;
    Fmin = f;
    funcount = gradcount = 1;
    mn_value(&ctx, b, g); passes++;
    iter++;
    ilast = gradcount;
    do {
        if (ilast == gradcount) {
            for (int i = 0; i < np; i++) {
                for (int j = 0; j < i; j++) B[i * np + j] = 0.0;
                B[i * np + i] = 1.0;
            }
        }
        for (int i = 0; i < np; i++) { Xv[i] = b[i]; c[i] = g[i]; }
        gradproj = 0.0;
        for (int i = 0; i < np; i++) {
            s = 0.0;
            for (int j = 0; j <= i; j++) s -= B[i * np + j] * g[j];
            for (int j = i + 1; j < np; j++) s -= B[j * np + i] * g[j];
            t[i] = s;
            gradproj += s * g[i];
        }
        if (gradproj < 0.0) { /* downhill */
            steplength = 1.0;
            accpoint = 0;
            do {
                count = 0;
                for (int i = 0; i < np; i++) {
                    b[i] = Xv[i] + steplength * t[i];
                    if (reltest + Xv[i] == reltest + b[i]) count++;
                }
                if (count < np) {
                    f = mn_value(&ctx, b, NULL); passes++;
                    funcount++;
                    accpoint = isfinite(f) && (f <= Fmin + gradproj * steplength * acctol);
                    if (!accpoint) steplength *= stepredn;
                }
            } while (!(count == np || accpoint));
            enough = (f > abstol) && fabs(f - Fmin) > reltol * (fabs(Fmin) + reltol);
            if (!enough) { count = np; Fmin = f; }
            if (count < np) {
                Fmin = f;
                mn_value(&ctx, b, g); passes++;
                gradcount++;
                iter++;
                D1 = 0.0;
                for (int i = 0; i < np; i++) {
                    t[i] = steplength * t[i];
                    c[i] = g[i] - c[i];
                    D1 += t[i] * c[i];
                }
                if (D1 > 0) {
                    D2 = 0.0;
                    for (int i = 0; i < np; i++) {
                        s = 0.0;
                        for (int j = 0; j <= i; j++) s += B[i * np + j] * c[j];
                        for (int j = i + 1; j < np; j++) s += B[j * np + i] * c[j];
                        Xv[i] = s;
                        D2 += s * c[i];
                    }
                    D2 = 1.0 + D2 / D1;
                    for (int i = 0; i < np; i++)
                        for (int j = 0; j <= i; j++)
                            B[i * np + j] += (D2 * t[i] * t[j] - Xv[i] * c[j] - t[i] * Xv[j]) / D1;
                } else {
                    ilast = gradcount;
                }
            } else {
                if (ilast < gradcount) { count = 0; ilast = gradcount; }
            }
        } else { /* uphill */
            count = 0;
            if (ilast == gradcount) count = np; else ilast = gradcount;
        }
        if (iter >= maxit) break;
        if (gradcount - ilast > 2 * np) ilast = gradcount;
    } while (count != np || ilast != gradcount);

    double value = Fmin;
    double edf = -1.0 * R; /* R/build_score_cache_mle.R:329: length(tmp$lev) is 0 for nnet.default */
    out->loglik = -value;                                  /* :328 */
    out->aic = 2 * value + 2 * edf;                        /* :330 */
    out->bic = 2 * value + edf * log((double)n);           /* :331 */
    out->iters = iter;
    out->passes = passes;
    out->status = !isfinite(value) ? ABN_FIT_NON_FINITE : (iter < maxit ? ABN_FIT_OK : ABN_FIT_NOT_CONVERGED);
    for (int i = 0; i < np && i < OR_MAXCOEF; i++) out->coef[i] = b[i];
    free(b); free(g); free(t); free(Xv); free(c); free(B);
}
