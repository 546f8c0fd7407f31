/* TEST INFRASTRUCTURE - part of the CPU oracle (see abn_oracle.h).
 * Laplace-approximated log marginal likelihood of one GLM node, method = "bayes". */
#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "abn_oracle.h"
#include "linalg.h"

typedef struct lap {
    const double *X, *y;
    long n;
    int m;          /* regression coefficients incl. intercept */
    int M;          /* unknowns: m (binomial, Poisson) or m + 1 (Gaussian: + precision) */
    int family;
    double pmean, psd, gshape, gscale;
    double *eta, *wrk; /* n-long scratch */
    int passes;        /* data passes (each df / fdf / g evaluation streams the rows once) */
} lap;

static void eta_of(const lap *L, const double *beta)
{
    OR_ROWS(i, L->n) L->eta[i] = 0.0;
    for (int a = 0; a < L->m; a++) {
        const double *xa = L->X + (size_t)a * L->n;
        double c = beta[a];
        OR_ROWS(i, L->n) L->eta[i] += xa[i] * c;
    }
}

static double prior_terms(const lap *L, const double *beta)
{
    /* term2 + term3 of laplace_g (node_binomial.c:607-624): log N(beta_j; mean, sd) */
    double t2 = 0.0, t3 = 0.0;
    for (int j = 0; j < L->m; j++) {
        t2 += -log(sqrt(2.0 * M_PI) * L->psd);
        double dlt = beta[j] - L->pmean;
        t3 += -0.5 * (dlt * dlt) / (L->psd * L->psd);
    }
    return t2 + t3;
}

/* ---- g ------------------------------------------------------------------------------------ */
static double lap_g(lap *L, const double *th)
{
    const long n = L->n;
    const double dn = (double)n;
    L->passes++;
    eta_of(L, th);
    if (L->family == ABN_BINOMIAL) {                      /* node_binomial.c:585-649 */
        double t1 = 0.0, s = 0.0;
        OR_ROWS(i, n) t1 += L->y[i] * L->eta[i];
        OR_ROWS(i, n) {
            double big = exp(L->eta[i]);
            s += (big != INFINITY) ? -log(1.0 + big) : -L->eta[i]; /* :633-639 */
        }
        return (-1.0 / dn) * (t1 + s + prior_terms(L, th));
    }
    if (L->family == ABN_POISSON) {                       /* node_poisson.c:517-571 */
        double t1 = 0.0, s = 0.0;
        OR_ROWS(i, n) t1 += L->y[i] * L->eta[i];
        OR_ROWS(i, n) s += -lgamma(L->y[i] + 1.0) - exp(L->eta[i]); /* gsl_sf_lnfact :561 */
        return (-1.0 / dn) * (t1 + s + prior_terms(L, th));
    }
    /* Gaussian, node_gaussian.c:503-579 */
    double tau = th[L->m];
    double yxb = 0.0, xbxb = 0.0, yy = 0.0;
    OR_ROWS(i, n) {
        yxb += L->y[i] * L->eta[i];
        xbxb += L->eta[i] * L->eta[i];
        yy += L->y[i] * L->y[i];
    }
    double term4 = -(tau / 2.0) * (-2.0 * yxb + xbxb + yy);                     /* :559-565 */
    double term5 = (dn / 2.0) * log(tau / (2.0 * M_PI));                        /* :568 */
    double term6 = -L->gshape * log(L->gscale) - lgamma(L->gshape)
                   + (L->gshape - 1) * log(tau) - (tau / L->gscale);            /* :570-573 */
    return (-1.0 / dn) * (prior_terms(L, th) + term4 + term5 + term6);
}

/* ---- grad g ------------------------------------------------------------------------------- */
static void lap_dg(lap *L, const double *th, double *dg)
{
    const long n = L->n;
    const int m = L->m;
    const double dn = (double)n;
    L->passes++;
    eta_of(L, th);
    double xty[OR_MAXM], xtr[OR_MAXM];
    for (int a = 0; a < m; a++) {
        const double *xa = L->X + (size_t)a * n;
        double s = 0.0;
        OR_ROWS(i, n) s += xa[i] * L->y[i];
        xty[a] = s;
    }
    if (L->family == ABN_BINOMIAL) {                      /* node_binomial.c:655-719 */
        OR_ROWS(i, n) {
            double big = exp(L->eta[i]);
            L->wrk[i] = (big != INFINITY) ? -big / (1 + big) : -1.0; /* :697-700 */
        }
    } else if (L->family == ABN_POISSON) {                /* node_poisson.c:576-630 */
        OR_ROWS(i, n) L->wrk[i] = -exp(L->eta[i]);
    } else {                                              /* node_gaussian.c:586-654 */
        OR_ROWS(i, n) L->wrk[i] = -L->eta[i];
    }
    for (int a = 0; a < m; a++) {
        const double *xa = L->X + (size_t)a * n;
        double s = 0.0;
        OR_ROWS(i, n) s += xa[i] * L->wrk[i];
        xtr[a] = s;
    }
    if (L->family == ABN_GAUSSIAN) {
        double tau = th[m];
        for (int a = 0; a < m; a++) {
            double prior = -(th[a] - L->pmean) / (L->psd * L->psd);
            dg[a] = (-1.0 / dn) * (prior + tau * (xty[a] + xtr[a]));             /* :625-637 */
        }
        double rss = 0.0;
        OR_ROWS(i, n) { double e = L->y[i] - L->eta[i]; rss += e * e; }
        dg[m] = (-1.0 / dn) * ((dn / (2.0 * tau)) + rss * (-0.5) + (L->gshape - 1.0) / tau
                               - (1.0 / L->gscale));                             /* :645-651 */
    } else {
        for (int a = 0; a < m; a++) {
            double prior = -(th[a] - L->pmean) / (L->psd * L->psd);
            dg[a] = (-1.0 / dn) * (prior + xtr[a] + xty[a]);
        }
    }
}

/* ---- hess g ------------------------------------------------------------------------------- */
static void lap_hess(lap *L, const double *th, double *H)
{
    const long n = L->n;
    const int m = L->m, M = L->M;
    const double dn = (double)n;
    L->passes++;
    eta_of(L, th);
    if (L->family == ABN_GAUSSIAN) {                      /* node_gaussian.c:660-738 */
        double tau = th[m];
        for (int j = 0; j < m; j++) {
            for (int k = 0; k < m; k++) {
                const double *xj = L->X + (size_t)j * n, *xk = L->X + (size_t)k * n;
                double s = 0.0;
                OR_ROWS(i, n) s += xj[i] * xk[i];
                H[j * M + k] = (j != k) ? (-1.0 / dn) * (-tau * s)
                                        : (-1.0 / dn) * (-tau * s - 1.0 / (L->psd * L->psd)); /* :700-709 */
            }
        }
        H[m * M + m] = (-1.0 / dn) * ((-dn / (2.0 * tau * tau)) - ((L->gshape - 1) / (tau * tau))); /* :713 */
        for (int j = 0; j < m; j++) {
            const double *xj = L->X + (size_t)j * n;
            double s = 0.0;
            OR_ROWS(i, n) s += xj[i] * (L->y[i] - L->eta[i]);
            H[m * M + j] = H[j * M + m] = (-1.0 / dn) * s;                        /* :718-735 */
        }
        return;
    }
    if (L->family == ABN_BINOMIAL) {                      /* node_binomial.c:759-770 */
        OR_ROWS(i, n) {
            double t1 = L->eta[i];
            double t2 = -2.0 * log(1 + exp(L->eta[i])) - log(dn);
            double v = exp(t1 + t2);
            L->wrk[i] = isnan(v) ? 0.0 : v;
        }
    } else {                                              /* node_poisson.c:665-667 */
        OR_ROWS(i, n) L->wrk[i] = exp(L->eta[i]) / dn;
    }
    for (int j = 0; j < m; j++) {
        for (int k = 0; k < m; k++) {
            const double *xj = L->X + (size_t)j * n, *xk = L->X + (size_t)k * n;
            double s = 0.0;
            OR_ROWS(i, n) s += (xj[i] * xk[i]) * L->wrk[i];
            if (j == k) s += 1.0 / (dn * L->psd * L->psd);  /* node_binomial.c:781-789 */
            H[j * M + k] = s;
        }
    }
}

/* ---- initial values ----------------------------------------------------------------------- */
static void lap_inits(lap *L, double *th)
{
    /* generate_inits_n (node_binomial.c:832-883), generate_inits_n_pois (node_poisson.c:705-763),
     * generate_gaus_inits (node_gaussian.c:743-808): least squares through inv(X'X) by LU */
    const long n = L->n;
    const int m = L->m;
    double XtX[OR_MAXM * OR_MAXM], inv[OR_MAXM * OR_MAXM], xty[OR_MAXM], b[OR_MAXM];
    int perm[OR_MAXM], sg;
    for (int j = 0; j < m; j++)
        for (int k = 0; k < m; k++) {
            const double *xj = L->X + (size_t)j * n, *xk = L->X + (size_t)k * n;
            double s = 0.0;
            OR_ROWS(i, n) s += xj[i] * xk[i];
            XtX[j * m + k] = s;
        }
    or_lu_decomp(XtX, m, perm, &sg);
    int err = or_lu_invert(XtX, perm, m, inv);
    for (int j = 0; j < m; j++) b[j] = 0.0;
    if (!err) {
        for (int j = 0; j < m; j++) {
            const double *xj = L->X + (size_t)j * n;
            double s = 0.0;
            if (L->family == ABN_POISSON)
                OR_ROWS(i, n) s += xj[i] * log(L->y[i] + 0.1);   /* node_poisson.c:743 */
            else
                OR_ROWS(i, n) s += xj[i] * L->y[i];
            xty[j] = s;
        }
        for (int j = 0; j < m; j++) {
            double s = 0.0;
            for (int k = 0; k < m; k++) s += inv[j * m + k] * xty[k];
            b[j] = s;
        }
        for (int j = 0; j < m; j++) th[j] = b[j];
    } else {
        for (int j = 0; j < m; j++) th[j] = (L->family == ABN_GAUSSIAN) ? 0.01 : 0.0; /* node_gaussian.c:790 */
    }
    if (L->family == ABN_GAUSSIAN) {
        /* node_gaussian.c:795-805. In the singular case the reference multiplies X by a stale
         * scratch vector; we use the coefficients just set (unpinned corner). */
        eta_of(L, err ? th : b);
        double rss = 0.0;
        OR_ROWS(i, n) { double e = L->y[i] - L->eta[i]; rss += e * e; }
        th[m] = 1.0 / (rss / ((double)n - (double)m));
    }
    L->passes++;
}

/* ---- Powell hybrid, as gsl_multiroot_fdfsolver_hybridsj (scale = 1) / hybridj (scale = 0) --- */
/* GSL is an unpinned system library that is not in the reference tree (configure.ac:6-14).
 * This restates the published algorithm (MINPACK HYBRJ as implemented in GSL's
 * multiroots/hybridj.c and dogleg.c). */
static double enorm(const double *f, int n) { double s = 0; for (int i = 0; i < n; i++) s += f[i] * f[i]; return sqrt(s); }
static double scaled_enorm(const double *d, const double *f, int n) { double s = 0; for (int i = 0; i < n; i++) { double u = d[i] * f[i]; s += u * u; } return sqrt(s); }

typedef struct hyb {
    int n, scale, iter, ncfail, ncsuc, nslow1, nslow2;
    double fnorm, delta;
    double x[OR_MAXM], f[OR_MAXM], dx[OR_MAXM];
    double J[OR_MAXM * OR_MAXM], q[OR_MAXM * OR_MAXM], r[OR_MAXM * OR_MAXM];
    double diag[OR_MAXM];
} hyb;

static void compute_diag(const double *J, int n, double *diag)
{
    for (int j = 0; j < n; j++) {
        double sum = 0;
        for (int i = 0; i < n; i++) sum += J[i * n + j] * J[i * n + j];
        if (sum == 0) sum = 1.0;
        diag[j] = sqrt(sum);
    }
}
static void update_diag(const double *J, int n, double *diag)
{
    for (int j = 0; j < n; j++) {
        double sum = 0;
        for (int i = 0; i < n; i++) sum += J[i * n + j] * J[i * n + j];
        if (sum == 0) sum = 1.0;
        double cnorm = sqrt(sum);
        if (cnorm > diag[j]) diag[j] = cnorm;
    }
}
static double compute_delta(const double *diag, const double *x, int n)
{
    double Dx = scaled_enorm(diag, x, n);
    double factor = 100;
    return (Dx > 0) ? factor * Dx : factor;
}

static void dogleg(const hyb *s, const double *qtf, double *p)
{
    const int n = s->n;
    const double *r = s->r, *diag = s->diag;
    double newton[OR_MAXM], grad[OR_MAXM];
    /* newton_direction: solve R p = qtf, negate (gsl_linalg_R_solve: plain back substitution) */
    for (int i = n - 1; i >= 0; i--) {
        double t = qtf[i];
        for (int j = i + 1; j < n; j++) t -= r[i * n + j] * newton[j];
        newton[i] = t / r[i * n + i];
    }
    for (int i = 0; i < n; i++) newton[i] = -newton[i];
    double qnorm = scaled_enorm(diag, newton, n);
    if (qnorm <= s->delta) { memcpy(p, newton, sizeof(double) * n); return; }
    /* gradient_direction */
    for (int j = 0; j < n; j++) {
        double sum = 0;
        for (int i = 0; i < n; i++) sum += r[i * n + j] * qtf[i];
        grad[j] = -sum / diag[j];
    }
    double gnorm = enorm(grad, n);
    if (gnorm == 0) {
        double alpha = s->delta / qnorm;
        for (int i = 0; i < n; i++) p[i] = alpha * newton[i];
        return;
    }
    for (int i = 0; i < n; i++) grad[i] = (grad[i] / gnorm) / diag[i]; /* minimum_step */
    for (int i = 0; i < n; i++) {                                     /* compute_Rg */
        double sum = 0;
        for (int j = i; j < n; j++) sum += r[i * n + j] * grad[j];
        p[i] = sum;
    }
    double temp = enorm(p, n);
    double sgnorm = (gnorm / temp) / temp;
    if (sgnorm > s->delta) {
        for (int i = 0; i < n; i++) p[i] = s->delta * grad[i];
        return;
    }
    double bnorm = enorm(qtf, n);
    double bg = bnorm / gnorm, bq = bnorm / qnorm, dq = s->delta / qnorm, dq2 = dq * dq;
    double sd = sgnorm / s->delta, sd2 = sd * sd;
    double t1 = bg * bq * sd, u = t1 - dq;
    double t2 = t1 - dq * sd2 + sqrt(u * u + (1 - dq2) * (1 - sd2));
    double alpha = dq * (1 - sd2) / t2, beta = (1 - alpha) * sgnorm;
    for (int i = 0; i < n; i++) p[i] = alpha * newton[i] + beta * grad[i];
}

static int all_fin(const double *v, int n) { for (int i = 0; i < n; i++) if (!isfinite(v[i])) return 0; return 1; }

static void hyb_set(hyb *s, lap *L, const double *x0, int scale)
{
    const int n = s->n = L->M;
    s->scale = scale;
    memcpy(s->x, x0, sizeof(double) * n);
    lap_dg(L, s->x, s->f);         /* wrapper_fdf: two evaluations in the reference, */
    lap_hess(L, s->x, s->J);       /* one fused data pass on the GPU                   */
    L->passes--;
    s->iter = 1;
    s->fnorm = enorm(s->f, n);
    s->ncfail = s->ncsuc = s->nslow1 = s->nslow2 = 0;
    memset(s->dx, 0, sizeof(double) * n);
    if (scale) compute_diag(s->J, n, s->diag); else for (int i = 0; i < n; i++) s->diag[i] = 1.0;
    s->delta = compute_delta(s->diag, s->x, n);
    or_qr(s->J, n, s->q, s->r);
}

enum { HY_SUCCESS = 0, HY_EBADFUNC = 9, HY_ENOPROG = 27, HY_ENOPROGJ = 28 };

static int hyb_iterate(hyb *s, lap *L)
{
    const int n = s->n;
    const double fnorm = s->fnorm;
    const double p1 = 0.1, p5 = 0.5, p001 = 0.001, p0001 = 0.0001;
    double qtf[OR_MAXM], x_trial[OR_MAXM], f_trial[OR_MAXM], df[OR_MAXM], rdx[OR_MAXM], qtdf[OR_MAXM], w[OR_MAXM], v[OR_MAXM];

    for (int j = 0; j < n; j++) { double sum = 0; for (int i = 0; i < n; i++) sum += s->q[i * n + j] * s->f[i]; qtf[j] = sum; }
    dogleg(s, qtf, s->dx);
    for (int i = 0; i < n; i++) x_trial[i] = s->x[i] + s->dx[i];
    double pnorm = scaled_enorm(s->diag, s->dx, n);
    if (s->iter == 1 && pnorm < s->delta) s->delta = pnorm;

    lap_dg(L, x_trial, f_trial);
    /* the reference's callbacks always return GSL_SUCCESS, so GSL_EBADFUNC never fires */
    for (int i = 0; i < n; i++) df[i] = f_trial[i] - s->f[i];
    double fnorm1 = enorm(f_trial, n);
    double actred = (fnorm1 < fnorm) ? 1 - (fnorm1 / fnorm) * (fnorm1 / fnorm) : -1;
    for (int i = 0; i < n; i++) { double sum = 0; for (int j = i; j < n; j++) sum += s->r[i * n + j] * s->dx[j]; rdx[i] = sum; }
    double fnorm1p = 0; for (int i = 0; i < n; i++) { double u = qtf[i] + rdx[i]; fnorm1p += u * u; } fnorm1p = sqrt(fnorm1p);
    double prered = (fnorm1p < fnorm) ? 1 - (fnorm1p / fnorm) * (fnorm1p / fnorm) : 0;
    double ratio = (prered > 0) ? actred / prered : 0;

    if (ratio < p1) { s->ncsuc = 0; s->ncfail++; s->delta *= p5; }
    else {
        s->ncfail = 0; s->ncsuc++;
        if (ratio >= p5 || s->ncsuc > 1) s->delta = fmax(s->delta, pnorm / p5);
        if (fabs(ratio - 1) <= p1) s->delta = pnorm / p5;
    }
    if (ratio >= p0001) {
        memcpy(s->x, x_trial, sizeof(double) * n);
        memcpy(s->f, f_trial, sizeof(double) * n);
        s->fnorm = fnorm1;
        s->iter++;
    }
    s->nslow1++;
    if (actred >= p001) s->nslow1 = 0;
    if (actred >= p1) s->nslow2 = 0;

    if (s->ncfail == 2) {
        lap_hess(L, s->x, s->J);
        s->nslow2++;
        if (s->iter == 1) {
            if (s->scale) compute_diag(s->J, n, s->diag);
            s->delta = compute_delta(s->diag, s->x, n);
        } else {
            if (s->scale) update_diag(s->J, n, s->diag);
        }
        or_qr(s->J, n, s->q, s->r);
        return HY_SUCCESS;
    }
    for (int j = 0; j < n; j++) { double sum = 0; for (int i = 0; i < n; i++) sum += s->q[i * n + j] * df[i]; qtdf[j] = sum; }
    for (int i = 0; i < n; i++) {
        w[i] = (qtdf[i] - rdx[i]) / pnorm;
        v[i] = s->diag[i] * s->diag[i] * s->dx[i] / pnorm;
    }
    or_qr_update(s->q, s->r, n, w, v);
    if (s->nslow2 == 5) return HY_ENOPROGJ;
    if (s->nslow1 == 10) return HY_ENOPROG;
    return HY_SUCCESS;
}

/* one solver attempt: the do { iterate; test_residual } while loop of node_binomial.c:100-115.
 * Returns 0 = GSL_SUCCESS, non-zero otherwise. */
static int hyb_solve(lap *L, double *th, int scale, int max_iters, double epsabs, int *iters)
{
    hyb s;
    lap_inits(L, th);
    hyb_set(&s, L, th, scale);
    int iter = 0, status;
    do {
        iter++;
        status = hyb_iterate(&s, L);
        if (status) break;
        double res = 0; for (int i = 0; i < s.n; i++) res += fabs(s.f[i]);
        status = (res < epsabs) ? 0 : -2; /* GSL_CONTINUE */
        if (isnan(res)) status = -2;
    } while (status == -2 && iter < max_iters);
    memcpy(th, s.x, sizeof(double) * s.n);
    *iters += iter;
    return status;
}

/* full Newton on grad g = 0 with the analytic Hessian, stop rule of gsl_multiroot_test_residual */
static int newton_solve(lap *L, double *th, int max_iters, double epsabs, int *iters)
{
    const int M = L->M;
    double dg[OR_MAXM], H[OR_MAXM * OR_MAXM], step[OR_MAXM];
    int perm[OR_MAXM], sg;
    lap_inits(L, th);
    for (int it = 0; it <= max_iters; it++) {
        lap_dg(L, th, dg);
        double res = 0; for (int i = 0; i < M; i++) res += fabs(dg[i]);
        if (res < epsabs) return 0;
        if (it == max_iters || !all_fin(dg, M)) break;
        lap_hess(L, th, H);
        L->passes--; /* fused with the gradient pass on the GPU */
        or_lu_decomp(H, M, perm, &sg);
        if (or_lu_singular(H, M)) break;
        or_lu_solve(H, perm, M, dg, step);
        for (int i = 0; i < M; i++) th[i] -= step[i];
        (*iters)++;
    }
    return 1;
}

static void lap_score(lap *L, int max_iters, double epsabs, int solver, or_fit *out)
{
    double th[OR_MAXM];
    int status, iters = 0;
    memset(out, 0, sizeof(*out));
    if (solver == ABN_SOLVER_HYBRID) {
        /* binomial / Gaussian: hybridsj then hybridj (node_binomial.c:90,127; node_gaussian.c:95,126);
         * Poisson: hybridj then hybridsj, each tried with two (identical) inits
         * (node_poisson.c:96-99,132-135) */
        int first = (L->family == ABN_POISSON) ? 0 : 1;
        int tries = (L->family == ABN_POISSON) ? 2 : 1;
        status = 1;
        for (int t = 0; t < tries && status; t++) status = hyb_solve(L, th, first, max_iters, epsabs, &iters);
        for (int t = 0; t < tries && status; t++) status = hyb_solve(L, th, !first, max_iters, epsabs, &iters);
    } else {
        status = newton_solve(L, th, max_iters, epsabs, &iters);
    }
    /* node_binomial.c:172-193 */
    double H[OR_MAXM * OR_MAXM];
    int perm[OR_MAXM], sg;
    double gvalue = lap_g(L, th);
    lap_hess(L, th, H);
    L->passes--; /* g and hess share the final pass on the GPU */
    or_lu_decomp(H, L->M, perm, &sg);
    double mydet = or_lu_lndet(H, L->M);
    double dn = (double)L->n, dm = (double)L->M;
    double logscore = -dn * gvalue - 0.5 * mydet + (dm / 2.0) * log((2.0 * M_PI) / dn);
    out->status = status ? ABN_FIT_NOT_CONVERGED : ABN_FIT_OK;
    if (isnan(logscore)) { logscore = NAN; out->status = ABN_FIT_NON_FINITE; }
    out->loglik = logscore;
    out->aic = out->bic = NAN;
    out->iters = iters;
    out->passes = L->passes;
    memcpy(out->coef, th, sizeof(double) * L->M);
}

static void run(const double *X, const double *y, long n, int m, int family, double pmean, double psd,
                double gshape, double gscale, int max_iters, double epsabs, int solver, or_fit *out)
{
    lap L = {X, y, n, m, family == ABN_GAUSSIAN ? m + 1 : m, family, pmean, psd, gshape, gscale, NULL, NULL, 0};
    L.eta = malloc(sizeof(double) * (n > 0 ? n : 1));
    L.wrk = malloc(sizeof(double) * (n > 0 ? n : 1));
    lap_score(&L, max_iters, epsabs, solver, out);
    free(L.eta);
    free(L.wrk);
}

void or_bayes_binomial(const double *X, const double *y, long n, int m, double prior_mean, double prior_sd,
                       int max_iters, double epsabs, int solver, or_fit *out)
{
    run(X, y, n, m, ABN_BINOMIAL, prior_mean, prior_sd, 0, 0, max_iters, epsabs, solver, out);
}

void or_bayes_gaussian(const double *X, const double *y, long n, int m, double prior_mean, double prior_sd,
                       double gam_shape, double gam_scale, int max_iters, double epsabs, int solver, or_fit *out)
{
    run(X, y, n, m, ABN_GAUSSIAN, prior_mean, prior_sd, gam_shape, gam_scale, max_iters, epsabs, solver, out);
}

void or_bayes_poisson(const double *X, const double *y, long n, int m, double prior_mean, double prior_sd,
                      int max_iters, double epsabs, int solver, or_fit *out)
{
    run(X, y, n, m, ABN_POISSON, prior_mean, prior_sd, 0, 0, max_iters, epsabs, solver, out);
}
