/* TEST INFRASTRUCTURE - part of the CPU oracle (see abn_oracle.h).
 * Whole-cache driver: the per-candidate loops of the R drivers. */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "abn_oracle.h"
#include "linalg.h"

int or_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

static void set_result(abn_result *out, int64_t r, double mlik, double aic, double bic, double mdl,
                       int status, int iters)
{
    if (out->mlik) out->mlik[r] = mlik;
    if (out->aic) out->aic[r] = aic;
    if (out->bic) out->bic[r] = bic;
    if (out->mdl) out->mdl[r] = mdl;
    if (out->status) out->status[r] = status;
    if (out->iters) out->iters[r] = iters;
}

/* forLoopContent, non-grouped branch: R/build_score_cache_mle.R:251-343 */
static void one_mle(const abn_dataset *ds, const abn_candidates *cand, const abn_control *ctrl,
                    abn_result *out, int64_t r)
{
    const long n = (long)ds->n;
    const int child = cand->child[r];
    const int fam = ds->node_family[child];
    const int32_t *par = cand->parents + r * cand->max_parents;
    int any_multi = 0, kdata = 0;
    for (int j = 0; j < cand->max_parents; j++) {
        if (par[j] < 0) continue;
        if (ds->node_family[par[j]] == ABN_MULTINOMIAL) any_multi = 1;        /* :266 */
        kdata += ds->node_col0[par[j] + 1] - ds->node_col0[par[j]];
    }
    int k = kdata + (any_multi ? 0 : 1);
    /* fresh design copy per candidate, like cbind(1, data.df.multi[, sel]) :267-269 */
    double *X = malloc(sizeof(double) * (size_t)n * (k > 0 ? k : 1));
    int c = 0;
    if (!any_multi) { for (long i = 0; i < n; i++) X[i] = 1.0; c = 1; }
    for (int j = 0; j < cand->max_parents; j++) {
        if (par[j] < 0) continue;
        for (int col = ds->node_col0[par[j]]; col < ds->node_col0[par[j] + 1]; col++)
            memcpy(X + (size_t)(c++) * n, ds->X + (size_t)col * n, sizeof(double) * n);
    }
    const double *y = ds->X + (size_t)ds->node_col0[child] * n;
    or_fit fit;
    memset(&fit, 0, sizeof(fit));
    int num_na = 0;
    int R = or_rank_cpp(X, n, k);                                              /* :282 */
    const double *Xuse = X;
    int kuse = k;
    if (fam == ABN_BINOMIAL) {
        if (R != k) {                                                          /* :286-294 */
            while (kuse > 0 && or_rank_cpp(Xuse, n, kuse) != kuse) { Xuse += n; kuse--; num_na++; }
        }
        or_mle_binomial_br(Xuse, y, n, kuse, ctrl->max_iters, ctrl->tol, &fit); /* :296,:311 */
    } else if (fam == ABN_GAUSSIAN) {
        or_mle_gaussian(X, y, n, k, &fit);                                     /* :314 */
    } else if (fam == ABN_POISSON) {
        or_mle_poisson(X, y, n, k, ctrl->max_iters, ctrl->tol, &fit);          /* :316 */
    } else {
        const int L = ds->node_col0[child + 1] - ds->node_col0[child];
        int *yc = malloc(sizeof(int) * n);
        for (long i = 0; i < n; i++) {
            yc[i] = 0;
            for (int l = 0; l < L; l++) if (y[(size_t)l * n + i] != 0.0) { yc[i] = l; break; }
        }
        or_mle_multinomial(X, yc, n, k, L, R, &fit);                           /* :319-331 */
        free(yc);
    }
    double mdl = fit.bic + (1 + kdata - num_na) * log((double)ctrl->nvars);    /* :343 */
    set_result(out, r, fit.loglik, fit.aic, fit.bic, mdl, fit.status, fit.passes);
    free(X);
}

/* forLoopContentBayes GLM branch -> fit_single_node: R/build_score_cache_bayes.R:303-335,
 * src/fit_single_node.c:104-165, build_designmatrix* (src/node_binomial.c:497-572) */
static void one_bayes(const abn_dataset *ds, const abn_candidates *cand, const abn_control *ctrl,
                      abn_result *out, int64_t r)
{
    const long n = (long)ds->n;
    const int child = cand->child[r];
    const int fam = ds->node_family[child];
    const int32_t *par = cand->parents + r * cand->max_parents;
    int m = 1;
    for (int j = 0; j < cand->max_parents; j++) if (par[j] >= 0) m++;
    double *X = malloc(sizeof(double) * (size_t)n * m);
    for (long i = 0; i < n; i++) X[i] = 1.0;
    int c = 1;
    for (int j = 0; j < cand->max_parents; j++) {
        if (par[j] < 0) continue;
        memcpy(X + (size_t)(c++) * n, ds->X + (size_t)ds->node_col0[par[j]] * n, sizeof(double) * n);
    }
    const double *y = ds->X + (size_t)ds->node_col0[child] * n;
    or_fit fit;
    if (fam == ABN_BINOMIAL)
        or_bayes_binomial(X, y, n, m, ctrl->prior_mean, ctrl->prior_sd, ctrl->max_iters, ctrl->epsabs, ctrl->solver, &fit);
    else if (fam == ABN_GAUSSIAN)
        or_bayes_gaussian(X, y, n, m, ctrl->prior_mean, ctrl->prior_sd, ctrl->gam_shape, ctrl->gam_scale,
                          ctrl->max_iters, ctrl->epsabs, ctrl->solver, &fit);
    else
        or_bayes_poisson(X, y, n, m, ctrl->prior_mean, ctrl->prior_sd, ctrl->max_iters, ctrl->epsabs, ctrl->solver, &fit);
    set_result(out, r, fit.loglik, NAN, NAN, NAN, fit.status, fit.passes);
    free(X);
}

int or_score_cache(const abn_dataset *ds, const abn_candidates *cand, const abn_control *ctrl,
                   abn_result *out, int threads)
{
    if (!ds || !cand || !ctrl || !out) return ABN_ERR_ARG;
    if (ctrl->method == ABN_METHOD_BAYES) {
        for (int i = 0; i < ds->n_nodes; i++)
            if (ds->node_family[i] == ABN_MULTINOMIAL) return ABN_ERR_ARG; /* _bayes.R:30-32 */
    }
#ifdef _OPENMP
    if (threads <= 0) threads = omp_get_max_threads();
#else
    threads = 1;
#endif
    const int64_t rows = cand->rows;
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads)
    for (int64_t r = 0; r < rows; r++) {
        if (ctrl->method == ABN_METHOD_MLE) one_mle(ds, cand, ctrl, out, r);
        else one_bayes(ds, cand, ctrl, out, r);
    }
    return ABN_OK;
}
