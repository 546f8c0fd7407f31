/*
 * abn_oracle.h - CPU oracle for the score-cache hot path of abn (furrer-lab/abn v3.1.13).
 *
 * TEST INFRASTRUCTURE.  This is a plain-C restatement of the reference's algorithm, one
 * function per reference function, each citing the reference file:line it follows.  Only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load it, and only as the checker or the timed CPU baseline; nothing under abn_b200/ may
 * import, link or call it.
 *
 * Parity pin (see oracle/README.md, tests/test_oracle_golden.py): the reference itself cannot
 * be built here (needs R, Rcpp, RcppArmadillo, GSL - none present), so the oracle is pinned
 * against the reference's own golden caches: tests/testthat/testdata/buildscorecache_ex1.Rdata
 * (2 560 Bayes rows, enforced by the reference's test-build_score_cache_bayes.R:126-135),
 * inst/extdata/mycache.c1/c3/c4.rds, and the Gaussian / separated-binomial rows of
 * tests/testthat/testdata/mycache.Rdata.  Parts with no golden (binomial / Poisson MLE exact
 * values, multinomial children, rank-deficient designs) are "parity unpinned" and listed as
 * such in DESIGN.md.
 */
#ifndef ABN_ORACLE_H
#define ABN_ORACLE_H

#include <stdint.h>

#include "../include/abn_b200.h" /* abn_dataset / abn_candidates / abn_control / abn_result */

#ifdef __cplusplus
extern "C" {
#endif

#define OR_MAXCOEF 40

/* ---- enumeration ------------------------------------------------------------------------ */

/* src/buildcachematrix.c:15-141 (gsl_combination order, ban / retain filter).
 * banned / retained d x d column-major [child + parent*d]; which_nodes 0-based.
 * Two-call pattern like the reference's two passes: children == NULL counts only.
 * children are 0-based; node_defn is row-major rows x d (the layout of res[[2]] before
 * matrix(byrow=TRUE), R/build_score_cache_bayes.R:82). */
int64_t or_enumerate_bayes(int d, const int *banned, const int *retained, const int *max_parents,
                           const int *which_nodes, int n_which, int *children, int *node_defn);

/* R/build_score_cache_mle.R:432-518 (combn over d-1 items with a 0 inserted at the child's
 * position, then retain / ban row deletion). Same outputs as above. */
int64_t or_enumerate_mle(int d, const int *banned, const int *retained, int max_parents,
                         int *children, int *node_defn);

/* ---- single fits -------------------------------------------------------------------------- */

typedef struct or_fit {
    double loglik; /* mle: log-likelihood; bayes: log marginal likelihood                     */
    double aic, bic;
    int iters;     /* solver iterations executed                                               */
    int passes;    /* data passes a fused implementation needs (SURVEY 8d accounting)          */
    int status;    /* ABN_FIT_*                                                                */
    int rank;      /* mle: rank of the design as rank_cpp sees it                              */
    double coef[OR_MAXCOEF];
    double se[OR_MAXCOEF]; /* mle: the standard errors fitAbn derives from the fit's `varcov`, R/fitabn_mle.R:715-745 */
} or_fit;

/* MLE: X column-major n x k INCLUDING the intercept column when the model has one. */
void or_mle_gaussian(const double *X, const double *y, long n, int k, or_fit *out);   /* src/irls_gaussian_fast.cpp:12-99 */
void or_mle_binomial_br(const double *X, const double *y, long n, int k, int maxit, double tol,
                        or_fit *out);                                                 /* src/irls_binomial_fast_br.cpp:15-153 */
void or_mle_binomial(const double *X, const double *y, long n, int k, int maxit, double tol,
                     or_fit *out);                                                    /* src/irls_binomial_fast.cpp:14-150 */
void or_mle_poisson(const double *X, const double *y, long n, int k, int maxit, double tol,
                    or_fit *out);                                                     /* src/irls_poisson_fast.cpp:16-160 */
/* nnet.default(x=X, y=onehot, mask, size=0, skip=TRUE, softmax=TRUE, rang=0) + BFGS (vmmin),
 * R/build_score_cache_mle.R:319-331; ycode[i] in 0..L-1; R = rank(X) enters edf */
void or_mle_multinomial(const double *X, const int *ycode, long n, int k, int L, int R,
                        or_fit *out);
/* test knob: row order of every sum over the observations (0 ascending, 1 descending), see mle.c */
void or_set_sum_order(int order);
extern int or_sum_order_;
#define OR_ROWS(i, n) for (long c_ = 0, i = (or_sum_order_ ? (long)(n) - 1 : 0); c_ < (long)(n); c_++, i += (or_sum_order_ ? -1 : 1))
int or_rank_cpp(const double *X, long n, int k);                                      /* src/rank.cpp:15 */

/* Bayes Laplace: X column-major n x m INCLUDING the intercept (build_designmatrix*).
 * solver: ABN_SOLVER_NEWTON or ABN_SOLVER_HYBRID. */
void or_bayes_binomial(const double *X, const double *y, long n, int m, double prior_mean,
                       double prior_sd, int max_iters, double epsabs, int solver, or_fit *out); /* src/node_binomial.c:18-227 */
void or_bayes_gaussian(const double *X, const double *y, long n, int m, double prior_mean,
                       double prior_sd, double gam_shape, double gam_scale, int max_iters,
                       double epsabs, int solver, or_fit *out);                       /* src/node_gaussian.c:18-216 */
void or_bayes_poisson(const double *X, const double *y, long n, int m, double prior_mean,
                      double prior_sd, int max_iters, double epsabs, int solver, or_fit *out); /* src/node_poisson.c:20-226 */

/* ---- whole cache -------------------------------------------------------------------------- */

/* The per-candidate loop of R/build_score_cache_mle.R:615-703 (forLoopContent :251-352) or
 * R/build_score_cache_bayes.R:121-214 (fit_single_node, src/fit_single_node.c:25-177): for
 * every candidate assemble a fresh design copy, rank-check (mle), fit, score.  OpenMP
 * parallel-for over candidates = the reference's foreach %dopar% workers.
 * threads <= 0 uses omp_get_max_threads(). Returns 0 or ABN_ERR_*. */
int or_score_cache(const abn_dataset *ds, const abn_candidates *cand, const abn_control *ctrl,
                   abn_result *out, int threads);

int or_num_threads(void);

#ifdef __cplusplus
}
#endif
#endif
