/* The R .Call shim of INTEGRATION.md section 2 as a file: drop it into the reference's src/ next to
 * include/abn_b200.h. There is no R in this image, so the tests compile it against a small stand-in for R's C
 * API (tests/r_stub/) and call both entry points the way .Call would (tests/test_r_shim.py). */

/* src/abn_cuda_shim.c - .Call entry points over libabn_b200.so (include/abn_b200.h) */
#include <R.h>
#include <Rinternals.h>
#include "abn_b200.h"

/* .Call("abn_cuda_enumerate", d, banned, retained, max.parents, which.nodes)
 * same arguments as buildcachematrix (src/buildcachematrix.c:15); returns
 * list(children = int[rows] (1-based), parents = int[width, rows] (1-based, 0 = none)) */
SEXP abn_cuda_enumerate(SEXP R_d, SEXP R_banned, SEXP R_retained, SEXP R_maxparents, SEXP R_which)
{
    const int d = asInteger(R_d), nw = LENGTH(R_which);
    int *which = (int *) R_alloc(nw, sizeof(int));
    for (int i = 0; i < nw; i++) which[i] = INTEGER(R_which)[i] - 1;
    char err[512] = "";
    int64_t rows = 0; int32_t width = 0;
    if (abn_b200_enumerate(d, INTEGER(R_banned), INTEGER(R_retained), INTEGER(R_maxparents), which, nw,
                           &rows, &width, NULL, NULL, err, sizeof err) != ABN_OK) error("%s", err);
    int32_t w = width > 0 ? width : 1;
    SEXP out = PROTECT(allocVector(VECSXP, 2));
    SEXP ch = PROTECT(allocVector(INTSXP, rows));
    SEXP pa = PROTECT(allocMatrix(INTSXP, w, rows));      /* column r = parents of candidate r */
    if (abn_b200_enumerate(d, INTEGER(R_banned), INTEGER(R_retained), INTEGER(R_maxparents), which, nw,
                           &rows, &w, INTEGER(ch), INTEGER(pa), err, sizeof err) != ABN_OK) error("%s", err);
    for (R_xlen_t i = 0; i < rows; i++) INTEGER(ch)[i] += 1;
    for (R_xlen_t i = 0; i < (R_xlen_t) rows * w; i++) INTEGER(pa)[i] += 1;   /* -1 -> 0 */
    SET_VECTOR_ELT(out, 0, ch); SET_VECTOR_ELT(out, 1, pa);
    UNPROTECT(3);
    return out;
}

/* .Call("abn_cuda_score_cache", X, node.family, node.col0, children, parents, method, ctrl, nvars, device)
 * X: double matrix n x n_cols, the prepared data (data.df.multi for mle, data.df for bayes);
 * node.family: get.var.types() codes; node.col0: 0-based first column per node, length d + 1;
 * children / parents: as returned by abn_cuda_enumerate; method 0 = mle, 1 = bayes;
 * ctrl = c(max.iters, tol, epsabs, mean, 1/sqrt(prec), loggam.shape, 1/loggam.inv.scale)
 * returns list(mlik, aic, bic, mdl, error.code, passes) */
SEXP abn_cuda_score_cache(SEXP R_X, SEXP R_family, SEXP R_col0, SEXP R_children, SEXP R_parents,
                          SEXP R_method, SEXP R_ctrl, SEXP R_nvars, SEXP R_device)
{
    abn_dataset ds = { nrows(R_X), LENGTH(R_family), ncols(R_X), REAL(R_X), INTEGER(R_family), INTEGER(R_col0) };
    const R_xlen_t rows = XLENGTH(R_children);
    const int w = nrows(R_parents);
    int *ch = (int *) R_alloc(rows, sizeof(int)), *pa = (int *) R_alloc((size_t) rows * w, sizeof(int));
    for (R_xlen_t i = 0; i < rows; i++) ch[i] = INTEGER(R_children)[i] - 1;
    for (R_xlen_t i = 0; i < rows * w; i++) pa[i] = INTEGER(R_parents)[i] - 1;
    abn_candidates cand = { rows, w, ch, pa };
    abn_control c;
    abn_b200_default_control(asInteger(R_method), &c);
    const double *v = REAL(R_ctrl);
    c.max_iters = (int) v[0]; c.tol = v[1]; c.epsabs = v[2]; c.prior_mean = v[3]; c.prior_sd = v[4];
    c.gam_shape = v[5]; c.gam_scale = v[6];
    c.nvars = asInteger(R_nvars);
    /* device: one ordinal, or a vector of ordinals = control$ncores' analogue: the candidate list is split over
     * these GPUs inside the library (abn_control.ngpus / devices) */
    const int ndev = LENGTH(R_device);
    if (ndev < 1 || ndev > ABN_MAX_DEVICES) error("abn_b200: 1..%d devices expected", ABN_MAX_DEVICES);
    c.device = INTEGER(R_device)[0];
    if (ndev > 1) {
        c.ngpus = ndev;
        for (int i = 0; i < ndev; i++) c.devices[i] = INTEGER(R_device)[i];
    }
    /* c.flags = ABN_FLAG_EXACT_PATH;   to follow the reference's IRLS pass for pass */
    const char *nm[] = { "mlik", "aic", "bic", "mdl", "error.code", "passes", "" };
    SEXP out = PROTECT(mkNamed(VECSXP, nm));
    for (int i = 0; i < 4; i++) SET_VECTOR_ELT(out, i, allocVector(REALSXP, rows));
    for (int i = 4; i < 6; i++) SET_VECTOR_ELT(out, i, allocVector(INTSXP, rows));
    abn_result res = { REAL(VECTOR_ELT(out, 0)), REAL(VECTOR_ELT(out, 1)), REAL(VECTOR_ELT(out, 2)),
                       REAL(VECTOR_ELT(out, 3)), INTEGER(VECTOR_ELT(out, 4)), INTEGER(VECTOR_ELT(out, 5)) };
    char err[512] = "";
    int rc = abn_b200_score_cache(&ds, &cand, &c, &res, err, sizeof err);   /* no R API inside */
    UNPROTECT(1);
    if (rc != ABN_OK) error("abn_b200: %s", err);      /* Rf_error last: no C++/CUDA state is live here */
    return out;
}
