/* TEST INFRASTRUCTURE - a ~100-line stand-in for the part of R's C API that integration/abn_cuda_shim.c uses,
 * so that the shim a maintainer drops into the reference's src/ is compiled and executed by this repository's
 * tests (there is no R in the image). SEXPs are tagged heap blocks; PROTECT / UNPROTECT only count; error()
 * longjmps back to stub_call (R's own error() longjmps to the top level). Nothing under abn_b200/ uses this. */
#ifndef ABN_R_STUB_RINTERNALS_H
#define ABN_R_STUB_RINTERNALS_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef ptrdiff_t R_xlen_t;
enum { INTSXP = 13, REALSXP = 14, STRSXP = 16, VECSXP = 19 };

typedef struct stub_sexp {
    int type;
    R_xlen_t length;
    int nrow, ncol;        /* 0 unless made by allocMatrix / stub_matrix */
    void *data;            /* int[], double[] or SEXP[] */
    const char **names;    /* VECSXP made by mkNamed */
} *SEXP;

SEXP allocVector(int type, R_xlen_t n);
SEXP allocMatrix(int type, int nrow, int ncol);
SEXP mkNamed(int type, const char **names);
SEXP PROTECT(SEXP s);
void UNPROTECT(int n);
int *INTEGER(SEXP s);
double *REAL(SEXP s);
SEXP VECTOR_ELT(SEXP s, R_xlen_t i);
SEXP SET_VECTOR_ELT(SEXP s, R_xlen_t i, SEXP v);
int LENGTH(SEXP s);
R_xlen_t XLENGTH(SEXP s);
int nrows(SEXP s);
int ncols(SEXP s);
int asInteger(SEXP s);
char *R_alloc(size_t n, int size);
/* as in R_ext/Error.h: error is a macro for Rf_error (a bare `error` would resolve to glibc's error(3)) */
void Rf_error(const char *fmt, ...) __attribute__((noreturn));
#define error Rf_error

/* ---- test-side helpers (not R API) ---- */
SEXP stub_int_vector(const int *v, R_xlen_t n);
SEXP stub_real_vector(const double *v, R_xlen_t n);
SEXP stub_int_matrix(const int *v, int nrow, int ncol);
SEXP stub_real_matrix_view(double *v, int nrow, int ncol); /* no copy: the shim must read R's own buffer */
const char *stub_last_error(void);
int stub_protect_depth(void);
const char *stub_name(SEXP s, int i);
/* calls fn(args...) the way .Call does, catching error(): NULL on error */
SEXP stub_call(void *fn, int nargs, SEXP *args);
void stub_free_all(void);

#ifdef __cplusplus
}
#endif
#endif
