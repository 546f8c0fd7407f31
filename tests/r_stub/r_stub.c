/* TEST INFRASTRUCTURE - implementation of the R API stand-in declared in Rinternals.h */
#include <setjmp.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "Rinternals.h"

static void *g_blocks[1 << 16];
static int g_nblocks = 0, g_protect = 0;
static char g_error[1024] = "";
static jmp_buf g_jmp;
static int g_jmp_set = 0;

static void *keep(void *p)
{
    if (g_nblocks < (1 << 16)) g_blocks[g_nblocks++] = p;
    return p;
}

void stub_free_all(void)
{
    for (int i = 0; i < g_nblocks; i++) free(g_blocks[i]);
    g_nblocks = 0;
    g_protect = 0;
}

static SEXP make(int type, R_xlen_t n, void *data)
{
    SEXP s = (SEXP)keep(calloc(1, sizeof(*s)));
    s->type = type;
    s->length = n;
    const size_t el = type == INTSXP ? sizeof(int) : (type == REALSXP ? sizeof(double) : sizeof(SEXP));
    s->data = data ? data : keep(calloc(n > 0 ? (size_t)n : 1, el));
    return s;
}

SEXP allocVector(int type, R_xlen_t n) { return make(type, n, NULL); }
SEXP allocMatrix(int type, int nrow, int ncol)
{
    SEXP s = make(type, (R_xlen_t)nrow * ncol, NULL);
    s->nrow = nrow;
    s->ncol = ncol;
    return s;
}
SEXP mkNamed(int type, const char **names)
{
    int n = 0;
    while (names[n][0]) n++;
    SEXP s = make(type, n, NULL);
    /* R copies the names into a STRSXP; the caller's array may live on its stack */
    const char **copy = (const char **)keep(calloc((size_t)n + 1, sizeof(char *)));
    for (int i = 0; i < n; i++) {
        char *c = (char *)keep(malloc(strlen(names[i]) + 1));
        strcpy(c, names[i]);
        copy[i] = c;
    }
    s->names = copy;
    return s;
}
SEXP PROTECT(SEXP s) { g_protect++; return s; }
void UNPROTECT(int n) { g_protect -= n; }
int *INTEGER(SEXP s) { if (s->type != INTSXP) error("INTEGER() on a non-integer SEXP"); return (int *)s->data; }
double *REAL(SEXP s) { if (s->type != REALSXP) error("REAL() on a non-double SEXP"); return (double *)s->data; }
SEXP VECTOR_ELT(SEXP s, R_xlen_t i) { return ((SEXP *)s->data)[i]; }
SEXP SET_VECTOR_ELT(SEXP s, R_xlen_t i, SEXP v) { ((SEXP *)s->data)[i] = v; return v; }
int LENGTH(SEXP s) { return (int)s->length; }
R_xlen_t XLENGTH(SEXP s) { return s->length; }
int nrows(SEXP s) { return s->nrow ? s->nrow : (int)s->length; }
int ncols(SEXP s) { return s->nrow ? s->ncol : 1; }
int asInteger(SEXP s) { return s->type == INTSXP ? ((int *)s->data)[0] : (int)((double *)s->data)[0]; }
char *R_alloc(size_t n, int size) { return (char *)keep(calloc(n > 0 ? n : 1, (size_t)size)); }

void Rf_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
    if (g_jmp_set) longjmp(g_jmp, 1);
    fprintf(stderr, "r_stub: error() outside stub_call: %s\n", g_error);
    abort();
}

SEXP stub_int_vector(const int *v, R_xlen_t n)
{
    SEXP s = make(INTSXP, n, NULL);
    if (n > 0) memcpy(s->data, v, (size_t)n * sizeof(int));
    return s;
}
SEXP stub_real_vector(const double *v, R_xlen_t n)
{
    SEXP s = make(REALSXP, n, NULL);
    if (n > 0) memcpy(s->data, v, (size_t)n * sizeof(double));
    return s;
}
SEXP stub_int_matrix(const int *v, int nrow, int ncol)
{
    SEXP s = stub_int_vector(v, (R_xlen_t)nrow * ncol);
    s->nrow = nrow;
    s->ncol = ncol;
    return s;
}
SEXP stub_real_matrix_view(double *v, int nrow, int ncol)
{
    SEXP s = (SEXP)keep(calloc(1, sizeof(*s)));
    s->type = REALSXP;
    s->length = (R_xlen_t)nrow * ncol;
    s->nrow = nrow;
    s->ncol = ncol;
    s->data = v; /* owned by the caller */
    return s;
}
const char *stub_last_error(void) { return g_error; }
int stub_protect_depth(void) { return g_protect; }
const char *stub_name(SEXP s, int i) { return s->names ? s->names[i] : ""; }

typedef SEXP (*fn5_t)(SEXP, SEXP, SEXP, SEXP, SEXP);
typedef SEXP (*fn9_t)(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP);

SEXP stub_call(void *fn, int nargs, SEXP *a)
{
    g_error[0] = 0;
    SEXP out = NULL;
    g_jmp_set = 1;
    if (setjmp(g_jmp) == 0) {
        if (nargs == 5) out = ((fn5_t)fn)(a[0], a[1], a[2], a[3], a[4]);
        else if (nargs == 9) out = ((fn9_t)fn)(a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7], a[8]);
        else snprintf(g_error, sizeof(g_error), "stub_call: unsupported arity %d", nargs);
    } else {
        out = NULL;
        g_protect = 0; /* R unwinds the protect stack on error */
    }
    g_jmp_set = 0;
    return out;
}
