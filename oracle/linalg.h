/* TEST INFRASTRUCTURE - part of the CPU oracle (see oracle/README.md). Not shipped, not linked
 * into libabn_b200.so.  Small dense linear algebra standing in for the pieces of GSL / LAPACK
 * (Armadillo) that the reference calls on its k x k systems. All matrices row-major m x m
 * with leading dimension m unless said otherwise. */
#ifndef ABN_ORACLE_LINALG_H
#define ABN_ORACLE_LINALG_H

#define OR_MAXM 40 /* widest system the oracle handles (multinomial child: (L-1)*k params) */

/* gsl_linalg_LU_decomp: Gaussian elimination with partial pivoting, in place.
 * Returns 0; *signum gets the permutation sign. */
int or_lu_decomp(double *A, int m, int *perm, int *signum);
/* gsl_linalg_LU_lndet: sum log |u_ii| */
double or_lu_lndet(const double *LU, int m);
/* 1 if any u_ii == 0 (gsl's singular() test used by LU_invert / LU_solve) */
int or_lu_singular(const double *LU, int m);
/* gsl_linalg_LU_solve */
void or_lu_solve(const double *LU, const int *perm, int m, const double *b, double *x);
/* gsl_linalg_LU_invert; returns non-zero (GSL_EDOM) when singular */
int or_lu_invert(const double *LU, const int *perm, int m, double *inv);

/* Cholesky A = L L^T (lower, in place). Returns non-zero if a pivot is <= 0 or non-finite. */
int or_chol(double *A, int m);
void or_chol_solve(const double *L, int m, const double *b, double *x);

/* arma::solve(A, b, opts): likely_sympd tries Cholesky (dposv), falls back to LU (dgesv);
 * when both fail: allow_approx ? minimum-norm least squares via symmetric eigen-decomposition
 * : failure. Returns 0 ok, 1 = used approximate solution, -1 = no solution (x set to NaN). */
int or_arma_solve(const double *A, const double *b, int m, int allow_approx, double *x);

/* cyclic Jacobi eigen-decomposition of a symmetric matrix; eigenvalues in w, vectors in
 * columns of V (row-major). */
void or_jacobi_eig(const double *A, int m, double *w, double *V);

/* arma::rank(X) for a column-major n x k matrix: number of singular values greater than
 * max(n,k) * sigma_max * DBL_EPSILON, singular values by one-sided Jacobi. */
int or_rank(const double *X, long n, int k);

/* Householder QR of a square matrix, gsl_linalg_QR_decomp + QR_unpack: A = Q R */
void or_qr(const double *A, int m, double *Q, double *R);
/* gsl_linalg_QR_update: Q' R' = Q (R + w v^T), Givens based */
void or_qr_update(double *Q, double *R, int m, double *w, const double *v);

#endif
