/* TEST INFRASTRUCTURE - part of the CPU oracle (see oracle/README.md). */
#include "linalg.h"

#include <float.h>
#include <math.h>
#include <string.h>

int or_lu_decomp(double *A, int m, int *perm, int *signum)
{
    /* gsl_linalg_LU_decomp: for each column j pick the row with the largest |a_ij|, i >= j */
    *signum = 1;
    for (int i = 0; i < m; i++) perm[i] = i;
    for (int j = 0; j + 1 < m; j++) {
        double max = fabs(A[j * m + j]);
        int ip = j;
        for (int i = j + 1; i < m; i++) {
            double a = fabs(A[i * m + j]);
            if (a > max) { max = a; ip = i; }
        }
        if (ip != j) {
            for (int c = 0; c < m; c++) {
                double t = A[j * m + c]; A[j * m + c] = A[ip * m + c]; A[ip * m + c] = t;
            }
            int t = perm[j]; perm[j] = perm[ip]; perm[ip] = t;
            *signum = -*signum;
        }
        double ajj = A[j * m + j];
        if (ajj != 0.0) {
            for (int i = j + 1; i < m; i++) {
                double aij = A[i * m + j] / ajj;
                A[i * m + j] = aij;
                for (int c = j + 1; c < m; c++) A[i * m + c] -= aij * A[j * m + c];
            }
        }
    }
    return 0;
}

double or_lu_lndet(const double *LU, int m)
{
    double s = 0.0;
    for (int i = 0; i < m; i++) s += log(fabs(LU[i * m + i]));
    return s;
}

int or_lu_singular(const double *LU, int m)
{
    for (int i = 0; i < m; i++) if (LU[i * m + i] == 0.0) return 1;
    return 0;
}

void or_lu_solve(const double *LU, const int *perm, int m, const double *b, double *x)
{
    double y[OR_MAXM];
    for (int i = 0; i < m; i++) y[i] = b[perm[i]];
    for (int i = 0; i < m; i++) {
        double s = y[i];
        for (int j = 0; j < i; j++) s -= LU[i * m + j] * y[j];
        y[i] = s;
    }
    for (int i = m - 1; i >= 0; i--) {
        double s = y[i];
        for (int j = i + 1; j < m; j++) s -= LU[i * m + j] * y[j];
        y[i] = s / LU[i * m + i];
    }
    memcpy(x, y, sizeof(double) * m);
}

int or_lu_invert(const double *LU, const int *perm, int m, double *inv)
{
    if (or_lu_singular(LU, m)) return 1;
    double e[OR_MAXM], col[OR_MAXM];
    for (int j = 0; j < m; j++) {
        for (int i = 0; i < m; i++) e[i] = (i == j) ? 1.0 : 0.0;
        or_lu_solve(LU, perm, m, e, col);
        for (int i = 0; i < m; i++) inv[i * m + j] = col[i];
    }
    return 0;
}

int or_chol(double *A, int m)
{
    for (int j = 0; j < m; j++) {
        double s = A[j * m + j];
        for (int k = 0; k < j; k++) s -= A[j * m + k] * A[j * m + k];
        if (!(s > 0.0) || !isfinite(s)) return 1;
        double ljj = sqrt(s);
        A[j * m + j] = ljj;
        for (int i = j + 1; i < m; i++) {
            double t = A[i * m + j];
            for (int k = 0; k < j; k++) t -= A[i * m + k] * A[j * m + k];
            A[i * m + j] = t / ljj;
        }
    }
    return 0;
}

void or_chol_solve(const double *L, int m, const double *b, double *x)
{
    double y[OR_MAXM];
    for (int i = 0; i < m; i++) {
        double s = b[i];
        for (int j = 0; j < i; j++) s -= L[i * m + j] * y[j];
        y[i] = s / L[i * m + i];
    }
    for (int i = m - 1; i >= 0; i--) {
        double s = y[i];
        for (int j = i + 1; j < m; j++) s -= L[j * m + i] * y[j];
        y[i] = s / L[i * m + i];
    }
    memcpy(x, y, sizeof(double) * m);
}

void or_jacobi_eig(const double *A, int m, double *w, double *V)
{
    double a[OR_MAXM * OR_MAXM];
    memcpy(a, A, sizeof(double) * m * m);
    for (int i = 0; i < m; i++)
        for (int j = 0; j < m; j++) V[i * m + j] = (i == j) ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 60; sweep++) {
        double off = 0.0;
        for (int i = 0; i < m; i++)
            for (int j = i + 1; j < m; j++) off += a[i * m + j] * a[i * m + j];
        if (off == 0.0) break;
        for (int p = 0; p < m; p++) {
            for (int q = p + 1; q < m; q++) {
                double apq = a[p * m + q];
                if (apq == 0.0) continue;
                double theta = (a[q * m + q] - a[p * m + p]) / (2.0 * apq);
                double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
                for (int k = 0; k < m; k++) {
                    double akp = a[k * m + p], akq = a[k * m + q];
                    a[k * m + p] = c * akp - s * akq;
                    a[k * m + q] = s * akp + c * akq;
                }
                for (int k = 0; k < m; k++) {
                    double apk = a[p * m + k], aqk = a[q * m + k];
                    a[p * m + k] = c * apk - s * aqk;
                    a[q * m + k] = s * apk + c * aqk;
                }
                for (int k = 0; k < m; k++) {
                    double vkp = V[k * m + p], vkq = V[k * m + q];
                    V[k * m + p] = c * vkp - s * vkq;
                    V[k * m + q] = s * vkp + c * vkq;
                }
            }
        }
    }
    for (int i = 0; i < m; i++) w[i] = a[i * m + i];
}

int or_arma_solve(const double *A, const double *b, int m, int allow_approx, double *x)
{
    double W[OR_MAXM * OR_MAXM];
    int finite = 1;
    for (int i = 0; i < m * m; i++) if (!isfinite(A[i])) finite = 0;
    for (int i = 0; i < m; i++) if (!isfinite(b[i])) finite = 0;
    if (finite) {
        memcpy(W, A, sizeof(double) * m * m);
        if (or_chol(W, m) == 0) {
            or_chol_solve(W, m, b, x);
            return 0;
        }
        int perm[OR_MAXM], sg;
        memcpy(W, A, sizeof(double) * m * m);
        or_lu_decomp(W, m, perm, &sg);
        /* dgesv reports singular on an exact zero pivot; Armadillo additionally treats
         * rcond < eps as failure when approximation is allowed */
        double umax = 0.0, umin = DBL_MAX;
        for (int i = 0; i < m; i++) {
            double u = fabs(W[i * m + i]);
            if (u > umax) umax = u;
            if (u < umin) umin = u;
        }
        int bad = (umin == 0.0) || (allow_approx && umin < umax * DBL_EPSILON);
        if (!bad) {
            or_lu_solve(W, perm, m, b, x);
            return 0;
        }
        if (allow_approx) {
            /* minimum-norm least squares through the eigen-decomposition of the symmetric A */
            double w[OR_MAXM], V[OR_MAXM * OR_MAXM], S[OR_MAXM * OR_MAXM];
            for (int i = 0; i < m; i++)
                for (int j = 0; j < m; j++) S[i * m + j] = 0.5 * (A[i * m + j] + A[j * m + i]);
            or_jacobi_eig(S, m, w, V);
            double wmax = 0.0;
            for (int i = 0; i < m; i++) if (fabs(w[i]) > wmax) wmax = fabs(w[i]);
            double tol = m * wmax * DBL_EPSILON;
            for (int i = 0; i < m; i++) x[i] = 0.0;
            for (int e = 0; e < m; e++) {
                if (fabs(w[e]) <= tol) continue;
                double s = 0.0;
                for (int i = 0; i < m; i++) s += V[i * m + e] * b[i];
                s /= w[e];
                for (int i = 0; i < m; i++) x[i] += s * V[i * m + e];
            }
            return 1;
        }
    }
    for (int i = 0; i < m; i++) x[i] = NAN;
    return -1;
}

int or_rank(const double *X, long n, int k)
{
    /* one-sided (Hestenes) Jacobi on a copy of the columns; singular values = column norms */
    if (k == 0 || n == 0) return 0;
    double *U = (double *)__builtin_malloc(sizeof(double) * (size_t)n * k);
    if (!U) return -1;
    memcpy(U, X, sizeof(double) * (size_t)n * k);
    for (int sweep = 0; sweep < 40; sweep++) {
        int rotated = 0;
        for (int p = 0; p < k; p++) {
            for (int q = p + 1; q < k; q++) {
                double *up = U + (size_t)p * n, *uq = U + (size_t)q * n;
                double a = 0, b = 0, c = 0;
                for (long i = 0; i < n; i++) { a += up[i] * up[i]; b += uq[i] * uq[i]; c += up[i] * uq[i]; }
                if (fabs(c) <= 1e-15 * sqrt(a * b) || c == 0.0) continue;
                rotated = 1;
                double zeta = (b - a) / (2.0 * c);
                double t = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                double cs = 1.0 / sqrt(1.0 + t * t), sn = cs * t;
                for (long i = 0; i < n; i++) {
                    double x = up[i], y = uq[i];
                    up[i] = cs * x - sn * y;
                    uq[i] = sn * x + cs * y;
                }
            }
        }
        if (!rotated) break;
    }
    double sv[OR_MAXM], smax = 0.0;
    for (int p = 0; p < k; p++) {
        double a = 0;
        const double *up = U + (size_t)p * n;
        for (long i = 0; i < n; i++) a += up[i] * up[i];
        sv[p] = sqrt(a);
        if (sv[p] > smax) smax = sv[p];
    }
    __builtin_free(U);
    double tol = (double)(n > k ? n : k) * smax * DBL_EPSILON;
    int r = 0;
    for (int p = 0; p < k; p++) if (sv[p] > tol) r++;
    return r;
}

void or_qr(const double *A, int m, double *Q, double *R)
{
    /* Householder QR, the construction of gsl_linalg_QR_decomp / QR_unpack */
    memcpy(R, A, sizeof(double) * m * m);
    for (int i = 0; i < m; i++)
        for (int j = 0; j < m; j++) Q[i * m + j] = (i == j) ? 1.0 : 0.0;
    for (int j = 0; j < m; j++) {
        /* gsl_linalg_householder_transform on column j, rows j..m-1 */
        double xnorm = 0.0;
        for (int i = j + 1; i < m; i++) xnorm += R[i * m + j] * R[i * m + j];
        xnorm = sqrt(xnorm);
        if (xnorm == 0.0) continue; /* tau = 0 */
        double alpha = R[j * m + j];
        double beta = -(alpha >= 0.0 ? +1.0 : -1.0) * hypot(alpha, xnorm);
        double tau = (beta - alpha) / beta;
        double v[OR_MAXM];
        double s = alpha - beta;
        v[j] = 1.0;
        for (int i = j + 1; i < m; i++) v[i] = R[i * m + j] / s;
        /* R <- (I - tau v v^T) R */
        for (int c = j; c < m; c++) {
            double w = 0.0;
            for (int i = j; i < m; i++) w += v[i] * R[i * m + c];
            w *= tau;
            for (int i = j; i < m; i++) R[i * m + c] -= w * v[i];
        }
        /* Q <- Q (I - tau v v^T) */
        for (int r = 0; r < m; r++) {
            double w = 0.0;
            for (int i = j; i < m; i++) w += Q[r * m + i] * v[i];
            w *= tau;
            for (int i = j; i < m; i++) Q[r * m + i] -= w * v[i];
        }
        for (int i = j + 1; i < m; i++) R[i * m + j] = 0.0;
    }
}

static void givens(double a, double b, double *c, double *s)
{
    /* gsl_linalg_givens */
    if (b == 0) { *c = 1; *s = 0; }
    else if (fabs(b) > fabs(a)) {
        double t = -a / b;
        double s1 = 1.0 / sqrt(1 + t * t);
        *s = s1; *c = s1 * t;
    } else {
        double t = -b / a;
        double c1 = 1.0 / sqrt(1 + t * t);
        *c = c1; *s = c1 * t;
    }
}

static void apply_givens_qr(int m, double *Q, double *R, int i, int j, double c, double s)
{
    /* Q' = Q G */
    for (int k = 0; k < m; k++) {
        double qki = Q[k * m + i], qkj = Q[k * m + j];
        Q[k * m + i] = qki * c - qkj * s;
        Q[k * m + j] = qki * s + qkj * c;
    }
    /* R' = G^T R (note: upper triangular so fewer elements change) */
    int k0 = (i < j ? i : j);
    for (int k = k0; k < m; k++) {
        double rik = R[i * m + k], rjk = R[j * m + k];
        R[i * m + k] = c * rik - s * rjk;
        R[j * m + k] = s * rik + c * rjk;
    }
}

void or_qr_update(double *Q, double *R, int m, double *w, const double *v)
{
    /* gsl_linalg_QR_update (Golub & Van Loan 12.5): reduce w to +-|w| e_1 by Givens rotations
     * applied to R as well (making it upper Hessenberg), add w0 v^T to the first row, then
     * restore triangularity. */
    for (int k = m - 1; k > 0; k--) {
        double c, s;
        givens(w[k - 1], w[k], &c, &s);
        double wi = w[k - 1], wj = w[k];
        w[k - 1] = c * wi - s * wj;
        w[k] = s * wi + c * wj;
        apply_givens_qr(m, Q, R, k - 1, k, c, s);
    }
    double w0 = w[0];
    for (int j = 0; j < m; j++) R[0 * m + j] += w0 * v[j];
    for (int k = 1; k < m; k++) {
        double c, s;
        givens(R[(k - 1) * m + (k - 1)], R[k * m + (k - 1)], &c, &s);
        apply_givens_qr(m, Q, R, k - 1, k, c, s);
        R[k * m + (k - 1)] = 0.0;
    }
}
