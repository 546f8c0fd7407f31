// types.cuh - plain structs shared by the narrow (MAXK = 8, namespace abn) and wide (MAXK = 32,
// namespace abnw) builds of the solver headers. Nothing here depends on MAXK.
#pragma once

namespace abnc {

struct Control {
    int method, max_iters, solver, nvars;
    double tol, epsabs, prior_mean, prior_sd, gam_shape, gam_scale;
    int flags; // ABN_FLAG_* of include/abn_b200.h
};

// device-resident description of the data set
struct DataMeta {
    long long n;     // observations
    long long ld;    // leading dimension (elements) of the column-major matrix
    int n_nodes, n_cols;
    int D;           // Gram dimension = 1 + n_cols + n_pois
    const double *X; // column-major, ld x n_cols
    const double *G; // D x D row-major symmetric
    const int *node_family; // [n_nodes]
    const int *node_col0;   // [n_nodes + 1]
    const double *node_lfact; // [n_nodes] sum lgamma(y + 1) (Poisson nodes, else 0)
    const int *node_gw;       // [n_nodes] ordinal of a Poisson node among the Poisson nodes, -1 otherwise
    int n_pois;               // Poisson nodes: each owns one virtual Gram column, log(y + 0.1) (plan.cuh gram_logy)
    const int *pois_node;     // [n_pois] node of ordinal o (inverse of node_gw)
    // discrete columns (fit_cell.cuh): byte codes, column-major with ldq bytes per column (16-byte aligned), and
    // per column max code + 1 if every value is an integer in 0..255, else 0. Null on host builds / when disabled.
    const unsigned char *Xq;
    long long ldq;
    const int *col_card;
    // bit planes of the BINARY columns (col_card <= 2): bit r mod 32 of word r / 32 is the code of row r; ldb words
    // per column (a multiple of 4), bits beyond n and the planes of other columns are zero. Null: not built.
    const unsigned *Xb;
    long long ldb;
    // capped codes: col_tail[c * 8 + b] = rows of column c whose code is >= 2^b (b = 0..7); rows beyond a cap go
    // to the warp's overflow list cell_ovf + slot * cell_ovf_stride (row indices). Null: no capping.
    const unsigned *col_tail;
    unsigned *cell_ovf;
    long long cell_ovf_stride;
    // centred second moments for Gaussian children (plan.cuh gauss_centered): Gc = Zc'Zc over
    // Zc = [1 | X - col_mean], (1 + n_cols)^2 row-major, col_mean[j] = column mean as the raw Gram gives it.
    // Residual sums of squares formed from these do not cancel when |mean| >> sd. Null: raw moments are used.
    const double *Gc;
    const double *col_mean;
};

struct ResultPtrs {
    double *mlik, *aic, *bic, *mdl;
    int *status, *iters;
    double *coef;    // optional [rows x coef_stride]: fitted coefficients / posterior modes (plan.cuh write_coef)
    int coef_stride;
    double *se;      // optional, same shape: mle standard errors (plan.cuh se_from_information)
    double *sse;     // optional [rows]: ||y - X beta||^2 as the reference's IRLS functions report it (plan.cuh sse_of)
};

} // namespace abnc
