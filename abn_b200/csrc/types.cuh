// types.cuh - plain structs shared by the narrow (MAXK = 8, namespace abn) and wide (MAXK = 32,
// namespace abnw) builds of the solver headers. Nothing here depends on MAXK.
#pragma once

namespace abnc {

struct Control {
    int method, max_iters, solver, nvars;
    double tol, epsabs, prior_mean, prior_sd, gam_shape, gam_scale;
};

// device-resident description of the data set
struct DataMeta {
    long long n;     // observations
    long long ld;    // leading dimension (elements) of the column-major matrix
    int n_nodes, n_cols;
    int D;           // Gram dimension = 1 + n_cols + 2 n_pois
    const double *X; // column-major, ld x n_cols
    const double *G; // D x D row-major symmetric
    const int *node_family; // [n_nodes]
    const int *node_col0;   // [n_nodes + 1]
    const double *node_lfact; // [n_nodes] sum lgamma(y + 1) (Poisson nodes, else 0)
    const double *Gw;         // per Poisson node: Z0' diag(y + 0.1) Z0, Z0 = [1 | X], D0 x D0 each
    const int *node_gw;       // [n_nodes] ordinal of the node's block in Gw, -1 if not Poisson
    int D0;                   // 1 + n_cols
    int n_pois;               // Poisson nodes: each owns two virtual Gram columns (plan.cuh gram_logy / gram_wz0)
    const int *pois_node;     // [n_pois] node of ordinal o (inverse of node_gw)
};

struct ResultPtrs {
    double *mlik, *aic, *bic, *mdl;
    int *status, *iters;
};

} // namespace abnc
