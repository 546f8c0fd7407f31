/*
 * abn_b200.h - C ABI of the B200-native score-cache path of abn (furrer-lab/abn v3.1.13).
 *
 * This is the drop-in boundary: a plain-C shared library (libabn_b200.so, built from
 * abn_b200/csrc/ with nvcc for sm_100a).  No R, Rcpp, torch or C++ types cross it; every
 * pointer is caller-owned and never retained after the call returns (device state lives
 * behind the opaque abn_data handle).  INTEGRATION.md shows the R .Call shim a maintainer
 * adds to src/ and the R_CallMethodDef rows for src/RcppExports.cpp:122-138.
 *
 * Reference interfaces replaced (all paths relative to the reference tree):
 *
 *   abn_b200_enumerate      <- buildcachematrix()         src/buildcachematrix.c:15-141
 *                              + isretained()             src/buildcachematrix.c:150-162
 *                              + combn enumeration        R/build_score_cache_mle.R:432-518
 *   abn_b200_node_defn      <- node.defn matrix assembly  R/build_score_cache_bayes.R:82-85
 *   abn_b200_score_cache    <- the per-candidate loops    R/build_score_cache_mle.R:615-703
 *                                                         R/build_score_cache_bayes.R:121-214
 *                              and everything they call:
 *                                rank_cpp                 src/rank.cpp:15
 *                                irls_gaussian_cpp_fast   src/irls_gaussian_fast.cpp:12
 *                                irls_binomial_cpp_fast_br src/irls_binomial_fast_br.cpp:15
 *                                irls_poisson_cpp_fast    src/irls_poisson_fast.cpp:16
 *                                nnet.default softmax fit R/build_score_cache_mle.R:319-331
 *                                fit_single_node          src/fit_single_node.c:25
 *                                calc_node_Score_binary   src/node_binomial.c:18
 *                                calc_node_Score_gaus     src/node_gaussian.c:18
 *                                calc_node_Score_pois     src/node_poisson.c:20
 *   abn_b200_data_create /  <- make_data() (copies the whole data frame per candidate,
 *   abn_b200_data_destroy      src/utility.c:106-141); here it is done once per cache
 *   abn_b200_score          <- same as abn_b200_score_cache, on an already resident data set
 *
 * Conventions: all node / column indices that cross this ABI are 0-based unless a comment
 * says otherwise; matrices are column-major (R layout) unless a comment says otherwise;
 * all floating point is IEEE double.
 */
#ifndef ABN_B200_H
#define ABN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ABN_B200_VERSION 200 /* 0.2.0: abn_control grew ngpus / block / devices[], abn_timing grew kernel_fp64[] */

/* node families: the codes of get.var.types() (R/abn-internal.R:833) plus multinomial */
enum {
    ABN_BINOMIAL = 1,
    ABN_GAUSSIAN = 2,
    ABN_POISSON = 3,
    ABN_MULTINOMIAL = 4
};

enum {
    ABN_METHOD_MLE = 0,  /* buildScoreCache(method = "mle")   */
    ABN_METHOD_BAYES = 1 /* buildScoreCache(method = "bayes") */
};

/* Bayes root finder: what locates the mode of the Laplace integrand */
enum {
    ABN_SOLVER_NEWTON = 0, /* full Newton steps with the analytic Hessian, run to the epsabs test */
    ABN_SOLVER_HYBRID = 1  /* Powell hybrid (MINPACK HYBRJ as in GSL hybridsj / hybridj), the
                              reference's iterate sequence and its error.code = 1 outcomes */
};

/* per-fit status: the reference's error.code (src/node_binomial.c:117,152,192-193) */
enum {
    ABN_FIT_OK = 0,
    ABN_FIT_NOT_CONVERGED = 1, /* solver left without meeting its stop rule               */
    ABN_FIT_NON_FINITE = 2     /* score is NaN / a non-finite weight or Gram was met       */
};

/* return codes of the entry points */
enum {
    ABN_OK = 0,
    ABN_ERR_ARG = -1,     /* bad argument (message says which)                            */
    ABN_ERR_CUDA = -2,    /* CUDA runtime error (message carries cudaGetErrorString)      */
    ABN_ERR_NO_GPU = -3,  /* no CUDA device / kernel image not loadable - there is no CPU path */
    ABN_ERR_LIMIT = -4,   /* nothing could be scored (designs wider than ABN_MAX_K columns ...)  */
    ABN_ERR_ALLOC = -5
};

#define ABN_MAX_K 32 /* widest design matrix (intercept + expanded parent columns); designs of up to
                       ABN_FAST_K columns run in the register-resident kernels, wider ones in a generic kernel */
#define ABN_FAST_K 8

/*
 * The data set, already prepared as the R drivers prepare it before their loops
 * (R/build_score_cache_mle.R:398-415,563-599; R/build_score_cache_bayes.R:40-60):
 * Gaussian columns centred and scaled, binomial columns 0/1, Poisson columns counts,
 * multinomial nodes one-hot expanded with ALL levels kept.
 */
typedef struct abn_dataset {
    int64_t n;                  /* observations                                              */
    int32_t n_nodes;            /* d: variables / nodes                                      */
    int32_t n_cols;             /* stored columns = sum over nodes of their column counts    */
    const double *X;            /* column-major n x n_cols                                   */
    const int32_t *node_family; /* [n_nodes] ABN_BINOMIAL ...                                */
    const int32_t *node_col0;   /* [n_nodes + 1] node i owns columns node_col0[i] .. node_col0[i+1]-1
                                   (one column, or nlevels columns for a multinomial node)   */
} abn_dataset;

/*
 * The candidate list: one (child, parent set) per row, in the order of the cache.
 * parents is row-major rows x max_parents, each row ascending, padded with -1.
 */
typedef struct abn_candidates {
    int64_t rows;
    int32_t max_parents;    /* row width of parents[]                                       */
    const int32_t *child;   /* [rows] 0-based node id                                       */
    const int32_t *parents; /* [rows * max_parents] 0-based node ids, -1 = unused slot      */
} abn_candidates;

/* abn_control.flags */
#define ABN_FLAG_EXACT_PATH 1 /* mle: follow the reference's IRLS pass for pass - its own start values and its
                                 own stopping rule - instead of ending a fit once the remaining terms are
                                 below 1e-11 of |loglik| ("IRLS shortcuts"); ~2.5x the passes */

#define ABN_FLAG_NO_BIAS_REDUCTION 2 /* mle, binomial children: b' = y, i.e. irls_binomial_cpp_fast (src/irls_binomial_fast.cpp)
                                        instead of the bias-reduced irls_binomial_cpp_fast_br the score cache uses
                                        (fitAbn's first choice, R/fitabn_mle.R:585) */

#define ABN_FLAG_RESULT_ON_DEVICE 4 /* abn_b200_score only: the abn_result pointers (mlik, aic, bic, mdl, status, iters; not
                                       coef / se / sse) are DEVICE pointers on the handle's device; the kernels write the
                                       scores there and nothing is copied to the host - the caller (e.g. one rank of a
                                       multi-process run) hands the slab to its collective. The call still returns only
                                       after the scores are complete. */

#define ABN_MAX_DEVICES 16

/* build.control() scalars that reach the natives (R/build_score_cache.R:90-127,
 * R/build_score_cache_bayes.R:314-328) */
typedef struct abn_control {
    int32_t method;     /* ABN_METHOD_MLE / ABN_METHOD_BAYES                                */
    int32_t max_iters;  /* control$max.iters (100)                                           */
    double tol;         /* mle: control$tol (1e-8), ||x - x_old||_2 stop                     */
    double epsabs;      /* bayes: control$epsabs (1e-7), sum |grad g| stop                   */
    double prior_mean;  /* bayes: control$mean (0)                                           */
    double prior_sd;    /* bayes: 1/sqrt(control$prec) (prec = 0.001)                        */
    double gam_shape;   /* bayes: control$loggam.shape (1)                                   */
    double gam_scale;   /* bayes: 1/control$loggam.inv.scale (inv.scale = 5e-05)             */
    int32_t solver;     /* bayes: ABN_SOLVER_*                                               */
    int32_t device;     /* CUDA device ordinal this call runs on (one process per GPU)       */
    int32_t nvars;      /* MDL: number of variables in log(nvars), R/build_score_cache_mle.R:343 */
    int32_t flags;      /* ABN_FLAG_*; 0 by default                                          */
    /* Multi-GPU, the analogue of control$ncores / cluster.type (R/build_score_cache.R:111-112, the foreach loops of
     * R/build_score_cache_mle.R:636-681 and R/build_score_cache_bayes.R:149-193): with ngpus > 1
     * abn_b200_score_cache runs one host thread per listed device; every device holds the whole data set and
     * scores the blocks b = i, i + ngpus, ... of `block` consecutive candidates (block-cyclic: every device sees
     * the cache's own mix of families and parent counts, neighbours stay together), and the scores land in the
     * caller's arrays in cache order. Results are bit-identical to a one-device run. ngpus <= 1: `device` alone. */
    int32_t ngpus;
    int32_t block;      /* candidates per block of the split; 0 = 1024                       */
    int32_t devices[ABN_MAX_DEVICES];
} abn_control;

/* Output arrays, all caller-allocated with `rows` entries; any pointer may be NULL.
 * mle fills mlik (= log-likelihood), aic, bic, mdl; bayes fills mlik only. */
typedef struct abn_result {
    double *mlik;
    double *aic;
    double *bic;
    double *mdl;
    int32_t *status; /* ABN_FIT_*                                                            */
    int32_t *iters;  /* data passes / solver iterations executed, for the roofline account   */
    /* Optional (NULL = not wanted): the fitted coefficients (mle: irls_*_cpp_fast()$coefficients; bayes: the
     * posterior modes fit_single_node leaves in its `modes` vector, src/fit_single_node.c:182-194), rows x
     * coef_stride, row-major. Slot 0 = intercept, slot 1 + t = the t-th stored column of the row's parents in
     * candidate order (a multinomial parent owns one slot per level column); Gaussian child under bayes: the
     * slot after the last column holds the precision parameter. Columns dropped by the rank rule, absent
     * intercepts, multinomial children and failed designs are NaN. Slots >= coef_stride are not written. mle:
     * the default path ends a fit once its SCORE is settled to 1e-11, which can leave coefficients ~1e-6 short
     * along ill-conditioned directions; set ABN_FLAG_EXACT_PATH for the reference's own convergence (1e-8 step).
     * This is the first step towards fitAbn. */
    double *coef;
    int32_t coef_stride;
    int32_t reserved;
    /* Optional, needs coef (same shape and slots), mle only: the standard errors fitAbn(method = "mle") reports,
     * sqrt(diag(solve(varcov))) of the IRLS functions' `varcov` = X'WX of the last iteration for binomial and
     * Poisson children, sqrt(diag(mse * solve(X'X))) for Gaussian ones (R/fitabn_mle.R:712-745). Bayes: NaN. */
    double *se;
    /* Optional [rows], mle only: the `sse` of the reference's IRLS functions, ||y - X beta||^2 on the scale of the
     * linear predictor whatever the family (src/irls_binomial_fast.cpp:128-130); fitAbn's mse = sse / df. */
    double *sse;
} abn_result;

typedef struct abn_data abn_data; /* opaque: device-resident data set + its global Gram     */

/* fills *c with build.control() defaults for `method` */
void abn_b200_default_control(int32_t method, abn_control *c);

int32_t abn_b200_version(void);

/* number of CUDA devices visible; 0 with no driver. Never fails. */
int32_t abn_b200_device_count(void);

/*
 * Candidate enumeration, bit-exact with buildcachematrix (and hence with the mle combn path,
 * tests/testthat/test-build_score_cache.R:174-184): for every node in which_nodes, for
 * j = 0 .. max_parents[node], all j-subsets of {0..d-1} in lexicographic order, dropping
 * subsets that contain the node, contain a banned arc or miss a retained arc.
 *
 * banned / retained: d x d column-major 0/1, entry [child + parent * d] (utility.c:35).
 * Call once with children == parents == NULL to get *rows, then again with storage:
 * children [rows] 0-based; parents [rows * width] row-major, -1 padded, where
 * width = max over which_nodes of max_parents[node].
 */
int abn_b200_enumerate(int32_t d, const int32_t *banned, const int32_t *retained,
                       const int32_t *max_parents, const int32_t *which_nodes, int32_t n_which,
                       int64_t *rows, int32_t *width, int32_t *children, int32_t *parents,
                       char *err, size_t errlen);

/* expands the compact parent lists to the rows x d 0/1 integer matrix `node.defn`,
 * column-major as R stores it */
int abn_b200_node_defn(const abn_candidates *cand, int32_t d, int32_t *node_defn,
                       char *err, size_t errlen);

/* uploads the data set to ctrl-independent device storage on `device` and computes the
 * global Gram matrix once */
int abn_b200_data_create(const abn_dataset *ds, int32_t device, abn_data **out,
                         char *err, size_t errlen);
void abn_b200_data_destroy(abn_data *d);

/* scores `cand` on a resident data set; synchronous (results are in host memory on return) */
int abn_b200_score(abn_data *d, const abn_candidates *cand, const abn_control *ctrl,
                   abn_result *out, char *err, size_t errlen);

/*
 * A whole cache without a candidate upload: the list abn_b200_enumerate describes is generated on the
 * device - row r from r, by combinatorial unranking over the node's admissible parents - and scored in place, so
 * that only the d x d ban / retain matrices travel to the GPU (config 4: 260 MB of indices otherwise). Arguments as
 * for abn_b200_enumerate; `rows` and `width` are what that function reports (call it with children == NULL first:
 * it only counts), the output arrays of `out` have `rows` entries. children / parents (optional, both or neither,
 * host memory, same layout as abn_b200_enumerate fills) receive the list for the abnCache's children / node.defn.
 * Bit-identical to abn_b200_enumerate + abn_b200_score. Replaces buildcachematrix()'s gsl_combination walk
 * (src/buildcachematrix.c:41-130) together with the loops over its result.
 */
int abn_b200_score_enumerated(abn_data *d, const int32_t *banned, const int32_t *retained,
                              const int32_t *max_parents, const int32_t *which_nodes, int32_t n_which,
                              const abn_control *ctrl, abn_result *out, int64_t rows, int32_t width,
                              int32_t *children, int32_t *parents, char *err, size_t errlen);

/* one-shot: data_create + score + data_destroy */
int abn_b200_score_cache(const abn_dataset *ds, const abn_candidates *cand,
                         const abn_control *ctrl, abn_result *out, char *err, size_t errlen);

/* instrumentation of the last abn_b200_score on this handle: device milliseconds by phase (CUDA
 * events on the library's stream), kernel launches issued (bench.py's gpu_launches) and, per fit
 * kernel launched, its duration, fits and algorithmic bytes. */
#define ABN_MAX_KERNELS 64
typedef struct abn_timing {
    double h2d_ms;      /* candidate upload                                                  */
    double plan_ms;     /* design resolution / rank / work-list build                        */
    double fit_ms;      /* all fit kernels                                                   */
    double d2h_ms;      /* result download                                                   */
    int64_t launches;   /* kernels launched                                                  */
    int64_t algo_bytes; /* sum over fits of passes * n * (k_data + 1) * 8         */
    int32_t n_kernels;  /* fit kernels launched by the call, entries used below              */
    int32_t reserved;
    int32_t kernel_id[ABN_MAX_KERNELS];    /* 0 = generic wide kernel; 1 = Gram-only Gaussian kernel; otherwise
                                              2 + ((family_idx * (ABN_FAST_K+1) + k) * 2 + intercept),
                                              family_idx 0 = binomial, 1 = Poisson, L = 2..5:
                                              multinomial child with L levels               */
    int64_t kernel_fits[ABN_MAX_KERNELS];
    int64_t kernel_bytes[ABN_MAX_KERNELS]; /* algorithmic bytes of the fits that kernel ran  */
    double kernel_ms[ABN_MAX_KERNELS];
    int64_t kernel_fp64[ABN_MAX_KERNELS];  /* FP64 lane-instructions that kernel executed in its data passes: per pass
                                              mode (tallied on the device) rows x the FP64 instructions one row costs
                                              in that mode (counted from the SASS, profiles/sass_r2_*): the numerator
                                              of bench.py's roofline_fp64. 0 for kernels without row passes.       */
} abn_timing;
int abn_b200_last_timing(const abn_data *d, abn_timing *t);

#ifdef __cplusplus
}
#endif
#endif /* ABN_B200_H */
