// host_sim.cpp - TEST INFRASTRUCTURE. Drives the product's __host__ __device__ solver state
// machines (abn_b200/csrc/solver.cuh, rowops.cuh, plan.cuh) on the CPU with plain sequential
// row loops standing in for the CUDA data passes. It lets the -m "not gpu" suite check the
// exact control flow the kernels execute (step control, stop rules, flags, Gram-based start
// values and rank handling) against the oracle and the reference's golden caches in a
// container without a GPU. It is not part of libabn_b200.so and nothing in abn_b200/ loads it.
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../abn_b200/csrc/plan.cuh"
#include "../abn_b200/csrc/rowops.cuh"
#include "../abn_b200/csrc/solver.cuh"
#include "../abn_b200/csrc/rowops_mn.cuh"
#include "../abn_b200/csrc/solver_mn.cuh"
#include "../include/abn_b200.h"

using namespace ABN_NS;

namespace {

const double *host_log_table()
{
    static double tab[128];
    static bool done = false;
    if (!done) { for (int i = 0; i < 64; i++) fill_log_table(tab, i); done = true; }
    return tab;
}

template <int KIND>
void host_pass(const DataMeta &m, const FitDesc &d, const double *beta, int mode, PassAcc &a, bool bias_reduction = true)
{
    const int k = d.k;
    std::memset(&a, 0, sizeof(a));
    RowConst rc = make_row_const(bias_reduction ? k : 0, (double)m.n, h_exp2_table, host_log_table());
    const double *y = m.X + (size_t)d.ycol * m.ld;
    const bool need_h = (mode == PASS_FULL || mode == PASS_HG || mode == PASS_HESS);
    const bool need_g = (mode == PASS_FULL || mode == PASS_HG || mode == PASS_GRAD);
    const bool need_l = (mode == PASS_FULL || mode == PASS_LL);
    double x[MAXK];
    RowTrack tk = row_track_init();
    bool any_rho = false;
    for (long long i = 0; i < m.n; i++) {
        double eta = 0.0;
        for (int j = 0; j < k; j++) {
            x[j] = d.cols[j] < 0 ? 1.0 : m.X[(size_t)d.cols[j] * m.ld + i];
            eta = std::fma(x[j], beta[j], eta);
        }
        double w = 0.0, r = 0.0, l = 0.0;
        // the kernels decide per warp step (vote) whether rho is evaluated; row by row here
        const bool need_rho = eta >= BIN_RHO_ETA;
        any_rho |= need_rho;
        if (need_l && (need_h || need_g)) RowOp<KIND>::template apply<true, true, true>(eta, y[i], rc, w, r, l, tk, need_rho);
        else if (need_l) RowOp<KIND>::template apply<false, false, true>(eta, y[i], rc, w, r, l, tk, need_rho);
        else RowOp<KIND>::template apply<true, true, false>(eta, y[i], rc, w, r, l, tk, need_rho);
        if (need_h)
            for (int p = 0; p < k; p++)
                for (int q = p; q < k; q++) a.H[tri(p, q, k)] += w * x[p] * x[q];
        if (need_g)
            for (int p = 0; p < k; p++) a.g[p] += r * x[p];
        if (need_g && KIND == RK_BIN_MLE) { // X'(y - g) for the Taylor finish (solver.cuh "IRLS shortcuts")
            const double gi = 1.0 / (1.0 + std::exp(-eta));
            for (int p = 0; p < k; p++) a.gl[p] += (y[i] - gi) * x[p];
        }
        if (need_l) a.ll += l;
    }
    a.bad = RowOp<KIND>::bad_of(tk);
    a.has_gl = (need_g && KIND == RK_BIN_MLE && !any_rho) ? 1 : 0;
}

void run_pass(const DataMeta &m, const FitDesc &d, const Control &c, const FitState &s, int mode, PassAcc &a)
{
#if defined(ABN_TRACE_STEPS)
    fprintf(stderr, "PASS fam %d mode %d\n", d.family, mode);
#endif
    if (mode == PASS_PINIT) { host_pass<RK_PINIT>(m, d, s.x, PASS_HG, a); return; }
    if (d.family == POISSON) { host_pass<RK_POIS>(m, d, s.x, mode, a); return; }
    if (c.method == MLE) host_pass<RK_BIN_MLE>(m, d, s.x, mode, a, !(c.flags & 2));
    else host_pass<RK_BIN_BAYES>(m, d, s.x, mode, a);
}

// multinomial child: value + gradient pass, dispatched to the same mn_row_lean / mn_row<L, K, ICPT> the kernel uses
// (fit_mn_kernel.cuh: the lean pass first; the careful one for the rest of the fit once a linear predictor has left
// (-700, 700), and always under ABN_FLAG_EXACT_PATH)
static const double *host_exp256_table()
{
    static double tab[ABN_EXP256_N];
    static bool done = false;
    if (!done) { for (int j = 0; j < ABN_EXP256_N; j++) tab[j] = exp256_entry(h_exp2_256_table[j], j); done = true; }
    return tab;
}

template <int L, int K, bool ICPT>
void mn_host_pass(const DataMeta &m, const FitDesc &d, const double *w, MnAcc &a, bool &careful)
{
    constexpr int NP = (L - 1) * K;
    constexpr int KD = K - (ICPT ? 1 : 0);
    double W[NP], g[NP], f = 0.0;
    for (int i = 0; i < NP; i++) { W[i] = w[i]; g[i] = 0.0; }
    const double *y0 = m.X + (size_t)d.ycol * m.ld;
    if (!careful) {
        bool in_range = true;
        for (long long i = 0; i < m.n && in_range; i++) {
            double xd[KD > 0 ? KD : 1], o[L - 1];
            for (int j = 0; j < KD; j++) xd[j] = m.X[(size_t)d.cols[j + (ICPT ? 1 : 0)] * m.ld + i];
            mn_row_lean<L, K, ICPT>(xd, W, host_exp256_table(), host_log_table(), f, g, o);
            for (int l = 0; l < L - 1; l++) in_range &= std::fabs(o[l]) < 700.0; // false for NaN
        }
        if (in_range) {
            double xty[NP];
            for (int l = 1; l < L; l++)
                for (int j = 0; j < K; j++) xty[(l - 1) * K + j] = m.G[(size_t)gidx(d, j) * m.D + gram_col(d.ycol + l)];
            mn_lean_finish(NP, f, g, W, xty, &a.f, a.g);
            a.bad = 0;
            return;
        }
        careful = true;
        f = 0.0;
        for (int i = 0; i < NP; i++) g[i] = 0.0;
    }
    for (long long i = 0; i < m.n; i++) {
        double x[K], yl[L - 1];
        for (int j = 0; j < K; j++) x[j] = d.cols[j] < 0 ? 1.0 : m.X[(size_t)d.cols[j] * m.ld + i];
        for (int l = 1; l < L; l++) yl[l - 1] = y0[(size_t)l * m.ld + i];
        mn_row<L, K, ICPT>(x, yl, W, h_exp2_table, host_log_table(), f, g);
    }
    a.f = f;
    for (int i = 0; i < NP; i++) a.g[i] = g[i];
    a.bad = 0;
}

template <int L, int K>
void mn_pass_k(const DataMeta &m, const FitDesc &d, const double *w, MnAcc &a, bool &careful)
{
    if (d.cols[0] < 0) mn_host_pass<L, K, true>(m, d, w, a, careful); else mn_host_pass<L, K, false>(m, d, w, a, careful);
}

template <int L>
void mn_pass_l(const DataMeta &m, const FitDesc &d, const double *w, MnAcc &a, bool &careful)
{
    switch (d.k) {
    case 1: mn_pass_k<L, 1>(m, d, w, a, careful); break;
    case 2: mn_pass_k<L, 2>(m, d, w, a, careful); break;
    case 3: mn_pass_k<L, 3>(m, d, w, a, careful); break;
    case 4: mn_pass_k<L, 4>(m, d, w, a, careful); break;
    case 5: mn_pass_k<L, 5>(m, d, w, a, careful); break;
    case 6: mn_pass_k<L, 6>(m, d, w, a, careful); break;
    case 7: mn_pass_k<L, 7>(m, d, w, a, careful); break;
    case 8: mn_pass_k<L, 8>(m, d, w, a, careful); break;
    }
}

void mn_pass(const DataMeta &m, const FitDesc &d, const double *w, MnAcc &a, bool &careful)
{
    switch (d.nlev) {
    case 2: mn_pass_l<2>(m, d, w, a, careful); break;
    case 3: mn_pass_l<3>(m, d, w, a, careful); break;
    case 4: mn_pass_l<4>(m, d, w, a, careful); break;
    case 5: mn_pass_l<5>(m, d, w, a, careful); break;
    }
}

} // namespace

extern "C" int sim_score_cache(const abn_dataset *ds, const abn_candidates *cand, const abn_control *ctrl,
                               abn_result *out, int *buckets /* optional [rows] */)
{
    host_log_table(); // initialise before the parallel region
    const long long n = ds->n;
    int n_pois = 0;
    std::vector<int> gw_index(ds->n_nodes, -1), pois_node;
    for (int v = 0; v < ds->n_nodes; v++) if (ds->node_family[v] == ABN_POISSON) { gw_index[v] = n_pois++; pois_node.push_back(v); }
    const int D = 1 + ds->n_cols + n_pois;
    // Z = [1 | X | log(y + 0.1) of Poisson nodes], G = Z'Z
    std::vector<double> G((size_t)D * D, 0.0), lfact(ds->n_nodes, 0.0);
    std::vector<const double *> zc(D, nullptr);
    std::vector<std::vector<double>> extra;
    std::vector<double> ones(n, 1.0);
    zc[0] = ones.data();
    for (int c = 0; c < ds->n_cols; c++) zc[1 + c] = ds->X + (size_t)c * n;
    extra.resize(2 * ds->n_nodes);
    for (int v = 0; v < ds->n_nodes; v++) {
        if (ds->node_family[v] != ABN_POISSON) continue;
        const double *y = ds->X + (size_t)ds->node_col0[v] * n;
        extra[v].resize(n);
        for (long long i = 0; i < n; i++) {
            const double mu = y[i] + 0.1, eta0 = std::log(mu);
            extra[v][i] = eta0;
            lfact[v] += std::lgamma(y[i] + 1.0);
        }
        zc[1 + ds->n_cols + gw_index[v]] = extra[v].data();
    }
    for (int a = 0; a < D; a++) {
        if (!zc[a]) continue;
        for (int b = a; b < D; b++) {
            if (!zc[b]) continue;
            double s = 0.0;
            for (long long i = 0; i < n; i++) s += zc[a][i] * zc[b][i];
            G[(size_t)a * D + b] = G[(size_t)b * D + a] = s;
        }
    }
    DataMeta m{n, n, ds->n_nodes, ds->n_cols, D, ds->X, G.data(), ds->node_family, ds->node_col0, lfact.data(),
               gw_index.data(), n_pois, pois_node.data()};
    // centred second moments for the Gaussian children, as abn_b200_data_create builds them (types.cuh DataMeta::Gc)
    const int Dc = 1 + ds->n_cols;
    std::vector<double> shift(Dc, 0.0), Gc((size_t)Dc * Dc, 0.0);
    for (int a = 1; a < Dc; a++) shift[a] = G[a] / (double)n;
    for (int a = 0; a < Dc; a++)
        for (int b = a; b < Dc; b++) {
            double s = 0.0;
            for (long long i = 0; i < n; i++) s += (zc[a][i] - shift[a]) * (zc[b][i] - shift[b]);
            Gc[(size_t)a * Dc + b] = Gc[(size_t)b * Dc + a] = s;
        }
    if (!std::getenv("ABN_SIM_NO_GC")) m.Gc = Gc.data();
    m.col_mean = shift.data() + 1;
    Control c{ctrl->method, ctrl->max_iters, ctrl->solver, ctrl->nvars, ctrl->tol, ctrl->epsabs,
              ctrl->prior_mean, ctrl->prior_sd, ctrl->gam_shape, ctrl->gam_scale, ctrl->flags};

#pragma omp parallel for schedule(dynamic, 8)
    for (long long r = 0; r < cand->rows; r++) {
        FitDesc d;
        const int child = cand->child[r];
        resolve_design(m, c.method, child, cand->parents + r * cand->max_parents, cand->max_parents, d);
        if (buckets) buckets[r] = bucket_of(d);
        double ll = NAN, aic = NAN, bic = NAN, mdl = NAN;
        int status = FIT_NON_FINITE, passes = 0;
        ResultPtrs rp{nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, out->coef, out->coef ? out->coef_stride : 0,
                      out->coef ? out->se : nullptr, out->sse};
        if (rp.sse) rp.sse[r] = NAN;
        if (rp.coef) for (int j = 0; j < rp.coef_stride; j++) rp.coef[r * rp.coef_stride + j] = NAN;
        if (rp.se) for (int j = 0; j < rp.coef_stride; j++) rp.se[r * rp.coef_stride + j] = NAN;
        if (d.err || d.k == 0 || (d.family == MULTINOMIAL && (d.nlev < 2 || d.nlev > MN_MAXL || c.method != MLE || d.k > 8))) {
            // not supported by this build
        } else if (d.family == MULTINOMIAL) {
            MnState s;
            std::memset(&s, 0, sizeof(s));
            s.k = d.k; s.L = d.nlev; s.n = (double)m.n; s.rank = d.rank;
            MnAcc a;
            bool more = mn_begin(s);
            bool careful = (ctrl->flags & ABN_FLAG_EXACT_PATH) != 0;
            while (more) { mn_pass(m, d, s.b, a, careful); more = mn_consume(s, a); }
            ll = s.loglik; aic = s.aic; bic = s.bic; status = s.status; passes = s.passes;
        } else if (c.method == MLE && d.family == GAUSSIAN) {
            double x[MAXM], se[MAXM];
            double sse;
            gauss_mle_fit(m, d, &ll, &aic, &bic, &status, x, rp.se ? se : nullptr, &sse);
            write_coef(rp, r, d, x, d.k);
            if (rp.se) { ResultPtrs o = rp; o.coef = rp.se; write_coef(o, r, d, se, d.k); }
            if (rp.sse) rp.sse[r] = sse;
            passes = 1;
        } else {
            FitState s;
            std::memset(&s, 0, sizeof(s));
            state_setup(m, d, child, c, s);
            GaussAux aux;
            if (d.family == GAUSSIAN && gauss_centered_block(m, d, aux)) s.gaux = &aux; // as gauss_kernel
            if (c.method == BAYES) bayes_inits(m, d, child, s);
            PassAcc a;
            int mode = fit_begin(s, c);
            if (c.method == MLE && d.family == BINOMIAL) { // as fit_kernel: iteration 1 from the Gram matrix
                binomial_first_acc(m, d, a, !(c.flags & 2));
                s.have_ll = 1;
                mode = fit_consume(s, c, a);
                s.passes--;
            }
            if (c.method == MLE && d.family == POISSON && mode == PASS_PINIT && !(c.flags & 1)) { // Gram-only start values
                poisson_init_acc(m, d, child, a);
                mode = fit_consume(s, c, a);
                s.passes--;
            }
            while (mode != PASS_DONE) {
                if (d.family == GAUSSIAN) gauss_gram_acc(m, d, a);
                else run_pass(m, d, c, s, mode, a);
                mode = fit_consume(s, c, a);
            }
            ll = s.loglik; aic = s.aic; bic = s.bic; status = s.status;
            passes = d.family == GAUSSIAN ? 1 : s.passes;
            write_coef(rp, r, d, s.x, s.M);
            write_se(rp, r, d, s, c.method);
            if (rp.sse && c.method == MLE) rp.sse[r] = sse_of(m, d, s.x);
        }
        if (c.method == MLE) mdl = mdl_of(bic, d, c.nvars);
        if (out->mlik) out->mlik[r] = ll;
        if (out->aic) out->aic[r] = aic;
        if (out->bic) out->bic[r] = bic;
        if (out->mdl) out->mdl[r] = mdl;
        if (out->status) out->status[r] = status;
        if (out->iters) out->iters[r] = passes;
    }
    return 0;
}

// fast FP64 math of the kernels (fastmath.cuh), host build, for tests/test_fastmath.py
extern "C" void sim_fast_log(const double *x, double *out, long n)
{
    for (long i = 0; i < n; i++) out[i] = fast_log_ge1(x[i], host_log_table());
}

extern "C" void sim_fast_exp(const double *t, double *out, long n)
{
    for (long i = 0; i < n; i++) out[i] = fast_exp(t[i], h_exp2_table);
}

// the lean passes' exp, replicated-table form (fastmath.cuh fast_exp64r); lane 5's replica
extern "C" void sim_fast_exp64r(const double *t, double *out, long n)
{
    static double tab[ABN_EXP2_N * ABN_EXP64R_COPIES];
    static bool done = false;
    if (!done) { for (int i = 0; i < ABN_EXP2_N * ABN_EXP64R_COPIES; i++) tab[i] = exp64r_entry(h_exp2_table[i / ABN_EXP64R_COPIES], i / ABN_EXP64R_COPIES); done = true; }
    for (long i = 0; i < n; i++) out[i] = fast_exp64r(t[i], tab + 5);
}

// the lean passes' exp: 256-entry pre-scaled table (fastmath.cuh fast_exp256)
extern "C" void sim_fast_exp256(const double *t, double *out, long n)
{
    static double tab[ABN_EXP256_N];
    static bool done = false;
    if (!done) { for (int j = 0; j < ABN_EXP256_N; j++) tab[j] = exp256_entry(h_exp2_256_table[j], j); done = true; }
    for (long i = 0; i < n; i++) out[i] = fast_exp256(t[i], tab);
}

// ... and its single-reduction-step form (fast_exp256s)
extern "C" void sim_fast_exp256s(const double *t, double *out, long n)
{
    static double tab[ABN_EXP256_N];
    static bool done = false;
    if (!done) { for (int j = 0; j < ABN_EXP256_N; j++) tab[j] = exp256_entry(h_exp2_256_table[j], j); done = true; }
    for (long i = 0; i < n; i++) out[i] = fast_exp256s(t[i], tab);
}
