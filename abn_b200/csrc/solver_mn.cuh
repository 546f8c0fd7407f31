// solver_mn.cuh - multinomial child, method = "mle": the state machine of
//   nnet.default(x = X, y = onehot(child), mask, size = 0, skip = TRUE, softmax = TRUE, rang = 0)
// as called at R/build_score_cache_mle.R:319-331, i.e. softmax regression of the L-level child on
// the design X with reference level 1 and no extra bias, minimised by R's vmmin (BFGS, Nash's
// Algorithm 21) from zero weights with maxit = 100, abstol = 1e-4, reltol = 1e-8.
// nnet 7.3-20 (renv.lock) is not in the reference tree; this follows its published algorithm and
// R's src/appl/optim.c:vmmin. PARITY UNPINNED beyond the optimiser's own tolerance (the only
// reference test is equality with nnet::multinom itself, tests/testthat/test-build_score_cache_mle.R:248).
//
// One data pass evaluates the objective "value" = - sum_i log p_{i,y_i} and its gradient
// X'(P - Y) together, so every vmmin function evaluation (initial point, each line-search trial)
// is one pass and the gradient of an accepted point needs none.
#pragma once

#include "solver.cuh"

namespace ABN_NS {

constexpr int MN_MAXL = 5;                       // levels of a multinomial child handled on the GPU
constexpr int MN_MAXP = (MN_MAXL - 1) * MAXK;    // free weights

struct MnAcc {
    double f;           // value at the evaluation point
    double g[MN_MAXP];  // gradient, index (l-1)*k + j
    int bad;
};

enum MnPhase : int { MN_INIT = 1, MN_TRIAL };

struct MnState {
    int k, L, np;
    double n;
    int rank;                 // rank_cpp(X), enters edf
    double b[MN_MAXP];        // evaluation point of the next pass
    // vmmin
    double g[MN_MAXP], t[MN_MAXP], X[MN_MAXP], c[MN_MAXP];
    double B[MN_MAXP * (MN_MAXP + 1) / 2]; // lower triangle, row-major
    double Fmin, f, gradproj, steplength;
    int count, funcount, gradcount, ilast, iter, phase;
    int passes, status;
    double loglik, aic, bic;
};

ABN_HD double &mnB(MnState &s, int i, int j) { return s.B[i * (i + 1) / 2 + j]; } // j <= i

// everything vmmin does between two function evaluations; returns true when another pass (at s.b)
// is needed, false when the optimisation is over.
ABN_HDN bool mn_advance(MnState &s, bool have_trial, const MnAcc *a)
{
    const int n = s.np;
    const int maxit = 100;
    const double abstol = 1.0e-4, reltol = 1.0e-8, stepredn = 0.2, acctol = 0.0001, reltest = 10.0;
    bool resume_trial = have_trial;
    for (;;) {
        if (!resume_trial) {
            // top of vmmin's do-loop
            if (s.ilast == s.gradcount) {
                for (int i = 0; i < n; i++) {
                    for (int j = 0; j < i; j++) mnB(s, i, j) = 0.0;
                    mnB(s, i, i) = 1.0;
                }
            }
            for (int i = 0; i < n; i++) { s.X[i] = s.b[i]; s.c[i] = s.g[i]; }
            s.gradproj = 0.0;
            for (int i = 0; i < n; i++) {
                double sum = 0.0;
                for (int j = 0; j <= i; j++) sum -= mnB(s, i, j) * s.g[j];
                for (int j = i + 1; j < n; j++) sum -= mnB(s, j, i) * s.g[j];
                s.t[i] = sum;
                s.gradproj += sum * s.g[i];
            }
            if (!(s.gradproj < 0.0)) { // uphill search
                s.count = 0;
                if (s.ilast == s.gradcount) s.count = n; else s.ilast = s.gradcount;
                goto loop_end;
            }
            s.steplength = 1.0;
        }
        // line search: do { trial } while (!(count == n || accpoint))
        {
            bool accpoint = false;
            if (resume_trial) {
                // result of the pass at b = X + steplength * t
                resume_trial = false;
                s.f = a->f;
                s.funcount++;
                accpoint = isfinite(s.f) && (s.f <= s.Fmin + s.gradproj * s.steplength * acctol);
                if (!accpoint) s.steplength *= stepredn;
            } else {
                accpoint = false;
            }
            if (!accpoint) {
                s.count = 0;
                for (int i = 0; i < n; i++) {
                    s.b[i] = s.X[i] + s.steplength * s.t[i];
                    if (reltest + s.X[i] == reltest + s.b[i]) s.count++;
                }
                if (s.count < n) { s.phase = MN_TRIAL; return true; } // evaluate f (and g) at b
                // count == n: no change possible, fall through with the old f
            }
            const bool enough = (s.f > abstol) && fabs(s.f - s.Fmin) > reltol * (fabs(s.Fmin) + reltol);
            if (!enough) { s.count = n; s.Fmin = s.f; }
            if (s.count < n) { // making progress: gradient at b is already in the pass result
                s.Fmin = s.f;
                for (int i = 0; i < n; i++) s.g[i] = a->g[i];
                s.gradcount++;
                s.iter++;
                double D1 = 0.0;
                for (int i = 0; i < n; i++) {
                    s.t[i] = s.steplength * s.t[i];
                    s.c[i] = s.g[i] - s.c[i];
                    D1 += s.t[i] * s.c[i];
                }
                if (D1 > 0) {
                    double D2 = 0.0;
                    for (int i = 0; i < n; i++) {
                        double sum = 0.0;
                        for (int j = 0; j <= i; j++) sum += mnB(s, i, j) * s.c[j];
                        for (int j = i + 1; j < n; j++) sum += mnB(s, j, i) * s.c[j];
                        s.X[i] = sum;
                        D2 += sum * s.c[i];
                    }
                    D2 = 1.0 + D2 / D1;
                    for (int i = 0; i < n; i++)
                        for (int j = 0; j <= i; j++)
                            mnB(s, i, j) += (D2 * s.t[i] * s.t[j] - s.X[i] * s.c[j] - s.t[i] * s.X[j]) / D1;
                } else {
                    s.ilast = s.gradcount;
                }
            } else { // no progress
                if (s.ilast < s.gradcount) { s.count = 0; s.ilast = s.gradcount; }
            }
        }
    loop_end:
        if (s.iter >= maxit) break;
        if (s.gradcount - s.ilast > 2 * n) s.ilast = s.gradcount;
        if (!(s.count != n || s.ilast != s.gradcount)) break;
    }
    return false;
}

ABN_HDN void mn_finish(MnState &s)
{
    const double value = s.Fmin;
    const double edf = -1.0 * s.rank; // R/build_score_cache_mle.R:329: length(tmp$lev) is 0 for nnet.default
    s.loglik = -value;                                  // :328
    s.aic = 2 * value + 2 * edf;                        // :330
    s.bic = 2 * value + edf * log(s.n);                 // :331
    s.status = !isfinite(value) ? FIT_NON_FINITE : (s.iter < 100 ? FIT_OK : FIT_NOT_CONVERGED);
}

// returns true when a pass at s.b is needed
ABN_HDN bool mn_begin(MnState &s)
{
    s.np = (s.L - 1) * s.k;
    for (int i = 0; i < s.np; i++) s.b[i] = 0.0; // rang = 0: all weights start at zero
    s.passes = 0;
    s.status = FIT_OK;
    s.phase = MN_INIT;
    return true;
}

ABN_HDN bool mn_consume(MnState &s, const MnAcc &a)
{
    s.passes++;
    if (s.phase == MN_INIT) {
        // f = fminfn(b); funcount = gradcount = 1; fmingr(b, g); iter++; ilast = gradcount
        s.f = a.f;
        if (!isfinite(s.f)) { s.Fmin = s.f; mn_finish(s); s.status = FIT_NON_FINITE; return false; }
        s.Fmin = s.f;
        s.funcount = s.gradcount = 1;
        for (int i = 0; i < s.np; i++) s.g[i] = a.g[i];
        s.iter = 1;
        s.ilast = s.gradcount;
        s.count = 0;
        if (mn_advance(s, false, nullptr)) return true;
        mn_finish(s);
        return false;
    }
    if (mn_advance(s, true, &a)) return true;
    mn_finish(s);
    return false;
}

} // namespace ABN_NS
