// rowops.cuh - what one observation row contributes to a pass: (w, r, l) from (eta, y).
//
//   H += w x x^T,   g += r x,   ll += l
//
// One struct per (family, method); all FP64. Reference formulas (paths relative to the
// reference tree) are cited per op. __host__ __device__ so tests/host_sim.cpp reuses them.
// Each op is templated on what the pass needs (NEED_W: Hessian weight, NEED_R: score residual,
// NEED_L: log-likelihood term) so a gradient-only or loglik-only pass does not pay for the rest.
#pragma once

#include "fastmath.cuh"
#include "solver.cuh"

namespace ABN_NS {

enum RowKind : int {
    RK_BIN_MLE = 0,   // src/irls_binomial_fast_br.cpp:53-66,97
    RK_POIS = 1,      // src/irls_poisson_fast.cpp:62-74 and src/node_poisson.c:545-667
    RK_BIN_BAYES = 2, // src/node_binomial.c:626-640,690-700,759-770
    RK_PINIT = 3      // src/irls_poisson_fast.cpp:37-40
};

struct RowConst {
    double bp0, bp1;   // binomial MLE bias-reduced pseudo-responses for y = 0 / y = 1 (b' = bp0 + y (bp1 - bp0))
    const double *tab; // 2^(j/32) table for fast_exp (shared memory on the device)
    const double *lnmu; // log(y + 0.1), y = 0..63, for the Poisson start values (may be null)
    const double *logtab; // fast_log_ge1 table
};

// b' = (y + 0.5 k/n) / (1 + k/n), irls_binomial_fast_br.cpp:58; callers pass k = 0 for the plain b' = y of
// irls_binomial_fast.cpp (ABN_FLAG_NO_BIAS_REDUCTION)
ABN_HD RowConst make_row_const(int k, double n, const double *tab, const double *logtab, const double *lnmu = nullptr)
{
    RowConst c;
    double dfn = (double)k / n;
    c.bp0 = (0.0 + dfn * 0.5) / (1 + dfn);
    c.bp1 = (1.0 + dfn * 0.5) / (1 + dfn);
    c.tab = tab;
    c.lnmu = lnmu;
    c.logtab = logtab;
    return c;
}

constexpr double BIN_RHO_ETA = 12.0;

// Range bookkeeping of one data pass, kept on the HIGH WORD of eta with integer max (fastmath.cuh):
//   smax  signed max   - positive doubles order like their high words, negatives sort below them
//   umax  unsigned max - negative doubles (sign bit set) sort above every positive, by magnitude
// so "some eta >= a" is smax >= hi(a) for a > 0 and "some eta <= -b" is umax >= hi(-b); NaNs of
// either sign exceed every finite bound on one of the two. The bounds are compared on the high
// word only: a fit is flagged when eta comes within 2^-20 relative of the exact limit (the limits
// are where the reference's arithmetic overflows; nothing meaningful lives that close to them).
struct RowTrack {
    int smax;
    unsigned umax;
    int flag;
};

ABN_HD RowTrack row_track_init()
{
    RowTrack t;
    t.smax = (int)0x80000000u;
    t.umax = 0u;
    t.flag = 0;
    return t;
}

ABN_HD void row_track(RowTrack &t, double eta)
{
    const int h = hi_word(eta);
    t.smax = t.smax > h ? t.smax : h;
    t.umax = t.umax > (unsigned)h ? t.umax : (unsigned)h;
}

// eta clamped to [-700.0005, 700.0005] (the low word is kept: the bound moves by < 2^-11), NaN -> +-700
ABN_HD double clamp700(double t)
{
    const unsigned h = (unsigned)hi_word(t);
    unsigned a = h & 0x7fffffffu;
    a = a < 0x4085E000u ? a : 0x4085E000u; // hi(700.0)
    return with_hi_word(t, (int)(a | (h & 0x80000000u)));
}

// eta clamped from below only, same convention; positive values (and +NaN) pass through
ABN_HD double clamp_lo700(double t)
{
    const unsigned h = (unsigned)hi_word(t);
    return with_hi_word(t, (int)(h < 0xC085E000u ? h : 0xC085E000u)); // hi(-700.0)
}

template <int KIND>
struct RowOp;

template <>
struct RowOp<RK_BIN_MLE> {
    // e = exp(-eta); g = 1/(1+e); g' = e/(1+e)^2; W = g'^2/(g(1-g)); z = eta + (b' - g)/g'.
    // In exact arithmetic W == g', but the reference forms 1 - g in floating point, and for
    // saturated rows (|eta| > ~20) the rounding of g makes rho = W/g' differ from 1 by up to
    // 2^-53/(1-g). Because the bias-reduced pseudo-response keeps (b' - g) ~ k/2n on those rows,
    // that noise moves the fixed point of the reference's IRLS at the 1e-8 level; parity needs
    // the same rho. So: W = rho g', W z = W eta + rho (b' - g), i.e. the Newton form
    //   x_new = x + (X'WX)^-1 X'[rho (b' - g)]  ==  solve(X'WX, X'Wz)  (irls_binomial_fast_br.cpp:72-73)
    // with g = fl(1/(1+e)) correctly rounded like the reference's division.
    // loglik term: -y log(1+e^-eta) - (1-y) log(1+e^eta) = -log(1+e) - (1-y) eta   (:97)
    //
    // rho differs from 1 only through the rounding of g in 1 - g: |rho - 1| <= 2^-53 e^eta. The
    // kernels evaluate rho only for warp steps in which some row has eta >= BIN_RHO_ETA = 12 (a
    // warp vote); below that |rho - 1| < 2e-11 moves the coefficients by < 1e-10 of their scale,
    // three orders under the IRLS stop tolerance, and the score by ~1e-20 - while the saturated
    // rows, where the reference's noise is visible at 1e-8, always take the exact path.
    template <bool NEED_W, bool NEED_R, bool NEED_L>
    static ABN_HD void apply(double eta, double y, const RowConst &c, double &w, double &r, double &l, RowTrack &tk, bool need_rho = true)
    {
        // The reference's W or z is non-finite when g rounds to 1 (g(1-g) = 0: 1 + e == 1, i.e.
        // eta >= 53 ln 2) or when (1+e)^2 overflows (e >= sqrt(DBL_MAX), eta <= -354.89). Outside
        // that window the fit is flagged and whatever fast_exp returns is never used; inside it
        // fast_exp needs no range clamp.
        row_track(tk, eta);
        const double e = fast_exp(-eta, c.tab);
        const double d = 1.0 + e;
        if (NEED_W || NEED_R) {
            const double g = fast_rcp(d);
            const double gp = e * g * g;
            const double bpg = fma(y, c.bp1 - c.bp0, c.bp0) - g;
            if (need_rho) {
                const double rho = gp * fast_rcp_approx(g * (1.0 - g));
                w = rho * gp;
                r = rho * bpg;
            } else { // every row of this warp step has eta < BIN_RHO_ETA: |rho - 1| < 2e-11, see below
                w = gp;
                r = bpg;
            }
        }
        if (NEED_L) l = -fast_log_ge1(d, c.logtab) - (1.0 - y) * eta;
    }
    // eta >= 53 ln 2 = 36.7368005696771 (0x40425E4F 7B2737FA) or eta <= -354.891356446692 (0xC0762E42 ...)
    static ABN_HD int bad_of(const RowTrack &tk) { return (tk.smax >= 0x40425E4F) | (tk.umax >= 0xC0762E42u) | tk.flag; }
};

template <>
struct RowOp<RK_POIS> {
    template <bool NEED_W, bool NEED_R, bool NEED_L>
    static ABN_HD void apply(double eta, double y, const RowConst &c, double &w, double &r, double &l, RowTrack &tk, bool need_rho = true)
    {
        row_track(tk, eta);
        const double E = fast_exp(clamp_lo700(eta), c.tab); // flagged fits never use E; no upper clamp
        w = E;
        r = y - E;
        if (NEED_L) l = fma(y, eta, -E);
    }
    // exp(eta) = +Inf in the reference from eta >= 709.782712893384 (0x40862E42 FEFA39EF); NaN either sign
    static ABN_HD int bad_of(const RowTrack &tk) { return (tk.smax >= 0x40862E42) | (tk.umax > 0xFFF00000u) | tk.flag; }
};

template <>
struct RowOp<RK_BIN_BAYES> {
    // E = exp(eta): sigma = E/(1+E) (1 if E overflows, node_binomial.c:697-700),
    // hessian weight exp(eta - 2 log(1+E)) = sigma/(1+E) (:761-764), log term log(1+E) (eta if E
    // overflows, :633-639). For |eta| > 700 the clamped E gives the same doubles as those branches.
    template <bool NEED_W, bool NEED_R, bool NEED_L>
    static ABN_HD void apply(double eta, double y, const RowConst &c, double &w, double &r, double &l, RowTrack &tk, bool need_rho = true)
    {
        const double E = fast_exp(clamp700(eta), c.tab);
        const double d = 1.0 + E;
        if (NEED_W || NEED_R) {
            const double inv = fast_rcp_approx(d);
            const double sig = E * inv;
            w = sig * inv;
            r = y - sig;
        }
        if (NEED_L) l = (hi_word(eta) >= 0x4085E000) ? (y * eta - eta) : (y * eta - fast_log_ge1(d, c.logtab)); // eta >= 700
        (void)tk;
        (void)need_rho;
    }
    static ABN_HD int bad_of(const RowTrack &tk) { return tk.flag; }
};

template <>
struct RowOp<RK_PINIT> {
    // mu = y + 0.1; eta0 = log(mu); W0 = exp(eta0) (= mu up to an ulp); z0 = eta0 + (y - W0)/W0;
    // H = X'W0X, g = X'W0z0 with W0 z0 = W0 eta0 + (y - W0): no division needed. Counts are small
    // integers, so log(y + 0.1) comes from a 64-entry table; other values take the libm path.
    template <bool NEED_W, bool NEED_R, bool NEED_L>
    static ABN_HD void apply(double, double y, const RowConst &c, double &w, double &r, double &l, RowTrack &tk, bool need_rho = true)
    {
        const double mu = y + 0.1;
        const int yi = (int)y;
        double eta0;
        if (c.lnmu != nullptr && y == (double)yi && yi >= 0 && yi < 64) eta0 = c.lnmu[yi];
        else eta0 = log(mu);
        w = mu;
        r = fma(mu, eta0, y - mu);
        tk.flag |= !(mu > 0.0) | !(mu < 1.79769313486231570815e308);
        if (NEED_L) l = 0.0;
    }
    static ABN_HD int bad_of(const RowTrack &tk) { return tk.flag; }
};

} // namespace ABN_NS
