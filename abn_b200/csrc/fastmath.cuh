// fastmath.cuh - branch-free FP64 exp / reciprocal for the streaming row ops.
//
// The fit kernels are bound by instruction issue and the FP64 pipe, not by HBM (ncu: profiles/):
// libdevice's exp() / log() / IEEE division cost ~35 / ~45 / ~25 instructions each including
// range branches and slow paths that these loops never need. The versions here cost
//   exp   10 FP64 + 5 integer/LDS instructions   (max error 1.01 ulp, mean 0.33 ulp on |t| <= 700:
//                                                 tests/test_fastmath.py against long double)
//   rcp    1 MUFU + 6 FP64                        (correctly rounded 1/d for normal d: Markstein)
// __host__ __device__ with identical arithmetic (FMA on both sides) so tests/host_sim.cpp
// exercises the same bits the kernels produce, apart from the device's MUFU seed (the Newton
// steps converge to the same correctly rounded result).
#pragma once

#include <math.h>
#include <stdint.h>
#include <string.h>

#include "solver.cuh"

namespace abn {

// 2^(j/32), j = 0..31, correctly rounded (generated with 60-digit decimal arithmetic)
#define ABN_EXP2_TABLE                                                                              \
    {1.0, 1.0218971486541166, 1.0442737824274138, 1.0671404006768237, 1.0905077326652577,            \
     1.1143867425958924, 1.1387886347566916, 1.1637248587775775, 1.189207115002721,                  \
     1.215247359980469, 1.241857812073484, 1.2690509571917332, 1.2968395546510096,                   \
     1.3252366431597413, 1.3542555469368927, 1.383909881963832, 1.4142135623730951,                  \
     1.4451808069770467, 1.4768261459394993, 1.5091644275934228, 1.5422108254079407,                 \
     1.5759808451078865, 1.6104903319492543, 1.645755478153965, 1.681792830507429,                   \
     1.718619298122478, 1.7562521603732995, 1.7947090750031072, 1.8340080864093424,                  \
     1.8741676341103, 1.9152065613971474, 1.9571441241754002}

#if defined(__CUDACC__)
__constant__ double c_exp2_table[32] = ABN_EXP2_TABLE;
#endif
static const double h_exp2_table[32] = ABN_EXP2_TABLE;

ABN_HD double bits_to_double(uint64_t b)
{
#if defined(__CUDA_ARCH__)
    return __longlong_as_double((long long)b);
#else
    double d;
    memcpy(&d, &b, 8);
    return d;
#endif
}

ABN_HD uint64_t double_to_bits(double d)
{
#if defined(__CUDA_ARCH__)
    return (uint64_t)__double_as_longlong(d);
#else
    uint64_t b;
    memcpy(&b, &d, 8);
    return b;
#endif
}

// exp(t) for |t| <= 700 (callers clamp). `tab` points at the 32-entry table: shared memory on the
// device (lane-dependent index: a __constant__ read would serialise), h_exp2_table on the host.
//   t = (32 k + j) ln2/32 + r,  |r| <= ln2/64;  exp(t) = 2^k * 2^(j/32) * (1 + p(r))
ABN_HD double fast_exp(double t, const double *tab)
{
    const double MAGIC = 6755399441055744.0;        // 1.5 * 2^52: round-to-nearest-integer trick
    const double INV = 46.16624130844683;           // 32 / ln 2
    const double LN2_32_HI = 0.02166084938653512;   // ln2/32, low 21 mantissa bits zero: n * HI exact
    const double LN2_32_LO = 5.9631716539705866e-12;
    const double kd = fma(t, INV, MAGIC);
    const double nf = kd - MAGIC;                   // n = 32 k + j as a double
    double r = fma(nf, -LN2_32_HI, t);
    r = fma(nf, -LN2_32_LO, r);
    const int64_t n = (int64_t)(int32_t)(uint32_t)double_to_bits(kd); // low word of kd holds n
    const int j = (int)(n & 31);
    const int64_t k = n >> 5;
    // p(r) = r + r^2/2 + r^3/6 + r^4/24 + r^5/120 + r^6/720, |r| <= 0.0109: truncation < 4e-18
    double p = fma(r, 1.0 / 720.0, 1.0 / 120.0);
    p = fma(p, r, 1.0 / 24.0);
    p = fma(p, r, 1.0 / 6.0);
    p = fma(p, r, 0.5);
    p = fma(p * r, r, r);
    const double tj = tab[j];
    const double v = fma(tj, p, tj);                // in [1, 2) * (1 +- 0.011)
    return bits_to_double(double_to_bits(v) + ((uint64_t)k << 52)); // * 2^k, k in [-1010, 1010]
}

// 1 / d, correctly rounded for normal d with a normal result (Markstein: two Newton steps to
// within an ulp, then r = fma(-d, x, 1), x + x r rounds correctly unless d's mantissa is all ones)
ABN_HD double fast_rcp(double d)
{
#if defined(__CUDA_ARCH__)
    double x;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(d)); // MUFU.RCP64H, ~2^-20
    double e = fma(-d, x, 1.0);
    e = fma(e, e, e);               // (e + e^2): one step with cubic convergence  -> ~2^-60
    x = fma(x, e, x);
    e = fma(-d, x, 1.0);
    x = fma(x, e, x);
    e = fma(-d, x, 1.0);
    return fma(x, e, x);
#else
    return 1.0 / d;
#endif
}

// 1 / d to ~1 ulp (not necessarily correctly rounded): one Newton step less
ABN_HD double fast_rcp_approx(double d)
{
#if defined(__CUDA_ARCH__)
    double x;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(d));
    double e = fma(-d, x, 1.0);
    e = fma(e, e, e);
    x = fma(x, e, x);
    e = fma(-d, x, 1.0);
    return fma(x, e, x);
#else
    return 1.0 / d;
#endif
}

// log(x) for finite x >= 1 (every use in the row ops: 1 + e, 1 + E, a softmax denominator).
//   x = 2^e m, m in [1,2), i = top 6 mantissa bits, c_i = 1/(1 + i/64), r = m c_i - 1 in [0, 1/64)
//   log x = e ln2 + log(1 + i/64) + log1p(r)
// `tab` holds {c_i, log(1 + i/64)} pairs (64 x 2 doubles; shared memory on the device). Absolute
// error < 3e-16, relative error ~1 ulp away from x = 1 and exact-ish near it (c_0 = 1, r = m - 1).
ABN_HD void fill_log_table(double *tab, int i)
{
    const double mi = 1.0 + (double)i / 64.0;
    tab[2 * i] = 1.0 / mi;
    tab[2 * i + 1] = log(mi);
}

ABN_HD double fast_log_ge1(double x, const double *tab)
{
    const double LN2_HI = 6.93147180369123816490e-01, LN2_LO = 1.90821492927058770002e-10;
    const uint64_t b = double_to_bits(x);
    const int e = (int)(b >> 52) - 1023;
    const int i = (int)((b >> 46) & 63);
    const double m = bits_to_double((b & 0x000FFFFFFFFFFFFFull) | 0x3FF0000000000000ull);
    const double r = fma(m, tab[2 * i], -1.0);
    double p = fma(r, -1.0 / 8.0, 1.0 / 7.0);
    p = fma(p, r, -1.0 / 6.0);
    p = fma(p, r, 1.0 / 5.0);
    p = fma(p, r, -1.0 / 4.0);
    p = fma(p, r, 1.0 / 3.0);
    p = fma(p, r, -0.5);
    p = fma(p * r, r, r);                       // log1p(r)
    const double ed = (double)e;
    return fma(ed, LN2_HI, tab[2 * i + 1] + fma(ed, LN2_LO, p));
}

} // namespace abn
