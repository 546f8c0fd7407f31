// fastmath.cuh - branch-free FP64 exp / reciprocal for the streaming row ops.
//
// The fit kernels are bound by instruction issue and the FP64 pipe, not by HBM (ncu: profiles/):
// libdevice's exp() / log() / IEEE division cost ~35 / ~45 / ~25 instructions each including
// range branches and slow paths that these loops never need. The versions here cost
//   exp    9 FP64 + 5 integer/LDS instructions   (max error 1.32 ulp, mean 0.34 ulp on |t| <= 700:
//                                                 tests/test_fastmath.py against long double)
//   rcp    1 MUFU + 5 FP64                        (correctly rounded 1/d for normal d: Markstein)
// __host__ __device__ with identical arithmetic (FMA on both sides) so tests/host_sim.cpp
// exercises the same bits the kernels produce, apart from the device's MUFU seed (the Newton
// steps converge to the same correctly rounded result).
#pragma once

#include <math.h>
#include <stdint.h>
#include <string.h>

#include "solver.cuh"

namespace ABN_NS {

// 2^(j/64), j = 0..63, correctly rounded (generated with 60-digit decimal arithmetic)
#define ABN_EXP2_N 64
#define ABN_EXP2_TABLE \
    {1.0, 1.0108892860517005, 1.0218971486541166, 1.0330248790212284, \
     1.0442737824274138, 1.0556451783605572, 1.0671404006768237, 1.0787607977571199, \
     1.0905077326652577, 1.102382583307841, 1.1143867425958924, 1.1265216186082418, \
     1.1387886347566916, 1.1511892299529827, 1.1637248587775775, 1.1763969916502812, \
     1.189207115002721, 1.202156731452703, 1.215247359980469, 1.22848053610687, \
     1.241857812073484, 1.255380757024691, 1.2690509571917332, 1.2828700160787783, \
     1.2968395546510096, 1.3109612115247644, 1.3252366431597413, 1.339667524053303, \
     1.3542555469368927, 1.3690024229745905, 1.383909881963832, 1.3989796725383112, \
     1.4142135623730951, 1.42961333839197, 1.4451808069770467, 1.460917794180647, \
     1.4768261459394993, 1.4929077282912648, 1.5091644275934228, 1.5255981507445384, \
     1.5422108254079407, 1.559004400237837, 1.5759808451078865, 1.593142151342267, \
     1.6104903319492543, 1.6280274218573478, 1.645755478153965, 1.6636765803267364, \
     1.681792830507429, 1.7001063537185235, 1.718619298122478, 1.7373338352737062, \
     1.7562521603732995, 1.7753764925265212, 1.7947090750031072, 1.8142521755003989, \
     1.8340080864093424, 1.8539791250833855, 1.8741676341103, 1.8945759815869656, \
     1.9152065613971474, 1.9360617934922943, 1.9571441241754002, 1.978456026387951}

#if defined(__CUDACC__)
__constant__ double c_exp2_table[ABN_EXP2_N] = ABN_EXP2_TABLE;
#endif
static const double h_exp2_table[ABN_EXP2_N] = ABN_EXP2_TABLE;

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

// high 32 bits of a double (sign, exponent, top 20 mantissa bits) and a double with them replaced.
// The row ops do their range checks and clamps on this word with integer min / max: one ALU
// instruction each, where fmax() / a double compare costs an FP64-pipe DSETP plus selects.
ABN_HD int hi_word(double d)
{
#if defined(__CUDA_ARCH__)
    return __double2hiint(d);
#else
    return (int)(uint32_t)(double_to_bits(d) >> 32);
#endif
}

ABN_HD double with_hi_word(double d, int hi)
{
#if defined(__CUDA_ARCH__)
    return __hiloint2double(hi, __double2loint(d));
#else
    return bits_to_double(((uint64_t)(uint32_t)hi << 32) | (double_to_bits(d) & 0xffffffffull));
#endif
}

// exp(t) for |t| <= 700 (callers clamp). `tab` points at the 64-entry table: shared memory on the
// device (lane-dependent index: a __constant__ read would serialise), h_exp2_table on the host.
//   t = (64 k + j) ln2/64 + r,  |r| <= ln2/128;  exp(t) = 2^k * 2^(j/64) * (1 + p(r))
ABN_HD double fast_exp(double t, const double *tab)
{
    // Constants whose low 32 bits are zero are encoded as immediates of the FP64 instructions;
    // the others cost two register moves per use site. So: INV only steers the choice of n (any
    // n within +-0.03 of the ideal keeps |r| < 0.0056), LN2_64_HI is ln2/64 cut to its high word
    // (n * HI exact for |n| < 2^32), and the r^4 / r^5 coefficients are rounded to 21 bits
    // (error < 1e-8 * r^4 < 1e-17).
    const double MAGIC = 6755399441055744.0;        // 1.5 * 2^52: round-to-nearest-integer trick
    const double INV = 92.33245849609375;           // 64 / ln 2 to 21 bits (0x40571547 00000000)
    const double LN2_64_HI = 0.010830417275428772;  // 0x3f862e42 00000000
    const double LN2_64_LO = 7.420820373486988e-09; // ln2/64 - HI
    const double C4 = 0.041666656732559204;         // 1/24  (0x3fa55555 00000000)
    const double C5 = 0.00833333283662796;          // 1/120 (0x3f811111 00000000)
    const double kd = fma(t, INV, MAGIC);
    const double nf = kd - MAGIC;                   // n = 64 k + j as a double
    double r = fma(nf, -LN2_64_HI, t);              // exact
    r = fma(nf, -LN2_64_LO, r);
    const int n = (int)(uint32_t)double_to_bits(kd); // low word of kd holds n
    // p(r) = r + r^2/2 + r^3/6 + r^4/24 + r^5/120, |r| <= 0.0056: truncation < 5e-17
    double p = fma(r, C5, C4);
    p = fma(p, r, 1.0 / 6.0);
    p = fma(p, r, 0.5);
    p = fma(p * r, r, r);
    const double tj = tab[n & 63];
    const double v = fma(tj, p, tj);                // in [1, 2) * (1 +- 0.0056)
    // * 2^k, k = n >> 6 in [-1010, 1010]: add k to the exponent field, i.e. (n >> 6) << 20 to the high word
    return with_hi_word(v, hi_word(v) + (int)(((uint32_t)n << 14) & 0xfff00000u));
}

// ---- exp for the lean passes: 256-entry table with pre-scaled exponent words ----------------------
// Same scheme as fast_exp with n = 256 k + j, |r| <= ln2/512, so the polynomial stops at r^4 (the r^5
// term is < 4e-17) - 9 FP64 instructions - and with the integer side cut to three instructions:
// the shared-memory copy of the table stores 2^(j/256) with (j << 12) taken off its high word, so
//     hi(2^k 2^(j/256)) = hi(entry_j) + (n << 12)        (one shift-add; n << 12 = (k << 20) + (j << 12))
// and the scaling is applied to the table entry BEFORE the last fma, which then yields exp(t) itself.
// Valid for |t| <= 700 (callers check the range after the pass; nothing clamps here).
#define ABN_EXP256_N 256
#if defined(__CUDACC__)
__constant__ double c_exp2_256_table[ABN_EXP256_N] = {
#include "exp2_256_table.inc"
};
#endif
static const double h_exp2_256_table[ABN_EXP256_N] = {
#include "exp2_256_table.inc"
};

// entry j of the pre-scaled table from the plain one
ABN_HD double exp256_entry(double two_j_256, int j)
{
    return with_hi_word(two_j_256, hi_word(two_j_256) - (j << 12));
}

ABN_HD double fast_exp256(double t, const double *tab)
{
    const double MAGIC = 6755399441055744.0;         // 1.5 * 2^52
    const double INV = 369.329833984375;             // 256 / ln 2 to 21 bits (0x40771547 00000000)
    const double LN2_256_HI = 0.002707604318857193;  // 0x3f662e42 00000000
    const double LN2_256_LO = 1.855205093371747e-09; // ln2/256 - HI
    const double C4 = 0.041666656732559204;          // 1/24 (0x3fa55555 00000000)
    const double kd = fma(t, INV, MAGIC);
    const double nf = kd - MAGIC;
    double r = fma(nf, -LN2_256_HI, t);              // exact
    r = fma(nf, -LN2_256_LO, r);
    const int n = (int)(uint32_t)double_to_bits(kd);
    // p(r) = r + r^2/2 + r^3/6 + r^4/24, |r| <= 0.00136
    double p = fma(r, C4, 1.0 / 6.0);
    p = fma(p, r, 0.5);
    p = fma(p * r, r, r);
    const double tj = tab[n & 255];
    const double ts = with_hi_word(tj, hi_word(tj) + (int)((uint32_t)n << 12)); // 2^k 2^(j/256)
    return fma(ts, p, ts);
}

// The same with ONE reduction step: r = t - n * fl(ln2/256) in a single fma (the product is exact inside the fma).
// The double nearest to ln2/256 is off by |delta| <= 2.2e-19, so r - and exp(t) relatively - is off by n * delta
// <= 6e-14 at |t| = 700 and <= 2e-15 for |t| <= 20: this computes exp((1 + kappa) t) with |kappa| <= 8e-17, a
// perturbation that is uniform in t and therefore invisible to a maximum-likelihood fit (it rescales every linear
// predictor by the same 1 + kappa). 8 FP64 instructions. Used by the lean passes only.
ABN_HD double fast_exp256s(double t, const double *tab)
{
    const double MAGIC = 6755399441055744.0;
    const double INV = 369.329833984375;
    const double LN2_256 = 0.0027076061740622863;    // ln2 / 256 rounded to nearest
    const double C4 = 0.041666656732559204;
    const double kd = fma(t, INV, MAGIC);
    const double nf = kd - MAGIC;
    const double r = fma(nf, -LN2_256, t);
    const int n = (int)(uint32_t)double_to_bits(kd);
    double p = fma(r, C4, 1.0 / 6.0);
    p = fma(p, r, 0.5);
    p = fma(p * r, r, r);
    const double tj = tab[n & 255];
    const double ts = with_hi_word(tj, hi_word(tj) + (int)((uint32_t)n << 12));
    return fma(ts, p, ts);
}

// ---- exp for the lean passes, bank-conflict-free: 64-entry table, 16 interleaved replicas -------------------
// A lane-dependent 8-byte table read costs as many shared-memory wavefronts as lanes collide on a bank pair:
// ~9 per warp for a 256-entry table (ncu, round 2) - 40 % of all shared-memory traffic of the K = 4 Poisson
// pass, which saturates the SM's 128 B/cycle shared-memory pipe before its FP64 pipe. Here entry j of replica
// c = lane mod 16 sits at tab[16 j + c]: the 16 lanes of a half warp always read 16 different bank pairs, two
// wavefronts per warp whatever the indices. 64 entries (|r| <= ln2/128) need the r^5 term back: 10 FP64
// instructions instead of 9. Entries are pre-scaled like exp256_entry: hi(2^k 2^(j/64)) = hi(entry_j) + (n << 14).
#define ABN_EXP64R_COPIES 16
ABN_HD double exp64r_entry(double two_j_64, int j)
{
    return with_hi_word(two_j_64, hi_word(two_j_64) - (j << 14));
}

// tab_lane = table base + (lane mod 16)
ABN_HD double fast_exp64r(double t, const double *tab_lane)
{
    const double MAGIC = 6755399441055744.0;        // 1.5 * 2^52
    const double INV = 92.33245849609375;           // 64 / ln 2 to 21 bits
    const double LN2_64_HI = 0.010830417275428772;  // 0x3f862e42 00000000
    const double LN2_64_LO = 7.420820373486988e-09; // ln2/64 - HI
    const double C4 = 0.041666656732559204;         // 1/24  (21 bits)
    const double C5 = 0.00833333283662796;          // 1/120 (21 bits)
    const double kd = fma(t, INV, MAGIC);
    const double nf = kd - MAGIC;
    double r = fma(nf, -LN2_64_HI, t);
    r = fma(nf, -LN2_64_LO, r);
    const int n = (int)(uint32_t)double_to_bits(kd);
    double p = fma(r, C5, C4);
    p = fma(p, r, 1.0 / 6.0);
    p = fma(p, r, 0.5);
    p = fma(p * r, r, r);
    const double tj = tab_lane[(n & 63) * ABN_EXP64R_COPIES];
    const double ts = with_hi_word(tj, hi_word(tj) + (int)((uint32_t)n << 14));
    return fma(ts, p, ts);
}

// 1 / d, correctly rounded for normal d with a normal result. MUFU.RCP64H seeds ~2^-20; one step
// with cubic convergence (x (1 + e + e^2)) lands within an ulp; then Markstein's correction
// r = fma(-d, x, 1), x + x r rounds correctly unless d's mantissa is all ones.
ABN_HD double fast_rcp(double d)
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

// 1 / d to about an ulp (not necessarily correctly rounded): the cubic step alone
ABN_HD double fast_rcp_approx(double d)
{
#if defined(__CUDA_ARCH__)
    double x;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(d));
    double e = fma(-d, x, 1.0);
    e = fma(e, e, e);
    return fma(x, e, x);
#else
    return 1.0 / d;
#endif
}

// 1 / d with one Newton step on the MUFU seed: relative error <= seed^2 (~1e-12 .. 1e-14). For the lean binomial
// passes: g = 1 / (1 + e) enters only the sums S = sum g x and w = g - g^2, where such an error moves the fixed point
// of the IRLS by its own size - four orders under the parity bar - and the log term does not use it.
ABN_HD double fast_rcp_newton1(double d)
{
#if defined(__CUDA_ARCH__)
    double x;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(d));
    const double e = fma(-d, x, 1.0);
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

} // namespace ABN_NS
