// FP64 exp and constant-divisor division for the kernel-matrix element generators.
//
// Kernel-matrix assembly is FP64-pipe bound, not HBM bound, with the stock exp()/division (measured:
// 180-330 Gelem/s = 22-41 % of HBM write bandwidth).  These two helpers bring an RBF element
// down to ~20 FP64-pipe instructions.  Both compile for host and device so tests/test_fastmath.py can
// measure their error on the CPU (g++ -mfma) against long-double references.
//
//   gpbo_exp(x)          |error| <= 1 ulp for -708 <= x <= 708.  Results below 2^-1021 (x < -708) are flushed to 0
//                        (absolute error < 3.4e-308; the reference's np.exp returns sub-normals there);
//                        x > 708 returns +inf, NaN returns NaN.
//   gpbo_div(a, b, rb)   a / b for a divisor b whose correctly rounded reciprocal rb = 1/b is known;
//                        one Newton correction with FMA residual: correctly rounded except for rare
//                        1-ulp cases (Markstein).
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define GPBO_HD __host__ __device__ __forceinline__
#else
#define GPBO_HD inline
#endif

namespace gpbo {

GPBO_HD int64_t f64_bits(double v) {
#if defined(__CUDA_ARCH__)
    return __double_as_longlong(v);
#else
    int64_t b;
    memcpy(&b, &v, 8);
    return b;
#endif
}
GPBO_HD double bits_f64(int64_t b) {
#if defined(__CUDA_ARCH__)
    return __longlong_as_double(b);
#else
    double v;
    memcpy(&v, &b, 8);
    return v;
#endif
}

// Degree-11 minimax polynomial of exp on [-ln2/2, ln2/2] (own Remez fit, relative error 3.1e-18),
// Cody-Waite reduction with FMA, scaling by exponent-field addition.
GPBO_HD double gpbo_exp(double x) {
    const double SHIFT = 6755399441055744.0;               // 1.5 * 2^52: rint() by addition
    const double L2E = 1.4426950408889634;                 // log2(e)
    const double LN2_HI = 6.9314718055994529e-01;          // ln 2 rounded to double
    const double LN2_LO = 2.3190468138462996e-17;          // ln 2 - LN2_HI
    const double t = fma(x, L2E, SHIFT);
    const double k = t - SHIFT;
    double r = fma(k, -LN2_HI, x);
    r = fma(k, -LN2_LO, r);
    double p = 2.4994246136424405e-08;
    p = fma(p, r, 2.763236802746315e-07);
    p = fma(p, r, 2.7557623140145747e-06);
    p = fma(p, r, 2.4801486320566664e-05);
    p = fma(p, r, 0.0001984126943145065);
    p = fma(p, r, 0.001388888895141027);
    p = fma(p, r, 0.008333333333560176);
    p = fma(p, r, 0.041666666666492075);
    p = fma(p, r, 0.16666666666666166);
    p = fma(p, r, 0.5000000000000018);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    // low 32 bits of t hold k as a two's-complement integer
    const int64_t ki = (int64_t)(int32_t)(uint32_t)(f64_bits(t) & 0xffffffffLL);
    double res = bits_f64(f64_bits(p) + (ki << 52));
    // |x| > 708 (or NaN): decided on the integer pipe so the test costs no FP64 issue slot
    const int64_t xb = f64_bits(x);
    const uint32_t ahi = (uint32_t)(xb >> 32) & 0x7fffffffu;
    if (ahi > 0x40862000u) {
        const bool isnan_ = ahi > 0x7ff00000u || (ahi == 0x7ff00000u && (uint32_t)xb != 0u);
        res = isnan_ ? x : (xb < 0 ? 0.0 : bits_f64(0x7ff0000000000000LL));
    }
    return res;
}

// Device fast path for arguments x <= 0 (every kernel-matrix element: -d^2/2, -d^2/(2 ell^2), -sqrt(2 nu) d).
// Same arithmetic as gpbo_exp, with (i) the constants read as constant-bank operands of the DFMAs (as 64-bit
// immediates they cost two uniform-register moves each: 14 % of the issue slots of the assembly kernel, which is
// issue bound), (ii) the flush of results below 2^-1021 done with integer selects instead of a branch.
// NaN is NOT propagated (the result is 0): the abscissae are validated on the host (the *_host entry points reject
// non-finite t), hyper-parameters enter through sigma^2 which multiplies every element.
#define GPBO_EXPC_VALUES                                                                                         \
    6755399441055744.0, 1.4426950408889634, -6.9314718055994529e-01, -2.3190468138462996e-17,                   \
    2.4994246136424405e-08, 2.763236802746315e-07, 2.7557623140145747e-06, 2.4801486320566664e-05,              \
    0.0001984126943145065, 0.001388888895141027, 0.008333333333560176, 0.041666666666492075,                    \
    0.16666666666666166, 0.5000000000000018, 1.0, 1.0
#if defined(__CUDACC__)
static __device__ __constant__ double GPBO_EXPC[16] = {GPBO_EXPC_VALUES};
#endif

GPBO_HD double gpbo_exp_neg(double x) {
#if defined(__CUDA_ARCH__)
    const double t = fma(x, GPBO_EXPC[1], GPBO_EXPC[0]);
    const double k = t - GPBO_EXPC[0];
    double r = fma(k, GPBO_EXPC[2], x);
    r = fma(k, GPBO_EXPC[3], r);
    double p = GPBO_EXPC[4];
#pragma unroll
    for (int i = 5; i < 16; ++i) p = fma(p, r, GPBO_EXPC[i]);
    const int ki = __double2loint(t);
    const int hi = __double2hiint(p) + (ki << 20);
    const int lo = __double2loint(p);
    const bool flush = (unsigned)(__double2hiint(x) & 0x7fffffff) > 0x40862000u;     // |x| > 708 (or NaN / inf)
    return __hiloint2double(flush ? 0 : hi, flush ? 0 : lo);
#else
    // host restatement of the same operations, for tests/test_fastmath.py
    static const double C[16] = {GPBO_EXPC_VALUES};
    const double t = fma(x, C[1], C[0]);
    const double k = t - C[0];
    double r = fma(k, C[2], x);
    r = fma(k, C[3], r);
    double p = C[4];
    for (int i = 5; i < 16; ++i) p = fma(p, r, C[i]);
    const int64_t tb = f64_bits(t), pb = f64_bits(p), xb = f64_bits(x);
    const int32_t ki = (int32_t)(uint32_t)(tb & 0xffffffffLL);
    const int32_t hi = (int32_t)(uint32_t)((uint64_t)pb >> 32) + (int32_t)((uint32_t)ki << 20);
    const uint32_t lo = (uint32_t)(pb & 0xffffffffLL);
    const bool flush = ((uint32_t)((uint64_t)xb >> 32) & 0x7fffffffu) > 0x40862000u;
    return flush ? 0.0 : bits_f64((int64_t)(((uint64_t)(uint32_t)hi << 32) | lo));
#endif
}

// Table-driven variant for the stand-alone assembly kernels (kernels_predict.cuh), which are FP64-issue bound:
// exp(x) = 2^(k/64) * exp(r), k = rint(64 x / ln 2), |r| <= ln2/128 = 5.4e-3, so a degree-5 Taylor polynomial is exact
// to 3.5e-17 relative and the whole exponential costs 10 FP64-pipe instructions instead of 16; the 64 correctly rounded
// values 2^(j/64) come from a table the kernel copies into shared memory (an LDS, not an FP64 slot).
// |error| <= 1.5 ulp for -708 <= x <= 0 (measured on the host: tests/test_fastmath.py), results below
// 2^-1021 flushed to 0 like gpbo_exp_neg, exp(0) = 1 exactly.  `tab` is GPBO_EXP2_TAB or a shared-memory copy of it.
#define GPBO_EXP2_TAB_VALUES                                                                      \
    0x1.0000000000000p+0, 0x1.02c9a3e778061p+0, 0x1.059b0d3158574p+0, 0x1.0874518759bc8p+0,        \
    0x1.0b5586cf9890fp+0, 0x1.0e3ec32d3d1a2p+0, 0x1.11301d0125b51p+0, 0x1.1429aaea92de0p+0,        \
    0x1.172b83c7d517bp+0, 0x1.1a35beb6fcb75p+0, 0x1.1d4873168b9aap+0, 0x1.2063b88628cd6p+0,        \
    0x1.2387a6e756238p+0, 0x1.26b4565e27cddp+0, 0x1.29e9df51fdee1p+0, 0x1.2d285a6e4030bp+0,        \
    0x1.306fe0a31b715p+0, 0x1.33c08b26416ffp+0, 0x1.371a7373aa9cbp+0, 0x1.3a7db34e59ff7p+0,        \
    0x1.3dea64c123422p+0, 0x1.4160a21f72e2ap+0, 0x1.44e086061892dp+0, 0x1.486a2b5c13cd0p+0,        \
    0x1.4bfdad5362a27p+0, 0x1.4f9b2769d2ca7p+0, 0x1.5342b569d4f82p+0, 0x1.56f4736b527dap+0,        \
    0x1.5ab07dd485429p+0, 0x1.5e76f15ad2148p+0, 0x1.6247eb03a5585p+0, 0x1.6623882552225p+0,        \
    0x1.6a09e667f3bcdp+0, 0x1.6dfb23c651a2fp+0, 0x1.71f75e8ec5f74p+0, 0x1.75feb564267c9p+0,        \
    0x1.7a11473eb0187p+0, 0x1.7e2f336cf4e62p+0, 0x1.82589994cce13p+0, 0x1.868d99b4492edp+0,        \
    0x1.8ace5422aa0dbp+0, 0x1.8f1ae99157736p+0, 0x1.93737b0cdc5e5p+0, 0x1.97d829fde4e50p+0,        \
    0x1.9c49182a3f090p+0, 0x1.a0c667b5de565p+0, 0x1.a5503b23e255dp+0, 0x1.a9e6b5579fdbfp+0,        \
    0x1.ae89f995ad3adp+0, 0x1.b33a2b84f15fbp+0, 0x1.b7f76f2fb5e47p+0, 0x1.bcc1e904bc1d2p+0,        \
    0x1.c199bdd85529cp+0, 0x1.c67f12e57d14bp+0, 0x1.cb720dcef9069p+0, 0x1.d072d4a07897cp+0,        \
    0x1.d5818dcfba487p+0, 0x1.da9e603db3285p+0, 0x1.dfc97337b9b5fp+0, 0x1.e502ee78b3ff6p+0,        \
    0x1.ea4afa2a490dap+0, 0x1.efa1bee615a27p+0, 0x1.f50765b6e4540p+0, 0x1.fa7c1819e90d8p+0
// 1.5 * 2^52, 64 / ln 2, -(ln 2 / 64) hi / lo, Taylor coefficients 1/120, 1/24, 1/6, 1/2
#define GPBO_EXPT_VALUES                                                                                         \
    6755399441055744.0, 92.33248261689366, -0.010830424696249145, -3.623510646634843e-19,                       \
    8.3333333333333332e-03, 4.1666666666666664e-02, 1.6666666666666666e-01, 0.5
#if defined(__CUDACC__)
static __device__ __constant__ double GPBO_EXP2_TAB[64] = {GPBO_EXP2_TAB_VALUES};
static __device__ __constant__ double GPBO_EXPT[8] = {GPBO_EXPT_VALUES};
#endif

GPBO_HD double gpbo_exp_neg_tab(double x, const double* tab) {
#if defined(__CUDA_ARCH__)
    const double t = fma(x, GPBO_EXPT[1], GPBO_EXPT[0]);
    const double k = t - GPBO_EXPT[0];
    double r = fma(k, GPBO_EXPT[2], x);
    r = fma(k, GPBO_EXPT[3], r);
    const int ki = __double2loint(t);
    const double tj = tab[ki & 63];
    const double r2 = r * r;
    double p = fma(r, GPBO_EXPT[4], GPBO_EXPT[5]);
    p = fma(p, r, GPBO_EXPT[6]);
    p = fma(p, r, GPBO_EXPT[7]);
    const double v = fma(tj, fma(p, r2, r), tj);          // 2^(j/64) (1 + (r + r^2 (1/2 + r/6 + r^2/24 + r^3/120)))
    const int hi = __double2hiint(v) + ((ki >> 6) << 20);
    const int lo = __double2loint(v);
    const bool flush = (unsigned)(__double2hiint(x) & 0x7fffffff) > 0x40862000u;     // |x| > 708 (or NaN / inf)
    return __hiloint2double(flush ? 0 : hi, flush ? 0 : lo);
#else
    static const double C[8] = {GPBO_EXPT_VALUES};
    const double t = fma(x, C[1], C[0]);
    const double k = t - C[0];
    double r = fma(k, C[2], x);
    r = fma(k, C[3], r);
    const int64_t tb = f64_bits(t), xb = f64_bits(x);
    const int32_t ki = (int32_t)(uint32_t)(tb & 0xffffffffLL);
    const double r2 = r * r;
    double p = fma(r, C[4], C[5]);
    p = fma(p, r, C[6]);
    p = fma(p, r, C[7]);
    const double tj = tab[ki & 63];
    const double v = fma(tj, fma(p, r2, r), tj);
    const int64_t vb = f64_bits(v);
    const int32_t hi = (int32_t)(uint32_t)((uint64_t)vb >> 32) + (int32_t)((uint32_t)(ki >> 6) << 20);
    const uint32_t lo = (uint32_t)(vb & 0xffffffffLL);
    const bool flush = ((uint32_t)((uint64_t)xb >> 32) & 0x7fffffffu) > 0x40862000u;
    return flush ? 0.0 : bits_f64((int64_t)(((uint64_t)(uint32_t)hi << 32) | lo));
#endif
}

GPBO_HD double gpbo_div(double a, double b, double rb) {
    const double q = a * rb;
    const double e = fma(-q, b, a);
    return fma(e, rb, q);
}

}  // namespace gpbo
