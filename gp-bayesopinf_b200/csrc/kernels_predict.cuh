// Posterior-moment kernels and the stand-alone kernel-matrix assembly (K1/K6).
//
// The prediction TRSM (V = L^-1 K_*^T) is chol_panel_kernel<.., CROSS=true> in kernels_chol.cuh;
// this file holds what follows it: Schur complement (derivative covariance), row norms
// (predictive variance), fused assemble+GEMV means, and the HBM-bound assembly kernel.
#pragma once
#include <type_traits>
#include "kernels_chol.cuh"

namespace gpbo {

// mean[p][a] = sum_j k(t'_a, t_j) alpha_j, kernel matrix generated on the fly (never stored).
// kind 0: sklearn K(X*, X) alpha (_gpr.py:446-447); 1: kappa_zy alpha (gpkernels.py:485);
// 2: K_zy alpha (gpkernels.py:488).  One warp per output row.
__global__ void __launch_bounds__(NTHR)
mean_kernel(MatArgs a, CrossArgs cr, const double* __restrict__ alpha, double* __restrict__ out, long out_stride) {
    const int p = blockIdx.y;
    const int row = blockIdx.x * (NTHR / 32) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= cr.nrow) return;
    const PairParams q = a.pp[p];
    const double xr = cr.trow[(long)p * cr.lrow + row];
    const double* tsp = a.ts + (long)p * a.lda;
    const double* al = alpha + (long)p * a.lda;
    double s = 0.0;
    for (int j = lane; j < a.m; j += 32) s += cross_element(cr.kind, cr.fam, q, row, j, cr.nrow, a.m, xr, tsp[j]) * al[j];
    s = warp_sum(s);
    if (lane == 0) out[(long)p * out_stride + row] = s;
}

// std[p][a] = sqrt(max(0, (sigma^2 + chi) - sum_k V[k][a]^2))  (_gpr.py:480-500), V^T rows in X.
__global__ void __launch_bounds__(NTHR)
std_kernel(MatArgs a, const double* __restrict__ X, long x_stride, int nrow, double* __restrict__ out, long out_stride) {
    const int p = blockIdx.y;
    const int row = blockIdx.x * (NTHR / 32) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= nrow) return;
    const PairParams q = a.pp[p];
    const double* xr = X + (long)p * x_stride + (long)row * a.lda;
    double s = 0.0;
    for (int k = 2 * lane; k < a.lda; k += 64) {
        const double2 v = *reinterpret_cast<const double2*>(xr + k);
        s += v.x * v.x + v.y * v.y;
    }
    s = warp_sum(s);
    if (lane == 0) {
        double var = (q.sig2 + q.chi) - s;
        if (var < 0.0) var = 0.0;
        out[(long)p * out_stride + row] = sqrt(var);
    }
}

// K_zz as a function of the (scaled like cr.trow) distance d between two estimation points (gpkernels.py:641).
__device__ __forceinline__ double kzz_element(int fam, const PairParams& pr, double d) {
    const double d2 = d * d;
    if (fam == 0) {
        const double ell2 = pr.ell * pr.ell;
        const double kap = pr.sig2 * gpbo_exp_neg(-gpbo_div(d2, 2 * ell2, 0.5 * pr.inv_ell2));
        return gpbo_div((1 - gpbo_div(d2, ell2, pr.inv_ell2)) * kap, ell2, pr.inv_ell2);
    }
    // Matern: -kappa''(tau), rows are t' / ell
    const double K = fabs(d) * (fam == 3 ? 1.7320508075688772 : 2.23606797749979);
    const double ex = gpbo_exp_neg(-K);
    const double a2 = (fam == 3 ? 3.0 : 5.0) * pr.inv_ell2;
    return fam == 3 ? pr.sig2 * a2 * (1.0 - K) * ex : pr.sig2 * (a2 / 3.0) * (1.0 + K - K * K) * ex;
}

// Row N2: the reference's estimation grid is np.linspace (PDEs/main.py:101-105), so K_zz is Toeplitz --
// m' distinct values per GP instead of m'^2 exponentials.  One CTA per GP: decides whether the (unscaled) points are
// equispaced to rounding (|t_k - t_0 - k h| <= 8 eps max|t|) and tabulates K_zz by lag from the scaled rows the Schur
// kernel uses, tab[p][k] = K_zz(trow[k] - trow[0]).  flag[p] = 1: schur_kernel indexes the table by r - c.
__global__ void __launch_bounds__(NTHR)
kzz_table_kernel(const double* __restrict__ src, long src_stride, CrossArgs cr, const PairParams* __restrict__ pp,
                 double* __restrict__ tab, int* __restrict__ flag) {
    __shared__ double red[NTHR / 32];
    const int p = blockIdx.x, n = cr.nrow;
    const double* sp = src + (long)p * src_stride;
    const double* tr = cr.trow + (long)p * cr.lrow;
    const PairParams pr = pp[p];
    const double s0 = sp[0], s1 = sp[n - 1];
    const double h = n > 1 ? (s1 - s0) / (double)(n - 1) : 0.0;
    double dev = 0.0;
    for (int k = threadIdx.x; k < n; k += NTHR) {
        dev = fmax(dev, fabs((sp[k] - s0) - (double)k * h));
        tab[(long)p * cr.lrow + k] = kzz_element(cr.fam, pr, tr[k] - tr[0]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) dev = fmax(dev, __shfl_xor_sync(0xffffffffu, dev, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = dev;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < NTHR / 32; ++w) dev = fmax(dev, red[w]);
        const double scale = fmax(fabs(s0), fabs(s1));
        flag[p] = (n > 2 && h != 0.0 && dev <= 8.0 * 2.220446049250313e-16 * scale) ? 1 : 0;
    }
}

// C = K_zz - V^T V (gpkernels.py:491-493, 641): tile (I, J), I >= J, of the Schur complement of the
// joint covariance [[K_yy, K_zy^T], [K_zy, K_zz]]; written symmetric into the unpadded m' x m' output.
// kzz_tab / kzz_flag (may be null): Toeplitz table of kzz_table_kernel.
__global__ void __launch_bounds__(NTHR, 1)
schur_kernel(MatArgs a, const double* __restrict__ X, long x_stride, CrossArgs cr, int ntiles,
             double* __restrict__ C, long c_stride, const double* __restrict__ kzz_tab, const int* __restrict__ kzz_flag) {
    extern __shared__ __align__(16) double smem[];
    const ThreadCoord tc;
    const int p = blockIdx.x / ntiles, q = blockIdx.x % ntiles;
    int I = (int)((sqrt(8.0 * q + 1.0) - 1.0) * 0.5);
    while ((I + 1) * (I + 2) / 2 <= q) ++I;
    while (I * (I + 1) / 2 > q) --I;
    const int J = q - I * (I + 1) / 2;
    const double* XI = X + (long)p * x_stride + (long)I * TB * a.lda;
    const double* XJ = X + (long)p * x_stride + (long)J * TB * a.lda;
    __shared__ uint64_t bars[2 * NSTAGE];
    Ring ring;
    ring_init(ring, bars);
    Acc acc;
    acc_zero(acc);
    if (I == J) {
        gemm_nt_loop<true>(acc, [&](int kt) { return SliceSrc{XI + kt * BK, a.lda, nullptr, 0}; }, a.lda / BK, smem,
                           ring, tc);
    } else {
        gemm_nt_loop<false>(acc, [&](int kt) { return SliceSrc{XI + kt * BK, a.lda, XJ + kt * BK, a.lda}; },
                            a.lda / BK, smem, ring, tc);
    }
    const PairParams pr = a.pp[p];
    const double* tr = cr.trow + (long)p * cr.lrow;
    double* Cp = C + (long)p * c_stride;
    const int n = cr.nrow;
    const bool toeplitz = kzz_flag != nullptr && kzz_flag[p] != 0;
    const double* tab = kzz_tab + (long)p * cr.lrow;
    double xr[8], xc[4][2];
#pragma unroll
    for (int mi = 0; mi < 8; ++mi) xr[mi] = tr[I * TB + tc.row(mi)];
#pragma unroll
    for (int ni = 0; ni < 4; ++ni)
#pragma unroll
        for (int e = 0; e < 2; ++e) xc[ni][e] = tr[J * TB + tc.col(ni, e)];
#pragma unroll
    for (int mi = 0; mi < 8; ++mi)
#pragma unroll
        for (int ni = 0; ni < 4; ++ni)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int r = I * TB + tc.row(mi), c = J * TB + tc.col(ni, e);
                if (r < n && c < n && c <= r) {
                    const double kzz = toeplitz ? tab[r - c] : kzz_element(cr.fam, pr, xr[mi] - xc[ni][e]);
                    const double v = kzz - acc.v[mi][ni][e];
                    Cp[(long)r * n + c] = v;
                    if (r != c) Cp[(long)c * n + r] = v;
                }
            }
}

// rows[p][i] = src[i] / ell_p (order 0) or src[i] (order 1), zero padded to lrow.
__global__ void scale_rows_kernel(const double* __restrict__ src, long src_stride, int n, int lrow, int order,
                                  const PairParams* __restrict__ pp, double* __restrict__ dst) {
    const int p = blockIdx.x;
    const double ell = pp[p].ell;
    const double* s = src + (long)p * src_stride;
    for (int i = threadIdx.x; i < lrow; i += blockDim.x)
        dst[(long)p * lrow + i] = i < n ? (order == 0 ? s[i] / ell : s[i]) : 0.0;
}

// ---- stand-alone assembly to HBM (K1 / K6) ---------------------------------------------------
// kind: 0 K(theta) sklearn order (train, chi on the diagonal)      kernels.py:1559-1565
//       1 K_yy rbf_eval order (chi on the diagonal)                gpkernels.py:630,639
//       2 K(t1, t2) sklearn cross (no white noise)                 kernels.py:1566-1570,1418
//       3 kappa(t1, t2) rbf_eval                                   gpkernels.py:608-609
//       4 K_zy = -(t1-t2) kappa / ell^2                            gpkernels.py:640
//       5 K_zz = (1 - (t1-t2)^2/ell^2) kappa / ell^2               gpkernels.py:641
//       6 dK/dlog(ell) = sigma^2 R (x1-x2)^2                       kernels.py:1575-1577, 966-969
// The element generators keep each reference expression's operation order (x = t/ell is a true
// division done once per abscissa; divisions by ell^2 / 2 ell^2 use gpbo_div); exp is the table-driven
// gpbo_exp_neg_tab (10 FP64 instructions, <= 1.5 ulp; fastmath.h).
// An element costs ~20 FP64-pipe instructions, which at 64 FP64 op/clk/SM is within ~10 % of what the
// HBM write stream needs, so
//  * the general kernel (t1 != t2) writes each row segment with 16-byte stores, 512 contiguous bytes
//    per warp, one 32 x 512 tile per CTA;
//  * when t1 == t2 (train matrices, K_zz, dK/dlog ell) the symmetric kernel evaluates only tiles
//    I >= J (64 x 64), stores the tile from registers and its transpose through shared memory: half the
//    FP64 work, so the kernel is HBM-write bound.
struct AsmConsts {
    double sig2, ell, chi, ell2, two_ell2, r_ell2, r_two_ell2;
    double c_e, c_zy, c_zz0, c_zz1;     // folded constants of the cross kinds 3, 4, 5
    __device__ __forceinline__ AsmConsts() {}
    // Evaluated ONCE per matrix by asm_consts_kernel and read back by every thread of the assembly kernels: three
    // exp() and two divisions are ~100 FP64-pipe instructions, as much as 7 matrix elements -- done per thread that
    // was 30 % of the symmetric kernel's FP64 work (16 elements per thread) and 11 % of the general kernel's (64);
    // done by one thread per CTA it left the other 255 waiting at the first barrier (ncu: top stall).
    __device__ __forceinline__ void init(const double* theta) {
        sig2 = exp(theta[0]); ell = exp(theta[1]); chi = exp(theta[2]);
        ell2 = ell * ell; two_ell2 = 2 * ell2; r_ell2 = 1.0 / ell2; r_two_ell2 = 0.5 * r_ell2;
        c_e = -r_two_ell2; c_zy = -sig2 * r_ell2; c_zz0 = sig2 * r_ell2; c_zz1 = -(sig2 * r_ell2) * r_ell2;
    }
};

// FAM 0: RBF (the reference's kernel).  FAM 3 / 5: Matern nu = 3/2 / 5/2 -- an extension (the reference has no
// Matern kernel; BASELINE.json's north star names "RBF/Matern kernel-matrix assembly"): K follows scikit-learn's
// Matern (kernels.py:1601-1790: dists = |x1 - x2| with x = t / ell, K = dists * sqrt(2 nu), ...), the derivative
// cross-covariances are the analytic d/dt1 and d^2/dt1 dt2 of the same kernel.  Matern kinds 1 / 3 alias 0 / 2.
template <int FAM, int KIND>
__device__ __forceinline__ double assemble_element(const AsmConsts& k, bool diag, double x1, double x2,
                                                   const double* __restrict__ tab) {
    if (FAM == 0) {
        const double d = x1 - x2;
        const double d2 = d * d;
        if (KIND == 0) return diag ? k.sig2 + k.chi : k.sig2 * gpbo_exp_neg_tab(-0.5 * d2, tab);
        if (KIND == 1) return diag ? k.sig2 + k.chi : k.sig2 * gpbo_exp_neg_tab(-gpbo_div(d2, k.two_ell2, k.r_two_ell2), tab);
        if (KIND == 2) return k.sig2 * gpbo_exp_neg_tab(-0.5 * d2, tab);
        if (KIND == 6) return k.sig2 * (gpbo_exp_neg_tab(-0.5 * d2, tab) * d2);
        // Cross matrices (never inverted): the divisions by ell^2 are folded into per-matrix constants, 19-20 FP64
        // instructions per element instead of 22-30.  An element then differs from the reference's operation order
        // by a few ulp plus |d^2 / (2 ell^2)| ulp (rounding of the exponent argument), i.e. <= 2e-16 of the
        // matrix scale; the fused posterior-moment kernels keep the reference's order.
        const double e = gpbo_exp_neg_tab(d2 * k.c_e, tab);
        if (KIND == 3) return k.sig2 * e;
        if (KIND == 4) return (d * k.c_zy) * e;
        return fma(d2, k.c_zz1, k.c_zz0) * e;
    } else {
        // x1, x2 are t / ell for every Matern kind
        const double dx = x1 - x2;
        const double dist = fabs(dx);
        const double c = (FAM == 3) ? 1.7320508075688772 : 2.23606797749979;      // sqrt(2 nu)
        const double K = dist * c;
        const double e = gpbo_exp_neg_tab(-K, tab);
        if (KIND == 0 || KIND == 1) {
            if (diag) return k.sig2 + k.chi;
            return k.sig2 * ((FAM == 3) ? (1.0 + K) * e : (1.0 + K + K * K / 3.0) * e);
        }
        if (KIND == 2 || KIND == 3) return k.sig2 * ((FAM == 3) ? (1.0 + K) * e : (1.0 + K + K * K / 3.0) * e);
        if (KIND == 6) {   // dK / dlog(ell): kernels.py (nu = 1.5: 3 D exp(-sqrt(3 D)); nu = 2.5: 5/3 D (tmp + 1) exp(-tmp))
            const double D = dx * dx;
            return k.sig2 * ((FAM == 3) ? 3.0 * D * e : 5.0 / 3.0 * D * (K + 1.0) * e);
        }
        // derivatives with respect to the unscaled times: tau = ell * dx, a = c / ell, a^2 = 2 nu / ell^2
        const double a2 = (FAM == 3 ? 3.0 : 5.0) * k.r_ell2;
        const double tau = dx * k.ell;
        if (KIND == 4)    // d k / d t1
            return (FAM == 3) ? -k.sig2 * a2 * tau * e : -k.sig2 * (a2 / 3.0) * tau * (1.0 + K) * e;
        // KIND 5: d^2 k / d t1 d t2 = - k''(tau)
        return (FAM == 3) ? k.sig2 * a2 * (1.0 - K) * e : k.sig2 * (a2 / 3.0) * (1.0 + K - K * K) * e;
    }
}

// Pre-pass of every assembly call (one tiny launch): the per-matrix constants, and -- for the kinds whose reference
// expression works on x = t / ell (sklearn order: a true division, once per abscissa) -- the scaled abscissae, so that
// the assembly kernels themselves contain no division and no exp() outside the element generator.
// grid (ceil(max(n1, n2) / 256), B); xs1 / xs2: [B][n1] / [B][n2] (xs2 unused when t2 aliases t1).
__global__ void asm_prep_kernel(const double* __restrict__ theta, const double* __restrict__ t1, long t1_stride, int n1,
                                const double* __restrict__ t2, long t2_stride, int n2, bool scaled, bool same,
                                AsmConsts* __restrict__ consts, double* __restrict__ xs1, double* __restrict__ xs2) {
    const int p = blockIdx.y;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        AsmConsts k;
        k.init(theta + 3 * p);
        consts[p] = k;
    }
    if (!scaled) return;
    const double ell = exp(theta[3 * p + 1]);
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n1) xs1[(long)p * n1 + i] = t1[(long)p * t1_stride + i] / ell;
    if (!same && i < n2) xs2[(long)p * n2 + i] = t2[(long)p * t2_stride + i] / ell;
}

// K_ii = sigma^2 + chi of the train kinds 0 / 1 after the general kernel (which treats every element as off-diagonal,
// so its loop carries no row == column test; the symmetric kernel patches its diagonal tiles itself).
__global__ void asm_diag_kernel(const AsmConsts* __restrict__ consts, int n, double* __restrict__ out, long out_stride) {
    const int p = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[(long)p * out_stride + (long)i * n + i] = consts[p].sig2 + consts[p].chi;
}

template <int FAM, int KIND>
struct AsmScaled { static constexpr bool value = (FAM != 0) || (KIND == 0 || KIND == 2 || KIND == 6); };

constexpr int ASM_ROWS = 32;     // rows per CTA of the general kernel
// resident CTAs per SM the two assembly kernels are compiled for (register caps 32 / 40); -D overrides for A/B builds
#ifndef GPBO_GEN_MINB
#define GPBO_GEN_MINB 8
#endif
#ifndef GPBO_SYM2_MINB
#define GPBO_SYM2_MINB 6
#endif

// General kernel (t1 != t2).  blockDim.x = 64 ... 256 threads (a multiple of 32, chosen by the host so that a matrix with
// a short second dimension -- the reference's 3200 x 200 -- does not leave most of a 256-thread CTA idle); each thread
// owns two adjacent columns of the CTA's 32 rows and writes them with one 16-byte store per row (512 contiguous bytes
// per warp).  x1 / x2 are the abscissae as the element generator wants them (pre-scaled by asm_prep_kernel or raw).
// The kernel is instruction-issue bound (ncu: issue 76-78 %, FP64 pipe 50-53 %), so it is kept at 32 registers
// (8 CTAs / SM) and free of anything that is not per element.
template <int FAM, int KIND>
__global__ void __launch_bounds__(NTHR, GPBO_GEN_MINB)
assemble_general_kernel(const double* __restrict__ x1, long x1_stride, int n1, const double* __restrict__ x2,
                        long x2_stride, int n2, const AsmConsts* __restrict__ consts, double* __restrict__ out,
                        long out_stride) {
    __shared__ double x1s[ASM_ROWS];
    __shared__ double tab[64];
    for (int i = threadIdx.x; i < 64; i += blockDim.x) tab[i] = GPBO_EXP2_TAB[i];
    const int p = blockIdx.z;
    const int c0 = blockIdx.x * (2 * blockDim.x) + 2 * threadIdx.x;
    const int r0 = blockIdx.y * ASM_ROWS;
    if (threadIdx.x < ASM_ROWS) {
        const int r = r0 + threadIdx.x;
        x1s[threadIdx.x] = r < n1 ? x1[(long)p * x1_stride + r] : 0.0;
    }
    __syncthreads();
    if (c0 >= n2) return;
    const AsmConsts k = consts[p];
    const bool two = (c0 + 1 < n2);
    const double* a2 = x2 + (long)p * x2_stride;
    const double x2a = a2[c0], x2b = two ? a2[c0 + 1] : 0.0;
    double* o = out + (long)p * out_stride;
    const bool vec = two && ((n2 & 1) == 0) && ((reinterpret_cast<uintptr_t>(o) & 15) == 0);
    const int nr = min(ASM_ROWS, n1 - r0);
    double* orow = o + (long)r0 * n2 + c0;
#pragma unroll 4
    for (int rr = 0; rr < nr; ++rr, orow += n2) {
        const double xa = x1s[rr];
        const double va = assemble_element<FAM, KIND>(k, false, xa, x2a, tab);
        const double vb = assemble_element<FAM, KIND>(k, false, xa, x2b, tab);
        if (vec) {
            *reinterpret_cast<double2*>(orow) = make_double2(va, vb);
        } else {
            orow[0] = va;
            if (two) orow[1] = vb;
        }
    }
}

// (Two special kernels for matrices with a short second dimension -- the reference's m' >> m shape 3200 x 200 -- were
// measured and dropped in round 2b: a "flat" one treating the output as one contiguous array and a "narrow" one arranging
// the threads as row groups x column pairs, profiles/r02c_asm_ab.txt.  What that shape needed was less work outside the
// element loop -- the pre-pass above -- and a CTA no wider than the matrix.  A CTA sweeping several row blocks (set-up
// once per CTA instead of once per 32 rows) spills at the 32-register cap and runs at 0.45-0.60: dropped, same file.)

// Symmetric kernel (t1 is t2: train matrices, K_zz, dK/dlog ell): only tiles I >= J (64 x 64) are evaluated -- half the
// FP64 work.  The DIRECT tile is stored straight from the registers that computed it (thread (ty, tx) owns rows
// ty + 16 i, columns tx + 16 j: the 16 lanes of a half warp write one 128-byte line of a row) and only the
// TRANSPOSED tile is staged in shared memory (odd stride: conflict-free both ways); interior tiles take a path without
// bounds predicates.  The first-generation kernel (both tiles through shared memory, predicated scalar copy loops:
// ~36 thread instructions per element written, issue bound at 0.62-0.86 of the copy peak) ran at 0.68-0.73 where this
// one runs at 0.85-1.02 (profiles/r02d_asm_ab.txt).
constexpr int SYM_T = 64;            // tile edge of the symmetric kernel
constexpr int SYM_LD = SYM_T + 1;    // odd stride: conflict-free row and column access
constexpr int SYM2_SMEM = (SYM_T * SYM_LD + 2 * SYM_T) * 8;     // 34304 B -> 6 CTAs / SM

template <int FAM, int KIND>
__global__ void __launch_bounds__(NTHR, GPBO_SYM2_MINB)
assemble_sym2_kernel(const double* __restrict__ x, long x_stride, int n, const AsmConsts* __restrict__ consts,
                     double* __restrict__ out, long out_stride) {
    extern __shared__ __align__(16) double dsm[];
    double* ST = dsm;
    double* xr = ST + SYM_T * SYM_LD;
    double* xc = xr + SYM_T;
    __shared__ double tab[64];
    const int tid = threadIdx.x;
    if (tid >= 128 && tid < 192) tab[tid - 128] = GPBO_EXP2_TAB[tid - 128];
    const int p = blockIdx.y, q = blockIdx.x;
    int I = (int)((sqrt(8.0 * q + 1.0) - 1.0) * 0.5);
    while ((I + 1) * (I + 2) / 2 <= q) ++I;
    while (I * (I + 1) / 2 > q) --I;
    const int J = q - I * (I + 1) / 2;
    if (tid < 2 * SYM_T) {
        const int idx = (tid < SYM_T ? I * SYM_T + tid : J * SYM_T + tid - SYM_T);
        (tid < SYM_T ? xr[tid] : xc[tid - SYM_T]) = idx < n ? x[(long)p * x_stride + idx] : 0.0;
    }
    __syncthreads();
    const AsmConsts k = consts[p];
    const int ty = tid >> 4, tx = tid & 15;
    double x1[4], x2[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { x1[i] = xr[ty + 16 * i]; x2[i] = xc[tx + 16 * i]; }
    double* o = out + (long)p * out_stride;
    const bool interior = (I + 1) * SYM_T <= n;       // J <= I: the tile and its mirror lie inside the matrix
    double* od = o + (long)(I * SYM_T + ty) * n + J * SYM_T + tx;
    double* om = o + (long)(J * SYM_T + ty) * n + I * SYM_T + tx;     // mirror tile: rows J*64 + r, columns I*64 + c
    const int rmax = n - I * SYM_T, cmax = n - J * SYM_T;            // only read on the boundary path
    auto tile = [&](auto full_tag) {
        constexpr bool FULL = decltype(full_tag)::value;
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int r = ty + 16 * i, c = tx + 16 * j;
                const double v = assemble_element<FAM, KIND>(k, I == J && r == c, x1[i], x2[j], tab);
                if (FULL || (r < rmax && c < cmax)) od[(long)(16 * i) * n + 16 * j] = v;
                if (I != J) ST[c * SYM_LD + r] = (KIND == 4) ? -v : v;
            }
        if (I == J) return;
        __syncthreads();
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int r = ty + 16 * i, c = tx + 16 * j;
                if (FULL || (r < cmax && c < rmax)) om[(long)(16 * i) * n + 16 * j] = ST[r * SYM_LD + c];
            }
    };
    if (interior) tile(std::true_type{});
    else tile(std::false_type{});
}

}  // namespace gpbo
