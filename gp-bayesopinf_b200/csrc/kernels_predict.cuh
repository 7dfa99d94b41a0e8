// Posterior-moment kernels and the stand-alone kernel-matrix assembly (K1/K6).
//
// The prediction TRSM (V = L^-1 K_*^T) is chol_panel_kernel<.., CROSS=true> in kernels_chol.cuh;
// this file holds what follows it: Schur complement (derivative covariance), row norms
// (predictive variance), fused assemble+GEMV means, and the HBM-bound assembly kernel.
#pragma once
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
    for (int j = lane; j < a.m; j += 32) s += cross_element(cr.kind, q, row, j, cr.nrow, a.m, xr, tsp[j]) * al[j];
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

// C = K_zz - V^T V (gpkernels.py:491-493, 641): tile (I, J), I >= J, of the Schur complement of the
// joint covariance [[K_yy, K_zy^T], [K_zy, K_zz]]; written symmetric into the unpadded m' x m' output.
__global__ void __launch_bounds__(NTHR, 1)
schur_kernel(MatArgs a, const double* __restrict__ X, long x_stride, CrossArgs cr, int ntiles,
             double* __restrict__ C, long c_stride) {
    extern __shared__ __align__(16) double smem[];
    const ThreadCoord tc;
    const int p = blockIdx.x / ntiles, q = blockIdx.x % ntiles;
    int I = (int)((sqrt(8.0 * q + 1.0) - 1.0) * 0.5);
    while ((I + 1) * (I + 2) / 2 <= q) ++I;
    while (I * (I + 1) / 2 > q) --I;
    const int J = q - I * (I + 1) / 2;
    const double* XI = X + (long)p * x_stride + (long)I * TB * a.lda;
    const double* XJ = X + (long)p * x_stride + (long)J * TB * a.lda;
    Acc acc;
    acc_zero(acc);
    if (I == J) {
        auto f = [&](int kt) { return XI + kt * BK; };
        gemm_nt_loop<true>(acc, f, a.lda, f, a.lda, a.lda / BK, smem, tc);
    } else {
        gemm_nt_loop<false>(acc, [&](int kt) { return XI + kt * BK; }, a.lda, [&](int kt) { return XJ + kt * BK; },
                            a.lda, a.lda / BK, smem, tc);
    }
    const PairParams pr = a.pp[p];
    const double ell2 = pr.ell * pr.ell;
    const double* tr = cr.trow + (long)p * cr.lrow;
    double* Cp = C + (long)p * c_stride;
    const int n = cr.nrow;
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
                    const double d = xr[mi] - xc[ni][e];
                    const double d2 = d * d;
                    const double kap = pr.sig2 * exp(-d2 / (2 * ell2));
                    const double kzz = (1 - (d2 / ell2)) * kap / ell2;
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
// Each thread produces two adjacent columns (one 16-byte store); a warp covers 512 contiguous
// bytes of a row; a CTA covers ASM_ROWS rows x 512 columns.
constexpr int ASM_ROWS = 8;

__device__ __forceinline__ double assemble_element(int kind, double sig2, double ell, double chi, double ell2, int r,
                                                   int c, double x1, double x2) {
    const double d = x1 - x2;
    const double d2 = d * d;
    switch (kind) {
        case 0: return r == c ? sig2 + chi : sig2 * exp(-0.5 * d2);
        case 1: return r == c ? sig2 + chi : sig2 * exp(-d2 / (2 * ell2));
        case 2: return sig2 * exp(-0.5 * d2);
        case 3: return sig2 * exp(-d2 / (2 * ell2));
        case 4: return -d * (sig2 * exp(-d2 / (2 * ell2))) / ell2;
        case 5: return (1 - (d2 / ell2)) * (sig2 * exp(-d2 / (2 * ell2))) / ell2;
        default: return sig2 * (exp(-0.5 * d2) * d2);
    }
}

__global__ void __launch_bounds__(NTHR)
assemble_kernel(int kind, const double* __restrict__ t1, long t1_stride, int n1, const double* __restrict__ t2,
                long t2_stride, int n2, const double* __restrict__ theta, int B, double* __restrict__ out,
                long out_stride) {
    const int p = blockIdx.z;
    const double sig2 = exp(theta[3 * p]), ell = exp(theta[3 * p + 1]), chi = exp(theta[3 * p + 2]);
    const double ell2 = ell * ell;
    const bool scaled = (kind == 0 || kind == 2 || kind == 6);
    const int c0 = blockIdx.x * (2 * NTHR) + 2 * threadIdx.x;
    const int r0 = blockIdx.y * ASM_ROWS;
    if (c0 >= n2) return;
    const double* a1 = t1 + (long)p * t1_stride;
    const double* a2 = t2 + (long)p * t2_stride;
    const bool two = (c0 + 1 < n2);
    double x2a = a2[c0], x2b = two ? a2[c0 + 1] : 0.0;
    if (scaled) { x2a = x2a / ell; x2b = x2b / ell; }
    double* o = out + (long)p * out_stride;
    const bool vec = two && ((n2 & 1) == 0) && ((reinterpret_cast<uintptr_t>(o) & 15) == 0);
#pragma unroll
    for (int rr = 0; rr < ASM_ROWS; ++rr) {
        const int r = r0 + rr;
        if (r >= n1) break;
        double x1 = a1[r];
        if (scaled) x1 = x1 / ell;
        const double va = assemble_element(kind, sig2, ell, chi, ell2, r, c0, x1, x2a);
        if (vec) {
            const double vb = assemble_element(kind, sig2, ell, chi, ell2, r, c0 + 1, x1, x2b);
            *reinterpret_cast<double2*>(o + (long)r * n2 + c0) = make_double2(va, vb);
        } else {
            o[(long)r * n2 + c0] = va;
            if (two) o[(long)r * n2 + c0 + 1] = assemble_element(kind, sig2, ell, chi, ell2, r, c0 + 1, x1, x2b);
        }
    }
}

}  // namespace gpbo
