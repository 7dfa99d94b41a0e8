// sqrtW = (C + eta I)^(-1/2), the weight matrix of the reference's weighted least squares
// (codebase/gpkernels.py:496-504: eigh(C + eta I) -> V diag(lambda^-1/2) V^T, ValueError unless all lambda > 0).
//
// The GPU path does not diagonalise.  The symmetric inverse square root is the limit of the coupled
// Newton-Schulz iteration (Higham, Functions of Matrices, eq. 6.35), which is nothing but symmetric GEMMs and
// therefore runs on the same FP64 DMMA tile engine as the Cholesky:
//     A_s = (C + eta I) / s,  s = ||C + eta I||_inf  (so 0 < lambda(A_s) <= 1)
//     Y_0 = A_s, Z_0 = I;   T_k = Z_k Y_k;   Y_{k+1} = a_k Y_k - b_k Y_k T_k;   Z_{k+1} = a_k Z_k - b_k T_k Z_k
//     Y_k -> A_s^(1/2), Z_k -> A_s^(-1/2), T_k -> I (quadratically once ||I - T|| < 1);  sqrtW = Z / sqrt(s).
// Plain Newton-Schulz has a_k = 1.5, b_k = 0.5 and lifts a small eigenvalue of T by only 2.25 per step (31 steps
// for cond 1e11).  The steps are therefore SCALED (Chen & Chow 2014, stable scaling of Newton-Schulz): with x^2 an
// eigenvalue of T_k in [l_k^2, 1], the substitution (Y, Z) -> (mu Y, mu Z), mu_k = sqrt(3 / (1 + l_k + l_k^2)), gives
// a_k = 1.5 mu_k, b_k = 0.5 mu_k^3 and x -> mu x (3 - mu^2 x^2) / 2, which still maps (0, 1] into (0, 1] for every
// mu in [1, sqrt 3] (so a wrong bound l only costs speed) but lifts small eigenvalues of T by up to 6.75 per step:
// 13 steps for cond 1e11.  l_0 = sqrt(eta / s) (lambda_min(C + eta I) >= eta for a covariance C), l_{k+1} = the
// image of l_k; mu_k -> 1 as l_k -> 1, where the iteration is the plain quadratically convergent one.
// All iterates are polynomials in A_s, hence symmetric and commuting in exact arithmetic, so every product can be
// written as an "NT" product of row blocks.  Y and Z are kept exactly symmetric (tiles I >= J computed, mirrored);
// T = Z Y must NOT be symmetrised -- replacing it by its mirrored lower triangle perturbs the coupling and the
// iteration diverges once ||I - T|| ~ 1e-5 (observed at cond 5e4) -- so all T x T tiles of T are computed and
// written together with T^T (the operand layout Y T needs).  ~ log_2.25(cond) + 6 iterations (38 at cond 1e11).
// A matrix with an eigenvalue <= 0 makes Z blow up (non-finite residual) or never converge: status 1, which the
// Python layer maps to the reference's ValueError("inverse covariance not positive definite, increase eta").
// On the reference's three configurations the result satisfies || sqrtW (C + eta I) sqrtW - I ||_max 4x tighter
// than LAPACK's eigh route (the matrix is ill-posed element-wise).
#pragma once
#include "tile_engine.cuh"

namespace gpbo {

struct NsArgs {
    double* Y;      // [G][ld*ld]
    double* Z;
    double* T;
    double* TT;     // T^T
    double* Yn;
    double* Zn;
    long stride;    // ld * ld
    int ld;         // n padded to 128
    int n;
    double* part;   // [G][nT*nT] residual partials
};

// || C + eta I ||_inf per matrix (max absolute row sum, lower triangle mirrored); norm[] must be zeroed.
__global__ void __launch_bounds__(NTHR) ns_norm_kernel(const double* __restrict__ C, int n, double eta,
                                                       unsigned long long* __restrict__ norm) {
    const int p = blockIdx.y;
    const int row = blockIdx.x * (NTHR / 32) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= n) return;
    const double* Cp = C + (long)p * n * n;
    double s = 0.0;
    for (int c = lane; c < n; c += 32) {
        const double v = c <= row ? Cp[(long)row * n + c] : Cp[(long)c * n + row];
        s += fabs(c == row ? v + eta : v);
    }
    s = warp_sum(s);
    if (lane == 0) atomicMax(norm + p, (unsigned long long)__double_as_longlong(s));   // s >= 0: bit order = value order
}

// Y0 = (sym(C) + eta I) / s padded with the identity, Z0 = I.
__global__ void __launch_bounds__(NTHR) ns_init_kernel(const double* __restrict__ C, int n, double eta,
                                                       const unsigned long long* __restrict__ norm, NsArgs a) {
    const int p = blockIdx.y;
    const int r = blockIdx.x;
    const double s = __longlong_as_double((long long)norm[p]);
    const double inv_s = 1.0 / s;
    const double* Cp = C + (long)p * n * n;
    double* Yr = a.Y + (long)p * a.stride + (long)r * a.ld;
    double* Zr = a.Z + (long)p * a.stride + (long)r * a.ld;
    for (int c = threadIdx.x; c < a.ld; c += NTHR) {
        double y;
        if (r < n && c < n) {
            const double v = c <= r ? Cp[(long)r * n + c] : Cp[(long)c * n + r];
            y = (c == r ? v + eta : v) * inv_s;
        } else {
            y = r == c ? 1.0 : 0.0;
        }
        Yr[c] = y;
        Zr[c] = r == c ? 1.0 : 0.0;
    }
}

// which = first_which + blockIdx.y:
//   0: T  = Z Y (every tile; T^T written alongside),  part[p][tile] = sum (delta - T)^2 over the tile
//   1: Yn = ca Y - cb Y T  (tiles I >= J, mirrored)      2: Zn = ca Z - cb T Z  (tiles I >= J, mirrored)
// ntiles = nT * nT for which 0, nT (nT + 1) / 2 otherwise.
__global__ void __launch_bounds__(NTHR, 1) ns_gemm_kernel(NsArgs a, int ntiles, int first_which, double ca, double cb) {
    extern __shared__ __align__(16) double smem[];
    const ThreadCoord tc;
    const int p = blockIdx.x / ntiles, q = blockIdx.x % ntiles;
    const int which = first_which + blockIdx.y;
    int I, J;
    if (which == 0) {
        const int nT = a.ld / TB;
        I = q / nT;
        J = q % nT;
    } else {
        I = (int)((sqrt(8.0 * q + 1.0) - 1.0) * 0.5);
        while ((I + 1) * (I + 2) / 2 <= q) ++I;
        while (I * (I + 1) / 2 > q) --I;
        J = q - I * (I + 1) / 2;
    }
    const long off = (long)p * a.stride;
    // acc[r][c] = sum_k P[r][k] Q[c][k]:  Z Y = Z Y^T;  Y T = Y (T^T)^T;  T Z = T Z^T
    const double* P = (which == 0 ? a.Z : which == 1 ? a.Y : a.T) + off;
    const double* Q = (which == 0 ? a.Y : which == 1 ? a.TT : a.Z) + off;
    const double* X = (which == 1 ? a.Y : a.Z) + off;
    double* O = (which == 0 ? a.T : which == 1 ? a.Yn : a.Zn) + off;
    double* OT = a.TT + off;
    __shared__ uint64_t bars[2 * NSTAGE];
    Ring ring;
    ring_init(ring, bars);
    Acc acc;
    acc_zero(acc);
    const double* Prow = P + (long)I * TB * a.ld;
    const double* Qrow = Q + (long)J * TB * a.ld;
    gemm_nt_loop<false>(acc, [&](int kt) { return SliceSrc{Prow + kt * BK, a.ld, Qrow + kt * BK, a.ld}; }, a.ld / BK,
                        smem, ring, tc);
    double res[1] = {0.0};
#pragma unroll
    for (int mi = 0; mi < 8; ++mi)
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) {
            const int r = I * TB + tc.row(mi), c0 = J * TB + tc.col(ni, 0);
            double v[2] = {acc.v[mi][ni][0], acc.v[mi][ni][1]};
            if (which == 0) {
                *reinterpret_cast<double2*>(O + (long)r * a.ld + c0) = make_double2(v[0], v[1]);
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int c = c0 + e;
                    OT[(long)c * a.ld + r] = v[e];
                    const double d = (c == r ? 1.0 : 0.0) - v[e];
                    res[0] += d * d;
                }
            } else {
                const double2 x = *reinterpret_cast<const double2*>(X + (long)r * a.ld + c0);
                v[0] = ca * x.x - cb * v[0];
                v[1] = ca * x.y - cb * v[1];
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int c = c0 + e;
                    if (c <= r) {
                        O[(long)r * a.ld + c] = v[e];
                        if (c != r) O[(long)c * a.ld + r] = v[e];
                    }
                }
            }
        }
    if (which == 0) {
        __syncthreads();   // ring memory is re-used by block_sum
        double tot[1];
        block_sum<1>(res, smem, tot);
        if (tc.tid == 0) a.part[(long)p * ntiles + q] = tot[0];
    }
}

// resid[p] = sum of the tile partials (fixed order -> deterministic).
__global__ void __launch_bounds__(NTHR) ns_resid_kernel(const double* __restrict__ part, int ntiles, double* __restrict__ resid) {
    __shared__ double red[8];
    const int p = blockIdx.x;
    double v[1] = {0.0};
    for (int q = threadIdx.x; q < ntiles; q += NTHR) v[0] += part[(long)p * ntiles + q];
    double tot[1];
    block_sum<1>(v, red, tot);
    if (threadIdx.x == 0) resid[p] = tot[0];
}

// sqrtW[p] = Z[:n, :n] / sqrt(s)
__global__ void __launch_bounds__(NTHR) ns_out_kernel(NsArgs a, const unsigned long long* __restrict__ norm,
                                                      double* __restrict__ out) {
    const int p = blockIdx.y, r = blockIdx.x;
    const double f = rsqrt(__longlong_as_double((long long)norm[p]));
    const double* Zr = a.Z + (long)p * a.stride + (long)r * a.ld;
    double* o = out + (long)p * a.n * a.n + (long)r * a.n;
    for (int c = threadIdx.x; c < a.n; c += NTHR) o[c] = Zr[c] * f;
}

// ---- weighted least-squares products (codebase/wlstsq.py:183-188) -------------------------------
// For every GP g:  out_lhs[g] = sqrtW[g] @ lhs  (n x d, lhs shared by all GPs)  and  out_rhs[g] = sqrtW[g] @ rhs[g].
// HBM bound on the n x n weight matrices, which stay in device memory: one warp per output row streams the row of
// sqrtW once per chunk of WP_DC columns of lhs (the row stays in L1), lanes across k.
constexpr int WP_DC = 8;

__global__ void __launch_bounds__(NTHR)
weighted_products_kernel(const double* __restrict__ W, int n, const double* __restrict__ lhs, int d,
                         const double* __restrict__ rhs, double* __restrict__ out_lhs, double* __restrict__ out_rhs) {
    const int g = blockIdx.y;
    const int r = blockIdx.x * (NTHR / 32) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (r >= n) return;
    const double* Wr = W + ((long)g * n + r) * n;
    {
        const double* zg = rhs + (long)g * n;
        double s = 0.0;
        for (int k = lane; k < n; k += 32) s = fma(Wr[k], zg[k], s);
        s = warp_sum(s);
        if (lane == 0) out_rhs[(long)g * n + r] = s;
    }
    for (int c0 = 0; c0 < d; c0 += WP_DC) {
        double acc[WP_DC];
#pragma unroll
        for (int q = 0; q < WP_DC; ++q) acc[q] = 0.0;
        for (int k = lane; k < n; k += 32) {
            const double w = Wr[k];
            const double* Dk = lhs + (long)k * d + c0;
#pragma unroll
            for (int q = 0; q < WP_DC; ++q)
                if (c0 + q < d) acc[q] = fma(w, Dk[q], acc[q]);
        }
#pragma unroll
        for (int q = 0; q < WP_DC; ++q) {
            const double s = warp_sum(acc[q]);
            if (lane == 0 && c0 + q < d) out_lhs[((long)g * n + r) * d + c0 + q] = s;
        }
    }
}

}  // namespace gpbo
