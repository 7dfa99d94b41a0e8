// Posterior assembly of the operator-inference step: the linear algebra of
// PDEs/step3_estimate.py:75-95 on the outputs of the GP path, for a whole grid of regularisers at once.
//
// For every mode i (weights sqrtW_i from compute_lstsq_matrices, shared data matrix D, right-hand side z_i):
//     A_i = sqrtW_i D,  b_i = sqrtW_i z_i                      codebase/wlstsq.py:183-188   (weighted_products_kernel)
//     G_i = A_i^T A_i,  g_i = A_i^T b_i                                                      (gram_kernel)
// and for every regulariser lambda_k of the grid (step3_estimate.py:131-148 walks 81 of them, one solve each):
//     precision  P_ik = G_i + lambda_k^2 I                     step3_estimate.py:86-90
//     mean       mu_ik = argmin |A_i o - b_i|^2 + lambda_k^2 |o|^2 = P_ik^-1 g_i            step3_estimate.py:78-79
//     Cholesky   P_ik = C C^T -- what bayes.BayesianROM builds through scipy.stats.Covariance.from_precision
//                (codebase/bayes.py:283-287) to draw operator samples; a non-positive pivot is the reference's
//                "Matrix is not positive definite" -> that candidate is skipped (step3_estimate.py:92-95): status 1.
// The reference's solver (opinf.lstsq.L2Solver, un-vendored) takes the SVD of A_i; here the d x d normal equations are
// factored in shared memory and the mean gets ONE step of iterative refinement with the residual formed from A_i itself,
// r = A_i^T (b_i - A_i mu) - lambda^2 mu (corrected semi-normal equations), which restores the accuracy lost to
// cond(A)^2 as long as cond(P) eps < 1.  d <= POST_DMAX (operator rows of the reference's models have d = 21 ... 45).
// What step 3 does with each candidate afterwards -- integrating the ROM for every posterior draw -- needs `opinf` and is
// out of scope.
#pragma once
#include "common.cuh"

namespace gpbo {

constexpr int POST_DMAX = 128;      // largest operator-row length handled in shared memory
constexpr int GRAM_T = 16;          // output tile of gram_kernel
constexpr int GRAM_KC = 64;         // rows of A staged per step

// G[g] = A[g]^T A[g] (d x d), proj[g] = A[g]^T b[g].  A: [G][n][d] row-major, b: [G][n].
// grid (tiles_a * tiles_c, G), 256 threads = one 16 x 16 output tile; k runs over the n rows in a fixed order
// (deterministic).  Tiles with tc > ta are skipped (G is symmetric: the mirror is written by the lower tile).
__global__ void __launch_bounds__(256) gram_kernel(const double* __restrict__ A, const double* __restrict__ b, int n, int d,
                                                   double* __restrict__ G, double* __restrict__ proj) {
    __shared__ double sa[GRAM_KC][GRAM_T + 1];
    __shared__ double sc[GRAM_KC][GRAM_T + 1];
    __shared__ double sb[GRAM_KC];
    const int nt = (d + GRAM_T - 1) / GRAM_T;
    const int ta = blockIdx.x / nt, tc = blockIdx.x % nt;
    if (tc > ta) return;
    const int g = blockIdx.y;
    const double* Ag = A + (long)g * n * d;
    const double* bg = b + (long)g * n;
    const int la = threadIdx.x >> 4, lc = threadIdx.x & 15;     // output element (ta*16 + la, tc*16 + lc)
    double acc = 0.0, accb = 0.0;
    for (int k0 = 0; k0 < n; k0 += GRAM_KC) {
        for (int q = threadIdx.x; q < GRAM_KC * GRAM_T; q += 256) {
            const int kk = q >> 4, cc = q & 15;
            const int k = k0 + kk;
            const int ca = ta * GRAM_T + cc, ccol = tc * GRAM_T + cc;
            sa[kk][cc] = (k < n && ca < d) ? Ag[(long)k * d + ca] : 0.0;
            sc[kk][cc] = (k < n && ccol < d) ? Ag[(long)k * d + ccol] : 0.0;
        }
        if (threadIdx.x < GRAM_KC) sb[threadIdx.x] = (k0 + threadIdx.x < n) ? bg[k0 + threadIdx.x] : 0.0;
        __syncthreads();
#pragma unroll 8
        for (int kk = 0; kk < GRAM_KC; ++kk) {
            acc = fma(sa[kk][la], sc[kk][lc], acc);
            if (tc == 0 && lc == 0) accb = fma(sa[kk][la], sb[kk], accb);
        }
        __syncthreads();
    }
    const int r = ta * GRAM_T + la, c = tc * GRAM_T + lc;
    if (r < d && c < d) {
        double* Gg = G + (long)g * d * d;
        Gg[(long)r * d + c] = acc;
        if (ta != tc) Gg[(long)c * d + r] = acc;
    }
    if (tc == 0 && lc == 0 && r < d) proj[(long)g * d + r] = accb;
}

// In-shared Cholesky P = C C^T (lower, in place, stride ld) by the whole CTA; returns (uniformly) whether a pivot was
// <= 0 or not finite.  Right-looking, one column per step.
__device__ inline bool post_chol(double* P, int d, int ld, int* flag) {
    const int tid = threadIdx.x, nthr = blockDim.x;
    if (tid == 0) *flag = 0;
    __syncthreads();
    for (int j = 0; j < d; ++j) {
        if (tid == 0) {
            const double pjj = P[j * ld + j];
            if (!(pjj > 0.0) || !isfinite(pjj)) { *flag = 1; P[j * ld + j] = 1.0; }
            else P[j * ld + j] = sqrt(pjj);
        }
        __syncthreads();
        const double inv = 1.0 / P[j * ld + j];
        for (int i = j + 1 + tid; i < d; i += nthr) P[i * ld + j] *= inv;
        __syncthreads();
        // trailing lower triangle: element (i, c), j < c <= i
        const int rem = d - j - 1;
        for (int q = tid; q < rem * rem; q += nthr) {
            const int i = j + 1 + q / rem, c = j + 1 + q % rem;
            if (c <= i) P[i * ld + c] = fma(-P[i * ld + j], P[c * ld + j], P[i * ld + c]);
        }
        __syncthreads();
    }
    return *flag != 0;
}

// x <- (C C^T)^-1 x by forward and backward substitution (column oriented, all threads update the remaining entries).
__device__ inline void post_solve(const double* P, int d, int ld, double* x) {
    const int tid = threadIdx.x, nthr = blockDim.x;
    for (int j = 0; j < d; ++j) {
        if (tid == 0) x[j] = x[j] / P[j * ld + j];
        __syncthreads();
        const double xj = x[j];
        for (int i = j + 1 + tid; i < d; i += nthr) x[i] = fma(-P[i * ld + j], xj, x[i]);
        __syncthreads();
    }
    for (int j = d - 1; j >= 0; --j) {
        if (tid == 0) x[j] = x[j] / P[j * ld + j];
        __syncthreads();
        const double xj = x[j];
        for (int i = tid; i < j; i += nthr) x[i] = fma(-P[j * ld + i], xj, x[i]);
        __syncthreads();
    }
}

// One CTA per (mode g, regulariser k): precision, Cholesky, mean with one refinement step.
// grid (G, nreg), 256 threads, dynamic shared: (d * (d + 1) + 3 * d + 8) doubles.
// means: [nreg][G][d]; chol (may be NULL): [nreg][G][d][d] lower factor, strict upper part zero; status: [nreg][G];
// tmp: [nreg][G][n] scratch for b - A mu.
__global__ void __launch_bounds__(256) ridge_solve_kernel(const double* __restrict__ A, const double* __restrict__ b, int n,
                                                          int d, const double* __restrict__ Gm,
                                                          const double* __restrict__ proj, const double* __restrict__ regs,
                                                          double* __restrict__ means, double* __restrict__ chol,
                                                          int* __restrict__ status, double* __restrict__ tmp) {
    extern __shared__ __align__(16) double psm[];
    const int ld = d + 1;
    double* P = psm;
    double* mu = P + d * ld;
    double* rho = mu + d;
    double* gsh = rho + d;
    int* flag = reinterpret_cast<int*>(gsh + d);
    const int g = blockIdx.x, kreg = blockIdx.y, G = gridDim.x;
    const int tid = threadIdx.x, nthr = blockDim.x;
    const double lam2 = regs[kreg] * regs[kreg];
    const double* Gg = Gm + (long)g * d * d;
    for (int q = tid; q < d * d; q += nthr) {
        const int i = q / d, c = q % d;
        P[i * ld + c] = Gg[q] + (i == c ? lam2 : 0.0);
    }
    for (int i = tid; i < d; i += nthr) { gsh[i] = proj[(long)g * d + i]; mu[i] = gsh[i]; }
    __syncthreads();
    const bool bad = post_chol(P, d, ld, flag);
    const long slot = (long)kreg * G + g;
    if (tid == 0) status[slot] = bad ? 1 : 0;
    if (chol) {
        double* Cg = chol + slot * d * d;
        for (int q = tid; q < d * d; q += nthr) {
            const int i = q / d, c = q % d;
            Cg[q] = c <= i ? P[i * ld + c] : 0.0;
        }
    }
    double* mo = means + slot * d;
    if (bad) {      // the reference drops this candidate; the mean is not meaningful
        for (int i = tid; i < d; i += nthr) mo[i] = nan("");
        return;
    }
    post_solve(P, d, ld, mu);
    // refinement with the un-squared residual: t = b - A mu (n), rho = A^T t - lambda^2 mu (d)
    const double* Ag = A + (long)g * n * d;
    const double* bg = b + (long)g * n;
    double* tg = tmp + slot * n;
    for (int k = tid; k < n; k += nthr) {
        const double* Ak = Ag + (long)k * d;
        double s0 = 0.0, s1 = 0.0;
        int c = 0;
        for (; c + 1 < d; c += 2) { s0 = fma(Ak[c], mu[c], s0); s1 = fma(Ak[c + 1], mu[c + 1], s1); }
        if (c < d) s0 = fma(Ak[c], mu[c], s0);
        tg[k] = bg[k] - (s0 + s1);
    }
    __syncthreads();
    // thread group per column: 256 / 8 = 32 columns at a time, 8 threads stride the rows (fixed order, then a fixed
    // shuffle tree: deterministic)
    for (int c0 = 0; c0 < d; c0 += 32) {
        const int c = c0 + (tid >> 3), part = tid & 7;
        double s = 0.0;
        if (c < d)
            for (int k = part; k < n; k += 8) s = fma(Ag[(long)k * d + c], tg[k], s);
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        s += __shfl_xor_sync(0xffffffffu, s, 4);
        if (c < d && part == 0) rho[c] = s - lam2 * mu[c];
    }
    __syncthreads();
    post_solve(P, d, ld, rho);
    for (int i = tid; i < d; i += nthr) mo[i] = mu[i] + rho[i];
}

}  // namespace gpbo
