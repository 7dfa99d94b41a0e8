// Small-matrix path (m <= SMALL_MAX = 224): the reference's own experiment sizes (m = 10 ... 200,
// ODEs/experiments.sh:11-18, PDEs/experiments.sh:13-26, PDEsMulti/experiments.sh:6).
//
// At these sizes the blocked 128-tile path is latency bound (about ten launches and a host round trip per
// lock-step round).  Here ONE CTA evaluates a whole LML + gradient (sklearn _gpr.py:583-651) with the matrix
// resident in shared memory, in one launch, and the multi-start fit (sklearn _gpr.py:298-340, 658-668) is ONE
// persistent kernel: every CTA pulls (GP, start) pairs from a queue and runs the complete L-BFGS-B optimisation of
// a pair -- the same state machine as the host path (lbfgsb.h, __host__ __device__) on thread 0, the evaluation on
// all threads -- so there is no host round trip and no lock-step: stragglers only occupy their own CTA.
//
// Storage: the lower triangle, packed by rows (element (i, j), j <= i, at i (i + 1) / 2 + j), n = m rounded up to a
// multiple of 32 with identity padding; 224 x 224 packed = 197 KB of the 227 KB of shared memory per CTA.  Packed
// rows give conflict-free column walks: consecutive columns of one row are adjacent, and the starts of 16
// consecutive rows fall into 16 different 8-byte bank pairs (triangular numbers mod 16 are a permutation).
//   1. K(theta) into L                                  kernels.py:1559-1565, 1279-1292, 1407-1414
//   2. Cholesky in place, 32-wide panels                _gpr.py:589-593  (not PD -> status 1 -> (-inf, 0))
//   3. W = L^-1 in place (diagonal blocks by substitution in registers, block rows as two small GEMMs)
//   4. z = W y, alpha = W^T z                           _gpr.py:601
//   5. K^-1 = W^T W tile by tile in registers, fused with the three gradient traces (dK regenerated on the fly)
//                                                       _gpr.py:629-651, kernels.py:1575-1577, 966-969
//   6. LML = -1/2 y'alpha - sum log L_ii - m/2 log 2 pi _gpr.py:613-617
// All contractions are register-tiled FP64 FMA loops (6 x 6 or 2 x 12 outputs per thread): on B200 the FP64 FMA
// rate equals the DMMA rate (64 FMA / clk / SM), so the tensor pipe has nothing to add at one CTA per matrix.
#pragma once
#include "kernels_chol.cuh"
#include "lbfgsb.h"

namespace gpbo {

constexpr int SMALL_MAX = 224;     // largest training size of the in-shared path
constexpr int SMALL_DEFAULT = 96;  // sizes up to this take it by default (measured crossover with the blocked path)
constexpr int SB = 32;             // panel / block width

__host__ __device__ __forceinline__ int tri(int i) { return (i * (i + 1)) >> 1; }
inline int small_pad(int m) { return (m + SB - 1) / SB * SB; }
// packed matrix + x, y, alpha, z, dinv vectors + reduction scratch
inline size_t small_smem_bytes(int n) { return ((size_t)tri(n) + 5 * (size_t)n + 64) * sizeof(double); }

struct SmallProblem {
    const double* t;   // [G][m]
    const double* y;   // [G][m]
    int m;             // valid size
    int n;             // m rounded up to a multiple of 32
    long long* dbg;    // nullptr, or 16 counters: clock cycles of thread 0 per phase, summed over all CTAs (GPBO_SMALL_DBG)
};

// phase timing for tuning runs (GPBO_SMALL_DBG=1): thread 0 accumulates the cycles since the previous mark
struct SmallClock {
    long long* dbg;
    long long last;
};
__device__ __forceinline__ long long small_now() {
#ifdef __CUDA_ARCH__
    return clock64();
#else
    return 0;
#endif
}
__device__ __forceinline__ void small_mark(SmallClock& c, int phase) {
    if (c.dbg && threadIdx.x == 0) {
        const long long now = small_now();
        atomicAdd(reinterpret_cast<unsigned long long*>(c.dbg + phase), (unsigned long long)(now - c.last));
        c.last = now;
    }
}

struct SmallBox { double lo[3], hi[3]; };

template <int FAM>
__device__ __forceinline__ void small_kernel_pair(double sig2, double dx, double& kr, double& dkl) {
    // sigma^2 R_ij and dK_ij / dlog(ell) for scaled abscissae difference dx
    const double d2 = dx * dx;
    if (FAM == 0) {
        kr = sig2 * gpbo_exp_neg(-0.5 * d2);
        dkl = kr * d2;
    } else {
        const double K = fabs(dx) * (FAM == 3 ? 1.7320508075688772 : 2.23606797749979);
        const double ex = gpbo_exp_neg(-K);
        kr = sig2 * ((FAM == 3) ? (1.0 + K) * ex : (1.0 + K + K * K / 3.0) * ex);
        dkl = sig2 * ((FAM == 3) ? 3.0 * d2 * ex : 5.0 / 3.0 * d2 * (K + 1.0) * ex);
    }
}

// Deterministic block sum of NV values per thread (fixed shuffle tree, fixed warp order); red: >= NV * NT / 32 doubles.
template <int NT, int NV>
__device__ __forceinline__ void small_block_sum(double (&v)[NV], double* red, double (&out)[NV]) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const double s = warp_sum(v[i]);
        if (lane == 0) red[i * (NT / 32) + warp] = s;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < NT / 32; ++w) s += red[i * (NT / 32) + w];
        out[i] = s;
    }
    __syncthreads();
}

#ifndef GPBO_CHOL32_ALLTHREADS
// Cholesky of the 32 x 32 diagonal block at (o, o) of the packed matrix by warp 0 (warp_chol32, kernels_chol.cuh:
// registers + shuffles, no barrier inside).  dinv[o + i] = 1 / L_ii.  flag: one double of scratch.  Ends with
// __syncthreads(); returns (uniformly) whether a pivot was <= 0.  flag: 33 doubles of scratch.
__device__ __forceinline__ bool small_chol32(double* L, int o, double* dinv, double* flag, int nv = 32) {
    if (threadIdx.x < 32) {
        const int lane = threadIdx.x;
        const bool bad = warp_chol32(L + tri(o + lane) + o, dinv + o + lane, nv);
        if (lane == 0) *flag = bad ? 1.0 : 0.0;
    }
    __syncthreads();
    return *flag != 0.0;
}
#else
// Variant kept for A/B measurements (-DGPBO_CHOL32_ALLTHREADS): all threads of the CTA, one barrier per pivot step.
__device__ __forceinline__ bool small_chol32(double* L, int o, double* dinv, double* rsv, int /*nv*/ = 32) {
    const int NT = blockDim.x;
    const int TPR = NT / SB;
    const int tid = threadIdx.x;
    const int r = tid / TPR, cg = tid % TPR;
    double* Lr = L + tri(o + r) + o;
    bool bad = false;
    for (int j = 0; j < SB - 1; ++j) {
        double d = L[tri(o + j) + o + j];
        if (!(d > 0.0)) { bad = true; d = 1.0; }
        const double inv_d = __drcp_rn(d);
        if (tid == 0) rsv[j] = d;
        if (r > j) {
            const double w = Lr[j] * inv_d;
            for (int c = cg; c < SB; c += TPR)
                if (c > j && c <= r) Lr[c] = fma(-w, L[tri(o + c) + o + j], Lr[c]);
        }
        __syncthreads();
    }
    {
        double d = L[tri(o + SB - 1) + o + SB - 1];
        if (!(d > 0.0)) { bad = true; d = 1.0; }
        if (tid == 0) rsv[SB - 1] = d;
    }
    __syncthreads();
    if (tid < SB) rsv[tid] = rsqrt(rsv[tid]);
    __syncthreads();
    for (int c = cg; c < SB; c += TPR)
        if (c <= r) {
            const double v = Lr[c];
            Lr[c] = (c == r && !(v > 0.0)) ? 1.0 : v * rsv[c];
        }
    if (tid < SB) dinv[o + tid] = rsv[tid];
    __syncthreads();
    return bad;
}
#endif

// One LML + gradient evaluation by the whole CTA.  sm: dynamic shared memory (small_smem_bytes(n)).
// Thread 0 returns the results in out[0..3] = (lml, grad) and *st = 0 / 1 (not positive definite: lml = -inf, grad = 0),
// which are valid for every thread after the trailing __syncthreads().
template <int NT, int FAM>
__device__ void small_eval(const SmallProblem& pr, int gp, double th0, double th1, double th2, double* sm,
                           double* out, int* st) {
    constexpr int NW = NT / 32;
    constexpr int TY = NT / 16;                   // thread grid TY x 16 of the register-tiled contractions
    constexpr int CR = 6 * TY;                    // rows per chunk of the 6 x 6 tiles
    constexpr int RI = SB / TY;                   // rows per thread of a 32-row block (2 x 12 tiles)
    constexpr int NMAXT = (NT == 64) ? 64 : SMALL_MAX;
    constexpr int JJ = (NMAXT - SB) / 16;         // column groups of the widest block row
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ty = tid >> 4, tx = tid & 15;
    const int m = pr.m, n = pr.n, nblk = n / SB;
    double* L = sm;
    double* x = L + tri(n);
    double* yv = x + n;
    double* al = yv + n;
    double* z = al + n;
    double* dinv = z + n;
    double* red = dinv + n;                       // 64 doubles

    SmallClock clk{pr.dbg, pr.dbg ? small_now() : 0};
    const double sig2 = exp(th0), ell = exp(th1), chi = exp(th2);
    for (int i = tid; i < n; i += NT) {
        x[i] = i < m ? pr.t[(long)gp * m + i] / ell : 0.0;        // kernels.py:1559 (X / length_scale)
        yv[i] = i < m ? pr.y[(long)gp * m + i] : 0.0;
    }
    __syncthreads();

    // 1. K(theta): rows i and n-1-i together (n + 1 elements) so that every warp gets the same share
    for (int rp = warp; rp < n / 2; rp += NW) {
        const int ia = rp, ib = n - 1 - rp;
        for (int e = lane; e <= n; e += 32) {
            const int i = e <= ia ? ia : ib;
            const int j = e <= ia ? e : e - ia - 1;
            double v;
            if (i >= m || j >= m) v = i == j ? 1.0 : 0.0;
            else if (i == j) v = sig2 + chi;
            else if (FAM == 0) v = sig2 * gpbo_exp_neg(-0.5 * ((x[i] - x[j]) * (x[i] - x[j])));
            else v = matern_value<(FAM == 0 ? 3 : FAM)>(sig2, x[i] - x[j]);
            L[tri(i) + j] = v;
        }
    }
    __syncthreads();
    small_mark(clk, 0);

    // 2. Cholesky, right-looking with 32-wide panels
    bool bad = false;
    for (int kb = 0; kb < nblk; ++kb) {
        const int o = kb * SB;
        bad |= small_chol32(L, o, dinv, red, min(SB, m - o));
        small_mark(clk, 1);
        const int R0 = o + SB, nb = n - R0;
        if (nb <= 0) break;
        // panel: X L_kk^T = A by forward substitution, one thread per row (row in registers, L_kk broadcast)
        if (tid < nb) {
            double* Ar = L + tri(R0 + tid) + o;
            double xr[SB];
#pragma unroll
            for (int c = 0; c < SB; ++c) xr[c] = Ar[c];
#pragma unroll
            for (int c = 0; c < SB; ++c) {
                const double* Lk = L + tri(o + c) + o;
                double s0 = xr[c], s1 = 0.0;
#pragma unroll
                for (int k = 0; k + 1 < c; k += 2) {
                    s0 = fma(-xr[k], Lk[k], s0);
                    s1 = fma(-xr[k + 1], Lk[k + 1], s1);
                }
                if (c & 1) s0 = fma(-xr[c - 1], Lk[c - 1], s0);
                xr[c] = (s0 + s1) * dinv[o + c];
            }
#pragma unroll
            for (int c = 0; c < SB; ++c) Ar[c] = xr[c];
        }
        __syncthreads();
        small_mark(clk, 2);
        // trailing update of the lower triangle: A[r][c] -= sum_k X[r][k] X[c][k], 6 x 6 register tiles
        for (int rc0 = 0; rc0 < nb; rc0 += CR)
            for (int cc0 = 0; cc0 < rc0 + CR && cc0 < nb; cc0 += 96) {
                double acc[6][6];
                int pr_[6], pc_[6];
#pragma unroll
                for (int ii = 0; ii < 6; ++ii) {
                    const int r = rc0 + ty + TY * ii, c = cc0 + tx + 16 * ii;
                    pr_[ii] = r < nb ? tri(R0 + r) + o : -1;
                    pc_[ii] = c < nb ? tri(R0 + c) + o : -1;
#pragma unroll
                    for (int jj = 0; jj < 6; ++jj) acc[ii][jj] = 0.0;
                }
#pragma unroll 4
                for (int k = 0; k < SB; ++k) {
                    double xr[6], xc[6];
#pragma unroll
                    for (int ii = 0; ii < 6; ++ii) {
                        xr[ii] = pr_[ii] >= 0 ? L[pr_[ii] + k] : 0.0;
                        xc[ii] = pc_[ii] >= 0 ? L[pc_[ii] + k] : 0.0;
                    }
#pragma unroll
                    for (int ii = 0; ii < 6; ++ii)
#pragma unroll
                        for (int jj = 0; jj < 6; ++jj) acc[ii][jj] = fma(xr[ii], xc[jj], acc[ii][jj]);
                }
#pragma unroll
                for (int ii = 0; ii < 6; ++ii)
#pragma unroll
                    for (int jj = 0; jj < 6; ++jj) {
                        const int r = rc0 + ty + TY * ii, c = cc0 + tx + 16 * jj;
                        if (r < nb && c <= r) L[tri(R0 + r) + R0 + c] -= acc[ii][jj];
                    }
            }
        __syncthreads();
        small_mark(clk, 3);
    }
    if (bad) {                                    // uniform: every thread saw the same pivots
        if (tid == 0) { out[0] = -INFINITY; out[1] = 0.0; out[2] = 0.0; out[3] = 0.0; *st = 1; }
        __syncthreads();
        return;
    }

    // 3a. inverses of the 32 x 32 diagonal blocks, one warp per block, lane = column, the column in registers
    for (int b = warp; b < nblk; b += NW) {
        const int o = b * SB, jc = lane;
        double w[SB];
#pragma unroll
        for (int rr = 0; rr < SB; ++rr) {
            const double* Lr = L + tri(o + rr) + o;
            double s0 = 0.0, s1 = 0.0;
#pragma unroll
            for (int k = 0; k + 1 < rr; k += 2) {
                s0 = fma(Lr[k], w[k], s0);            // w[k] = 0 for k < jc
                s1 = fma(Lr[k + 1], w[k + 1], s1);
            }
            if (rr & 1) s0 = fma(Lr[rr - 1], w[rr - 1], s0);
            const double di = dinv[o + rr];
            w[rr] = rr < jc ? 0.0 : (rr == jc ? di : -(s0 + s1) * di);
        }
        __syncwarp();                             // every lane has finished reading L_bb
#pragma unroll
        for (int rr = 0; rr < SB; ++rr)
            if (rr >= jc) L[tri(o + rr) + o + jc] = w[rr];
    }
    __syncthreads();
    small_mark(clk, 4);
    // 3b. block rows: W_i,: = -W_ii (L_i,: W_<i,<i), both products with 2 x 12 (RI x JJ) register tiles
    for (int bi = 1; bi < nblk; ++bi) {
        const int o = bi * SB;
        double acc[RI][JJ];
        int prow[RI];
#pragma unroll
        for (int ii = 0; ii < RI; ++ii) {
            prow[ii] = tri(o + ty + TY * ii);
#pragma unroll
            for (int jj = 0; jj < JJ; ++jj) acc[ii][jj] = 0.0;
        }
        // G[r][c] = sum_{k = c}^{o - 1} L[o + r][k] W[k][c]
        for (int kb16 = 0; kb16 < o / 16; ++kb16)
#pragma unroll 4
            for (int kk = 0; kk < 16; ++kk) {
                const int k = kb16 * 16 + kk;
                const double* Wk = L + tri(k);
                double lr[RI];
#pragma unroll
                for (int ii = 0; ii < RI; ++ii) lr[ii] = L[prow[ii] + k];
#pragma unroll
                for (int jj = 0; jj < JJ; ++jj)
                    if (jj <= kb16) {
                        const int c = tx + 16 * jj;
                        const double wv = c <= k ? Wk[c] : 0.0;
#pragma unroll
                        for (int ii = 0; ii < RI; ++ii) acc[ii][jj] = fma(lr[ii], wv, acc[ii][jj]);
                    }
            }
        __syncthreads();                          // all of L_i,[0,o) has been read
#pragma unroll
        for (int ii = 0; ii < RI; ++ii)
#pragma unroll
            for (int jj = 0; jj < JJ; ++jj)
                if (16 * jj < o) L[prow[ii] + tx + 16 * jj] = acc[ii][jj];
        __syncthreads();
        // W[o + r][c] = -sum_{r' <= r} W_ii[r][r'] G[r'][c]
#pragma unroll
        for (int ii = 0; ii < RI; ++ii)
#pragma unroll
            for (int jj = 0; jj < JJ; ++jj) acc[ii][jj] = 0.0;
#pragma unroll 4
        for (int rq = 0; rq < SB; ++rq) {
            const double* Gq = L + tri(o + rq);
            double a[RI];
#pragma unroll
            for (int ii = 0; ii < RI; ++ii) a[ii] = rq <= ty + TY * ii ? L[prow[ii] + o + rq] : 0.0;
#pragma unroll
            for (int jj = 0; jj < JJ; ++jj)
                if (16 * jj < o) {
                    const double gv = Gq[tx + 16 * jj];
#pragma unroll
                    for (int ii = 0; ii < RI; ++ii) acc[ii][jj] = fma(a[ii], gv, acc[ii][jj]);
                }
        }
        __syncthreads();                          // all of G has been read
#pragma unroll
        for (int ii = 0; ii < RI; ++ii)
#pragma unroll
            for (int jj = 0; jj < JJ; ++jj)
                if (16 * jj < o) L[prow[ii] + tx + 16 * jj] = -acc[ii][jj];
        __syncthreads();
    }

    small_mark(clk, 5);
    // 4. z = W y (a warp per row), alpha = W^T z (a thread per column)
    for (int i = warp; i < n; i += NW) {
        const double* Wi = L + tri(i);
        double s = 0.0;
        for (int j = lane; j <= i; j += 32) s = fma(Wi[j], yv[j], s);
        s = warp_sum(s);
        if (lane == 0) z[i] = s;
    }
    __syncthreads();
    for (int j = tid; j < n; j += NT) {
        double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
        int k = j;
        for (; k + 3 < n; k += 4) {
            s0 = fma(L[tri(k) + j], z[k], s0);
            s1 = fma(L[tri(k + 1) + j], z[k + 1], s1);
            s2 = fma(L[tri(k + 2) + j], z[k + 2], s2);
            s3 = fma(L[tri(k + 3) + j], z[k + 3], s3);
        }
        for (; k < n; ++k) s0 = fma(L[tri(k) + j], z[k], s0);
        al[j] = (s0 + s1) + (s2 + s3);
    }
    __syncthreads();

    small_mark(clk, 6);
    // 5. K^-1 = W^T W in 6 x 6 register tiles fused with the gradient traces; 6. LML
    double s[5] = {0.0, 0.0, 0.0, 0.0, 0.0};     // s0, s1, s2, y'alpha, sum log L_ii
    for (int rc0 = 0; rc0 < n; rc0 += CR)
        for (int cc0 = 0; cc0 < rc0 + CR && cc0 < n; cc0 += 96) {
            double acc[6][6];
            int ir[6], jc[6];
#pragma unroll
            for (int ii = 0; ii < 6; ++ii) {
                ir[ii] = rc0 + ty + TY * ii;
                jc[ii] = cc0 + tx + 16 * ii;
#pragma unroll
                for (int jj = 0; jj < 6; ++jj) acc[ii][jj] = 0.0;
            }
#pragma unroll 2
            for (int k = rc0; k < n; ++k) {
                const double* Wk = L + tri(k);
                double wi[6], wj[6];
#pragma unroll
                for (int ii = 0; ii < 6; ++ii) {
                    wi[ii] = ir[ii] <= k ? Wk[ir[ii]] : 0.0;
                    wj[ii] = jc[ii] <= k ? Wk[jc[ii]] : 0.0;
                }
#pragma unroll
                for (int ii = 0; ii < 6; ++ii)
#pragma unroll
                    for (int jj = 0; jj < 6; ++jj) acc[ii][jj] = fma(wi[ii], wj[jj], acc[ii][jj]);
            }
#pragma unroll
            for (int ii = 0; ii < 6; ++ii)
#pragma unroll
                for (int jj = 0; jj < 6; ++jj) {
                    const int i = ir[ii], j = jc[jj];
                    if (i < m && j <= i) {
                        const double w = al[i] * al[j] - acc[ii][jj];
                        if (i == j) {
                            s[0] += w * sig2;
                            s[2] += w;
                        } else {
                            double kr, dkl;
                            small_kernel_pair<FAM>(sig2, x[i] - x[j], kr, dkl);
                            s[0] += 2.0 * (w * kr);
                            s[1] += 2.0 * (w * dkl);
                        }
                    }
                }
        }
    for (int i = tid; i < m; i += NT) {
        s[3] += yv[i] * al[i];
        s[4] -= log(dinv[i]);                     // log L_ii = -log(1 / L_ii)
    }
    double tot[5];
    small_block_sum<NT, 5>(s, red, tot);
    if (tid == 0) {
        const double val = -0.5 * tot[3] - tot[4] - 0.5 * m * 1.8378770664093453;   // log(2 pi), _gpr.py:613-617
        const bool ok = isfinite(val);
        out[0] = ok ? val : -INFINITY;
        out[1] = ok ? 0.5 * tot[0] : 0.0;
        out[2] = ok ? 0.5 * tot[1] : 0.0;
        out[3] = ok ? 0.5 * chi * tot[2] : 0.0;
        *st = ok ? 0 : 1;
    }
    __syncthreads();
    small_mark(clk, 7);
}

// Fixed-theta entry: one CTA per pair.
template <int NT, int FAM>
__global__ void __launch_bounds__(NT, NT == 64 ? 6 : 1)
small_lml_grad_kernel(SmallProblem pr, const double* __restrict__ theta, const int* __restrict__ gp_of, int B,
                      double* __restrict__ lml, double* __restrict__ grad, int* __restrict__ status) {
    extern __shared__ __align__(16) double sm[];
    __shared__ double s_out[4];
    __shared__ int s_st;
    for (int p = blockIdx.x; p < B; p += gridDim.x) {
        const int gp = gp_of ? gp_of[p] : p;
        small_eval<NT, FAM>(pr, gp, theta[3 * p], theta[3 * p + 1], theta[3 * p + 2], sm, s_out, &s_st);
        if (threadIdx.x == 0) {
            lml[p] = s_out[0];
            if (grad) { grad[3 * p] = s_out[1]; grad[3 * p + 1] = s_out[2]; grad[3 * p + 2] = s_out[3]; }
            if (status) status[p] = s_st;
        }
        __syncthreads();
    }
}

// The whole multi-start fit: persistent CTAs pull pairs from `counter` and run each pair's L-BFGS-B to termination.
template <int NT, int FAM>
__global__ void __launch_bounds__(NT, NT == 64 ? 6 : 1)
small_fit_kernel(SmallProblem pr, const double* __restrict__ starts, const int* __restrict__ gp_of, int B, SmallBox box,
                 LbOptions o, int* __restrict__ counter, double* __restrict__ theta_opt, double* __restrict__ fun,
                 int* __restrict__ nfev, int* __restrict__ nit, int* __restrict__ opt_status) {
    extern __shared__ __align__(16) double sm[];
    __shared__ __align__(8) unsigned char optbuf[sizeof(Lbfgsb)];
    __shared__ double s_out[4], s_theta[3];
    __shared__ int s_st, s_pair, s_run;
    Lbfgsb& opt = *reinterpret_cast<Lbfgsb*>(optbuf);
    for (;;) {
        if (threadIdx.x == 0) s_pair = atomicAdd(counter, 1);
        __syncthreads();
        const int p = s_pair;
        if (p >= B) break;
        const int gp = gp_of ? gp_of[p] : p;
        if (threadIdx.x == 0) {
            opt.init(starts + 3 * (long)p, box.lo, box.hi, o);
            s_theta[0] = opt.x[0]; s_theta[1] = opt.x[1]; s_theta[2] = opt.x[2];
        }
        __syncthreads();
        for (;;) {
            small_eval<NT, FAM>(pr, gp, s_theta[0], s_theta[1], s_theta[2], sm, s_out, &s_st);
            if (threadIdx.x == 0) {
                const long long c0 = pr.dbg ? small_now() : 0;
                double gneg[3] = {-s_out[1], -s_out[2], -s_out[3]};
                opt.feed(-s_out[0], gneg);        // obj_func = (-lml, -grad), _gpr.py:300-307
                if (pr.dbg) {
                    atomicAdd(reinterpret_cast<unsigned long long*>(pr.dbg + 8), (unsigned long long)(small_now() - c0));
                    atomicAdd(reinterpret_cast<unsigned long long*>(pr.dbg + 9), 1ULL);
                }
                s_run = opt.running() ? 1 : 0;
                s_theta[0] = opt.x[0]; s_theta[1] = opt.x[1]; s_theta[2] = opt.x[2];
            }
            __syncthreads();
            if (!s_run) break;
        }
        if (threadIdx.x == 0) {
            theta_opt[3 * (long)p] = opt.x[0]; theta_opt[3 * (long)p + 1] = opt.x[1]; theta_opt[3 * (long)p + 2] = opt.x[2];
            fun[p] = opt.f;
            nfev[p] = opt.nfev; nit[p] = opt.nit; opt_status[p] = opt.status;
        }
        __syncthreads();
    }
}

}  // namespace gpbo
