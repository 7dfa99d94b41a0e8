// FP64 DMMA tile engine: 128x128 output tile per CTA, 8 warps (2x4), warp tile 64x32,
// k-slices of 16 staged with cp.async into a 4-deep shared ring (see common.cuh).
#pragma once
#include "common.cuh"

namespace gpbo {

struct ThreadCoord {
    int tid, lane, warp, g, c, wm, wn;
    __device__ __forceinline__ ThreadCoord() {
        tid = threadIdx.x; lane = tid & 31; warp = tid >> 5;
        g = lane >> 2; c = lane & 3; wm = warp >> 2; wn = warp & 3;
    }
    // accumulator element (mi, ni, e) -> tile row / col
    __device__ __forceinline__ int row(int mi) const { return wm * 64 + mi * 8 + g; }
    __device__ __forceinline__ int col(int ni, int e) const { return wn * 32 + ni * 8 + 2 * c + e; }
};

// Stage one 128 x 16 operand slice (rows `ld` apart in global memory) into padded shared rows.
__device__ __forceinline__ void stage_slice(double* s, const double* __restrict__ g, long ld, int tid) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int q = tid + i * NTHR;
        const int r = q >> 3, cc = q & 7;
        cp_async16(s + r * LDT + cc * 2, g + (long)r * ld + cc * 2);
    }
}

// 4 k-steps of DMMA on one staged pair of slices.  sa/sb already point at this thread's
// first fragment element: s?[(w?*.. + g) * stride + c].
template <int SA_STRIDE, int SB_STRIDE>
__device__ __forceinline__ void mma_slice(Acc& acc, const double* sa, const double* sb) {
#pragma unroll
    for (int ks = 0; ks < BK / 4; ++ks) {
        double a[8], b[4];
#pragma unroll
        for (int mi = 0; mi < 8; ++mi) a[mi] = sa[mi * 8 * SA_STRIDE + ks * 4];
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) b[ni] = sb[ni * 8 * SB_STRIDE + ks * 4];
#pragma unroll
        for (int mi = 0; mi < 8; ++mi)
#pragma unroll
            for (int ni = 0; ni < 4; ++ni) dmma884(acc.v[mi][ni], a[mi], b[ni]);
    }
}

// acc += A * B^T over nk k-slices.  fa(kt)/fb(kt) give the global address of (row 0, k = 16*kt)
// of each operand's 128 rows.  SAME: both operands are the same rows (diagonal tiles) -> stage once.
template <bool SAME, class FA, class FB>
__device__ __forceinline__ void gemm_nt_loop(Acc& acc, FA fa, long lda, FB fb, long ldb, int nk, double* smem,
                                             const ThreadCoord& tc) {
    if (nk <= 0) return;
    double* sA = smem;
    double* sB = smem + NSTAGE * STAGE_DBL;
#pragma unroll
    for (int s = 0; s < NSTAGE - 1; ++s) {
        if (s < nk) {
            stage_slice(sA + s * STAGE_DBL, fa(s), lda, tc.tid);
            if (!SAME) stage_slice(sB + s * STAGE_DBL, fb(s), ldb, tc.tid);
        }
        cp_async_commit();
    }
    for (int kt = 0; kt < nk; ++kt) {
        cp_async_wait<NSTAGE - 2>();
        __syncthreads();
        const int nx = kt + NSTAGE - 1;
        if (nx < nk) {
            const int st = nx % NSTAGE;
            stage_slice(sA + st * STAGE_DBL, fa(nx), lda, tc.tid);
            if (!SAME) stage_slice(sB + st * STAGE_DBL, fb(nx), ldb, tc.tid);
        }
        cp_async_commit();
        const int cs = kt % NSTAGE;
        const double* sa = sA + cs * STAGE_DBL + (tc.wm * 64 + tc.g) * LDT + tc.c;
        const double* sb = (SAME ? sA : sB) + cs * STAGE_DBL + (tc.wn * 32 + tc.g) * LDT + tc.c;
        mma_slice<LDT, LDT>(acc, sa, sb);
    }
    cp_async_wait<0>();
    __syncthreads();
}

// Store the accumulator tile into a shared 128x128 tile with row stride LD.
template <int LD>
__device__ __forceinline__ void acc_to_smem(const Acc& acc, double* S, const ThreadCoord& tc) {
#pragma unroll
    for (int mi = 0; mi < 8; ++mi)
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) {
            double* p = S + tc.row(mi) * LD + tc.col(ni, 0);
            p[0] = acc.v[mi][ni][0];
            p[1] = acc.v[mi][ni][1];
        }
}

// out += S * Dm^T : S is a shared 128x128 tile [r][k] (stride LDS); Dm is a 128x128 global block
// (row stride 128) whose rows [n][k] are streamed through `ring` (ESTAGE slices).
__device__ __forceinline__ void epi_product_SxDt(Acc& out, const double* S, const double* __restrict__ Dm,
                                                 double* ring, const ThreadCoord& tc) {
    constexpr int nk = TB / BK;
#pragma unroll
    for (int s = 0; s < ESTAGE - 1; ++s) {
        stage_slice(ring + s * STAGE_DBL, Dm + s * BK, TB, tc.tid);
        cp_async_commit();
    }
    for (int kt = 0; kt < nk; ++kt) {
        cp_async_wait<ESTAGE - 2>();
        __syncthreads();
        const int nx = kt + ESTAGE - 1;
        if (nx < nk) stage_slice(ring + (nx % ESTAGE) * STAGE_DBL, Dm + nx * BK, TB, tc.tid);
        cp_async_commit();
        const double* sa = S + (tc.wm * 64 + tc.g) * LDS + kt * BK + tc.c;
        const double* sb = ring + (kt % ESTAGE) * STAGE_DBL + (tc.wn * 32 + tc.g) * LDT + tc.c;
        mma_slice<LDS, LDT>(out, sa, sb);
    }
    cp_async_wait<0>();
    __syncthreads();
}

// out += Dm * G : Dm is a 128x128 global block [r][k] streamed through `ring`;
// G is a shared 128x128 tile stored [k][n] (stride LDS).
__device__ __forceinline__ void epi_product_DxG(Acc& out, const double* __restrict__ Dm, const double* G,
                                                double* ring, const ThreadCoord& tc) {
    constexpr int nk = TB / BK;
#pragma unroll
    for (int s = 0; s < ESTAGE - 1; ++s) {
        stage_slice(ring + s * STAGE_DBL, Dm + s * BK, TB, tc.tid);
        cp_async_commit();
    }
    for (int kt = 0; kt < nk; ++kt) {
        cp_async_wait<ESTAGE - 2>();
        __syncthreads();
        const int nx = kt + ESTAGE - 1;
        if (nx < nk) stage_slice(ring + (nx % ESTAGE) * STAGE_DBL, Dm + nx * BK, TB, tc.tid);
        cp_async_commit();
        const double* sa = ring + (kt % ESTAGE) * STAGE_DBL + (tc.wm * 64 + tc.g) * LDT + tc.c;
        // B fragment: element (k = kt*16 + ks*4 + c, n = wn*32 + ni*8 + g) of G[k][n]
        const double* sb = G + (kt * BK + tc.c) * LDS + tc.wn * 32 + tc.g;
#pragma unroll
        for (int ks = 0; ks < BK / 4; ++ks) {
            double a[8], b[4];
#pragma unroll
            for (int mi = 0; mi < 8; ++mi) a[mi] = sa[mi * 8 * LDT + ks * 4];
#pragma unroll
            for (int ni = 0; ni < 4; ++ni) b[ni] = sb[ks * 4 * LDS + ni * 8];
#pragma unroll
            for (int mi = 0; mi < 8; ++mi)
#pragma unroll
                for (int ni = 0; ni < 4; ++ni) dmma884(out.v[mi][ni], a[mi], b[ni]);
        }
    }
    cp_async_wait<0>();
    __syncthreads();
}

}  // namespace gpbo
