// FP64 DMMA tile engine: 128x128 output tile per CTA, 8 warps (2x4), warp tile 64x32,
// k-slices of 16 staged with cp.async into a 4-deep shared ring (see common.cuh).
//
// Stage hand-over uses two mbarriers per stage instead of __syncthreads():
//   full[s]  (count 256): every thread issues its share of the slice with cp.async and then
//                         cp.async.mbarrier.arrive.noinc -> the phase completes when all copies landed;
//   empty[s] (count 8):   each warp arrives after its last DMMA that reads the stage.
// A warp refills the stage that was consumed one slice earlier in the MIDDLE of the current slice's
// DMMA block, so the wait on empty[] is practically never blocking and warps may drift by half a slice.
// Measured on B200 (tools/mma_sweep.cu, K^-1 tile product at m = 4096): 86.5 % of the DMMA issue peak with
// the __syncthreads ring, 95.9 % with this one (profiles/r01_mma_sweep.txt).
#pragma once
#include <type_traits>

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

// ---- the stage ring ---------------------------------------------------------------------------
// `count` = slices pushed through the ring so far by this CTA; it is carried across successive loops of
// one kernel so that barrier phase parities stay consistent without re-initialising the barriers.
struct Ring {
    uint64_t* full;
    uint64_t* empty;
    int count;
};

// bars: 2 * NSTAGE mbarriers in (static) shared memory.  Ends with __syncthreads().
__device__ __forceinline__ void ring_init(Ring& r, uint64_t* bars) {
    r.full = bars;
    r.empty = bars + NSTAGE;
    r.count = 0;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < NSTAGE; ++s) {
            mbar_init(&bars[s], NTHR);
            mbar_init(&bars[NSTAGE + s], NTHR / 32);
        }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::);
    }
    __syncthreads();
}

// Runs nk slices through the ring.  stage(st, kt): issue this thread's cp.async for slice kt into stage st.
// compute(st, kt, half): the DMMA work of slice kt (half = integral_constant 0 / 1: first / second BK/8 k-steps).
template <class STAGE, class COMPUTE>
__device__ __forceinline__ void ring_pipeline(Ring& ring, int nk, STAGE stage, COMPUTE compute) {
    if (nk <= 0) return;
    const int lane = threadIdx.x & 31;
    const int base = ring.count;
    auto push = [&](int kt) {
        const int gi = base + kt;
        const int st = gi % NSTAGE;
        if (gi >= NSTAGE) mbar_wait(&ring.empty[st], ((gi / NSTAGE) - 1) & 1);
        stage(st, kt);
        mbar_cp_arrive(&ring.full[st]);
    };
#pragma unroll
    for (int s = 0; s < NSTAGE - 1; ++s)
        if (s < nk) push(s);
    for (int kt = 0; kt < nk; ++kt) {
        const int gi = base + kt;
        const int cs = gi % NSTAGE;
        mbar_wait(&ring.full[cs], (gi / NSTAGE) & 1);
        compute(cs, kt, std::integral_constant<int, 0>{});
        const int nx = kt + NSTAGE - 1;
        if (nx < nk) push(nx);
        compute(cs, kt, std::integral_constant<int, 1>{});
        __syncwarp();
        if (lane == 0) mbar_arrive(&ring.empty[cs]);
    }
    ring.count = base + nk;
}

// Stage one 128 x 16 operand slice (rows `ld` apart in global memory) into padded shared rows.
__device__ __forceinline__ void stage_slice(double* s, const double* __restrict__ g, long ld, int tid) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int q = tid + i * NTHR;
        const int r = q >> 3, cc = q & 7;
        cp_async16(s + r * LDT + cc * 2, g + (long)r * ld + cc * 2);
    }
}

// Two of the four k-steps (HALF = 0: ks 0,1; HALF = 1: ks 2,3) of DMMA on one staged pair of slices.
// sa/sb already point at this thread's first fragment element: s?[(w?*.. + g) * stride + c].
template <int SA_STRIDE, int SB_STRIDE, int HALF>
__device__ __forceinline__ void mma_half(Acc& acc, const double* sa, const double* sb) {
#pragma unroll
    for (int ks = HALF * (BK / 8); ks < (HALF + 1) * (BK / 8); ++ks) {
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

// Where slice kt of the two operands lives in global memory: 128 rows each, `ld*` doubles apart,
// 16 contiguous doubles per row starting at a / b.
struct SliceSrc {
    const double* a;
    long lda;
    const double* b;
    long ldb;
};

// acc += A * B^T over nk k-slices; src(kt) -> SliceSrc.  SAME: both operands are the same rows
// (diagonal tiles) -> staged once (src.b ignored).  On return every warp has finished its DMMAs but other
// warps may still be reading the ring: callers __syncthreads() before re-using `smem` for something else.
template <bool SAME, class SRC>
__device__ __forceinline__ void gemm_nt_loop(Acc& acc, SRC src, int nk, double* smem, Ring& ring,
                                             const ThreadCoord& tc) {
    double* sA = smem;
    double* sB = smem + NSTAGE * STAGE_DBL;
    const int oa = (tc.wm * 64 + tc.g) * LDT + tc.c;
    const int ob = (tc.wn * 32 + tc.g) * LDT + tc.c;
    ring_pipeline(
        ring, nk,
        [&](int st, int kt) {
            const SliceSrc s = src(kt);
            stage_slice(sA + st * STAGE_DBL, s.a, s.lda, tc.tid);
            if (!SAME) stage_slice(sB + st * STAGE_DBL, s.b, s.ldb, tc.tid);
        },
        [&](int st, int, auto half) {
            const double* sa = sA + st * STAGE_DBL + oa;
            const double* sb = (SAME ? sA : sB) + st * STAGE_DBL + ob;
            mma_half<LDT, LDT, decltype(half)::value>(acc, sa, sb);
        });
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

// out += S * Dm^T : S is a shared 128x128 tile [r][k] (stride LDS), complete and visible to the CTA;
// Dm is a 128x128 global block (row stride 128) whose rows [n][k] are streamed through `stages`.
__device__ __forceinline__ void epi_product_SxDt(Acc& out, const double* S, const double* __restrict__ Dm,
                                                 double* stages, Ring& ring, const ThreadCoord& tc) {
    const double* sa0 = S + (tc.wm * 64 + tc.g) * LDS + tc.c;
    const int ob = (tc.wn * 32 + tc.g) * LDT + tc.c;
    ring_pipeline(
        ring, TB / BK, [&](int st, int kt) { stage_slice(stages + st * STAGE_DBL, Dm + kt * BK, TB, tc.tid); },
        [&](int st, int kt, auto half) {
            mma_half<LDS, LDT, decltype(half)::value>(out, sa0 + kt * BK, stages + st * STAGE_DBL + ob);
        });
}

// out += Dm * G : Dm is a 128x128 global block [r][k] streamed through `stages`;
// G is a shared 128x128 tile stored [k][n] (stride LDS), complete and visible to the CTA.
__device__ __forceinline__ void epi_product_DxG(Acc& out, const double* __restrict__ Dm, const double* G,
                                                double* stages, Ring& ring, const ThreadCoord& tc) {
    const int oa = (tc.wm * 64 + tc.g) * LDT + tc.c;
    ring_pipeline(
        ring, TB / BK, [&](int st, int kt) { stage_slice(stages + st * STAGE_DBL, Dm + kt * BK, TB, tc.tid); },
        [&](int st, int kt, auto half) {
            constexpr int H = decltype(half)::value;
            const double* sa = stages + st * STAGE_DBL + oa;
            // B fragment: element (k = kt*16 + ks*4 + c, n = wn*32 + ni*8 + g) of G[k][n]
            const double* sb = G + (kt * BK + tc.c) * LDS + tc.wn * 32 + tc.g;
#pragma unroll
            for (int ks = H * (BK / 8); ks < (H + 1) * (BK / 8); ++ks) {
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
        });
}

}  // namespace gpbo
