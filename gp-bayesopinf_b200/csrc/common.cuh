// Shared constants and device helpers for the sm_100a kernels of the step2_fitgps hot path.
//
// Tile geometry: every dense contraction on the path is expressed as an
// "NT" product  acc[128x128] += A[128 x k] * B[128 x k]^T  whose operands are rows of
// row-major FP64 matrices with k contiguous.  Operand k-slices (16 doubles = one 128-byte
// line per row) are staged global->shared with cp.async (LDGSTS) into a 4-deep ring whose stages are
// handed over with mbarriers (full: cp.async.mbarrier.arrive of all 256 threads; empty: one arrive per
// warp), so the k-loop has no CTA-wide barrier, and fed to the FP64 tensor pipe with
// mma.sync.m8n8k4.f64 (SASS: DMMA.8x8x4).  tcgen05.mma has no f64 kind on sm_100a, so the DMMA pipe
// is the FP64 tensor path on B200.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

#include "fastmath.h"

namespace gpbo {

constexpr int TB = 128;           // tile edge (rows and columns of an output tile)
constexpr int BK = 16;            // k-slice per pipeline stage (16 doubles = 128 B per row)
constexpr int LDT = BK + 4;       // padded shared row stride of a k-slice: conflict-free LDS.64 fragment reads
constexpr int LDS = TB + 4;       // padded shared row stride of a staged 128x128 tile used as an MMA operand
constexpr int LDP = TB + 1;       // odd stride of the 128x128 tile used by the in-shared potf2/trtri
constexpr int NSTAGE = 4;         // cp.async ring depth (main loop and epilogue product share the ring barriers)
constexpr int NTHR = 256;         // 8 warps: 2 (rows) x 4 (cols), warp tile 64 x 32
constexpr int STAGE_DBL = TB * LDT;
constexpr int MAIN_SMEM = NSTAGE * 2 * STAGE_DBL * 8;                 // 163840 B
constexpr int EPI_SMEM = TB * LDS * 8 + NSTAGE * STAGE_DBL * 8;       // 217088 B
constexpr int TILE_SMEM = EPI_SMEM > MAIN_SMEM ? EPI_SMEM : MAIN_SMEM;

// Per (GP, start) pair hyper-parameters in natural units, produced by prep_pairs_kernel.
struct PairParams {
    double sig2;      // sigma^2  = exp(theta[0])
    double ell;       // ell      = exp(theta[1])
    double chi;       // chi      = exp(theta[2])
    double inv_ell2;  // 1 / ell^2 (correctly rounded reciprocal of ell*ell, for gpbo_div)
    int gp;           // index of the GP (mode) whose (t, y) this pair uses
    int pad;
};

struct Acc {
    double v[8][4][2];   // [m-fragment][n-fragment][element]; row = wm*64+mi*8+g, col = wn*32+ni*8+2c+e
};

__device__ __forceinline__ void acc_zero(Acc& a) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) { a.v[i][j][0] = 0.0; a.v[i][j][1] = 0.0; }
}

// D(8x8) += A(8x4, row) * B(4x8, col) on the FP64 tensor pipe.
__device__ __forceinline__ void dmma884(double (&d)[2], double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d[0]), "+d"(d[1])
                 : "d"(a), "d"(b));
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    uint32_t s = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

// ---- mbarrier (shared::cta) ------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* b, int count) {
    uint32_t a = (uint32_t)__cvta_generic_to_shared(b);
    asm volatile("mbarrier.init.shared.b64 [%0], %1;\n" ::"r"(a), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* b) {
    uint32_t a = (uint32_t)__cvta_generic_to_shared(b);
    asm volatile("{ .reg .b64 t; mbarrier.arrive.shared.b64 t, [%0]; }\n" ::"r"(a) : "memory");
}
// arrive once all cp.async issued so far by this thread have landed (does not bump the pending count)
__device__ __forceinline__ void mbar_cp_arrive(uint64_t* b) {
    uint32_t a = (uint32_t)__cvta_generic_to_shared(b);
    asm volatile("cp.async.mbarrier.arrive.noinc.shared.b64 [%0];\n" ::"r"(a) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, int parity) {
    uint32_t a = (uint32_t)__cvta_generic_to_shared(b);
    asm volatile(
        "{\n .reg .pred p;\n WAIT_%=:\n mbarrier.try_wait.parity.shared.b64 p, [%0], %1;\n"
        " @p bra DONE_%=;\n bra WAIT_%=;\n DONE_%=:\n}\n" ::"r"(a), "r"(parity)
        : "memory");
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Deterministic block-wide sum of up to NV values per thread (fixed shuffle tree + fixed warp order).
template <int NV>
__device__ __forceinline__ void block_sum(double (&v)[NV], double* red /* >= NV*8 doubles */, double (&out)[NV]) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        double s = warp_sum(v[i]);
        if (lane == 0) red[i * 8 + warp] = s;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < NTHR / 32; ++w) s += red[i * 8 + w];
        out[i] = s;
    }
    __syncthreads();
}

}  // namespace gpbo
