// Main-loop sweep for the FP64 DMMA tile engine (scratch tool, not product code).
//
// Replays the K^-1 tile product of lauum_grad_kernel (tile (I,J) of W^T W over P matrices of
// 4096 x 4096, T = 32) with different staging schemes and prints TFLOP/s per variant, next to the
// DMMA issue-rate peak measured with the same CTA shape (8 warps, 1 CTA / SM).
//
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -o mma_sweep mma_sweep.cu
//   ./mma_sweep [P]
#include <cuda_runtime.h>
#include <stdint.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

constexpr int TB = 128;
constexpr int NTHR = 256;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

struct Acc { double v[8][4][2]; };

__device__ __forceinline__ void dmma884(double (&d)[2], double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d[0]), "+d"(d[1]) : "d"(a), "d"(b));
}
__device__ __forceinline__ void cp_async16(void* s, const void* g) {
    uint32_t a = (uint32_t)__cvta_generic_to_shared(s);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(a), "l"(g));
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N> __device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

__device__ __forceinline__ void mbar_init(uint64_t* b, int count) {
    uint32_t a = (uint32_t)__cvta_generic_to_shared(b);
    asm volatile("mbarrier.init.shared.b64 [%0], %1;\n" ::"r"(a), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* b) {
    uint32_t a = (uint32_t)__cvta_generic_to_shared(b);
    asm volatile("{ .reg .b64 t; mbarrier.arrive.shared.b64 t, [%0]; }\n" ::"r"(a) : "memory");
}
__device__ __forceinline__ void mbar_cp_arrive(uint64_t* b) {
    uint32_t a = (uint32_t)__cvta_generic_to_shared(b);
    asm volatile("cp.async.mbarrier.arrive.noinc.shared.b64 [%0];\n" ::"r"(a) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, int parity) {
    uint32_t a = (uint32_t)__cvta_generic_to_shared(b);
    asm volatile(
        "{\n .reg .pred p;\n WAIT_%=:\n mbarrier.try_wait.parity.shared.b64 p, [%0], %1;\n @p bra DONE_%=;\n bra WAIT_%=;\n DONE_%=:\n}\n"
        ::"r"(a), "r"(parity) : "memory");
}

template <int BK>
__device__ __forceinline__ void stage_slice(double* s, const double* __restrict__ g, long ld, int tid) {
    constexpr int LDT = BK + 4;
    constexpr int CH = BK / 2;                 // 16-byte chunks per row
#pragma unroll
    for (int i = 0; i < TB * CH / NTHR; ++i) {
        const int q = tid + i * NTHR;
        const int r = q / CH, cc = q % CH;
        cp_async16(s + r * LDT + cc * 2, g + (long)r * ld + cc * 2);
    }
}

template <int LDT>
__device__ __forceinline__ void mma_ks(Acc& acc, const double* sa, const double* sb, int ks) {
    double a[8], b[4];
#pragma unroll
    for (int mi = 0; mi < 8; ++mi) a[mi] = sa[mi * 8 * LDT + ks * 4];
#pragma unroll
    for (int ni = 0; ni < 4; ++ni) b[ni] = sb[ni * 8 * LDT + ks * 4];
#pragma unroll
    for (int mi = 0; mi < 8; ++mi)
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) dmma884(acc.v[mi][ni], a[mi], b[ni]);
}

// MODE 0: wait + __syncthreads, refill, compute (current product scheme)
// MODE 1: wait + __syncthreads, first k-step, refill, remaining k-steps
// MODE 2: mbarrier full/empty ring, no CTA-wide barrier in the loop; refill in the middle of the slice
template <int BK, int NSTAGE, int MODE>
__device__ __forceinline__ void gemm_loop(Acc& acc, const double* A, const double* B, long ld, int nk, double* smem,
                                          uint64_t* bars, int& it0) {
    constexpr int LDT = BK + 4;
    constexpr int STAGE = TB * LDT;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, c = lane & 3, wm = warp >> 2, wn = warp & 3;
    double* sA = smem;
    double* sB = smem + NSTAGE * STAGE;
    if (nk <= 0) return;
    if (MODE == 2) {
        uint64_t* full = bars;
        uint64_t* empty = bars + NSTAGE;
        // it0: number of slices already pushed through the ring by earlier calls (keeps parities consistent)
        for (int s = 0; s < NSTAGE - 1; ++s) {
            if (s < nk) {
                const int gi = it0 + s;
                const int st = gi % NSTAGE;
                if (gi >= NSTAGE) mbar_wait(&empty[st], ((gi / NSTAGE) - 1) & 1);
                stage_slice<BK>(sA + st * STAGE, A + s * BK, ld, tid);
                stage_slice<BK>(sB + st * STAGE, B + s * BK, ld, tid);
                mbar_cp_arrive(&full[st]);
            }
        }
        for (int kt = 0; kt < nk; ++kt) {
            const int gi = it0 + kt;
            const int cs = gi % NSTAGE;
            mbar_wait(&full[cs], (gi / NSTAGE) & 1);
            const double* sa = sA + cs * STAGE + (wm * 64 + g) * LDT + c;
            const double* sb = sB + cs * STAGE + (wn * 32 + g) * LDT + c;
#pragma unroll
            for (int ks = 0; ks < BK / 8; ++ks) mma_ks<LDT>(acc, sa, sb, ks);
            const int nx = kt + NSTAGE - 1;
            if (nx < nk) {
                const int gn = it0 + nx;
                const int st = gn % NSTAGE;
                if (gn >= NSTAGE) mbar_wait(&empty[st], ((gn / NSTAGE) - 1) & 1);
                stage_slice<BK>(sA + st * STAGE, A + nx * BK, ld, tid);
                stage_slice<BK>(sB + st * STAGE, B + nx * BK, ld, tid);
                mbar_cp_arrive(&full[st]);
            }
#pragma unroll
            for (int ks = BK / 8; ks < BK / 4; ++ks) mma_ks<LDT>(acc, sa, sb, ks);
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[cs]);
        }
        it0 += nk;
        return;
    }
#pragma unroll
    for (int s = 0; s < NSTAGE - 1; ++s) {
        if (s < nk) {
            stage_slice<BK>(sA + s * STAGE, A + s * BK, ld, tid);
            stage_slice<BK>(sB + s * STAGE, B + s * BK, ld, tid);
        }
        cp_commit();
    }
    for (int kt = 0; kt < nk; ++kt) {
        cp_wait<NSTAGE - 2>();
        __syncthreads();
        const int cs = kt % NSTAGE;
        const double* sa = sA + cs * STAGE + (wm * 64 + g) * LDT + c;
        const double* sb = sB + cs * STAGE + (wn * 32 + g) * LDT + c;
        const int nx = kt + NSTAGE - 1;
        if (MODE == 1) mma_ks<LDT>(acc, sa, sb, 0);
        if (nx < nk) {
            const int st = nx % NSTAGE;
            stage_slice<BK>(sA + st * STAGE, A + nx * BK, ld, tid);
            stage_slice<BK>(sB + st * STAGE, B + nx * BK, ld, tid);
        }
        cp_commit();
#pragma unroll
        for (int ks = (MODE == 1 ? 1 : 0); ks < BK / 4; ++ks) mma_ks<LDT>(acc, sa, sb, ks);
    }
    cp_wait<0>();
    __syncthreads();
}

template <int BK, int NSTAGE, int MODE>
__global__ void __launch_bounds__(NTHR, 1) lauum_like(const double* __restrict__ Abase, long mat_stride, int lda, int T,
                                                      int ntiles, double* out) {
    extern __shared__ __align__(16) double smem[];
    __shared__ uint64_t bars[2 * NSTAGE];
    const int p = blockIdx.x / ntiles, q = blockIdx.x % ntiles;
    int I = (int)((sqrt(8.0 * q + 1.0) - 1.0) * 0.5);
    while ((I + 1) * (I + 2) / 2 <= q) ++I;
    while (I * (I + 1) / 2 > q) --I;
    const int J = q - I * (I + 1) / 2;
    const double* Ap = Abase + (long)p * mat_stride;
    const double* UI = Ap + (long)I * TB * lda;
    const double* UJ = Ap + (long)J * TB * lda;
    if (MODE == 2) {
        if (threadIdx.x == 0) {
            for (int s = 0; s < NSTAGE; ++s) { mbar_init(&bars[s], NTHR); mbar_init(&bars[NSTAGE + s], NTHR / 32); }
            asm volatile("fence.mbarrier_init.release.cluster;\n" ::);
        }
        __syncthreads();
    }
    Acc acc;
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) { acc.v[i][j][0] = 0.0; acc.v[i][j][1] = 0.0; }
    int it0 = 0;
    const int nk = (T - I) * (TB / BK);
    gemm_loop<BK, NSTAGE, MODE>(acc, UI + I * TB, UJ + I * TB, lda, nk, smem, bars, it0);
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) s += acc.v[i][j][0] + acc.v[i][j][1];
    if (s == 1234.5678) out[blockIdx.x] = s;
}

__global__ void fill(double* p, size_t n) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) p[i] = 1e-3 * (double)((i * 2654435761u) & 1023) - 0.5;
}

__global__ void __launch_bounds__(NTHR, 1) dmma_peak(int iters, double* out) {
    double acc[32][2];
#pragma unroll
    for (int i = 0; i < 32; ++i) { acc[i][0] = 0.0; acc[i][1] = 0.0; }
    const double a = 1e-3 * (threadIdx.x & 7), b = 1e-3 * (threadIdx.x & 3);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 32; ++i) dmma884(acc[i], a, b);
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 32; ++i) s += acc[i][0] + acc[i][1];
    if (s == 12345.678) out[0] = s;
}

template <int BK, int NSTAGE, int MODE>
void run(const char* name, const double* A, int P, int m, double* out, double peak) {
    const int T = m / TB, ntiles = T * (T + 1) / 2;
    const int smem = NSTAGE * 2 * TB * (BK + 4) * 8;
    auto k = lauum_like<BK, NSTAGE, MODE>;
    CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    k<<<P * ntiles, NTHR, smem>>>(A, (long)m * m, m, T, ntiles, out);   // warm-up
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        CK(cudaEventRecord(e0));
        k<<<P * ntiles, NTHR, smem>>>(A, (long)m * m, m, T, ntiles, out);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    // flops: sum over tiles of 2*128*128*(T-I)*128
    double fl = 0;
    for (int I = 0; I < T; ++I) fl += (double)(I + 1) * 2.0 * TB * TB * (double)(T - I) * TB;
    fl *= P;
    const double tf = fl / (best * 1e-3) / 1e12;
    printf("{\"variant\": \"%s\", \"BK\": %d, \"NSTAGE\": %d, \"MODE\": %d, \"smem\": %d, \"ms\": %.3f, \"tflops\": %.3f, \"frac_of_peak\": %.4f}\n",
           name, BK, NSTAGE, MODE, smem, best, tf, tf / peak);
    fflush(stdout);
}

int main(int argc, char** argv) {
    const int P = argc > 1 ? atoi(argv[1]) : 148;
    const int m = 4096;
    double* A; double* out;
    CK(cudaMalloc(&A, (size_t)P * m * m * 8));
    CK(cudaMalloc(&out, (size_t)P * 1024 * 8));
    fill<<<148 * 8, 256>>>(A, (size_t)P * m * m);
    CK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    dmma_peak<<<148, NTHR>>>(2000, out);
    CK(cudaEventRecord(e0));
    dmma_peak<<<148, NTHR>>>(40000, out);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    const double peak = 148.0 * 8 * 40000.0 * 32 * 512 / (ms * 1e-3) / 1e12;
    printf("{\"variant\": \"dmma_peak_8warps_1cta\", \"ms\": %.3f, \"tflops\": %.3f}\n", ms, peak);
    run<16, 4, 0>("bk16_s4_sync", A, P, m, out, peak);
    run<16, 4, 1>("bk16_s4_sync_late_refill", A, P, m, out, peak);
    run<16, 3, 0>("bk16_s3_sync", A, P, m, out, peak);
    run<32, 3, 0>("bk32_s3_sync", A, P, m, out, peak);
    run<32, 3, 1>("bk32_s3_sync_late_refill", A, P, m, out, peak);
    run<16, 4, 2>("bk16_s4_mbar", A, P, m, out, peak);
    run<16, 5, 2>("bk16_s5_mbar", A, P, m, out, peak);
    run<16, 3, 2>("bk16_s3_mbar", A, P, m, out, peak);
    run<32, 3, 2>("bk32_s3_mbar", A, P, m, out, peak);
    run<8, 8, 2>("bk8_s8_mbar", A, P, m, out, peak);
    run<8, 6, 2>("bk8_s6_mbar", A, P, m, out, peak);
    return 0;
}
