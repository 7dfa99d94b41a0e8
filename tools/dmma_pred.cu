// Does a predicated-off / branched-over DMMA free FP64 tensor-pipe time?  (scratch probe)
#include <cuda_runtime.h>
#include <cstdio>
__device__ __forceinline__ void dmma(double (&d)[2], double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(d[0]), "+d"(d[1]) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma_br(double (&d)[2], double a, double b, int live) {
    asm volatile("{\n .reg .pred p;\n setp.ne.s32 p, %4, 0;\n @!p bra.uni SKIP_%=;\n"
                 " mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n SKIP_%=:\n}\n"
                 : "+d"(d[0]), "+d"(d[1]) : "d"(a), "d"(b), "r"(live));
}
// MODE 0: all 32 live; 1: half predicated off (if); 2: half skipped with bra.uni per DMMA; 3: half skipped, one branch around 16
template <int MODE>
__global__ void __launch_bounds__(256, 1) k(int iters, int mask, double* out) {
    double acc[32][2];
#pragma unroll
    for (int i = 0; i < 32; ++i) { acc[i][0] = 0; acc[i][1] = 0; }
    const double a = 1e-3 * (threadIdx.x & 7), b = 1e-3 * (threadIdx.x & 3);
    for (int it = 0; it < iters; ++it) {
        if (MODE == 3) {
#pragma unroll
            for (int i = 0; i < 16; ++i) dmma(acc[i], a, b);
            if (mask & 1) {
#pragma unroll
                for (int i = 16; i < 32; ++i) dmma(acc[i], a, b);
            }
        } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                const int live = (MODE == 0) ? 1 : ((mask >> (i & 1)) & 1) ^ 1 ? 1 : 0;   // mask=2: even live, odd dead
                if (MODE == 2) dmma_br(acc[i], a, b, live);
                else if (live) dmma(acc[i], a, b);
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 32; ++i) s += acc[i][0] + acc[i][1];
    if (s == 123.456) out[0] = s;
}
template <int MODE> void run(const char* name, int mask, double* out) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<148, 256>>>(1000, mask, out);
    cudaEventRecord(e0); k<MODE><<<148, 256>>>(20000, mask, out); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("%-28s %.3f ms\n", name, ms);
}
int main() { double* out; cudaMalloc(&out, 64);
    run<0>("all 32 live", 0, out); run<1>("16 predicated off", 2, out); run<2>("16 skipped bra.uni each", 2, out);
    run<3>("16 skipped one branch", 0, out); run<3>("mode3 all live", 1, out); return 0; }
