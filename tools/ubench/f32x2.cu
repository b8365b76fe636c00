// Micro-benchmark: issue throughput of packed FP32 (FMUL2 / FADD2 / FFMA2, sm_100) against scalar FMUL / FADD,
// alone and mixed with ALU-pipe integer work.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/f32x2 tools/ubench/f32x2.cu
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ITERS = 4096, ILP = 8;

template <int MODE>
__global__ void __launch_bounds__(1024) k(float* out, float seed) {
    float2 a[ILP];
    unsigned u[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) { a[i] = make_float2(seed + i + threadIdx.x, seed - i); u[i] = threadIdx.x * 7 + i; }
    const float2 m = make_float2(1.0000001f, 0.9999999f);
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            if (MODE == 0) { a[i].x = __fmul_rn(a[i].x, m.x); a[i].y = __fmul_rn(a[i].y, m.y); }          // 2 scalar FMUL
            if (MODE == 1) { a[i] = __fmul2_rn(a[i], m); }                                                  // 1 FMUL2
            if (MODE == 2) { a[i].x = __fadd_rn(a[i].x, m.x); a[i].y = __fadd_rn(a[i].y, m.y); }          // 2 scalar FADD
            if (MODE == 3) { a[i] = __fadd2_rn(a[i], m); }                                                  // 1 FADD2
            if (MODE == 4) { a[i] = __fmul2_rn(a[i], m); u[i] = (u[i] ^ (u[i] >> 3)) + 0x9e3779b9u; }      // FMUL2 + 2 ALU (LOP3/SHF + IADD)
            if (MODE == 5) { a[i].x = __fmul_rn(a[i].x, m.x); a[i].y = __fmul_rn(a[i].y, m.y); u[i] = (u[i] ^ (u[i] >> 3)) + 0x9e3779b9u; }
            if (MODE == 6) { a[i] = __ffma2_rn(a[i], m, m); }                                               // 1 FFMA2
            if (MODE == 7) { a[i].x = __fmaf_rn(a[i].x, m.x, m.x); a[i].y = __fmaf_rn(a[i].y, m.y, m.y); } // 2 FFMA
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += a[i].x + a[i].y + (float)u[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char* name, float* d) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<148 * 2, 1024>>>(d, 1.0f);
    cudaEventRecord(e0);
    k<MODE><<<148 * 2, 1024>>>(d, 1.0f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double pairs = 148.0 * 2 * 1024 * (double)ITERS * ILP;   // (x, y) element pairs processed
    printf("%-28s %8.3f ms  %8.1f G pair-ops/s  (%.2f pair-ops/clk/SM at 1.9 GHz)\n", name, ms, pairs / ms * 1e-6, pairs / (ms * 1e-3) / 148 / 1.9e9);
}

int main() {
    float* d; cudaMalloc(&d, 148 * 2 * 1024 * 4);
    run<0>("2x FMUL", d); run<1>("FMUL2", d); run<2>("2x FADD", d); run<3>("FADD2", d);
    run<7>("2x FFMA", d); run<6>("FFMA2", d);
    run<5>("2x FMUL + 3 ALU", d); run<4>("FMUL2 + 3 ALU", d);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
