// FP32 pipe micro-benchmark for B200: FFMA vs FFMA2 (f32x2) vs non-FMA FADD/FMUL issue rates.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp32_peak tools/fp32_peak.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, int iters, float s) {
    float a[16];
    unsigned long long p[8];
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = threadIdx.x * 0.001f + i;
#pragma unroll
    for (int i = 0; i < 8; ++i) asm("mov.b64 %0, {%1,%2};" : "=l"(p[i]) : "f"(a[2 * i]), "f"(a[2 * i + 1]));
    unsigned long long ss;
    asm("mov.b64 %0, {%1,%2};" : "=l"(ss) : "f"(s), "f"(s));
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            if (MODE == 0) {
#pragma unroll
                for (int i = 0; i < 16; ++i) a[i] = fmaf(a[i], s, s);
            } else if (MODE == 1) {
#pragma unroll
                for (int i = 0; i < 8; ++i) asm("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(p[i]) : "l"(ss));
            } else if (MODE == 2) {
#pragma unroll
                for (int i = 0; i < 16; ++i) a[i] = __fadd_rn(__fmul_rn(a[i], s), s);
            } else {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    asm("mul.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(ss));
                    asm("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(ss));
                }
            }
        }
    }
    float r = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) r += a[i];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        float lo, hi;
        asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(p[i]));
        r += lo + hi;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <int MODE>
void run(const char* name, double flop_per_inner) {
    float* out;
    cudaMalloc(&out, 148 * 8 * 256 * sizeof(float));
    const int iters = 20000;
    k<MODE><<<148 * 8, 256>>>(out, 100, 1.0001f);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<MODE><<<148 * 8, 256>>>(out, iters, 1.0001f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    double flops = (double)148 * 8 * 256 * iters * 8 * flop_per_inner;
    printf("%-28s %8.3f ms  %7.2f TFLOP/s\n", name, ms, flops / ms * 1e-9);
    cudaFree(out);
}

int main() {
    run<0>("FFMA  (16 indep chains)", 16 * 2);
    run<1>("FFMA2 (8 packed chains)", 16 * 2);
    run<2>("FMUL+FADD scalar", 16 * 2);
    run<3>("FMUL2+FADD2 packed", 16 * 2);
    return 0;
}
