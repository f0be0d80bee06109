// FP32-pipe peak probe: the roofline denominator of the CUDA-core kernels (KNN, fused LFA GEMMs).
// MEASURED_PEAKS.json holds HBM and bf16 tensor peaks only, so bench.py measures this one live, the same
// way (a kernel that does nothing but the instruction in question, timed with CUDA events).
//   mode 0: scalar FFMA, 16 independent chains per thread      (2 flop / lane / issue)
//   mode 1: packed FFMA2 (fma.rn.f32x2), 8 independent chains  (4 flop / lane / issue — Blackwell)
#include "common.cuh"

namespace r3d {

template <int MODE>
__global__ void __launch_bounds__(256) fp32_probe_kernel(float* out, int iters, float s) {
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
            } else {
#pragma unroll
                for (int i = 0; i < 8; ++i) asm("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(p[i]) : "l"(ss));
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

}  // namespace r3d

using namespace r3d;

// out: device buffer of r3d_fp32_probe_floats() floats.  *flops receives the flop count of the launch.
extern "C" size_t r3d_fp32_probe_floats(void) { return (size_t)kNumSMs * 8 * 256; }

extern "C" int r3d_fp32_probe(int mode, int iters, float* out, double* flops, r3d_stream_t stream) {
    if (!out || iters <= 0 || mode < 0 || mode > 1) return R3D_EINVAL;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (mode == 0)
        fp32_probe_kernel<0><<<kNumSMs * 8, 256, 0, st>>>(out, iters, 1.0001f);
    else
        fp32_probe_kernel<1><<<kNumSMs * 8, 256, 0, st>>>(out, iters, 1.0001f);
    R3D_LAUNCH_CHECK("fp32_probe_kernel");
    if (flops) *flops = (double)kNumSMs * 8 * 256 * (double)iters * 8 * 32;
    return R3D_OK;
}
