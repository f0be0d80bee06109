// Element-wise kernels of the training step that are not part of a layer (sm_100a): the residual sum + LeakyReLU that
// closes a LocalFeatureAggregation block, and the Adam update over the flat parameter buffer.  Both are pure HBM
// streams: 16-byte accesses, grid = a multiple of the SM count.
//
// Reference: randlanet/utils/modules.py:325 (`self.lrelu(self.mlp2(x) + self.shortcut(input))`, LeakyReLU(0.01)) and
// trainer.py:78-81 (torch.optim.Adam, lr 1e-2, default betas / eps, no weight decay, no amsgrad).
#include "common.cuh"

namespace r3d {

__device__ __forceinline__ float lrelu(float v, float slope) { return v > 0.f ? v : v * slope; }

__global__ void __launch_bounds__(256) add_lrelu_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                        float* __restrict__ y, long long n, float slope) {
    const long long n4 = n / 4;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        const float4 u = reinterpret_cast<const float4*>(a)[i], v = reinterpret_cast<const float4*>(b)[i];
        reinterpret_cast<float4*>(y)[i] = make_float4(lrelu(u.x + v.x, slope), lrelu(u.y + v.y, slope),
                                                      lrelu(u.z + v.z, slope), lrelu(u.w + v.w, slope));
    }
    for (long long i = n4 * 4 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        y[i] = lrelu(a[i] + b[i], slope);
}

// d = dy * LeakyReLU'(a + b); the sign of the sum is the sign of y (slope > 0)
__global__ void __launch_bounds__(256) add_lrelu_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ y,
                                                            float* __restrict__ d, long long n, float slope) {
    const long long n4 = n / 4;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        const float4 g = reinterpret_cast<const float4*>(dy)[i], v = reinterpret_cast<const float4*>(y)[i];
        reinterpret_cast<float4*>(d)[i] = make_float4(v.x > 0.f ? g.x : g.x * slope, v.y > 0.f ? g.y : g.y * slope,
                                                      v.z > 0.f ? g.z : g.z * slope, v.w > 0.f ? g.w : g.w * slope);
    }
    for (long long i = n4 * 4 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        d[i] = y[i] > 0.f ? dy[i] : dy[i] * slope;
}

// step counter of a capturable optimiser: a device scalar advanced by a one-thread launch in front of the update, so
// that the whole step can be replayed from a CUDA graph
__global__ void adam_tick_kernel(float* __restrict__ step) { *step += 1.0f; }

// torch.optim.Adam's arithmetic per element (fp32 state; the hyper-parameters arrive as doubles and 1 - beta is formed in
// double before rounding to fp32, as torch's fused kernel does: 1 - float(0.999) is off by 1.3e-5 relative):
//   m = m + (g - m)(1 - b1);  v = b2 v + (1 - b2) g^2;  p -= (lr / (1 - b1^t)) m / (sqrt(v) / sqrt(1 - b2^t) + eps)
__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                   float* __restrict__ v, long long n, const float* __restrict__ lr_dev,
                                                   double lr_host, double beta1, double beta2, double eps_d,
                                                   const float* __restrict__ step) {
    const double t = (double)*step;
    const double lr = lr_dev ? (double)*lr_dev : lr_host;
    const float omb1 = (float)(1.0 - beta1), b2 = (float)beta2, omb2 = (float)(1.0 - beta2), eps = (float)eps_d;
    const float bc2_sqrt = (float)sqrt(1.0 - pow(beta2, t));
    const float step_size = (float)(lr / (1.0 - pow(beta1, t)));
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const float gi = g[i];
        const float mi = m[i] + (gi - m[i]) * omb1;
        const float vi = b2 * v[i] + omb2 * gi * gi;
        m[i] = mi;
        v[i] = vi;
        p[i] -= step_size * mi / (sqrtf(vi) / bc2_sqrt + eps);
    }
}

static int stream_grid(long long work) {
    long long blocks = (work + 255) / 256;
    const long long cap = 8LL * kNumSMs;
    return (int)(blocks < 1 ? 1 : (blocks > cap ? cap : blocks));
}

}  // namespace r3d

using namespace r3d;

extern "C" int r3d_add_lrelu(const float* a, const float* b, float* y, long long n, float slope, r3d_stream_t stream) {
    if (n < 0 || !(slope > 0.f)) return R3D_EINVAL;
    if (n == 0) return R3D_OK;
    if (!a || !b || !y) return R3D_EINVAL;
    if (!is_aligned(a, 16) || !is_aligned(b, 16) || !is_aligned(y, 16)) return R3D_EALIGN;
    add_lrelu_kernel<<<stream_grid(n / 4 + 1), 256, 0, static_cast<cudaStream_t>(stream)>>>(a, b, y, n, slope);
    R3D_LAUNCH_CHECK("add_lrelu_kernel");
    return R3D_OK;
}

extern "C" int r3d_add_lrelu_bwd(const float* dy, const float* y, float* d, long long n, float slope, r3d_stream_t stream) {
    if (n < 0 || !(slope > 0.f)) return R3D_EINVAL;
    if (n == 0) return R3D_OK;
    if (!dy || !y || !d) return R3D_EINVAL;
    if (!is_aligned(dy, 16) || !is_aligned(y, 16) || !is_aligned(d, 16)) return R3D_EALIGN;
    add_lrelu_bwd_kernel<<<stream_grid(n / 4 + 1), 256, 0, static_cast<cudaStream_t>(stream)>>>(dy, y, d, n, slope);
    R3D_LAUNCH_CHECK("add_lrelu_bwd_kernel");
    return R3D_OK;
}

extern "C" int r3d_adam_step(float* p, const float* g, float* m, float* v, long long n, const float* lr_dev, double lr_host,
                             double beta1, double beta2, double eps, float* step, r3d_stream_t stream) {
    if (n < 0) return R3D_EINVAL;
    if (!p || !g || !m || !v || !step) return R3D_EINVAL;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    adam_tick_kernel<<<1, 1, 0, st>>>(step);
    R3D_LAUNCH_CHECK("adam_tick_kernel");
    if (n == 0) return R3D_OK;
    adam_kernel<<<stream_grid(n), 256, 0, st>>>(p, g, m, v, n, lr_dev, lr_host, beta1, beta2, eps, step);
    R3D_LAUNCH_CHECK("adam_kernel");
    return R3D_OK;
}
