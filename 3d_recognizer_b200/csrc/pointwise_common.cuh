// Argument block and row addressing shared by the per-point layer kernels (pointwise.cu: FP32 CUDA-core kernels,
// pointwise_tc.cu: tcgen05 3xTF32 kernel).
#pragma once
#include "common.cuh"

namespace r3d {

// Fused train-mode BatchNorm (+ activation) behind the layer: after the batch statistics are complete (grid barrier of
// a cooperative launch) every thread normalises the outputs it still holds in registers.  y == nullptr: not fused.
struct PwBnArgs {
    float* y;                  // (M, cout) dense: act(a z + c)
    const float* gamma;
    const float* beta;
    const float* bias;         // nullable conv bias (only enters the running mean)
    float eps;
    float momentum;
    float* running_mean;       // nullable
    float* running_var;
    long long* num_batches;    // nullable
    float* save;               // (3, cout): a, mean, rstd
    int act;
    float slope;
};

struct PwArgs {
    const float* xa;
    long long xa_bstride;
    int ca;
    const int32_t* gidx;       // nullable; row n of cloud b reads xa row gidx[b*gidx_bstride + n]
    long long gidx_bstride;    // 0 => the same index vector for every cloud
    const float* xb;           // nullable second source
    long long xb_bstride;
    int cb;
    const float* wT;           // (ca+cb, cout)
    const float* scale;        // nullable (cout)
    const float* shift;        // nullable (cout)
    int act;                   // 0 none, 1 relu, 2 leaky relu
    float slope;
    float* y;
    long long y_bstride;
    int y_ld;                  // floats between consecutive output rows (>= cout): lets a layer write a channel slice
    int cout;
    int B;
    int n;                     // rows per cloud
    int transpose_out;         // 1: write y as (B, cout, n) — the reference's logits layout (modules.py:611)
    double* stats;             // nullable (2*cout): += per-channel sum and sum of squares of the written values
    int w_out_in;              // 0: wT is (cin, cout);  1: the weight is stored (cout, cin) (conv / Linear layout)
    PwBnArgs bn;               // fused BatchNorm of the train-mode forward (r3d_pointwise_bn); zero otherwise
};

// BatchNorm affine of channel c from the finished batch statistics (same arithmetic as bn_apply_kernel); the writer
// thread also stores save[] and updates the running statistics.
__device__ __forceinline__ void pw_bn_coeffs(const PwArgs& a, int c, long long M, bool writer, float& ca, float& cc) {
    const double mean = __ldcg(a.stats + c) / (double)M;
    double var = __ldcg(a.stats + a.cout + c) / (double)M - mean * mean;
    var = var > 0.0 ? var : 0.0;
    const float rstd = (float)(1.0 / sqrt(var + (double)a.bn.eps));
    ca = a.bn.gamma[c] * rstd;
    cc = a.bn.beta[c] - ca * (float)mean;
    if (writer) {
        a.bn.save[c] = ca;
        a.bn.save[a.cout + c] = (float)mean;
        a.bn.save[2 * a.cout + c] = rstd;
        if (a.bn.running_mean) {
            const double unbiased = var * ((double)M / (double)(M > 1 ? M - 1 : 1));
            a.bn.running_mean[c] = (1.f - a.bn.momentum) * a.bn.running_mean[c] +
                                   a.bn.momentum * ((float)mean + (a.bn.bias ? a.bn.bias[c] : 0.f));
            a.bn.running_var[c] = (1.f - a.bn.momentum) * a.bn.running_var[c] + a.bn.momentum * (float)unbiased;
        }
    }
}

__device__ __forceinline__ float apply_act(float v, int act, float slope) {
    if (act == 1) return fmaxf(v, 0.f);
    if (act == 2) return v > 0.f ? v : v * slope;
    return v;
}

__device__ __forceinline__ const float* src_row(const PwArgs& a, int b, int n, bool second) {
    if (second) return a.xb + (size_t)b * a.xb_bstride + (size_t)n * a.cb;
    int r = n;
    if (a.gidx) r = a.gidx[(size_t)b * a.gidx_bstride + n];
    return a.xa + (size_t)b * a.xa_bstride + (size_t)r * a.ca;
}

// pointwise_tc.cu: true if the tensor-core kernel can take this layer
bool pw_tc_eligible(const PwArgs& a, bool force);
int bn_fused_mask();     // r3d_bn_set_fused (pointwise_train.cu)
int pw_tc_launch(const PwArgs& a, cudaStream_t st);

}  // namespace r3d
