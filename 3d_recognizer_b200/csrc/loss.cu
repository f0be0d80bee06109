// Focal-Tversky / Dice loss of the training step in three launches (sm_100a).
//
// Reference: randlanet/utils/losses.py:66-86 through the factory trainer.py:245-269 ("dice" = alpha 0.5, gamma 1;
// "tversky" = 0.7, 1; "focal_tversky" = 0.7, 4/3; the background class 0 is neglected).  With p = softmax(logits) over
// the class axis and m_c = [label == c]:
//     TP_c = sum m_c p_c,  FN_c = |m_c| - TP_c,  FP_c = sum p_c - TP_c,
//     TI_c = (TP_c + eps) / (TP_c + alpha FN_c + (1 - alpha) FP_c + eps),  loss = mean_c (1 - TI_c)^gamma.
// As tensor ops this is ~20 launches forward and ~20 backward on the critical chain of a 3 ms step.  Here:
//   tversky_reduce_kernel    per class: sum m_c p_c, sum p_c, |m_c|  (fp64 atomics, one flush per CTA)
//   tversky_finish_kernel    loss, and the coefficients gTP_c = dloss/dTP_c, gSP_c = dloss/d(sum p_c)
//   tversky_bwd_kernel       dlogits = g * p (dp - sum_k p_k dp_k),  dp_c = m_c gTP_c + gSP_c
// logits are addressed through (batch, class, point) strides, so the (B,N,C)-major tensor the network produces is read
// in place (modules.py:611 hands the loss a transposed view).
#include "common.cuh"

namespace r3d {

constexpr int kLossMaxC = 16;

__global__ void __launch_bounds__(256) tversky_reduce_kernel(const float* __restrict__ logits, long long sb, long long sc,
                                                             long long sn, const int64_t* __restrict__ labels, int B,
                                                             int C, int N, double* __restrict__ acc /* (3,C) */) {
    __shared__ double red[3][kLossMaxC];
    if (threadIdx.x < 3 * kLossMaxC) (&red[0][0])[threadIdx.x] = 0.0;
    __syncthreads();
    float tp[kLossMaxC], sp[kLossMaxC], cnt[kLossMaxC];
#pragma unroll
    for (int c = 0; c < kLossMaxC; ++c) tp[c] = sp[c] = cnt[c] = 0.f;
    const long long total = (long long)B * N;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int b = (int)(i / N), n = (int)(i % N);
        const float* x = logits + b * sb + n * sn;
        float v[kLossMaxC], mx = -3.4e38f;
#pragma unroll
        for (int c = 0; c < kLossMaxC; ++c)
            if (c < C) {
                v[c] = x[c * sc];
                mx = fmaxf(mx, v[c]);
            }
        float s = 0.f;
#pragma unroll
        for (int c = 0; c < kLossMaxC; ++c)
            if (c < C) {
                v[c] = expf(v[c] - mx);
                s += v[c];
            }
        const float inv = 1.f / s;
        const int lab = (int)labels[i];
#pragma unroll
        for (int c = 0; c < kLossMaxC; ++c)
            if (c < C) {
                const float p = v[c] * inv;
                sp[c] += p;
                if (lab == c) {
                    tp[c] += p;
                    cnt[c] += 1.f;
                }
            }
    }
#pragma unroll
    for (int c = 0; c < kLossMaxC; ++c) {
        if (c >= C) break;
        float a = tp[c], b2 = sp[c], d = cnt[c];
        for (int o = 16; o > 0; o >>= 1) {
            a += __shfl_xor_sync(0xffffffffu, a, o);
            b2 += __shfl_xor_sync(0xffffffffu, b2, o);
            d += __shfl_xor_sync(0xffffffffu, d, o);
        }
        if ((threadIdx.x & 31) == 0) {
            atomicAdd(&red[0][c], (double)a);
            atomicAdd(&red[1][c], (double)b2);
            atomicAdd(&red[2][c], (double)d);
        }
    }
    __syncthreads();
    if (threadIdx.x < C) {
        atomicAdd(acc + threadIdx.x, red[0][threadIdx.x]);
        atomicAdd(acc + C + threadIdx.x, red[1][threadIdx.x]);
        atomicAdd(acc + 2 * C + threadIdx.x, red[2][threadIdx.x]);
    }
}

__global__ void tversky_finish_kernel(const double* __restrict__ acc, int C, int first, double alpha, double gamma,
                                      double eps, float* __restrict__ loss, float* __restrict__ coef /* (2,C) */) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const int nc = C - first;
    double total = 0.0;
    for (int c = 0; c < C; ++c) {
        coef[c] = 0.f;
        coef[C + c] = 0.f;
        if (c < first) continue;
        const double tp = acc[c], sp = acc[C + c], cnt = acc[2 * C + c];
        const double den = tp + alpha * (cnt - tp) + (1.0 - alpha) * (sp - tp) + eps;
        const double ti = (tp + eps) / den;
        const double one = 1.0 - ti;
        total += gamma == 1.0 ? one : pow(one, gamma);
        const double dterm = gamma == 1.0 ? -1.0 : -gamma * pow(one, gamma - 1.0);        // d term / d ti
        // d ti / d tp = (den - (tp + eps) * d den/d tp) / den^2 with d den/d tp = 1 - alpha - (1 - alpha) = 0
        const double dti_dtp = 1.0 / den;
        const double dti_dsp = -(tp + eps) * (1.0 - alpha) / (den * den);
        coef[c] = (float)(dterm * dti_dtp / nc);
        coef[C + c] = (float)(dterm * dti_dsp / nc);
    }
    *loss = (float)(total / nc);
}

__global__ void __launch_bounds__(256) tversky_bwd_kernel(const float* __restrict__ logits, long long sb, long long sc,
                                                          long long sn, const int64_t* __restrict__ labels, int B, int C,
                                                          int N, const float* __restrict__ coef,
                                                          const float* __restrict__ gout, float* __restrict__ dlogits) {
    __shared__ float cf[2 * kLossMaxC];
    if (threadIdx.x < 2 * C) cf[threadIdx.x] = coef[threadIdx.x];
    __syncthreads();
    const float g = gout ? *gout : 1.f;
    const long long total = (long long)B * N;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int b = (int)(i / N), n = (int)(i % N);
        const long long off = b * sb + n * sn;
        float v[kLossMaxC], mx = -3.4e38f;
#pragma unroll
        for (int c = 0; c < kLossMaxC; ++c)
            if (c < C) {
                v[c] = logits[off + c * sc];
                mx = fmaxf(mx, v[c]);
            }
        float s = 0.f;
#pragma unroll
        for (int c = 0; c < kLossMaxC; ++c)
            if (c < C) {
                v[c] = expf(v[c] - mx);
                s += v[c];
            }
        const float inv = 1.f / s;
        const int lab = (int)labels[i];
        float dot = 0.f, dp[kLossMaxC];
#pragma unroll
        for (int c = 0; c < kLossMaxC; ++c)
            if (c < C) {
                v[c] *= inv;
                dp[c] = (lab == c ? cf[c] : 0.f) + cf[C + c];
                dot = fmaf(v[c], dp[c], dot);
            }
#pragma unroll
        for (int c = 0; c < kLossMaxC; ++c)
            if (c < C) dlogits[off + c * sc] = g * v[c] * (dp[c] - dot);
    }
}

}  // namespace r3d

using namespace r3d;

extern "C" int r3d_tversky_loss_fwd(const float* logits, long long sb, long long sc, long long sn, const int64_t* labels,
                                    int B, int C, int N, int first_class, float alpha, float gamma, float eps, double* acc,
                                    float* loss, float* coef, r3d_stream_t stream) {
    if (B <= 0 || N <= 0 || C <= 0 || first_class < 0 || first_class >= C) return R3D_EINVAL;
    if (C > kLossMaxC) return R3D_EUNSUPPORTED;
    if (!logits || !labels || !acc || !loss || !coef) return R3D_EINVAL;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    long long blocks = ((long long)B * N + 255) / 256;
    if (blocks > (long long)kNumSMs * 8) blocks = (long long)kNumSMs * 8;
    tversky_reduce_kernel<<<(unsigned)blocks, 256, 0, st>>>(logits, sb, sc, sn, labels, B, C, N, acc);
    R3D_LAUNCH_CHECK("tversky_reduce_kernel");
    tversky_finish_kernel<<<1, 32, 0, st>>>(acc, C, first_class, (double)alpha, (double)gamma, (double)eps, loss, coef);
    R3D_LAUNCH_CHECK("tversky_finish_kernel");
    return R3D_OK;
}

extern "C" int r3d_tversky_loss_bwd(const float* logits, long long sb, long long sc, long long sn, const int64_t* labels,
                                    int B, int C, int N, const float* coef, const float* gout, float* dlogits,
                                    r3d_stream_t stream) {
    if (B <= 0 || N <= 0 || C <= 0) return R3D_EINVAL;
    if (C > kLossMaxC) return R3D_EUNSUPPORTED;
    if (!logits || !labels || !coef || !dlogits) return R3D_EINVAL;
    long long blocks = ((long long)B * N + 255) / 256;
    if (blocks > (long long)kNumSMs * 8) blocks = (long long)kNumSMs * 8;
    tversky_bwd_kernel<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(logits, sb, sc, sn, labels, B, C, N,
                                                                                       coef, gout, dlogits);
    R3D_LAUNCH_CHECK("tversky_bwd_kernel");
    return R3D_OK;
}
