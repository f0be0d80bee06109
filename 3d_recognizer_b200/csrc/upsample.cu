// Feature up-sampling from a coarse to a fine point set, fused with the decoder's skip concatenation (sm_100a).
//
// Reference: randlanet/utils/modules.py:343-414 (UpSampler: nearest_neighbor_interpolation = torch.gather at the 1-NN;
// nearest_neighbors_averaging = gather of K=8 neighbours, weights (1+eps)/(dist^p+eps) normalised over K, weighted sum)
// and the decoder loop :596-602 (up-sample with "nni", torch.cat with the encoder skip, SharedMLP).  There the gather,
// the weights, the product, the sum and the concat are separate tensor ops (and in the backward a split, an index_add
// and a copy); here one launch each way:
//
//   forward   out[b,q, 0:F]    = sum_k w[b,q,k] * feat[b, idx[b,q,k], :]
//             out[b,q, F:F+Fs] = skip[b,q,:]                                   (skip nullable)
//   backward  dfeat[b, idx[b,q,k], :] += w[b,q,k] * dout[b,q,0:F]              (vector reductions to global memory)
//             dskip[b,q,:]      = dout[b,q,F:F+Fs]
//
// HBM-bound gather/scatter: a thread owns 4 adjacent channels of one fine point (16-byte accesses; the K weights are
// recomputed per thread from K distances that the whole row's threads read as one broadcast).  `channel_major` writes
// (B,F,N2) instead — the layout Model.upsample returns (model.py:123-144) — with one thread per fine point so that the
// stores stay coalesced for the few class channels.
#include "common.cuh"
#include "lfa_common.cuh"

namespace r3d {

constexpr int kUpMaxK = 16;

enum { kUpFirst = 0, kUpInverseDistance = 1, kUpMean = 2 };

// normalised weights of one fine point (modules.py:396-399: eps 1e-7; ** 1.0 and ** 2 as the multiplications torch does)
template <typename IdxT>
__device__ __forceinline__ void up_weights(const float* __restrict__ dist, long long row, int K, int weighting, float power,
                                           float* w) {
    if (weighting == kUpInverseDistance) {
        const float eps = 1e-7f;
        float sum = 0.f;
        for (int k = 0; k < K; ++k) {
            const float d = dist[row * K + k];
            const float dp = power == 1.f ? d : (power == 2.f ? d * d : powf(d, power));
            w[k] = (1.0f + eps) / (dp + eps);
            sum += w[k];
        }
        for (int k = 0; k < K; ++k) w[k] = w[k] / sum;
    } else if (weighting == kUpMean) {
        for (int k = 0; k < K; ++k) w[k] = 1.0f / (float)K;
    } else {
        for (int k = 0; k < K; ++k) w[k] = k == 0 ? 1.f : 0.f;
    }
}

template <typename IdxT, int V>
__global__ void __launch_bounds__(256)
    upsample_rows_kernel(const float* __restrict__ feat, long long feat_bs, int feat_ld, int F, const IdxT* __restrict__ idx,
                         const float* __restrict__ dist, int K, int weighting, float power, const float* __restrict__ skip,
                         long long skip_bs, int skip_ld, int Fs, float* __restrict__ out, long long out_bs, int out_ld, int B,
                         int N1, int N2) {
    const int cols = (F + Fs) / V;
    const long long total = (long long)B * N2 * cols;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const long long row = t / cols;
        const int c = (int)(t % cols) * V;
        const int b = (int)(row / N2), q = (int)(row % N2);
        float acc[V];
        if (c < F) {
            const float* fb = feat + b * feat_bs + c;
            if (weighting == kUpFirst) {
                const long long j = (long long)idx[row * K];
                if constexpr (V == 4) {
                    const float4 v = *reinterpret_cast<const float4*>(fb + j * feat_ld);
                    acc[0] = v.x, acc[1] = v.y, acc[2] = v.z, acc[3] = v.w;
                } else {
                    acc[0] = fb[j * feat_ld];
                }
            } else {
                float w[kUpMaxK];
                up_weights<IdxT>(dist, row, K, weighting, power, w);
#pragma unroll
                for (int v = 0; v < V; ++v) acc[v] = 0.f;
                for (int k = 0; k < K; ++k) {
                    const long long j = (long long)idx[row * K + k];
                    if constexpr (V == 4) {
                        const float4 v = *reinterpret_cast<const float4*>(fb + j * feat_ld);
                        acc[0] = fmaf(w[k], v.x, acc[0]), acc[1] = fmaf(w[k], v.y, acc[1]);
                        acc[2] = fmaf(w[k], v.z, acc[2]), acc[3] = fmaf(w[k], v.w, acc[3]);
                    } else {
                        acc[0] = fmaf(w[k], fb[j * feat_ld], acc[0]);
                    }
                }
            }
        } else {
            const float* sp = skip + b * skip_bs + (long long)q * skip_ld + (c - F);
            if constexpr (V == 4) {
                const float4 v = *reinterpret_cast<const float4*>(sp);
                acc[0] = v.x, acc[1] = v.y, acc[2] = v.z, acc[3] = v.w;
            } else {
                acc[0] = sp[0];
            }
        }
        float* op = out + b * out_bs + (long long)q * out_ld + c;
        if constexpr (V == 4)
            *reinterpret_cast<float4*>(op) = make_float4(acc[0], acc[1], acc[2], acc[3]);
        else
            op[0] = acc[0];
    }
}

// channel-major output (B,F,N2): one thread per fine point, loop over the (few) channels
template <typename IdxT>
__global__ void __launch_bounds__(256)
    upsample_cm_kernel(const float* __restrict__ feat, long long feat_bs, int feat_ld, int F, const IdxT* __restrict__ idx,
                       const float* __restrict__ dist, int K, int weighting, float power, float* __restrict__ out,
                       long long out_bs, int out_ld, int B, int N1, int N2) {
    const long long total = (long long)B * N2;
    for (long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x; row < total;
         row += (long long)gridDim.x * blockDim.x) {
        const int b = (int)(row / N2), q = (int)(row % N2);
        float w[kUpMaxK];
        long long j[kUpMaxK];
        up_weights<IdxT>(dist, row, K, weighting, power, w);
        const int kk = weighting == kUpFirst ? 1 : K;
        for (int k = 0; k < kk; ++k) j[k] = (long long)idx[row * K + k];
        const float* fb = feat + b * feat_bs;
        float* ob = out + b * out_bs + q;
        for (int c = 0; c < F; ++c) {
            float acc;
            if (weighting == kUpFirst) {
                acc = fb[j[0] * feat_ld + c];
            } else {
                acc = 0.f;
                for (int k = 0; k < kk; ++k) acc = fmaf(w[k], fb[j[k] * feat_ld + c], acc);
            }
            ob[(long long)c * out_ld] = acc;
        }
    }
}

template <typename IdxT, int V>
__global__ void __launch_bounds__(256)
    upsample_bwd_kernel(const float* __restrict__ dout, long long dout_bs, int dout_ld, const IdxT* __restrict__ idx,
                        const float* __restrict__ dist, int K, int weighting, float power, float* __restrict__ dfeat,
                        long long dfeat_bs, int dfeat_ld, int F, float* __restrict__ dskip, long long dskip_bs, int dskip_ld,
                        int Fs, int B, int N1, int N2) {
    const int cols = (F + Fs) / V;
    const long long total = (long long)B * N2 * cols;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const long long row = t / cols;
        const int c = (int)(t % cols) * V;
        const int b = (int)(row / N2), q = (int)(row % N2);
        const float* gp = dout + b * dout_bs + (long long)q * dout_ld + c;
        float g[V];
        if constexpr (V == 4) {
            const float4 v = *reinterpret_cast<const float4*>(gp);
            g[0] = v.x, g[1] = v.y, g[2] = v.z, g[3] = v.w;
        } else {
            g[0] = gp[0];
        }
        if (c >= F) {
            if (dskip == nullptr) continue;
            float* sp = dskip + b * dskip_bs + (long long)q * dskip_ld + (c - F);
            if constexpr (V == 4)
                *reinterpret_cast<float4*>(sp) = make_float4(g[0], g[1], g[2], g[3]);
            else
                sp[0] = g[0];
            continue;
        }
        float w[kUpMaxK];
        up_weights<IdxT>(dist, row, K, weighting, power, w);
        const int kk = weighting == kUpFirst ? 1 : K;
        float* fb = dfeat + b * dfeat_bs + c;
        for (int k = 0; k < kk; ++k) {
            const long long j = (long long)idx[row * K + k];
            if constexpr (V == 4)
                red_add_v4(fb + j * dfeat_ld, make_float4(w[k] * g[0], w[k] * g[1], w[k] * g[2], w[k] * g[3]));
            else
                atomicAdd(fb + j * dfeat_ld, w[k] * g[0]);
        }
    }
}

static inline int up_blocks(long long total) {
    long long blocks = (total + 255) / 256;
    const long long cap = 8LL * kNumSMs;
    return (int)(blocks > cap ? cap : (blocks < 1 ? 1 : blocks));
}

static inline bool up_vec4(const void* p, long long bs, int ld, int n) {
    return p == nullptr || (n % 4 == 0 && ld % 4 == 0 && bs % 4 == 0 && (reinterpret_cast<uintptr_t>(p) & 15) == 0);
}

}  // namespace r3d

using namespace r3d;

static int up_check(const void* idx, const float* dist, int K, int weighting, int B, int N1, int N2, int F, int Fs) {
    if (B < 0 || N1 < 0 || N2 < 0 || F < 0 || Fs < 0 || K <= 0) return R3D_EINVAL;
    if (weighting < kUpFirst || weighting > kUpMean) return R3D_EINVAL;
    if (K > kUpMaxK) return R3D_EKMAX;
    if (B == 0 || N2 == 0 || F + Fs == 0) return 1;
    if (!idx || (weighting == kUpInverseDistance && !dist)) return R3D_EINVAL;
    if (F > 0 && N1 == 0) return R3D_ENOT_ENOUGH;
    return R3D_OK;
}

extern "C" int r3d_upsample(const float* feat, long long feat_bstride, int feat_ld, int F, const void* idx, int idx64,
                            const float* dist, int K, int weighting, float power, const float* skip,
                            long long skip_bstride, int skip_ld, int Fs, float* out, long long out_bstride, int out_ld,
                            int channel_major, int B, int N1, int N2, r3d_stream_t stream) {
    const int rc = up_check(idx, dist, K, weighting, B, N1, N2, F, Fs);
    if (rc != R3D_OK) return rc < 0 ? rc : R3D_OK;
    if (!feat || !out || (Fs > 0 && !skip)) return R3D_EINVAL;
    if (channel_major && Fs > 0) return R3D_EUNSUPPORTED;
    if (feat_ld == 0) feat_ld = F;
    if (feat_bstride == 0) feat_bstride = (long long)N1 * feat_ld;
    if (skip_ld == 0) skip_ld = Fs;
    if (skip_bstride == 0) skip_bstride = (long long)N2 * skip_ld;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (channel_major) {
        if (out_ld == 0) out_ld = N2;
        if (out_bstride == 0) out_bstride = (long long)F * out_ld;
        const int blocks = up_blocks((long long)B * N2);
        if (idx64)
            upsample_cm_kernel<int64_t><<<blocks, 256, 0, s>>>(feat, feat_bstride, feat_ld, F, (const int64_t*)idx, dist, K,
                                                               weighting, power, out, out_bstride, out_ld, B, N1, N2);
        else
            upsample_cm_kernel<int32_t><<<blocks, 256, 0, s>>>(feat, feat_bstride, feat_ld, F, (const int32_t*)idx, dist, K,
                                                               weighting, power, out, out_bstride, out_ld, B, N1, N2);
        R3D_LAUNCH_CHECK("upsample_cm_kernel");
        return R3D_OK;
    }
    if (out_ld == 0) out_ld = F + Fs;
    if (out_bstride == 0) out_bstride = (long long)N2 * out_ld;
    const bool v4 = up_vec4(feat, feat_bstride, feat_ld, F) && up_vec4(skip, skip_bstride, skip_ld, Fs) &&
                    up_vec4(out, out_bstride, out_ld, F + Fs);
    const int blocks = up_blocks((long long)B * N2 * ((F + Fs) / (v4 ? 4 : 1)));
#define R3D_UP_LAUNCH(T, V)                                                                                              \
    upsample_rows_kernel<T, V><<<blocks, 256, 0, s>>>(feat, feat_bstride, feat_ld, F, (const T*)idx, dist, K, weighting, \
                                                      power, skip, skip_bstride, skip_ld, Fs, out, out_bstride, out_ld, \
                                                      B, N1, N2)
    if (idx64) {
        if (v4) R3D_UP_LAUNCH(int64_t, 4); else R3D_UP_LAUNCH(int64_t, 1);
    } else {
        if (v4) R3D_UP_LAUNCH(int32_t, 4); else R3D_UP_LAUNCH(int32_t, 1);
    }
#undef R3D_UP_LAUNCH
    R3D_LAUNCH_CHECK("upsample_rows_kernel");
    return R3D_OK;
}

extern "C" int r3d_upsample_bwd(const float* dout, long long dout_bstride, int dout_ld, const void* idx, int idx64,
                                const float* dist, int K, int weighting, float power, float* dfeat,
                                long long dfeat_bstride, int dfeat_ld, int F, float* dskip, long long dskip_bstride,
                                int dskip_ld, int Fs, int B, int N1, int N2, r3d_stream_t stream) {
    const int rc = up_check(idx, dist, K, weighting, B, N1, N2, F, Fs);
    if (rc != R3D_OK) return rc < 0 ? rc : R3D_OK;
    if (!dout || (F > 0 && !dfeat)) return R3D_EINVAL;
    if (dout_ld == 0) dout_ld = F + Fs;
    if (dout_bstride == 0) dout_bstride = (long long)N2 * dout_ld;
    if (dfeat_ld == 0) dfeat_ld = F;
    if (dfeat_bstride == 0) dfeat_bstride = (long long)N1 * dfeat_ld;
    if (dskip_ld == 0) dskip_ld = Fs;
    if (dskip_bstride == 0) dskip_bstride = (long long)N2 * dskip_ld;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const bool v4 = up_vec4(dout, dout_bstride, dout_ld, F + Fs) && up_vec4(dfeat, dfeat_bstride, dfeat_ld, F) &&
                    up_vec4(dskip, dskip_bstride, dskip_ld, Fs) && F % 4 == 0;
    const int blocks = up_blocks((long long)B * N2 * ((F + Fs) / (v4 ? 4 : 1)));
#define R3D_UP_LAUNCH(T, V)                                                                                           \
    upsample_bwd_kernel<T, V><<<blocks, 256, 0, s>>>(dout, dout_bstride, dout_ld, (const T*)idx, dist, K, weighting,  \
                                                     power, dfeat, dfeat_bstride, dfeat_ld, F, dskip, dskip_bstride, \
                                                     dskip_ld, Fs, B, N1, N2)
    if (idx64) {
        if (v4) R3D_UP_LAUNCH(int64_t, 4); else R3D_UP_LAUNCH(int64_t, 1);
    } else {
        if (v4) R3D_UP_LAUNCH(int32_t, 4); else R3D_UP_LAUNCH(int32_t, 1);
    }
#undef R3D_UP_LAUNCH
    R3D_LAUNCH_CHECK("upsample_bwd_kernel");
    return R3D_OK;
}
