// Row-form LocSE + attentive pooling for shapes the fused kernels are not instantiated for (sm_100a).
//
// The fused kernels (lfa.cu, lfa_cl*.cu) are templates over the block width d in {16,...,256} and K in {16,32}.  The
// reference accepts any n_neighbors and any even layer size (randlanet/utils/modules.py:298-325, 484-500); for those
// settings the block runs in ROW FORM: the (B*N*K, C) neighbourhood rows are materialised and the per-point layer
// kernels (pointwise.cu, pointwise_train.cu) do mlp_rpe1/2 and the score Linear over them.  This file holds the three
// pieces that are not per-point layers, for any K >= 1 and any width:
//
//   r3d_lfa_rpe_rows        rows[b,n,k,0:10] = [p_i, p_j, p_i - p_j, |p_i - p_j|]          (modules.py:170-186)
//   r3d_lfa_gather_concat   X[b,n,k,:] = [r[b,n,k,:] ; feat[b, idx[b,n,k], :]]              (modules.py:200-208)
//   r3d_lfa_attn_pool       pooled[b,n,c] = sum_k softmax_k(S[b,n,k,c]) X[b,n,k,c]          (modules.py:246-252)
// and the backward of the last two.  Slower than the fused path by the memory traffic of the materialised rows
// (8 d K bytes per point and tensor); results agree with it to fp32 round-off (tests/test_lfa_rows_gpu.py).
#include "common.cuh"

namespace r3d {

__global__ void __launch_bounds__(256) rpe_rows_kernel(const float* __restrict__ xyz, long long xs,
                                                       const int32_t* __restrict__ idx, float* __restrict__ out,
                                                       int N, int K, long long rows) {
    const long long per_cloud = (long long)N * K;
    for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += (long long)gridDim.x * blockDim.x) {
        const long long b = r / per_cloud;
        const int n = (int)((r % per_cloud) / K);
        const float* c = xyz + b * xs;
        const int j = idx[r];
        const float ix = c[(size_t)n * 3], iy = c[(size_t)n * 3 + 1], iz = c[(size_t)n * 3 + 2];
        const float jx = c[(size_t)j * 3], jy = c[(size_t)j * 3 + 1], jz = c[(size_t)j * 3 + 2];
        // the KNN contract's rounding sequence: |p_i - p_j| is the square root of the search's d2 bit for bit
        const float dx = __fsub_rn(ix, jx), dy = __fsub_rn(iy, jy), dz = __fsub_rn(iz, jz);
        const float dist = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz)));
        float* o = out + r * 10;
        o[0] = ix; o[1] = iy; o[2] = iz; o[3] = jx; o[4] = jy; o[5] = jz; o[6] = dx; o[7] = dy; o[8] = dz; o[9] = dist;
    }
}

// one thread per (row, channel of the 2h-wide output); channels of a row are consecutive threads (coalesced both ways)
__global__ void __launch_bounds__(256) gather_concat_kernel(const float* __restrict__ r, const float* __restrict__ feat,
                                                            long long fs, const int32_t* __restrict__ idx,
                                                            float* __restrict__ out, int N, int K, int h, long long rows) {
    const long long per_cloud = (long long)N * K;
    const long long total = rows * 2 * h;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long row = i / (2 * h);
        const int c = (int)(i % (2 * h));
        float v;
        if (c < h) {
            v = r[row * h + c];
        } else {
            const long long b = row / per_cloud;
            v = feat[b * fs + (long long)idx[row] * h + (c - h)];
        }
        out[i] = v;
    }
}

__global__ void __launch_bounds__(256) gather_concat_bwd_kernel(const float* __restrict__ dout,
                                                                const int32_t* __restrict__ idx, float* __restrict__ dr,
                                                                float* __restrict__ dfeat, long long dfs, int N, int K,
                                                                int h, long long rows) {
    const long long per_cloud = (long long)N * K;
    const long long total = rows * 2 * h;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long row = i / (2 * h);
        const int c = (int)(i % (2 * h));
        const float v = dout[i];
        if (c < h) {
            if (dr) dr[row * h + c] = v;
        } else if (dfeat) {
            const long long b = row / per_cloud;
            atomicAdd(dfeat + b * dfs + (long long)idx[row] * h + (c - h), v);
        }
    }
}

// one thread per (point, channel); the K rows of a point are d floats apart
__global__ void __launch_bounds__(256) attn_pool_kernel(const float* __restrict__ S, const float* __restrict__ X,
                                                        float* __restrict__ pooled, int K, int d, long long points) {
    const long long total = points * d;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long p = i / d;
        const int c = (int)(i % d);
        const float* s = S + p * K * d + c;
        const float* x = X + p * K * d + c;
        float m = s[0];
        for (int k = 1; k < K; ++k) m = fmaxf(m, s[(size_t)k * d]);
        float den = 0.f, num = 0.f;
        for (int k = 0; k < K; ++k) {
            const float e = expf(s[(size_t)k * d] - m);
            den += e;
            num = fmaf(e, x[(size_t)k * d], num);
        }
        pooled[i] = num / den;
    }
}

// dX = g A (the direct term; the score Linear adds dS Ws through its own backward), dS = A g (X - pooled)
__global__ void __launch_bounds__(256) attn_pool_bwd_kernel(const float* __restrict__ S, const float* __restrict__ X,
                                                            const float* __restrict__ dpooled, float* __restrict__ dS,
                                                            float* __restrict__ dX, int K, int d, long long points) {
    const long long total = points * d;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long p = i / d;
        const int c = (int)(i % d);
        const size_t base = (size_t)p * K * d + c;
        float m = S[base];
        for (int k = 1; k < K; ++k) m = fmaxf(m, S[base + (size_t)k * d]);
        float den = 0.f, num = 0.f;
        for (int k = 0; k < K; ++k) {
            const float e = expf(S[base + (size_t)k * d] - m);
            den += e;
            num = fmaf(e, X[base + (size_t)k * d], num);
        }
        const float inv = 1.0f / den, pooled = num * inv, g = dpooled[i];
        for (int k = 0; k < K; ++k) {
            const size_t o = base + (size_t)k * d;
            const float ga = g * (expf(S[o] - m) * inv);
            dX[o] = ga;
            dS[o] = ga * (X[o] - pooled);
        }
    }
}

static int grid_for(long long work) {
    long long blocks = (work + 255) / 256;
    const long long cap = 16LL * kNumSMs;
    return (int)(blocks < 1 ? 1 : (blocks > cap ? cap : blocks));
}

}  // namespace r3d

using namespace r3d;

extern "C" int r3d_lfa_rpe_rows(const float* xyz, long long xyz_bstride, const int32_t* idx, float* out, int B, int N,
                                int K, r3d_stream_t stream) {
    if (B < 0 || N < 0 || K <= 0) return R3D_EINVAL;
    if (B == 0 || N == 0) return R3D_OK;
    if (!xyz || !idx || !out) return R3D_EINVAL;
    if (xyz_bstride == 0) xyz_bstride = (long long)N * 3;
    const long long rows = (long long)B * N * K;
    rpe_rows_kernel<<<grid_for(rows), 256, 0, static_cast<cudaStream_t>(stream)>>>(xyz, xyz_bstride, idx, out, N, K, rows);
    R3D_LAUNCH_CHECK("rpe_rows_kernel");
    return R3D_OK;
}

extern "C" int r3d_lfa_gather_concat(const float* r, const float* feat, long long feat_bstride, const int32_t* idx,
                                     float* out, int B, int N, int K, int h, r3d_stream_t stream) {
    if (B < 0 || N < 0 || K <= 0 || h <= 0) return R3D_EINVAL;
    if (B == 0 || N == 0) return R3D_OK;
    if (!r || !feat || !idx || !out) return R3D_EINVAL;
    if (feat_bstride == 0) feat_bstride = (long long)N * h;
    const long long rows = (long long)B * N * K;
    gather_concat_kernel<<<grid_for(rows * 2 * h), 256, 0, static_cast<cudaStream_t>(stream)>>>(r, feat, feat_bstride, idx,
                                                                                                out, N, K, h, rows);
    R3D_LAUNCH_CHECK("gather_concat_kernel");
    return R3D_OK;
}

extern "C" int r3d_lfa_gather_concat_bwd(const float* dout, const int32_t* idx, float* dr, float* dfeat,
                                         long long dfeat_bstride, int B, int N, int K, int h, r3d_stream_t stream) {
    if (B < 0 || N < 0 || K <= 0 || h <= 0) return R3D_EINVAL;
    if (B == 0 || N == 0) return R3D_OK;
    if (!dout || !idx) return R3D_EINVAL;
    if (dfeat_bstride == 0) dfeat_bstride = (long long)N * h;
    const long long rows = (long long)B * N * K;
    gather_concat_bwd_kernel<<<grid_for(rows * 2 * h), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        dout, idx, dr, dfeat, dfeat_bstride, N, K, h, rows);
    R3D_LAUNCH_CHECK("gather_concat_bwd_kernel");
    return R3D_OK;
}

extern "C" int r3d_lfa_attn_pool(const float* S, const float* X, float* pooled, long long points, int K, int d,
                                 r3d_stream_t stream) {
    if (points < 0 || K <= 0 || d <= 0) return R3D_EINVAL;
    if (points == 0) return R3D_OK;
    if (!S || !X || !pooled) return R3D_EINVAL;
    attn_pool_kernel<<<grid_for(points * d), 256, 0, static_cast<cudaStream_t>(stream)>>>(S, X, pooled, K, d, points);
    R3D_LAUNCH_CHECK("attn_pool_kernel");
    return R3D_OK;
}

extern "C" int r3d_lfa_attn_pool_bwd(const float* S, const float* X, const float* dpooled, float* dS, float* dX,
                                     long long points, int K, int d, r3d_stream_t stream) {
    if (points < 0 || K <= 0 || d <= 0) return R3D_EINVAL;
    if (points == 0) return R3D_OK;
    if (!S || !X || !dpooled || !dS || !dX) return R3D_EINVAL;
    attn_pool_bwd_kernel<<<grid_for(points * d), 256, 0, static_cast<cudaStream_t>(stream)>>>(S, X, dpooled, dS, dX, K, d,
                                                                                              points);
    R3D_LAUNCH_CHECK("attn_pool_bwd_kernel");
    return R3D_OK;
}
