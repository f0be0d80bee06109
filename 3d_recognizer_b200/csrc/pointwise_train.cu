// Train-mode companions of the per-point layer kernel (pointwise.cu), sm_100a.
//
// A SharedMLP in training mode (randlanet/utils/modules.py:60-104: 1x1 conv -> BatchNorm2d(eps 1e-6, momentum
// 0.99) with BATCH statistics -> activation) over M = B*N rows runs as
//   forward : r3d_pointwise_stats  z = W x, per-channel sum / sum of squares (fp64) in the GEMM epilogue
//             r3d_bn_apply         y = act(a z + c), a = gamma * rstd, c = beta - a * mean; running statistics
//   backward: r3d_bn_bwd_reduce    s1 = sum du, s2 = sum du * zhat   (du = dy * act'(a z + c))
//             r3d_bn_bwd_dz        dz = a (du - s1/M - zhat s2/M)    (BatchNorm backward, batch statistics)
//             r3d_pointwise        dx = dz W                         (the forward GEMM kernel on the transposed weight)
//             r3d_rowreduce_gemm   dW = dz^T x                       (row-reduction GEMM, split over row chunks)
// All tensors are dense row-major (M, C); C is a multiple of 4 for the BatchNorm kernels.
// Backward reference: autograd of modules.py:92-104 as driven by trainer.py:115-119.
#include <cstdlib>
#include "common.cuh"

#include <cooperative_groups.h>
#include <atomic>

namespace r3d {

int bn_fused_mask();     // r3d_bn_set_fused, also read by pointwise.cu

constexpr int kBnMaxC = 1024;

__device__ __forceinline__ float act_fwd(float v, int act, float slope) {
    if (act == 1) return fmaxf(v, 0.f);
    if (act == 2) return v > 0.f ? v : v * slope;
    return v;
}
__device__ __forceinline__ float act_grad(float u, int act, float slope) {
    if (act == 1) return u > 0.f ? 1.f : 0.f;
    if (act == 2) return u > 0.f ? 1.f : slope;
    return 1.f;
}

// ---------------------------------------------------------------------------------------- bn_apply
// save: (3, C) = a (gamma * rstd), mean (of z, without the conv bias), rstd
__global__ void __launch_bounds__(256) bn_apply_kernel(const float* __restrict__ z, const double* __restrict__ stats,
                                                       long long M, int C, const float* __restrict__ gamma,
                                                       const float* __restrict__ beta, const float* __restrict__ bias,
                                                       float eps, float momentum, float* __restrict__ running_mean,
                                                       float* __restrict__ running_var,
                                                       long long* __restrict__ num_batches, int act, float slope,
                                                       float* __restrict__ y, float* __restrict__ save) {
    __shared__ float sa[kBnMaxC], sc[kBnMaxC];
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        const double mean = stats[c] / (double)M;
        double var = stats[C + c] / (double)M - mean * mean;
        var = var > 0.0 ? var : 0.0;
        const float rstd = (float)(1.0 / sqrt(var + (double)eps));
        const float a = gamma[c] * rstd;
        sa[c] = a;
        sc[c] = beta[c] - a * (float)mean;
        if (blockIdx.x == 0) {
            save[c] = a;
            save[C + c] = (float)mean;
            save[2 * C + c] = rstd;
            if (running_mean) {
                const double unbiased = var * ((double)M / (double)(M > 1 ? M - 1 : 1));
                running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * ((float)mean + (bias ? bias[c] : 0.f));
                running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
            }
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0 && num_batches) *num_batches += 1;
    __syncthreads();
    const long long total4 = M * C / 4;
    const int c4 = C / 4;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % c4) * 4;
        float4 v = reinterpret_cast<const float4*>(z)[i];
        v.x = act_fwd(fmaf(v.x, sa[c + 0], sc[c + 0]), act, slope);
        v.y = act_fwd(fmaf(v.y, sa[c + 1], sc[c + 1]), act, slope);
        v.z = act_fwd(fmaf(v.z, sa[c + 2], sc[c + 2]), act, slope);
        v.w = act_fwd(fmaf(v.w, sa[c + 3], sc[c + 3]), act, slope);
        reinterpret_cast<float4*>(y)[i] = v;
    }
}

// ----------------------------------------------------------------------------------- bn_bwd_reduce
// stats2 (2C fp64): [c] += sum du, [C + c] += sum du * zhat
__device__ __forceinline__ void bn_bwd_reduce_body(const float* __restrict__ dy, const float* __restrict__ z,
                                                   long long M, int C, const float* __restrict__ save,
                                                   const float* __restrict__ beta, int act, float slope,
                                                   double* __restrict__ stats2, float (*red)[kBnMaxC]) {
    const int c4n = C / 4;                       // column quads
    const int rows_per_pass = blockDim.x / c4n;  // >= 1 (C <= 1024)
    const int q = threadIdx.x % c4n, rl = threadIdx.x / c4n;
    float a[4], mean[4], rstd[4], cc[4];
    const bool active = rl < rows_per_pass;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        a[j] = save[q * 4 + j];
        mean[j] = save[C + q * 4 + j];
        rstd[j] = save[2 * C + q * 4 + j];
        cc[j] = beta[q * 4 + j] - a[j] * mean[j];
    }
    float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
    if (active) {
        for (long long r = (long long)blockIdx.x * rows_per_pass + rl; r < M; r += (long long)gridDim.x * rows_per_pass) {
            const float4 g = reinterpret_cast<const float4*>(dy + r * C)[q];
            const float4 v = reinterpret_cast<const float4*>(z + r * C)[q];
            const float gv[4] = {g.x, g.y, g.z, g.w}, zv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float du = gv[j] * act_grad(fmaf(zv[j], a[j], cc[j]), act, slope);
                s1[j] += du;
                s2[j] = fmaf(du, (zv[j] - mean[j]) * rstd[j], s2[j]);
            }
        }
    }
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) (&red[0][0])[i < C ? i : kBnMaxC + (i - C)] = 0.f;
    __syncthreads();
    // narrow layers: the lanes of a warp that hold the same column quad (lane % c4n) combine by shuffles first --
    // with C = 8 all 256 threads would otherwise queue on 16 shared-memory addresses
    bool writer = active;
    if (c4n <= 16 && (c4n & (c4n - 1)) == 0) {
        for (int o = 16; o >= c4n; o >>= 1) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                s1[j] += __shfl_xor_sync(0xffffffffu, s1[j], o);
                s2[j] += __shfl_xor_sync(0xffffffffu, s2[j], o);
            }
        }
        writer = (threadIdx.x & 31) < c4n;
    }
    if (writer) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            atomicAdd(&red[0][q * 4 + j], s1[j]);
            atomicAdd(&red[1][q * 4 + j], s2[j]);
        }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        atomicAdd(stats2 + c, (double)red[0][c]);
        atomicAdd(stats2 + C + c, (double)red[1][c]);
    }
}

__global__ void __launch_bounds__(256) bn_bwd_reduce_kernel(const float* __restrict__ dy, const float* __restrict__ z,
                                                            long long M, int C, const float* __restrict__ save,
                                                            const float* __restrict__ beta, int act, float slope,
                                                            double* __restrict__ stats2) {
    __shared__ float red[2][kBnMaxC];
    bn_bwd_reduce_body(dy, z, M, C, save, beta, act, slope, stats2, red);
}

// --------------------------------------------------------------------------------------- bn_bwd_dz
struct BnDzSmem {
    float sa[kBnMaxC], sm[kBnMaxC], sr[kBnMaxC], sc[kBnMaxC], m1[kBnMaxC], m2[kBnMaxC];
};

__device__ __forceinline__ void bn_bwd_dz_body(const float* __restrict__ dy, const float* __restrict__ z, long long M,
                                               int C, const float* __restrict__ save, const float* __restrict__ beta,
                                               int act, float slope, const double* stats2, float* __restrict__ dz,
                                               float* __restrict__ dgb, BnDzSmem& S, float* absmax = nullptr) {
    float *sa = S.sa, *sm = S.sm, *sr = S.sr, *sc = S.sc, *m1 = S.m1, *m2 = S.m2;
    if (dgb && blockIdx.x == 0)
        for (int c = threadIdx.x; c < 2 * C; c += blockDim.x) dgb[c] = (float)stats2[c];
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        sa[c] = save[c];
        sm[c] = save[C + c];
        sr[c] = save[2 * C + c];
        sc[c] = beta[c] - sa[c] * sm[c];
        m1[c] = (float)(stats2[c] / (double)M);
        m2[c] = (float)(stats2[C + c] / (double)M);
    }
    __syncthreads();
    const long long total4 = M * C / 4;
    const int c4 = C / 4;
    float mx = 0.f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % c4) * 4;
        const float4 g = reinterpret_cast<const float4*>(dy)[i];
        const float4 v = reinterpret_cast<const float4*>(z)[i];
        const float gv[4] = {g.x, g.y, g.z, g.w}, zv[4] = {v.x, v.y, v.z, v.w};
        float o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float du = gv[j] * act_grad(fmaf(zv[j], sa[c + j], sc[c + j]), act, slope);
            const float zh = (zv[j] - sm[c + j]) * sr[c + j];
            o[j] = sa[c + j] * (du - m1[c + j] - zh * m2[c + j]);
            mx = fmaxf(mx, fabsf(o[j]));
        }
        reinterpret_cast<float4*>(dz)[i] = make_float4(o[0], o[1], o[2], o[3]);
    }
    if (absmax) {        // max |dz|: the operand scale of the tensor-core weight-gradient kernel (r3d_pc_wgrad)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        if ((threadIdx.x & 31) == 0 && mx > 0.f) atomicMax(reinterpret_cast<int*>(absmax), __float_as_int(mx));
    }
}

__global__ void __launch_bounds__(256) bn_bwd_dz_kernel(const float* __restrict__ dy, const float* __restrict__ z,
                                                        long long M, int C, const float* __restrict__ save,
                                                        const float* __restrict__ beta, int act, float slope,
                                                        const double* __restrict__ stats2, float* __restrict__ dz,
                                                        float* __restrict__ dgb, float* __restrict__ absmax) {
    __shared__ BnDzSmem S;
    bn_bwd_dz_body(dy, z, M, C, save, beta, act, slope, stats2, dz, dgb, S, absmax);
}

// Both passes in ONE cooperative launch with a grid barrier between them: for the layers of a small cloud each pass
// is a ~5 us kernel whose cost is its launch, and the second pass re-reads dy and z from L2 anyway.  The grid is
// sized to be co-resident (r3d_bn_bwd); larger tensors take the two-launch path.
__global__ void __launch_bounds__(256) bn_bwd_fused_kernel(const float* __restrict__ dy, const float* __restrict__ z,
                                                           long long M, int C, const float* __restrict__ save,
                                                           const float* __restrict__ beta, int act, float slope,
                                                           double* stats2, float* __restrict__ dz,
                                                           float* __restrict__ dgb, float* __restrict__ absmax) {
    __shared__ union {
        float red[2][kBnMaxC];
        BnDzSmem dzs;
    } S;
    bn_bwd_reduce_body(dy, z, M, C, save, beta, act, slope, stats2, S.red);
    __threadfence();
    cooperative_groups::this_grid().sync();
    bn_bwd_dz_body(dy, z, M, C, save, beta, act, slope, stats2, dz, dgb, S.dzs, absmax);
}

// ------------------------------------------------------------------------------------ rowreduce_gemm
// out[ca][cb] += sum_m A[m][ca] * B[m][cb]   (A: M x Ca, B: M x Cb, dense row-major; out: Ca x Cb, ld_out)
// CTA tile 64 x 64 outputs, 256 threads (4 x 4 per thread), rows chunked over gridDim.z; fp32 atomics.
constexpr int kRrTile = 64;
constexpr int kRrKC = 16;

__global__ void __launch_bounds__(256) rowreduce_gemm_kernel(const float* __restrict__ A, int Ca,
                                                             const float* __restrict__ Bm, int Cb, long long M,
                                                             long long rows_per_cta, float* __restrict__ out,
                                                             int ld_out) {
    __shared__ __align__(16) float As[kRrKC][kRrTile];
    __shared__ __align__(16) float Bs[kRrKC][kRrTile];
    const int tid = threadIdx.x;
    const int a0 = blockIdx.x * kRrTile, b0 = blockIdx.y * kRrTile;
    const long long r_lo = (long long)blockIdx.z * rows_per_cta;
    const long long r_hi = min(M, r_lo + rows_per_cta);
    const int ta = tid / 16, tb = tid % 16;   // thread tile: rows ta*4.., cols tb*4..
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (long long r0 = r_lo; r0 < r_hi; r0 += kRrKC) {
        // stage 16 rows x 64 columns of both operands (zero-filled outside)
        for (int i = tid; i < kRrKC * kRrTile; i += 256) {
            const int k = i / kRrTile, c = i % kRrTile;
            const long long r = r0 + k;
            As[k][c] = (r < r_hi && a0 + c < Ca) ? A[r * Ca + a0 + c] : 0.f;
            Bs[k][c] = (r < r_hi && b0 + c < Cb) ? Bm[r * Cb + b0 + c] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < kRrKC; ++k) {
            const float4 av = *reinterpret_cast<const float4*>(&As[k][ta * 4]);
            const float4 bv = *reinterpret_cast<const float4*>(&Bs[k][tb * 4]);
            const float a_[4] = {av.x, av.y, av.z, av.w}, b_[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a_[i], b_[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int ca = a0 + ta * 4 + i, cb = b0 + tb * 4 + j;
            if (ca < Ca && cb < Cb) atomicAdd(out + (size_t)ca * ld_out + cb, acc[i][j]);
        }
}

// Pipelined variant for channel counts that are multiples of 4: 32-row chunks, float4 loads, the loads of chunk
// c+1 in flight during the FMAs of chunk c, one barrier per chunk (the scalar version above spent 42 us per
// launch on load latency in the small-cloud training step).
constexpr int kRrKC2 = 32;

__global__ void __launch_bounds__(256) rowreduce_gemm_fast_kernel(const float* __restrict__ A, int Ca,
                                                                  const float* __restrict__ Bm, int Cb, long long M,
                                                                  long long rows_per_cta, float* __restrict__ out,
                                                                  int ld_out) {
    __shared__ __align__(16) float As[2][kRrKC2][kRrTile];
    __shared__ __align__(16) float Bs[2][kRrKC2][kRrTile];
    const int tid = threadIdx.x;
    const int a0 = blockIdx.x * kRrTile, b0 = blockIdx.y * kRrTile;
    const long long r_lo = (long long)blockIdx.z * rows_per_cta;
    const long long r_hi = min(M, r_lo + rows_per_cta);
    const int ta = tid / 16, tb = tid % 16;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    float4 ra[2], rb[2];
    auto load = [&](long long r0) {
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const int i = tid + k * 256;
            const int row = i / 16, c = (i % 16) * 4;
            const long long r = r0 + row;
            ra[k] = (r < r_hi && a0 + c < Ca) ? *reinterpret_cast<const float4*>(A + r * Ca + a0 + c)
                                              : make_float4(0.f, 0.f, 0.f, 0.f);
            rb[k] = (r < r_hi && b0 + c < Cb) ? *reinterpret_cast<const float4*>(Bm + r * Cb + b0 + c)
                                              : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    };
    auto store = [&](int buf) {
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const int i = tid + k * 256;
            *reinterpret_cast<float4*>(&As[buf][i / 16][(i % 16) * 4]) = ra[k];
            *reinterpret_cast<float4*>(&Bs[buf][i / 16][(i % 16) * 4]) = rb[k];
        }
    };
    load(r_lo);
    store(0);
    __syncthreads();
    int buf = 0;
    for (long long r0 = r_lo; r0 < r_hi; r0 += kRrKC2, buf ^= 1) {
        const bool more = r0 + kRrKC2 < r_hi;
        if (more) load(r0 + kRrKC2);
#pragma unroll
        for (int k = 0; k < kRrKC2; ++k) {
            const float4 av = *reinterpret_cast<const float4*>(&As[buf][k][ta * 4]);
            const float4 bv = *reinterpret_cast<const float4*>(&Bs[buf][k][tb * 4]);
            const float a_[4] = {av.x, av.y, av.z, av.w}, b_[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a_[i], b_[j], acc[i][j]);
        }
        if (more) store(buf ^ 1);
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int ca = a0 + ta * 4 + i, cb = b0 + tb * 4 + j;
            if (ca < Ca && cb < Cb) atomicAdd(out + (size_t)ca * ld_out + cb, acc[i][j]);
        }
}

// Large-tile variant for wide layers (both channel counts >= 128): 128 x 128 outputs per CTA, 8 x 8 per thread, 16-row
// chunks double-buffered with register prefetch — four times the FMA : LDS ratio of the 64 x 64 kernel.
constexpr int kRrBig = 128;
constexpr int kRrBigKC = 16;

__global__ void __launch_bounds__(256) rowreduce_gemm_big_kernel(const float* __restrict__ A, int Ca,
                                                                 const float* __restrict__ Bm, int Cb, long long M,
                                                                 long long rows_per_cta, float* __restrict__ out,
                                                                 int ld_out) {
    __shared__ __align__(16) float As[2][kRrBigKC][kRrBig];
    __shared__ __align__(16) float Bs[2][kRrBigKC][kRrBig];
    const int tid = threadIdx.x;
    const int a0 = blockIdx.x * kRrBig, b0 = blockIdx.y * kRrBig;
    const long long r_lo = (long long)blockIdx.z * rows_per_cta;
    const long long r_hi = min(M, r_lo + rows_per_cta);
    const int ta = tid / 16, tb = tid % 16;   // rows ta*4.. and 64 + ta*4.., cols tb*4.. and 64 + tb*4..
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
    float4 ra[2], rb[2];
    auto load = [&](long long r0) {
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const int i = tid + k * 256;                 // 16 rows x 32 quads
            const int row = i / 32, c = (i % 32) * 4;
            const long long r = r0 + row;
            ra[k] = (r < r_hi && a0 + c < Ca) ? *reinterpret_cast<const float4*>(A + r * Ca + a0 + c)
                                              : make_float4(0.f, 0.f, 0.f, 0.f);
            rb[k] = (r < r_hi && b0 + c < Cb) ? *reinterpret_cast<const float4*>(Bm + r * Cb + b0 + c)
                                              : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    };
    auto store = [&](int buf) {
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const int i = tid + k * 256;
            *reinterpret_cast<float4*>(&As[buf][i / 32][(i % 32) * 4]) = ra[k];
            *reinterpret_cast<float4*>(&Bs[buf][i / 32][(i % 32) * 4]) = rb[k];
        }
    };
    load(r_lo);
    store(0);
    __syncthreads();
    int buf = 0;
    for (long long r0 = r_lo; r0 < r_hi; r0 += kRrBigKC, buf ^= 1) {
        const bool more = r0 + kRrBigKC < r_hi;
        if (more) load(r0 + kRrBigKC);
#pragma unroll
        for (int k = 0; k < kRrBigKC; ++k) {
            const float4 a0v = *reinterpret_cast<const float4*>(&As[buf][k][ta * 4]);
            const float4 a1v = *reinterpret_cast<const float4*>(&As[buf][k][64 + ta * 4]);
            const float4 b0v = *reinterpret_cast<const float4*>(&Bs[buf][k][tb * 4]);
            const float4 b1v = *reinterpret_cast<const float4*>(&Bs[buf][k][64 + tb * 4]);
            const float a_[8] = {a0v.x, a0v.y, a0v.z, a0v.w, a1v.x, a1v.y, a1v.z, a1v.w};
            const float b_[8] = {b0v.x, b0v.y, b0v.z, b0v.w, b1v.x, b1v.y, b1v.z, b1v.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a_[i], b_[j], acc[i][j]);
        }
        if (more) store(buf ^ 1);
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int ca = a0 + (i < 4 ? ta * 4 + i : 64 + ta * 4 + (i - 4));
            const int cb = b0 + (j < 4 ? tb * 4 + j : 64 + tb * 4 + (j - 4));
            if (ca < Ca && cb < Cb) atomicAdd(out + (size_t)ca * ld_out + cb, acc[i][j]);
        }
}

// Narrow variant (both channel counts <= 64 and multiples of 4, at least one of them <= 32): the 64 x 64 tile of the
// kernel above would spend 4-16x its FMAs and half its load slots on padding.  Here the output tile is TA x TB, a
// thread still owns a 4 x 4 block, and the 256 threads form NS = 256 / ((TA/4)(TB/4)) row SLICES: slice s takes rows
// s, s + NS, ... of every 64-row chunk.  Chunks are double-buffered in shared memory by float4 loads that carry no
// padding; the slices are summed through shared memory at the end (one atomic per element and CTA).  HBM-streaming:
// 2.6 M rows of a 16 x 16 layer in ~0.1 ms instead of 0.57 ms.
constexpr int kRrNarrowRows = 64;

template <int TA, int TB>
__global__ void __launch_bounds__(256) rowreduce_gemm_narrow_kernel(const float* __restrict__ A, int Ca,
                                                                    const float* __restrict__ Bm, int Cb, long long M,
                                                                    long long rows_per_cta, float* __restrict__ out,
                                                                    int ld_out) {
    constexpr int NTT = (TA / 4) * (TB / 4);          // threads per row slice
    constexpr int NS = 256 / NTT;                     // row slices
    constexpr int FA = TA / 4, FB = TB / 4;           // float4 per row
    constexpr int ROWS = (FA + FB > 16) ? kRrNarrowRows / 2 : kRrNarrowRows;   // rows per chunk (48 KB static limit)
    constexpr int F4 = ROWS * (FA + FB);              // float4 per chunk
    constexpr int LPT = (F4 + 255) / 256;             // loads per thread and chunk
    __shared__ __align__(16) float4 buf[2][F4];       // [row][FA | FB]
    __shared__ float red[TA * TB];
    const int tid = threadIdx.x;
    const int slice = tid / NTT, tt = tid % NTT;
    const int ta = tt / FB, tb = tt % FB;
    const long long r_lo = (long long)blockIdx.x * rows_per_cta;
    const long long r_hi = min(M, r_lo + rows_per_cta);
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    float4 rg[LPT];
    auto load = [&](long long r0) {
#pragma unroll
        for (int k = 0; k < LPT; ++k) {
            const int i = tid + k * 256;
            const int row = i / (FA + FB), c = i % (FA + FB);
            const long long r = r0 + row;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (i < F4 && r < r_hi) {
                if (c < FA) {
                    if (c * 4 < Ca) v = *reinterpret_cast<const float4*>(A + r * Ca + c * 4);
                } else if ((c - FA) * 4 < Cb) {
                    v = *reinterpret_cast<const float4*>(Bm + r * Cb + (c - FA) * 4);
                }
            }
            rg[k] = v;
        }
    };
    auto store = [&](int b) {
#pragma unroll
        for (int k = 0; k < LPT; ++k) {
            const int i = tid + k * 256;
            if (i < F4) buf[b][i] = rg[k];
        }
    };
    for (int i = tid; i < TA * TB; i += 256) red[i] = 0.f;
    load(r_lo);
    store(0);
    __syncthreads();
    int b = 0;
    for (long long r0 = r_lo; r0 < r_hi; r0 += ROWS, b ^= 1) {
        const bool more = r0 + ROWS < r_hi;
        if (more) load(r0 + ROWS);
#pragma unroll
        for (int k = slice; k < ROWS; k += NS) {
            const float4 av = buf[b][k * (FA + FB) + ta];
            const float4 bv = buf[b][k * (FA + FB) + FA + tb];
            const float a_[4] = {av.x, av.y, av.z, av.w}, b_[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a_[i], b_[j], acc[i][j]);
        }
        if (more) store(b ^ 1);
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) atomicAdd(&red[(ta * 4 + i) * TB + tb * 4 + j], acc[i][j]);
    __syncthreads();
    for (int i = tid; i < TA * TB; i += 256) {
        const int ca = i / TB, cb = i % TB;
        if (ca < Ca && cb < Cb) atomicAdd(out + (size_t)ca * ld_out + cb, red[i]);
    }
}

template <int TA, int TB>
static void launch_rr_narrow(const float* A, int Ca, const float* Bm, int Cb, long long M, float* out, int ld_out,
                             cudaStream_t st) {
    long long ctas = (long long)kNumSMs * 4;
    const long long max_ctas = (M + 4 * kRrNarrowRows - 1) / (4 * kRrNarrowRows);     // at least 256 rows per CTA
    if (ctas > max_ctas) ctas = max_ctas;
    if (ctas < 1) ctas = 1;
    long long rows_per_cta = (M + ctas - 1) / ctas;
    rows_per_cta = (rows_per_cta + kRrNarrowRows - 1) / kRrNarrowRows * kRrNarrowRows;
    ctas = (M + rows_per_cta - 1) / rows_per_cta;
    rowreduce_gemm_narrow_kernel<TA, TB><<<(unsigned)ctas, 256, 0, st>>>(A, Ca, Bm, Cb, M, rows_per_cta, out, ld_out);
}

// Thin variant (one side <= 8 channels, the other <= 32: fc_start's 8x3, the class logits' 2x32): HBM-streaming.
// A lane owns one channel of the wide side, loops over the narrow side; rows are strided over all warps of the grid;
// one shared-memory reduction per CTA, then one atomic per element and CTA.
__global__ void __launch_bounds__(256) rowreduce_gemm_thin_kernel(const float* __restrict__ A, int Ca,
                                                                  const float* __restrict__ Bm, int Cb, long long M,
                                                                  float* __restrict__ out, int ld_out) {
    // W = wide operand (lanes), Nw = its channels; S = narrow operand, Ns <= 8
    const bool a_wide = Ca >= Cb;
    const float* Wd = a_wide ? A : Bm;
    const float* Sd = a_wide ? Bm : A;
    const int Nw = a_wide ? Ca : Cb, Ns = a_wide ? Cb : Ca;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long gw = (long long)blockIdx.x * 8 + warp, nw = (long long)gridDim.x * 8;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
#pragma unroll 4
    for (long long r = gw; r < M; r += nw) {
        const float w = lane < Nw ? Wd[r * Nw + lane] : 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j)
            if (j < Ns) acc[j] = fmaf(w, Sd[r * Ns + j], acc[j]);
    }
    __shared__ float red[8][32];
    for (int i = threadIdx.x; i < 256; i += 256) (&red[0][0])[i] = 0.f;
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 8; ++j)
        if (j < Ns) atomicAdd(&red[j][lane], acc[j]);
    __syncthreads();
    const int i = threadIdx.x;            // 256 = 8 x 32 elements
    const int j = i / 32, l = i % 32;
    if (j < Ns && l < Nw) {
        const int ca = a_wide ? l : j, cb = a_wide ? j : l;
        atomicAdd(out + (size_t)ca * ld_out + cb, red[j][l]);
    }
}

}  // namespace r3d

using namespace r3d;

static int grid_for(long long work_items, int per_block) {
    long long blocks = (work_items + per_block - 1) / per_block;
    const long long cap = (long long)kNumSMs * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

extern "C" int r3d_bn_apply(const float* z, const double* stats, long long M, int C, const float* gamma,
                            const float* beta, const float* bias, float eps, float momentum, float* running_mean,
                            float* running_var, long long* num_batches, int act, float slope, float* y, float* save,
                            r3d_stream_t stream) {
    if (M < 0 || C <= 0 || act < 0 || act > 2) return R3D_EINVAL;
    if (C > kBnMaxC || (C % 4) != 0) return R3D_EUNSUPPORTED;
    if (M == 0) return R3D_OK;
    if (!z || !stats || !gamma || !beta || !y || !save) return R3D_EINVAL;
    if (!is_aligned(z, 16) || !is_aligned(y, 16)) return R3D_EALIGN;
    bn_apply_kernel<<<grid_for(M * C / 4, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        z, stats, M, C, gamma, beta, bias, eps, momentum, running_mean, running_var, num_batches, act, slope, y, save);
    R3D_LAUNCH_CHECK("bn_apply_kernel");
    return R3D_OK;
}

extern "C" int r3d_bn_bwd_reduce(const float* dy, const float* z, long long M, int C, const float* save,
                                 const float* beta, int act, float slope, double* stats2, r3d_stream_t stream) {
    if (M < 0 || C <= 0 || act < 0 || act > 2) return R3D_EINVAL;
    if (C > kBnMaxC || (C % 4) != 0) return R3D_EUNSUPPORTED;
    if (M == 0) return R3D_OK;
    if (!dy || !z || !save || !beta || !stats2) return R3D_EINVAL;
    if (!is_aligned(dy, 16) || !is_aligned(z, 16)) return R3D_EALIGN;
    const int rows_per_pass = 256 / (C / 4);
    bn_bwd_reduce_kernel<<<grid_for(M, rows_per_pass * 2), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        dy, z, M, C, save, beta, act, slope, stats2);
    R3D_LAUNCH_CHECK("bn_bwd_reduce_kernel");
    return R3D_OK;
}

static int bn_bwd_dz_launch(const float* dy, const float* z, long long M, int C, const float* save, const float* beta,
                            int act, float slope, const double* stats2, float* dz, float* dgb, float* absmax,
                            r3d_stream_t stream) {
    if (M < 0 || C <= 0 || act < 0 || act > 2) return R3D_EINVAL;
    if (C > kBnMaxC || (C % 4) != 0) return R3D_EUNSUPPORTED;
    if (M == 0) return R3D_OK;
    if (!dy || !z || !save || !beta || !stats2 || !dz) return R3D_EINVAL;
    if (!is_aligned(dy, 16) || !is_aligned(z, 16) || !is_aligned(dz, 16)) return R3D_EALIGN;
    bn_bwd_dz_kernel<<<grid_for(M * C / 4, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(dy, z, M, C, save, beta, act,
                                                                                             slope, stats2, dz, dgb, absmax);
    R3D_LAUNCH_CHECK("bn_bwd_dz_kernel");
    return R3D_OK;
}

extern "C" int r3d_bn_bwd_dz(const float* dy, const float* z, long long M, int C, const float* save, const float* beta,
                             int act, float slope, const double* stats2, float* dz, float* dgb, r3d_stream_t stream) {
    return bn_bwd_dz_launch(dy, z, M, C, save, beta, act, slope, stats2, dz, dgb, nullptr, stream);
}

// Single-launch (cooperative, grid barrier) variants of the train-mode BatchNorm: bit 0 forward (r3d_pointwise_bn),
// bit 1 backward (r3d_bn_bwd).  Default 0: measured on the 8 x 2 500-point step (same box, CUDA-graph replay, two
// rounds) both fused 2.99 ms, forward only 2.86, backward only 2.80-2.91, none 2.77-2.89 -- a cooperative launch
// needs its whole grid resident at once and so does not overlap the side-stream kernels the step relies on, which
// costs more than the saved launch.  Returns the previous mask; a negative argument only queries.
static std::atomic<int> g_bn_fused{0};
int r3d::bn_fused_mask() { return g_bn_fused.load(); }
extern "C" int r3d_bn_set_fused(int mask) {
    if (mask < 0 || mask > 3) return g_bn_fused.load();
    return g_bn_fused.exchange(mask);
}

// r3d_bn_bwd + absmax_dz (nullable; caller-zeroed device scalar) = atomic max with max |dz|
extern "C" int r3d_bn_bwd_absmax(const float* dy, const float* z, long long M, int C, const float* save, const float* beta,
                                 int act, float slope, double* stats2, float* dz, float* dgb, float* absmax_dz,
                                 r3d_stream_t stream) {
    if (M < 0 || C <= 0 || act < 0 || act > 2) return R3D_EINVAL;
    if (C > kBnMaxC || (C % 4) != 0) return R3D_EUNSUPPORTED;
    if (M == 0) return R3D_OK;
    if (!dy || !z || !save || !beta || !stats2 || !dz) return R3D_EINVAL;
    if (!is_aligned(dy, 16) || !is_aligned(z, 16) || !is_aligned(dz, 16)) return R3D_EALIGN;
    // one cooperative launch while the whole grid can be co-resident (<= 4 CTAs per SM asked for) and the tensor is
    // small enough for the launch cost to matter; two launches otherwise
    static int max_blocks_per_sm = -1;
    if (max_blocks_per_sm < 0) {
        int n = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, bn_bwd_fused_kernel, 256, 0) != cudaSuccess) n = 0;
        max_blocks_per_sm = n;
    }
    const int rows_per_pass = 256 / (C / 4) > 0 ? 256 / (C / 4) : 1;
    const long long want = (M + (long long)rows_per_pass * 2 - 1) / ((long long)rows_per_pass * 2);
    const long long resident = (long long)kNumSMs * (max_blocks_per_sm < 4 ? max_blocks_per_sm : 4);
    if ((bn_fused_mask() & 2) && M * C <= (1LL << 22) && resident > 0) {
        const long long blocks = want < resident ? want : resident;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)(blocks < 1 ? 1 : blocks));
        cfg.blockDim = dim3(256);
        cfg.dynamicSmemBytes = 0;
        cfg.stream = static_cast<cudaStream_t>(stream);
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeCooperative;
        attr[0].val.cooperative = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        R3D_CUDA_TRY(cudaLaunchKernelEx(&cfg, bn_bwd_fused_kernel, dy, z, M, C, save, beta, act, slope, stats2, dz, dgb, absmax_dz));
        R3D_LAUNCH_CHECK("bn_bwd_fused_kernel");
        return R3D_OK;
    }
    int rc = r3d_bn_bwd_reduce(dy, z, M, C, save, beta, act, slope, stats2, stream);
    if (rc != R3D_OK) return rc;
    return bn_bwd_dz_launch(dy, z, M, C, save, beta, act, slope, stats2, dz, dgb, absmax_dz, stream);
}

extern "C" int r3d_bn_bwd(const float* dy, const float* z, long long M, int C, const float* save, const float* beta,
                          int act, float slope, double* stats2, float* dz, float* dgb, r3d_stream_t stream) {
    return r3d_bn_bwd_absmax(dy, z, M, C, save, beta, act, slope, stats2, dz, dgb, nullptr, stream);
}

// ------------------------------------------------------------------------------ rowreduce_gemm_rows8
// Weight gradients of the narrowest layers over millions of rows (level 0 of a large batch: 8x8, 8x3, 2x32, 8x16, 32x8,
// 64x8 at 2.6 M rows): pure streaming, a few FMAs per loaded float.  The thin kernel above puts one ROW on a warp (8 of
// 32 lanes busy at 8 channels, 0.6-2.5 TB/s).  Here a thread owns an 8 x 8 tile of the output and walks rows: the
// T = ceil(Ca/8) ceil(Cb/8) threads of a row are adjacent (a warp reads 32/T consecutive rows: contiguous memory), 64
// accumulators in registers, a shuffle tree over the lanes that share a tile at the end, then shared-memory and global
// atomics once per CTA.
__global__ void __launch_bounds__(256) rowreduce_gemm_rows8_kernel(const float* __restrict__ A, int Ca,
                                                                   const float* __restrict__ Bm, int Cb, long long M,
                                                                   float* __restrict__ out, int ld_out, int na, int nb) {
    const int T = na * nb;                                   // threads per row: a power of two <= 32
    const int tile = threadIdx.x % T, ia = tile / nb, ib = tile % nb;
    const int rows_per_block = 256 / T;
    const int a0 = ia * 8, b0 = ib * 8;
    const bool vec_a = (Ca % 4 == 0) && a0 + 8 <= Ca, vec_b = (Cb % 4 == 0) && b0 + 8 <= Cb;
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
    for (long long r = (long long)blockIdx.x * rows_per_block + threadIdx.x / T; r < M;
         r += (long long)gridDim.x * rows_per_block) {
        float a[8], b[8];
        const float* ap = A + r * Ca + a0;
        const float* bp = Bm + r * Cb + b0;
        if (vec_a) {
            const float4 u = __ldg(reinterpret_cast<const float4*>(ap)), v = __ldg(reinterpret_cast<const float4*>(ap + 4));
            a[0] = u.x, a[1] = u.y, a[2] = u.z, a[3] = u.w, a[4] = v.x, a[5] = v.y, a[6] = v.z, a[7] = v.w;
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = a0 + i < Ca ? __ldg(ap + i) : 0.f;
        }
        if (vec_b) {
            const float4 u = __ldg(reinterpret_cast<const float4*>(bp)), v = __ldg(reinterpret_cast<const float4*>(bp + 4));
            b[0] = u.x, b[1] = u.y, b[2] = u.z, b[3] = u.w, b[4] = v.x, b[5] = v.y, b[6] = v.z, b[7] = v.w;
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) b[j] = b0 + j < Cb ? __ldg(bp + j) : 0.f;
        }
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __shared__ float red[32][64];                            // [tile][8 x 8]
    for (int i = threadIdx.x; i < 32 * 64; i += 256) (&red[0][0])[i] = 0.f;
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            float v = acc[i][j];
            for (int o = 16; o >= T; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);   // lanes with the same tile
            if ((threadIdx.x & 31) < T) atomicAdd(&red[tile][i * 8 + j], v);
        }
    __syncthreads();
    for (int e = threadIdx.x; e < T * 64; e += 256) {
        const int t = e / 64, i = (e % 64) / 8, j = e % 8;
        const int ca = (t / nb) * 8 + i, cb = (t % nb) * 8 + j;
        if (ca < Ca && cb < Cb) atomicAdd(out + (size_t)ca * ld_out + cb, red[t][i * 8 + j]);
    }
}

extern "C" int r3d_rowreduce_gemm(const float* A, int Ca, const float* Bm, int Cb, long long M, float* out, int ld_out,
                                  r3d_stream_t stream) {
    if (M < 0 || Ca <= 0 || Cb <= 0) return R3D_EINVAL;
    if (M == 0) return R3D_OK;
    if (!A || !Bm || !out) return R3D_EINVAL;
    if (ld_out == 0) ld_out = Cb;
    {
        // a few channels each way over very many rows: thread-owned 8 x 8 output tiles (see rowreduce_gemm_rows8_kernel)
        const int na = (Ca + 7) / 8, nb = (Cb + 7) / 8, T = na * nb;
        static const bool rows8 = getenv("R3D_RR_ROWS8_OFF") == nullptr;       // tuning hook
        if (rows8 && M >= 131072 && T <= 8 && (T & (T - 1)) == 0 && is_aligned(A, 16) && is_aligned(Bm, 16)) {
            const int rows_per_block = 256 / T;
            long long blocks = (M + (long long)rows_per_block * 16 - 1) / ((long long)rows_per_block * 16);
            if (blocks > (long long)kNumSMs * 6) blocks = (long long)kNumSMs * 6;
            rowreduce_gemm_rows8_kernel<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(A, Ca, Bm, Cb, M, out,
                                                                                                       ld_out, na, nb);
            R3D_LAUNCH_CHECK("rowreduce_gemm_rows8_kernel");
            return R3D_OK;
        }
    }
    if ((Ca <= 8 && Cb <= 32) || (Cb <= 8 && Ca <= 32)) {
        long long blocks = (M + 8 * 8 - 1) / (8 * 8);             // >= 8 rows per warp: the row loop is latency-bound
        if (blocks > (long long)kNumSMs * 8) blocks = (long long)kNumSMs * 8;
        if (blocks < 1) blocks = 1;
        rowreduce_gemm_thin_kernel<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(A, Ca, Bm, Cb, M, out,
                                                                                                  ld_out);
        R3D_LAUNCH_CHECK("rowreduce_gemm_thin_kernel");
        return R3D_OK;
    }
    const bool vec = (Ca % 4) == 0 && (Cb % 4) == 0 && is_aligned(A, 16) && is_aligned(Bm, 16);
    if (vec && Ca <= 64 && Cb <= 64 && (Ca <= 32 || Cb <= 32) && M >= 4096) {
        cudaStream_t st = static_cast<cudaStream_t>(stream);
        const int ta = Ca <= 16 ? 16 : (Ca <= 32 ? 32 : 64), tb = Cb <= 16 ? 16 : (Cb <= 32 ? 32 : 64);
        if (ta == 16 && tb == 16) launch_rr_narrow<16, 16>(A, Ca, Bm, Cb, M, out, ld_out, st);
        else if (ta == 16 && tb == 32) launch_rr_narrow<16, 32>(A, Ca, Bm, Cb, M, out, ld_out, st);
        else if (ta == 32 && tb == 16) launch_rr_narrow<32, 16>(A, Ca, Bm, Cb, M, out, ld_out, st);
        else if (ta == 32 && tb == 32) launch_rr_narrow<32, 32>(A, Ca, Bm, Cb, M, out, ld_out, st);
        else if (ta == 16 && tb == 64) launch_rr_narrow<16, 64>(A, Ca, Bm, Cb, M, out, ld_out, st);
        else if (ta == 64 && tb == 16) launch_rr_narrow<64, 16>(A, Ca, Bm, Cb, M, out, ld_out, st);
        else if (ta == 32 && tb == 64) launch_rr_narrow<32, 64>(A, Ca, Bm, Cb, M, out, ld_out, st);
        else launch_rr_narrow<64, 32>(A, Ca, Bm, Cb, M, out, ld_out, st);
        R3D_LAUNCH_CHECK("rowreduce_gemm_narrow_kernel");
        return R3D_OK;
    }
    const bool big = vec && Ca >= 128 && Cb >= 128;
    const int tile = big ? kRrBig : kRrTile;
    const int ga = ceil_div(Ca, tile), gb = ceil_div(Cb, tile);
    // enough row chunks to fill the machine, at least 128 rows each
    long long chunks = ((long long)kNumSMs * (big ? 2 : 4) + ga * gb - 1) / (ga * gb);
    const long long max_chunks = (M + 2 * kRrKC2 - 1) / (2 * kRrKC2);   // at least 64 rows per CTA
    if (chunks > max_chunks) chunks = max_chunks;
    if (chunks < 1) chunks = 1;
    if (chunks > 65535) chunks = 65535;
    long long rows_per_cta = (M + chunks - 1) / chunks;
    rows_per_cta = (rows_per_cta + kRrKC2 - 1) / kRrKC2 * kRrKC2;
    chunks = (M + rows_per_cta - 1) / rows_per_cta;
    dim3 grid(ga, gb, (unsigned)chunks);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (big)
        rowreduce_gemm_big_kernel<<<grid, 256, 0, st>>>(A, Ca, Bm, Cb, M, rows_per_cta, out, ld_out);
    else if (vec)
        rowreduce_gemm_fast_kernel<<<grid, 256, 0, st>>>(A, Ca, Bm, Cb, M, rows_per_cta, out, ld_out);
    else
        rowreduce_gemm_kernel<<<grid, 256, 0, st>>>(A, Ca, Bm, Cb, M, rows_per_cta, out, ld_out);
    R3D_LAUNCH_CHECK("rowreduce_gemm_kernel");
    return R3D_OK;
}
