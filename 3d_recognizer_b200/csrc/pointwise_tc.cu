// Per-point layer on the tcgen05 tensor cores:  y = act(scale * (W [xa[g] ; xb]) + shift)  for wide layers.
//
// "The shared-MLP 1x1 convolutions go on tcgen05 tensor cores only where channel widths make them a real dense
// contraction" (north_star): C_in >= 32 and C_out a multiple of 32 (encoder levels 1-3, bottleneck, decoder).  Same
// operator, arguments and epilogue (affine, activation, BatchNorm batch statistics) as pw_gemm_fast_kernel in
// pointwise.cu (randlanet/utils/modules.py:60-104 and the gather / concat call sites :359-363, :600-602).
// fp32 parity is kept by the 3xTF32 scheme of tc_gemm.cu (hi/lo operand split, three kind::tf32 MMAs per K step,
// fp32 accumulation in TMEM; measured error ~1e-6 relative).
//
// CTA = 128 threads = 128 rows, up to 256 output channels (gridDim.y column blocks).  16-channel K blocks, two stages
// of 32 KB + N*256 B: two or more CTAs per SM so that one CTA's global loads overlap another's MMAs; the loads of
// block kb+1 are issued into registers before the wait on block kb's stage.  Thread t owns row t: it resolves its
// source row pointers once (gather index, batch stride, concat), stages its row's K block (split hi/lo) and W rows
// t, t+128, and in the epilogue reads its row of accumulators with tcgen05.ld (TMEM lane = row).
#include "pointwise_common.cuh"
#include "tc_common.cuh"

namespace r3d {

constexpr int kPwTcKB = 16;         // input channels per stage
constexpr int kPwTcMaxN = 256;

__global__ void __launch_bounds__(128) pw_tc_kernel(PwArgs a, int tmem_cols) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int n0 = blockIdx.y * kPwTcMaxN;
    const int N = min(kPwTcMaxN, a.cout - n0);          // multiple of 32
    const int Kc = a.ca + a.cb;
    float4* Ahi = reinterpret_cast<float4*>(smem_raw);  // [2][4][128]
    float4* Alo = Ahi + 2 * 4 * 128;
    float4* Whi = Alo + 2 * 4 * 128;                    // [2][4][N]
    float4* Wlo = Whi + 2 * 4 * N;
    double* csum = reinterpret_cast<double*>(Wlo + 2 * 4 * N);  // [2][256] fp64: run-to-run reproducible statistics
    uint64_t* bars = reinterpret_cast<uint64_t*>(csum + 2 * kPwTcMaxN);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const long long M = (long long)a.B * a.n;
    const long long m = (long long)blockIdx.x * 128 + tid;
    const bool live = m < M;

    if (tid == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        mbar_init(&bars[2], 1);
        mbar_fence_init();
    }
    if (warp == 0) tmem_alloc_warp(tmem_slot, (uint32_t)tmem_cols);
    if (a.stats)
        for (int i = tid; i < 2 * kPwTcMaxN; i += 128) csum[i] = 0.0;
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem = *tmem_slot;
    const uint32_t idesc = umma_idesc_tf32(128, N);

    // source rows of this thread
    const float* rowA = nullptr;
    const float* rowB = nullptr;
    int bcloud = 0, nrow = 0;
    if (live) {
        bcloud = (int)(m / a.n);
        nrow = (int)(m % a.n);
        rowA = src_row(a, bcloud, nrow, false);
        if (a.cb > 0) rowB = src_row(a, bcloud, nrow, true);
    }
    const int wrows = (N + 127) / 128;     // W rows staged by this thread: n = tid, tid + 128

    float4 ra[4], rw[2][4];
    auto prefetch = [&](int k0) {
#pragma unroll
        for (int kq = 0; kq < 4; ++kq) {
            const int k = k0 + kq * 4;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (live && k < Kc) v = *reinterpret_cast<const float4*>(k < a.ca ? rowA + k : rowB + (k - a.ca));
            ra[kq] = v;
        }
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int n = tid + 128 * r;
            if (r < wrows && n < N) {
#pragma unroll
                for (int kq = 0; kq < 4; ++kq) {
                    const int k = k0 + kq * 4;
                    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (k < Kc) {
                        if (a.w_out_in) {
                            v = *reinterpret_cast<const float4*>(a.wT + (size_t)(n0 + n) * Kc + k);
                        } else {
                            const float* w = a.wT + (size_t)k * a.cout + n0 + n;
                            v = make_float4(w[0], w[a.cout], w[2 * (size_t)a.cout], w[3 * (size_t)a.cout]);
                        }
                    }
                    rw[r][kq] = v;
                }
            }
        }
    };

    const int nkb = (Kc + kPwTcKB - 1) / kPwTcKB;
    prefetch(0);
    for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb & 1;
        if (kb >= 2) mbar_wait(&bars[s], (uint32_t)((kb / 2 - 1) & 1));
#pragma unroll
        for (int kq = 0; kq < 4; ++kq) {
            float4 hi, lo;
            split_tf32(ra[kq], hi, lo);
            Ahi[(s * 4 + kq) * 128 + tid] = hi;
            Alo[(s * 4 + kq) * 128 + tid] = lo;
        }
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int n = tid + 128 * r;
            if (r < wrows && n < N) {
#pragma unroll
                for (int kq = 0; kq < 4; ++kq) {
                    float4 hi, lo;
                    split_tf32(rw[r][kq], hi, lo);
                    Whi[(s * 4 + kq) * N + n] = hi;
                    Wlo[(s * 4 + kq) * N + n] = lo;
                }
            }
        }
        fence_async_smem();
        __syncthreads();
        if (tid == 0) {
            tc_fence_after_sync();
#pragma unroll
            for (int ks = 0; ks < kPwTcKB / 8; ++ks) {
                const uint64_t ah = umma_desc(Ahi + (s * 4 + 2 * ks) * 128, 128, 8);
                const uint64_t al = umma_desc(Alo + (s * 4 + 2 * ks) * 128, 128, 8);
                const uint64_t wh = umma_desc(Whi + (s * 4 + 2 * ks) * N, (uint32_t)N, 8);
                const uint64_t wl = umma_desc(Wlo + (s * 4 + 2 * ks) * N, (uint32_t)N, 8);
                umma_tf32(tmem, ah, wh, idesc, (kb > 0 || ks > 0) ? 1u : 0u);
                umma_tf32(tmem, ah, wl, idesc, 1u);
                umma_tf32(tmem, al, wh, idesc, 1u);
            }
            umma_commit(&bars[s]);
            if (kb == nkb - 1) umma_commit(&bars[2]);
        }
        if (kb + 1 < nkb) prefetch((kb + 1) * kPwTcKB);
    }
    mbar_wait(&bars[2], 0);
    tc_fence_after_sync();

    // ---- epilogue: affine + activation, row store, batch statistics
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    float* yr = live ? a.y + (size_t)bcloud * a.y_bstride + (size_t)nrow * a.y_ld + n0 : nullptr;
#pragma unroll 1
    for (int c0 = 0; c0 < N; c0 += 32) {
        float v[32];
        tmem_ld32(tmem + lane_base + (uint32_t)c0, v);
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            const int c = n0 + c0 + j;
            const float sc = a.scale ? a.scale[c] : 1.f, sh = a.shift ? a.shift[c] : 0.f;
            v[j] = apply_act(fmaf(v[j], sc, sh), a.act, a.slope);
        }
        if (live) {
#pragma unroll
            for (int j = 0; j < 32; j += 4)
                *reinterpret_cast<float4*>(yr + c0 + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        }
        if (a.stats) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                float s1 = live ? v[j] : 0.f, s2 = live ? v[j] * v[j] : 0.f;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    s1 += __shfl_xor_sync(0xffffffffu, s1, o);
                    s2 += __shfl_xor_sync(0xffffffffu, s2, o);
                }
                if (lane == 0) {
                    atomicAdd(&csum[c0 + j], (double)s1);
                    atomicAdd(&csum[kPwTcMaxN + c0 + j], (double)s2);
                }
            }
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (a.stats) {
        for (int i = tid; i < N; i += 128) {
            atomicAdd(a.stats + n0 + i, csum[i]);
            atomicAdd(a.stats + a.cout + n0 + i, csum[kPwTcMaxN + i]);
        }
    }
    if (warp == 0) tmem_dealloc_warp(tmem, (uint32_t)tmem_cols);
}

// Where the tensor-core kernel wins (tools/pointwise_bench.py, profiles/r01_pointwise_tc_vs_fp32.txt): every CTA
// re-splits its W block into TF32 hi/lo and runs a two-stage load -> split -> MMA chain, so it needs (a) enough rows
// to fill the machine several times over and (b) either a long contraction (C_in >= 256: 59 vs 39 TFLOP/s at
// 512 -> 256) or a narrow output (C_out <= 32, where the FP32 kernel is HBM/latency-bound); mid-size layers
// (64 -> 128, 128 -> 256) stay on the FP32 kernels, which are 15-25 % faster there.  Few-row layers stay on FP32 in
// any case: they are latency-bound, and a train-mode BatchNorm over a handful of rows (eps 1e-6) amplifies the
// ~2e-6 error of the tensor-core accumulation a thousandfold on near-constant channels (measured: 5e-3 on the
// bottleneck's dgamma at 20 rows).
constexpr long long kPwTcMinRows = 32768;

bool pw_tc_eligible(const PwArgs& a, bool force) {
    const int Kc = a.ca + a.cb;
    if ((long long)a.B * a.n < (force ? 4096 : kPwTcMinRows)) return false;
    if (!force && !(Kc >= 256 || a.cout <= 32)) return false;
    return (a.ca % 4 == 0) && (a.cb % 4 == 0) && Kc >= 32 && a.cout >= 32 && (a.cout % 32) == 0 && !a.transpose_out &&
           (a.y_ld % 4) == 0 && (a.y_bstride % 4) == 0;
}

int pw_tc_launch(const PwArgs& a, cudaStream_t st) {
    const long long M = (long long)a.B * a.n;
    const int Nmax = a.cout < kPwTcMaxN ? a.cout : kPwTcMaxN;
    int cols = 32;
    while (cols < Nmax) cols <<= 1;
    const size_t smem = (size_t)(2 * 2 * 4 * 128 + 2 * 2 * 4 * Nmax) * sizeof(float4) + 2 * kPwTcMaxN * sizeof(double) +
                        3 * sizeof(uint64_t) + 16;
    R3D_CUDA_TRY(cudaFuncSetAttribute(pw_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((unsigned)((M + 127) / 128), (unsigned)ceil_div(a.cout, kPwTcMaxN));
    pw_tc_kernel<<<grid, 128, smem, st>>>(a, cols);
    R3D_LAUNCH_CHECK("pw_tc_kernel");
    return R3D_OK;
}

}  // namespace r3d
