// Probe / unit test of the split-fp16 tcgen05 building blocks (tc16_common.cuh): one CTA computes
//     D (M x N) = A (M x K) B (N x K)^T          M in {64, 128}, N a multiple of 16 <= 256, K a multiple of 16 <= 256
// with each operand staged in shared memory either K-major or MN-major (no swizzle), three kind::f16 MMAs per K step
// (hi hi + hi lo + lo hi), and dumps the RAW TMEM accumulator — all 128 lanes x N columns — so that the host can
// check the lane mapping of M = 64 as well as the values.  Test infrastructure of the tensor-core LFA kernels.
#include "tc16_common.cuh"

namespace r3d {

__global__ void __launch_bounds__(128, 1) tc16_probe_kernel(const float* __restrict__ A, const float* __restrict__ B,
                                                            float* __restrict__ out, int M, int N, int K, int a_mn,
                                                            int b_mn, float sa, float sb) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int a_bytes = (M / 8) * (K / 8) * 128, b_bytes = (N / 8) * (K / 8) * 128;
    unsigned char* Ahi = smem_raw;
    unsigned char* Alo = Ahi + a_bytes;
    unsigned char* Bhi = Alo + a_bytes;
    unsigned char* Blo = Bhi + b_bytes;
    uint64_t* bar = reinterpret_cast<uint64_t*>(Blo + b_bytes);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
    const int tid = threadIdx.x, warp = tid >> 5;

    if (tid == 0) {
        mbar_init(bar, 1);
        mbar_fence_init();
    }
    if (warp == 0) tmem_alloc_warp(tmem_slot, 256);
    auto stage = [&](const float* src, int MN, int mn_major, float s, unsigned char* hi, unsigned char* lo) {
        for (int e = tid; e < MN * K; e += 128) {
            const int mn = e / K, k = e % K;
            const float v = src[e] * s;
            const __half h = __float2half_rn(v);
            const __half l = __float2half_rn(v - __half2float(h));
            const int core = (k / 8) * (MN / 8) + (mn / 8);
            const int off = core * 128 + (mn_major ? (k % 8) * 16 + (mn % 8) * 2 : (mn % 8) * 16 + (k % 8) * 2);
            *reinterpret_cast<__half*>(hi + off) = h;
            *reinterpret_cast<__half*>(lo + off) = l;
        }
    };
    stage(A, M, a_mn, sa, Ahi, Alo);
    stage(B, N, b_mn, sb, Bhi, Blo);
    fence_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem = *tmem_slot;
    // zero the whole dump window first so that lanes an M = 64 instruction does not write read as 0
    {
        uint32_t z[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) z[j] = 0x7fc00000u;      // NaN pattern: untouched cells are recognisable
        for (int c0 = 0; c0 < N; c0 += 16) tmem_st16(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, z);
        tmem_st_wait();
    }
    tc_fence_before_sync();
    __syncthreads();
    if (tid == 0) {
        tc_fence_after_sync();
        const uint32_t idesc = umma_idesc_f16(M, N, a_mn, b_mn);
        const uint32_t a_lbo = (uint32_t)(M / 8) * 128, b_lbo = (uint32_t)(N / 8) * 128;
        for (int ks = 0; ks < K / 16; ++ks) {
            const uint64_t ah = umma_desc_b(smem_u32(Ahi) + ks * 2 * a_lbo, a_lbo, 128);
            const uint64_t al = umma_desc_b(smem_u32(Alo) + ks * 2 * a_lbo, a_lbo, 128);
            const uint64_t bh = umma_desc_b(smem_u32(Bhi) + ks * 2 * b_lbo, b_lbo, 128);
            const uint64_t bl = umma_desc_b(smem_u32(Blo) + ks * 2 * b_lbo, b_lbo, 128);
            umma_f16(tmem, ah, bh, idesc, ks > 0 ? 1u : 0u);
            umma_f16(tmem, ah, bl, idesc, 1u);
            umma_f16(tmem, al, bh, idesc, 1u);
        }
        umma_commit(bar);
    }
    mbar_wait(bar, 0);
    tc_fence_after_sync();
    const float inv = 1.0f / (sa * sb);
    for (int c0 = 0; c0 < N; c0 += 16) {
        uint32_t r[16];
        tmem_ld16_nowait(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, r);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) out[(size_t)tid * N + c0 + j] = __uint_as_float(r[j]) * inv;
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc_warp(tmem, 256);
}

}  // namespace r3d

using namespace r3d;

extern "C" int r3d_tc16_probe(const float* A, const float* B, float* out, int M, int N, int K, int a_mn_major,
                              int b_mn_major, float scale_a, float scale_b, r3d_stream_t stream) {
    if (!A || !B || !out) return R3D_EINVAL;
    if ((M != 64 && M != 128) || N < 16 || N > 256 || (N % 16) != 0 || K < 16 || K > 256 || (K % 16) != 0)
        return R3D_EUNSUPPORTED;
    const size_t smem = (size_t)2 * ((M / 8) * (K / 8) * 128 + (N / 8) * (K / 8) * 128) + 64;
    R3D_CUDA_TRY(cudaFuncSetAttribute(tc16_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    tc16_probe_kernel<<<1, 128, smem, static_cast<cudaStream_t>(stream)>>>(A, B, out, M, N, K, a_mn_major, b_mn_major,
                                                                            scale_a, scale_b);
    R3D_LAUNCH_CHECK("tc16_probe_kernel");
    return R3D_OK;
}
