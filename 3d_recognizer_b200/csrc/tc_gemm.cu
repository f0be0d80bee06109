// tcgen05 (5th-generation tensor core) FP32-accurate GEMM for sm_100a:  C (M,N) = A (M,K) W (N,K)^T
//
// fp32 parity (1e-4 on logits after four encoder levels) rules out plain TF32 inputs (10-bit mantissa, ~1e-3
// per product).  Each operand is split on the fly into hi = x with the low 13 mantissa bits cleared (exactly a
// TF32 number) and lo = x - hi (exact in fp32), and three kind::tf32 MMAs are accumulated in TMEM:
//     A W^T ~= Ahi Whi^T + Ahi Wlo^T + Alo Whi^T          (the dropped Alo Wlo^T term is O(2^-22))
// which restores ~fp32 accuracy at a third of the TF32 rate (`terms` = 1 issues only the first product).
//
// Structure (one CTA = 128 rows of A, all N <= 256 columns, 128 threads):
//   * all threads load a 32-wide K block of A and W from global memory, split it and store it in shared memory in
//     the canonical K-major no-swizzle UMMA layout ([k/4][row][4] : core matrices of 8 rows x 16 bytes, LBO = rows,
//     SBO = 8 in 16-byte units), two stages;
//   * fence.proxy.async + barrier, then ONE thread issues 4 x 3 tcgen05.mma (M=128, N, K=8) on shared-memory
//     descriptors and commits them to the stage's mbarrier (tcgen05.commit), which frees the stage for the refill
//     two blocks later while the tensor core keeps running;
//   * accumulators live in TMEM (N columns x 128 lanes); the epilogue reads them back with tcgen05.ld (warp w owns
//     lanes 32w..32w+31 = rows) and stores C.
#include "tc_common.cuh"

namespace r3d {

constexpr int kTcThreads = 128;
constexpr int kTcKB = 32;          // K block per stage
constexpr int kTcMaxN = 256;

// smem: Ahi[2][8][128] float4, Alo same, Whi[2][8][Np] float4, Wlo same, then barriers + tmem slot
__global__ void __launch_bounds__(kTcThreads, 1) tc_gemm_kernel(const float* __restrict__ A, const float* __restrict__ W,
                                                                float* __restrict__ C, int M, int N, int K, int terms,
                                                                int tmem_cols) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float4* Ahi = reinterpret_cast<float4*>(smem_raw);
    float4* Alo = Ahi + 2 * 8 * 128;
    float4* Whi = Alo + 2 * 8 * 128;
    float4* Wlo = Whi + 2 * 8 * N;
    uint64_t* bars = reinterpret_cast<uint64_t*>(Wlo + 2 * 8 * N);     // [0],[1]: stage free; [2]: all done
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int m0 = blockIdx.x * 128;

    if (tid == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        mbar_init(&bars[2], 1);
        mbar_fence_init();
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"((uint32_t)tmem_cols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d = *tmem_slot;
    const uint32_t idesc = umma_idesc_tf32(128, N);

    const int nkb = (K + kTcKB - 1) / kTcKB;
    for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb & 1;
        if (kb >= 2) mbar_wait(&bars[s], (uint32_t)((kb / 2 - 1) & 1));   // MMAs that read stage s have completed
        const int k0 = kb * kTcKB;
        // ---- A block: thread = row
        {
            const int m = m0 + tid;
            const float* row = A + (size_t)min(m, M - 1) * K;
#pragma unroll
            for (int kq = 0; kq < 8; ++kq) {
                const int k = k0 + kq * 4;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (m < M && k < K) v = *reinterpret_cast<const float4*>(row + k);    // K % 4 == 0
                float4 hi, lo;
                split_tf32(v, hi, lo);
                Ahi[(s * 8 + kq) * 128 + tid] = hi;
                Alo[(s * 8 + kq) * 128 + tid] = lo;
            }
        }
        // ---- W block: thread = output channel
        for (int n = tid; n < N; n += kTcThreads) {
            const float* row = W + (size_t)n * K;
#pragma unroll
            for (int kq = 0; kq < 8; ++kq) {
                const int k = k0 + kq * 4;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (k < K) v = *reinterpret_cast<const float4*>(row + k);
                float4 hi, lo;
                split_tf32(v, hi, lo);
                Whi[(s * 8 + kq) * N + n] = hi;
                Wlo[(s * 8 + kq) * N + n] = lo;
            }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the tensor core
        __syncthreads();
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
            for (int ks = 0; ks < kTcKB / 8; ++ks) {
                const uint64_t ah = umma_desc(Ahi + (s * 8 + 2 * ks) * 128, 128, 8);
                const uint64_t al = umma_desc(Alo + (s * 8 + 2 * ks) * 128, 128, 8);
                const uint64_t wh = umma_desc(Whi + (s * 8 + 2 * ks) * N, (uint32_t)N, 8);
                const uint64_t wl = umma_desc(Wlo + (s * 8 + 2 * ks) * N, (uint32_t)N, 8);
                umma_tf32(tmem_d, ah, wh, idesc, (kb > 0 || ks > 0) ? 1u : 0u);
                if (terms == 3) {
                    umma_tf32(tmem_d, ah, wl, idesc, 1u);
                    umma_tf32(tmem_d, al, wh, idesc, 1u);
                }
            }
            umma_commit(&bars[s]);
            if (kb == nkb - 1) umma_commit(&bars[2]);
        }
    }
    mbar_wait(&bars[2], 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

    // ---- epilogue: TMEM lane = row (warp w owns lanes 32w..32w+31), 32 columns per load
    const int m = m0 + warp * 32 + lane;
    for (int c0 = 0; c0 < N; c0 += 32) {
        uint32_t r[32];
        const uint32_t taddr = tmem_d + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
              "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
              "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
              "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
            : "r"(taddr)
            : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (m < M) {
            float* out = C + (size_t)m * N + c0;
#pragma unroll
            for (int j = 0; j < 32; j += 4)
                if (c0 + j < N)
                    *reinterpret_cast<float4*>(out + j) = make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]),
                                                                      __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"((uint32_t)tmem_cols) : "memory");
}

}  // namespace r3d

using namespace r3d;

// C (M,N) = A (M,K) W (N,K)^T on the tcgen05 tensor cores; terms = 3 (fp32-accurate 3xTF32) or 1 (plain TF32).
// N a multiple of 16 in [16, 256] (and of 32 for the epilogue), K a multiple of 4; A, W, C dense and 16-byte aligned.
extern "C" int r3d_tc_gemm(const float* A, const float* W, float* C, int M, int N, int K, int terms, r3d_stream_t stream) {
    if (M < 0 || N <= 0 || K <= 0 || (terms != 1 && terms != 3)) return R3D_EINVAL;
    if (M == 0) return R3D_OK;
    if (!A || !W || !C) return R3D_EINVAL;
    if (N > kTcMaxN || (N % 32) != 0 || (K % 4) != 0) return R3D_EUNSUPPORTED;
    if (!is_aligned(A, 16) || !is_aligned(W, 16) || !is_aligned(C, 16)) return R3D_EALIGN;
    int cols = 32;
    while (cols < N) cols <<= 1;
    const size_t smem = (size_t)(2 * 2 * 8 * 128 + 2 * 2 * 8 * N) * sizeof(float4) + 3 * sizeof(uint64_t) + 16;
    R3D_CUDA_TRY(cudaFuncSetAttribute(tc_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    tc_gemm_kernel<<<ceil_div(M, 128), kTcThreads, smem, static_cast<cudaStream_t>(stream)>>>(A, W, C, M, N, K, terms, cols);
    R3D_LAUNCH_CHECK("tc_gemm_kernel");
    return R3D_OK;
}
