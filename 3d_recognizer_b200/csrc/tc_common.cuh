// tcgen05 helpers shared by the tensor-core kernels (tc_gemm.cu, lfa_tc.cu): UMMA shared-memory / instruction
// descriptors (bit layouts as in CUTLASS cute/arch/mma_sm100_desc.hpp), MMA issue, commit, TMEM alloc / load, and the
// hi/lo split of the 3xTF32 scheme.
#pragma once
#include "common.cuh"

namespace r3d {

__device__ __forceinline__ uint64_t umma_desc(const void* smem, uint32_t lbo16, uint32_t sbo16) {
    const uint32_t addr = smem_u32(smem);
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3FFF);            // start address, 16-byte units
    d |= (uint64_t)(lbo16 & 0x3FFF) << 16;            // leading-dimension byte offset (K direction), 16-byte units
    d |= (uint64_t)(sbo16 & 0x3FFF) << 32;            // stride byte offset (8-row groups), 16-byte units
    d |= (uint64_t)1 << 46;                           // descriptor version (Blackwell)
    return d;                                         // base offset 0, layout type 0 = no swizzle
}

__device__ __forceinline__ uint32_t umma_idesc_tf32(int M, int N) {
    uint32_t d = 0;
    d |= 1u << 4;                 // accumulator format F32
    d |= 2u << 7;                 // A format TF32
    d |= 2u << 10;                // B format TF32
    d |= (uint32_t)(N >> 3) << 17;
    d |= (uint32_t)(M >> 4) << 24;
    return d;                     // K-major A and B, no negate, dense
}

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
// Round-to-nearest TF32 (cvt.rna): hi = rn(x), lo = rn(x - hi).  Plain masking (truncation) leaves the tensor core to
// truncate lo as well, and truncation errors all point towards zero: they add up linearly over K (measured 9e-6 at
// K = 1024) instead of as a random walk.
__device__ __forceinline__ float rn_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}
__device__ __forceinline__ void split_tf32(float4 v, float4& hi, float4& lo) {
    hi.x = rn_tf32(v.x);
    hi.y = rn_tf32(v.y);
    hi.z = rn_tf32(v.z);
    hi.w = rn_tf32(v.w);
    lo = make_float4(rn_tf32(v.x - hi.x), rn_tf32(v.y - hi.y), rn_tf32(v.z - hi.z), rn_tf32(v.w - hi.w));
}
__device__ __forceinline__ void split_tf32_1(float v, float& hi, float& lo) {
    hi = rn_tf32(v);
    lo = rn_tf32(v - hi);
}

// one warp allocates `cols` (power of two >= 32) TMEM columns and publishes the base address in *slot (shared memory)
__device__ __forceinline__ void tmem_alloc_warp(uint32_t* slot, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_warp(uint32_t base, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(cols) : "memory");
}
__device__ __forceinline__ void tc_fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// 32 consecutive accumulator columns of this thread's TMEM lane (row) -> registers
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
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
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
}

}  // namespace r3d
