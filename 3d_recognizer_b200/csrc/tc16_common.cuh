// tcgen05 helpers of the split-fp16 ("fp16x2") tensor-core kernels (lfa_cl.cu, tc16_probe.cu).
//
// Numerics.  An fp32 operand x is carried as two fp16 numbers, hi = rn16(s x) and lo = rn16(s x - hi), with s a power
// of two chosen per tensor so that |s x| stays far below the fp16 maximum.  hi + lo holds 22 mantissa bits of s x
// (absolute floor 2^-25, the fp16 subnormal half-spacing), the same as the hi/lo pair of a 3xTF32 split, in HALF the
// shared-memory bytes (4 per value instead of 8) and at TWICE the tensor-core rate (kind::f16 vs kind::tf32):
//     A B^T ~= (Ahi Bhi^T + Ahi Blo^T + Alo Bhi^T) / (sA sB)         fp32 accumulation in TMEM
// The dropped lo x lo term is O(2^-22) relative.  Products of two fp16 numbers are exact in fp32.
//
// Operand tiles live in shared memory as 128-byte core matrices (8 x 16 bytes, no swizzle).  One tile serves both
// major-nesses: a "unit" of 16 bytes holds 8 consecutive elements along dimension U at one index of dimension V, and
// 8 consecutive V indices are adjacent; read as K-major with K = U or as MN-major with MN = U (UMMA canonical
// INTERLEAVE layouts, cute/atom/mma_traits_sm100.hpp).  In both readings the descriptor's LBO is the byte stride
// between core matrices along K and SBO the stride along M/N.
#pragma once
#include <cuda_fp16.h>
#include "tc_common.cuh"

namespace r3d {

// kind::f16 instruction descriptor: fp16 A and B, fp32 accumulator; a_mn / b_mn = 1 reads the operand MN-major.
__device__ __forceinline__ uint32_t umma_idesc_f16(int M, int N, int a_mn, int b_mn) {
    uint32_t d = 0;
    d |= 1u << 4;                      // accumulator format F32
    d |= (uint32_t)(a_mn & 1) << 15;   // A major-ness (0 = K-major)
    d |= (uint32_t)(b_mn & 1) << 16;   // B major-ness
    d |= (uint32_t)(N >> 3) << 17;
    d |= (uint32_t)(M >> 4) << 24;
    return d;                          // A/B format 0 = F16, no negate, dense
}

__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}

// descriptor from a shared-memory byte address and byte strides
__device__ __forceinline__ uint64_t umma_desc_b(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}

// (hi, lo) halves of two scaled fp32 values, packed as half2 bit patterns (low 16 bits = first value)
__device__ __forceinline__ void split16_2(float a, float b, uint32_t& hi, uint32_t& lo) {
    const __half2 h = __floats2half2_rn(a, b);
    const float2 hf = __half22float2(h);
    const __half2 l = __floats2half2_rn(a - hf.x, b - hf.y);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    lo = *reinterpret_cast<const uint32_t*>(&l);
}
__device__ __forceinline__ float2 join16_2(uint32_t hi, uint32_t lo) {
    const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&hi));
    const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&lo));
    return make_float2(a.x + b.x, a.y + b.y);
}

// 16 consecutive accumulator columns of this thread's TMEM lane -> registers (no wait: call tmem_ld_wait)
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// registers -> 16 consecutive columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}

// named barrier over `count` threads (count a multiple of 32); id 1..15 (0 is __syncthreads)
__device__ __forceinline__ void named_bar_sync(int id, int count) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

}  // namespace r3d
