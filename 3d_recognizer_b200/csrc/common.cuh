// Shared helpers for the sm_100a kernels of the RandLA-Net hot path.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/r3d_b200.h"

namespace r3d {

void set_cuda_error(cudaError_t e, const char* where);
void count_launch();  // every kernel launch of the library is counted (r3d_launch_count)

#define R3D_CUDA_TRY(expr)                                  \
    do {                                                    \
        cudaError_t _e = (expr);                            \
        if (_e != cudaSuccess) {                            \
            ::r3d::set_cuda_error(_e, #expr);               \
            return R3D_ECUDA;                               \
        }                                                   \
    } while (0)

#define R3D_LAUNCH_CHECK(what)                              \
    do {                                                    \
        ::r3d::count_launch();                              \
        cudaError_t _e = cudaGetLastError();                \
        if (_e != cudaSuccess) {                            \
            ::r3d::set_cuda_error(_e, what);                \
            return R3D_ECUDA;                               \
        }                                                   \
    } while (0)

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
static inline bool is_aligned(const void* p, size_t a) { return (reinterpret_cast<uintptr_t>(p) % a) == 0; }
static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

constexpr int kNumSMs = 148;  // B200

// ---- mbarrier + 1-D TMA bulk copy (cp.async.bulk -> SASS UBLKCP) ---------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}
// global -> shared bulk copy; dst/src 16-byte aligned, bytes a multiple of 16; completes on `bar`.
__device__ __forceinline__ void tma_bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

}  // namespace r3d
