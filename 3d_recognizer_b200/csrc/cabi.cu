// C-ABI plumbing shared by all entry points: version, error strings, per-thread CUDA error text.
#include "common.cuh"

#include <atomic>
#include <cstdio>
#include <cstring>

namespace r3d {
static thread_local char tls_cuda_error[256] = "";
static std::atomic<unsigned long long> g_launches{0};

void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

void set_cuda_error(cudaError_t e, const char* where) {
    std::snprintf(tls_cuda_error, sizeof(tls_cuda_error), "%s: %s (%s)", where, cudaGetErrorString(e),
                  cudaGetErrorName(e));
    cudaGetLastError();  // clear the sticky-less error state for the next call
}
}  // namespace r3d

extern "C" int r3d_abi_version(void) { return R3D_ABI_VERSION; }

extern "C" unsigned long long r3d_launch_count(void) { return r3d::g_launches.load(); }

extern "C" const char* r3d_last_cuda_error(void) { return r3d::tls_cuda_error; }

extern "C" const char* r3d_error_string(int code) {
    switch (code) {
        case R3D_OK: return "ok";
        case R3D_EINVAL: return "invalid argument";
        case R3D_ENOT_ENOUGH: return "Not enough points in support to find the requested neighboors";
        case R3D_EKMAX: return "K exceeds the compiled maximum";
        case R3D_EALIGN: return "pointer not aligned as documented";
        case R3D_EWORKSPACE: return "workspace too small";
        case R3D_ECUDA: return "CUDA runtime error";
        case R3D_EUNSUPPORTED: return "shape not supported by the sm_100a kernels";
        default: return "unknown error code";
    }
}
