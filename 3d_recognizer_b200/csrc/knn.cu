// Exact brute-force K-nearest-neighbour search for sm_100a.
//
// Replaces the reference's native extension knn_tpk.knn (randlanet/utils/src/knn.cpp:43-61,
// neighbors.h:281-322, nanoflann.hpp:1367) and the sqrt of KNN.forward (randlanet/utils/modules.py:149).
//
// Contract (include/r3d_b200.h): d2 = fl(fl(fl(dx*dx)+fl(dy*dy))+fl(dz*dz)), d = query - support
// (nanoflann.hpp:488-497 built without FMA), neighbours ordered by (d2, index) ascending.
//
// Design
//   * xyz_to_soa_kernel  packs (B,N,3) into three padded rows per cloud (x.., y.., z..), so a tile of
//     support points is three contiguous, 16-byte aligned segments — the shape a 1-D TMA bulk copy
//     (cp.async.bulk, SASS UBLKCP) wants.  Pad points carry +inf and can never be admitted.
//   * knn_kernel: one thread owns Q queries and scans ALL support points in ascending index order.
//     Support tiles (kTile points) are staged in shared memory by a double-buffered TMA bulk copy
//     signalled through mbarriers; every lane reads the same support point (smem broadcast), so one
//     LDS.128 feeds 4 points x Q queries.
//   * The running K-best of every query lives in a per-thread column of shared memory
//     ([k][q][thread], conflict free); only the admission threshold stays in a register.  Admission
//     is strict '<' against the current K-th d2 and insertion is stable, which yields the (d2, index)
//     order without comparing indices (the scan is ascending).
//   * FP32-pipe budget.  The hot loop is issue bound, so it evaluates a CHEAP d2 with FMAs
//     (3 sub + 1 mul + 2 fma) — packed two points per instruction with the Blackwell f32x2 ops in
//     variant 2 — takes the min over 4 points (FMNMX3) and tests it against a threshold inflated by
//     2^-20 relative (the FMA form differs from the contract form by < 3 ulp).  Only groups that pass
//     (rare: K*ln(N/K) admissions per query) re-evaluate the contract d2 with separate roundings
//     (__fmul_rn/__fadd_rn, never contracted) and run the exact admission test.  The prefilter is
//     conservative, so results are bit-identical to variant 0, which evaluates the contract form
//     for every pair.
#include "common.cuh"

#include <atomic>
#include <cstdlib>
#include <math_constants.h>

namespace r3d {

constexpr int kKnnThreads = 256;
constexpr int kTile = 1024;  // support points per smem stage (3 rows x 4 KB)

static std::atomic<int> g_knn_variant{3};   // 3 = dot-form prefilter + warp-cooperative admission (K = 16, 32), else 2
static std::atomic<int> g_knn_algorithm{0};  // 0 auto, 1 tiled brute force, 2 uniform grid

// knn_grid.cu
size_t knn_grid_workspace_bytes(int B, int Ns, int Nq);
void knn_grid_set_density(float v);
void knn_grid_set_walk(int thread_per_query);
int knn_grid_run(const float* support, long long s_stride, const float* query, long long q_stride, int B, int Ns,
                 int Nq, int K, int64_t* idx64, int32_t* idx32, float* dist, float* dist_sq, void* workspace,
                 cudaStream_t st);

// clouds at least this large go to the grid search in auto mode (measured crossover: profiles/r01_knn_sweep.log)
constexpr int kGridMinSupport = 2048;

// ------------------------------------------------------------------------------------------- pack
__global__ void xyz_to_soa_kernel(const float* __restrict__ xyz, long long batch_stride, float* __restrict__ soa,
                                  int N, int Np) {
    const int b = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= Np) return;
    float x = CUDART_INF_F, y = CUDART_INF_F, z = CUDART_INF_F;
    if (i < N) {
        const float* p = xyz + (size_t)b * batch_stride + (size_t)i * 3;
        x = p[0];
        y = p[1];
        z = p[2];
    }
    float* row = soa + (size_t)b * 3 * Np;
    row[i] = x;
    row[Np + i] = y;
    row[2 * Np + i] = z;
}

// --------------------------------------------------------------------------------- packed helpers
__device__ __forceinline__ uint64_t pk2(float lo, float hi) {
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void upk2(uint64_t v, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t sub2(uint64_t a, uint64_t b) {
    uint64_t r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) {
    uint64_t r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ float min3(float a, float b, float c) {
    float r;
    asm("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
}

// the contract distance: three separate roundings, never contracted into FMAs
__device__ __forceinline__ float d2_contract(float qx, float qy, float qz, float sx, float sy, float sz) {
    const float dx = __fsub_rn(qx, sx), dy = __fsub_rn(qy, sy), dz = __fsub_rn(qz, sz);
    return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}
// cheap form for the prefilter (differs from the contract form by < 3 ulp; all terms >= 0)
__device__ __forceinline__ float d2_cheap(float qx, float qy, float qz, float sx, float sy, float sz) {
    const float dx = qx - sx, dy = qy - sy, dz = qz - sz;
    return fmaf(dz, dz, fmaf(dy, dy, dx * dx));
}
__device__ __forceinline__ float inflate(float thr) {
    // > 3 ulp relative plus an absolute term that covers the subnormal range
    return fmaf(thr, 0x1p-20f, thr) + 1e-40f;
}

// stable insertion into an ascending column of shared memory; returns the new K-th d2
__device__ __noinline__ float list_insert(float* ld, int* li, int K, int stride, float d, int id) {
    int pos = K - 1;
    while (pos > 0) {
        const float pd = ld[(pos - 1) * stride];
        if (!(pd > d)) break;
        ld[pos * stride] = pd;
        li[pos * stride] = li[(pos - 1) * stride];
        --pos;
    }
    ld[pos * stride] = d;
    li[pos * stride] = id;
    return ld[(K - 1) * stride];
}

// Same insertion without a loop-carried shared-memory dependency, for a compile-time list length KT: all KT
// entries are read with independent loads, the insertion point is a count, and the shifted tail is written back
// with predicated stores.  (ncu of the loop version: the dependent LDS chain of up to K iterations, paid by the
// whole warp for its slowest lane, dominated small clouds — 206 us for 8 x 625 points.)
template <int KT>
__device__ __noinline__ float list_insert_t(float* ld, int* li, int stride, float d, int id) {
    float vd[KT];
    int vi[KT];
#pragma unroll
    for (int k = 0; k < KT; ++k) {
        vd[k] = ld[k * stride];
        vi[k] = li[k * stride];
    }
    int pos = 0;
#pragma unroll
    for (int k = 0; k < KT; ++k) pos += (vd[k] <= d) ? 1 : 0;   // stable: an equal d2 keeps the earlier (lower) index first
#pragma unroll
    for (int k = KT - 1; k >= 1; --k) {
        if (k > pos) {
            ld[k * stride] = vd[k - 1];
            li[k * stride] = vi[k - 1];
        }
    }
    ld[pos * stride] = d;     // pos <= KT-1 because the caller admitted d < vd[KT-1]
    li[pos * stride] = id;
    return (pos == KT - 1) ? d : vd[KT - 2];
}

// ------------------------------------------------------------------------------------- the kernel
// VARIANT 0: contract d2 for every pair.  1: FMA prefilter, scalar.  2: FMA prefilter, packed f32x2.
// K1: K == 1, best candidate kept in registers (decoder / post-process 1-NN, modules.py:358).
// KT: compile-time K (16 or 32) for the loop-free insertion, 0 = any K (loop version).
template <int VARIANT, int Q, bool K1, int KT = 0>
__global__ void __launch_bounds__(kKnnThreads, 2) knn_kernel(const float* __restrict__ sup_soa, int Nsp,
                                                          const float* __restrict__ query, long long q_stride,
                                                          int Ns, int Nq, int K,
                                                          int64_t* __restrict__ idx64, int32_t* __restrict__ idx32,
                                                          float* __restrict__ dist, float* __restrict__ dist_sq) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* tile = reinterpret_cast<float*>(smem_raw);                         // [2][3][kTile]
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + 2 * 3 * kTile * sizeof(float));
    float* list_d = reinterpret_cast<float*>(smem_raw + 2 * 3 * kTile * sizeof(float) + 16);
    int* list_i = reinterpret_cast<int*>(list_d + (size_t)(K1 ? 0 : K) * Q * kKnnThreads);

    const int tid = threadIdx.x;
    const int b = blockIdx.y;
    const float* sup = sup_soa + (size_t)b * 3 * Nsp;
    const int num_tiles = (Ns + kTile - 1) / kTile;
    constexpr int stride = Q * kKnnThreads;

    auto issue = [&](int t) {
        const int s = t & 1;
        const int n = min(kTile, Ns - t * kTile);
        const uint32_t bytes = (uint32_t)((n + 3) & ~3) * sizeof(float);
        mbar_expect_tx(&bars[s], 3 * bytes);
#pragma unroll
        for (int c = 0; c < 3; ++c)
            tma_bulk_g2s(tile + (s * 3 + c) * kTile, sup + (size_t)c * Nsp + (size_t)t * kTile, bytes, &bars[s]);
    };

    if (tid == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        mbar_fence_init();
    }
    __syncthreads();
    if (tid == 0) {
        issue(0);
        if (num_tiles > 1) issue(1);
    }

    // ---- per-thread query state
    float qx[Q], qy[Q], qz[Q], thr[Q], thr_hi[Q];
    int best_i[Q];
#pragma unroll
    for (int q = 0; q < Q; ++q) {
        int qi = blockIdx.x * stride + q * kKnnThreads + tid;
        qi = min(qi, Nq - 1);
        const float* p = query + (size_t)b * q_stride + (size_t)qi * 3;
        qx[q] = p[0];
        qy[q] = p[1];
        qz[q] = p[2];
        thr[q] = CUDART_INF_F;
        thr_hi[q] = CUDART_INF_F;
        best_i[q] = -1;
        if (!K1) {
            for (int k = 0; k < K; ++k) {
                list_d[k * stride + q * kKnnThreads + tid] = CUDART_INF_F;
                list_i[k * stride + q * kKnnThreads + tid] = -1;
            }
        }
    }

    for (int t = 0; t < num_tiles; ++t) {
        const int s = t & 1;
        mbar_wait(&bars[s], (t >> 1) & 1);
        const float* xs = tile + (s * 3 + 0) * kTile;
        const float* ys = tile + (s * 3 + 1) * kTile;
        const float* zs = tile + (s * 3 + 2) * kTile;
        const int n = min(kTile, Ns - t * kTile);
        const int n4 = (n + 3) & ~3;
        const int base = t * kTile;

#pragma unroll 2
        for (int j = 0; j < n4; j += 4) {
            const float4 x4 = *reinterpret_cast<const float4*>(xs + j);
            const float4 y4 = *reinterpret_cast<const float4*>(ys + j);
            const float4 z4 = *reinterpret_cast<const float4*>(zs + j);
#pragma unroll
            for (int q = 0; q < Q; ++q) {
                float m;
                if (VARIANT == 2) {
                    const uint64_t qxx = pk2(qx[q], qx[q]), qyy = pk2(qy[q], qy[q]), qzz = pk2(qz[q], qz[q]);
                    const uint64_t dx01 = sub2(qxx, pk2(x4.x, x4.y)), dx23 = sub2(qxx, pk2(x4.z, x4.w));
                    const uint64_t dy01 = sub2(qyy, pk2(y4.x, y4.y)), dy23 = sub2(qyy, pk2(y4.z, y4.w));
                    const uint64_t dz01 = sub2(qzz, pk2(z4.x, z4.y)), dz23 = sub2(qzz, pk2(z4.z, z4.w));
                    const uint64_t a01 = fma2(dz01, dz01, fma2(dy01, dy01, mul2(dx01, dx01)));
                    const uint64_t a23 = fma2(dz23, dz23, fma2(dy23, dy23, mul2(dx23, dx23)));
                    float a0, a1, a2, a3;
                    upk2(a01, a0, a1);
                    upk2(a23, a2, a3);
                    m = min3(fminf(a0, a1), a2, a3);
                } else if (VARIANT == 1) {
                    const float a0 = d2_cheap(qx[q], qy[q], qz[q], x4.x, y4.x, z4.x);
                    const float a1 = d2_cheap(qx[q], qy[q], qz[q], x4.y, y4.y, z4.y);
                    const float a2 = d2_cheap(qx[q], qy[q], qz[q], x4.z, y4.z, z4.z);
                    const float a3 = d2_cheap(qx[q], qy[q], qz[q], x4.w, y4.w, z4.w);
                    m = min3(fminf(a0, a1), a2, a3);
                } else {
                    const float a0 = d2_contract(qx[q], qy[q], qz[q], x4.x, y4.x, z4.x);
                    const float a1 = d2_contract(qx[q], qy[q], qz[q], x4.y, y4.y, z4.y);
                    const float a2 = d2_contract(qx[q], qy[q], qz[q], x4.z, y4.z, z4.z);
                    const float a3 = d2_contract(qx[q], qy[q], qz[q], x4.w, y4.w, z4.w);
                    m = min3(fminf(a0, a1), a2, a3);
                }
                if (m < thr_hi[q]) {
                    // rare path: exact admission, ascending index order inside the group
                    for (int u = 0; u < 4; ++u) {
                        const float d = d2_contract(qx[q], qy[q], qz[q], xs[j + u], ys[j + u], zs[j + u]);
                        if (d < thr[q]) {
                            if (K1) {
                                thr[q] = d;
                                best_i[q] = base + j + u;
                            } else {
                                if (KT > 1)
                                    thr[q] = list_insert_t<(KT > 1 ? KT : 2)>(list_d + q * kKnnThreads + tid,
                                                                              list_i + q * kKnnThreads + tid, stride, d,
                                                                              base + j + u);
                                else
                                    thr[q] = list_insert(list_d + q * kKnnThreads + tid,
                                                         list_i + q * kKnnThreads + tid, K, stride, d, base + j + u);
                            }
                            thr_hi[q] = (VARIANT == 0) ? thr[q] : inflate(thr[q]);
                        }
                    }
                }
            }
        }
        __syncthreads();  // everyone is done with stage s
        if (tid == 0 && t + 2 < num_tiles) issue(t + 2);
    }

    // ---- write-out
#pragma unroll
    for (int q = 0; q < Q; ++q) {
        const int qi = blockIdx.x * stride + q * kKnnThreads + tid;
        if (qi >= Nq) continue;
        const size_t o = ((size_t)b * Nq + qi) * K;
        for (int k = 0; k < K; ++k) {
            const float d = K1 ? thr[q] : list_d[k * stride + q * kKnnThreads + tid];
            const int id = K1 ? best_i[q] : list_i[k * stride + q * kKnnThreads + tid];
            if (idx64) idx64[o + k] = id;
            if (idx32) idx32[o + k] = id;
            if (dist) dist[o + k] = __fsqrt_rn(d);
            if (dist_sq) dist_sq[o + k] = d;
        }
    }
}

// ------------------------------------------------------------------------ VARIANT 3: dot-form prefilter
// The prefilter only has to be CONSERVATIVE, not accurate, so it can use the cheapest form there is:
//     |q - s|^2 - |q|^2 = |s|^2 - 2 q.s            3 FMAs per pair (6 FP32-pipe operations in the difference form)
// on coordinates taken relative to the cloud's first point (so that |q|, |s| are of the size of the cloud, not of its
// distance from the origin).  The pack kernel stores (-2 s~, |s~|^2) as four rows; a thread compares the minimum over 4
// points with  tq = thr (1 + 2^-20) + 2^-18 (|q~|^2 + max|s~|^2) - |q~|^2, which bounds every rounding error of the form
// (centring 2^-24 relative per coordinate, |s~|^2, |q~|^2 and the FMA chain < 2^-19 (|q~|^2 + |s~|^2) together; derivation
// in DESIGN.md §4.1).  Survivors are re-evaluated with the contract arithmetic on the ORIGINAL coordinates, so the
// results are bit-identical to variant 0.
// The admission itself is WARP-COOPERATIVE: a per-lane branch made the whole warp wait for one lane's insertion chain
// (ncu: 21 of 32 lanes active on average).  Here the warp ballots the prefilter and serves the hit lanes one after the
// other with all 32 lanes: lanes 0-3 evaluate the contract d2 of the group's four points, and the insertion into the hit
// lane's list (row-major [query][K+1] in shared memory: conflict-free for both the owner and the warp) is one
// compare + ballot + shifted store.
__global__ void xyz_to_dot_kernel(const float* __restrict__ xyz, long long batch_stride, float* __restrict__ soa,
                                  float* __restrict__ s2max, int N, int Np) {
    const int b = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const float* cloud = xyz + (size_t)b * batch_stride;
    const float cx = cloud[0], cy = cloud[1], cz = cloud[2];
    float x = 0.f, y = 0.f, z = 0.f, w = CUDART_INF_F;
    if (i < N) {
        const float* p = cloud + (size_t)i * 3;
        const float sx = p[0] - cx, sy = p[1] - cy, sz = p[2] - cz;
        x = -2.f * sx, y = -2.f * sy, z = -2.f * sz;
        w = fminf(fmaf(sz, sz, fmaf(sy, sy, sx * sx)), 3.0e38f);
    }
    if (i < Np) {
        float* row = soa + (size_t)b * 4 * Np;
        row[i] = x, row[Np + i] = y, row[2 * Np + i] = z, row[3 * Np + i] = w;
    }
    float m = i < N ? w : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(reinterpret_cast<int*>(s2max + b), __float_as_int(m));
}

template <int Q, int KT, int NT>
__global__ void __launch_bounds__(NT, 512 / NT)
    knn_dot_kernel(const float* __restrict__ soa, int Nsp, const float* __restrict__ s2max,
                   const float* __restrict__ support, long long s_stride, const float* __restrict__ query,
                   long long q_stride, int Ns, int Nq, int64_t* __restrict__ idx64, int32_t* __restrict__ idx32,
                   float* __restrict__ dist, float* __restrict__ dist_sq) {
    constexpr int KP = KT + 1;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* tile = reinterpret_cast<float*>(smem_raw);                         // [2][4][kTile]
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + 2 * 4 * kTile * sizeof(float));
    float* list_d = reinterpret_cast<float*>(smem_raw + 2 * 4 * kTile * sizeof(float) + 16);   // [Q][threads][KP]
    int* list_i = reinterpret_cast<int*>(list_d + (size_t)KP * Q * NT);

    const int tid = threadIdx.x, lane = tid & 31;
    const int b = blockIdx.y;
    const float* rows = soa + (size_t)b * 4 * Nsp;
    const float* sup = support + (size_t)b * s_stride;
    const int num_tiles = (Ns + kTile - 1) / kTile;
    constexpr int stride = Q * NT;

    auto issue = [&](int t) {
        const int s = t & 1;
        const int n = min(kTile, Ns - t * kTile);
        const uint32_t bytes = (uint32_t)((n + 3) & ~3) * sizeof(float);
        mbar_expect_tx(&bars[s], 4 * bytes);
#pragma unroll
        for (int c = 0; c < 4; ++c)
            tma_bulk_g2s(tile + (s * 4 + c) * kTile, rows + (size_t)c * Nsp + (size_t)t * kTile, bytes, &bars[s]);
    };
    if (tid == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        mbar_fence_init();
    }
    __syncthreads();
    if (tid == 0) {
        issue(0);
        if (num_tiles > 1) issue(1);
    }

    const float cx = sup[0], cy = sup[1], cz = sup[2];
    const float smax = s2max[b];
    float qx[Q], qy[Q], qz[Q], thr[Q], tq[Q], off[Q];      // off = 2^-18 (|q~|^2 + max|s~|^2) - |q~|^2
    uint64_t qxx[Q], qyy[Q], qzz[Q];
#pragma unroll
    for (int q = 0; q < Q; ++q) {
        int qi = blockIdx.x * stride + q * NT + tid;
        qi = min(qi, Nq - 1);
        const float* p = query + (size_t)b * q_stride + (size_t)qi * 3;
        qx[q] = p[0], qy[q] = p[1], qz[q] = p[2];
        const float ux = qx[q] - cx, uy = qy[q] - cy, uz = qz[q] - cz;
        const float q2 = fmaf(uz, uz, fmaf(uy, uy, ux * ux));
        off[q] = fmaf(0x1p-18f, q2 + smax, -q2);
        qxx[q] = pk2(ux, ux), qyy[q] = pk2(uy, uy), qzz[q] = pk2(uz, uz);
        thr[q] = CUDART_INF_F;
        tq[q] = CUDART_INF_F;
        float* cd = list_d + (size_t)(q * NT + tid) * KP;
        int* ci = list_i + (size_t)(q * NT + tid) * KP;
        for (int k = 0; k < KT; ++k) cd[k] = CUDART_INF_F, ci[k] = -1;
    }
    __syncwarp();

    for (int t = 0; t < num_tiles; ++t) {
        const int s = t & 1;
        mbar_wait(&bars[s], (t >> 1) & 1);
        const float* xs = tile + (s * 4 + 0) * kTile;
        const float* ys = tile + (s * 4 + 1) * kTile;
        const float* zs = tile + (s * 4 + 2) * kTile;
        const float* ws = tile + (s * 4 + 3) * kTile;
        const int n = min(kTile, Ns - t * kTile);
        const int n4 = (n + 3) & ~3;
        const int base = t * kTile;

        // G groups of 4 points x Q queries per iteration in ONE straight-line block — 2 G Q = 8 independent FMA chains —
        // followed by a single warp vote: a vote + branch per group and query serialised the chains (measured: 40 % of
        // the issue slots used, no faster than the difference form)
        constexpr int G = Q == 1 ? 4 : 2;
        for (int j = 0; j < n4; j += 4 * G) {
            float m[Q][G];
            bool any = false;
#pragma unroll
            for (int g = 0; g < G; ++g) {
                const float4 x4 = *reinterpret_cast<const float4*>(xs + j + 4 * g);
                const float4 y4 = *reinterpret_cast<const float4*>(ys + j + 4 * g);
                const float4 z4 = *reinterpret_cast<const float4*>(zs + j + 4 * g);
                float4 w4 = *reinterpret_cast<const float4*>(ws + j + 4 * g);
                if (g >= 1 && j + 4 * g >= n4) w4 = make_float4(CUDART_INF_F, CUDART_INF_F, CUDART_INF_F, CUDART_INF_F);
                const uint64_t x01 = pk2(x4.x, x4.y), x23 = pk2(x4.z, x4.w), y01 = pk2(y4.x, y4.y), y23 = pk2(y4.z, y4.w);
                const uint64_t z01 = pk2(z4.x, z4.y), z23 = pk2(z4.z, z4.w), w01 = pk2(w4.x, w4.y), w23 = pk2(w4.z, w4.w);
#pragma unroll
                for (int q = 0; q < Q; ++q) {
                    const uint64_t a01 = fma2(qzz[q], z01, fma2(qyy[q], y01, fma2(qxx[q], x01, w01)));
                    const uint64_t a23 = fma2(qzz[q], z23, fma2(qyy[q], y23, fma2(qxx[q], x23, w23)));
                    float a0, a1, a2, a3;
                    upk2(a01, a0, a1);
                    upk2(a23, a2, a3);
                    m[q][g] = min3(fminf(a0, a1), a2, a3);
                    any |= m[q][g] < tq[q];
                }
            }
            if (!__any_sync(0xffffffffu, any)) continue;
#pragma unroll
            for (int g = 0; g < G; ++g) {
#pragma unroll
                for (int q = 0; q < Q; ++q) {
                    unsigned hits = __ballot_sync(0xffffffffu, m[q][g] < tq[q]);
                    if (!hits) continue;
                    const int jg = j + 4 * g;
                    // contract d2 of the group's four points, for whichever lane is being served: lanes 0-3 hold the points
                    float px = 0.f, py = 0.f, pz = 0.f;
                    const int cand = base + jg + (lane & 3);
                    const bool cand_ok = cand < Ns;
                    if (lane < 4 && cand_ok) {
                        const float* p = sup + (size_t)cand * 3;
                        px = p[0], py = p[1], pz = p[2];
                    }
                    while (hits) {
                        const int L = __ffs(hits) - 1;
                        hits &= hits - 1;
                        const float lx = __shfl_sync(0xffffffffu, qx[q], L), ly = __shfl_sync(0xffffffffu, qy[q], L);
                        const float lz = __shfl_sync(0xffffffffu, qz[q], L);
                        float lthr = __shfl_sync(0xffffffffu, thr[q], L);
                        const float dc = (lane < 4 && cand_ok) ? d2_contract(lx, ly, lz, px, py, pz) : CUDART_INF_F;
                        float* cd = list_d + (size_t)(q * NT + (tid - lane + L)) * KP;
                        int* ci = list_i + (size_t)(q * NT + (tid - lane + L)) * KP;
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const float d = __shfl_sync(0xffffffffu, dc, u);
                            if (d < lthr) {                                   // warp-uniform
                                const float vd = lane < KT ? cd[lane] : CUDART_INF_F;
                                const int vi = lane < KT ? ci[lane] : -1;
                                // the list is ascending: the entries <= d are a prefix (an equal d2 keeps the lower index first)
                                const int pos = __popc(__ballot_sync(0xffffffffu, lane < KT && vd <= d));
                                __syncwarp();
                                if (lane >= pos && lane + 1 < KT) cd[lane + 1] = vd, ci[lane + 1] = vi;
                                if (lane == pos) cd[pos] = d, ci[pos] = base + jg + u;
                                const float prev = __shfl_sync(0xffffffffu, vd, KT >= 2 ? KT - 2 : 0);
                                lthr = (pos == KT - 1) ? d : prev;
                                __syncwarp();
                            }
                        }
                        if (lane == L) {
                            thr[q] = lthr;
                            tq[q] = fmaf(lthr, 0x1p-20f, lthr) + off[q];
                            if (!(tq[q] == tq[q])) tq[q] = CUDART_INF_F;   // inf - inf: keep everything on the exact path
                        }
                    }
                }
            }
        }
        __syncthreads();  // everyone is done with stage s
        if (tid == 0 && t + 2 < num_tiles) issue(t + 2);
    }

#pragma unroll
    for (int q = 0; q < Q; ++q) {
        const int qi = blockIdx.x * stride + q * NT + tid;
        if (qi >= Nq) continue;
        const size_t o = ((size_t)b * Nq + qi) * KT;
        const float* cd = list_d + (size_t)(q * NT + tid) * KP;
        const int* ci = list_i + (size_t)(q * NT + tid) * KP;
        for (int k = 0; k < KT; ++k) {
            const float d = cd[k];
            const int id = ci[k];
            if (idx64) idx64[o + k] = id;
            if (idx32) idx32[o + k] = id;
            if (dist) dist[o + k] = __fsqrt_rn(d);
            if (dist_sq) dist_sq[o + k] = d;
        }
    }
}

// ------------------------------------------------------------------------- small clouds (Ns < 2048)
// The down-sampled levels and the decoder of a 2 500-point cloud search 39..625 support points for a few thousand
// queries: one thread per query leaves most SMs idle and every thread walks the whole support with a dependent
// insertion chain (96 us per launch in the training step's launch list).  Here a WARP owns a query: lane l computes
// the contract d2 of support points l, l+32, l+64, .. once and keeps them in registers as 64-bit keys
// (d2 bits << 32 | index; d2 >= 0, so the unsigned order of the keys IS the (d2, index) order of the contract), and
// the K nearest are extracted by K rounds of "smallest key above the last one": a branch-free scan of the lane's
// registers and two warp redux.min.u32 (d2 bits, then index among the lanes that hold that d2).  No insertion,
// no divergence, results bit-identical to the other back-ends.
constexpr int kSmallThreads = 256;
constexpr int kSmallMaxNs = 2047;

template <int C>
__global__ void __launch_bounds__(kSmallThreads) knn_small_kernel(const float* __restrict__ support, long long s_stride,
                                                                  const float* __restrict__ query, long long q_stride,
                                                                  int Ns, int Nq, int K, int qpw,
                                                                  int64_t* __restrict__ idx64, int32_t* __restrict__ idx32,
                                                                  float* __restrict__ dist, float* __restrict__ dist_sq) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* sx = reinterpret_cast<float*>(smem_raw);            // [3][C*32]
    constexpr int NP = C * 32;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, b = blockIdx.y;
    const float* sup = support + (size_t)b * s_stride;
    for (int i = tid; i < Ns; i += kSmallThreads) {
        sx[i] = sup[(size_t)i * 3 + 0];
        sx[NP + i] = sup[(size_t)i * 3 + 1];
        sx[2 * NP + i] = sup[(size_t)i * 3 + 2];
    }
    __syncthreads();
    const int q0 = (blockIdx.x * (kSmallThreads / 32) + warp) * qpw;
    for (int qi = q0; qi < min(q0 + qpw, Nq); ++qi) {
        const float* qp = query + (size_t)b * q_stride + (size_t)qi * 3;
        const float qx = qp[0], qy = qp[1], qz = qp[2];
        unsigned long long key[C];
#pragma unroll
        for (int c = 0; c < C; ++c) {
            const int j = c * 32 + lane;
            const float d = d2_contract(qx, qy, qz, sx[j], sx[NP + j], sx[2 * NP + j]);
            key[c] = j < Ns ? ((unsigned long long)__float_as_uint(d) << 32) | (unsigned)j : ~0ull;
        }
        unsigned long long lo = 0, mine = 0;
        const size_t o = ((size_t)b * Nq + qi) * K;
        for (int k = 0; k < K; ++k) {
            unsigned long long best = ~0ull;
#pragma unroll
            for (int c = 0; c < C; ++c)
                if (key[c] >= lo && key[c] < best) best = key[c];
            const unsigned hi = (unsigned)(best >> 32);
            const unsigned mhi = __reduce_min_sync(0xffffffffu, hi);
            const unsigned mlo = __reduce_min_sync(0xffffffffu, hi == mhi ? (unsigned)best : 0xffffffffu);
            const unsigned long long m = ((unsigned long long)mhi << 32) | mlo;
            lo = m + 1;
            if ((k & 31) == lane) mine = m;
            if ((k & 31) == 31 || k == K - 1) {       // coalesced store of up to 32 results
                const int kk = (k & ~31) + lane;
                if (kk <= k) {
                    const float d = __uint_as_float((unsigned)(mine >> 32));
                    const int id = (int)(unsigned)mine;
                    if (idx64) idx64[o + kk] = id;
                    if (idx32) idx32[o + kk] = id;
                    if (dist) dist[o + kk] = __fsqrt_rn(d);
                    if (dist_sq) dist_sq[o + kk] = d;
                }
            }
        }
    }
}

template <int C>
static int launch_knn_small_c(const float* support, long long s_stride, const float* query, long long q_stride, int B,
                              int Ns, int Nq, int K, int64_t* idx64, int32_t* idx32, float* dist, float* dist_sq,
                              cudaStream_t st) {
    const size_t smem = (size_t)3 * C * 32 * sizeof(float);
    // queries per warp: enough CTAs for ~4 per SM, at most 8 queries behind one staging of the cloud per warp
    long long qpw = ((long long)B * Nq) / (8LL * 148 * 4);
    qpw = qpw < 1 ? 1 : (qpw > 8 ? 8 : qpw);
    dim3 grid(ceil_div(Nq, (kSmallThreads / 32) * (int)qpw), B);
    knn_small_kernel<C><<<grid, kSmallThreads, smem, st>>>(support, s_stride, query, q_stride, Ns, Nq, K, (int)qpw, idx64,
                                                           idx32, dist, dist_sq);
    R3D_LAUNCH_CHECK("knn_small_kernel");
    return R3D_OK;
}

static int launch_knn_small(const float* support, long long s_stride, const float* query, long long q_stride, int B,
                            int Ns, int Nq, int K, int64_t* idx64, int32_t* idx32, float* dist, float* dist_sq,
                            cudaStream_t st) {
#define R3D_SMALL(CC)                                                                                                   \
    if (Ns <= CC * 32)                                                                                                  \
        return launch_knn_small_c<CC>(support, s_stride, query, q_stride, B, Ns, Nq, K, idx64, idx32, dist, dist_sq, st);
    R3D_SMALL(2)
    R3D_SMALL(5)
    R3D_SMALL(10)
    R3D_SMALL(20)
    R3D_SMALL(32)
    R3D_SMALL(64)
#undef R3D_SMALL
    return R3D_EUNSUPPORTED;
}

static size_t knn_smem_bytes(int K, int Q, bool k1) {
    return 2 * 3 * kTile * sizeof(float) + 16 + (k1 ? 0 : (size_t)K * Q * kKnnThreads * 8);
}

template <int VARIANT, int Q, bool K1, int KT = 0>
static int launch_knn(const float* sup_soa, int Nsp, const float* query, long long q_stride, int B, int Ns, int Nq,
                      int K, int64_t* idx64, int32_t* idx32, float* dist, float* dist_sq, cudaStream_t st) {
    auto kern = knn_kernel<VARIANT, Q, K1, KT>;
    const size_t smem = knn_smem_bytes(K, Q, K1);
    R3D_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid(ceil_div(Nq, Q * kKnnThreads), B);
    kern<<<grid, kKnnThreads, smem, st>>>(sup_soa, Nsp, query, q_stride, Ns, Nq, K, idx64, idx32, dist, dist_sq);
    R3D_LAUNCH_CHECK("knn_kernel");
    return R3D_OK;
}

template <int VARIANT>
static int dispatch_q(int Q, bool k1, const float* sup_soa, int Nsp, const float* query, long long q_stride, int B,
                      int Ns, int Nq, int K, int64_t* idx64, int32_t* idx32, float* dist, float* dist_sq,
                      cudaStream_t st) {
#define R3D_KNN_ARGS sup_soa, Nsp, query, q_stride, B, Ns, Nq, K, idx64, idx32, dist, dist_sq, st
#define R3D_KNN_CASE(QQ)                                                              \
    if (Q == QQ) {                                                                    \
        if (k1) return launch_knn<VARIANT, QQ, true>(R3D_KNN_ARGS);                   \
        if (K == 16) return launch_knn<VARIANT, QQ, false, 16>(R3D_KNN_ARGS);         \
        if (K == 32 && QQ == 1) return launch_knn<VARIANT, QQ, false, 32>(R3D_KNN_ARGS); \
        return launch_knn<VARIANT, QQ, false>(R3D_KNN_ARGS);                          \
    }
    R3D_KNN_CASE(1)
    R3D_KNN_CASE(2)
    R3D_KNN_CASE(4)
#undef R3D_KNN_CASE
#undef R3D_KNN_ARGS
    return R3D_EINVAL;
}

template <int Q, int KT, int NT = kKnnThreads>
static int launch_knn_dot(const float* soa, int Nsp, const float* s2max, const float* support, long long s_stride,
                          const float* query, long long q_stride, int B, int Ns, int Nq, int64_t* idx64, int32_t* idx32,
                          float* dist, float* dist_sq, cudaStream_t st) {
    auto kern = knn_dot_kernel<Q, KT, NT>;
    const size_t smem = 2 * 4 * kTile * sizeof(float) + 16 + (size_t)(KT + 1) * Q * NT * 8;
    R3D_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid(ceil_div(Nq, Q * NT), B);
    kern<<<grid, NT, smem, st>>>(soa, Nsp, s2max, support, s_stride, query, q_stride, Ns, Nq, idx64, idx32, dist,
                                          dist_sq);
    R3D_LAUNCH_CHECK("knn_dot_kernel");
    return R3D_OK;
}

static int pick_q(int B, int Nq, int K) {
    // two CTAs per SM (16 warps) hide the FP32-pipe and shared-memory latencies: K*Q*256*8 B of lists
    // + 24 KB of tiles must fit half of the 227 KB
    int qmax = (K <= 8) ? 4 : (K <= 16) ? 2 : 1;
    // do not starve the 148 SMs when there are few queries
    const long long total = (long long)B * Nq;
    int q = qmax;
    while (q > 1 && total < (long long)kNumSMs * kKnnThreads * q) q >>= 1;
    return q;
}

}  // namespace r3d

using namespace r3d;

extern "C" size_t r3d_knn_workspace_bytes(int B, int Ns, int Nq, int K) {
    (void)Nq;
    (void)K;
    if (B <= 0 || Ns <= 0) return 256;
    const size_t Nsp = (size_t)((Ns + 3) & ~3);
    const size_t brute = align_up((size_t)B * 4 * Nsp * sizeof(float), 256) + align_up((size_t)B * sizeof(float), 256) + 256;
    const size_t grid = knn_grid_workspace_bytes(B, Ns, Nq > 0 ? Nq : 1) + 256;
    return brute > grid ? brute : grid;
}

extern "C" int r3d_knn_set_variant(int variant) {
    if (variant < 0 || variant > 3) return g_knn_variant.load();
    return g_knn_variant.exchange(variant);
}

// tuning hook for tools/knn_bench.py: average support points per grid cell (0 = built-in default)
extern "C" int r3d_knn_set_grid_density(float points_per_cell) {
    knn_grid_set_density(points_per_cell);
    return R3D_OK;
}

extern "C" int r3d_knn_set_algorithm(int algorithm) {
    if (algorithm < 0 || algorithm > 4 || algorithm == 3) return g_knn_algorithm.load();
    knn_grid_set_walk(algorithm == 4 ? 1 : (algorithm == 2 ? 0 : -1));
    return g_knn_algorithm.exchange(algorithm);
}

// Which back-end r3d_knn runs for a shape under the current algorithm setting: 1 tiled brute force, 2 uniform grid,
// 3 warp-per-query register scan.  Auto: the grid from 2048 support points up; below, the register scan while its
// O(Nq Ns K / 32) work stays under the grid search's fixed cost (measured: 15 ps per query x candidate-per-lane x K
// against ~25 us for the grid build and walk, profiles/r01_knn_small.txt).
extern "C" int r3d_knn_plan(int B, int Ns, int Nq, int K) {
    const int algo = g_knn_algorithm.load();
    if (algo == 1 || algo == 2) return algo;
    if (algo == 4) return 2;
    if (Ns >= kGridMinSupport) return 2;
    const long long c = (Ns + 31) / 32;
    if (c <= 10 || K == 1 || (long long)B * Nq * c * K <= 1500000LL) return 3;
    return 2;
}

extern "C" int r3d_knn(const float* support, long long support_batch_stride, const float* query,
                       long long query_batch_stride, int B, int Ns, int Nq, int K, int64_t* idx64, int32_t* idx32,
                       float* dist, float* dist_sq, void* workspace, size_t workspace_bytes, r3d_stream_t stream) {
    if (B < 0 || Ns < 0 || Nq < 0 || K <= 0) return R3D_EINVAL;
    if (K > R3D_KNN_KMAX) return R3D_EKMAX;
    if (Ns < K) return R3D_ENOT_ENOUGH;  // knn.cpp:15-17
    if (B == 0 || Nq == 0) return R3D_OK;
    if (!support || !query || !workspace) return R3D_EINVAL;
    if (!is_aligned(workspace, 256) || !is_aligned(support, 4) || !is_aligned(query, 4)) return R3D_EALIGN;
    if (workspace_bytes < r3d_knn_workspace_bytes(B, Ns, Nq, K)) return R3D_EWORKSPACE;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (support_batch_stride == 0) support_batch_stride = (long long)Ns * 3;
    if (query_batch_stride == 0) query_batch_stride = (long long)Nq * 3;
    if (support_batch_stride < (long long)Ns * 3 || query_batch_stride < (long long)Nq * 3) return R3D_EINVAL;

    const int plan = r3d_knn_plan(B, Ns, Nq, K);
    if (plan == 2)
        return knn_grid_run(support, support_batch_stride, query, query_batch_stride, B, Ns, Nq, K, idx64, idx32, dist,
                            dist_sq, workspace, st);
    if (plan == 3)
        return launch_knn_small(support, support_batch_stride, query, query_batch_stride, B, Ns, Nq, K, idx64, idx32, dist,
                                dist_sq, st);

    const int Nsp = (Ns + 3) & ~3;
    float* soa = static_cast<float*>(workspace);
    if (g_knn_variant.load() == 3 && (K == 16 || K == 32)) {
        float* s2max = soa + align_up((size_t)B * 4 * Nsp * sizeof(float), 256) / sizeof(float);
        R3D_CUDA_TRY(cudaMemsetAsync(s2max, 0, (size_t)B * sizeof(float), st));
        dim3 grid(ceil_div(Nsp, 256), B);
        xyz_to_dot_kernel<<<grid, 256, 0, st>>>(support, support_batch_stride, soa, s2max, Ns, Nsp);
        R3D_LAUNCH_CHECK("xyz_to_dot_kernel");
        int Q = pick_q(B, Nq, K);
        if (const char* e = getenv("R3D_KNN_DOT_Q")) Q = atoi(e);       // tuning hook (tools/knn_bench.py)
#define R3D_DOT_ARGS soa, Nsp, s2max, support, support_batch_stride, query, query_batch_stride, B, Ns, Nq, idx64, idx32, dist, dist_sq, st
        if (K == 32) {
            if (Q >= 2) return launch_knn_dot<2, 32>(R3D_DOT_ARGS);
            return launch_knn_dot<1, 32>(R3D_DOT_ARGS);
        }
        if (Q >= 4) return launch_knn_dot<4, 16>(R3D_DOT_ARGS);
        if (Q >= 2) return launch_knn_dot<2, 16>(R3D_DOT_ARGS);
        return launch_knn_dot<1, 16>(R3D_DOT_ARGS);
#undef R3D_DOT_ARGS
    }
    {
        dim3 grid(ceil_div(Nsp, 256), B);
        xyz_to_soa_kernel<<<grid, 256, 0, st>>>(support, support_batch_stride, soa, Ns, Nsp);
        R3D_LAUNCH_CHECK("xyz_to_soa_kernel");
    }
    const bool k1 = (K == 1);
    const int Q = k1 ? ((long long)B * Nq >= (long long)kNumSMs * kKnnThreads * 8 ? 4 : pick_q(B, Nq, 8))
                     : pick_q(B, Nq, K);
    switch (g_knn_variant.load()) {
        case 0: return dispatch_q<0>(Q, k1, soa, Nsp, query, query_batch_stride, B, Ns, Nq, K, idx64, idx32, dist,
                                       dist_sq, st);
        case 1: return dispatch_q<1>(Q, k1, soa, Nsp, query, query_batch_stride, B, Ns, Nq, K, idx64, idx32, dist,
                                       dist_sq, st);
        default: return dispatch_q<2>(Q, k1, soa, Nsp, query, query_batch_stride, B, Ns, Nq, K, idx64, idx32, dist,
                                       dist_sq, st);
    }
}

// ------------------------------------------------------------------------------------------ host-buffer drop-in
// Per host thread: two non-blocking streams (search, copy-back), a grow-only device arena and two events.  Nothing here
// touches the legacy default stream, allocates per call or synchronises the device: a search issued from one thread
// does not serialise against a training step running in another (reference: main.py:71-89 predicts on the Tk thread
// while train.py:108-115 trains in a child).
namespace {
struct KnnHostCtx {
    int device = -1;
    cudaStream_t run = nullptr, copy = nullptr;
    cudaEvent_t ready[2] = {nullptr, nullptr};
    unsigned char* arena = nullptr;
    size_t arena_bytes = 0;
    void release() {
        if (device < 0) return;
        // the CUDA context may already be gone at thread exit: errors are ignored
        if (arena) cudaFree(arena);
        if (run) cudaStreamDestroy(run);
        if (copy) cudaStreamDestroy(copy);
        for (auto& e : ready)
            if (e) cudaEventDestroy(e);
        *this = KnnHostCtx();
    }
    KnnHostCtx() = default;
    KnnHostCtx(const KnnHostCtx&) = default;
    KnnHostCtx& operator=(const KnnHostCtx&) = default;
    ~KnnHostCtx() {
        if (device >= 0) {
            if (arena) cudaFree(arena);
            if (run) cudaStreamDestroy(run);
            if (copy) cudaStreamDestroy(copy);
            for (auto& e : ready)
                if (e) cudaEventDestroy(e);
        }
    }
};
thread_local KnnHostCtx g_knn_host;

int knn_host_prepare(KnnHostCtx& c, size_t bytes) {
    int dev = 0;
    R3D_CUDA_TRY(cudaGetDevice(&dev));
    if (c.device != dev) {
        c.release();                  // the thread moved to another device: drop the old resources
        R3D_CUDA_TRY(cudaStreamCreateWithFlags(&c.run, cudaStreamNonBlocking));
        R3D_CUDA_TRY(cudaStreamCreateWithFlags(&c.copy, cudaStreamNonBlocking));
        for (auto& e : c.ready) R3D_CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        c.device = dev;
    }
    if (c.arena_bytes < bytes) {
        if (c.arena) {
            R3D_CUDA_TRY(cudaStreamSynchronize(c.run));
            R3D_CUDA_TRY(cudaStreamSynchronize(c.copy));
            R3D_CUDA_TRY(cudaFree(c.arena));
            c.arena = nullptr;
            c.arena_bytes = 0;
        }
        const size_t want = align_up(bytes + bytes / 4, (size_t)1 << 20);
        R3D_CUDA_TRY(cudaMalloc(&c.arena, want));
        c.arena_bytes = want;
    }
    return R3D_OK;
}
}  // namespace

// Host buffers may be pageable or page-locked.  The search runs in chunks of clouds (or of queries when there are few
// clouds); chunk i's results travel back on the copy stream while chunk i+1 is searched.  Into page-locked outputs
// (what the Python host allocates) that copy is a direct DMA at PCIe rate.
extern "C" int r3d_knn_host(const float* support, const float* query, int B, int Ns, int Nq, int K, int64_t* idx64,
                            float* dist_sq) {
    if (B < 0 || Ns < 0 || Nq < 0 || K <= 0) return R3D_EINVAL;
    if (K > R3D_KNN_KMAX) return R3D_EKMAX;
    if (Ns < K) return R3D_ENOT_ENOUGH;
    if (B == 0 || Nq == 0) return R3D_OK;
    if (!support || !query || !idx64 || !dist_sq) return R3D_EINVAL;
    const size_t sb = (size_t)B * Ns * 3 * sizeof(float), qb = (size_t)B * Nq * 3 * sizeof(float);
    const size_t ib = (size_t)B * Nq * K * sizeof(int64_t), db = (size_t)B * Nq * K * sizeof(float);
    const bool self = (support == query && Ns == Nq);
    // chunking: whole clouds when there are many, query ranges of one cloud otherwise
    const long long total_q = (long long)B * Nq;
    int nchunk = (int)(total_q / 196608);
    nchunk = nchunk < 1 ? 1 : (nchunk > 8 ? 8 : nchunk);
    const bool by_cloud = B >= nchunk * 2 || Nq < 4096;
    const int cb = by_cloud ? ceil_div(B, nchunk) : 1;                 // clouds per chunk
    const int cq = by_cloud ? Nq : ceil_div(Nq, nchunk);               // queries per chunk
    const size_t wb = r3d_knn_workspace_bytes(cb, Ns, cq, K);
    const size_t o_s = 0, o_q = align_up(sb, 256), o_i = o_q + (self ? 0 : align_up(qb, 256)),
                 o_d = o_i + align_up(ib, 256), o_w = o_d + align_up(db, 256);
    KnnHostCtx& c = g_knn_host;
    const int prc = knn_host_prepare(c, o_w + wb);
    if (prc != R3D_OK) return prc;
    unsigned char* dev = c.arena;
    float* d_s = reinterpret_cast<float*>(dev + o_s);
    float* d_q = self ? d_s : reinterpret_cast<float*>(dev + o_q);
    int64_t* d_i = reinterpret_cast<int64_t*>(dev + o_i);
    float* d_d = reinterpret_cast<float*>(dev + o_d);
    int rc = R3D_OK;
    cudaError_t e = cudaMemcpyAsync(d_s, support, sb, cudaMemcpyHostToDevice, c.run);
    if (e == cudaSuccess && !self) e = cudaMemcpyAsync(d_q, query, qb, cudaMemcpyHostToDevice, c.run);
    int n = 0;
    for (int b0 = 0; b0 < B && e == cudaSuccess && rc == R3D_OK; b0 += cb) {
        const int nb = (B - b0 < cb) ? B - b0 : cb;
        for (int q0 = 0; q0 < Nq && e == cudaSuccess && rc == R3D_OK; q0 += cq, ++n) {
            const int nq = (Nq - q0 < cq) ? Nq - q0 : cq;
            const size_t off = ((size_t)b0 * Nq + q0) * K;             // first result of this chunk (contiguous: nb == 1 or nq == Nq)
            const size_t cnt = (size_t)nb * nq * K;
            rc = r3d_knn(d_s + (size_t)b0 * Ns * 3, 0, d_q + ((size_t)b0 * Nq + q0) * 3, by_cloud ? 0 : (long long)Nq * 3,
                         nb, Ns, nq, K, d_i + off, nullptr, nullptr, d_d + off, dev + o_w, wb, c.run);
            if (rc != R3D_OK) break;
            cudaEvent_t ev = c.ready[n & 1];
            e = cudaEventRecord(ev, c.run);
            if (e == cudaSuccess) e = cudaStreamWaitEvent(c.copy, ev, 0);
            if (e == cudaSuccess) e = cudaMemcpyAsync(idx64 + off, d_i + off, cnt * sizeof(int64_t), cudaMemcpyDeviceToHost, c.copy);
            if (e == cudaSuccess) e = cudaMemcpyAsync(dist_sq + off, d_d + off, cnt * sizeof(float), cudaMemcpyDeviceToHost, c.copy);
        }
    }
    // always drain both streams: the arena is reused by the next call
    const cudaError_t e1 = cudaStreamSynchronize(c.run), e2 = cudaStreamSynchronize(c.copy);
    if (e == cudaSuccess) e = (e1 != cudaSuccess) ? e1 : e2;
    if (e != cudaSuccess) {
        set_cuda_error(e, "r3d_knn_host");
        return R3D_ECUDA;
    }
    return rc;
}
