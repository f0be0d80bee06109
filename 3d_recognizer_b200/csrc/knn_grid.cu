// Exact K-nearest-neighbour search over a uniform grid (sm_100a) — the large-cloud back-end of r3d_knn.
//
// Same contract as the tiled brute-force kernel in knn.cu (include/r3d_b200.h): contract-rounded d2,
// neighbours ordered by (d2, index), bit-identical output.  It replaces the reference's KD-tree search
// (randlanet/utils/src/neighbors.h:281-322: nanoflann build + knnSearch per query) the B200 way: instead
// of a pointer-chasing tree, the support cloud is counting-sorted into cubic cells and every query
// thread visits cells in growing Chebyshev rings around its own cell until the K-th distance is provably
// smaller than the distance to anything unvisited.  O(N*K) work instead of the O(N^2) of brute force.
//
// Pipeline (all on the caller's stream, memory from the caller's workspace):
//   grid_bbox_kernel      per cloud: bounding box of the support
//   grid_setup_kernel     per cloud: cell size s and grid dims g (<= cell budget), header
//   grid_count_kernel     cell histogram of the support  (+ the same for the queries when they differ)
//   grid_scan_kernel      exclusive scan of the histogram(s), one CTA per cloud
//   grid_scatter_kernel   support -> float4 {x,y,z,index} in cell order; queries -> visiting order
//   knn_grid_kernel       one thread per query, queries in cell order (neighbouring lanes walk the same
//                         cells: their loads coalesce in L1), running K-best in a shared-memory column
//
// Exactness.  A point whose computed cell is c has true grid coordinate within [c - e, c + 1 + e],
// e < 2.5e-4 (fp32 rounding of (x - lo) * inv_s for g <= 1024).  After ring r every unvisited point is
// outside the block of cells [c-r, c+r]^3, so its true distance to the query is at least the distance
// `bound` from the query to the nearest block face that is not the grid border (faces on the border have
// nothing behind them), minus e*s.  The search stops when the list is full and
// kth_d2 < (0.99999 * (bound - margin))^2; the slack covers the 3-ulp error of the computed d2.
// Insertion is lexicographic on (d2, index) because cells are not visited in index order.
#include "common.cuh"

#include <math_constants.h>

namespace r3d {

struct GridHdr {
    float lo[3];
    float s;        // cell edge
    float inv_s;
    int g[3];       // cells per axis
    float margin;   // absolute slack of the stop test
    int pad[3];
};
static_assert(sizeof(GridHdr) == 48, "GridHdr layout");

constexpr int kGridMaxAxis = 1024;
static float g_pts_per_cell = 0.f;   // tuning hook (r3d_knn_set_grid_density); 0 = default
static int g_grid_walk = -1;         // -1 by batch size, 0 one warp per query, 1 one thread per query (r3d_knn_set_algorithm)

__device__ __forceinline__ float d2_contract_g(float qx, float qy, float qz, float sx, float sy, float sz) {
    const float dx = __fsub_rn(qx, sx), dy = __fsub_rn(qy, sy), dz = __fsub_rn(qz, sz);
    return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

__device__ __forceinline__ int cell_coord(float x, float lo, float inv_s, int g) {
    const int c = (int)floorf((x - lo) * inv_s);
    return min(g - 1, max(0, c));
}

// ----------------------------------------------------------------------------------------- bbox
__global__ void __launch_bounds__(1024) grid_bbox_kernel(const float* __restrict__ xyz, long long bstride, int N,
                                                         float* __restrict__ bbox /* (B,6) */) {
    const int b = blockIdx.x;
    const float* p = xyz + (size_t)b * bstride;
    float lo[3] = {CUDART_INF_F, CUDART_INF_F, CUDART_INF_F}, hi[3] = {-CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F};
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float v = p[(size_t)i * 3 + c];
            lo[c] = fminf(lo[c], v);
            hi[c] = fmaxf(hi[c], v);
        }
    }
    __shared__ float red[32][6];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        for (int o = 16; o > 0; o >>= 1) {
            lo[c] = fminf(lo[c], __shfl_xor_sync(0xffffffffu, lo[c], o));
            hi[c] = fmaxf(hi[c], __shfl_xor_sync(0xffffffffu, hi[c], o));
        }
    }
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            red[w][c] = lo[c];
            red[w][3 + c] = hi[c];
        }
    }
    __syncthreads();
    if (threadIdx.x < 6) {
        const int c = threadIdx.x;
        float v = red[0][c];
        for (int i = 1; i < (int)(blockDim.x >> 5); ++i) v = (c < 3) ? fminf(v, red[i][c]) : fmaxf(v, red[i][c]);
        bbox[b * 6 + c] = v;
    }
}

// ---------------------------------------------------------------------------------------- setup
__global__ void grid_setup_kernel(const float* __restrict__ bbox, GridHdr* __restrict__ hdr, int B, int N,
                                  int cell_budget, float pts_per_cell) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const float* bb = bbox + b * 6;
    float ext[3], emax = 0.f, amax = 0.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        ext[c] = bb[3 + c] - bb[c];
        emax = fmaxf(emax, ext[c]);
        amax = fmaxf(amax, fmaxf(fabsf(bb[c]), fabsf(bb[3 + c])));
    }
    if (!(emax > 0.f)) emax = 1.f;  // all points identical
    // volume with degenerate (flat) axes padded to 1/64 of the largest extent
    float vol = 1.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) vol *= fmaxf(ext[c], emax * (1.f / 64.f));
    float s = cbrtf(vol * pts_per_cell / (float)N);
    s = fmaxf(s, emax / (float)(kGridMaxAxis - 1));
    int g[3];
    for (int it = 0; it < 8; ++it) {
        long long prod = 1;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            g[c] = min(kGridMaxAxis, (int)floorf(ext[c] / s) + 1);
            prod *= g[c];
        }
        if (prod <= cell_budget) break;
        s *= cbrtf((float)prod / (float)cell_budget) * 1.02f;
    }
    GridHdr h;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        h.lo[c] = bb[c];
        h.g[c] = g[c];
    }
    h.s = s;
    h.inv_s = 1.f / s;
    h.margin = 1e-3f * s + 1e-6f * (amax + emax);
    h.pad[0] = h.pad[1] = h.pad[2] = 0;
    hdr[b] = h;
}

// ------------------------------------------------------------------------------ count / scan / scatter
__global__ void __launch_bounds__(256) grid_count_kernel(const float* __restrict__ xyz, long long bstride, int N,
                                                         const GridHdr* __restrict__ hdr, int* __restrict__ counts,
                                                         int cells_alloc, int* __restrict__ cell_of) {
    const int b = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const GridHdr h = hdr[b];
    const float* p = xyz + (size_t)b * bstride + (size_t)i * 3;
    const int cx = cell_coord(p[0], h.lo[0], h.inv_s, h.g[0]);
    const int cy = cell_coord(p[1], h.lo[1], h.inv_s, h.g[1]);
    const int cz = cell_coord(p[2], h.lo[2], h.inv_s, h.g[2]);
    const int c = (cz * h.g[1] + cy) * h.g[0] + cx;
    cell_of[(size_t)b * N + i] = c;
    atomicAdd(&counts[(size_t)b * cells_alloc + c], 1);
}

// exclusive scan of counts[b][0..cells) in place; counts[b][cells] = total.  One CTA per cloud.
__global__ void __launch_bounds__(1024) grid_scan_kernel(int* __restrict__ counts, int cells_alloc,
                                                         const GridHdr* __restrict__ hdr) {
    const int b = blockIdx.x;
    const GridHdr h = hdr[b];
    const int cells = h.g[0] * h.g[1] * h.g[2];
    int* c = counts + (size_t)b * cells_alloc;
    __shared__ int warp_sum[32];
    __shared__ int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const int per = 4;
    for (int base = 0; base < cells; base += 1024 * per) {
        int v[per], s = 0;
        const int i0 = base + threadIdx.x * per;
#pragma unroll
        for (int u = 0; u < per; ++u) {
            v[u] = (i0 + u < cells) ? c[i0 + u] : 0;
            s += v[u];
        }
        int inc = s;
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, inc, o);
            if ((threadIdx.x & 31) >= o) inc += t;
        }
        if ((threadIdx.x & 31) == 31) warp_sum[threadIdx.x >> 5] = inc;
        __syncthreads();
        if (threadIdx.x < 32) {
            int w = warp_sum[threadIdx.x];
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, w, o);
                if (threadIdx.x >= o) w += t;
            }
            warp_sum[threadIdx.x] = w;
        }
        __syncthreads();
        int excl = carry + inc - s + ((threadIdx.x >> 5) ? warp_sum[(threadIdx.x >> 5) - 1] : 0);
#pragma unroll
        for (int u = 0; u < per; ++u) {
            if (i0 + u < cells) c[i0 + u] = excl;
            excl += v[u];
        }
        __syncthreads();
        if (threadIdx.x == 1023) carry = excl;
        __syncthreads();
    }
    if (threadIdx.x == 0) c[cells] = carry;
}

// support: pts[start[c] + slot] = {x,y,z,index}; queries: order[start[c] + slot] = index
__global__ void __launch_bounds__(256) grid_scatter_kernel(const float* __restrict__ xyz, long long bstride, int N,
                                                           const int* __restrict__ cell_of,
                                                           const int* __restrict__ starts, int* __restrict__ cursor,
                                                           int cells_alloc, float4* __restrict__ pts,
                                                           int* __restrict__ order) {
    const int b = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const int c = cell_of[(size_t)b * N + i];
    const int slot = atomicAdd(&cursor[(size_t)b * cells_alloc + c], 1);
    const int pos = starts[(size_t)b * cells_alloc + c] + slot;
    if (pts) {
        const float* p = xyz + (size_t)b * bstride + (size_t)i * 3;
        pts[(size_t)b * N + pos] = make_float4(p[0], p[1], p[2], __int_as_float(i));
    }
    if (order) order[(size_t)b * N + pos] = i;
}

// ----------------------------------------------------------------------------------------- query
constexpr int kGridThreads = 128;

struct KBest {
    float* d;   // column of this thread: d[k * kGridThreads]
    int* id;
    int K;
    float thr;  // K-th d2 (inf until the list is full)
    int thr_id;
};

__device__ __forceinline__ void kbest_insert(KBest& L, float d, int id) {
    // lexicographic (d, id) admission against the current K-th entry
    if (!(d < L.thr || (d == L.thr && id < L.thr_id))) return;
    int pos = L.K - 1;
    while (pos > 0) {
        const float pd = L.d[(pos - 1) * kGridThreads];
        const int pi = L.id[(pos - 1) * kGridThreads];
        if (!(pd > d || (pd == d && pi > id))) break;
        L.d[pos * kGridThreads] = pd;
        L.id[pos * kGridThreads] = pi;
        --pos;
    }
    L.d[pos * kGridThreads] = d;
    L.id[pos * kGridThreads] = id;
    L.thr = L.d[(L.K - 1) * kGridThreads];
    L.thr_id = L.id[(L.K - 1) * kGridThreads];
}

// loop-free lexicographic insertion for a compile-time list length (see knn.cu list_insert_t)
template <int KT>
__device__ __noinline__ void kbest_insert_t(KBest& L, float d, int id) {
    if (!(d < L.thr || (d == L.thr && id < L.thr_id))) return;
    float vd[KT];
    int vi[KT];
#pragma unroll
    for (int k = 0; k < KT; ++k) {
        vd[k] = L.d[k * kGridThreads];
        vi[k] = L.id[k * kGridThreads];
    }
    int pos = 0;
#pragma unroll
    for (int k = 0; k < KT; ++k) pos += (vd[k] < d || (vd[k] == d && vi[k] < id)) ? 1 : 0;
#pragma unroll
    for (int k = KT - 1; k >= 1; --k) {
        if (k > pos) {
            L.d[k * kGridThreads] = vd[k - 1];
            L.id[k * kGridThreads] = vi[k - 1];
        }
    }
    L.d[pos * kGridThreads] = d;
    L.id[pos * kGridThreads] = id;
    L.thr = (pos == KT - 1) ? d : vd[KT - 2];
    L.thr_id = (pos == KT - 1) ? id : vi[KT - 2];
}

template <bool K1, int KT = 0>
__global__ void __launch_bounds__(kGridThreads) knn_grid_kernel(
    const float4* __restrict__ pts, const int* __restrict__ starts, int cells_alloc, const GridHdr* __restrict__ hdr,
    const float* __restrict__ query, long long q_stride, const int* __restrict__ qorder /* nullable: self search */,
    int Ns, int Nq, int K, int64_t* __restrict__ idx64, int32_t* __restrict__ idx32, float* __restrict__ dist,
    float* __restrict__ dist_sq) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* list_d = reinterpret_cast<float*>(smem_raw);
    int* list_i = reinterpret_cast<int*>(list_d + (size_t)(K1 ? 0 : K) * kGridThreads);

    const int b = blockIdx.y;
    const int t = blockIdx.x * kGridThreads + threadIdx.x;
    if (t >= Nq) return;
    const GridHdr h = hdr[b];
    const float4* P = pts + (size_t)b * Ns;
    const int* S = starts + (size_t)b * cells_alloc;

    // queries are visited in cell order; a self search reads the query (and its index) from the sorted support
    int qi;
    float qx, qy, qz;
    if (qorder) {
        qi = qorder[(size_t)b * Nq + t];
        const float* q = query + (size_t)b * q_stride + (size_t)qi * 3;
        qx = q[0]; qy = q[1]; qz = q[2];
    } else {
        const float4 q = P[t];
        qx = q.x; qy = q.y; qz = q.z;
        qi = __float_as_int(q.w);
    }
    const int cx = cell_coord(qx, h.lo[0], h.inv_s, h.g[0]);
    const int cy = cell_coord(qy, h.lo[1], h.inv_s, h.g[1]);
    const int cz = cell_coord(qz, h.lo[2], h.inv_s, h.g[2]);

    KBest L{list_d + threadIdx.x, list_i + threadIdx.x, K, CUDART_INF_F, 0x7fffffff};
    float best_d = CUDART_INF_F;
    int best_i = 0x7fffffff;
    if (!K1) {
        for (int k = 0; k < K; ++k) {
            L.d[k * kGridThreads] = CUDART_INF_F;
            L.id[k * kGridThreads] = 0x7fffffff;
        }
    }

    auto consider = [&](const float4 p) {
        const float d = d2_contract_g(qx, qy, qz, p.x, p.y, p.z);
        const int id = __float_as_int(p.w);
        if (K1) {
            if (d < best_d || (d == best_d && id < best_i)) {
                best_d = d;
                best_i = id;
            }
        } else if (KT > 1) {
            kbest_insert_t<(KT > 1 ? KT : 2)>(L, d, id);
        } else {
            kbest_insert(L, d, id);
        }
    };
    // candidates four at a time: the loads are independent of the insertions, so issuing them together hides the
    // L1/L2 latency that a one-at-a-time loop pays per candidate (small clouds run one warp per scheduler)
    auto scan = [&](int lo, int hi) {
        int j = lo;
        for (; j + 4 <= hi; j += 4) {
            const float4 p0 = __ldg(&P[j]), p1 = __ldg(&P[j + 1]), p2 = __ldg(&P[j + 2]), p3 = __ldg(&P[j + 3]);
            consider(p0);
            consider(p1);
            consider(p2);
            consider(p3);
        }
        for (; j < hi; ++j) consider(__ldg(&P[j]));
    };

    const int gx = h.g[0], gy = h.g[1], gz = h.g[2];
    const int rmax = max(max(gx, gy), gz);
    for (int r = 0; r <= rmax; ++r) {
        const int z0 = max(cz - r, 0), z1 = min(cz + r, gz - 1);
        const int y0 = max(cy - r, 0), y1 = min(cy + r, gy - 1);
        const int x0 = max(cx - r, 0), x1 = min(cx + r, gx - 1);
        for (int z = z0; z <= z1; ++z) {
            const bool zface = (z == cz - r) || (z == cz + r);
            for (int y = y0; y <= y1; ++y) {
                const int row = (z * gy + y) * gx;
                if (zface || y == cy - r || y == cy + r) {
                    scan(S[row + x0], S[row + x1 + 1]);
                } else {
                    if (cx - r >= 0) scan(S[row + cx - r], S[row + cx - r + 1]);
                    if (cx + r < gx && r > 0) scan(S[row + cx + r], S[row + cx + r + 1]);
                }
            }
        }
        // ---- stop test: distance from the query to the nearest block face with cells behind it
        float bound = CUDART_INF_F;
        if (cx - r > 0) bound = fminf(bound, qx - (h.lo[0] + (float)(cx - r) * h.s));
        if (cx + r < gx - 1) bound = fminf(bound, (h.lo[0] + (float)(cx + r + 1) * h.s) - qx);
        if (cy - r > 0) bound = fminf(bound, qy - (h.lo[1] + (float)(cy - r) * h.s));
        if (cy + r < gy - 1) bound = fminf(bound, (h.lo[1] + (float)(cy + r + 1) * h.s) - qy);
        if (cz - r > 0) bound = fminf(bound, qz - (h.lo[2] + (float)(cz - r) * h.s));
        if (cz + r < gz - 1) bound = fminf(bound, (h.lo[2] + (float)(cz + r + 1) * h.s) - qz);
        if (bound == CUDART_INF_F) break;  // the block covers the whole grid
        bound = (bound - h.margin) * 0.99999f;
        const float kth = K1 ? best_d : L.thr;
        if (bound > 0.f && kth < bound * bound) break;
    }

    const size_t o = ((size_t)b * Nq + qi) * K;
    for (int k = 0; k < K; ++k) {
        const float d = K1 ? best_d : L.d[k * kGridThreads];
        const int id = K1 ? best_i : L.id[k * kGridThreads];
        if (idx64) idx64[o + k] = id;
        if (idx32) idx32[o + k] = id;
        if (dist) dist[o + k] = __fsqrt_rn(d);
        if (dist_sq) dist_sq[o + k] = d;
    }
}

// ------------------------------------------------------------------------ query, one warp per query
// The thread-per-query walk above is latency-bound: every candidate goes through a dependent shared-memory insertion
// and a batch of 20 000 queries (8 clouds of 2 500 points) fills half the SMs with four warps each (230 us in the
// training step's launch list).  Here a WARP walks the rings of one query: the lanes fetch 32 candidates at a time
// (the block's cell rows are contiguous runs of the cell-sorted support; a shuffle binary search over the runs'
// prefix sums maps lane -> point), candidates are 64-bit keys (d2 bits << 32 | index: d2 >= 0, so unsigned key order
// is the contract's (d2, index) order), and the running K-best is a sorted list DISTRIBUTED over the lanes (entry p
// lives in lane p % 32): inserting a key is one ballot (how many entries are smaller) and one shuffle-up.  Same rings,
// same stop test, same results as the walk above.
constexpr int kGridWarpThreads = 256;

template <int KS>
__global__ void __launch_bounds__(kGridWarpThreads) knn_grid_warp_kernel(
    const float4* __restrict__ pts, const int* __restrict__ starts, int cells_alloc, const GridHdr* __restrict__ hdr,
    const float* __restrict__ query, long long q_stride, const int* __restrict__ qorder /* nullable: self search */,
    int Ns, int Nq, int K, int64_t* __restrict__ idx64, int32_t* __restrict__ idx32, float* __restrict__ dist,
    float* __restrict__ dist_sq) {
    constexpr unsigned FULL = 0xffffffffu;
    constexpr unsigned long long NONE = ~0ull;
    const int b = blockIdx.y;
    const int lane = threadIdx.x & 31;
    const long long t = (long long)blockIdx.x * (kGridWarpThreads / 32) + (threadIdx.x >> 5);
    if (t >= Nq) return;                                   // warp-uniform
    const GridHdr h = hdr[b];
    const float4* P = pts + (size_t)b * Ns;
    const int* S = starts + (size_t)b * cells_alloc;

    int qi;
    float qx, qy, qz;
    if (qorder) {
        qi = qorder[(size_t)b * Nq + t];
        const float* q = query + (size_t)b * q_stride + (size_t)qi * 3;
        qx = q[0]; qy = q[1]; qz = q[2];
    } else {
        const float4 q = P[t];
        qx = q.x; qy = q.y; qz = q.z;
        qi = __float_as_int(q.w);
    }
    const int gx = h.g[0], gy = h.g[1], gz = h.g[2];
    const int cx = cell_coord(qx, h.lo[0], h.inv_s, gx);
    const int cy = cell_coord(qy, h.lo[1], h.inv_s, gy);
    const int cz = cell_coord(qz, h.lo[2], h.inv_s, gz);

    unsigned long long lst[KS];
#pragma unroll
    for (int s = 0; s < KS; ++s) lst[s] = NONE;
    unsigned long long kth = NONE;                          // key of entry K-1 (NONE until the list is full)
    bool fresh = true;                                      // nothing inserted yet
    const int kslot = (K - 1) >> 5, klane = (K - 1) & 31;

    const int rmax = max(max(gx, gy), gz);
    for (int r = 1; r <= rmax + 1; ++r) {
        const int z0 = max(cz - r, 0), z1 = min(cz + r, gz - 1);
        const int y0 = max(cy - r, 0), y1 = min(cy + r, gy - 1);
        const int x0 = max(cx - r, 0), x1 = min(cx + r, gx - 1);
        const int ny = y1 - y0 + 1, nrows = ny * (z1 - z0 + 1);
        const bool block = (r == 1);                        // the first pass takes the whole 3x3x3 block, later ones a shell
        const int rows_per_batch = block ? 32 : 16;
        for (int rb = 0; rb < nrows; rb += rows_per_batch) {
            // ---- this lane's run of the cell-sorted support
            int beg = 0, len = 0;
            const int i = rb + (block ? lane : (lane >> 1));
            if (i < nrows) {
                const int z = z0 + i / ny, y = y0 + i % ny;
                const int row = (z * gy + y) * gx;
                if (block || z == cz - r || z == cz + r || y == cy - r || y == cy + r) {
                    if (block || !(lane & 1)) {
                        beg = S[row + x0];
                        len = S[row + x1 + 1] - beg;
                    }
                } else {
                    const int xx = (lane & 1) ? cx + r : cx - r;
                    if (xx >= 0 && xx < gx) {
                        beg = S[row + xx];
                        len = S[row + xx + 1] - beg;
                    }
                }
            }
            int incl = len;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int v = __shfl_up_sync(FULL, incl, o);
                if (lane >= o) incl += v;
            }
            const int total = __shfl_sync(FULL, incl, 31);
            const int shift = beg - (incl - len);           // point index = candidate number + shift, within this run
            for (int cb = 0; cb < total; cb += 32) {
                const int tt = cb + lane;
                int sgm = 0;                                // first run whose inclusive prefix exceeds tt
#pragma unroll
                for (int step = 16; step >= 1; step >>= 1) {
                    const int v = __shfl_sync(FULL, incl, (sgm + step - 1) & 31);
                    if (tt >= v) sgm += step;
                }
                const int sh = __shfl_sync(FULL, shift, sgm & 31);
                unsigned long long key = NONE;
                if (tt < total) {
                    const float4 p = __ldg(&P[tt + sh]);
                    const float d = d2_contract_g(qx, qy, qz, p.x, p.y, p.z);
                    key = ((unsigned long long)__float_as_uint(d) << 32) | (unsigned)__float_as_int(p.w);
                }
                unsigned mask = __ballot_sync(FULL, key < kth);
                if (KS == 1 && (fresh || __popc(mask) >= 8)) {
                    // many admissible candidates (always true for the first chunk of a query): sort the chunk across
                    // the lanes (15-stage bitonic network) and merge it with the list -- the element-wise minimum of
                    // the list and the reversed chunk is a bitonic sequence holding the 32 smallest keys of both,
                    // which five more stages sort.  ~200 instructions instead of ~25 per insertion.
                    if (!(key < kth)) key = NONE;
#pragma unroll
                    for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
                        for (int j = k >> 1; j > 0; j >>= 1) {
                            const unsigned long long other = __shfl_xor_sync(FULL, key, j);
                            const bool keep_min = ((lane & k) == 0) == ((lane & j) == 0);
                            key = ((other < key) == keep_min) ? other : key;
                        }
                    }
                    if (fresh) {
                        lst[0] = key;
                    } else {
                        const unsigned long long rev = __shfl_sync(FULL, key, 31 - lane);
                        unsigned long long m = rev < lst[0] ? rev : lst[0];
#pragma unroll
                        for (int j = 16; j > 0; j >>= 1) {
                            const unsigned long long other = __shfl_xor_sync(FULL, m, j);
                            const bool keep_min = (lane & j) == 0;
                            m = ((other < m) == keep_min) ? other : m;
                        }
                        lst[0] = m;
                    }
                    fresh = false;
                    kth = __shfl_sync(FULL, lst[0], klane);
                    continue;
                }
                fresh = false;
                while (mask) {
                    const int src = __ffs(mask) - 1;
                    mask &= mask - 1;
                    const unsigned long long x = __shfl_sync(FULL, key, src);
                    if (x < kth) {                          // kth shrinks while the chunk is inserted
                        int pos = 0;
#pragma unroll
                        for (int s = 0; s < KS; ++s) pos += __popc(__ballot_sync(FULL, lst[s] < x));
#pragma unroll
                        for (int s = KS - 1; s >= 0; --s) {
                            unsigned long long up = __shfl_up_sync(FULL, lst[s], 1);
                            if (s > 0) {
                                const unsigned long long carry = __shfl_sync(FULL, lst[s - 1], 31);
                                if (lane == 0) up = carry;
                            }
                            const int gp = s * 32 + lane;
                            if (gp > pos) lst[s] = up;
                            else if (gp == pos) lst[s] = x;
                        }
                        unsigned long long tail = lst[0];
#pragma unroll
                        for (int s = 1; s < KS; ++s)
                            if (kslot == s) tail = lst[s];
                        kth = __shfl_sync(FULL, tail, klane);
                    }
                }
            }
        }
        // ---- stop test: distance from the query to the nearest block face with cells behind it
        float bound = CUDART_INF_F;
        if (cx - r > 0) bound = fminf(bound, qx - (h.lo[0] + (float)(cx - r) * h.s));
        if (cx + r < gx - 1) bound = fminf(bound, (h.lo[0] + (float)(cx + r + 1) * h.s) - qx);
        if (cy - r > 0) bound = fminf(bound, qy - (h.lo[1] + (float)(cy - r) * h.s));
        if (cy + r < gy - 1) bound = fminf(bound, (h.lo[1] + (float)(cy + r + 1) * h.s) - qy);
        if (cz - r > 0) bound = fminf(bound, qz - (h.lo[2] + (float)(cz - r) * h.s));
        if (cz + r < gz - 1) bound = fminf(bound, (h.lo[2] + (float)(cz + r + 1) * h.s) - qz);
        if (bound == CUDART_INF_F) break;  // the block covers the whole grid
        bound = (bound - h.margin) * 0.99999f;
        const float kd = kth == NONE ? CUDART_INF_F : __uint_as_float((unsigned)(kth >> 32));
        if (bound > 0.f && kd < bound * bound) break;
    }

    const size_t o = ((size_t)b * Nq + qi) * K;
#pragma unroll
    for (int s = 0; s < KS; ++s) {
        const int k = s * 32 + lane;
        if (k < K) {
            const float d = __uint_as_float((unsigned)(lst[s] >> 32));
            const int id = (int)(unsigned)lst[s];
            if (idx64) idx64[o + k] = id;
            if (idx32) idx32[o + k] = id;
            if (dist) dist[o + k] = __fsqrt_rn(d);
            if (dist_sq) dist_sq[o + k] = d;
        }
    }
}

// ------------------------------------------------------------------------------------- host side
struct GridPlan {
    int cells_alloc;    // ints per cloud in a count/start array (cell budget + 1, rounded up)
    size_t off_bbox, off_hdr, off_counts, off_cursor, off_cellof, off_pts, off_qcounts, off_qcursor, off_qcellof,
        off_qorder, total;
};

static GridPlan grid_plan(int B, int Ns, int Nq, bool self) {
    GridPlan p;
    long long budget = 1;
    while (budget < (long long)Ns / 2) budget <<= 1;
    if (budget < 64) budget = 64;
    if (budget > (1 << 21)) budget = 1 << 21;
    p.cells_alloc = (int)budget + 32;
    size_t o = 0;
    auto take = [&](size_t bytes) {
        const size_t at = o;
        o += align_up(bytes, 256);
        return at;
    };
    p.off_bbox = take((size_t)B * 6 * sizeof(float));
    p.off_hdr = take((size_t)B * sizeof(GridHdr));
    p.off_counts = take((size_t)B * p.cells_alloc * sizeof(int));
    p.off_cursor = take((size_t)B * p.cells_alloc * sizeof(int));
    p.off_cellof = take((size_t)B * Ns * sizeof(int));
    p.off_pts = take((size_t)B * Ns * sizeof(float4));
    if (!self) {
        p.off_qcounts = take((size_t)B * p.cells_alloc * sizeof(int));
        p.off_qcursor = take((size_t)B * p.cells_alloc * sizeof(int));
        p.off_qcellof = take((size_t)B * Nq * sizeof(int));
        p.off_qorder = take((size_t)B * Nq * sizeof(int));
    } else {
        p.off_qcounts = p.off_qcursor = p.off_qcellof = p.off_qorder = 0;
    }
    p.total = o;
    return p;
}

size_t knn_grid_workspace_bytes(int B, int Ns, int Nq) {
    // sized for the cross-search case (a self search needs less)
    return grid_plan(B, Ns, Nq, false).total;
}

int knn_grid_run(const float* support, long long s_stride, const float* query, long long q_stride, int B, int Ns,
                 int Nq, int K, int64_t* idx64, int32_t* idx32, float* dist, float* dist_sq, void* workspace,
                 cudaStream_t st) {
    const bool self = (support == query && Ns == Nq && s_stride == q_stride);
    const GridPlan p = grid_plan(B, Ns, Nq, self);
    unsigned char* w = static_cast<unsigned char*>(workspace);
    float* bbox = reinterpret_cast<float*>(w + p.off_bbox);
    GridHdr* hdr = reinterpret_cast<GridHdr*>(w + p.off_hdr);
    int* counts = reinterpret_cast<int*>(w + p.off_counts);
    int* cursor = reinterpret_cast<int*>(w + p.off_cursor);
    int* cellof = reinterpret_cast<int*>(w + p.off_cellof);
    float4* pts = reinterpret_cast<float4*>(w + p.off_pts);
    const int budget = p.cells_alloc - 32;

    // counts and cursor are adjacent: one memset clears both
    R3D_CUDA_TRY(cudaMemsetAsync(counts, 0, (size_t)(p.off_cellof - p.off_counts), st));
    grid_bbox_kernel<<<B, 1024, 0, st>>>(support, s_stride, Ns, bbox);
    R3D_LAUNCH_CHECK("grid_bbox_kernel");
    grid_setup_kernel<<<ceil_div(B, 128), 128, 0, st>>>(bbox, hdr, B, Ns, budget, g_pts_per_cell > 0.f ? g_pts_per_cell : (K >= 24 ? 4.f : 2.f));
    R3D_LAUNCH_CHECK("grid_setup_kernel");
    {
        dim3 grid(ceil_div(Ns, 256), B);
        grid_count_kernel<<<grid, 256, 0, st>>>(support, s_stride, Ns, hdr, counts, p.cells_alloc, cellof);
        R3D_LAUNCH_CHECK("grid_count_kernel");
        grid_scan_kernel<<<B, 1024, 0, st>>>(counts, p.cells_alloc, hdr);
        R3D_LAUNCH_CHECK("grid_scan_kernel");
        grid_scatter_kernel<<<grid, 256, 0, st>>>(support, s_stride, Ns, cellof, counts, cursor, p.cells_alloc, pts,
                                                  nullptr);
        R3D_LAUNCH_CHECK("grid_scatter_kernel");
    }
    int* qorder = nullptr;
    if (!self) {
        int* qcounts = reinterpret_cast<int*>(w + p.off_qcounts);
        int* qcursor = reinterpret_cast<int*>(w + p.off_qcursor);
        int* qcellof = reinterpret_cast<int*>(w + p.off_qcellof);
        qorder = reinterpret_cast<int*>(w + p.off_qorder);
        R3D_CUDA_TRY(cudaMemsetAsync(qcounts, 0, (size_t)(p.off_qcellof - p.off_qcounts), st));
        dim3 grid(ceil_div(Nq, 256), B);
        grid_count_kernel<<<grid, 256, 0, st>>>(query, q_stride, Nq, hdr, qcounts, p.cells_alloc, qcellof);
        R3D_LAUNCH_CHECK("grid_count_kernel(q)");
        grid_scan_kernel<<<B, 1024, 0, st>>>(qcounts, p.cells_alloc, hdr);
        R3D_LAUNCH_CHECK("grid_scan_kernel(q)");
        grid_scatter_kernel<<<grid, 256, 0, st>>>(query, q_stride, Nq, qcellof, qcounts, qcursor, p.cells_alloc,
                                                  nullptr, qorder);
        R3D_LAUNCH_CHECK("grid_scatter_kernel(q)");
    }
    // One warp per query wins while the batch is too small to fill the SMs with independent query threads; beyond,
    // the thread-per-query walk has the higher throughput (measured, profiles/r01_knn_small.txt: 8 x 2500 K=16
    // 42 vs 144 us, 64 x 40960 3.8 vs 3.1 ms; from K=24 up the warp wins at every size: 1M x 1M K=32 3.0 vs 4.4 ms;
    // K=1 has no list to share: thread walk).
    const long long nq_total = (long long)B * Nq;
    const bool warp_walk = g_grid_walk >= 0 ? g_grid_walk == 0 : (K > 1 && (K >= 24 || nq_total <= 160000LL));
    if (warp_walk) {
        dim3 wgrid(ceil_div(Nq, kGridWarpThreads / 32), B);
        if (K <= 32)
            knn_grid_warp_kernel<1><<<wgrid, kGridWarpThreads, 0, st>>>(pts, counts, p.cells_alloc, hdr, query, q_stride,
                                                                        qorder, Ns, Nq, K, idx64, idx32, dist, dist_sq);
        else
            knn_grid_warp_kernel<2><<<wgrid, kGridWarpThreads, 0, st>>>(pts, counts, p.cells_alloc, hdr, query, q_stride,
                                                                        qorder, Ns, Nq, K, idx64, idx32, dist, dist_sq);
        R3D_LAUNCH_CHECK("knn_grid_warp_kernel");
        return R3D_OK;
    }
    dim3 grid(ceil_div(Nq, kGridThreads), B);
    if (K == 1) {
        knn_grid_kernel<true><<<grid, kGridThreads, 0, st>>>(pts, counts, p.cells_alloc, hdr, query, q_stride, qorder,
                                                             Ns, Nq, K, idx64, idx32, dist, dist_sq);
    } else {
        const size_t smem = (size_t)K * kGridThreads * 8;
        // the ring walk meets candidates roughly in order of distance, so insertions land near the list tail and
        // the short shift loop beats the loop-free version (measured: 1M x 1M K=16 1.99 ms vs 3.85 ms)
        auto kern = knn_grid_kernel<false, 0>;
        R3D_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<grid, kGridThreads, smem, st>>>(pts, counts, p.cells_alloc, hdr, query, q_stride, qorder, Ns, Nq, K,
                                               idx64, idx32, dist, dist_sq);
    }
    R3D_LAUNCH_CHECK("knn_grid_kernel");
    return R3D_OK;
}


void knn_grid_set_density(float v) { g_pts_per_cell = v; }
void knn_grid_set_walk(int thread_per_query) { g_grid_walk = thread_per_query; }

}  // namespace r3d
