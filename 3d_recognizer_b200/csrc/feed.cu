// Batch assembly from clouds cached in HBM: sub/up-sampling gather, normalisation and augmentation in ONE launch per
// batch (sm_100a), plus a device-side random subset draw.
//
// Reference: randlanet/utils/dataset.py:61-97 (PointCloudPreprocessor.preprocess: sample -> normalise -> augment, per
// item, numpy, inside a single-process DataLoader) and randlanet/utils/augmentation.py:24-167 (jitter scaled by the mean
// radius and clipped; scale about the centre; Rz Ry Rx rotation about the centre; shift scaled by the mean radius).
// Each of those steps recomputes the cloud's centre / mean radius with a full numpy pass; here one CTA owns one cloud of
// the batch and runs the whole chain over its rows (L2-resident between passes), with fp64 block reductions:
//
//   pass 0  gather rows (xyz + features) and labels at the sample indices;  sum x            -> centre c0
//   pass 1  d = |x - c0|: sum d, sum d^2, max d                                              -> radius (mean/max/stdev)
//   pass 2  x' = (x - c0) / radius;  x1 = x' + clip(r_jit sigma z, +-limit);  sum x1         -> centre c1
//   pass 3  sum |x1 - c1|                                                                    -> mean radius r1
//   pass 4  x4 = s R (x1 - c1) + c1 + s r1 u                                                 (scale, rotate, shift)
//
// (scaling and rotating about the centre leave the centre where it is and multiply the mean radius by s, which is what
// the reference's per-step recomputation finds, to round-off).  The per-cloud random numbers (s, three angles, u) are
// drawn by the host in the reference's order; the per-point standard normals z either come from the host too (the
// reference's numpy stream: results equal to the reference's to round-off) or from Philox4x32-10 on the device
// (statistically equivalent, a different stream; what a multi-GPU run uses so that the host draws 7 numbers per cloud).
#include "common.cuh"

namespace r3d {

constexpr int kFeedThreads = 1024;

// ---- Philox4x32-10 (Salmon et al., SC'11): counter-based, no state
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
        k.x += 0x9E3779B9u;
        k.y += 0xBB67AE85u;
    }
    return c;
}

__device__ __forceinline__ float u01(uint32_t x) { return ((float)x + 1.0f) * 2.3283064365386963e-10f; }  // (0, 1]

// three standard normals for point i of cloud b of batch `counter` (Box-Muller on the four Philox words)
__device__ __forceinline__ float3 philox_normal3(unsigned long long seed, unsigned long long counter, int b, int i) {
    const uint4 r = philox4x32_10(make_uint4((uint32_t)i, (uint32_t)b, (uint32_t)counter, (uint32_t)(counter >> 32)),
                                  make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
    const float ra = sqrtf(-2.0f * logf(u01(r.x))), rb = sqrtf(-2.0f * logf(u01(r.z)));
    float sa, ca, sb, cb;
    sincospif(2.0f * u01(r.y), &sa, &ca);
    sincospif(2.0f * u01(r.w), &sb, &cb);
    (void)sb;
    return make_float3(ra * ca, ra * sa, rb * cb);
}

// sum of three doubles (and max of a fourth) over the CTA, result to every thread
struct Red4 {
    double a, b, c, m;
};

__device__ __forceinline__ Red4 block_reduce(Red4 v, Red4* sh) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        v.a += __shfl_xor_sync(0xffffffffu, v.a, o);
        v.b += __shfl_xor_sync(0xffffffffu, v.b, o);
        v.c += __shfl_xor_sync(0xffffffffu, v.c, o);
        v.m = fmax(v.m, __shfl_xor_sync(0xffffffffu, v.m, o));
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    __syncthreads();                                   // sh may still be read from the previous reduction
    if (lane == 0) sh[warp] = v;
    __syncthreads();
    Red4 t = lane < nw ? sh[lane] : Red4{0.0, 0.0, 0.0, 0.0};
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        t.a += __shfl_xor_sync(0xffffffffu, t.a, o);
        t.b += __shfl_xor_sync(0xffffffffu, t.b, o);
        t.c += __shfl_xor_sync(0xffffffffu, t.c, o);
        t.m = fmax(t.m, __shfl_xor_sync(0xffffffffu, t.m, o));
    }
    return t;
}

enum { kNormNone = 0, kNormMean = 1, kNormMax = 2, kNormStdev = 3, kNormCentre = 4 };  // 4: any other string (radius 1)

__global__ void __launch_bounds__(kFeedThreads)
    feed_batch_kernel(const float* __restrict__ points, int ld, const int64_t* __restrict__ labels_in,
                      const int64_t* __restrict__ row_start, const int32_t* __restrict__ sample_idx, int n,
                      int normalization, const float* __restrict__ aug, const float* __restrict__ noise,
                      float jitter_sigma, float jitter_limit, unsigned long long seed, unsigned long long counter,
                      float* __restrict__ out, int64_t* __restrict__ labels_out) {
    __shared__ Red4 sh[32];
    const int b = blockIdx.x;
    const long long base = row_start[b];
    const int32_t* sidx = sample_idx + (long long)b * n;
    float* o = out + (long long)b * n * ld;

    // pass 0: gather
    Red4 acc{0.0, 0.0, 0.0, 0.0};
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const long long src = base + sidx[i];
        const float* p = points + src * ld;
        float* q = o + (long long)i * ld;
        for (int c = 0; c < ld; ++c) q[c] = p[c];
        acc.a += p[0], acc.b += p[1], acc.c += p[2];
        if (labels_out) labels_out[(long long)b * n + i] = labels_in[src];
    }
    if (normalization == kNormNone && aug == nullptr) return;
    Red4 t = block_reduce(acc, sh);
    const double c0x = t.a / n, c0y = t.b / n, c0z = t.c / n;

    // pass 1: distances to the centre
    acc = Red4{0.0, 0.0, 0.0, 0.0};
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const float* q = o + (long long)i * ld;
        const double dx = q[0] - c0x, dy = q[1] - c0y, dz = q[2] - c0z;
        const double d = sqrt(dx * dx + dy * dy + dz * dz);
        acc.a += d, acc.b += d * d;
        acc.m = fmax(acc.m, d);
    }
    t = block_reduce(acc, sh);
    const double mean_d = t.a / n;
    double radius = 1.0;                                 // dataset.py:84-92
    if (normalization == kNormMean) radius = mean_d;
    if (normalization == kNormMax) radius = t.m;
    if (normalization == kNormStdev) radius = sqrt(fmax(t.b / n - mean_d * mean_d, 0.0));
    const bool normalise = normalization != kNormNone;
    // after normalisation the cloud is centred at 0 with mean radius mean_d / radius (augmentation.py:24-33)
    const double r_jit = normalise ? mean_d / radius : mean_d;
    const double inv_radius = 1.0 / radius;
    if (aug == nullptr) {
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            float* q = o + (long long)i * ld;
            q[0] = (float)((q[0] - c0x) * inv_radius), q[1] = (float)((q[1] - c0y) * inv_radius);
            q[2] = (float)((q[2] - c0z) * inv_radius);
        }
        return;
    }

    // pass 2: normalise + jitter (augmentation.py:36-55)
    const double amp = r_jit * (double)jitter_sigma;
    const double lim = (double)jitter_limit;
    acc = Red4{0.0, 0.0, 0.0, 0.0};
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        float* q = o + (long long)i * ld;
        double x = q[0], y = q[1], z = q[2];
        if (normalise) x = (x - c0x) * inv_radius, y = (y - c0y) * inv_radius, z = (z - c0z) * inv_radius;
        float3 g;
        if (noise) {
            const float* np_ = noise + ((long long)b * n + i) * 3;
            g = make_float3(np_[0], np_[1], np_[2]);
        } else {
            g = philox_normal3(seed, counter, b, i);
        }
        x += fmin(fmax(amp * g.x, -lim), lim), y += fmin(fmax(amp * g.y, -lim), lim);
        z += fmin(fmax(amp * g.z, -lim), lim);
        q[0] = (float)x, q[1] = (float)y, q[2] = (float)z;
        acc.a += q[0], acc.b += q[1], acc.c += q[2];
    }
    t = block_reduce(acc, sh);
    const double c1x = t.a / n, c1y = t.b / n, c1z = t.c / n;

    // pass 3: mean radius of the jittered cloud
    acc = Red4{0.0, 0.0, 0.0, 0.0};
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const float* q = o + (long long)i * ld;
        const double dx = q[0] - c1x, dy = q[1] - c1y, dz = q[2] - c1z;
        acc.a += sqrt(dx * dx + dy * dy + dz * dz);
    }
    t = block_reduce(acc, sh);
    const double r1 = t.a / n;

    // pass 4: scale, rotate (R = Rz Ry Rx, rows applied to the centred point), shift (augmentation.py:58-140)
    const float* a = aug + b * 7;
    const double s = a[0];
    const double cx = cos((double)a[1]), sx = sin((double)a[1]), cy = cos((double)a[2]), sy = sin((double)a[2]);
    const double cz = cos((double)a[3]), sz = sin((double)a[3]);
    const double R00 = cz * cy, R01 = cz * sy * sx - sz * cx, R02 = cz * sy * cx + sz * sx;
    const double R10 = sz * cy, R11 = sz * sy * sx + cz * cx, R12 = sz * sy * cx - cz * sx;
    const double R20 = -sy, R21 = cy * sx, R22 = cy * cx;
    const double shift = s * r1;
    const double tx = c1x + shift * a[4], ty = c1y + shift * a[5], tz = c1z + shift * a[6];
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        float* q = o + (long long)i * ld;
        const double x = (q[0] - c1x) * s, y = (q[1] - c1y) * s, z = (q[2] - c1z) * s;
        q[0] = (float)(R00 * x + R01 * y + R02 * z + tx);
        q[1] = (float)(R10 * x + R11 * y + R12 * z + ty);
        q[2] = (float)(R20 * x + R21 * y + R22 * z + tz);
    }
}

// ---- uniform random subset of n out of N points, order-preserving (preprocessing.py:35-62 draws the same kind of
// subset with numpy's global stream, in random order; the network permutes its input itself, modules.py:571).
// Each point gets a 32-bit Philox key; the n smallest keys win (ties to the lower index): a radix select over the keys
// (four 8-bit rounds, keys recomputed instead of stored), then a scan-based compaction.  N <= n: every point once, the
// remaining n - N draws with replacement.
__device__ __forceinline__ uint32_t subset_key(unsigned long long seed, unsigned long long counter, int b, int i) {
    return philox4x32_10(make_uint4((uint32_t)i, (uint32_t)b, (uint32_t)counter, (uint32_t)(counter >> 32) ^ 0x5ab5e7u),
                         make_uint2((uint32_t)seed, (uint32_t)(seed >> 32))).x;
}

__global__ void __launch_bounds__(kFeedThreads)
    sample_subset_kernel(const int32_t* __restrict__ sizes, int n, unsigned long long seed, unsigned long long counter,
                         int32_t* __restrict__ out) {
    __shared__ unsigned int hist[256];
    __shared__ unsigned int s_prefix, s_need, s_base_sel, s_base_eq;
    __shared__ unsigned int warp_a[32], warp_b[32];
    const int b = blockIdx.x;
    const int N = sizes[b];
    int32_t* o = out + (long long)b * n;
    if (N <= n) {
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            if (i < N) {
                o[i] = i;
            } else {
                const uint32_t r = subset_key(seed, counter ^ 0x9e3779b97f4a7c15ull, b, i);
                o[i] = (int32_t)(((unsigned long long)r * (unsigned long long)N) >> 32);
            }
        }
        return;
    }
    // radix select: the n-th smallest key T, and how many keys equal to T are taken
    if (threadIdx.x == 0) s_prefix = 0u, s_need = (unsigned int)n;
    unsigned int mask = 0u;
    for (int shift = 24; shift >= 0; shift -= 8) {
        for (int i = threadIdx.x; i < 256; i += blockDim.x) hist[i] = 0u;
        __syncthreads();
        const unsigned int prefix = s_prefix;
        for (int i = threadIdx.x; i < N; i += blockDim.x) {
            const uint32_t k = subset_key(seed, counter, b, i);
            if ((k & mask) == prefix) atomicAdd(&hist[(k >> shift) & 255u], 1u);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned int need = s_need, bin = 0;
            for (; bin < 256; ++bin) {
                if (hist[bin] >= need) break;
                need -= hist[bin];
            }
            s_need = need;                                    // still to take inside this bin
            s_prefix = prefix | (bin << shift);
        }
        mask |= 255u << shift;
        __syncthreads();
    }
    const uint32_t T = s_prefix;
    const unsigned int need_eq = s_need;
    if (threadIdx.x == 0) s_base_sel = 0u, s_base_eq = 0u;
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    for (int start = 0; start < N; start += blockDim.x) {
        const int i = start + threadIdx.x;
        const uint32_t k = i < N ? subset_key(seed, counter, b, i) : 0xffffffffu;
        const bool less = i < N && k < T, eq = i < N && k == T;
        // rank among the keys equal to T (ties go to the lower index)
        const unsigned int eq_ballot = __ballot_sync(0xffffffffu, eq);
        if (lane == 0) warp_a[warp] = __popc(eq_ballot);
        __syncthreads();
        unsigned int eq_rank = s_base_eq + __popc(eq_ballot & ((1u << lane) - 1u));
        for (int w = 0; w < warp; ++w) eq_rank += warp_a[w];
        const bool sel = less || (eq && eq_rank < need_eq);
        const unsigned int sel_ballot = __ballot_sync(0xffffffffu, sel);
        if (lane == 0) warp_b[warp] = __popc(sel_ballot);
        __syncthreads();
        unsigned int pos = s_base_sel + __popc(sel_ballot & ((1u << lane) - 1u));
        for (int w = 0; w < warp; ++w) pos += warp_b[w];
        if (sel && pos < (unsigned int)n) o[pos] = i;
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned int ta = 0, tb = 0;
            for (int w = 0; w < nw; ++w) ta += warp_a[w], tb += warp_b[w];
            s_base_eq += ta, s_base_sel += tb;
        }
        __syncthreads();
    }
}

}  // namespace r3d

using namespace r3d;

extern "C" int r3d_feed_batch(const float* points, int ld, const int64_t* labels, const int64_t* row_start,
                              const int32_t* sample_idx, int n, int normalization, const float* aug, const float* noise,
                              float jitter_sigma, float jitter_limit, unsigned long long seed, unsigned long long counter,
                              float* out, int64_t* labels_out, int B, r3d_stream_t stream) {
    if (B < 0 || n < 0 || ld < 3) return R3D_EINVAL;
    if (normalization < kNormNone || normalization > kNormCentre) return R3D_EINVAL;
    if (B == 0 || n == 0) return R3D_OK;
    if (!points || !row_start || !sample_idx || !out || (labels_out && !labels)) return R3D_EINVAL;
    feed_batch_kernel<<<B, kFeedThreads, 0, static_cast<cudaStream_t>(stream)>>>(
        points, ld, labels, row_start, sample_idx, n, normalization, aug, noise, jitter_sigma, jitter_limit, seed, counter,
        out, labels_out);
    R3D_LAUNCH_CHECK("feed_batch_kernel");
    return R3D_OK;
}

extern "C" int r3d_sample_subset(const int32_t* sizes, int n, unsigned long long seed, unsigned long long counter,
                                 int32_t* out, int B, r3d_stream_t stream) {
    if (B < 0 || n < 0) return R3D_EINVAL;
    if (B == 0 || n == 0) return R3D_OK;
    if (!sizes || !out) return R3D_EINVAL;
    sample_subset_kernel<<<B, kFeedThreads, 0, static_cast<cudaStream_t>(stream)>>>(sizes, n, seed, counter, out);
    R3D_LAUNCH_CHECK("sample_subset_kernel");
    return R3D_OK;
}
