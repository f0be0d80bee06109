// Building blocks of the tensor-core ("channel-lane") fused LocSE + attentive-pooling kernels (lfa_cl.cu forward,
// lfa_cl_bwd.cu backward / moments).  Design:
//
//  * TRANSPOSED contraction.  The score Linear of AttentivePooling (modules.py:234-237, 246-252) is issued as
//        S^T (channels x rows) = Ws (channels x channels) . X^T (channels x rows)
//    so that the accumulator's TMEM LANE is the CHANNEL and its COLUMNS are the (point, neighbour) rows.  A thread then
//    owns one channel: the softmax over a point's K neighbours and the weighted sum are in-thread loops over K
//    consecutive columns (no shuffles), per-channel parameters live in registers, the reductions over rows of the
//    backward are in-thread too, and the neighbour-feature scatter is coalesced across the lanes of a warp.
//  * 128 VIRTUAL CHANNELS.  A tcgen05 MMA wants M = 128 lanes.  Layers narrower than 128 channels stack SUB = 128/d
//    sub-tiles, each with its own rows, on the lanes; the weight operand becomes block-diagonal (SUB copies of Ws) and
//    the row operand holds, for virtual channel v = (sub, c) and row slot n, channel c of row n of sub-tile sub.  The
//    zero blocks cost tensor-core time only, which is not the bound of these layers.  Lane order: lanes 0..63 carry the
//    position-encoding halves (r) of all sub-tiles, lanes 64..127 the gathered-feature halves (F), so that warps are
//    uniform in role.
//  * SPLIT-FP16 operands (tc16_common.cuh): fp32 accuracy at half the shared-memory footprint of 3xTF32.
//  * One operand layout for every role: 16-byte units of 8 consecutive rows of one channel, 8 channels adjacent
//    (128-byte core matrix), read K-major or MN-major as the GEMM at hand needs (S = X Ws^T, dX = dS Ws, dWs = dS^T X
//    all read the same two tiles).
//  * Warp specialisation: NG worker groups of 128 threads (lane = TMEM lane) take tiles round-robin and each run
//    row-info -> produce operand -> [MMA] -> epilogue; ONE extra warp's elected thread issues every MMA, polling the
//    groups' "operand ready" mbarriers, and commits completion to the group's "accumulator ready" mbarrier.  While one
//    group waits for the tensor core the others gather / convert / reduce.
#pragma once
#include "lfa_common.cuh"
#include "tc16_common.cuh"

namespace r3d {

constexpr int kClLanes = 128;       // virtual channels = TMEM lanes = threads of a worker group
constexpr int kClR = 64;            // row slots per sub-tile = MMA N = TMEM columns of an accumulator
constexpr float kClSx = 16.0f;      // fixed scale of activation operands: |x| < 4094 (status bit 0 reports a violation)
constexpr int kClRinfo = 8;         // floats per row of the row-info table: p_j (3), |p_i - p_j|, two neighbour offsets, 2 spare

template <int D, int K>
struct ClCfg {
    static_assert(D == 16 || D == 32 || D == 64 || D == 128, "width");
    static_assert(K == 16 || K == 32, "neighbours");
    static constexpr int H = D / 2;
    static constexpr int SUB = kClLanes / D;          // sub-tiles stacked on the lanes
    static constexpr int R = kClR;
    static constexpr int PTS = R / K;                 // points per sub-tile
    static constexpr int TPTS = SUB * PTS;            // points per tile
    static constexpr int ROWS = SUB * R;              // (point, neighbour) rows per tile
    static constexpr int SUB_RI = R * kClRinfo + PTS * 4 + 4;   // floats per sub-tile: rows, then p_i of its points (+ bank skew)
    static constexpr int RINFO_FLOATS = SUB * SUB_RI;
    static constexpr int OP_BYTES = kClLanes * R * 2; // one fp16 plane (hi or lo) of a row operand: 16 KB
    static constexpr int OP_CS = (R / 8) * 128;       // byte stride between channel groups (8 channels) of a row operand
    static constexpr int W_BYTES = kClLanes * kClLanes * 2;   // one plane of the 128 x 128 virtual weight matrix
    static constexpr int W_IS = (kClLanes / 8) * 128; // byte stride between input-channel groups of a weight plane
};

// lane -> (part, sub-tile, channel inside the part); real channel = part * H + c
template <int D>
struct ClLane {
    int part, sub, c;
    __device__ __forceinline__ explicit ClLane(int l) {
        constexpr int H = D / 2;
        part = l >> 6;
        sub = (l & 63) / H;
        c = (l & 63) % H;
    }
    __device__ __forceinline__ int channel() const { return part * (D / 2) + c; }
};

// power-of-two scale that brings absmax into [2^(target-1), 2^target)
__device__ __forceinline__ float cl_pow2_scale_to(float absmax, int target) {
    if (!(absmax > 0.f) || !isfinite(absmax)) return 1.0f;
    int e;
    frexpf(absmax, &e);                // absmax = m 2^e, m in [0.5, 1)
    e = target - e;
    e = e > 100 ? 100 : (e < -100 ? -100 : e);
    return exp2f((float)e);
}
// weights: absmax -> [2^11, 2^12)
__device__ __forceinline__ float cl_pow2_scale(float absmax) { return cl_pow2_scale_to(absmax, 12); }

// block-wide max of |w[i]|, i < n (every thread of the CTA calls it; `red` = 33 floats of shared memory)
__device__ __forceinline__ float cl_block_absmax(const float* __restrict__ w, int n, float* red) {
    float m = 0.f;
    for (int i = threadIdx.x; i < n; i += blockDim.x) m = fmaxf(m, fabsf(w[i]));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x < 32) {
        float v = (threadIdx.x < (blockDim.x + 31) / 32) ? red[threadIdx.x] : 0.f;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
        if (threadIdx.x == 0) red[32] = v;
    }
    __syncthreads();
    return red[32];
}

// Virtual (block-diagonal) image of the d x d score weight, w (D,D) [out][in], as the hi/lo planes of an operand whose
// 16-byte units hold 8 consecutive INPUT channels of one output channel:
//   element (vo, vi) at (vi/8) * W_IS + (vo/8) * 128 + (vo%8) * 16 + (vi%8) * 2,   v <-> real channel part*H + c of a sub-tile.
// Read K-major it is the A operand of S^T = Ws X^T (M = vo, K = vi); read MN-major that of dX^T = Ws^T dS^T (M = vi, K = vo).
template <int D>
__device__ __forceinline__ void cl_build_weight_image(const float* __restrict__ w, float scale, unsigned char* hi,
                                                      unsigned char* lo) {
    constexpr int W_IS = (kClLanes / 8) * 128;
    for (int e = threadIdx.x; e < kClLanes * kClLanes; e += blockDim.x) {
        const int vo = e / kClLanes, vi = e % kClLanes;
        float v = 0.f;
        const ClLane<D> lo_(vo), li_(vi);
        if (lo_.sub == li_.sub) v = w[lo_.channel() * D + li_.channel()];
        v *= scale;
        const __half hh = __float2half_rn(v);
        const __half ll = __float2half_rn(v - __half2float(hh));
        const int off = (vi >> 3) * W_IS + (vo >> 3) * 128 + (vo & 7) * 16 + (vi & 7) * 2;
        *reinterpret_cast<__half*>(hi + off) = hh;
        *reinterpret_cast<__half*>(lo + off) = ll;
    }
}
// Same for the h x h mlp_rpe2 weight, w (H,H) [out][in], on the 64 r lanes: a COMPACT 64 x 64 image (8 KB per plane),
//   element (vo, vi) at (vi/8) * 1024 + (vo/8) * 128 + (vo%8) * 16 + (vi%8) * 2,   vo, vi < 64.
// MMAs read it with M = 128: rows 64..127 alias whatever follows in shared memory and produce garbage in TMEM lanes
// 64..127, which nobody reads (the bytes behind the lo plane must be mapped shared memory).
constexpr int kClW2Is = 1024;
constexpr int kClW2Bytes = 8 * kClW2Is;     // one plane
template <int D>
__device__ __forceinline__ void cl_build_w2_image(const float* __restrict__ w, float scale, unsigned char* hi,
                                                  unsigned char* lo) {
    constexpr int H = D / 2;
    for (int e = threadIdx.x; e < 64 * 64; e += blockDim.x) {
        const int vo = e >> 6, vi = e & 63;
        float v = (vo / H == vi / H) ? w[(vo % H) * H + (vi % H)] * scale : 0.f;
        const __half hh = __float2half_rn(v);
        const __half ll = __float2half_rn(v - __half2float(hh));
        const int off = (vi >> 3) * kClW2Is + (vo >> 3) * 128 + (vo & 7) * 16 + (vi & 7) * 2;
        *reinterpret_cast<__half*>(hi + off) = hh;
        *reinterpret_cast<__half*>(lo + off) = ll;
    }
}

// ---------------------------------------------------------------------------------------------- row info
// Row-info table of a tile, per sub-tile: R rows of {p_j.xyz, |p_i - p_j|, feature offset of the neighbour (uint32 bits),
// gradient offset of the neighbour (uint32 bits; 0xffffffff marks a padding row past the last point), -, -} followed by
// the PTS points' {p_i.xyz, -}.  The relative position p_i - p_j is re-formed by the readers with the same single
// rounded subtraction as rpe_of_row, so the encoding stays bit-identical to the KNN's distances.
template <int D, int K>
__device__ __forceinline__ void cl_row_info(float* __restrict__ rinfo, const float* __restrict__ xyz, long long xyz_bstride,
                                            const int32_t* __restrict__ idx, long long feat_bstride,
                                            long long dfeat_bstride, int N, long long npts, long long tile, int l) {
    using C = ClCfg<D, K>;
    // Thread = row, IT rows per thread (4 at d = 16).  A row is a chain of two dependent global loads (neighbour index,
    // then its coordinates) that nothing else of the group overlaps, so the chains of a thread's rows are issued
    // together: all indices, then all coordinates, then the table stores (ncu source view of the d = 16 backward: 20 % of
    // the warp samples sat on the coordinate loads when the rows were taken one after the other).
    constexpr int IT = (C::ROWS + kClLanes - 1) / kClLanes;
    const bool small = npts <= 0x7fffffffLL;
    int pj[IT], pi[IT], bb[IT];
    bool valid[IT];
#pragma unroll
    for (int it = 0; it < IT; ++it) {
        const int row = l + it * kClLanes;
        const int sub = row / C::R, n = row % C::R;
        long long gp = tile * C::TPTS + sub * C::PTS + n / K;
        valid[it] = gp < npts;
        if (!valid[it]) gp = npts - 1;                       // padding rows recompute the last point (never written)
        bb[it] = small ? (int)((unsigned)gp / (unsigned)N) : (int)(gp / N);
        pi[it] = (int)(gp - (long long)bb[it] * N);
        pj[it] = (row < C::ROWS) ? idx[gp * K + n % K] : 0;
    }
    float ci[IT][3], cj[IT][3];
#pragma unroll
    for (int it = 0; it < IT; ++it) {
        const float* c = xyz + (size_t)bb[it] * xyz_bstride;
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            ci[it][q] = c[(size_t)pi[it] * 3 + q];
            cj[it][q] = c[(size_t)pj[it] * 3 + q];
        }
    }
#pragma unroll
    for (int it = 0; it < IT; ++it) {
        const int row = l + it * kClLanes;
        if (row >= C::ROWS) break;
        const int sub = row / C::R, n = row % C::R;
        const int p = n / K, k = n % K;
        // the KNN contract's rounding sequence (rpe_of_row): |p_i - p_j| equals sqrt of the search's d2 bit for bit
        const float dx = __fsub_rn(ci[it][0], cj[it][0]), dy = __fsub_rn(ci[it][1], cj[it][1]),
                    dz = __fsub_rn(ci[it][2], cj[it][2]);
        const float dist = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz)));
        float* base = rinfo + sub * C::SUB_RI;
        const uint32_t off = (uint32_t)((long long)bb[it] * feat_bstride + (long long)pj[it] * C::H);
        const uint32_t doff = valid[it] ? (uint32_t)((long long)bb[it] * dfeat_bstride + (long long)pj[it] * C::H) : 0xffffffffu;
        float4* dst = reinterpret_cast<float4*>(base + n * kClRinfo);
        dst[0] = make_float4(cj[it][0], cj[it][1], cj[it][2], dist);
        dst[1] = make_float4(__uint_as_float(off), __uint_as_float(doff), 0.f, 0.f);
        if (k == 0)
            *reinterpret_cast<float4*>(base + C::R * kClRinfo + p * 4) = make_float4(ci[it][0], ci[it][1], ci[it][2], 0.f);
    }
}

// the ten encoding channels of row n of a sub-tile, in the register layout cl_mlp1 takes
template <int D, int K>
__device__ __forceinline__ void cl_rpe_row(const float* __restrict__ ri, int n, float4& q0, float4& q1, float4& q2) {
    using C = ClCfg<D, K>;
    const float4 pt = *reinterpret_cast<const float4*>(ri + C::R * kClRinfo + (n / K) * 4);
    const float4 rw = *reinterpret_cast<const float4*>(ri + n * kClRinfo);
    q0 = make_float4(pt.x, pt.y, pt.z, rw.x);
    q1 = make_float4(rw.y, rw.z, __fsub_rn(pt.x, rw.x), __fsub_rn(pt.y, rw.y));
    q2 = make_float4(__fsub_rn(pt.z, rw.z), rw.w, 0.f, 0.f);
}
// neighbour offsets of row n: feature row (element offset), gradient row (0xffffffff: padding row)
__device__ __forceinline__ uint32_t cl_feat_off(const float* __restrict__ ri, int n) { return __float_as_uint(ri[n * kClRinfo + 4]); }
__device__ __forceinline__ uint32_t cl_grad_off(const float* __restrict__ ri, int n) { return __float_as_uint(ri[n * kClRinfo + 5]); }

// r1 = relu(a1 (W1 . rpe) + b1) for one row, W1 row in registers (same FMA order as rpe_mlp1 of lfa_common.cuh)
__device__ __forceinline__ float cl_mlp1(const float (&w)[10], float a1, float b1, const float4& q0, const float4& q1,
                                         const float4& q2) {
    float z = w[0] * q0.x;
    z = fmaf(w[1], q0.y, z); z = fmaf(w[2], q0.z, z); z = fmaf(w[3], q0.w, z);
    z = fmaf(w[4], q1.x, z); z = fmaf(w[5], q1.y, z); z = fmaf(w[6], q1.z, z);
    z = fmaf(w[7], q1.w, z); z = fmaf(w[8], q2.x, z); z = fmaf(w[9], q2.y, z);
    return fmaxf(fmaf(z, a1, b1), 0.f);
}

// 8 scaled values of one channel (rows ng*8 .. ng*8+7) -> the hi/lo planes of a row operand
__device__ __forceinline__ void cl_store_unit(unsigned char* hi, unsigned char* lo, int unit_off, const float (&v)[8]) {
    uint4 h, l;
    split16_2(v[0], v[1], h.x, l.x);
    split16_2(v[2], v[3], h.y, l.y);
    split16_2(v[4], v[5], h.z, l.z);
    split16_2(v[6], v[7], h.w, l.w);
    *reinterpret_cast<uint4*>(hi + unit_off) = h;
    *reinterpret_cast<uint4*>(lo + unit_off) = l;
}
__device__ __forceinline__ void cl_load_unit(const unsigned char* hi, const unsigned char* lo, int unit_off, float (&v)[8]) {
    const uint4 h = *reinterpret_cast<const uint4*>(hi + unit_off);
    const uint4 l = *reinterpret_cast<const uint4*>(lo + unit_off);
    float2 t;
    t = join16_2(h.x, l.x); v[0] = t.x; v[1] = t.y;
    t = join16_2(h.y, l.y); v[2] = t.x; v[3] = t.y;
    t = join16_2(h.z, l.z); v[4] = t.x; v[5] = t.y;
    t = join16_2(h.w, l.w); v[6] = t.x; v[7] = t.y;
}

// byte offset of the unit (channel lane l, row group ng) inside a row-operand plane
template <int D, int K>
__device__ __forceinline__ int cl_unit_off(int l, int ng) {
    return (l >> 3) * ClCfg<D, K>::OP_CS + ng * 128 + (l & 7) * 16;
}

// --------------------------------------------------------------------------------------------- MMA issue
// D[tmem] (+)= A B^T over `ksteps` K steps of 16, three split products per step.
//   a_*: hi/lo plane addresses of the A operand, a_lbo / a_sbo its core-matrix strides along K / along M
//   b_*: likewise for B.   The per-step advance along K is two core matrices: 2 * lbo.
template <int UNROLL = 2>
__device__ __forceinline__ void cl_mma_3x(uint32_t tmem_d, uint32_t a_hi, uint32_t a_lo, uint32_t a_lbo, uint32_t a_sbo,
                                          uint32_t b_hi, uint32_t b_lo, uint32_t b_lbo, uint32_t b_sbo, uint32_t idesc,
                                          int ksteps, bool accumulate) {
    // The issuer's instruction stream is on the critical path of every group (ncu: a third of the workers' time is spent
    // waiting for the tensor core) and, fully unrolled with all five call sites of the backward, it is ~30 KB of SASS
    // competing with the worker groups for the instruction cache.  So: descriptors are built once and advanced by adding
    // the K-step increment to their 14-bit address field (shared-memory addresses stay below 2^18, no carry into the
    // stride fields), and the loop is unrolled by two only (UNROLL: kernels with a single call site take four).
    uint64_t ah = umma_desc_b(a_hi, a_lbo, a_sbo), al = umma_desc_b(a_lo, a_lbo, a_sbo);
    uint64_t bh = umma_desc_b(b_hi, b_lbo, b_sbo), bl = umma_desc_b(b_lo, b_lbo, b_sbo);
    const uint64_t ia = (uint64_t)((2 * a_lbo) >> 4), ib = (uint64_t)((2 * b_lbo) >> 4);
#pragma unroll UNROLL
    for (int ks = 0; ks < ksteps; ++ks) {
        umma_f16(tmem_d, ah, bh, idesc, (accumulate || ks > 0) ? 1u : 0u);
        umma_f16(tmem_d, ah, bl, idesc, 1u);
        umma_f16(tmem_d, al, bh, idesc, 1u);
        ah += ia;
        al += ia;
        bh += ib;
        bl += ib;
    }
}

}  // namespace r3d
