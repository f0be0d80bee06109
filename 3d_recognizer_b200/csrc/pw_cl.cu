// Per-point layers of the training step on the tcgen05 tensor cores, split-fp16 operands (tc16_common.cuh): the
// forward / input-gradient GEMM and the weight-gradient row reduction of SharedMLP (randlanet/utils/modules.py:60-104 as
// driven by trainer.py:115-119) for the mid-width layers, where the FP32 CUDA-core GEMMs of pointwise.cu /
// pointwise_train.cu ran at 17-35 TFLOP/s (profiles/bench_train40960_n1.json).
//
// One operand tile for every role, straight from row-major memory: 64 rows x C channels as 16-byte units of 8 consecutive
// CHANNELS of one row, 8 rows adjacent (128-byte core matrix), channel groups 1 KB apart.
//   * read K-major (K = channels) it is the B operand of   Y^T (out x rows)  = W (out x in) . X^T        [pc_gemm_kernel]
//     (TMEM lane = output channel: per-channel affine / activation / BatchNorm batch sums are in-thread, and a warp's
//     store of one row is 128 contiguous bytes);
//   * read MN-major (MN = channels, K = rows) two such tiles are both operands of   dW (p x q) = P^T Q   [pc_wgrad_kernel].
// Persistent CTAs, warp-specialised: 8 loader warps (global fp32 -> scale -> hi/lo fp16 -> shared memory; they also run
// the epilogues) and one MMA warp, mbarrier rings between them.
//
// Scales (powers of two).  pc_gemm: one per ROW, from the row's own absmax (a row scale multiplies one accumulator
// column and is divided out in the epilogue), weights per CTA from their absmax.  pc_wgrad contracts over rows, so each
// operand carries ONE scale, from an absmax the caller provides (written by the kernel that produced the tensor, or by
// r3d_absmax); an absmax that is too large by a factor up to 2^10 costs no accuracy (fp16x2 keeps 22 bits down to 2^-18
// of the scaled maximum).
// The tensor core truncates when it adds to its accumulator (~1 ulp per MMA, tools/tc16_probe_test.py): pc_wgrad sums
// at most kPcFold stages (24 MMAs) in a first-level accumulator and folds it with round-to-nearest fp32 adds into a
// second-level one (TMEM -> registers -> TMEM), like the weight-gradient sums of lfa_cl_bwd.cu.
#include "lfa_cl_common.cuh"

namespace r3d {

constexpr int kPcRows = 64;                    // rows per stage
constexpr int kPcLoaders = 256;                // 8 loader / epilogue warps
constexpr int kPcThreads = kPcLoaders + 32;    // + the MMA warp
constexpr int kPcCS = (kPcRows / 8) * 128;     // byte stride between channel groups of a stage plane
constexpr int kPcWIS = (128 / 8) * 128;        // byte stride between input-channel groups of a weight plane
constexpr int kPcFold = 2;                     // stages per first-level accumulator of pc_wgrad (24 MMAs: deficit <= 7e-7)

__device__ __forceinline__ void pc_split_store(unsigned char* hi, unsigned char* lo, int off, const float4& a, const float4& b,
                                               float s) {
    uint4 h, l;
    split16_2(a.x * s, a.y * s, h.x, l.x);
    split16_2(a.z * s, a.w * s, h.y, l.y);
    split16_2(b.x * s, b.y * s, h.z, l.z);
    split16_2(b.z * s, b.w * s, h.w, l.w);
    *reinterpret_cast<uint4*>(hi + off) = h;
    *reinterpret_cast<uint4*>(lo + off) = l;
}
__device__ __forceinline__ float pc_absmax8(const float4& a, const float4& b) {
    return fmaxf(fmaxf(fmaxf(fabsf(a.x), fabsf(a.y)), fmaxf(fabsf(a.z), fabsf(a.w))),
                 fmaxf(fmaxf(fabsf(b.x), fabsf(b.y)), fmaxf(fabsf(b.z), fabsf(b.w))));
}
__device__ __forceinline__ void pc_red_add(float* p, float v) {
    asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}
__device__ __forceinline__ void pc_atomic_absmax(float* slot, float v) {      // v >= 0: integer order = float order
    if (slot && v > 0.f) atomicMax(reinterpret_cast<int*>(slot), __float_as_int(v));
}

// =====================================================================================================================
// Y (M, cout) = epilogue(X (M, cin) . W^T): forward (W = weight (cout,cin)) and input gradient (W = weight^T)
// =====================================================================================================================
struct PcGemmArgs {
    const float* x;          // (M, cin) rows, ldx floats apart
    long long ldx;
    const float* w;          // element (out o, in i) at w[o * w_so + i * w_si]
    long long w_so, w_si;
    const float* scale;      // (cout) nullable
    const float* shift;      // (cout) nullable
    int act;                 // 0 none, 1 relu, 2 leaky relu
    float slope;
    float* y;                // (M, cout) rows, ldy floats apart
    long long ldy;
    double* stats;           // (2 cout) += sum y, sum y^2 (before scale / shift / act: the conv output), nullable
    float* absmax_x;         // += max |x| (atomic max), nullable
    int accumulate;          // 1: y holds the partial sum of earlier input-channel blocks; add it before the epilogue
    long long M;
    int cin, cout;
    long long nstages;       // ceil(M / 64)
};

constexpr int kPcGemmStages = 4;

// shared memory, sized by cin (two CTAs per SM up to cin = 64): W hi | W lo | stages (hi, lo, per-row 1/scale) | barriers
struct PcGemmSmem {
    int x_bytes, stage, off_st, off_bars;
    size_t bytes;
    __host__ __device__ explicit PcGemmSmem(int cin) {
        x_bytes = (cin >> 3) * kPcCS;
        stage = 2 * x_bytes + kPcRows * 4;
        off_st = 2 * (cin >> 3) * kPcWIS;
        off_bars = off_st + kPcGemmStages * stage;
        bytes = (size_t)off_bars + 4 * kPcGemmStages * 8 + 16 + 34 * 4;
    }
};

// NSETS sets of 8 loader warps take the CTA's stages alternately (NSETS = 2 where shared memory allows one CTA per SM only:
// the chain load -> convert -> store -> epilogue of a warp is latency-bound, more warps in flight is what hides it)
template <int NSETS>
__global__ void __launch_bounds__(NSETS * kPcLoaders + 32, NSETS == 1 ? 2 : 1) pc_gemm_kernel(PcGemmArgs a) {
    constexpr int NT = NSETS * kPcLoaders + 32, MMA_WARP = NSETS * 8;
    const PcGemmSmem L(a.cin);
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* Whi = smem;
    unsigned char* Wlo = smem + L.off_st / 2;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + L.off_bars);
    uint64_t* empty = full + kPcGemmStages;
    uint64_t* accfull = empty + kPcGemmStages;
    uint64_t* accfree = accfull + kPcGemmStages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accfree + kPcGemmStages);
    float* red = reinterpret_cast<float*>(tmem_slot + 4);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int o0 = blockIdx.y * 128;
    const int nout = min(128, a.cout - o0);
    const int chunks = a.cin >> 3;

    if (tid == 0) {
        for (int s = 0; s < kPcGemmStages; ++s) {
            mbar_init(&full[s], 8);
            mbar_init(&empty[s], 1);
            mbar_init(&accfull[s], 1);
            mbar_init(&accfree[s], 8);
        }
        mbar_fence_init();
    }
    if (warp == MMA_WARP) tmem_alloc_warp(tmem_slot, 256);
    // weight block -> absmax -> K-major image: element (o, i) at (i/8) W_IS + (o/8) 128 + (o%8) 16 + (i%8) 2
    float wmax = 0.f;
    for (int e = tid; e < nout * a.cin; e += NT) {
        const int o = e / a.cin, i = e % a.cin;
        wmax = fmaxf(wmax, fabsf(a.w[(o0 + o) * a.w_so + i * a.w_si]));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) wmax = fmaxf(wmax, __shfl_xor_sync(0xffffffffu, wmax, o));
    if (lane == 0) red[warp] = wmax;
    __syncthreads();
    wmax = 0.f;
    for (int i = 0; i < NT / 32; ++i) wmax = fmaxf(wmax, red[i]);
    const float sw = cl_pow2_scale(wmax);
    for (int e = tid; e < 128 * a.cin; e += NT) {
        const int o = e / a.cin, i = e % a.cin;
        const float v = o < nout ? a.w[(o0 + o) * a.w_so + i * a.w_si] * sw : 0.f;
        const __half hh = __float2half_rn(v);
        const __half ll = __float2half_rn(v - __half2float(hh));
        const int off = (i >> 3) * kPcWIS + (o >> 3) * 128 + (o & 7) * 16 + (i & 7) * 2;
        *reinterpret_cast<__half*>(Whi + off) = hh;
        *reinterpret_cast<__half*>(Wlo + off) = ll;
    }
    fence_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem = *tmem_slot;
    // stages of this CTA: blockIdx.x, blockIdx.x + gridDim.x, ...
    const long long first = blockIdx.x, stride = gridDim.x;
    const long long mine = first < a.nstages ? (a.nstages - first + stride - 1) / stride : 0;

    if (warp == MMA_WARP) {
        // ================================================================ MMA issuer
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_f16(128, kPcRows, 0, 0);
            const int ksteps = a.cin >> 4;
            for (long long it = 0; it < mine; ++it) {
                const int s = (int)(it % kPcGemmStages);
                const uint32_t ph = (uint32_t)((it / kPcGemmStages) & 1);
                mbar_wait(&accfree[s], ph ^ 1u);          // epilogue of the previous use of this accumulator is done
                mbar_wait(&full[s], ph);
                tc_fence_after_sync();
                const uint32_t xhi = smem_u32(smem + L.off_st + s * L.stage);
                cl_mma_3x<4>(tmem + s * kPcRows, smem_u32(Whi), smem_u32(Wlo), kPcWIS, 128, xhi, xhi + L.x_bytes, kPcCS, 128,
                          idesc, ksteps, false);
                umma_commit(&empty[s]);
                umma_commit(&accfull[s]);
            }
        }
    } else {
        // ================================================================ loaders + epilogue
        const int r8 = lane & 7, cq = lane >> 3;
        const int set = warp >> 3, wr = warp & 7;            // loader set, row group inside a stage
        const int q = wr & 3, hh = wr >> 2;                  // epilogue: TMEM lane quarter, column half
        const int oc = o0 + q * 32 + lane;                   // this thread's output channel
        const bool oc_ok = oc < a.cout;
        const float e_scale = (a.scale && oc_ok) ? a.scale[oc] : 1.f;
        const float e_shift = (a.shift && oc_ok) ? a.shift[oc] : 0.f;
        const float inv_sw = 1.0f / sw;
        double s1 = 0.0, s2 = 0.0;
        float xmax = 0.f;

        auto epilogue = [&](long long it) {
            const int s = (int)(it % kPcGemmStages);
            const uint32_t ph = (uint32_t)((it / kPcGemmStages) & 1);
            const long long row0 = (first + it * stride) * kPcRows + hh * 32;
            const float* inv_row = reinterpret_cast<const float*>(smem + L.off_st + s * L.stage + 2 * L.x_bytes) + hh * 32;
            mbar_wait(&accfull[s], ph);
            tc_fence_after_sync();
            uint32_t r0[16], r1[16];
            const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + s * kPcRows + hh * 32;
            tmem_ld16_nowait(taddr, r0);
            tmem_ld16_nowait(taddr + 16, r1);
            tmem_ld_wait();
            float ps1 = 0.f, ps2 = 0.f;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                const long long row = row0 + j;
                float v = __uint_as_float(j < 16 ? r0[j & 15] : r1[j & 15]) * (inv_row[j] * inv_sw);
                if (row < a.M && oc_ok) {
                    if (a.accumulate) v += a.y[row * a.ldy + oc];
                    ps1 += v;
                    ps2 = fmaf(v, v, ps2);
                    v = fmaf(v, e_scale, e_shift);
                    if (a.act == 1) v = fmaxf(v, 0.f);
                    if (a.act == 2) v = v > 0.f ? v : v * a.slope;
                    a.y[row * a.ldy + oc] = v;
                }
            }
            s1 += (double)ps1;
            s2 += (double)ps2;
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(&accfree[s]);
        };

        // this thread's units of its row (row group = warp): chunks cq, cq + 4, cq + 8, cq + 12.  The loads of stage
        // it + 1 are issued right after stage it is handed to the MMA warp, so they fly during the epilogue
        float4 v[4][2];
        auto load_stage = [&](long long it) {
            const long long row = (first + it * stride) * kPcRows + wr * 8 + r8;
            const float* src = a.x + row * a.ldx;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int c = k * 4 + cq;
                if (c < chunks && row < a.M) {
                    v[k][0] = __ldg(reinterpret_cast<const float4*>(src + c * 8));
                    v[k][1] = __ldg(reinterpret_cast<const float4*>(src + c * 8 + 4));
                } else {
                    v[k][0] = v[k][1] = make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
        };
        if (set < mine) load_stage(set);
        long long it = set;
        for (; it < mine; it += NSETS) {
            const int s = (int)(it % kPcGemmStages);
            const uint32_t ph = (uint32_t)((it / kPcGemmStages) & 1);
            unsigned char* xhi = smem + L.off_st + s * L.stage;
            unsigned char* xlo = xhi + L.x_bytes;
            float* inv_row = reinterpret_cast<float*>(xlo + L.x_bytes);
            float m = 0.f;
#pragma unroll
            for (int k = 0; k < 4; ++k) m = fmaxf(m, pc_absmax8(v[k][0], v[k][1]));
            m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 8));
            m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 16));          // the row's absmax, on its four lanes
            xmax = fmaxf(xmax, m);
            const float sx = cl_pow2_scale(m);
            // the stage slot is free once the MMAs that read it are done AND its previous epilogue has read the row scales
            mbar_wait(&empty[s], ph ^ 1u);
            mbar_wait(&accfree[s], ph ^ 1u);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int c = k * 4 + cq;
                if (c < chunks) pc_split_store(xhi, xlo, c * kPcCS + wr * 128 + r8 * 16, v[k][0], v[k][1], sx);
            }
            if (cq == 0) inv_row[wr * 8 + r8] = 1.0f / sx;
            fence_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&full[s]);
            if (it + NSETS < mine) load_stage(it + NSETS);
            if (it >= 2) epilogue(it - 2);           // NSETS = 1: two stages behind; NSETS = 2: this set's previous stage
        }
        // `it` = this set's first stage beyond the end: the stages whose epilogue is still to run are it - 2 (both set
        // counts) and, for one set, it - 1
        if (it >= 2 && it - 2 < mine) epilogue(it - 2);
        if (NSETS == 1 && it >= 1 && it - 1 < mine) epilogue(it - 1);

        if (a.stats && oc_ok) {
            atomicAdd(&a.stats[oc], s1);
            atomicAdd(&a.stats[a.cout + oc], s2);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) xmax = fmaxf(xmax, __shfl_xor_sync(0xffffffffu, xmax, o));
        if (lane == 0 && blockIdx.y == 0) pc_atomic_absmax(a.absmax_x, xmax);
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == MMA_WARP) tmem_dealloc_warp(tmem, 256);
}

// =====================================================================================================================
// out (cp, cq) += P^T Q over M rows
// =====================================================================================================================
struct PcWgradArgs {
    const float* p;          // (M, cp) rows, ldp floats apart: the operand on the TMEM lanes
    long long ldp;
    const float* q;          // (M, cq) rows, ldq floats apart: the operand on the TMEM columns
    long long ldq;
    const float* absmax_p;   // device scalars: upper bounds of |p|, |q|
    const float* absmax_q;
    float* out;              // element (i, j) of P^T Q at out[i * so_p + j * so_q], +=
    long long so_p, so_q;
    long long M;
    int cp, cq;
    int tiles_q;             // column blocks of 128
    long long rows_per_slice;     // multiple of 64
};

constexpr int kPcWgradStages = 3;

struct PcWgradSmem {
    static constexpr int PLANE = 16 * kPcCS;                          // 128 channels x 64 rows fp16
    static constexpr int STAGE = 4 * PLANE;                           // P hi, P lo, Q hi, Q lo
    static constexpr int OFF_BARS = kPcWgradStages * STAGE;
    static constexpr size_t BYTES = (size_t)OFF_BARS + (2 * kPcWgradStages + 4) * 8 + 16;
};

__global__ void __launch_bounds__(kPcThreads, 1) pc_wgrad_kernel(PcWgradArgs a) {
    using S = PcWgradSmem;
    extern __shared__ __align__(1024) unsigned char smem[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + S::OFF_BARS);
    uint64_t* empty = full + kPcWgradStages;
    uint64_t* accfull = empty + kPcWgradStages;      // [2]
    uint64_t* accfree = accfull + 2;                 // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accfree + 2);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int p0 = (blockIdx.y / a.tiles_q) * 128, q0 = (blockIdx.y % a.tiles_q) * 128;
    const int np = min(128, a.cp - p0), nq = min(128, a.cq - q0);
    const int chunks_p = np >> 3, chunks_q = nq >> 3;
    const int NQ = (nq + 31) & ~31;                  // MMA N (a multiple of 32: two column halves of whole 16-column groups)
    const long long row_begin = (long long)blockIdx.x * a.rows_per_slice;
    const long long row_end = min(a.M, row_begin + a.rows_per_slice);
    const long long nst = row_begin < row_end ? (row_end - row_begin + kPcRows - 1) / kPcRows : 0;
    const long long nwin = (nst + kPcFold - 1) / kPcFold;

    if (tid == 0) {
        for (int s = 0; s < kPcWgradStages; ++s) {
            mbar_init(&full[s], 8);
            mbar_init(&empty[s], 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&accfull[b], 1);
            mbar_init(&accfree[b], 8);
        }
        mbar_fence_init();
    }
    if (warp == 8) tmem_alloc_warp(tmem_slot, 512);
    // channel groups beyond the operands' widths stay zero for the whole kernel
    for (int i = tid; i < kPcWgradStages * S::STAGE / 16; i += kPcThreads) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    fence_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem = *tmem_slot;
    const float sp = cl_pow2_scale(*a.absmax_p), sq = cl_pow2_scale(*a.absmax_q);

    if (warp == 8) {
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_f16(128, NQ, 1, 1);
            for (long long it = 0; it < nst; ++it) {
                const int s = (int)(it % kPcWgradStages);
                const uint32_t ph = (uint32_t)((it / kPcWgradStages) & 1);
                const long long win = it / kPcFold;
                const int buf = (int)(win & 1);
                if (it % kPcFold == 0) {
                    mbar_wait(&accfree[buf], (uint32_t)(((win >> 1) & 1) ^ 1));
                    tc_fence_after_sync();
                }
                mbar_wait(&full[s], ph);
                tc_fence_after_sync();
                const uint32_t base = smem_u32(smem + s * S::STAGE);
                // contraction over the stage's 64 rows: K-step = 2 row groups (256 B), LBO = row-group stride, SBO = channel-group stride
                cl_mma_3x<4>(tmem + buf * 128, base, base + S::PLANE, 128, kPcCS, base + 2 * S::PLANE, base + 3 * S::PLANE, 128, kPcCS,
                          idesc, kPcRows / 16, it % kPcFold != 0);
                umma_commit(&empty[s]);
                if (it % kPcFold == kPcFold - 1 || it == nst - 1) umma_commit(&accfull[buf]);
            }
        }
    } else {
        const int r8 = lane & 7, cq = lane >> 3;
        const int qd = warp & 3, hh = warp >> 2;
        const int ncols = NQ / 2;                            // columns of this warp's half (multiple of 16)
        const uint32_t lane_base = tmem + ((uint32_t)(qd * 32) << 16);

        // fold window `win` (first-level accumulator win & 1) into the second level at columns [256, 256 + NQ); a warp
        // folds, and at the end stores, its own (lane quarter, column half) only
        auto fold = [&](long long win) {
            const int buf = (int)(win & 1);
            mbar_wait(&accfull[buf], (uint32_t)((win >> 1) & 1));
            tc_fence_after_sync();
            for (int c = hh * ncols; c < (hh + 1) * ncols; c += 16) {
                uint32_t x[16], y[16];
                tmem_ld16_nowait(lane_base + buf * 128 + c, x);
                if (win > 0) tmem_ld16_nowait(lane_base + 256 + c, y);
                tmem_ld_wait();
                if (win > 0) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) x[j] = __float_as_uint(__uint_as_float(x[j]) + __uint_as_float(y[j]));
                }
                tmem_st16(lane_base + 256 + c, x);
            }
            tmem_st_wait();
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(&accfree[buf]);
        };

        float4 vp[4][2], vq[4][2];
        auto load_stage = [&](long long it) {
            const long long row = row_begin + it * kPcRows + warp * 8 + r8;
            const bool live = row < row_end;
            const float* ps = a.p + row * a.ldp + p0;
            const float* qs = a.q + row * a.ldq + q0;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int c = k * 4 + cq;
                const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
                vp[k][0] = (c < chunks_p && live) ? __ldg(reinterpret_cast<const float4*>(ps + c * 8)) : z;
                vp[k][1] = (c < chunks_p && live) ? __ldg(reinterpret_cast<const float4*>(ps + c * 8 + 4)) : z;
                vq[k][0] = (c < chunks_q && live) ? __ldg(reinterpret_cast<const float4*>(qs + c * 8)) : z;
                vq[k][1] = (c < chunks_q && live) ? __ldg(reinterpret_cast<const float4*>(qs + c * 8 + 4)) : z;
            }
        };
        if (nst > 0) load_stage(0);
        for (long long it = 0; it < nst; ++it) {
            const int s = (int)(it % kPcWgradStages);
            const uint32_t ph = (uint32_t)((it / kPcWgradStages) & 1);
            unsigned char* base = smem + s * S::STAGE;
            mbar_wait(&empty[s], ph ^ 1u);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int c = k * 4 + cq;
                const int off = c * kPcCS + warp * 128 + r8 * 16;
                if (c < chunks_p) pc_split_store(base, base + S::PLANE, off, vp[k][0], vp[k][1], sp);
                if (c < chunks_q) pc_split_store(base + 2 * S::PLANE, base + 3 * S::PLANE, off, vq[k][0], vq[k][1], sq);
            }
            fence_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&full[s]);
            if (it + 1 < nst) load_stage(it + 1);         // in flight during the fold and the wait for the next slot
            // fold the window before the previous one's successor starts: window w - 1 once stage 1 of window w is produced
            if (it % kPcFold == 1 && it / kPcFold >= 1) fold(it / kPcFold - 1);
        }
        // windows not folded inside the loop (window j is folded there when stage (j + 1) kPcFold + 1 exists)
        for (long long w = (nst >= 2 ? (nst - 2) / kPcFold : 0); w < nwin; ++w) fold(w);
        if (nst > 0) {
            const float unscale = 1.0f / (sp * sq);
            const int i = p0 + qd * 32 + lane;
            for (int c = hh * ncols; c < (hh + 1) * ncols; c += 16) {
                uint32_t x[16];
                tmem_ld16_nowait(lane_base + 256 + c, x);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const float v = __uint_as_float(x[j]) * unscale;
                    if (i < a.cp && c + j < nq && v != 0.f)
                        pc_red_add(a.out + (long long)i * a.so_p + (long long)(q0 + c + j) * a.so_q, v);
                }
            }
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 8) tmem_dealloc_warp(tmem, 512);
}

}  // namespace r3d

using namespace r3d;

extern "C" int r3d_pc_gemm_supported(int cin, int cout, long long M) {
    return cin >= 16 && cin <= 4096 && cin % 16 == 0 && cout >= 8 && M >= 1;
}

// one launch per block of <= 128 input channels (the weight image of a block stays in shared memory); later blocks add
// to the partial sums in y, the last one applies statistics / affine / activation
extern "C" int r3d_pc_gemm(const float* x, long long ldx, const float* w, long long w_so, long long w_si, const float* scale,
                           const float* shift, int act, float slope, float* y, long long ldy, double* stats,
                           float* absmax_x, long long M, int cin, int cout, r3d_stream_t stream) {
    if (M < 0 || cin <= 0 || cout <= 0 || act < 0 || act > 2) return R3D_EINVAL;
    if (!r3d_pc_gemm_supported(cin, cout, M > 0 ? M : 1)) return R3D_EUNSUPPORTED;
    if (M == 0) return R3D_OK;
    if (!x || !w || !y) return R3D_EINVAL;
    if (ldx == 0) ldx = cin;
    if (ldy == 0) ldy = cout;
    if (!is_aligned(x, 16) || ldx % 4 != 0) return R3D_EALIGN;
    const int tiles_o = (cout + 127) / 128;
    for (int k0 = 0; k0 < cin; k0 += 128) {
        const int kc = cin - k0 < 128 ? cin - k0 : 128;
        const bool last = k0 + kc >= cin;
        PcGemmArgs a{x + k0, ldx, w + (long long)k0 * w_si, w_so, w_si, last ? scale : nullptr, last ? shift : nullptr,
                     last ? act : 0, slope, y, ldy, last ? stats : nullptr, absmax_x, k0 > 0 ? 1 : 0, M, kc, cout,
                     (M + kPcRows - 1) / kPcRows};
        const PcGemmSmem L(kc);
        const int per_sm = L.bytes <= 110 * 1024 ? 2 : 1;      // two CTAs of 8 loader warps, or one of 16
        long long gx = (long long)per_sm * kNumSMs / tiles_o;
        if (gx < 1) gx = 1;
        if (gx > a.nstages) gx = a.nstages;
        if (per_sm == 2) {
            R3D_CUDA_TRY(cudaFuncSetAttribute(pc_gemm_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PcGemmSmem(128).bytes));
            pc_gemm_kernel<1><<<dim3((unsigned)gx, (unsigned)tiles_o), kPcLoaders + 32, L.bytes, static_cast<cudaStream_t>(stream)>>>(a);
        } else {
            R3D_CUDA_TRY(cudaFuncSetAttribute(pc_gemm_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PcGemmSmem(128).bytes));
            pc_gemm_kernel<2><<<dim3((unsigned)gx, (unsigned)tiles_o), 2 * kPcLoaders + 32, L.bytes, static_cast<cudaStream_t>(stream)>>>(a);
        }
        R3D_LAUNCH_CHECK("pc_gemm_kernel");
    }
    return R3D_OK;
}

extern "C" int r3d_pc_wgrad_supported(int ca, int cb, long long M) {
    return ca % 8 == 0 && cb % 8 == 0 && ca >= 8 && cb >= 8 && M >= 1;
}

// out (ca, cb; ld_out; caller-zeroed) += A^T B like r3d_rowreduce_gemm; absmax_a / absmax_b: device scalars >= max |A|, |B|
extern "C" int r3d_pc_wgrad(const float* A, long long lda, int ca, const float* Bm, long long ldb, int cb, long long M,
                            const float* absmax_a, const float* absmax_b, float* out, int ld_out, r3d_stream_t stream) {
    if (M < 0 || ca <= 0 || cb <= 0) return R3D_EINVAL;
    if (!r3d_pc_wgrad_supported(ca, cb, M > 0 ? M : 1)) return R3D_EUNSUPPORTED;
    if (M == 0) return R3D_OK;
    if (!A || !Bm || !out || !absmax_a || !absmax_b) return R3D_EINVAL;
    if (lda == 0) lda = ca;
    if (ldb == 0) ldb = cb;
    if (!is_aligned(A, 16) || !is_aligned(Bm, 16) || lda % 4 != 0 || ldb % 4 != 0) return R3D_EALIGN;
    // the wider operand goes on the 128 TMEM lanes (fewer zero-padded lanes), the other on the columns
    const bool swap = cb > ca;
    PcWgradArgs a;
    a.p = swap ? Bm : A, a.ldp = swap ? ldb : lda, a.cp = swap ? cb : ca;
    a.q = swap ? A : Bm, a.ldq = swap ? lda : ldb, a.cq = swap ? ca : cb;
    a.absmax_p = swap ? absmax_b : absmax_a, a.absmax_q = swap ? absmax_a : absmax_b;
    a.out = out;
    a.so_p = swap ? 1 : ld_out, a.so_q = swap ? ld_out : 1;
    a.M = M;
    const int tiles_p = (a.cp + 127) / 128;
    a.tiles_q = (a.cq + 127) / 128;
    const int tiles = tiles_p * a.tiles_q;
    long long slices = kNumSMs / tiles;
    if (slices < 1) slices = 1;
    const long long stages = (M + kPcRows - 1) / kPcRows;
    if (slices > stages) slices = stages;
    a.rows_per_slice = (stages + slices - 1) / slices * kPcRows;
    slices = (M + a.rows_per_slice - 1) / a.rows_per_slice;
    R3D_CUDA_TRY(cudaFuncSetAttribute(pc_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PcWgradSmem::BYTES));
    pc_wgrad_kernel<<<dim3((unsigned)slices, (unsigned)tiles), kPcThreads, PcWgradSmem::BYTES, static_cast<cudaStream_t>(stream)>>>(a);
    R3D_LAUNCH_CHECK("pc_wgrad_kernel");
    return R3D_OK;
}
