// Tensor-core fused LocSE + attentive pooling for the WIDEST level, d = 256 (h = 128): forward and the two backward
// kernels that dominate its cost.  Same channel-lane design as lfa_cl.cu / lfa_cl_bwd.cu (lfa_cl_common.cuh), adapted
// to what one SM can hold at this width:
//   * 256 channels do not fit 128 TMEM lanes, and dWs (256 x 256 fp32) alone would fill all 512 TMEM columns.  So a CTA
//     owns one HALF of the output channels (blockIdx.y): its lanes are 128 score channels o, its weight image is
//     Ws[o-half][all 256 inputs] (128 KB as split-fp16 planes, resident), its sum accumulator dWs[o-half][:] is 256
//     columns.  S^T[o-half] = Ws[o-half] X^T needs the FULL row operand (both CTAs of a pair build it), dS^T exists for
//     the own half only, and dX^T = Ws^T dS^T is a PARTIAL sum over the own half of o: everything downstream of dX
//     (ReLU mask, G1 sums, BatchNorm sums, neighbour scatter) is linear, so both halves simply add their parts.
//   * rows per tile R = 32 (MMA N = 32): X^T planes 32 KB + dS^T planes 16 KB next to the 128 KB of weights.
//   * thread l produces r channel l AND F channel l of the row operand, owns score channel half*128 + l in the
//     softmax epilogue, and in the dX epilogue reduces r channel l (mask, G1 / BatchNorm sums) and scatters F channel l.
//   * stage 2 (r2 = relu(a2 (W2 r1) + c2)) does not recompute mlp_rpe2 here — its 128 x 128 weight image does not fit
//     as well — but reads r2 rows from `rmat` (rows x 128, written once per step by r3d_lfa_r1_rows + a per-point
//     layer; 335 MB at the config-D size of this level, where rows are few).
//   * the sum accumulator has a single TMEM level; it is flushed to global memory (vector atomics) every kWideFlush
//     tiles to bound the tensor core's accumulate-truncation error (see lfa_cl_bwd.cu).
// MODE 0 forward, MODE 1 backward of a stage-1 launch, MODE 2 pass 1 of the train-mode backward of a stage-2 launch
// (du2 partial per half -> r3d_lfa_du2_combine adds the halves into the layout r3d_lfa_bn2_bwd reads).
#include "lfa_cl_common.cuh"

namespace r3d {

constexpr int kWD = 256, kWH = 128, kWR = 32;
constexpr int kWOpCs = (kWR / 8) * 128;                 // 512: byte stride between channel groups of a row operand
constexpr int kWXBytes = kWD * kWR * 2;                 // one plane of X^T: 16 KB
constexpr int kWSBytes = kWH * kWR * 2;                 // one plane of dS^T: 8 KB
constexpr int kWWBytes = kWH * kWD * 2;                 // one plane of the weight image: 64 KB
constexpr int kWWIs = 16 * 128;                         // byte stride between input-channel groups of the weight image
constexpr int kWRow = 12;                                // floats per row of this kernel's row table: rpe[10], two offsets
constexpr int kWRinfo = kWR * kWRow;                    // floats
constexpr int kWideFlush = 32;

struct LfaWideArgs {
    const float* xyz;
    long long xyz_bstride;
    const int32_t* idx;
    const float* feat;
    long long feat_bstride;
    const float* w_rpe1;
    const float* a_rpe1;
    const float* b_rpe1;
    const float* rmat;        // (B*N*K, 128) r2 rows, stage 2; nullptr: r1 = mlp_rpe1 evaluated here
    const float* w_score;     // (256,256) [out][in]
    float* pooled;            // MODE 0
    const float* dpooled;     // MODE 1, 2
    float* dfeat;
    long long dfeat_bstride;
    float* dw_score;
    double* g1;               // MODE 1
    float* du2_part;          // MODE 2: [half][row][128]
    double* sum_du2;          // MODE 2: (2,128)
    const float* scal;        // [0] absmax |dpooled|
    int* status;
    int N, K;
    long long npts, ntiles;
};

template <int MODE, int NG>
struct WideSmem {
    static constexpr int GROUP_BYTES = 2 * kWXBytes + (MODE ? 2 * kWSBytes : 0) + kWRinfo * 4;
    static constexpr int OFF_GROUPS = 2 * kWWBytes;
    static constexpr int OFF_BARS = OFF_GROUPS + NG * GROUP_BYTES;
    static constexpr size_t BYTES = (size_t)OFF_BARS + 2 * NG * 8 + 16 + 34 * 4;
    // TMEM: per group S^T / dX^T slab 0 (R) + dX^T slab 1 (R); one shared sum accumulator (256)
    static constexpr int COLS = NG * 2 * kWR + (MODE ? kWD : 0);
    static constexpr uint32_t TMEM_COLS = COLS <= 64 ? 64 : (COLS <= 128 ? 128 : (COLS <= 256 ? 256 : 512));
};

__device__ __forceinline__ int wide_unit_off(int v, int ng) { return (v >> 3) * kWOpCs + ng * 128 + (v & 7) * 16; }
__device__ __forceinline__ float ex2_approx_w(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ void red_add_f32_w(float* p, float v) {
    asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}

template <int K, int MODE, int NG>
__global__ void __launch_bounds__((NG * 4 + 1) * 32, 1) lfa_cl_wide_kernel(LfaWideArgs a) {
    using S = WideSmem<MODE, NG>;
    constexpr int R = kWR, PTS = kWR / K;
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* Whi = smem;
    unsigned char* Wlo = smem + kWWBytes;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + S::OFF_BARS);
    uint64_t* done = full + NG;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + NG);
    float* red = reinterpret_cast<float*>(tmem_slot + 4);
    const int tid = threadIdx.x, warp = tid >> 5;
    const int half = blockIdx.y;

    if (tid == 0) {
        for (int g = 0; g < NG; ++g) {
            mbar_init(&full[g], 1);
            mbar_init(&done[g], 1);
        }
        mbar_fence_init();
    }
    if (warp == NG * 4) tmem_alloc_warp(tmem_slot, S::TMEM_COLS);
    // weight image of this half: element (vo = o - 128 half, vi) at (vi/8) * kWWIs + (vo/8) * 128 + (vo%8) * 16 + (vi%8) * 2
    const float sw = cl_pow2_scale(cl_block_absmax(a.w_score, kWD * kWD, red));
    for (int e = tid; e < kWH * kWD; e += blockDim.x) {
        const int vo = e / kWD, vi = e % kWD;
        const float v = a.w_score[(size_t)(half * kWH + vo) * kWD + vi] * sw;
        const __half hh = __float2half_rn(v);
        const __half ll = __float2half_rn(v - __half2float(hh));
        const int off = (vi >> 3) * kWWIs + (vo >> 3) * 128 + (vo & 7) * 16 + (vi & 7) * 2;
        *reinterpret_cast<__half*>(Whi + off) = hh;
        *reinterpret_cast<__half*>(Wlo + off) = ll;
    }
    float sg = 1.f;
    if (MODE) sg = cl_pow2_scale_to(a.scal[0], 2);
    fence_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem = *tmem_slot;
    const uint32_t col_sum = (uint32_t)(NG * 2 * R);

    if (warp == NG * 4) {
        if ((tid & 31) == 0) {
            constexpr int NR = MODE ? 2 : 1;
            int left[NG], step[NG];
            uint32_t ph[NG];
            int total = 0;
#pragma unroll
            for (int g = 0; g < NG; ++g) {
                const long long first = (long long)blockIdx.x + (long long)g * gridDim.x;
                const long long stride = (long long)NG * gridDim.x;
                const long long nt = first < a.ntiles ? (a.ntiles - first + stride - 1) / stride : 0;
                left[g] = (int)nt * NR;
                step[g] = 0;
                ph[g] = 0;
                total += left[g];
            }
            const uint32_t id_fwd = umma_idesc_f16(128, R, 0, 1);
            const uint32_t id_dx = umma_idesc_f16(128, R, 1, 1);
            const uint32_t id_sum = umma_idesc_f16(128, kWD, 0, 0);
            int nsum = 0;                                            // GEMM3 issues so far (single group in MODE 1, 2)
            while (total > 0) {
#pragma unroll
                for (int g = 0; g < NG; ++g) {
                    if (left[g] > 0 && mbar_try_wait(&full[g], ph[g])) {
                        tc_fence_after_sync();
                        const uint32_t xhi = smem_u32(smem + S::OFF_GROUPS + g * S::GROUP_BYTES);
                        const uint32_t xlo = xhi + kWXBytes;
                        const uint32_t shi = xlo + kWXBytes, slo = shi + kWSBytes;
                        const uint32_t acc0 = tmem + (uint32_t)(g * 2 * R), acc1 = acc0 + R;
                        if (!MODE || (step[g] & 1) == 0) {
                            // S^T[o-half] = Ws[o-half] X^T, K = 256
                            cl_mma_3x(acc0, smem_u32(Whi), smem_u32(Wlo), kWWIs, 128, xhi, xlo, kWOpCs, 128, id_fwd, 16, false);
                        } else {
                            // dWs[o-half] += dS^T X (K = rows);  dX^T slabs: r channels (vi 0..127), F channels (128..255)
                            cl_mma_3x(tmem + col_sum, shi, slo, 128, kWOpCs, xhi, xlo, 128, kWOpCs, id_sum, R / 16,
                                      (nsum % kWideFlush) != 0);
                            ++nsum;
                            cl_mma_3x(acc0, smem_u32(Whi), smem_u32(Wlo), 128, kWWIs, shi, slo, kWOpCs, 128, id_dx, 8,
                                      half == 0);
                            cl_mma_3x(acc1, smem_u32(Whi) + 16 * kWWIs, smem_u32(Wlo) + 16 * kWWIs, 128, kWWIs, shi, slo,
                                      kWOpCs, 128, id_dx, 8, half == 1);
                        }
                        umma_commit(&done[g]);
                        ph[g] ^= 1u;
                        ++step[g];
                        --left[g];
                        --total;
                    }
                }
            }
        }
    } else {
        const int g = tid >> 7, l = tid & 127;
        unsigned char* Xhi = smem + S::OFF_GROUPS + g * S::GROUP_BYTES;
        unsigned char* Xlo = Xhi + kWXBytes;
        unsigned char* Shi = Xlo + kWXBytes;
        unsigned char* Slo = Shi + kWSBytes;
        float* ri = reinterpret_cast<float*>(Xhi + 2 * kWXBytes + (MODE ? 2 * kWSBytes : 0));
        float w1[10];
#pragma unroll
        for (int q = 0; q < 10; ++q) w1[q] = a.w_rpe1[l * 10 + q];
        const float a1s = a.a_rpe1[l] * kClSx, b1s = a.b_rpe1[l] * kClSx;
        const uint32_t lane_field = (uint32_t)((l >> 5) * 32) << 16;
        const uint32_t tacc0 = tmem + lane_field + (uint32_t)(g * 2 * R), tacc1 = tacc0 + R;
        const uint32_t tsum = tmem + lane_field + col_sum;
        const float cs = 1.4426950408889634f / (kClSx * sw);
        const float inv_sx = 1.0f / kClSx;
        const int vo = half * kWH + l;                       // this lane's score channel = its virtual channel in X^T
        float amax = 0.f;
        double gacc[11];
#pragma unroll
        for (int q = 0; q < 11; ++q) gacc[q] = 0.0;
        uint32_t done_phase = 0;
        int ntile = 0;

        auto flush = [&]() {
            // dWs[o][:] += accumulator row of this lane (vector atomics), o = half * 128 + l
            const float unscale = 1.0f / (sg * kClSx);
            float* out = a.dw_score + (size_t)vo * kWD;
            tc_fence_after_sync();
#pragma unroll 1
            for (int c0 = 0; c0 < kWD; c0 += 16) {
                uint32_t u[16];
                tmem_ld16_nowait(tsum + (uint32_t)c0, u);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 16; j += 4)
                    red_add_v4(out + c0 + j, make_float4(__uint_as_float(u[j]) * unscale, __uint_as_float(u[j + 1]) * unscale,
                                                         __uint_as_float(u[j + 2]) * unscale, __uint_as_float(u[j + 3]) * unscale));
            }
            tc_fence_before_sync();
        };

        for (int it = 0;; ++it) {
            const long long tile = (long long)blockIdx.x + (long long)(it * NG + g) * gridDim.x;
            if (tile >= a.ntiles) break;
            // ---- A: row info (32 rows)
            if (l < R) {
                const int p = l / K, k = l % K;
                long long gp = tile * PTS + p;
                const bool valid = gp < a.npts;
                if (!valid) gp = a.npts - 1;
                const int b = (int)(gp / a.N);
                const int pi = (int)(gp - (long long)b * a.N);
                const int pj = a.idx[gp * K + k];
                float rpe[10];
                rpe_of_row(a.xyz + (size_t)b * a.xyz_bstride, pi, pj, rpe);
                float4* dst = reinterpret_cast<float4*>(ri + l * kWRow);
                dst[0] = make_float4(rpe[0], rpe[1], rpe[2], rpe[3]);
                dst[1] = make_float4(rpe[4], rpe[5], rpe[6], rpe[7]);
                const uint32_t off = (uint32_t)((long long)b * a.feat_bstride + (long long)pj * kWH);
                const uint32_t doff = valid ? (uint32_t)((long long)b * a.dfeat_bstride + (long long)pj * kWH) : 0xffffffffu;
                dst[2] = make_float4(rpe[8], rpe[9], __uint_as_float(off), __uint_as_float(doff));
            }
            named_bar_sync(1 + g, kClLanes);
            // ---- B: X^T: r channel l (mlp_rpe1 here, or r2 rows from rmat) and F channel l, all 32 rows
            {
                float fv[R], rv[R];
                const float* fb = a.feat + l;
#pragma unroll
                for (int j = 0; j < R; ++j) fv[j] = fb[__float_as_uint(ri[j * kWRow + 10])];
                if (a.rmat != nullptr) {
                    const float* rb = a.rmat + (size_t)tile * R * kWH + l;
                    const long long rows_left = a.npts * K - tile * R;          // rows of this tile that exist
#pragma unroll
                    for (int j = 0; j < R; ++j) rv[j] = (j < rows_left) ? rb[(size_t)j * kWH] * kClSx : 0.f;
                } else {
#pragma unroll
                    for (int j = 0; j < R; ++j) {
                        const float4* q = reinterpret_cast<const float4*>(ri + j * kWRow);
                        rv[j] = cl_mlp1(w1, a1s, b1s, q[0], q[1], q[2]);
                    }
                }
#pragma unroll
                for (int u = 0; u < R / 8; ++u) {
                    float v[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        v[j] = rv[u * 8 + j];
                        amax = fmaxf(amax, v[j]);
                    }
                    cl_store_unit(Xhi, Xlo, wide_unit_off(l, u), v);
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        v[j] = fv[u * 8 + j] * kClSx;
                        amax = fmaxf(amax, fabsf(v[j]));
                    }
                    cl_store_unit(Xhi, Xlo, wide_unit_off(kWH + l, u), v);
                }
            }
            fence_async_smem();
            tc_fence_before_sync();
            named_bar_sync(1 + g, kClLanes);
            if (l == 0) mbar_arrive(&full[g]);

            // ---- C: softmax over K, pooled; backward: dS and g A
            mbar_wait(&done[g], done_phase);
            done_phase ^= 1u;
            tc_fence_after_sync();
#pragma unroll 1
            for (int p = 0; p < PTS; ++p) {
                float s[K], x[K];
#pragma unroll
                for (int k0 = 0; k0 < K; k0 += 16) {
                    uint32_t u[16];
                    tmem_ld16_nowait(tacc0 + (uint32_t)(p * K + k0), u);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 16; ++j) s[k0 + j] = __uint_as_float(u[j]);
                }
#pragma unroll
                for (int k0 = 0; k0 < K; k0 += 8) {
                    float t[8];
                    cl_load_unit(Xhi, Xlo, wide_unit_off(vo, (p * K + k0) / 8), t);
#pragma unroll
                    for (int j = 0; j < 8; ++j) x[k0 + j] = t[j];
                }
                float m = s[0];
#pragma unroll
                for (int k = 1; k < K; ++k) m = fmaxf(m, s[k]);
                const float mc = m * cs;
                float den = 0.f, num = 0.f;
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    s[k] = ex2_approx_w(fmaf(s[k], cs, -mc));
                    den += s[k];
                    num = fmaf(s[k], x[k], num);
                }
                const long long gp = tile * PTS + p;
                if (!MODE) {
                    if (gp < a.npts) a.pooled[gp * kWD + vo] = (num * inv_sx) / den;
                } else {
                    const float inv = 1.0f / den;
                    const float pooled = num * inv;
                    const float gv = (gp < a.npts) ? a.dpooled[gp * kWD + vo] * sg : 0.f;
                    const float gi = gv * inv;
                    const uint32_t tself = (half == 0 ? tacc0 : tacc1) + (uint32_t)(p * K);
#pragma unroll
                    for (int k0 = 0; k0 < K; k0 += 16) {
                        uint32_t u[16];
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            const float ga = gi * s[k0 + j];
                            x[k0 + j] = ga * (x[k0 + j] - pooled) * inv_sx;
                            u[j] = __float_as_uint(ga * sw);
                        }
                        tmem_st16(tself + (uint32_t)k0, u);        // initial value of the slab that holds this lane's channel
                    }
#pragma unroll
                    for (int k0 = 0; k0 < K; k0 += 8) {
                        float t[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) t[j] = x[k0 + j];
                        cl_store_unit(Shi, Slo, wide_unit_off(l, (p * K + k0) / 8), t);
                    }
                }
            }
            if (MODE) {
                tmem_st_wait();
                fence_async_smem();
                tc_fence_before_sync();
                named_bar_sync(1 + g, kClLanes);
                if (l == 0) mbar_arrive(&full[g]);
                // ---- D: partial dX^T over this half's score channels: r channel l (slab 0), F channel l (slab 1)
                mbar_wait(&done[g], done_phase);
                done_phase ^= 1u;
                tc_fence_after_sync();
                const float inv2 = 1.0f / (sg * sw);
                float part[11];
#pragma unroll
                for (int q = 0; q < 11; ++q) part[q] = 0.f;
                float* df = a.dfeat + l;
                float* du2_row = (MODE == 2) ? a.du2_part + ((size_t)half * a.npts * K + (size_t)tile * R) * kWH + l : nullptr;
                const long long rows_left = a.npts * K - tile * R;
#pragma unroll 1
                for (int c0 = 0; c0 < R; c0 += 16) {
                    uint32_t u0[16], u1[16];
                    tmem_ld16_nowait(tacc0 + (uint32_t)c0, u0);
                    tmem_ld16_nowait(tacc1 + (uint32_t)c0, u1);
                    tmem_ld_wait();
#pragma unroll
                    for (int hlf = 0; hlf < 2; ++hlf) {
                        float xr[8];
                        cl_load_unit(Xhi, Xlo, wide_unit_off(l, c0 / 8 + hlf), xr);
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const int n = c0 + hlf * 8 + j;
                            const float4* q = reinterpret_cast<const float4*>(ri + n * kWRow);
                            const float4 q2 = q[2];
                            const uint32_t doff = __float_as_uint(q2.w);
                            if (doff != 0xffffffffu) red_add_f32_w(df + doff, __uint_as_float(u1[hlf * 8 + j]) * inv2);
                            const float du = (xr[j] > 0.f) ? __uint_as_float(u0[hlf * 8 + j]) * inv2 : 0.f;
                            if (MODE == 1) {
                                const float4 q0 = q[0], q1 = q[1];
                                part[0] = fmaf(du, q0.x, part[0]); part[1] = fmaf(du, q0.y, part[1]);
                                part[2] = fmaf(du, q0.z, part[2]); part[3] = fmaf(du, q0.w, part[3]);
                                part[4] = fmaf(du, q1.x, part[4]); part[5] = fmaf(du, q1.y, part[5]);
                                part[6] = fmaf(du, q1.z, part[6]); part[7] = fmaf(du, q1.w, part[7]);
                                part[8] = fmaf(du, q2.x, part[8]); part[9] = fmaf(du, q2.y, part[9]);
                                part[10] += du;
                            } else {
                                part[0] += du;
                                part[1] = fmaf(du, xr[j] * inv_sx, part[1]);
                                if (n < rows_left) du2_row[(size_t)n * kWH] = du;
                            }
                        }
                    }
                }
#pragma unroll
                for (int q = 0; q < 11; ++q) gacc[q] += (double)part[q];
                tc_fence_before_sync();
                ++ntile;
                if (ntile % kWideFlush == 0) flush();
                named_bar_sync(1 + g, kClLanes);          // row info is rewritten by the next tile
            } else {
                tc_fence_before_sync();
            }
        }
        if (MODE) {
            if (ntile % kWideFlush != 0) flush();
            if (MODE == 1) {
#pragma unroll
                for (int q = 0; q < 11; ++q) atomicAdd(a.g1 + l * 16 + q, gacc[q]);
            } else {
                atomicAdd(a.sum_du2 + l, gacc[0]);
                atomicAdd(a.sum_du2 + kWH + l, gacc[1]);
            }
        }
        if (a.status != nullptr && !(amax < 65504.f)) atomicOr(a.status, 1);
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == NG * 4) tmem_dealloc_warp(tmem, S::TMEM_COLS);
}

template <int K, int MODE, int NG>
static int launch_wide(const LfaWideArgs& a, cudaStream_t st) {
    auto kern = lfa_cl_wide_kernel<K, MODE, NG>;
    constexpr size_t smem = WideSmem<MODE, NG>::BYTES;
    static_assert(smem <= 232448, "tile does not fit shared memory");
    R3D_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int gx = (int)(a.ntiles < kNumSMs / 2 ? a.ntiles : kNumSMs / 2);
    kern<<<dim3(gx, 2), (NG * 4 + 1) * 32, smem, st>>>(a);
    R3D_LAUNCH_CHECK("lfa_cl_wide_kernel");
    return R3D_OK;
}

// r1 rows: out (B*N*K, h) = relu(a1 (W1 rpe) + b1) for every (point, neighbour) row — the materialised input of mlp_rpe2
// for the widest level (see file header).  One CTA = 32 rows, thread = channel (h <= 128: one pass).
__global__ void __launch_bounds__(128) lfa_r1_rows_kernel(const float* __restrict__ xyz, long long xyz_bstride,
                                                          const int32_t* __restrict__ idx, const float* __restrict__ w1,
                                                          const float* __restrict__ a1, const float* __restrict__ b1,
                                                          float* __restrict__ out, int N, int K, int h, long long rows) {
    __shared__ __align__(16) float rp[32 * 12];
    const long long row0 = (long long)blockIdx.x * 32;
    const int t = threadIdx.x;
    if (t < 32 && row0 + t < rows) {
        const long long row = row0 + t;
        const long long gp = row / K;
        const int b = (int)(gp / N), pi = (int)(gp - (long long)b * N);
        float rpe[10];
        rpe_of_row(xyz + (size_t)b * xyz_bstride, pi, idx[row], rpe);
#pragma unroll
        for (int q = 0; q < 10; ++q) rp[t * 12 + q] = rpe[q];
    }
    __syncthreads();
    for (int c = t; c < h; c += 128) {
        float w[10];
#pragma unroll
        for (int q = 0; q < 10; ++q) w[q] = w1[c * 10 + q];
        const float sa = a1[c], sb = b1[c];
        for (int j = 0; j < 32 && row0 + j < rows; ++j) {
            const float4* q = reinterpret_cast<const float4*>(rp + j * 12);
            out[(size_t)(row0 + j) * h + c] = cl_mlp1(w, sa, sb, q[0], q[1], q[2]);
        }
    }
}

// du2 (layout of r3d_lfa_bn2_bwd: [b][tile][h][P*K], P points per tile) = part[0] + part[1], part [half][row][h].
// One CTA per (cloud, tile): its P*K rows are read row by row (coalesced along the channels), summed, transposed
// through shared memory and written channel by channel (coalesced along the rows).
__global__ void __launch_bounds__(256) lfa_du2_combine_kernel(const float* __restrict__ part, float* __restrict__ out, int B,
                                                              int N, int K, int h, int P, long long rows) {
    extern __shared__ float tile[];                    // [P*K][h + 1]
    const int T = (N + P - 1) / P;
    const int b = blockIdx.x / T, t = blockIdx.x % T;
    const int PK = P * K, ld = h + 1;
    const long long row0 = ((long long)b * N + (long long)t * P) * K;
    const int nrows = min(PK, (N - t * P) * K);        // rows of this tile that exist
    for (int i = threadIdx.x; i < PK * h; i += 256) {
        const int r = i / h, c = i % h;
        float v = 0.f;
        if (r < nrows) v = part[(row0 + r) * h + c] + part[(rows + row0 + r) * h + c];
        tile[r * ld + c] = v;
    }
    __syncthreads();
    float* o = out + (size_t)blockIdx.x * h * PK;
    for (int i = threadIdx.x; i < PK * h; i += 256) {
        const int c = i / PK, r = i % PK;
        o[i] = tile[r * ld + c];
    }
}

}  // namespace r3d

using namespace r3d;

extern "C" int r3d_lfa_r1_rows(const float* xyz, long long xyz_bstride, const int32_t* idx, const float* w_rpe1,
                               const float* a_rpe1, const float* b_rpe1, float* out, int B, int N, int K, int h,
                               r3d_stream_t stream) {
    if (B < 0 || N < 0 || K <= 0 || h <= 0) return R3D_EINVAL;
    if (B == 0 || N == 0) return R3D_OK;
    if (!xyz || !idx || !w_rpe1 || !a_rpe1 || !b_rpe1 || !out) return R3D_EINVAL;
    if (xyz_bstride == 0) xyz_bstride = (long long)N * 3;
    const long long rows = (long long)B * N * K;
    lfa_r1_rows_kernel<<<(unsigned)((rows + 31) / 32), 128, 0, static_cast<cudaStream_t>(stream)>>>(
        xyz, xyz_bstride, idx, w_rpe1, a_rpe1, b_rpe1, out, N, K, h, rows);
    R3D_LAUNCH_CHECK("lfa_r1_rows_kernel");
    return R3D_OK;
}

extern "C" int r3d_lfa_du2_combine(const float* part, float* out, int B, int N, int K, int h, int tile_points,
                                   r3d_stream_t stream) {
    if (B < 0 || N < 0 || K <= 0 || h <= 0 || tile_points <= 0) return R3D_EINVAL;
    if (B == 0 || N == 0) return R3D_OK;
    if (!part || !out) return R3D_EINVAL;
    const int T = (N + tile_points - 1) / tile_points;
    const size_t smem = (size_t)tile_points * K * (h + 1) * sizeof(float);
    if (smem > 200 * 1024) return R3D_EUNSUPPORTED;
    R3D_CUDA_TRY(cudaFuncSetAttribute(lfa_du2_combine_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    lfa_du2_combine_kernel<<<(unsigned)(B * T), 256, smem, static_cast<cudaStream_t>(stream)>>>(
        part, out, B, N, K, h, tile_points, (long long)B * N * K);
    R3D_LAUNCH_CHECK("lfa_du2_combine_kernel");
    return R3D_OK;
}

extern "C" int r3d_lfa_tc_wide(int mode, const float* xyz, long long xyz_bstride, const int32_t* idx, const float* feat,
                               long long feat_bstride, const float* w_rpe1, const float* a_rpe1, const float* b_rpe1,
                               const float* rmat, const float* w_score, float* pooled, const float* dpooled, float* dfeat,
                               long long dfeat_bstride, float* dw_score, double* g1, float* du2_part, double* sum_du2,
                               const float* scal, int* status, int B, int N, int K, int d, r3d_stream_t stream) {
    if (mode < 0 || mode > 2) return R3D_EINVAL;
    if (B < 0 || N < 0 || K <= 0) return R3D_EINVAL;
    if (d != kWD || (K != 16 && K != 32)) return R3D_EUNSUPPORTED;
    if (B == 0 || N == 0) return R3D_OK;
    if (!xyz || !idx || !feat || !w_rpe1 || !a_rpe1 || !b_rpe1 || !w_score) return R3D_EINVAL;
    if (mode == 0 && !pooled) return R3D_EINVAL;
    if (mode >= 1 && (!dpooled || !dfeat || !dw_score || !scal)) return R3D_EINVAL;
    if (mode == 1 && !g1) return R3D_EINVAL;
    if (mode == 2 && (!rmat || !du2_part || !sum_du2)) return R3D_EINVAL;
    if (xyz_bstride == 0) xyz_bstride = (long long)N * 3;
    if (feat_bstride == 0) feat_bstride = (long long)N * kWH;
    if (dfeat_bstride == 0) dfeat_bstride = (long long)N * kWH;
    if ((long long)(B - 1) * feat_bstride + (long long)N * kWH >= (1ll << 32) - 1 ||
        (long long)(B - 1) * dfeat_bstride + (long long)N * kWH >= (1ll << 32) - 1)
        return R3D_EUNSUPPORTED;
    if (dw_score && !is_aligned(dw_score, 16)) return R3D_EALIGN;
    LfaWideArgs a{xyz, xyz_bstride, idx, feat, feat_bstride, w_rpe1, a_rpe1, b_rpe1, rmat, w_score, pooled, dpooled, dfeat,
                  dfeat_bstride, dw_score, g1, du2_part, sum_du2, scal, status, N, K, (long long)B * N, 0};
    const int pts = kWR / K;
    a.ntiles = (a.npts + pts - 1) / pts;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (K == 16) {
        if (mode == 0) return launch_wide<16, 0, 2>(a, st);
        if (mode == 1) return launch_wide<16, 1, 1>(a, st);
        return launch_wide<16, 2, 1>(a, st);
    }
    if (mode == 0) return launch_wide<32, 0, 2>(a, st);
    if (mode == 1) return launch_wide<32, 1, 1>(a, st);
    return launch_wide<32, 2, 1>(a, st);
}
