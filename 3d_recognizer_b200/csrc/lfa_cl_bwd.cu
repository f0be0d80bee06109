// Training-mode companions of the tensor-core fused LocSE + attentive-pooling kernel (lfa_cl.cu), same "channel-lane"
// design (lfa_cl_common.cuh): TMEM lane = virtual channel, columns = (point, neighbour) rows, split-fp16 operands, one
// MMA warp + NG worker groups per persistent CTA.  One kernel template, four modes:
//
//   MODE 1  backward of a stage-1 launch (autograd of modules.py:316-319 as driven by trainer.py:115-119)
//             recompute X^T, S^T = Ws X^T, A = softmax_K(S), pooled;  dS = A g (X - pooled);
//             dX^T = g A + Ws^T dS^T   (g A is stored to TMEM as the accumulator's initial value);
//             dWs += dS^T X            (accumulated ACROSS TILES in TMEM, see below);
//             lanes of the feature half scatter-add dX to dfeat[idx] (coalesced red.global.add.f32),
//             lanes of the encoding half form du1 = dX [r1 > 0] and G1 += du1 (x) [rpe, 1] in registers.
//   MODE 2  pass 1 of the train-mode backward of a stage-2 launch (batch-statistics BatchNorm behind mlp_rpe2): as
//             MODE 1 on X = [r2 ; p1[idx]] with r2 = relu(a2 (W2 r1) + c2) recomputed through a second MMA, but the
//             encoding lanes stop at du2 = dX [r2 > 0]: they write it per tile and accumulate sum du2, sum du2 r2.
//   MODE 3  pass 2: dz2 = a2 (du2 - m1 - zhat2 m2) per row, dW2 += dz2^T r1 (TMEM), dr1^T = W2^T dz2^T,
//             du1 = dr1 [r1 > 0], G1 += du1 (x) [rpe, 1].
//   MODE 4  second moments of r1 for mlp_rpe2's batch statistics: M += r1^T r1 (TMEM), s += sum r1.
//
// Weight-gradient-like sums (dWs, dW2, M) are contractions over ROWS: both operands are the row-operand planes read
// K-major with K = rows, and the accumulator stays in TMEM across the tiles of a group.  The tensor core truncates when
// it adds to the accumulator (measured ~1 ulp per MMA, tools/tc16_probe_test.py), which over thousands of MMAs would
// bias the sum by ~1e-4; so each group owns a first-level accumulator that it folds, every kClFlush tiles, into a
// second-level accumulator (fp32 round-to-nearest adds in registers, TMEM -> registers -> TMEM) shared by the CTA's
// groups under a shared-memory lock.  At the end the CTA adds its totals to global memory with one atomic per element.
//
// Scales of the split-fp16 operands (powers of two): activations kClSx (fixed), weights from their absmax (per CTA),
// dS from absmax |dpooled| (scal[0], written by r3d_absmax before the launch), dz2 from absmax |du2| (scal[1], pass 1).
#include "lfa_cl_common.cuh"

namespace r3d {

// unroll factors of the softmax / dS loop over a sub-tile's points (phase C) and of the scatter loop (phase D): the fully
// unrolled tile loop is ~52 KB of SASS, more than the instruction caches hold with three warp roles in flight
#ifndef R3D_CL_UNROLL_C
#define R3D_CL_UNROLL_C 1
#endif
#ifndef R3D_CL_UNROLL_D
#define R3D_CL_UNROLL_D 1
#endif
constexpr int kClFlush = 8;       // tiles of a group between two folds of its first-level accumulator

struct LfaClBwdArgs {
    const float* xyz;
    long long xyz_bstride;
    const int32_t* idx;
    const float* feat;
    long long feat_bstride;
    const float* w_rpe1;    // (h,10)
    const float* a_rpe1;
    const float* b_rpe1;
    const float* w_rpe2;    // (h,h) [out][in]
    const float* a_rpe2;
    const float* b_rpe2;
    const float* w_score;   // (d,d) [out][in]
    const float* dpooled;   // (B,N,d)
    float* dfeat;           // (B,N,h) +=
    long long dfeat_bstride;
    float* dw_score;        // (d,d) +=
    double* g1;             // (h,16) +=
    float* du2_tiles;       // MODE 2 out / MODE 3 in: [tile][64 r lanes][64 rows]
    double* sum_du2;        // (2,h) +=                                   MODE 2
    const float* bn2;       // (5,h) a2, mean2, rstd2, m1, m2             MODE 3
    double* dw2;            // (h,h) +=                                   MODE 3
    double* m_r1;           // (h,h) +=                                   MODE 4
    double* s_r1;           // (h,16): [:,10] +=                          MODE 4
    float* scal;            // [0] absmax |dpooled| (in), [1] absmax |du2| (MODE 2: atomic max, MODE 3: in)
    int* status;
    int N;
    long long npts;
    long long ntiles;
};

template <int D, int K, int MODE, int NG>
struct LfaClBwdSmem {
    using C = ClCfg<D, K>;
    static constexpr bool HAS_W = MODE <= 2;
    // mlp_rpe2 inside pass 1 (MODE 2): an MMA against the W2 image for d >= 64; for the narrow levels (h <= 16) the
    // h x h product runs on the CUDA cores from an fp32 staging copy of r1 (no image, no extra MMA round, and the
    // shared memory it saves is what lets two worker groups fit at d = 16)
    static constexpr bool U2_MMA = (MODE == 2 && D > 32) || MODE == 3;
    static constexpr bool HAS_W2 = U2_MMA;
    static constexpr bool HAS_DS = MODE <= 3;
    static constexpr int OFF_W2 = HAS_W ? 2 * C::W_BYTES : 0;
    static constexpr int OFF_GROUPS = OFF_W2 + (HAS_W2 ? 2 * kClW2Bytes : 0);
    static constexpr int GROUP_BYTES = (HAS_DS ? 4 : 2) * C::OP_BYTES + C::RINFO_FLOATS * 4;
    static constexpr int OFF_BARS = OFF_GROUPS + NG * GROUP_BYTES;
    static constexpr int OFF_EXTRA = OFF_BARS + 2 * NG * 8 + 16 + 34 * 4 + 16;      // MODE 4: centre[64] f32, S'[64] f64, n[8] f64
    static constexpr size_t BYTES = (size_t)OFF_EXTRA + (MODE == 4 ? 64 * 4 + 64 * 8 + 8 * 8 : 0);
    // TMEM columns: per group a working accumulator (R) and a first-level sum accumulator (ACC1), one shared second level
    static constexpr int ACC1 = MODE <= 2 ? 128 : 64;
    static constexpr int COL_ACC2 = NG * (C::R + ACC1);
    static constexpr int COLS = COL_ACC2 + ACC1;
    static constexpr uint32_t TMEM_COLS = COLS <= 128 ? 128 : (COLS <= 256 ? 256 : 512);
    static_assert(COLS <= 512, "TMEM");
};

__device__ __forceinline__ float ex2_approx_b(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ void red_add_f32(float* p, float v) {
    asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}

template <int D, int K, int MODE, int NG>
__global__ void __launch_bounds__((NG * 4 + 1) * 32, 1) lfa_cl_bwd_kernel(LfaClBwdArgs a) {
    using C = ClCfg<D, K>;
    using S = LfaClBwdSmem<D, K, MODE, NG>;
    constexpr int H = C::H, R = C::R;
    constexpr int ACC1 = S::ACC1;
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* Whi = smem;
    unsigned char* Wlo = smem + C::W_BYTES;
    unsigned char* W2hi = smem + S::OFF_W2;
    unsigned char* W2lo = W2hi + kClW2Bytes;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + S::OFF_BARS);
    uint64_t* done = full + NG;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + NG);
    int* lock = reinterpret_cast<int*>(tmem_slot + 1);
    float* red = reinterpret_cast<float*>(tmem_slot + 4);
    float* centre = reinterpret_cast<float*>(smem + S::OFF_EXTRA);            // MODE 4: per virtual r lane
    double* csum = reinterpret_cast<double*>(smem + S::OFF_EXTRA + 64 * 4);   //         sum of centred r1 per lane
    double* cnum = csum + 64;                                                 //         valid rows per sub-tile

    const int tid = threadIdx.x, warp = tid >> 5;

    if (tid == 0) {
        for (int g = 0; g < NG; ++g) {
            mbar_init(&full[g], 1);
            mbar_init(&done[g], 1);
        }
        *lock = 0;
        mbar_fence_init();
    }
    if (MODE == 4 && tid < 64) {
        if (tid < 8) red[tid] = 0.f;
        centre[tid] = 0.f;
        csum[tid] = 0.0;
        if (tid < 8) cnum[tid] = 0.0;
    }
    if (warp == NG * 4) tmem_alloc_warp(tmem_slot, S::TMEM_COLS);
    float sw = 1.f, sw2 = 1.f;
    if (S::HAS_W) {
        sw = cl_pow2_scale(cl_block_absmax(a.w_score, D * D, red));
        cl_build_weight_image<D>(a.w_score, sw, Whi, Wlo);
    }
    if (S::HAS_W2) {
        sw2 = cl_pow2_scale(cl_block_absmax(a.w_rpe2, H * H, red));
        cl_build_w2_image<D>(a.w_rpe2, sw2, W2hi, W2lo);
    }
    // gradient-side operand scale: dS (MODE 1, 2) from absmax |dpooled| -> [2, 4); dz2 (MODE 3) from absmax |du2| max |a2| -> [32, 64)
    float sg = 1.f;
    if (MODE <= 2) sg = cl_pow2_scale_to(a.scal[0], 2);
    if (MODE == 3) sg = cl_pow2_scale_to(a.scal[1] * cl_block_absmax(a.bn2, H, red), 6);
    fence_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem = *tmem_slot;

    if (warp == NG * 4) {
        // ============================================================ MMA issuer
        if ((tid & 31) == 0) {
            constexpr int NR = (MODE == 1) ? 2 : (MODE == 2 ? (S::U2_MMA ? 3 : 2) : (MODE == 3 ? 2 : 1));     // MMA rounds per tile
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
            const uint32_t id_fwd = umma_idesc_f16(kClLanes, R, 0, 1);        // weights K-major x rows MN-major
            const uint32_t id_dx = umma_idesc_f16(kClLanes, R, 1, 1);         // weights MN-major x rows MN-major
            const uint32_t id_sum = umma_idesc_f16(kClLanes, ACC1, 0, 0);     // rows K-major x rows K-major (K = rows)
            while (total > 0) {
#pragma unroll
                for (int g = 0; g < NG; ++g) {
                    if (left[g] > 0 && mbar_try_wait(&full[g], ph[g])) {
                        tc_fence_after_sync();
                        const uint32_t xhi = smem_u32(smem + S::OFF_GROUPS + g * S::GROUP_BYTES);
                        const uint32_t xlo = xhi + C::OP_BYTES;
                        const uint32_t shi = xlo + C::OP_BYTES, slo = shi + C::OP_BYTES;      // dS / dz2 planes
                        const uint32_t acc = tmem + (uint32_t)(g * (R + ACC1));
                        const uint32_t acc1 = acc + R;
                        const int tile_no = step[g] / NR, round = step[g] % NR;
                        const bool fresh = (tile_no % kClFlush) == 0;                          // first tile after a fold
                        const uint32_t w2h = smem_u32(W2hi), w2l = smem_u32(W2lo);
                        if (MODE == 1 || MODE == 2) {
                            if (MODE == 2 && S::U2_MMA && round == 0) {
                                cl_mma_3x(acc, w2h, w2l, kClW2Is, 128, xhi, xlo, C::OP_CS, 128, id_fwd, 4, false);      // U2
                            } else if (round == NR - 2) {
                                cl_mma_3x(acc, smem_u32(Whi), smem_u32(Wlo), C::W_IS, 128, xhi, xlo, C::OP_CS, 128, id_fwd, 8,
                                          false);                                                                       // S
                            } else {
                                // dWs += dS^T X  (K = rows);  dX^T = gA + Ws^T dS^T
                                cl_mma_3x(acc1, shi, slo, 128, C::OP_CS, xhi, xlo, 128, C::OP_CS, id_sum, R / 16, !fresh);
                                cl_mma_3x(acc, smem_u32(Whi), smem_u32(Wlo), 128, C::W_IS, shi, slo, C::OP_CS, 128, id_dx, 8,
                                          true);
                            }
                        } else if (MODE == 3) {
                            if (round == 0) {
                                cl_mma_3x(acc, w2h, w2l, kClW2Is, 128, xhi, xlo, C::OP_CS, 128, id_fwd, 4, false);      // U2
                            } else {
                                cl_mma_3x(acc1, shi, slo, 128, C::OP_CS, xhi, xlo, 128, C::OP_CS, id_sum, R / 16, !fresh);  // dW2
                                cl_mma_3x(acc, w2h, w2l, 128, kClW2Is, shi, slo, C::OP_CS, 128, id_dx, 4, false);       // dr1
                            }
                        } else {
                            cl_mma_3x(acc1, xhi, xlo, 128, C::OP_CS, xhi, xlo, 128, C::OP_CS, id_sum, R / 16, !fresh);  // M
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
        // ============================================================ worker groups
        const int g = tid >> 7, l = tid & 127;
        const ClLane<D> ln(l);
        const int lr = l & 63;                      // the r lane / F lane pair this thread produces rows for
        const int hh = l >> 6;                      // ... on row groups [4 hh, 4 hh + 4)
        const int rc = lr % H;                      // channel inside the half
        unsigned char* Xhi = smem + S::OFF_GROUPS + g * S::GROUP_BYTES;
        unsigned char* Xlo = Xhi + C::OP_BYTES;
        unsigned char* Shi = Xlo + C::OP_BYTES;
        unsigned char* Slo = Shi + C::OP_BYTES;
        float* rinfo = reinterpret_cast<float*>(Xhi + (S::HAS_DS ? 4 : 2) * C::OP_BYTES);
        const float* ri = rinfo + ln.sub * C::SUB_RI;            // (l & 63) / H == ln.sub for both parts
        float w1[10];
#pragma unroll
        for (int q = 0; q < 10; ++q) w1[q] = a.w_rpe1[rc * 10 + q];
        const float a1s = a.a_rpe1[rc] * kClSx, b1s = a.b_rpe1[rc] * kClSx;
        float a2s = 0.f, b2s = 0.f;
        float w2r[S::U2_MMA ? 1 : H];                 // CUDA-core mlp_rpe2 (narrow levels): this channel's weight row
        if (MODE == 2) {
            a2s = S::U2_MMA ? a.a_rpe2[rc] / sw2 : a.a_rpe2[rc];     // staging holds r1 * sx, so does the result
            b2s = a.b_rpe2[rc] * kClSx;
            if (!S::U2_MMA) {
#pragma unroll
                for (int i = 0; i < H; ++i) w2r[i] = a.w_rpe2[rc * H + i];
            }
        }
        float bn_a2 = 0.f, bn_mean = 0.f, bn_rstd = 0.f, bn_m1 = 0.f, bn_m2 = 0.f;
        if (MODE == 3) {
            bn_a2 = a.bn2[rc];
            bn_mean = a.bn2[H + rc];
            bn_rstd = a.bn2[2 * H + rc];
            bn_m1 = a.bn2[3 * H + rc];
            bn_m2 = a.bn2[4 * H + rc];
        }
        const uint32_t lane_field = (uint32_t)((l >> 5) * 32) << 16;
        const uint32_t tacc = tmem + lane_field + (uint32_t)(g * (R + ACC1));
        const uint32_t tacc1 = tacc + R;
        const uint32_t tacc2 = tmem + lane_field + (uint32_t)S::COL_ACC2;
        const float cs = 1.4426950408889634f / (kClSx * sw);
        const float inv_sx = 1.0f / kClSx;
        float amax = 0.f, du2max = 0.f;
        double gacc[11];                       // G1 row of this lane's channel (MODE 1, 3) / sums (MODE 2, 4)
#pragma unroll
        for (int q = 0; q < 11; ++q) gacc[q] = 0.0;
        uint32_t done_phase = 0;
        int ntile = 0;

        // second-level accumulator starts at zero
        if (g == 0) {
            uint32_t z[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) z[j] = 0u;
            for (int c0 = 0; c0 < ACC1; c0 += 16) tmem_st16(tacc2 + (uint32_t)c0, z);
            tmem_st_wait();
            tc_fence_before_sync();
        }
        named_bar_sync(15, NG * kClLanes);
        tc_fence_after_sync();

        // MODE 4: the moments are taken of CENTRED values r1 - c (c = this CTA's estimate of the channel mean from its
        // first tile) and converted back in fp64 at the end: M = M' + c S'^T + S' c^T + n c c^T.  The tensor core's
        // accumulate truncation is a systematic relative bias of the accumulated matrix; on raw second moments it is
        // amplified by E[z^2] / var(z) in the BatchNorm variance that is formed from them, on centred ones it is not.
        float cme = 0.f;
        double nvalid = 0.0;
        if (MODE == 4) {
            if (g == 0 && (long long)blockIdx.x < a.ntiles) {
                cl_row_info<D, K>(rinfo, a.xyz, a.xyz_bstride, a.idx, a.feat_bstride, a.dfeat_bstride, a.N, a.npts,
                                  (long long)blockIdx.x, l);
                named_bar_sync(1, kClLanes);
                float sum = 0.f, cnt = 0.f;
                for (int j = 0; j < 32; ++j) {
                    if (cl_grad_off(ri, hh * 32 + j) != 0xffffffffu) {
                        float4 q0, q1, q2;
                        cl_rpe_row<D, K>(ri, hh * 32 + j, q0, q1, q2);
                        sum += cl_mlp1(w1, a1s, b1s, q0, q1, q2);
                        cnt += 1.f;
                    }
                }
                atomicAdd(&centre[lr], sum);                    // both row halves of the lane
                if (rc == 0) atomicAdd(&red[ln.sub], cnt);      // rows of the sub-tile (red[0..7] is free by now)
            }
            if (g == 0) {
                named_bar_sync(1, kClLanes);
                if (l < 64) {
                    const float n = red[l / H];
                    centre[l] = n > 0.f ? centre[l] / n : 0.f;  // still scaled by sx, like the operand values
                }
            }
            named_bar_sync(15, NG * kClLanes);
            cme = centre[lr];
        }

        auto fold = [&]() {
            // acc2 += acc1 under the CTA lock (all 128 lanes of the group; its MMAs have completed)
            if (l == 0) {
                while (atomicCAS(lock, 0, 1) != 0) {
                }
                __threadfence_block();
            }
            named_bar_sync(1 + g, kClLanes);
            tc_fence_after_sync();
#pragma unroll 1
            for (int c0 = 0; c0 < ACC1; c0 += 16) {
                uint32_t x1[16], x2[16];
                tmem_ld16_nowait(tacc1 + (uint32_t)c0, x1);
                tmem_ld16_nowait(tacc2 + (uint32_t)c0, x2);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 16; ++j) x2[j] = __float_as_uint(__uint_as_float(x1[j]) + __uint_as_float(x2[j]));
                tmem_st16(tacc2 + (uint32_t)c0, x2);
            }
            tmem_st_wait();
            tc_fence_before_sync();
            named_bar_sync(1 + g, kClLanes);
            if (l == 0) {
                __threadfence_block();
                atomicExch(lock, 0);
            }
        };

        for (int it = 0;; ++it) {
            const long long tile = (long long)blockIdx.x + (long long)(it * NG + g) * gridDim.x;
            if (tile >= a.ntiles) break;
            // ---- A: row info
            cl_row_info<D, K>(rinfo, a.xyz, a.xyz_bstride, a.idx, a.feat_bstride, a.dfeat_bstride, a.N, a.npts, tile, l);
            named_bar_sync(1 + g, kClLanes);
            // ---- B: row operand.  Thread pair (l, l ^ 64) shares r lane lr and F lane 64 + lr, half the rows each.
            {
                float fv[32];
                if (MODE <= 2) {
                    const float* fb = a.feat + rc;
#pragma unroll
                    for (int j = 0; j < 32; ++j) fv[j] = fb[cl_feat_off(ri, hh * 32 + j)];
                }
                float rs = 0.f;
#pragma unroll 1
                for (int u = 0; u < 4; ++u) {
                    const int ng = hh * 4 + u;
                    float v[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        float4 q0, q1, q2;
                        cl_rpe_row<D, K>(ri, ng * 8 + j, q0, q1, q2);
                        v[j] = cl_mlp1(w1, a1s, b1s, q0, q1, q2);
                        if (MODE == 4) {
                            const bool valid = cl_grad_off(ri, ng * 8 + j) != 0xffffffffu;
                            v[j] = valid ? v[j] - cme : 0.f;                  // centred; padding rows count for nothing
                            nvalid += valid ? 1.0 : 0.0;
                        }
                        rs += v[j];
                    }
                    amax = fmaxf(amax, fmaxf(fmaxf(fmaxf(fabsf(v[0]), fabsf(v[1])), fmaxf(fabsf(v[2]), fabsf(v[3]))),
                                             fmaxf(fmaxf(fabsf(v[4]), fabsf(v[5])), fmaxf(fabsf(v[6]), fabsf(v[7])))));
                    if (MODE == 2 && !S::U2_MMA) {
                        // r1 (x sx) of this channel -> fp32 staging [sub][row][h] in the (still unused) dS planes
                        float* stg = reinterpret_cast<float*>(Shi) + ((size_t)ln.sub * R + ng * 8) * H + rc;
#pragma unroll
                        for (int j = 0; j < 8; ++j) stg[j * H] = v[j];
                    } else {
                        cl_store_unit(Xhi, Xlo, cl_unit_off<D, K>(lr, ng), v);
                    }
                }
                if (MODE == 4) gacc[10] += (double)(rs * inv_sx);
                if (MODE <= 2) {
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        float v[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            v[j] = fv[u * 8 + j] * kClSx;
                            amax = fmaxf(amax, fabsf(v[j]));
                        }
                        cl_store_unit(Xhi, Xlo, cl_unit_off<D, K>(64 + lr, hh * 4 + u), v);
                    }
                }
            }
            if (MODE == 2 && !S::U2_MMA) {
                // ---- B': r2 = relu(a2 (W2 r1) + c2) on the CUDA cores: every thread reads the h staged r1 values of
                // its rows (broadcast among the lanes of a sub-tile) and writes its channel's units of X^T
                named_bar_sync(1 + g, kClLanes);
                const float* stg = reinterpret_cast<const float*>(Shi) + (size_t)ln.sub * R * H;
#pragma unroll 1
                for (int u = 0; u < 4; ++u) {
                    const int ng = hh * 4 + u;
                    float v[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float4* rr = reinterpret_cast<const float4*>(stg + (size_t)(ng * 8 + j) * H);
                        float acc = 0.f;
#pragma unroll
                        for (int i4 = 0; i4 < H / 4; ++i4) {
                            const float4 t = rr[i4];
                            acc = fmaf(w2r[i4 * 4 + 0], t.x, acc);
                            acc = fmaf(w2r[i4 * 4 + 1], t.y, acc);
                            acc = fmaf(w2r[i4 * 4 + 2], t.z, acc);
                            acc = fmaf(w2r[i4 * 4 + 3], t.w, acc);
                        }
                        v[j] = fmaxf(fmaf(acc, a2s, b2s), 0.f);
                        amax = fmaxf(amax, v[j]);
                    }
                    cl_store_unit(Xhi, Xlo, cl_unit_off<D, K>(lr, ng), v);
                }
                named_bar_sync(1 + g, kClLanes);          // staging is read: the dS planes may be written again (epilogue C)
            }
            fence_async_smem();
            tc_fence_before_sync();
            named_bar_sync(1 + g, kClLanes);
            if (l == 0) mbar_arrive(&full[g]);

            if (MODE == 2 && S::U2_MMA) {
                // ---- B2: r2 = relu(a2 (W2 r1) + c2) over the r half of X^T (TMEM lanes 0..63 = warps 0, 1 of the group)
                mbar_wait(&done[g], done_phase);
                done_phase ^= 1u;
                tc_fence_after_sync();
                if (ln.part == 0) {
#pragma unroll 1
                    for (int c0 = 0; c0 < R; c0 += 16) {
                        uint32_t u[16];
                        tmem_ld16_nowait(tacc + (uint32_t)c0, u);
                        tmem_ld_wait();
                        float v[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) v[j] = fmaxf(fmaf(__uint_as_float(u[j]), a2s, b2s), 0.f);
                        cl_store_unit(Xhi, Xlo, cl_unit_off<D, K>(l, c0 / 8), v);
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            amax = fmaxf(amax, v[j]);
                            v[j] = fmaxf(fmaf(__uint_as_float(u[8 + j]), a2s, b2s), 0.f);
                        }
                        cl_store_unit(Xhi, Xlo, cl_unit_off<D, K>(l, c0 / 8 + 1), v);
#pragma unroll
                        for (int j = 0; j < 8; ++j) amax = fmaxf(amax, v[j]);
                    }
                }
                fence_async_smem();
                tc_fence_before_sync();
                named_bar_sync(1 + g, kClLanes);
                if (l == 0) mbar_arrive(&full[g]);
            }

            if (MODE == 1 || MODE == 2) {
                // ---- C: softmax, pooled, dS -> operand planes, g A -> TMEM (initial value of dX^T)
                // the upstream gradient of this lane's channel for the tile's points: in flight while the MMA runs
                float gpre[C::PTS];
#pragma unroll
                for (int p = 0; p < C::PTS; ++p) {
                    const long long gp = tile * C::TPTS + ln.sub * C::PTS + p;
                    gpre[p] = (gp < a.npts) ? a.dpooled[gp * D + ln.channel()] : 0.f;
                }
                mbar_wait(&done[g], done_phase);
                done_phase ^= 1u;
                tc_fence_after_sync();
                const float gsw = sg * sw;
                uint32_t un[K];                                        // scores of the NEXT point, loaded while this one is reduced
#pragma unroll
                for (int k0 = 0; k0 < K; k0 += 16) tmem_ld16_nowait(tacc + (uint32_t)k0, *reinterpret_cast<uint32_t(*)[16]>(un + k0));
                tmem_ld_wait();
#pragma unroll 1
                for (int p = 0; p < C::PTS; ++p) {
                    float s[K], x[K];
#pragma unroll
                    for (int k = 0; k < K; ++k) s[k] = __uint_as_float(un[k]);
                    if (p + 1 < C::PTS) {
#pragma unroll
                        for (int k0 = 0; k0 < K; k0 += 16)
                            tmem_ld16_nowait(tacc + (uint32_t)((p + 1) * K + k0), *reinterpret_cast<uint32_t(*)[16]>(un + k0));
                    }
#pragma unroll
                    for (int k0 = 0; k0 < K; k0 += 8) {
                        float t[8];
                        cl_load_unit(Xhi, Xlo, cl_unit_off<D, K>(l, (p * K + k0) / 8), t);
#pragma unroll
                        for (int j = 0; j < 8; ++j) x[k0 + j] = t[j];
                    }
                    float m = fmaxf(s[0], s[1]);
#pragma unroll
                    for (int k = 2; k < K; k += 2) m = fmaxf(m, fmaxf(s[k], s[k + 1]));
                    const float mc = m * cs;
                    float den0 = 0.f, den1 = 0.f, num0 = 0.f, num1 = 0.f;
#pragma unroll
                    for (int k = 0; k < K; k += 2) {
                        s[k] = ex2_approx_b(fmaf(s[k], cs, -mc));
                        s[k + 1] = ex2_approx_b(fmaf(s[k + 1], cs, -mc));
                        den0 += s[k];
                        den1 += s[k + 1];
                        num0 = fmaf(s[k], x[k], num0);
                        num1 = fmaf(s[k + 1], x[k + 1], num1);
                    }
                    const float inv = 1.0f / (den0 + den1);
                    const float pooled = (num0 + num1) * inv;             // scaled by sx like x
                    // the point loop is rolled (code size): select this point's gradient with static indices so that
                    // gpre stays in registers and its loads stay in flight across the accumulator wait (as a local-memory
                    // array the loads had to land before the wait: 11 % of the warp samples of the d = 64 pass 1)
                    float gsel = gpre[0];
#pragma unroll
                    for (int q = 1; q < C::PTS; ++q) gsel = (p == q) ? gpre[q] : gsel;
                    const float gi = gsel * sg * inv;
#pragma unroll
                    for (int k0 = 0; k0 < K; k0 += 16) {
                        uint32_t u[16];
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            const float ga = gi * s[k0 + j];              // sg * g * A
                            x[k0 + j] = ga * (x[k0 + j] - pooled) * inv_sx;   // sg * dS
                            u[j] = __float_as_uint(ga * sw);
                        }
                        tmem_st16(tacc + (uint32_t)(p * K + k0), u);
                    }
#pragma unroll
                    for (int k0 = 0; k0 < K; k0 += 8) {
                        float t[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) t[j] = x[k0 + j];
                        cl_store_unit(Shi, Slo, cl_unit_off<D, K>(l, (p * K + k0) / 8), t);
                    }
                    if (p + 1 < C::PTS) tmem_ld_wait();
                }
                tmem_st_wait();
                fence_async_smem();
                tc_fence_before_sync();
                named_bar_sync(1 + g, kClLanes);
                if (l == 0) mbar_arrive(&full[g]);

                // ---- D: dX^T
                mbar_wait(&done[g], done_phase);
                done_phase ^= 1u;
                tc_fence_after_sync();
                const float inv2 = 1.0f / gsw;
                if (ln.part == 1) {
                    // feature half: dfeat[idx[row]] += dX, lanes = consecutive channels
                    float* df = a.dfeat + ln.c;
                    uint32_t un[16];
                    tmem_ld16_nowait(tacc, un);
                    tmem_ld_wait();
#pragma unroll 1
                    for (int c0 = 0; c0 < R; c0 += 16) {
                        uint32_t u[16];
#pragma unroll
                        for (int j = 0; j < 16; ++j) u[j] = un[j];
                        if (c0 + 16 < R) tmem_ld16_nowait(tacc + (uint32_t)(c0 + 16), un);
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            const uint32_t off = cl_grad_off(ri, c0 + j);
                            if (off != 0xffffffffu) red_add_f32(df + off, __uint_as_float(u[j]) * inv2);
                        }
                        if (c0 + 16 < R) tmem_ld_wait();
                    }
                } else {
                    float part[11];
#pragma unroll
                    for (int q = 0; q < 11; ++q) part[q] = 0.f;
                    float* du2_row = (MODE == 2) ? a.du2_tiles + ((size_t)tile * 64 + l) * R : nullptr;
#pragma unroll 1
                    for (int c0 = 0; c0 < R; c0 += 16) {
                        uint32_t u[16];
                        tmem_ld16_nowait(tacc + (uint32_t)c0, u);
                        tmem_ld_wait();
#pragma unroll
                        for (int hlf = 0; hlf < 2; ++hlf) {
                            float xr[8], du[8];
                            cl_load_unit(Xhi, Xlo, cl_unit_off<D, K>(l, c0 / 8 + hlf), xr);
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                const float dx = __uint_as_float(u[hlf * 8 + j]) * inv2;
                                du[j] = (xr[j] > 0.f) ? dx : 0.f;
                            }
                            if (MODE == 1) {
#pragma unroll
                                for (int j = 0; j < 8; ++j) {
                                    float4 q0, q1, q2;
                                    cl_rpe_row<D, K>(ri, c0 + hlf * 8 + j, q0, q1, q2);
                                    part[0] = fmaf(du[j], q0.x, part[0]); part[1] = fmaf(du[j], q0.y, part[1]);
                                    part[2] = fmaf(du[j], q0.z, part[2]); part[3] = fmaf(du[j], q0.w, part[3]);
                                    part[4] = fmaf(du[j], q1.x, part[4]); part[5] = fmaf(du[j], q1.y, part[5]);
                                    part[6] = fmaf(du[j], q1.z, part[6]); part[7] = fmaf(du[j], q1.w, part[7]);
                                    part[8] = fmaf(du[j], q2.x, part[8]); part[9] = fmaf(du[j], q2.y, part[9]);
                                    part[10] += du[j];
                                }
                            } else {
#pragma unroll
                                for (int j = 0; j < 8; ++j) {
                                    part[0] += du[j];
                                    part[1] = fmaf(du[j], xr[j] * inv_sx, part[1]);
                                    du2max = fmaxf(du2max, fabsf(du[j]));
                                }
                                float* dst = du2_row + c0 + hlf * 8;
                                *reinterpret_cast<float4*>(dst) = make_float4(du[0], du[1], du[2], du[3]);
                                *reinterpret_cast<float4*>(dst + 4) = make_float4(du[4], du[5], du[6], du[7]);
                            }
                        }
                    }
#pragma unroll
                    for (int q = 0; q < 11; ++q) gacc[q] += (double)part[q];
                }
                tc_fence_before_sync();
            }

            if (MODE == 3) {
                // ---- C: dz2 = a2 (du2 - m1 - zhat2 m2) on the r lanes, zero on padding rows
                mbar_wait(&done[g], done_phase);
                done_phase ^= 1u;
                tc_fence_after_sync();
                if (ln.part == 0) {
                    const float* du2_row = a.du2_tiles + ((size_t)tile * 64 + l) * R;
                    const float iu = 1.0f / (kClSx * sw2);
#pragma unroll 1
                    for (int c0 = 0; c0 < R; c0 += 16) {
                        uint32_t u[16];
                        tmem_ld16_nowait(tacc + (uint32_t)c0, u);
                        tmem_ld_wait();
#pragma unroll
                        for (int hlf = 0; hlf < 2; ++hlf) {
                            const float4 d0 = *reinterpret_cast<const float4*>(du2_row + c0 + hlf * 8);
                            const float4 d1 = *reinterpret_cast<const float4*>(du2_row + c0 + hlf * 8 + 4);
                            const float dv[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
                            float t[8];
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                const float zh = (__uint_as_float(u[hlf * 8 + j]) * iu - bn_mean) * bn_rstd;
                                const bool valid = cl_grad_off(ri, c0 + hlf * 8 + j) != 0xffffffffu;
                                t[j] = valid ? bn_a2 * (dv[j] - bn_m1 - zh * bn_m2) * sg : 0.f;
                            }
                            cl_store_unit(Shi, Slo, cl_unit_off<D, K>(l, c0 / 8 + hlf), t);
                        }
                    }
                }
                fence_async_smem();
                tc_fence_before_sync();
                named_bar_sync(1 + g, kClLanes);
                if (l == 0) mbar_arrive(&full[g]);
                // ---- D: du1 = dr1 [r1 > 0], G1 += du1 (x) [rpe, 1]
                mbar_wait(&done[g], done_phase);
                done_phase ^= 1u;
                tc_fence_after_sync();
                if (ln.part == 0) {
                    const float inv2 = 1.0f / (sg * sw2);
                    float part[11];
#pragma unroll
                    for (int q = 0; q < 11; ++q) part[q] = 0.f;
#pragma unroll 1
                    for (int c0 = 0; c0 < R; c0 += 16) {
                        uint32_t u[16];
                        tmem_ld16_nowait(tacc + (uint32_t)c0, u);
                        tmem_ld_wait();
#pragma unroll
                        for (int hlf = 0; hlf < 2; ++hlf) {
                            float xr[8];
                            cl_load_unit(Xhi, Xlo, cl_unit_off<D, K>(l, c0 / 8 + hlf), xr);
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                const float du = (xr[j] > 0.f) ? __uint_as_float(u[hlf * 8 + j]) * inv2 : 0.f;
                                float4 q0, q1, q2;
                                cl_rpe_row<D, K>(ri, c0 + hlf * 8 + j, q0, q1, q2);
                                part[0] = fmaf(du, q0.x, part[0]); part[1] = fmaf(du, q0.y, part[1]);
                                part[2] = fmaf(du, q0.z, part[2]); part[3] = fmaf(du, q0.w, part[3]);
                                part[4] = fmaf(du, q1.x, part[4]); part[5] = fmaf(du, q1.y, part[5]);
                                part[6] = fmaf(du, q1.z, part[6]); part[7] = fmaf(du, q1.w, part[7]);
                                part[8] = fmaf(du, q2.x, part[8]); part[9] = fmaf(du, q2.y, part[9]);
                                part[10] += du;
                            }
                        }
                    }
#pragma unroll
                    for (int q = 0; q < 11; ++q) gacc[q] += (double)part[q];
                }
                tc_fence_before_sync();
            }

            if (MODE == 4) {
                mbar_wait(&done[g], done_phase);          // the MMA has read X: the planes may be overwritten
                done_phase ^= 1u;
                tc_fence_after_sync();
            }
            ++ntile;
            if (ntile % kClFlush == 0) fold();
            // the epilogues read the row-info table: nobody may start the next tile's table before everyone is done
            if (MODE <= 3) named_bar_sync(1 + g, kClLanes);
        }
        if (ntile % kClFlush != 0) fold();
        if (a.status != nullptr && !(amax < 65504.f)) atomicOr(a.status, 1);

        // ---- per-lane sums -> global
        if (MODE == 1 || MODE == 3) {
            if (ln.part == 0) {
#pragma unroll
                for (int q = 0; q < 11; ++q) atomicAdd(a.g1 + rc * 16 + q, gacc[q]);
            }
        } else if (MODE == 2) {
            if (ln.part == 0) {
                atomicAdd(a.sum_du2 + rc, gacc[0]);
                atomicAdd(a.sum_du2 + H + rc, gacc[1]);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) du2max = fmaxf(du2max, __shfl_xor_sync(0xffffffffu, du2max, o));
            if ((l & 31) == 0) atomicMax(reinterpret_cast<unsigned int*>(a.scal + 1), __float_as_uint(du2max));
        } else {
            // centred sums of this CTA per virtual lane, valid rows per sub-tile (shared-memory fp64 atomics: a few hundred)
            atomicAdd(&csum[lr], gacc[10]);
            if (rc == 0) atomicAdd(&cnum[ln.sub], nvalid);
        }

        // ---- second-level accumulator -> global (group 0, after every group's last fold)
        named_bar_sync(15, NG * kClLanes);
        tc_fence_after_sync();
        if (g == 0) {
            // lane = virtual output channel; only the columns of its own sub-tile's block are real.  tcgen05.ld takes a
            // warp-uniform address, so every lane reads all columns and keeps its block.
            if (MODE <= 2) {
                const float unscale = 1.0f / (sg * kClSx);
                float* out = a.dw_score + (size_t)ln.channel() * D;
#pragma unroll 1
                for (int c0 = 0; c0 < kClLanes; c0 += 16) {
                    uint32_t u[16];
                    tmem_ld16_nowait(tacc2 + (uint32_t)c0, u);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const ClLane<D> li(c0 + j);
                        if (li.sub == ln.sub) atomicAdd(out + li.channel(), __uint_as_float(u[j]) * unscale);
                    }
                }
            } else if (ln.part == 0) {
                const float unscale = (MODE == 3) ? 1.0f / (sg * kClSx) : 1.0f / (kClSx * kClSx);
                double* out = (MODE == 3 ? a.dw2 : a.m_r1) + (size_t)rc * H;
                // MODE 4: back from centred to raw moments in fp64 (values below are true-scale: centre and sums / sx)
                const double ci = (MODE == 4) ? (double)centre[l] / kClSx : 0.0, si = (MODE == 4) ? csum[l] : 0.0;
                const double nsub = (MODE == 4) ? cnum[ln.sub] : 0.0;
                if (MODE == 4) atomicAdd(a.s_r1 + rc * 16 + 10, si + nsub * ci);
#pragma unroll 1
                for (int c0 = 0; c0 < 64; c0 += 16) {
                    uint32_t u[16];
                    tmem_ld16_nowait(tacc2 + (uint32_t)c0, u);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        if ((c0 + j) / H == ln.sub) {
                            double v = (double)(__uint_as_float(u[j]) * unscale);
                            if (MODE == 4) {
                                const double cj = (double)centre[c0 + j] / kClSx, sj = csum[c0 + j];
                                v += ci * sj + si * cj + nsub * ci * cj;
                            }
                            atomicAdd(out + (c0 + j) % H, v);
                        }
                }
            }
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == NG * 4) tmem_dealloc_warp(tmem, S::TMEM_COLS);
}

// absmax |x| of n floats -> *out (atomic max on the bit pattern; *out zero-filled by the caller)
__global__ void __launch_bounds__(256) absmax_kernel(const float* __restrict__ x, long long n, float* out) {
    float m = 0.f;
    const long long n4 = n / 4;
    const float4* x4 = reinterpret_cast<const float4*>(x);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        const float4 v = x4[i];
        m = fmaxf(m, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
    }
    for (long long i = n4 * 4 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        m = fmaxf(m, fabsf(x[i]));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(reinterpret_cast<unsigned int*>(out), __float_as_uint(m));
}

template <int D, int K, int MODE, int NG>
static int launch_cl_bwd(const LfaClBwdArgs& a, cudaStream_t st) {
    auto kern = lfa_cl_bwd_kernel<D, K, MODE, NG>;
    constexpr size_t smem = LfaClBwdSmem<D, K, MODE, NG>::BYTES;
    static_assert(smem <= 232448, "tile does not fit shared memory");
    R3D_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int grid = (int)(a.ntiles < kNumSMs ? a.ntiles : kNumSMs);
    kern<<<grid, (NG * 4 + 1) * 32, smem, st>>>(a);
    R3D_LAUNCH_CHECK("lfa_cl_bwd_kernel");
    return R3D_OK;
}

// worker groups per CTA by mode and width (shared-memory budget: weight images + per-group operand planes + row info)
constexpr int cl_bwd_groups(int mode, int d) { return mode <= 2 ? 2 : (mode == 3 ? (d >= 64 ? 3 : 2) : 3); }

template <int MODE>
static int dispatch_cl_bwd(LfaClBwdArgs& a, int K, int d, cudaStream_t st) {
#define R3D_CLB_CASE(DD, KK)                                                      \
    if (d == DD && K == KK) {                                                     \
        a.ntiles = (a.npts + ClCfg<DD, KK>::TPTS - 1) / ClCfg<DD, KK>::TPTS;      \
        return launch_cl_bwd<DD, KK, MODE, cl_bwd_groups(MODE, DD)>(a, st);       \
    }
    R3D_CLB_CASE(128, 16) R3D_CLB_CASE(64, 16) R3D_CLB_CASE(32, 16) R3D_CLB_CASE(16, 16)
    R3D_CLB_CASE(128, 32) R3D_CLB_CASE(64, 32) R3D_CLB_CASE(32, 32) R3D_CLB_CASE(16, 32)
#undef R3D_CLB_CASE
    return R3D_EUNSUPPORTED;
}

}  // namespace r3d

using namespace r3d;

extern "C" long long r3d_lfa_tc_du2_floats(int B, int N, int K, int d) {
    if (B <= 0 || N <= 0 || K <= 0 || d <= 0 || d > 128 || 128 % d != 0 || 64 % K != 0) return 0;
    const long long tpts = (128 / d) * (64 / K);
    const long long ntiles = ((long long)B * N + tpts - 1) / tpts;
    return ntiles * 64 * 64;
}

extern "C" int r3d_absmax(const float* x, long long n, float* out, r3d_stream_t stream) {
    if (n < 0 || !out || (n > 0 && !x)) return R3D_EINVAL;
    if (n == 0) return R3D_OK;
    if (!is_aligned(x, 16)) return R3D_EALIGN;
    long long blocks = (n / 4 + 255) / 256;
    if (blocks > 4 * kNumSMs) blocks = 4 * kNumSMs;
    if (blocks < 1) blocks = 1;
    absmax_kernel<<<(int)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(x, n, out);
    R3D_LAUNCH_CHECK("absmax_kernel");
    return R3D_OK;
}

extern "C" int r3d_lfa_tc_bwd(int mode, const float* xyz, long long xyz_bstride, const int32_t* idx, const float* feat,
                              long long feat_bstride, const float* w_rpe1, const float* a_rpe1, const float* b_rpe1,
                              const float* w_rpe2, const float* a_rpe2, const float* b_rpe2, const float* w_score,
                              const float* dpooled, float* dfeat, long long dfeat_bstride, float* dw_score, double* g1,
                              float* du2_tiles, double* sum_du2, const float* bn2, double* dw2, double* m_r1, double* s_r1,
                              float* scal, int* status, int B, int N, int K, int d, r3d_stream_t stream) {
    if (mode < 1 || mode > 4) return R3D_EINVAL;
    if (B < 0 || N < 0 || K <= 0 || d <= 0) return R3D_EINVAL;
    if (B == 0 || N == 0) return R3D_OK;
    if (!xyz || !idx || !w_rpe1 || !a_rpe1 || !b_rpe1) return R3D_EINVAL;
    if (mode <= 2 && (!feat || !w_score || !dpooled || !dfeat || !dw_score || !scal)) return R3D_EINVAL;
    if (mode == 1 && !g1) return R3D_EINVAL;
    if (mode == 2 && (!w_rpe2 || !a_rpe2 || !b_rpe2 || !du2_tiles || !sum_du2)) return R3D_EINVAL;
    if (mode == 3 && (!w_rpe2 || !du2_tiles || !bn2 || !dw2 || !g1 || !scal)) return R3D_EINVAL;
    if (mode == 4 && (!m_r1 || !s_r1)) return R3D_EINVAL;
    const int h = d / 2;
    if (xyz_bstride == 0) xyz_bstride = (long long)N * 3;
    if (feat_bstride == 0) feat_bstride = (long long)N * h;
    if (dfeat_bstride == 0) dfeat_bstride = (long long)N * h;
    if ((long long)(B - 1) * feat_bstride + (long long)N * h >= (1ll << 32) - 1 ||
        (long long)(B - 1) * dfeat_bstride + (long long)N * h >= (1ll << 32) - 1)
        return R3D_EUNSUPPORTED;
    if (du2_tiles && !is_aligned(du2_tiles, 16)) return R3D_EALIGN;
    LfaClBwdArgs a{xyz, xyz_bstride, idx, feat, feat_bstride, w_rpe1, a_rpe1, b_rpe1, w_rpe2, a_rpe2, b_rpe2, w_score,
                   dpooled, dfeat, dfeat_bstride, dw_score, g1, du2_tiles, sum_du2, bn2, dw2, m_r1, s_r1, scal, status,
                   N, (long long)B * N, 0};
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    switch (mode) {
        case 1: return dispatch_cl_bwd<1>(a, K, d, st);
        case 2: return dispatch_cl_bwd<2>(a, K, d, st);
        case 3: return dispatch_cl_bwd<3>(a, K, d, st);
        default: return dispatch_cl_bwd<4>(a, K, d, st);
    }
}
