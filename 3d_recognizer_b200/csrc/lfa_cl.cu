// Fused LocSE + attentive pooling, FORWARD, on the tcgen05 tensor cores — "channel-lane" kernel (see
// lfa_cl_common.cuh for the design).  Same operator as lfa.cu (randlanet/utils/modules.py:316-323):
//   rpe (:170-186) -> r1 = relu(bn(W1 rpe)) (:317) [-> r2 = relu(bn(W2 r1)) (:321), stage 2] -> x = [r ; F[idx]] (:209-221)
//   -> s = Ws x, a = softmax over K, pooled = sum_K a * x (:246-252).
//
// Persistent CTA = NG worker groups of 128 threads + one MMA warp.  Per tile (SUB sub-tiles of R = 64 rows) a group
//   A  builds the row-info table (rpe[10] + neighbour offset per row; thread = row),
//   B  produces the row operand X^T (thread = virtual channel: r lanes evaluate mlp_rpe1 from the table with their W1
//      row in registers, F lanes gather feat[idx] — coalesced across the lanes), split into fp16 hi/lo planes,
//   B2 (stage 2) reads U2^T = W2 r1^T from TMEM, applies BatchNorm affine + ReLU and overwrites the r half of X^T,
//   D  reads S^T = Ws X^T from TMEM (lane = channel, columns = rows), softmax over each point's K columns and the
//      weighted sum in registers, writes pooled[point][channel].
// The MMA thread issues U2 / S for whichever group has its operand ready (M = 128, N = 64, K = 16 per instruction,
// three split products per K step) and commits to the group's mbarrier.
#include "lfa_cl_common.cuh"

namespace r3d {

struct LfaClArgs {
    const float* xyz;
    long long xyz_bstride;
    const int32_t* idx;
    const float* feat;
    long long feat_bstride;
    const float* w_rpe1;    // (h,10)
    const float* a_rpe1;
    const float* b_rpe1;
    const float* w_rpe2;    // (h,h) [out][in]   (stage 2)
    const float* a_rpe2;
    const float* b_rpe2;
    const float* w_score;   // (d,d) [out][in]
    float* pooled;          // (B,N,d)
    int* status;            // nullable: bit 0 <- an activation left the fp16x2 range
    int N;
    long long npts;         // B * N
    long long ntiles;
};

template <int D, int K, int STAGE, int NG>
struct LfaClFwdSmem {
    using C = ClCfg<D, K>;
    static constexpr int W2_BYTES = (STAGE == 2) ? kClW2Bytes : 0;              // compact 64 x 64 image, one plane
    static constexpr int GROUP_BYTES = 2 * C::OP_BYTES + C::RINFO_FLOATS * 4;
    static constexpr int OFF_W2 = 2 * C::W_BYTES;
    static constexpr int OFF_GROUPS = OFF_W2 + 2 * W2_BYTES;
    static constexpr int OFF_BARS = OFF_GROUPS + NG * GROUP_BYTES;
    static constexpr size_t BYTES = (size_t)OFF_BARS + 2 * NG * 8 + 16 + 34 * 4;
};

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

template <int D, int K, int STAGE, int NG>
__global__ void __launch_bounds__((NG * 4 + 1) * 32, 1) lfa_cl_fwd_kernel(LfaClArgs a) {
    using C = ClCfg<D, K>;
    using S = LfaClFwdSmem<D, K, STAGE, NG>;
    constexpr int H = C::H, R = C::R;
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* Whi = smem;
    unsigned char* Wlo = smem + C::W_BYTES;
    unsigned char* W2hi = smem + S::OFF_W2;
    unsigned char* W2lo = W2hi + S::W2_BYTES;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + S::OFF_BARS);
    uint64_t* done = full + NG;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + NG);
    float* red = reinterpret_cast<float*>(tmem_slot + 4);

    const int tid = threadIdx.x, warp = tid >> 5;
    constexpr uint32_t TMEM_COLS = (NG * R <= 64) ? 64 : ((NG * R <= 128) ? 128 : ((NG * R <= 256) ? 256 : 512));

    if (tid == 0) {
        for (int g = 0; g < NG; ++g) {
            mbar_init(&full[g], 1);
            mbar_init(&done[g], 1);
        }
        mbar_fence_init();
    }
    if (warp == NG * 4) tmem_alloc_warp(tmem_slot, TMEM_COLS);
    // weights: power-of-two scales from their absmax, block-diagonal fp16 hi/lo images
    const float sw = cl_pow2_scale(cl_block_absmax(a.w_score, D * D, red));
    cl_build_weight_image<D>(a.w_score, sw, Whi, Wlo);
    float sw2 = 1.f;
    if (STAGE == 2) {
        sw2 = cl_pow2_scale(cl_block_absmax(a.w_rpe2, H * H, red));
        cl_build_w2_image<D>(a.w_rpe2, sw2, W2hi, W2lo);
    }
    fence_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem = *tmem_slot;

    if (warp == NG * 4) {
        // ============================================================ MMA issuer
        if ((tid & 31) == 0) {
            constexpr int NR = (STAGE == 2) ? 2 : 1;                 // MMA rounds per tile
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
            const uint32_t idesc = umma_idesc_f16(kClLanes, R, 0, 1);     // A K-major (weights), B MN-major (rows)
            while (total > 0) {
#pragma unroll
                for (int g = 0; g < NG; ++g) {
                    if (left[g] > 0 && mbar_try_wait(&full[g], ph[g])) {
                        tc_fence_after_sync();
                        const uint32_t xhi = smem_u32(smem + S::OFF_GROUPS + g * S::GROUP_BYTES);
                        const uint32_t xlo = xhi + C::OP_BYTES;
                        const uint32_t dcol = tmem + (uint32_t)(g * R);
                        if (STAGE == 2 && (step[g] & 1) == 0)
                            cl_mma_3x(dcol, smem_u32(W2hi), smem_u32(W2lo), kClW2Is, 128, xhi, xlo, C::OP_CS, 128, idesc, 4,
                                      false);
                        else
                            cl_mma_3x(dcol, smem_u32(Whi), smem_u32(Wlo), C::W_IS, 128, xhi, xlo, C::OP_CS, 128, idesc, 8,
                                      false);
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
        unsigned char* Xhi = smem + S::OFF_GROUPS + g * S::GROUP_BYTES;
        unsigned char* Xlo = Xhi + C::OP_BYTES;
        float* rinfo = reinterpret_cast<float*>(Xlo + C::OP_BYTES);
        const float* ri = rinfo + ln.sub * C::SUB_RI;
        const int lr = l & 63, hh = l >> 6;          // producer role: r lane lr and F lane 64 + lr, row groups [4 hh, 4 hh + 4)
        const int rc = lr % H;
        float w1[10], a2s = 0.f, b2s = 0.f;
#pragma unroll
        for (int q = 0; q < 10; ++q) w1[q] = a.w_rpe1[rc * 10 + q];
        const float a1s = a.a_rpe1[rc] * kClSx, b1s = a.b_rpe1[rc] * kClSx;
        if (STAGE == 2) {
            a2s = a.a_rpe2[rc] / sw2;                  // u2 = acc / (sx sw2); r2 * sx = relu(acc * a2/sw2 + b2 * sx)
            b2s = a.b_rpe2[rc] * kClSx;
        }
        const uint32_t tbase = tmem + ((uint32_t)((l >> 5) * 32) << 16) + (uint32_t)(g * R);
        const float cs = 1.4426950408889634f / (kClSx * sw);     // log2(e) / (sx sw): scores leave the MMA scaled
        const float inv_sx = 1.0f / kClSx;
        float amax = 0.f;
        uint32_t done_phase = 0;
        for (int it = 0;; ++it) {
            const long long tile = (long long)blockIdx.x + (long long)(it * NG + g) * gridDim.x;
            if (tile >= a.ntiles) break;
            // ---- A: row info
            cl_row_info<D, K>(rinfo, a.xyz, a.xyz_bstride, a.idx, a.feat_bstride, 0, a.N, a.npts, tile, l);
            named_bar_sync(1 + g, kClLanes);
            // ---- B: row operand X^T (scaled by sx).  Thread pair (l, l ^ 64) shares r lane lr and F lane 64 + lr, half
            // the rows each: every thread carries the same mix of gather latency and mlp_rpe1 arithmetic.
            {
                float fv[32];
                const float* fb = a.feat + rc;
#pragma unroll
                for (int j = 0; j < 32; ++j) fv[j] = fb[cl_feat_off(ri, hh * 32 + j)];
#pragma unroll 1
                for (int u = 0; u < 4; ++u) {
                    const int ng = hh * 4 + u;
                    float v[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        float4 q0, q1, q2;
                        cl_rpe_row<D, K>(ri, ng * 8 + j, q0, q1, q2);
                        v[j] = cl_mlp1(w1, a1s, b1s, q0, q1, q2);
                    }
                    amax = fmaxf(amax, fmaxf(fmaxf(fmaxf(v[0], v[1]), fmaxf(v[2], v[3])), fmaxf(fmaxf(v[4], v[5]), fmaxf(v[6], v[7]))));
                    cl_store_unit(Xhi, Xlo, cl_unit_off<D, K>(lr, ng), v);
                }
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
            fence_async_smem();
            tc_fence_before_sync();
            named_bar_sync(1 + g, kClLanes);
            if (l == 0) mbar_arrive(&full[g]);
            if (STAGE == 2) {
                // ---- B2: r2 = relu(a2 (W2 r1) + b2) over the r half of X^T
                mbar_wait(&done[g], done_phase);
                done_phase ^= 1u;
                tc_fence_after_sync();
                if (ln.part == 0) {
#pragma unroll 1
                    for (int c0 = 0; c0 < R; c0 += 16) {
                        uint32_t u[16];
                        tmem_ld16_nowait(tbase + (uint32_t)c0, u);
                        tmem_ld_wait();
                        float v[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            v[j] = fmaxf(fmaf(__uint_as_float(u[j]), a2s, b2s), 0.f);
                            amax = fmaxf(amax, v[j]);
                        }
                        cl_store_unit(Xhi, Xlo, cl_unit_off<D, K>(l, c0 / 8), v);
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            v[j] = fmaxf(fmaf(__uint_as_float(u[8 + j]), a2s, b2s), 0.f);
                            amax = fmaxf(amax, v[j]);
                        }
                        cl_store_unit(Xhi, Xlo, cl_unit_off<D, K>(l, c0 / 8 + 1), v);
                    }
                }
                fence_async_smem();
                tc_fence_before_sync();
                named_bar_sync(1 + g, kClLanes);
                if (l == 0) mbar_arrive(&full[g]);
            }
            // ---- D: softmax over K + weighted sum, thread = channel
            mbar_wait(&done[g], done_phase);
            done_phase ^= 1u;
            tc_fence_after_sync();
            uint32_t un[K];                                            // scores of the NEXT point, loaded while this one is reduced
#pragma unroll
            for (int k0 = 0; k0 < K; k0 += 16) tmem_ld16_nowait(tbase + (uint32_t)k0, *reinterpret_cast<uint32_t(*)[16]>(un + k0));
            tmem_ld_wait();
#pragma unroll
            for (int p = 0; p < C::PTS; ++p) {
                float s[K], x[K];
#pragma unroll
                for (int k = 0; k < K; ++k) s[k] = __uint_as_float(un[k]);
                if (p + 1 < C::PTS) {
#pragma unroll
                    for (int k0 = 0; k0 < K; k0 += 16)
                        tmem_ld16_nowait(tbase + (uint32_t)((p + 1) * K + k0), *reinterpret_cast<uint32_t(*)[16]>(un + k0));
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
                    const float e0 = ex2_approx(fmaf(s[k], cs, -mc)), e1 = ex2_approx(fmaf(s[k + 1], cs, -mc));
                    den0 += e0;
                    den1 += e1;
                    num0 = fmaf(e0, x[k], num0);
                    num1 = fmaf(e1, x[k + 1], num1);
                }
                const long long gp = tile * C::TPTS + ln.sub * C::PTS + p;
                if (gp < a.npts) a.pooled[gp * D + ln.channel()] = ((num0 + num1) * inv_sx) / (den0 + den1);
                if (p + 1 < C::PTS) tmem_ld_wait();
            }
            tc_fence_before_sync();
        }
        if (a.status != nullptr && !(amax < 65504.f)) atomicOr(a.status, 1);
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == NG * 4) tmem_dealloc_warp(tmem, TMEM_COLS);
}

template <int D, int K, int STAGE, int NG>
static int launch_cl_fwd(const LfaClArgs& a, cudaStream_t st) {
    auto kern = lfa_cl_fwd_kernel<D, K, STAGE, NG>;
    constexpr size_t smem = LfaClFwdSmem<D, K, STAGE, NG>::BYTES;
    static_assert(smem <= 232448, "tile does not fit shared memory");
    R3D_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int grid = (int)(a.ntiles < kNumSMs ? a.ntiles : kNumSMs);
    kern<<<grid, (NG * 4 + 1) * 32, smem, st>>>(a);
    R3D_LAUNCH_CHECK("lfa_cl_fwd_kernel");
    return R3D_OK;
}

}  // namespace r3d

using namespace r3d;

extern "C" int r3d_lfa_pool_tc(int stage, const float* xyz, long long xyz_bstride, const int32_t* idx, const float* feat,
                               long long feat_bstride, const float* w_rpe1, const float* a_rpe1, const float* b_rpe1,
                               const float* w_rpe2, const float* a_rpe2, const float* b_rpe2, const float* w_score,
                               float* pooled, int* status, int B, int N, int K, int d, r3d_stream_t stream) {
    if (stage != 1 && stage != 2) return R3D_EINVAL;
    if (B < 0 || N < 0 || K <= 0 || d <= 0) return R3D_EINVAL;
    if (B == 0 || N == 0) return R3D_OK;
    if (!xyz || !idx || !feat || !w_rpe1 || !a_rpe1 || !b_rpe1 || !w_score || !pooled) return R3D_EINVAL;
    if (stage == 2 && (!w_rpe2 || !a_rpe2 || !b_rpe2)) return R3D_EINVAL;
    const int h = d / 2;
    if (xyz_bstride == 0) xyz_bstride = (long long)N * 3;
    if (feat_bstride == 0) feat_bstride = (long long)N * h;
    // neighbour feature offsets are carried as 32-bit element offsets in the row-info table
    if ((long long)(B - 1) * feat_bstride + (long long)N * h >= (1ll << 32)) return R3D_EUNSUPPORTED;
    LfaClArgs a{xyz, xyz_bstride, idx, feat, feat_bstride, w_rpe1, a_rpe1, b_rpe1, w_rpe2, a_rpe2, b_rpe2, w_score,
                pooled, status, N, (long long)B * N, 0};
    cudaStream_t st = static_cast<cudaStream_t>(stream);
#define R3D_CL_CASE(DD, KK, NG1, NG2)                                                       \
    if (d == DD && K == KK) {                                                               \
        a.ntiles = (a.npts + ClCfg<DD, KK>::TPTS - 1) / ClCfg<DD, KK>::TPTS;                \
        return stage == 1 ? launch_cl_fwd<DD, KK, 1, NG1>(a, st) : launch_cl_fwd<DD, KK, 2, NG2>(a, st); \
    }
    R3D_CL_CASE(128, 16, 4, 4) R3D_CL_CASE(64, 16, 4, 3) R3D_CL_CASE(32, 16, 4, 3) R3D_CL_CASE(16, 16, 3, 3)
    R3D_CL_CASE(128, 32, 4, 4) R3D_CL_CASE(64, 32, 4, 3) R3D_CL_CASE(32, 32, 4, 3) R3D_CL_CASE(16, 32, 3, 3)
#undef R3D_CL_CASE
    return R3D_EUNSUPPORTED;
}
