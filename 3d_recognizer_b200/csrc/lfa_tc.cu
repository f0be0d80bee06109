// Fused LocSE + attentive pooling, forward, with the score GEMM (and mlp_rpe2) on the tcgen05 tensor cores.
//
// Same operator as lfa.cu (randlanet/utils/modules.py:316-323) for the widths where the d x d score Linear is a
// real dense contraction (d = 64, 128).  fp32 parity forbids plain TF32 (SURVEY.md §7 hard part 2), so the kernel
// uses the 3xTF32 scheme of tc_gemm.cu: operands split into hi/lo, three kind::tf32 MMAs per K step, fp32
// accumulation in TMEM (measured error ~1e-6 relative, the same as an fp32 FMA chain).
//
// CTA = 128 threads = 128 (point, neighbour) rows (128/K points).  Thread t owns row t everywhere:
//   prologue  gathers its neighbour, evaluates the position encoding and mlp_rpe1, and writes its row of
//             X = [r1 ; F[idx]] — already split into hi/lo — straight into the canonical K-major no-swizzle UMMA
//             layout ([channel/4][row][4 channels], 16-byte units: LBO = 128, SBO = 8);
//   GEMMs     weights are staged per 32-channel block ([k/4][out][4], LBO = N, SBO = 8; split on the fly, two
//             stages); one thread issues the MMAs (M=128, N=d or h, K=8) and commits them to mbarriers, so staging
//             block kb+1 overlaps the tensor core on block kb.  Stage 2 first runs W2 r1 into a second TMEM region,
//             applies BatchNorm affine + ReLU in registers and overwrites the r1 half of X in place;
//   epilogue  reads its row of scores from TMEM (tcgen05.ld, lane = row), does the softmax over the K rows of a
//             point with warp shuffles (the rows of a point are K consecutive lanes) and the weighted sum.
#include "lfa_common.cuh"
#include "tc_common.cuh"

namespace r3d {

constexpr int kTcRows = 128;

struct LfaTcArgs {
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
    int B, N;
};

template <int D, int K>
struct LfaTcSmem {
    static constexpr int H = D / 2;
    static constexpr int X_F4 = (D / 4) * kTcRows;            // float4 per hi (or lo) copy of X
    static constexpr int W_F4 = 8 * D;                        // float4 per staged weight block (32 channels x D outputs)
    static constexpr int PTS = kTcRows / K;
    static constexpr size_t BYTES = (size_t)(2 * X_F4 + 2 * 2 * W_F4) * 16 + (size_t)(H * 16 + PTS * D) * 4 + 4 * 8 + 16;
};

template <int D, int K, int STAGE>
__global__ void __launch_bounds__(kTcRows, 1) lfa_pool_tc_kernel(LfaTcArgs a) {
    using S = LfaTcSmem<D, K>;
    constexpr int H = S::H;
    constexpr int PTS = S::PTS;
    constexpr uint32_t TMEM_COLS = (D + (STAGE == 2 ? H : 0)) <= 64 ? 64 : ((D + (STAGE == 2 ? H : 0)) <= 128 ? 128 : 256);
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float4* Xhi = reinterpret_cast<float4*>(smem_raw);        // [D/4][128]
    float4* Xlo = Xhi + S::X_F4;
    float4* Whi = Xlo + S::X_F4;                              // [2][8][N]
    float4* Wlo = Whi + 2 * S::W_F4;
    float* Pw1 = reinterpret_cast<float*>(Wlo + 2 * S::W_F4); // [H][12]
    float* Pa1 = Pw1 + H * 12;
    float* Pb1 = Pa1 + H;
    float* Pa2 = Pb1 + H;
    float* Pb2 = Pa2 + H;
    float* outs = Pb2 + H;                                    // [PTS][D]
    uint64_t* bars = reinterpret_cast<uint64_t*>(outs + PTS * D);   // [0],[1] stage free, [2] GEMM done
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int b = blockIdx.y;
    const int p0 = blockIdx.x * PTS;

    if (tid == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        mbar_init(&bars[2], 1);
        mbar_fence_init();
    }
    if (warp == 0) tmem_alloc_warp(tmem_slot, TMEM_COLS);
    load_rpe1_params<H, kTcRows>(Pw1, Pa1, Pb1, a.w_rpe1, a.a_rpe1, a.b_rpe1, tid);
    if (STAGE == 2) {
        for (int i = tid; i < H; i += kTcRows) {
            Pa2[i] = a.a_rpe2[i];
            Pb2[i] = a.b_rpe2[i];
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem = *tmem_slot;

    // ------------------------------------------------------------------ prologue: row tid of X = [r1 ; F[idx]]
    {
        const int p = tid / K, k = tid % K;
        const int pi = min(p0 + p, a.N - 1);
        const int pj = a.idx[((size_t)b * a.N + pi) * K + k];
        const float* xyz_b = a.xyz + (size_t)b * a.xyz_bstride;
        float rpe[10];
        rpe_of_row(xyz_b, pi, pj, rpe);
#pragma unroll 2
        for (int c = 0; c < H; c += 4) {
            float4 v;
            v.x = rpe_mlp1(Pw1, Pa1, Pb1, c + 0, rpe);
            v.y = rpe_mlp1(Pw1, Pa1, Pb1, c + 1, rpe);
            v.z = rpe_mlp1(Pw1, Pa1, Pb1, c + 2, rpe);
            v.w = rpe_mlp1(Pw1, Pa1, Pb1, c + 3, rpe);
            float4 hi, lo;
            split_tf32(v, hi, lo);
            Xhi[(c / 4) * kTcRows + tid] = hi;
            Xlo[(c / 4) * kTcRows + tid] = lo;
        }
        const float* frow = a.feat + (size_t)b * a.feat_bstride + (size_t)pj * H;
#pragma unroll 4
        for (int c = 0; c < H; c += 4) {
            const float4 v = *reinterpret_cast<const float4*>(frow + c);
            float4 hi, lo;
            split_tf32(v, hi, lo);
            Xhi[((H + c) / 4) * kTcRows + tid] = hi;
            Xlo[((H + c) / 4) * kTcRows + tid] = lo;
        }
    }

    // ------------------------------------------------------------------ streamed 3xTF32 GEMM: TMEM[col0..col0+N) = X[:, :Kred] W^T
    uint32_t blk = 0;        // weight blocks staged so far (ring position / barrier phases)
    uint32_t done_phase = 0;
    auto gemm = [&](const float* __restrict__ Wg, int N, int Kred, uint32_t col0) {
        const uint32_t idesc = umma_idesc_tf32(kTcRows, N);
        const int nkb = Kred / 32;
        for (int kb = 0; kb < nkb; ++kb, ++blk) {
            const uint32_t s = blk & 1u;
            if (blk >= 2) mbar_wait(&bars[s], ((blk >> 1) - 1u) & 1u);    // the MMAs that read this stage are done
            for (int n = tid; n < N; n += kTcRows) {
                const float* row = Wg + (size_t)n * Kred + kb * 32;
#pragma unroll
                for (int kq = 0; kq < 8; ++kq) {
                    const float4 v = *reinterpret_cast<const float4*>(row + kq * 4);
                    float4 hi, lo;
                    split_tf32(v, hi, lo);
                    Whi[(s * 8 + kq) * N + n] = hi;
                    Wlo[(s * 8 + kq) * N + n] = lo;
                }
            }
            fence_async_smem();
            __syncthreads();
            if (tid == 0) {
                tc_fence_after_sync();
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                    const int kq = kb * 8 + 2 * ks;
                    const uint64_t ah = umma_desc(Xhi + kq * kTcRows, kTcRows, 8);
                    const uint64_t al = umma_desc(Xlo + kq * kTcRows, kTcRows, 8);
                    const uint64_t wh = umma_desc(Whi + (s * 8 + 2 * ks) * N, (uint32_t)N, 8);
                    const uint64_t wl = umma_desc(Wlo + (s * 8 + 2 * ks) * N, (uint32_t)N, 8);
                    umma_tf32(tmem + col0, ah, wh, idesc, (kb > 0 || ks > 0) ? 1u : 0u);
                    umma_tf32(tmem + col0, ah, wl, idesc, 1u);
                    umma_tf32(tmem + col0, al, wh, idesc, 1u);
                }
                umma_commit(&bars[s]);
                if (kb == nkb - 1) umma_commit(&bars[2]);
            }
        }
        mbar_wait(&bars[2], done_phase);
        done_phase ^= 1u;
        tc_fence_after_sync();
    };

    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;

    // ------------------------------------------------------------------ stage 2: r2 = relu(a2 (W2 r1) + b2) over r1, in place
    if (STAGE == 2) {
        gemm(a.w_rpe2, H, H, D);
#pragma unroll 1
        for (int c0 = 0; c0 < H; c0 += 32) {
            float v[32];
            tmem_ld32(tmem + lane_base + (uint32_t)(D + c0), v);
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
                float4 r;
                r.x = fmaxf(fmaf(v[j + 0], Pa2[c0 + j + 0], Pb2[c0 + j + 0]), 0.f);
                r.y = fmaxf(fmaf(v[j + 1], Pa2[c0 + j + 1], Pb2[c0 + j + 1]), 0.f);
                r.z = fmaxf(fmaf(v[j + 2], Pa2[c0 + j + 2], Pb2[c0 + j + 2]), 0.f);
                r.w = fmaxf(fmaf(v[j + 3], Pa2[c0 + j + 3], Pb2[c0 + j + 3]), 0.f);
                float4 hi, lo;
                split_tf32(r, hi, lo);
                Xhi[((c0 + j) / 4) * kTcRows + tid] = hi;
                Xlo[((c0 + j) / 4) * kTcRows + tid] = lo;
            }
        }
        tc_fence_before_sync();   // order the TMEM reads before the barrier inside the next gemm()
    }

    // ------------------------------------------------------------------ scores S = X Ws^T, softmax over K, weighted sum
    gemm(a.w_score, D, D, 0);
    const int pl = tid / K;       // point of this row inside the CTA
#pragma unroll 1
    for (int c0 = 0; c0 < D; c0 += 32) {
        float sc[32];
        tmem_ld32(tmem + lane_base + (uint32_t)c0, sc);
#pragma unroll
        for (int j4 = 0; j4 < 32; j4 += 4) {
            const float4 xh = Xhi[((c0 + j4) / 4) * kTcRows + tid];
            const float4 xl = Xlo[((c0 + j4) / 4) * kTcRows + tid];
            const float xv[4] = {xh.x + xl.x, xh.y + xl.y, xh.z + xl.z, xh.w + xl.w};
            float res[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                float m = sc[j4 + u];
#pragma unroll
                for (int o = 1; o < K; o <<= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
                const float e = __expf(sc[j4 + u] - m);
                float se = e, sx = e * xv[u];
#pragma unroll
                for (int o = 1; o < K; o <<= 1) {
                    se += __shfl_xor_sync(0xffffffffu, se, o);
                    sx += __shfl_xor_sync(0xffffffffu, sx, o);
                }
                res[u] = sx / se;
            }
            if (tid % K == 0) *reinterpret_cast<float4*>(outs + pl * D + c0 + j4) = make_float4(res[0], res[1], res[2], res[3]);
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    for (int i = tid; i < PTS * D / 4; i += kTcRows) {
        const int p = i / (D / 4), c4 = i % (D / 4);
        if (p0 + p < a.N)
            reinterpret_cast<float4*>(a.pooled + ((size_t)b * a.N + p0 + p) * D)[c4] = reinterpret_cast<const float4*>(outs)[i];
    }
    if (warp == 0) tmem_dealloc_warp(tmem, TMEM_COLS);
}

template <int D, int K, int STAGE>
static int launch_tc(const LfaTcArgs& a, cudaStream_t st) {
    auto kern = lfa_pool_tc_kernel<D, K, STAGE>;
    constexpr size_t smem = LfaTcSmem<D, K>::BYTES;
    static_assert(smem <= 232448, "tile does not fit shared memory");
    R3D_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid(ceil_div(a.N, LfaTcSmem<D, K>::PTS), a.B);
    kern<<<grid, kTcRows, smem, st>>>(a);
    R3D_LAUNCH_CHECK("lfa_pool_tc_kernel");
    return R3D_OK;
}

}  // namespace r3d

using namespace r3d;

extern "C" int r3d_lfa_pool_tc(int stage, const float* xyz, long long xyz_bstride, const int32_t* idx, const float* feat,
                               long long feat_bstride, const float* w_rpe1, const float* a_rpe1, const float* b_rpe1,
                               const float* w_rpe2, const float* a_rpe2, const float* b_rpe2, const float* w_score,
                               float* pooled, int B, int N, int K, int d, r3d_stream_t stream) {
    if (stage != 1 && stage != 2) return R3D_EINVAL;
    if (B < 0 || N < 0 || K <= 0 || d <= 0) return R3D_EINVAL;
    if (B == 0 || N == 0) return R3D_OK;
    if (!xyz || !idx || !feat || !w_rpe1 || !a_rpe1 || !b_rpe1 || !w_score || !pooled) return R3D_EINVAL;
    if (stage == 2 && (!w_rpe2 || !a_rpe2 || !b_rpe2)) return R3D_EINVAL;
    const int h = d / 2;
    if (xyz_bstride == 0) xyz_bstride = (long long)N * 3;
    if (feat_bstride == 0) feat_bstride = (long long)N * h;
    if (!is_aligned(feat, 16) || !is_aligned(pooled, 16) || !is_aligned(w_score, 16) ||
        (w_rpe2 && !is_aligned(w_rpe2, 16)) || (feat_bstride % 4) != 0)
        return R3D_EALIGN;
    LfaTcArgs a{xyz, xyz_bstride, idx, feat, feat_bstride, w_rpe1, a_rpe1, b_rpe1, w_rpe2, a_rpe2, b_rpe2, w_score,
                pooled, B, N};
    cudaStream_t st = static_cast<cudaStream_t>(stream);
#define R3D_TC_CASE(DD, KK)                                                                     \
    if (d == DD && K == KK) return stage == 1 ? launch_tc<DD, KK, 1>(a, st) : launch_tc<DD, KK, 2>(a, st);
    R3D_TC_CASE(64, 16) R3D_TC_CASE(128, 16) R3D_TC_CASE(64, 32) R3D_TC_CASE(128, 32)
#undef R3D_TC_CASE
    return R3D_EUNSUPPORTED;
}
