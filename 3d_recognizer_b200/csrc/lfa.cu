// Fused LocalSpatialEncoding + AttentivePooling for one half of a LocalFeatureAggregation block (sm_100a).
//
// Reference op sequence replaced (randlanet/utils/modules.py):
//   :170-186  RelativePositionEncoding   rpe = [p_i, p_j, p_i - p_j, |p_i - p_j|]      (B,10,N,K)
//   :317      mlp_rpe1                   r1 = relu(bn(W1 rpe + b1))                      (B,h,N,K)
//   :321      mlp_rpe2 (STAGE 2 only)    r2 = relu(bn(W2 r1 + b2))   -- fed by r1, not by rpe
//   :209-221  PointFeatureAugmentation   x  = [r ; F[idx]]  (F = mlp1 output, or pool1 output in stage 2)
//   :246-252  AttentivePooling           s = Ws x ; a = softmax_K(s) ; pooled = sum_K a * x (B,d,N,1)
// The pooling MLP (:253) is a per-point layer and runs in pointwise.cu.
//
// None of the (B,C,N,K) intermediates of the reference touches HBM: a CTA owns PTS points, builds their
// K x d neighbourhood matrices in shared memory (channel-major, [c][point*K + k]), runs the d x d score
// GEMM (and the h x h rpe GEMM of stage 2) on the FP32 pipe with a 16 x 8 register tile per thread, and
// finishes softmax + weighted sum in registers.  HBM traffic per point is the compulsory
// 12 B xyz + 4K B indices + K*4h B gathered features (L2-resident) + 4d B output.
//
// BatchNorm arrives as per-channel (scale, shift): eval mode folds running statistics
// (eps 1e-6, modules.py:87), train mode passes the batch statistics computed by lfa_rpe_stats_kernel.
//
// Thread mapping (128 threads, CG = d/8 column groups, RH = K/16 row halves):
//   thread -> (point p, row-half rh, column group g); score tile = rows [p*K + rh*16, +16) x
//   cols {g*4..g*4+3} U {d/2 + g*4..+3}.  A operand (neighbourhood matrix) is read with LDS.128
//   (broadcast inside a point), B operand (weights, streamed through a double-buffered ring by 1-D TMA
//   bulk copies) with conflict-free LDS.128.
#include "lfa_common.cuh"

namespace r3d {

constexpr int kLfaThreads = 128;
// 8 KB weight-ring stages keep the forward tile under 113 KB: two CTAs per SM, so one CTA's gather prologue
// overlaps the other's GEMM (ncu of the 16 KB version: 1 CTA/SM, issue slots 41 % busy)
constexpr int kLfaFwdStage = 2048;
// narrow layers (d = 16: a 16x16 score matrix) get 4-row register tiles: 8 threads per point instead of 2, so a
// 128-thread CTA needs a 4x smaller shared-memory tile and 4-8 CTAs fit an SM (ncu: 2 warps/SM with 16-row tiles)

struct LfaArgs {
    const float* xyz;        // (B,N,3)
    long long xyz_bstride;
    const int32_t* idx;      // (B,N,K)
    const float* feat;       // (B,N,h) neighbour feature source, dense rows of h floats
    long long feat_bstride;
    const float* w_rpe1;     // (h,10)  [out][in]
    const float* a_rpe1;     // (h) scale
    const float* b_rpe1;     // (h) shift
    const float* w_rpe2T;    // (h,h)   [in][out]   (stage 2)
    const float* a_rpe2;
    const float* b_rpe2;
    const float* w_scoreT;   // (d,d)   [in][out]
    float* pooled;           // (B,N,d)
    int B, N;
};

template <int D, int K>
struct LfaFwdSmem {
    using C = LfaCfg<D, K, kLfaThreads, kLfaFwdStage, lfa_rows_per_thread(D)>;
    static constexpr int P_FLOATS = C::H * 12 + 4 * C::H;     // w_rpe1 padded to 12 per channel + a1,b1,a2,b2
    static constexpr size_t BYTES = (size_t)(C::X_FLOATS + 2 * C::WSTAGE + P_FLOATS) * sizeof(float) + 16;
};

template <int D, int K, int STAGE>
__global__ void __launch_bounds__(kLfaThreads, (D <= 16 ? 5 : (D <= 64 ? 3 : 2))) lfa_pool_kernel(LfaArgs a) {
    using C = LfaCfg<D, K, kLfaThreads, kLfaFwdStage, lfa_rows_per_thread(D)>;
    constexpr int H = C::H;
    constexpr int RT = C::RT;
    extern __shared__ __align__(128) float smem[];
    float* X = smem;                               // [D][ROWS_PAD]
    float* ring = X + C::X_FLOATS;                 // [2][WSTAGE]
    float* Pw1 = ring + 2 * C::WSTAGE;             // [H][12]
    float* Pa1 = Pw1 + H * 12;
    float* Pb1 = Pa1 + H;
    float* Pa2 = Pb1 + H;
    float* Pb2 = Pa2 + H;
    uint64_t* bars = reinterpret_cast<uint64_t*>(Pb2 + H);

    const int tid = threadIdx.x;
    const int b = blockIdx.y;
    const int p0 = blockIdx.x * C::PTS;

    if (tid == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        mbar_fence_init();
    }
    // parameters -> smem
    load_rpe1_params<H, kLfaThreads>(Pw1, Pa1, Pb1, a.w_rpe1, a.a_rpe1, a.b_rpe1, tid);
    if (STAGE == 2) {
        for (int i = tid; i < H; i += kLfaThreads) {
            Pa2[i] = a.a_rpe2[i];
            Pb2[i] = a.b_rpe2[i];
        }
    }
    __syncthreads();

    // ------------------------------------------------------------------ prologue: build X = [r1 ; F[idx]]
    // work item = (row, channel part); rows are (point, neighbour) pairs of this CTA
    constexpr int NSPLIT = (C::ROWS >= kLfaThreads) ? 1 : kLfaThreads / C::ROWS;
    constexpr int CH_PER = H / NSPLIT;
    static_assert(H % NSPLIT == 0 && CH_PER % 4 == 0 || NSPLIT == 1, "channel split");
    const float* xyz_b = a.xyz + (size_t)b * a.xyz_bstride;
    const float* feat_b = a.feat + (size_t)b * a.feat_bstride;
    for (int item = tid; item < C::ROWS * NSPLIT; item += kLfaThreads) {
        const int row = item % C::ROWS, part = item / C::ROWS;
        const int p = row / K, k = row % K;
        const int pi = min(p0 + p, a.N - 1);
        const int pj = a.idx[((size_t)b * a.N + pi) * K + k];
        float rpe[10];
        rpe_of_row(xyz_b, pi, pj, rpe);
        float* xcol = X + p * C::PSTRIDE + k;
        const int c_lo = part * CH_PER, c_hi = c_lo + CH_PER;
        for (int ch = c_lo; ch < c_hi; ++ch) xcol[(size_t)ch * C::ROWS_PAD] = rpe_mlp1(Pw1, Pa1, Pb1, ch, rpe);
        const float* frow = feat_b + (size_t)pj * H;
        for (int c = c_lo; c < c_hi; c += 4) {
            const float4 t = *reinterpret_cast<const float4*>(frow + c);
            xcol[(size_t)(H + c + 0) * C::ROWS_PAD] = t.x;
            xcol[(size_t)(H + c + 1) * C::ROWS_PAD] = t.y;
            xcol[(size_t)(H + c + 2) * C::ROWS_PAD] = t.z;
            xcol[(size_t)(H + c + 3) * C::ROWS_PAD] = t.w;
        }
    }
    __syncthreads();

    // ------------------------------------------------------------------ thread tile coordinates
    const int rh = tid % C::RH;
    const int g = (tid / C::RH) % C::CG;
    const int p = tid / C::TPP;
    const int row0 = p * C::PSTRIDE + rh * RT;
    WPipe pipe{ring, bars, 0u, C::WSTAGE};

    // ------------------------------------------------------------------ stage 2: r2 = relu(a2 * (W2 r1) + b2), in place
    if (STAGE == 2) {
        float acc2[RT][4];
#pragma unroll
        for (int r = 0; r < RT; ++r)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc2[r][j] = 0.f;
        gemm_stream<1, kLfaThreads, RT>(acc2, X, C::ROWS_PAD, row0, H, a.w_rpe2T, H, 0, g, pipe, tid);
        // every thread is past the barrier that ends gemm_stream: r1 may be overwritten
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int col = g * 4 + j;
            const float sa = Pa2[col], sb = Pb2[col];
            float* dst = X + (size_t)col * C::ROWS_PAD + row0;
#pragma unroll
            for (int v = 0; v < RT / 4; ++v) {
                float4 t;
                t.x = fmaxf(fmaf(acc2[4 * v + 0][j], sa, sb), 0.f);
                t.y = fmaxf(fmaf(acc2[4 * v + 1][j], sa, sb), 0.f);
                t.z = fmaxf(fmaf(acc2[4 * v + 2][j], sa, sb), 0.f);
                t.w = fmaxf(fmaf(acc2[4 * v + 3][j], sa, sb), 0.f);
                *reinterpret_cast<float4*>(dst + 4 * v) = t;
            }
        }
        __syncthreads();
    }

    // ------------------------------------------------------------------ score GEMM  S = X^T Ws^T
    float acc[RT][8];
#pragma unroll
    for (int r = 0; r < RT; ++r)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[r][j] = 0.f;
    gemm_stream<2, kLfaThreads, RT>(acc, X, C::ROWS_PAD, row0, D, a.w_scoreT, D, D / 2, g, pipe, tid);

    // ------------------------------------------------------------------ softmax over K + weighted sum
    float outv[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int col = (j < 4) ? (g * 4 + j) : (D / 2 + g * 4 + (j - 4));
        float m = acc[0][j];
#pragma unroll
        for (int r = 1; r < RT; ++r) m = fmaxf(m, acc[r][j]);
#pragma unroll
        for (int o = 1; o < C::RH; o <<= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        const float* xc = X + (size_t)col * C::ROWS_PAD + row0;
        float se = 0.f, sx = 0.f;
#pragma unroll
        for (int v = 0; v < RT / 4; ++v) {
            const float4 t = *reinterpret_cast<const float4*>(xc + 4 * v);
            const float e0 = __expf(acc[4 * v + 0][j] - m), e1 = __expf(acc[4 * v + 1][j] - m);
            const float e2 = __expf(acc[4 * v + 2][j] - m), e3 = __expf(acc[4 * v + 3][j] - m);
            se += (e0 + e1) + (e2 + e3);
            sx = fmaf(e0, t.x, sx); sx = fmaf(e1, t.y, sx); sx = fmaf(e2, t.z, sx); sx = fmaf(e3, t.w, sx);
        }
#pragma unroll
        for (int o = 1; o < C::RH; o <<= 1) {
            se += __shfl_xor_sync(0xffffffffu, se, o);
            sx += __shfl_xor_sync(0xffffffffu, sx, o);
        }
        outv[j] = sx / se;
    }
    if (rh == 0 && p0 + p < a.N) {
        float* out = a.pooled + ((size_t)b * a.N + p0 + p) * D;
        *reinterpret_cast<float4*>(out + g * 4) = make_float4(outv[0], outv[1], outv[2], outv[3]);
        *reinterpret_cast<float4*>(out + D / 2 + g * 4) = make_float4(outv[4], outv[5], outv[6], outv[7]);
    }
}

template <int D, int K, int STAGE>
static int launch_lfa(const LfaArgs& a, cudaStream_t st) {
    using C = LfaCfg<D, K, kLfaThreads, kLfaFwdStage, lfa_rows_per_thread(D)>;
    auto kern = lfa_pool_kernel<D, K, STAGE>;
    constexpr size_t smem = LfaFwdSmem<D, K>::BYTES;
    R3D_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid(ceil_div(a.N, C::PTS), a.B);
    kern<<<grid, kLfaThreads, smem, st>>>(a);
    R3D_LAUNCH_CHECK("lfa_pool_kernel");
    return R3D_OK;
}

template <int STAGE>
static int dispatch_lfa(int d, int K, const LfaArgs& a, cudaStream_t st) {
#define R3D_LFA_CASE(DD, KK) \
    if (d == DD && K == KK) return launch_lfa<DD, KK, STAGE>(a, st);
    R3D_LFA_CASE(16, 16) R3D_LFA_CASE(32, 16) R3D_LFA_CASE(64, 16) R3D_LFA_CASE(128, 16) R3D_LFA_CASE(256, 16)
    R3D_LFA_CASE(16, 32) R3D_LFA_CASE(32, 32) R3D_LFA_CASE(64, 32) R3D_LFA_CASE(128, 32) R3D_LFA_CASE(256, 32)
#undef R3D_LFA_CASE
    return R3D_EUNSUPPORTED;
}

}  // namespace r3d

using namespace r3d;

extern "C" int r3d_lfa_pool(int stage, const float* xyz, long long xyz_bstride, const int32_t* idx, const float* feat,
                            long long feat_bstride, const float* w_rpe1, const float* a_rpe1, const float* b_rpe1,
                            const float* w_rpe2T, const float* a_rpe2, const float* b_rpe2, const float* w_scoreT,
                            float* pooled, int B, int N, int K, int d, r3d_stream_t stream) {
    if (stage != 1 && stage != 2) return R3D_EINVAL;
    if (B < 0 || N < 0 || K <= 0 || d <= 0) return R3D_EINVAL;
    if (B == 0 || N == 0) return R3D_OK;
    if (!xyz || !idx || !feat || !w_rpe1 || !a_rpe1 || !b_rpe1 || !w_scoreT || !pooled) return R3D_EINVAL;
    if (stage == 2 && (!w_rpe2T || !a_rpe2 || !b_rpe2)) return R3D_EINVAL;
    const int h = d / 2;
    if (xyz_bstride == 0) xyz_bstride = (long long)N * 3;
    if (feat_bstride == 0) feat_bstride = (long long)N * h;
    if (!is_aligned(feat, 16) || !is_aligned(pooled, 16) || !is_aligned(w_scoreT, 16) ||
        (w_rpe2T && !is_aligned(w_rpe2T, 16)) || (feat_bstride % 4) != 0)
        return R3D_EALIGN;
    LfaArgs a{xyz, xyz_bstride, idx, feat, feat_bstride, w_rpe1, a_rpe1, b_rpe1, w_rpe2T, a_rpe2, b_rpe2,
              w_scoreT, pooled, B, N};
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    return stage == 1 ? dispatch_lfa<1>(d, K, a, st) : dispatch_lfa<2>(d, K, a, st);
}
