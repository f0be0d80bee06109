// Training-mode companions of the fused LocSE + attentive-pooling kernel (lfa.cu), sm_100a:
//
//   lfa_pool_bwd_kernel   backward of one r3d_lfa_pool launch.  Mirrors the forward: the CTA rebuilds its
//                         neighbourhood matrix X = [r ; F[idx]] and the scores S = X Ws^T in shared memory /
//                         registers, then
//                             A   = softmax_K(S),  pooled = sum_K A*X
//                             dS  = A * g * (X - pooled)            (g = d loss / d pooled)
//                             dX  = g * A + dS Ws                   (second streamed GEMM)
//                             dWs += dS^T X                         (row reduction, reduce_gemm)
//                             dF[idx] += dX[:, h:]                  (vector atomics, red.global.add.v4.f32)
//                             du  = dX[:, :h] * [r > 0]             -> G1 += du1^T [rpe, 1]      (stage 1)
//                         and, in stage 2, through mlp_rpe2 first:  G2 += du2^T [r1 | rpe, 1],
//                             dr1 = du2 (a2 . W2),  du1 = dr1 * [r1 > 0].
//                         The host turns G1/G2 into weight, scale and shift gradients (engine.py).
//   lfa_moments_kernel    first and second moments of the position encoding (16x16, fp64) and of
//                         r1 = relu(a1 (W1 rpe) + b1) (h x h, fp64) over all (point, neighbour) rows: the
//                         train-mode BatchNorm statistics of mlp_rpe1 / mlp_rpe2 (modules.py:86-90) follow
//                         from them in closed form (mean = W mu, var = diag(W Cov W^T)), which keeps every
//                         kernel single-pass and lets autograd differentiate the statistics.
//   lfa_moments_bwd_kernel  backward of the r1 moments: dr1 = g_sum + (G + G^T) r1, du1 = dr1 * [r1 > 0],
//                         G1 += du1^T [rpe, 1].
//
// Reference: autograd of randlanet/utils/modules.py:298-325 driven by trainer.py:115-119.
#include "lfa_common.cuh"

namespace r3d {

struct LfaBwdArgs {
    const float* xyz;
    long long xyz_bstride;
    const int32_t* idx;
    const float* feat;
    long long feat_bstride;
    const float* w_rpe1;      // (h,10)
    const float* a_rpe1;
    const float* b_rpe1;
    const float* w_rpe2T;     // (h,h) [in][out]           (stage 2)
    const float* a_rpe2;
    const float* b_rpe2;
    const float* w_rpe2s;     // (h,h) [out][in] * a2[out] (stage 2; backward GEMM operand)
    const float* w_scoreT;    // (d,d) [in][out]
    const float* w_score;     // (d,d) [out][in]
    const float* dpooled;     // (B,N,d)
    float* dfeat;             // (B,N,h)   += (pre-zeroed by the caller)
    long long dfeat_bstride;
    float* dw_score;          // (d,d) [out][in] +=
    double* g1;               // (h,16): [:, :10] += du1^T rpe, [:,10] += sum du1   (fp64: these sums cancel against the
    double* g2m;              // (h,h)  += du2^T r1        (stage 2)               BatchNorm mean/variance terms)
    double* g2c;              // (h,16): [:,10] += sum du2 (stage 2)
    // train-mode stage 2 (batch-statistics BatchNorm behind mlp_rpe2): instead of going on to r1, the kernel
    // hands du2 to the two-pass BatchNorm backward (lfa_moments_kernel MODE 3)
    float* du2_tiles;         // nullable; [b][tile][h][PTS*K] du2 of every tile
    double* sum_du2;          // (2,h): += sum du2, += sum du2 * r2
    int B, N;
};

// FINE = 1: half the rows per thread, i.e. twice the threads per point and half the points per CTA.  For the wide
// levels of a SMALL batch (8 clouds x 39..156 points) the default tiles give 80..160 CTAs of 4 warps -- one CTA per SM,
// one or two waves, every warp carrying a 16-row register tile; the fine tiles double the CTAs, halve each warp's
// work and fit two CTAs per SM.
template <int D, int K, int NT, int FINE = 0>
using LfaBwdCfg = LfaCfg<D, K, NT, (D <= 64 ? 2048 : 4096), (FINE ? lfa_rows_per_thread(D) / 2 : lfa_rows_per_thread(D))>;

template <int D, int K, int NT, int FINE = 0>
struct LfaBwdSmem {
    using C = LfaBwdCfg<D, K, NT, FINE>;
    static constexpr int FLOATS = 2 * C::X_FLOATS + kRpeRows * C::ROWS_PAD + 2 * C::WSTAGE + C::H * 16 + C::ROWS +
                                  C::PTS * D;
    static constexpr size_t BYTES = (size_t)FLOATS * sizeof(float) + 16;
};

template <int D, int K, int NT, int STAGE, int FINE = 0>
__global__ void __launch_bounds__(NT, (D <= 16 ? 3 : (D <= 64 ? 2 : (FINE ? 2 : 1)))) lfa_pool_bwd_kernel(LfaBwdArgs a) {
    using C = LfaBwdCfg<D, K, NT, FINE>;
    constexpr int H = C::H;
    constexpr int RT = C::RT;
    constexpr int RP = C::ROWS_PAD;
    extern __shared__ __align__(128) float smem[];
    float* X = smem;                                  // [D][RP]   neighbourhood matrix, later du
    float* G = X + C::X_FLOATS;                       // [D][RP]   dS, later r1 / du1 (stage 2)
    float* RPE = G + C::X_FLOATS;                     // [16][RP]  rpe rows, ones row, zero rows
    float* ring = RPE + kRpeRows * RP;                // [2][WSTAGE]
    float* Pw1 = ring + 2 * C::WSTAGE;                // [H][12]
    float* Pa1 = Pw1 + H * 12;
    float* Pb1 = Pa1 + H;
    float* Pa2 = Pb1 + H;
    float* Pb2 = Pa2 + H;
    int* idxs = reinterpret_cast<int*>(Pb2 + H);      // [ROWS] neighbour index per row (-1: padding point)
    float* gp = reinterpret_cast<float*>(idxs + C::ROWS);  // [PTS][D] dpooled tile
    uint64_t* bars = reinterpret_cast<uint64_t*>(gp + C::PTS * D);

    const int tid = threadIdx.x;
    const int b = blockIdx.y;
    const int p0 = blockIdx.x * C::PTS;

    if (tid == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        mbar_fence_init();
    }
    load_rpe1_params<H, NT>(Pw1, Pa1, Pb1, a.w_rpe1, a.a_rpe1, a.b_rpe1, tid);
    if (STAGE == 2) {
        for (int i = tid; i < H; i += NT) {
            Pa2[i] = a.a_rpe2[i];
            Pb2[i] = a.b_rpe2[i];
        }
    }
    for (int i = tid; i < C::PTS * D; i += NT) {
        const int p = i / D, c = i % D;
        gp[i] = (p0 + p < a.N) ? a.dpooled[((size_t)b * a.N + p0 + p) * D + c] : 0.f;
    }
    __syncthreads();

    // ------------------------------------------------------------------ rebuild X = [r1 ; F[idx]], RPE
    constexpr int NSPLIT = (C::ROWS >= NT) ? 1 : NT / C::ROWS;
    constexpr int CH_PER = H / NSPLIT;
    static_assert(NSPLIT == 1 || (H % NSPLIT == 0 && CH_PER % 4 == 0), "channel split");
    const float* xyz_b = a.xyz + (size_t)b * a.xyz_bstride;
    const float* feat_b = a.feat + (size_t)b * a.feat_bstride;
    for (int item = tid; item < C::ROWS * NSPLIT; item += NT) {
        const int row = item % C::ROWS, part = item / C::ROWS;
        const int p = row / K, k = row % K;
        const bool valid = p0 + p < a.N;
        const int pi = min(p0 + p, a.N - 1);
        const int pj = a.idx[((size_t)b * a.N + pi) * K + k];
        float rpe[10];
        rpe_of_row(xyz_b, pi, pj, rpe);
        const int off = p * C::PSTRIDE + k;
        if (part == 0) {
#pragma unroll
            for (int m = 0; m < 10; ++m) RPE[m * RP + off] = valid ? rpe[m] : 0.f;
            RPE[10 * RP + off] = valid ? 1.f : 0.f;
#pragma unroll
            for (int m = 11; m < kRpeRows; ++m) RPE[m * RP + off] = 0.f;
            idxs[row] = valid ? pj : -1;
        }
        float* xcol = X + off;
        const int c_lo = part * CH_PER, c_hi = c_lo + CH_PER;
        for (int ch = c_lo; ch < c_hi; ++ch) xcol[(size_t)ch * RP] = rpe_mlp1(Pw1, Pa1, Pb1, ch, rpe);
        const float* frow = feat_b + (size_t)pj * H;
        for (int c = c_lo; c < c_hi; c += 4) {
            const float4 t = *reinterpret_cast<const float4*>(frow + c);
            xcol[(size_t)(H + c + 0) * RP] = t.x;
            xcol[(size_t)(H + c + 1) * RP] = t.y;
            xcol[(size_t)(H + c + 2) * RP] = t.z;
            xcol[(size_t)(H + c + 3) * RP] = t.w;
        }
    }
    __syncthreads();

    const int rh = tid % C::RH;
    const int g = (tid / C::RH) % C::CG;
    const int p = tid / C::TPP;
    const int row0 = p * C::PSTRIDE + rh * RT;
    WPipe pipe{ring, bars, 0u, C::WSTAGE};

    // ------------------------------------------------------------------ stage 2: r2 = relu(a2 (W2 r1) + b2) in place
    if (STAGE == 2) {
        float acc2[RT][4];
#pragma unroll
        for (int r = 0; r < RT; ++r)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc2[r][j] = 0.f;
        gemm_stream<1, NT, RT>(acc2, X, RP, row0, H, a.w_rpe2T, H, 0, g, pipe, tid);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int col = g * 4 + j;
            const float sa = Pa2[col], sb = Pb2[col];
            float* dst = X + (size_t)col * RP + row0;
#pragma unroll
            for (int r = 0; r < RT; ++r) dst[r] = fmaxf(fmaf(acc2[r][j], sa, sb), 0.f);
        }
        __syncthreads();
    }

    // ------------------------------------------------------------------ scores, softmax, dS and the direct part of dX
    float acc[RT][8];
#pragma unroll
    for (int r = 0; r < RT; ++r)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[r][j] = 0.f;
    gemm_stream<2, NT, RT>(acc, X, RP, row0, D, a.w_scoreT, D, D / 2, g, pipe, tid);

#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int col = (j < 4) ? (g * 4 + j) : (D / 2 + g * 4 + (j - 4));
        float m = acc[0][j];
#pragma unroll
        for (int r = 1; r < RT; ++r) m = fmaxf(m, acc[r][j]);
#pragma unroll
        for (int o = 1; o < C::RH; o <<= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        const float* xc = X + (size_t)col * RP + row0;
        float xv[RT];
        float se = 0.f, sx = 0.f;
#pragma unroll
        for (int r = 0; r < RT; ++r) {
            xv[r] = xc[r];
            const float e = __expf(acc[r][j] - m);
            acc[r][j] = e;
            se += e;
            sx = fmaf(e, xv[r], sx);
        }
#pragma unroll
        for (int o = 1; o < C::RH; o <<= 1) {
            se += __shfl_xor_sync(0xffffffffu, se, o);
            sx += __shfl_xor_sync(0xffffffffu, sx, o);
        }
        const float inv = 1.f / se;
        const float pooled = sx * inv;
        const float gv = gp[p * D + col];
        float* gc = G + (size_t)col * RP + row0;
#pragma unroll
        for (int r = 0; r < RT; ++r) {
            const float ga = gv * (acc[r][j] * inv);      // g * A  = direct part of dX
            gc[r] = ga * (xv[r] - pooled);                // dS
            acc[r][j] = ga;
        }
    }
    __syncthreads();

    // ------------------------------------------------------------------ dX = g*A + dS Ws ;  dWs += dS^T X
    gemm_stream<2, NT, RT>(acc, G, RP, row0, D, a.w_score, D, D / 2, g, pipe, tid);
    reduce_gemm<D, D, NT, C::PTS, K, C::PSTRIDE, float>(G, X, RP, a.dw_score, D, D, tid);
    __syncthreads();

    // ------------------------------------------------------------------ feature half: scatter-add to dF[idx]
    {
        float* df_b = a.dfeat + (size_t)b * a.dfeat_bstride;
#pragma unroll
        for (int r = 0; r < RT; ++r) {
            const int pj = idxs[p * K + rh * RT + r];
            if (pj >= 0)
                red_add_v4(df_b + (size_t)pj * H + g * 4, make_float4(acc[r][4], acc[r][5], acc[r][6], acc[r][7]));
        }
    }
    // ------------------------------------------------------------------ encoding half: du = dX * [r > 0], in place over r
    float pr2[4] = {0.f, 0.f, 0.f, 0.f}, ps2[4] = {0.f, 0.f, 0.f, 0.f};   // per-thread sum du2 * r2, sum du2
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        float* xc = X + (size_t)(g * 4 + j) * RP + row0;
#pragma unroll
        for (int r = 0; r < RT; ++r) {
            const float rv = xc[r];
            const float du = (rv > 0.f) ? acc[r][j] : 0.f;
            pr2[j] = fmaf(du, rv, pr2[j]);
            ps2[j] += du;
            xc[r] = du;
        }
    }
    __syncthreads();

    if (STAGE == 2 && a.du2_tiles) {
        // ---- train mode: du2 tile -> global (coalesced, channel-major as in shared memory), BatchNorm sums
        float* red = gp;                                   // [2][H] scratch (the dpooled tile is no longer needed)
        for (int i = tid; i < 2 * H; i += NT) red[i] = 0.f;
        __syncthreads();
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            atomicAdd(&red[g * 4 + j], ps2[j]);
            atomicAdd(&red[H + g * 4 + j], pr2[j]);
        }
        float* tile = a.du2_tiles + ((size_t)b * gridDim.x + blockIdx.x) * H * C::ROWS;
        for (int i = tid; i < H * C::ROWS / 4; i += NT) {
            const int c = i / (C::ROWS / 4), q = i % (C::ROWS / 4);
            const int p = (q * 4) / K, k = (q * 4) % K;
            reinterpret_cast<float4*>(tile)[i] = *reinterpret_cast<const float4*>(X + (size_t)c * RP + p * C::PSTRIDE + k);
        }
        __syncthreads();
        for (int i = tid; i < 2 * H; i += NT) atomicAdd(a.sum_du2 + i, (double)red[i]);
        return;
    }

    if (STAGE == 1) {
        reduce_gemm<H, kRpeRows, NT, C::PTS, K, C::PSTRIDE, double>(X, RPE, RP, a.g1, kRpeRows, 11, tid);
        return;
    }

    // ------------------------------------------------------------------ stage 2: through mlp_rpe2 back to r1
    // r1 again (from the buffered encoding) into G[:H]
    for (int item = tid; item < C::ROWS * NSPLIT; item += NT) {
        const int row = item % C::ROWS, part = item / C::ROWS;
        const int off = (row / K) * C::PSTRIDE + (row % K);
        float rpe[10];
#pragma unroll
        for (int m = 0; m < 10; ++m) rpe[m] = RPE[m * RP + off];
        const int c_lo = part * CH_PER, c_hi = c_lo + CH_PER;
        for (int ch = c_lo; ch < c_hi; ++ch) G[(size_t)ch * RP + off] = rpe_mlp1(Pw1, Pa1, Pb1, ch, rpe);
    }
    __syncthreads();
    reduce_gemm<H, H, NT, C::PTS, K, C::PSTRIDE, double>(X, G, RP, a.g2m, H, H, tid);
    reduce_gemm<H, kRpeRows, NT, C::PTS, K, C::PSTRIDE, double>(X, RPE, RP, a.g2c, kRpeRows, 11, tid);
    {
        float acc2[RT][4];
#pragma unroll
        for (int r = 0; r < RT; ++r)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc2[r][j] = 0.f;
        gemm_stream<1, NT, RT>(acc2, X, RP, row0, H, a.w_rpe2s, H, 0, g, pipe, tid);
        // every thread is past the barrier that ends gemm_stream, i.e. past its reads of r1 in reduce_gemm
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float* gc = G + (size_t)(g * 4 + j) * RP + row0;
#pragma unroll
            for (int r = 0; r < RT; ++r) gc[r] = (gc[r] > 0.f) ? acc2[r][j] : 0.f;
        }
    }
    __syncthreads();
    reduce_gemm<H, kRpeRows, NT, C::PTS, K, C::PSTRIDE, double>(G, RPE, RP, a.g1, kRpeRows, 11, tid);
}

// ---------------------------------------------------------------------------------------- moments
struct LfaMomArgs {
    const float* xyz;
    long long xyz_bstride;
    const int32_t* idx;
    const float* w_rpe1;
    const float* a_rpe1;
    const float* b_rpe1;
    double* m_rpe;      // (16,16): [j][c] += sum rpe_j rpe_c (row/col 10 = the ones channel)    MODE 0
    double* m_r1;       // (h,h)  += sum r1 r1^T                                                  MODE 1
    double* s_r1;       // (h,16): [:,10] += sum r1                                               MODE 1
    const float* gsym;  // (h,h) G + G^T, [c][j] (symmetric)                                      MODE 2 (backward)
    const float* gsum;  // (h)   d loss / d sum r1                                                MODE 2
    double* g1;         // (h,16) += du1^T [rpe, 1]                                               MODE 2, 3
    // MODE 3: two-pass BatchNorm backward of mlp_rpe2 (batch statistics)
    const float* du2_tiles;   // [b][tile][h][PTS*K] from lfa_pool_bwd_kernel
    const float* w_rpe2T;     // (h,h) [in][out]
    const float* w_rpe2;      // (h,h) [out][in]
    const float* bn2;         // (5,h): a2, mean2 (of W2 r1), rstd2, m1 = mean(du2), m2 = mean(du2 * zhat2)
    double* dw2;              // (h,h) [out][in] += dz2^T r1
    int B, N;
};

template <int D, int K, int NT, int FINE = 0>
struct LfaMomSmem {
    using C = LfaBwdCfg<D, K, NT, FINE>;
    static constexpr int FLOATS = 2 * C::H * C::ROWS_PAD + kRpeRows * C::ROWS_PAD + 2 * C::WSTAGE + C::H * 14 + 4;
    static constexpr size_t BYTES = (size_t)FLOATS * sizeof(float) + 16;
};

// MODE 0: rpe moments.  MODE 1: r1 moments.  MODE 2: backward of the r1 moments.
// MODE 3: BatchNorm backward of mlp_rpe2 in the standard two-pass form: with the batch sums of pass 1 (m1, m2),
//         z2 = W2 r1, dz2 = a2 (du2 - m1 - zhat2 m2), dW2 += dz2^T r1, dr1 = dz2 W2, du1 = dr1 * [r1 > 0],
//         G1 += du1^T [rpe, 1].  Mean and variance terms are subtracted PER ROW before any row sum is taken: going
//         through the moments instead (MODE 2) subtracts two row sums that cancel to ~1/sqrt(rows) of their size and
//         lost 3-4 digits at 65 k rows.
template <int D, int K, int NT, int MODE, int FINE = 0>
__global__ void __launch_bounds__(NT, (D <= 16 ? 4 : (D <= 64 ? 3 : (FINE ? 2 : 1)))) lfa_moments_kernel(LfaMomArgs a) {
    using C = LfaBwdCfg<D, K, NT, FINE>;
    constexpr int H = C::H;
    constexpr int RT = C::RT;
    constexpr int RP = C::ROWS_PAD;
    extern __shared__ __align__(128) float smem[];
    float* R1 = smem;                        // [H][RP]
    float* DU = R1 + H * RP;                 // [H][RP]  (MODE 2)
    float* RPE = DU + H * RP;                // [16][RP]
    float* ring = RPE + kRpeRows * RP;       // [2][WSTAGE]
    float* Pw1 = ring + 2 * C::WSTAGE;       // [H][12]
    float* Pa1 = Pw1 + H * 12;
    float* Pb1 = Pa1 + H;
    uint64_t* bars = reinterpret_cast<uint64_t*>(Pb1 + H);

    const int tid = threadIdx.x;
    const int b = blockIdx.y;
    const int p0 = blockIdx.x * C::PTS;
    if (MODE >= 2 && tid == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        mbar_fence_init();
    }
    if (MODE != 0) load_rpe1_params<H, NT>(Pw1, Pa1, Pb1, a.w_rpe1, a.a_rpe1, a.b_rpe1, tid);
    __syncthreads();

    constexpr int NSPLIT = (C::ROWS >= NT) ? 1 : NT / C::ROWS;
    constexpr int CH_PER = H / NSPLIT;
    const float* xyz_b = a.xyz + (size_t)b * a.xyz_bstride;
    for (int item = tid; item < C::ROWS * NSPLIT; item += NT) {
        const int row = item % C::ROWS, part = item / C::ROWS;
        const int p = row / K, k = row % K;
        const bool valid = p0 + p < a.N;
        const int pi = min(p0 + p, a.N - 1);
        const int pj = a.idx[((size_t)b * a.N + pi) * K + k];
        float rpe[10];
        rpe_of_row(xyz_b, pi, pj, rpe);
        const int off = p * C::PSTRIDE + k;
        if (part == 0) {
#pragma unroll
            for (int m = 0; m < 10; ++m) RPE[m * RP + off] = valid ? rpe[m] : 0.f;
            RPE[10 * RP + off] = valid ? 1.f : 0.f;
#pragma unroll
            for (int m = 11; m < kRpeRows; ++m) RPE[m * RP + off] = 0.f;
        }
        if (MODE != 0) {
            const int c_lo = part * CH_PER, c_hi = c_lo + CH_PER;
            for (int ch = c_lo; ch < c_hi; ++ch)
                R1[(size_t)ch * RP + off] = valid ? rpe_mlp1(Pw1, Pa1, Pb1, ch, rpe) : 0.f;
        }
    }
    __syncthreads();

    if (MODE == 0) {
        reduce_gemm<kRpeRows, kRpeRows, NT, C::PTS, K, C::PSTRIDE, double>(RPE, RPE, RP, a.m_rpe, kRpeRows, 11, tid);
    } else if (MODE == 1) {
        reduce_gemm<H, H, NT, C::PTS, K, C::PSTRIDE, double>(R1, R1, RP, a.m_r1, H, H, tid);
        reduce_gemm<H, kRpeRows, NT, C::PTS, K, C::PSTRIDE, double>(R1, RPE, RP, a.s_r1, kRpeRows, 11, tid);
    } else if (MODE == 3) {
        const int rh = tid % C::RH;
        const int g = (tid / C::RH) % C::CG;
        const int p = tid / C::TPP;
        const int row0 = p * C::PSTRIDE + rh * RT;
        WPipe pipe{ring, bars, 0u, C::WSTAGE};
        // du2 tile -> DU (padding floats of the shared-memory rows are never read)
        const float* tile = a.du2_tiles + ((size_t)b * gridDim.x + blockIdx.x) * H * C::ROWS;
        for (int i = tid; i < H * C::ROWS / 4; i += NT) {
            const int c = i / (C::ROWS / 4), q = i % (C::ROWS / 4);
            const int pp = (q * 4) / K, k = (q * 4) % K;
            *reinterpret_cast<float4*>(DU + (size_t)c * RP + pp * C::PSTRIDE + k) = reinterpret_cast<const float4*>(tile)[i];
        }
        float acc2[RT][4];
#pragma unroll
        for (int r = 0; r < RT; ++r)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc2[r][j] = 0.f;
        gemm_stream<1, NT, RT>(acc2, R1, RP, row0, H, a.w_rpe2T, H, 0, g, pipe, tid);   // z2 = W2 r1 (barrier inside)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int col = g * 4 + j;
            const float a2 = a.bn2[col], mean2 = a.bn2[H + col], rstd2 = a.bn2[2 * H + col];
            const float m1 = a.bn2[3 * H + col], m2 = a.bn2[4 * H + col];
            float* dc = DU + (size_t)col * RP + row0;
            const float* valid = RPE + 10 * RP + row0;      // the ones channel: 0 on padding points
#pragma unroll
            for (int r = 0; r < RT; ++r) {
                const float zh = (acc2[r][j] - mean2) * rstd2;
                dc[r] = valid[r] * (a2 * (dc[r] - m1 - zh * m2));
            }
        }
        __syncthreads();
        reduce_gemm<H, H, NT, C::PTS, K, C::PSTRIDE, double>(DU, R1, RP, a.dw2, H, H, tid);
#pragma unroll
        for (int r = 0; r < RT; ++r)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc2[r][j] = 0.f;
        gemm_stream<1, NT, RT>(acc2, DU, RP, row0, H, a.w_rpe2, H, 0, g, pipe, tid);    // dr1 = dz2 W2
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float* rc = R1 + (size_t)(g * 4 + j) * RP + row0;
            float* dc = DU + (size_t)(g * 4 + j) * RP + row0;
#pragma unroll
            for (int r = 0; r < RT; ++r) dc[r] = (rc[r] > 0.f) ? acc2[r][j] : 0.f;
        }
        __syncthreads();
        reduce_gemm<H, kRpeRows, NT, C::PTS, K, C::PSTRIDE, double>(DU, RPE, RP, a.g1, kRpeRows, 11, tid);
    } else {
        const int rh = tid % C::RH;
        const int g = (tid / C::RH) % C::CG;
        const int p = tid / C::TPP;
        const int row0 = p * C::PSTRIDE + rh * RT;
        WPipe pipe{ring, bars, 0u, C::WSTAGE};
        float acc2[RT][4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float gs = a.gsum[g * 4 + j];
#pragma unroll
            for (int r = 0; r < RT; ++r) acc2[r][j] = gs;
        }
        gemm_stream<1, NT, RT>(acc2, R1, RP, row0, H, a.gsym, H, 0, g, pipe, tid);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float* rc = R1 + (size_t)(g * 4 + j) * RP + row0;
            float* dc = DU + (size_t)(g * 4 + j) * RP + row0;
#pragma unroll
            for (int r = 0; r < RT; ++r) dc[r] = (rc[r] > 0.f) ? acc2[r][j] : 0.f;   // padding rows: r1 == 0
        }
        __syncthreads();
        reduce_gemm<H, kRpeRows, NT, C::PTS, K, C::PSTRIDE, double>(DU, RPE, RP, a.g1, kRpeRows, 11, tid);
    }
}


// Moments of the position encoding (r3d_lfa_moments mode 0) as an HBM-streaming kernel: they depend on the coordinates
// and the neighbour lists only, any width.  A thread walks rows with the grid stride, accumulates the 66 distinct
// products of e = [rpe(10), 1] over at most 16 rows in fp32, then the warp reduces them in fp64 and adds them to the
// CTA's fp64 shared-memory slices; one global fp64 atomic per entry and CTA at the end.  (The tile kernel above rebuilt
// a channel-major shared-memory tile and ran a 16 x 16 reduce_gemm over it: 2.0 ms for 42 M rows; this: 0.9 ms.)
__global__ void __launch_bounds__(256) lfa_rpe_moments_kernel(const float* __restrict__ xyz, long long xyz_bstride,
                                                              const int32_t* __restrict__ idx, double* __restrict__ m_rpe,
                                                              int N, int K, long long rows) {
    // one fp64 slice per warp, entry i owned by lane i % 32 of that warp: plain read-modify-write, no atomics (fp64
    // shared-memory atomics are CAS loops; eight warps retrying on the same 66 addresses were most of this kernel's time)
    __shared__ double acc64[8][66];
    for (int i = threadIdx.x; i < 8 * 66; i += blockDim.x) (&acc64[0][0])[i] = 0.0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long stride = (long long)gridDim.x * blockDim.x;
    long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    // every thread of a warp runs the same number of batches (the warp reduction below is collective)
    const long long warp_first = row - lane;
    const bool small = rows <= 0x7fffffffLL;
    for (long long base = warp_first; base < rows; base += 16 * stride) {
        float p[66];
#pragma unroll
        for (int i = 0; i < 66; ++i) p[i] = 0.f;
        for (int it = 0; it < 16; ++it) {
            const long long r = base + lane + (long long)it * stride;
            if (r < rows) {
                // 64-bit divisions are ~100-instruction sequences, two of them per row cost more than the 66 FMAs
                int b, pi;
                if (small) {
                    const unsigned gp = (unsigned)r / (unsigned)K;
                    b = (int)(gp / (unsigned)N);
                    pi = (int)(gp - (unsigned)b * (unsigned)N);
                } else {
                    const long long gp = r / K;
                    b = (int)(gp / N);
                    pi = (int)(gp - (long long)b * N);
                }
                float e[11];
                float rpe[10];
                rpe_of_row(xyz + (size_t)b * xyz_bstride, pi, idx[r], rpe);
#pragma unroll
                for (int q = 0; q < 10; ++q) e[q] = rpe[q];
                e[10] = 1.f;
                int o = 0;
#pragma unroll
                for (int i = 0; i < 11; ++i)
#pragma unroll
                    for (int j = i; j < 11; ++j) p[o] = fmaf(e[i], e[j], p[o]), ++o;
            }
        }
#pragma unroll
        for (int i = 0; i < 66; ++i) {
            double v = (double)p[i];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == (i & 31)) acc64[warp][i] += v;
        }
    }
    __syncthreads();
    if (threadIdx.x < 66) {
        // entry t of the upper triangle -> (i, j), i <= j
        int t = threadIdx.x, i = 0;
        while (t >= 11 - i) {
            t -= 11 - i;
            ++i;
        }
        const int j = i + t;
        double v = 0.0;
#pragma unroll
        for (int w = 0; w < 8; ++w) v += acc64[w][threadIdx.x];
        atomicAdd(m_rpe + i * kRpeRows + j, v);
        if (i != j) atomicAdd(m_rpe + j * kRpeRows + i, v);
    }
}

// ------------------------------------------------------------------------------------- launchers
template <int D, int K, int NT, int STAGE, int FINE = 0>
static int launch_bwd(const LfaBwdArgs& a, cudaStream_t st) {
    using C = LfaBwdCfg<D, K, NT, FINE>;
    auto kern = lfa_pool_bwd_kernel<D, K, NT, STAGE, FINE>;
    constexpr size_t smem = LfaBwdSmem<D, K, NT, FINE>::BYTES;
    static_assert(smem <= 232448, "backward tile does not fit shared memory");
    R3D_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid(ceil_div(a.N, C::PTS), a.B);
    kern<<<grid, NT, smem, st>>>(a);
    R3D_LAUNCH_CHECK("lfa_pool_bwd_kernel");
    return R3D_OK;
}

// Fine tiles (see LfaBwdCfg) when the default grid would not even fill two waves of the wide-level kernels.
template <int D, int K>
static int tile_points() { return LfaBwdCfg<D, K, 128>::PTS; }
static int default_tile_points(int K, int d);

static bool lfa_fine_tiles(int d, int K, int B, int N) {
    // measured on the 8 x 2 500-point step: d = 128 (8 x 156 points) 130 -> 107 us; d = 256 (8 x 39) 117 -> 130 us (every
    // CTA streams the whole 3 x 256 KB of weights: halving the tile doubles that), so only d = 128 takes the fine tiles
    if (d != 128) return false;
    const int pts = default_tile_points(K, d);
    if (pts < 2) return false;
    return (long long)ceil_div(N, pts) * B <= 2LL * kNumSMs;
}

template <int STAGE>
static int dispatch_bwd(int d, int K, const LfaBwdArgs& a, cudaStream_t st) {
    if (lfa_fine_tiles(d, K, a.B, a.N)) {
#define R3D_FINE(DD, KK) \
    if (d == DD && K == KK) return launch_bwd<DD, KK, 128, STAGE, 1>(a, st);
        R3D_FINE(128, 16) R3D_FINE(128, 32)
#undef R3D_FINE
    }
#define R3D_CASE(DD, KK, NT) \
    if (d == DD && K == KK) return launch_bwd<DD, KK, NT, STAGE>(a, st);
    R3D_CASE(16, 16, 128) R3D_CASE(32, 16, 128) R3D_CASE(64, 16, 128) R3D_CASE(128, 16, 128) R3D_CASE(256, 16, 128)
    R3D_CASE(16, 32, 128) R3D_CASE(32, 32, 128) R3D_CASE(64, 32, 128) R3D_CASE(128, 32, 128) R3D_CASE(256, 32, 128)
#undef R3D_CASE
    return R3D_EUNSUPPORTED;
}

template <int D, int K, int NT, int MODE, int FINE = 0>
static int launch_mom(const LfaMomArgs& a, cudaStream_t st) {
    using C = LfaBwdCfg<D, K, NT, FINE>;
    auto kern = lfa_moments_kernel<D, K, NT, MODE, FINE>;
    constexpr size_t smem = LfaMomSmem<D, K, NT, FINE>::BYTES;
    static_assert(smem <= 232448, "moments tile does not fit shared memory");
    R3D_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid(ceil_div(a.N, C::PTS), a.B);
    kern<<<grid, NT, smem, st>>>(a);
    R3D_LAUNCH_CHECK("lfa_moments_kernel");
    return R3D_OK;
}

template <int MODE>
static int dispatch_mom(int d, int K, const LfaMomArgs& a, cudaStream_t st) {
    if (MODE == 3 && lfa_fine_tiles(d, K, a.B, a.N)) {      // pass 2 reads pass 1's du2 tiles: same tiling
#define R3D_FINE(DD, KK) \
    if (d == DD && K == KK) return launch_mom<DD, KK, 128, MODE, 1>(a, st);
        R3D_FINE(128, 16) R3D_FINE(128, 32)
#undef R3D_FINE
    }
#define R3D_CASE(DD, KK, NT) \
    if (d == DD && K == KK) return launch_mom<DD, KK, NT, MODE>(a, st);
    R3D_CASE(16, 16, 128) R3D_CASE(32, 16, 128) R3D_CASE(64, 16, 128) R3D_CASE(128, 16, 128) R3D_CASE(256, 16, 128)
    R3D_CASE(16, 32, 128) R3D_CASE(32, 32, 128) R3D_CASE(64, 32, 128) R3D_CASE(128, 32, 128) R3D_CASE(256, 32, 128)
#undef R3D_CASE
    return R3D_EUNSUPPORTED;
}

// points per CTA tile of the backward / moment kernels (layout of du2_tiles)
static int default_tile_points(int K, int d) {
#define R3D_CASE(DD, KK) \
    if (d == DD && K == KK) return tile_points<DD, KK>();
    R3D_CASE(16, 16) R3D_CASE(32, 16) R3D_CASE(64, 16) R3D_CASE(128, 16) R3D_CASE(256, 16)
    R3D_CASE(16, 32) R3D_CASE(32, 32) R3D_CASE(64, 32) R3D_CASE(128, 32) R3D_CASE(256, 32)
#undef R3D_CASE
    return R3D_EUNSUPPORTED;
}

}  // namespace r3d

using namespace r3d;

extern "C" int r3d_lfa_pool_bwd(int stage, const float* xyz, long long xyz_bstride, const int32_t* idx,
                                const float* feat, long long feat_bstride, const float* w_rpe1, const float* a_rpe1,
                                const float* b_rpe1, const float* w_rpe2T, const float* a_rpe2, const float* b_rpe2,
                                const float* w_rpe2s, const float* w_scoreT, const float* w_score,
                                const float* dpooled, float* dfeat, long long dfeat_bstride, float* dw_score,
                                double* g1, double* g2m, double* g2c, int B, int N, int K, int d, r3d_stream_t stream) {
    if (stage != 1 && stage != 2) return R3D_EINVAL;
    if (B < 0 || N < 0 || K <= 0 || d <= 0) return R3D_EINVAL;
    if (B == 0 || N == 0) return R3D_OK;
    if (!xyz || !idx || !feat || !w_rpe1 || !a_rpe1 || !b_rpe1 || !w_scoreT || !w_score || !dpooled || !dfeat ||
        !dw_score || !g1)
        return R3D_EINVAL;
    if (stage == 2 && (!w_rpe2T || !a_rpe2 || !b_rpe2 || !w_rpe2s || !g2m || !g2c)) return R3D_EINVAL;
    const int h = d / 2;
    if (xyz_bstride == 0) xyz_bstride = (long long)N * 3;
    if (feat_bstride == 0) feat_bstride = (long long)N * h;
    if (dfeat_bstride == 0) dfeat_bstride = (long long)N * h;
    if (!is_aligned(feat, 16) || !is_aligned(dfeat, 16) || !is_aligned(w_scoreT, 16) || !is_aligned(w_score, 16) ||
        (w_rpe2T && !is_aligned(w_rpe2T, 16)) || (w_rpe2s && !is_aligned(w_rpe2s, 16)) || (feat_bstride % 4) != 0 ||
        (dfeat_bstride % 4) != 0)
        return R3D_EALIGN;
    LfaBwdArgs a{xyz, xyz_bstride, idx, feat, feat_bstride, w_rpe1, a_rpe1, b_rpe1, w_rpe2T, a_rpe2, b_rpe2,
                 w_rpe2s, w_scoreT, w_score, dpooled, dfeat, dfeat_bstride, dw_score, g1, g2m, g2c, nullptr, nullptr,
                 B, N};
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    return stage == 1 ? dispatch_bwd<1>(d, K, a, st) : dispatch_bwd<2>(d, K, a, st);
}

extern "C" int r3d_lfa_tile_points(int K, int d) { return default_tile_points(K, d); }

extern "C" int r3d_lfa_tile_points_for(int K, int d, int B, int N) {
    const int pts = default_tile_points(K, d);
    if (pts <= 0) return pts;
    const bool fine_built = d == 128 && (K == 16 || K == 32);
    return (fine_built && lfa_fine_tiles(d, K, B, N)) ? pts / 2 : pts;
}

extern "C" int r3d_lfa_pool2_bwd_train(const float* xyz, long long xyz_bstride, const int32_t* idx, const float* feat,
                                       long long feat_bstride, const float* w_rpe1, const float* a_rpe1,
                                       const float* b_rpe1, const float* w_rpe2T, const float* a_rpe2,
                                       const float* b_rpe2, const float* w_scoreT, const float* w_score,
                                       const float* dpooled, float* dfeat, long long dfeat_bstride, float* dw_score,
                                       float* du2_tiles, double* sum_du2, int B, int N, int K, int d,
                                       r3d_stream_t stream) {
    if (B < 0 || N < 0 || K <= 0 || d <= 0) return R3D_EINVAL;
    if (B == 0 || N == 0) return R3D_OK;
    if (!xyz || !idx || !feat || !w_rpe1 || !a_rpe1 || !b_rpe1 || !w_rpe2T || !a_rpe2 || !b_rpe2 || !w_scoreT ||
        !w_score || !dpooled || !dfeat || !dw_score || !du2_tiles || !sum_du2)
        return R3D_EINVAL;
    const int h = d / 2;
    if (xyz_bstride == 0) xyz_bstride = (long long)N * 3;
    if (feat_bstride == 0) feat_bstride = (long long)N * h;
    if (dfeat_bstride == 0) dfeat_bstride = (long long)N * h;
    if (!is_aligned(feat, 16) || !is_aligned(dfeat, 16) || !is_aligned(w_scoreT, 16) || !is_aligned(w_score, 16) ||
        !is_aligned(w_rpe2T, 16) || !is_aligned(du2_tiles, 16) || (feat_bstride % 4) != 0 || (dfeat_bstride % 4) != 0)
        return R3D_EALIGN;
    LfaBwdArgs a{xyz, xyz_bstride, idx, feat, feat_bstride, w_rpe1, a_rpe1, b_rpe1, w_rpe2T, a_rpe2, b_rpe2, nullptr,
                 w_scoreT, w_score, dpooled, dfeat, dfeat_bstride, dw_score, nullptr, nullptr, nullptr, du2_tiles,
                 sum_du2, B, N};
    return dispatch_bwd<2>(d, K, a, static_cast<cudaStream_t>(stream));
}

extern "C" int r3d_lfa_bn2_bwd(const float* xyz, long long xyz_bstride, const int32_t* idx, const float* w_rpe1,
                               const float* a_rpe1, const float* b_rpe1, const float* du2_tiles, const float* w_rpe2T,
                               const float* w_rpe2, const float* bn2, double* g1, double* dw2, int B, int N, int K,
                               int d, r3d_stream_t stream) {
    if (B < 0 || N < 0 || K <= 0 || d <= 0) return R3D_EINVAL;
    if (B == 0 || N == 0) return R3D_OK;
    if (!xyz || !idx || !w_rpe1 || !a_rpe1 || !b_rpe1 || !du2_tiles || !w_rpe2T || !w_rpe2 || !bn2 || !g1 || !dw2)
        return R3D_EINVAL;
    if (!is_aligned(du2_tiles, 16) || !is_aligned(w_rpe2T, 16) || !is_aligned(w_rpe2, 16)) return R3D_EALIGN;
    if (xyz_bstride == 0) xyz_bstride = (long long)N * 3;
    LfaMomArgs a{xyz, xyz_bstride, idx, w_rpe1, a_rpe1, b_rpe1, nullptr, nullptr, nullptr, nullptr, nullptr, g1,
                 du2_tiles, w_rpe2T, w_rpe2, bn2, dw2, B, N};
    return dispatch_mom<3>(d, K, a, static_cast<cudaStream_t>(stream));
}

extern "C" int r3d_lfa_moments(int mode, const float* xyz, long long xyz_bstride, const int32_t* idx,
                               const float* w_rpe1, const float* a_rpe1, const float* b_rpe1, double* m_rpe,
                               double* m_r1, double* s_r1, const float* gsym, const float* gsum, double* g1, int B,
                               int N, int K, int d, r3d_stream_t stream) {
    if (mode < 0 || mode > 2 || B < 0 || N < 0 || K <= 0 || d <= 0) return R3D_EINVAL;
    if (B == 0 || N == 0) return R3D_OK;
    if (!xyz || !idx) return R3D_EINVAL;
    if (mode == 0 && !m_rpe) return R3D_EINVAL;
    if (mode != 0 && (!w_rpe1 || !a_rpe1 || !b_rpe1)) return R3D_EINVAL;
    if (mode == 1 && (!m_r1 || !s_r1)) return R3D_EINVAL;
    if (mode == 2 && (!gsym || !gsum || !g1 || !is_aligned(gsym, 16))) return R3D_EINVAL;
    if (xyz_bstride == 0) xyz_bstride = (long long)N * 3;
    LfaMomArgs a{xyz, xyz_bstride, idx, w_rpe1, a_rpe1, b_rpe1, m_rpe, m_r1, s_r1, gsym, gsum, g1,
                 nullptr, nullptr, nullptr, nullptr, nullptr, B, N};
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (mode == 0) {
        const long long rows = (long long)B * N * K;
        long long blocks = (rows + 256 * 16 - 1) / (256 * 16);
        if (blocks > (long long)kNumSMs * 8) blocks = (long long)kNumSMs * 8;
        lfa_rpe_moments_kernel<<<(unsigned)blocks, 256, 0, st>>>(xyz, xyz_bstride, idx, m_rpe, N, K, rows);
        R3D_LAUNCH_CHECK("lfa_rpe_moments_kernel");
        return R3D_OK;
    }
    if (mode == 1) return dispatch_mom<1>(d, K, a, st);
    return dispatch_mom<2>(d, K, a, st);
}
