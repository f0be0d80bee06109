// Building blocks shared by the fused LocSE + attentive-pooling kernels (lfa.cu forward, lfa_bwd.cu
// backward and BatchNorm-moment kernels).
//
// A CTA owns PTS points = PTS*K (point, neighbour) rows.  Row-indexed operands live in shared memory
// CHANNEL-MAJOR: buf[c * ROWS_PAD + p * PSTRIDE + k].  PSTRIDE pads every point to K+4 floats so that the
// 16-row register tiles of different points start in different banks; ROWS_PAD == 4 (mod 32) so that
// consecutive channels are 4 banks apart (conflict-free LDS.128 across channels in reduce_gemm).
#pragma once
#include "common.cuh"

#include <math_constants.h>

namespace r3d {

// rows of the per-thread register tile by layer width (see lfa.cu)
__host__ __device__ constexpr int lfa_rows_per_thread(int d) { return d <= 16 ? 4 : (d <= 64 ? 8 : 16); }

constexpr int kRpeRows = 16;  // rpe buffer: 10 encoding channels, a row of ones (col 10), 5 zero rows

template <int D, int K, int THREADS = 128, int WSTAGE_MAX = 4096, int ROWS_PER_THREAD = 16>
struct LfaCfg {
    static constexpr int NT = THREADS;
    static constexpr int H = D / 2;
    static constexpr int RT = ROWS_PER_THREAD;          // rows of the register tile (16; 4 for narrow layers: more threads per point)
    static constexpr int RH = K / RT;                   // row groups per point
    static constexpr int CG = D / 8;                    // column groups (8 columns per thread)
    static constexpr int TPP = RH * CG;                 // threads per point
    static constexpr int PTS = THREADS / TPP;           // points per CTA
    static constexpr int PAD = (CG >= 32) ? 0 : 4;
    static constexpr int PSTRIDE = K + PAD;
    static constexpr int ROWS_BASE = PTS * PSTRIDE;
    static constexpr int ROWS_PAD = ROWS_BASE + ((4 - ROWS_BASE % 32) + 32) % 32;
    static constexpr int ROWS = PTS * K;
    static constexpr int X_FLOATS = D * ROWS_PAD;
    static constexpr int WSTAGE = (D * D < WSTAGE_MAX) ? D * D : WSTAGE_MAX;   // floats per weight-ring stage
    static_assert(K % 16 == 0 && K >= 16 && K <= 64 && K % RT == 0 && RT % 4 == 0, "K must be a multiple of 16 up to 64");
    static_assert(D % 8 == 0 && TPP <= THREADS && THREADS % TPP == 0 && THREADS % 32 == 0, "unsupported width");
};

// ---------------------------------------------------------------- weight streaming (1-D TMA bulk copies)
// (rows x width) row-major matrix in global memory -> ring of two shared-memory stages.
struct WPipe {
    float* ring;
    uint64_t* bars;
    uint32_t count;   // chunks consumed so far by this CTA (selects stage and mbarrier phase)
    int stage_floats;
};

__device__ __forceinline__ void wpipe_issue(const WPipe& p, uint32_t chunk_no, const float* src, uint32_t floats) {
    const uint32_t s = chunk_no & 1u;
    mbar_expect_tx(&p.bars[s], floats * 4u);
    tma_bulk_g2s(p.ring + s * p.stage_floats, src, floats * 4u, &p.bars[s]);
}

// acc[RT][4*NC] += A[RT rows][Kred] * W[Kred][cols].  A in shared memory channel-major with leading
// dimension lda (rows row0..row0+RT-1 of channel kk at A[kk*lda + row0 ..]); W (Kred x width) row-major in
// global memory.  Thread columns: for q < NC: q*qstride + g*4 + {0..3}.  Ends with a CTA barrier.
template <int NC, int NT, int RT>
__device__ __forceinline__ void gemm_stream(float (&acc)[RT][4 * NC], const float* __restrict__ A, int lda,
                                            int row0, int Kred, const float* __restrict__ Wg, int width,
                                            int qstride, int g, WPipe& pipe, int tid) {
    const int kc_max = pipe.stage_floats / width;
    const int kc = kc_max < Kred ? kc_max : Kred;  // reduction rows per chunk
    const int nchunks = (Kred + kc - 1) / kc;
    if (tid == 0) wpipe_issue(pipe, pipe.count, Wg, (uint32_t)(kc * width));
    for (int ch = 0; ch < nchunks; ++ch) {
        const uint32_t cur = pipe.count + ch;
        if (tid == 0 && ch + 1 < nchunks) {
            const int rows_next = min(kc, Kred - (ch + 1) * kc);
            wpipe_issue(pipe, cur + 1, Wg + (size_t)(ch + 1) * kc * width, (uint32_t)(rows_next * width));
        }
        mbar_wait(&pipe.bars[cur & 1u], (cur >> 1) & 1u);
        const float* Wst = pipe.ring + (cur & 1u) * pipe.stage_floats;
        const int rows_here = min(kc, Kred - ch * kc);
        const float* Ap = A + (size_t)(ch * kc) * lda + row0;
#pragma unroll 2
        for (int kk = 0; kk < rows_here; ++kk) {
            float av[RT];
#pragma unroll
            for (int v = 0; v < RT / 4; ++v) {
                const float4 t = *reinterpret_cast<const float4*>(Ap + (size_t)kk * lda + 4 * v);
                av[4 * v + 0] = t.x; av[4 * v + 1] = t.y; av[4 * v + 2] = t.z; av[4 * v + 3] = t.w;
            }
            float wv[4 * NC];
#pragma unroll
            for (int q = 0; q < NC; ++q) {
                const float4 t = *reinterpret_cast<const float4*>(Wst + kk * width + q * qstride + g * 4);
                wv[4 * q + 0] = t.x; wv[4 * q + 1] = t.y; wv[4 * q + 2] = t.z; wv[4 * q + 3] = t.w;
            }
#pragma unroll
            for (int r = 0; r < RT; ++r)
#pragma unroll
                for (int j = 0; j < 4 * NC; ++j) acc[r][j] = fmaf(av[r], wv[j], acc[r][j]);
        }
        __syncthreads();
    }
    pipe.count += nchunks;
}

// out[j * ld_out + c] += sum over the CTA's rows of A[j][row] * Bm[c][row]   (j < NJ, c < NCOLS)
// Both operands channel-major in shared memory with leading dimension RP; rows are visited point by
// point (PTS points of K rows, PSTRIDE apart) so the padding floats are never read.  Results are added
// to global memory with one atomic per element and CTA (OutT = float or double).
// Mapping: lane -> channel(s) c = lane % CL + 32*u, warps (and lane / CL when NCOLS < 32) -> groups of 4 j.
template <int NJ, int NCOLS, int NT, int PTS, int K, int PSTRIDE, typename OutT>
__device__ __forceinline__ void reduce_gemm(const float* __restrict__ A, const float* __restrict__ Bm, int RP,
                                            OutT* __restrict__ out, int ld_out, int c_limit, int tid) {
    constexpr int CL = NCOLS < 32 ? NCOLS : 32;
    constexpr int TC = NCOLS / CL;
    constexpr int JS = 32 / CL;                       // j-groups per warp
    constexpr int NW = NT / 32;
    constexpr int JT = (NJ + 3) / 4;                  // j tiles of 4
    const int lane = tid & 31, warp = tid >> 5;
    const int c0 = lane % CL, jsub = lane / CL;
    for (int jt = warp * JS + jsub; jt < JT; jt += NW * JS) {
        const int j0 = jt * 4;
        float acc[4][TC];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int u = 0; u < TC; ++u) acc[a][u] = 0.f;
        for (int p = 0; p < PTS; ++p) {
#pragma unroll
            for (int v = 0; v < K / 4; ++v) {
                const int off = p * PSTRIDE + 4 * v;
                float4 av[4];
#pragma unroll
                for (int a = 0; a < 4; ++a)
                    av[a] = (j0 + a < NJ) ? *reinterpret_cast<const float4*>(A + (size_t)(j0 + a) * RP + off)
                                          : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int u = 0; u < TC; ++u) {
                    const float4 b = *reinterpret_cast<const float4*>(Bm + (size_t)(c0 + 32 * u) * RP + off);
#pragma unroll
                    for (int a = 0; a < 4; ++a) {
                        acc[a][u] = fmaf(av[a].x, b.x, acc[a][u]);
                        acc[a][u] = fmaf(av[a].y, b.y, acc[a][u]);
                        acc[a][u] = fmaf(av[a].z, b.z, acc[a][u]);
                        acc[a][u] = fmaf(av[a].w, b.w, acc[a][u]);
                    }
                }
            }
        }
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int u = 0; u < TC; ++u) {
                const int c = c0 + 32 * u;
                if (j0 + a < NJ && c < c_limit) atomicAdd(out + (size_t)(j0 + a) * ld_out + c, (OutT)acc[a][u]);
            }
    }
}

// ------------------------------------------------------------------------------ row construction
// Relative position encoding of one (point, neighbour) row with the KNN contract's rounding sequence, so
// that |p_i - p_j| equals sqrt of the KNN d2 bit for bit (modules.py:170-186).
__device__ __forceinline__ void rpe_of_row(const float* __restrict__ xyz_b, int pi, int pj, float (&rpe)[10]) {
    const float ix = xyz_b[(size_t)pi * 3 + 0], iy = xyz_b[(size_t)pi * 3 + 1], iz = xyz_b[(size_t)pi * 3 + 2];
    const float jx = xyz_b[(size_t)pj * 3 + 0], jy = xyz_b[(size_t)pj * 3 + 1], jz = xyz_b[(size_t)pj * 3 + 2];
    const float dx = __fsub_rn(ix, jx), dy = __fsub_rn(iy, jy), dz = __fsub_rn(iz, jz);
    const float dist = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz)));
    rpe[0] = ix; rpe[1] = iy; rpe[2] = iz; rpe[3] = jx; rpe[4] = jy; rpe[5] = jz;
    rpe[6] = dx; rpe[7] = dy; rpe[8] = dz; rpe[9] = dist;
}

// r1[ch] = relu(a[ch] * (W1[ch] . rpe) + b[ch]); Pw1 holds W1 padded to 12 floats per channel
__device__ __forceinline__ float rpe_mlp1(const float* __restrict__ Pw1, const float* __restrict__ Pa1,
                                          const float* __restrict__ Pb1, int ch, const float (&rpe)[10]) {
    const float4 w0 = *reinterpret_cast<const float4*>(Pw1 + ch * 12);
    const float4 w1 = *reinterpret_cast<const float4*>(Pw1 + ch * 12 + 4);
    const float4 w2 = *reinterpret_cast<const float4*>(Pw1 + ch * 12 + 8);
    float z = w0.x * rpe[0];
    z = fmaf(w0.y, rpe[1], z); z = fmaf(w0.z, rpe[2], z); z = fmaf(w0.w, rpe[3], z);
    z = fmaf(w1.x, rpe[4], z); z = fmaf(w1.y, rpe[5], z); z = fmaf(w1.z, rpe[6], z);
    z = fmaf(w1.w, rpe[7], z); z = fmaf(w2.x, rpe[8], z); z = fmaf(w2.y, rpe[9], z);
    return fmaxf(fmaf(z, Pa1[ch], Pb1[ch]), 0.f);
}

// mlp_rpe1 parameters -> shared memory (Pw1 [H][12], Pa1 [H], Pb1 [H])
template <int H, int NT>
__device__ __forceinline__ void load_rpe1_params(float* Pw1, float* Pa1, float* Pb1, const float* __restrict__ w,
                                                 const float* __restrict__ a, const float* __restrict__ b, int tid) {
    for (int i = tid; i < H * 12; i += NT) {
        const int ch = i / 12, m = i % 12;
        Pw1[i] = (m < 10) ? w[ch * 10 + m] : 0.f;
    }
    for (int i = tid; i < H; i += NT) {
        Pa1[i] = a[i];
        Pb1[i] = b[i];
    }
}

// 16-byte vector reduction to global memory (sm_90+): out[0..3] += v
__device__ __forceinline__ void red_add_v4(float* out, float4 v) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(out), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                 : "memory");
}

}  // namespace r3d
