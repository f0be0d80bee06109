// Per-point ("shared") MLP layers for sm_100a: y = act(scale * (W [xa ; xb]) + shift).
//
// Replaces every SharedMLP / Linear on single points of the reference
// (randlanet/utils/modules.py:60-104 SharedMLP; call sites :314 mlp1, :325 mlp2 + shortcut, :253 pool
// MLPs, :565-566 fc_start + bn_start, :591 bottleneck, :594-605 decoder, :610 fc_end) together with the
// tensor shuffling the reference does around them:
//   * row gather on source A  — the decoder's 1-NN up-sampling gather (modules.py:359-363), the
//     permutation at the start (:572-573) and the inverse permutation at the end (:608);
//   * concatenation of a second source B — the decoder skip concat (modules.py:600-602) and, with the
//     BN scales folded into the weights, the residual sum mlp2(p2) + shortcut(input) (:325).
// Eval-mode BatchNorm (eps 1e-6) is an affine map, folded by the host into (scale, shift) or directly
// into W; train-mode layers run with scale = NULL/identity and write the pre-BN tensor.
//
// Layout: point-major.  Source rows are contiguous C floats; clouds are `*_bstride` floats apart so
// that a prefix view [:, :n] of a longer cloud needs no copy.  Weights come transposed, wT (Cin, Cout).
//
// Two kernels:
//   pw_gemm_kernel<TM,TN>  register-tiled FP32 GEMM (8x8 per thread, operands staged in shared memory
//                          transposed so both are read with LDS.128) for Cout >= 16;
//   pw_small_kernel        one thread per row for Cout <= 16 (fc_start 3->8, last decoder stage,
//                          class logits): HBM-bound streaming.
#include "pointwise_common.cuh"

#include <atomic>
#include <cooperative_groups.h>
#include <mutex>
#include <unordered_map>

namespace r3d {

static std::atomic<int> g_pw_tensor_cores{1};

// ------------------------------------------------------------------------------------ GEMM kernel
constexpr int kPwKC = 16;  // input channels per staged chunk

template <int TM, int TN>
__global__ void __launch_bounds__((TM / 8) * (TN / 8)) pw_gemm_kernel(PwArgs a) {
    constexpr int NT = (TM / 8) * (TN / 8);
    constexpr int TMP = TM + 4;
    __shared__ __align__(16) float As[kPwKC][TMP];
    __shared__ __align__(16) float Ws[kPwKC][TN];

    const int tid = threadIdx.x;
    const int tr = tid / (TN / 8);   // row group
    const int tc = tid % (TN / 8);   // col group: cols tc*4 + {0..3} and TN/2 + tc*4 + {0..3}
    const long long M = (long long)a.B * a.n;
    const long long m0 = (long long)blockIdx.x * TM;
    const int n0 = blockIdx.y * TN;
    const int cin = a.ca + a.cb;

    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    // each thread stages A elements (row r, 4 consecutive channels): index i -> c4 = i % 4, r = i / 4
    constexpr int A_ITEMS = TM * (kPwKC / 4);
    constexpr int W_ITEMS = kPwKC * TN / 4;

    for (int c0 = 0; c0 < cin; c0 += kPwKC) {
        // ---- stage A chunk (transposed into [c][row])
        for (int i = tid; i < A_ITEMS; i += NT) {
            const int c4 = i % (kPwKC / 4), r = i / (kPwKC / 4);
            const long long m = m0 + r;
            float v[4] = {0.f, 0.f, 0.f, 0.f};
            if (m < M) {
                const int b = (int)(m / a.n), n = (int)(m % a.n);
                const int c = c0 + c4 * 4;
                // a 4-channel group never straddles the A/B boundary when ca % 4 == 0 (checked by the host);
                // otherwise fall back to scalar selection
                if ((a.ca & 3) == 0 && (a.cb & 3) == 0) {
                    if (c < cin) {
                        const bool second = c >= a.ca;
                        const float* row = src_row(a, b, n, second);
                        const float4 t = *reinterpret_cast<const float4*>(row + (second ? c - a.ca : c));
                        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
                    }
                } else {
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int cc = c + u;
                        if (cc < cin) {
                            const bool second = cc >= a.ca;
                            v[u] = src_row(a, b, n, second)[second ? cc - a.ca : cc];
                        }
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) As[c4 * 4 + u][r] = v[u];
        }
        // ---- stage W chunk
        if (a.w_out_in) {
            // weight rows are output channels: read along cin, store transposed
            for (int i = tid; i < TN * kPwKC; i += NT) {
                const int col = i / kPwKC, kk = i % kPwKC;
                const int c = c0 + kk, oc = n0 + col;
                Ws[kk][col] = (c < cin && oc < a.cout) ? a.wT[(size_t)oc * cin + c] : 0.f;
            }
        } else {
            for (int i = tid; i < W_ITEMS; i += NT) {
                const int kk = i / (TN / 4), j4 = i % (TN / 4);
                const int c = c0 + kk, col = n0 + j4 * 4;
                float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
                if (c < cin) {
                    const float* wr = a.wT + (size_t)c * a.cout;
                    if ((a.cout & 3) == 0 && col + 3 < a.cout) {
                        t = *reinterpret_cast<const float4*>(wr + col);
                    } else {
                        if (col + 0 < a.cout) t.x = wr[col + 0];
                        if (col + 1 < a.cout) t.y = wr[col + 1];
                        if (col + 2 < a.cout) t.z = wr[col + 2];
                        if (col + 3 < a.cout) t.w = wr[col + 3];
                    }
                }
                *reinterpret_cast<float4*>(&Ws[kk][j4 * 4]) = t;
            }
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < kPwKC; ++kk) {
            const float4 a0 = *reinterpret_cast<const float4*>(&As[kk][tr * 8]);
            const float4 a1 = *reinterpret_cast<const float4*>(&As[kk][tr * 8 + 4]);
            const float4 w0 = *reinterpret_cast<const float4*>(&Ws[kk][tc * 4]);
            const float4 w1 = *reinterpret_cast<const float4*>(&Ws[kk][TN / 2 + tc * 4]);
            const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], wv[j], acc[i][j]);
        }
        __syncthreads();
    }

    // ---- epilogue
    // fp64 shared accumulators: the order of the atomics then only matters at 1e-16, so the batch statistics — and
    // with them which side of a ReLU kink a borderline activation falls on — are reproducible from run to run
    __shared__ double csum[2][TN];
    if (a.stats) {
        for (int i = tid; i < 2 * TN; i += NT) (&csum[0][0])[i] = 0.0;
        __syncthreads();
    }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int col = n0 + h * (TN / 2) + tc * 4;
        float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
        float sc[4], sh[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const bool ok = col + j < a.cout;
            sc[j] = (a.scale && ok) ? a.scale[col + j] : 1.f;
            sh[j] = (a.shift && ok) ? a.shift[col + j] : 0.f;
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const long long m = m0 + tr * 8 + i;
            if (m >= M) continue;
            const int b = (int)(m / a.n), n = (int)(m % a.n);
            float o[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                o[j] = apply_act(fmaf(acc[i][h * 4 + j], sc[j], sh[j]), a.act, a.slope);
                s1[j] += o[j];
                s2[j] = fmaf(o[j], o[j], s2[j]);
            }
            if (a.transpose_out) {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (col + j < a.cout) a.y[(size_t)b * a.y_bstride + (size_t)(col + j) * a.n + n] = o[j];
            } else {
                float* yr = a.y + (size_t)b * a.y_bstride + (size_t)n * a.y_ld;
                if (col + 3 < a.cout && (a.y_ld & 3) == 0 && (a.y_bstride & 3) == 0) {
                    *reinterpret_cast<float4*>(yr + col) = make_float4(o[0], o[1], o[2], o[3]);
                } else {
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if (col + j < a.cout) yr[col + j] = o[j];
                }
            }
        }
        if (a.stats) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                atomicAdd(&csum[0][h * (TN / 2) + tc * 4 + j], (double)s1[j]);
                atomicAdd(&csum[1][h * (TN / 2) + tc * 4 + j], (double)s2[j]);
            }
        }
    }
    if (a.stats) {
        __syncthreads();
        for (int i = tid; i < TN; i += NT) {
            if (n0 + i < a.cout) {
                atomicAdd(a.stats + n0 + i, csum[0][i]);
                atomicAdd(a.stats + a.cout + n0 + i, csum[1][i]);
            }
        }
    }
}

// ------------------------------------------------------------------------- pipelined GEMM kernel
// Same math as pw_gemm_kernel for channel counts that are multiples of 4, restructured for latency: source row
// pointers (gather, batch stride) are resolved once per thread, and the global loads of chunk c+1 are issued
// into registers before the FMAs of chunk c (double-buffered shared memory, ONE barrier per chunk).  The launch
// list of the small-cloud training step showed the unpipelined kernel at 61 us per launch (24 % of the step):
// 16..64 chunks of load -> barrier -> 16 k-steps -> barrier on a handful of CTAs.
// RT = register tile edge (8: 8x8 per thread; 4: 4x4 per thread, more threads/CTAs for small row counts).
// SPLITK > 1: a thread-block cluster of SPLITK CTAs (gridDim.z) shares one output tile, each CTA contracting 1/SPLITK
// of the input channels; the partial tiles are summed through distributed shared memory by the cluster's rank-0 CTA,
// which then runs the epilogue.  For few-row wide layers the chain of dependent load -> FMA steps is the whole cost
// of the kernel: a cluster of 4 cuts it 4x and puts 4x more SMs to work.
template <int TM, int TN, int RT, int KC = kPwKC, int SPLITK = 1>
__global__ void __launch_bounds__((TM / RT) * (TN / RT)) pw_gemm_fast_kernel(PwArgs a) {
    constexpr int NT = (TM / RT) * (TN / RT);
    constexpr int TMP = TM + 4;
    constexpr int A_PER = TM * (KC / 4) / NT;
    constexpr int W_PER = KC * TN / 4 / NT;
    static_assert(A_PER >= 1 && W_PER >= 1 && (TM * (KC / 4)) % NT == 0 && (KC * TN / 4) % NT == 0, "tile/threads");
    __shared__ __align__(16) float As[2][KC][TMP];
    __shared__ __align__(16) float Ws[2][KC][TN];

    const int tid = threadIdx.x;
    const int tr = tid / (TN / RT);
    const int tc = tid % (TN / RT);
    const long long M = (long long)a.B * a.n;
    const long long m0 = (long long)blockIdx.x * TM;
    const int n0 = blockIdx.y * TN;
    const int cin = a.ca + a.cb;

    // ---- per-thread source rows
    const float* rowA[A_PER];
    const float* rowB[A_PER];
#pragma unroll
    for (int k = 0; k < A_PER; ++k) {
        const int r = (tid + k * NT) / (KC / 4);
        const long long m = m0 + r;
        rowA[k] = nullptr;
        rowB[k] = nullptr;
        if (m < M) {
            const int b = (int)(m / a.n), n = (int)(m % a.n);
            rowA[k] = src_row(a, b, n, false);
            if (a.cb > 0) rowB[k] = src_row(a, b, n, true);
        }
    }
    float4 ra[A_PER], rw[W_PER];
    auto load = [&](int c0) {
#pragma unroll
        for (int k = 0; k < A_PER; ++k) {
            const int c = c0 + ((tid + k * NT) % (KC / 4)) * 4;
            float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
            if (rowA[k] && c < cin) t = *reinterpret_cast<const float4*>(c < a.ca ? rowA[k] + c : rowB[k] + (c - a.ca));
            ra[k] = t;
        }
#pragma unroll
        for (int k = 0; k < W_PER; ++k) {
            const int i = tid + k * NT;
            float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
            if (a.w_out_in) {                       // (cout, cin): float4 along cin
                const int col = i / (KC / 4), c = c0 + (i % (KC / 4)) * 4;
                if (c < cin && n0 + col < a.cout) t = *reinterpret_cast<const float4*>(a.wT + (size_t)(n0 + col) * cin + c);
            } else {                                // (cin, cout): float4 along cout
                const int kk = i / (TN / 4), col = n0 + (i % (TN / 4)) * 4;
                if (c0 + kk < cin) {
                    const float* wr = a.wT + (size_t)(c0 + kk) * a.cout;
                    if ((a.cout & 3) == 0 && col + 3 < a.cout) {
                        t = *reinterpret_cast<const float4*>(wr + col);
                    } else {
                        if (col + 0 < a.cout) t.x = wr[col + 0];
                        if (col + 1 < a.cout) t.y = wr[col + 1];
                        if (col + 2 < a.cout) t.z = wr[col + 2];
                        if (col + 3 < a.cout) t.w = wr[col + 3];
                    }
                }
            }
            rw[k] = t;
        }
    };
    auto store = [&](int buf) {
#pragma unroll
        for (int k = 0; k < A_PER; ++k) {
            const int i = tid + k * NT;
            const int c4 = i % (KC / 4), r = i / (KC / 4);
            As[buf][c4 * 4 + 0][r] = ra[k].x;
            As[buf][c4 * 4 + 1][r] = ra[k].y;
            As[buf][c4 * 4 + 2][r] = ra[k].z;
            As[buf][c4 * 4 + 3][r] = ra[k].w;
        }
#pragma unroll
        for (int k = 0; k < W_PER; ++k) {
            const int i = tid + k * NT;
            if (a.w_out_in) {
                const int col = i / (KC / 4), k4 = (i % (KC / 4)) * 4;
                Ws[buf][k4 + 0][col] = rw[k].x;
                Ws[buf][k4 + 1][col] = rw[k].y;
                Ws[buf][k4 + 2][col] = rw[k].z;
                Ws[buf][k4 + 3][col] = rw[k].w;
            } else {
                const int kk = i / (TN / 4), j4 = i % (TN / 4);
                *reinterpret_cast<float4*>(&Ws[buf][kk][j4 * 4]) = rw[k];
            }
        }
    };

    float acc[RT][RT];
#pragma unroll
    for (int i = 0; i < RT; ++i)
#pragma unroll
        for (int j = 0; j < RT; ++j) acc[i][j] = 0.f;

    // this CTA's share of the contraction (whole chunks)
    const int chunks_all = (cin + KC - 1) / KC;
    const int chunks_per = (chunks_all + SPLITK - 1) / SPLITK;
    const int ch_begin = (SPLITK > 1 ? (int)blockIdx.z : 0) * chunks_per;
    const int ch_end = ch_begin + chunks_per < chunks_all ? ch_begin + chunks_per : chunks_all;
    const int nchunks = ch_end > ch_begin ? ch_end - ch_begin : 0;
    const int cbase = ch_begin * KC;
    if (nchunks > 0) {
        load(cbase);
        store(0);
    }
    __syncthreads();
    for (int ch = 0; ch < nchunks; ++ch) {
        const int buf = ch & 1;
        if (ch + 1 < nchunks) load(cbase + (ch + 1) * KC);
#pragma unroll
        for (int kk = 0; kk < KC; ++kk) {
            float av[RT], wv[RT];
            if (RT == 8) {
                const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][kk][tr * 8]);
                const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][kk][tr * 8 + 4]);
                const float4 w0 = *reinterpret_cast<const float4*>(&Ws[buf][kk][tc * 4]);
                const float4 w1 = *reinterpret_cast<const float4*>(&Ws[buf][kk][TN / 2 + tc * 4]);
                av[0] = a0.x; av[1] = a0.y; av[2] = a0.z; av[3] = a0.w;
                av[RT - 4] = a1.x; av[RT - 3] = a1.y; av[RT - 2] = a1.z; av[RT - 1] = a1.w;
                wv[0] = w0.x; wv[1] = w0.y; wv[2] = w0.z; wv[3] = w0.w;
                wv[RT - 4] = w1.x; wv[RT - 3] = w1.y; wv[RT - 2] = w1.z; wv[RT - 1] = w1.w;
            } else {
                const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][kk][tr * 4]);
                const float4 w0 = *reinterpret_cast<const float4*>(&Ws[buf][kk][tc * 4]);
                av[0] = a0.x; av[1] = a0.y; av[2] = a0.z; av[3] = a0.w;
                wv[0] = w0.x; wv[1] = w0.y; wv[2] = w0.z; wv[3] = w0.w;
            }
#pragma unroll
            for (int i = 0; i < RT; ++i)
#pragma unroll
                for (int j = 0; j < RT; ++j) acc[i][j] = fmaf(av[i], wv[j], acc[i][j]);
        }
        if (ch + 1 < nchunks) store(buf ^ 1);
        __syncthreads();
    }

    if (SPLITK > 1) {
        // partial tiles -> rank 0 through distributed shared memory (the staging buffers are free now)
        namespace cg = cooperative_groups;
        cg::cluster_group cluster = cg::this_cluster();
        static_assert(SPLITK == 1 || RT * RT * NT <= 2 * KC * TN, "partial tile must fit the W staging buffer");
        float* part = &Ws[0][0][0];
        const unsigned rank = cluster.block_rank();
        if (rank != 0) {
#pragma unroll
            for (int i = 0; i < RT; ++i)
#pragma unroll
                for (int j = 0; j < RT; ++j) part[(i * RT + j) * NT + tid] = acc[i][j];
        }
        cluster.sync();
        if (rank == 0) {
            for (unsigned r = 1; r < (unsigned)SPLITK; ++r) {
                const float* remote = cluster.map_shared_rank(part, r);
#pragma unroll
                for (int i = 0; i < RT; ++i)
#pragma unroll
                    for (int j = 0; j < RT; ++j) acc[i][j] += remote[(i * RT + j) * NT + tid];
            }
        }
        cluster.sync();          // the partial tiles stay mapped until rank 0 has read them
        if (rank != 0) return;
    }

    // ---- epilogue
    constexpr int NH = RT / 4;   // column quads per thread
    static_assert(2 * (TM / RT) * TN <= 2 * KC * TMP, "statistics partials must fit the A staging buffer");
    float* part = &As[0][0][0];  // free: the main loop (and the split-K exchange) ended with a barrier
#pragma unroll
    for (int h = 0; h < NH; ++h) {
        const int lcol = (RT == 8 ? h * (TN / 2) : 0) + tc * 4;
        const int col = n0 + lcol;
        float sc[4], sh[4], s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const bool ok = col + j < a.cout;
            sc[j] = (a.scale && ok) ? a.scale[col + j] : 1.f;
            sh[j] = (a.shift && ok) ? a.shift[col + j] : 0.f;
        }
#pragma unroll
        for (int i = 0; i < RT; ++i) {
            const long long m = m0 + tr * RT + i;
            if (m >= M) continue;
            const int b = (int)(m / a.n), n = (int)(m % a.n);
            float o[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                o[j] = apply_act(fmaf(acc[i][h * 4 + j], sc[j], sh[j]), a.act, a.slope);
                s1[j] += o[j];
                s2[j] = fmaf(o[j], o[j], s2[j]);
            }
            if (a.transpose_out) {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (col + j < a.cout) a.y[(size_t)b * a.y_bstride + (size_t)(col + j) * a.n + n] = o[j];
            } else {
                float* yr = a.y + (size_t)b * a.y_bstride + (size_t)n * a.y_ld;
                if (col + 3 < a.cout && (a.y_ld & 3) == 0 && (a.y_bstride & 3) == 0) {
                    *reinterpret_cast<float4*>(yr + col) = make_float4(o[0], o[1], o[2], o[3]);
                } else {
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if (col + j < a.cout) yr[col + j] = o[j];
                }
            }
        }
        if (a.stats) {
            // per-thread partial sums (its RT rows) into the free A staging buffer: [2][TM / RT][TN] floats.  No
            // shared-memory atomics: fp64 ones are compare-and-swap loops and 16 threads share every column.
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                part[(0 * (TM / RT) + tr) * TN + lcol + j] = s1[j];
                part[(1 * (TM / RT) + tr) * TN + lcol + j] = s2[j];
            }
        }
    }
    if (a.stats) {
        __syncthreads();
        for (int i = tid; i < TN; i += NT) {
            if (n0 + i < a.cout) {
                double t1 = 0.0, t2 = 0.0;            // fp64 across the row groups: run-to-run reproducible
#pragma unroll 4
                for (int t = 0; t < TM / RT; ++t) {
                    t1 += (double)part[(0 * (TM / RT) + t) * TN + i];
                    t2 += (double)part[(1 * (TM / RT) + t) * TN + i];
                }
                atomicAdd(a.stats + n0 + i, t1);
                atomicAdd(a.stats + a.cout + n0 + i, t2);
            }
        }
    }
    if (SPLITK == 1 && a.bn.y) {
        // fused train-mode BatchNorm: the accumulators are the conv outputs z (no affine / activation in this mode)
        __threadfence();
        cooperative_groups::this_grid().sync();
#pragma unroll
        for (int h = 0; h < NH; ++h) {
            const int col = n0 + (RT == 8 ? h * (TN / 2) : 0) + tc * 4;
            float ca[4], cc[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                ca[j] = cc[j] = 0.f;
                if (col + j < a.cout) pw_bn_coeffs(a, col + j, M, blockIdx.x == 0 && tr == 0, ca[j], cc[j]);
            }
#pragma unroll
            for (int i = 0; i < RT; ++i) {
                const long long m = m0 + tr * RT + i;
                if (m >= M) continue;
                float o[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) o[j] = apply_act(fmaf(acc[i][h * 4 + j], ca[j], cc[j]), a.bn.act, a.bn.slope);
                float* yr = a.bn.y + (size_t)m * a.cout;
                if (col + 3 < a.cout && (a.cout & 3) == 0) {
                    *reinterpret_cast<float4*>(yr + col) = make_float4(o[0], o[1], o[2], o[3]);
                } else {
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if (col + j < a.cout) yr[col + j] = o[j];
                }
            }
        }
        if (blockIdx.x == 0 && blockIdx.y == 0 && tid == 0 && a.bn.num_batches) *a.bn.num_batches += 1;
    }
}

// ----------------------------------------------------------------------------------- small kernel
// Cout <= 16: one thread per output row; weights (cin x cout) live in shared memory.
constexpr int kPwSmallMaxCout = 16;
constexpr int kPwSmallMaxW = 4096;  // floats of weights staged in smem (cin*cout <= 4096)

__global__ void __launch_bounds__(256) pw_small_kernel(PwArgs a) {
    __shared__ float Ws[kPwSmallMaxW];
    __shared__ double cta_sum[2][kPwSmallMaxCout];
    if (threadIdx.x < 2 * kPwSmallMaxCout) (&cta_sum[0][0])[threadIdx.x] = 0.0;
    const int cin = a.ca + a.cb;
    for (int i = threadIdx.x; i < cin * a.cout; i += blockDim.x)
        Ws[i] = a.w_out_in ? a.wT[(size_t)(i % a.cout) * cin + i / a.cout] : a.wT[i];
    __syncthreads();
    const long long M = (long long)a.B * a.n;
    const long long m_raw = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = m_raw < M;
    if (!live && !a.stats) return;
    const long long m = live ? m_raw : M - 1;
    const int b = (int)(m / a.n), n = (int)(m % a.n);
    float acc[kPwSmallMaxCout];
#pragma unroll
    for (int j = 0; j < kPwSmallMaxCout; ++j) acc[j] = 0.f;
    const float* ra = src_row(a, b, n, false);
    if ((a.ca & 3) == 0) {                       // rows are 16-byte aligned (checked by the host): vector loads
        for (int c = 0; c < a.ca; c += 4) {
            const float4 x4 = *reinterpret_cast<const float4*>(ra + c);
            const float xv[4] = {x4.x, x4.y, x4.z, x4.w};
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int j = 0; j < kPwSmallMaxCout; ++j)
                    if (j < a.cout) acc[j] = fmaf(xv[u], Ws[(c + u) * a.cout + j], acc[j]);
        }
    } else {
        for (int c = 0; c < a.ca; ++c) {
            const float x = ra[c];
#pragma unroll
            for (int j = 0; j < kPwSmallMaxCout; ++j)
                if (j < a.cout) acc[j] = fmaf(x, Ws[c * a.cout + j], acc[j]);
        }
    }
    if (a.cb > 0) {
        const float* rb = src_row(a, b, n, true);
        if ((a.cb & 3) == 0) {
            for (int c = 0; c < a.cb; c += 4) {
                const float4 x4 = *reinterpret_cast<const float4*>(rb + c);
                const float xv[4] = {x4.x, x4.y, x4.z, x4.w};
#pragma unroll
                for (int u = 0; u < 4; ++u)
#pragma unroll
                    for (int j = 0; j < kPwSmallMaxCout; ++j)
                        if (j < a.cout) acc[j] = fmaf(xv[u], Ws[(a.ca + c + u) * a.cout + j], acc[j]);
            }
        } else {
            for (int c = 0; c < a.cb; ++c) {
                const float x = rb[c];
#pragma unroll
                for (int j = 0; j < kPwSmallMaxCout; ++j)
                    if (j < a.cout) acc[j] = fmaf(x, Ws[(a.ca + c) * a.cout + j], acc[j]);
            }
        }
    }
    // whole output rows as 16-byte stores when the layout allows (a row is cout contiguous floats)
    const bool vec_out = !a.transpose_out && (a.cout & 3) == 0 && (a.y_ld & 3) == 0 && (a.y_bstride & 3) == 0;
    float outv[kPwSmallMaxCout];
#pragma unroll
    for (int j = 0; j < kPwSmallMaxCout; ++j) {
        if (j >= a.cout) break;
        const float sc = a.scale ? a.scale[j] : 1.f, sh = a.shift ? a.shift[j] : 0.f;
        const float o = apply_act(fmaf(acc[j], sc, sh), a.act, a.slope);
        outv[j] = o;
        if (live && !vec_out) {
            if (a.transpose_out)
                a.y[(size_t)b * a.y_bstride + (size_t)j * a.n + n] = o;
            else
                a.y[(size_t)b * a.y_bstride + (size_t)n * a.y_ld + j] = o;
        }
        if (a.stats) {
            float s1 = live ? o : 0.f, s2 = live ? o * o : 0.f;
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                s1 += __shfl_xor_sync(0xffffffffu, s1, off);
                s2 += __shfl_xor_sync(0xffffffffu, s2, off);
            }
            if ((threadIdx.x & 31) == 0) {
                atomicAdd(&cta_sum[0][j], (double)s1);
                atomicAdd(&cta_sum[1][j], (double)s2);
            }
        }
    }
    if (a.stats) {          // one global (fp64) atomic per channel and CTA, not per warp: the 2*cout addresses are hot
        __syncthreads();
        if (threadIdx.x < a.cout) {
            atomicAdd(a.stats + threadIdx.x, cta_sum[0][threadIdx.x]);
            atomicAdd(a.stats + a.cout + threadIdx.x, cta_sum[1][threadIdx.x]);
        }
    }
    if (live && vec_out) {
        float* yr = a.y + (size_t)b * a.y_bstride + (size_t)n * a.y_ld;
#pragma unroll
        for (int j = 0; j < kPwSmallMaxCout; j += 4)
            if (j < a.cout) *reinterpret_cast<float4*>(yr + j) = make_float4(outv[j], outv[j + 1], outv[j + 2], outv[j + 3]);
    }
    if (a.bn.y) {
        // fused train-mode BatchNorm (cooperative launch): outv holds this row's conv outputs z
        __threadfence();
        cooperative_groups::this_grid().sync();
        float* yr = a.bn.y + (size_t)m * a.cout;
#pragma unroll
        for (int j = 0; j < kPwSmallMaxCout; ++j) {
            if (j >= a.cout) break;
            float ca, cc;
            pw_bn_coeffs(a, j, M, blockIdx.x == 0 && threadIdx.x == 0, ca, cc);
            outv[j] = apply_act(fmaf(outv[j], ca, cc), a.bn.act, a.bn.slope);
            if (live && (a.cout & 3) != 0) yr[j] = outv[j];
        }
        if (live && (a.cout & 3) == 0) {
#pragma unroll
            for (int j = 0; j < kPwSmallMaxCout; j += 4)
                if (j < a.cout) *reinterpret_cast<float4*>(yr + j) = make_float4(outv[j], outv[j + 1], outv[j + 2], outv[j + 3]);
        }
        if (blockIdx.x == 0 && threadIdx.x == 0 && a.bn.num_batches) *a.bn.num_batches += 1;
    }
}

}  // namespace r3d

using namespace r3d;

extern "C" int r3d_pointwise_stats(const float* xa, long long xa_bstride, int ca, const int32_t* gidx,
                                   long long gidx_bstride, const float* xb, long long xb_bstride, int cb,
                                   const float* wT, const float* scale, const float* shift, int act, float slope,
                                   float* y, long long y_bstride, int y_ld, int cout, int B, int n, int transpose_out,
                                   double* stats, int w_out_in, r3d_stream_t stream);


// ------------------------------------------------------------------------------------ streaming rows kernel
// Dense narrow layers on many rows (the level-0 layers of a large batch: 2.6 M rows of 8..64 channels): HBM-streaming.
// The GEMM kernels above stage 16-channel chunks behind barriers and the one-thread-per-row kernel reads rows with a
// 32..256-byte stride between lanes; both sit at 1.2-1.5 TB/s on these shapes.  Here a CTA of 128 threads takes blocks of
// 128 rows: the block is read with fully coalesced float4 loads into a padded shared-memory tile, thread t computes row
// t from the tile (own row: conflict-free LDS.128; weights: broadcast LDS.128), the outputs go back through a second
// tile and leave with coalesced float4 stores; the batch statistics of a train-mode BatchNorm are column sums of that
// tile, kept per channel in fp64 registers across the CTA's blocks and flushed with one atomic per channel and CTA.
template <int CIN, int COUT>
struct PwRowsSmem {
    static constexpr int XLD = CIN + 4, YLD = COUT + 4;
    static constexpr size_t BYTES = (size_t)(128 * XLD + 128 * YLD + CIN * COUT + 2 * COUT) * sizeof(float);
};

template <int CIN, int COUT>
__global__ void __launch_bounds__(128) pw_rows_kernel(PwArgs a) {
    using S = PwRowsSmem<CIN, COUT>;
    extern __shared__ __align__(16) float sm[];
    float* Xs = sm;                              // [128][XLD]
    float* Ys = Xs + 128 * S::XLD;               // [128][YLD]
    float* Ws = Ys + 128 * S::YLD;               // [CIN][COUT]
    float* Sc = Ws + CIN * COUT;                 // scale, shift
    const int tid = threadIdx.x;
    for (int i = tid; i < CIN * COUT; i += 128)
        Ws[i] = a.w_out_in ? a.wT[(size_t)(i % COUT) * CIN + i / COUT] : a.wT[i];
    for (int i = tid; i < COUT; i += 128) {
        Sc[i] = a.scale ? a.scale[i] : 1.f;
        Sc[COUT + i] = a.shift ? a.shift[i] : 0.f;
    }
    __syncthreads();
    const long long M = (long long)a.B * a.n;
    const long long nblk = (M + 127) / 128;
    double s1 = 0.0, s2 = 0.0;                   // thread c < COUT: statistics of channel c
    for (long long blk = blockIdx.x; blk < nblk; blk += gridDim.x) {
        const long long m0 = blk * 128;
        const int rows = (int)((M - m0 < 128) ? (M - m0) : 128);
        // ---- coalesced load of the block
        const float4* src = reinterpret_cast<const float4*>(a.xa + m0 * CIN);
#pragma unroll
        for (int k = 0; k < CIN / 4; ++k) {
            const int i = tid + k * 128;
            const int r = i / (CIN / 4), c4 = i % (CIN / 4);
            const float4 v = (r < rows) ? src[i] : make_float4(0.f, 0.f, 0.f, 0.f);
            *reinterpret_cast<float4*>(Xs + r * S::XLD + c4 * 4) = v;
        }
        __syncthreads();
        // ---- row tid
        float acc[COUT];
#pragma unroll
        for (int j = 0; j < COUT; ++j) acc[j] = 0.f;
        const float* xr = Xs + tid * S::XLD;
#pragma unroll
        for (int c = 0; c < CIN; c += 4) {
            const float4 x4 = *reinterpret_cast<const float4*>(xr + c);
            const float xv[4] = {x4.x, x4.y, x4.z, x4.w};
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int j = 0; j < COUT; j += 4) {
                    const float4 w = *reinterpret_cast<const float4*>(Ws + (c + u) * COUT + j);
                    acc[j] = fmaf(xv[u], w.x, acc[j]);
                    acc[j + 1] = fmaf(xv[u], w.y, acc[j + 1]);
                    acc[j + 2] = fmaf(xv[u], w.z, acc[j + 2]);
                    acc[j + 3] = fmaf(xv[u], w.w, acc[j + 3]);
                }
        }
        float* yr = Ys + tid * S::YLD;
#pragma unroll
        for (int j = 0; j < COUT; j += 4) {
            float4 o;
            o.x = apply_act(fmaf(acc[j], Sc[j], Sc[COUT + j]), a.act, a.slope);
            o.y = apply_act(fmaf(acc[j + 1], Sc[j + 1], Sc[COUT + j + 1]), a.act, a.slope);
            o.z = apply_act(fmaf(acc[j + 2], Sc[j + 2], Sc[COUT + j + 2]), a.act, a.slope);
            o.w = apply_act(fmaf(acc[j + 3], Sc[j + 3], Sc[COUT + j + 3]), a.act, a.slope);
            if (tid >= rows) o = make_float4(0.f, 0.f, 0.f, 0.f);          // rows past the end count for nothing
            *reinterpret_cast<float4*>(yr + j) = o;
        }
        __syncthreads();
        // ---- coalesced store, column sums
        float4* dst = reinterpret_cast<float4*>(a.y + m0 * COUT);
#pragma unroll
        for (int k = 0; k < COUT / 4; ++k) {
            const int i = tid + k * 128;
            const int r = i / (COUT / 4), c4 = i % (COUT / 4);
            if (r < rows) dst[i] = *reinterpret_cast<const float4*>(Ys + r * S::YLD + c4 * 4);
        }
        if (a.stats && tid < COUT) {
            float p1 = 0.f, p2 = 0.f;
#pragma unroll 8
            for (int r = 0; r < 128; ++r) {
                const float v = Ys[r * S::YLD + tid];
                p1 += v;
                p2 = fmaf(v, v, p2);
            }
            s1 += (double)p1;
            s2 += (double)p2;
        }
        __syncthreads();
    }
    if (a.stats && tid < COUT) {
        atomicAdd(a.stats + tid, s1);
        atomicAdd(a.stats + COUT + tid, s2);
    }
}

static bool pw_rows_eligible(const PwArgs& a) {
    const long long M = (long long)a.B * a.n;
    auto ok = [](int c) { return c == 8 || c == 16 || c == 32 || c == 64; };
    return M >= 131072 && a.cb == 0 && !a.gidx && !a.transpose_out && !a.bn.y && ok(a.ca) && ok(a.cout) &&
           a.ca * a.cout <= 2048 && a.y_ld == a.cout && a.xa_bstride == (long long)a.n * a.ca &&
           a.y_bstride == (long long)a.n * a.cout;
}

template <int CIN, int COUT>
static int pw_rows_launch(const PwArgs& a, cudaStream_t st) {
    auto kern = pw_rows_kernel<CIN, COUT>;
    constexpr size_t smem = PwRowsSmem<CIN, COUT>::BYTES;
    R3D_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const long long nblk = ((long long)a.B * a.n + 127) / 128;
    const long long grid = nblk < (long long)kNumSMs * 4 ? nblk : (long long)kNumSMs * 4;
    kern<<<(unsigned)grid, 128, smem, st>>>(a);
    R3D_LAUNCH_CHECK("pw_rows_kernel");
    return R3D_OK;
}

static int pw_rows_run(const PwArgs& a, cudaStream_t st) {
#define R3D_ROWS(CI, CO) \
    if (a.ca == CI && a.cout == CO) return pw_rows_launch<CI, CO>(a, st);
    R3D_ROWS(8, 8) R3D_ROWS(8, 16) R3D_ROWS(8, 32) R3D_ROWS(8, 64) R3D_ROWS(16, 8) R3D_ROWS(16, 16) R3D_ROWS(16, 32)
    R3D_ROWS(16, 64) R3D_ROWS(32, 8) R3D_ROWS(32, 16) R3D_ROWS(32, 32) R3D_ROWS(32, 64) R3D_ROWS(64, 8) R3D_ROWS(64, 16)
    R3D_ROWS(64, 32)
#undef R3D_ROWS
    return R3D_EUNSUPPORTED;
}

// ------------------------------------------------------------------------------------- launch helpers
// Launches a per-point kernel.  When the fused BatchNorm tail is wanted (a.bn.y) the launch must be cooperative (grid
// barrier) and the whole grid co-resident; otherwise a.bn.y is cleared and the caller runs r3d_bn_apply afterwards.
template <typename Kern>
static int pw_launch(Kern kern, dim3 grid, int threads, PwArgs& a, cudaStream_t st, bool* fused) {
    if (a.bn.y) {
        static std::mutex mu;
        static std::unordered_map<const void*, int> per_sm_cache;
        int per_sm = 0;
        {
            std::lock_guard<std::mutex> lock(mu);
            auto it = per_sm_cache.find(reinterpret_cast<const void*>(kern));
            if (it == per_sm_cache.end()) {
                if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, 0) != cudaSuccess) per_sm = 0;
                per_sm_cache[reinterpret_cast<const void*>(kern)] = per_sm;
            } else {
                per_sm = it->second;
            }
        }
        if ((long long)grid.x * grid.y * grid.z <= (long long)per_sm * kNumSMs) {
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = grid;
            cfg.blockDim = dim3(threads);
            cfg.dynamicSmemBytes = 0;
            cfg.stream = st;
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeCooperative;
            attr[0].val.cooperative = 1;
            cfg.attrs = attr;
            cfg.numAttrs = 1;
            R3D_CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, a));
            if (fused) *fused = true;
            return R3D_OK;
        }
        a.bn.y = nullptr;
    }
    kern<<<grid, threads, 0, st>>>(a);
    return R3D_OK;
}

// ------------------------------------------------------------------------------------------ pw_expand
// C_in <= 4 into a wider layer over very many rows (the input gradient of the class-logits layer, 2 -> 32 at 2.6 M
// rows): a write stream.  A thread owns 4 output channels of one row — the row's inputs are a broadcast load for the
// cout / 4 adjacent threads, stores are 16 bytes and contiguous across the warp.  (The GEMM tile kernel ran this shape
// at 0.85 TB/s.)
__global__ void __launch_bounds__(256) pw_expand_kernel(PwArgs a) {
    const int c4 = a.cout / 4;
    const int cin = a.ca;
    const long long total = (long long)a.B * a.n * c4;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const long long row = t / c4;
        const int c = (int)(t % c4) * 4;
        const int b = (int)(row / a.n), n = (int)(row % a.n);
        const float* x = a.xa + b * a.xa_bstride + (long long)n * cin;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int i = 0; i < cin; ++i) {
            const float v = x[i];
            float4 w;
            if (a.w_out_in)
                w = make_float4(a.wT[(c + 0) * cin + i], a.wT[(c + 1) * cin + i], a.wT[(c + 2) * cin + i], a.wT[(c + 3) * cin + i]);
            else
                w = *reinterpret_cast<const float4*>(a.wT + (long long)i * a.cout + c);
            acc.x = fmaf(v, w.x, acc.x), acc.y = fmaf(v, w.y, acc.y), acc.z = fmaf(v, w.z, acc.z), acc.w = fmaf(v, w.w, acc.w);
        }
        float o[4] = {acc.x, acc.y, acc.z, acc.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float v = o[j];
            if (a.scale) v *= a.scale[c + j];
            if (a.shift) v += a.shift[c + j];
            if (a.act == 1) v = fmaxf(v, 0.f);
            if (a.act == 2) v = v > 0.f ? v : v * a.slope;
            o[j] = v;
        }
        *reinterpret_cast<float4*>(a.y + b * a.y_bstride + (long long)n * a.y_ld + c) = make_float4(o[0], o[1], o[2], o[3]);
    }
}

static bool pw_expand_eligible(const PwArgs& a) {
    const long long M = (long long)a.B * a.n;
    return M >= 131072 && a.ca <= 4 && a.cb == 0 && !a.gidx && !a.transpose_out && !a.stats && !a.bn.y && a.cout % 4 == 0 &&
           a.cout > kPwSmallMaxCout && a.y_ld % 4 == 0 && a.y_bstride % 4 == 0 && is_aligned(a.y, 16) &&
           is_aligned(a.wT, 16);
}

static int pw_run(PwArgs a, cudaStream_t st, bool* fused) {
    const int ca = a.ca, cb = a.cb, cout = a.cout;
    const long long M = (long long)a.B * a.n;
    if (fused) *fused = false;
    if (pw_rows_eligible(a)) return pw_rows_run(a, st);
    if (pw_expand_eligible(a)) {
        long long blocks = (M * (cout / 4) + 255) / 256;
        if (blocks > (long long)kNumSMs * 16) blocks = (long long)kNumSMs * 16;
        pw_expand_kernel<<<(unsigned)blocks, 256, 0, st>>>(a);
        R3D_LAUNCH_CHECK("pw_expand_kernel");
        return R3D_OK;
    }
    if (cout <= kPwSmallMaxCout && (long long)(ca + cb) * cout <= kPwSmallMaxW) {
        int rc = pw_launch(pw_small_kernel, dim3((unsigned)((M + 255) / 256)), 256, a, st, fused);
        if (rc != R3D_OK) return rc;
        R3D_LAUNCH_CHECK("pw_small_kernel");
        return R3D_OK;
    }
    const bool aligned = (ca % 4 == 0) && (cb % 4 == 0);
    if (g_pw_tensor_cores.load() && pw_tc_eligible(a, g_pw_tensor_cores.load() == 2)) {
        a.bn.y = nullptr;
        return pw_tc_launch(a, st);
    }
    int rc = R3D_OK;
    if (!aligned) {
        a.bn.y = nullptr;
        dim3 grid((unsigned)((M + 127) / 128), (cout + 63) / 64);
        pw_gemm_kernel<128, 64><<<grid, 128, 0, st>>>(a);
    } else if (M <= 2048 && ca + cb >= 128) {
        // a few hundred rows of a wide layer (the bottom of the encoder / decoder of a small cloud): the kernel is a
        // chain of load -> barrier -> FMA steps on a handful of CTAs, each step exposed to the full L2 latency.
        // 32-channel steps halve the chain, 32-row tiles put 2-4x more CTAs on the machine.
        const long long tiles32 = ((M + 31) / 32) * ((cout + 31) / 32);
        if (tiles32 <= 2 * kNumSMs && ca + cb >= 256) {
            // long contraction on few tiles: clusters of 4 CTAs split the input channels (see SPLITK above)
            a.bn.y = nullptr;
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3((unsigned)((M + 31) / 32), (cout + 31) / 32, 4);
            cfg.blockDim = dim3(64);
            cfg.dynamicSmemBytes = 0;
            cfg.stream = st;
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeClusterDimension;
            attr[0].val.clusterDim.x = 1;
            attr[0].val.clusterDim.y = 1;
            attr[0].val.clusterDim.z = 4;
            cfg.attrs = attr;
            cfg.numAttrs = 1;
            R3D_CUDA_TRY(cudaLaunchKernelEx(&cfg, pw_gemm_fast_kernel<32, 32, 4, 32, 4>, a));
        } else if (tiles32 <= 2 * kNumSMs) {
            rc = pw_launch(pw_gemm_fast_kernel<32, 32, 4, 32>, dim3((unsigned)((M + 31) / 32), (cout + 31) / 32), 64, a, st,
                           fused);
        } else {
            rc = pw_launch(pw_gemm_fast_kernel<32, 64, 4, 32>, dim3((unsigned)((M + 31) / 32), (cout + 63) / 64), 128, a, st,
                           fused);
        }
    } else if (M <= 64 * 4 * kNumSMs / ((cout + 63) / 64)) {
        // few rows: 64x64 tiles of 256 threads (4x4 per thread) put more CTAs and warps on the machine
        rc = pw_launch(pw_gemm_fast_kernel<64, 64, 4>, dim3((unsigned)((M + 63) / 64), (cout + 63) / 64), 256, a, st, fused);
    } else if (cout <= 32) {
        rc = pw_launch(pw_gemm_fast_kernel<256, 32, 8>, dim3((unsigned)((M + 255) / 256), (cout + 31) / 32), 128, a, st,
                       fused);
    } else if (cout <= 64 || M < 64 * 148) {
        rc = pw_launch(pw_gemm_fast_kernel<128, 64, 8>, dim3((unsigned)((M + 127) / 128), (cout + 63) / 64), 128, a, st,
                       fused);
    } else {
        rc = pw_launch(pw_gemm_fast_kernel<64, 128, 8>, dim3((unsigned)((M + 63) / 64), (cout + 127) / 128), 128, a, st,
                       fused);
    }
    if (rc != R3D_OK) return rc;
    R3D_LAUNCH_CHECK("pw_gemm_kernel");
    return R3D_OK;
}

extern "C" int r3d_pointwise(const float* xa, long long xa_bstride, int ca, const int32_t* gidx,
                             long long gidx_bstride, const float* xb, long long xb_bstride, int cb, const float* wT,
                             const float* scale, const float* shift, int act, float slope, float* y,
                             long long y_bstride, int y_ld, int cout, int B, int n, int transpose_out,
                             r3d_stream_t stream) {
    return r3d_pointwise_stats(xa, xa_bstride, ca, gidx, gidx_bstride, xb, xb_bstride, cb, wT, scale, shift, act, slope,
                               y, y_bstride, y_ld, cout, B, n, transpose_out, nullptr, 0, stream);
}

extern "C" int r3d_pointwise_stats(const float* xa, long long xa_bstride, int ca, const int32_t* gidx,
                                   long long gidx_bstride, const float* xb, long long xb_bstride, int cb,
                                   const float* wT, const float* scale, const float* shift, int act, float slope,
                                   float* y, long long y_bstride, int y_ld, int cout, int B, int n, int transpose_out,
                                   double* stats, int w_out_in, r3d_stream_t stream) {
    if (B < 0 || n < 0 || ca <= 0 || cb < 0 || cout <= 0 || act < 0 || act > 2) return R3D_EINVAL;
    if (B == 0 || n == 0) return R3D_OK;
    if (!xa || !wT || !y || (cb > 0 && !xb)) return R3D_EINVAL;
    if (y_ld == 0) y_ld = cout;
    if (y_ld < cout) return R3D_EINVAL;
    if (xa_bstride == 0) xa_bstride = (long long)n * ca;
    if (xb_bstride == 0) xb_bstride = (long long)n * cb;
    if (y_bstride == 0) y_bstride = transpose_out ? (long long)cout * n : (long long)n * y_ld;
    if (!is_aligned(xa, 16) || (xb && !is_aligned(xb, 16)) || !is_aligned(wT, 16) || !is_aligned(y, 16))
        return R3D_EALIGN;
    // vector loads need every source row 16-byte aligned
    if ((ca % 4 == 0 && (xa_bstride % 4)) || (cb > 0 && cb % 4 == 0 && (xb_bstride % 4))) return R3D_EALIGN;
    PwArgs a{xa, xa_bstride, ca, gidx, gidx_bstride, xb, xb_bstride, cb, wT, scale, shift, act, slope,
             y, y_bstride, y_ld, cout, B, n, transpose_out, stats, w_out_in};
    return pw_run(a, static_cast<cudaStream_t>(stream), nullptr);
}

// Train-mode SharedMLP forward on dense rows: z = W x with batch statistics, then y = act(BatchNorm_batch(z)).  One
// cooperative launch (grid barrier between the statistics and the normalisation, outputs still in registers) when
// the layer's grid is co-resident and its kernel has the fused tail; r3d_pointwise_stats + r3d_bn_apply otherwise.
extern "C" int r3d_pointwise_bn(const float* x, long long M, int cin, const float* w, int cout, double* stats,
                                const float* gamma, const float* beta, const float* bias, float eps, float momentum,
                                float* running_mean, float* running_var, long long* num_batches, int act,
                                float slope, float* z, float* y, float* save, r3d_stream_t stream) {
    if (M < 0 || M > 0x7fffffffLL || cin <= 0 || cout <= 0 || act < 0 || act > 2) return R3D_EINVAL;
    if (M == 0) return R3D_OK;
    if (!x || !w || !stats || !gamma || !beta || !z || !y || !save) return R3D_EINVAL;
    if (!is_aligned(x, 16) || !is_aligned(w, 16) || !is_aligned(z, 16) || !is_aligned(y, 16)) return R3D_EALIGN;
    PwArgs a{x, (long long)M * cin, cin, nullptr, 0, nullptr, 0, 0, w, nullptr, nullptr, 0, 0.f,
             z, (long long)M * cout, cout, cout, 1, (int)M, 0, stats, 1};
    a.bn = PwBnArgs{y, gamma, beta, bias, eps, momentum, running_mean, running_var, num_batches, save, act, slope};
    if (!(bn_fused_mask() & 1)) a.bn.y = nullptr;
    bool fused = false;
    const int rc = pw_run(a, static_cast<cudaStream_t>(stream), &fused);
    if (rc != R3D_OK || fused) return rc;
    return r3d_bn_apply(z, stats, M, cout, gamma, beta, bias, eps, momentum, running_mean, running_var, num_batches, act,
                        slope, y, save, stream);
}

// Which kernel r3d_pointwise runs for a layer under the current settings: 0 pw_small_kernel (thin layers, weights in
// shared memory), 1 pw_gemm_kernel (channel counts not multiples of 4), 2 pw_gemm_fast_kernel (pipelined FP32 GEMM,
// any tile shape), 3 pw_tc_kernel (tcgen05 3xTF32).  Dense layout assumed (y_ld = cout).
extern "C" int r3d_pointwise_plan(int ca, int cb, int cout, long long rows, int transpose_out) {
    {
        PwArgs r{};
        r.ca = ca;
        r.cb = cb;
        r.cout = cout;
        r.B = 1;
        r.n = (int)(rows > 0x7fffffffLL ? 0x7fffffffLL : rows);
        r.transpose_out = transpose_out;
        r.y_ld = cout;
        r.xa_bstride = (long long)r.n * ca;
        r.y_bstride = (long long)r.n * cout;
        if (pw_rows_eligible(r)) return 4;
        r.y = reinterpret_cast<float*>(16);      // alignment checks only
        r.wT = reinterpret_cast<const float*>(16);
        if (pw_expand_eligible(r)) return 5;
    }
    if (cout <= kPwSmallMaxCout && (long long)(ca + cb) * cout <= kPwSmallMaxW) return 0;
    PwArgs a{};
    a.ca = ca;
    a.cb = cb;
    a.cout = cout;
    a.B = 1;
    a.n = (int)(rows > 0x7fffffffLL ? 0x7fffffffLL : rows);
    a.transpose_out = transpose_out;
    a.y_ld = cout;
    a.y_bstride = 0;
    if (g_pw_tensor_cores.load() && pw_tc_eligible(a, g_pw_tensor_cores.load() == 2)) return 3;
    return ((ca % 4 == 0) && (cb % 4 == 0)) ? 2 : 1;
}

// 1 (default): the layers the tcgen05 3xTF32 kernel is measured to win on (pw_tc_eligible) run there; 0: FP32
// CUDA-core kernels everywhere; 2: every layer the tensor-core kernel supports (C_in >= 32, C_out a multiple of 32,
// >= 4096 rows) -- benchmarks and tests.  Returns the previous value.
extern "C" int r3d_pointwise_set_tensor_cores(int on) {
    if (on < 0 || on > 2) return g_pw_tensor_cores.load();
    return g_pw_tensor_cores.exchange(on);
}
