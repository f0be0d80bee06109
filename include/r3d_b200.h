/*
 * r3d_b200.h — C ABI of the B200-native RandLA-Net hot path (3d_recognizer drop-in).
 *
 * One shared library, lib/libr3d_b200.so, built from 3d_recognizer_b200/csrc/ for sm_100a.
 * Plain pointers and sizes only; no torch / pybind types.  The reference's FFI for this path
 * is the pybind11 module `knn_tpk` (randlanet/utils/src/bindings.cpp:5-7); everything else on the
 * path is Python calling torch ops (randlanet/utils/modules.py), so the remaining entry points
 * mirror the reference's operator boundaries one to one (file:line cited per function).
 *
 * Conventions (all functions):
 *   - return 0 on success or a negative R3D_E* code; r3d_error_string() names it.  The Python
 *     host (3d_recognizer_b200/_cabi.py) maps codes to the reference's exception types.
 *   - "*_host" entry points take HOST buffers, own their device memory and synchronise: they are
 *     the drop-in for a CPU-tensor FFI call such as knn_tpk.knn.
 *   - every other entry point takes DEVICE pointers on the current device, never allocates,
 *     frees or synchronises, and launches on `stream` (a cudaStream_t passed as void*).  The
 *     caller allocates outputs and the workspace (size from the matching *_workspace_bytes).
 *   - tensors are dense row-major with the layouts written beside each argument; float = fp32.
 *   - no global mutable state: safe from several host threads and from a spawned process
 *     (reference: train.py:108-115 trains in a spawned child, main.py:71-89 predicts on the Tk thread).
 */
#ifndef R3D_B200_H
#define R3D_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define R3D_ABI_VERSION 1

#define R3D_OK 0
#define R3D_EINVAL (-1)        /* null pointer / negative size / bad enum          -> ValueError   */
#define R3D_ENOT_ENOUGH (-2)   /* Ns < K  (knn.cpp:15-17 TORCH_CHECK)               -> RuntimeError */
#define R3D_EKMAX (-3)         /* K larger than the compiled maximum (R3D_KNN_KMAX) -> ValueError   */
#define R3D_EALIGN (-4)        /* pointer not aligned as documented                 -> ValueError   */
#define R3D_EWORKSPACE (-5)    /* workspace too small                               -> ValueError   */
#define R3D_ECUDA (-6)         /* CUDA runtime error (see r3d_last_cuda_error)      -> RuntimeError */
#define R3D_EUNSUPPORTED (-7)  /* shape outside what the kernels are built for      -> ValueError   */

#define R3D_KNN_KMAX 64

typedef void* r3d_stream_t; /* cudaStream_t */

int r3d_abi_version(void);
const char* r3d_error_string(int code);
/* text of the last CUDA error seen by THIS thread inside the library ("" if none) */
const char* r3d_last_cuda_error(void);
/* number of kernels this library has launched in this process so far (bench.py's gpu_launches) */
unsigned long long r3d_launch_count(void);

/* ------------------------------------------------------------------------------------------ KNN
 * Replaces knn_tpk.knn(support, querry, k) (bindings.cpp:5-7 -> knn.cpp:43-61 -> neighbors.h:281-322
 * -> nanoflann.hpp:1367) and folds KNN.forward's sqrt (modules.py:149).
 * Exact search.  d2 = fl(fl(fl(dx*dx)+fl(dy*dy))+fl(dz*dz)) with d = query - support
 * (nanoflann.hpp:488-497 compiled without FMA); neighbours ordered by (d2, index) ascending, i.e.
 * exact ties go to the LOWER support index, also at the K-th boundary.  Coordinates must be finite.
 *
 *   support (B,Ns,3)  query (B,Nq,3)           fp32; query may alias support (self search).
 *       *_batch_stride: floats between consecutive clouds (0 = dense, i.e. N*3).  A stride larger
 *       than N*3 addresses the first N points of longer clouds — the "random down-sampling = prefix
 *       of the permuted cloud" views of RandLANet.forward (modules.py:586-589, :596-597) — without a copy.
 *   idx64   (B,Nq,K)  int64   nullable        (the reference's index dtype)
 *   idx32   (B,Nq,K)  int32   nullable        (what the fused kernels below consume)
 *   dist    (B,Nq,K)  fp32    nullable        sqrt(d2)   — what KNN.forward returns
 *   dist_sq (B,Nq,K)  fp32    nullable        d2         — what knn_tpk.knn returns
 * Two back-ends with identical results: the tiled brute-force scan (TMA-staged support tiles) and, for
 * large clouds, a uniform-grid search (counting sort into cubic cells + ring walk per query).
 * workspace: r3d_knn_workspace_bytes(), 256-byte aligned.
 */
size_t r3d_knn_workspace_bytes(int B, int Ns, int Nq, int K);
int r3d_knn(const float* support, long long support_batch_stride, const float* query,
            long long query_batch_stride, int B, int Ns, int Nq, int K,
            int64_t* idx64, int32_t* idx32, float* dist, float* dist_sq,
            void* workspace, size_t workspace_bytes, r3d_stream_t stream);
/* host-buffer drop-in for knn_tpk.knn: (idx int64, d2 fp32), both (B,Nq,K), caller-allocated, pageable or page-locked
 * (into page-locked outputs the results arrive by direct DMA, chunk by chunk, while the next chunk is searched).  Uses
 * per-thread non-blocking streams and a cached grow-only device arena: no allocation per call after the first, never the
 * legacy default stream, no device-wide synchronisation. */
int r3d_knn_host(const float* support, const float* query, int B, int Ns, int Nq, int K,
                 int64_t* idx64, float* dist_sq);
/* tuning hook for benchmarks/tests: 0 exact scalar, 1 FMA-prefilter scalar, 2 FMA-prefilter
 * packed f32x2 (default).  All variants return identical results.  Returns the previous value. */
int r3d_knn_set_variant(int variant);
/* search algorithm: 0 auto (see r3d_knn_plan), 1 tiled brute force, 2 uniform grid (one warp walks the rings of a
 * query), 4 uniform grid with the one-thread-per-query walk.  All back-ends return identical results.  Returns the
 * previous value. */
int r3d_knn_set_algorithm(int algorithm);
/* the back-end r3d_knn runs for this shape under the current setting: 1 tiled brute force, 2 uniform grid,
 * 3 warp-per-query register scan (auto mode only: small supports, Ns < 2048) */
int r3d_knn_plan(int B, int Ns, int Nq, int K);
/* tuning hook: average number of support points per grid cell of the uniform-grid search (0 = built-in default) */
int r3d_knn_set_grid_density(float points_per_cell);

/* ------------------------------------------------------- fused LocSE + attentive pooling (one LFA half)
 * Replaces, for one half of LocalFeatureAggregation.forward (modules.py:316-319 = stage 1, :321-323 =
 * stage 2), the op sequence RelativePositionEncoding (modules.py:170-186) -> mlp_rpe1 (:317)
 * [-> mlp_rpe2 (:321), stage 2] -> PointFeatureAugmentation gather + concat (:209-221) ->
 * AttentivePooling score Linear + softmax over K + weighted sum (:246-252).  The pooling MLP (:253) is
 * a per-point layer: r3d_pointwise.  Nothing of shape (B,C,N,K) is written to memory.
 *
 *   xyz      (B,N,3)   fp32, clouds xyz_bstride floats apart (0 = dense)
 *   idx      (B,N,K)   int32 neighbour indices (r3d_knn idx32)
 *   feat     (B,N,h)   fp32 rows gathered at the neighbours (h = d/2): mlp1 output (stage 1) or the
 *                      pool-1 output (stage 2); clouds feat_bstride floats apart (0 = dense)
 *   w_rpe1   (h,10)    mlp_rpe1 conv weight, [out][in]
 *   a_rpe1,b_rpe1 (h)  per-channel affine applied after it: r1 = relu(a*(W rpe)+b)  (BatchNorm + conv
 *                      bias folded by the host: running stats in eval mode, batch stats in train mode)
 *   w_rpe2T  (h,h)     mlp_rpe2 weight TRANSPOSED, [in][out]; a_rpe2,b_rpe2 (h)     (stage 2 only)
 *   w_scoreT (d,d)     score Linear weight TRANSPOSED, [in][out]
 *   pooled   (B,N,d)   out: sum_k softmax_k(s)[k,c] * x[k,c],  x = [r ; feat[idx]]
 * Supported: d in {16,32,64,128,256}, K in {16,32}; other shapes -> R3D_EUNSUPPORTED.
 * feat, pooled, w_rpe2T, w_scoreT 16-byte aligned. */
int r3d_lfa_pool(int stage, const float* xyz, long long xyz_bstride, const int32_t* idx, const float* feat,
                 long long feat_bstride, const float* w_rpe1, const float* a_rpe1, const float* b_rpe1,
                 const float* w_rpe2T, const float* a_rpe2, const float* b_rpe2, const float* w_scoreT,
                 float* pooled, int B, int N, int K, int d, r3d_stream_t stream);

/* r3d_lfa_pool on the tcgen05 tensor cores (csrc/lfa_cl.cu, "channel-lane" kernel): the score GEMM (and mlp_rpe2) run
 * as kind::f16 MMAs on split-fp16 operands (hi/lo halves of power-of-two scaled values, three products per K step,
 * fp32 accumulation in TMEM: fp32-level accuracy) in transposed form — TMEM lane = channel, columns = rows — so that
 * the softmax over K and the weighted sum are in-thread; persistent warp-specialised CTAs (worker groups + one MMA
 * warp).  Same operator and arguments as r3d_lfa_pool, except that the weights come in their stored [out][in] layout:
 * w_rpe2 (h,h), w_score (d,d).  status (device int, nullable): bit 0 is OR-ed in when an activation left the range the
 * fixed operand scale covers (|x| >= 4094; never seen behind a BatchNorm) — results are then invalid and the host
 * raises.  Supported: d in {16,32,64,128}, K in {16,32}; else R3D_EUNSUPPORTED. */
int r3d_lfa_pool_tc(int stage, const float* xyz, long long xyz_bstride, const int32_t* idx, const float* feat,
                    long long feat_bstride, const float* w_rpe1, const float* a_rpe1, const float* b_rpe1,
                    const float* w_rpe2, const float* a_rpe2, const float* b_rpe2, const float* w_score,
                    float* pooled, int* status, int B, int N, int K, int d, r3d_stream_t stream);

/* Training-mode companions of r3d_lfa_pool_tc on the tensor cores (csrc/lfa_cl_bwd.cu; same channel-lane design,
 * weight-gradient-like sums accumulated in TMEM across the tiles of a persistent CTA).  One entry point, four modes;
 * arguments a mode does not use may be NULL.  Weights in their stored [out][in] layouts; outputs ACCUMULATED into
 * caller-zeroed buffers exactly as documented for the FP32 entry points below:
 *   mode 1  = r3d_lfa_pool_bwd(stage 1):        dfeat, dw_score, g1            (needs scal[0])
 *   mode 2  = r3d_lfa_pool2_bwd_train (pass 1): dfeat, dw_score, du2_tiles, sum_du2; scal[1] <- max |du2| (atomic max)
 *   mode 3  = r3d_lfa_bn2_bwd (pass 2):         g1, dw2 from du2_tiles, bn2    (needs scal[1])
 *   mode 4  = r3d_lfa_moments(mode 1):          m_r1, s_r1
 * du2_tiles: r3d_lfa_tc_du2_floats(B,N,K,d) floats, 16-byte aligned, layout private to modes 2/3.
 * scal (2 floats, device): [0] = max |dpooled| written by r3d_absmax before a mode 1/2 launch (the power-of-two scale of
 * the gradient-side fp16x2 operands is derived from it on the device; [1] zero-filled before mode 2).
 * status as for r3d_lfa_pool_tc.  Supported: d in {16,32,64,128}, K in {16,32}; else R3D_EUNSUPPORTED. */
long long r3d_lfa_tc_du2_floats(int B, int N, int K, int d);
int r3d_absmax(const float* x, long long n, float* out, r3d_stream_t stream);
int r3d_lfa_tc_bwd(int mode, const float* xyz, long long xyz_bstride, const int32_t* idx, const float* feat,
                   long long feat_bstride, const float* w_rpe1, const float* a_rpe1, const float* b_rpe1,
                   const float* w_rpe2, const float* a_rpe2, const float* b_rpe2, const float* w_score,
                   const float* dpooled, float* dfeat, long long dfeat_bstride, float* dw_score, double* g1,
                   float* du2_tiles, double* sum_du2, const float* bn2, double* dw2, double* m_r1, double* s_r1,
                   float* scal, int* status, int B, int N, int K, int d, r3d_stream_t stream);

/* The widest level, d = 256 (csrc/lfa_cl_wide.cu): a CTA owns one HALF of the score channels (weights of that half
 * resident as split-fp16 planes, dWs half in TMEM), the two halves add their partial dX.  mode 0 forward (pooled), mode 1
 * = r3d_lfa_pool_bwd(stage 1) (dfeat, dw_score, g1), mode 2 = pass 1 of the train-mode stage-2 backward (dfeat,
 * dw_score, sum_du2 and du2_part = the two halves' partial du2, [half][B*N*K][128]; r3d_lfa_du2_combine adds them into
 * the tile layout r3d_lfa_bn2_bwd reads, tile_points = r3d_lfa_tile_points_for).  rmat (B*N*K,128), nullable: the r
 * half of the row operand read from memory instead of evaluating mlp_rpe1 — stage 2 passes r2 = relu(a2 (W2 r1) + c2),
 * built from r3d_lfa_r1_rows (out (B*N*K,h) = r1 rows) and a per-point layer.  scal[0] = max |dpooled| (r3d_absmax).
 * Supported: d = 256, K in {16,32}. */
int r3d_lfa_r1_rows(const float* xyz, long long xyz_bstride, const int32_t* idx, const float* w_rpe1, const float* a_rpe1,
                    const float* b_rpe1, float* out, int B, int N, int K, int h, r3d_stream_t stream);
int r3d_lfa_du2_combine(const float* part, float* out, int B, int N, int K, int h, int tile_points, r3d_stream_t stream);
int r3d_lfa_tc_wide(int mode, const float* xyz, long long xyz_bstride, const int32_t* idx, const float* feat,
                    long long feat_bstride, const float* w_rpe1, const float* a_rpe1, const float* b_rpe1,
                    const float* rmat, const float* w_score, float* pooled, const float* dpooled, float* dfeat,
                    long long dfeat_bstride, float* dw_score, double* g1, float* du2_part, double* sum_du2,
                    const float* scal, int* status, int B, int N, int K, int d, r3d_stream_t stream);

/* Backward of one r3d_lfa_pool launch (autograd of modules.py:316-323 as driven by trainer.py:115-119).
 * Inputs as in the forward plus
 *   w_rpe2s  (h,h)  [out][in] mlp_rpe2 weight with row j scaled by a_rpe2[j]         (stage 2)
 *   w_score  (d,d)  [out][in] score weight (the forward's w_scoreT transposed back)
 *   dpooled  (B,N,d) gradient of the loss with respect to `pooled`
 * Outputs, all ACCUMULATED into (the caller zero-fills them):
 *   dfeat    (B,N,h)  gradient w.r.t. `feat` (scatter-add over neighbour lists, fp32 atomics)
 *   dw_score (d,d)    [out][in]
 * g1, g2m, g2c are fp64 (their sums cancel against the BatchNorm mean/variance terms of the moment path):
 *   g1       (h,16)   cols 0..9 = sum_rows du1 * rpe, col 10 = sum_rows du1, where du1 is the gradient at
 *                     the input of mlp_rpe1's ReLU; the host forms dW1 = a1 (.) g1[:, :10],
 *                     da1 = rowsum(W1 (.) g1[:, :10]), db1 = g1[:,10]
 *   g2m      (h,h)    sum_rows du2 r1^T      (stage 2: dW2 = a2 (.) g2m, da2 = rowsum(W2 (.) g2m))
 *   g2c      (h,16)   col 10 = sum_rows du2  (stage 2: db2) */
int r3d_lfa_pool_bwd(int stage, const float* xyz, long long xyz_bstride, const int32_t* idx, const float* feat,
                     long long feat_bstride, const float* w_rpe1, const float* a_rpe1, const float* b_rpe1,
                     const float* w_rpe2T, const float* a_rpe2, const float* b_rpe2, const float* w_rpe2s,
                     const float* w_scoreT, const float* w_score, const float* dpooled, float* dfeat,
                     long long dfeat_bstride, float* dw_score, double* g1, double* g2m, double* g2c, int B, int N,
                     int K, int d, r3d_stream_t stream);

/* Train-mode backward of a STAGE-2 launch, split in the standard two BatchNorm passes (batch statistics of mlp_rpe2):
 *   r3d_lfa_pool2_bwd_train  pass 1: as r3d_lfa_pool_bwd(stage 2) up to du2 (gradient at mlp_rpe2's ReLU input),
 *                            which is written per CTA tile to du2_tiles ([b][tile][h][P*K], P = r3d_lfa_tile_points_for)
 *                            together with sum_du2 (2,h) fp64 += (sum du2, sum du2 * r2); dfeat, dw_score as before.
 *   r3d_lfa_bn2_bwd          pass 2: with bn2 (5,h) = a2, mean2, rstd2, mean(du2), mean(du2*zhat2):
 *                            dz2 = a2 (du2 - m1 - zhat2 m2), dw2 (h,h) fp64 += dz2^T r1, dr1 = dz2 W2,
 *                            du1 = dr1 [r1 > 0], g1 (h,16) fp64 += du1^T [rpe, 1].
 * Subtracting the mean/variance terms per row (not as row sums) keeps fp32 accuracy at any cloud size. */
int r3d_lfa_tile_points(int K, int d);
/* P for a given batch: small batches of the wide levels (d >= 128, <= 2 waves of default tiles) run half-size tiles */
int r3d_lfa_tile_points_for(int K, int d, int B, int N);
int r3d_lfa_pool2_bwd_train(const float* xyz, long long xyz_bstride, const int32_t* idx, const float* feat,
                            long long feat_bstride, const float* w_rpe1, const float* a_rpe1, const float* b_rpe1,
                            const float* w_rpe2T, const float* a_rpe2, const float* b_rpe2, const float* w_scoreT,
                            const float* w_score, const float* dpooled, float* dfeat, long long dfeat_bstride,
                            float* dw_score, float* du2_tiles, double* sum_du2, int B, int N, int K, int d,
                            r3d_stream_t stream);
int r3d_lfa_bn2_bwd(const float* xyz, long long xyz_bstride, const int32_t* idx, const float* w_rpe1,
                    const float* a_rpe1, const float* b_rpe1, const float* du2_tiles, const float* w_rpe2T,
                    const float* w_rpe2, const float* bn2, double* g1, double* dw2, int B, int N, int K, int d,
                    r3d_stream_t stream);

/* Moments for the train-mode BatchNorm of mlp_rpe1 / mlp_rpe2 (modules.py:86-90 statistics over all
 * B*N*K positions), accumulated in fp64 into caller-zeroed buffers:
 *   mode 0: m_rpe (16,16) += sum_rows e e^T with e = [rpe(10), 1, 0...]: row 10 holds sum rpe, [10][10] the
 *           row count.  mean/var of W1 rpe follow in closed form (W mu, diag(W Cov W^T)).
 *   mode 1: m_r1 (h,h) += sum_rows r1 r1^T and s_r1 (h,16)[:,10] += sum_rows r1, r1 = relu(a1 (W1 rpe) + b1).
 *   mode 2: backward of mode 1: with gsym (h,h) = G + G^T (G = d loss/d m_r1) and gsum (h) = d loss/d sum r1,
 *           g1 (h,16) += du1^T [rpe, 1] as in r3d_lfa_pool_bwd. */
int r3d_lfa_moments(int mode, const float* xyz, long long xyz_bstride, const int32_t* idx, const float* w_rpe1,
                    const float* a_rpe1, const float* b_rpe1, double* m_rpe, double* m_r1, double* s_r1,
                    const float* gsym, const float* gsum, double* g1, int B, int N, int K, int d,
                    r3d_stream_t stream);

/* Train-mode BatchNorm of y = W x (+bias) from the input moments S = sum x (stride s_stride), M = sum x x^T
 * (ldm), R rows (r3d_lfa_moments): a = gamma/sqrt(var+eps), c = beta - a (W mu) with mean = W mu,
 * var = diag(W Cov W^T).  W (cout,cin) [out][in], cin <= 128.  Updates running_mean/var/num_batches (nullable)
 * like BatchNorm2d; save (5,cout) fp64 is scratch for the backward.
 * Backward: ga, gc (cout, fp64) -> dW (cout,cin, fp64), dgamma, dbeta; scal (2,cout) fp64 scratch; when dM (cin,cin) and
 * dS (cin) are given (fp64) they receive the gradient w.r.t. the moments (mlp_rpe2: r1's moments depend on
 * mlp_rpe1's parameters; they feed r3d_lfa_moments mode 2). */
int r3d_bn_from_moments(const float* W, int cout, int cin, const double* S, int s_stride, const double* M, int ldm,
                        double R, const float* gamma, const float* beta, const float* bias, float eps,
                        float momentum, float* running_mean, float* running_var, long long* num_batches,
                        float* a_out, float* c_out, double* save, r3d_stream_t stream);
int r3d_bn_from_moments_bwd(const float* W, int cout, int cin, const double* S, int s_stride, const double* M,
                            int ldm, double R, const float* gamma, const double* save, const double* ga,
                            const double* gc, double* dW, float* dgamma, float* dbeta, double* scal, double* dM,
                            double* dS, r3d_stream_t stream);

/* mlp_rpe1's parameter gradients in one launch (train mode): G (cout, ldg > cin) fp64 = per output channel the sums
 * over all (point, neighbour) rows of du (x) x (columns 0..cin-1) and of du (column cin), as accumulated by
 * r3d_lfa_pool_bwd / r3d_lfa_bn2_bwd of BOTH halves of a block into one g1 buffer; S, M, R, gamma, save as in
 * r3d_bn_from_moments.  dW (cout,cin) = a G + BatchNorm terms, dgamma, dbeta: fp32.  Replaces the tensor-op
 * composition of modules.py:60-104 backward for mlp_rpe1 (autograd of Conv2d + BatchNorm2d over B*N*K rows). */
int r3d_lfa_rpe1_grads(const float* W, int cout, int cin, const double* S, int s_stride, const double* M, int ldm,
                       double R, const float* gamma, const double* save, const double* G, int ldg, float* dW,
                       float* dgamma, float* dbeta, r3d_stream_t stream);
/* Coefficients of r3d_lfa_bn2_bwd from pass 1's batch sums: sums (2,h) fp64 (r3d_lfa_pool2_bwd_train), a2, c2 the
 * forward's fp32 affine, save (5,h) fp64 the forward's statistics (r3d_bn_from_moments), rows = B*N*K.
 * bn2 (5,h) fp32 = a2, mean, rstd, mean du2, mean du2*zhat2;  dgamma = sum du2*zhat2, dbeta = sum du2 (mlp_rpe2's). */
int r3d_lfa_bn2_coeffs(const double* sums, const float* a2, const float* c2, const double* save, double rows, int h,
                       float* bn2, float* dgamma, float* dbeta, r3d_stream_t stream);

/* ------------------------------------------------------------------ row-form LFA block (any K, any width)
 * The fused kernels above are built for d in {16,...,256} and K in {16,32}; the reference takes any n_neighbors and any
 * layer size (modules.py:298-325, 484-500).  Other settings run the same operators in ROW FORM: the (B*N*K, C)
 * neighbourhood rows are materialised, mlp_rpe1/2 and the score Linear run as per-point layers (r3d_pointwise*) over
 * them, and these entry points supply what is not a per-point layer (csrc/lfa_rows.cu):
 *   r3d_lfa_rpe_rows           out (B*N*K,10) = [p_i, p_j, p_i - p_j, |p_i - p_j|]             (modules.py:170-186)
 *   r3d_lfa_gather_concat      out (B*N*K,2h) = [r (B*N*K,h) ; feat[b, idx, :] (h)]             (modules.py:200-208)
 *   r3d_lfa_gather_concat_bwd  dr (B*N*K,h, nullable) = dout[:, :h];  dfeat (B,N,h, nullable) += scatter of dout[:, h:]
 *   r3d_lfa_attn_pool          pooled (points,d) = sum_k softmax_k(S)[k,c] X[k,c]; S, X (points*K,d)  (modules.py:246-252)
 *   r3d_lfa_attn_pool_bwd      dX = g A (direct term only: the score Linear's own backward adds dS Ws), dS = A g (X - pooled)
 * idx int32 (B,N,K); feat/dfeat clouds *_bstride floats apart (0 = dense); K >= 1, h, d >= 1. */
int r3d_lfa_rpe_rows(const float* xyz, long long xyz_bstride, const int32_t* idx, float* out, int B, int N, int K,
                     r3d_stream_t stream);
int r3d_lfa_gather_concat(const float* r, const float* feat, long long feat_bstride, const int32_t* idx, float* out,
                          int B, int N, int K, int h, r3d_stream_t stream);
int r3d_lfa_gather_concat_bwd(const float* dout, const int32_t* idx, float* dr, float* dfeat, long long dfeat_bstride,
                              int B, int N, int K, int h, r3d_stream_t stream);
int r3d_lfa_attn_pool(const float* S, const float* X, float* pooled, long long points, int K, int d, r3d_stream_t stream);
int r3d_lfa_attn_pool_bwd(const float* S, const float* X, const float* dpooled, float* dS, float* dX, long long points,
                          int K, int d, r3d_stream_t stream);

/* ------------------------------------------------------------------ element-wise pieces of the training step
 * r3d_add_lrelu      y = LeakyReLU_slope(a + b): the residual sum that closes an LFA block (modules.py:325, slope 0.01)
 * r3d_add_lrelu_bwd  d = dy * LeakyReLU'(a + b), from the sign of y (slope > 0); d is the gradient of BOTH summands
 * r3d_adam_step      torch.optim.Adam's update (trainer.py:78-81: no weight decay, no amsgrad) over flat fp32 buffers
 *                    p, g, m, v (n elements).  step: device scalar (fp32), advanced by one inside the call, so that the
 *                    update can be replayed from a CUDA graph; lr_dev (nullable): device scalar that overrides lr_host
 *                    (a scheduler updates it in place).
 * n elements, pointers of the add kernels 16-byte aligned. */
int r3d_add_lrelu(const float* a, const float* b, float* y, long long n, float slope, r3d_stream_t stream);
int r3d_add_lrelu_bwd(const float* dy, const float* y, float* d, long long n, float slope, r3d_stream_t stream);
int r3d_adam_step(float* p, const float* g, float* m, float* v, long long n, const float* lr_dev, double lr_host,
                  double beta1, double beta2, double eps, float* step, r3d_stream_t stream);

/* ------------------------------------------------------------------------- Focal-Tversky / Dice loss
 * randlanet/utils/losses.py:66-86 via trainer.py:245-269: p = softmax over classes, TI_c = (TP_c + eps) /
 * (TP_c + alpha FN_c + (1 - alpha) FP_c + eps), loss = mean over classes >= first_class of (1 - TI_c)^gamma.
 * logits (B,C,N) fp32 addressed through element strides (sb, sc, sn); labels (B,N) int64 dense; C <= 16.
 * fwd: acc (3,C) fp64 caller-zeroed scratch; writes *loss and coef (2,C) fp32 for the backward.
 * bwd: dlogits (same strides as logits) = *gout (nullable: 1) * d loss / d logits. */
int r3d_tversky_loss_fwd(const float* logits, long long sb, long long sc, long long sn, const int64_t* labels, int B,
                         int C, int N, int first_class, float alpha, float gamma, float eps, double* acc, float* loss,
                         float* coef, r3d_stream_t stream);
int r3d_tversky_loss_bwd(const float* logits, long long sb, long long sc, long long sn, const int64_t* labels, int B,
                         int C, int N, const float* coef, const float* gout, float* dlogits, r3d_stream_t stream);

/* ------------------------------------------------------------------------------ training metrics
 * randlanet/utils/metrics.py:8-59 (accuracy, iou) per batch, trainer.py:121-131: counts (C,C) int64 +=
 * [label][prediction] with prediction = arg max over classes (lowest index on ties); overall / per-class accuracy and
 * IoU follow on the host from the matrix.  One launch, no host synchronisation; logits strided like the loss; C <= 16. */
int r3d_confusion_counts(const float* logits, long long sb, long long sc, long long sn, const int64_t* labels, int B,
                         int C, int N, long long* counts, r3d_stream_t stream);

/* ------------------------------------------------------- per-point layers on the tensor cores (training)
 * SharedMLP's 1x1 convolution (randlanet/utils/modules.py:60-104) on dense rows, tcgen05 with split-fp16 operands
 * (fp32-accurate: ~1e-6 relative), for the mid-width layers of the training step:
 *   r3d_pc_gemm    y (M,cout; ldy) = act(scale * (x W^T) + shift), x (M,cin; ldx), element (o,i) of W at
 *                  w[o*w_so + i*w_si] (forward: the (cout,cin) weight, strides (cin,1); input gradient: the same
 *                  array with strides (1,cout_layer)); stats (2 cout fp64, nullable, caller-zeroed) += per-channel sums
 *                  of x W^T and of its square; absmax_x (nullable) = atomic max with max |x|.  cin % 16 == 0; one launch per block
 *                  of 128 input channels (later blocks add to the partial sums in y).
 *   r3d_pc_wgrad   out (ca,cb; ld_out; caller-zeroed) += A^T B, A (M,ca; lda), B (M,cb; ldb) like r3d_rowreduce_gemm;
 *                  absmax_a / absmax_b: device scalars that bound |A|, |B| (r3d_absmax, or absmax_x of r3d_pc_gemm).
 *                  ca, cb multiples of 8.
 * *_supported return 1 when the shape is served (else the calls return R3D_EUNSUPPORTED). */
int r3d_pc_gemm_supported(int cin, int cout, long long M);
int r3d_pc_gemm(const float* x, long long ldx, const float* w, long long w_so, long long w_si, const float* scale,
                const float* shift, int act, float slope, float* y, long long ldy, double* stats, float* absmax_x,
                long long M, int cin, int cout, r3d_stream_t stream);
int r3d_pc_wgrad_supported(int ca, int cb, long long M);
int r3d_pc_wgrad(const float* A, long long lda, int ca, const float* Bm, long long ldb, int cb, long long M,
                 const float* absmax_a, const float* absmax_b, float* out, int ld_out, r3d_stream_t stream);

/* --------------------------------------------------------------------------------- feature up-sampling
 * randlanet/utils/modules.py:343-414 (UpSampler.nearest_neighbor_interpolation / nearest_neighbors_averaging) fused
 * with the decoder's skip concat (:600-602):
 *   out[b,q,0:F] = sum_k w[b,q,k] feat[b, idx[b,q,k], :],   out[b,q,F:F+Fs] = skip[b,q,:]   (skip nullable, Fs = 0)
 * weighting 0: the first neighbour only (nni, K = 1); 1: w = (1+1e-7)/(dist^power + 1e-7) normalised over K (nna / idw
 * power 1, isdw power 2; dist = Euclidean distances (B,N2,K), not squared); 2: plain mean over K.
 * idx (B,N2,K) int32 or int64 (idx64 = 1) from r3d_knn; K <= 16.  feat (B,N1,F), skip (B,N2,Fs), out (B,N2,F+Fs) rows
 * with leading dimensions *_ld and cloud strides *_bstride in floats (0 = dense).  channel_major = 1 writes (B,F,N2)
 * instead (the layout Model.upsample returns, model.py:123-144; Fs must be 0).
 * _bwd: dfeat (B,N1,F) += scattered w * dout[:, :, 0:F] (caller zero-fills), dskip (nullable) = dout[:, :, F:]. */
int r3d_upsample(const float* feat, long long feat_bstride, int feat_ld, int F, const void* idx, int idx64,
                 const float* dist, int K, int weighting, float power, const float* skip, long long skip_bstride,
                 int skip_ld, int Fs, float* out, long long out_bstride, int out_ld, int channel_major, int B, int N1,
                 int N2, r3d_stream_t stream);
int r3d_upsample_bwd(const float* dout, long long dout_bstride, int dout_ld, const void* idx, int idx64,
                     const float* dist, int K, int weighting, float power, float* dfeat, long long dfeat_bstride,
                     int dfeat_ld, int F, float* dskip, long long dskip_bstride, int dskip_ld, int Fs, int B, int N1,
                     int N2, r3d_stream_t stream);

/* ------------------------------------------------------------------------------------- data feeding
 * randlanet/utils/dataset.py:61-97 (PointCloudPreprocessor.preprocess) + augmentation.py:24-167, one launch per batch
 * over clouds cached in device memory:
 *   points (rows, ld = 3+F) fp32 and labels (rows) int64: all clouds of the dataset back to back; row_start (B) int64 =
 *   first row of each cloud of THIS batch; sample_idx (B,n) int32 = rows within the cloud (preprocessing.py:35-62, or
 *   r3d_sample_subset).  normalization 0 none, 1 mean, 2 max, 3 stdev, 4 centre only (dataset.py:81-92).
 *   aug (B,7) fp32 nullable = per cloud [scale, angle_x, angle_y, angle_z, shift_x, shift_y, shift_z] as the reference
 *   draws them (augmentation.py:73, :99-102, :154); null: no augmentation.  noise (B,n,3) fp32 nullable = standard
 *   normals of the jitter (augmentation.py:48-52); null: Philox4x32-10 keyed by (seed, counter, cloud, point).
 *   jitter_sigma / jitter_limit = AugmentationSettings.jitter_variance / jitter_limit.
 *   out (B,n,ld) fp32 = [augmented xyz, features]; labels_out (B,n) int64 nullable.
 * r3d_sample_subset: sizes (B) int32 device = points per cloud -> out (B,n) int32: a uniform random subset of n points
 * in ascending order (N > n), or every point once followed by n - N uniform draws with replacement (N <= n). */
int r3d_feed_batch(const float* points, int ld, const int64_t* labels, const int64_t* row_start,
                   const int32_t* sample_idx, int n, int normalization, const float* aug, const float* noise,
                   float jitter_sigma, float jitter_limit, unsigned long long seed, unsigned long long counter,
                   float* out, int64_t* labels_out, int B, r3d_stream_t stream);
int r3d_sample_subset(const int32_t* sizes, int n, unsigned long long seed, unsigned long long counter, int32_t* out,
                      int B, r3d_stream_t stream);

/* ----------------------------------------------------------------------------- per-point MLP layer
 * y[b,n,:] = act(scale * (W [xa[b, g(n), :] ; xb[b,n,:]]) + shift)
 * Replaces SharedMLP / Linear on single points (modules.py:60-104; call sites :314, :325, :253, :565-566,
 * :591, :594-605, :610) fused with the tensor shuffling around them: the row gather g() is the decoder's
 * 1-NN up-sampling (modules.py:359-363) or the (inverse) point permutation (:572-573, :608); the second
 * source is the decoder skip concat (:600-602) or, with folded BN scales, the residual sum of :325.
 *   xa (B,*,ca) rows of ca floats, clouds xa_bstride apart; gidx int32 nullable, cloud b uses
 *      gidx[b*gidx_bstride + n] (gidx_bstride 0 = one index vector shared by all clouds)
 *   xb (B,n,cb) nullable; wT (ca+cb, cout) = weight transposed; scale, shift (cout) nullable
 *   act 0 none, 1 relu, 2 leaky relu(slope)
 *   y  (B,n,y_ld >= cout) rows, clouds y_bstride apart; transpose_out=1 writes (B,cout,n) instead
 *      (the logits layout, modules.py:611).  0 strides mean dense. */
int r3d_pointwise(const float* xa, long long xa_bstride, int ca, const int32_t* gidx, long long gidx_bstride,
                  const float* xb, long long xb_bstride, int cb, const float* wT, const float* scale,
                  const float* shift, int act, float slope, float* y, long long y_bstride, int y_ld, int cout,
                  int B, int n, int transpose_out, r3d_stream_t stream);

/* r3d_pointwise that also accumulates, per output channel, the sum and the sum of squares (fp64, stats[2*cout],
 * caller-zeroed, nullable) of the values it writes: the batch statistics of a train-mode BatchNorm
 * (modules.py:86-90).  w_out_in = 1: `wT` points at the weight in its stored (cout, cin) layout (no transposed copy). */
int r3d_pointwise_stats(const float* xa, long long xa_bstride, int ca, const int32_t* gidx, long long gidx_bstride,
                        const float* xb, long long xb_bstride, int cb, const float* wT, const float* scale,
                        const float* shift, int act, float slope, float* y, long long y_bstride, int y_ld, int cout,
                        int B, int n, int transpose_out, double* stats, int w_out_in, r3d_stream_t stream);

/* Wide layers (C_in >= 32, C_out a multiple of 32, channel counts multiples of 4) run on the tcgen05 tensor cores
 * with the 3xTF32 split (csrc/pointwise_tc.cu, fp32-level accuracy); everything else on the FP32 CUDA-core kernels.
 * on = 0 forces the CUDA-core kernels everywhere, 1 (default) uses the tensor-core kernel where it is measured to win
 * (>= 32768 rows and C_in >= 256 or C_out <= 32; profiles/r01_pointwise_tc_vs_fp32.txt), 2 wherever it is supported
 * (>= 4096 rows; tests, benchmarks).  Returns the previous value. */
int r3d_pointwise_set_tensor_cores(int on);
/* the kernel r3d_pointwise runs for a dense layer under the current settings: 0 pw_small_kernel, 1 pw_gemm_kernel,
 * 2 pw_gemm_fast_kernel, 3 pw_tc_kernel (tcgen05), 4 pw_rows_kernel (HBM-streaming: dense narrow layers, >= 131072 rows) */
int r3d_pointwise_plan(int ca, int cb, int cout, long long rows, int transpose_out);

/* ------------------------------------------------------------- train-mode BatchNorm of a per-point layer
 * Forward tail of SharedMLP in training mode (modules.py:92-104): z (M,C) = conv output WITHOUT bias, stats from
 * r3d_pointwise_stats.  y = act(a z + c) with a = gamma*rstd, c = beta - a*mean (the conv bias cancels against
 * the batch mean); running_mean/var (nullable) are updated like BatchNorm2d (momentum, unbiased variance; the
 * bias is added back to the mean), *num_batches (nullable) is incremented; save (3,C) = a, mean, rstd. */
int r3d_bn_apply(const float* z, const double* stats, long long M, int C, const float* gamma, const float* beta,
                 const float* bias, float eps, float momentum, float* running_mean, float* running_var,
                 long long* num_batches, int act, float slope, float* y, float* save, r3d_stream_t stream);
/* Train-mode SharedMLP forward in one call: z (M,cout) = x (M,cin) W^T with W (cout,cin), stats (2*cout fp64,
 * caller-zeroed) += batch sums, then y = act(BatchNorm_batch(z)) and the running statistics / save (3,cout) exactly as
 * r3d_bn_apply.  Runs as r3d_pointwise_stats + r3d_bn_apply, or (r3d_bn_set_fused bit 0) as ONE cooperative launch
 * with a grid barrier when the layer's grid is co-resident; results are identical.  C % 4 == 0 as for r3d_bn_apply. */
int r3d_pointwise_bn(const float* x, long long M, int cin, const float* w, int cout, double* stats, const float* gamma,
                     const float* beta, const float* bias, float eps, float momentum, float* running_mean,
                     float* running_var, long long* num_batches, int act, float slope, float* z, float* y, float* save,
                     r3d_stream_t stream);
/* Enables the single-launch (cooperative) variants: bit 0 r3d_pointwise_bn, bit 1 r3d_bn_bwd.  Default 0 (measured
 * slower inside the multi-stream training step, see csrc/pointwise_train.cu).  Returns the previous mask. */
int r3d_bn_set_fused(int mask);
/* Backward: du = dy * act'(a z + c); stats2[2C] (fp64, caller-zeroed) += (sum du, sum du*zhat) = (dbeta, dgamma);
 * then dz = a (du - mean(du) - zhat * mean(du*zhat)); dgb (2C fp32, nullable) receives (dbeta, dgamma) rounded. */
int r3d_bn_bwd_reduce(const float* dy, const float* z, long long M, int C, const float* save, const float* beta,
                      int act, float slope, double* stats2, r3d_stream_t stream);
int r3d_bn_bwd_dz(const float* dy, const float* z, long long M, int C, const float* save, const float* beta, int act,
                  float slope, const double* stats2, float* dz, float* dgb, r3d_stream_t stream);
/* Both passes: r3d_bn_bwd_reduce + r3d_bn_bwd_dz, or (r3d_bn_set_fused bit 1) one cooperative launch with a grid
 * barrier between the passes while the tensor is small (M*C <= 4 Mi elements).  Same arguments and results. */
int r3d_bn_bwd(const float* dy, const float* z, long long M, int C, const float* save, const float* beta, int act,
               float slope, double* stats2, float* dz, float* dgb, r3d_stream_t stream);
/* r3d_bn_bwd that also reports max |dz|: absmax_dz (nullable, caller-zeroed device scalar) = atomic max.  The operand
 * scale of r3d_pc_wgrad. */
int r3d_bn_bwd_absmax(const float* dy, const float* z, long long M, int C, const float* save, const float* beta, int act,
                      float slope, double* stats2, float* dz, float* dgb, float* absmax_dz, r3d_stream_t stream);
/* Weight gradient of a per-point layer: out (Ca,Cb; ld_out, caller-zeroed) += A^T B for A (M,Ca), B (M,Cb). */
int r3d_rowreduce_gemm(const float* A, int Ca, const float* Bm, int Cb, long long M, float* out, int ld_out,
                       r3d_stream_t stream);

/* ------------------------------------------------------------------------ tcgen05 tensor-core GEMM
 * C (M,N) = A (M,K) W (N,K)^T with tcgen05.mma kind::tf32, accumulators in TMEM.  terms = 3: fp32-accurate
 * 3xTF32 (operands split hi/lo on the fly, three MMAs per K step); terms = 1: plain TF32.
 * N a multiple of 32 in [32,256], K a multiple of 4; A, W, C dense, 16-byte aligned.  (csrc/tc_gemm.cu) */
int r3d_tc_gemm(const float* A, const float* W, float* C, int M, int N, int K, int terms, r3d_stream_t stream);

/* Unit test of the split-fp16 ("fp16x2") tensor-core building blocks used by the fused LFA kernels (csrc/tc16_common.cuh):
 * ONE CTA computes D (M,N) = A (M,K) B (N,K)^T with kind::f16 MMAs on hi/lo fp16 halves of the scaled operands,
 * each operand staged K-major (0) or MN-major (1) in shared memory, and writes the raw accumulator window
 * out (128 TMEM lanes, N) divided by scale_a*scale_b (cells no MMA wrote read as NaN).  M in {64,128},
 * N, K multiples of 16 <= 256.  A, B, out dense fp32 device buffers. */
int r3d_tc16_probe(const float* A, const float* B, float* out, int M, int N, int K, int a_mn_major, int b_mn_major,
                   float scale_a, float scale_b, r3d_stream_t stream);

/* ------------------------------------------------------------------------------ FP32 peak probe
 * Roofline denominator of the CUDA-core kernels, measured live by bench.py (MEASURED_PEAKS.json has
 * HBM and bf16 tensor peaks only).  mode 0 = scalar FFMA, 1 = packed FFMA2 (f32x2).  `out` is a device
 * buffer of r3d_fp32_probe_floats() floats; *flops (host, nullable) receives the launch's flop count. */
size_t r3d_fp32_probe_floats(void);
int r3d_fp32_probe(int mode, int iters, float* out, double* flops, r3d_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* R3D_B200_H */
