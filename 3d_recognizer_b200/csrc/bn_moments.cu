// Train-mode BatchNorm of a 1x1 convolution expressed through the MOMENTS of its input (sm_100a).
//
// For y = W x (+ bias) over R rows with input sums S = sum x and M = sum x x^T (fp64, from
// r3d_lfa_moments): mean_y = W mu, var_y = diag(W Cov W^T), mu = S/R, Cov = M/R - mu mu^T, so that
// BatchNorm(y) = a (W x) + c with a = gamma / sqrt(var_y + eps), c = beta - a (W mu) (the bias cancels).
// This is how the fused LocSE kernels get the batch statistics of mlp_rpe1 / mlp_rpe2
// (randlanet/utils/modules.py:86-90, statistics over all B*N*K positions) without materialising y.
//   bnm_fwd_kernel   one CTA per output channel: a, c, running statistics, and the scalars backward needs
//   bnm_bwd_kernel   one CTA per output channel: dW, dgamma, dbeta and per-channel scalars gq, gwmu
//   bnm_bwd_moments_kernel   dM = W^T diag(gq) W, dS = W^T gwmu / R   (only when the moments depend on parameters)
// Replaces ~60 tiny tensor-op launches per LFA block and step (profiles/r01_launches_train2500.txt).
#include "common.cuh"

namespace r3d {

constexpr int kBnmThreads = 128;
constexpr int kBnmMaxCin = 128;

__device__ __forceinline__ double block_sum(double v, double* red) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
    for (int i = 0; i < kBnmThreads / 32; ++i) t += red[i];
    __syncthreads();
    return t;
}

// save (5, cout) doubles: wmu, var, rstd, a, (unused)
__global__ void __launch_bounds__(kBnmThreads) bnm_fwd_kernel(
    const float* __restrict__ W, int cin, const double* __restrict__ S, int s_stride, const double* __restrict__ M,
    int ldm, double R, const float* __restrict__ gamma, const float* __restrict__ beta, const float* __restrict__ bias,
    float eps, float momentum, float* __restrict__ running_mean, float* __restrict__ running_var,
    long long* __restrict__ num_batches, float* __restrict__ a_out, float* __restrict__ c_out,
    double* __restrict__ save, int cout) {
    __shared__ double red[kBnmThreads / 32];
    const int j = blockIdx.x, i = threadIdx.x;
    const float* w = W + (size_t)j * cin;
    double q_part = 0.0, m_part = 0.0;
    if (i < cin) {
        double t = 0.0;
        for (int k = 0; k < cin; ++k) t += M[(size_t)i * ldm + k] * (double)w[k];
        q_part = (double)w[i] * t;
        m_part = (double)w[i] * S[(size_t)i * s_stride];
    }
    const double q = block_sum(q_part, red);
    const double wmu = block_sum(m_part, red) / R;
    if (i == 0) {
        double var = q / R - wmu * wmu;
        var = var > 0.0 ? var : 0.0;
        const double rstd = 1.0 / sqrt(var + (double)eps);
        const double a = (double)gamma[j] * rstd;
        a_out[j] = (float)a;
        c_out[j] = (float)((double)beta[j] - a * wmu);
        save[j] = wmu;
        save[cout + j] = var;
        save[2 * cout + j] = rstd;
        save[3 * cout + j] = a;
        if (running_mean) {
            const double unbiased = var * (R / (R > 1.0 ? R - 1.0 : 1.0));
            running_mean[j] = (1.f - momentum) * running_mean[j] + momentum * (float)(wmu + (bias ? (double)bias[j] : 0.0));
            running_var[j] = (1.f - momentum) * running_var[j] + momentum * (float)unbiased;
        }
        if (j == 0 && num_batches) *num_batches += 1;
    }
}

// scal (2, cout) doubles: gq, gwmu
__global__ void __launch_bounds__(kBnmThreads) bnm_bwd_kernel(
    const float* __restrict__ W, int cin, const double* __restrict__ S, int s_stride, const double* __restrict__ M,
    int ldm, double R, const float* __restrict__ gamma, const double* __restrict__ save, const double* __restrict__ ga,
    const double* __restrict__ gc, double* __restrict__ dW, float* __restrict__ dgamma, float* __restrict__ dbeta,
    double* __restrict__ scal, int cout) {
    const int j = blockIdx.x, i = threadIdx.x;
    const float* w = W + (size_t)j * cin;
    const double wmu = save[j], rstd = save[2 * cout + j], a = save[3 * cout + j];
    const double gc_j = gc[j];
    const double ga1 = ga[j] - gc_j * wmu;                 // total gradient at a_j
    const double gvar = ga1 * (double)gamma[j] * (-0.5) * rstd * rstd * rstd;
    const double gq = gvar / R;
    const double gwmu = -gc_j * a - 2.0 * wmu * gvar;
    if (i < cin) {
        double t = 0.0;
        for (int k = 0; k < cin; ++k) t += (M[(size_t)i * ldm + k] + M[(size_t)k * ldm + i]) * (double)w[k];
        dW[(size_t)j * cin + i] = gq * t + gwmu * S[(size_t)i * s_stride] / R;
    }
    if (i == 0) {
        dgamma[j] = (float)(ga1 * rstd);
        dbeta[j] = (float)gc_j;
        scal[j] = gq;
        scal[cout + j] = gwmu;
    }
}

// dM[i][k] = sum_j gq_j W[j][i] W[j][k];  dS[i] = sum_j gwmu_j W[j][i] / R
__global__ void __launch_bounds__(kBnmThreads) bnm_bwd_moments_kernel(const float* __restrict__ W, int cin, int cout,
                                                                      double R, const double* __restrict__ scal,
                                                                      double* __restrict__ dM, double* __restrict__ dS) {
    const int i = blockIdx.x, k = threadIdx.x;
    if (k >= cin) return;
    double t = 0.0, s = 0.0;
    for (int j = 0; j < cout; ++j) {
        const double wji = (double)W[(size_t)j * cin + i];
        t += scal[j] * wji * (double)W[(size_t)j * cin + k];
        if (k == 0) s += scal[cout + j] * wji;
    }
    dM[(size_t)i * cin + k] = t;
    if (k == 0) dS[i] = s / R;
}

// Parameter gradients of mlp_rpe1 (train mode) in one launch: G (cout, ldg) fp64 holds, per output channel, the sums
// over all (point, neighbour) rows of du (x) x in columns 0..cin-1 and of du in column cin (accumulated by the fused
// LocSE backward kernels of BOTH halves of the block, u = a (W x) + c).  With ga = sum_i W_i G_i (gradient at a),
// gc = G[cin] (gradient at c) this is bnm_bwd_kernel plus the direct term a G:  dW = a G + (BatchNorm terms).
__global__ void __launch_bounds__(kBnmThreads) bnm_rpe1_grads_kernel(
    const float* __restrict__ W, int cin, const double* __restrict__ S, int s_stride, const double* __restrict__ M,
    int ldm, double R, const float* __restrict__ gamma, const double* __restrict__ save, const double* __restrict__ G,
    int ldg, float* __restrict__ dW, float* __restrict__ dgamma, float* __restrict__ dbeta, int cout) {
    __shared__ double red[kBnmThreads / 32];
    const int j = blockIdx.x, i = threadIdx.x;
    const float* w = W + (size_t)j * cin;
    const double* g = G + (size_t)j * ldg;
    const double ga = block_sum(i < cin ? (double)w[i] * g[i] : 0.0, red);
    const double wmu = save[j], rstd = save[2 * cout + j], a = save[3 * cout + j];
    const double gc_j = g[cin];
    const double ga1 = ga - gc_j * wmu;
    const double gvar = ga1 * (double)gamma[j] * (-0.5) * rstd * rstd * rstd;
    const double gq = gvar / R;
    const double gwmu = -gc_j * a - 2.0 * wmu * gvar;
    if (i < cin) {
        double t = 0.0;
        for (int k = 0; k < cin; ++k) t += (M[(size_t)i * ldm + k] + M[(size_t)k * ldm + i]) * (double)w[k];
        dW[(size_t)j * cin + i] = (float)(a * g[i] + gq * t + gwmu * S[(size_t)i * s_stride] / R);
    }
    if (i == 0) {
        dgamma[j] = (float)(ga1 * rstd);
        dbeta[j] = (float)gc_j;
    }
}

// Coefficients of the second BatchNorm-backward pass of mlp_rpe2 (r3d_lfa_bn2_bwd) from the batch sums of pass 1:
// sums (2,h) fp64 = (sum du2, sum du2 r2), r2 = relu(a2 z2 + c2) the forward's activations, save (5,h) the forward's
// statistics (bnm_fwd_kernel).  sum du2 zhat2 = rstd ((sum du2 r2 - c2 sum du2) / a2 - mean sum du2)  (du2 != 0 only
// where r2 = a2 z2 + c2).  bn2 (5,h) fp32 = a2, mean, rstd, mean du2, mean du2 zhat2;  dgamma = sum du2 zhat2, dbeta = sum du2.
__global__ void bn2_coeffs_kernel(const double* __restrict__ sums, const float* __restrict__ a2, const float* __restrict__ c2,
                                  const double* __restrict__ save, double rows, int h, float* __restrict__ bn2,
                                  float* __restrict__ dgamma, float* __restrict__ dbeta) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= h) return;
    const double a = (double)a2[i], c = (double)c2[i], mu = save[i], rs = save[2 * h + i];
    const double s_du = sums[i], s_dur = sums[h + i];
    const double s_duz = a == 0.0 ? 0.0 : rs * ((s_dur - c * s_du) / a - mu * s_du);
    bn2[i] = (float)a;
    bn2[h + i] = (float)mu;
    bn2[2 * h + i] = (float)rs;
    bn2[3 * h + i] = (float)(s_du / rows);
    bn2[4 * h + i] = (float)(s_duz / rows);
    dgamma[i] = (float)s_duz;
    dbeta[i] = (float)s_du;
}

}  // namespace r3d

using namespace r3d;

extern "C" int r3d_bn_from_moments(const float* W, int cout, int cin, const double* S, int s_stride, const double* M,
                                   int ldm, double R, const float* gamma, const float* beta, const float* bias,
                                   float eps, float momentum, float* running_mean, float* running_var,
                                   long long* num_batches, float* a_out, float* c_out, double* save,
                                   r3d_stream_t stream) {
    if (cout <= 0 || cin <= 0 || !(R > 0.0)) return R3D_EINVAL;
    if (cin > kBnmMaxCin) return R3D_EUNSUPPORTED;
    if (!W || !S || !M || !gamma || !beta || !a_out || !c_out || !save) return R3D_EINVAL;
    bnm_fwd_kernel<<<cout, kBnmThreads, 0, static_cast<cudaStream_t>(stream)>>>(
        W, cin, S, s_stride, M, ldm, R, gamma, beta, bias, eps, momentum, running_mean, running_var, num_batches,
        a_out, c_out, save, cout);
    R3D_LAUNCH_CHECK("bnm_fwd_kernel");
    return R3D_OK;
}

extern "C" int r3d_bn_from_moments_bwd(const float* W, int cout, int cin, const double* S, int s_stride,
                                       const double* M, int ldm, double R, const float* gamma, const double* save,
                                       const double* ga, const double* gc, double* dW, float* dgamma, float* dbeta,
                                       double* scal, double* dM, double* dS, r3d_stream_t stream) {
    if (cout <= 0 || cin <= 0 || !(R > 0.0)) return R3D_EINVAL;
    if (cin > kBnmMaxCin) return R3D_EUNSUPPORTED;
    if (!W || !S || !M || !gamma || !save || !ga || !gc || !dW || !dgamma || !dbeta || !scal) return R3D_EINVAL;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    bnm_bwd_kernel<<<cout, kBnmThreads, 0, st>>>(W, cin, S, s_stride, M, ldm, R, gamma, save, ga, gc, dW, dgamma, dbeta,
                                                 scal, cout);
    R3D_LAUNCH_CHECK("bnm_bwd_kernel");
    if (dM && dS) {
        bnm_bwd_moments_kernel<<<cin, kBnmThreads, 0, st>>>(W, cin, cout, R, scal, dM, dS);
        R3D_LAUNCH_CHECK("bnm_bwd_moments_kernel");
    }
    return R3D_OK;
}

extern "C" int r3d_lfa_rpe1_grads(const float* W, int cout, int cin, const double* S, int s_stride, const double* M,
                                  int ldm, double R, const float* gamma, const double* save, const double* G, int ldg,
                                  float* dW, float* dgamma, float* dbeta, r3d_stream_t stream) {
    if (cout <= 0 || cin <= 0 || ldg <= cin || !(R > 0.0)) return R3D_EINVAL;
    if (cin > kBnmMaxCin) return R3D_EUNSUPPORTED;
    if (!W || !S || !M || !gamma || !save || !G || !dW || !dgamma || !dbeta) return R3D_EINVAL;
    bnm_rpe1_grads_kernel<<<cout, kBnmThreads, 0, static_cast<cudaStream_t>(stream)>>>(W, cin, S, s_stride, M, ldm, R, gamma,
                                                                                      save, G, ldg, dW, dgamma, dbeta, cout);
    R3D_LAUNCH_CHECK("bnm_rpe1_grads_kernel");
    return R3D_OK;
}

extern "C" int r3d_lfa_bn2_coeffs(const double* sums, const float* a2, const float* c2, const double* save, double rows,
                                  int h, float* bn2, float* dgamma, float* dbeta, r3d_stream_t stream) {
    if (h <= 0 || !(rows > 0.0)) return R3D_EINVAL;
    if (!sums || !a2 || !c2 || !save || !bn2 || !dgamma || !dbeta) return R3D_EINVAL;
    bn2_coeffs_kernel<<<ceil_div(h, 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(sums, a2, c2, save, rows, h, bn2,
                                                                                      dgamma, dbeta);
    R3D_LAUNCH_CHECK("bn2_coeffs_kernel");
    return R3D_OK;
}
