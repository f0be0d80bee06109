// Confusion counts of one batch in ONE launch and without a host synchronisation (sm_100a).
//
// Reference: randlanet/utils/metrics.py:8-59 (accuracy, iou) as called per batch by the trainer (trainer.py:121-131):
// torch.max over the class axis, then per class 2-3 masked sums each followed by `.cpu().item()` — 2 + 3C device
// round trips per batch.  Everything those functions return is a function of the C x C confusion matrix
// counts[label][prediction]; this kernel adds a batch's matrix into a device slot, and the host reads all slots of an
// epoch at once (3d_recognizer_b200/metrics.py turns them into the reference's per-batch values).
// prediction = arg max_c logits[b, c, n], the lowest class index on ties.  logits addressed through strides like the
// loss kernels (the network hands over a transposed view).
#include "common.cuh"

namespace r3d {

constexpr int kMetricsMaxC = 16;

__global__ void __launch_bounds__(256) confusion_kernel(const float* __restrict__ logits, long long sb, long long sc,
                                                        long long sn, const int64_t* __restrict__ labels, int B, int C,
                                                        int N, unsigned long long* __restrict__ counts /* (C,C) */) {
    __shared__ unsigned int cm[kMetricsMaxC * kMetricsMaxC];
    for (int i = threadIdx.x; i < C * C; i += blockDim.x) cm[i] = 0u;
    __syncthreads();
    const long long total = (long long)B * N;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int b = (int)(i / N), n = (int)(i % N);
        const float* x = logits + b * sb + n * sn;
        float best = x[0];
        int arg = 0;
        for (int c = 1; c < C; ++c) {
            const float v = x[c * sc];
            if (v > best) {
                best = v;
                arg = c;
            }
        }
        const int lab = (int)labels[i];
        if (lab >= 0 && lab < C) atomicAdd(&cm[lab * C + arg], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < C * C; i += blockDim.x)
        if (cm[i]) atomicAdd(&counts[i], (unsigned long long)cm[i]);
}

}  // namespace r3d

using namespace r3d;

extern "C" int r3d_confusion_counts(const float* logits, long long sb, long long sc, long long sn, const int64_t* labels,
                                    int B, int C, int N, long long* counts, r3d_stream_t stream) {
    if (B < 0 || N < 0 || C <= 0) return R3D_EINVAL;
    if (C > kMetricsMaxC) return R3D_EUNSUPPORTED;
    if (B == 0 || N == 0) return R3D_OK;
    if (!logits || !labels || !counts) return R3D_EINVAL;
    const long long total = (long long)B * N;
    long long blocks = (total + 255) / 256;
    if (blocks > 2 * kNumSMs) blocks = 2 * kNumSMs;
    confusion_kernel<<<(int)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        logits, sb, sc, sn, labels, B, C, N, reinterpret_cast<unsigned long long*>(counts));
    R3D_LAUNCH_CHECK("confusion_kernel");
    return R3D_OK;
}
