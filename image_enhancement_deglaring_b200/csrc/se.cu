// ChannelAttention (squeeze-and-excitation) of OptimizedUNet, src/optimized_model.py:161-202:
//   scale[n][c] = sigmoid(W2 . silu(W1 . mean_hw(x[n])))         (both Linear layers bias-free, hidden = max(C/16, 8))
// The global average comes for free from the `act_sum` epilogue of the conv that pools the same tensor; this kernel is
// the two tiny mat-vecs (one CTA per sample); the scale itself is applied on the consumer's load (dg_src.scale).
#include "common.cuh"

namespace dg {

__global__ void se_scale_kernel(const double* __restrict__ act_sum, double inv_plane, const float* __restrict__ w1,
                                const float* __restrict__ w2, int C, int hidden, float* __restrict__ scale) {
    extern __shared__ float ssm[];  // mean[C], h[hidden]
    float* mean = ssm;
    float* h = ssm + C;
    const int n = blockIdx.x;
    for (int c = threadIdx.x; c < C; c += blockDim.x) mean[c] = (float)(act_sum[(size_t)n * C + c] * inv_plane);
    __syncthreads();
    for (int j = threadIdx.x; j < hidden; j += blockDim.x) {
        float a = 0.f;
        for (int c = 0; c < C; ++c) a = fmaf(w1[j * C + c], mean[c], a);
        h[j] = silu_f(a);
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float a = 0.f;
        for (int j = 0; j < hidden; ++j) a = fmaf(w2[c * hidden + j], h[j], a);
        scale[(size_t)n * C + c] = 1.f / (1.f + __expf(-a));
    }
}

int se_scale_launch(const double* act_sum, double plane, const float* w1, const float* w2, int N, int C, int hidden,
                    float* scale, cudaStream_t st) {
    se_scale_kernel<<<N, 128, (size_t)(C + hidden) * sizeof(float), st>>>(act_sum, 1.0 / plane, w1, w2, C, hidden, scale);
    count_launch();
    return check_launch("se_scale");
}

}  // namespace dg
