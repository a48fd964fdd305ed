// Output head: GroupNorm apply + SiLU of the last block's raw output, 1x1 conv + bias, fp32 NCHW
// output; optional fused L1-loss partial sum.  Memory-bound (0.8 FLOP/B): one thread per pixel,
// 128-bit channel-chunk loads, coalesced planar stores.
// Reference: src/model.py:57,131 (output_conv), src/optimized_model.py:74,158 (output),
// optimized_train.py:439 (nn.L1Loss forward).
#include "common.cuh"

namespace dg {

constexpr int HEAD_THREADS = 256;
constexpr int HEAD_MAX_OC = 4;

template <typename T>
__global__ void __launch_bounds__(HEAD_THREADS) head_kernel(const dg_head_args p) {
    extern __shared__ float hsm[];
    const int C = p.src.channels;
    float* coef = hsm;            // [C][2]
    float* wsm = hsm + 2 * C;     // [OC][C]
    const int n = blockIdx.y;
    const int HW = p.H * p.W;
    for (int c = threadIdx.x; c < C; c += HEAD_THREADS) {
        float a = 1.f, b = 0.f;
        if (p.src.coef != nullptr) {
            a = __ldg(p.src.coef + (size_t)(n * C + c) * 2);
            b = __ldg(p.src.coef + (size_t)(n * C + c) * 2 + 1);
        } else if (p.src.stats != nullptr) {
            gn_coef(p.src.stats, p.src.gamma, p.src.beta, n, C, p.src.groups, c, (double)HW, p.eps, a, b);
        }
        coef[2 * c] = a;
        coef[2 * c + 1] = b;
    }
    for (int i = threadIdx.x; i < p.cout * C; i += HEAD_THREADS) wsm[i] = p.weight[i];
    __syncthreads();

    const T* raw = reinterpret_cast<const T*>(p.src.raw);
    const bool vec = (C % 8 == 0) && ((reinterpret_cast<uintptr_t>(p.src.raw) & 15) == 0);
    double l1 = 0.0;
    for (int pix = blockIdx.x * HEAD_THREADS + threadIdx.x; pix < HW; pix += gridDim.x * HEAD_THREADS) {
        float o[HEAD_MAX_OC];
#pragma unroll
        for (int j = 0; j < HEAD_MAX_OC; ++j) o[j] = (j < p.cout) ? p.bias[j] : 0.f;
        for (int c0 = 0; c0 < C; c0 += 8) {
            const int cn = min(8, C - c0);
            float v[8];
            load8<T>(raw + ((size_t)n * HW + pix) * C + c0, cn, vec, v);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                if (k < cn) {
                    float y = v[k] * coef[2 * (c0 + k)] + coef[2 * (c0 + k) + 1];
                    if (p.src.silu) y = silu_f(y);
#pragma unroll
                    for (int j = 0; j < HEAD_MAX_OC; ++j)
                        if (j < p.cout) o[j] = fmaf(y, wsm[j * C + c0 + k], o[j]);
                }
            }
        }
#pragma unroll
        for (int j = 0; j < HEAD_MAX_OC; ++j) {
            if (j < p.cout) {
                const size_t oi = ((size_t)n * p.cout + j) * HW + pix;
                p.out[oi] = o[j];
                if (p.target != nullptr) l1 += (double)fabsf(o[j] - p.target[oi]);
            }
        }
    }
    if (p.target != nullptr && p.l1_sum != nullptr) {
        for (int o = 16; o > 0; o >>= 1) l1 += __shfl_xor_sync(0xffffffffu, l1, o);
        __shared__ double red[HEAD_THREADS / 32];
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = l1;
        __syncthreads();
        if (threadIdx.x == 0) {
            double t = 0.0;
            for (int i = 0; i < HEAD_THREADS / 32; ++i) t += red[i];
            atomicAdd(p.l1_sum, t);
        }
    }
}

int head_launch(const dg_head_args& a, cudaStream_t stream) {
    if (a.cout < 1 || a.cout > HEAD_MAX_OC) {
        set_error("head: out_channels %d not in 1..%d", a.cout, HEAD_MAX_OC);
        return 3;
    }
    // >= 8 pixels per thread when the image allows it: the per-block prologue (GroupNorm coefficients in double,
    // weights to shared memory) was most of the kernel at one pixel per thread (profile r1b: 11.6 warp-inst/pixel)
    const int HW = a.H * a.W;
    int bx = (HW + HEAD_THREADS * 8 - 1) / (HEAD_THREADS * 8);
    if (bx < 1) bx = 1;
    if (bx > 1024) bx = 1024;
    dim3 grid(bx, a.N);
    const size_t smem = (size_t)(2 + a.cout) * a.src.channels * sizeof(float);
    switch (a.dtype) {
        case DG_F32: head_kernel<float><<<grid, HEAD_THREADS, smem, stream>>>(a); break;
        case DG_F16: head_kernel<__half><<<grid, HEAD_THREADS, smem, stream>>>(a); break;
        case DG_BF16: head_kernel<__nv_bfloat16><<<grid, HEAD_THREADS, smem, stream>>>(a); break;
        default: set_error("head: bad dtype %d", a.dtype); return 2;
    }
    count_launch();
    return check_launch("head1x1");
}

}  // namespace dg
