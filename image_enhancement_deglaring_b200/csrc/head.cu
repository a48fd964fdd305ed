// Output head: GroupNorm apply + SiLU of the last block's raw output, 1x1 conv + bias, fp32 NCHW
// output; optional fused L1-loss partial sum.  Memory-bound (0.8 FLOP/B): one thread per pixel,
// 128-bit channel-chunk loads, coalesced planar stores.
// Reference: src/model.py:57,131 (output_conv), src/optimized_model.py:74,158 (output),
// optimized_train.py:439 (nn.L1Loss forward).
#include <type_traits>

#include "common.cuh"

namespace dg {

constexpr int HEAD_THREADS = 256;
constexpr int HEAD_MAX_OC = 4;

// out_kind 1: the /infer post-processing `(np.clip(y, 0, 1) * 255).astype(np.uint8)` (api/app.py:190-193): fp32 multiply, truncation
__device__ __forceinline__ unsigned char quant_u8(float y) {
    return (unsigned char)__float2uint_rz(__fmul_rn(fminf(fmaxf(y, 0.f), 1.f), 255.f));
}

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
                if (p.out_kind == 1) reinterpret_cast<unsigned char*>(p.out)[oi] = quant_u8(o[j]);
                else reinterpret_cast<float*>(p.out)[oi] = o[j];
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

// ---- fast path: 16-bit storage, C = 8 or 16 channels, one output channel, no fused loss ------------------------------
// The generic kernel above spends ~300 instructions per pixel on runtime channel loops and shared-memory coefficient
// reads (profile r1d: 86 % issue-slot utilisation at 23 % of HBM peak).  Here the affine, the 1x1 weights and the bias
// live in registers, SiLU is h + h*tanh(h), and each thread has four 128-bit loads in flight: ~45 instructions per pixel.
template <typename T, int C>
__global__ void __launch_bounds__(HEAD_THREADS) head_fast_kernel(const dg_head_args p) {
    constexpr int NC8 = C / 8;
    __shared__ float2 cfs[C];
    __shared__ float ws[C];
    const int n = blockIdx.y;
    const int HW = p.H * p.W;
    pdl_launch_dependents();
    pdl_wait();
    if (threadIdx.x < C) {
        float a, b;
        if (p.src.coef != nullptr) {
            a = __ldg(p.src.coef + (size_t)(n * C + threadIdx.x) * 2);
            b = __ldg(p.src.coef + (size_t)(n * C + threadIdx.x) * 2 + 1);
        } else {
            gn_coef(p.src.stats, p.src.gamma, p.src.beta, n, C, p.src.groups, threadIdx.x, (double)HW, p.eps, a, b);
        }
        cfs[threadIdx.x] = make_float2(0.5f * a, 0.5f * b);  // silu(y) = h + h*tanh(h), h = y/2
        ws[threadIdx.x] = p.weight[threadIdx.x];
    }
    __syncthreads();
    float2 cf[C];
    float w[C];
#pragma unroll
    for (int c = 0; c < C; ++c) { cf[c] = cfs[c]; w[c] = ws[c]; }
    const float bias = p.bias[0];
    const uint4* raw = reinterpret_cast<const uint4*>(p.src.raw) + (size_t)n * HW * NC8;
    float* out = reinterpret_cast<float*>(p.out) + (size_t)n * HW;
    unsigned char* out8 = reinterpret_cast<unsigned char*>(p.out) + (size_t)n * HW;
    const bool u8 = p.out_kind == 1;
    constexpr int PB = 4;  // pixels in flight per thread
    for (int pix0 = blockIdx.x * HEAD_THREADS * PB + threadIdx.x; pix0 < HW; pix0 += gridDim.x * HEAD_THREADS * PB) {
        uint4 q[PB][NC8];
#pragma unroll
        for (int b = 0; b < PB; ++b) {
            const int pix = pix0 + b * HEAD_THREADS;
            if (pix < HW) {
#pragma unroll
                for (int k = 0; k < NC8; ++k) q[b][k] = __ldg(raw + (size_t)pix * NC8 + k);
            }
        }
#pragma unroll
        for (int b = 0; b < PB; ++b) {
            const int pix = pix0 + b * HEAD_THREADS;
            if (pix < HW) {
                float o = bias;
#pragma unroll
                for (int k = 0; k < NC8; ++k) {
                    const uint32_t wd[4] = {q[b][k].x, q[b][k].y, q[b][k].z, q[b][k].w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        float2 v;
                        if constexpr (sizeof(T) == 2 && std::is_same<T, __half>::value) v = __half22float2(*reinterpret_cast<const __half2*>(&wd[e]));
                        else v = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&wd[e]));
                        const int c = k * 8 + 2 * e;
                        float h0 = fmaf(v.x, cf[c].x, cf[c].y), h1 = fmaf(v.y, cf[c + 1].x, cf[c + 1].y);
                        float t0, t1;
                        asm("tanh.approx.f32 %0, %1;" : "=f"(t0) : "f"(h0));
                        asm("tanh.approx.f32 %0, %1;" : "=f"(t1) : "f"(h1));
                        o = fmaf(fmaf(h0, t0, h0), w[c], o);
                        o = fmaf(fmaf(h1, t1, h1), w[c + 1], o);
                    }
                }
                if (u8) out8[pix] = quant_u8(o);
                else out[pix] = o;
            }
        }
    }
}

template <typename T>
static bool head_fast(const dg_head_args& a, cudaStream_t stream) {
    const int C = a.src.channels;
    if (a.cout != 1 || a.target != nullptr || !a.src.silu || (a.src.stats == nullptr && a.src.coef == nullptr)) return false;
    if ((C != 8 && C != 16) || (reinterpret_cast<uintptr_t>(a.src.raw) & 15)) return false;
    const int HW = a.H * a.W;
    int bx = (HW + HEAD_THREADS * 8 - 1) / (HEAD_THREADS * 8);
    if (bx < 1) bx = 1;
    dim3 grid(bx, a.N);
    if (C == 8) launch_kernel(head_fast_kernel<T, 8>, grid, dim3(HEAD_THREADS), (size_t)0, stream, a);
    else launch_kernel(head_fast_kernel<T, 16>, grid, dim3(HEAD_THREADS), (size_t)0, stream, a);
    return true;
}

int head_launch(const dg_head_args& a, cudaStream_t stream) {
    if (a.cout < 1 || a.cout > HEAD_MAX_OC) {
        set_error("head: out_channels %d not in 1..%d", a.cout, HEAD_MAX_OC);
        return 3;
    }
    if (a.out_kind != 0 && a.out_kind != 1) { set_error("head: bad out_kind %d", a.out_kind); return 2; }
    if (a.out_kind == 1 && a.target != nullptr) { set_error("head: the fused L1 sum needs the fp32 output"); return 2; }
    // >= 8 pixels per thread when the image allows it: the per-block prologue (GroupNorm coefficients in double,
    // weights to shared memory) was most of the kernel at one pixel per thread (profile r1b: 11.6 warp-inst/pixel)
    const int HW = a.H * a.W;
    int bx = (HW + HEAD_THREADS * 8 - 1) / (HEAD_THREADS * 8);
    if (bx < 1) bx = 1;
    if (bx > 1024) bx = 1024;
    dim3 grid(bx, a.N);
    const size_t smem = (size_t)(2 + a.cout) * a.src.channels * sizeof(float);
    if (a.N <= 65535 && ((a.dtype == DG_F16 && head_fast<__half>(a, stream)) ||
                         (a.dtype == DG_BF16 && head_fast<__nv_bfloat16>(a, stream)))) {
        count_launch();
        return check_launch("head_fast");
    }
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
