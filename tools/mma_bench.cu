// Micro-benchmarks that size the design (run on the B200 box, results recorded in DESIGN.md):
//   1. mma.sync.m16n8k16 f16 (HMMA) issue rate per SM  -> is the legacy tensor path fast enough for N=8 layers?
//   2. MUFU throughput: ex2+rcp SiLU vs tanh.approx.f32 vs tanh.approx.f16x2
//   3. ldmatrix.x4 shared-memory throughput
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdio>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__global__ void hmma_kernel(int iters, float* out) {
    unsigned a0 = threadIdx.x, a1 = threadIdx.x * 3, a2 = 7, a3 = 9, b0 = 5, b1 = 11;
    float c[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0.f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
            asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3])
                         : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    if (s == 123.456f) out[0] = s;
}

template <int MODE>
__global__ void mufu_kernel(int iters, float* out) {
    float x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = threadIdx.x * 0.001f + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) {
                x[i] = x[i] / (1.f + __expf(-x[i]));            // ex2 + IEEE divide
            } else if (MODE == 1) {
                x[i] = __fdividef(x[i], 1.f + __expf(-x[i]));   // ex2 + rcp
            } else if (MODE == 2) {
                float h = 0.5f * x[i], t;
                asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
                x[i] = fmaf(h, t, h);
            } else {
                unsigned h = __float_as_uint(x[i]), t;
                asm("tanh.approx.f16x2 %0, %1;" : "=r"(t) : "r"(h));
                x[i] = __uint_as_float(t ^ h);
            }
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += x[i];
    if (s == 123.456f) out[0] = s;
}

__global__ void ldmatrix_kernel(int iters, float* out) {
    __shared__ __align__(16) unsigned char sm[32768];
    for (int i = threadIdx.x; i < 32768 / 4; i += blockDim.x) reinterpret_cast<unsigned*>(sm)[i] = i;
    __syncthreads();
    unsigned base = (unsigned)__cvta_generic_to_shared(sm) + (threadIdx.x & 31) * 16 + (threadIdx.x >> 5) * 512;
    unsigned acc = 0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            unsigned r0, r1, r2, r3;
            asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                         : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(base + ((i * 4096 + it * 16) & 16383)));
            acc += r0 ^ r1 ^ r2 ^ r3;
        }
    }
    if (acc == 0x12345u) out[0] = acc;
}

int main() {
    cudaDeviceProp p;
    CK(cudaGetDeviceProperties(&p, 0));
    printf("device %s SMs %d L2 %d MB smem/SM %zu clock %d kHz\n", p.name, p.multiProcessorCount, p.l2CacheSize >> 20,
           p.sharedMemPerMultiprocessor, p.clockRate);
    float* out;
    CK(cudaMalloc(&out, 4));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms;
    const int sms = p.multiProcessorCount;
    for (int warps = 4; warps <= 32; warps *= 2) {
        const int iters = 20000;
        hmma_kernel<<<sms, warps * 32>>>(100, out);
        cudaEventRecord(e0);
        hmma_kernel<<<sms, warps * 32>>>(iters, out);
        cudaEventRecord(e1);
        CK(cudaEventSynchronize(e1));
        cudaEventElapsedTime(&ms, e0, e1);
        double macs = (double)sms * warps * iters * 8 * 2048.0;
        printf("HMMA m16n8k16 f16: %2d warps/SM: %.1f TFLOP/s dense-equivalent, %.0f MAC/clk/SM @1.9GHz-nominal, %.3f ms\n", warps,
               2 * macs / ms / 1e9, macs / (ms * 1e-3) / sms / 1.9e9, ms);
    }
    const char* names[4] = {"silu ex2+div", "silu ex2+rcp", "silu tanh.f32", "tanh.f16x2 (2 elts/op)"};
    for (int mode = 0; mode < 4; ++mode) {
        const int iters = 4000, warps = 32;
        auto launch = [&](int it) {
            if (mode == 0) mufu_kernel<0><<<sms * 2, warps * 32>>>(it, out);
            if (mode == 1) mufu_kernel<1><<<sms * 2, warps * 32>>>(it, out);
            if (mode == 2) mufu_kernel<2><<<sms * 2, warps * 32>>>(it, out);
            if (mode == 3) mufu_kernel<3><<<sms * 2, warps * 32>>>(it, out);
        };
        launch(10);
        cudaEventRecord(e0);
        launch(iters);
        cudaEventRecord(e1);
        CK(cudaEventSynchronize(e1));
        cudaEventElapsedTime(&ms, e0, e1);
        double ops = (double)sms * 2 * warps * 32 * iters * 8;
        printf("%-24s: %.2f Gop/s chip, %.1f ops/clk/SM @1.9GHz-nominal\n", names[mode], ops / ms / 1e6, ops / (ms * 1e-3) / sms / 1.9e9);
    }
    {
        const int iters = 4000, warps = 16;
        ldmatrix_kernel<<<sms, warps * 32>>>(10, out);
        cudaEventRecord(e0);
        ldmatrix_kernel<<<sms, warps * 32>>>(iters, out);
        cudaEventRecord(e1);
        CK(cudaEventSynchronize(e1));
        cudaEventElapsedTime(&ms, e0, e1);
        double bytes = (double)sms * warps * iters * 8 * 512.0;
        printf("ldmatrix.x4: %.1f TB/s chip, %.1f B/clk/SM @1.9GHz-nominal\n", bytes / ms / 1e9, bytes / (ms * 1e-3) / sms / 1.9e9);
    }
    return 0;
}
