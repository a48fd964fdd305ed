// tcgen05.mma issue / execution rate with the conv kernels' operand layout (run on the B200 box):
//   no-swizzle K-major descriptors over 16-bit channel planes, M = 128, N = 32..256, K = 16 per instruction,
//   one thread per CTA issuing `reps` back-to-back MMAs over a ring of distinct A start addresses (tap-style shifts),
//   then one commit.  Prints clk per MMA for (a) precomputed descriptors (pure hardware rate: the larger of the
//   tensor time 128*N/256... and the shared-memory read time (4 KB + 32 N B) / 128 B/clk) and (b) descriptors rebuilt
//   with 64-bit arithmetic per instruction (what a naive issue loop costs).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/umma_rate tools/umma_rate.cu
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((addr & 0x3FFFF) >> 4) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) |
           ((uint64_t)1 << 46);
}

template <int MODE>   // 0: precomputed descriptor + 32-bit add of the start address; 1: full rebuild per MMA
__global__ void __launch_bounds__(128, 1) rate_kernel(int N, int reps, int planes_px, long long* out, int nacc, int run) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 200 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;  // 1.0h
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;");
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t taddr = tmem_base;
    if (tid == 0) {
        const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint32_t a0 = smem_u32(smem), plane = (uint32_t)planes_px * 16u;
        const uint32_t b0 = a0 + 128 * 1024;
        const uint64_t da0 = make_desc(a0, plane, 128), db0 = make_desc(b0, (uint32_t)N * 16u, 128);
        const long long t0 = clock64();
        if (MODE == 0) {
            uint32_t off = 0;
            for (int i = 0; i < reps; i += nacc * run) {       // `run` consecutive MMAs per accumulator, `nacc` accumulators round-robin
                uint32_t dc = taddr;
                for (int a = 0; a < nacc; ++a, dc += (uint32_t)N)
                    for (int r = 0; r < run; ++r) {
                        const uint64_t da = da0 + off;           // start-address field only (units of 16 B), no carry out of 14 bits
                        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                                     "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                                     ::"r"(dc), "l"(da), "l"(db0), "r"(idesc), "r"(i ? 1u : 0u));
                        off = (off + 1) & 63;
                    }
            }
        } else {
            for (int i = 0; i < reps; ++i) {
                const int m = i & 3, j = (i >> 2) & 1, tap = (i >> 3) % 9;
                const uint64_t da = make_desc(a0 + (uint32_t)((tap / 3) * 66 + tap % 3) * 16u + (uint32_t)m * 2048u + (uint32_t)(2 * j) * plane, plane, 128);
                const uint64_t db = make_desc(b0 + (uint32_t)(2 * j) * N * 16u, (uint32_t)N * 16u, 128);
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                             "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                             ::"r"(taddr + (uint32_t)(m * N)), "l"(da), "l"(db), "r"(idesc), "r"(i > 7 ? 1u : 0u) : "memory");
            }
        }
        const long long t1 = clock64();
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)));
        uint32_t done = 0;
        for (long long spin = 0; spin < (1LL << 28) && !done; ++spin)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(done) : "r"(smem_u32(&bar)), "r"(0u));
        const long long t2 = clock64();
        if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; out[2] = done; }
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(taddr));
    }
}

int main() {
    long long* d;
    CK(cudaMalloc(&d, 64));
    CK(cudaFuncSetAttribute(rate_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CK(cudaFuncSetAttribute(rate_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    const int reps = 4096;
    // accumulator dependence: `run` consecutive MMAs hit the same TMEM accumulator, then the next of `nacc` accumulators
    for (int N : {32, 64, 128})
        for (int nacc : {1, 2, 4, 8})
            for (int run : {1, 2, 4}) {
                if (nacc * N > 512) continue;
                for (int rep = 0; rep < 2; ++rep) {
                    rate_kernel<0><<<148, 128, 200 * 1024>>>(N, reps, 1290, d, nacc, run);
                    CK(cudaDeviceSynchronize());
                }
                long long h[3];
                CK(cudaMemcpy(h, d, 24, cudaMemcpyDeviceToHost));
                printf("N %3d accumulators %d run %d: %.1f clk/MMA\n", N, nacc, run, (double)h[1] / reps);
            }
    for (int mode = 0; mode < 2; ++mode)
        for (int grid : {1, 148})
            for (int N : {32, 64, 128, 256}) {
                if (4 * N > 512 && mode == 1) continue;
                for (int rep = 0; rep < 2; ++rep) {
                    if (mode == 0) rate_kernel<0><<<grid, 128, 200 * 1024>>>(N, reps, 650, d, 2, 1);
                    else rate_kernel<1><<<grid, 128, 200 * 1024>>>(N, reps, 650, d, 2, 1);
                    CK(cudaDeviceSynchronize());
                }
                long long h[3];
                CK(cudaMemcpy(h, d, 24, cudaMemcpyDeviceToHost));
                printf("mode %d grid %3d N %3d: issue %.1f clk/MMA, issue+drain %.1f clk/MMA  (tensor floor %d, smem floor %d) done=%lld\n",
                       mode, grid, N, (double)h[0] / reps, (double)h[1] / reps, N / 2, (4096 + 32 * N) / 128, h[2]);
            }
    return 0;
}
