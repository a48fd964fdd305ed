// Unit test of the tcgen05 building blocks the deep-layer conv kernel relies on (run on the B200 box):
//   * no-swizzle K-major shared-memory descriptors over the "channel plane" layout [K/8][rows][8 halfs]
//     (core matrix = 8 rows x 16 B contiguous; LBO = plane stride, SBO = 128 B),
//   * a row-SHIFTED A start address (the implicit-GEMM tap shift: +16 B per pixel, not a multiple of 128 B),
//   * tcgen05.mma kind::f16 M=128, N=64, fp32 accumulation in TMEM, commit -> mbarrier, tcgen05.ld epilogue.
// D[128 x N] = A[shift .. shift+128, :K] * B[N x K]^T is checked against the host.  (The swapped convention, LBO = 128 /
// SBO = plane stride, faults on sm_100a -- measured once, the branch is gone.)
// Every wait is a bounded spin, so a wrong descriptor cannot hang the GPU.
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

constexpr int M = 128, N = 64, K = 64, ROWS = 160;
constexpr int KP = K / 8;  // planes

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((addr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;  // descriptor version (sm_100)
    return d;                // base offset 0, LBO mode 0, layout type 0 = no swizzle
}

__global__ void __launch_bounds__(128) umma_test_kernel(const __half* __restrict__ A, const __half* __restrict__ B, float* __restrict__ D,
                                                          int shift, int* status) {
    __shared__ __align__(128) __half sA[KP * ROWS * 8];
    __shared__ __align__(128) __half sB[KP * N * 8];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // planes layout: element (row r, k) at plane k/8, row r, slot k%8
    for (int i = tid; i < ROWS * K; i += 128) {
        const int r = i / K, k = i % K;
        sA[((k / 8) * ROWS + r) * 8 + (k % 8)] = A[r * K + k];
    }
    for (int i = tid; i < N * K; i += 128) {
        const int n = i / K, k = i % K;
        sB[((k / 8) * N + n) * 8 + (k % 8)] = B[n * K + k];
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"(smem_u32(&tmem_base)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    // make the generic-proxy shared-memory writes visible to the async proxy (tensor core reads)
    asm volatile("fence.proxy.async.shared::cta;");
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t taddr = tmem_base;

    if (tid == 0) {
        // instruction descriptor: D=F32, A=B=F16, both K-major, N>>3 at bit 17, M>>4 at bit 24
        const uint32_t idesc = (1u << 4) | (0u << 7) | (0u << 10) | (0u << 15) | (0u << 16) | ((uint32_t)(N >> 3) << 17) |
                               ((uint32_t)(M >> 4) << 24);
        const uint32_t a_plane = ROWS * 16, b_plane = N * 16;  // bytes between K-adjacent core matrices
        for (int ks = 0; ks < K / 16; ++ks) {
            const uint32_t a_addr = smem_u32(sA) + (2 * ks) * a_plane + shift * 16;
            const uint32_t b_addr = smem_u32(sB) + (2 * ks) * b_plane;
            const uint64_t da = make_desc(a_addr, a_plane, 128);   // LBO = plane stride (K-adjacent core matrices), SBO = 128 B
            const uint64_t db = make_desc(b_addr, b_plane, 128);
            const uint32_t acc = ks > 0 ? 1u : 0u;
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                         "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                         ::"r"(taddr), "l"(da), "l"(db), "r"(idesc), "r"(acc));
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)));
    }
    // bounded wait on the commit barrier (phase 0)
    bool done = false;
    for (int it = 0; it < 2000000 && !done; ++it) {
        uint32_t ok;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(smem_u32(&bar)));
        done = ok != 0;
    }
    if (!done) {
        if (tid == 0) *status = 1;  // timed out
    } else {
        asm volatile("tcgen05.fence::after_thread_sync;");
        // warp w reads TMEM lanes 32w..32w+31: thread = one row of D, 64 fp32 columns, 16 at a time
        for (int c0 = 0; c0 < N; c0 += 16) {
            uint32_t r[16];
            const uint32_t ta = taddr + ((uint32_t)(warp * 32) << 16) + c0;
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                         : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                           "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                         : "r"(ta));
            asm volatile("tcgen05.wait::ld.sync.aligned;");
            for (int j = 0; j < 16; ++j) D[(warp * 32 + lane) * N + c0 + j] = __uint_as_float(r[j]);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(taddr));
}

int main() {
    std::vector<__half> hA(ROWS * K), hB(N * K);
    std::vector<float> fA(ROWS * K), fB(N * K);
    srand(1);
    for (int i = 0; i < ROWS * K; ++i) { float v = (rand() % 2001 - 1000) / 1000.f; hA[i] = __float2half(v); fA[i] = __half2float(hA[i]); }
    for (int i = 0; i < N * K; ++i) { float v = (rand() % 2001 - 1000) / 1000.f; hB[i] = __float2half(v); fB[i] = __half2float(hB[i]); }
    __half *dA, *dB; float* dD; int* dS;
    CK(cudaMalloc(&dA, hA.size() * 2)); CK(cudaMalloc(&dB, hB.size() * 2)); CK(cudaMalloc(&dD, M * N * 4)); CK(cudaMalloc(&dS, 4));
    CK(cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice));
    int rc = 0;
    {
        for (int shift : {0, 3, 17}) {
            CK(cudaMemset(dD, 0, M * N * 4)); CK(cudaMemset(dS, 0, 4));
            umma_test_kernel<<<1, 128>>>(dA, dB, dD, shift, dS);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("shift=%d: kernel error %s\n", shift, cudaGetErrorString(e)); return 2; }
            std::vector<float> hD(M * N); int st = 0;
            CK(cudaMemcpy(hD.data(), dD, M * N * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(&st, dS, 4, cudaMemcpyDeviceToHost));
            double maxerr = 0;
            for (int m = 0; m < M; ++m)
                for (int n = 0; n < N; ++n) {
                    double ref = 0;
                    for (int k = 0; k < K; ++k) ref += (double)fA[(m + shift) * K + k] * fB[n * K + k];
                    const double err = fabs(ref - hD[m * N + n]);
                    if (err > maxerr) maxerr = err;
                }
            printf("LBO=plane,SBO=128  shift %2d: status %d  max|err| %.3e  %s\n", shift, st, maxerr,
                   (st == 0 && maxerr < 1e-3) ? "OK" : "MISMATCH");
            if (st != 0 || maxerr >= 1e-3) rc = 1;
        }
    }
    return rc;
}
