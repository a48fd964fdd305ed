// tcgen05 implicit-GEMM 3x3 conv for the DEEP layers (C_out >= 32..64, K = 9*C_in >= 288): 5th-generation tensor cores with
// the accumulator in tensor memory.  Same fusion contract as conv3x3_tc.cu (GroupNorm apply + SiLU [+ AvgPool | identity
// `up` half of a concat] on load, raw NHWC output + GroupNorm statistics in the epilogue).
//
// Why here and not on the narrow layers: a UMMA reads its 128 x 16 A tile from shared memory for every instruction, so
// with N = C_out = 8..16 it is shared-memory-bound at 1/8..1/4 of the tensor rate, while from N = 64 on the 4 KB A tile is
// amortised over 64+ output channels and the measured HMMA ceiling of the legacy path (990 MAC/clk/SM, tools/mma_bench.cu)
// becomes the limit instead (these layers ran at 38-47 % tensor-pipe utilisation on mma.sync, profiles/r01_ncu_full_summary).
//
// Implicit GEMM without im2col: the activated haloed tile lives in shared memory as 16-bit channel planes
// [C/8][rows*32 pixels][8 ch]; 8 consecutive pixels of a plane are one contiguous 128-byte core matrix, so a no-swizzle
// K-major descriptor (LBO = plane stride, SBO = 128 B) describes 128 output pixels x 16 channels, and a vertical tap is a
// start-address offset (validated by tools/umma_test.cu).  M-rows must be consecutive pixels, so horizontal taps use three
// copies of the tile, pre-shifted by kx (copy_k[r][c] = tile(r, c + k)); tiles are TH x 32 pixels = TH/4 accumulators of
// 128 lanes x C_out fp32 columns in TMEM.  Weights use the HMMA path's packing ([chunk][k-half][C_out][8]) unchanged:
// it is exactly the K-major no-swizzle B layout (LBO = C_out*16, SBO = 128).
//
// One elected thread issues all 9 * C_in/16 UMMAs per accumulator and commits them to an mbarrier; all warps then read
// their TMEM lanes (tcgen05.ld 32x32b), round, store, and reduce the statistics with a halving butterfly.
#include "tc_common.cuh"

namespace dg {

namespace {
constexpr int UM_THREADS = 256;
constexpr int UM_TW = 32;
enum { UM_SAME = 0, UM_POOL = 1, UM_CAT2 = 3 };

struct UmArgs {
    const void* src0; const double* st0; const float* g0; const float* b0; const float* cf0; int groups0;
    const void* src1; const double* st1; const float* g1; const float* b1; const float* cf1; int groups1;
    const void* wgt;
    void* out; double* out_stats;
    int N, H, W; float eps;
};

constexpr int um_pow2_cols(int c) { return c <= 32 ? 32 : (c <= 64 ? 64 : (c <= 128 ? 128 : (c <= 256 ? 256 : 512))); }

template <int CIN_, int COUT_, int MODE_>
struct UGeo {
    static constexpr int CIN = CIN_, COUT = COUT_, MODE = MODE_, TH = 4;   // tile = 4 rows x 32 pixels = one 128-lane accumulator
    static constexpr int NC8 = CIN / 8, ROWS = TH + 2, HW = UM_TW + 2;
    // one channel plane of one shifted copy, +16 B so that the NC8 planes a warp writes for one pixel land in different
    // bank groups (a 128-byte-multiple stride made every staging store an 8-way bank conflict)
    static constexpr int PLANE_BYTES = ROWS * UM_TW * 16 + 16;
    static constexpr int COPY_BYTES = NC8 * PLANE_BYTES;
    static constexpr int ACT_BYTES = 3 * COPY_BYTES;           // one staged tile (three kx-shifted copies)
    static constexpr int KSTEPS = CIN / 16, NCHUNK = 9 * KSTEPS;
    static constexpr int WGT_BYTES = NCHUNK * COUT * 32;
    static constexpr int NCOEF = MODE == UM_CAT2 ? COUT : CIN;
    static constexpr int TMEM_COLS = um_pow2_cols(2 * COUT);   // two accumulator stages
    static constexpr int CHW = COUT / 2 / 16;                  // 16-column epilogue chunks per warp (a warp owns half the columns)
    static constexpr int OFF_ACT = 0;                          // two tile buffers
    static constexpr int OFF_WGT = OFF_ACT + 2 * ACT_BYTES;
    static constexpr int OFF_COEF = OFF_WGT + WGT_BYTES;
    static constexpr int OFF_STAT = OFF_COEF + NCOEF * 8;
    static constexpr int OFF_BAR = OFF_STAT + 8 * (COUT / 2) * 2 * 4;
    static constexpr int SMEM_BYTES = OFF_BAR + 32;
    static_assert(CIN % 16 == 0 && COUT % 32 == 0 && COUT <= 256, "UMMA shape");
    static_assert(UM_THREADS % NC8 == 0, "chunk ownership");
    static_assert(SMEM_BYTES <= 227 * 1024, "shared memory");
    static_assert(MODE != UM_CAT2 || CIN == 2 * COUT, "CAT2: (up C, skip C) -> C");
};

__device__ __forceinline__ uint64_t um_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((addr & 0x3FFFF) >> 4) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) |
           ((uint64_t)1 << 46);
}

// halving butterfly: 16 per-lane values -> every lane ends with the 32-lane total of ONE of them (index from lane bits 4..1)
__device__ __forceinline__ float butterfly16(float (&v)[16], int lane) {
#pragma unroll
    for (int half = 8, bit = 16; half >= 1; half >>= 1, bit >>= 1) {
        const bool hi = (lane & bit) != 0;
#pragma unroll
        for (int i = 0; i < half; ++i) {
            const float keep = hi ? v[i + half] : v[i];
            const float give = hi ? v[i] : v[i + half];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, give, bit);
        }
    }
    return v[0] + __shfl_xor_sync(0xffffffffu, v[0], 1);
}

// Persistent, software-pipelined: while the tensor core runs the 9*C_in/16 UMMAs of tile t+1 (asynchronously, into the other
// TMEM accumulator stage, from the other shared-memory tile buffer), the CUDA cores drain tile t (TMEM -> HBM + statistics) and
// then stage tile t+2.  Weights are loaded once per CTA; GroupNorm coefficients and the statistics flush are per image.
template <typename T, typename G, int ACT>
__global__ void __launch_bounds__(UM_THREADS, 1) conv3x3_umma_kernel(const UmArgs p) {
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* wgt = smem + G::OFF_WGT;
    float2* coef = reinterpret_cast<float2*>(smem + G::OFF_COEF);
    float* statw = reinterpret_cast<float*>(smem + G::OFF_STAT);   // [8 warps][COUT/2][2]
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + G::OFF_BAR);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + G::OFF_BAR + 16);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int H = p.H, W = p.W;
    constexpr int FACT = ACT == ACT_HALF2 ? ACT_TANH : ACT;
    const int tiles_x = (W + UM_TW - 1) / UM_TW, tiles_y = (H + G::TH - 1) / G::TH;
    const int tiles_per_img = tiles_x * tiles_y;
    const long long total = (long long)tiles_per_img * p.N;
    const int t0 = (int)(total * blockIdx.x / gridDim.x), t1 = (int)(total * (blockIdx.x + 1) / gridDim.x);

    // ---- (0) weights -> shared memory (once), TMEM allocation, barrier init ----------------------------------------------
    {
        const unsigned char* src = reinterpret_cast<const unsigned char*>(p.wgt);
        const uint32_t dst = smem_u32(wgt);
        for (int i = tid * 16; i < G::WGT_BYTES; i += UM_THREADS * 16) cp_async16(dst + i, src + i);
        cp_async_commit();
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[0])));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[1])));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(G::TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t taddr = *tmem_slot;
    const uint32_t act_u = smem_u32(smem + G::OFF_ACT), wgt_u = smem_u32(wgt);

    int coef_n = -1;
    auto ensure_coefs = [&](int n) {
        if (n == coef_n) return;
        __syncthreads();
        const bool cat = G::MODE == UM_CAT2;
        const double plane = G::MODE == UM_POOL ? (double)(2 * H) * (2 * W) : (double)H * W;
        for (int c = tid; c < G::NCOEF; c += UM_THREADS) {
            float a, b;
            const float* cf = cat ? p.cf1 : p.cf0;
            if (cf) { a = __ldg(cf + (size_t)(n * G::NCOEF + c) * 2); b = __ldg(cf + (size_t)(n * G::NCOEF + c) * 2 + 1); }
            else if (cat) gn_coef(p.st1, p.g1, p.b1, n, G::NCOEF, p.groups1, c, plane, p.eps, a, b);
            else gn_coef(p.st0, p.g0, p.b0, n, G::NCOEF, p.groups0, c, plane, p.eps, a, b);
            if constexpr (FACT != ACT_EXACT) { a *= 0.5f; b *= 0.5f; }
            coef[c] = make_float2(a, b);
        }
        coef_n = n;
        __syncthreads();
    };

    // ---- stage the activated haloed tile `tile` into the three kx-shifted copies of buffer `buf` ----------------------------
    auto stage = [&](int tile, int buf) {
        const int n = tile / tiles_per_img;
        const int trem = tile - n * tiles_per_img;
        const int y0 = (trem / tiles_x) * G::TH, x0 = (trem % tiles_x) * UM_TW;
        const int c8 = tid % G::NC8;
        const bool ident = G::MODE == UM_CAT2 && c8 < G::COUT / 8;              // `up` half of the concat: plain copy
        const int cc8 = (G::MODE == UM_CAT2 && !ident) ? c8 - G::COUT / 8 : c8;  // chunk index inside its source
        constexpr int Cs = G::MODE == UM_CAT2 ? G::COUT : G::CIN;                // channels of the source tensor
        float2 cf[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) cf[k] = ident ? make_float2(0.f, 0.f) : coef[cc8 * 8 + k];
        const unsigned char* srcb = reinterpret_cast<const unsigned char*>((G::MODE == UM_CAT2 && !ident) ? p.src1 : p.src0);
        const int Hs = G::MODE == UM_POOL ? 2 * H : H, Ws = G::MODE == UM_POOL ? 2 * W : W;
        srcb += (size_t)n * Hs * Ws * Cs * 2 + cc8 * 16;
        const uint32_t rowb = (uint32_t)Ws * Cs * 2;
        constexpr int NPIX = G::ROWS * G::HW;
        constexpr int PSTRIDE = UM_THREADS / G::NC8;
        constexpr int NSLOT = (NPIX + PSTRIDE - 1) / PSTRIDE;
        constexpr int MAXB = G::MODE == UM_POOL ? 2 : 4;
        constexpr int ITERS = (NSLOT + MAXB - 1) / MAXB;
        constexpr int BATCH = (NSLOT + ITERS - 1) / ITERS;
        unsigned char* dstp = smem + G::OFF_ACT + (size_t)buf * G::ACT_BYTES + (size_t)c8 * G::PLANE_BYTES;
        int hp = tid / G::NC8;
#pragma unroll 1
        for (int it = 0; it < ITERS; ++it) {
            uint4 q[BATCH][G::MODE == UM_POOL ? 4 : 1];
            int rr[BATCH], cc[BATCH];
            bool ok[BATCH];
#pragma unroll
            for (int b = 0; b < BATCH; ++b) {
                const int r = hp / G::HW, c = hp - r * G::HW;
                const int gy = y0 + r - 1, gx = x0 + c - 1;
                rr[b] = hp < NPIX ? r : -1;
                cc[b] = c;
                ok[b] = hp < NPIX && (unsigned)gy < (unsigned)H && (unsigned)gx < (unsigned)W;
                if (ok[b]) {
                    if constexpr (G::MODE == UM_POOL) {
                        const unsigned char* base = srcb + ((uint32_t)(2 * gy) * rowb + (uint32_t)(2 * gx) * (Cs * 2));
                        q[b][0] = __ldg(reinterpret_cast<const uint4*>(base));
                        q[b][1] = __ldg(reinterpret_cast<const uint4*>(base + Cs * 2));
                        q[b][2] = __ldg(reinterpret_cast<const uint4*>(base + rowb));
                        q[b][3] = __ldg(reinterpret_cast<const uint4*>(base + rowb + Cs * 2));
                    } else {
                        q[b][0] = __ldg(reinterpret_cast<const uint4*>(srcb + ((uint32_t)gy * rowb + (uint32_t)gx * (Cs * 2))));
                    }
                }
                hp += PSTRIDE;
            }
#pragma unroll
            for (int b = 0; b < BATCH; ++b) {
                if (rr[b] < 0) continue;
                uint4 o = make_uint4(0u, 0u, 0u, 0u);
                if (ok[b]) {
                    if (ident) {
                        o = q[b][0];
                    } else {
                        float y[8];
                        act8<T, FACT>(q[b][0], cf, y);
                        if constexpr (G::MODE == UM_POOL) {
                            float t[8];
#pragma unroll
                            for (int j = 1; j < 4; ++j) {
                                act8<T, FACT>(q[b][j], cf, t);
#pragma unroll
                                for (int k = 0; k < 8; ++k) y[k] += t[k];
                            }
#pragma unroll
                            for (int k = 0; k < 8; ++k) y[k] *= 0.25f;
                        }
                        o = pack8<T>(y);
                    }
                }
                // haloed column c (tile x = c - 1) lands in copy k at column c - k, for the copies where that is inside [0, 32)
                unsigned char* d = dstp + (uint32_t)(rr[b] * UM_TW + cc[b]) * 16;
                if (cc[b] < UM_TW) *reinterpret_cast<uint4*>(d) = o;
                if (cc[b] >= 1 && cc[b] <= UM_TW) *reinterpret_cast<uint4*>(d + G::COPY_BYTES - 16) = o;
                if (cc[b] >= 2) *reinterpret_cast<uint4*>(d + 2 * G::COPY_BYTES - 32) = o;
            }
        }
    };

    // ---- one thread issues the whole K loop of a tile into accumulator stage s, from tile buffer s ------------------------------
    auto issue = [&](int s) {
        constexpr uint32_t IDESC = (1u << 4) | ((std::is_same<T, __half>::value ? 0u : 1u) << 7) |
                                   ((std::is_same<T, __half>::value ? 0u : 1u) << 10) | ((uint32_t)(G::COUT >> 3) << 17) |
                                   ((uint32_t)(128 >> 4) << 24);
        const uint32_t abuf = act_u + s * G::ACT_BYTES;
#pragma unroll 1
        for (int tap = 0; tap < 9; ++tap) {
            const int ky = tap / 3, kx = tap - ky * 3;
            const uint32_t a_tap = abuf + kx * G::COPY_BYTES + (uint32_t)(ky * UM_TW) * 16;
#pragma unroll
            for (int j = 0; j < G::KSTEPS; ++j) {
                const uint64_t da = um_desc(a_tap + 2 * j * G::PLANE_BYTES, G::PLANE_BYTES, 128);
                const uint64_t db = um_desc(wgt_u + (uint32_t)((tap * G::KSTEPS + j) * G::COUT * 32), G::COUT * 16, 128);
                const uint32_t accum = (tap | j) ? 1u : 0u;
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                             "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                             ::"r"(taddr + s * G::COUT), "l"(da), "l"(db), "r"(IDESC), "r"(accum));
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bars[s])));
    };

    // per-lane running statistics: after the butterfly, even lanes own one channel of each of this warp's column chunks
    float run_s[G::CHW], run_q[G::CHW];
#pragma unroll
    for (int i = 0; i < G::CHW; ++i) run_s[i] = run_q[i] = 0.f;
    int stats_n = -1;
    const int colhalf = warp >> 2;  // warps 0-3: columns [0, COUT/2), warps 4-7: [COUT/2, COUT) of the same TMEM lanes
    auto flush_stats = [&](int n) {
        if ((lane & 1) == 0) {
            const int sub = ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
#pragma unroll
            for (int i = 0; i < G::CHW; ++i) {
                statw[(warp * (G::COUT / 2) + i * 16 + sub) * 2] = run_s[i];
                statw[(warp * (G::COUT / 2) + i * 16 + sub) * 2 + 1] = run_q[i];
            }
        }
#pragma unroll
        for (int i = 0; i < G::CHW; ++i) run_s[i] = run_q[i] = 0.f;
        __syncthreads();
        if (p.out_stats != nullptr)
            for (int c = tid; c < 2 * G::COUT; c += UM_THREADS) {
                const int ch = c >> 1, half = ch / (G::COUT / 2), cl = ch - half * (G::COUT / 2);
                double t = 0.0;
#pragma unroll
                for (int w = 0; w < 4; ++w) t += (double)statw[((half * 4 + w) * (G::COUT / 2) + cl) * 2 + (c & 1)];
                atomicAdd(p.out_stats + (size_t)n * G::COUT * 2 + c, t);
            }
        __syncthreads();
    };

    // ---- TMEM -> registers -> HBM + statistics for `tile`, accumulator stage s ------------------------------------------------
    auto epilogue = [&](int tile, int s, uint32_t parity) {
        const int n = tile / tiles_per_img;
        const int trem = tile - n * tiles_per_img;
        const int y0 = (trem / tiles_x) * G::TH, x0 = (trem % tiles_x) * UM_TW;
        if (n != stats_n) {
            if (stats_n >= 0) flush_stats(stats_n);
            stats_n = n;
        }
        {
            uint32_t done = 0;
            for (int spin = 0; spin < (1 << 26) && !done; ++spin)
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                             : "=r"(done) : "r"(smem_u32(&bars[s])), "r"(parity));
            if (!done) __trap();  // never observed; a bounded wait keeps a broken descriptor from hanging the device
        }
        asm volatile("tcgen05.fence::after_thread_sync;");
        const int gy = y0 + (warp & 3), gx = x0 + lane;
        const bool valid = gy < H && gx < W;
        T* o = reinterpret_cast<T*>(p.out) + ((size_t)(n * H + gy) * W + gx) * G::COUT + colhalf * (G::COUT / 2);
#pragma unroll
        for (int i = 0; i < G::CHW; ++i) {
            uint32_t r[16];
            const uint32_t ta = taddr + ((uint32_t)((warp & 3) * 32) << 16) + s * G::COUT + colhalf * (G::COUT / 2) + i * 16;
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                         : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                           "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                         : "r"(ta));
            asm volatile("tcgen05.wait::ld.sync.aligned;");
            float v[16], q2[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                v[j] = valid ? __uint_as_float(r[j]) : 0.f;
                q2[j] = v[j] * v[j];
            }
            if (valid) {
                *reinterpret_cast<uint4*>(o + i * 16) =
                    make_uint4(pack2<T>(v[0], v[1]), pack2<T>(v[2], v[3]), pack2<T>(v[4], v[5]), pack2<T>(v[6], v[7]));
                *reinterpret_cast<uint4*>(o + i * 16 + 8) =
                    make_uint4(pack2<T>(v[8], v[9]), pack2<T>(v[10], v[11]), pack2<T>(v[12], v[13]), pack2<T>(v[14], v[15]));
            }
            run_s[i] += butterfly16(v, lane);
            run_q[i] += butterfly16(q2, lane);
        }
        asm volatile("tcgen05.fence::before_thread_sync;");
    };

    // ---- software pipeline ----------------------------------------------------------------------------------------------------------
    const int T1 = t1 - t0;
    if (T1 > 0) {
        ensure_coefs(t0 / tiles_per_img);
        stage(t0, 0);
        cp_async_wait<0>();  // weights
        asm volatile("fence.proxy.async.shared::cta;");
        __syncthreads();
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;");
            issue(0);
        }
#pragma unroll 1
        for (int k = 0; k < T1; ++k) {
            const int tile = t0 + k, s = k & 1;
            if (k + 1 < T1) {
                ensure_coefs((tile + 1) / tiles_per_img);
                stage(tile + 1, s ^ 1);            // buffer s^1 was last read by the UMMAs of tile k-1, whose commit we waited on
                asm volatile("fence.proxy.async.shared::cta;");
                __syncthreads();                   // also: every warp has finished the epilogue of tile k-1 (TMEM stage s^1 is free)
                if (tid == 0) {
                    asm volatile("tcgen05.fence::after_thread_sync;");
                    issue(s ^ 1);
                }
            }
            epilogue(tile, s, (uint32_t)((k >> 1) & 1));
        }
        flush_stats(stats_n);
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(G::TMEM_COLS));
    }
}

template <typename T, typename G, int ACT>
int launch_um(const UmArgs& a, cudaStream_t st) {
    auto kern = conv3x3_umma_kernel<T, G, ACT>;
    static bool done = false;
    if (!done) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, G::SMEM_BYTES);
        if (e != cudaSuccess) { set_error("cudaFuncSetAttribute(%d B): %s", G::SMEM_BYTES, cudaGetErrorString(e)); return 4; }
        done = true;
    }
    static int sms = 0;
    if (sms == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    }
    const long long tiles = (long long)((a.W + UM_TW - 1) / UM_TW) * ((a.H + G::TH - 1) / G::TH) * a.N;
    if (tiles > 0x7fffffffLL) { set_error("conv3x3 umma: too many tiles"); return 3; }
    const int grid = tiles < sms ? (int)tiles : sms;   // one persistent CTA per SM (the tile buffers fill its shared memory)
    kern<<<grid, UM_THREADS, G::SMEM_BYTES, st>>>(a);
    count_launch();
    return check_launch("conv3x3_umma");
}

template <typename T, int ACT>
int dispatch_um(const UmArgs& a, int mode, int cin, int cout, cudaStream_t st, bool* handled) {
    *handled = true;
    if (mode == UM_SAME && cin == 64 && cout == 64) return launch_um<T, UGeo<64, 64, UM_SAME>, ACT>(a, st);   // enc4.3, dec4.3
    if (mode == UM_POOL && cin == 32 && cout == 64) return launch_um<T, UGeo<32, 64, UM_POOL>, ACT>(a, st);   // enc4.0
    if (mode == UM_CAT2 && cin == 64 && cout == 32) return launch_um<T, UGeo<64, 32, UM_CAT2>, ACT>(a, st);   // dec3.0
    *handled = false;
    return 0;
}
}  // namespace

int conv3x3_umma_launch(const dg_conv3x3_args& a, cudaStream_t stream, bool* handled) {
    *handled = false;
    if (a.dtype != DG_F16 && a.dtype != DG_BF16) return 0;
    // Opt-in (path bit 6).  Measured on B200 (profiles/r01_ncu_umma.txt): correct, but for THIS 486k-parameter net it only
    // ties the mma.sync kernel on 64->64 (74 vs 76 us) and loses on 64->32 (322 vs 208 us): with C <= 128 the per-pixel
    // CUDA-core work of the fused prologue/epilogue (activation, three shifted tile copies, statistics butterfly) at 8 warps
    // per SM bounds the kernel while the tensor pipe idles at 12 %.  It is the template for the wide (features_start=64)
    // variant, where MMA work grows with C^2 and prologue work with C.
    if (a.weight_tc == nullptr || a.act_sum != nullptr || a.N > 65535 || !(a.path & 64)) return 0;
    const dg_src& s0 = a.src[0];
    UmArgs u;
    memset(&u, 0, sizeof(u));
    int mode, cin;
    if (a.nsrc == 1 && (s0.xform == DG_X_SAME || s0.xform == DG_X_POOL2)) {
        if (s0.stats == nullptr || !s0.silu || s0.scale != nullptr) return 0;
        mode = s0.xform == DG_X_SAME ? UM_SAME : UM_POOL;
        cin = s0.channels;
    } else if (a.nsrc == 2 && s0.xform == DG_X_SAME && a.src[1].xform == DG_X_SAME && s0.stats == nullptr && !s0.silu &&
               s0.scale == nullptr) {
        const dg_src& s1 = a.src[1];
        if (s1.stats == nullptr || !s1.silu || s1.scale || s0.channels != a.cout || s1.channels != a.cout) return 0;
        mode = UM_CAT2;
        cin = 2 * a.cout;
        u.src1 = s1.raw; u.st1 = s1.stats; u.g1 = s1.gamma; u.b1 = s1.beta; u.cf1 = s1.coef; u.groups1 = s1.groups;
    } else {
        return 0;
    }
    if ((reinterpret_cast<uintptr_t>(s0.raw) | reinterpret_cast<uintptr_t>(a.out) | reinterpret_cast<uintptr_t>(a.weight_tc) |
         reinterpret_cast<uintptr_t>(u.src1)) & 15)
        return 0;
    u.src0 = s0.raw; u.st0 = s0.stats; u.g0 = s0.gamma; u.b0 = s0.beta; u.cf0 = s0.coef; u.groups0 = s0.groups;
    u.wgt = a.weight_tc;
    u.out = a.out; u.out_stats = a.out_stats;
    u.N = a.N; u.H = a.H; u.W = a.W; u.eps = a.eps;
    const int flavour = (a.path >> 2) & 3;
    if (a.dtype == DG_F16)
        return flavour == 1 ? dispatch_um<__half, ACT_EXACT>(u, mode, cin, a.cout, stream, handled)
                            : dispatch_um<__half, ACT_TANH>(u, mode, cin, a.cout, stream, handled);
    return flavour == 1 ? dispatch_um<__nv_bfloat16, ACT_EXACT>(u, mode, cin, a.cout, stream, handled)
                        : dispatch_um<__nv_bfloat16, ACT_TANH>(u, mode, cin, a.cout, stream, handled);
}

}  // namespace dg
