// Data gradient of the 3x3 convs on the tensor cores (16-bit storage tiers):
//   dA[n, y, x, ci] = sum over (ky, kx, co) of  dR[n, y+1-ky, x+1-kx, co] * W[co][ci][ky][kx]
// i.e. a 3x3 conv of dR (fp32 NHWC, identity source, zero padding) with the taps flipped and Cin / Cout swapped
// (autograd of src/model.py:93,96).  As an implicit GEMM: M = pixels, K = (tap', co), N = ci.  A fragments come from the
// haloed dR tile staged as bf16 channel planes (the forward kernels' layout, tap = +16 B/pixel shift); B fragments are read
// with ldmatrix.trans straight from the FORWARD weights' tensor-core packing ([chunk][k-half][Cout][8], bf16): a matrix
// there is 8 co rows x 8 ci columns, and transposed it is exactly the (k = co, n = ci) fragment of tap' = 8 - tap.
// bf16 operands (dR ~ 1/numel underflows fp16), fp32 accumulate, fp32 NHWC output.  Replaces the generic CUDA-core conv
// on flipped fp32 weights for these layers: 6.4 ms of an 18 ms batch-32 training step before this kernel.
#include "tc_common.cuh"

namespace dg {

namespace {

constexpr int DGR_THREADS = 256;

constexpr int dgr_pad_plane(int pix, int nc8) {
    const int want = nc8 >= 8 ? 1 : (nc8 <= 1 ? 0 : 8 / nc8);
    if (nc8 <= 1) return pix;
    int p = pix;
    while (p % 8 != want) ++p;
    return p;
}

struct DgradArgs {
    const float* dR;   // [N,H,W,CK]
    const void* wtc;   // forward packing of W [CK = Cout_fwd][CN = Cin_fwd][3][3] in bf16 (dg_pack_conv3x3_tc)
    float* out;        // [N,H,W,CN]
    int N, H, W;
    const void* dRb;   // optional: dR as bf16 [N,H,W,CK] (see gn_bwd_apply): the tile is copied, not converted
    int dry;           // probe only
    float* out2;       // optional (plain epilogue): channels [CN/2, CN) go HERE as a compact [N,H,W,CN/2] tensor and channels
                       // [0, CN/2) to `out`, also compact -- the two halves of a concat gradient have different consumers
                       // (ConvTranspose backward / skip activation backward), which otherwise read every other 32 bytes
    // optional fused activation backward (the conv's only input is the activated output of ONE producer conv, src/model.py:93-98):
    // out = G = dA * silu'(GroupNorm(raw_prev)) instead of dA, and P[n][c] += (sum G, sum G * xhat) -- what act_bwd_vec would
    // compute from a materialised dA (backward.cu), without writing and re-reading it
    const void* raw_prev;      // [N,H,W,CN] storage dtype, NULL = plain data gradient
    const double* stats_prev;  // [N,CN,2]
    const float* gamma_prev; const float* beta_prev;
    double* P;                 // [N,CN,2], accumulated
    int groups_prev, dtype_prev;
    float eps;
};

__device__ __forceinline__ float dgr_silu_grad(float y) {
    const float s = 1.f / (1.f + __expf(-y));
    return s * (1.f + y * (1.f - s));
}

// the producer description shared by the two data-gradient kernels
struct ActFuse {
    const void* raw; const double* stats; const float* gamma; const float* beta; double* P;
    int groups, dtype; float eps;
};

// (mean, rstd, gamma, beta) of channel ch0 + threadIdx.x of image n (threads < NB); published by the caller's next barrier
template <int CN>
__device__ __forceinline__ void act_fuse_coef(const ActFuse& f, int n, int ch0, double plane, float4* pcoef) {
    const int c = ch0 + threadIdx.x;
    const int cpg = CN / f.groups, g0 = (c / cpg) * cpg;
    double s1 = 0.0, s2 = 0.0;
    for (int k = 0; k < cpg; ++k) {
        s1 += f.stats[(size_t)(n * CN + g0 + k) * 2];
        s2 += f.stats[(size_t)(n * CN + g0 + k) * 2 + 1];
    }
    const double cnt = plane * cpg;
    double inv = (double)__frcp_rn((float)cnt);
    inv = inv * (2.0 - cnt * inv);
    inv = inv * (2.0 - cnt * inv);
    const double mean = s1 * inv;
    double var = fma(s2, inv, -mean * mean);
    if (var < 0.0) var = 0.0;
    const double xv = var + (double)f.eps;
    double r = (double)rsqrtf((float)xv);
    r = r * (1.5 - 0.5 * xv * r * r);
    r = r * (1.5 - 0.5 * xv * r * r);
    pcoef[threadIdx.x] = make_float4((float)mean, (float)r, f.gamma[c], f.beta[c]);
}

// Epilogue of both kernels when the activation backward is fused: out = G = acc * silu'(GroupNorm(raw)), P += (sum G, sum G*xhat).
// acc[m][j][.] is the mma accumulator of m-tile (warp + 8 m) = 16 pixels of row mt / SEGS, n-tile j of this CTA's NB channels.
template <int CN, int NB, int MPW, int SEGS>
__device__ __forceinline__ void act_fuse_epilogue(const ActFuse& f, float* out, float (&acc)[MPW][NB / 8][4], const float4* pcoef,
                                                  float (*pslot)[NB][2], int n, int H, int W, int y0, int x0, int ch0) {
    constexpr int NB8 = NB / 8;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, q = lane & 3;
    float p1[NB8][2], p2[NB8][2];
#pragma unroll
    for (int j = 0; j < NB8; ++j) p1[j][0] = p1[j][1] = p2[j][0] = p2[j][1] = 0.f;
#pragma unroll
    for (int m = 0; m < MPW; ++m) {
        const int mt = warp + 8 * m;
        const int gy = y0 + mt / SEGS;
        if (gy >= H) continue;
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
            const int gx = x0 + (mt % SEGS) * 16 + g + 8 * hf;
            if (gx >= W) continue;
            const size_t e = ((size_t)(n * H + gy) * W + gx) * CN + ch0 + 2 * q;
            float* o = out + e;
#pragma unroll
            for (int j = 0; j < NB8; ++j) {
                float r0, r1;
                if (f.dtype == DG_F16) {
                    const float2 v = __half22float2(*reinterpret_cast<const __half2*>(reinterpret_cast<const __half*>(f.raw) + e + j * 8));
                    r0 = v.x; r1 = v.y;
                } else {
                    const float2 v = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(reinterpret_cast<const __nv_bfloat16*>(f.raw) + e + j * 8));
                    r0 = v.x; r1 = v.y;
                }
                const float4 c0 = pcoef[j * 8 + 2 * q], c1 = pcoef[j * 8 + 2 * q + 1];
                const float xh0 = (r0 - c0.x) * c0.y, xh1 = (r1 - c1.x) * c1.y;
                const float g0 = acc[m][j][2 * hf] * dgr_silu_grad(xh0 * c0.z + c0.w);
                const float g1 = acc[m][j][2 * hf + 1] * dgr_silu_grad(xh1 * c1.z + c1.w);
                *reinterpret_cast<float2*>(o + j * 8) = make_float2(g0, g1);
                p1[j][0] += g0; p1[j][1] += g1;
                p2[j][0] = fmaf(g0, xh0, p2[j][0]); p2[j][1] = fmaf(g1, xh1, p2[j][1]);
            }
        }
    }
    // lanes with equal q hold the same channels: reduce over g, then over the 8 warps in a fixed order
#pragma unroll
    for (int j = 0; j < NB8; ++j)
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            float a = p1[j][k], b = p2[j][k];
#pragma unroll
            for (int o = 4; o < 32; o <<= 1) {
                a += __shfl_xor_sync(0xffffffffu, a, o);
                b += __shfl_xor_sync(0xffffffffu, b, o);
            }
            if (lane < 4) { pslot[warp][j * 8 + 2 * lane + k][0] = a; pslot[warp][j * 8 + 2 * lane + k][1] = b; }
        }
    __syncthreads();
    for (int i = tid; i < 2 * NB; i += DGR_THREADS) {
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < 8; ++w) t += (double)pslot[w][i >> 1][i & 1];
        atomicAdd(f.P + ((size_t)n * CN + ch0 + (i >> 1)) * 2 + (i & 1), t);
    }
}

__device__ __forceinline__ void dgr_ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void dgr_ldsm_x2_t(uint32_t addr, uint32_t& r0, uint32_t& r1) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(addr));
}

// CK: channels of dR (K side), CN: channels of the result, NB: result channels per CTA (grid.z = CN / NB)
template <int CK, int CN, int NB, int TH, int TW>
struct DgrGeo {
    static constexpr int PH = TH + 2, PW = TW + 2, KC8 = CK / 8, NB8 = NB / 8;
    static constexpr int PLANE = dgr_pad_plane(PH * PW, KC8);
    static constexpr bool PAIR = CK == 8;                    // K chunk of 16 = two taps x 8 channels
    static constexpr int KCH = PAIR ? 1 : CK / 16;           // K chunks per tap
    static constexpr int SEGS = TW / 16, MTILES = TH * SEGS, MPW = MTILES / 8;  // 8 m-warps, every warp all NB8 n-tiles
    static constexpr int A_BYTES = KC8 * PLANE * 16;
    static constexpr int W_BYTES = 9 * NB8 * CK * 16;        // [tap_f][n8][co][8 ci]
    static constexpr int SMEM = A_BYTES + W_BYTES;
    static_assert(MTILES % 8 == 0 && MPW * NB8 * 4 <= 64, "accumulator budget");
    static_assert(CK % 8 == 0 && (PAIR || CK % 16 == 0) && CN % NB == 0 && NB % 8 == 0 && TW % 16 == 0, "shape");
    static_assert(NB8 == 1 || NB8 % 2 == 0, "n-tiles come in ldmatrix.x4 pairs");
    static_assert(DGR_THREADS % KC8 == 0, "chunk ownership");
};

template <int CK, int CN, int NB, int TH, int TW>
__global__ void __launch_bounds__(DGR_THREADS) dgrad_tc_kernel(const DgradArgs p) {
    using G = DgrGeo<CK, CN, NB, TH, TW>;
    using BF = __nv_bfloat16;
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned char* act = smem;
    unsigned char* wsm = smem + G::A_BYTES;
    __shared__ float4 pcoef[NB];          // fused activation backward: (mean, rstd, gamma, beta) of this CTA's channels
    __shared__ float pslot[8][NB][2];     // per-warp (sum G, sum G*xhat)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = blockIdx.y;
    const int tiles_x = (p.W + TW - 1) / TW;
    const int y0 = (blockIdx.x / tiles_x) * TH, x0 = (blockIdx.x % tiles_x) * TW;
    const int nb8_0 = blockIdx.z * G::NB8;   // first 8-channel block of the result handled here
    const int H = p.H, W = p.W;

    // ---- weights of this CTA's result channels: for every forward tap and n8 block, CK rows of 16 bytes ----------------
    {
        const unsigned char* src = reinterpret_cast<const unsigned char*>(p.wtc);
        const uint32_t dst = smem_u32(wsm);
        constexpr int PIECES = 9 * G::NB8 * CK;
        for (int i = tid; i < PIECES; i += DGR_THREADS) {
            const int co = i % CK, t2 = i / CK;
            const int j8 = t2 % G::NB8, tap = t2 / G::NB8;
            const int n8g = nb8_0 + j8;
            int chunk, khalf;
            if constexpr (CN >= 16) { chunk = tap * (CN / 16) + (n8g >> 1); khalf = n8g & 1; }
            else { chunk = tap >> 1; khalf = tap & 1; }
            cp_async16(dst + (uint32_t)i * 16, src + ((size_t)(chunk * 2 + khalf) * CK + co) * 16);
        }
        cp_async_commit();
    }
    const ActFuse fuse{p.raw_prev, p.stats_prev, p.gamma_prev, p.beta_prev, p.P, p.groups_prev, p.dtype_prev, p.eps};
    if (p.raw_prev != nullptr && tid < NB) act_fuse_coef<CN>(fuse, n, blockIdx.z * NB, (double)p.H * p.W, pcoef);   // published by the barrier below
    // ---- haloed dR tile -> bf16 channel planes, zero outside the image -----------------------------------------------
    {
        const int c8 = tid % G::KC8;
        const float* gsrc = p.dR + (size_t)n * H * W * CK + c8 * 8;
        unsigned char* dst = act + (size_t)c8 * G::PLANE * 16;
        // four items (eight 128-bit loads) in flight per thread: nothing else is live yet, registers are free here
        constexpr int STEP = DGR_THREADS / G::KC8, NPIX = G::PH * G::PW, SB = 4;
        if (p.dRb != nullptr) {
            const unsigned char* bsrc = reinterpret_cast<const unsigned char*>(p.dRb) + ((size_t)n * H * W * CK + c8 * 8) * 2;
#pragma unroll 1
            for (int pix0 = tid / G::KC8; pix0 < NPIX; pix0 += SB * STEP) {
                uint4 q[SB];
#pragma unroll
                for (int k = 0; k < SB; ++k) {
                    const int pix = pix0 + k * STEP;
                    const int r = pix / G::PW, c = pix - r * G::PW;
                    const int gy = y0 + r - 1, gx = x0 + c - 1;
                    q[k] = make_uint4(0u, 0u, 0u, 0u);
                    if (pix < NPIX && (unsigned)gy < (unsigned)H && (unsigned)gx < (unsigned)W)
                        q[k] = __ldg(reinterpret_cast<const uint4*>(bsrc + ((size_t)gy * W + gx) * CK * 2));
                }
#pragma unroll
                for (int k = 0; k < SB; ++k) {
                    const int pix = pix0 + k * STEP;
                    if (pix < NPIX) *reinterpret_cast<uint4*>(dst + (size_t)pix * 16) = q[k];
                }
            }
        } else {
#pragma unroll 1
        for (int pix0 = tid / G::KC8; pix0 < NPIX; pix0 += SB * STEP) {
            float4 va[SB], vb[SB];
            bool ok[SB];
#pragma unroll
            for (int k = 0; k < SB; ++k) {
                const int pix = pix0 + k * STEP;
                const int r = pix / G::PW, c = pix - r * G::PW;
                const int gy = y0 + r - 1, gx = x0 + c - 1;
                ok[k] = pix < NPIX && (unsigned)gy < (unsigned)H && (unsigned)gx < (unsigned)W;
                if (ok[k]) {
                    va[k] = __ldg(reinterpret_cast<const float4*>(gsrc + ((size_t)gy * W + gx) * CK));
                    vb[k] = __ldg(reinterpret_cast<const float4*>(gsrc + ((size_t)gy * W + gx) * CK) + 1);
                }
            }
#pragma unroll
            for (int k = 0; k < SB; ++k) {
                const int pix = pix0 + k * STEP;
                if (pix >= NPIX) continue;
                uint4 o = make_uint4(0u, 0u, 0u, 0u);
                if (ok[k]) o = make_uint4(pack2<BF>(va[k].x, va[k].y), pack2<BF>(va[k].z, va[k].w), pack2<BF>(vb[k].x, vb[k].y),
                                          pack2<BF>(vb[k].z, vb[k].w));
                *reinterpret_cast<uint4*>(dst + (size_t)pix * 16) = o;
            }
        }
        }
    }
    cp_async_wait<0>();
    __syncthreads();

    // ---- main loop ---------------------------------------------------------------------------------------------------
    const uint32_t act_u = smem_u32(act), wsm_u = smem_u32(wsm);
    float acc[G::MPW][G::NB8][4];
    uint32_t a_pix[G::MPW];
#pragma unroll
    for (int m = 0; m < G::MPW; ++m) {
        const int mt = warp + 8 * m;
        a_pix[m] = (uint32_t)(((mt / G::SEGS) * G::PW + (mt % G::SEGS) * 16 + (lane & 15)) * 16);
#pragma unroll
        for (int j = 0; j < G::NB8; ++j) acc[m][j][0] = acc[m][j][1] = acc[m][j][2] = acc[m][j][3] = 0.f;
    }
    // B lane addressing inside one (tap, n8) block of CK rows: matrix = lane >> 3
    //   regular: (rows +0..7, n8 j), (rows +8..15, n8 j), (rows +0..7, n8 j+1), (rows +8..15, n8 j+1)
    //   PAIR:    (tap lo, n8 j), (tap hi, n8 j), (tap lo, n8 j+1), (tap hi, n8 j+1)   -- tap offset added per chunk
    constexpr uint32_t BLK = CK * 16;  // bytes of one (tap, n8) block
    if constexpr (G::PAIR) {
#pragma unroll
        for (int ch = 0; ch < 5; ++ch) {
            const int tp_lo = 2 * ch, tp_hi = (2 * ch + 1 < 9) ? 2 * ch + 1 : 2 * ch;   // taps of the dgrad conv (shift of dR)
            const int o_lo = ((tp_lo / 3) * G::PW + (tp_lo % 3)) * 16, o_hi = ((tp_hi / 3) * G::PW + (tp_hi % 3)) * 16;
            const uint32_t a_off = (uint32_t)(o_lo + (lane >> 4) * (o_hi - o_lo));
            const int tf = 8 - (((lane >> 3) & 1) ? tp_hi : tp_lo);   // forward tap whose weights multiply this shift
            uint32_t bf[G::NB8][2];
            if constexpr (G::NB8 == 1) {
                dgr_ldsm_x2_t(wsm_u + (uint32_t)(tf * G::NB8) * BLK + (uint32_t)((lane & 7) * 16), bf[0][0], bf[0][1]);
            } else {
#pragma unroll
                for (int jp = 0; jp < G::NB8 / 2; ++jp)
                    dgr_ldsm_x4_t(wsm_u + (uint32_t)(tf * G::NB8 + 2 * jp + (lane >> 4)) * BLK + (uint32_t)((lane & 7) * 16),
                                  bf[2 * jp][0], bf[2 * jp][1], bf[2 * jp + 1][0], bf[2 * jp + 1][1]);
            }
            if (ch == 4) {   // the last chunk holds tap 8 only: its upper K half must not contribute
#pragma unroll
                for (int j = 0; j < G::NB8; ++j) bf[j][1] = 0u;
            }
#pragma unroll
            for (int m = 0; m < G::MPW; ++m) {
                uint32_t a0, a1, a2, a3;
                ldsm_x4(act_u + a_pix[m] + a_off, a0, a1, a2, a3);
#pragma unroll
                for (int j = 0; j < G::NB8; ++j) mma16816<BF>(acc[m][j], a0, a1, a2, a3, bf[j][0], bf[j][1]);
            }
        }
    } else {
#pragma unroll 1
        for (int tp = 0; tp < 9; ++tp) {
            const int tf = 8 - tp;
            const uint32_t a_tap = (uint32_t)(((tp / 3) * G::PW + (tp % 3)) * 16);
#pragma unroll
            for (int kc = 0; kc < G::KCH; ++kc) {
                const uint32_t a_off = a_tap + (uint32_t)((2 * kc + (lane >> 4)) * G::PLANE * 16);
                uint32_t bf[G::NB8][2];
                const uint32_t row = (uint32_t)((kc * 16 + ((lane >> 3) & 1) * 8 + (lane & 7)) * 16);
                if constexpr (G::NB8 == 1) {
                    dgr_ldsm_x2_t(wsm_u + (uint32_t)(tf * G::NB8) * BLK + row, bf[0][0], bf[0][1]);
                } else {
#pragma unroll
                    for (int jp = 0; jp < G::NB8 / 2; ++jp)
                        dgr_ldsm_x4_t(wsm_u + (uint32_t)(tf * G::NB8 + 2 * jp + (lane >> 4)) * BLK + row, bf[2 * jp][0], bf[2 * jp][1],
                                      bf[2 * jp + 1][0], bf[2 * jp + 1][1]);
                }
#pragma unroll
                for (int m = 0; m < G::MPW; ++m) {
                    uint32_t a0, a1, a2, a3;
                    ldsm_x4(act_u + a_pix[m] + a_off, a0, a1, a2, a3);
#pragma unroll
                    for (int j = 0; j < G::NB8; ++j) mma16816<BF>(acc[m][j], a0, a1, a2, a3, bf[j][0], bf[j][1]);
                }
            }
        }
    }
    // ---- epilogue: fp32 NHWC ------------------------------------------------------------------------------------------
    const int g = lane >> 2, q = lane & 3;
    if (p.raw_prev != nullptr) {
        act_fuse_epilogue<CN, NB, G::MPW, G::SEGS>(fuse, p.out, acc, pcoef, pslot, n, H, W, y0, x0, blockIdx.z * NB);
        return;
    }
#pragma unroll
    for (int m = 0; m < G::MPW; ++m) {
        const int mt = warp + 8 * m;
        const int gy = y0 + mt / G::SEGS;
        if (gy >= H) continue;
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
            const int gx = x0 + (mt % G::SEGS) * 16 + g + 8 * hf;
            if (gx >= W) continue;
            if (p.out2 != nullptr) {
                constexpr int HC = CN / 2;
                const size_t pix = (size_t)(n * H + gy) * W + gx;
#pragma unroll
                for (int j = 0; j < G::NB8; ++j) {
                    const int ch = nb8_0 * 8 + j * 8 + 2 * q;
                    float* o = ch < HC ? p.out + pix * HC + ch : p.out2 + pix * HC + (ch - HC);
                    *reinterpret_cast<float2*>(o) = make_float2(acc[m][j][2 * hf], acc[m][j][2 * hf + 1]);
                }
                continue;
            }
            float* o = p.out + ((size_t)(n * H + gy) * W + gx) * CN + nb8_0 * 8 + 2 * q;
#pragma unroll
            for (int j = 0; j < G::NB8; ++j)
                *reinterpret_cast<float2*>(o + j * 8) = make_float2(acc[m][j][2 * hf], acc[m][j][2 * hf + 1]);
        }
    }
}

template <int CK, int CN, int NB, int TH, int TW>
int launch_dgr(const DgradArgs& a, cudaStream_t st) {
    using G = DgrGeo<CK, CN, NB, TH, TW>;
    auto kern = dgrad_tc_kernel<CK, CN, NB, TH, TW>;
    static bool done = false;
    if (!done) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, G::SMEM);
        if (e != cudaSuccess) { set_error("dgrad_tc: cudaFuncSetAttribute(%d B): %s", G::SMEM, cudaGetErrorString(e)); return 4; }
        done = true;
    }
    if (a.dry) return 0;
    dim3 grid(((a.W + TW - 1) / TW) * ((a.H + TH - 1) / TH), a.N, CN / NB);
    kern<<<grid, DGR_THREADS, G::SMEM, st>>>(a);
    count_launch();
    return check_launch("dgrad_tc");
}


// ---- ConvTranspose2d(k=2, s=2) data gradient (src/model.py:47-53 backward) ----------------------------------------------
//   dLow[n, i, j, ci] = sum over (a, b, co) of  dUp[n, 2i+a, 2j+b, co] * Wt[ci][co][a][b]
// GEMM with M = low pixels, K = (position, co) = 4 CU, N = ci.  A = the up half of the concat gradient gathered per position
// into bf16 planes [pos * CU/8 + co8][low pixel][8]; B = the forward ConvTranspose packing ([ci/16][k-half][pos*CU + co][8],
// bf16) read with ldmatrix.trans: rows (pos, co) = K, 8 ci per row = N.
struct CtDgradArgs {
    const float* dCat; int stride;   // [N, 2Hl, 2Wl, stride], up half = channels 0..CU
    const void* wtc;                 // dg_pack_convt2x2_tc(..., DG_BF16)
    float* out;                      // [N, Hl, Wl, CL]
    int N, Hl, Wl;
    ActFuse act;                     // act.raw != NULL: fused activation backward of the low-resolution producer (out = G, act.P += sums)
};

template <int CL, int CU, int NB, int TH, int TW>
__global__ void __launch_bounds__(DGR_THREADS) convt_dgrad_tc_kernel(const CtDgradArgs p) {
    using BF = __nv_bfloat16;
    constexpr int KP = 4 * CU / 8;                 // A planes: (pos, co8)
    constexpr int KCH = 4 * CU / 16;               // K chunks
    constexpr int NB8 = NB / 8, CTN = 4 * CU;
    constexpr int SEGS = TW / 16, MTILES = TH * SEGS, MPW = MTILES / 8;
    constexpr int PLANE = dgr_pad_plane(TH * TW, KP);
    constexpr int A_BYTES = KP * PLANE * 16, W_BYTES = NB8 * CTN * 16;   // weights: [n8][k = pos*CU + co][8 ci]
    static_assert(MTILES % 8 == 0 && MPW * NB8 * 4 <= 64 && CL % NB == 0 && (NB8 == 1 || NB8 % 2 == 0), "shape");
    static_assert(DGR_THREADS % (CU / 8) == 0, "chunk ownership");
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned char* act = smem;
    unsigned char* wsm = smem + A_BYTES;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = blockIdx.y;
    const int Hl = p.Hl, Wl = p.Wl, W = 2 * Wl;
    const int tiles_x = (Wl + TW - 1) / TW;
    const int y0 = (blockIdx.x / tiles_x) * TH, x0 = (blockIdx.x % tiles_x) * TW;
    const int nb8_0 = blockIdx.z * NB8;
    __shared__ float4 pcoef[NB];
    __shared__ float pslot[8][NB][2];
    if (p.act.raw != nullptr && tid < NB) act_fuse_coef<CL>(p.act, n, blockIdx.z * NB, (double)Hl * Wl, pcoef);   // published by the barrier below
    {   // packed weights of this CTA's ci blocks: n8 block j -> (chunk = n8g / 2, k-half = n8g & 1), CTN rows of 16 bytes
        const unsigned char* src = reinterpret_cast<const unsigned char*>(p.wtc);
        const uint32_t dst = smem_u32(wsm);
        for (int i = tid; i < NB8 * CTN; i += DGR_THREADS) {
            const int k = i % CTN, j8 = i / CTN;
            const int n8g = nb8_0 + j8;
            cp_async16(dst + (uint32_t)i * 16, src + ((size_t)((n8g >> 1) * 2 + (n8g & 1)) * CTN + k) * 16);
        }
        cp_async_commit();
    }
    {   // gradient planes
        constexpr int CU8 = CU / 8;
        const int c8 = tid % CU8;
        const float* gsrc = p.dCat + (size_t)n * (2 * Hl) * W * p.stride + c8 * 8;
#pragma unroll 4
        for (int it = tid / CU8; it < 4 * TH * TW; it += DGR_THREADS / CU8) {
            const int pos = it / (TH * TW), pix = it - pos * (TH * TW);
            const int r = pix / TW, c = pix - r * TW;
            const int gy = y0 + r, gx = x0 + c;
            uint4 o = make_uint4(0u, 0u, 0u, 0u);
            if (gy < Hl && gx < Wl) {
                const float* q = gsrc + ((size_t)(2 * gy + (pos >> 1)) * W + 2 * gx + (pos & 1)) * p.stride;
                const float4 a = __ldg(reinterpret_cast<const float4*>(q));
                const float4 b = __ldg(reinterpret_cast<const float4*>(q) + 1);
                o = make_uint4(pack2<BF>(a.x, a.y), pack2<BF>(a.z, a.w), pack2<BF>(b.x, b.y), pack2<BF>(b.z, b.w));
            }
            *reinterpret_cast<uint4*>(act + ((size_t)(pos * CU8 + c8) * PLANE + pix) * 16) = o;
        }
    }
    cp_async_wait<0>();
    __syncthreads();
    const uint32_t act_u = smem_u32(act), wsm_u = smem_u32(wsm);
    float acc[MPW][NB8][4];
    uint32_t a_pix[MPW];
#pragma unroll
    for (int m = 0; m < MPW; ++m) {
        a_pix[m] = (uint32_t)(((warp + 8 * m) * 16 + (lane & 15)) * 16);
#pragma unroll
        for (int j = 0; j < NB8; ++j) acc[m][j][0] = acc[m][j][1] = acc[m][j][2] = acc[m][j][3] = 0.f;
    }
#pragma unroll 2
    for (int kc = 0; kc < KCH; ++kc) {
        uint32_t bf[NB8][2];
        const uint32_t row = (uint32_t)((kc * 16 + ((lane >> 3) & 1) * 8 + (lane & 7)) * 16);
        if constexpr (NB8 == 1) {
            dgr_ldsm_x2_t(wsm_u + row, bf[0][0], bf[0][1]);
        } else {
#pragma unroll
            for (int jp = 0; jp < NB8 / 2; ++jp)
                dgr_ldsm_x4_t(wsm_u + (uint32_t)((2 * jp + (lane >> 4)) * CTN * 16) + row, bf[2 * jp][0], bf[2 * jp][1], bf[2 * jp + 1][0],
                              bf[2 * jp + 1][1]);
        }
        const uint32_t a_off = (uint32_t)((2 * kc + (lane >> 4)) * PLANE * 16);
#pragma unroll
        for (int m = 0; m < MPW; ++m) {
            uint32_t a0, a1, a2, a3;
            ldsm_x4(act_u + a_pix[m] + a_off, a0, a1, a2, a3);
#pragma unroll
            for (int j = 0; j < NB8; ++j) mma16816<BF>(acc[m][j], a0, a1, a2, a3, bf[j][0], bf[j][1]);
        }
    }
    if (p.act.raw != nullptr) {
        act_fuse_epilogue<CL, NB, MPW, SEGS>(p.act, p.out, acc, pcoef, pslot, n, Hl, Wl, y0, x0, blockIdx.z * NB);
        return;
    }
    const int g = lane >> 2, q = lane & 3;
#pragma unroll
    for (int m = 0; m < MPW; ++m) {
        const int mt = warp + 8 * m;
        const int gy = y0 + mt / SEGS;
        if (gy >= Hl) continue;
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
            const int gx = x0 + (mt % SEGS) * 16 + g + 8 * hf;
            if (gx >= Wl) continue;
            float* o = p.out + ((size_t)(n * Hl + gy) * Wl + gx) * CL + nb8_0 * 8 + 2 * q;
#pragma unroll
            for (int j = 0; j < NB8; ++j)
                *reinterpret_cast<float2*>(o + j * 8) = make_float2(acc[m][j][2 * hf], acc[m][j][2 * hf + 1]);
        }
    }
}

template <int CL, int CU, int NB, int TH, int TW>
int launch_ctdgr(const CtDgradArgs& a, cudaStream_t st) {
    constexpr int KP = 4 * CU / 8;
    constexpr int SMEM = KP * dgr_pad_plane(TH * TW, KP) * 16 + (NB / 8) * 4 * CU * 16;
    auto kern = convt_dgrad_tc_kernel<CL, CU, NB, TH, TW>;
    static bool done = false;
    if (!done) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
        if (e != cudaSuccess) { set_error("convt_dgrad_tc: cudaFuncSetAttribute(%d B): %s", SMEM, cudaGetErrorString(e)); return 4; }
        done = true;
    }
    dim3 grid(((a.Wl + TW - 1) / TW) * ((a.Hl + TH - 1) / TH), a.N, CL / NB);
    kern<<<grid, DGR_THREADS, SMEM, st>>>(a);
    count_launch();
    return check_launch("convt_dgrad_tc");
}

}  // namespace

// dLow [N,H/2,W/2,Cl] = ConvTranspose2d data gradient of the up half (channels 0..Cu of dCat [N,H,W,stride]);
// wtc_bf16 = dg_pack_convt2x2_tc(packed Wt, Cl, Cu, DG_BF16)
int convt_dgrad_tc_launch(const float* dCat, int stride, const void* wtc_bf16, float* dLow, int N, int H, int W, int Cl, int Cu,
                          cudaStream_t st, bool* handled, const DgradAct* act) {
    *handled = false;
    if (wtc_bf16 == nullptr || N < 1 || N > 65535 || ((H | W) & 1) || (stride & 3)) return 0;
    if ((reinterpret_cast<uintptr_t>(dCat) | reinterpret_cast<uintptr_t>(dLow) | reinterpret_cast<uintptr_t>(wtc_bf16)) & 15) return 0;
    CtDgradArgs a{dCat, stride, wtc_bf16, dLow, N, H / 2, W / 2, ActFuse{nullptr, nullptr, nullptr, nullptr, nullptr, 1, DG_F16, 1e-5f}};
    if (act != nullptr) {
        if ((act->dtype != DG_F16 && act->dtype != DG_BF16) || (reinterpret_cast<uintptr_t>(act->raw) & 3) || act->groups < 1 ||
            Cl % act->groups != 0)
            return 0;
        a.act = ActFuse{act->raw, act->stats, act->gamma, act->beta, act->P, act->groups, act->dtype, act->eps};
    }
    *handled = true;
    if (Cl == 128 && Cu == 64) return launch_ctdgr<128, 64, 64, 4, 32>(a, st);   // upconv4
    if (Cl == 64 && Cu == 32) return launch_ctdgr<64, 32, 64, 8, 32>(a, st);     // upconv3
    if (Cl == 32 && Cu == 16) return launch_ctdgr<32, 16, 32, 8, 32>(a, st);     // upconv2
    if (Cl == 16 && Cu == 8) return launch_ctdgr<16, 8, 16, 8, 32>(a, st);       // upconv1
    *handled = false;
    return 0;
}

namespace {
}  // namespace

// out[N,H,W,cn] = conv3x3(dR[N,H,W,ck], flipped / transposed W); wtc_bf16 = dg_pack_conv3x3_tc(W packed, cin = cn, cout = ck, DG_BF16)
// `act` (optional): fuse the producer's activation backward into the epilogue -- see DgradArgs
int conv3x3_dgrad_tc_launch(const float* dR, const void* wtc_bf16, float* out, int N, int H, int W, int ck, int cn,
                            cudaStream_t st, bool* handled, const DgradAct* act, const void* dR_bf16, bool dry, float* out2) {
    *handled = false;
    if (wtc_bf16 == nullptr || N < 1 || N > 65535) return 0;
    if ((reinterpret_cast<uintptr_t>(dR) | reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(wtc_bf16)) & 15) return 0;
    if (dR_bf16 != nullptr && (reinterpret_cast<uintptr_t>(dR_bf16) & 15)) return 0;
    if (out2 != nullptr && ((reinterpret_cast<uintptr_t>(out2) & 15) || (cn & 15) || act != nullptr)) return 0;
    DgradArgs a{dR, wtc_bf16, out, N, H, W, dR_bf16, dry ? 1 : 0, out2, nullptr, nullptr, nullptr, nullptr, nullptr, 1, DG_F16, 1e-5f};
    if (act != nullptr) {
        if ((act->dtype != DG_F16 && act->dtype != DG_BF16) || (reinterpret_cast<uintptr_t>(act->raw) & 3) || act->groups < 1 ||
            cn % act->groups != 0)
            return 0;
        a.raw_prev = act->raw; a.stats_prev = act->stats; a.gamma_prev = act->gamma; a.beta_prev = act->beta; a.P = act->P;
        a.groups_prev = act->groups; a.dtype_prev = act->dtype; a.eps = act->eps;
    }
    *handled = true;
#define DG_DGR(CK_, CN_, NB_, TH_, TW_) if (ck == CK_ && cn == CN_) return launch_dgr<CK_, CN_, NB_, TH_, TW_>(a, st);
    DG_DGR(8, 8, 8, 16, 32)        // enc1.3, dec1.3
    DG_DGR(16, 8, 8, 16, 32)       // enc2.0
    DG_DGR(16, 16, 16, 16, 32)     // enc2.3, dec2.3
    DG_DGR(32, 16, 16, 16, 32)     // enc3.0
    DG_DGR(32, 32, 32, 8, 32)      // enc3.3, dec3.3
    DG_DGR(64, 32, 32, 8, 32)      // enc4.0
    DG_DGR(64, 64, 32, 8, 32)      // enc4.3, dec4.3
    DG_DGR(128, 64, 32, 8, 32)     // bottleneck.0
    DG_DGR(128, 128, 32, 8, 32)    // bottleneck.3
    DG_DGR(64, 128, 32, 8, 32)     // dec4.0
    DG_DGR(32, 64, 32, 8, 32)      // dec3.0
    DG_DGR(16, 32, 32, 16, 32)     // dec2.0
    DG_DGR(8, 16, 16, 16, 32)      // dec1.0
#undef DG_DGR
    *handled = false;
    return 0;
}

}  // namespace dg
