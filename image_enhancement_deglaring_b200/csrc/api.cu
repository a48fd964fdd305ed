// C-ABI entry points + the native LightweightUNet orchestrator (see include/deglare.h).
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>

#include <nvtx3/nvToolsExt.h>

#include "common.cuh"

namespace dg {

static thread_local char g_err[512] = "";
static std::atomic<uint64_t> g_launches{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
static std::atomic<int> g_pdl{1};
bool pdl_enabled() { return g_pdl.load(std::memory_order_relaxed) != 0; }
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: launch failed: %s", what, cudaGetErrorString(e));
        return 10;
    }
    return 0;
}

// NVTX ranges per fused kernel of the forward / backward (the reference has no tracing at all).  NVTX v3 is
// header-only and a no-op unless a tool (Nsight Systems / Compute) injects itself; DG_NVTX=0 turns the calls off entirely.
static bool nvtx_on() {
    static const bool on = [] { const char* e = getenv("DG_NVTX"); return !(e && e[0] == '0'); }();
    return on;
}
struct NvtxRange {
    bool live;
    explicit NvtxRange(const char* name) : live(nvtx_on()) { if (live) nvtxRangePushA(name); }
    ~NvtxRange() { if (live) nvtxRangePop(); }
};
static const char* const kConvNames[18] = {"enc1.0", "enc1.3", "enc2.0", "enc2.3", "enc3.0", "enc3.3", "enc4.0", "enc4.3", "bottleneck.0",
                                           "bottleneck.3", "up4+dec4.0", "dec4.3", "up3+dec3.0", "dec3.3", "up2+dec2.0", "dec2.3",
                                           "up1+dec1.0", "dec1.3"};

static int validate_src(const dg_src& s, const char* who) {
    if (s.raw == nullptr) { set_error("%s: null source pointer", who); return 2; }
    if (s.channels < 1) { set_error("%s: bad channel count %d", who, s.channels); return 2; }
    if (s.stats != nullptr) {
        if (s.gamma == nullptr || s.beta == nullptr) { set_error("%s: GroupNorm affine missing", who); return 2; }
        if (s.groups < 1 || s.channels % s.groups != 0) {
            set_error("%s: %d channels not divisible into %d groups", who, s.channels, s.groups);
            return 2;
        }
    }
    if (s.xform == DG_X_CONVT2 && (s.ct_w == nullptr || s.ct_b == nullptr || s.ct_cout < 1)) {
        set_error("%s: ConvTranspose parameters missing", who);
        return 2;
    }
    if (s.xform < DG_X_SAME || s.xform > DG_X_IMAGE_U8) { set_error("%s: bad xform %d", who, s.xform); return 2; }
    return 0;
}

// ---- LightweightUNet plan -----------------------------------------------------------------
struct LwPlan {
    int f[5];
    int conv_c[18], conv_h[18], conv_w[18];
    size_t raw_off[18], stats_off[18], counter_off[18], coef_off[18];
    size_t up_off[4];  // scratch for a materialised upconvK output (deep decoder levels, 16-bit storage)
    size_t stats_bytes, total_bytes;  // stats_bytes = the zero-initialised prefix (statistics + arrival counters)
};

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
static inline size_t dtype_size(int dt) { return dt == DG_F32 ? 4 : 2; }
static inline int block_level(int b) { return b < 5 ? b : 8 - b; }

static int make_plan(const dg_lw_params* p, int N, int H, int W, LwPlan* pl) {
    if (p == nullptr) { set_error("null params"); return 2; }
    if (N < 1 || H < 16 || W < 16 || (H % 16) || (W % 16)) {
        // 500x500 raises at the first torch.cat in the reference (src/model.py:116); here it is an error too
        set_error("input %dx%dx%d: H and W must be positive multiples of 16", N, H, W);
        return 3;
    }
    if (p->dtype < DG_F32 || p->dtype > DG_BF16) { set_error("bad dtype %d", p->dtype); return 2; }
    if (p->features_start < 1 || p->in_channels < 1 || p->out_channels < 1) { set_error("bad channel configuration"); return 2; }
    for (int l = 0; l < 5; ++l) pl->f[l] = p->features_start << l;
    size_t off = 0;
    for (int i = 0; i < 18; ++i) {
        const int lvl = block_level(i / 2);
        pl->conv_c[i] = pl->f[lvl];
        pl->conv_h[i] = H >> lvl;
        pl->conv_w[i] = W >> lvl;
        pl->stats_off[i] = off;
        off += (size_t)N * pl->conv_c[i] * 2 * sizeof(double);
    }
    for (int i = 0; i < 18; ++i) {
        pl->counter_off[i] = off;
        off += align_up((size_t)N * sizeof(int32_t), 8);
    }
    pl->stats_bytes = off;
    off = align_up(off, 256);
    for (int i = 0; i < 18; ++i) {
        pl->coef_off[i] = off;
        off += align_up((size_t)N * pl->conv_c[i] * 2 * sizeof(float), 256);
    }
    for (int i = 0; i < 18; ++i) {
        pl->raw_off[i] = off;
        off += align_up((size_t)N * pl->conv_h[i] * pl->conv_w[i] * pl->conv_c[i] * dtype_size(p->dtype), 256);
    }
    for (int u = 0; u < 4; ++u) {
        const int lvl = 3 - u;
        pl->up_off[u] = off;
        if (p->dtype != DG_F32)   // all four: inference materialises the deep ones, the tensor-core wgrad wants every `up`
            off += align_up((size_t)N * (H >> lvl) * (W >> lvl) * pl->f[lvl] * dtype_size(p->dtype), 256);
    }
    pl->total_bytes = off;
    return 0;
}

// Decoder levels whose ConvTranspose output is materialised by a stand-alone tensor-core GEMM (16-bit tiers) and read by the
// consuming conv as an identity source: every level with C_up >= 32 -- upconv4 / upconv3 of the shipped model, ALL levels of the
// wide variant (features_start = 64) -- plus upconv2 with path bit 5.  The concat itself is never stored.
static inline bool lw_up_materialised(const dg_lw_params* p, const LwPlan& pl, int u) {
    return pl.f[3 - u] >= 32 || (u == 2 && (p->path & 32));
}

// n0: first image of the sub-batch the descriptor is for (every per-image tensor is contiguous over the batch)
static dg_src gn_src(const dg_lw_params* p, const LwPlan& pl, char* ws, int conv_idx, int xform, int n0 = 0) {
    dg_src s;
    memset(&s, 0, sizeof(s));
    const int b = conv_idx / 2, j = conv_idx % 2;
    const size_t img = (size_t)pl.conv_h[conv_idx] * pl.conv_w[conv_idx] * pl.conv_c[conv_idx] * dtype_size(p->dtype);
    s.raw = ws + pl.raw_off[conv_idx] + (size_t)n0 * img;
    s.stats = reinterpret_cast<const double*>(ws + pl.stats_off[conv_idx]) + (size_t)n0 * pl.conv_c[conv_idx] * 2;
    // path bit 4: producer-side GroupNorm finalisation.  Measured SLOWER on B200 (2.87 vs 2.74 ms/batch-64 forward): the
    // threadfence + ticket keeps every producer CTA resident until its statistics atomics have landed, which costs more
    // than the consumers' double-precision prologue saves.  Kept selectable for experiments; off by default.
    s.coef = (p->path & 16) ? reinterpret_cast<const float*>(ws + pl.coef_off[conv_idx]) + (size_t)n0 * pl.conv_c[conv_idx] * 2 : nullptr;
    s.gamma = p->gn_w[b][j];
    s.beta = p->gn_b[b][j];
    s.channels = pl.conv_c[conv_idx];
    s.groups = p->groups[b];
    s.xform = xform;
    s.silu = 1;
    return s;
}

// One forward over images [n0, n0 + nn) of a batch whose workspace was laid out for the whole batch (make_plan), on `stream`.
// io: bit 0 = x is uint8 (normalised by 1/255 on load), bit 1 = y is uint8 (clip + quantise in the head)
static int lw_forward_range(const dg_lw_params* p, const LwPlan& pl, char* ws, const void* x, void* y, int n0, int nn, int H, int W,
                            const float* target, double* l1_sum, cudaStream_t stream, cudaEvent_t* evs, int io) {
    int rc = 0;
    const size_t esz = dtype_size(p->dtype);
    if (evs) cudaEventRecord(evs[0], stream);
    for (int i = 0; i < 18; ++i) {
        const int b = i / 2;
        NvtxRange range(kConvNames[i]);
        dg_conv3x3_args a;
        memset(&a, 0, sizeof(a));
        a.dtype = p->dtype;
        a.N = nn; a.H = pl.conv_h[i]; a.W = pl.conv_w[i];
        a.cout = pl.conv_c[i];
        a.weight = p->conv_w[b][i % 2];
        a.weight_tc = p->conv_w_tc[b][i % 2];
        a.out = ws + pl.raw_off[i] + (size_t)n0 * a.H * a.W * a.cout * esz;
        a.out_stats = reinterpret_cast<double*>(ws + pl.stats_off[i]) + (size_t)n0 * a.cout * 2;
        a.out_coef = (p->path & 16) ? reinterpret_cast<float*>(ws + pl.coef_off[i]) + (size_t)n0 * a.cout * 2 : nullptr;
        a.out_counter = reinterpret_cast<int32_t*>(ws + pl.counter_off[i]) + n0;
        a.out_gamma = p->gn_w[b][i % 2];
        a.out_beta = p->gn_b[b][i % 2];
        a.out_groups = p->groups[b];
        a.eps = 1e-5f;
        a.path = p->path;
        a.nsrc = 1;
        if (i == 0) {
            memset(&a.src[0], 0, sizeof(dg_src));
            a.src[0].raw = static_cast<const char*>(x) + (size_t)n0 * p->in_channels * H * W * ((io & 1) ? 1 : sizeof(float));
            a.src[0].channels = p->in_channels;
            a.src[0].groups = 1;
            a.src[0].xform = (io & 1) ? DG_X_IMAGE_U8 : DG_X_IMAGE;
        } else if (i % 2 == 1) {
            a.src[0] = gn_src(p, pl, ws, i - 1, DG_X_SAME, n0);
        } else if (b < 5) {
            a.src[0] = gn_src(p, pl, ws, i - 1, DG_X_POOL2, n0);        // pool1..4, src/model.py:107-112
        } else {
            const int lvl = block_level(b), u = b - 5;
            a.src[0] = gn_src(p, pl, ws, i - 1, DG_X_CONVT2, n0);      // upconv4..1, src/model.py:115-127
            a.src[0].ct_w = p->up_w[u];
            a.src[0].ct_b = p->up_b[u];
            a.src[0].ct_w_tc = p->up_w_tc[u];
            a.src[0].ct_cout = pl.f[lvl];
            a.src[1] = gn_src(p, pl, ws, 2 * lvl + 1, DG_X_SAME, n0);  // skip: torch.cat((up, skip), 1)
            a.nsrc = 2;
            a.weight_comp = p->dec_comp[u];
            // deep levels (upconv4, upconv3; upconv2 with path bit 5): run the transposed conv as its own tensor-core GEMM and
            // feed its output as an identity source -- see convt_tc.cu for why this beats fusing there
            // composite weights present for a level the tcgen05 decoder mode covers (C_up >= 16): ConvTranspose, concat and conv run
            // as one low-resolution conv in conv3x3_t5.cu -- no stand-alone ConvTranspose, no `up` tensor
            const bool dec_t5 = p->dec_comp[u] != nullptr && !(p->path & (128 | 1024)) && pl.f[lvl] >= 16;
            const bool unfuse = p->dtype != DG_F32 && (p->path & 3) != 1 && lw_up_materialised(p, pl, u) && !dec_t5;
            if (unfuse) {
                bool handled = false;
                void* up = ws + pl.up_off[u] + (size_t)n0 * a.H * a.W * pl.f[lvl] * esz;
                rc = convt_tc_launch(a.src[0], p->dtype, nn, a.H, a.W, up, 1e-5f, p->path, stream, &handled);
                if (rc) return rc;
                if (handled) {
                    memset(&a.src[0], 0, sizeof(dg_src));
                    a.src[0].raw = up;
                    a.src[0].channels = pl.f[lvl];
                    a.src[0].groups = 1;
                    a.src[0].xform = DG_X_SAME;
                }
            }
        }
        rc = dg_conv3x3_fused(&a, reinterpret_cast<dg_stream_t>(stream));
        if (rc) return rc;
        if (evs) cudaEventRecord(evs[i + 1], stream);
    }
    NvtxRange head_range("head");
    dg_head_args h;
    memset(&h, 0, sizeof(h));
    h.src = gn_src(p, pl, ws, 17, DG_X_SAME, n0);
    h.dtype = p->dtype;
    h.N = nn; h.H = H; h.W = W;
    h.cout = p->out_channels;
    h.weight = p->head_w;
    h.bias = p->head_b;
    const size_t out_img = (size_t)p->out_channels * H * W;
    h.out = static_cast<char*>(y) + (size_t)n0 * out_img * ((io & 2) ? 1 : sizeof(float));
    h.target = target ? target + (size_t)n0 * out_img : nullptr;
    h.l1_sum = l1_sum;
    h.eps = 1e-5f;
    h.out_kind = (io & 2) ? 1 : 0;
    rc = dg_head1x1(&h, reinterpret_cast<dg_stream_t>(stream));
    if (evs) cudaEventRecord(evs[19], stream);
    return rc;
}

// Two halves of a large batch run concurrently on two private streams (forked from / joined to the caller's stream with
// events): images are independent, and the tail of every kernel of one half -- the partial last wave, during which most SMs
// idle until the next kernel may start -- is filled by the other half's kernels.  Measured on B200, batch 64 x 512x512 fp16:
// 2.25 -> 2.09 ms per forward (two streams; four: 2.12); batch 32: 1.20 -> 1.08 ms, batch 16: 0.675 -> 0.599 ms, batch 8: no
// change (hence the default threshold of 16).  Outputs are bit-identical (nothing depends on the batch size).
struct FwdFork {
    static constexpr int MAXF = 4;
    bool ready = false;
    int device = -1;
    cudaStream_t s[MAXF];
    cudaEvent_t fork, join[MAXF];
};
// slot 0: calls on the caller's stream (dg_lw_forward / dg_lw_backward); slots 1..4: the chunks in flight of the host pipeline,
// which must not share fork streams (that would serialise the chunks again)
static FwdFork g_forks[5];
static std::atomic<int> g_split{16};   // minimum batch for the two-stream split; 0 disables (dg_set_batch_split)

static int fork_init(int slot = 0) {
    FwdFork& g_fork = g_forks[slot];
    int dev = 0;
    cudaGetDevice(&dev);
    if (g_fork.ready && g_fork.device == dev) return 0;
    cudaError_t e = cudaSuccess;
    for (int k = 0; k < FwdFork::MAXF && e == cudaSuccess; ++k) e = cudaStreamCreateWithFlags(&g_fork.s[k], cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&g_fork.fork, cudaEventDisableTiming);
    for (int k = 0; k < FwdFork::MAXF && e == cudaSuccess; ++k) e = cudaEventCreateWithFlags(&g_fork.join[k], cudaEventDisableTiming);
    if (e != cudaSuccess) { set_error("forward fork streams: %s", cudaGetErrorString(e)); return 10; }
    g_fork.ready = true;
    g_fork.device = dev;
    return 0;
}

static int lw_forward(const dg_lw_params* p, const void* x, void* y, int N, int H, int W, void* workspace,
                      size_t ws_bytes, const float* target, double* l1_sum, cudaStream_t stream,
                      cudaEvent_t* evs = nullptr, int io = 0, int fork_slot = 0) {
    LwPlan pl;
    int rc = make_plan(p, N, H, W, &pl);
    if (rc) return rc;
    if (workspace == nullptr || ws_bytes < pl.total_bytes) {
        set_error("workspace too small: %zu < %zu", ws_bytes, pl.total_bytes);
        return 4;
    }
    if (x == nullptr || y == nullptr) { set_error("null input/output"); return 2; }
    char* ws = static_cast<char*>(workspace);
    cudaError_t e = cudaMemsetAsync(ws, 0, pl.stats_bytes, stream);
    if (e != cudaSuccess) { set_error("memset: %s", cudaGetErrorString(e)); return 10; }
    const int split = g_split.load();
    if (evs == nullptr && split > 0 && N >= split && N >= 2) {
        if ((rc = fork_init(fork_slot))) return rc;
        FwdFork& F = g_forks[fork_slot];
        cudaEventRecord(F.fork, stream);
        const int half = N / 2;
        for (int k = 0; k < 2; ++k) {
            cudaStreamWaitEvent(F.s[k], F.fork, 0);
            rc = lw_forward_range(p, pl, ws, x, y, k ? half : 0, k ? N - half : half, H, W, target, l1_sum, F.s[k], nullptr, io);
            cudaEventRecord(F.join[k], F.s[k]);
            cudaStreamWaitEvent(stream, F.join[k], 0);   // joined even on error, so the caller's stream stays ordered
            if (rc) return rc;
        }
        return 0;
    }
    return lw_forward_range(p, pl, ws, x, y, 0, N, H, W, target, l1_sum, stream, evs, io);
}

// ---- LightweightUNet backward ------------------------------------------------------------------
// Flat gradient layout = the module's parameters() order and the parameters' own memory layouts:
//   per block: conv0.weight [Co][Ci][3][3], gn0.weight, gn0.bias, conv1.weight, gn1.weight, gn1.bias
//   order: enc1..enc4, bottleneck, (upconvK.weight [Ci][Co][2][2], upconvK.bias, decK block) for K=4..1, head w, head b
struct GradLayout {
    size_t conv_w[9][2], gn_w[9][2], gn_b[9][2], up_w[4], up_b[4], head_w, head_b, total;
};

static void make_grad_layout(const dg_lw_params* p, const LwPlan& pl, GradLayout* g) {
    size_t off = 0;
    auto block = [&](int b) {
        const int lvl = block_level(b);
        const int c = pl.f[lvl];
        const int cin0 = b == 0 ? p->in_channels : (b < 5 ? pl.f[lvl - 1] : 2 * c);
        g->conv_w[b][0] = off; off += (size_t)c * cin0 * 9;
        g->gn_w[b][0] = off; off += c;
        g->gn_b[b][0] = off; off += c;
        g->conv_w[b][1] = off; off += (size_t)c * c * 9;
        g->gn_w[b][1] = off; off += c;
        g->gn_b[b][1] = off; off += c;
    };
    for (int b = 0; b < 5; ++b) block(b);
    for (int b = 5; b < 9; ++b) {
        const int lvl = block_level(b), u = b - 5;
        g->up_w[u] = off; off += (size_t)pl.f[lvl + 1] * pl.f[lvl] * 4;
        g->up_b[u] = off; off += pl.f[lvl];
        block(b);
    }
    g->head_w = off; off += (size_t)p->out_channels * pl.f[0];
    g->head_b = off; off += p->out_channels;
    g->total = off;
}

struct BwdPlan {
    size_t p_off[18], g_off[18], t_off[18], gb_off[18], low_off[4], coef_off, p_bytes, total;   // gb: dR as bf16 (16-bit tiers)
    int cin_tot[18], maxc;
};

static void make_bwd_plan(const dg_lw_params* p, const LwPlan& pl, int N, BwdPlan* bp) {
    size_t off = 0;
    for (int i = 0; i < 18; ++i) { bp->p_off[i] = off; off += (size_t)N * pl.conv_c[i] * 2 * sizeof(double); }
    bp->p_bytes = off;
    off = align_up(off, 256);
    int maxc = 0;
    for (int i = 0; i < 18; ++i) {
        const int b = i / 2, lvl = block_level(b);
        bp->cin_tot[i] = i == 0 ? p->in_channels : (i % 2 ? pl.conv_c[i] : (b < 5 ? pl.f[lvl - 1] : 2 * pl.f[lvl]));
        bp->g_off[i] = off;
        off += align_up((size_t)N * pl.conv_h[i] * pl.conv_w[i] * pl.conv_c[i] * sizeof(float), 256);
        bp->t_off[i] = off;
        if (i > 0) off += align_up((size_t)N * pl.conv_h[i] * pl.conv_w[i] * bp->cin_tot[i] * sizeof(float), 256);
        if (pl.conv_c[i] > maxc) maxc = pl.conv_c[i];
    }
    for (int u = 0; u < 4; ++u) {
        const int lvl = 3 - u;  // upconv4 produces level 3 from level 4
        bp->low_off[u] = off;
        off += align_up((size_t)N * (pl.conv_h[2 * lvl + 2]) * (pl.conv_w[2 * lvl + 2]) * pl.f[lvl + 1] * sizeof(float), 256);
    }
    for (int i = 0; i < 18; ++i) {
        bp->gb_off[i] = off;
        if (i > 0 && p->dtype != DG_F32) off += align_up((size_t)N * pl.conv_h[i] * pl.conv_w[i] * pl.conv_c[i] * 2, 256);
    }
    bp->coef_off = off;
    bp->maxc = maxc;
    off += align_up((size_t)N * maxc * 2 * sizeof(float), 256);
    bp->total = off;
}

// the forward conv i's argument block for images [n0, n0 + nn) (x = the WHOLE batch's input)
static void fwd_conv_args(const dg_lw_params* p, const LwPlan& pl, char* ws, const float* x, int nn, int i, dg_conv3x3_args* a,
                          int n0 = 0) {
    const int b = i / 2;
    memset(a, 0, sizeof(*a));
    a->dtype = p->dtype;
    a->N = nn; a->H = pl.conv_h[i]; a->W = pl.conv_w[i];
    a->cout = pl.conv_c[i];
    a->weight = p->conv_w[b][i % 2];
    a->weight_tc = p->conv_w_tc[b][i % 2];
    a->out = ws + pl.raw_off[i] + (size_t)n0 * a->H * a->W * a->cout * dtype_size(p->dtype);
    a->out_stats = reinterpret_cast<double*>(ws + pl.stats_off[i]) + (size_t)n0 * a->cout * 2;
    a->eps = 1e-5f;
    a->path = p->path;
    a->nsrc = 1;
    if (i == 0) {
        a->src[0].raw = x + (size_t)n0 * p->in_channels * pl.conv_h[0] * pl.conv_w[0];
        a->src[0].channels = p->in_channels;
        a->src[0].groups = 1;
        a->src[0].xform = DG_X_IMAGE;
    } else if (i % 2 == 1) {
        a->src[0] = gn_src(p, pl, ws, i - 1, DG_X_SAME, n0);
    } else if (b < 5) {
        a->src[0] = gn_src(p, pl, ws, i - 1, DG_X_POOL2, n0);       // pool1..4, src/model.py:107-112
    } else {
        const int lvl = block_level(b), u = b - 5;
        a->src[0] = gn_src(p, pl, ws, i - 1, DG_X_CONVT2, n0);     // upconv4..1, src/model.py:115-127
        a->src[0].ct_w = p->up_w[u];
        a->src[0].ct_b = p->up_b[u];
        a->src[0].ct_w_tc = p->up_w_tc[u];
        a->src[0].ct_cout = pl.f[lvl];
        a->src[1] = gn_src(p, pl, ws, 2 * lvl + 1, DG_X_SAME, n0); // skip: torch.cat((up, skip), 1)
        a->nsrc = 2;
    }
}

// nn.L1Loss backward fused into the head backward: forward output, target, device scalar dLoss (NULL = 1) and 1 / numel
struct L1Seed { const float* y; const float* target; const float* scale; float inv_numel; };

static int lw_backward_range(const dg_lw_params* p, const struct LwPlan& pl, const struct BwdPlan& bp, const struct GradLayout& gl,
                             const float* x, const float* grad_y, int n0, int N, int H, int W, char* fw, char* bw, float* grads,
                             cudaStream_t st, const L1Seed* l1);

static int lw_backward(const dg_lw_params* p, const float* x, const float* grad_y, int N, int H, int W, void* fwd_ws,
                       size_t fwd_bytes, void* bwd_ws, size_t bwd_bytes, float* grads, cudaStream_t st, const L1Seed* l1 = nullptr) {
    LwPlan pl;
    int rc = make_plan(p, N, H, W, &pl);
    if (rc) return rc;
    BwdPlan bp;
    make_bwd_plan(p, pl, N, &bp);
    GradLayout gl;
    make_grad_layout(p, pl, &gl);
    if (fwd_ws == nullptr || fwd_bytes < pl.total_bytes || bwd_ws == nullptr || bwd_bytes < bp.total) {
        set_error("backward: workspace too small (fwd %zu/%zu, bwd %zu/%zu)", fwd_bytes, pl.total_bytes, bwd_bytes, bp.total);
        return 4;
    }
    if (x == nullptr || (grad_y == nullptr && l1 == nullptr) || grads == nullptr) { set_error("backward: null pointer"); return 2; }
    // conv_w_flip / up_w_t feed the CUDA-core data-gradient kernels only: required where a tensor-core packing is absent
    for (int b = 0; b < 9; ++b)
        for (int j = 0; j < 2; ++j)
            if ((2 * b + j) > 0 && p->conv_w_flip[b][j] == nullptr && p->conv_w_tc_bf16[b][j] == nullptr) {
                set_error("backward: neither flipped weights nor a bf16 tensor-core packing for conv %d", 2 * b + j);
                return 2;
            }
    for (int u = 0; u < 4; ++u)
        if (p->up_w_t[u] == nullptr && p->up_w_tc_bf16[u] == nullptr) { set_error("backward: transposed ConvTranspose weights missing"); return 2; }
    char* fw = static_cast<char*>(fwd_ws);
    char* bw = static_cast<char*>(bwd_ws);
    cudaError_t e = cudaMemsetAsync(bw, 0, bp.p_bytes, st);
    if (e == cudaSuccess) e = cudaMemsetAsync(grads, 0, gl.total * sizeof(float), st);
    if (e != cudaSuccess) { set_error("memset: %s", cudaGetErrorString(e)); return 10; }
    const int split = g_split.load();
    if (split > 0 && N >= split && N >= 2) {   // concurrent sub-batches, as in lw_forward (DG_BWD_FORKS = 2..4 of them; default 2)
        if ((rc = fork_init())) return rc;
        FwdFork& F = g_forks[0];
        static const int want = [] { const char* e = getenv("DG_BWD_FORKS"); const int v = e ? atoi(e) : 2; return v < 2 ? 2 : (v > FwdFork::MAXF ? FwdFork::MAXF : v); }();
        const int nf = N >= 8 * want ? want : 2;
        cudaEventRecord(F.fork, st);
        int n0 = 0;
        for (int k = 0; k < nf; ++k) {
            const int nn = (N - n0) / (nf - k);
            cudaStreamWaitEvent(F.s[k], F.fork, 0);
            rc = lw_backward_range(p, pl, bp, gl, x, grad_y, n0, nn, H, W, fw, bw, grads, F.s[k], l1);
            cudaEventRecord(F.join[k], F.s[k]);
            cudaStreamWaitEvent(st, F.join[k], 0);
            n0 += nn;
            if (rc) return rc;
        }
        return 0;
    }
    return lw_backward_range(p, pl, bp, gl, x, grad_y, 0, N, H, W, fw, bw, grads, st, l1);
}

// Backward over images [n0, n0 + N) of a batch of N_total; every buffer was laid out for the whole batch (`N` below is the
// sub-batch size).  Parameter gradients are accumulated atomically, so concurrent sub-batches add up.
static int lw_backward_range(const dg_lw_params* p, const LwPlan& pl, const BwdPlan& bp, const GradLayout& gl, const float* x,
                             const float* grad_y, int n0, int N, int H, int W, char* fw, char* bw, float* grads, cudaStream_t st,
                             const L1Seed* l1) {
    int rc = 0;
    const size_t esz = dtype_size(p->dtype);
    bool t_split[18] = {};   // decoder conv i: T(i) holds the two halves of the concat gradient as two compact tensors
    auto hwc = [&](int i, int c) { return (size_t)pl.conv_h[i] * pl.conv_w[i] * c; };
    auto G = [&](int i) { return reinterpret_cast<float*>(bw + bp.g_off[i]) + (size_t)n0 * hwc(i, pl.conv_c[i]); };
    auto T = [&](int i) { return reinterpret_cast<float*>(bw + bp.t_off[i]) + (size_t)n0 * hwc(i, bp.cin_tot[i]); };
    auto P = [&](int i) { return reinterpret_cast<double*>(bw + bp.p_off[i]) + (size_t)n0 * pl.conv_c[i] * 2; };
    auto raw = [&](int i) { return static_cast<const void*>(fw + pl.raw_off[i] + (size_t)n0 * hwc(i, pl.conv_c[i]) * esz); };
    auto stats = [&](int i) { return reinterpret_cast<const double*>(fw + pl.stats_off[i]) + (size_t)n0 * pl.conv_c[i] * 2; };
    auto act_bwd = [&](int j, const float* da, int sa, int oa, const float* db, int sb, int ob) {
        return act_bwd_launch(p->dtype, raw(j), stats(j), p->gn_w[j / 2][j % 2], p->gn_b[j / 2][j % 2], da, sa, oa, db, sb, ob, G(j),
                              P(j), N, pl.conv_h[j], pl.conv_w[j], pl.conv_c[j], p->groups[j / 2], 1e-5f, st);
    };
    // head: src/model.py:131 backward
    const size_t o0 = (size_t)n0 * p->out_channels * H * W;
    rc = head_bwd_launch(p->dtype, raw(17), stats(17), p->gn_w[8][1], p->gn_b[8][1],
                         grad_y ? grad_y + o0 : nullptr, p->head_w, G(17), P(17),
                         grads + gl.head_w, grads + gl.head_b, N, H, W, pl.conv_c[17], p->out_channels, p->groups[8], 1e-5f, st,
                         l1 ? l1->y + o0 : nullptr, l1 ? l1->target + o0 : nullptr, l1 ? l1->scale : nullptr, l1 ? l1->inv_numel : 0.f);
    if (rc) return rc;
    for (int i = 17; i >= 0; --i) {
        const int b = i / 2, j = i % 2, C = pl.conv_c[i], Hi = pl.conv_h[i], Wi = pl.conv_w[i];
        char rname[32];
        snprintf(rname, sizeof(rname), "bwd %s", kConvNames[i]);
        NvtxRange range(rname);
        // dW_i: same sources as the forward conv, correlated with dR_i; written in the parameter's [Co][Ci][3][3] layout
        dg_conv3x3_args a;
        fwd_conv_args(p, pl, fw, x, N, i, &a, n0);
        // G_i -> dR_i, dgamma / dbeta.  Where BOTH consumers of dR_i are tensor-core kernels (they round it to bf16 anyway) it is
        // written once as bf16 instead of fp32 in place: they read half the bytes and copy the tile instead of converting it.
        void* dRb = nullptr;
        if (i > 0 && p->dtype != DG_F32 && (p->path & 3) != 1 && p->conv_w_tc_bf16[b][j] != nullptr) {
            dg_conv3x3_args probe = a;
            if (probe.nsrc == 2) {   // the wgrad kernel sees a decoder conv with its ConvTranspose source materialised
                memset(&probe.src[0], 0, sizeof(dg_src));
                probe.src[0].raw = fw + pl.up_off[b - 5];
                probe.src[0].channels = pl.f[block_level(b)];
                probe.src[0].groups = 1;
                probe.src[0].xform = DG_X_SAME;
            }
            bool wg_ok = false, dgr_ok = false;
            conv3x3_wgrad_tc_launch(probe, G(i), grads + gl.conv_w[b][j], 1, 9, 9 * bp.cin_tot[i], st, &wg_ok, nullptr, /*dry*/ true);
            conv3x3_dgrad_tc_launch(G(i), p->conv_w_tc_bf16[b][j], G(i), N, Hi, Wi, C, bp.cin_tot[i], st, &dgr_ok, nullptr, nullptr, true);
            if (wg_ok && dgr_ok) dRb = bw + bp.gb_off[i] + (size_t)n0 * hwc(i, C) * 2;
        }
        bool dr_bf16 = false;
        rc = gn_bwd_apply_launch(p->dtype, raw(i), stats(i), p->gn_w[b][j], P(i), G(i), grads + gl.gn_w[b][j], grads + gl.gn_b[b][j],
                                 N, Hi, Wi, C, p->groups[b], 1e-5f, st, dRb, &dr_bf16);
        if (rc) return rc;
        const void* dRb_in = dr_bf16 ? dRb : nullptr;   // when set, G(i) still holds G, NOT dR: only the bf16 copy is valid
        bool wg_done = false;
        if (p->dtype != DG_F32 && (p->path & 3) != 1 && i > 0) {
            // tensor-core wgrad (wgrad_tc.cu); a decoder conv reads the MATERIALISED ConvTranspose output: the forward left it
            // in the workspace for the deep levels, for the fused levels it is rebuilt here by the same stand-alone kernel
            bool ok = true;
            if (a.nsrc == 2) {
                const int lvl = block_level(b), u = b - 5;
                void* up = fw + pl.up_off[u] + (size_t)n0 * Hi * Wi * pl.f[lvl] * esz;
                const bool have_up = lw_up_materialised(p, pl, u);
                if (!have_up) {
                    rc = convt_tc_launch(a.src[0], p->dtype, N, Hi, Wi, up, 1e-5f, p->path, st, &ok);
                    if (rc) return rc;
                }
                if (ok) {
                    memset(&a.src[0], 0, sizeof(dg_src));
                    a.src[0].raw = up;
                    a.src[0].channels = pl.f[lvl];
                    a.src[0].groups = 1;
                    a.src[0].xform = DG_X_SAME;
                }
            }
            if (ok) {
                rc = conv3x3_wgrad_tc_launch(a, G(i), grads + gl.conv_w[b][j], 1, 9, 9 * bp.cin_tot[i], st, &wg_done, dRb_in);
                if (rc) return rc;
            }
            if (!wg_done) fwd_conv_args(p, pl, fw, x, N, i, &a, n0);  // restore the fused description for the generic kernel
        }
        if (i == 0 && (p->path & 3) != 1 && p->in_channels == 1) {   // first conv: dedicated streaming kernel (backward.cu)
            rc = first_wgrad_launch(x + (size_t)n0 * p->in_channels * H * W, G(0), grads + gl.conv_w[0][0], N, Hi, Wi, C, st, &wg_done);
            if (rc) return rc;
        }
        if (!wg_done && dRb_in != nullptr) { set_error("backward: conv %d wgrad probe / launch mismatch", i); return 11; }
        if (!wg_done) {
            rc = conv3x3_wgrad_launch(a, G(i), grads + gl.conv_w[b][j], /*tap*/ 1, /*ci*/ 9, /*co*/ 9 * bp.cin_tot[i], st);
            if (rc) return rc;
        }
        if (i == 0) break;
        // dA_in = conv3x3(dR_i, flipped weights): tensor cores where covered (dgrad_tc.cu) ...
        bool dg_done = false;
        bool act_fused = false;
        if (p->dtype != DG_F32 && (p->path & 3) != 1 && p->conv_w_tc_bf16[b][j] != nullptr) {
            if (j == 1) {
                // the second conv of a block: its input is the activated output of the first and nothing else consumes that, so
                // the data gradient goes straight through the producer's SiLU' / statistics sums into G(i-1), P(i-1)
                DgradAct act{raw(i - 1), stats(i - 1), p->gn_w[b][0], p->gn_b[b][0], P(i - 1), p->groups[b], p->dtype, 1e-5f};
                rc = conv3x3_dgrad_tc_launch(G(i), p->conv_w_tc_bf16[b][j], G(i - 1), N, Hi, Wi, C, bp.cin_tot[i], st, &dg_done, &act, dRb_in);
                if (rc) return rc;
                act_fused = dg_done;
            }
            if (!dg_done && j == 0 && b >= 5) {
                // decoder conv: its input gradient is a concat gradient whose halves go to different consumers -- written as two
                // compact tensors (up half at T(i), skip half right behind it) instead of one interleaved [.., 2C]
                rc = conv3x3_dgrad_tc_launch(G(i), p->conv_w_tc_bf16[b][j], T(i), N, Hi, Wi, C, bp.cin_tot[i], st, &dg_done, nullptr, dRb_in,
                                             false, T(i) + (size_t)N * hwc(i, C));
                if (rc) return rc;
                t_split[i] = dg_done;
            }
            if (!dg_done) {
                rc = conv3x3_dgrad_tc_launch(G(i), p->conv_w_tc_bf16[b][j], T(i), N, Hi, Wi, C, bp.cin_tot[i], st, &dg_done, nullptr, dRb_in);
                if (rc) return rc;
            }
        }
        // ... the tcgen05 kernel where the mma.sync data gradient has no instance (wider variants): dR rounded to bf16, the
        // taps-flipped weights in the tensor-core packing, fp32 result (conv3x3_t5.cu, T5_IDENT) ...
        if (!dg_done && dRb_in == nullptr && p->dtype != DG_F32 && (p->path & 3) != 1 && p->conv_w_flip_tc_bf16[b][j] != nullptr) {
            rc = conv3x3_dgrad_t5_launch(G(i), nullptr, bw + bp.gb_off[i] + (size_t)n0 * hwc(i, C) * 2, p->conv_w_flip_tc_bf16[b][j], T(i),
                                         N, Hi, Wi, C, bp.cin_tot[i], st, &dg_done);
            if (rc) return rc;
        }
        // ... else the forward generic kernel on an identity fp32 source
        dg_conv3x3_args d;
        memset(&d, 0, sizeof(d));
        d.dtype = DG_F32;
        d.N = N; d.H = Hi; d.W = Wi;
        d.cout = bp.cin_tot[i];
        d.weight = p->conv_w_flip[b][j];
        d.out = T(i);
        d.eps = 1e-5f;
        d.path = 1;
        d.nsrc = 1;
        d.src[0].raw = G(i);
        d.src[0].channels = C;
        d.src[0].groups = 1;
        d.src[0].xform = DG_X_SAME;
        if (!dg_done && dRb_in != nullptr) { set_error("backward: conv %d dgrad probe / launch mismatch", i); return 11; }
        if (!dg_done) {
            if (d.weight == nullptr) { set_error("backward: conv %d needs the CUDA-core dgrad but conv_w_flip is NULL", i); return 2; }
            rc = conv3x3_generic_launch(d, st);
            if (rc) return rc;
        }
        if (j == 1) {
            if (!act_fused) rc = act_bwd(i - 1, T(i), C, 0, nullptr, 0, 0);
        } else if (b < 5) {
            // producer = skip tensor of level b-1: gradient from the decoder's concat (skip half) + this pooled path
            const int lvl = b - 1, dconv = 2 * (8 - lvl);
            if (t_split[dconv]) rc = act_bwd(i - 1, T(dconv) + (size_t)N * hwc(dconv, pl.f[lvl]), pl.f[lvl], 0, T(i), pl.f[lvl], 0);
            else rc = act_bwd(i - 1, T(dconv), 2 * pl.f[lvl], pl.f[lvl], T(i), pl.f[lvl], 0);
        } else {
            const int lvl = block_level(b), u = b - 5;
            float* dlow = reinterpret_cast<float*>(bw + bp.low_off[u]) + (size_t)n0 * (Hi / 2) * (Wi / 2) * pl.f[lvl + 1];
            // the low-resolution producer feeds nothing but this ConvTranspose: with the tensor-core data gradient its activation
            // backward is fused into that kernel's epilogue, which then writes G(i-1) / P(i-1) directly
            const bool try_fuse = p->dtype != DG_F32 && (p->path & 3) != 1 && p->up_w_tc_bf16[u] != nullptr;
            DgradAct lact{raw(i - 1), stats(i - 1), p->gn_w[b - 1][1], p->gn_b[b - 1][1], P(i - 1), p->groups[b - 1], p->dtype, 1e-5f};
            bool low_fused = false;
            rc = convt_bwd_launch(p->dtype, T(i), t_split[i] ? pl.f[lvl] : 2 * pl.f[lvl], p->up_w[u], p->up_w_t[u], raw(i - 1), stats(i - 1), p->gn_w[b - 1][1],
                                  p->gn_b[b - 1][1], try_fuse ? G(i - 1) : dlow, grads + gl.up_w[u], grads + gl.up_b[u],
                                  reinterpret_cast<float*>(bw + bp.coef_off) + (size_t)n0 * bp.maxc * 2, N, Hi, Wi, pl.f[lvl + 1], pl.f[lvl],
                                  p->groups[b - 1], 1e-5f, st, (p->path & 3) != 1 ? p->up_w_tc_bf16[u] : nullptr,
                                  try_fuse ? &lact : nullptr, try_fuse ? &low_fused : nullptr,
                                  (p->path & 3) != 1 ? p->up_w_dgrad_tc_bf16[u] : nullptr,
                                  p->dtype != DG_F32 ? bw + bp.gb_off[i] + (size_t)n0 * hwc(i, C) * 2 : nullptr);   // conv i's bf16 dR scratch is free again
            if (rc) return rc;
            if (low_fused) continue;   // G(i-1), P(i-1) are done
            if (try_fuse) {            // the kernel declined: G(i-1) holds the plain gradient; act_bwd works in place on it
                rc = act_bwd(i - 1, G(i - 1), pl.f[lvl + 1], 0, nullptr, 0, 0);
                if (rc) return rc;
                continue;
            }
            rc = act_bwd(i - 1, dlow, pl.f[lvl + 1], 0, nullptr, 0, 0);
        }
        if (rc) return rc;
    }
    return 0;
}

// ---- host-buffer pipeline state --------------------------------------------------------------
struct HostPipe {
    bool ready = false;
    cudaEvent_t start;   // recorded on the CALLER's stream: weight packing etc. enqueued there is ordered before the pipeline
    static constexpr int NS = 4;  // chunks in flight (measured: 4 x 8-image chunks 2.84 ms per 64 images, 8 in flight 2.98, 2 in flight 3.3): compute streams, staging buffers and workspaces
    cudaStream_t s_in, s_cmp[NS], s_out;
    cudaEvent_t in_done[NS], cmp_done[NS], out_done[NS];
    // asynchronous submissions (dg_lw_infer_host_submit / _wait): the chunk counter runs ACROSS calls, so the first chunks of
    // call k+1 are copied in and computed while the last chunks of call k are still computing / being copied out
    static constexpr int NT = 8;  // tickets (calls) that may be outstanding
    unsigned long long seq = 0, tickets = 0;
    cudaEvent_t ticket_done[NT];
    const void* ws_in_flight = nullptr;   // scratch / geometry of the calls in flight: a different one drains the pipeline first
    int chunk_in_flight = 0, h_in_flight = 0, w_in_flight = 0;
};
static HostPipe g_pipes[16];   // one pipeline (streams, events) per device ordinal
static std::mutex g_pipe_mutex;  // the *_host entry points serialise: they share the pipeline and the caller's scratch

static int pipe_init(HostPipe** out) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 16) { set_error("infer_host: device ordinal %d out of range", dev); return 2; }
    HostPipe& g_pipe = g_pipes[dev];
    *out = &g_pipe;
    if (g_pipe.ready) return 0;
    cudaError_t e;
    if ((e = cudaStreamCreateWithFlags(&g_pipe.s_in, cudaStreamNonBlocking)) != cudaSuccess ||
        (e = cudaStreamCreateWithFlags(&g_pipe.s_out, cudaStreamNonBlocking)) != cudaSuccess) {
        set_error("stream create: %s", cudaGetErrorString(e));
        return 10;
    }
    for (int i = 0; i < HostPipe::NS; ++i) {
        if ((e = cudaStreamCreateWithFlags(&g_pipe.s_cmp[i], cudaStreamNonBlocking)) != cudaSuccess) {
            set_error("stream create: %s", cudaGetErrorString(e));
            return 10;
        }
        cudaEventCreateWithFlags(&g_pipe.in_done[i], cudaEventDisableTiming);
        cudaEventCreateWithFlags(&g_pipe.cmp_done[i], cudaEventDisableTiming);
        cudaEventCreateWithFlags(&g_pipe.out_done[i], cudaEventDisableTiming);
    }
    if ((e = cudaEventCreateWithFlags(&g_pipe.start, cudaEventDisableTiming)) != cudaSuccess) {
        set_error("event create: %s", cudaGetErrorString(e));
        return 10;
    }
    for (int i = 0; i < HostPipe::NT; ++i)
        if ((e = cudaEventCreateWithFlags(&g_pipe.ticket_done[i], cudaEventDisableTiming | cudaEventBlockingSync)) != cudaSuccess) {
            set_error("event create: %s", cudaGetErrorString(e));
            return 10;
        }
    g_pipe.ready = true;
    return 0;
}

}  // namespace dg

using namespace dg;

extern "C" {

const char* dg_last_error_string(void) { return g_err; }
int dg_version(void) { return 100; }
int dg_set_batch_split(int min_batch) {
    const int old = g_split.load();
    g_split.store(min_batch < 0 ? 0 : min_batch);
    return old;
}
int dg_set_pdl(int enabled) {
    const int old = g_pdl.exchange(enabled ? 1 : 0);
    return old;
}
uint64_t dg_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

int dg_conv3x3_fused(const dg_conv3x3_args* a, dg_stream_t stream) {
    if (a == nullptr) { set_error("conv3x3: null args"); return 2; }
    if (a->nsrc < 1 || a->nsrc > 2) { set_error("conv3x3: nsrc %d", a->nsrc); return 2; }
    if (a->N < 1 || a->H < 1 || a->W < 1 || a->cout < 1) { set_error("conv3x3: bad shape"); return 3; }
    if (a->weight == nullptr || a->out == nullptr) { set_error("conv3x3: null weight/out"); return 2; }
    for (int s = 0; s < a->nsrc; ++s) {
        int rc = validate_src(a->src[s], "conv3x3");
        if (rc) return rc;
        const int xf = a->src[s].xform;
        if ((xf == DG_X_UP2 || xf == DG_X_CONVT2) && ((a->H | a->W) & 1)) {
            set_error("conv3x3: up-sampled source needs even H, W (got %dx%d)", a->H, a->W);
            return 3;
        }
    }
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if ((a->path & 3) != 1) {
        bool handled = false;
        int rc = conv_first_tc_launch(*a, st, &handled);
        if (rc) return rc;
        if (handled) return 0;
        rc = conv3x3_t5_launch(*a, st, &handled);     // warp-specialised tcgen05 + TMEM kernel: every C_out >= 32 layer
        if (rc) return rc;
        if (handled) return 0;
        rc = conv3x3_umma_launch(*a, st, &handled);   // round-1 tcgen05 kernel (opt-in, path bit 6)
        if (rc) return rc;
        if (handled) return 0;
        rc = conv3x3_dec_launch(*a, st, &handled);    // upconv1 + dec1.0 with the ConvTranspose folded into the conv taps
        if (rc) return rc;
        if (handled) return 0;
        rc = conv3x3_ring_launch(*a, st, &handled);   // persistent TMA-fed kernel for the 8 -> 8 full-resolution layers
        if (rc) return rc;
        if (handled) return 0;
        rc = conv3x3_tc_launch(*a, st, &handled);
        if (rc) return rc;
        if (handled) return 0;
        if ((a->path & 3) == 2) { set_error("conv3x3: tensor-core path does not cover this configuration"); return 3; }
    }
    return conv3x3_generic_launch(*a, st);
}

int dg_conv3x3_wgrad(const dg_conv3x3_args* a, const float* dR, float* dW, int32_t s_tap, int32_t s_ci, int32_t s_co,
                     dg_stream_t stream) {
    if (a == nullptr || dR == nullptr || dW == nullptr) { set_error("wgrad: null pointer"); return 2; }
    if (a->nsrc < 1 || a->nsrc > 2) { set_error("wgrad: nsrc %d", a->nsrc); return 2; }
    for (int s = 0; s < a->nsrc; ++s) {
        int rc = validate_src(a->src[s], "wgrad");
        if (rc) return rc;
    }
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if ((a->path & 3) != 1) {
        bool handled = false;
        int rc = conv3x3_wgrad_tc_launch(*a, dR, dW, s_tap, s_ci, s_co, st, &handled);
        if (rc || handled) return rc;
        if ((a->path & 3) == 2) { set_error("wgrad: no tensor-core kernel for this configuration"); return 3; }
    }
    return conv3x3_wgrad_launch(*a, dR, dW, s_tap, s_ci, s_co, st);
}

int dg_conv3x3_dgrad(const float* dR, const void* weight_tc_bf16, float* dX, int32_t N, int32_t H, int32_t W, int32_t cin,
                     int32_t cout, dg_stream_t stream) {
    if (dR == nullptr || weight_tc_bf16 == nullptr || dX == nullptr) { set_error("dgrad: null pointer"); return 2; }
    bool handled = false;
    int rc = conv3x3_dgrad_tc_launch(dR, weight_tc_bf16, dX, N, H, W, cout, cin, reinterpret_cast<cudaStream_t>(stream), &handled);
    if (rc == 0 && !handled) { set_error("dgrad: no tensor-core kernel for %d -> %d channels (or unaligned pointers)", cin, cout); return 3; }
    return rc;
}

int dg_convt2x2_dgrad(const float* dCat, int32_t stride, const void* ct_w_tc_bf16, float* dLow, int32_t N, int32_t H, int32_t W,
                      int32_t cl, int32_t cu, dg_stream_t stream) {
    if (dCat == nullptr || ct_w_tc_bf16 == nullptr || dLow == nullptr) { set_error("convT dgrad: null pointer"); return 2; }
    bool handled = false;
    int rc = convt_dgrad_tc_launch(dCat, stride, ct_w_tc_bf16, dLow, N, H, W, cl, cu, reinterpret_cast<cudaStream_t>(stream), &handled);
    if (rc == 0 && !handled) { set_error("convT dgrad: no tensor-core kernel for %d -> %d channels / this shape", cl, cu); return 3; }
    return rc;
}

int dg_conv3x3_dgrad_wide(const float* dR, const void* weight_flip_tc_bf16, float* dX, void* scratch_bf16, int32_t N, int32_t H,
                          int32_t W, int32_t cin, int32_t cout, dg_stream_t stream) {
    if (dR == nullptr || weight_flip_tc_bf16 == nullptr || dX == nullptr || scratch_bf16 == nullptr) { set_error("wide dgrad: null pointer"); return 2; }
    if (N < 1 || H < 1 || W < 1) { set_error("wide dgrad: bad shape"); return 3; }
    bool handled = false;
    int rc = conv3x3_dgrad_t5_launch(dR, nullptr, scratch_bf16, weight_flip_tc_bf16, dX, N, H, W, cout, cin,
                                     reinterpret_cast<cudaStream_t>(stream), &handled);
    if (rc == 0 && !handled) { set_error("wide dgrad: no tcgen05 plan for %d -> %d channels at %dx%d (or unaligned pointers)", cin, cout, H, W); return 3; }
    return rc;
}

int dg_image_metrics(const float* output, const float* target, int32_t N, int32_t H, int32_t W, int32_t clip01, double data_range,
                     double* acc, dg_stream_t stream) {
    if (output == nullptr || target == nullptr || acc == nullptr) { set_error("metrics: null pointer"); return 2; }
    return image_metrics_launch(output, target, N, H, W, clip01, data_range, acc, reinterpret_cast<cudaStream_t>(stream));
}

int dg_head1x1(const dg_head_args* a, dg_stream_t stream) {
    if (a == nullptr) { set_error("head: null args"); return 2; }
    int rc = validate_src(a->src, "head");
    if (rc) return rc;
    if (a->src.xform != DG_X_SAME) { set_error("head: source must be same-resolution"); return 3; }
    if (a->weight == nullptr || a->bias == nullptr || a->out == nullptr) { set_error("head: null pointer"); return 2; }
    return head_launch(*a, reinterpret_cast<cudaStream_t>(stream));
}

int dg_tc_conv3x3_bytes(int32_t cin, int32_t cout, size_t* bytes) {
    if (bytes == nullptr) { set_error("null bytes"); return 2; }
    return tc_conv3x3_bytes(cin, cout, bytes);
}
int dg_pack_conv3x3_tc(const float* w, void* out, int32_t cin, int32_t cout, int32_t dtype, dg_stream_t stream) {
    if (w == nullptr || out == nullptr) { set_error("pack: null pointer"); return 2; }
    return pack_conv3x3_tc(w, out, cin, cout, dtype, reinterpret_cast<cudaStream_t>(stream));
}
int dg_tc_convt2x2_bytes(int32_t cin, int32_t cout, size_t* bytes) {
    if (bytes == nullptr) { set_error("null bytes"); return 2; }
    return tc_convt_bytes(cin, cout, bytes);
}
int dg_pack_convt2x2_tc(const float* w, void* out, int32_t cin, int32_t cout, int32_t dtype, dg_stream_t stream) {
    if (w == nullptr || out == nullptr) { set_error("pack: null pointer"); return 2; }
    return pack_convt_tc(w, out, cin, cout, dtype, reinterpret_cast<cudaStream_t>(stream));
}

int dg_dec_composite_bytes(int32_t cl, int32_t cu, size_t* bytes) {
    if (bytes == nullptr) { set_error("null bytes"); return 2; }
    return dec_composite_bytes(cl, cu, bytes);
}
int dg_pack_dec_composite(const float* ct_w, const float* ct_b, const float* conv_w, void* out, int32_t cl, int32_t cu,
                          int32_t dtype, dg_stream_t stream) {
    if (!ct_w || !ct_b || !conv_w || !out) { set_error("pack: null pointer"); return 2; }
    return pack_dec_composite(ct_w, ct_b, conv_w, out, cl, cu, dtype, reinterpret_cast<cudaStream_t>(stream));
}

int dg_lw_workspace_bytes(const dg_lw_params* p, int32_t N, int32_t H, int32_t W, size_t* bytes) {
    LwPlan pl;
    int rc = make_plan(p, N, H, W, &pl);
    if (rc) return rc;
    if (bytes) *bytes = pl.total_bytes;
    return 0;
}

int dg_lw_layout(const dg_lw_params* p, int32_t N, int32_t H, int32_t W, int32_t idx, size_t* raw_offset,
                 size_t* stats_offset, int32_t* channels, int32_t* h, int32_t* w) {
    LwPlan pl;
    int rc = make_plan(p, N, H, W, &pl);
    if (rc) return rc;
    if (idx < 0 || idx >= 18) { set_error("layout: conv index %d", idx); return 2; }
    if (raw_offset) *raw_offset = pl.raw_off[idx];
    if (stats_offset) *stats_offset = pl.stats_off[idx];
    if (channels) *channels = pl.conv_c[idx];
    if (h) *h = pl.conv_h[idx];
    if (w) *w = pl.conv_w[idx];
    return 0;
}

int dg_lw_forward(const dg_lw_params* p, const float* x, float* y, int32_t N, int32_t H, int32_t W, void* workspace,
                  size_t workspace_bytes, const float* target, double* l1_sum, dg_stream_t stream) {
    return lw_forward(p, x, y, N, H, W, workspace, workspace_bytes, target, l1_sum,
                      reinterpret_cast<cudaStream_t>(stream));
}

int dg_convt2x2_fused(const dg_src* src, int32_t dtype, int32_t N, int32_t H, int32_t W, void* out, float eps, int32_t path,
                      dg_stream_t stream) {
    if (src == nullptr || out == nullptr || N < 1 || H < 2 || W < 2) { set_error("convt2x2: bad arguments"); return 2; }
    int rc = validate_src(*src, "convt2x2");
    if (rc) return rc;
    bool handled = false;
    rc = convt_tc_launch(*src, dtype, N, H, W, out, eps, path, reinterpret_cast<cudaStream_t>(stream), &handled);
    if (rc) return rc;
    if (!handled) { set_error("convt2x2: configuration not covered by the tensor-core kernel (use the fused DG_X_CONVT2 source)"); return 3; }
    return 0;
}

int dg_band_stats(const double* kernel_stats, const void* rows, int32_t dtype, int32_t W, int32_t C, int32_t halo_top0, int32_t own0,
                  int32_t own1, int32_t halo_bottom1, double* out, dg_stream_t stream) {
    if (!kernel_stats || !rows || !out || W < 1 || halo_top0 < 0 || halo_top0 > own0 || own0 > own1 || own1 > halo_bottom1) {
        set_error("band_stats: bad arguments");
        return 2;
    }
    return band_stats_launch(kernel_stats, rows, dtype, W, C, halo_top0, own0, own1, halo_bottom1, out, reinterpret_cast<cudaStream_t>(stream));
}

int dg_gn_affine(const double* parts, int32_t nparts, size_t part_stride, const float* gamma, const float* beta, int32_t C, int32_t groups,
                 double plane, float eps, float* coef, dg_stream_t stream) {
    if (!parts || nparts < 1 || !gamma || !beta || !coef || C < 1 || groups < 1 || C % groups || plane <= 0 ||
        (nparts > 1 && part_stride < (size_t)2 * C)) {
        set_error("gn_affine: bad arguments");
        return 2;
    }
    return gn_affine_launch(parts, nparts, part_stride, gamma, beta, C, groups, plane, eps, coef, reinterpret_cast<cudaStream_t>(stream));
}

int dg_channel_attention(const double* act_sum, double plane, const float* w1, const float* w2, int32_t N, int32_t C,
                         int32_t hidden, float* scale, dg_stream_t stream) {
    if (!act_sum || !w1 || !w2 || !scale || N < 1 || C < 1 || hidden < 1 || plane <= 0) {
        set_error("channel_attention: bad arguments");
        return 2;
    }
    return se_scale_launch(act_sum, plane, w1, w2, N, C, hidden, scale, reinterpret_cast<cudaStream_t>(stream));
}

// ---- per-op backward entry points (OptimizedUNet training is orchestrated above the C-ABI from these) ----------------------
static int check_gn(const char* who, const void* raw, const double* stats, const float* gamma, const float* beta, int N, int H, int W,
                    int C, int groups) {
    if (raw == nullptr || stats == nullptr || gamma == nullptr || beta == nullptr) { set_error("%s: null pointer", who); return 2; }
    if (N < 1 || N > 65535 || H < 1 || W < 1 || C < 1) { set_error("%s: bad shape (N must be 1..65535: images are grid.y)", who); return 3; }
    if (groups < 1 || C % groups != 0) { set_error("%s: %d channels not divisible into %d groups", who, C, groups); return 2; }
    return 0;
}

int dg_head1x1_bwd(const dg_head_args* a, const float* grad_y, float* G, double* P, float* dW, float* dB, dg_stream_t stream) {
    if (a == nullptr || grad_y == nullptr || G == nullptr || P == nullptr || dW == nullptr || dB == nullptr) {
        set_error("head backward: null pointer");
        return 2;
    }
    int rc = validate_src(a->src, "head backward");
    if (rc) return rc;
    if (a->src.xform != DG_X_SAME || a->src.stats == nullptr || !a->src.silu || a->src.scale != nullptr || a->src.coef != nullptr) {
        set_error("head backward: the source must be a same-resolution GroupNorm + SiLU tensor described by its statistics");
        return 3;
    }
    if (a->weight == nullptr) { set_error("head backward: null weight"); return 2; }
    return head_bwd_launch(a->dtype, a->src.raw, a->src.stats, a->src.gamma, a->src.beta, grad_y, a->weight, G, P, dW, dB, a->N, a->H,
                           a->W, a->src.channels, a->cout, a->src.groups, a->eps, reinterpret_cast<cudaStream_t>(stream));
}

int dg_act_bwd(int32_t dtype, const void* raw, const double* stats, const float* gamma, const float* beta, int32_t groups,
               const float* dA_a, int32_t stride_a, int32_t off_a, const float* dA_b, int32_t stride_b, int32_t off_b, float* G,
               double* P, int32_t N, int32_t H, int32_t W, int32_t C, float eps, dg_stream_t stream) {
    int rc = check_gn("act backward", raw, stats, gamma, beta, N, H, W, C, groups);
    if (rc) return rc;
    if (G == nullptr || P == nullptr || (dA_a == nullptr && dA_b == nullptr)) { set_error("act backward: null pointer"); return 2; }
    if (dA_b != nullptr && ((H | W) & 1)) { set_error("act backward: pooled gradient needs even H, W"); return 3; }
    if ((dA_a != nullptr && (off_a < 0 || off_a + C > stride_a)) || (dA_b != nullptr && (off_b < 0 || off_b + C > stride_b))) {
        set_error("act backward: channel window outside the gradient tensor");
        return 3;
    }
    return act_bwd_launch(dtype, raw, stats, gamma, beta, dA_a, stride_a, off_a, dA_b, stride_b, off_b, G, P, N, H, W, C, groups, eps,
                          reinterpret_cast<cudaStream_t>(stream));
}

int dg_gn_bwd_apply(int32_t dtype, const void* raw, const double* stats, const float* gamma, int32_t groups, const double* P, float* G,
                    float* dgamma, float* dbeta, int32_t N, int32_t H, int32_t W, int32_t C, float eps, dg_stream_t stream) {
    int rc = check_gn("GroupNorm backward", raw, stats, gamma, gamma, N, H, W, C, groups);
    if (rc) return rc;
    if (P == nullptr || G == nullptr || (dgamma == nullptr) != (dbeta == nullptr)) { set_error("GroupNorm backward: null pointer"); return 2; }
    return gn_bwd_apply_launch(dtype, raw, stats, gamma, P, G, dgamma, dbeta, N, H, W, C, groups, eps,
                               reinterpret_cast<cudaStream_t>(stream));
}

int dg_grad_gather(const float* a, int32_t stride_a, int32_t off_a, const float* a_scale, const float* b, int32_t stride_b,
                   int32_t off_b, const float* u, int32_t stride_u, int32_t off_u, const float* add, float* out, int32_t N, int32_t H,
                   int32_t W, int32_t C, dg_stream_t stream) {
    if (out == nullptr || (a == nullptr && b == nullptr && u == nullptr)) { set_error("grad gather: null pointer"); return 2; }
    if (N < 1 || H < 1 || W < 1 || C < 1) { set_error("grad gather: bad shape"); return 3; }
    if (b != nullptr && ((H | W) & 1)) { set_error("grad gather: pooled gradient needs even H, W"); return 3; }
    if ((a != nullptr && (off_a < 0 || off_a + C > stride_a)) || (b != nullptr && (off_b < 0 || off_b + C > stride_b)) ||
        (u != nullptr && (off_u < 0 || off_u + C > stride_u)) || (a == nullptr && a_scale != nullptr)) {
        set_error("grad gather: channel window outside a gradient tensor");
        return 3;
    }
    return grad_gather_launch(a, stride_a, off_a, a_scale, b, stride_b, off_b, u, stride_u, off_u, add, out, N, H, W, C,
                              reinterpret_cast<cudaStream_t>(stream));
}

int dg_scale_bwd_sum(int32_t dtype, const void* raw, const double* stats, const float* gamma, const float* beta, int32_t groups,
                     const float* d, int32_t stride_d, int32_t off_d, double* dscale, int32_t N, int32_t H, int32_t W, int32_t C,
                     float eps, dg_stream_t stream) {
    int rc = check_gn("scale backward", raw, stats, gamma, beta, N, H, W, C, groups);
    if (rc) return rc;
    if (d == nullptr || dscale == nullptr) { set_error("scale backward: null pointer"); return 2; }
    if (off_d < 0 || off_d + C > stride_d) { set_error("scale backward: channel window outside the gradient tensor"); return 3; }
    if (C > 2048) { set_error("scale backward: %d channels > 2048", C); return 3; }
    return scale_bwd_sum_launch(dtype, raw, stats, gamma, beta, d, stride_d, off_d, dscale, N, H, W, C, groups, eps,
                                reinterpret_cast<cudaStream_t>(stream));
}

int dg_channel_attention_bwd(const double* act_sum, double plane, const float* w1, const float* w2, const double* dscale, int32_t N,
                             int32_t C, int32_t hidden, float* add, float* dw1, float* dw2, dg_stream_t stream) {
    if (act_sum == nullptr || w1 == nullptr || w2 == nullptr || dscale == nullptr || add == nullptr || dw1 == nullptr || dw2 == nullptr) {
        set_error("attention backward: null pointer");
        return 2;
    }
    if (N < 1 || C < 1 || hidden < 1 || plane <= 0.0 || (size_t)(2 * C + 3 * hidden) * sizeof(float) > 48 * 1024) {
        set_error("attention backward: bad shape");
        return 3;
    }
    return se_bwd_launch(act_sum, plane, w1, w2, dscale, N, C, hidden, add, dw1, dw2, reinterpret_cast<cudaStream_t>(stream));
}

int dg_lw_num_params(const dg_lw_params* p, size_t* count) {
    LwPlan pl;
    int rc = make_plan(p, 1, 16, 16, &pl);
    if (rc) return rc;
    GradLayout gl;
    make_grad_layout(p, pl, &gl);
    if (count) *count = gl.total;
    return 0;
}

int dg_lw_backward_workspace_bytes(const dg_lw_params* p, int32_t N, int32_t H, int32_t W, size_t* bytes) {
    LwPlan pl;
    int rc = make_plan(p, N, H, W, &pl);
    if (rc) return rc;
    BwdPlan bp;
    make_bwd_plan(p, pl, N, &bp);
    if (bytes) *bytes = bp.total;
    return 0;
}

int dg_lw_backward(const dg_lw_params* p, const float* x, const float* grad_y, int32_t N, int32_t H, int32_t W,
                   void* fwd_workspace, size_t fwd_bytes, void* bwd_workspace, size_t bwd_bytes, float* grads,
                   dg_stream_t stream) {
    return lw_backward(p, x, grad_y, N, H, W, fwd_workspace, fwd_bytes, bwd_workspace, bwd_bytes, grads,
                       reinterpret_cast<cudaStream_t>(stream));
}

int dg_lw_backward_l1(const dg_lw_params* p, const float* x, const float* y, const float* target, const float* loss_grad,
                      int32_t N, int32_t H, int32_t W, void* fwd_workspace, size_t fwd_bytes, void* bwd_workspace, size_t bwd_bytes,
                      float* grads, dg_stream_t stream) {
    if (y == nullptr || target == nullptr) { set_error("backward_l1: null output / target"); return 2; }
    if (p == nullptr) { set_error("null params"); return 2; }
    L1Seed l1{y, target, loss_grad, 1.0f / ((float)N * (float)p->out_channels * (float)H * (float)W)};
    return lw_backward(p, x, nullptr, N, H, W, fwd_workspace, fwd_bytes, bwd_workspace, bwd_bytes, grads,
                       reinterpret_cast<cudaStream_t>(stream), &l1);
}

int dg_l1_loss_sum(const float* y, const float* target, size_t count, double* sum, dg_stream_t stream) {
    if (!y || !target || !sum || count == 0) { set_error("l1_loss_sum: bad arguments"); return 2; }
    return l1_sum_launch(y, target, count, sum, reinterpret_cast<cudaStream_t>(stream));
}

int dg_adamw_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, size_t count, double* scratch,
                  float max_norm, float lr, float beta1, float beta2, float eps, float weight_decay, int32_t step,
                  float grad_scale, dg_stream_t stream) {
    if (!params || !grads || !exp_avg || !exp_avg_sq || !scratch || count == 0 || step < 1) {
        set_error("adamw: bad arguments");
        return 2;
    }
    return adamw_launch(params, grads, exp_avg, exp_avg_sq, count, scratch, max_norm, lr, beta1, beta2, eps, weight_decay,
                        step, grad_scale, reinterpret_cast<cudaStream_t>(stream));
}

int dg_adamw_step_graph(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, size_t count, double* scratch,
                        float max_norm, const float* lr_dev, float beta1, float beta2, float eps, float weight_decay,
                        int32_t* step_dev, float grad_scale, dg_stream_t stream) {
    if (!params || !grads || !exp_avg || !exp_avg_sq || !scratch || !lr_dev || !step_dev || count == 0) {
        set_error("adamw (device step): bad arguments");
        return 2;
    }
    return adamw_dev_launch(params, grads, exp_avg, exp_avg_sq, count, scratch, max_norm, lr_dev, beta1, beta2, eps, weight_decay,
                            step_dev, grad_scale, reinterpret_cast<cudaStream_t>(stream));
}

int dg_lw_profile(const dg_lw_params* p, const float* x, float* y, int32_t N, int32_t H, int32_t W, void* workspace,
                  size_t workspace_bytes, dg_stream_t stream, float* ms19) {
    if (ms19 == nullptr) { set_error("profile: null output"); return 2; }
    cudaEvent_t evs[20];
    for (int i = 0; i < 20; ++i) cudaEventCreate(&evs[i]);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    int rc = lw_forward(p, x, y, N, H, W, workspace, workspace_bytes, nullptr, nullptr, st, evs);
    cudaError_t e = cudaStreamSynchronize(st);
    if (rc == 0 && e != cudaSuccess) { set_error("profile: %s", cudaGetErrorString(e)); rc = 10; }
    if (rc == 0)
        for (int i = 0; i < 19; ++i) cudaEventElapsedTime(&ms19[i], evs[i], evs[i + 1]);
    for (int i = 0; i < 20; ++i) cudaEventDestroy(evs[i]);
    return rc;
}

int dg_lw_host_scratch_bytes(const dg_lw_params* p, int32_t chunk, int32_t H, int32_t W, size_t* bytes) {
    LwPlan pl;
    int rc = make_plan(p, chunk, H, W, &pl);
    if (rc) return rc;
    const size_t in_b = align_up((size_t)chunk * p->in_channels * H * W * sizeof(float), 256);
    const size_t out_b = align_up((size_t)chunk * p->out_channels * H * W * sizeof(float), 256);
    if (bytes) *bytes = dg::HostPipe::NS * (in_b + out_b + align_up(pl.total_bytes, 256));
    return 0;
}

// Chunk i: H2D on s_in -> forward on s_cmp[i % NS] with workspace i % NS -> D2H on s_out.  Several compute streams let the
// under-filled deep layers of one chunk (16 images x 8 tiles < 148 SMs) overlap the next chunk's wide layers.
// ticket == nullptr: blocking call (returns when host_y is complete).  ticket != nullptr: everything is enqueued, *ticket names
// the call for dg_lw_infer_host_wait, and the pipeline keeps running across calls.
static int infer_host_impl(const dg_lw_params* p, const void* host_x, void* host_y, int N, int H, int W, int chunk,
                           void* dev_ws, size_t dev_ws_bytes, int io, cudaStream_t caller, int64_t* ticket = nullptr) {
    if (chunk < 1) { set_error("infer_host: chunk %d", chunk); return 2; }
    if (chunk > N) chunk = N;
    LwPlan pl;
    int rc = make_plan(p, chunk, H, W, &pl);
    if (rc) return rc;
    size_t need = 0;
    dg_lw_host_scratch_bytes(p, chunk, H, W, &need);
    if (dev_ws == nullptr || dev_ws_bytes < need) { set_error("infer_host: scratch %zu < %zu", dev_ws_bytes, need); return 4; }
    if (host_x == nullptr || host_y == nullptr) { set_error("infer_host: null host buffer"); return 2; }
    std::lock_guard<std::mutex> lock(dg::g_pipe_mutex);
    dg::HostPipe* pipe = nullptr;
    if ((rc = dg::pipe_init(&pipe))) return rc;
    const size_t esz_in = (io & 1) ? 1 : sizeof(float), esz_out = (io & 2) ? 1 : sizeof(float);
    const size_t in_img = (size_t)p->in_channels * H * W, out_img = (size_t)p->out_channels * H * W;
    const size_t in_b = align_up(chunk * in_img * sizeof(float), 256);
    const size_t out_b = align_up(chunk * out_img * sizeof(float), 256);
    const size_t ws_b = align_up(pl.total_bytes, 256);
    char* base = static_cast<char*>(dev_ws);
    constexpr int NS = dg::HostPipe::NS;
    char *dx[NS], *dy[NS], *ws[NS];
    for (int k = 0; k < NS; ++k) {
        dx[k] = base + k * in_b;
        dy[k] = base + NS * in_b + k * out_b;
        ws[k] = base + NS * (in_b + out_b) + k * ws_b;
    }
    dg::HostPipe& P = *pipe;
    if (P.ws_in_flight != dev_ws || P.chunk_in_flight != chunk || P.h_in_flight != H || P.w_in_flight != W) {
        // other scratch or geometry than the calls still in flight: their slots are not ours -- drain first
        cudaStreamSynchronize(P.s_out);
        for (int k = 0; k < NS; ++k) cudaStreamSynchronize(P.s_cmp[k]);
        P.seq = 0;
        P.ws_in_flight = dev_ws; P.chunk_in_flight = chunk; P.h_in_flight = H; P.w_in_flight = W;
    }
    // everything the caller enqueued on ITS stream (weight packing after a parameter update, the previous consumer of dev_ws) is
    // ordered before the first kernel of the pipeline
    cudaError_t ce = cudaEventRecord(P.start, caller);
    for (int k = 0; k < NS && ce == cudaSuccess; ++k) ce = cudaStreamWaitEvent(P.s_cmp[k], P.start, 0);
    if (ce != cudaSuccess) { set_error("infer_host: %s", cudaGetErrorString(ce)); return 10; }
    const char* hx = static_cast<const char*>(host_x);
    char* hy = static_cast<char*>(host_y);
    // chunk schedule: a half-size first and last chunk shorten the pipeline fill (first H2D) and drain (last D2H)
    const int head = (ticket == nullptr && N >= 3 * chunk && chunk >= 2) ? chunk / 2 : 0;
    const int body = N - 2 * head;
    const int nchunks = (body + chunk - 1) / chunk + (head ? 2 : 0);
    int n0 = 0;
    for (int i = 0; i < nchunks; ++i, ++P.seq) {
        const int b = (int)(P.seq % NS);
        const bool reuse = P.seq >= (unsigned long long)NS;   // slot b was used by chunk seq - NS (possibly of an earlier call)
        int nn;
        if (head && (i == 0 || i == nchunks - 1)) nn = head;
        else { const int left = N - head - n0; nn = left < chunk ? left : chunk; }  // body: what the tail chunk leaves
        if (reuse) cudaStreamWaitEvent(P.s_in, P.cmp_done[b], 0);  // dx[b] consumed by chunk seq-NS
        ce = cudaMemcpyAsync(dx[b], hx + (size_t)n0 * in_img * esz_in, nn * in_img * esz_in, cudaMemcpyHostToDevice, P.s_in);
        if (ce != cudaSuccess) { cudaDeviceSynchronize(); set_error("infer_host: H2D copy: %s", cudaGetErrorString(ce)); return 10; }
        cudaEventRecord(P.in_done[b], P.s_in);
        cudaStreamWaitEvent(P.s_cmp[b], P.in_done[b], 0);
        if (reuse) cudaStreamWaitEvent(P.s_cmp[b], P.out_done[b], 0);  // dy[b] drained by chunk seq-NS
        rc = dg::lw_forward(p, dx[b], dy[b], nn, H, W, ws[b], pl.total_bytes, nullptr, nullptr, P.s_cmp[b], nullptr, io, 1 + b);
        if (rc) { cudaDeviceSynchronize(); return rc; }
        cudaEventRecord(P.cmp_done[b], P.s_cmp[b]);
        cudaStreamWaitEvent(P.s_out, P.cmp_done[b], 0);
        ce = cudaMemcpyAsync(hy + (size_t)n0 * out_img * esz_out, dy[b], nn * out_img * esz_out, cudaMemcpyDeviceToHost, P.s_out);
        if (ce != cudaSuccess) { cudaDeviceSynchronize(); set_error("infer_host: D2H copy: %s", cudaGetErrorString(ce)); return 10; }
        cudaEventRecord(P.out_done[b], P.s_out);
        n0 += nn;
    }
    if (ticket != nullptr) {   // s_out is one stream: its last copy done = every chunk of this call computed and copied out
        const unsigned long long t = P.tickets++;
        ce = cudaEventRecord(P.ticket_done[t % dg::HostPipe::NT], P.s_out);
        if (ce != cudaSuccess) { cudaDeviceSynchronize(); set_error("infer_host: %s", cudaGetErrorString(ce)); return 10; }
        *ticket = (int64_t)t;
        return 0;
    }
    cudaError_t e = cudaStreamSynchronize(P.s_out);
    for (int k = 0; k < NS && e == cudaSuccess; ++k) e = cudaStreamSynchronize(P.s_cmp[k]);
    if (e != cudaSuccess) { set_error("infer_host: %s", cudaGetErrorString(e)); return 10; }
    return 0;
}

int dg_lw_infer_host(const dg_lw_params* p, const float* host_x, float* host_y, int32_t N, int32_t H, int32_t W,
                     int32_t chunk, void* dev_ws, size_t dev_ws_bytes, dg_stream_t stream) {
    return infer_host_impl(p, host_x, host_y, N, H, W, chunk, dev_ws, dev_ws_bytes, 0, reinterpret_cast<cudaStream_t>(stream));
}

int dg_lw_infer_host_u8(const dg_lw_params* p, const uint8_t* host_x, uint8_t* host_y, int32_t N, int32_t H, int32_t W,
                        int32_t chunk, void* dev_ws, size_t dev_ws_bytes, dg_stream_t stream) {
    return infer_host_impl(p, host_x, host_y, N, H, W, chunk, dev_ws, dev_ws_bytes, 3, reinterpret_cast<cudaStream_t>(stream));
}

int dg_lw_infer_host_submit(const dg_lw_params* p, const void* host_x, void* host_y, int32_t N, int32_t H, int32_t W, int32_t chunk,
                            void* dev_ws, size_t dev_ws_bytes, int32_t u8, dg_stream_t stream, int64_t* ticket) {
    if (ticket == nullptr) { set_error("infer_host_submit: null ticket"); return 2; }
    {   // at most NT calls outstanding: the event about to be re-recorded must have completed
        int dev = 0;
        cudaGetDevice(&dev);
        if (dev >= 0 && dev < 16 && dg::g_pipes[dev].ready && dg::g_pipes[dev].tickets >= (unsigned long long)dg::HostPipe::NT)
            cudaEventSynchronize(dg::g_pipes[dev].ticket_done[dg::g_pipes[dev].tickets % dg::HostPipe::NT]);
    }
    return infer_host_impl(p, host_x, host_y, N, H, W, chunk, dev_ws, dev_ws_bytes, u8 ? 3 : 0, reinterpret_cast<cudaStream_t>(stream), ticket);
}

int dg_lw_infer_host_wait(int64_t ticket) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 16 || !dg::g_pipes[dev].ready) { set_error("infer_host_wait: no pipeline on this device"); return 2; }
    dg::HostPipe& P = dg::g_pipes[dev];
    if (ticket < 0 || (unsigned long long)ticket >= P.tickets) { set_error("infer_host_wait: unknown ticket %lld", (long long)ticket); return 2; }
    if (P.tickets - (unsigned long long)ticket > (unsigned long long)dg::HostPipe::NT) return 0;   // long retired (submit waited on it)
    cudaError_t e = cudaEventSynchronize(P.ticket_done[(unsigned long long)ticket % dg::HostPipe::NT]);
    if (e != cudaSuccess) { set_error("infer_host_wait: %s", cudaGetErrorString(e)); return 10; }
    return 0;
}

int dg_lw_forward_u8(const dg_lw_params* p, const uint8_t* x, uint8_t* y, int32_t N, int32_t H, int32_t W, void* workspace,
                     size_t workspace_bytes, dg_stream_t stream) {
    return dg::lw_forward(p, x, y, N, H, W, workspace, workspace_bytes, nullptr, nullptr,
                          reinterpret_cast<cudaStream_t>(stream), nullptr, 3);
}

}  // extern "C"
