// On-device validation metrics: per-image mean squared error and mean structural similarity of
// output vs target, the two quantities optimized_train.py:92-122 and evaluate.py:254-272 obtain per image on the host from
// skimage.metrics.peak_signal_noise_ratio / structural_similarity(data_range=1.0) -- SSIM with its defaults for float
// images: 7x7 uniform window, K1 = 0.01, K2 = 0.03, sample covariance (49/48), the 3-pixel border cropped, i.e. the mean
// over every window that lies fully inside the image.  One CTA = a 16x16 block of window centres: the 22x22 input patches
// of both images go to shared memory, the five window sums (x, y, xx, yy, xy) are formed separably in double, and one double
// atomicAdd per CTA accumulates sum(S).  The MSE kernel is a plain grid-stride reduction.
#include "common.cuh"

namespace dg {

namespace {
constexpr int MT = 16, MW = 7, MP = MT + MW - 1;  // centres per tile side, window, patch side (static smem: 19 KB)

__global__ void __launch_bounds__(256) ssim_kernel(const float* __restrict__ out, const float* __restrict__ tgt, int H, int W,
                                                    int clip01, double c1, double c2, double* __restrict__ acc) {
    __shared__ float xs[MP][MP + 1], ys[MP][MP + 1];
    __shared__ double hs[5][MP][MT + 1];
    __shared__ double red[8];
    const int n = blockIdx.z;
    const int r0 = blockIdx.y * MT, c0 = blockIdx.x * MT;   // first window's top-left corner = first centre - 3
    const int VH = H - MW + 1, VW = W - MW + 1;               // windows per column / row
    const float* o = out + (size_t)n * H * W;
    const float* t = tgt + (size_t)n * H * W;
    for (int i = threadIdx.x; i < MP * MP; i += 256) {
        const int r = i / MP, c = i - r * MP;
        const int gy = r0 + r, gx = c0 + c;
        float a = 0.f, b = 0.f;
        if (gy < H && gx < W) {
            a = __ldg(t + (size_t)gy * W + gx);    // im1 = target, im2 = output (argument order of the reference's calls)
            b = __ldg(o + (size_t)gy * W + gx);
            if (clip01) b = fminf(fmaxf(b, 0.f), 1.f);
        }
        xs[r][c] = a;
        ys[r][c] = b;
    }
    __syncthreads();
    // horizontal 7-sums: hs[q][r][c] for patch row r, window column c
    for (int i = threadIdx.x; i < MP * MT; i += 256) {
        const int r = i / MT, c = i - r * MT;
        double sx = 0, sy = 0, sxx = 0, syy = 0, sxy = 0;
#pragma unroll
        for (int k = 0; k < MW; ++k) {
            const float a = xs[r][c + k], b = ys[r][c + k];
            // products in float32, as skimage forms im1 * im1 before filtering
            sx += (double)a; sy += (double)b; sxx += (double)(a * a); syy += (double)(b * b); sxy += (double)(a * b);
        }
        hs[0][r][c] = sx; hs[1][r][c] = sy; hs[2][r][c] = sxx; hs[3][r][c] = syy; hs[4][r][c] = sxy;
    }
    __syncthreads();
    double local = 0.0;
    for (int i = threadIdx.x; i < MT * MT; i += 256) {
        const int r = i / MT, c = i - r * MT;
        if (r0 + r >= VH || c0 + c >= VW) continue;
        double s[5] = {0, 0, 0, 0, 0};
#pragma unroll
        for (int k = 0; k < MW; ++k)
#pragma unroll
            for (int q = 0; q < 5; ++q) s[q] += hs[q][r + k][c];
        const double inv = 1.0 / (MW * MW), cov = (double)(MW * MW) / (MW * MW - 1);
        const double ux = s[0] * inv, uy = s[1] * inv;
        const double vx = cov * (s[2] * inv - ux * ux), vy = cov * (s[3] * inv - uy * uy), vxy = cov * (s[4] * inv - ux * uy);
        local += ((2 * ux * uy + c1) * (2 * vxy + c2)) / ((ux * ux + uy * uy + c1) * (vx + vy + c2));
    }
    for (int k = 16; k > 0; k >>= 1) local += __shfl_xor_sync(0xffffffffu, local, k);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = local;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int w = 0; w < 8; ++w) s += red[w];
        atomicAdd(acc + 2 * n + 1, s);
    }
}

__global__ void __launch_bounds__(256) mse_kernel(const float* __restrict__ out, const float* __restrict__ tgt, int HW, int clip01,
                                                   double* __restrict__ acc) {
    __shared__ double red[8];
    const int n = blockIdx.y;
    const float* o = out + (size_t)n * HW;
    const float* t = tgt + (size_t)n * HW;
    double local = 0.0;
    for (int i = blockIdx.x * 256 + threadIdx.x; i < HW; i += gridDim.x * 256) {
        float b = __ldg(o + i);
        if (clip01) b = fminf(fmaxf(b, 0.f), 1.f);
        const double d = (double)__ldg(t + i) - (double)b;
        local += d * d;
    }
    for (int k = 16; k > 0; k >>= 1) local += __shfl_xor_sync(0xffffffffu, local, k);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = local;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int w = 0; w < 8; ++w) s += red[w];
        atomicAdd(acc + 2 * n, s);
    }
}
}  // namespace

// acc [N][2] doubles, zero on entry: (sum of squared errors, sum of S over all full windows)
int image_metrics_launch(const float* out, const float* tgt, int N, int H, int W, int clip01, double data_range, double* acc,
                         cudaStream_t st) {
    if (H < MW || W < MW) { set_error("metrics: image %dx%d smaller than the 7x7 SSIM window", H, W); return 3; }
    if (N < 1 || N > 65535) { set_error("metrics: batch %d", N); return 3; }
    const double c1 = (0.01 * data_range) * (0.01 * data_range), c2 = (0.03 * data_range) * (0.03 * data_range);
    dim3 grid((W - MW + 1 + MT - 1) / MT, (H - MW + 1 + MT - 1) / MT, N);
    ssim_kernel<<<grid, 256, 0, st>>>(out, tgt, H, W, clip01, c1, c2, acc);
    count_launch();
    int blocks = (H * W + 256 * 8 - 1) / (256 * 8);
    if (blocks > 256) blocks = 256;
    mse_kernel<<<dim3(blocks, N), 256, 0, st>>>(out, tgt, H * W, clip01, acc);
    count_launch();
    return check_launch("image_metrics");
}

}  // namespace dg
