// Image pre/post-processing either side of the network, on the device.  Integer work, bit-exact with the
// libraries the reference calls on the host:
//   f1  api/app.py:143-150,199-203   PIL convert('L') (Convert.c rgb2l, 16-bit fixed point) + Image.resize(LANCZOS)
//                                    (Resample.c: two passes, 22-bit fixed-point taps, uint8 rounding after each pass)
//   f2  src/optimized_dataset.py:104-123   triptych split (a column offset), cv2.cvtColor RGB2GRAY (15-bit fixed point) and
//                                    cv2.resize INTER_LINEAR on uint8 (11-bit taps, the library's two-stage truncation)
//       src/optimized_dataset.py:159-172   HorizontalFlip / RandomBrightnessContrast / GaussNoise given the sampled parameters
// The tap tables are computed on the host (double / float arithmetic exactly as the libraries do) and passed in; the kernels are
// plain gather-accumulate loops: a 4096x4096 RGB input is 50 MB read once, the 512x512 result 0.26 MB -- HBM-bound by construction.
#include <stdint.h>

#include "common.cuh"

namespace dg {

namespace {

constexpr int PIL_BITS = 22;   // Resample.c PRECISION_BITS = 32 - 8 - 2

__device__ __forceinline__ int pil_l(const uint8_t* p) {            // Convert.c L24 >> 16
    return (p[0] * 19595 + p[1] * 38470 + p[2] * 7471 + 0x8000) >> 16;
}
__device__ __forceinline__ int cv_gray(const uint8_t* p) {          // color_rgb.simd.hpp RGB2Gray<uchar>, gray_shift = 15
    return (p[0] * 9798 + p[1] * 19235 + p[2] * 3735 + (1 << 14)) >> 15;
}
__device__ __forceinline__ uint8_t clip8(int v) { return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v)); }

// Horizontal LANCZOS pass over source rows [row0, row0 + rows) (the rows the vertical pass reads); the 'L' conversion of an RGB(A)
// source happens per tap (bit-identical to converting the whole image first).  ident != 0: no resampling along x, copy / convert only.
__global__ void pil_h_kernel(const uint8_t* __restrict__ src, int C, int in_h, int in_w, int row0, int rows, uint8_t* __restrict__ dst,
                             int out_w, const int32_t* __restrict__ bounds, const int32_t* __restrict__ kk, int ksize, int ident) {
    const int xx = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y, n = blockIdx.z;
    if (xx >= out_w || r >= rows) return;
    const uint8_t* row = src + ((size_t)n * in_h + row0 + r) * in_w * C;
    int v;
    if (ident) {
        v = C >= 3 ? pil_l(row + (size_t)xx * C) : row[xx];
    } else {
        const int x0 = bounds[2 * xx], cnt = bounds[2 * xx + 1];
        const int32_t* k = kk + (size_t)xx * ksize;
        int ss = 1 << (PIL_BITS - 1);
        if (C >= 3) {
            for (int x = 0; x < cnt; ++x) ss += pil_l(row + (size_t)(x0 + x) * C) * k[x];
        } else {
            for (int x = 0; x < cnt; ++x) ss += row[x0 + x] * k[x];
        }
        v = ss >> PIL_BITS;
    }
    dst[((size_t)n * rows + r) * out_w + xx] = clip8(v);
}

__global__ void pil_v_kernel(const uint8_t* __restrict__ src, int rows, int row0, uint8_t* __restrict__ dst, int out_h, int out_w,
                             const int32_t* __restrict__ bounds, const int32_t* __restrict__ kk, int ksize) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, yy = blockIdx.y, n = blockIdx.z;
    if (x >= out_w) return;
    const int y0 = bounds[2 * yy] - row0, cnt = bounds[2 * yy + 1];
    const int32_t* k = kk + (size_t)yy * ksize;
    const uint8_t* col = src + ((size_t)n * rows + y0) * out_w + x;
    int ss = 1 << (PIL_BITS - 1);
    for (int y = 0; y < cnt; ++y) ss += col[(size_t)y * out_w] * k[y];
    dst[((size_t)n * out_h + yy) * out_w + x] = clip8(ss >> PIL_BITS);
}

// cv2.resize INTER_LINEAR (uint8) of the panel [x_off, x_off + in_w) of an image of full width in_wf; RGB sources are converted with
// the cv2 gray formula per tap.  area2 != 0: the library's exact-2x fast path (INTER_AREA, (a + b + c + d + 2) >> 2).
__global__ void cv2_resize_kernel(const uint8_t* __restrict__ src, int C, int in_h, int in_wf, int x_off, uint8_t* __restrict__ dst,
                                  int out_h, int out_w, const int32_t* __restrict__ xofs, const int32_t* __restrict__ xab,
                                  const int32_t* __restrict__ yofs, const int32_t* __restrict__ yab, int area2) {
    const int dx = blockIdx.x * blockDim.x + threadIdx.x, dy = blockIdx.y, n = blockIdx.z;
    if (dx >= out_w) return;
    const uint8_t* img = src + (size_t)n * in_h * in_wf * C;
    auto px = [&](int y, int x) -> int {
        const uint8_t* p = img + ((size_t)y * in_wf + x_off + x) * C;
        return C >= 3 ? cv_gray(p) : (int)p[0];
    };
    int v;
    if (area2) {
        v = (px(2 * dy, 2 * dx) + px(2 * dy, 2 * dx + 1) + px(2 * dy + 1, 2 * dx) + px(2 * dy + 1, 2 * dx + 1) + 2) >> 2;
    } else {
        const int sx = xofs[dx], a0 = xab[2 * dx], a1 = xab[2 * dx + 1];
        const int sx1 = xab[2 * dx + 1] ? sx + 1 : sx;   // a clamped column has weight 0 on its (possibly out-of-range) neighbour
        int sy0 = yofs[dy], sy1 = sy0 + 1;
        const int b0 = yab[2 * dy], b1 = yab[2 * dy + 1];
        sy0 = sy0 < 0 ? 0 : (sy0 > in_h - 1 ? in_h - 1 : sy0);
        sy1 = sy1 < 0 ? 0 : (sy1 > in_h - 1 ? in_h - 1 : sy1);
        const int r0 = px(sy0, sx) * a0 + px(sy0, sx1) * a1;   // HResizeLinear: int, scale 2^11
        const int r1 = px(sy1, sx) * a0 + px(sy1, sx1) * a1;
        v = (((b0 * (r0 >> 4)) >> 16) + ((b1 * (r1 >> 4)) >> 16) + 2) >> 2;   // VResizeLinear<uchar, int, short>
    }
    dst[((size_t)n * out_h + dy) * out_w + dx] = clip8(v);
}

// counter-based generator for the noise field: two rounds of a 64-bit mix of (seed, sample, pixel) -> two uniforms -> Box-Muller
__device__ __forceinline__ uint64_t mix64(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

// image / mask uint8 [N][H][W] -> float32 [N][1][H][W] = x / 255 (optimized_dataset.py:126-127), then HorizontalFlip of both and,
// on the image only, clip(alpha * x + beta, 0, 1) and / or clip(x + N(0, sigma), 0, 1).  params [N][4] = flip, alpha, beta, sigma.
__global__ void augment_kernel(const uint8_t* __restrict__ image, const uint8_t* __restrict__ mask, float* __restrict__ image_out,
                               float* __restrict__ mask_out, int H, int W, const float* __restrict__ params, uint64_t seed) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, n = blockIdx.z;
    if (x >= W) return;
    const float flip = params[4 * n], alpha = params[4 * n + 1], beta = params[4 * n + 2], sigma = params[4 * n + 3];
    const int sx = flip != 0.f ? W - 1 - x : x;
    const size_t si = ((size_t)n * H + y) * W + sx, di = ((size_t)n * H + y) * W + x;
    float v = __fdiv_rn((float)image[si], 255.0f);
    if (alpha != 1.f || beta != 0.f) {
        v = __fadd_rn(__fmul_rn(v, alpha), beta);   // two roundings, as numpy: no fused multiply-add
        v = fminf(fmaxf(v, 0.f), 1.f);
    }
    if (sigma > 0.f) {
        const uint64_t h = mix64(mix64(seed ^ ((uint64_t)n << 40)) + di);
        const float u1 = ((float)(uint32_t)(h >> 40) + 1.0f) * (1.0f / 16777216.0f);   // (0, 1]
        const float u2 = (float)(uint32_t)((h >> 8) & 0xFFFFFFu) * (1.0f / 16777216.0f);
        const float g = sqrtf(-2.0f * logf(u1)) * cospif(2.0f * u2);
        v = fminf(fmaxf(v + sigma * g, 0.f), 1.f);
    }
    image_out[di] = v;
    if (mask != nullptr && mask_out != nullptr) mask_out[di] = __fdiv_rn((float)mask[si], 255.0f);
}

int launched(const char* what) {
    const int rc = check_launch(what);
    if (rc == 0) count_launch();
    return rc;
}

}  // namespace

}  // namespace dg

using namespace dg;

extern "C" {

int dg_pil_resize_u8(const uint8_t* src, int32_t channels, int32_t N, int32_t in_h, int32_t in_w, uint8_t* dst, int32_t out_h,
                     int32_t out_w, const int32_t* bounds_h, const int32_t* kk_h, int32_t ksize_h, const int32_t* bounds_v,
                     const int32_t* kk_v, int32_t ksize_v, int32_t row0, int32_t rows, uint8_t* tmp, size_t tmp_bytes,
                     dg_stream_t stream) {
    if (src == nullptr || dst == nullptr) { set_error("pil_resize: null pointer"); return 2; }
    if (channels != 1 && channels != 3 && channels != 4) { set_error("pil_resize: %d channels (1, 3 or 4)", channels); return 2; }
    if (N < 1 || in_h < 1 || in_w < 1 || out_h < 1 || out_w < 1) { set_error("pil_resize: bad shape"); return 2; }
    const bool need_h = out_w != in_w, need_v = out_h != in_h;
    if (need_h && (bounds_h == nullptr || kk_h == nullptr || ksize_h < 1)) { set_error("pil_resize: horizontal taps missing"); return 2; }
    if (need_v && (bounds_v == nullptr || kk_v == nullptr || ksize_v < 1)) { set_error("pil_resize: vertical taps missing"); return 2; }
    if (!need_v) { row0 = 0; rows = in_h; }
    if (row0 < 0 || rows < 1 || row0 + rows > in_h) { set_error("pil_resize: bad source row range [%d, %d)", row0, row0 + rows); return 2; }
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    uint8_t* hout = need_v ? tmp : dst;
    if (need_v && (tmp == nullptr || tmp_bytes < (size_t)N * rows * out_w)) {
        set_error("pil_resize: scratch too small: %zu < %zu", tmp_bytes, (size_t)N * rows * out_w);
        return 4;
    }
    pil_h_kernel<<<dim3((out_w + 127) / 128, rows, N), 128, 0, st>>>(src, channels, in_h, in_w, row0, rows, hout, out_w, bounds_h, kk_h,
                                                                      ksize_h, need_h ? 0 : 1);
    int rc = launched("pil_h");
    if (rc || !need_v) return rc;
    pil_v_kernel<<<dim3((out_w + 127) / 128, out_h, N), 128, 0, st>>>(tmp, rows, row0, dst, out_h, out_w, bounds_v, kk_v, ksize_v);
    return launched("pil_v");
}

int dg_cv2_resize_u8(const uint8_t* src, int32_t channels, int32_t N, int32_t in_h, int32_t in_w_full, int32_t x_off, int32_t in_w,
                     uint8_t* dst, int32_t out_h, int32_t out_w, const int32_t* xofs, const int32_t* xab, const int32_t* yofs,
                     const int32_t* yab, dg_stream_t stream) {
    if (src == nullptr || dst == nullptr) { set_error("cv2_resize: null pointer"); return 2; }
    if (channels != 1 && channels != 3) { set_error("cv2_resize: %d channels (1 or 3)", channels); return 2; }
    if (N < 1 || in_h < 1 || in_w < 1 || out_h < 1 || out_w < 1 || x_off < 0 || x_off + in_w > in_w_full) {
        set_error("cv2_resize: bad shape / panel");
        return 2;
    }
    const int area2 = in_w == 2 * out_w && in_h == 2 * out_h;
    if (!area2 && (xofs == nullptr || xab == nullptr || yofs == nullptr || yab == nullptr)) { set_error("cv2_resize: taps missing"); return 2; }
    cv2_resize_kernel<<<dim3((out_w + 127) / 128, out_h, N), 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        src, channels, in_h, in_w_full, x_off, dst, out_h, out_w, xofs, xab, yofs, yab, area2);
    return launched("cv2_resize");
}

int dg_augment(const uint8_t* image, const uint8_t* mask, float* image_out, float* mask_out, int32_t N, int32_t H, int32_t W,
               const float* params, uint64_t seed, dg_stream_t stream) {
    if (image == nullptr || image_out == nullptr || params == nullptr) { set_error("augment: null pointer"); return 2; }
    if (N < 1 || H < 1 || W < 1) { set_error("augment: bad shape"); return 2; }
    augment_kernel<<<dim3((W + 127) / 128, H, N), 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(image, mask, image_out, mask_out, H, W,
                                                                                                   params, seed);
    return launched("augment");
}

}  // extern "C"
