// Row-sharded whole-image inference (whole_image.py; SURVEY 8e "definition B"): the two small device steps either side of the
// per-conv GroupNorm all-reduce, so that a band's forward stays a short chain of launches instead of ~25 tiny tensor ops per layer.
//   band_stats : partial sums of the rows a band OWNS = the conv kernel's epilogue statistics (over every row it computed) minus
//                the halo rows it computed as well -- nn.GroupNorm (src/model.py:94,97) normalises over the whole image, each
//                rank contributes the rows it owns exactly once
//   gn_affine  : every rank's partial (sum, sum of squares), added in rank order -> the finished per-channel affine (a, b) consumers take as dg_src.coef,
//                with the SAME double-precision chain the kernels run themselves (common.cuh:gn_coef)
// One CTA each, fixed summation order: the result does not depend on the launch.
#include "common.cuh"

namespace dg {

namespace {

constexpr int BS_THREADS = 1024;

template <typename T>
__global__ void __launch_bounds__(BS_THREADS) band_stats_kernel(const double* __restrict__ kstats, const T* __restrict__ t, int W, int C,
                                                                int a0, int a1, int b0, int b1, double* __restrict__ out) {
    extern __shared__ double red[];   // [active][2]
    const int active = (BS_THREADS / C) * C;            // threads that keep ONE channel over their stride loop
    const int tid = threadIdx.x;
    double s1 = 0.0, s2 = 0.0;
    if (tid < active) {
        const size_t row = (size_t)W * C;
        for (int r = a0; r < b1; ++r) {
            if (r >= a1 && r < b0) { r = b0 - 1; continue; }   // rows [a1, b0) are the band's own
            const T* p = t + (size_t)r * row;
            for (size_t i = tid; i < row; i += active) {
                const float v = Store<T>::to_f(p[i]);
                s1 += (double)v;
                s2 += (double)v * (double)v;
            }
        }
        red[2 * tid] = s1;
        red[2 * tid + 1] = s2;
    }
    __syncthreads();
    if (tid < C) {
        double h1 = 0.0, h2 = 0.0;
        for (int k = tid; k < active; k += C) { h1 += red[2 * k]; h2 += red[2 * k + 1]; }
        out[2 * tid] = kstats[2 * tid] - h1;
        out[2 * tid + 1] = kstats[2 * tid + 1] - h2;
    }
}

__global__ void gn_affine_kernel(const double* __restrict__ parts, int nparts, size_t stride, const float* __restrict__ gamma,
                                 const float* __restrict__ beta, int C, int groups, double plane, float eps, float* __restrict__ coef) {
    extern __shared__ double tot[];   // [C][2]: the partial sums of all ranks, added in rank order (the same bits on every rank)
    for (int k = threadIdx.x; k < 2 * C; k += blockDim.x) {
        double v = 0.0;
        for (int r = 0; r < nparts; ++r) v += parts[(size_t)r * stride + k];
        tot[k] = v;
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float a, b;
        gn_coef(tot, gamma, beta, 0, C, groups, c, plane, eps, a, b);
        coef[2 * c] = a;
        coef[2 * c + 1] = b;
    }
}

}  // namespace

int band_stats_launch(const double* kstats, const void* t, int dtype, int W, int C, int a0, int a1, int b0, int b1, double* out,
                      cudaStream_t st) {
    if (C < 1 || C > BS_THREADS) { set_error("band_stats: %d channels (1..%d supported)", C, BS_THREADS); return 3; }
    const size_t smem = (size_t)BS_THREADS * 2 * sizeof(double);
    switch (dtype) {
        case DG_F32: band_stats_kernel<float><<<1, BS_THREADS, smem, st>>>(kstats, static_cast<const float*>(t), W, C, a0, a1, b0, b1, out); break;
        case DG_F16: band_stats_kernel<__half><<<1, BS_THREADS, smem, st>>>(kstats, static_cast<const __half*>(t), W, C, a0, a1, b0, b1, out); break;
        case DG_BF16: band_stats_kernel<__nv_bfloat16><<<1, BS_THREADS, smem, st>>>(kstats, static_cast<const __nv_bfloat16*>(t), W, C, a0, a1, b0, b1, out); break;
        default: set_error("band_stats: bad dtype %d", dtype); return 2;
    }
    count_launch();
    return check_launch("band_stats");
}

int gn_affine_launch(const double* parts, int nparts, size_t stride, const float* gamma, const float* beta, int C, int groups, double plane,
                     float eps, float* coef, cudaStream_t st) {
    if (C > 2048) { set_error("gn_affine: %d channels (<= 2048 supported)", C); return 3; }
    gn_affine_kernel<<<1, C < 256 ? ((C + 31) / 32) * 32 : 256, (size_t)C * 2 * sizeof(double), st>>>(parts, nparts, stride, gamma, beta, C,
                                                                                                     groups, plane, eps, coef);
    count_launch();
    return check_launch("gn_affine");
}

}  // namespace dg
