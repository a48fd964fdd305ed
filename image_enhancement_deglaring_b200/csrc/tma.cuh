// TMA (cp.async.bulk.tensor) plumbing: tensor-map encoding through the driver entry point (no -lcuda link dependency),
// mbarrier helpers and the tensor-tile load instruction.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace dg {

typedef CUresult (*PFN_tensorMapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                            const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                            CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline PFN_tensorMapEncodeTiled tma_encoder() {
    static PFN_tensorMapEncodeTiled fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_tensorMapEncodeTiled>(p);
    }
    return fn;
}

// NHWC activation tensor [N][H][W][C] of 16-bit values with C * 2 == 16 bytes per pixel (C = 8), viewed as a 3-D tensor of
// 8-byte elements {2W, H, N}: a box of (TW+2) x (TH+2) pixels is then TH+2 contiguous rows of 16 (TW+2) bytes -- the halo tile
// lands in shared memory exactly as the 16-byte-per-pixel channel plane the ldmatrix addressing wants, and pixels outside the
// image arrive as zeros (out-of-bound fill), i.e. as the conv's zero padding of a RAW tile.
// `epp` = 8-byte elements per pixel (2 for C = 8, 4 for C = 16: the latter lands pixel-major, 32 bytes per pixel).
inline bool tma_map_nhwc(CUtensorMap* map, const void* base, int N, int H, int W, int epp, int box_w, int box_h) {
    PFN_tensorMapEncodeTiled enc = tma_encoder();
    if (enc == nullptr) return false;
    if (epp * box_w > 256) return false;   // box dimensions are limited to 256 elements
    const cuuint64_t dims[3] = {(cuuint64_t)epp * W, (cuuint64_t)H, (cuuint64_t)N};
    const cuuint64_t strides[2] = {(cuuint64_t)W * epp * 8, (cuuint64_t)H * W * epp * 8};
    const cuuint32_t box[3] = {(cuuint32_t)epp * box_w, (cuuint32_t)box_h, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    return enc(map, CU_TENSOR_MAP_DATA_TYPE_INT64, 3, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

inline bool tma_map_nhwc16(CUtensorMap* map, const void* base, int N, int H, int W, int box_w, int box_h) {
    return tma_map_nhwc(map, base, N, H, W, 2, box_w, box_h);
}

#ifdef __CUDACC__
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(bar) : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}
#endif

}  // namespace dg
