// Tensor-core (HMMA) implicit-GEMM path of dg_conv3x3_fused for 16-bit storage.
// Stage B -- not yet implemented: every configuration falls through to the generic CUDA-core kernel.
#include "common.cuh"

namespace dg {
int conv3x3_tc_launch(const dg_conv3x3_args& a, cudaStream_t stream, bool* handled) {
    (void)a; (void)stream;
    *handled = false;
    return 0;
}
}  // namespace dg
