// tcgen05 / TMEM / TMA implicit-GEMM convolution (bf16 operands, fp32 accumulate).
// Placeholder until the tensor-core path lands: reports "unsupported" so callers use conv_simt.
#include "common.cuh"
#include "kernels.h"

namespace vae2 {
int conv_tc_supported(const ConvGeom&) { return 0; }
int conv_fwd_tc(const void*, const void*, const float*, void*, const ConvGeom&, float*, cudaStream_t) {
    return VAE2_ERR_UNSUPPORTED;
}
}  // namespace vae2
