// tcgen05 / TMEM implicit-GEMM convolution (bf16 operands, fp32 accumulate).  PLACEHOLDER: the
// dispatcher reports "unsupported" so the executor falls back to the CUDA-core kernels.
#include "ops.cuh"

namespace spk {
bool conv_tc_supported(const ConvArgs &, int) { return false; }
int launch_conv_tc(const ConvArgs &, int, int, cudaStream_t) {
    set_error("tcgen05 conv path not built");
    return SPK_ERR_UNSUPPORTED;
}
}  // namespace spk
