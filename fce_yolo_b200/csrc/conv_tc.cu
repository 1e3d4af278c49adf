// tcgen05 implicit-GEMM convolution - placeholder until the kernel lands (next commit).
#include "common.cuh"
namespace fce {
bool conv2d_tc_supported(const fce_conv_desc*, const void*, const void*, const void*, const void*) { return false; }
int conv2d_tc(const fce_conv_desc*, const void*, const void*, const float*, const void*, void*, cudaStream_t) {
    return FCE_ERR_UNSUPPORTED;
}
}  // namespace fce
