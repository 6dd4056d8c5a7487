// tcgen05 (kind::tf32) implicit-GEMM convolution -- placeholder until the kernels land:
// every entry reports SE_ERR_UNSUPPORTED so that the dispatcher uses the fp32 FFMA kernels.
#include "common.cuh"
namespace se {
int conv_fwd_tc(const se_conv_desc*, const float*, const float*, const float*, const float*, float*, int, double*, cudaStream_t) { return SE_ERR_UNSUPPORTED; }
int conv_dgrad_tc(const se_conv_desc*, const float*, const float*, float*, float, cudaStream_t) { return SE_ERR_UNSUPPORTED; }
int conv_wgrad_tc(const se_conv_desc*, const float*, const float*, float*, float*, cudaStream_t) { return SE_ERR_UNSUPPORTED; }
// bit 0 conv fwd, bit 1 conv dgrad, bit 2 conv wgrad, bit 3 pairwise: which tcgen05 kernels are compiled in
int tc_capabilities() { return 8; }
int init_conv_tc() { return SE_OK; }
}  // namespace se
