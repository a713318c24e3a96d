// tcgen05 / TMA convolution kernel for the 3x3x3 stride-1 layers -- placeholder until the kernel lands.
#include "conv_common.cuh"
int nnd_conv_tc_supported(const ConvGeom&, const ConvEpilogue&) { return 0; }
int nnd_conv_tc(const __nv_bfloat16*, const __nv_bfloat16*, const ConvGeom&, const ConvEpilogue&, cudaStream_t) { return NND_ERR_ARG; }
