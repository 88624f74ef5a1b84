// Tensor-core (tcgen05) path -- placeholder until the UMMA kernels land.
#pragma once
#include "model.h"
#include "fp32_kernels.cuh"

namespace hfg {

inline void tc_pack_conv(hfg_handle*, ConvLayer&, const HostTensor&, const HostTensor&) {}
inline void tc_pack_up(hfg_handle*, UpLayer&, const HostTensor&, const HostTensor&) {}
inline void tc_pack_post(hfg_handle*, const HostTensor&, const HostTensor&) {}
inline size_t tc_workspace_bytes(const hfg_handle*, int, int, int) {
    throw StatusError(HFG_ERR_UNSUPPORTED, "tensor-core modes not built yet");
}
inline void tc_forward(hfg_handle*, const float*, int, int, float*, char*, int, cudaStream_t,
                       float* const*) {
    throw StatusError(HFG_ERR_UNSUPPORTED, "tensor-core modes not built yet");
}
inline void configure_kernels(hfg_handle*) {
    const int smem = 100 * 1024;
    check_cuda(cudaFuncSetAttribute(conv_tile_fp32<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem), "attr");
    check_cuda(cudaFuncSetAttribute(conv_tile_fp32<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem), "attr");
    check_cuda(cudaFuncSetAttribute(conv_tile_fp32<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem), "attr");
}

}  // namespace hfg
