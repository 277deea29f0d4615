// Internal C++ interface of the tcgen05 1x1 adaptation conv (see adapt_gemm.cu).
#pragma once

#include "common.cuh"

namespace dbev {

// x_cl: [batch, hw, c_in] (channels-last), w: [c_out, c_in], y: [batch, c_out, hw] (NCHW)
int adapt_conv1x1_forward(const float* x_cl, const float* w, const float* bias, int batch, int c_in,
                          int c_out, int hw, float* y, cudaStream_t stream);

}  // namespace dbev
