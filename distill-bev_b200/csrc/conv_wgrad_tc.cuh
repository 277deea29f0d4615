// Weight gradient of a dense 2D convolution on tcgen05 (MN-major TF32 operands); see conv_wgrad_tc.cu.
#pragma once

#include "common.cuh"

namespace dbev {

size_t conv_wgrad_tc_workspace_bytes(int n_img, int ho, int wo, int c_in, int c_out, int kh, int kw, int stride);

int conv_wgrad_tc(const float* x_nhwc, int n_img, int h, int w, int c_in, int x_ld, const float* dy_nhwc, int ho,
                  int wo, int c_out, int dy_ld, int kh, int kw, int stride, int pad, float* dw, int accumulate,
                  void* workspace, size_t workspace_bytes, cudaStream_t stream);

}  // namespace dbev
