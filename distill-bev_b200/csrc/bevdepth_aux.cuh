// shift_feature (BEV temporal alignment) and get_depth_loss of BEVDepth4D; see bevdepth_aux.cu.
#pragma once

#include "common.cuh"

namespace dbev {

int shift_feature_forward(const float* in, const float* tf, int n, int C, int h, int w, float* out,
                          cudaStream_t stream);
int shift_feature_backward(const float* grad_out, const float* tf, int n, int C, int h, int w, float* grad_in,
                           cudaStream_t stream);
size_t depth_loss_ws_bytes();
int depth_loss_forward(const float* logits, const float* depth_gt, int BN, int D, int HW, float dmin,
                       float dstep, float loss_weight, float* loss, void* ws, size_t ws_bytes,
                       cudaStream_t stream);
int depth_loss_backward(const float* logits, const float* depth_gt, int BN, int D, int HW, float dmin,
                        float dstep, float loss_weight, const float* grad_loss, float* grad_logits,
                        cudaStream_t stream);

}  // namespace dbev
