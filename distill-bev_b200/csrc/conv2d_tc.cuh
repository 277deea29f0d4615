// Dense conv + folded BN + ReLU on tcgen05 for the frozen teacher's SECOND / SECONDFPN; see conv2d_tc.cu.
#pragma once

#include "common.cuh"

namespace dbev {

int conv2d_tc_forward(const float* x_nhwc, int n_img, int h, int w, int c_in, const float* w_packed,
                      int c_out, int kh, int kw, int stride, int pad, const float* scale,
                      const float* shift, int relu, float* out, int out_h, int out_w, int out_ld,
                      int out_c_off, int out_mul, int out_add_y, int out_add_x, int out_nchw,
                      int out_groups, cudaStream_t stream);

int conv2d_tc_forward_ex(const float* x_nhwc, int n_img, int h, int w, int c_in, int x_ld, const float* w_packed,
                         int c_out, int n_col_blocks, int kh, int kw, int stride, int pad, const float* scale,
                         const float* shift, int relu, float* out, int out_h, int out_w, int out_ld,
                         int out_c_off, int out_mul, int out_add_y, int out_add_x, int out_nchw,
                         int out_groups, int force_ho, int force_wo, int accumulate, cudaStream_t stream);

}  // namespace dbev
