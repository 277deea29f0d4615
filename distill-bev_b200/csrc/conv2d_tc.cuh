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

// dx[n, 2*ho, 2*wo, dx_c_off + ...] (+)= input gradient of a 3x3 / stride 2 / pad 1 convolution, all four parity classes in
// one launch. dy NHWC [n, ho, wo, c_out_fwd] (channel stride dy_ld); w_mode2 = pack_conv_weights(..., mode 2) of the
// whole layer (c_in_fwd_total input channels); the launch computes col_width * n_col_blocks channels from dx_c_off.
int conv2d_tc_dgrad_s2(const float* dy_nhwc, int n_img, int ho, int wo, int c_out_fwd, int dy_ld, const float* w_mode2,
                       int c_in_fwd_total, int col_width, int n_col_blocks, float* dx, int dx_ld, int dx_c_off, int accumulate,
                       cudaStream_t stream);

}  // namespace dbev
