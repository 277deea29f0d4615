// BatchNorm (batch statistics) / ReLU / residual / bilinear upsampling / weight packing for the trained
// student BEV encoder; see bev_encoder_ops.cu.
#pragma once

#include "common.cuh"

namespace dbev {

size_t channel_stats_workspace_bytes(long long rows, int C);

int bn_batch_stats(const float* y, int y_ld, long long rows, int C, const float* gamma, const float* beta, float eps,
                   float momentum, float* running_mean, float* running_var, float* out4c, void* workspace,
                   size_t workspace_bytes, cudaStream_t stream);

int channel_sums(const float* y, int y_ld, long long rows, int C, float* out, int accumulate, void* workspace,
                 size_t workspace_bytes, cudaStream_t stream);

int bn_act_forward(const float* y, int y_ld, const float* ab, const float* residual, int res_ld, long long rows, int C, int relu,
                   float* out, int out_ld, unsigned char* relu_mask, cudaStream_t stream);

int bn_backward(const float* dz, int dz_ld, const float* z, int z_ld, const float* y, int y_ld, const float* fwd4c, long long rows,
                int C, float* bwd4c, float* dy, int dy_ld, float* g_out, int g_ld, int g_accumulate, const unsigned char* relu_mask,
                void* workspace, size_t workspace_bytes, cudaStream_t stream);

int relu_mask_backward(const float* dz, int dz_ld, const float* z, int z_ld, long long rows, int C, float* g_out, int g_ld,
                       int accumulate, const unsigned char* relu_mask, cudaStream_t stream);

int upsample_bilinear_forward(const float* in, int in_ld, int n_img, int h, int w, int C, int H, int W, float* out, int out_ld,
                              cudaStream_t stream);

int upsample_bilinear_backward(const float* dout, int dout_ld, int n_img, int h, int w, int C, int H, int W, float* din, int din_ld,
                               int accumulate, cudaStream_t stream);

int pack_conv_weights(const float* w, int c_out, int c_in, int kh, int kw, int mode, float* out, cudaStream_t stream);

int pack_conv_weights_train(const float* w, int c_out, int c_in, int kh, int kw, int dgrad_mode, float* out_fwd, float* out_dgrad,
                            cudaStream_t stream);

// Every layer in one launch. jobs_dev: n_jobs records of 8 int64 in device memory
//   {w, out_fwd (0 = skip), out_dgrad (0 = skip), C_out, C_in, KH * 256 + KW, dgrad_mode, first tile}
// with tiles = ceil(C_out / 32) * ceil(C_in / 32) per job numbered consecutively; total_tiles = their sum.
int pack_conv_weights_batch(const long long* jobs_dev, int n_jobs, int total_tiles, cudaStream_t stream);

}  // namespace dbev
