// Internal C++ interface of the distillation-loss kernels (see distill_loss.cu).
#pragma once

#include "../../include/distill_bev_b200.h"
#include "common.cuh"

namespace dbev {

using FgdConfig = dbev_fgd_config;

int fgd_foreground_mask(const float* boxes, int box_dim, const int* box_offsets, int max_boxes,
                        int batch, int H, int W, float voxel_x, float voxel_y, float osf,
                        float pc_min_x, float pc_min_y, int cell_center, int transpose_mask,
                        float* fg, float* fg_scale, int* fg_count, cudaStream_t stream);

int heatmap_class_max(const float* hm, int batch, int K, int H, int W, int apply_clip_sigmoid,
                      float* out, cudaStream_t stream);

// fp_scale_mode 'dfs' (bevdet_distill.py:926-966): per-component scale 1 / (FIFO pops incl. re-queued cells)
size_t fgd_fp_dfs_ws_bytes(int batch, int H, int W);
int fgd_fp_dfs_scale(const float* fp, int batch, int H, int W, float* scale, void* ws, size_t ws_bytes,
                     cudaStream_t stream);

int fgd_fp_mask(const float* gt_max, int Sg, const float* teacher_max, int St,
                const float* student_max, int Ss, const float* fg, int R, int batch, int mode,
                float thres, float gt_thres, float* fp, int* fp_count, cudaStream_t stream);

size_t fgd_state_bytes(const FgdConfig& c);

int fgd_loss_forward(const FgdConfig& c, const float* student, const float* teacher,
                     const float* fg, const float* fg_scale, const int* fg_count, const float* fp,
                     const int* fp_count, const float* conv_w, const float* conv_b, void* state,
                     size_t state_bytes, float* losses, cudaStream_t stream);

int fgd_loss_backward(const FgdConfig& c, const float* student, const float* teacher,
                      const float* conv_w, const float* conv_b, void* state, size_t state_bytes,
                      const float* grad_losses, float* grad_student, float* grad_conv_w,
                      float* grad_conv_b, float* grad_channel_sum,
                      cudaStream_t stream);

// '1x1conv' adaptation fused with the loss (adapt_loss_tc.cu): x_cl [B, HW, C_in] channels-last student feature,
// adapt_w [C, C_in], adapt_b [C] or null; the adapted map is never materialised.
bool fgd_adapt_fused_supported(const FgdConfig& c, int c_in);

int fgd_adapt_loss_forward(const FgdConfig& c, const float* x_cl, int c_in, const float* adapt_w, const float* adapt_b,
                           const float* teacher, const float* fg, const float* fg_scale, const int* fg_count,
                           const float* fp, const int* fp_count, const float* conv_w, const float* conv_b, void* state,
                           size_t state_bytes, float* losses, cudaStream_t stream);

int fgd_adapt_loss_backward(const FgdConfig& c, const float* x_cl, int c_in, const float* adapt_w, const float* adapt_b,
                            const float* teacher, const float* conv_w, const float* conv_b, void* state,
                            size_t state_bytes, const float* grad_losses, float* grad_adapted_cl, float* grad_conv_w,
                            float* grad_conv_b, float* grad_channel_sum, cudaStream_t stream);

}  // namespace dbev
