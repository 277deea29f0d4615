// extern "C" boundary: forwards include/distill_bev_b200.h to the kernels.
#include "../../include/distill_bev_b200.h"

#include <stdarg.h>
#include <string.h>

#include "adapt_gemm.cuh"
#include "affinity.cuh"
#include "bev_encoder_ops.cuh"
#include "bev_pool.cuh"
#include "bevdepth_aux.cuh"
#include "center_targets.cuh"
#include "conv2d_tc.cuh"
#include "conv_wgrad_tc.cuh"
#include "distill_loss.cuh"
#include "ms_deform_attn.cuh"
#include "pillar.cuh"
#include "pillar_hard.cuh"
#include "sca_rebatch.cuh"
#include "sort.cuh"
#include "spconv.cuh"
#include "voxel_encoders.cuh"
#include "voxelize.cuh"

namespace dbev {

static thread_local char g_last_error[512] = "";

void set_last_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
  va_end(ap);
}

}  // namespace dbev

using namespace dbev;

extern "C" {

int dbev_abi_version(void) { return DBEV_ABI_VERSION; }
const char* dbev_last_error(void) { return g_last_error; }
const char* dbev_build_arch(void) { return "sm_100a"; }

int dbev_bev_pool_forward(int b, int d, int h, int w, int n, int c, int n_intervals,
                          const float* x, const int* geom_feats, const int* interval_starts,
                          const int* interval_lengths, float* out, int zero_out, void* stream) {
  return bev_pool_interval_forward(b, d, h, w, n, c, n_intervals, x, geom_feats, interval_starts,
                                   interval_lengths, out, zero_out, (cudaStream_t)stream);
}

int dbev_bev_pool_backward(int b, int d, int h, int w, int n, int c, int n_intervals,
                           const float* out_grad, const int* geom_feats,
                           const int* interval_starts, const int* interval_lengths, float* x_grad,
                           int zero_x_grad, void* stream) {
  return bev_pool_interval_backward(b, d, h, w, n, c, n_intervals, out_grad, geom_feats,
                                    interval_starts, interval_lengths, x_grad, zero_x_grad,
                                    (cudaStream_t)stream);
}

size_t dbev_bev_plan_workspace_bytes(long long n_points, long long n_cells) {
  return bev_plan_ws_bytes(n_points, n_cells);
}

long long dbev_bev_plan_max_items(long long n_points, long long n_cells, int nfast,
                                  int rows_per_item) {
  return bev_plan_max_items(n_points, n_cells, nfast < 1 ? 1 : nfast, rows_per_item);
}

int dbev_bev_plan_from_geom(const float* geom, long long n_points, int batch,
                            const float* off_host3, const float* dx_host3,
                            const float* nx_float_host3, const int* nx_int_host3, int fast_axis,
                            int rows_per_item, uint32_t* order, int* cell_start, int* cell_end,
                            int* items, long long max_items, int* n_items, int* point_cell,
                            void* workspace, size_t workspace_bytes, void* stream) {
  return bev_plan_from_geom(geom, n_points, batch, off_host3, dx_host3, nx_float_host3,
                            nx_int_host3, fast_axis, rows_per_item, order, cell_start, cell_end,
                            (int4*)items, max_items, n_items, point_cell, workspace,
                            workspace_bytes, (cudaStream_t)stream);
}

int dbev_bev_pool_point_backward(const float* grad_cl, const int* point_cell, long long n_points, int C,
                                 float* x_grad, void* stream) {
  return bev_pool_point_backward(grad_cl, point_cell, n_points, C, x_grad, (cudaStream_t)stream);
}

int dbev_lift_splat_forward(const float* depth, const float* feat_cl, int C, int D, int fhw,
                            const uint32_t* order, const int* cell_start, const int* cell_end,
                            const int* items, const int* n_items, int batch, int nz, int nslow,
                            int nfast, long long stride_b, long long stride_z, long long stride_c,
                            float* out, void* stream) {
  return lift_splat_forward(depth, feat_cl, C, D, fhw, order, cell_start, cell_end,
                            (const int4*)items, n_items, batch, nz, nslow, nfast, stride_b,
                            stride_z, stride_c, out, (cudaStream_t)stream);
}

int dbev_lift_splat_backward(const float* grad_cl, const float* depth, const float* feat_cl,
                             const int* point_cell, long long n_pixels, int C, int D, int fhw,
                             float* d_depth, float* d_feat_cl, void* stream) {
  return lift_splat_backward(grad_cl, depth, feat_cl, point_cell, n_pixels, C, D, fhw, d_depth,
                             d_feat_cl, (cudaStream_t)stream);
}

int dbev_bev_point_cells(const float* geom, long long n_points, int batch, const float* off_host3,
                         const float* dx_host3, const float* nx_float_host3, const int* nx_int_host3,
                         int fast_axis, int* point_cell, void* stream) {
  return bev_point_cells(geom, n_points, batch, off_host3, dx_host3, nx_float_host3, nx_int_host3, fast_axis,
                         point_cell, (cudaStream_t)stream);
}

int dbev_bev_point_cells_frames(const float* geom, long long n_points, int batch, int frames, const float* off_host3,
                                const float* dx_host3, const float* nx_float_host3, const int* nx_int_host3,
                                int fast_axis, int* point_cell, void* stream) {
  return bev_point_cells(geom, n_points, batch, off_host3, dx_host3, nx_float_host3, nx_int_host3, fast_axis,
                         point_cell, (cudaStream_t)stream, frames);
}

int dbev_lift_splat_atomic_forward(const float* depth, const float* feat_cl, const int* point_cell,
                                   long long n_pixels, int C, int D, int fhw, long long n_cells,
                                   float* out_cl, void* stream) {
  return lift_splat_atomic_forward(depth, feat_cl, point_cell, n_pixels, C, D, fhw, n_cells, out_cl,
                                   (cudaStream_t)stream);
}

int dbev_transpose_batched(const float* in, float* out, int batch, int rows, int cols,
                           void* stream) {
  return transpose_batched(in, out, batch, rows, cols, (cudaStream_t)stream);
}

int dbev_bev_plan_from_coords(const void* coords, int coords_is_i64, long long n_points,
                              int batch, int n0, int n1, int nz, int fast_axis, int rows_per_item,
                              uint32_t* order, int* cell_start, int* cell_end, int* items,
                              long long max_items, int* n_items, void* workspace,
                              size_t workspace_bytes, void* stream) {
  return bev_plan_from_coords(coords, coords_is_i64, n_points, batch, n0, n1, nz, fast_axis,
                              rows_per_item, order, cell_start, cell_end, (int4*)items, max_items,
                              n_items, workspace, workspace_bytes, (cudaStream_t)stream);
}

int dbev_bev_pool_gather_forward(const float* x, int C, const uint32_t* order,
                                 const int* cell_start, const int* cell_end, const int* items,
                                 const int* n_items, int batch, int nz, int nslow, int nfast,
                                 long long stride_b, long long stride_z, long long stride_c,
                                 float* out, void* stream) {
  return bev_pool_gather_forward(x, C, order, cell_start, cell_end, (const int4*)items, n_items,
                                 batch, nz, nslow, nfast, stride_b, stride_z, stride_c, out,
                                 (cudaStream_t)stream);
}

int dbev_bev_pool_gather_backward(const float* out_grad, int C, const uint32_t* order,
                                  const int* cell_start, const int* cell_end, const int* items,
                                  const int* n_items, int batch, int nz, int nslow, int nfast,
                                  long long stride_b, long long stride_z, long long stride_c,
                                  float* x_grad, void* stream) {
  return bev_pool_gather_backward(out_grad, C, order, cell_start, cell_end, (const int4*)items,
                                  n_items, batch, nz, nslow, nfast, stride_b, stride_z, stride_c,
                                  x_grad, (cudaStream_t)stream);
}

int dbev_voxel_grid_size(const float* voxel_size_host3, const float* coors_range_host6,
                         int* grid_xyz_host3) {
  return voxel_grid_size(voxel_size_host3, coors_range_host6, grid_xyz_host3);
}

int dbev_dynamic_voxelize(const float* points, int n, int nfeat, const float* voxel_size_host3,
                          const float* coors_range_host6, int* coors, void* stream) {
  return dynamic_voxelize(points, n, nfeat, voxel_size_host3, coors_range_host6, coors,
                          (cudaStream_t)stream);
}

size_t dbev_hard_voxelize_workspace_bytes(long long n) { return hard_voxelize_ws_bytes(n); }

int dbev_hard_voxelize(const float* points, int n, int nfeat, const float* voxel_size_host3,
                       const float* coors_range_host6, int max_points, int max_voxels,
                       float* voxels, int* coors, int* num_points_per_voxel, int* voxel_num,
                       void* workspace, size_t workspace_bytes, void* stream) {
  return hard_voxelize(points, n, nfeat, voxel_size_host3, coors_range_host6, max_points,
                       max_voxels, voxels, coors, num_points_per_voxel, voxel_num, workspace,
                       workspace_bytes, (cudaStream_t)stream);
}

size_t dbev_dynamic_scatter_workspace_bytes(long long n) { return dynamic_scatter_ws_bytes(n); }

int dbev_dynamic_scatter_forward(const float* feats, const int* coors, int n, int nfeat, int ncol,
                                 const int* dims_host, int reduce_type, float* reduced_feats,
                                 int* out_coors, int* coors_map, int* reduce_count, int* num_out,
                                 void* workspace, size_t workspace_bytes, void* stream) {
  return dynamic_scatter_forward(feats, coors, n, nfeat, ncol, dims_host, reduce_type,
                                 reduced_feats, out_coors, coors_map, reduce_count, num_out,
                                 workspace, workspace_bytes, (cudaStream_t)stream);
}

int dbev_dynamic_scatter_backward(const float* grad_reduced, const float* feats,
                                  const float* reduced, const int* coors_map,
                                  const int* reduce_count, long long n, long long m, int nfeat,
                                  int reduce_type, float* grad_feats, int* reduce_from_ws,
                                  void* stream) {
  return dynamic_scatter_backward(grad_reduced, feats, reduced, coors_map, reduce_count, n, m,
                                  nfeat, reduce_type, grad_feats, reduce_from_ws,
                                  (cudaStream_t)stream);
}

int dbev_fgd_foreground_mask(const float* boxes, int box_dim, const int* box_offsets,
                             int max_boxes_per_sample, int batch, int H, int W, float voxel_x,
                             float voxel_y, float out_size_factor, float pc_min_x, float pc_min_y,
                             int cell_center, int transpose_mask, float* fg, float* fg_scale,
                             int* fg_count, void* stream) {
  return fgd_foreground_mask(boxes, box_dim, box_offsets, max_boxes_per_sample, batch, H, W,
                             voxel_x, voxel_y, out_size_factor, pc_min_x, pc_min_y, cell_center,
                             transpose_mask, fg, fg_scale, fg_count, (cudaStream_t)stream);
}

int dbev_heatmap_class_max(const float* heatmaps, int batch, int K, int H, int W,
                           int apply_clip_sigmoid, float* out, void* stream) {
  return heatmap_class_max(heatmaps, batch, K, H, W, apply_clip_sigmoid, out, (cudaStream_t)stream);
}

size_t dbev_fgd_fp_dfs_workspace_bytes(int batch, int H, int W) { return fgd_fp_dfs_ws_bytes(batch, H, W); }

int dbev_fgd_fp_dfs_scale(const float* fp, int batch, int H, int W, float* scale, void* workspace,
                          size_t workspace_bytes, void* stream) {
  return fgd_fp_dfs_scale(fp, batch, H, W, scale, workspace, workspace_bytes, (cudaStream_t)stream);
}

int dbev_fgd_fp_mask(const float* gt_max, int Sg, const float* teacher_max, int St,
                     const float* student_max, int Ss, const float* fg, int R, int batch,
                     int mode, float thres, float gt_thres, float* fp, int* fp_count,
                     void* stream) {
  return fgd_fp_mask(gt_max, Sg, teacher_max, St, student_max, Ss, fg, R, batch, mode, thres,
                     gt_thres, fp, fp_count, (cudaStream_t)stream);
}

size_t dbev_fgd_state_bytes(const dbev_fgd_config* cfg) {
  if (!cfg) return 0;
  return fgd_state_bytes(*cfg);
}

int dbev_fgd_loss_forward(const dbev_fgd_config* cfg, const float* student, const float* teacher,
                          const float* fg, const float* fg_scale, const int* fg_count,
                          const float* fp, const int* fp_count, const float* conv_w,
                          const float* conv_b, void* state, size_t state_bytes, float* losses,
                          void* stream) {
  DBEV_CHECK_ARG(cfg != nullptr, "fgd: null config");
  return fgd_loss_forward(*cfg, student, teacher, fg, fg_scale, fg_count, fp, fp_count, conv_w,
                          conv_b, state, state_bytes, losses, (cudaStream_t)stream);
}

int dbev_fgd_loss_backward(const dbev_fgd_config* cfg, const float* student, const float* teacher,
                           const float* conv_w, const float* conv_b, void* state,
                           size_t state_bytes, const float* grad_losses, float* grad_student,
                           float* grad_conv_w, float* grad_conv_b, float* grad_channel_sum,
                           void* stream) {
  DBEV_CHECK_ARG(cfg != nullptr, "fgd: null config");
  return fgd_loss_backward(*cfg, student, teacher, conv_w, conv_b, state, state_bytes, grad_losses,
                           grad_student, grad_conv_w, grad_conv_b, grad_channel_sum,
                           (cudaStream_t)stream);
}

int dbev_fgd_adapt_supported(const dbev_fgd_config* cfg, int c_in) {
  return cfg != nullptr && fgd_adapt_fused_supported(*cfg, c_in) ? 1 : 0;
}

int dbev_fgd_adapt_loss_forward(const dbev_fgd_config* cfg, const float* x_cl, int c_in, const float* adapt_w,
                                const float* adapt_b, const float* teacher, const float* fg, const float* fg_scale,
                                const int* fg_count, const float* fp, const int* fp_count, const float* conv_w,
                                const float* conv_b, void* state, size_t state_bytes, float* losses, void* stream) {
  DBEV_CHECK_ARG(cfg != nullptr, "fgd: null config");
  return fgd_adapt_loss_forward(*cfg, x_cl, c_in, adapt_w, adapt_b, teacher, fg, fg_scale, fg_count, fp, fp_count, conv_w,
                                conv_b, state, state_bytes, losses, (cudaStream_t)stream);
}

int dbev_fgd_adapt_loss_backward(const dbev_fgd_config* cfg, const float* x_cl, int c_in, const float* adapt_w,
                                 const float* adapt_b, const float* teacher, const float* conv_w, const float* conv_b,
                                 void* state, size_t state_bytes, const float* grad_losses, float* grad_adapted_cl,
                                 float* grad_conv_w, float* grad_conv_b, float* grad_channel_sum, void* stream) {
  DBEV_CHECK_ARG(cfg != nullptr, "fgd: null config");
  return fgd_adapt_loss_backward(*cfg, x_cl, c_in, adapt_w, adapt_b, teacher, conv_w, conv_b, state, state_bytes,
                                 grad_losses, grad_adapted_cl, grad_conv_w, grad_conv_b, grad_channel_sum,
                                 (cudaStream_t)stream);
}

size_t dbev_pillar_encode_workspace_bytes(long long n) { return pillar_encode_ws_bytes(n); }

int dbev_pillar_encode(const float* points, const int* batch_offsets, const int* coors_in,
                       int batch, int n, int nfeat, const float* voxel_size_host3,
                       const float* coors_range_host6, float x_offset, float y_offset,
                       const float* weight, int nout, const float* bn_scale,
                       const float* bn_shift, float* voxel_feats, int* voxel_coors,
                       int* num_voxels, int* point_coors, void* workspace, size_t workspace_bytes,
                       void* stream) {
  return pillar_encode(points, batch_offsets, coors_in, batch, n, nfeat, voxel_size_host3,
                       coors_range_host6, x_offset, y_offset, weight, nout, bn_scale, bn_shift,
                       voxel_feats, voxel_coors, num_voxels, point_coors, nullptr, 0, 0, workspace,
                       workspace_bytes, (cudaStream_t)stream);
}

int dbev_pillar_canvas(const float* points, const int* batch_offsets, const int* coors_in,
                       int batch, int n, int nfeat, const float* voxel_size_host3,
                       const float* coors_range_host6, float x_offset, float y_offset,
                       const float* weight, int nout, const float* bn_scale,
                       const float* bn_shift, int channels_last, int zero_canvas, float* canvas,
                       int* num_voxels, void* workspace, size_t workspace_bytes, void* stream) {
  DBEV_CHECK_ARG(canvas != nullptr, "pillar_canvas: null canvas");
  return pillar_encode(points, batch_offsets, coors_in, batch, n, nfeat, voxel_size_host3,
                       coors_range_host6, x_offset, y_offset, weight, nout, bn_scale, bn_shift,
                       nullptr, nullptr, num_voxels, nullptr, canvas, channels_last, zero_canvas,
                       workspace, workspace_bytes, (cudaStream_t)stream);
}

int dbev_pillar_scatter(const float* voxel_feats, const int* coors, const int* m_dev, int m_max,
                        int C, int batch, int ny, int nx, int channels_last, int zero_canvas,
                        float* canvas, void* stream) {
  return pillar_scatter(voxel_feats, coors, m_dev, m_max, C, batch, ny, nx, channels_last,
                        zero_canvas, canvas, (cudaStream_t)stream);
}

int dbev_lss_geometry(const float* frustum, int pts_per_cam, const float* rots,
                      const float* trans, const float* intrins, const float* post_rots,
                      const float* post_trans, int n_cams, float* mats_ws, float* geom,
                      void* stream) {
  return lss_geometry(frustum, pts_per_cam, rots, trans, intrins, post_rots, post_trans, n_cams,
                      mats_ws, geom, (cudaStream_t)stream);
}

int dbev_adapt_conv1x1_forward(const float* x_cl, const float* w, const float* bias, int batch,
                               int c_in, int c_out, int hw, float* y, void* stream) {
  return adapt_conv1x1_forward(x_cl, w, bias, batch, c_in, c_out, hw, y, (cudaStream_t)stream);
}

static SpConvGeom geom_from_host(const int* g) {
  SpConvGeom r;
  for (int i = 0; i < 3; ++i) {
    r.k[i] = g[i], r.s[i] = g[3 + i], r.p[i] = g[6 + i], r.d[i] = g[9 + i];
    r.in_shape[i] = g[12 + i], r.out_shape[i] = g[15 + i];
  }
  r.batch = g[18];
  return r;
}

long long dbev_spconv_max_out(long long n_in, const int* geom_host19) {
  return spconv_max_out(n_in, geom_from_host(geom_host19));
}

size_t dbev_spconv_workspace_bytes(long long n_in, long long max_out) {
  return spconv_ws_bytes(n_in, max_out);
}

int dbev_spconv_table(const int* in_coors, int n_in, const int* out_coors, int n_out,
                      const int* geom_host19, int* nbr, void* workspace, size_t workspace_bytes,
                      void* stream) {
  return spconv_table(in_coors, n_in, out_coors, n_out, geom_from_host(geom_host19), nbr,
                      workspace, workspace_bytes, (cudaStream_t)stream);
}

int dbev_spconv_out_candidates(const int* in_coors, int n_in, const int* geom_host19,
                               uint32_t* out_keys, long long max_out, int* n_out,
                               void* workspace, size_t workspace_bytes, void* stream) {
  return spconv_out_candidates(in_coors, n_in, geom_from_host(geom_host19), out_keys, max_out,
                               n_out, workspace, workspace_bytes, (cudaStream_t)stream);
}

int dbev_spconv_out_table(const int* in_coors, int n_in, const int* geom_host19,
                          uint32_t* out_keys, int n_out, int* out_coors, int* nbr,
                          void* workspace, size_t workspace_bytes, void* stream) {
  return spconv_out_table(in_coors, n_in, geom_from_host(geom_host19), out_keys, n_out, out_coors,
                          nbr, workspace, workspace_bytes, (cudaStream_t)stream);
}

int dbev_spconv_pairs_from_table(const int* nbr, int kvol, int n_out, int pair_stride,
                                 int* indice_pairs, int* indice_num, void* stream) {
  return spconv_pairs_from_table(nbr, kvol, n_out, pair_stride, indice_pairs, indice_num,
                                 (cudaStream_t)stream);
}

int dbev_spconv_table_from_pairs(const int* indice_pairs, const int* indice_num, int kvol,
                                 int pair_stride, int inverse, int n_out, int* nbr, void* stream) {
  return spconv_table_from_pairs(indice_pairs, indice_num, kvol, pair_stride, inverse, n_out, nbr,
                                 (cudaStream_t)stream);
}

int dbev_spconv_forward(const float* in_feats, int c_in, const float* weight, int c_out,
                        const int* nbr, int kvol, int n_out, const float* scale,
                        const float* shift, const float* residual, int relu, float* out,
                        void* stream) {
  return spconv_forward(in_feats, c_in, weight, c_out, nbr, kvol, n_out, scale, shift, residual,
                        relu, out, (cudaStream_t)stream);
}

int dbev_shift_feature_forward(const float* in, const float* tf, int n, int C, int h, int w, float* out,
                               void* stream) {
  return shift_feature_forward(in, tf, n, C, h, w, out, (cudaStream_t)stream);
}

int dbev_shift_feature_backward(const float* grad_out, const float* tf, int n, int C, int h, int w,
                                float* grad_in, void* stream) {
  return shift_feature_backward(grad_out, tf, n, C, h, w, grad_in, (cudaStream_t)stream);
}

size_t dbev_depth_loss_workspace_bytes(void) { return depth_loss_ws_bytes(); }

int dbev_depth_loss_forward(const float* logits, const float* depth_gt, int BN, int D, int HW, float dmin,
                            float dstep, float loss_weight, float* loss, void* workspace,
                            size_t workspace_bytes, void* stream) {
  return depth_loss_forward(logits, depth_gt, BN, D, HW, dmin, dstep, loss_weight, loss, workspace,
                            workspace_bytes, (cudaStream_t)stream);
}

int dbev_depth_loss_backward(const float* logits, const float* depth_gt, int BN, int D, int HW, float dmin,
                             float dstep, float loss_weight, const float* grad_loss, float* grad_logits,
                             void* stream) {
  return depth_loss_backward(logits, depth_gt, BN, D, HW, dmin, dstep, loss_weight, grad_loss, grad_logits,
                             (cudaStream_t)stream);
}

int dbev_ms_deform_attn_forward(const float* value, const long long* spatial_shapes,
                                const long long* level_start, const float* sampling_loc,
                                const float* attn_weight, int bs, int num_keys, int heads, int dim,
                                int num_queries, int levels, int points, float* out, void* stream) {
  return ms_deform_attn_forward(value, spatial_shapes, level_start, sampling_loc, attn_weight, bs, num_keys,
                                heads, dim, num_queries, levels, points, out, (cudaStream_t)stream);
}

int dbev_ms_deform_attn_backward(const float* value, const long long* spatial_shapes,
                                 const long long* level_start, const float* sampling_loc,
                                 const float* attn_weight, const float* grad_out, int bs, int num_keys,
                                 int heads, int dim, int num_queries, int levels, int points,
                                 float* grad_value, float* grad_loc, float* grad_attn, void* stream) {
  return ms_deform_attn_backward(value, spatial_shapes, level_start, sampling_loc, attn_weight, grad_out, bs,
                                 num_keys, heads, dim, num_queries, levels, points, grad_value, grad_loc,
                                 grad_attn, (cudaStream_t)stream);
}

int dbev_center_targets(const float* boxes, int box_dim, const int* labels, const int* offsets, int batch,
                        const int* class_task_host, const int* class_in_task_host, int num_classes,
                        int num_tasks, int max_objs, int H, int W, float voxel_x, float voxel_y,
                        float out_size_factor, float pc_min_x, float pc_min_y, float gaussian_overlap,
                        int min_radius, int norm_bbox, float* heatmap, float* anno_box,
                        long long* ind, unsigned char* mask, void* stream) {
  return center_targets(boxes, box_dim, labels, offsets, batch, class_task_host, class_in_task_host,
                        num_classes, num_tasks, max_objs, H, W, voxel_x, voxel_y, out_size_factor, pc_min_x,
                        pc_min_y, gaussian_overlap, min_radius, norm_bbox, heatmap, anno_box, ind, mask,
                        (cudaStream_t)stream);
}

int dbev_conv2d_tc_forward(const float* x_nhwc, int n, int h, int w, int c_in, const float* w_packed,
                           int c_out, int kh, int kw, int stride, int pad, const float* scale,
                           const float* shift, int relu, float* out, int out_h, int out_w,
                           int out_ld, int out_c_off, int out_mul, int out_add_y, int out_add_x,
                           int out_nchw, void* stream) {
  return conv2d_tc_forward(x_nhwc, n, h, w, c_in, w_packed, c_out, kh, kw, stride, pad, scale, shift,
                           relu, out, out_h, out_w, out_ld, out_c_off, out_mul, out_add_y, out_add_x,
                           out_nchw, 1, (cudaStream_t)stream);
}

int dbev_conv2d_tc_forward_grouped(const float* x_nhwc, int n, int h, int w, int c_in, const float* w_packed,
                                   int c_out, int kh, int kw, int stride, int pad, const float* scale,
                                   const float* shift, int relu, float* out, int out_h, int out_w,
                                   int out_ld, int out_c_off, int out_mul, int out_add_y, int out_add_x,
                                   int out_nchw, int out_groups, void* stream) {
  return conv2d_tc_forward(x_nhwc, n, h, w, c_in, w_packed, c_out, kh, kw, stride, pad, scale, shift,
                           relu, out, out_h, out_w, out_ld, out_c_off, out_mul, out_add_y, out_add_x,
                           out_nchw, out_groups, (cudaStream_t)stream);
}

int dbev_conv2d_tc_forward_ex(const float* x_nhwc, int n, int h, int w, int c_in, int x_ld, const float* w_packed,
                              int c_out, int n_col_blocks, int kh, int kw, int stride, int pad, const float* scale,
                              const float* shift, int relu, float* out, int out_h, int out_w,
                              int out_ld, int out_c_off, int out_mul, int out_add_y, int out_add_x,
                              int out_nchw, int out_groups, int force_ho, int force_wo, int accumulate,
                              void* stream) {
  return conv2d_tc_forward_ex(x_nhwc, n, h, w, c_in, x_ld, w_packed, c_out, n_col_blocks, kh, kw, stride, pad, scale, shift,
                              relu, out, out_h, out_w, out_ld, out_c_off, out_mul, out_add_y, out_add_x,
                              out_nchw, out_groups, force_ho, force_wo, accumulate, (cudaStream_t)stream);
}

size_t dbev_conv_wgrad_tc_workspace_bytes(int n, int ho, int wo, int c_in, int c_out, int kh, int kw, int stride) {
  return conv_wgrad_tc_workspace_bytes(n, ho, wo, c_in, c_out, kh, kw, stride);
}

int dbev_conv_wgrad_tc(const float* x_nhwc, int n, int h, int w, int c_in, int x_ld, const float* dy_nhwc,
                       int ho, int wo, int c_out, int dy_ld, int kh, int kw, int stride, int pad, float* dw,
                       int accumulate, void* workspace, size_t workspace_bytes, void* stream) {
  return conv_wgrad_tc(x_nhwc, n, h, w, c_in, x_ld, dy_nhwc, ho, wo, c_out, dy_ld, kh, kw, stride, pad, dw,
                       accumulate, workspace, workspace_bytes, (cudaStream_t)stream);
}

int dbev_conv2d_tc_dgrad_s2(const float* dy_nhwc, int n, int ho, int wo, int c_out, int dy_ld, const float* w_mode2,
                            int c_in_total, int col_width, int n_col_blocks, float* dx, int dx_ld, int dx_c_off, int accumulate,
                            void* stream) {
  return conv2d_tc_dgrad_s2(dy_nhwc, n, ho, wo, c_out, dy_ld, w_mode2, c_in_total, col_width, n_col_blocks, dx, dx_ld, dx_c_off,
                            accumulate, (cudaStream_t)stream);
}

int dbev_pack_conv_weights(const float* w, int c_out, int c_in, int kh, int kw, int mode, float* out, void* stream) {
  return pack_conv_weights(w, c_out, c_in, kh, kw, mode, out, (cudaStream_t)stream);
}

int dbev_pack_conv_weights_train(const float* w, int c_out, int c_in, int kh, int kw, int dgrad_mode, float* out_fwd,
                                 float* out_dgrad, void* stream) {
  return pack_conv_weights_train(w, c_out, c_in, kh, kw, dgrad_mode, out_fwd, out_dgrad, (cudaStream_t)stream);
}

int dbev_pack_conv_weights_batch(const long long* jobs_dev, int n_jobs, int total_tiles, void* stream) {
  return pack_conv_weights_batch(jobs_dev, n_jobs, total_tiles, (cudaStream_t)stream);
}

size_t dbev_channel_stats_workspace_bytes(long long rows, int C) { return channel_stats_workspace_bytes(rows, C); }

int dbev_bn_batch_stats(const float* y, int y_ld, long long rows, int C, const float* gamma, const float* beta,
                        float eps, float momentum, float* running_mean, float* running_var, float* out4c,
                        void* workspace, size_t workspace_bytes, void* stream) {
  return bn_batch_stats(y, y_ld, rows, C, gamma, beta, eps, momentum, running_mean, running_var, out4c, workspace,
                        workspace_bytes, (cudaStream_t)stream);
}

int dbev_channel_sums(const float* y, int y_ld, long long rows, int C, float* out, int accumulate,
                      void* workspace, size_t workspace_bytes, void* stream) {
  return channel_sums(y, y_ld, rows, C, out, accumulate, workspace, workspace_bytes, (cudaStream_t)stream);
}

int dbev_bn_act_forward(const float* y, int y_ld, const float* ab, const float* residual, int res_ld,
                        long long rows, int C, int relu, float* out, int out_ld, unsigned char* relu_mask, void* stream) {
  return bn_act_forward(y, y_ld, ab, residual, res_ld, rows, C, relu, out, out_ld, relu_mask, (cudaStream_t)stream);
}

int dbev_bn_backward(const float* dz, int dz_ld, const float* z, int z_ld, const float* y, int y_ld,
                     const float* fwd4c, long long rows, int C, float* bwd4c, float* dy, int dy_ld,
                     float* g_out, int g_ld, int g_accumulate, const unsigned char* relu_mask, void* workspace,
                     size_t workspace_bytes, void* stream) {
  return bn_backward(dz, dz_ld, z, z_ld, y, y_ld, fwd4c, rows, C, bwd4c, dy, dy_ld, g_out, g_ld, g_accumulate, relu_mask,
                     workspace, workspace_bytes, (cudaStream_t)stream);
}

int dbev_relu_mask_backward(const float* dz, int dz_ld, const float* z, int z_ld, long long rows, int C,
                            float* g_out, int g_ld, int accumulate, const unsigned char* relu_mask, void* stream) {
  return relu_mask_backward(dz, dz_ld, z, z_ld, rows, C, g_out, g_ld, accumulate, relu_mask, (cudaStream_t)stream);
}

int dbev_upsample_bilinear_forward(const float* in, int in_ld, int n, int h, int w, int C, int H, int W,
                                   float* out, int out_ld, void* stream) {
  return upsample_bilinear_forward(in, in_ld, n, h, w, C, H, W, out, out_ld, (cudaStream_t)stream);
}

int dbev_upsample_bilinear_backward(const float* dout, int dout_ld, int n, int h, int w, int C, int H, int W,
                                    float* din, int din_ld, int accumulate, void* stream) {
  return upsample_bilinear_backward(dout, dout_ld, n, h, w, C, H, W, din, din_ld, accumulate, (cudaStream_t)stream);
}

int dbev_sca_gather_rows(const float* in, const int* idx, const float* scale, int bs, int cams, int max_len, int nq, int C,
                         long long in_cam_stride, long long in_batch_stride, long long in_query_stride, float* out, void* stream) {
  return sca_gather_rows(in, idx, scale, bs, cams, max_len, nq, C, in_cam_stride, in_batch_stride, in_query_stride, out,
                         (cudaStream_t)stream);
}

int dbev_sca_reduce_rows(const float* in, const int* pos, const float* scale, int bs, int cams, int max_len, int nq, int C,
                         float* out, void* stream) {
  return sca_reduce_rows(in, pos, scale, bs, cams, max_len, nq, C, out, (cudaStream_t)stream);
}

int dbev_hard_pillar_encode(const float* voxels, const int* num_points, const int* coors, const int* m_dev, int m_max,
                            int max_points, int nfeat, const float* voxel_size_xy_host2, float x_offset, float y_offset,
                            const float* weight, int nout, const float* bn_scale, const float* bn_shift, int legacy,
                            float* out, void* stream) {
  return hard_pillar_encode(voxels, num_points, coors, m_dev, m_max, max_points, nfeat, voxel_size_xy_host2, x_offset,
                            y_offset, weight, nout, bn_scale, bn_shift, legacy, out, (cudaStream_t)stream);
}

int dbev_spconv_tc_supported(int c_in, int c_out, int kvol) {
  return spconv_tc_supported(c_in, c_out, kvol) ? 1 : 0;
}

int dbev_spconv_pack_weights(const float* weight, int kvol, int c_in, int c_out, float* wt_hi,
                             float* wt_lo, void* stream) {
  return spconv_pack_weights(weight, kvol, c_in, c_out, wt_hi, wt_lo, (cudaStream_t)stream);
}

int dbev_spconv_forward_tc(const float* in_feats, int c_in, const float* wt_hi, const float* wt_lo,
                           int c_out, const int* nbr, int kvol, int n_out, const float* scale,
                           const float* shift, const float* residual, int relu, float* out,
                           void* stream) {
  return spconv_forward_tc(in_feats, c_in, wt_hi, wt_lo, c_out, nbr, kvol, n_out, scale, shift,
                           residual, relu, out, (cudaStream_t)stream);
}

int dbev_spconv_dense(const float* feats, const int* coors, int m, int C, int batch, int Z,
                      int Y, int X, float* dense, void* stream) {
  return spconv_dense(feats, coors, m, C, batch, Z, Y, X, dense, (cudaStream_t)stream);
}

int dbev_hard_simple_vfe(const float* voxels, const int* num_points, long long m, int max_points,
                         int nfeat, int num_features, float* out, void* stream) {
  return hard_simple_vfe(voxels, num_points, m, max_points, nfeat, num_features, out,
                         (cudaStream_t)stream);
}

int dbev_dynvoxel_coords(const float* points, int n, int nfeat, const int* batch_offsets,
                         int batch, const float* pc_range_host6, const float* voxel_size_host3,
                         int check_flag, int* coors, void* stream) {
  return dynvoxel_coords(points, n, nfeat, batch_offsets, batch, pc_range_host6, voxel_size_host3,
                         check_flag, coors, (cudaStream_t)stream);
}

int dbev_dynvoxel_virtual_rows(const float* points, int n, int nfeat, float* rows24,
                               void* stream) {
  return dynvoxel_virtual_rows(points, n, nfeat, rows24, (cudaStream_t)stream);
}

int dbev_dynvoxel_virtual_fix(const float* mean24, const int* m_dev, int m_max, float* out23,
                              void* stream) {
  return dynvoxel_virtual_fix(mean24, m_dev, m_max, out23, (cudaStream_t)stream);
}

size_t dbev_affinity_select_workspace_bytes(int batch, int hw) {
  return affinity_select_ws_bytes(batch, hw);
}

int dbev_affinity_select(const float* mask_a, const float* mask_b, int batch, int hw,
                         int* row_cell, int* row_offsets, void* workspace, size_t workspace_bytes,
                         void* stream) {
  return affinity_select(mask_a, mask_b, batch, hw, row_cell, row_offsets, workspace,
                         workspace_bytes, (cudaStream_t)stream);
}

int dbev_affinity_gather_rows(const float* feat, const int* row_cell, const int* row_offsets,
                              int batch, int C, int hw, int k_total, float* rows, void* stream) {
  return affinity_gather_rows(feat, row_cell, row_offsets, batch, C, hw, k_total, rows,
                              (cudaStream_t)stream);
}

size_t dbev_affinity_partial_floats(const int* row_offsets_host, int batch) {
  return affinity_partial_floats(row_offsets_host, batch);
}

int dbev_affinity_forward(const float* t_rows, const float* s_rows, const int* row_offsets_host,
                          int batch, int C, int kind, float beta, float weight, float* partial,
                          float* loss, void* stream) {
  return affinity_forward(t_rows, s_rows, row_offsets_host, batch, C, kind, beta, weight, partial,
                          loss, (cudaStream_t)stream);
}

int dbev_affinity_backward(const float* t_rows, const float* s_rows, const int* row_offsets_host,
                           int batch, int C, int kind, float beta, float weight,
                           const float* grad_loss, float* d_s_rows, void* stream) {
  return affinity_backward(t_rows, s_rows, row_offsets_host, batch, C, kind, beta, weight,
                           grad_loss, d_s_rows, (cudaStream_t)stream);
}

int dbev_affinity_scatter_rows(const float* d_rows, const int* row_cell, const int* row_offsets,
                               int batch, int C, int hw, int k_total, float* grad, void* stream) {
  return affinity_scatter_rows(d_rows, row_cell, row_offsets, batch, C, hw, k_total, grad,
                               (cudaStream_t)stream);
}

size_t dbev_sort_workspace_bytes(long long n) {
  return radix_sort_ws_bytes(n) + 2 * align_up((size_t)(n > 0 ? n : 1) * 4) + 1024;
}

int dbev_sort_keys_iota(const uint32_t* keys_in, int n, int num_bits, uint32_t* keys_out,
                        uint32_t* order_out, void* workspace, size_t workspace_bytes,
                        void* stream) {
  DBEV_CHECK_ARG(n >= 0 && num_bits >= 1 && num_bits <= 32, "sort: bad n=%d num_bits=%d", n,
                 num_bits);
  if (n == 0) return DBEV_OK;
  cudaStream_t s = (cudaStream_t)stream;
  Workspace w(workspace, workspace_bytes);
  uint32_t* kt = w.take<uint32_t>(n);
  uint32_t* vt = w.take<uint32_t>(n);
  if (!w.ok()) {
    set_last_error("sort: workspace too small");
    return DBEV_ERR_WORKSPACE;
  }
  size_t consumed = align_up(w.used);
  const int passes = (num_bits + kRadixBits - 1) / kRadixBits;
  // the result must land in keys_out/order_out: seed so parity works out
  uint32_t* keys[2];
  uint32_t* vals[2];
  keys[passes & 1] = keys_out;
  keys[(passes & 1) ^ 1] = kt;
  vals[passes & 1] = order_out;
  vals[(passes & 1) ^ 1] = vt;
  DBEV_CUDA(cudaMemcpyAsync(keys[0], keys_in, (size_t)n * 4, cudaMemcpyDeviceToDevice, s));
  int sel = 0;
  return radix_sort_pairs(keys, vals, true, n, num_bits, (char*)workspace + consumed,
                          workspace_bytes - consumed, s, &sel);
}

size_t dbev_scan_workspace_bytes(long long n) { return scan_ws_bytes(n); }

int dbev_exclusive_scan_i32(const int* in, int* out, int n, int* total_out, void* workspace,
                            size_t workspace_bytes, void* stream) {
  return exclusive_scan_i32(in, out, n, total_out, workspace, workspace_bytes,
                            (cudaStream_t)stream);
}

}  // extern "C"
