/*
 * distill_bev_b200.h — C-ABI of the B200-native DistillBEV hot path.
 *
 * This is the drop-in boundary: plain pointers and sizes, no
 * torch types. Every pointer that is not marked "host" is a DEVICE pointer on
 * the current CUDA device; `stream` is a cudaStream_t passed as void* (the
 * caller's current stream — the library never uses another stream and never
 * synchronises the device). Outputs and workspaces are allocated by the
 * caller (PyTorch's caching allocator in the Python shim).
 *
 * Error convention (replaces the reference's TORCH_CHECK / AT_ERROR C++
 * exceptions, mmdet3d/ops/voxel/src/voxelization.h:118,139): every function
 * returns 0 on success or a DBEV_ERR_* code; dbev_last_error() returns the
 * message of the last failure on the calling thread. The Python shim raises
 * RuntimeError with that text.
 *
 * Each entry point cites the reference interface it replaces (paths relative
 * to the reference checkout, qcraftai/distill-bev @ 3e8f6a4).
 */
#ifndef DISTILL_BEV_B200_H_
#define DISTILL_BEV_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DBEV_OK 0
#define DBEV_ERR_INVALID 1   /* bad argument */
#define DBEV_ERR_CUDA 2      /* CUDA runtime / launch failure */
#define DBEV_ERR_WORKSPACE 3 /* workspace too small */

#define DBEV_ABI_VERSION 1

int dbev_abi_version(void);
const char* dbev_last_error(void);
/* "sm_100a" — the only architecture this library carries code for. */
const char* dbev_build_arch(void);

/* ------------------------------------------------------------------------ *
 * bev_pool — reference-ABI launchers
 * Replace:  void bev_pool(...) / void bev_pool_grad(...)
 *           mmdet3d/ops/bev_pool/src/bev_pool.cpp:5-9 (declarations),
 *           mmdet3d/ops/bev_pool/src/bev_pool_cuda.cu:86-98 (launchers),
 *           bound to Python by bev_pool_forward/backward (bev_pool.cpp:22-87).
 * Same argument meaning: x[n,c] rows PRE-SORTED by rank, geom_feats[n,4] int32
 * (x, y, z, b), interval_starts/lengths[n_intervals]; out is [b,d,h,w,c]
 * addressed as out[g3][g2][g0][g1][:]. The reference launcher expects `out`
 * zero-filled by its caller (bev_pool.cpp:40); pass zero_out=1 to have the
 * fill enqueued here instead. Unlike the reference these run on `stream`, not
 * on the legacy default stream (bev_pool_cuda.cu:88,95).
 * ------------------------------------------------------------------------ */
int dbev_bev_pool_forward(int b, int d, int h, int w, int n, int c, int n_intervals,
                          const float* x, const int* geom_feats, const int* interval_starts,
                          const int* interval_lengths, float* out, int zero_out, void* stream);

int dbev_bev_pool_backward(int b, int d, int h, int w, int n, int c, int n_intervals,
                           const float* out_grad, const int* geom_feats,
                           const int* interval_starts, const int* interval_lengths, float* x_grad,
                           int zero_x_grad, void* stream);

/* ------------------------------------------------------------------------ *
 * bev_pool — B200 plan/gather path (what the Python shims actually call)
 *
 * A "plan" is the sorted view of the frustum points over the BEV grid:
 *   order[n_points]        point ids, stably sorted by output cell
 *   cell_start[n_cells+1]  first position in order[] of every cell
 *   cell_end[n_cells+1]    one past the last; slot n_cells = dropped points
 *   items[max_items][4]    work list: (tile, c0 | c1 << 8, row_lo, row_hi) = cells
 *                          [c0, c1) of a 32-cell tile and their rows of order[];
 *                          each item holds <= ~rows_per_item rows so that warps
 *                          get equal work although BEV occupancy is extremely
 *                          skewed (16 B aligned); *n_items = count
 * with cell = ((b*nz + iz)*nslow + islow)*nfast + ifast. fast_axis selects
 * which of the first two coordinates is the output's fastest axis:
 *   0 -> x fastest: final[b, iz*C + c, iy, ix]  (voxel_pooling,
 *        mmdet3d/models/necks/view_transformer_mine.py:176-179)
 *   1 -> y fastest: out[b, c, z, x, y]          (bev_pool + permute,
 *        mmdet3d/ops/bev_pool/bev_pool.py:96)
 * The plan depends on geometry only (camera calibration + augmentation), so
 * it can be cached across calls that share it. rows_per_item <= 0 selects the
 * default (192).
 * ------------------------------------------------------------------------ */
size_t dbev_bev_plan_workspace_bytes(long long n_points, long long n_cells);
/* required capacity (entries) of items[] */
long long dbev_bev_plan_max_items(long long n_points, long long n_cells, int nfast,
                                  int rows_per_item);

/* Replaces the index math + mask + rank + argsort of voxel_pooling
 * (view_transformer_mine.py:150-168): geom[n_points,3] fp32 ego-frame xyz,
 * off = bx - dx/2, dx, nx as fp32 (host float[3]) and nx.to(long) (host
 * int[3]); points are batch-major, n_points/batch per sample. */
int dbev_bev_plan_from_geom(const float* geom, long long n_points, int batch,
                            const float* off_host3, const float* dx_host3,
                            const float* nx_float_host3, const int* nx_int_host3, int fast_axis,
                            int rows_per_item, uint32_t* order, int* cell_start, int* cell_end,
                            int* items, long long max_items, int* n_items, int* point_cell,
                            void* workspace, size_t workspace_bytes, void* stream);
/* point_cell (nullable): [n_points] cell id of every point, -1 when dropped; needed by
 * dbev_lift_splat_backward. */

/* Replaces the rank + argsort + kept/where prelude of bev_pool()
 * (mmdet3d/ops/bev_pool/bev_pool.py:86-93,40-46). coords[n,4] = (c0,c1,c2,b),
 * int64 (coords_is_i64=1) or int32; grid n0 x n1 x nz per sample. Rows whose
 * coordinates fall outside the grid are dropped (the reference would write
 * out of bounds). */
int dbev_bev_plan_from_coords(const void* coords, int coords_is_i64, long long n_points,
                              int batch, int n0, int n1, int nz, int fast_axis, int rows_per_item,
                              uint32_t* order, int* cell_start, int* cell_end, int* items,
                              long long max_items, int* n_items, void* workspace,
                              size_t workspace_bytes, void* stream);

/* out[b*sB + iz*sZ + c*sC + islow*nfast + ifast] = sum of x[p, c] over the
 * cell's points, 0 for empty cells (every output element is written once:
 * no zero-fill, no permute copy). x is [n_points, C] fp32, rows 16 B aligned
 * when C % 4 == 0. Replaces cumsum/select/diff/scatter + cat(unbind)
 * (view_transformer_mine.py:171-179) and bev_pool_kernel + permute. */
int dbev_bev_pool_gather_forward(const float* x, int C, const uint32_t* order,
                                 const int* cell_start, const int* cell_end, const int* items,
                                 const int* n_items, int batch, int nz, int nslow, int nfast,
                                 long long stride_b, long long stride_z, long long stride_c,
                                 float* out, void* stream);

/* x_grad[p, :] = out_grad[cell(p), :], zero rows for dropped points (every
 * row written once). Replaces QuickCumsum.backward
 * (view_transformer_mine.py:48-56) / bev_pool_grad_kernel. */
int dbev_bev_pool_gather_backward(const float* out_grad, int C, const uint32_t* order,
                                  const int* cell_start, const int* cell_end, const int* items,
                                  const int* n_items, int batch, int nz, int nslow, int nfast,
                                  long long stride_b, long long stride_z, long long stride_c,
                                  float* x_grad, void* stream);

/* Point-centric form of the same backward for plans that carry point_cell (dbev_bev_plan_from_geom):
 * x_grad[p, :] = grad_cl[point_cell[p], :], zero rows for dropped points; grad_cl[n_cells, C] = the BEV
 * gradient in cells-major rows (dbev_transpose_batched of the NCHW gradient). Sequential row writes
 * instead of a scatter through the sorted order. C % 4 == 0. */
int dbev_bev_pool_point_backward(const float* grad_cl, const int* point_cell, long long n_points, int C,
                                 float* x_grad, void* stream);

/* ------------------------------------------------------------------------ *
 * Fused lift + splat. Replaces, in one kernel, the outer
 * product volume = depth.unsqueeze(1) * img_feat.unsqueeze(2), its permute to
 * channels-last (view_transformer_mine.py:333-335 = bevdet_distill_more.py:413-416)
 * and voxel_pooling (:141-181): the [B,N,D,fH,fW,C] volume (64-510 MB per
 * sample and frame) is never written or read.
 *   depth    [BN, D, fH, fW] fp32 depth distribution (flat index = point id)
 *   feat_cl  [BN * fH * fW, C] fp32 image features, channels-last
 * plan: from dbev_bev_plan_from_geom on the same frustum (n_points = BN*D*fH*fW).
 * ------------------------------------------------------------------------ */
int dbev_lift_splat_forward(const float* depth, const float* feat_cl, int C, int D, int fhw,
                            const uint32_t* order, const int* cell_start, const int* cell_end,
                            const int* items, const int* n_items, int batch, int nz, int nslow,
                            int nfast, long long stride_b, long long stride_z, long long stride_c,
                            float* out, void* stream);

/* grad_cl [n_cells, C]: BEV gradient in cells-major layout (dbev_transpose_batched of the
 * NCHW gradient). Writes d_depth[BN*D*fH*fW] and d_feat_cl[BN*fH*fW, C] completely. */
int dbev_lift_splat_backward(const float* grad_cl, const float* depth, const float* feat_cl,
                             const int* point_cell, long long n_pixels, int C, int D, int fhw,
                             float* d_depth, float* d_feat_cl, void* stream);

/* Sort-free variant (opt-in; fp32 summation order not fixed, everything else identical).
 * dbev_bev_point_cells: point_cell[n_points] = output cell of every frustum point, -1 when dropped —
 * the index math of voxel_pooling (view_transformer_mine.py:150-161) without the argsort.
 * dbev_lift_splat_atomic_forward: out_cl[n_cells, C] (channels-last BEV rows, zero-filled here) +=
 * depth[p] * feat_cl[pixel(p), :] with 16-byte vector reductions that resolve in L2; the feature row
 * of a pixel is read once for all its D depth bins. Backward = dbev_lift_splat_backward. */
int dbev_bev_point_cells(const float* geom, long long n_points, int batch, const float* off_host3,
                         const float* dx_host3, const float* nx_float_host3, const int* nx_int_host3,
                         int fast_axis, int* point_cell, void* stream);
/* Same with the sample-frames of a sample numbered LAST: batch = samples * frames, sample-frame b = s*frames + f,
 * cell = ((s*ny + y)*nx + x)*frames + f. A channels-last map out_cl[n_cells, C] indexed by these cells is
 * [s][y][x][f][C], i.e. the frames concatenated along the channels (torch.cat(bev_feat_list, dim=1),
 * bevdet.py:300-320) - the student BEV encoder reads it as [samples, ny, nx, frames*C] with no concat pass. */
int dbev_bev_point_cells_frames(const float* geom, long long n_points, int batch, int frames, const float* off_host3,
                                const float* dx_host3, const float* nx_float_host3, const int* nx_int_host3,
                                int fast_axis, int* point_cell, void* stream);
int dbev_lift_splat_atomic_forward(const float* depth, const float* feat_cl, const int* point_cell,
                                   long long n_pixels, int C, int D, int fhw, long long n_cells,
                                   float* out_cl, void* stream);

/* out[b][c][r] = in[b][r][c] for b < batch: NCHW <-> channels-last helper. */
int dbev_transpose_batched(const float* in, float* out, int batch, int rows, int cols,
                           void* stream);

/* ------------------------------------------------------------------------ *
 * LiDAR voxelization — replaces the pybind module mmdet3d.ops.voxel.voxel_layer
 * (mmdet3d/ops/voxel/src/voxelization.cpp:6-11, dispatch voxelization.h:58-140).
 * voxel_size_host3 = (vx, vy, vz), coors_range_host6 = (xmin, ymin, zmin, xmax,
 * ymax, zmax) are HOST float arrays; NDim is fixed to 3 like every caller in
 * the reference. Coordinates are stored (z, y, x) (voxelization_cpu.cpp:30).
 * ------------------------------------------------------------------------ */

/* grid = round((max - min) / voxel) in fp32 -> host int[3] (x, y, z)
 * (voxelization_cpu.cpp:120-123, voxelize.py:121-126). */
int dbev_voxel_grid_size(const float* voxel_size_host3, const float* coors_range_host6,
                         int* grid_xyz_host3);

/* dynamic_voxelize(points, coors, voxel_size, coors_range, NDim=3)
 * (voxelization.h:82-93; CPU :146-171, CUDA voxelization_cuda.cu:485-528).
 * points[n, nfeat] fp32 -> coors[n, 3] int32, (-1, -1, -1) for points outside
 * the range (the CPU build's convention). No device synchronisation (the
 * reference calls cudaDeviceSynchronize, voxelization_cuda.cu:524). */
int dbev_dynamic_voxelize(const float* points, int n, int nfeat, const float* voxel_size_host3,
                          const float* coors_range_host6, int* coors, void* stream);

/* hard_voxelize(points, voxels, coors, num_points_per_voxel, voxel_size,
 * coors_range, max_points, max_voxels, NDim=3, deterministic=true) -> voxel_num
 * (voxelization.h:58-80; CPU :45-144; CUDA voxelization_cuda.cu:231-402).
 * Caller pre-allocates voxels[max_voxels, max_points, nfeat], coors[max_voxels,3],
 * num_points_per_voxel[max_voxels] (voxelize.py:57-62). Voxel order = first
 * appearance in point order, in-voxel order = point order, both capped
 * exactly like the reference. Every slot of the first *voxel_num voxels is
 * written (points or zeros), so the buffers need not be zero-filled; entries
 * past *voxel_num are left untouched. *voxel_num is a DEVICE int (the Python
 * shim reads it back, the one sync the reference API implies). The
 * deterministic result is also a valid outcome of deterministic=false
 * (voxelization_cuda.cu:404-483), so both flags map here. */
size_t dbev_hard_voxelize_workspace_bytes(long long n);
int dbev_hard_voxelize(const float* points, int n, int nfeat, const float* voxel_size_host3,
                       const float* coors_range_host6, int max_points, int max_voxels,
                       float* voxels, int* coors, int* num_points_per_voxel, int* voxel_num,
                       void* workspace, size_t workspace_bytes, void* stream);

/* dynamic_point_to_voxel_forward(feats, coors, reduce_type) ->
 * [reduced_feats, out_coors, coors_map, reduce_count]
 * (voxelization.h:107-120, scatter_points_cuda.cu:183-239).
 * feats[n, nfeat] fp32; coors[n, ncol] int32 with ncol = 3 (z, y, x) or 4
 * (batch, z, y, x — one call for the whole batch instead of the reference's
 * per-sample Python loop, scatter_points.py:86-97); rows with any negative
 * component are dropped (coors_map = -1). dims_host[ncol] = exclusive upper
 * bound of every column. Output voxels come in lexicographic coordinate order
 * (what at::unique_dim returns). reduce_type: 0 sum, 1 mean, 2 max
 * (reduce_t, voxelization.h:4). Outputs are caller-allocated for the worst
 * case (n rows); *num_out is a DEVICE int. */
size_t dbev_dynamic_scatter_workspace_bytes(long long n);
int dbev_dynamic_scatter_forward(const float* feats, const int* coors, int n, int nfeat, int ncol,
                                 const int* dims_host, int reduce_type, float* reduced_feats,
                                 int* out_coors, int* coors_map, int* reduce_count, int* num_out,
                                 void* workspace, size_t workspace_bytes, void* stream);

/* dynamic_point_to_voxel_backward(grad_feats, grad_reduced_feats, feats,
 * reduced_feats, coors_idx, reduce_count, reduce_type)
 * (voxelization.h:122-138, scatter_points_cuda.cu:241-308). grad_feats[n, nfeat]
 * is fully written. reduce_from_ws: m * nfeat ints, needed for max only. */
int dbev_dynamic_scatter_backward(const float* grad_reduced, const float* feats,
                                  const float* reduced, const int* coors_map,
                                  const int* reduce_count, long long n, long long m, int nfeat,
                                  int reduce_type, float* grad_feats, int* reduce_from_ws,
                                  void* stream);

/* ------------------------------------------------------------------------ *
 * Teacher pillar encoder (eval mode) and pseudo-image scatter.
 * ------------------------------------------------------------------------ */

/* DynamicPillarFeatureNet.forward with one PFN layer, eval-mode BN
 * (mmdet3d/models/voxel_encoders/pillar_encoder.py:282-338): cluster mean,
 * f_cluster / f_center decorations, Linear(nfeat+5 -> nout, no bias) + BN + ReLU,
 * max over the pillar - fused, the decorated / per-point features never reach HBM.
 *   points[n, nfeat] of the whole batch back to back; either coors_in[n,4] = (b,z,y,x)
 *   (the reference's call shape) or batch_offsets[batch+1] (device int; voxelization is
 *   then done on the fly and optionally written to point_coors[n,4]).
 *   weight[nout, nfeat+5] = pfn_layers[0][0].weight; bn_scale/shift[nout] = folded BN.
 *   x_offset = vx/2 + pc_min_x, y_offset likewise (host doubles rounded to float, :88-89).
 * Outputs (sized for n rows): voxel_feats[M, nout], voxel_coors[M, 4] in lexicographic
 * (b,z,y,x) order = the order pfn_scatter returns; *num_voxels is a DEVICE int. */
size_t dbev_pillar_encode_workspace_bytes(long long n);
int dbev_pillar_encode(const float* points, const int* batch_offsets, const int* coors_in,
                       int batch, int n, int nfeat, const float* voxel_size_host3,
                       const float* coors_range_host6, float x_offset, float y_offset,
                       const float* weight, int nout, const float* bn_scale,
                       const float* bn_shift, float* voxel_feats, int* voxel_coors,
                       int* num_voxels, int* point_coors, void* workspace, size_t workspace_bytes,
                       void* stream);

/* Voxel encoder + middle encoder of the pillar teacher in one pass
 * (DynamicCenterPoint.extract_pts_feat, mmdet3d/models/detectors/dynamic_centerpoint.py:43-93:
 * pts_voxel_encoder -> pts_middle_encoder): dbev_pillar_encode whose pillar rows are stored
 * straight into the BEV canvas [batch, nout, ny, nx] (channels_last as for dbev_pillar_scatter),
 * so the [M, nout] table is never written or re-read. Needs a single z bin. Same workspace as
 * dbev_pillar_encode; *num_voxels (DEVICE int) = number of non-empty pillars. */
int dbev_pillar_canvas(const float* points, const int* batch_offsets, const int* coors_in,
                       int batch, int n, int nfeat, const float* voxel_size_host3,
                       const float* coors_range_host6, float x_offset, float y_offset,
                       const float* weight, int nout, const float* bn_scale,
                       const float* bn_shift, int channels_last, int zero_canvas, float* canvas,
                       int* num_voxels, void* workspace, size_t workspace_bytes, void* stream);

/* PointPillarsScatter.forward_batch (mmdet3d/models/middle_encoders/pillar_scatter.py:62-102):
 * canvas[b, :, y, x] = voxel_feats[m, :] for coors[m] = (b, z, y, x). The canvas is
 * [batch, C, ny, nx] (channels_last = 0) or the same tensor in NHWC memory order
 * (channels_last = 1: one contiguous row per pillar). m_dev (nullable): device count
 * clamp; zero_canvas != 0 enqueues the zero fill. One launch for the whole batch. */
int dbev_pillar_scatter(const float* voxel_feats, const int* coors, const int* m_dev, int m_max,
                        int C, int batch, int ny, int nx, int channels_last, int zero_canvas,
                        float* canvas, void* stream);

/* get_geometry (mmdet3d/models/necks/view_transformer_mine.py:111-139): frustum[pts_per_cam, 3]
 * (x_px, y_px, depth), per camera rots/intrins/post_rots [n_cams,3,3], trans/post_trans
 * [n_cams,3] -> geom[n_cams * pts_per_cam, 3] ego-frame xyz. mats_ws: n_cams * 18 floats. */
int dbev_lss_geometry(const float* frustum, int pts_per_cam, const float* rots,
                      const float* trans, const float* intrins, const float* post_rots,
                      const float* post_trans, int n_cams, float* mats_ws, float* geom,
                      void* stream);

/* ------------------------------------------------------------------------ *
 * '1x1conv' student adaptation layer on tcgen05 tensor cores (TF32 inputs, fp32 accumulate
 * in TMEM, operands fed by TMA straight from the NCHW tensors): replaces
 * nn.Conv2d(student_channel, teacher_channel, 1) of BEVDetDistill
 * (mmdet3d/models/detectors/bevdet_distill.py:216-351, applied at :1004).
 *   x_cl[batch, hw, c_in] CHANNELS-LAST (the memory of a torch.channels_last NCHW tensor; an
 *   NCHW-contiguous tensor goes through dbev_transpose_batched first), w[c_out, c_in],
 *   bias[c_out] (nullable) -> y[batch, c_out, hw] NCHW-contiguous.
 * c_in % 32 == 0, c_out in {128, 256, 384, 512}, hw % 4 == 0, 16-byte aligned pointers.
 * ------------------------------------------------------------------------ */
int dbev_adapt_conv1x1_forward(const float* x_cl, const float* w, const float* bias, int batch,
                               int c_in, int c_out, int hw, float* y, void* stream);

/* ------------------------------------------------------------------------ *
 * BEVDepth4D steps around the view transform.
 * ------------------------------------------------------------------------ */

/* shift_feature (mmdet3d/models/detectors/bevdet.py:267-321): out[n,c,y,x] = bilinear sample
 * (zeros padding, align_corners=True) of in[n,c] at tf[n] . (x, y, 1), tf[n] a 3x3 row-major affine
 * map in feature-pixel units (= inv(feat2bev) . l02l1 . feat2bev, :313); the reference's
 * [n,h,w,3,1] grid tensor, per-pixel matmul and normalisation pass are not materialised.
 * Backward = gradient w.r.t. `in` (bilinear scatter with float atomics; grad_in fully written). */
int dbev_shift_feature_forward(const float* in, const float* tf, int n, int C, int h, int w, float* out,
                               void* stream);
int dbev_shift_feature_backward(const float* grad_out, const float* tf, int n, int C, int h, int w,
                                float* grad_in, void* stream);

/* get_depth_loss (bevdet.py:397-417): loss_weight * mean over [BN, D, HW] of
 * (depth_gt != 0) * BCE(sigmoid(logits), one_hot(clip(floor((depth_gt - dmin) / dstep), 0, D))),
 * logits [BN, D, HW], depth_gt [BN, HW]; the one-hot tensor is never built. A ground-truth bin
 * equal to D (F.one_hot raises in the reference) counts as "no positive class". loss / grad_loss:
 * DEVICE float[1]; backward writes grad_logits completely. */
size_t dbev_depth_loss_workspace_bytes(void);
int dbev_depth_loss_forward(const float* logits, const float* depth_gt, int BN, int D, int HW, float dmin,
                            float dstep, float loss_weight, float* loss, void* workspace,
                            size_t workspace_bytes, void* stream);
int dbev_depth_loss_backward(const float* logits, const float* depth_gt, int BN, int D, int HW, float dmin,
                             float dstep, float loss_weight, const float* grad_loss, float* grad_logits,
                             void* stream);

/* ------------------------------------------------------------------------ *
 * Multi-scale deformable attention of the BEVFormer student. Replaces mmcv
 * 1.6.0 `_ext.ms_deform_attn_forward / ms_deform_attn_backward` as called by
 * MultiScaleDeformableAttnFunction_fp32
 * (mmdet3d/models/transformer_modules/multi_scale_deformable_attn_function.py:90-165); semantics =
 * mmcv's multi_scale_deformable_attn_pytorch (third party: parity unpinned). value [bs, num_keys,
 * heads, dim]; spatial_shapes [levels, 2] (h, w) and level_start [levels] int64 DEVICE tensors;
 * sampling_loc [bs, nq, heads, levels, points, 2] in [0,1]; attn_weight [bs, nq, heads, levels,
 * points]; out [bs, nq, heads*dim]. Backward writes all three gradients completely (grad_value by
 * float atomics, like mmcv).
 * ------------------------------------------------------------------------ */
int dbev_ms_deform_attn_forward(const float* value, const long long* spatial_shapes,
                                const long long* level_start, const float* sampling_loc,
                                const float* attn_weight, int bs, int num_keys, int heads, int dim,
                                int num_queries, int levels, int points, float* out, void* stream);
int dbev_ms_deform_attn_backward(const float* value, const long long* spatial_shapes,
                                 const long long* level_start, const float* sampling_loc,
                                 const float* attn_weight, const float* grad_out, int bs, int num_keys,
                                 int heads, int dim, int num_queries, int levels, int points,
                                 float* grad_value, float* grad_loc, float* grad_attn, void* stream);

/* ------------------------------------------------------------------------ *
 * CenterHead.get_targets on the device (SURVEY.md §8f rank 3;
 * mmdet3d/models/dense_heads/centerpoint_head.py:400-445, 447-611, core/utils/gaussian.py:6-87):
 * boxes[total, box_dim] = (gravity-centre-less) LiDAR boxes (x, y, z_bottom, dx, dy, dz, yaw, vx, vy)
 * of all samples back to back, labels[total] int32, offsets[batch+1] (device ints);
 * class_task_host / class_in_task_host[num_classes]: task of every class and its position inside
 * the task (HOST ints, <= 32 classes). Outputs (fully written): heatmap[batch, num_classes, H, W]
 * (the tasks' heat maps are its channel slices), anno_box[batch, num_tasks, max_objs, 10],
 * ind[batch, num_tasks, max_objs] int64, mask[batch, num_tasks, max_objs] uint8.
 * ------------------------------------------------------------------------ */
int dbev_center_targets(const float* boxes, int box_dim, const int* labels, const int* offsets, int batch,
                        const int* class_task_host, const int* class_in_task_host, int num_classes,
                        int num_tasks, int max_objs, int H, int W, float voxel_x, float voxel_y,
                        float out_size_factor, float pc_min_x, float pc_min_y, float gaussian_overlap,
                        int min_radius, int norm_bbox, float* heatmap, float* anno_box,
                        long long* ind, unsigned char* mask, void* stream);

/* ------------------------------------------------------------------------ *
 * Dense conv + eval BatchNorm + ReLU of the FROZEN LiDAR teacher's BEV backbone / neck on
 * tcgen05 tensor cores (TF32 inputs, fp32 accumulate): replaces the cuDNN conv + BatchNorm2d +
 * ReLU kernel triples of SECOND.forward (mmdet3d/models/backbones/second.py:80-93) and
 * SECONDFPN.forward (mmdet3d/models/necks/second_fpn.py:77-93) by one implicit-GEMM kernel per
 * layer. NHWC fp32 activations ([n, h, w, c_in], the memory of a torch.channels_last tensor);
 * w_packed[c_out][(ky*kw + kx)*c_in + ci]; scale / shift [c_out] = folded BN (nullable); the
 * filter taps are TMA boxes of the input (padding = TMA zero fill, stride 2 = TMA element
 * strides): no im2col buffer. c_in % 32 == 0, c_out in {64, 128, 256}, kernel 1..3, stride 1 or
 * 2, padding 0 or 1. Output pixel (oy, ox) channel c goes to
 *   out[((n*out_h + oy*out_mul + out_add_y)*out_w + ox*out_mul + out_add_x)*out_ld + out_c_off + c]
 * so a layer can write a channel slice of the FPN concat (out_c_off, out_ld) and a k2/s2
 * transposed conv is four 1x1 launches with out_mul = 2 and (out_add_y, out_add_x) = (dy, dx).
 * out_nchw != 0 stores the same element at out[((n*out_ld + out_c_off + c)*out_h + y)*out_w + x]
 * instead (NCHW; what the distillation-loss kernels read), saving a layout-conversion pass.
 * ------------------------------------------------------------------------ */
int dbev_conv2d_tc_forward(const float* x_nhwc, int n, int h, int w, int c_in, const float* w_packed,
                           int c_out, int kh, int kw, int stride, int pad, const float* scale,
                           const float* shift, int relu, float* out, int out_h, int out_w,
                           int out_ld, int out_c_off, int out_mul, int out_add_y, int out_add_x,
                           int out_nchw, void* stream);

/* Same, with the C_out columns split into out_groups blocks of C_out / out_groups real channels: block g is
 * written at lattice x + g (channel = column % (C_out / out_groups); scale / shift are per real channel).
 * With out_mul = 2, out_groups = 2 one launch computes both x taps of a k2/s2 ConvTranspose2d row
 * (second_fpn.py:41-46: the deconv of the stride-4 branch) from ONE read of the input tile. */
int dbev_conv2d_tc_forward_grouped(const float* x_nhwc, int n, int h, int w, int c_in, const float* w_packed,
                                   int c_out, int kh, int kw, int stride, int pad, const float* scale,
                                   const float* shift, int relu, float* out, int out_h, int out_w,
                                   int out_ld, int out_c_off, int out_mul, int out_add_y, int out_add_x,
                                   int out_nchw, int out_groups, void* stream);

/* Extended form used by the TRAINING path (input gradients are convolutions of dy with re-packed weights):
 * x_ld = channel stride of the input rows (>= c_in; the input may be a channel slice of a wider NHWC
 * tensor); force_ho / force_wo > 0 override the output size (taps outside the input read zeros: the four
 * output-parity classes of a stride-2 input gradient are stride-1 convolutions padded on one side);
 * accumulate != 0 adds to `out` (TMA reduce-add store) - the residual branches of BasicBlock
 * (mmdet3d/models/bricks/res_block.py:70-99) sum their input gradients this way; n_col_blocks > 1: the layer has
 * n_col_blocks * c_out output channels (c_out in {64, 128, 256} is the tile width) and a work item is
 * (pixel tile, column block), so small BEV maps (16 x 16 .. 64 x 64) still fill the 148 SMs in one launch. */
int dbev_conv2d_tc_forward_ex(const float* x_nhwc, int n, int h, int w, int c_in, int x_ld, const float* w_packed,
                              int c_out, int n_col_blocks, int kh, int kw, int stride, int pad, const float* scale,
                              const float* shift, int relu, float* out, int out_h, int out_w,
                              int out_ld, int out_c_off, int out_mul, int out_add_y, int out_add_x,
                              int out_nchw, int out_groups, int force_ho, int force_wo, int accumulate,
                              void* stream);

/* ------------------------------------------------------------------------ *
 * Training of the student's BEV encoder: ResNetForBEVDet.forward
 * (mmdet3d/models/backbones/resnet.py:51-62), BasicBlock (bricks/res_block.py:70-99), FPN_LSS.forward
 * (necks/lss_fpn.py:62-72) and the adaptation convs (detectors/bevdet_distill.py:261-345). The reference
 * runs these as nn.Conv2d / nn.BatchNorm2d (training mode) / nn.ReLU / nn.Upsample through cuDNN + ATen
 * with autograd; these entry points are what replaces aten::convolution_backward, native_batch_norm
 * (+ backward), upsample_bilinear2d (+ backward). NHWC fp32, row strides (`*_ld`, in floats) so that an
 * operand can be a channel slice of a wider tensor (torch.cat is never materialised separately).
 * ------------------------------------------------------------------------ */

/* dW[c_out][c_in][kh][kw] (torch layout) (+)= sum_pixels dy (x) x on tcgen05 (TF32, MN-major operands,
 * split-K over pixel strips, partial sums combined in fixed order). c_in, c_out % 128 == 0; 3x3 / pad 1 /
 * stride 1|2 or 1x1 / stride 1. workspace: dbev_conv_wgrad_tc_workspace_bytes, 256-byte aligned. */
size_t dbev_conv_wgrad_tc_workspace_bytes(int n, int ho, int wo, int c_in, int c_out, int kh, int kw, int stride);
int dbev_conv_wgrad_tc(const float* x_nhwc, int n, int h, int w, int c_in, int x_ld, const float* dy_nhwc,
                       int ho, int wo, int c_out, int dy_ld, int kh, int kw, int stride, int pad, float* dw,
                       int accumulate, void* workspace, size_t workspace_bytes, void* stream);

/* Input gradient of a 3x3 / stride 2 / pad 1 convolution (ResNetForBEVDet's stride-2 conv1 / downsample,
 * backbones/resnet.py:26-34) in ONE launch: dx[n, 2*ho, 2*wo, dx_ld] channels [dx_c_off, dx_c_off + col_width *
 * n_col_blocks) (+)= conv_transpose(dy, W). dy NHWC [n, ho, wo, c_out] (channel stride dy_ld); w_mode2 =
 * dbev_pack_conv_weights(mode 2) of the whole layer with c_in_total input channels; col_width in {64, 128, 256}.
 * The four parity classes of the input pixels are work items of the same launch (dbev_conv2d_tc_forward_ex with
 * force_ho / out_mul = 2 runs them as four launches). */
int dbev_conv2d_tc_dgrad_s2(const float* dy_nhwc, int n, int ho, int wo, int c_out, int dy_ld, const float* w_mode2,
                            int c_in_total, int col_width, int n_col_blocks, float* dx, int dx_ld, int dx_c_off, int accumulate,
                            void* stream);
/* torch weight [c_out][c_in][kh][kw] -> K-major matrices for the conv kernels' TMA loads. mode 0: forward
 * [c_out][(ky*kw+kx)*c_in + ci]; mode 1: stride-1 input gradient [c_in][flipped tap * c_out + co];
 * mode 2 (3x3): the four parity-class matrices of a stride-2 input gradient at float offsets
 * {0, 1, 3, 5} * c_in*c_out, class (a, b): [c_in][(ty*(1+b) + tx)*c_out + co], ky = a+1-2ty, kx = b+1-2tx. */
int dbev_pack_conv_weights(const float* w, int c_out, int c_in, int kh, int kw, int mode, float* out, void* stream);
/* Forward matrix (mode 0 layout) and input-gradient matrix (dgrad_mode 1 or 2 layout) of one layer in ONE pass over
 * the weights (shared-memory tiles: every global access is a 128-byte run); either output may be NULL. */
int dbev_pack_conv_weights_train(const float* w, int c_out, int c_in, int kh, int kw, int dgrad_mode, float* out_fwd,
                                 float* out_dgrad, void* stream);
/* The same for every layer of a network in ONE launch (the per-step re-packing after the optimizer update).
 * jobs_dev: n_jobs records of 8 int64 in DEVICE memory {w, out_fwd (0 = skip), out_dgrad (0 = skip), c_out, c_in,
 * kh * 256 + kw, dgrad_mode, first_tile}; a job has ceil(c_out/32) * ceil(c_in/32) tiles, numbered consecutively over
 * the jobs; total_tiles = their sum. Filters up to 3x3. */
int dbev_pack_conv_weights_batch(const long long* jobs_dev, int n_jobs, int total_tiles, void* stream);

/* Workspace of the per-channel reductions below (per-block partial sums, combined in block order). */
size_t dbev_channel_stats_workspace_bytes(long long rows, int C);
/* nn.BatchNorm2d training forward, statistics part: out4c[4][C] = (a, b, mean, invstd) with
 * a = gamma*invstd, b = beta - mean*a over the `rows` pixels; running_mean / running_var (nullable pair)
 * are updated like torch (momentum, unbiased variance). */
int dbev_bn_batch_stats(const float* y, int y_ld, long long rows, int C, const float* gamma, const float* beta,
                        float eps, float momentum, float* running_mean, float* running_var, float* out4c,
                        void* workspace, size_t workspace_bytes, void* stream);
/* out[C] (+)= column sums of y[rows, C] (conv bias gradient). */
int dbev_channel_sums(const float* y, int y_ld, long long rows, int C, float* out, int accumulate,
                      void* workspace, size_t workspace_bytes, void* stream);
/* out = relu?(a*y + b (+ residual)); ab[2][C] nullable (identity affine). */
int dbev_bn_act_forward(const float* y, int y_ld, const float* ab, const float* residual, int res_ld,
                        long long rows, int C, int relu, float* out, int out_ld, unsigned char* relu_mask, void* stream);
/* relu_mask (nullable, [rows][C/4] bytes, written when relu != 0): bit k of byte (row, quad) = pre-activation of
 * channel 4*quad + k > 0. The backward entry points below take it instead of z - 1/16 of the bytes of re-reading z.
 * Backward of z = relu?(BN(y) (+ identity)): g = dz * (z > 0) (z and relu_mask both NULL: no ReLU);
 * bwd4c[4][C] = (dgamma, dbeta, mean g, mean g*yhat); dy = a*(g - mean g - yhat*mean(g*yhat));
 * g_out (nullable) receives g (the identity-branch gradient), added when g_accumulate. */
int dbev_bn_backward(const float* dz, int dz_ld, const float* z, int z_ld, const float* y, int y_ld,
                     const float* fwd4c, long long rows, int C, float* bwd4c, float* dy, int dy_ld,
                     float* g_out, int g_ld, int g_accumulate, const unsigned char* relu_mask, void* workspace,
                     size_t workspace_bytes, void* stream);
/* g_out (+)= dz * (z > 0) without a BatchNorm (z or relu_mask). */
int dbev_relu_mask_backward(const float* dz, int dz_ld, const float* z, int z_ld, long long rows, int C,
                            float* g_out, int g_ld, int accumulate, const unsigned char* relu_mask, void* stream);
/* nn.Upsample(mode='bilinear', align_corners=True) on NHWC, ATen's index arithmetic; the backward is a
 * gather (no atomics, deterministic). */
int dbev_upsample_bilinear_forward(const float* in, int in_ld, int n, int h, int w, int C, int H, int W,
                                   float* out, int out_ld, void* stream);
int dbev_upsample_bilinear_backward(const float* dout, int dout_ld, int n, int h, int w, int C, int H, int W,
                                    float* din, int din_ld, int accumulate, void* stream);

/* ------------------------------------------------------------------------ *
 * Per-camera re-batching of BEV queries in BEVFormer's SpatialCrossAttention.forward
 * (mmdet3d/models/transformer_modules/spatial_cross_attention.py:128-167; the reference uses Python double loops over
 * batch x camera and nonzero() index lists). idx[cams, max_len] = k-th query seen by the camera or -1;
 * pos[cams, nq] = rank of the query in the camera's list or -1; scale[bs, nq] nullable.
 *   gather: out[bs, cams, max_len, C] = scale * in[b, idx, :] (0 for padding); `in` rows addressed by
 *           cam*in_cam_stride + b*in_batch_stride + q*in_query_stride (floats): [bs, nq, C] (camera stride 0) and
 *           [cams, bs, nq, C] sources both work
 *   reduce: out[bs, nq, C] = scale * sum over cameras of in[b, cam, pos, :] (fixed camera order: deterministic)
 * forward re-batch = gather (backward: reduce); forward slot accumulation / count = reduce with scale = 1 / count
 * (backward: gather with the same scale).
 * ------------------------------------------------------------------------ */
int dbev_sca_gather_rows(const float* in, const int* idx, const float* scale, int bs, int cams, int max_len, int nq, int C,
                         long long in_cam_stride, long long in_batch_stride, long long in_query_stride, float* out, void* stream);
int dbev_sca_reduce_rows(const float* in, const int* pos, const float* scale, int bs, int cams, int max_len, int nq, int C,
                         float* out, void* stream);

/* ------------------------------------------------------------------------ *
 * PillarFeatureNet.forward in eval mode (mmdet3d/models/voxel_encoders/pillar_encoder.py:95-162 + PFNLayer
 * utils.py:107-181; one PFN layer, cluster + voxel-centre decorations, max pooling: the shipped pillar teacher,
 * configs/_base_/models/centerpoint_02pillar_second_secfpn_nus.py:6-13). voxels[m_max, max_points, F] / num_points[m_max]
 * / coors[m_max, 4] (b, z, y, x) as hard_voxelize + the batch padding produce them; m_dev (nullable) = device voxel count;
 * weight[nout, F + 5] (nn.Linear), bn_scale / bn_shift = folded BatchNorm1d; legacy != 0 reproduces the in-place
 * centre offset of legacy=True (:132-139). out[m_max, nout].
 * ------------------------------------------------------------------------ */
int dbev_hard_pillar_encode(const float* voxels, const int* num_points, const int* coors, const int* m_dev, int m_max,
                            int max_points, int nfeat, const float* voxel_size_xy_host2, float x_offset, float y_offset,
                            const float* weight, int nout, const float* bn_scale, const float* bn_shift, int legacy,
                            float* out, void* stream);

/* ------------------------------------------------------------------------ *
 * Cross-modal feature distillation loss (FGD-style), BEVDetDistill
 * (mmdet3d/models/detectors/bevdet_distill.py). The reference has no native
 * interface for this path: it is ~20 torch kernels plus numpy/numba on the
 * host. These entry points are what a maintainer binds instead of the bodies
 * of foreground_scale_mask (:755-843), add_fp_as_fg (:846-970) and the
 * attention / mask / loss part of fgd_distill_loss (:1084-1293).
 * ------------------------------------------------------------------------ */

/* foreground_scale_mask (:755-843; BEVFormer variant bevformer_distill.py:391-482).
 * boxes[total, box_dim] fp32 rows (x, y, z, x_size, y_size, z_size, yaw, ...) of all
 * samples back to back, box_offsets[batch + 1] (device int). The sample point of
 * BEV cell (i, j) is (i*voxel_x*osf + pc_min_x, j*voxel_y*osf + pc_min_y) in fp32
 * (+ half a cell when cell_center != 0); osf = grid_size // W (BEVDet) or
 * grid_size / W (BEVFormer). Outputs fg[batch,H,W] in {0,1}, fg_scale[batch,H,W] =
 * sqrt(cell_area / (w*l)) of the first containing box, fg_count[batch] (int). The
 * background scale 1 / (H*W - fg_count) is applied inside dbev_fgd_loss_forward. */
int dbev_fgd_foreground_mask(const float* boxes, int box_dim, const int* box_offsets,
                             int max_boxes_per_sample, int batch, int H, int W, float voxel_x,
                             float voxel_y, float out_size_factor, float pc_min_x, float pc_min_y,
                             int cell_center, int transpose_mask, float* fg, float* fg_scale,
                             int* fg_count, void* stream);

/* max over the class axis of heatmaps[batch, K, H, W]; apply_clip_sigmoid != 0 applies
 * clamp(sigmoid(x), 1e-4, 1 - 1e-4) first (models/utils/clip_sigmoid.py:17; :855-858). */
int dbev_heatmap_class_max(const float* heatmaps, int batch, int K, int H, int W,
                           int apply_clip_sigmoid, float* out, void* stream);

/* add_fp_as_fg (:846-925) on class-max maps: gt_max[batch,Sg,Sg], teacher_max[batch,St,St],
 * student_max[batch,Ss,Ss] (may be NULL for mode 0), fg[batch,R,R]. mode: 0 teacher,
 * 1 student, 2 teacher_selected_student, 3 teacher+teacher_selected_student (:893-903).
 * Resampling between resolutions = max-pool / repeat as in the reference. Outputs
 * fp[batch,R,R] in {0,1} (already cleared where fg != 0) and fp_count[batch] (int);
 * fp_scale_mode 'average' (1 / fp_count) is applied inside dbev_fgd_loss_forward. */
int dbev_fgd_fp_mask(const float* gt_max, int Sg, const float* teacher_max, int St,
                     const float* student_max, int Ss, const float* fg, int R, int batch,
                     int mode, float thres, float gt_thres, float* fp, int* fp_count,
                     void* stream);

/* fp_scale_mode 'dfs' of add_fp_as_fg (bevdet_distill.py:926-966): fp[batch,H,W] in {0,1} -> scale[batch,H,W],
 * every 4-connected FP component gets 1 / n where n = the number of FIFO pops of the reference's flood fill
 * (cells re-queued before they are visited count again: a 2x3 block gives 1/9). workspace: device memory,
 * 8-byte aligned, dbev_fgd_fp_dfs_workspace_bytes(). The loss entry points take the map as
 * fp = scale * fp_count (their 'average' factor 1 / fp_count cancels). */
size_t dbev_fgd_fp_dfs_workspace_bytes(int batch, int H, int W);
int dbev_fgd_fp_dfs_scale(const float* fp, int batch, int H, int W, float* scale, void* workspace,
                          size_t workspace_bytes, void* stream);

typedef struct dbev_fgd_config {
  int B, C, H, W;
  float spatial_t;             /* distill_params['spatial_t'] */
  float channel_t;             /* distill_params['channel_t'] */
  float spatial_student_ratio; /* distill_params['spatial_student_ratio'] */
  float w_fg, w_bg, w_fp, w_channel, w_spatial; /* loss weights (fp_weight for w_fp) */
  int spatial_att;     /* 0 'teacher', 1 'teacher_student' */
  int spatial_mask;    /* distill_params['spatial_mask'] */
  int channel_mask;    /* distill_params['channel_mask'] */
  int scale_mask;      /* 0 none, 1 'combine_gt', 2 'separate_gt', 3 'bg_only' */
  int use_fp;          /* fp_as_foreground != 'none' and epoch >= fp_epoch */
} dbev_fgd_config;

/* bytes of the opaque state that links forward and backward */
size_t dbev_fgd_state_bytes(const dbev_fgd_config* cfg);

/* Losses of fgd_distill_loss (:1252-1293) for already-adapted student[B,C,H,W] and
 * teacher[B,C,H,W] (NCHW fp32, H*W % 4 == 0): losses[5] = kd_fg_feat_loss,
 * kd_bg_feat_loss, kd_fp_bg_feat_loss, kd_channel_loss, kd_spatial_loss (terms the
 * configuration disables are 0). conv_w[9] / conv_b[1]: spatial_wise_adaptations
 * Conv2d(1,1,3,padding=1) (:348-351), device pointers. 3 tensor reads in total. */
int dbev_fgd_loss_forward(const dbev_fgd_config* cfg, const float* student, const float* teacher,
                          const float* fg, const float* fg_scale, const int* fg_count,
                          const float* fp, const int* fp_count, const float* conv_w,
                          const float* conv_b, void* state, size_t state_bytes, float* losses,
                          void* stream);

/* Gradient of sum_k grad_losses[k] * losses[k] w.r.t. student (attention masks and the
 * teacher are detached as in the reference, :1100,1104,1108), conv_w[9] and conv_b[1].
 * grad_channel_sum (optional, [C]) receives sum over (b, h, w) of grad_student: the bias gradient
 * of the 1x1 channel_wise_adaptations conv that produced `student` (:1004), for free. */
int dbev_fgd_loss_backward(const dbev_fgd_config* cfg, const float* student, const float* teacher,
                           const float* conv_w, const float* conv_b, void* state,
                           size_t state_bytes, const float* grad_losses, float* grad_student,
                           float* grad_conv_w, float* grad_conv_b, float* grad_channel_sum,
                           void* stream);

/* The '1x1conv' channel_wise_adaptations layer (nn.Conv2d(C_in, C, 1), bevdet_distill.py:216-351, applied at :1004)
 * fused with the loss: adapted = x W^T + bias is computed tile by tile on the tensor cores (TF32) and consumed in the
 * GEMM epilogue - the adapted map [B,C,H,W] is never written or re-read. x_cl[B, H*W, C_in] is the student feature in
 * channels-last memory (torch.channels_last), adapt_w[C, C_in], adapt_b[C] or NULL, teacher[B,C,H,W] NCHW; the other
 * arguments, losses[5] and the state are those of dbev_fgd_loss_forward. dbev_fgd_adapt_supported: C_in % 32 == 0 and
 * C % 32 == 0 up to 256 or C % 64 == 0 up to 512 (else the entry points return DBEV_ERR_INVALID_ARGUMENT). */
int dbev_fgd_adapt_supported(const dbev_fgd_config* cfg, int c_in);
int dbev_fgd_adapt_loss_forward(const dbev_fgd_config* cfg, const float* x_cl, int c_in, const float* adapt_w,
                                const float* adapt_b, const float* teacher, const float* fg, const float* fg_scale,
                                const int* fg_count, const float* fp, const int* fp_count, const float* conv_w,
                                const float* conv_b, void* state, size_t state_bytes, float* losses, void* stream);
/* Backward: grad_adapted_cl[B, H*W, C] (channels-last rows) = d sum_k grad_losses[k] losses[k] / d adapted, recomputing
 * the adapted tiles; grad_channel_sum[C] (optional) = its sum over (b, h, w) = the gradient of adapt_b; grad_conv_w[9],
 * grad_conv_b[1] as in dbev_fgd_loss_backward. The gradients of x and adapt_w are the two GEMMs
 * dbev_conv2d_tc_forward_ex (transposed filter) and dbev_conv_wgrad_tc over grad_adapted_cl. */
int dbev_fgd_adapt_loss_backward(const dbev_fgd_config* cfg, const float* x_cl, int c_in, const float* adapt_w,
                                 const float* adapt_b, const float* teacher, const float* conv_w, const float* conv_b,
                                 void* state, size_t state_bytes, const float* grad_losses, float* grad_adapted_cl,
                                 float* grad_conv_w, float* grad_conv_b, float* grad_channel_sum, void* stream);

/* ------------------------------------------------------------------------ *
 * Sparse 3D convolution of the sparse LiDAR teachers (LidarFormer / MVPFormer middle
 * encoder). Replaces the pybind module mmdet3d.ops.spconv.sparse_conv_ext
 * (mmdet3d/ops/spconv/src/all.cc:21-51): get_indice_pairs_3d
 * (include/spconv/spconv_ops.h:28-141, geometry.h:24-84,141-199,259-311) and
 * indice_conv_fp32 (spconv_ops.h:261-361), plus SparseConvTensor.dense
 * (mmdet3d/ops/spconv/structure.py:53-64).
 *
 * The rulebook is an OUTPUT-major neighbour table nbr[kvol][n_out] (int32): row of the
 * input voxel that feeds output o through kernel offset k, or -1. Offsets are numbered
 * like the reference (geometry.h:64-66): k = (kz*Ky + ky)*Kx + kx with
 * in = out*stride - padding + k*dilation per axis; weight is [Kz,Ky,Kx,Cin,Cout]
 * (conv.py:109-110). Coordinates are int32 rows (batch, z, y, x), 16-byte aligned.
 * geom_host[19] = ksize[3], stride[3], padding[3], dilation[3], in_shape[3],
 * out_shape[3] (all z,y,x) and batch — HOST ints. For a submanifold conv pass
 * stride 1, padding = ksize/2 (spconv_ops.h:75-78) and out_shape = in_shape.
 * ------------------------------------------------------------------------ */
long long dbev_spconv_max_out(long long n_in, const int* geom_host19);
size_t dbev_spconv_workspace_bytes(long long n_in, long long max_out);

/* Neighbour table for known output coordinates (SubMConv3d: out_coors = in_coors;
 * replaces getIndicePairsSubM, geometry.h:259-311). */
int dbev_spconv_table(const int* in_coors, int n_in, const int* out_coors, int n_out,
                      const int* geom_host19, int* nbr, void* workspace, size_t workspace_bytes,
                      void* stream);

/* SparseConv3d step 1: the distinct output cells (getValidOutPos, geometry.h:24-84) as
 * unsorted linear keys out_keys[max_out] and their count *n_out (DEVICE int). The caller
 * reads *n_out back (the reference synchronises in the same place, spconv_ops.h:130-137). */
int dbev_spconv_out_candidates(const int* in_coors, int n_in, const int* geom_host19,
                               uint32_t* out_keys, long long max_out, int* n_out,
                               void* workspace, size_t workspace_bytes, void* stream);

/* SparseConv3d step 2: sorts the keys (lexicographic (b,z,y,x) output order = what
 * torch::_unique gives the reference's CUDA path, spconv_ops.h:131), writes
 * out_coors[n_out,4] and nbr[kvol][n_out]. out_keys is clobbered. */
int dbev_spconv_out_table(const int* in_coors, int n_in, const int* geom_host19,
                          uint32_t* out_keys, int n_out, int* out_coors, int* nbr,
                          void* workspace, size_t workspace_bytes, void* stream);

/* Neighbour table -> the reference's rulebook tensors indice_pairs[kvol,2,pair_stride]
 * (-1 filled) and indice_num[kvol] (what get_indice_pairs returns, spconv_ops.h:56-59),
 * pairs of every offset in ascending output order; and back (inverse != 0 swaps the two
 * rows, as indice_conv's `inverse` flag does, spconv_ops.h:323-352). */
int dbev_spconv_pairs_from_table(const int* nbr, int kvol, int n_out, int pair_stride,
                                 int* indice_pairs, int* indice_num, void* stream);
int dbev_spconv_table_from_pairs(const int* indice_pairs, const int* indice_num, int kvol,
                                 int pair_stride, int inverse, int n_out, int* nbr, void* stream);

/* out[o,:] = act((sum_k in_feats[nbr[k][o],:] . weight[k]) * scale + shift + residual[o,:])
 * One kernel instead of kvol x (gather, mm, scatter-add) (spconv_ops.h:308-356); scale /
 * shift [c_out] fold the eval-mode BatchNorm1d that follows every sparse conv of
 * SparseEncoder (sparse_encoder.py:97-128) and a conv bias (conv.py:223-224); residual
 * [n_out,c_out] is SparseBasicBlock's identity (sparse_block.py:116-117). All three are
 * nullable; relu != 0 applies max(., 0). c_out in {16, 32, 64, 128}. */
int dbev_spconv_forward(const float* in_feats, int c_in, const float* weight, int c_out,
                        const int* nbr, int kvol, int n_out, const float* scale,
                        const float* shift, const float* residual, int relu, float* out,
                        void* stream);

/* Tensor-core variant of dbev_spconv_forward for the wide layers: tcgen05 implicit GEMM per tile
 * of 128 outputs (gathered rows written into the swizzled smem operand layout, weights by TMA,
 * accumulators in TMEM), 3xTF32 split so that results match the reference's fp32 cuBLAS mm
 * (spconv_ops.h:333) to ~1e-6. dbev_spconv_tc_supported: c_in, c_out in {32, 64, 128} and
 * kvol in {27, 3}. Weights are pre-packed once per layer (frozen teacher) by
 * dbev_spconv_pack_weights: weight [kvol, c_in, c_out] -> wt_hi / wt_lo [kvol, c_out, c_in]. */
int dbev_spconv_tc_supported(int c_in, int c_out, int kvol);
int dbev_spconv_pack_weights(const float* weight, int kvol, int c_in, int c_out, float* wt_hi,
                             float* wt_lo, void* stream);
int dbev_spconv_forward_tc(const float* in_feats, int c_in, const float* wt_hi, const float* wt_lo,
                           int c_out, const int* nbr, int kvol, int n_out, const float* scale,
                           const float* shift, const float* residual, int relu, float* out,
                           void* stream);

/* SparseConvTensor.dense() + view(N, C*D, H, W) (structure.py:53-64,
 * sparse_encoder.py:122-126): dense[b, c*Z + z, y, x] = feats[m, c]; the whole tensor is
 * written (zero fill included). */
int dbev_spconv_dense(const float* feats, const int* coors, int m, int C, int batch, int Z,
                      int Y, int X, float* dense, void* stream);

/* ------------------------------------------------------------------------ *
 * Voxel encoders of the sparse teachers.
 * ------------------------------------------------------------------------ */

/* HardSimpleVFE.forward (mmdet3d/models/voxel_encoders/voxel_encoder.py:29-45):
 * out[m, :nf] = voxels[m, :, :nf].sum(1) / num_points[m]. */
int dbev_hard_simple_vfe(const float* voxels, const int* num_points, long long m, int max_points,
                         int nfeat, int num_features, float* out, void* stream);

/* voxelization() of DynamicVoxelEncoder (voxel_encoders/dynamic_voxel_encoder.py:8-17): keep
 * lo <= p <= hi (both ends inclusive), coords = trunc((p - lo) / voxel) in fp32, stored
 * (batch, z, y, x); dropped points get -1 in every column. points of the whole batch back to
 * back, batch_offsets[batch+1] device ints. The per-voxel mean (coords.unique(dim=0) +
 * scatter_mean) is dbev_dynamic_scatter_forward(reduce_type = mean, ncol = 4).
 * check_flag != 0 (voxelization_virtual, :19-68) also drops points whose flag column
 * points[:, -2] is not 1 / 0 / -1. */
int dbev_dynvoxel_coords(const float* points, int n, int nfeat, const int* batch_offsets,
                         int batch, const float* pc_range_host6, const float* voxel_size_host3,
                         int check_flag, int* coors, void* stream);

/* voxelization_virtual (:27-50): 17-column MVP points -> the 24-channel rows whose per-voxel
 * mean the reference takes; dbev_dynvoxel_virtual_fix (:56-66) re-normalises voxels mixing
 * real and painted/virtual points and drops the indicator channel: mean24[m,24] -> out23[m,23]
 * (m_dev: nullable DEVICE count clamp). */
int dbev_dynvoxel_virtual_rows(const float* points, int n, int nfeat, float* rows24, void* stream);
int dbev_dynvoxel_virtual_fix(const float* mean24, const int* m_dev, int m_max, float* out23,
                              void* stream);

/* ------------------------------------------------------------------------ *
 * Affinity distillation loss — BEVDetDistill.affinity_distill_loss, list branch
 * (mmdet3d/models/detectors/bevdet_distill.py:735-748) and the masked-cell gather that
 * feeds it (:1294-1321):
 *   loss = sum_b weight * mean_{K_b x K_b} criterion(T_b T_b^T - S_b S_b^T)
 * criterion kind: 0 SmoothL1(beta), 1 L1, 2 MSE (mmdet losses, reduction 'mean'). The
 * K x K gram matrices are never materialised.
 * ------------------------------------------------------------------------ */
size_t dbev_affinity_select_workspace_bytes(int batch, int hw);
/* Cells with mask_a != 0 (or mask_b != 0, nullable) of every sample in ascending cell
 * order: row_cell[batch*hw] (compact), row_offsets[batch+1] — DEVICE ints. */
int dbev_affinity_select(const float* mask_a, const float* mask_b, int batch, int hw,
                         int* row_cell, int* row_offsets, void* workspace, size_t workspace_bytes,
                         void* stream);
/* rows[r, :] = feat[b, :, cell(r)] for feat [batch, C, hw] NCHW (the feat[c][mask] gather). */
int dbev_affinity_gather_rows(const float* feat, const int* row_cell, const int* row_offsets,
                              int batch, int C, int hw, int k_total, float* rows, void* stream);
/* number of floats of `partial` for dbev_affinity_forward */
size_t dbev_affinity_partial_floats(const int* row_offsets_host, int batch);
/* t_rows / s_rows [k_total, C]; row_offsets_host[batch+1] HOST ints; loss: DEVICE float[1]. */
int dbev_affinity_forward(const float* t_rows, const float* s_rows, const int* row_offsets_host,
                          int batch, int C, int kind, float beta, float weight, float* partial,
                          float* loss, void* stream);
/* d_s_rows[k_total, C] = d(loss * *grad_loss)/d(s_rows); grad_loss: DEVICE float[1]. */
int dbev_affinity_backward(const float* t_rows, const float* s_rows, const int* row_offsets_host,
                           int batch, int C, int kind, float beta, float weight,
                           const float* grad_loss, float* d_s_rows, void* stream);
/* grad[batch, C, hw] = 0 except grad[b, :, cell(r)] = d_rows[r, :]. */
int dbev_affinity_scatter_rows(const float* d_rows, const int* row_cell, const int* row_offsets,
                               int batch, int C, int hw, int k_total, float* grad, void* stream);

/* ------------------------------------------------------------------------ *
 * Primitives exposed for testing (stable LSD radix sort, exclusive scan).
 * They stand in for argsort / at::unique_dim / cumsum on the reference path.
 * ------------------------------------------------------------------------ */
size_t dbev_sort_workspace_bytes(long long n);
/* Sorts (keys_in[i], i) by key bits [0,num_bits): keys_out / order_out. */
int dbev_sort_keys_iota(const uint32_t* keys_in, int n, int num_bits, uint32_t* keys_out,
                        uint32_t* order_out, void* workspace, size_t workspace_bytes, void* stream);
size_t dbev_scan_workspace_bytes(long long n);
int dbev_exclusive_scan_i32(const int* in, int* out, int n, int* total_out, void* workspace,
                            size_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif

#endif /* DISTILL_BEV_B200_H_ */
