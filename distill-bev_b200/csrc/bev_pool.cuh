// Internal C++ interface of the bev_pool kernels; the public boundary is the
// C-ABI in include/distill_bev_b200.h (capi.cu forwards to these).
#pragma once

#include "common.cuh"

namespace dbev {

size_t bev_plan_ws_bytes(long long n_points, long long n_cells);
long long bev_plan_max_items(long long n_points, long long n_cells, int nfast, int rows_per_item);

int bev_plan_from_geom(const float* geom, long long n_points, int batch, const float off[3],
                       const float dx[3], const float nx_f[3], const int nx_i[3], int fast_axis,
                       int rows_per_item, uint32_t* order, int* cell_start, int* cell_end,
                       int4* items, long long max_items, int* n_items, int* point_cell, void* ws,
                       size_t ws_bytes, cudaStream_t stream);

int bev_plan_from_coords(const void* coords, int coords_i64, long long n_points, int batch, int n0,
                         int n1, int nz, int fast_axis, int rows_per_item, uint32_t* order,
                         int* cell_start, int* cell_end, int4* items, long long max_items,
                         int* n_items, void* ws, size_t ws_bytes, cudaStream_t stream);

int bev_pool_gather_forward(const float* x, int C, const uint32_t* order, const int* cell_start,
                            const int* cell_end, const int4* items, const int* n_items, int batch,
                            int nz, int nslow, int nfast, long long sB, long long sZ, long long sC,
                            float* out, cudaStream_t stream);

int bev_pool_gather_backward(const float* out_grad, int C, const uint32_t* order,
                             const int* cell_start, const int* cell_end, const int4* items,
                             const int* n_items, int batch, int nz, int nslow, int nfast,
                             long long sB, long long sZ, long long sC, float* x_grad,
                             cudaStream_t stream);

int lift_splat_forward(const float* depth, const float* feat_cl, int C, int D, int fhw,
                       const uint32_t* order, const int* cell_start, const int* cell_end,
                       const int4* items, const int* n_items, int batch, int nz, int nslow,
                       int nfast, long long sB, long long sZ, long long sC, float* out,
                       cudaStream_t stream);

int lift_splat_backward(const float* g_cl, const float* depth, const float* feat_cl,
                        const int* point_cell, long long n_pix, int C, int D, int fhw,
                        float* d_depth, float* d_feat_cl, cudaStream_t stream);

int transpose_batched(const float* in, float* out, int batch, int rows, int cols,
                      cudaStream_t stream);

int bev_pool_interval_forward(int b, int d, int h, int w, int n, int c, int n_intervals,
                              const float* x, const int* geom_feats, const int* interval_starts,
                              const int* interval_lengths, float* out, int zero_out,
                              cudaStream_t stream);

int bev_pool_interval_backward(int b, int d, int h, int w, int n, int c, int n_intervals,
                               const float* out_grad, const int* geom_feats,
                               const int* interval_starts, const int* interval_lengths,
                               float* x_grad, int zero_x_grad, cudaStream_t stream);

// x_grad[p, :] = grad_cl[point_cell[p], :] (zero rows where point_cell < 0); grad_cl = BEV gradient in
// cells-major rows [n_cells, C].
int bev_pool_point_backward(const float* grad_cl, const int* point_cell, long long n_points, int C,
                            float* x_grad, cudaStream_t stream);

// point_cell[p] = output cell of frustum point p (or -1) straight from the geometry: no sort.
int bev_point_cells(const float* geom, long long n_points, int batch, const float* off, const float* dx,
                    const float* nx_f, const int* nx_i, int fast_axis, int* point_cell,
                    cudaStream_t stream, int frames = 1);
// Sort-free lift + splat into a channels-last BEV map out_cl[n_cells, C] (zero-filled here) with
// vector float reductions; summation order is not fixed.
int lift_splat_atomic_forward(const float* depth, const float* feat_cl, const int* point_cell,
                              long long n_pixels, int C, int D, int fhw, long long n_cells, float* out_cl,
                              cudaStream_t stream);

}  // namespace dbev
