// Internal C++ interface of the voxelization kernels (see voxelize.cu).
#pragma once

#include "common.cuh"

namespace dbev {

int voxel_grid_size(const float* voxel_size, const float* coors_range, int* grid_xyz);

int dynamic_voxelize(const float* points, int n, int nfeat, const float* voxel_size,
                     const float* coors_range, int* coors, cudaStream_t stream);

int segment_sorted_keys(const uint32_t* skeys, const uint32_t* sidx, int n, uint32_t sentinel,
                        int* flags, int* excl, int* head_pos, int* nseg, void* sub, size_t sub_bytes,
                        cudaStream_t stream);

size_t hard_voxelize_ws_bytes(long long n);
int hard_voxelize(const float* points, int n, int nfeat, const float* voxel_size,
                  const float* coors_range, int max_points, int max_voxels, float* voxels,
                  int* coors, int* num_points_per_voxel, int* voxel_num, void* ws, size_t ws_bytes,
                  cudaStream_t stream);

size_t dynamic_scatter_ws_bytes(long long n);
int dynamic_scatter_forward(const float* feats, const int* coors, int n, int nfeat, int ncol,
                            const int* dims, int reduce_type, float* reduced_feats, int* out_coors,
                            int* coors_map, int* reduce_count, int* num_out, void* ws,
                            size_t ws_bytes, cudaStream_t stream);
int dynamic_scatter_backward(const float* grad_reduced, const float* feats, const float* reduced,
                             const int* coors_map, const int* reduce_count, long long n,
                             long long m, int nfeat, int reduce_type, float* grad_feats,
                             int* reduce_from_ws, cudaStream_t stream);

}  // namespace dbev
