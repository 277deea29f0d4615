// Voxel encoders of the sparse LiDAR teachers (HardSimpleVFE, DynamicVoxelEncoder); see
// voxel_encoders.cu for the reference lines.
#pragma once

#include "common.cuh"

namespace dbev {

int hard_simple_vfe(const float* voxels, const int* num_points, long long m, int maxp, int F,
                    int num_features, float* out, cudaStream_t stream);

int dynvoxel_coords(const float* points, int n, int F, const int* batch_offsets, int batch,
                    const float* pc_range_host6, const float* voxel_size_host3, int check_flag,
                    int* coors, cudaStream_t stream);

int dynvoxel_virtual_rows(const float* points, int n, int F, float* rows24, cudaStream_t stream);

int dynvoxel_virtual_fix(const float* mean24, const int* m_dev, int m_max, float* out23,
                         cudaStream_t stream);

}  // namespace dbev
