// PillarFeatureNet (hard voxels, eval mode) fused kernel; see pillar_hard.cu.
#pragma once

#include "common.cuh"

namespace dbev {

int hard_pillar_encode(const float* voxels, const int* num_points, const int* coors, const int* m_dev, int m_max,
                       int max_points, int nfeat, const float* voxel_size_xy, float x_offset, float y_offset,
                       const float* weight, int nout, const float* bn_scale, const float* bn_shift, int legacy,
                       float* out, cudaStream_t stream);

}  // namespace dbev
