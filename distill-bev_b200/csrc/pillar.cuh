// Internal C++ interface of the pillar-encoder / scatter / geometry kernels (see pillar.cu).
#pragma once

#include "common.cuh"

namespace dbev {

size_t pillar_encode_ws_bytes(long long n);

int pillar_encode(const float* points, const int* batch_offsets, const int* coors_in, int batch,
                  int n, int nfeat, const float* voxel_size, const float* coors_range,
                  float x_offset, float y_offset, const float* weight, int nout,
                  const float* bn_scale, const float* bn_shift, float* voxel_feats,
                  int* voxel_coors, int* num_voxels, int* point_coors, float* canvas,
                  int canvas_channels_last, int zero_canvas, void* ws, size_t ws_bytes,
                  cudaStream_t stream);

int pillar_scatter(const float* voxel_feats, const int* coors, const int* m_dev, int m_max, int C,
                   int batch, int ny, int nx, int channels_last, int zero_canvas, float* canvas,
                   cudaStream_t stream);

int lss_geometry(const float* frustum, int pts_per_cam, const float* rots, const float* trans,
                 const float* intrins, const float* post_rots, const float* post_trans, int n_cams,
                 float* mats_ws, float* geom, cudaStream_t stream);

}  // namespace dbev
