// Per-camera query re-batching of BEVFormer's SpatialCrossAttention; see sca_rebatch.cu.
#pragma once

#include "common.cuh"

namespace dbev {

int sca_gather_rows(const float* in, const int* idx, const float* scale, int bs, int cams, int max_len, int nq, int C,
                    long long in_cam_stride, long long in_batch_stride, long long in_query_stride, float* out, cudaStream_t stream);

int sca_reduce_rows(const float* in, const int* pos, const float* scale, int bs, int cams, int max_len, int nq, int C, float* out,
                    cudaStream_t stream);

}  // namespace dbev
