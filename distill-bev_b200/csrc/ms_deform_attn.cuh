// Multi-scale deformable attention (BEVFormer student); see ms_deform_attn.cu.
#pragma once

#include "common.cuh"

namespace dbev {

int ms_deform_attn_forward(const float* value, const long long* spatial_shapes,
                           const long long* level_start, const float* sampling_loc,
                           const float* attn_weight, int bs, int num_keys, int heads, int dim,
                           int num_queries, int levels, int points, float* out, cudaStream_t stream);
int ms_deform_attn_backward(const float* value, const long long* spatial_shapes,
                            const long long* level_start, const float* sampling_loc,
                            const float* attn_weight, const float* grad_out, int bs, int num_keys,
                            int heads, int dim, int num_queries, int levels, int points,
                            float* grad_value, float* grad_loc, float* grad_attn, cudaStream_t stream);

}  // namespace dbev
