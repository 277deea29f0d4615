// Affinity distillation loss (bevdet_distill.py:703-752, :1294-1321); see affinity.cu.
#pragma once

#include "common.cuh"

namespace dbev {

size_t affinity_select_ws_bytes(int B, int HW);
int affinity_select(const float* mask_a, const float* mask_b, int B, int HW, int* row_cell,
                    int* row_offsets, void* ws, size_t ws_bytes, cudaStream_t stream);
int affinity_gather_rows(const float* feat, const int* row_cell, const int* row_offsets, int B,
                         int C, int HW, int k_total, float* rows, cudaStream_t stream);
size_t affinity_partial_floats(const int* row_offsets_host, int B);
int affinity_forward(const float* t_rows, const float* s_rows, const int* row_offsets_host, int B,
                     int C, int kind, float beta, float weight, float* partial, float* loss,
                     cudaStream_t stream);
int affinity_backward(const float* t_rows, const float* s_rows, const int* row_offsets_host, int B,
                      int C, int kind, float beta, float weight, const float* grad_loss,
                      float* d_s_rows, cudaStream_t stream);
int affinity_scatter_rows(const float* d_rows, const int* row_cell, const int* row_offsets, int B,
                          int C, int HW, int k_total, float* grad, cudaStream_t stream);

}  // namespace dbev
