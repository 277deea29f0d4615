// Sparse 3D convolution of the LiDAR teacher's middle encoder.
//
// Reference: mmdet3d/ops/spconv (vendored spconv v1):
//   * rulebook      getIndicePair            include/spconv/spconv_ops.h:28-141,
//                   getValidOutPos / getIndicePairsConv / getIndicePairsSubM
//                                             include/spconv/geometry.h:24-84,141-199,259-311
//   * convolution   indiceConv               include/spconv/spconv_ops.h:261-361
//                   (27 x gather -> mm -> scatter-add, one host sync per call, :271)
//   * densify       SparseConvTensor.dense   structure.py:53-64
//
// B200 design: the rulebook is kept OUTPUT-major as a neighbour table nbr[kvol][n_out]
// (input row feeding output o through kernel offset k, or -1). One kernel then computes
// out[o,:] = sum_k in[nbr[k][o],:] * W[k] for a tile of outputs: no [nHot, C] gather /
// scatter buffers, no float atomics (fixed summation order: k ascending), no host sync,
// and the BatchNorm (eval) / residual / ReLU that follow every sparse conv in
// SparseEncoder (middle_encoders/sparse_encoder.py:97-128, ops/sparse_block.py:101-121)
// are applied in the epilogue. Coordinates are found through an open-addressing hash of
// the linearised (b,z,y,x) instead of the reference's dense batch*Z*Y*X int grid
// (420 MB per sample at 41x1600x1600, spconv_ops.h:60-62).
#pragma once

#include "common.cuh"

namespace dbev {

struct SpConvGeom {
  int k[3], s[3], p[3], d[3];  // kernel / stride / padding / dilation in (z, y, x) order
  int in_shape[3];             // input spatial shape (z, y, x)
  int out_shape[3];            // output spatial shape (z, y, x)
  int batch;
  int kvol() const { return k[0] * k[1] * k[2]; }
};

// upper bound of the number of distinct outputs n_in inputs can produce
long long spconv_max_out(long long n_in, const SpConvGeom& g);

size_t spconv_ws_bytes(long long n_in, long long max_out);

// nbr[kvol][n_out] for given output coordinates (submanifold conv: out_coors == in_coors).
int spconv_table(const int* in_coors, int n_in, const int* out_coors, int n_out,
                 const SpConvGeom& g, int* nbr, void* ws, size_t ws_bytes, cudaStream_t stream);

// Distinct output cells of a strided SparseConv (unsorted linear keys) and their count.
int spconv_out_candidates(const int* in_coors, int n_in, const SpConvGeom& g, uint32_t* out_keys,
                          long long max_out, int* n_out_dev, void* ws, size_t ws_bytes,
                          cudaStream_t stream);

// Sort the n_out keys (-> lexicographic (b,z,y,x) order, what torch::_unique gives the
// reference's GPU path, spconv_ops.h:131), decode them and build the neighbour table.
int spconv_out_table(const int* in_coors, int n_in, const SpConvGeom& g, uint32_t* out_keys,
                     int n_out, int* out_coors, int* nbr, void* ws, size_t ws_bytes,
                     cudaStream_t stream);

// reference rulebook format <-> neighbour table
int spconv_pairs_from_table(const int* nbr, int kvol, int n_out, int pair_stride,
                            int* indice_pairs, int* indice_num, cudaStream_t stream);
int spconv_table_from_pairs(const int* indice_pairs, const int* indice_num, int kvol,
                            int pair_stride, int inverse, int n_out, int* nbr,
                            cudaStream_t stream);

// out[o,:] = act((sum_k in[nbr[k][o],:] . W[k]) * scale + shift + residual[o,:])
int spconv_forward(const float* in_feats, int c_in, const float* weight, int c_out,
                   const int* nbr, int kvol, int n_out, const float* scale, const float* shift,
                   const float* residual, int relu, float* out, cudaStream_t stream);

// Tensor-core path (spconv_tc.cu): tcgen05 3xTF32 implicit GEMM for C_in, C_out in {32,64,128}.
bool spconv_tc_supported(int c_in, int c_out, int kvol);
// weight [kvol][c_in][c_out] -> wt_hi / wt_lo [kvol][c_out][c_in] (TF32 head / remainder)
int spconv_pack_weights(const float* weight, int kvol, int c_in, int c_out, float* wt_hi,
                        float* wt_lo, cudaStream_t stream);
int spconv_forward_tc(const float* in_feats, int c_in, const float* wt_hi, const float* wt_lo,
                      int c_out, const int* nbr, int kvol, int n_out, const float* scale,
                      const float* shift, const float* residual, int relu, float* out,
                      cudaStream_t stream);

// dense[b, c*Z + z, y, x] = feats[m, c]; the whole tensor is written (zero fill included).
int spconv_dense(const float* feats, const int* coors, int m, int C, int batch, int Z, int Y,
                 int X, float* dense, cudaStream_t stream);

}  // namespace dbev
