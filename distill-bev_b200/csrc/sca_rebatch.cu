// Per-camera re-batching of BEV queries for BEVFormer's spatial cross attention:
//   mmdet3d/models/transformer_modules/spatial_cross_attention.py:128-167  SpatialCrossAttention.forward
// The reference finds, per camera, the BEV queries whose reference points project into that camera
// (bev_mask[cam][0].sum(-1).nonzero() - batch element 0's mask for every batch element), copies them and their
// reference points into zero-padded [bs, cams, max_len, ...] tensors with a Python double loop over (batch, camera),
// runs the deformable attention on the re-batched queries, adds the results back into `slots` with another double loop
// (slots[j, index] += queries[j, i, :len]) and divides by the number of cameras that see the query.
// Here the index lists live on the device (idx[cam][k] = k-th hit query of the camera or -1, pos[cam][q] = rank of
// query q among the camera's hits or -1) and both directions are row kernels, 16-byte vectors, no atomics:
//   gather   out[b, cam, k, :] = scale[b, idx] * in[b, idx[cam][k], :]            (0 for padding)
//   reduce   out[b, q, :]      = scale[b, q] * sum_cam in[b, cam, pos[cam][q], :]  (cameras in order: deterministic)
// forward re-batch = gather, its backward = reduce; forward slots = reduce with scale = 1 / count, its backward = gather
// with the same scale. HBM-bound: one read + one write of the rows.
#include "sca_rebatch.cuh"

namespace dbev {

namespace {

__global__ void sca_gather_rows_kernel(const float* __restrict__ in, const int* __restrict__ idx, const float* __restrict__ scale,
                                       int bs, int cams, int max_len, int nq, int C, long long in_cstride, long long in_bstride,
                                       long long in_qstride, float* __restrict__ out) {
  const int quads = C >> 2;
  const long long total = (long long)bs * cams * max_len * quads;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int qd = (int)(i % quads);
    long long r = i / quads;
    const int k = (int)(r % max_len);
    r /= max_len;
    const int cam = (int)(r % cams), b = (int)(r / cams);
    const int q = idx[cam * max_len + k];
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (q >= 0) {
      v = *reinterpret_cast<const float4*>(in + cam * in_cstride + b * in_bstride + q * in_qstride + 4 * qd);
      if (scale) {
        const float s = scale[(long long)b * nq + q];
        v.x *= s, v.y *= s, v.z *= s, v.w *= s;
      }
    }
    *reinterpret_cast<float4*>(out + (((long long)b * cams + cam) * max_len + k) * C + 4 * qd) = v;
  }
}

__global__ void sca_reduce_rows_kernel(const float* __restrict__ in, const int* __restrict__ pos, const float* __restrict__ scale,
                                       int bs, int cams, int max_len, int nq, int C, float* __restrict__ out) {
  const int quads = C >> 2;
  const long long total = (long long)bs * nq * quads;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int qd = (int)(i % quads);
    const long long r = i / quads;
    const int q = (int)(r % nq), b = (int)(r / nq);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int cam = 0; cam < cams; ++cam) {
      const int k = pos[cam * nq + q];
      if (k >= 0) {
        const float4 v = *reinterpret_cast<const float4*>(in + (((long long)b * cams + cam) * max_len + k) * C + 4 * qd);
        acc.x += v.x, acc.y += v.y, acc.z += v.z, acc.w += v.w;
      }
    }
    if (scale) {
      const float s = scale[(long long)b * nq + q];
      acc.x *= s, acc.y *= s, acc.z *= s, acc.w *= s;
    }
    *reinterpret_cast<float4*>(out + ((long long)b * nq + q) * C + 4 * qd) = acc;
  }
}

int grid_for(long long total) {
  long long b = (total + 255) / 256;
  const long long cap = (long long)kNumSMs * 16;
  return (int)(b < cap ? (b > 0 ? b : 1) : cap);
}

}  // namespace

int sca_gather_rows(const float* in, const int* idx, const float* scale, int bs, int cams, int max_len, int nq, int C,
                    long long in_cam_stride, long long in_batch_stride, long long in_query_stride, float* out, cudaStream_t stream) {
  DBEV_CHECK_ARG(bs > 0 && cams > 0 && max_len >= 0 && nq > 0 && C > 0 && C % 4 == 0 && in_query_stride % 4 == 0 &&
                     in_batch_stride % 4 == 0,
                 "sca_gather_rows: C and the strides must be multiples of 4");
  if (max_len == 0) return DBEV_OK;
  sca_gather_rows_kernel<<<grid_for((long long)bs * cams * max_len * (C / 4)), 256, 0, stream>>>(
      in, idx, scale, bs, cams, max_len, nq, C, in_cam_stride, in_batch_stride, in_query_stride, out);
  DBEV_CHECK_LAUNCH("sca_gather_rows_kernel");
  return DBEV_OK;
}

int sca_reduce_rows(const float* in, const int* pos, const float* scale, int bs, int cams, int max_len, int nq, int C, float* out,
                    cudaStream_t stream) {
  DBEV_CHECK_ARG(bs > 0 && cams > 0 && max_len >= 0 && nq > 0 && C > 0 && C % 4 == 0, "sca_reduce_rows: C must be a multiple of 4");
  sca_reduce_rows_kernel<<<grid_for((long long)bs * nq * (C / 4)), 256, 0, stream>>>(in, pos, scale, bs, cams, max_len, nq, C, out);
  DBEV_CHECK_LAUNCH("sca_reduce_rows_kernel");
  return DBEV_OK;
}

}  // namespace dbev
