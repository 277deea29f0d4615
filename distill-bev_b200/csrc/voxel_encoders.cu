// Voxel encoders of the sparse LiDAR teachers.
//
//   * HardSimpleVFE.forward       mmdet3d/models/voxel_encoders/voxel_encoder.py:29-45
//   * voxelization / voxelization_virtual (DynamicVoxelEncoder, MVPFormer teacher)
//                                  mmdet3d/models/voxel_encoders/dynamic_voxel_encoder.py:8-17,19-68
// The per-voxel reduction itself is dbev_dynamic_scatter_forward (voxelize.cu): these kernels
// only produce what it consumes and post-process what it returns. All HBM-bound streaming.
#include "voxel_encoders.cuh"

namespace dbev {

namespace {

// points_mean = voxels[:, :, :nf].sum(dim=1) / num_points  (voxel_encoder.py:42-44; the sum runs
// over all max_points slots, padded slots hold zeros).
__global__ void hard_simple_vfe_kernel(const float* __restrict__ voxels,
                                       const int* __restrict__ num_points, long long m, int maxp,
                                       int F, int nf, float* __restrict__ out) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= m * nf) return;
  const long long v = t / nf;
  const int f = (int)(t - v * nf);
  const float* p = voxels + v * maxp * F + f;
  float s = 0.f;
  for (int j = 0; j < maxp; ++j) s += p[(long long)j * F];
  out[t] = s / (float)num_points[v];
}

struct DynVoxParams {
  float lo[3], hi[3], vs[3];  // (x, y, z)
};

__device__ __forceinline__ int find_sample(const int* __restrict__ offsets, int batch, int i) {
  int b = 0;
  while (b + 1 < batch && i >= offsets[b + 1]) ++b;
  return b;
}

// keep = lo <= p <= hi on all three axes (both ends inclusive, :9-11); coords (z, y, x) =
// trunc((p - lo) / vs) computed in fp32 then converted like .to(torch.int64) (:13).
__global__ void dynvoxel_coords_kernel(const float* __restrict__ points, int n, int F,
                                       const int* __restrict__ offsets, int batch,
                                       DynVoxParams prm, int check_flag,
                                       int* __restrict__ coors) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float x = points[(long long)i * F + 0], y = points[(long long)i * F + 1],
              z = points[(long long)i * F + 2];
  bool keep = x >= prm.lo[0] && x <= prm.hi[0] && y >= prm.lo[1] && y <= prm.hi[1] &&
              z >= prm.lo[2] && z <= prm.hi[2];
  if (check_flag) {  // virtual variant: only real (1) / painted (0) / virtual (-1) points exist
    const float flag = points[(long long)i * F + F - 2];
    keep = keep && (flag == 1.f || flag == 0.f || flag == -1.f);
  }
  int4 c = make_int4(-1, -1, -1, -1);
  if (keep) {
    c.x = find_sample(offsets, batch, i);
    c.y = (int)__fdiv_rn(__fsub_rn(z, prm.lo[2]), prm.vs[2]);
    c.z = (int)__fdiv_rn(__fsub_rn(y, prm.lo[1]), prm.vs[1]);
    c.w = (int)__fdiv_rn(__fsub_rn(x, prm.lo[0]), prm.vs[0]);
  }
  reinterpret_cast<int4*>(coors)[i] = c;
}

// voxelization_virtual (:27-50): every point becomes a 24-channel row; real points (flag
// points[:, -2] == 1) fill channels 0-5 = (x, y, z, c3, c4, last) and set channel 23; painted
// (flag 0) / virtual (flag -1) points fill 6-20 = points[:, :-2], 21 = flag, 22 = 1 / 0.
__global__ void dynvoxel_virtual_rows_kernel(const float* __restrict__ points, int n, int F,
                                             float* __restrict__ rows) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)n * 24) return;
  const int i = (int)(t / 24), ch = (int)(t % 24);
  const float* p = points + (long long)i * F;
  const float flag = p[F - 2];
  float v = 0.f;
  if (flag == 1.f) {
    if (ch < 5) v = p[ch];
    else if (ch == 5) v = p[F - 1];
    else if (ch == 23) v = 1.f;
  } else if (flag == 0.f || flag == -1.f) {
    if (ch >= 6 && ch < 21) v = p[ch - 6];
    else if (ch == 21) v = flag;
    else if (ch == 22) v = (flag == 0.f) ? 1.f : 0.f;
  }
  rows[t] = v;
}

// indicator = mean of channel 23; voxels that mix real and painted/virtual points are
// re-normalised (:58-66): [:6] /= indicator, [6:] /= (1 - indicator); channel 23 is dropped.
__global__ void dynvoxel_virtual_fix_kernel(const float* __restrict__ mean24, const int* m_dev,
                                            int m_max, float* __restrict__ out23) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int m = m_dev ? min(*m_dev, m_max) : m_max;
  if (t >= (long long)m * 23) return;
  const int v = (int)(t / 23), ch = (int)(t % 23);
  const float ind = mean24[(long long)v * 24 + 23];
  float x = mean24[(long long)v * 24 + ch];
  if (ind > 0.f && ind < 1.f) x = (ch < 6) ? x / ind : x / (1.f - ind);
  out23[t] = x;
}

}  // namespace

int hard_simple_vfe(const float* voxels, const int* num_points, long long m, int maxp, int F,
                    int num_features, float* out, cudaStream_t stream) {
  DBEV_CHECK_ARG(m >= 0 && maxp >= 1 && F >= 1 && num_features >= 1 && num_features <= F,
                 "hard_simple_vfe: bad sizes");
  if (m == 0) return DBEV_OK;
  hard_simple_vfe_kernel<<<ceil_div(m * num_features, 256), 256, 0, stream>>>(
      voxels, num_points, m, maxp, F, num_features, out);
  DBEV_CHECK_LAUNCH("hard_simple_vfe_kernel");
  return DBEV_OK;
}

int dynvoxel_coords(const float* points, int n, int F, const int* batch_offsets, int batch,
                    const float* pc_range_host6, const float* voxel_size_host3, int check_flag,
                    int* coors, cudaStream_t stream) {
  DBEV_CHECK_ARG(n >= 0 && F >= 3 && batch >= 1, "dynvoxel_coords: bad sizes");
  DBEV_CHECK_ARG(((uintptr_t)coors & 15) == 0, "dynvoxel_coords: coors must be 16-byte aligned");
  if (n == 0) return DBEV_OK;
  DynVoxParams prm;
  for (int a = 0; a < 3; ++a) {
    prm.lo[a] = pc_range_host6[a];
    prm.hi[a] = pc_range_host6[3 + a];
    prm.vs[a] = voxel_size_host3[a];
    DBEV_CHECK_ARG(prm.vs[a] > 0.f, "dynvoxel_coords: voxel size must be positive");
  }
  dynvoxel_coords_kernel<<<ceil_div(n, 256), 256, 0, stream>>>(points, n, F, batch_offsets, batch,
                                                               prm, check_flag, coors);
  DBEV_CHECK_LAUNCH("dynvoxel_coords_kernel");
  return DBEV_OK;
}

int dynvoxel_virtual_rows(const float* points, int n, int F, float* rows24, cudaStream_t stream) {
  DBEV_CHECK_ARG(n >= 0 && F == 17, "dynvoxel_virtual_rows: points must have 17 columns (got %d)", F);
  if (n == 0) return DBEV_OK;
  dynvoxel_virtual_rows_kernel<<<ceil_div((long long)n * 24, 256), 256, 0, stream>>>(points, n, F,
                                                                                     rows24);
  DBEV_CHECK_LAUNCH("dynvoxel_virtual_rows_kernel");
  return DBEV_OK;
}

int dynvoxel_virtual_fix(const float* mean24, const int* m_dev, int m_max, float* out23,
                         cudaStream_t stream) {
  DBEV_CHECK_ARG(m_max >= 0, "dynvoxel_virtual_fix: bad sizes");
  if (m_max == 0) return DBEV_OK;
  dynvoxel_virtual_fix_kernel<<<ceil_div((long long)m_max * 23, 256), 256, 0, stream>>>(
      mean24, m_dev, m_max, out23);
  DBEV_CHECK_LAUNCH("dynvoxel_virtual_fix_kernel");
  return DBEV_OK;
}

}  // namespace dbev
