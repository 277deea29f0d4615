// Teacher pillar path for B200: fused DynamicPillarFeatureNet (eval) and
// PointPillarsScatter, plus the LSS frustum geometry kernel.
//
// Reference behaviour reproduced here:
//   DynamicPillarFeatureNet.forward  mmdet3d/models/voxel_encoders/pillar_encoder.py:282-338
//     cluster_scatter (mean)         :303-304  -> DynamicScatter, scatter_points.py:53-107
//     map_voxel_center_to_point      :243-280  (dense [C, B*ny*nx] canvas temp)
//     decorations f_cluster/f_center :307-318, x_offset = vx/2 + pc_min (:88-89)
//     PFN layer Linear(no bias)+BN1d+ReLU :221-233, pfn_scatter (max) :330
//   PointPillarsScatter.forward_batch mmdet3d/models/middle_encoders/pillar_scatter.py:62-102
//   get_geometry                     mmdet3d/models/necks/view_transformer_mine.py:111-139
//
// Design (not a port). The reference runs two unique_dim sorts, two atomic
// scatters, a dense canvas round trip to broadcast the pillar mean back to the
// points, a cuBLAS GEMM with K = 10, BN, ReLU and a per-sample Python loop for
// the final scatter. Here ONE sort of the pillar keys gives contiguous
// segments; eight lanes per pillar compute the mean, decorate every
// point, applies the 10x64 linear + folded BN + ReLU from shared memory and
// keeps the running max - the decorated [N, 10] and the [N, 64] point features
// are never written to HBM. Reads: points once (+ ids); writes: [M, 64].
#include "pillar.cuh"

#include <math.h>

#include "sort.cuh"
#include "voxelize.cuh"

namespace dbev {

namespace {

constexpr int kMaxIn = 32;  // raw + decorations

struct PillarGeom {
  float vx, vy, x_offset, y_offset;
  int nfeat;      // raw point features
  int nin;        // nfeat + 3 (cluster) + 2 (center)
  int nout;       // PFN output channels
  int nx, ny, nz; // grid (keys are ((b*nz + z)*ny + y)*nx + x)
};

// EIGHT LANES per pillar, eight output channels per lane and 64-channel round (a pillar holds ~1-3
// points). Compared with one thread per channel this keeps 8x more pillars in flight per SM,
// which is what hides the head_pos -> sidx -> point chain of dependent loads; compared with one
// thread per pillar it keeps the warp's divergence over "points in this pillar" to 4 pillars.
// Weights and the folded BN constants are shared-memory broadcasts (two LDS.128 per input).
constexpr int kLanesPerPillar = 8;
constexpr int kChPerLane = 8;

template <int RAW_MAX>
__global__ void __launch_bounds__(256)
pillar_encode_kernel(const float* __restrict__ points, const uint32_t* __restrict__ skeys,
                     const uint32_t* __restrict__ sidx, const int* __restrict__ head_pos,
                     const int* __restrict__ nseg_ptr, PillarGeom g, const float* __restrict__ weight,
                     const float* __restrict__ bn_scale, const float* __restrict__ bn_shift,
                     float* __restrict__ voxel_feats, int* __restrict__ voxel_coors,
                     float* __restrict__ canvas, int canvas_cl) {
  extern __shared__ __align__(16) float sh[];
  const int nout_p = (g.nout + 3) & ~3;   // padded row so every lane's 8 weights are 16B aligned
  float* w_s = sh;                        // [nin][nout_p]
  float* sc_s = sh + g.nin * nout_p;      // [nout_p]
  float* sf_s = sc_s + nout_p;            // [nout_p]
  for (int i = threadIdx.x; i < g.nin * nout_p; i += blockDim.x) {
    const int c = i % nout_p, k = i / nout_p;  // nn.Linear weight is [nout, nin]
    w_s[i] = c < g.nout ? weight[c * g.nin + k] : 0.f;
  }
  for (int i = threadIdx.x; i < nout_p; i += blockDim.x) {
    sc_s[i] = i < g.nout ? bn_scale[i] : 0.f;
    sf_s[i] = i < g.nout ? bn_shift[i] : 0.f;
  }
  __syncthreads();
  const int nseg = *nseg_ptr;
  const int sub = threadIdx.x & (kLanesPerPillar - 1);
  const int group = (blockIdx.x * blockDim.x + threadIdx.x) / kLanesPerPillar;
  const int ngroups = gridDim.x * blockDim.x / kLanesPerPillar;
  for (int seg = group; seg < nseg; seg += ngroups) {
    const int s = head_pos[seg], e = head_pos[seg + 1];
    const uint32_t key = skeys[s];
    float mx = 0.f, my = 0.f, mz = 0.f;   // cluster_scatter (mean) in point order
    for (int j = s; j < e; ++j) {
      const float* p = points + (size_t)sidx[j] * g.nfeat;
      mx += p[0]; my += p[1]; mz += p[2];
    }
    const float cnt = (float)(e - s);
    mx /= cnt; my /= cnt; mz /= cnt;
    uint32_t r = key;
    const int cx = (int)(r % (uint32_t)g.nx); r /= (uint32_t)g.nx;
    const int cy = (int)(r % (uint32_t)g.ny); r /= (uint32_t)g.ny;
    const int cz = (int)(r % (uint32_t)g.nz); r /= (uint32_t)g.nz;
    if (sub == 0 && voxel_coors)
      *reinterpret_cast<int4*>(voxel_coors + (size_t)seg * 4) = make_int4((int)r, cz, cy, cx);
    const float ctr_x = __fadd_rn(__fmul_rn((float)cx, g.vx), g.x_offset);
    const float ctr_y = __fadd_rn(__fmul_rn((float)cy, g.vy), g.y_offset);
    for (int c0 = 0; c0 < g.nout; c0 += kLanesPerPillar * kChPerLane) {
      const int cb = c0 + sub * kChPerLane;
      if (cb >= g.nout) continue;
      float best[kChPerLane];
#pragma unroll
      for (int i = 0; i < kChPerLane; ++i) best[i] = -INFINITY;
      for (int j = s; j < e; ++j) {
        const float* p = points + (size_t)sidx[j] * g.nfeat;
        float pv[RAW_MAX];
#pragma unroll
        for (int k = 0; k < RAW_MAX; ++k) pv[k] = k < g.nfeat ? p[k] : 0.f;
        const float dv[5] = {pv[0] - mx, pv[1] - my, pv[2] - mz, pv[0] - ctr_x, pv[1] - ctr_y};
        float acc[kChPerLane];
#pragma unroll
        for (int i = 0; i < kChPerLane; ++i) acc[i] = 0.f;
#pragma unroll
        for (int k = 0; k < RAW_MAX + 5; ++k) {
          const bool raw = k < RAW_MAX;
          if (raw && k >= g.nfeat) continue;
          const float v = raw ? pv[k < RAW_MAX ? k : 0] : dv[k < RAW_MAX ? 0 : k - RAW_MAX];
          const float* wr = w_s + (raw ? k : g.nfeat + k - RAW_MAX) * nout_p + cb;
          const float4 w0 = *reinterpret_cast<const float4*>(wr);
          const float4 w1 = cb + 4 < nout_p ? *reinterpret_cast<const float4*>(wr + 4) : make_float4(0, 0, 0, 0);
          acc[0] += v * w0.x; acc[1] += v * w0.y; acc[2] += v * w0.z; acc[3] += v * w0.w;
          acc[4] += v * w1.x; acc[5] += v * w1.y; acc[6] += v * w1.z; acc[7] += v * w1.w;
        }
#pragma unroll
        for (int i = 0; i < kChPerLane; ++i) {
          const int c = min(cb + i, nout_p - 1);
          best[i] = fmaxf(best[i], fmaxf(acc[i] * sc_s[c] + sf_s[c], 0.f));  // folded BN, ReLU, max
        }
      }
      // destinations: the [M, nout] pillar table and / or straight into the BEV canvas
      // (PointPillarsScatter fused: with nz == 1 the key IS the canvas cell)
      float* rows[2] = {voxel_feats ? voxel_feats + (size_t)seg * g.nout + cb : nullptr,
                        (canvas && canvas_cl) ? canvas + (size_t)key * g.nout + cb : nullptr};
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        float* o = rows[t];
        if (!o) continue;
        if ((g.nout & 3) == 0) {
          *reinterpret_cast<float4*>(o) = make_float4(best[0], best[1], best[2], best[3]);
          if (cb + 4 < g.nout) *reinterpret_cast<float4*>(o + 4) = make_float4(best[4], best[5], best[6], best[7]);
        } else {
#pragma unroll
          for (int i = 0; i < kChPerLane; ++i)
            if (cb + i < g.nout) o[i] = best[i];
        }
      }
      if (canvas && !canvas_cl) {  // NCHW canvas: one strided store per channel
        const size_t plane = (size_t)g.ny * g.nx;
        float* o = canvas + ((size_t)r * g.nout + cb) * plane + (size_t)cy * g.nx + cx;
#pragma unroll
        for (int i = 0; i < kChPerLane; ++i)
          if (cb + i < g.nout) o[(size_t)i * plane] = best[i];
      }
    }
  }
}

// keys of batched points: dynamic voxelize on the fly (batch id from offsets)
__global__ void __launch_bounds__(256)
pillar_keys_kernel(const float* __restrict__ points, const int* __restrict__ batch_offsets,
                   int batch, int n, int nfeat, float vx, float vy, float vz, float xmin, float ymin,
                   float zmin, int gx, int gy, int gz, uint32_t sentinel,
                   uint32_t* __restrict__ keys, int* __restrict__ coors) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int b = 0;
  while (b + 1 < batch && i >= batch_offsets[b + 1]) ++b;
  const float* p = points + (size_t)i * nfeat;
  const int cx = (int)floorf(__fdiv_rn(__fsub_rn(p[0], xmin), vx));
  const int cy = (int)floorf(__fdiv_rn(__fsub_rn(p[1], ymin), vy));
  const int cz = (int)floorf(__fdiv_rn(__fsub_rn(p[2], zmin), vz));
  const bool ok = cx >= 0 && cx < gx && cy >= 0 && cy < gy && cz >= 0 && cz < gz;
  keys[i] = ok ? (uint32_t)((((long long)b * gz + cz) * gy + cy) * gx + cx) : sentinel;
  if (coors) {
    coors[(size_t)i * 4 + 0] = b;
    coors[(size_t)i * 4 + 1] = ok ? cz : -1;
    coors[(size_t)i * 4 + 2] = ok ? cy : -1;
    coors[(size_t)i * 4 + 3] = ok ? cx : -1;
  }
}

// keys from caller-provided coors [n, 4] = (b, z, y, x); any negative component drops the point
__global__ void __launch_bounds__(256)
pillar_keys_from_coors_kernel(const int* __restrict__ coors, int n, int batch, int gx, int gy, int gz,
                              uint32_t sentinel, uint32_t* __restrict__ keys) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int b = coors[(size_t)i * 4 + 0], z = coors[(size_t)i * 4 + 1], y = coors[(size_t)i * 4 + 2],
            x = coors[(size_t)i * 4 + 3];
  const bool ok = b >= 0 && b < batch && z >= 0 && z < gz && y >= 0 && y < gy && x >= 0 && x < gx;
  keys[i] = ok ? (uint32_t)((((long long)b * gz + z) * gy + y) * gx + x) : sentinel;
}

// canvas[b, :, y, x] = voxel_feats[m, :]; channels_last: canvas is [B, ny, nx, C]
__global__ void __launch_bounds__(256)
pillar_scatter_kernel(const float* __restrict__ voxel_feats, const int* __restrict__ coors,
                      const int* __restrict__ m_ptr, int m_max, int C, int ny, int nx,
                      int channels_last, float* __restrict__ canvas) {
  const int m = m_ptr ? min(*m_ptr, m_max) : m_max;
  const long long total = (long long)m * C;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(t / C), c = (int)(t % C);
    const int b = coors[(size_t)v * 4 + 0], y = coors[(size_t)v * 4 + 2], x = coors[(size_t)v * 4 + 3];
    if (b < 0 || y < 0 || y >= ny || x < 0 || x >= nx) continue;
    const float val = voxel_feats[t];
    if (channels_last) canvas[(((size_t)b * ny + y) * nx + x) * C + c] = val;
    else canvas[(((size_t)b * C + c) * ny + y) * nx + x] = val;
  }
}

// ---- LSS geometry ------------------------------------------------------------

// per camera: A = inv(post_rots), M = rots * inv(intrins) (3x3 each, fp64 inside)
__global__ void lss_camera_mats_kernel(const float* __restrict__ rots, const float* __restrict__ intrins,
                                       const float* __restrict__ post_rots, int n_cams,
                                       float* __restrict__ mats /*[n][18]*/) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_cams) return;
  auto inv3 = [](const float* m, double* o) {
    const double a = m[0], b = m[1], c = m[2], d = m[3], e = m[4], f = m[5], g = m[6], h = m[7], k = m[8];
    const double det = a * (e * k - f * h) - b * (d * k - f * g) + c * (d * h - e * g);
    const double id = 1.0 / det;
    o[0] = (e * k - f * h) * id; o[1] = (c * h - b * k) * id; o[2] = (b * f - c * e) * id;
    o[3] = (f * g - d * k) * id; o[4] = (a * k - c * g) * id; o[5] = (c * d - a * f) * id;
    o[6] = (d * h - e * g) * id; o[7] = (b * g - a * h) * id; o[8] = (a * e - b * d) * id;
  };
  double A[9], I[9];
  inv3(post_rots + (size_t)i * 9, A);
  inv3(intrins + (size_t)i * 9, I);
  float If[9];
  for (int k = 0; k < 9; ++k) {
    mats[(size_t)i * 18 + k] = (float)A[k];
    If[k] = (float)I[k];  // the reference rounds inv(intrins) to fp32 before the product
  }
  const float* R = rots + (size_t)i * 9;
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) {
      float acc = R[r * 3 + 0] * If[0 * 3 + c];
      acc += R[r * 3 + 1] * If[1 * 3 + c];
      acc += R[r * 3 + 2] * If[2 * 3 + c];
      mats[(size_t)i * 18 + 9 + r * 3 + c] = acc;
    }
}

__global__ void __launch_bounds__(256)
lss_geometry_kernel(const float* __restrict__ frustum, const float* __restrict__ mats,
                    const float* __restrict__ trans, const float* __restrict__ post_trans,
                    long long n_points, int pts_per_cam, float* __restrict__ geom) {
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_points) return;
  const int cam = (int)(p / pts_per_cam);
  const int f = (int)(p % pts_per_cam);
  const float* m = mats + (size_t)cam * 18;
  const float* pt = post_trans + (size_t)cam * 3;
  const float x0 = frustum[f * 3 + 0] - pt[0], y0 = frustum[f * 3 + 1] - pt[1],
              z0 = frustum[f * 3 + 2] - pt[2];
  float x1 = m[0] * x0 + m[1] * y0 + m[2] * z0;
  float y1 = m[3] * x0 + m[4] * y0 + m[5] * z0;
  const float z1 = m[6] * x0 + m[7] * y0 + m[8] * z0;
  x1 *= z1;
  y1 *= z1;
  const float* t = trans + (size_t)cam * 3;
  geom[p * 3 + 0] = (m[9] * x1 + m[10] * y1 + m[11] * z1) + t[0];
  geom[p * 3 + 1] = (m[12] * x1 + m[13] * y1 + m[14] * z1) + t[1];
  geom[p * 3 + 2] = (m[15] * x1 + m[16] * y1 + m[17] * z1) + t[2];
}

}  // namespace

// ---------------------------------------------------------------------------

size_t pillar_encode_ws_bytes(long long n) {
  return 8 * align_up((size_t)(n + 2) * 4) + radix_sort_ws_bytes(n) + scan_ws_bytes(n) + 8192;
}

int pillar_encode(const float* points, const int* batch_offsets, const int* coors_in, int batch,
                  int n, int nfeat, const float* voxel_size, const float* coors_range,
                  float x_offset, float y_offset, const float* weight, int nout,
                  const float* bn_scale, const float* bn_shift, float* voxel_feats,
                  int* voxel_coors, int* num_voxels, int* point_coors, float* canvas,
                  int canvas_channels_last, int zero_canvas, void* ws, size_t ws_bytes,
                  cudaStream_t stream) {
  DBEV_CHECK_ARG(n >= 0 && batch > 0 && nfeat >= 3 && nfeat + 5 <= kMaxIn,
                 "pillar_encode: bad sizes n=%d batch=%d nfeat=%d", n, batch, nfeat);
  DBEV_CHECK_ARG(canvas != nullptr || (voxel_feats != nullptr && voxel_coors != nullptr),
                 "pillar_encode: no output given");
  DBEV_CHECK_ARG(nout > 0, "pillar_encode: output channels must be positive (got %d)", nout);
  int grid3[3];
  int rc = voxel_grid_size(voxel_size, coors_range, grid3);
  if (rc != DBEV_OK) return rc;
  const unsigned long long nkeys = (unsigned long long)batch * grid3[0] * grid3[1] * grid3[2];
  DBEV_CHECK_ARG(nkeys < 0xfffffff0ULL, "pillar_encode: key space exceeds 32 bits");
  if (canvas) {
    DBEV_CHECK_ARG(grid3[2] == 1, "pillar_canvas: pillars need a single z bin (got nz=%d)", grid3[2]);
    if (zero_canvas)
      DBEV_CUDA(cudaMemsetAsync(canvas, 0, (size_t)nkeys * nout * sizeof(float), stream));
  }
  if (n == 0) {
    DBEV_CUDA(cudaMemsetAsync(num_voxels, 0, sizeof(int), stream));
    return DBEV_OK;
  }
  const uint32_t sentinel = (uint32_t)nkeys;
  Workspace w(ws, ws_bytes);
  uint32_t* keys0 = w.take<uint32_t>(n);
  uint32_t* keys1 = w.take<uint32_t>(n);
  uint32_t* vals0 = w.take<uint32_t>(n);
  uint32_t* vals1 = w.take<uint32_t>(n);
  int* flags = w.take<int>(n);
  int* excl = w.take<int>(n);
  int* head_pos = w.take<int>((size_t)n + 2);
  if (!w.ok()) {
    set_last_error("pillar_encode: workspace too small");
    return DBEV_ERR_WORKSPACE;
  }
  const size_t consumed = align_up(w.used);
  void* sub = (char*)ws + consumed;
  const size_t sub_bytes = ws_bytes > consumed ? ws_bytes - consumed : 0;
  const int grid = ceil_div(n, 256);
  DBEV_CHECK_ARG(coors_in != nullptr || batch_offsets != nullptr,
                 "pillar_encode: need either coors [n,4] or batch_offsets [batch+1]");
  if (coors_in)
    pillar_keys_from_coors_kernel<<<grid, 256, 0, stream>>>(coors_in, n, batch, grid3[0], grid3[1],
                                                            grid3[2], sentinel, keys0);
  else
    pillar_keys_kernel<<<grid, 256, 0, stream>>>(points, batch_offsets, batch, n, nfeat,
                                                 voxel_size[0], voxel_size[1], voxel_size[2],
                                                 coors_range[0], coors_range[1], coors_range[2],
                                                 grid3[0], grid3[1], grid3[2], sentinel, keys0,
                                                 point_coors);
  uint32_t* keys[2] = {keys0, keys1};
  uint32_t* vals[2] = {vals0, vals1};
  int sel = 0;
  rc = radix_sort_pairs(keys, vals, true, n, bits_for(nkeys + 1), sub, sub_bytes, stream, &sel);
  if (rc != DBEV_OK) return rc;
  rc = segment_sorted_keys(keys[sel], vals[sel], n, sentinel, flags, excl, head_pos, num_voxels, sub,
                           sub_bytes, stream);
  if (rc != DBEV_OK) return rc;
  PillarGeom g;
  g.vx = voxel_size[0]; g.vy = voxel_size[1];
  g.x_offset = x_offset;  // vx / 2 + pc_min_x, evaluated by the caller in double (pillar_encoder.py:88-89)
  g.y_offset = y_offset;
  g.nfeat = nfeat; g.nin = nfeat + 5; g.nout = nout;
  g.nx = grid3[0]; g.ny = grid3[1]; g.nz = grid3[2];
  const int nout_p = (nout + 3) & ~3;
  const size_t smem = ((size_t)g.nin * nout_p + 2 * nout_p) * sizeof(float);
  DBEV_CHECK_ARG(smem <= 48 * 1024, "pillar_encode: weights (%zu B) exceed 48 KB of shared memory", smem);
  const int egrid = min(kNumSMs * 8, ceil_div((long long)n * kLanesPerPillar, 256));
#define DBEV_LAUNCH_ENCODE(RAW)                                                                       \
  pillar_encode_kernel<RAW><<<egrid, 256, smem, stream>>>(points, keys[sel], vals[sel], head_pos,     \
                                                          num_voxels, g, weight, bn_scale, bn_shift,  \
                                                          voxel_feats, voxel_coors, canvas,           \
                                                          canvas_channels_last)
  if (nfeat <= 5) DBEV_LAUNCH_ENCODE(5);
  else if (nfeat <= 8) DBEV_LAUNCH_ENCODE(8);
  else DBEV_LAUNCH_ENCODE(kMaxIn - 5);
#undef DBEV_LAUNCH_ENCODE
  DBEV_CHECK_LAUNCH("pillar_encode_kernel");
  return DBEV_OK;
}

int pillar_scatter(const float* voxel_feats, const int* coors, const int* m_dev, int m_max, int C,
                   int batch, int ny, int nx, int channels_last, int zero_canvas, float* canvas,
                   cudaStream_t stream) {
  DBEV_CHECK_ARG(m_max >= 0 && C > 0 && batch > 0 && ny > 0 && nx > 0, "pillar_scatter: bad sizes");
  if (zero_canvas)
    DBEV_CUDA(cudaMemsetAsync(canvas, 0, (size_t)batch * C * ny * nx * sizeof(float), stream));
  if (m_max == 0) return DBEV_OK;
  pillar_scatter_kernel<<<kNumSMs * 8, 256, 0, stream>>>(voxel_feats, coors, m_dev, m_max, C, ny, nx,
                                                        channels_last, canvas);
  DBEV_CHECK_LAUNCH("pillar_scatter_kernel");
  return DBEV_OK;
}

int lss_geometry(const float* frustum, int pts_per_cam, const float* rots, const float* trans,
                 const float* intrins, const float* post_rots, const float* post_trans, int n_cams,
                 float* mats_ws, float* geom, cudaStream_t stream) {
  DBEV_CHECK_ARG(pts_per_cam > 0 && n_cams > 0, "lss_geometry: bad sizes");
  lss_camera_mats_kernel<<<ceil_div(n_cams, 64), 64, 0, stream>>>(rots, intrins, post_rots, n_cams,
                                                                  mats_ws);
  const long long n = (long long)n_cams * pts_per_cam;
  lss_geometry_kernel<<<ceil_div(n, 256), 256, 0, stream>>>(frustum, mats_ws, trans, post_trans, n,
                                                            pts_per_cam, geom);
  DBEV_CHECK_LAUNCH("lss_geometry_kernel");
  return DBEV_OK;
}

}  // namespace dbev
