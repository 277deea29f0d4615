// PillarFeatureNet (hard-voxel variant) in eval mode as ONE kernel:
//   mmdet3d/models/voxel_encoders/pillar_encoder.py:95-162   PillarFeatureNet.forward
//   mmdet3d/models/voxel_encoders/utils.py:107-181           PFNLayer (Linear no bias -> BatchNorm1d -> ReLU -> max over points)
// the voxel encoder of the shipped pillar teacher (configs/_base_/models/centerpoint_02pillar_second_secfpn_nus.py:6-13:
// in_channels 5, feat_channels [64], legacy=False). The reference runs ~12 torch kernels over the padded
// [M, max_points, F + 5] tensor; here eight lanes own a pillar: they read its max_points x F block once, decorate every
// point in registers (offset to the pillar mean of ALL max_points rows / num_points, offset to the pillar centre),
// zero the padded rows, apply the 10 -> 64 linear layer with the folded BatchNorm + ReLU and keep the running maximum.
// Reproduced quirks: padded rows enter the maximum as relu(bn_shift) (they are zeroed BEFORE the linear layer,
// :149-153); with legacy=True the centre offsets overwrite x, y of the raw features as well (:132-139, f_center is a view).
// HBM-bound: M * (max_points * F * 4 + 20 + nout * 4) bytes.
#include "pillar_hard.cuh"

namespace dbev {

namespace {

constexpr int kHpLanes = 8;      // lanes per pillar
constexpr int kHpCh = 8;         // output channels per lane and pass
constexpr int kHpMaxIn = 32;     // raw features + 5 decorations

struct HardPillarArgs {
  const float* voxels; const int* num_points; const int* coors; const int* m_dev; int m_max;
  int max_points, nfeat, nout, legacy;
  float vx, vy, x_offset, y_offset;
  const float* weight; const float* bn_scale; const float* bn_shift;
  float* out;
};

__global__ void __launch_bounds__(256) hard_pillar_encode_kernel(HardPillarArgs a) {
  extern __shared__ __align__(16) float sh[];
  const int nin = a.nfeat + 5;
  float* w_s = sh;                       // [nin][nout]
  float* sc_s = sh + nin * a.nout;
  float* sf_s = sc_s + a.nout;
  for (int i = threadIdx.x; i < nin * a.nout; i += blockDim.x) {
    const int c = i % a.nout, k = i / a.nout;           // nn.Linear weight is [nout, nin]
    w_s[i] = a.weight[c * nin + k];
  }
  for (int i = threadIdx.x; i < a.nout; i += blockDim.x) sc_s[i] = a.bn_scale[i], sf_s[i] = a.bn_shift[i];
  __syncthreads();
  const int m_total = a.m_dev ? min(*a.m_dev, a.m_max) : a.m_max;
  const int sub = threadIdx.x & (kHpLanes - 1);
  const int group = (blockIdx.x * blockDim.x + threadIdx.x) / kHpLanes;
  const int ngroups = gridDim.x * blockDim.x / kHpLanes;
  for (int m = group; m < m_total; m += ngroups) {
    const float* v = a.voxels + (size_t)m * a.max_points * a.nfeat;
    const int n = a.num_points[m];
    float sx = 0.f, sy = 0.f, sz = 0.f;              // sum over ALL rows (padded rows are whatever the caller left there)
    for (int t = 0; t < a.max_points; ++t) {
      sx += v[t * a.nfeat], sy += v[t * a.nfeat + 1], sz += v[t * a.nfeat + 2];
    }
    const float cnt = (float)n;
    const float mx = sx / cnt, my = sy / cnt, mz = sz / cnt;
    const int cx = a.coors[(size_t)m * 4 + 3], cy = a.coors[(size_t)m * 4 + 2];
    const float ctr_x = __fadd_rn(__fmul_rn((float)cx, a.vx), a.x_offset);
    const float ctr_y = __fadd_rn(__fmul_rn((float)cy, a.vy), a.y_offset);
    for (int c0 = 0; c0 < a.nout; c0 += kHpLanes * kHpCh) {
      const int cb = c0 + sub * kHpCh;
      if (cb >= a.nout) continue;
      float best[kHpCh];
#pragma unroll
      for (int i = 0; i < kHpCh; ++i) best[i] = -INFINITY;
      const int real = n < a.max_points ? n : a.max_points;
      for (int t = 0; t < real; ++t) {
        const float* p = v + t * a.nfeat;
        float acc[kHpCh];
#pragma unroll
        for (int i = 0; i < kHpCh; ++i) acc[i] = 0.f;
        const float fx = p[0] - ctr_x, fy = p[1] - ctr_y;
        for (int k = 0; k < nin; ++k) {
          float f;
          if (k < a.nfeat) f = (a.legacy && k == 0) ? fx : ((a.legacy && k == 1) ? fy : p[k]);
          else if (k == a.nfeat) f = p[0] - mx;
          else if (k == a.nfeat + 1) f = p[1] - my;
          else if (k == a.nfeat + 2) f = p[2] - mz;
          else f = k == a.nfeat + 3 ? fx : fy;
          const float* wr = w_s + k * a.nout + cb;
#pragma unroll
          for (int i = 0; i < kHpCh; ++i)
            if (cb + i < a.nout) acc[i] = fmaf(f, wr[i], acc[i]);
        }
#pragma unroll
        for (int i = 0; i < kHpCh; ++i)
          if (cb + i < a.nout) best[i] = fmaxf(best[i], fmaxf(fmaf(acc[i], sc_s[cb + i], sf_s[cb + i]), 0.f));
      }
      if (real < a.max_points) {           // padded rows: zero features -> relu(shift)
#pragma unroll
        for (int i = 0; i < kHpCh; ++i)
          if (cb + i < a.nout) best[i] = fmaxf(best[i], fmaxf(sf_s[cb + i], 0.f));
      }
      float* o = a.out + (size_t)m * a.nout + cb;
#pragma unroll
      for (int i = 0; i < kHpCh; ++i)
        if (cb + i < a.nout) o[i] = best[i];
    }
  }
}

}  // namespace

int hard_pillar_encode(const float* voxels, const int* num_points, const int* coors, const int* m_dev, int m_max,
                       int max_points, int nfeat, const float* voxel_size_xy, float x_offset, float y_offset,
                       const float* weight, int nout, const float* bn_scale, const float* bn_shift, int legacy,
                       float* out, cudaStream_t stream) {
  DBEV_CHECK_ARG(m_max >= 0 && max_points > 0 && nfeat >= 3 && nfeat + 5 <= kHpMaxIn && nout > 0,
                 "hard_pillar_encode: bad sizes (F=%d, nout=%d)", nfeat, nout);
  if (m_max == 0) return DBEV_OK;
  HardPillarArgs a;
  a.voxels = voxels, a.num_points = num_points, a.coors = coors, a.m_dev = m_dev, a.m_max = m_max;
  a.max_points = max_points, a.nfeat = nfeat, a.nout = nout, a.legacy = legacy ? 1 : 0;
  a.vx = voxel_size_xy[0], a.vy = voxel_size_xy[1], a.x_offset = x_offset, a.y_offset = y_offset;
  a.weight = weight, a.bn_scale = bn_scale, a.bn_shift = bn_shift, a.out = out;
  const size_t smem = ((size_t)(nfeat + 5) * nout + 2 * nout) * sizeof(float);
  DBEV_CHECK_ARG(smem <= 48 * 1024, "hard_pillar_encode: weights (%zu B) exceed 48 KB of shared memory", smem);
  const int grid = min(kNumSMs * 8, ceil_div((long long)m_max * kHpLanes, 256));
  hard_pillar_encode_kernel<<<grid, 256, smem, stream>>>(a);
  DBEV_CHECK_LAUNCH("hard_pillar_encode_kernel");
  return DBEV_OK;
}

}  // namespace dbev
