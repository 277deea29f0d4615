// Multi-scale deformable attention for the BEVFormer student.
// Reference call sites: MultiScaleDeformableAttnFunction_fp32
//   mmdet3d/models/transformer_modules/multi_scale_deformable_attn_function.py:90-165 ->
//   mmcv 1.6.0 `_ext.ms_deform_attn_forward / _backward` (third party, not in the reference tree);
//   the published semantics = mmcv.ops.multi_scale_deform_attn.multi_scale_deformable_attn_pytorch
//   (F.grid_sample, bilinear, zeros padding, align_corners=False per level, weighted sum) — parity
//   is checked against that formula, UNPINNED by any reference-side fixture.
//   value [bs, num_keys, heads, dim]; spatial_shapes [L, 2] (h, w); level_start [L];
//   sampling_locations [bs, nq, heads, L, P, 2] in [0, 1]; attention_weights [bs, nq, heads, L, P]
//   -> output [bs, nq, heads * dim].
// One warp-lane per channel: the `dim` channels of a (query, head) are contiguous in `value`, so the
// four bilinear taps are coalesced row reads; backward reduces the location / weight gradients over
// the channels with shuffles and accumulates grad_value with float atomics (as mmcv does).
#include "ms_deform_attn.cuh"

namespace dbev {

namespace {

struct MsdaDims {
  int bs, nk, heads, dim, nq, L, P;
};

__global__ void msda_fwd_kernel(const float* __restrict__ value, const long long* __restrict__ shapes,
                                const long long* __restrict__ starts, const float* __restrict__ loc,
                                const float* __restrict__ attn, MsdaDims d, float* __restrict__ out) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)d.bs * d.nq * d.heads * d.dim;
  if (t >= total) return;
  const int c = (int)(t % d.dim);
  const int h = (int)((t / d.dim) % d.heads);
  const long long bq = t / ((long long)d.dim * d.heads);      // b * nq + q
  const int b = (int)(bq / d.nq);
  const long long lw = (bq * d.heads + h) * d.L * d.P;        // attention weights of this (b, q, h)
  const float* vb = value + (long long)b * d.nk * d.heads * d.dim + (long long)h * d.dim + c;
  const long long vstride = (long long)d.heads * d.dim;
  float acc = 0.f;
  for (int l = 0; l < d.L; ++l) {
    const int H = (int)shapes[2 * l], W = (int)shapes[2 * l + 1];
    const float* vl = vb + starts[l] * vstride;
    for (int p = 0; p < d.P; ++p) {
      const float lx = loc[(lw + l * d.P + p) * 2], ly = loc[(lw + l * d.P + p) * 2 + 1];
      const float w = attn[lw + l * d.P + p];
      const float x = lx * W - 0.5f, y = ly * H - 0.5f;          // grid_sample, align_corners=False
      if (y > -1.f && x > -1.f && y < H && x < W) {
        const int x0 = (int)floorf(x), y0 = (int)floorf(y);
        const float ax = x - x0, ay = y - y0;
        float v = 0.f;
        if (y0 >= 0 && x0 >= 0) v += (1.f - ay) * (1.f - ax) * vl[((long long)y0 * W + x0) * vstride];
        if (y0 >= 0 && x0 + 1 < W) v += (1.f - ay) * ax * vl[((long long)y0 * W + x0 + 1) * vstride];
        if (y0 + 1 < H && x0 >= 0) v += ay * (1.f - ax) * vl[((long long)(y0 + 1) * W + x0) * vstride];
        if (y0 + 1 < H && x0 + 1 < W) v += ay * ax * vl[((long long)(y0 + 1) * W + x0 + 1) * vstride];
        acc += w * v;
      }
    }
  }
  out[t] = acc;
}

// one warp per (b, q, head); lanes stride over the channels
__global__ void __launch_bounds__(256)
msda_bwd_kernel(const float* __restrict__ value, const long long* __restrict__ shapes,
                const long long* __restrict__ starts, const float* __restrict__ loc,
                const float* __restrict__ attn, const float* __restrict__ gout, MsdaDims d,
                float* __restrict__ gvalue, float* __restrict__ gloc, float* __restrict__ gattn) {
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const long long total = (long long)d.bs * d.nq * d.heads;
  if (warp >= total) return;
  const int h = (int)(warp % d.heads);
  const long long bq = warp / d.heads;
  const int b = (int)(bq / d.nq);
  const long long lw = warp * d.L * d.P;
  const long long vstride = (long long)d.heads * d.dim;
  const long long voff = (long long)b * d.nk * vstride + (long long)h * d.dim;
  const float* go = gout + warp * d.dim;
  for (int l = 0; l < d.L; ++l) {
    const int H = (int)shapes[2 * l], W = (int)shapes[2 * l + 1];
    const long long lbase = voff + starts[l] * vstride;
    for (int p = 0; p < d.P; ++p) {
      const long long i = lw + l * d.P + p;
      const float lx = loc[i * 2], ly = loc[i * 2 + 1], w = attn[i];
      const float x = lx * W - 0.5f, y = ly * H - 0.5f;
      float g_w = 0.f, g_x = 0.f, g_y = 0.f;
      if (y > -1.f && x > -1.f && y < H && x < W) {
        const int x0 = (int)floorf(x), y0 = (int)floorf(y);
        const float ax = x - x0, ay = y - y0;
        const bool v00 = y0 >= 0 && x0 >= 0, v01 = y0 >= 0 && x0 + 1 < W;
        const bool v10 = y0 + 1 < H && x0 >= 0, v11 = y0 + 1 < H && x0 + 1 < W;
        const long long o00 = lbase + ((long long)y0 * W + x0) * vstride;
        const long long o01 = o00 + vstride, o10 = o00 + (long long)W * vstride, o11 = o10 + vstride;
        for (int c = lane; c < d.dim; c += 32) {
          const float g = go[c];
          const float a00 = v00 ? value[o00 + c] : 0.f, a01 = v01 ? value[o01 + c] : 0.f;
          const float a10 = v10 ? value[o10 + c] : 0.f, a11 = v11 ? value[o11 + c] : 0.f;
          g_w += g * ((1.f - ay) * ((1.f - ax) * a00 + ax * a01) + ay * ((1.f - ax) * a10 + ax * a11));
          g_x += g * w * ((1.f - ay) * (a01 - a00) + ay * (a11 - a10));
          g_y += g * w * ((1.f - ax) * (a10 - a00) + ax * (a11 - a01));
          const float gw = g * w;
          if (v00) atomicAdd(gvalue + o00 + c, gw * (1.f - ay) * (1.f - ax));
          if (v01) atomicAdd(gvalue + o01 + c, gw * (1.f - ay) * ax);
          if (v10) atomicAdd(gvalue + o10 + c, gw * ay * (1.f - ax));
          if (v11) atomicAdd(gvalue + o11 + c, gw * ay * ax);
        }
      }
      g_w = warp_sum(g_w);
      g_x = warp_sum(g_x);
      g_y = warp_sum(g_y);
      if (lane == 0) {
        gattn[i] = g_w;
        gloc[i * 2] = g_x * W;       // d x / d loc_x = W
        gloc[i * 2 + 1] = g_y * H;
      }
    }
  }
}

}  // namespace

int ms_deform_attn_forward(const float* value, const long long* spatial_shapes,
                           const long long* level_start, const float* sampling_loc,
                           const float* attn_weight, int bs, int num_keys, int heads, int dim,
                           int num_queries, int levels, int points, float* out, cudaStream_t stream) {
  DBEV_CHECK_ARG(bs > 0 && num_keys > 0 && heads > 0 && dim > 0 && num_queries > 0 && levels > 0 && points > 0,
                 "ms_deform_attn: bad sizes");
  MsdaDims d{bs, num_keys, heads, dim, num_queries, levels, points};
  const long long total = (long long)bs * num_queries * heads * dim;
  msda_fwd_kernel<<<ceil_div(total, 256), 256, 0, stream>>>(value, spatial_shapes, level_start, sampling_loc,
                                                           attn_weight, d, out);
  DBEV_CHECK_LAUNCH("msda_fwd_kernel");
  return DBEV_OK;
}

int ms_deform_attn_backward(const float* value, const long long* spatial_shapes,
                            const long long* level_start, const float* sampling_loc,
                            const float* attn_weight, const float* grad_out, int bs, int num_keys,
                            int heads, int dim, int num_queries, int levels, int points,
                            float* grad_value, float* grad_loc, float* grad_attn, cudaStream_t stream) {
  DBEV_CHECK_ARG(bs > 0 && num_keys > 0 && heads > 0 && dim > 0 && num_queries > 0 && levels > 0 && points > 0,
                 "ms_deform_attn: bad sizes");
  MsdaDims d{bs, num_keys, heads, dim, num_queries, levels, points};
  DBEV_CUDA(cudaMemsetAsync(grad_value, 0, (size_t)bs * num_keys * heads * dim * sizeof(float), stream));
  const long long warps = (long long)bs * num_queries * heads;
  msda_bwd_kernel<<<ceil_div(warps * 32, 256), 256, 0, stream>>>(value, spatial_shapes, level_start, sampling_loc,
                                                                attn_weight, grad_out, d, grad_value, grad_loc,
                                                                grad_attn);
  DBEV_CHECK_LAUNCH("msda_bwd_kernel");
  return DBEV_OK;
}

}  // namespace dbev
