// Memory-bound companions of the tcgen05 conv kernels for TRAINING the student's BEV encoder
// (SURVEY.md §8 row S1), NHWC fp32 throughout:
//   BasicBlock  conv-BN-ReLU-conv-BN (+identity / downsample) -ReLU   mmdet3d/models/bricks/res_block.py:70-99
//   FPN_LSS     bilinear x4 (align_corners) | concat -> 2 x (conv3x3-BN-ReLU) -> bilinear x2 -> conv3x3-BN-ReLU -> conv1x1
//               mmdet3d/models/necks/lss_fpn.py:62-72
// The reference runs nn.BatchNorm2d in training mode (batch statistics, running-stat update), nn.ReLU,
// nn.Upsample and torch.cat as separate cuDNN / ATen kernels with autograd. Here:
//   channel_stats          per-channel sum / sum of squares of a [P, C] matrix (two-level, fixed order, fp64 combine),
//                          finalised by a second small kernel into BatchNorm's affine (a = gamma * invstd, b = beta - mean * a),
//                          the saved mean / invstd and the running-stat update (momentum, unbiased variance);
//                          also used for conv-bias gradients (column sums)
//   bn_act                 z = relu?(a * y + b (+ residual)), optionally written into a channel slice of a wider tensor
//   bn_bwd_reduce / apply  g = dz * (z > 0); dgamma = sum g * yhat, dbeta = sum g;
//                          dy = a * (g - mean(g) - yhat * mean(g * yhat)); g is also the identity-branch gradient
//   upsample_bilinear      align_corners=True forward (same index arithmetic as ATen's upsample_bilinear2d) and a
//                          gather-form backward (each input pixel sums its contributing outputs: no atomics)
//   pack_conv_weights      torch [C_out][C_in][KH][KW] -> the K-major matrices the conv kernels read by TMA:
//                          forward [C_out][tap][C_in], input gradient [C_in][flipped tap][C_out], and the four
//                          output-parity matrices of a stride-2 input gradient
// All are HBM-bound: algorithmic bytes = one read of each input + one write of each output.
#include "bev_encoder_ops.cuh"

namespace dbev {

namespace {

constexpr int kStatThreads = 256;
constexpr int kPackJobWords = 8;
constexpr int kFinalWarps = 32;   // channel_stats_final_kernel: 32 warps x 14 loads in flight cover the 444 partial rows in one round

// ------------------------------------------------------------------------------------------ column statistics
// partial[blk][2][C]: block blk sums rows [blk * rows_per_block, ...); MODE 0: (sum y, sum y^2);
// MODE 1: (sum g, sum g * yhat) with g = dz * (z > 0 or 1), yhat = (y - mean) * invstd
struct StatArgs {
  const float* y; int y_ld;
  const float* dz; int dz_ld;       // MODE 1
  const float* z; int z_ld;         // MODE 1, may be null (no ReLU after the BatchNorm)
  const unsigned char* mask;        // MODE 1, alternative to z: the forward's ReLU mask, one byte per channel quad
  const float* mean_invstd;         // MODE 1: [2][C]
  long long rows; int C; int rows_per_block;
  float* partial;
  // finalisation
  const float* gamma; const float* beta; float eps; float momentum;
  float* running_mean; float* running_var;
  float* out;                       // MODE 0: [4][C] = a, b, mean, invstd; MODE 1: [4][C] = dgamma, dbeta, mean g, mean g*yhat
  int plain_sum;                    // MODE 0 only: out[0][C] = column sums (bias gradient), nothing else
  int accumulate;                   // plain_sum: add to out
};

template <int MODE>
__device__ __forceinline__ void stat_accum(const StatArgs& a, long long r, int q, const float4& mu, const float4& is, float4& s1,
                                           float4& s2) {
  const float4 v = ld_stream_f4(a.y + r * a.y_ld + 4 * q);
  if (MODE == 0) {
    s1.x += v.x, s1.y += v.y, s1.z += v.z, s1.w += v.w;
    s2.x += v.x * v.x, s2.y += v.y * v.y, s2.z += v.z * v.z, s2.w += v.w * v.w;
  } else {
    float4 g = ld_stream_f4(a.dz + r * a.dz_ld + 4 * q);
    if (a.mask) {
      const unsigned m = __ldg(a.mask + r * (a.C >> 2) + q);
      g.x = (m & 1u) ? g.x : 0.f, g.y = (m & 2u) ? g.y : 0.f, g.z = (m & 4u) ? g.z : 0.f, g.w = (m & 8u) ? g.w : 0.f;
    } else if (a.z) {
      const float4 zz = ld_stream_f4(a.z + r * a.z_ld + 4 * q);
      g.x = zz.x > 0.f ? g.x : 0.f, g.y = zz.y > 0.f ? g.y : 0.f, g.z = zz.z > 0.f ? g.z : 0.f, g.w = zz.w > 0.f ? g.w : 0.f;
    }
    s1.x += g.x, s1.y += g.y, s1.z += g.z, s1.w += g.w;
    s2.x += g.x * ((v.x - mu.x) * is.x), s2.y += g.y * ((v.y - mu.y) * is.y);
    s2.z += g.z * ((v.z - mu.z) * is.z), s2.w += g.w * ((v.w - mu.w) * is.w);
  }
}

template <int MODE>
__global__ void __launch_bounds__(kStatThreads) channel_stats_kernel(StatArgs a) {
  const int quads = a.C >> 2;
  const int rgroups = blockDim.x / quads;
  const int q = threadIdx.x % quads, rg = threadIdx.x / quads;
  __shared__ float4 red[kStatThreads];      // 4 KB, used for both statistics in turn (small enough to be co-resident
                                            // with a persistent conv CTA of another stream, conv2d_tc.cu smem_reserve())
  float4 s1 = make_float4(0.f, 0.f, 0.f, 0.f), s2 = s1, t1 = s1, t2 = s1;
  const long long r0 = (long long)blockIdx.x * a.rows_per_block;
  long long r1 = r0 + a.rows_per_block;
  if (r1 > a.rows) r1 = a.rows;
  float4 mu = s1, is = s1;
  if (MODE == 1) {
    mu = *reinterpret_cast<const float4*>(a.mean_invstd + 4 * q);
    is = *reinterpret_cast<const float4*>(a.mean_invstd + a.C + 4 * q);
  }
  // four independent accumulator pairs: four rows' loads in flight per thread (3 blocks x 256 threads per SM -> ~50 KB
  // in flight per SM, what the HBM latency needs; with two it ran at 0.41 of the HBM peak). Fixed order -> deterministic.
  float4 u1 = t1, u2 = t1, v1 = t1, v2 = t1;
  long long r = r0 + rg;
  for (; r + 3 * rgroups < r1; r += 4 * rgroups) {
    stat_accum<MODE>(a, r, q, mu, is, s1, s2);
    stat_accum<MODE>(a, r + rgroups, q, mu, is, t1, t2);
    stat_accum<MODE>(a, r + 2 * rgroups, q, mu, is, u1, u2);
    stat_accum<MODE>(a, r + 3 * rgroups, q, mu, is, v1, v2);
  }
  for (; r < r1; r += rgroups) stat_accum<MODE>(a, r, q, mu, is, s1, s2);
  s1.x += t1.x, s1.y += t1.y, s1.z += t1.z, s1.w += t1.w;
  s2.x += t2.x, s2.y += t2.y, s2.z += t2.z, s2.w += t2.w;
  u1.x += v1.x, u1.y += v1.y, u1.z += v1.z, u1.w += v1.w;
  u2.x += v2.x, u2.y += v2.y, u2.z += v2.z, u2.w += v2.w;
  s1.x += u1.x, s1.y += u1.y, s1.z += u1.z, s1.w += u1.w;
  s2.x += u2.x, s2.y += u2.y, s2.z += u2.z, s2.w += u2.w;
  red[threadIdx.x] = s1;
  __syncthreads();
  if (rg == 0)
    for (int g = 1; g < rgroups; ++g) {
      const float4 u1 = red[g * quads + q];
      s1.x += u1.x, s1.y += u1.y, s1.z += u1.z, s1.w += u1.w;
    }
  __syncthreads();
  red[threadIdx.x] = s2;
  __syncthreads();
  if (rg == 0) {
    for (int g = 1; g < rgroups; ++g) {
      const float4 u2 = red[g * quads + q];
      s2.x += u2.x, s2.y += u2.y, s2.z += u2.z, s2.w += u2.w;
    }
    float* p = a.partial + (long long)blockIdx.x * 2 * a.C;
    *reinterpret_cast<float4*>(p + 4 * q) = s1;
    *reinterpret_cast<float4*>(p + a.C + 4 * q) = s2;
  }
}

// Combines the per-block partial sums in block order with fp64 accumulation (deterministic) and finalises:
// a block owns 16 channels; lanes 0-15 / 16-31 of every warp carry the first / second statistic, the kFinalWarps warps
// each sum their share of the partial rows with all loads in flight at once (one round of load latency, not a chain)
template <int MODE>
__global__ void __launch_bounds__(kFinalWarps * 32) channel_stats_final_kernel(StatArgs a, int n_blocks) {
  __shared__ double part[kFinalWarps][32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int stat = lane >> 4;
  const int c = blockIdx.x * 16 + (lane & 15);
  double t = 0.0;
  if (c < a.C) {
    const int b0 = (int)((long long)n_blocks * warp / kFinalWarps), b1 = (int)((long long)n_blocks * (warp + 1) / kFinalWarps);
    const float* p = a.partial + stat * a.C + c;
    const long long st = 2LL * a.C;
    int b = b0;
    for (; b < b1; b += 14) {                 // all of a warp's rows in flight at once (predicated), summed in row order
      float v[14];
#pragma unroll
      for (int k = 0; k < 14; ++k) v[k] = b + k < b1 ? __ldcg(p + (long long)(b + k) * st) : 0.f;
#pragma unroll
      for (int k = 0; k < 14; ++k) t += (double)v[k];
    }
  }
  part[warp][lane] = t;
  __syncthreads();
  if (warp != 0) return;
  double tot = 0.0;
#pragma unroll
  for (int k = 0; k < kFinalWarps; ++k) tot += part[k][lane];
  // lane l (< 16) needs the second statistic of its channel from lane l + 16
  const double other = __shfl_down_sync(0xffffffffu, tot, 16);
  if (lane >= 16 || c >= a.C) return;
  const double t1 = tot, t2 = other;
  const double n = (double)a.rows;
  if (MODE == 0) {
    if (a.plain_sum) {
      a.out[c] = (a.accumulate ? a.out[c] : 0.f) + (float)t1;
    } else {
      const double mean = t1 / n;
      double var = t2 / n - mean * mean;      // biased, as BatchNorm normalises with
      if (var < 0.0) var = 0.0;
      const float invstd = (float)(1.0 / sqrt(var + (double)a.eps));
      const float ga = a.gamma ? a.gamma[c] : 1.f, be = a.beta ? a.beta[c] : 0.f;
      const float sc = ga * invstd;
      a.out[c] = sc;
      a.out[a.C + c] = be - (float)mean * sc;
      a.out[2 * a.C + c] = (float)mean;
      a.out[3 * a.C + c] = invstd;
      if (a.running_mean) {                    // torch: running = (1 - m) * running + m * batch (unbiased variance)
        const double unb = n > 1.0 ? var * n / (n - 1.0) : var;
        a.running_mean[c] = (1.f - a.momentum) * a.running_mean[c] + a.momentum * (float)mean;
        a.running_var[c] = (1.f - a.momentum) * a.running_var[c] + a.momentum * (float)unb;
      }
    }
  } else {
    a.out[c] = (float)t2;                      // dgamma
    a.out[a.C + c] = (float)t1;                // dbeta
    a.out[2 * a.C + c] = (float)(t1 / n);
    a.out[3 * a.C + c] = (float)(t2 / n);
  }
}

// ------------------------------------------------------------------------------------------ elementwise passes
__global__ void bn_act_kernel(const float* __restrict__ y, int y_ld, const float* __restrict__ ab, const float* __restrict__ res,
                              int res_ld, long long rows, int C, int relu, float* __restrict__ out, int out_ld,
                              unsigned char* __restrict__ mask) {
  const int quads = C >> 2;
  const long long total = rows * quads;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / quads;
    const int q = (int)(i - r * quads);
    float4 v = ld_stream_f4(y + r * y_ld + 4 * q);
    if (ab) {
      const float4 sc = __ldg(reinterpret_cast<const float4*>(ab + 4 * q));
      const float4 sh = __ldg(reinterpret_cast<const float4*>(ab + C + 4 * q));
      v.x = v.x * sc.x + sh.x, v.y = v.y * sc.y + sh.y, v.z = v.z * sc.z + sh.z, v.w = v.w * sc.w + sh.w;
    }
    if (res) {
      const float4 t = ld_stream_f4(res + r * res_ld + 4 * q);
      v.x += t.x, v.y += t.y, v.z += t.z, v.w += t.w;
    }
    if (relu) {
      // the backward's ReLU mask as one byte per quad: 1/16 of re-reading the output (out > 0 <=> pre-activation > 0)
      if (mask) mask[i] = (unsigned char)((v.x > 0.f ? 1 : 0) | (v.y > 0.f ? 2 : 0) | (v.z > 0.f ? 4 : 0) | (v.w > 0.f ? 8 : 0));
      v.x = fmaxf(v.x, 0.f), v.y = fmaxf(v.y, 0.f), v.z = fmaxf(v.z, 0.f), v.w = fmaxf(v.w, 0.f);
    }
    *reinterpret_cast<float4*>(out + r * out_ld + 4 * q) = v;
  }
}

// dy = a * (g - m1 - yhat * m2), g = dz * (z > 0); optional g_out (identity-branch gradient): store or add
__global__ void bn_bwd_apply_kernel(const float* __restrict__ dz, int dz_ld, const float* __restrict__ z, int z_ld,
                                    const float* __restrict__ y, int y_ld, const float* __restrict__ fwd /* a,b,mean,invstd */,
                                    const float* __restrict__ bwd /* dgamma,dbeta,m1,m2 */, long long rows, int C,
                                    float* __restrict__ dy, int dy_ld, float* __restrict__ g_out, int g_ld, int g_accumulate,
                                    const unsigned char* __restrict__ mask) {
  const int quads = C >> 2;
  const long long total = rows * quads;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / quads;
    const int q = (int)(i - r * quads);
    float4 g = ld_stream_f4(dz + r * dz_ld + 4 * q);
    if (mask) {
      const unsigned m = __ldg(mask + i);
      g.x = (m & 1u) ? g.x : 0.f, g.y = (m & 2u) ? g.y : 0.f, g.z = (m & 4u) ? g.z : 0.f, g.w = (m & 8u) ? g.w : 0.f;
    } else if (z) {
      const float4 zz = ld_stream_f4(z + r * z_ld + 4 * q);
      g.x = zz.x > 0.f ? g.x : 0.f, g.y = zz.y > 0.f ? g.y : 0.f, g.z = zz.z > 0.f ? g.z : 0.f, g.w = zz.w > 0.f ? g.w : 0.f;
    }
    if (g_out) {
      float4* gp = reinterpret_cast<float4*>(g_out + r * g_ld + 4 * q);
      float4 o = g;
      if (g_accumulate) {
        const float4 p = *gp;
        o.x += p.x, o.y += p.y, o.z += p.z, o.w += p.w;
      }
      *gp = o;
    }
    if (dy) {
      const float4 v = ld_stream_f4(y + r * y_ld + 4 * q);
      const float4 sc = __ldg(reinterpret_cast<const float4*>(fwd + 4 * q));
      const float4 mu = __ldg(reinterpret_cast<const float4*>(fwd + 2 * C + 4 * q));
      const float4 is = __ldg(reinterpret_cast<const float4*>(fwd + 3 * C + 4 * q));
      const float4 m1 = __ldg(reinterpret_cast<const float4*>(bwd + 2 * C + 4 * q));
      const float4 m2 = __ldg(reinterpret_cast<const float4*>(bwd + 3 * C + 4 * q));
      float4 o;
      o.x = sc.x * (g.x - m1.x - (v.x - mu.x) * is.x * m2.x);
      o.y = sc.y * (g.y - m1.y - (v.y - mu.y) * is.y * m2.y);
      o.z = sc.z * (g.z - m1.z - (v.z - mu.z) * is.z * m2.z);
      o.w = sc.w * (g.w - m1.w - (v.w - mu.w) * is.w * m2.w);
      *reinterpret_cast<float4*>(dy + r * dy_ld + 4 * q) = o;
    }
  }
}

// ------------------------------------------------------------------------------------------ bilinear upsampling
// ATen upsample_bilinear2d, align_corners=True: src = dst * (in - 1) / (out - 1) in fp32, i0 = (int)src,
// lambda1 = src - i0, i1 = i0 + (i0 < in - 1)
__device__ __forceinline__ void bilin_src(int dst, float scale, int in_size, int& i0, int& i1, float& l0, float& l1) {
  const float src = scale * (float)dst;
  i0 = (int)src;
  if (i0 > in_size - 1) i0 = in_size - 1;
  i1 = i0 + (i0 < in_size - 1 ? 1 : 0);
  l1 = src - (float)i0;
  l0 = 1.f - l1;
}

// grid.y = (image, group of kUpRows output rows); a thread = (output column, channel quad), quads fastest: row weights
// are block-uniform. A block walks its rows one after the other: consecutive output rows read the same two input rows,
// which then come from L1 (a block's column range of an input row is a few KB) instead of L2.
__global__ void upsample_bilinear_fwd_kernel(const float* __restrict__ in, int in_ld, int n_img, int h, int w, int C, int H, int W,
                                             float sy, float sx, float* __restrict__ out, int out_ld, int kUpRows) {
  const int quads = C >> 2;
  const int groups = (H + kUpRows - 1) / kUpRows;
  const int n = blockIdx.y / groups, oy_begin = (blockIdx.y - n * groups) * kUpRows;
  const int oy_end = oy_begin + kUpRows < H ? oy_begin + kUpRows : H;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < W * quads; i += gridDim.x * blockDim.x) {
    const int ox = i / quads, q = i - ox * quads;
    int x0, x1;
    float lx0, lx1;
    bilin_src(ox, sx, w, x0, x1, lx0, lx1);
    for (int oy = oy_begin; oy < oy_end; ++oy) {
      int y0, y1;
      float ly0, ly1;
      bilin_src(oy, sy, h, y0, y1, ly0, ly1);
      const float* r0 = in + ((long long)n * h + y0) * w * in_ld;
      const float* r1 = in + ((long long)n * h + y1) * w * in_ld;
      const float4 v00 = __ldg(reinterpret_cast<const float4*>(r0 + (long long)x0 * in_ld + 4 * q));
      const float4 v01 = __ldg(reinterpret_cast<const float4*>(r0 + (long long)x1 * in_ld + 4 * q));
      const float4 v10 = __ldg(reinterpret_cast<const float4*>(r1 + (long long)x0 * in_ld + 4 * q));
      const float4 v11 = __ldg(reinterpret_cast<const float4*>(r1 + (long long)x1 * in_ld + 4 * q));
      float4 o;
      o.x = ly0 * (lx0 * v00.x + lx1 * v01.x) + ly1 * (lx0 * v10.x + lx1 * v11.x);
      o.y = ly0 * (lx0 * v00.y + lx1 * v01.y) + ly1 * (lx0 * v10.y + lx1 * v11.y);
      o.z = ly0 * (lx0 * v00.z + lx1 * v01.z) + ly1 * (lx0 * v10.z + lx1 * v11.z);
      o.w = ly0 * (lx0 * v00.w + lx1 * v01.w) + ly1 * (lx0 * v10.w + lx1 * v11.w);
      st_stream_f4(out + (((long long)n * H + oy) * W + ox) * out_ld + 4 * q, o);
    }
  }
}

// outputs whose (i0, i1) can contain input index i: src = s * o in (i - 1, i + 1); weights recomputed exactly as the
// forward does. At most kMaxTaps candidates per axis (scale factors >= 1: 2 * ceil(1 / s) + 2).
constexpr int kMaxTaps = 12;
__device__ __forceinline__ int bilin_taps(int i, float s, float rs, int in_size, int out_size, int* o_idx, float* o_w) {
  int lo = (int)floorf((float)(i - 1) * rs) - 1, hi = (int)ceilf((float)(i + 1) * rs) + 1;
  if (s == 0.f) lo = 0, hi = out_size - 1;
  lo = lo < 0 ? 0 : lo, hi = hi > out_size - 1 ? out_size - 1 : hi;
  int n = 0;
  for (int o = lo; o <= hi; ++o) {
    int i0, i1;
    float l0, l1;
    bilin_src(o, s, in_size, i0, i1, l0, l1);
    const float wgt = (i0 == i ? l0 : 0.f) + (i1 == i ? l1 : 0.f);
    if (wgt != 0.f && n < kMaxTaps) o_idx[n] = o, o_w[n] = wgt, ++n;
  }
  return n;
}

// gather form of the transpose: input pixel (iy, ix) sums the output gradients that read it. grid.y = (image, group of
// kUpRows input rows); a thread = (input column, channel quad) and walks the group's rows: neighbouring input rows read
// overlapping output rows, which then hit L1 instead of L2 (the kernel was L2-bound: every output element is read by up
// to four input pixels)
__global__ void upsample_bilinear_bwd_kernel(const float* __restrict__ dout, int dout_ld, int n_img, int h, int w, int C, int H,
                                             int W, float sy, float sx, float* __restrict__ din, int din_ld, int accumulate,
                                             int kUpRows) {
  const int quads = C >> 2;
  const int groups = (h + kUpRows - 1) / kUpRows;
  const int n = blockIdx.y / groups, iy_begin = (blockIdx.y - n * groups) * kUpRows;
  const int iy_end = iy_begin + kUpRows < h ? iy_begin + kUpRows : h;
  const float ry = sy > 0.f ? 1.f / sy : 0.f, rx = sx > 0.f ? 1.f / sx : 0.f;
  int oys[kMaxTaps], oxs[kMaxTaps];
  float wys[kMaxTaps], wxs[kMaxTaps];
  const float* b = dout + (long long)n * H * W * dout_ld;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < w * quads; i += gridDim.x * blockDim.x) {
    const int ix = i / quads, q = i - ix * quads;
    const int nx = bilin_taps(ix, sx, rx, w, W, oxs, wxs);
    for (int iy = iy_begin; iy < iy_end; ++iy) {
      const int ny = bilin_taps(iy, sy, ry, h, H, oys, wys);
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int a = 0; a < ny; ++a) {
        const float* row = b + (long long)oys[a] * W * dout_ld + 4 * q;
        for (int c = 0; c < nx; ++c) {
          const float4 g = __ldg(reinterpret_cast<const float4*>(row + (long long)oxs[c] * dout_ld));
          const float wgt = wys[a] * wxs[c];
          acc.x += wgt * g.x, acc.y += wgt * g.y, acc.z += wgt * g.z, acc.w += wgt * g.w;
        }
      }
      float4* o = reinterpret_cast<float4*>(din + (((long long)n * h + iy) * w + ix) * din_ld + 4 * q);
      if (accumulate) {
        const float4 prev = *o;
        acc.x += prev.x, acc.y += prev.y, acc.z += prev.z, acc.w += prev.w;
      }
      *o = acc;
    }
  }
}

// ------------------------------------------------------------------------------------------ weight packing
// mode 0: fwd  [co][(ky*KW + kx)*CI + ci]
// mode 1: dgrad (stride 1)  [ci][((KH-1-ky)*KW + (KW-1-kx))*CO + co]
// mode 2: dgrad (stride 2, 3x3, pad 1): four matrices, class (a, b) = parity of the input row / column:
//         [ci][(ty*nkx + tx)*CO + co] with ky = a + 1 - 2*ty, kx = b + 1 - 2*tx (nky = 1 + a, nkx = 1 + b);
//         class order (0,0), (0,1), (1,0), (1,1) at float offsets 0, 1, 3, 5 (x CI*CO)
__global__ void pack_conv_weights_kernel(const float* __restrict__ w, int CO, int CI, int KH, int KW, int mode, float* __restrict__ out) {
  const long long total = (long long)CO * CI * KH * KW;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    // i indexes the OUTPUT (coalesced writes); decode per mode
    if (mode == 0) {
      const int ci = (int)(i % CI);
      long long t = i / CI;
      const int tap = (int)(t % (KH * KW)), co = (int)(t / (KH * KW));
      out[i] = w[((long long)co * CI + ci) * KH * KW + tap];
    } else if (mode == 1) {
      const int co = (int)(i % CO);
      long long t = i / CO;
      const int tap = (int)(t % (KH * KW)), ci = (int)(t / (KH * KW));
      const int ky = KH - 1 - tap / KW, kx = KW - 1 - tap % KW;
      out[i] = w[((long long)co * CI + ci) * KH * KW + ky * KW + kx];
    } else {
      const long long per = (long long)CI * CO;
      const long long cls_off[5] = {0, per, 3 * per, 5 * per, 9 * per};
      int cls = 0;
      while (i >= cls_off[cls + 1]) ++cls;
      const int a = cls >> 1, b = cls & 1, nkx = 1 + b, ntap = (1 + a) * nkx;
      const long long j = i - cls_off[cls];
      const int co = (int)(j % CO);
      long long t = j / CO;
      const int tap = (int)(t % ntap), ci = (int)(t / ntap);
      const int ty = tap / nkx, tx = tap % nkx;
      const int ky = a + 1 - 2 * ty, kx = b + 1 - 2 * tx;
      out[i] = w[((long long)co * CI + ci) * 9 + ky * 3 + kx];
    }
  }
}

// Both matrices of a layer in one pass (what the training step uses): a block owns a 32 (C_out) x 32 (C_in) tile of
// the filter, staged in shared memory so that the global reads (taps fastest) and both sets of global writes
// (C_in fastest / C_out fastest) are 128-byte runs. dgrad_mode 1: stride-1 matrix, 2: the four parity matrices.
__device__ __forceinline__ void pack_conv_weights_tile(const float* __restrict__ w, int CO, int CI, int KH, int KW,
                                                       int dgrad_mode, float* __restrict__ out_f,
                                                       float* __restrict__ out_d, int co0, int ci0, float* tile) {
  const int taps = KH * KW;
  const int pitch = 32 * taps + 1;                    // tile = [32 co][pitch]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // flat, coalesced copy of the 32 rows (each 32 * taps contiguous floats); loads issued in batches of 12 so that
  // the (cold) global latency is paid three times, not 36 times
  const int row_len = 32 * taps, total = 32 * row_len;
  for (int base_i = threadIdx.x; base_i < total; base_i += 256 * 12) {
    float v[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) {
      const int i = base_i + k * 256;
      v[k] = 0.f;
      if (i < total) {
        const int co = i / row_len, j = i - co * row_len;
        if (co0 + co < CO && ci0 + j / taps < CI) v[k] = __ldg(w + ((long long)(co0 + co) * CI + ci0) * taps + j);
      }
    }
#pragma unroll
    for (int k = 0; k < 12; ++k) {
      const int i = base_i + k * 256;
      if (i < total) {
        const int co = i / row_len, j = i - co * row_len;
        tile[co * pitch + j] = v[k];
      }
    }
  }
  __syncthreads();
  if (out_f) {                                        // [co][tap][ci]: lane = ci
    for (int p = warp; p < 32 * taps; p += 8) {
      const int co = p / taps, tap = p - co * taps;
      if (co0 + co < CO && ci0 + lane < CI)
        out_f[((long long)(co0 + co) * taps + tap) * CI + ci0 + lane] = tile[co * pitch + lane * taps + tap];
    }
  }
  if (out_d) {                                        // [ci][tap'][co]: lane = co
    const long long per = (long long)CI * CO;
    for (int p = warp; p < 32 * taps; p += 8) {
      const int ci = p / taps, tap = p - ci * taps;
      const int ky = tap / KW, kx = tap - ky * KW;
      if (co0 + lane >= CO || ci0 + ci >= CI) continue;
      const float v = tile[lane * pitch + ci * taps + tap];
      if (dgrad_mode == 1) {
        const int ft = (KH - 1 - ky) * KW + (KW - 1 - kx);
        out_d[((long long)(ci0 + ci) * taps + ft) * CO + co0 + lane] = v;
      } else {
        const int a = ky == 1 ? 0 : 1, ty = ky == 0 ? 1 : 0, b = kx == 1 ? 0 : 1, tx = kx == 0 ? 1 : 0;
        const int cls = a * 2 + b, nkx = 1 + b, ntap = (1 + a) * nkx;
        const long long off = (cls == 0 ? 0 : (cls == 1 ? 1 : (cls == 2 ? 3 : 5))) * per;
        out_d[off + ((long long)(ci0 + ci) * ntap + ty * nkx + tx) * CO + co0 + lane] = v;
      }
    }
  }
}

__global__ void __launch_bounds__(256) pack_conv_weights_tile_kernel(const float* __restrict__ w, int CO, int CI, int KH, int KW,
                                                                     int dgrad_mode, float* __restrict__ out_f,
                                                                     float* __restrict__ out_d) {
  extern __shared__ float tile[];
  pack_conv_weights_tile(w, CO, CI, KH, KW, dgrad_mode, out_f, out_d, blockIdx.y * 32, blockIdx.x * 32, tile);
}

// every layer of a network in ONE launch: jobs = n_jobs records of kPackJobWords int64 {w, out_fwd, out_dgrad, C_out, C_in,
// KH * 256 + KW, dgrad_mode, first tile}; block b works on tile b - first_tile of the job that owns it
__global__ void __launch_bounds__(256) pack_conv_weights_batch_kernel(const long long* __restrict__ jobs, int n_jobs) {
  extern __shared__ float tile[];
  int j = 0;
  while (j + 1 < n_jobs && (long long)blockIdx.x >= __ldg(jobs + (j + 1) * kPackJobWords + 7)) ++j;
  const long long* r = jobs + j * kPackJobWords;
  const int CO = (int)__ldg(r + 3), CI = (int)__ldg(r + 4), khkw = (int)__ldg(r + 5);
  const int t = (int)(blockIdx.x - __ldg(r + 7)), tiles_ci = (CI + 31) / 32;
  pack_conv_weights_tile(reinterpret_cast<const float*>(__ldg(r + 0)), CO, CI, khkw >> 8, khkw & 255, (int)__ldg(r + 6),
                         reinterpret_cast<float*>(__ldg(r + 1)), reinterpret_cast<float*>(__ldg(r + 2)), (t / tiles_ci) * 32,
                         (t % tiles_ci) * 32, tile);
}

// rows a block of the upsampling kernels walks: 8 when that still leaves >= 1024 blocks, else fewer (small maps)
int upsample_rows_per_block(long long blocks_x, int n_img, int rows) {
  int r = 8;
  while (r > 1 && blocks_x * n_img * ceil_div(rows, r) < 1024) r >>= 1;
  return r;
}

int grid_for(long long total, int threads) {
  long long b = (total + threads - 1) / threads;
  const long long cap = (long long)kNumSMs * 16;
  return (int)(b < cap ? (b > 0 ? b : 1) : cap);
}

int stat_blocks(long long rows, int* rows_per_block) {
  int blocks = kNumSMs * 3;     // 256-thread blocks, four row loads in flight per thread
  if (blocks > rows) blocks = (int)rows;
  int rpb = (int)((rows + blocks - 1) / blocks);
  blocks = (int)((rows + rpb - 1) / rpb);
  *rows_per_block = rpb;
  return blocks;
}

}  // namespace

size_t channel_stats_workspace_bytes(long long rows, int C) {
  int rpb;
  const int blocks = stat_blocks(rows, &rpb);
  return (size_t)blocks * 2 * C * sizeof(float);
}

static int check_stat_shape(long long rows, int C, const char* what) {
  DBEV_CHECK_ARG(rows > 0 && C >= 4 && C % 4 == 0 && C / 4 <= kStatThreads, "%s: C must be a multiple of 4, <= %d (got %d)", what,
                 4 * kStatThreads, C);
  return DBEV_OK;
}

int bn_batch_stats(const float* y, int y_ld, long long rows, int C, const float* gamma, const float* beta, float eps,
                   float momentum, float* running_mean, float* running_var, float* out4c, void* workspace,
                   size_t workspace_bytes, cudaStream_t stream) {
  if (int rc = check_stat_shape(rows, C, "bn_batch_stats")) return rc;
  DBEV_CHECK_ARG(workspace && workspace_bytes >= channel_stats_workspace_bytes(rows, C), "bn_batch_stats: workspace too small");
  DBEV_CHECK_ARG((running_mean == nullptr) == (running_var == nullptr), "bn_batch_stats: running_mean / running_var go together");
  StatArgs a = {};
  int rpb;
  const int blocks = stat_blocks(rows, &rpb);
  a.y = y, a.y_ld = y_ld, a.rows = rows, a.C = C, a.rows_per_block = rpb;
  a.partial = (float*)workspace;
  a.gamma = gamma, a.beta = beta, a.eps = eps, a.momentum = momentum, a.running_mean = running_mean, a.running_var = running_var;
  a.out = out4c;
  const int quads = C / 4, threads = quads * (kStatThreads / quads);
  channel_stats_kernel<0><<<blocks, threads, 0, stream>>>(a);
  DBEV_CHECK_LAUNCH("channel_stats_kernel<0>");
  channel_stats_final_kernel<0><<<ceil_div(C, 16), kFinalWarps * 32, 0, stream>>>(a, blocks);
  DBEV_CHECK_LAUNCH("channel_stats_final_kernel<0>");
  return DBEV_OK;
}

int channel_sums(const float* y, int y_ld, long long rows, int C, float* out, int accumulate, void* workspace,
                 size_t workspace_bytes, cudaStream_t stream) {
  if (int rc = check_stat_shape(rows, C, "channel_sums")) return rc;
  DBEV_CHECK_ARG(workspace && workspace_bytes >= channel_stats_workspace_bytes(rows, C), "channel_sums: workspace too small");
  StatArgs a = {};
  int rpb;
  const int blocks = stat_blocks(rows, &rpb);
  a.y = y, a.y_ld = y_ld, a.rows = rows, a.C = C, a.rows_per_block = rpb;
  a.partial = (float*)workspace;
  a.out = out, a.plain_sum = 1, a.accumulate = accumulate;
  const int quads = C / 4, threads = quads * (kStatThreads / quads);
  channel_stats_kernel<0><<<blocks, threads, 0, stream>>>(a);
  DBEV_CHECK_LAUNCH("channel_stats_kernel<0>");
  channel_stats_final_kernel<0><<<ceil_div(C, 16), kFinalWarps * 32, 0, stream>>>(a, blocks);
  DBEV_CHECK_LAUNCH("channel_stats_final_kernel<0>");
  return DBEV_OK;
}

int bn_act_forward(const float* y, int y_ld, const float* ab, const float* residual, int res_ld, long long rows, int C, int relu,
                   float* out, int out_ld, unsigned char* relu_mask, cudaStream_t stream) {
  DBEV_CHECK_ARG(rows > 0 && C % 4 == 0 && y_ld % 4 == 0 && out_ld % 4 == 0 && (!residual || res_ld % 4 == 0),
                 "bn_act_forward: channel counts / strides must be multiples of 4");
  bn_act_kernel<<<grid_for(rows * (C / 4), 256), 256, 0, stream>>>(y, y_ld, ab, residual, res_ld, rows, C, relu, out, out_ld,
                                                                   relu ? relu_mask : nullptr);
  DBEV_CHECK_LAUNCH("bn_act_kernel");
  return DBEV_OK;
}

int bn_backward(const float* dz, int dz_ld, const float* z, int z_ld, const float* y, int y_ld, const float* fwd4c, long long rows,
                int C, float* bwd4c, float* dy, int dy_ld, float* g_out, int g_ld, int g_accumulate, const unsigned char* relu_mask,
                void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  if (int rc = check_stat_shape(rows, C, "bn_backward")) return rc;
  DBEV_CHECK_ARG(workspace && workspace_bytes >= channel_stats_workspace_bytes(rows, C), "bn_backward: workspace too small");
  StatArgs a = {};
  int rpb;
  const int blocks = stat_blocks(rows, &rpb);
  a.y = y, a.y_ld = y_ld, a.dz = dz, a.dz_ld = dz_ld, a.z = z, a.z_ld = z_ld, a.mask = relu_mask, a.mean_invstd = fwd4c + 2 * C;
  a.rows = rows, a.C = C, a.rows_per_block = rpb;
  a.partial = (float*)workspace;
  a.out = bwd4c;
  const int quads = C / 4, threads = quads * (kStatThreads / quads);
  channel_stats_kernel<1><<<blocks, threads, 0, stream>>>(a);
  DBEV_CHECK_LAUNCH("channel_stats_kernel<1>");
  channel_stats_final_kernel<1><<<ceil_div(C, 16), kFinalWarps * 32, 0, stream>>>(a, blocks);
  DBEV_CHECK_LAUNCH("channel_stats_final_kernel<1>");
  bn_bwd_apply_kernel<<<grid_for(rows * quads, 256), 256, 0, stream>>>(dz, dz_ld, z, z_ld, y, y_ld, fwd4c, bwd4c, rows, C, dy, dy_ld,
                                                                       g_out, g_ld, g_accumulate, relu_mask);
  DBEV_CHECK_LAUNCH("bn_bwd_apply_kernel");
  return DBEV_OK;
}

int relu_mask_backward(const float* dz, int dz_ld, const float* z, int z_ld, long long rows, int C, float* g_out, int g_ld,
                       int accumulate, const unsigned char* relu_mask, cudaStream_t stream) {
  DBEV_CHECK_ARG(rows > 0 && C % 4 == 0, "relu_mask_backward: C must be a multiple of 4");
  DBEV_CHECK_ARG(z || relu_mask, "relu_mask_backward: needs the forward output z or its ReLU mask");
  bn_bwd_apply_kernel<<<grid_for(rows * (C / 4), 256), 256, 0, stream>>>(dz, dz_ld, z, z_ld, nullptr, 0, nullptr, nullptr, rows, C,
                                                                         nullptr, 0, g_out, g_ld, accumulate, relu_mask);
  DBEV_CHECK_LAUNCH("bn_bwd_apply_kernel");
  return DBEV_OK;
}

int upsample_bilinear_forward(const float* in, int in_ld, int n_img, int h, int w, int C, int H, int W, float* out, int out_ld,
                              cudaStream_t stream) {
  DBEV_CHECK_ARG(n_img > 0 && h > 0 && w > 0 && H > 0 && W > 0 && C % 4 == 0, "upsample_bilinear: bad shape");
  const float sy = H > 1 ? (float)(h - 1) / (float)(H - 1) : 0.f, sx = W > 1 ? (float)(w - 1) / (float)(W - 1) : 0.f;
  const int kUpRows = upsample_rows_per_block(ceil_div((long long)W * (C / 4), 256), n_img, H);
  DBEV_CHECK_ARG((long long)n_img * ceil_div(H, kUpRows) <= 65535, "upsample_bilinear: too many output rows for one launch");
  upsample_bilinear_fwd_kernel<<<dim3((unsigned)ceil_div((long long)W * (C / 4), 256), (unsigned)(n_img * ceil_div(H, kUpRows))), 256, 0, stream>>>(
      in, in_ld, n_img, h, w, C, H, W, sy, sx, out, out_ld, kUpRows);
  DBEV_CHECK_LAUNCH("upsample_bilinear_fwd_kernel");
  return DBEV_OK;
}

int upsample_bilinear_backward(const float* dout, int dout_ld, int n_img, int h, int w, int C, int H, int W, float* din, int din_ld,
                               int accumulate, cudaStream_t stream) {
  DBEV_CHECK_ARG(n_img > 0 && h > 0 && w > 0 && H > 0 && W > 0 && C % 4 == 0, "upsample_bilinear: bad shape");
  const float sy = H > 1 ? (float)(h - 1) / (float)(H - 1) : 0.f, sx = W > 1 ? (float)(w - 1) / (float)(W - 1) : 0.f;
  const int kUpRows = upsample_rows_per_block(ceil_div((long long)w * (C / 4), 128), n_img, h);
  DBEV_CHECK_ARG((long long)n_img * ceil_div(h, kUpRows) <= 65535, "upsample_bilinear: too many input rows for one launch");
  DBEV_CHECK_ARG(H >= h && W >= w && H <= 4 * h + 4 && W <= 4 * w + 4, "upsample_bilinear_backward: scale factors 1..4 only");
  upsample_bilinear_bwd_kernel<<<dim3((unsigned)ceil_div((long long)w * (C / 4), 128), (unsigned)(n_img * ceil_div(h, kUpRows))), 128, 0, stream>>>(
      dout, dout_ld, n_img, h, w, C, H, W, sy, sx, din, din_ld, accumulate, kUpRows);
  DBEV_CHECK_LAUNCH("upsample_bilinear_bwd_kernel");
  return DBEV_OK;
}

int pack_conv_weights(const float* w, int c_out, int c_in, int kh, int kw, int mode, float* out, cudaStream_t stream) {
  DBEV_CHECK_ARG(c_out > 0 && c_in > 0 && kh > 0 && kw > 0 && mode >= 0 && mode <= 2, "pack_conv_weights: bad arguments");
  DBEV_CHECK_ARG(mode != 2 || (kh == 3 && kw == 3), "pack_conv_weights: mode 2 is for 3x3 filters");
  pack_conv_weights_kernel<<<grid_for((long long)c_out * c_in * kh * kw, 256), 256, 0, stream>>>(w, c_out, c_in, kh, kw, mode, out);
  DBEV_CHECK_LAUNCH("pack_conv_weights_kernel");
  return DBEV_OK;
}

int pack_conv_weights_train(const float* w, int c_out, int c_in, int kh, int kw, int dgrad_mode, float* out_fwd, float* out_dgrad,
                            cudaStream_t stream) {
  DBEV_CHECK_ARG(c_out > 0 && c_in > 0 && kh > 0 && kw > 0 && kh * kw <= 9, "pack_conv_weights_train: filters up to 3x3");
  DBEV_CHECK_ARG(dgrad_mode == 1 || (dgrad_mode == 2 && kh == 3 && kw == 3), "pack_conv_weights_train: dgrad_mode 1, or 2 for 3x3 filters");
  const size_t smem = (size_t)32 * (32 * kh * kw + 1) * sizeof(float);
  pack_conv_weights_tile_kernel<<<dim3((unsigned)ceil_div(c_in, 32), (unsigned)ceil_div(c_out, 32)), 256, smem, stream>>>(
      w, c_out, c_in, kh, kw, dgrad_mode, out_fwd, out_dgrad);
  DBEV_CHECK_LAUNCH("pack_conv_weights_tile_kernel");
  return DBEV_OK;
}

int pack_conv_weights_batch(const long long* jobs_dev, int n_jobs, int total_tiles, cudaStream_t stream) {
  DBEV_CHECK_ARG(jobs_dev && n_jobs > 0 && total_tiles > 0, "pack_conv_weights_batch: empty job table");
  const size_t smem = (size_t)32 * (32 * 9 + 1) * sizeof(float);
  pack_conv_weights_batch_kernel<<<(unsigned)total_tiles, 256, smem, stream>>>(jobs_dev, n_jobs);
  DBEV_CHECK_LAUNCH("pack_conv_weights_batch_kernel");
  return DBEV_OK;
}

}  // namespace dbev
