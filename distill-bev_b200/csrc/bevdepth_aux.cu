// Steps on either side of the view transform in BEVDepth4D:
//   * shift_feature   mmdet3d/models/detectors/bevdet.py:267-321 — warps the adjacent frame's BEV
//     feature into the current ego frame: the reference materialises a [n,h,w,3,1] grid, one batched
//     matmul per pixel, a normalisation pass and F.grid_sample; here one kernel evaluates the affine
//     map and the bilinear taps per output pixel (grid never written).
//   * get_depth_loss  mmdet3d/models/detectors/bevdet.py:397-417 — BCE between sigmoid(depth logits)
//     and the one-hot depth bin of the sparse LiDAR depth map; the reference builds the one-hot
//     [B,N,H,W,D] tensor, permutes it, and runs sigmoid + weighted BCE as separate kernels; here the
//     bin index is compared on the fly (no one-hot), forward and backward are one pass each.
#include "bevdepth_aux.cuh"

namespace dbev {

namespace {

// F.grid_sample(align_corners=True) coordinate round trip, in the reference's fp32 order of operations
__device__ __forceinline__ float unnorm(float g, int size) {
  const float nrm = __fsub_rn(__fmul_rn(__fdiv_rn(g, (float)(size - 1)), 2.0f), 1.0f);   // bevdet.py:318
  return __fmul_rn(__fdiv_rn(__fadd_rn(nrm, 1.f), 2.f), (float)(size - 1));              // grid_sampler_unnormalize
}

struct Taps {
  int x0, y0;
  float w00, w01, w10, w11;  // (y0,x0), (y0,x1), (y1,x0), (y1,x1)
};

__device__ __forceinline__ Taps make_taps(const float* __restrict__ tf, int n, int x, int y, int h, int w) {
  const float* t = tf + n * 9;
  const float gx = t[0] * (float)x + t[1] * (float)y + t[2];
  const float gy = t[3] * (float)x + t[4] * (float)y + t[5];
  const float ix = unnorm(gx, w), iy = unnorm(gy, h);
  const float fx = floorf(ix), fy = floorf(iy);
  Taps r;
  r.x0 = (int)fx, r.y0 = (int)fy;
  const float ax = ix - fx, ay = iy - fy;   // == ix - ix_nw
  r.w00 = (1.f - ax) * (1.f - ay);
  r.w01 = ax * (1.f - ay);
  r.w10 = (1.f - ax) * ay;
  r.w11 = ax * ay;
  return r;
}

// thread = output pixel, loop over channels: neighbouring threads read neighbouring taps of one plane
__global__ void shift_feature_fwd_kernel(const float* __restrict__ in, const float* __restrict__ tf, int n_img,
                                         int C, int h, int w, float* __restrict__ out) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long hw = (long long)h * w;
  if (t >= n_img * hw) return;
  const int n = (int)(t / hw), p = (int)(t % hw), y = p / w, x = p % w;
  const Taps k = make_taps(tf, n, x, y, h, w);
  const bool vx0 = k.x0 >= 0 && k.x0 < w, vx1 = k.x0 + 1 >= 0 && k.x0 + 1 < w;
  const bool vy0 = k.y0 >= 0 && k.y0 < h, vy1 = k.y0 + 1 >= 0 && k.y0 + 1 < h;
  const float* base = in + (long long)n * C * hw + (long long)k.y0 * w + k.x0;
  float* o = out + (long long)n * C * hw + p;
  for (int c = 0; c < C; ++c, base += hw, o += hw) {
    float v = 0.f;
    if (vy0 && vx0) v += base[0] * k.w00;
    if (vy0 && vx1) v += base[1] * k.w01;
    if (vy1 && vx0) v += base[w] * k.w10;
    if (vy1 && vx1) v += base[w + 1] * k.w11;
    *o = v;
  }
}

__global__ void shift_feature_bwd_kernel(const float* __restrict__ gout, const float* __restrict__ tf, int n_img,
                                         int C, int h, int w, float* __restrict__ gin) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long hw = (long long)h * w;
  if (t >= n_img * hw) return;
  const int n = (int)(t / hw), p = (int)(t % hw), y = p / w, x = p % w;
  const Taps k = make_taps(tf, n, x, y, h, w);
  const bool vx0 = k.x0 >= 0 && k.x0 < w, vx1 = k.x0 + 1 >= 0 && k.x0 + 1 < w;
  const bool vy0 = k.y0 >= 0 && k.y0 < h, vy1 = k.y0 + 1 >= 0 && k.y0 + 1 < h;
  float* base = gin + (long long)n * C * hw + (long long)k.y0 * w + k.x0;
  const float* g = gout + (long long)n * C * hw + p;
  for (int c = 0; c < C; ++c, base += hw, g += hw) {
    const float v = *g;
    if (vy0 && vx0) atomicAdd(base, v * k.w00);
    if (vy0 && vx1) atomicAdd(base + 1, v * k.w01);
    if (vy1 && vx0) atomicAdd(base + w, v * k.w10);
    if (vy1 && vx1) atomicAdd(base + w + 1, v * k.w11);
  }
}

// element (bn, d, p): y = (bin(bn, p) == d), weight = (depth_gt != 0); BCE with the -100 log clamp of
// F.binary_cross_entropy. One partial sum per CTA (fixed order), reduced by the final kernel.
__device__ __forceinline__ int depth_bin(float gt, float dmin, float dstep, int D) {
  float b = floorf(__fdiv_rn(__fsub_rn(gt, dmin), dstep));
  b = fminf(fmaxf(b, 0.f), (float)D);
  return (int)b;   // == D: no class (F.one_hot would raise in the reference)
}

__global__ void __launch_bounds__(256)
depth_loss_fwd_kernel(const float* __restrict__ logits, const float* __restrict__ gt, int BN, int D, int HW,
                      float dmin, float dstep, double* __restrict__ partial) {
  __shared__ double red[256];
  const long long total = (long long)BN * D * HW;
  double acc = 0.0;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const int p = (int)(t % HW);
    const int d = (int)((t / HW) % D);
    const int bn = (int)(t / ((long long)HW * D));
    const float g = gt[(long long)bn * HW + p];
    if (g == 0.f) continue;
    const float x = logits[t];
    const float prob = 1.f / (1.f + expf(-x));
    const bool pos = depth_bin(g, dmin, dstep, D) == d;
    const float l = pos ? fmaxf(logf(prob), -100.f) : fmaxf(logf(1.f - prob), -100.f);
    acc -= (double)l;
  }
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[blockIdx.x] = red[0];
}

__global__ void depth_loss_final_kernel(const double* __restrict__ partial, int n, double scale,
                                        float* __restrict__ loss) {
  double s = 0.0;
  for (int i = 0; i < n; ++i) s += partial[i];
  *loss = (float)(s * scale);
}

__global__ void depth_loss_bwd_kernel(const float* __restrict__ logits, const float* __restrict__ gt, int BN, int D,
                                      int HW, float dmin, float dstep, float scale,
                                      const float* __restrict__ grad_loss, float* __restrict__ grad) {
  const long long total = (long long)BN * D * HW;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  const int p = (int)(t % HW);
  const int d = (int)((t / HW) % D);
  const int bn = (int)(t / ((long long)HW * D));
  const float g = gt[(long long)bn * HW + p];
  float r = 0.f;
  if (g != 0.f) {
    const float x = logits[t];
    const float prob = 1.f / (1.f + expf(-x));
    const bool pos = depth_bin(g, dmin, dstep, D) == d;
    // d/dx of -[y log p + (1-y) log(1-p)] = p - y (outside the clamped region)
    r = (prob - (pos ? 1.f : 0.f)) * scale * grad_loss[0];
  }
  grad[t] = r;
}

}  // namespace

int shift_feature_forward(const float* in, const float* tf, int n, int C, int h, int w, float* out,
                          cudaStream_t stream) {
  DBEV_CHECK_ARG(n > 0 && C > 0 && h > 1 && w > 1, "shift_feature: bad sizes");
  shift_feature_fwd_kernel<<<ceil_div((long long)n * h * w, 256), 256, 0, stream>>>(in, tf, n, C, h, w, out);
  DBEV_CHECK_LAUNCH("shift_feature_fwd_kernel");
  return DBEV_OK;
}

int shift_feature_backward(const float* grad_out, const float* tf, int n, int C, int h, int w, float* grad_in,
                           cudaStream_t stream) {
  DBEV_CHECK_ARG(n > 0 && C > 0 && h > 1 && w > 1, "shift_feature: bad sizes");
  DBEV_CUDA(cudaMemsetAsync(grad_in, 0, (size_t)n * C * h * w * sizeof(float), stream));
  shift_feature_bwd_kernel<<<ceil_div((long long)n * h * w, 256), 256, 0, stream>>>(grad_out, tf, n, C, h, w,
                                                                                    grad_in);
  DBEV_CHECK_LAUNCH("shift_feature_bwd_kernel");
  return DBEV_OK;
}

size_t depth_loss_ws_bytes() { return (size_t)kNumSMs * 4 * sizeof(double); }

int depth_loss_forward(const float* logits, const float* depth_gt, int BN, int D, int HW, float dmin,
                       float dstep, float loss_weight, float* loss, void* ws, size_t ws_bytes,
                       cudaStream_t stream) {
  DBEV_CHECK_ARG(BN > 0 && D > 0 && HW > 0 && dstep > 0.f, "depth_loss: bad sizes");
  DBEV_CHECK_ARG(ws && ws_bytes >= depth_loss_ws_bytes(), "depth_loss: workspace too small");
  const int grid = kNumSMs * 4;
  depth_loss_fwd_kernel<<<grid, 256, 0, stream>>>(logits, depth_gt, BN, D, HW, dmin, dstep, (double*)ws);
  depth_loss_final_kernel<<<1, 1, 0, stream>>>((const double*)ws, grid,
                                               (double)loss_weight / ((double)BN * D * HW), loss);
  DBEV_CHECK_LAUNCH("depth_loss_fwd_kernel");
  return DBEV_OK;
}

int depth_loss_backward(const float* logits, const float* depth_gt, int BN, int D, int HW, float dmin,
                        float dstep, float loss_weight, const float* grad_loss, float* grad_logits,
                        cudaStream_t stream) {
  DBEV_CHECK_ARG(BN > 0 && D > 0 && HW > 0 && dstep > 0.f, "depth_loss: bad sizes");
  const long long total = (long long)BN * D * HW;
  depth_loss_bwd_kernel<<<ceil_div(total, 256), 256, 0, stream>>>(
      logits, depth_gt, BN, D, HW, dmin, dstep, loss_weight / (float)((double)total), grad_loss, grad_logits);
  DBEV_CHECK_LAUNCH("depth_loss_bwd_kernel");
  return DBEV_OK;
}

}  // namespace dbev
