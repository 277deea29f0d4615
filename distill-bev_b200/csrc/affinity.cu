// Affinity distillation loss.
//
// Reference: BEVDetDistill.affinity_distill_loss, list branch
//   mmdet3d/models/detectors/bevdet_distill.py:735-748, fed by the masked-cell gather at
//   :1294-1321 (feat[c][mask] for every channel -> [K, C] rows per sample):
//     loss = sum_b  weight * mean_{K_b x K_b} criterion(T_b T_b^T, S_b S_b^T)
//   with criterion = mmdet SmoothL1Loss(beta=1) / L1Loss / MSELoss (reduction 'mean').
// The K x K gram matrices are never written: a CTA computes a 64 x 64 tile of BOTH grams in
// registers, applies the criterion to the difference and keeps one partial sum (fixed order, no
// float atomics). The matrices are symmetric, so only tiles on or above the diagonal are visited.
// Backward recomputes the tile row to get dL/dA_s and multiplies it into the student rows:
//   dS = 2 * G S,  G = -weight * grad * criterion'(A_t - A_s) / K_b^2.
#include "affinity.cuh"

#include "sort.cuh"

namespace dbev {

namespace {

constexpr int kT = 64;    // gram tile edge
constexpr int kKC = 16;   // channels staged per step
constexpr int kLd = 68;   // padded row length of the k-major staging buffers

__device__ __forceinline__ float crit_value(float d, int kind, float beta) {
  const float a = fabsf(d);
  if (kind == 1) return a;
  if (kind == 2) return d * d;
  return a < beta ? 0.5f * a * a / beta : a - 0.5f * beta;
}

// derivative with respect to d = A_t - A_s
__device__ __forceinline__ float crit_deriv(float d, int kind, float beta) {
  if (kind == 1) return d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f);
  if (kind == 2) return 2.f * d;
  const float a = fabsf(d);
  return a < beta ? d / beta : (d > 0.f ? 1.f : -1.f);
}

__global__ void aff_flags_kernel(const float* __restrict__ mask_a, const float* __restrict__ mask_b,
                                 long long n, int* __restrict__ flags) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  bool f = mask_a[i] != 0.f;
  if (mask_b) f = f || (mask_b[i] != 0.f);
  flags[i] = f ? 1 : 0;
}

__global__ void aff_compact_kernel(const int* __restrict__ flags, const int* __restrict__ excl,
                                   const int* __restrict__ total, int B, int HW,
                                   int* __restrict__ row_cell, int* __restrict__ row_offsets) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long n = (long long)B * HW;
  if (i <= B) row_offsets[i] = (i == B) ? *total : excl[i * HW];
  if (i >= n) return;
  if (flags[i]) row_cell[excl[i]] = (int)(i % HW);
}

// rows[r][c] = feat[b][c][cell(r)]
__global__ void aff_gather_rows_kernel(const float* __restrict__ feat,
                                       const int* __restrict__ row_cell,
                                       const int* __restrict__ row_offsets, int B, int C, int HW,
                                       int k_total, float* __restrict__ rows) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)k_total * C) return;
  const int r = (int)(t / C), c = (int)(t % C);
  int b = 0;
  while (b + 1 < B && r >= row_offsets[b + 1]) ++b;
  rows[t] = feat[((long long)b * C + c) * HW + row_cell[r]];
}

// grad[b][c][cell(r)] = d_rows[r][c]; the rest of grad was zero-filled.
__global__ void aff_scatter_rows_kernel(const float* __restrict__ d_rows,
                                        const int* __restrict__ row_cell,
                                        const int* __restrict__ row_offsets, int B, int C, int HW,
                                        int k_total, float* __restrict__ grad) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)k_total * C) return;
  const int r = (int)(t / C), c = (int)(t % C);
  int b = 0;
  while (b + 1 < B && r >= row_offsets[b + 1]) ++b;
  grad[((long long)b * C + c) * HW + row_cell[r]] = d_rows[t];
}

// Stage kKC channels of 64 rows k-major: dst[kk][r] = rows[(r0 + r) * C + c0 + kk] (zero tail).
__device__ __forceinline__ void stage_rows(const float* __restrict__ rows, int r0, int r_end, int C,
                                           int c0, float (*dst)[kLd]) {
  const int t = threadIdx.x;
  const int r = t >> 2, q = (t & 3) * 4;
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (r0 + r < r_end && c0 + q < C) v = __ldg(reinterpret_cast<const float4*>(rows + (long long)(r0 + r) * C + c0 + q));
  dst[q + 0][r] = v.x;
  dst[q + 1][r] = v.y;
  dst[q + 2][r] = v.z;
  dst[q + 3][r] = v.w;
}

struct GramSmem {
  float ti[kKC][kLd], tj[kKC][kLd], si[kKC][kLd], sj[kKC][kLd];
};

// d[a][b] = (T_i T_j^T - S_i S_j^T)[ty*4+a][tx*4+b] for the tile pair (i0, j0).
__device__ __forceinline__ void gram_diff_tile(const float* __restrict__ t_rows,
                                               const float* __restrict__ s_rows, int i0, int j0,
                                               int r_end, int C, GramSmem& sm, float d[4][4]) {
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float at[4][4], as[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) at[a][b] = as[a][b] = 0.f;
  for (int c0 = 0; c0 < C; c0 += kKC) {
    stage_rows(t_rows, i0, r_end, C, c0, sm.ti);
    stage_rows(t_rows, j0, r_end, C, c0, sm.tj);
    stage_rows(s_rows, i0, r_end, C, c0, sm.si);
    stage_rows(s_rows, j0, r_end, C, c0, sm.sj);
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < kKC; ++kk) {
      const float4 a4 = *reinterpret_cast<const float4*>(&sm.ti[kk][ty * 4]);
      const float4 b4 = *reinterpret_cast<const float4*>(&sm.tj[kk][tx * 4]);
      const float4 c4 = *reinterpret_cast<const float4*>(&sm.si[kk][ty * 4]);
      const float4 e4 = *reinterpret_cast<const float4*>(&sm.sj[kk][tx * 4]);
      const float av[4] = {a4.x, a4.y, a4.z, a4.w}, bv[4] = {b4.x, b4.y, b4.z, b4.w};
      const float cv[4] = {c4.x, c4.y, c4.z, c4.w}, ev[4] = {e4.x, e4.y, e4.z, e4.w};
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          at[a][b] = fmaf(av[a], bv[b], at[a][b]);
          as[a][b] = fmaf(cv[a], ev[b], as[a][b]);
        }
    }
    __syncthreads();
  }
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) d[a][b] = at[a][b] - as[a][b];
}

struct AffOffsets {
  int off[65];  // row offsets of up to 64 samples (host copy, passed by value)
};

__global__ void __launch_bounds__(256)
aff_fwd_kernel(const float* __restrict__ t_rows, const float* __restrict__ s_rows, AffOffsets ofs,
               int C, int max_tiles, int kind, float beta, float* __restrict__ partial) {
  const int b = blockIdx.z, ti = blockIdx.y, tj = blockIdx.x;
  const int r0 = ofs.off[b], r_end = ofs.off[b + 1];
  const int tiles = (r_end - r0 + kT - 1) / kT;
  float* slot = partial + ((long long)b * max_tiles + ti) * max_tiles + tj;
  if (ti >= tiles || tj >= tiles || ti > tj) {
    if (threadIdx.x == 0) *slot = 0.f;
    return;
  }
  __shared__ GramSmem sm;
  __shared__ float red[8];
  float d[4][4];
  gram_diff_tile(t_rows, s_rows, r0 + ti * kT, r0 + tj * kT, r_end, C, sm, d);
  float s = 0.f;
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int bb = 0; bb < 4; ++bb) s += crit_value(d[a][bb], kind, beta);  // padded rows give d = 0
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float tot = 0.f;
    for (int w = 0; w < 8; ++w) tot += red[w];
    *slot = (ti == tj) ? tot : 2.f * tot;
  }
}

// loss = weight * sum_b (sum of the sample's partials) / K_b^2, samples in order.
__global__ void __launch_bounds__(256)
aff_final_kernel(const float* __restrict__ partial, AffOffsets ofs, int B, int max_tiles,
                 float weight, float* __restrict__ loss) {
  __shared__ float red[8];
  __shared__ float total;
  if (threadIdx.x == 0) total = 0.f;
  __syncthreads();
  for (int b = 0; b < B; ++b) {
    const int K = ofs.off[b + 1] - ofs.off[b];
    if (K <= 0) continue;
    const int tiles = (K + kT - 1) / kT;
    const float* p = partial + (long long)b * max_tiles * max_tiles;
    float s = 0.f;
    for (int e = threadIdx.x; e < tiles * tiles; e += 256) s += p[(e / tiles) * max_tiles + (e % tiles)];
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
      float tot = 0.f;
      for (int w = 0; w < 8; ++w) tot += red[w];
      total += weight * (tot / ((float)K * (float)K));
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) *loss = total;
}

// CTA (tile row ti, 128-channel chunk, sample): d_rows[i, chunk] = sum_j G[i, j] * S[j, chunk].
__global__ void __launch_bounds__(256)
aff_bwd_kernel(const float* __restrict__ t_rows, const float* __restrict__ s_rows, AffOffsets ofs,
               int C, int kind, float beta, float weight, const float* __restrict__ grad_loss,
               float* __restrict__ d_rows) {
  const int b = blockIdx.z, ti = blockIdx.x, cc = blockIdx.y * 128;
  const int r0 = ofs.off[b], r_end = ofs.off[b + 1];
  const int K = r_end - r0;
  const int tiles = (K + kT - 1) / kT;
  if (ti >= tiles) return;
  __shared__ union {
    GramSmem g;
    float s_stage[32][128];
  } sm;
  __shared__ float G_s[kT][kT + 1];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const float coef = -2.f * weight * (*grad_loss) / ((float)K * (float)K);
  float acc[4][8];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int q = 0; q < 8; ++q) acc[a][q] = 0.f;
  const int i0 = r0 + ti * kT;
  const int cw = min(128, C - cc);  // channels of this chunk (multiple of 4)
  for (int tj = 0; tj < tiles; ++tj) {
    const int j0 = r0 + tj * kT;
    float d[4][4];
    gram_diff_tile(t_rows, s_rows, i0, j0, r_end, C, sm.g, d);
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int bb = 0; bb < 4; ++bb) G_s[ty * 4 + a][tx * 4 + bb] = coef * crit_deriv(d[a][bb], kind, beta);
    __syncthreads();
    for (int h = 0; h < 2; ++h) {  // two halves of 32 student rows
      for (int e = threadIdx.x; e < 32 * 32; e += 256) {
        const int jr = e >> 5, c4 = (e & 31) * 4;
        const int row = j0 + h * 32 + jr;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row < r_end && c4 < cw)
          v = __ldg(reinterpret_cast<const float4*>(s_rows + (long long)row * C + cc + c4));
        *reinterpret_cast<float4*>(&sm.s_stage[jr][c4]) = v;
      }
      __syncthreads();
#pragma unroll 4
      for (int jj = 0; jj < 32; ++jj) {
        const float4 s0 = *reinterpret_cast<const float4*>(&sm.s_stage[jj][tx * 4]);
        const float4 s1 = *reinterpret_cast<const float4*>(&sm.s_stage[jj][64 + tx * 4]);
#pragma unroll
        for (int a = 0; a < 4; ++a) {
          const float g = G_s[ty * 4 + a][h * 32 + jj];
          acc[a][0] = fmaf(g, s0.x, acc[a][0]);
          acc[a][1] = fmaf(g, s0.y, acc[a][1]);
          acc[a][2] = fmaf(g, s0.z, acc[a][2]);
          acc[a][3] = fmaf(g, s0.w, acc[a][3]);
          acc[a][4] = fmaf(g, s1.x, acc[a][4]);
          acc[a][5] = fmaf(g, s1.y, acc[a][5]);
          acc[a][6] = fmaf(g, s1.z, acc[a][6]);
          acc[a][7] = fmaf(g, s1.w, acc[a][7]);
        }
      }
      __syncthreads();
    }
  }
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int row = i0 + ty * 4 + a;
    if (row >= r_end) continue;
    float* o = d_rows + (long long)row * C + cc;
    if (tx * 4 < cw) *reinterpret_cast<float4*>(o + tx * 4) = make_float4(acc[a][0], acc[a][1], acc[a][2], acc[a][3]);
    if (64 + tx * 4 < cw)
      *reinterpret_cast<float4*>(o + 64 + tx * 4) = make_float4(acc[a][4], acc[a][5], acc[a][6], acc[a][7]);
  }
}

int fill_offsets(const int* row_offsets_host, int B, AffOffsets* o, int* max_k) {
  DBEV_CHECK_ARG(B >= 1 && B <= 64, "affinity: batch must be in [1, 64]");
  *max_k = 0;
  for (int b = 0; b <= B; ++b) o->off[b] = row_offsets_host[b];
  for (int b = 0; b < B; ++b) {
    const int k = o->off[b + 1] - o->off[b];
    DBEV_CHECK_ARG(k >= 0, "affinity: row offsets must be non-decreasing");
    if (k > *max_k) *max_k = k;
  }
  return DBEV_OK;
}

}  // namespace

size_t affinity_select_ws_bytes(int B, int HW) {
  const long long n = (long long)B * HW;
  return align_up((size_t)n * 4) * 2 + 256 + scan_ws_bytes(n) + 1024;
}

int affinity_select(const float* mask_a, const float* mask_b, int B, int HW, int* row_cell,
                    int* row_offsets, void* ws, size_t ws_bytes, cudaStream_t stream) {
  DBEV_CHECK_ARG(B >= 1 && HW >= 1, "affinity_select: bad sizes");
  const long long n = (long long)B * HW;
  DBEV_CHECK_ARG(n < (1ll << 31), "affinity_select: B*HW too large");
  Workspace w(ws, ws_bytes);
  int* flags = w.take<int>(n);
  int* excl = w.take<int>(n);
  int* total = w.take<int>(1);
  if (!w.ok()) {
    set_last_error("affinity_select: workspace too small");
    return DBEV_ERR_WORKSPACE;
  }
  const size_t consumed = align_up(w.used);
  const int grid = ceil_div(n + 1, 256);
  aff_flags_kernel<<<grid, 256, 0, stream>>>(mask_a, mask_b, n, flags);
  int rc = exclusive_scan_i32(flags, excl, (int)n, total, (char*)ws + consumed,
                              ws_bytes > consumed ? ws_bytes - consumed : 0, stream);
  if (rc != DBEV_OK) return rc;
  aff_compact_kernel<<<grid, 256, 0, stream>>>(flags, excl, total, B, HW, row_cell, row_offsets);
  DBEV_CHECK_LAUNCH("aff_compact_kernel");
  return DBEV_OK;
}

int affinity_gather_rows(const float* feat, const int* row_cell, const int* row_offsets, int B,
                         int C, int HW, int k_total, float* rows, cudaStream_t stream) {
  DBEV_CHECK_ARG(B >= 1 && C >= 1 && HW >= 1 && k_total >= 0, "affinity_gather_rows: bad sizes");
  if (k_total == 0) return DBEV_OK;
  aff_gather_rows_kernel<<<ceil_div((long long)k_total * C, 256), 256, 0, stream>>>(
      feat, row_cell, row_offsets, B, C, HW, k_total, rows);
  DBEV_CHECK_LAUNCH("aff_gather_rows_kernel");
  return DBEV_OK;
}

size_t affinity_partial_floats(const int* row_offsets_host, int B) {
  int max_k = 0;
  for (int b = 0; b < B; ++b) {
    const int k = row_offsets_host[b + 1] - row_offsets_host[b];
    if (k > max_k) max_k = k;
  }
  const size_t t = (size_t)((max_k + kT - 1) / kT);
  return (size_t)B * t * t + 1;
}

int affinity_forward(const float* t_rows, const float* s_rows, const int* row_offsets_host, int B,
                     int C, int kind, float beta, float weight, float* partial, float* loss,
                     cudaStream_t stream) {
  DBEV_CHECK_ARG(C >= 4 && C % 4 == 0, "affinity_forward: C must be a multiple of 4");
  DBEV_CHECK_ARG(kind >= 0 && kind <= 2 && beta > 0.f, "affinity_forward: bad criterion");
  DBEV_CHECK_ARG(((uintptr_t)t_rows & 15) == 0 && ((uintptr_t)s_rows & 15) == 0,
                 "affinity_forward: rows must be 16-byte aligned");
  AffOffsets ofs;
  int max_k = 0;
  int rc = fill_offsets(row_offsets_host, B, &ofs, &max_k);
  if (rc != DBEV_OK) return rc;
  const int max_tiles = (max_k + kT - 1) / kT;
  if (max_tiles > 0) {
    dim3 grid(max_tiles, max_tiles, B);
    aff_fwd_kernel<<<grid, 256, 0, stream>>>(t_rows, s_rows, ofs, C, max_tiles, kind, beta, partial);
    DBEV_CHECK_LAUNCH("aff_fwd_kernel");
  }
  aff_final_kernel<<<1, 256, 0, stream>>>(partial, ofs, B, max_tiles, weight, loss);
  DBEV_CHECK_LAUNCH("aff_final_kernel");
  return DBEV_OK;
}

int affinity_backward(const float* t_rows, const float* s_rows, const int* row_offsets_host, int B,
                      int C, int kind, float beta, float weight, const float* grad_loss,
                      float* d_s_rows, cudaStream_t stream) {
  DBEV_CHECK_ARG(C >= 4 && C % 4 == 0, "affinity_backward: C must be a multiple of 4");
  DBEV_CHECK_ARG(kind >= 0 && kind <= 2 && beta > 0.f, "affinity_backward: bad criterion");
  AffOffsets ofs;
  int max_k = 0;
  int rc = fill_offsets(row_offsets_host, B, &ofs, &max_k);
  if (rc != DBEV_OK) return rc;
  const int max_tiles = (max_k + kT - 1) / kT;
  if (max_tiles == 0) return DBEV_OK;
  dim3 grid(max_tiles, ceil_div(C, 128), B);
  aff_bwd_kernel<<<grid, 256, 0, stream>>>(t_rows, s_rows, ofs, C, kind, beta, weight, grad_loss,
                                           d_s_rows);
  DBEV_CHECK_LAUNCH("aff_bwd_kernel");
  return DBEV_OK;
}

int affinity_scatter_rows(const float* d_rows, const int* row_cell, const int* row_offsets, int B,
                          int C, int HW, int k_total, float* grad, cudaStream_t stream) {
  DBEV_CHECK_ARG(B >= 1 && C >= 1 && HW >= 1 && k_total >= 0, "affinity_scatter_rows: bad sizes");
  DBEV_CUDA(cudaMemsetAsync(grad, 0, (size_t)B * C * HW * sizeof(float), stream));
  if (k_total == 0) return DBEV_OK;
  aff_scatter_rows_kernel<<<ceil_div((long long)k_total * C, 256), 256, 0, stream>>>(
      d_rows, row_cell, row_offsets, B, C, HW, k_total, grad);
  DBEV_CHECK_LAUNCH("aff_scatter_rows_kernel");
  return DBEV_OK;
}

}  // namespace dbev
