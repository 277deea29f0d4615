// Sparse 3D convolution: rulebook (neighbour table) + output-major gather convolution.
// See spconv.cuh for the design and the reference lines each piece follows.
#include "spconv.cuh"

#include "sort.cuh"

namespace dbev {

namespace {

constexpr uint32_t kEmpty = 0xFFFFFFFFu;

struct Geom {  // device copy of SpConvGeom
  int kz, ky, kx, sz, sy, sx, pz, py, px, dz, dy, dx;
  int iz, iy, ix, oz, oy, ox, batch;
};

Geom to_dev(const SpConvGeom& g) {
  Geom r;
  r.kz = g.k[0], r.ky = g.k[1], r.kx = g.k[2];
  r.sz = g.s[0], r.sy = g.s[1], r.sx = g.s[2];
  r.pz = g.p[0], r.py = g.p[1], r.px = g.p[2];
  r.dz = g.d[0], r.dy = g.d[1], r.dx = g.d[2];
  r.iz = g.in_shape[0], r.iy = g.in_shape[1], r.ix = g.in_shape[2];
  r.oz = g.out_shape[0], r.oy = g.out_shape[1], r.ox = g.out_shape[2];
  r.batch = g.batch;
  return r;
}

uint32_t pow2_at_least(unsigned long long v) {
  uint32_t p = 64;
  while (p < v) p <<= 1;
  return p;
}

__device__ __forceinline__ uint32_t hash_slot(uint32_t key, uint32_t mask) {
  uint32_t h = key * 0x9E3779B1u;
  h ^= h >> 15;
  h *= 0x85EBCA77u;
  h ^= h >> 13;
  return h & mask;
}

__device__ __forceinline__ int hash_find(const uint32_t* __restrict__ hkeys,
                                         const int* __restrict__ hvals, uint32_t mask,
                                         uint32_t key) {
  uint32_t slot = hash_slot(key, mask);
  while (true) {
    const uint32_t k = hkeys[slot];
    if (k == key) return hvals[slot];
    if (k == kEmpty) return -1;
    slot = (slot + 1) & mask;
  }
}

// key -> row index map of the active input voxels
__global__ void sp_hash_insert_kernel(const int4* __restrict__ coors, int n, int batch, int Z,
                                      int Y, int X, uint32_t* hkeys, int* hvals, uint32_t mask) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int4 c = coors[i];  // (b, z, y, x)
  if ((unsigned)c.x >= (unsigned)batch || (unsigned)c.y >= (unsigned)Z ||
      (unsigned)c.z >= (unsigned)Y || (unsigned)c.w >= (unsigned)X)
    return;
  const uint32_t key = (uint32_t)(((c.x * Z + c.y) * Y + c.z)) * (uint32_t)X + (uint32_t)c.w;
  uint32_t slot = hash_slot(key, mask);
  while (true) {
    const uint32_t prev = atomicCAS(&hkeys[slot], kEmpty, key);
    if (prev == kEmpty || prev == key) {
      hvals[slot] = i;  // duplicates: one of them wins (the reference keeps the last, geometry.h:289)
      return;
    }
    slot = (slot + 1) & mask;
  }
}

// nbr[k][o] = row of the input at out*stride - padding + k*dilation (geometry.h:64-66 solved
// for the input position), -1 when that cell is inactive or outside the input grid.
__global__ void sp_table_kernel(const int4* __restrict__ out_coors, int n_out, Geom g,
                                const uint32_t* __restrict__ hkeys,
                                const int* __restrict__ hvals, uint32_t mask,
                                int* __restrict__ nbr) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int kvol = g.kz * g.ky * g.kx;
  if (t >= (long long)kvol * n_out) return;
  const int k = (int)(t / n_out);
  const int o = (int)(t - (long long)k * n_out);
  const int4 c = out_coors[o];
  const int kx = k % g.kx, ky = (k / g.kx) % g.ky, kz = k / (g.kx * g.ky);
  const int z = c.y * g.sz - g.pz + kz * g.dz;
  const int y = c.z * g.sy - g.py + ky * g.dy;
  const int x = c.w * g.sx - g.px + kx * g.dx;
  int r = -1;
  if ((unsigned)c.x < (unsigned)g.batch && (unsigned)z < (unsigned)g.iz &&
      (unsigned)y < (unsigned)g.iy && (unsigned)x < (unsigned)g.ix) {
    const uint32_t key = (uint32_t)(((c.x * g.iz + z) * g.iy + y)) * (uint32_t)g.ix + (uint32_t)x;
    r = hash_find(hkeys, hvals, mask, key);
  }
  nbr[t] = r;
}

// Every (input, kernel offset) pair names one output cell when (in + p - k*d) is a multiple of
// the stride (getValidOutPos, geometry.h:24-84). Distinct cells are collected through a hash set.
__global__ void sp_candidates_kernel(const int4* __restrict__ coors, int n_in, Geom g, int cz, int cy,
                                     int cx, uint32_t* hset, uint32_t mask,
                                     uint32_t* __restrict__ out_keys, long long max_out, int* counter) {
  // One thread per (input, j): only offsets with (in + p - k) % stride == 0 can name an output, at
  // most cz*cy*cx = prod ceil(K/stride) of the K^3 per input (8 of 27 for a 3x3x3 stride-2 conv).
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int per = cz * cy * cx;
  if (t >= (long long)per * n_in) return;
  const int j = (int)(t / n_in);
  const int i = (int)(t - (long long)j * n_in);
  const int4 c = coors[i];
  if ((unsigned)c.x >= (unsigned)g.batch || (unsigned)c.y >= (unsigned)g.iz ||
      (unsigned)c.z >= (unsigned)g.iy || (unsigned)c.w >= (unsigned)g.ix)
    return;
  const int jx = j % cx, jy = (j / cx) % cy, jz = j / (cx * cy);
  const int kz = g.sz == 1 ? jz : (c.y + g.pz) % g.sz + jz * g.sz;
  const int ky = g.sy == 1 ? jy : (c.z + g.py) % g.sy + jy * g.sy;
  const int kx = g.sx == 1 ? jx : (c.w + g.px) % g.sx + jx * g.sx;
  if (kz >= g.kz || ky >= g.ky || kx >= g.kx) return;
  const int nz = c.y + g.pz - kz * g.dz, ny = c.z + g.py - ky * g.dy, nx = c.w + g.px - kx * g.dx;
  if (nz < 0 || ny < 0 || nx < 0) return;
  if (nz % g.sz || ny % g.sy || nx % g.sx) return;
  const int oz = nz / g.sz, oy = ny / g.sy, ox = nx / g.sx;
  if (oz >= g.oz || oy >= g.oy || ox >= g.ox) return;
  const uint32_t key = (uint32_t)(((c.x * g.oz + oz) * g.oy + oy)) * (uint32_t)g.ox + (uint32_t)ox;
  uint32_t slot = hash_slot(key, mask);
  while (true) {
    const uint32_t prev = atomicCAS(&hset[slot], kEmpty, key);
    if (prev == key) return;
    if (prev == kEmpty) {
      const int pos = atomicAdd(counter, 1);
      if (pos < max_out) out_keys[pos] = key;
      return;
    }
    slot = (slot + 1) & mask;
  }
}

__global__ void sp_decode_keys_kernel(const uint32_t* __restrict__ keys, int n, int Z, int Y, int X,
                                      int4* __restrict__ coors) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t k = keys[i];
  int4 c;
  c.w = (int)(k % (uint32_t)X);
  k /= (uint32_t)X;
  c.z = (int)(k % (uint32_t)Y);
  k /= (uint32_t)Y;
  c.y = (int)(k % (uint32_t)Z);
  c.x = (int)(k / (uint32_t)Z);
  coors[i] = c;
}

// One CTA per kernel offset: stream-compact the (in, out) pairs in ascending output order.
__global__ void __launch_bounds__(1024) sp_pairs_from_table_kernel(const int* __restrict__ nbr,
                                                                   int n_out, int pair_stride,
                                                                   int* __restrict__ pairs,
                                                                   int* __restrict__ indice_num) {
  const int k = blockIdx.x;
  const int* row = nbr + (long long)k * n_out;
  int* in_row = pairs + (long long)k * 2 * pair_stride;
  int* out_row = in_row + pair_stride;
  __shared__ int warp_tot[32];
  __shared__ int base_s;
  if (threadIdx.x == 0) base_s = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int o0 = 0; o0 < n_out; o0 += 1024) {
    const int o = o0 + threadIdx.x;
    const int v = o < n_out ? row[o] : -1;
    const unsigned bal = __ballot_sync(0xffffffffu, v >= 0);
    if (lane == 0) warp_tot[wid] = __popc(bal);
    __syncthreads();
    int before = 0;
    for (int w = 0; w < wid; ++w) before += warp_tot[w];
    const int pos = base_s + before + __popc(bal & lanemask_lt());
    if (v >= 0 && pos < pair_stride) {
      in_row[pos] = v;
      out_row[pos] = o;
    }
    __syncthreads();
    if (threadIdx.x == 1023) base_s = pos + (v >= 0 ? 1 : 0);
    __syncthreads();
  }
  if (threadIdx.x == 0) indice_num[k] = base_s;
}

__global__ void sp_table_from_pairs_kernel(const int* __restrict__ pairs,
                                           const int* __restrict__ indice_num, int pair_stride,
                                           int inverse, int n_out, int* __restrict__ nbr) {
  const int k = blockIdx.y;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= indice_num[k] || j >= pair_stride) return;
  const int* in_row = pairs + (long long)k * 2 * pair_stride;
  const int* out_row = in_row + pair_stride;
  const int src = inverse ? out_row[j] : in_row[j];
  const int dst = inverse ? in_row[j] : out_row[j];
  if ((unsigned)dst < (unsigned)n_out) nbr[(long long)k * n_out + dst] = src;
}

// ---------------------------------------------------------------------------------------------
// Output-major gather convolution, fp32 FMA. A CTA owns 64 consecutive outputs x all C_out; per
// kernel offset whose neighbour column is non-empty for this tile it stages the 64 gathered rows
// (32 input channels at a time) and the matching W[k] slab in shared memory. Thread (ty, tx)
// accumulates R = 64/TY rows x 4 output channels.
// ---------------------------------------------------------------------------------------------
constexpr int kConvTM = 64;
constexpr int kConvKC = 32;

template <int COUT>
__global__ void __launch_bounds__(256)
sp_conv_fma_kernel(const float* __restrict__ in_feats, int c_in, const float* __restrict__ w,
                   const int* __restrict__ nbr, int n_out, int kvol,
                   const float* __restrict__ scale, const float* __restrict__ shift,
                   const float* __restrict__ residual, int relu, float* __restrict__ out) {
  constexpr int TX = COUT / 4, TY = 256 / TX, R = kConvTM / TY;
  static_assert(R >= 1 && R * TY == kConvTM, "tile shape");
  __shared__ float A_s[kConvTM][kConvKC + 1];
  __shared__ __align__(16) float W_s[kConvKC][COUT];
  __shared__ int idx_s[kConvTM];

  const int tid = threadIdx.x;
  const int tx = tid % TX, ty = tid / TX;
  const int o0 = blockIdx.x * kConvTM;

  float acc[R][4];
#pragma unroll
  for (int j = 0; j < R; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;

  for (int k = 0; k < kvol; ++k) {
    int mine = -1;
    if (tid < kConvTM) {
      const int o = o0 + tid;
      mine = o < n_out ? nbr[(long long)k * n_out + o] : -1;
      idx_s[tid] = mine;
    }
    if (!__syncthreads_or(mine >= 0)) continue;  // no output of this tile sees offset k
    for (int c0 = 0; c0 < c_in; c0 += kConvKC) {
      const int kc = min(kConvKC, c_in - c0);
      for (int e = tid; e < kConvTM * kConvKC; e += 256) {
        const int r = e / kConvKC, c = e % kConvKC;
        const int idx = idx_s[r];
        A_s[r][c] = (idx >= 0 && c < kc) ? __ldg(in_feats + (long long)idx * c_in + c0 + c) : 0.f;
      }
      for (int e = tid; e < kConvKC * TX; e += 256) {
        const int ci = e / TX, c4 = e % TX;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (ci < kc)
          v = __ldg(reinterpret_cast<const float4*>(w + ((long long)k * c_in + c0 + ci) * COUT) + c4);
        *reinterpret_cast<float4*>(&W_s[ci][c4 * 4]) = v;
      }
      __syncthreads();
#pragma unroll 8
      for (int ci = 0; ci < kConvKC; ++ci) {
        const float4 wv = *reinterpret_cast<const float4*>(&W_s[ci][tx * 4]);
#pragma unroll
        for (int j = 0; j < R; ++j) {
          const float a = A_s[ty + j * TY][ci];
          acc[j][0] = fmaf(a, wv.x, acc[j][0]);
          acc[j][1] = fmaf(a, wv.y, acc[j][1]);
          acc[j][2] = fmaf(a, wv.z, acc[j][2]);
          acc[j][3] = fmaf(a, wv.w, acc[j][3]);
        }
      }
      __syncthreads();
    }
  }

  float4 sc = make_float4(1.f, 1.f, 1.f, 1.f), sh = make_float4(0.f, 0.f, 0.f, 0.f);
  if (scale) sc = __ldg(reinterpret_cast<const float4*>(scale) + tx);
  if (shift) sh = __ldg(reinterpret_cast<const float4*>(shift) + tx);
#pragma unroll
  for (int j = 0; j < R; ++j) {
    const int o = o0 + ty + j * TY;
    if (o >= n_out) continue;
    float4 v;
    v.x = fmaf(acc[j][0], sc.x, sh.x);
    v.y = fmaf(acc[j][1], sc.y, sh.y);
    v.z = fmaf(acc[j][2], sc.z, sh.z);
    v.w = fmaf(acc[j][3], sc.w, sh.w);
    if (residual) {
      const float4 rr = __ldg(reinterpret_cast<const float4*>(residual + (long long)o * COUT) + tx);
      v.x += rr.x, v.y += rr.y, v.z += rr.z, v.w += rr.w;
    }
    if (relu) v.x = fmaxf(v.x, 0.f), v.y = fmaxf(v.y, 0.f), v.z = fmaxf(v.z, 0.f), v.w = fmaxf(v.w, 0.f);
    *reinterpret_cast<float4*>(out + (long long)o * COUT + tx * 4) = v;
  }
}

// ---------------------------------------------------------------------------------------------
// Narrow layers (C_in * C_out <= 512: 5->16, 16->16, 16->32 of the LidarFormer encoder). These
// are gather-bound, and a LiDAR voxel has only ~4-6 of its 27 neighbours, so a tile kernel that
// walks all 27 offsets wastes >80 % of its work. Here FOUR lanes own one output row (each lane
// C_out/4 channels); the row's present offsets are a 27-bit mask and the lanes walk only the set
// bits, so every group of the warp does useful work in every iteration. All kvol weight slabs
// live in shared memory (slab stride padded by 4 words: groups of a warp read different slabs).
// ---------------------------------------------------------------------------------------------
template <int COUT>
__global__ void __launch_bounds__(256)
sp_conv_rows_kernel(const float* __restrict__ in_feats, int c_in, const float* __restrict__ w,
                    const int* __restrict__ nbr, int n_out, int kvol,
                    const float* __restrict__ scale, const float* __restrict__ shift,
                    const float* __restrict__ residual, int relu, float* __restrict__ out) {
  constexpr int OPL = COUT / 4;  // output channels per lane
  extern __shared__ __align__(16) float W_s[];
  const int slab = c_in * COUT + 4;
  for (int e = threadIdx.x; e < kvol * c_in * (COUT / 4); e += blockDim.x) {
    const int k = e / (c_in * (COUT / 4)), r = e % (c_in * (COUT / 4));
    *reinterpret_cast<float4*>(&W_s[k * slab + r * 4]) =
        __ldg(reinterpret_cast<const float4*>(w + (long long)k * c_in * COUT) + r);
  }
  __syncthreads();
  const int q = threadIdx.x & 3;
  const int group = (blockIdx.x * blockDim.x + threadIdx.x) >> 2;
  const int group_in_warp = (threadIdx.x & 31) >> 2;
  const int n_groups = (gridDim.x * blockDim.x) >> 2;
  const bool vec = (c_in & 3) == 0;
  for (int base = group - group_in_warp; base < n_out; base += n_groups) {
    const int row = base + group_in_warp;
    const bool valid = row < n_out;
    unsigned m = 0;
    if (valid)
      for (int k = q; k < kvol; k += 4)
        if (__ldg(nbr + (long long)k * n_out + row) >= 0) m |= 1u << k;
    m |= __shfl_xor_sync(0xffffffffu, m, 1);
    m |= __shfl_xor_sync(0xffffffffu, m, 2);
    float acc[OPL];
#pragma unroll
    for (int j = 0; j < OPL; ++j) acc[j] = 0.f;
    while (__any_sync(0xffffffffu, m != 0)) {
      if (m != 0) {
        const int k = __ffs(m) - 1;
        m &= m - 1;
        const int idx = __ldg(nbr + (long long)k * n_out + row);
        const float* src = in_feats + (long long)idx * c_in;
        const float* wk = W_s + k * slab + q * OPL;
        if (vec) {
          for (int c4 = 0; c4 < c_in; c4 += 4) {
            const float4 a = __ldg(reinterpret_cast<const float4*>(src + c4));
            const float av[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
            for (int u = 0; u < 4; ++u) {
#pragma unroll
              for (int j = 0; j < OPL; j += 4) {
                const float4 wv = *reinterpret_cast<const float4*>(wk + (c4 + u) * COUT + j);
                acc[j + 0] = fmaf(av[u], wv.x, acc[j + 0]);
                acc[j + 1] = fmaf(av[u], wv.y, acc[j + 1]);
                acc[j + 2] = fmaf(av[u], wv.z, acc[j + 2]);
                acc[j + 3] = fmaf(av[u], wv.w, acc[j + 3]);
              }
            }
          }
        } else {
          for (int ci = 0; ci < c_in; ++ci) {
            const float a = __ldg(src + ci);
#pragma unroll
            for (int j = 0; j < OPL; j += 4) {
              const float4 wv = *reinterpret_cast<const float4*>(wk + ci * COUT + j);
              acc[j + 0] = fmaf(a, wv.x, acc[j + 0]);
              acc[j + 1] = fmaf(a, wv.y, acc[j + 1]);
              acc[j + 2] = fmaf(a, wv.z, acc[j + 2]);
              acc[j + 3] = fmaf(a, wv.w, acc[j + 3]);
            }
          }
        }
      }
    }
    if (valid) {
#pragma unroll
      for (int j = 0; j < OPL; j += 4) {
        const int c = q * OPL + j;
        float4 v = make_float4(acc[j], acc[j + 1], acc[j + 2], acc[j + 3]);
        if (scale) {
          const float4 sc = __ldg(reinterpret_cast<const float4*>(scale + c));
          v.x *= sc.x, v.y *= sc.y, v.z *= sc.z, v.w *= sc.w;
        }
        if (shift) {
          const float4 sh = __ldg(reinterpret_cast<const float4*>(shift + c));
          v.x += sh.x, v.y += sh.y, v.z += sh.z, v.w += sh.w;
        }
        if (residual) {
          const float4 rr = __ldg(reinterpret_cast<const float4*>(residual + (long long)row * COUT + c));
          v.x += rr.x, v.y += rr.y, v.z += rr.z, v.w += rr.w;
        }
        if (relu) v.x = fmaxf(v.x, 0.f), v.y = fmaxf(v.y, 0.f), v.z = fmaxf(v.z, 0.f), v.w = fmaxf(v.w, 0.f);
        *reinterpret_cast<float4*>(out + (long long)row * COUT + c) = v;
      }
    }
  }
}

// dense[b][c*Z + z][y][x] = feats[m][c]. 32 consecutive (sorted) voxels per CTA: rows are read
// coalesced into shared memory, then lane <-> voxel so that neighbouring x land in one sector.
__global__ void __launch_bounds__(256)
sp_dense_kernel(const float* __restrict__ feats, const int4* __restrict__ coors, int m, int C,
                int batch, int Z, int Y, int X, float* __restrict__ dense) {
  extern __shared__ float tile[];  // [32][C + 1]
  __shared__ long long base_s[32];
  const int m0 = blockIdx.x * 32;
  const int ld = C + 1;
  for (int e = threadIdx.x; e < 32 * C; e += blockDim.x) {
    const int r = e / C, c = e % C;
    tile[r * ld + c] = (m0 + r < m) ? feats[(long long)(m0 + r) * C + c] : 0.f;
  }
  if (threadIdx.x < 32) {
    long long b = -1;
    if (m0 + threadIdx.x < m) {
      const int4 c = coors[m0 + threadIdx.x];
      if ((unsigned)c.x < (unsigned)batch && (unsigned)c.y < (unsigned)Z &&
          (unsigned)c.z < (unsigned)Y && (unsigned)c.w < (unsigned)X)
        b = (((long long)c.x * C * Z + c.y) * Y + c.z) * X + c.w;
    }
    base_s[threadIdx.x] = b;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const long long b = base_s[lane];
  const long long cs = (long long)Z * Y * X;
  if (b < 0) return;
  for (int c = wid; c < C; c += 8) dense[b + c * cs] = tile[lane * ld + c];
}

}  // namespace

long long spconv_max_out(long long n_in, const SpConvGeom& g) {
  long long per = 1;
  for (int i = 0; i < 3; ++i) {
    const int cnt = (g.d[i] == 1) ? (g.k[i] + g.s[i] - 1) / g.s[i] : g.k[i];
    per *= cnt < 1 ? 1 : cnt;
  }
  long long cells = (long long)g.batch * g.out_shape[0] * g.out_shape[1] * g.out_shape[2];
  long long bound = n_in * per;
  return bound < cells ? bound : cells;
}

size_t spconv_ws_bytes(long long n_in, long long max_out) {
  const long long n_big = n_in > max_out ? n_in : max_out;
  size_t b = 0;
  b += align_up((size_t)pow2_at_least(2ull * (unsigned long long)n_big) * 8);  // hash keys + vals
  b += align_up((size_t)max_out * 4) * 3;                                      // sort ping-pong
  b += radix_sort_ws_bytes(max_out > 0 ? max_out : 1);
  return b + 4096;
}

static int check_geom(const SpConvGeom& g, const char* what) {
  for (int i = 0; i < 3; ++i) {
    DBEV_CHECK_ARG(g.k[i] >= 1 && g.s[i] >= 1 && g.d[i] >= 1 && g.p[i] >= 0, "%s: bad geometry", what);
    DBEV_CHECK_ARG(g.s[i] == 1 || g.d[i] == 1, "%s: stride and dilation both > 1 (conv.py:92-93)", what);
    DBEV_CHECK_ARG(g.in_shape[i] >= 1 && g.out_shape[i] >= 1, "%s: bad spatial shape", what);
  }
  DBEV_CHECK_ARG(g.batch >= 1, "%s: batch must be >= 1", what);
  DBEV_CHECK_ARG(g.kvol() <= 4096, "%s: kernel volume > 4096 (spconv_ops.h:52)", what);
  const unsigned long long ci =
      (unsigned long long)g.batch * g.in_shape[0] * g.in_shape[1] * g.in_shape[2];
  const unsigned long long co =
      (unsigned long long)g.batch * g.out_shape[0] * g.out_shape[1] * g.out_shape[2];
  DBEV_CHECK_ARG(ci < 0xFFFFFFFFull && co < 0xFFFFFFFFull,
                 "%s: batch*Z*Y*X must stay below 2^32-1 (32-bit cell keys)", what);
  return DBEV_OK;
}

static int build_input_hash(const int* in_coors, int n_in, const SpConvGeom& g, Workspace& w,
                            uint32_t** hkeys_out, int** hvals_out, uint32_t* mask_out,
                            cudaStream_t stream) {
  const uint32_t cap = pow2_at_least(2ull * (unsigned long long)(n_in > 0 ? n_in : 1));
  uint32_t* hkeys = w.take<uint32_t>(cap);
  int* hvals = w.take<int>(cap);
  if (!w.ok()) {
    set_last_error("spconv: workspace too small");
    return DBEV_ERR_WORKSPACE;
  }
  DBEV_CUDA(cudaMemsetAsync(hkeys, 0xFF, (size_t)cap * 4, stream));
  if (n_in > 0) {
    sp_hash_insert_kernel<<<ceil_div(n_in, 256), 256, 0, stream>>>(
        (const int4*)in_coors, n_in, g.batch, g.in_shape[0], g.in_shape[1], g.in_shape[2], hkeys,
        hvals, cap - 1);
    DBEV_CHECK_LAUNCH("sp_hash_insert_kernel");
  }
  *hkeys_out = hkeys, *hvals_out = hvals, *mask_out = cap - 1;
  return DBEV_OK;
}

int spconv_table(const int* in_coors, int n_in, const int* out_coors, int n_out,
                 const SpConvGeom& g, int* nbr, void* ws, size_t ws_bytes, cudaStream_t stream) {
  int rc = check_geom(g, "spconv_table");
  if (rc != DBEV_OK) return rc;
  DBEV_CHECK_ARG(n_in >= 0 && n_out >= 0, "spconv_table: negative count");
  DBEV_CHECK_ARG(((uintptr_t)in_coors & 15) == 0 && ((uintptr_t)out_coors & 15) == 0,
                 "spconv_table: coordinate rows must be 16-byte aligned");
  if (n_out == 0) return DBEV_OK;
  Workspace w(ws, ws_bytes);
  uint32_t* hkeys;
  int* hvals;
  uint32_t mask;
  rc = build_input_hash(in_coors, n_in, g, w, &hkeys, &hvals, &mask, stream);
  if (rc != DBEV_OK) return rc;
  const long long total = (long long)g.kvol() * n_out;
  sp_table_kernel<<<ceil_div(total, 256), 256, 0, stream>>>((const int4*)out_coors, n_out,
                                                            to_dev(g), hkeys, hvals, mask, nbr);
  DBEV_CHECK_LAUNCH("sp_table_kernel");
  return DBEV_OK;
}

int spconv_out_candidates(const int* in_coors, int n_in, const SpConvGeom& g, uint32_t* out_keys,
                          long long max_out, int* n_out_dev, void* ws, size_t ws_bytes,
                          cudaStream_t stream) {
  int rc = check_geom(g, "spconv_out_candidates");
  if (rc != DBEV_OK) return rc;
  DBEV_CHECK_ARG(n_in >= 0 && max_out >= spconv_max_out(n_in, g),
                 "spconv_out_candidates: max_out below dbev_spconv_max_out()");
  DBEV_CUDA(cudaMemsetAsync(n_out_dev, 0, sizeof(int), stream));
  if (n_in == 0) return DBEV_OK;
  Workspace w(ws, ws_bytes);
  const uint32_t cap = pow2_at_least(2ull * (unsigned long long)max_out);
  uint32_t* hset = w.take<uint32_t>(cap);
  if (!w.ok()) {
    set_last_error("spconv_out_candidates: workspace too small");
    return DBEV_ERR_WORKSPACE;
  }
  DBEV_CUDA(cudaMemsetAsync(hset, 0xFF, (size_t)cap * 4, stream));
  int cnt[3];
  for (int i = 0; i < 3; ++i) cnt[i] = g.s[i] == 1 ? g.k[i] : (g.k[i] + g.s[i] - 1) / g.s[i];
  const long long total = (long long)cnt[0] * cnt[1] * cnt[2] * n_in;
  sp_candidates_kernel<<<ceil_div(total, 256), 256, 0, stream>>>(
      (const int4*)in_coors, n_in, to_dev(g), cnt[0], cnt[1], cnt[2], hset, cap - 1, out_keys, max_out,
      n_out_dev);
  DBEV_CHECK_LAUNCH("sp_candidates_kernel");
  return DBEV_OK;
}

int spconv_out_table(const int* in_coors, int n_in, const SpConvGeom& g, uint32_t* out_keys,
                     int n_out, int* out_coors, int* nbr, void* ws, size_t ws_bytes,
                     cudaStream_t stream) {
  int rc = check_geom(g, "spconv_out_table");
  if (rc != DBEV_OK) return rc;
  if (n_out == 0) return DBEV_OK;
  Workspace w(ws, ws_bytes);
  uint32_t* hkeys;
  int* hvals;
  uint32_t mask;
  rc = build_input_hash(in_coors, n_in, g, w, &hkeys, &hvals, &mask, stream);
  if (rc != DBEV_OK) return rc;
  uint32_t* keys1 = w.take<uint32_t>(n_out);
  uint32_t* vals0 = w.take<uint32_t>(n_out);
  uint32_t* vals1 = w.take<uint32_t>(n_out);
  if (!w.ok()) {
    set_last_error("spconv_out_table: workspace too small");
    return DBEV_ERR_WORKSPACE;
  }
  const size_t consumed = align_up(w.used);
  uint32_t* keys[2] = {out_keys, keys1};
  uint32_t* vals[2] = {vals0, vals1};
  int sel = 0;
  const unsigned long long cells =
      (unsigned long long)g.batch * g.out_shape[0] * g.out_shape[1] * g.out_shape[2];
  rc = radix_sort_pairs(keys, vals, true, n_out, bits_for(cells), (char*)ws + consumed,
                        ws_bytes > consumed ? ws_bytes - consumed : 0, stream, &sel);
  if (rc != DBEV_OK) return rc;
  sp_decode_keys_kernel<<<ceil_div(n_out, 256), 256, 0, stream>>>(
      keys[sel], n_out, g.out_shape[0], g.out_shape[1], g.out_shape[2], (int4*)out_coors);
  DBEV_CHECK_LAUNCH("sp_decode_keys_kernel");
  const long long total = (long long)g.kvol() * n_out;
  sp_table_kernel<<<ceil_div(total, 256), 256, 0, stream>>>((const int4*)out_coors, n_out,
                                                            to_dev(g), hkeys, hvals, mask, nbr);
  DBEV_CHECK_LAUNCH("sp_table_kernel");
  return DBEV_OK;
}

int spconv_pairs_from_table(const int* nbr, int kvol, int n_out, int pair_stride,
                            int* indice_pairs, int* indice_num, cudaStream_t stream) {
  DBEV_CHECK_ARG(kvol >= 1 && n_out >= 0 && pair_stride >= 0, "spconv_pairs_from_table: bad sizes");
  if (pair_stride > 0)
    DBEV_CUDA(cudaMemsetAsync(indice_pairs, 0xFF, (size_t)kvol * 2 * pair_stride * 4, stream));
  sp_pairs_from_table_kernel<<<kvol, 1024, 0, stream>>>(nbr, n_out, pair_stride, indice_pairs,
                                                        indice_num);
  DBEV_CHECK_LAUNCH("sp_pairs_from_table_kernel");
  return DBEV_OK;
}

int spconv_table_from_pairs(const int* indice_pairs, const int* indice_num, int kvol,
                            int pair_stride, int inverse, int n_out, int* nbr,
                            cudaStream_t stream) {
  DBEV_CHECK_ARG(kvol >= 1 && n_out >= 0 && pair_stride >= 0, "spconv_table_from_pairs: bad sizes");
  if (n_out == 0) return DBEV_OK;
  DBEV_CUDA(cudaMemsetAsync(nbr, 0xFF, (size_t)kvol * n_out * 4, stream));
  if (pair_stride == 0) return DBEV_OK;
  dim3 grid(ceil_div(pair_stride, 256), kvol);
  sp_table_from_pairs_kernel<<<grid, 256, 0, stream>>>(indice_pairs, indice_num, pair_stride,
                                                       inverse, n_out, nbr);
  DBEV_CHECK_LAUNCH("sp_table_from_pairs_kernel");
  return DBEV_OK;
}

int spconv_forward(const float* in_feats, int c_in, const float* weight, int c_out,
                   const int* nbr, int kvol, int n_out, const float* scale, const float* shift,
                   const float* residual, int relu, float* out, cudaStream_t stream) {
  DBEV_CHECK_ARG(c_in >= 1 && kvol >= 1 && n_out >= 0, "spconv_forward: bad sizes");
  DBEV_CHECK_ARG(c_out == 16 || c_out == 32 || c_out == 64 || c_out == 128,
                 "spconv_forward: c_out must be 16, 32, 64 or 128 (got %d)", c_out);
  DBEV_CHECK_ARG(((uintptr_t)weight & 15) == 0 && ((uintptr_t)out & 15) == 0 &&
                     ((uintptr_t)residual & 15) == 0 && ((uintptr_t)scale & 15) == 0 &&
                     ((uintptr_t)shift & 15) == 0,
                 "spconv_forward: pointers must be 16-byte aligned");
  if (n_out == 0) return DBEV_OK;
  if (kvol <= 32 && (c_out == 16 || c_out == 32) && c_in * c_out <= 512) {
    // narrow, gather-bound layers: lane-group-per-row kernel, persistent grid
    const size_t smem = (size_t)kvol * (c_in * c_out + 4) * sizeof(float);
    int dev = 0, sms = 0;
    DBEV_CUDA(cudaGetDevice(&dev));
    DBEV_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int per_sm = smem > 56 * 1024 ? 2 : (smem > 36 * 1024 ? 4 : 6);
    int grid_r = sms * per_sm;
    const int need = ceil_div((long long)n_out * 4, 256);
    if (grid_r > need) grid_r = need;
    if (c_out == 16) {
      DBEV_CUDA(cudaFuncSetAttribute(sp_conv_rows_kernel<16>,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      sp_conv_rows_kernel<16><<<grid_r, 256, smem, stream>>>(in_feats, c_in, weight, nbr, n_out, kvol,
                                                             scale, shift, residual, relu, out);
    } else {
      DBEV_CUDA(cudaFuncSetAttribute(sp_conv_rows_kernel<32>,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      sp_conv_rows_kernel<32><<<grid_r, 256, smem, stream>>>(in_feats, c_in, weight, nbr, n_out, kvol,
                                                             scale, shift, residual, relu, out);
    }
    DBEV_CHECK_LAUNCH("sp_conv_rows_kernel");
    return DBEV_OK;
  }
  const int grid = ceil_div(n_out, kConvTM);
#define DBEV_SP_LAUNCH(CO)                                                                   \
  sp_conv_fma_kernel<CO><<<grid, 256, 0, stream>>>(in_feats, c_in, weight, nbr, n_out, kvol, \
                                                   scale, shift, residual, relu, out)
  switch (c_out) {
    case 16: DBEV_SP_LAUNCH(16); break;
    case 32: DBEV_SP_LAUNCH(32); break;
    case 64: DBEV_SP_LAUNCH(64); break;
    default: DBEV_SP_LAUNCH(128); break;
  }
#undef DBEV_SP_LAUNCH
  DBEV_CHECK_LAUNCH("sp_conv_fma_kernel");
  return DBEV_OK;
}

int spconv_dense(const float* feats, const int* coors, int m, int C, int batch, int Z, int Y,
                 int X, float* dense, cudaStream_t stream) {
  DBEV_CHECK_ARG(m >= 0 && C >= 1 && batch >= 1 && Z >= 1 && Y >= 1 && X >= 1,
                 "spconv_dense: bad sizes");
  DBEV_CHECK_ARG(((uintptr_t)coors & 15) == 0, "spconv_dense: coors must be 16-byte aligned");
  DBEV_CUDA(cudaMemsetAsync(dense, 0, (size_t)batch * C * Z * Y * X * sizeof(float), stream));
  if (m == 0) return DBEV_OK;
  const size_t smem = (size_t)32 * (C + 1) * sizeof(float);
  DBEV_CHECK_ARG(smem <= 48 * 1024, "spconv_dense: C too large (%d)", C);
  sp_dense_kernel<<<ceil_div(m, 32), 256, smem, stream>>>(feats, (const int4*)coors, m, C, batch,
                                                          Z, Y, X, dense);
  DBEV_CHECK_LAUNCH("sp_dense_kernel");
  return DBEV_OK;
}

}  // namespace dbev
