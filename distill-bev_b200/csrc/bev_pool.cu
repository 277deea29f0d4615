// LSS "splat" (frustum voxel pooling / bev_pool) for B200.
//
// Reference behaviour reproduced here:
//   voxel_pooling        mmdet3d/models/necks/view_transformer_mine.py:141-181
//   QuickCumsum          mmdet3d/models/necks/view_transformer_mine.py:30-56
//   bev_pool (python)    mmdet3d/ops/bev_pool/bev_pool.py:83-97
//   bev_pool_kernel      mmdet3d/ops/bev_pool/src/bev_pool_cuda.cu:20-42
//   bev_pool_grad_kernel mmdet3d/ops/bev_pool/src/bev_pool_cuda.cu:61-84
//
// Design (not a port): the reference sorts the FEATURES (argsort + 3 gathers,
// then cumsum / boolean select / diff / scatter: >= 12 passes over [n, C]).
// Here only 4-byte point ids are sorted (a "plan": order[], cell_start[],
// cell_end[]); the feature rows are then read exactly once, gathered by id in
// 16-byte vectors, summed per BEV cell in a fixed order, transposed through
// shared memory and written exactly once in the caller's final NCHW layout
// (zeros for empty cells included: no zero-fill pass, no permute copy).
// HBM traffic = n*C*4 (rows) + n*4 (ids) + cells*8 + out  ~ algorithmic bytes.
#include "bev_pool.cuh"

#include <cstdlib>

#include <atomic>

#include "sort.cuh"

namespace dbev {

namespace {

constexpr int kPoolBlock = 128;
constexpr int kPoolWarps = kPoolBlock / 32;
constexpr int kTileCells = 32;
constexpr int kTilePitch = kTileCells + 1;
constexpr int kDefaultRowsPerItem = 192;

// ---- plan: cell keys -------------------------------------------------------

// view_transformer_mine.py:150 -> ((geom - (bx - dx/2)) / dx).long(), then the
// in-bounds test of :157-159. trunc-toward-zero semantics: a quotient in
// (-1, 0) lands in cell 0 and is KEPT. For a float q, 0 <= trunc(q) < nx is
// equivalent to q > -1 && q < nx (NaN fails both, as it does in the
// reference where it converts to INT64_MIN).
__device__ __forceinline__ bool geom_to_index(float g, float off, float dx, float nx_f, int* idx) {
  float q = __fdiv_rn(__fsub_rn(g, off), dx);
  if (!(q > -1.0f && q < nx_f)) return false;
  *idx = (int)q;  // cvt.rzi
  return true;
}

__global__ void __launch_bounds__(256)
bev_keys_from_geom_kernel(const float* __restrict__ geom, long long n, long long pts_per_batch,
                          float off0, float off1, float off2, float dx0, float dx1, float dx2,
                          float nxf0, float nxf1, float nxf2, int n0, int n1, int nz,
                          int fast_axis, uint32_t sentinel, uint32_t* __restrict__ keys,
                          int* __restrict__ point_cell, int frames = 1) {
  long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  const float g0 = geom[p * 3 + 0], g1 = geom[p * 3 + 1], g2 = geom[p * 3 + 2];
  int i0, i1, i2;
  bool ok = geom_to_index(g0, off0, dx0, nxf0, &i0);
  ok = geom_to_index(g1, off1, dx1, nxf1, &i1) && ok;
  ok = geom_to_index(g2, off2, dx2, nxf2, &i2) && ok;
  // the float compare is against the float nx (reference); the canvas is sized
  // by nx.to(long). Guard the integer canvas too.
  ok = ok && i0 < n0 && i1 < n1 && i2 < nz;
  uint32_t key = sentinel;
  if (ok) {
    const int b = (int)(p / pts_per_batch);
    const int islow = fast_axis == 0 ? i1 : i0;
    const int ifast = fast_axis == 0 ? i0 : i1;
    const int nslow = fast_axis == 0 ? n1 : n0;
    const int nfast = fast_axis == 0 ? n0 : n1;
    // frames > 1 (BEVDepth4D): sample-frame b = s * frames + f; the frame index becomes the FASTEST part of the cell
    // number, so a channels-last BEV map indexed by it is [s][y][x][f][C] = the frames concatenated along the
    // channels (torch.cat(bev_feat_list, dim=1), bevdet.py:300-320) with no concat pass
    const int smp = b / frames, fr = b - smp * frames;
    key = (uint32_t)((((((long long)smp * nz + i2) * nslow + islow) * nfast + ifast) * frames) + fr);
  }
  if (keys) keys[p] = key;
  if (point_cell) point_cell[p] = ok ? (int)key : -1;
}

template <typename T>
__global__ void __launch_bounds__(256)
bev_keys_from_coords_kernel(const T* __restrict__ coords, long long n, int nb, int n0, int n1,
                            int nz, int fast_axis, uint32_t sentinel,
                            uint32_t* __restrict__ keys) {
  long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  const long long c0 = (long long)coords[p * 4 + 0], c1 = (long long)coords[p * 4 + 1],
                  c2 = (long long)coords[p * 4 + 2], c3 = (long long)coords[p * 4 + 3];
  uint32_t key = sentinel;
  if (c0 >= 0 && c0 < n0 && c1 >= 0 && c1 < n1 && c2 >= 0 && c2 < nz && c3 >= 0 && c3 < nb) {
    const long long islow = fast_axis == 0 ? c1 : c0;
    const long long ifast = fast_axis == 0 ? c0 : c1;
    const int nslow = fast_axis == 0 ? n1 : n0;
    const int nfast = fast_axis == 0 ? n0 : n1;
    key = (uint32_t)(((c3 * nz + c2) * nslow + islow) * nfast + ifast);
  }
  keys[p] = key;
}

// sorted keys -> [start, end) of every cell; slot ncells is the dropped tail.
__global__ void __launch_bounds__(256)
bev_bounds_kernel(const uint32_t* __restrict__ skeys, long long n, int* __restrict__ cell_start,
                  int* __restrict__ cell_end) {
  long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const uint32_t k = skeys[j];
  if (j == 0 || skeys[j - 1] != k) cell_start[k] = (int)j;
  if (j == n - 1 || skeys[j + 1] != k) cell_end[k] = (int)(j + 1);
}

// ---- pooling ---------------------------------------------------------------

struct PoolGeom {
  int C;
  int nfast, nslow, nbz;  // nbz = B * nz
  int nz;
  int tiles_per_row;
  long long sB, sZ, sC;  // output strides (elements) of batch, z and channel
};

__device__ __forceinline__ void f4_add(float4& a, const float4& b) {
  a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
}

// ---- work items --------------------------------------------------------------
// BEV occupancy is extremely skewed (cells next to the ego vehicle collect
// hundreds of frustum points, far cells one or two), so a tile of 32 cells is
// split into "items": runs of consecutive cells holding at most ~rows_per_item
// rows (a single heavier cell is an item of its own). Items are part of the
// plan (geometry only) and carry their row range, so a consumer needs no
// dependent lookups: item -> {order[row_lo..row_hi), 32 cell bounds} -> rows.
//
// ONE WARP owns one item at a time (grid-stride over the item list) and never
// synchronises with other warps: no block barriers, every warp is at a
// different phase so load latency of one overlaps the epilogue of another.
// Inside the warp the item's rows are divided evenly over "workers" (groups of
// LPR lanes, LPR*4 = channel block <= 64), so each lane group streams the same
// number of rows no matter how they fall into cells; U rows per worker are in
// flight at a time (16-byte vector loads, L1 bypassed).

__device__ __forceinline__ int tile_ncell(int nfast, int ftile) {
  return min(kTileCells, nfast - ftile * kTileCells);
}

// groups of one tile: greedy split, a new group starts when the next cell
// would push a non-empty group beyond rows_per_item. counts != null: count;
// items != null: emit {tile, c0 | c1 << 8, row_lo, row_hi}.
__global__ void __launch_bounds__(256)
bev_items_kernel(const int* __restrict__ cell_start, const int* __restrict__ cell_end,
                 long long ntiles, int nfast, int tiles_per_row, int rows_per_item,
                 int* __restrict__ counts, const int* __restrict__ offsets,
                 int4* __restrict__ items, int max_items) {
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= ntiles) return;
  const int ftile = (int)(t % tiles_per_row);
  const long long rowid = t / tiles_per_row;
  const int ncell = tile_ncell(nfast, ftile);
  const long long cell0 = rowid * nfast + (long long)ftile * kTileCells;
  int groups = 0, c0 = 0, rows = 0, lo = 0, hi = 0;
  const int base = offsets ? offsets[t] : 0;
  for (int c = 0; c < ncell; ++c) {
    const int s = cell_start[cell0 + c], e = cell_end[cell0 + c];
    const int len = e - s;
    if (rows > 0 && rows + len > rows_per_item) {
      if (items && base + groups < max_items)
        items[base + groups] = make_int4((int)t, c0 | (c << 8), lo, hi);
      ++groups;
      c0 = c;
      rows = 0;
    }
    if (len > 0) {
      if (rows == 0) lo = s;
      hi = e;
      rows += len;
    }
  }
  if (rows == 0) lo = hi = 0;
  if (items && base + groups < max_items)
    items[base + groups] = make_int4((int)t, c0 | (ncell << 8), lo, hi);
  ++groups;
  if (counts) counts[t] = groups;
}

struct ItemCtx {
  long long cell0;   // first cell id of the tile
  long long obase;   // output offset of (tile, channel 0, cell 0)
  int c0, c1, row_lo, row_hi;
};

__device__ __forceinline__ ItemCtx decode_item(const int4& item, const PoolGeom& g) {
  ItemCtx ic;
  const long long t = item.x;
  ic.c0 = item.y & 0xff;
  ic.c1 = item.y >> 8;
  ic.row_lo = item.z;
  ic.row_hi = item.w;
  const int ftile = (int)(t % g.tiles_per_row);
  const long long rowid = t / g.tiles_per_row;
  const int f0 = ftile * kTileCells;
  ic.cell0 = rowid * g.nfast + f0;
  const long long bz = rowid / g.nslow;
  const int islow = (int)(rowid % g.nslow);
  ic.obase = (bz / g.nz) * g.sB + (bz % g.nz) * g.sZ + (long long)islow * g.nfast + f0;
  return ic;
}

// LIFT = true fuses the "lift" (view_transformer_mine.py:333-335): the row of point p is
// depth[p] * feat[pixel(p), :] computed on the fly from depth[BN, D, fH, fW] (flat index = p)
// and channels-last feat[BN * fH * fW, C]; the [B,N,D,fH,fW,C] volume never exists.
struct LiftArgs {
  const float* depth;
  FastDiv dfhw;  // D * fH * fW
  FastDiv fhw;   // fH * fW
};

// Dynamic work distribution for the persistent gather kernels: warps pull item indices from a
// device counter (items differ 100x in row count, a static round-robin leaves a long tail). Each
// launch takes the next of kSchedSlots counter pairs {next item, warps done}; the last warp to
// finish re-zeroes its pair, so a slot is clean again long before the host wraps around to it.
constexpr int kEagerSlots = 256;      // round-robin for eager launches
constexpr int kSchedSlots = 4096;     // the rest: one slot per launch recorded into a CUDA graph, never reused
__device__ int g_sched[kSchedSlots * 2];

struct WorkQueue {
  int* ctr;
  __device__ __forceinline__ explicit WorkQueue(int slot) : ctr(g_sched + 2 * slot) {}
  __device__ __forceinline__ long long next(int lane) const {
    int v = 0;
    if (lane == 0) v = atomicAdd(ctr, 1);
    return __shfl_sync(0xffffffffu, v, 0);
  }
  __device__ __forceinline__ void finish(int lane, int total_warps) const {
    if (lane == 0 && atomicAdd(ctr + 1, 1) == total_warps - 1) {
      ctr[0] = 0;
      ctr[1] = 0;
    }
  }
};

// MINB = CTAs per SM the register allocation must allow. The materialised-row variant (LIFT =
// false) is latency-bound and gains from the 6th resident CTA (80 instead of 96 registers, a
// 60-byte spill outside the row loop): 0.2355 -> 0.2295 ms on the configs[1] shape.
// CL = channels-last output (out[cell][C], g.sC == 1): a finished cell is ONE 16-byte store per lane straight from the
// accumulator (a coalesced row) - no shared-memory tile, no transposing epilogue; what the student BEV encoder's NHWC
// conv kernels read, and the layout in which the gather reaches its best fraction of the HBM roofline.
template <int LPR, bool LIFT, int MINB = 1, bool CL = false>
__global__ void __launch_bounds__(kPoolBlock, MINB)
bev_pool_gather_fwd_kernel(const float* __restrict__ x, const uint32_t* __restrict__ order,
                           const int* __restrict__ cell_start, const int* __restrict__ cell_end,
                           const int4* __restrict__ items, const int* __restrict__ n_items_ptr,
                           float* __restrict__ out, PoolGeom g, LiftArgs la, int sched_slot) {
  extern __shared__ float smem[];
  constexpr int NW = 32 / LPR;   // workers per warp
  constexpr int CB = LPR * 4;    // channels per block
  constexpr int U = 8;           // rows in flight per worker
  __shared__ int cstart_s[kPoolWarps][kTileCells], cend_s[kPoolWarps][kTileCells];
  __shared__ int side_cell_s[kPoolWarps][NW * 2];

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int sub = lane % LPR, worker = lane / LPR;
  float* tile = smem + (size_t)warp * ((CL ? 0 : CB * kTilePitch) + NW * 2 * CB);  // [CB][kTilePitch] (not in CL mode)
  float* side = tile + (CL ? 0 : CB * kTilePitch);                                  // [NW*2][CB]
  int* cs = cstart_s[warp];
  int* ce = cend_s[warp];
  int* side_cell = side_cell_s[warp];

  const int nblocks = (g.C + CB - 1) / CB;
  const long long n_virtual = (long long)(*n_items_ptr) * nblocks;
  const WorkQueue queue(sched_slot);

  for (;;) {
    const long long vit = queue.next(lane);
    if (vit >= n_virtual) break;
    const int4 item = items[vit / nblocks];
    const int cb = (int)(vit % nblocks) * CB;  // first channel of this block
    const ItemCtx ic = decode_item(item, g);
    const bool lane_active = cb + sub * 4 < g.C;

    __syncwarp();  // previous item's epilogue has finished with this warp's smem
    {
      const bool in = lane >= ic.c0 && lane < ic.c1;
      cs[lane] = in ? cell_start[ic.cell0 + lane] : 0;
      ce[lane] = in ? cell_end[ic.cell0 + lane] : 0;
      for (int i = lane; i < NW * 2; i += 32) side_cell[i] = -1;
    }
    __syncwarp();

    const int len = ic.row_hi - ic.row_lo;
    const int per = (len + NW - 1) / NW;
    const int ws = min(ic.row_lo + worker * per, ic.row_hi);
    const int we = min(ws + per, ic.row_hi);

    if (ws < we) {
      int cur = ic.c0;
      while (ce[cur] <= ws) ++cur;  // cell that owns row ws (empty cells have end 0)
      int cur_end = ce[cur];
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      const float* xb = x + cb + sub * 4;

      auto flush = [&](int cell) {
        const bool complete = cs[cell] >= ws && ce[cell] <= we;
        float* dst;
        int pitch;
        if (CL && complete) {
          if (lane_active)
            *reinterpret_cast<float4*>(out + (ic.cell0 + cell) * g.C + cb + sub * 4) = acc;
          return;
        }
        if (complete) {
          dst = tile + cell;
          pitch = kTilePitch;
        } else {
          const int slot = worker * 2 + ((cs[cell] < ws) ? 0 : 1);
          dst = side + slot * CB;
          pitch = 1;
          if (sub == 0) side_cell[slot] = cell;
        }
        const int c = sub * 4;
        dst[(c + 0) * pitch] = acc.x;
        dst[(c + 1) * pitch] = acc.y;
        dst[(c + 2) * pitch] = acc.z;
        dst[(c + 3) * pitch] = acc.w;
      };

      if constexpr (!LIFT) {
        // streaming gather of materialised rows (HBM-bound): every lane keeps its own point ids,
        // fewest registers -> most warps in flight
        // software pipeline: the point ids of batch i+1 are fetched while the rows of
        // batch i are in flight, so only one global latency per batch is exposed
        uint32_t pid[U];
  #pragma unroll
        for (int u = 0; u < U; ++u) pid[u] = (ws + u < we) ? __ldg(order + ws + u) : 0u;
        for (int pos = ws; pos < we; pos += U) {
          float4 v[U];
  #pragma unroll
          for (int u = 0; u < U; ++u) {
            if (lane_active && pos + u < we) {
              v[u] = ld_stream_f4(xb + (size_t)pid[u] * g.C);
            }
          }
  #pragma unroll
          for (int u = 0; u < U; ++u) pid[u] = (pos + U + u < we) ? __ldg(order + pos + U + u) : 0u;
  #pragma unroll
          for (int u = 0; u < U; ++u) {
            const int r = pos + u;
            if (r < we) {
              if (r >= cur_end) {
                flush(cur);
                ++cur;
                while (ce[cur] <= r) ++cur;
                cur_end = ce[cur];
                acc = make_float4(0.f, 0.f, 0.f, 0.f);
              }
              if (lane_active) f4_add(acc, v[u]);
            }
          }
        }
      } else {
        // Cooperative addressing: per batch of RB rows every lane of the worker resolves ONE row
        // (point id -> source row, + depth weight when LIFT) and the 16 lanes then exchange them
        // with shuffles, instead of all lanes redoing the same integer work per row. Software
        // pipeline: ids are fetched two batches ahead, offsets / depth weights one batch ahead.
        constexpr int K = (LPR >= 8) ? 1 : 8 / LPR;  // rows resolved per lane and batch
        constexpr int RB = LPR * K;                   // rows per batch (8 or 16)
        const uint32_t c4 = (uint32_t)(g.C >> 2);
        const float4* xb4 = reinterpret_cast<const float4*>(xb);
        const unsigned wmask = (LPR == 32) ? 0xffffffffu : (((1u << LPR) - 1u) << (worker * LPR));
        auto load_ids = [&](int base, uint32_t (&pp)[K]) {
  #pragma unroll
          for (int k = 0; k < K; ++k) {
            const int r = base + k * LPR + sub;
            pp[k] = (r < we) ? __ldg(order + r) : 0u;
          }
        };
        auto resolve = [&](const uint32_t (&pp)[K], int base, uint32_t (&off)[K], float (&wg)[K]) {
  #pragma unroll
          for (int k = 0; k < K; ++k) {
            if (LIFT) {
              const uint32_t bn = la.dfhw.div(pp[k]);
              const uint32_t pix = la.fhw.mod(pp[k] - bn * la.dfhw.d);
              off[k] = bn * la.fhw.d + pix;  // row of channels-last feat
              wg[k] = (base + k * LPR + sub < we) ? __ldg(la.depth + pp[k]) : 0.f;
            } else {
              off[k] = pp[k];
              wg[k] = 0.f;
            }
          }
        };
        uint32_t pid[K], off_c[K];
        float w_c[K];
        load_ids(ws, pid);
        resolve(pid, ws, off_c, w_c);
        load_ids(ws + RB, pid);
        for (int pos = ws; pos < we; pos += RB) {
          uint32_t off_n[K];
          float w_n[K];
          resolve(pid, pos + RB, off_n, w_n);
          load_ids(pos + 2 * RB, pid);
  #pragma unroll
          for (int j0 = 0; j0 < RB; j0 += U) {
            float4 v[U];
  #pragma unroll
            for (int u = 0; u < U; ++u) {
              const int j = j0 + u;
              const uint32_t o = __shfl_sync(wmask, off_c[j / LPR], j % LPR, LPR);
              float w = 0.f;
              if (LIFT) w = __shfl_sync(wmask, w_c[j / LPR], j % LPR, LPR);
              if (lane_active && pos + j < we) {
                if (LIFT) {
                  const float4 f = __ldg(xb4 + (size_t)o * c4);
                  // rounded product first, like the reference's materialised volume (no FMA contraction)
                  v[u] = make_float4(__fmul_rn(w, f.x), __fmul_rn(w, f.y), __fmul_rn(w, f.z), __fmul_rn(w, f.w));
                } else {
                  v[u] = ld_stream_f4(reinterpret_cast<const float*>(xb4 + (size_t)o * c4));
                }
              }
            }
  #pragma unroll
            for (int u = 0; u < U; ++u) {
              const int r = pos + j0 + u;
              if (r < we) {
                if (r >= cur_end) {
                  flush(cur);
                  ++cur;
                  while (ce[cur] <= r) ++cur;
                  cur_end = ce[cur];
                  acc = make_float4(0.f, 0.f, 0.f, 0.f);
                }
                if (lane_active) f4_add(acc, v[u]);
              }
            }
          }
  #pragma unroll
          for (int k = 0; k < K; ++k) {
            off_c[k] = off_n[k];
            w_c[k] = w_n[k];
          }
        }
      }
      flush(cur);
    }
    __syncwarp();
    if constexpr (CL) {
      // boundary partial sums (a cell shared by the workers has its partials in consecutive slots, added in slot
      // order: reproducible), written as whole rows; then zero rows for the empty cells of the item
      const int cmax = min(CB, g.C - cb);
      float a0 = 0.f, a1 = 0.f;
      int prev = -1;
      for (int w = 0; w <= NW * 2; ++w) {
        const int cell = w < NW * 2 ? side_cell[w] : -2;
        if (cell == -1) continue;
        if (cell != prev && prev >= 0) {
          float* orow = out + (ic.cell0 + prev) * g.C + cb;
          if (lane < cmax) orow[lane] = a0;
          if (lane + 32 < cmax) orow[lane + 32] = a1;
          if (CB > 64)
            for (int c = lane + 64; c < cmax; c += 32) {
              float t = 0.f;
              for (int v = 0; v < NW * 2; ++v)
                if (side_cell[v] == prev) t += side[v * CB + c];
              orow[c] = t;
            }
          a0 = a1 = 0.f;
        }
        if (cell >= 0) {
          if (lane < CB) a0 += side[w * CB + lane];
          if (lane + 32 < CB) a1 += side[w * CB + lane + 32];
        }
        prev = cell;
      }
      for (int cell = ic.c0 + worker; cell < ic.c1; cell += NW)
        if (ce[cell] <= cs[cell] && lane_active)
          *reinterpret_cast<float4*>(out + (ic.cell0 + cell) * g.C + cb + sub * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
    } else {
      // boundary partial sums: a cell shared by several workers has its partials in
      // consecutive slots; first assigns, the rest add, in slot order (reproducible)
      {
        int prev = -1;
        for (int w = 0; w < NW * 2; ++w) {
          const int cell = side_cell[w];
          if (cell >= 0) {
            for (int c = lane; c < CB; c += 32) {
              const float v = side[w * CB + c];
              if (cell == prev) tile[c * kTilePitch + cell] += v;
              else tile[c * kTilePitch + cell] = v;
            }
            prev = cell;
          }
        }
      }
      __syncwarp();
      // epilogue: lanes = (cell, channel group); empty cells are written as zero here
      {
        const int span = ic.c1 - ic.c0;
        int npow = 1;
        while (npow < span) npow <<= 1;
        const int cpar = 32 / npow;
        const int cell = ic.c0 + lane % npow;
        const bool wr = cell < ic.c1;
        const bool empty = wr ? (ce[cell] <= cs[cell]) : true;
        float* ob = out + ic.obase + cell;
        const int cmax = min(CB, g.C - cb);
        for (int c = lane / npow; c < cmax; c += cpar)
          if (wr) ob[(long long)(cb + c) * g.sC] = empty ? 0.f : tile[c * kTilePitch + cell];
      }
    }
    }
  queue.finish(lane, gridDim.x * kPoolWarps);
}

// scalar fallback for C % 4 != 0 (one row per warp step, lanes over channels)
__global__ void __launch_bounds__(kPoolBlock)
bev_pool_gather_fwd_generic_kernel(const float* __restrict__ x, const uint32_t* __restrict__ order,
                                   const int* __restrict__ cell_start,
                                   const int* __restrict__ cell_end, float* __restrict__ out,
                                   PoolGeom g) {
  extern __shared__ float tile[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int ftile = blockIdx.x % g.tiles_per_row;
  const long long rowid = blockIdx.x / g.tiles_per_row;
  const int f0 = ftile * kTileCells;
  const int ncell = min(kTileCells, g.nfast - f0);
  const long long cell0 = rowid * g.nfast + f0;
  for (int ci = warp; ci < ncell; ci += kPoolWarps) {
    const int start = cell_start[cell0 + ci], end = cell_end[cell0 + ci];
    for (int cb = 0; cb < g.C; cb += 32) {
      const int c = cb + lane;
      float acc = 0.f;
      for (int i = start; i < end; ++i) {
        const uint32_t p = order[i];
        if (c < g.C) acc += x[(size_t)p * g.C + c];
      }
      if (c < g.C) tile[c * kTilePitch + ci] = acc;
    }
  }
  __syncthreads();
  const long long bz = rowid / g.nslow;
  const int islow = (int)(rowid % g.nslow);
  const long long b = bz / g.nz, iz = bz % g.nz;
  float* obase = out + b * g.sB + iz * g.sZ + (long long)islow * g.nfast + f0;
  for (int c = warp; c < g.C; c += kPoolWarps)
    if (lane < ncell) obase[c * g.sC + lane] = tile[c * kTilePitch + lane];
}

// backward of the gather pool: every point of a cell receives the cell's
// gradient row (bev_pool_cuda.cu:61-84; QuickCumsum.backward
// view_transformer_mine.py:48-56). Same items / warp-autonomous workers as
// the forward; rows of dropped points are written as zero by
// bev_pool_zero_tail_kernel, so x_grad needs no prior memset.
template <int LPR>
__global__ void __launch_bounds__(kPoolBlock)
bev_pool_gather_bwd_kernel(const float* __restrict__ out_grad, const uint32_t* __restrict__ order,
                           const int* __restrict__ cell_start, const int* __restrict__ cell_end,
                           const int4* __restrict__ items, const int* __restrict__ n_items_ptr,
                           float* __restrict__ x_grad, PoolGeom g, int sched_slot) {
  extern __shared__ float smem[];
  constexpr int NW = 32 / LPR;
  constexpr int CB = LPR * 4;
  __shared__ int cend_s[kPoolWarps][kTileCells];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int sub = lane % LPR, worker = lane / LPR;
  float* tile = smem + (size_t)warp * (CB * kTilePitch);
  int* ce = cend_s[warp];
  const int nblocks = (g.C + CB - 1) / CB;
  const long long n_virtual = (long long)(*n_items_ptr) * nblocks;
  const WorkQueue queue(sched_slot);

  for (;;) {
    const long long vit = queue.next(lane);
    if (vit >= n_virtual) break;
    const int4 item = items[vit / nblocks];
    const int cb = (int)(vit % nblocks) * CB;
    const ItemCtx ic = decode_item(item, g);
    if (ic.row_hi <= ic.row_lo) continue;  // nothing to write for an all-empty item
    const bool lane_active = cb + sub * 4 < g.C;

    __syncwarp();
    {
      const bool in = lane >= ic.c0 && lane < ic.c1;
      ce[lane] = in ? cell_end[ic.cell0 + lane] : 0;
      const int span = ic.c1 - ic.c0;
      int npow = 1;
      while (npow < span) npow <<= 1;
      const int cpar = 32 / npow;
      const int cell = ic.c0 + lane % npow;
      const float* gb = out_grad + ic.obase + cell;
      const int cmax = min(CB, g.C - cb);
      if (cell < ic.c1)
        for (int c = lane / npow; c < cmax; c += cpar)
          tile[c * kTilePitch + cell] = gb[(long long)(cb + c) * g.sC];
    }
    __syncwarp();

    const int len = ic.row_hi - ic.row_lo;
    const int per = (len + NW - 1) / NW;
    const int ws = min(ic.row_lo + worker * per, ic.row_hi);
    const int we = min(ws + per, ic.row_hi);
    if (ws < we) {
      int cur = ic.c0;
      while (ce[cur] <= ws) ++cur;
      int cur_end = ce[cur];
      float4 val;
      auto load_cell = [&](int cell) {
        const int c = sub * 4;
        val = make_float4(tile[(c + 0) * kTilePitch + cell], tile[(c + 1) * kTilePitch + cell],
                          tile[(c + 2) * kTilePitch + cell], tile[(c + 3) * kTilePitch + cell]);
      };
      load_cell(cur);
      float* xb = x_grad + cb + sub * 4;
      constexpr int U = 8;
      uint32_t pid[U];
#pragma unroll
      for (int u = 0; u < U; ++u) pid[u] = (ws + u < we) ? __ldg(order + ws + u) : 0u;
      for (int pos = ws; pos < we; pos += U) {
        uint32_t cur_pid[U];
#pragma unroll
        for (int u = 0; u < U; ++u) cur_pid[u] = pid[u];
#pragma unroll
        for (int u = 0; u < U; ++u) pid[u] = (pos + U + u < we) ? __ldg(order + pos + U + u) : 0u;
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int r = pos + u;
          if (r < we) {
            if (r >= cur_end) {
              ++cur;
              while (ce[cur] <= r) ++cur;
              cur_end = ce[cur];
              load_cell(cur);
            }
            if (lane_active) st_stream_f4(xb + (size_t)cur_pid[u] * g.C, val);
          }
        }
      }
    }
  }
  queue.finish(lane, gridDim.x * kPoolWarps);
}

__global__ void __launch_bounds__(kPoolBlock)
bev_pool_gather_bwd_generic_kernel(const float* __restrict__ out_grad,
                                   const uint32_t* __restrict__ order,
                                   const int* __restrict__ cell_start,
                                   const int* __restrict__ cell_end, float* __restrict__ x_grad,
                                   PoolGeom g) {
  extern __shared__ float tile[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int ftile = blockIdx.x % g.tiles_per_row;
  const long long rowid = blockIdx.x / g.tiles_per_row;
  const int f0 = ftile * kTileCells;
  const int ncell = min(kTileCells, g.nfast - f0);
  const long long cell0 = rowid * g.nfast + f0;
  const long long bz = rowid / g.nslow;
  const int islow = (int)(rowid % g.nslow);
  const long long b = bz / g.nz, iz = bz % g.nz;
  const float* gbase = out_grad + b * g.sB + iz * g.sZ + (long long)islow * g.nfast + f0;
  for (int c = warp; c < g.C; c += kPoolWarps)
    if (lane < ncell) tile[c * kTilePitch + lane] = gbase[c * g.sC + lane];
  __syncthreads();
  for (int ci = warp; ci < ncell; ci += kPoolWarps) {
    const int start = cell_start[cell0 + ci], end = cell_end[cell0 + ci];
    for (int i = start; i < end; ++i) {
      const uint32_t p = order[i];
      for (int c = lane; c < g.C; c += 32) x_grad[(size_t)p * g.C + c] = tile[c * kTilePitch + ci];
    }
  }
}

// out[b][c][r] = in[b][r][c]  (32x32 tiles through shared memory, both sides coalesced)
__global__ void __launch_bounds__(256)
transpose_batched_kernel(const float* __restrict__ in, float* __restrict__ out, int rows, int cols) {
  __shared__ float t[32][33];
  const size_t boff = (size_t)blockIdx.z * rows * cols;
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int r = r0 + ty + k * 8, c = c0 + tx;
    if (r < rows && c < cols) t[ty + k * 8][tx] = in[boff + (size_t)r * cols + c];
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int c = c0 + ty + k * 8, r = r0 + tx;
    if (r < rows && c < cols) out[boff + (size_t)c * rows + r] = t[tx][ty + k * 8];
  }
}

// backward of the fused lift+splat. One group of LPR lanes per image pixel
// (bn, h, w) walks the D depth bins of its ray:
//   d_depth[p]     = <feat[pix, :], g[cell(p), :]>
//   d_feat[pix, :] = sum_d depth[p] * g[cell(p), :]
// g is the BEV gradient in cells-major layout [n_cells, C] (transposed once),
// so every access is a contiguous row; everything is read from L2.
// Cooperative addressing as in the forward: per batch of DB depth bins each lane of the group
// fetches (cell, depth weight) of ITS bins and the lanes exchange them with shuffles; the per-bin
// dot products are reduced with one transposed butterfly per batch (DB - K shuffles instead of
// DB * log2(LPR)), which leaves bin `sub * K + k` on lane `sub` - the lane that then stores it.
template <int LPR>
__global__ void __launch_bounds__(256)
lift_splat_bwd_kernel(const float* __restrict__ g_cl, const float* __restrict__ depth,
                      const float* __restrict__ feat, const int* __restrict__ point_cell,
                      long long n_pix, int C, int D, int fhw, float* __restrict__ d_depth,
                      float* __restrict__ d_feat) {
  constexpr int CB = LPR * 4;
  constexpr int GPB = 256 / LPR;                 // pixel groups per CTA
  constexpr int K = (LPR >= 8) ? 1 : 8 / LPR;    // bins owned per lane and batch
  constexpr int DB = LPR * K;                    // depth bins per batch (8, 16 or 32)
  constexpr int U = 8;                           // gradient rows in flight per group
  const int sub = threadIdx.x % LPR;
  const long long pixrow = (long long)blockIdx.x * GPB + threadIdx.x / LPR;
  const bool valid = pixrow < n_pix;  // whole groups are valid or not; shuffles stay in-group
  const long long bn = valid ? pixrow / fhw : 0;
  const int pix = valid ? (int)(pixrow % fhw) : 0;
  const int nblocks = (C + CB - 1) / CB;
  const unsigned gmask = (LPR == 32) ? 0xffffffffu
                                     : (((1u << LPR) - 1u) << ((threadIdx.x & 31) / LPR * LPR));
  const size_t p0 = (size_t)bn * D * fhw + pix;  // point id of depth bin 0 of this ray
  const float4* g4 = reinterpret_cast<const float4*>(g_cl);
  const uint32_t c4 = (uint32_t)(C >> 2);
  for (int cbk = 0; cbk < nblocks; ++cbk) {
    const int c = cbk * CB + sub * 4;
    const bool act = valid && c < C;
    float4 f = make_float4(0.f, 0.f, 0.f, 0.f), acc = f;
    if (act) f = __ldg(reinterpret_cast<const float4*>(feat + (size_t)pixrow * C + c));
    for (int d0 = 0; d0 < D; d0 += DB) {
      int cellv[K];
      float wv[K];
#pragma unroll
      for (int k = 0; k < K; ++k) {
        const int d = d0 + sub * K + k;
        const bool in = valid && d < D;
        cellv[k] = in ? __ldg(point_cell + p0 + (size_t)d * fhw) : -1;
        wv[k] = in ? __ldg(depth + p0 + (size_t)d * fhw) : 0.f;
      }
      float dot[DB];
#pragma unroll
      for (int j0 = 0; j0 < DB; j0 += U) {
        float4 gv[U];
        float wj[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int j = j0 + u;
          const int cj = __shfl_sync(gmask, cellv[j % K], j / K, LPR);
          wj[u] = __shfl_sync(gmask, wv[j % K], j / K, LPR);
          gv[u] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (cj >= 0 && act) gv[u] = __ldg(g4 + (size_t)cj * c4 + (c >> 2));
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          acc.x += wj[u] * gv[u].x; acc.y += wj[u] * gv[u].y;
          acc.z += wj[u] * gv[u].z; acc.w += wj[u] * gv[u].w;
          dot[j0 + u] = (f.x * gv[u].x + f.y * gv[u].y) + (f.z * gv[u].z + f.w * gv[u].w);
        }
      }
      // transposed butterfly: after the step with distance s, lanes with bit s set keep the
      // upper half of the remaining bins (same pairing order as a per-bin xor butterfly)
#pragma unroll
      for (int s = LPR / 2, half = DB / 2; s > 0; s >>= 1, half >>= 1) {
        const bool up = (sub & s) != 0;
#pragma unroll
        for (int i = 0; i < DB / 2; ++i) {
          if (i < half) {
            const float send = up ? dot[i] : dot[i + half];
            const float keep = up ? dot[i + half] : dot[i];
            dot[i] = keep + __shfl_xor_sync(gmask, send, s);
          }
        }
      }
#pragma unroll
      for (int k = 0; k < K; ++k) {
        const int d = d0 + sub * K + k;
        if (valid && d < D) {
          float* dd = d_depth + p0 + (size_t)d * fhw;
          if (cbk == 0) *dd = dot[k];
          else *dd += dot[k];  // same lane wrote it in the previous channel block
        }
      }
    }
    if (act) *reinterpret_cast<float4*>(d_feat + (size_t)pixrow * C + c) = acc;
  }
}

// rows of points that fell outside the grid get a zero gradient
__global__ void __launch_bounds__(256)
bev_pool_zero_tail_kernel(const uint32_t* __restrict__ order, const int* __restrict__ tail_start,
                          const int* __restrict__ tail_end, float* __restrict__ x_grad, int C) {
  const int start = *tail_start, end = *tail_end;
  const int warps_total = gridDim.x * (blockDim.x >> 5);
  const int gw = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  for (int i = start + gw; i < end; i += warps_total) {
    float* row = x_grad + (size_t)order[i] * C;
    for (int c = lane; c < C; c += 32) row[c] = 0.f;
  }
}

// ---- reference-ABI kernels (pre-sorted rows + interval lists) ---------------

// out is [b, d, h, w, c] channels-last exactly as bev_pool_cuda.cu:32-34
// addresses it; one warp per interval streams the interval's contiguous rows.
template <int LPR, int CHUNKS>
__global__ void __launch_bounds__(kPoolBlock)
bev_pool_interval_fwd_kernel(int d, int h, int w, int c, int n_intervals,
                             const float* __restrict__ x, const int* __restrict__ geom,
                             const int* __restrict__ istart, const int* __restrict__ ilen,
                             float* __restrict__ out) {
  constexpr int RPS = 32 / LPR;
  const int lane = threadIdx.x & 31;
  const int sub = lane % LPR, rslot = lane / LPR;
  const int iv = blockIdx.x * kPoolWarps + (threadIdx.x >> 5);
  if (iv >= n_intervals) return;
  const int start = istart[iv], len = ilen[iv];
  float4 acc[CHUNKS];
  bool lane_active[CHUNKS];
#pragma unroll
  for (int k = 0; k < CHUNKS; ++k) {
    acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    lane_active[k] = (k * LPR + sub) * 4 < c;
  }
  const float* xb = x + (size_t)start * c + sub * 4;
#pragma unroll 4
  for (int j = rslot; j < len; j += RPS) {
#pragma unroll
    for (int k = 0; k < CHUNKS; ++k)
      if (lane_active[k]) f4_add(acc[k], ld_stream_f4(xb + (size_t)j * c + k * LPR * 4));
  }
#pragma unroll
  for (int o = LPR; o < 32; o <<= 1) {
#pragma unroll
    for (int k = 0; k < CHUNKS; ++k) {
      acc[k].x += __shfl_xor_sync(0xffffffffu, acc[k].x, o);
      acc[k].y += __shfl_xor_sync(0xffffffffu, acc[k].y, o);
      acc[k].z += __shfl_xor_sync(0xffffffffu, acc[k].z, o);
      acc[k].w += __shfl_xor_sync(0xffffffffu, acc[k].w, o);
    }
  }
  if (rslot == 0) {
    const int* gf = geom + (size_t)start * 4;
    float* o = out + (((size_t)gf[3] * d + gf[2]) * h + gf[0]) * (size_t)w * c + (size_t)gf[1] * c +
               sub * 4;
#pragma unroll
    for (int k = 0; k < CHUNKS; ++k)
      if (lane_active[k]) *reinterpret_cast<float4*>(o + k * LPR * 4) = acc[k];
  }
}

template <int LPR, int CHUNKS>
__global__ void __launch_bounds__(kPoolBlock)
bev_pool_interval_bwd_kernel(int d, int h, int w, int c, int n_intervals,
                             const float* __restrict__ out_grad, const int* __restrict__ geom,
                             const int* __restrict__ istart, const int* __restrict__ ilen,
                             float* __restrict__ x_grad) {
  constexpr int RPS = 32 / LPR;
  const int lane = threadIdx.x & 31;
  const int sub = lane % LPR, rslot = lane / LPR;
  const int iv = blockIdx.x * kPoolWarps + (threadIdx.x >> 5);
  if (iv >= n_intervals) return;
  const int start = istart[iv], len = ilen[iv];
  const int* gf = geom + (size_t)start * 4;
  const float* o = out_grad + (((size_t)gf[3] * d + gf[2]) * h + gf[0]) * (size_t)w * c +
                   (size_t)gf[1] * c + sub * 4;
  float4 val[CHUNKS];
  bool lane_active[CHUNKS];
#pragma unroll
  for (int k = 0; k < CHUNKS; ++k) {
    lane_active[k] = (k * LPR + sub) * 4 < c;
    if (lane_active[k]) val[k] = *reinterpret_cast<const float4*>(o + k * LPR * 4);
  }
  float* xb = x_grad + (size_t)start * c + sub * 4;
  for (int j = rslot; j < len; j += RPS) {
#pragma unroll
    for (int k = 0; k < CHUNKS; ++k)
      if (lane_active[k]) st_stream_f4(xb + (size_t)j * c + k * LPR * 4, val[k]);
  }
}

// scalar version of the two interval kernels for C % 4 != 0
template <bool FWD>
__global__ void __launch_bounds__(kPoolBlock)
bev_pool_interval_generic_kernel(int d, int h, int w, int c, int n_intervals,
                                 const float* __restrict__ src, const int* __restrict__ geom,
                                 const int* __restrict__ istart, const int* __restrict__ ilen,
                                 float* __restrict__ dst) {
  const int lane = threadIdx.x & 31;
  const int iv = blockIdx.x * kPoolWarps + (threadIdx.x >> 5);
  if (iv >= n_intervals) return;
  const int start = istart[iv], len = ilen[iv];
  const int* gf = geom + (size_t)start * 4;
  const size_t cell = ((((size_t)gf[3] * d + gf[2]) * h + gf[0]) * (size_t)w + gf[1]) * c;
  for (int ch = lane; ch < c; ch += 32) {
    if (FWD) {
      float acc = 0.f;
      for (int j = 0; j < len; ++j) acc += src[(size_t)(start + j) * c + ch];
      dst[cell + ch] = acc;
    } else {
      const float v = src[cell + ch];
      for (int j = 0; j < len; ++j) dst[(size_t)(start + j) * c + ch] = v;
    }
  }
}

struct VecCfg {
  int lpr, chunks;
};

inline bool pick_vec_cfg(int C, VecCfg* cfg) {
  if (C <= 0 || C % 4 != 0 || C > 512) return false;
  int v = C / 4;
  int lpr = 1;
  while (lpr < v && lpr < 32) lpr <<= 1;
  cfg->lpr = lpr;
  cfg->chunks = (v + lpr - 1) / lpr;
  return true;
}

#define DBEV_DISPATCH_VEC(cfg, LAUNCH)                  \
  do {                                                  \
    if (cfg.chunks == 1) {                              \
      switch (cfg.lpr) {                                \
        case 1: { LAUNCH(1, 1); } break;                \
        case 2: { LAUNCH(2, 1); } break;                \
        case 4: { LAUNCH(4, 1); } break;                \
        case 8: { LAUNCH(8, 1); } break;                \
        case 16: { LAUNCH(16, 1); } break;              \
        default: { LAUNCH(32, 1); } break;              \
      }                                                 \
    } else if (cfg.chunks == 2) { LAUNCH(32, 2); }      \
    else if (cfg.chunks == 3) { LAUNCH(32, 3); }        \
    else { LAUNCH(32, 4); }                             \
  } while (0)

}  // namespace

// ---------------------------------------------------------------------------

static long long n_tiles_of(long long ncells, int nfast) {
  return (ncells / nfast) * ceil_div(nfast, kTileCells);
}

long long bev_plan_max_items(long long n_points, long long ncells, int nfast, int rows_per_item) {
  if (rows_per_item < 1) rows_per_item = kDefaultRowsPerItem;
  return n_tiles_of(ncells, nfast) + 2 * ((n_points + rows_per_item - 1) / rows_per_item) + 1;
}

size_t bev_plan_ws_bytes(long long n_points, long long ncells) {
  // keys x2 + payload + sort tables + per-tile counts/offsets + scan scratch
  const long long ntiles_ub = ncells;  // nfast >= 1
  return 3 * align_up((size_t)n_points * 4) + radix_sort_ws_bytes(n_points) +
         2 * align_up((size_t)ntiles_ub * 4) + scan_ws_bytes(ntiles_ub) + 2048;
}

struct PlanOut {
  uint32_t* order;
  int* cell_start;
  int* cell_end;
  int4* items;
  long long max_items;
  int* n_items;
  int rows_per_item;
  int nfast;
};

static int finish_plan(uint32_t* keys0, long long n, long long ncells, const PlanOut& po,
                       Workspace& w, void* ws, size_t ws_bytes, cudaStream_t stream) {
  const long long ntiles = n_tiles_of(ncells, po.nfast);
  uint32_t* keys1 = w.take<uint32_t>(n);
  uint32_t* vals_other = w.take<uint32_t>(n);
  int* counts = w.take<int>(ntiles);
  int* offsets = w.take<int>(ntiles);
  if (!w.ok()) {
    set_last_error("bev_plan: workspace too small (%zu bytes given, %zu needed so far)", ws_bytes,
                   w.used);
    return DBEV_ERR_WORKSPACE;
  }
  const size_t consumed = align_up(w.used);
  void* sub_ws = (char*)ws + consumed;
  const size_t sub_bytes = ws_bytes > consumed ? ws_bytes - consumed : 0;
  const int num_bits = bits_for((unsigned long long)ncells + 1);
  const int passes = (num_bits + kRadixBits - 1) / kRadixBits;
  uint32_t* keys[2] = {keys0, keys1};
  // arrange the ping-pong so the sorted payload lands in `order`
  uint32_t* vals[2];
  vals[passes & 1] = po.order;
  vals[(passes & 1) ^ 1] = vals_other;
  int sel = 0;
  int rc = radix_sort_pairs(keys, vals, /*vals_iota=*/true, (int)n, num_bits, sub_ws, sub_bytes,
                            stream, &sel);
  if (rc != DBEV_OK) return rc;
  DBEV_CUDA(cudaMemsetAsync(po.cell_start, 0, (size_t)(ncells + 1) * sizeof(int), stream));
  DBEV_CUDA(cudaMemsetAsync(po.cell_end, 0, (size_t)(ncells + 1) * sizeof(int), stream));
  if (n > 0) {
    bev_bounds_kernel<<<ceil_div(n, 256), 256, 0, stream>>>(keys[sel], n, po.cell_start,
                                                            po.cell_end);
    DBEV_CHECK_LAUNCH("bev_bounds_kernel");
  }
  // work items: count per tile, scan, emit
  const int tiles_per_row = ceil_div(po.nfast, kTileCells);
  const int rpi = po.rows_per_item < 1 ? kDefaultRowsPerItem : po.rows_per_item;
  bev_items_kernel<<<ceil_div(ntiles, 256), 256, 0, stream>>>(
      po.cell_start, po.cell_end, ntiles, po.nfast, tiles_per_row, rpi, counts, nullptr, nullptr, 0);
  DBEV_CHECK_LAUNCH("bev_items_kernel(count)");
  rc = exclusive_scan_i32(counts, offsets, (int)ntiles, po.n_items, sub_ws, sub_bytes, stream);
  if (rc != DBEV_OK) return rc;
  bev_items_kernel<<<ceil_div(ntiles, 256), 256, 0, stream>>>(
      po.cell_start, po.cell_end, ntiles, po.nfast, tiles_per_row, rpi, nullptr, offsets, po.items,
      (int)po.max_items);
  DBEV_CHECK_LAUNCH("bev_items_kernel(emit)");
  return DBEV_OK;
}

static int check_plan_out(const char* who, long long n_points, long long ncells, int nfast,
                          int rows_per_item, long long max_items) {
  DBEV_CHECK_ARG(ncells > 0 && ncells < 0x7fffffffLL && n_points < 0x7fffffffLL,
                 "%s: grid or point count exceeds 2^31", who);
  const long long need = bev_plan_max_items(n_points, ncells, nfast, rows_per_item);
  DBEV_CHECK_ARG(max_items >= need, "%s: items buffer holds %lld entries, %lld required", who,
                 max_items, need);
  return DBEV_OK;
}

int bev_plan_from_geom(const float* geom, long long n_points, int batch, const float off[3],
                       const float dx[3], const float nx_f[3], const int nx_i[3], int fast_axis,
                       int rows_per_item, uint32_t* order, int* cell_start, int* cell_end,
                       int4* items, long long max_items, int* n_items, int* point_cell, void* ws,
                       size_t ws_bytes, cudaStream_t stream) {
  DBEV_CHECK_ARG(batch > 0 && n_points >= 0 && n_points % batch == 0,
                 "bev_plan_from_geom: n_points (%lld) must be a multiple of batch (%d)", n_points,
                 batch);
  DBEV_CHECK_ARG(fast_axis == 0 || fast_axis == 1, "bev_plan_from_geom: fast_axis must be 0 or 1");
  DBEV_CHECK_ARG(nx_i[0] > 0 && nx_i[1] > 0 && nx_i[2] > 0, "bev_plan_from_geom: empty grid");
  const long long ncells = (long long)batch * nx_i[0] * nx_i[1] * nx_i[2];
  const int nfast = fast_axis == 0 ? nx_i[0] : nx_i[1];
  int rc = check_plan_out("bev_plan_from_geom", n_points, ncells, nfast, rows_per_item, max_items);
  if (rc != DBEV_OK) return rc;
  Workspace w(ws, ws_bytes);
  uint32_t* keys0 = w.take<uint32_t>(n_points);
  if (!w.ok()) {
    set_last_error("bev_plan_from_geom: workspace too small");
    return DBEV_ERR_WORKSPACE;
  }
  if (n_points > 0) {
    bev_keys_from_geom_kernel<<<ceil_div(n_points, 256), 256, 0, stream>>>(
        geom, n_points, n_points / batch, off[0], off[1], off[2], dx[0], dx[1], dx[2], nx_f[0],
        nx_f[1], nx_f[2], nx_i[0], nx_i[1], nx_i[2], fast_axis, (uint32_t)ncells, keys0,
        point_cell);
    DBEV_CHECK_LAUNCH("bev_keys_from_geom_kernel");
  }
  PlanOut po{order, cell_start, cell_end, items, max_items, n_items, rows_per_item, nfast};
  return finish_plan(keys0, n_points, ncells, po, w, ws, ws_bytes, stream);
}

int bev_plan_from_coords(const void* coords, int coords_i64, long long n_points, int batch, int n0,
                         int n1, int nz, int fast_axis, int rows_per_item, uint32_t* order,
                         int* cell_start, int* cell_end, int4* items, long long max_items,
                         int* n_items, void* ws, size_t ws_bytes, cudaStream_t stream) {
  DBEV_CHECK_ARG(fast_axis == 0 || fast_axis == 1, "bev_plan_from_coords: fast_axis must be 0 or 1");
  DBEV_CHECK_ARG(batch > 0 && n0 > 0 && n1 > 0 && nz > 0 && n_points >= 0,
                 "bev_plan_from_coords: empty grid");
  const long long ncells = (long long)batch * n0 * n1 * nz;
  const int nfast = fast_axis == 0 ? n0 : n1;
  int rc = check_plan_out("bev_plan_from_coords", n_points, ncells, nfast, rows_per_item, max_items);
  if (rc != DBEV_OK) return rc;
  Workspace w(ws, ws_bytes);
  uint32_t* keys0 = w.take<uint32_t>(n_points);
  if (!w.ok()) {
    set_last_error("bev_plan_from_coords: workspace too small");
    return DBEV_ERR_WORKSPACE;
  }
  if (n_points > 0) {
    if (coords_i64)
      bev_keys_from_coords_kernel<long long><<<ceil_div(n_points, 256), 256, 0, stream>>>(
          (const long long*)coords, n_points, batch, n0, n1, nz, fast_axis, (uint32_t)ncells, keys0);
    else
      bev_keys_from_coords_kernel<int><<<ceil_div(n_points, 256), 256, 0, stream>>>(
          (const int*)coords, n_points, batch, n0, n1, nz, fast_axis, (uint32_t)ncells, keys0);
    DBEV_CHECK_LAUNCH("bev_keys_from_coords_kernel");
  }
  PlanOut po{order, cell_start, cell_end, items, max_items, n_items, rows_per_item, nfast};
  return finish_plan(keys0, n_points, ncells, po, w, ws, ws_bytes, stream);
}

static int make_pool_geom(int C, int batch, int nz, int nslow, int nfast, long long sB,
                          long long sZ, long long sC, PoolGeom* g) {
  DBEV_CHECK_ARG(C > 0 && C <= 1024, "bev_pool: channel count %d unsupported (1..1024)", C);
  DBEV_CHECK_ARG(batch > 0 && nz > 0 && nslow > 0 && nfast > 0, "bev_pool: empty grid");
  g->C = C;
  g->nfast = nfast;
  g->nslow = nslow;
  g->nz = nz;
  g->nbz = batch * nz;
  g->tiles_per_row = ceil_div(nfast, kTileCells);
  g->sB = sB;
  g->sZ = sZ;
  g->sC = sC;
  long long tiles = (long long)g->nbz * nslow * g->tiles_per_row;
  DBEV_CHECK_ARG(tiles < 0x7fffffffLL, "bev_pool: grid too large");
  return DBEV_OK;
}

template <typename K>
static int persistent_grid(K kernel, size_t smem, int* grid) {
  if (smem > 48 * 1024)
    DBEV_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 0, dev = 0, sms = 0;
  DBEV_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kPoolBlock, smem));
  DBEV_CUDA(cudaGetDevice(&dev));
  DBEV_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  if (per_sm < 1) per_sm = 1;
  *grid = sms * per_sm;  // one resident wave; CTAs stride over the item list
  return DBEV_OK;
}

template <typename K>
static int set_smem(K kernel, size_t bytes) {
  if (bytes > 48 * 1024) {
    DBEV_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  }
  return DBEV_OK;
}

static int pick_lpr(int C) {  // lanes per row of one channel block (<= 64 channels)
  if (C % 4 != 0) return 0;
  int lpr = 1;
  while (lpr * 4 < C && lpr < 16) lpr <<= 1;
  return lpr;
}

#define DBEV_DISPATCH_LPR(lpr, LAUNCH)   \
  do {                                   \
    switch (lpr) {                       \
      case 1: { LAUNCH(1); } break;      \
      case 2: { LAUNCH(2); } break;      \
      case 4: { LAUNCH(4); } break;      \
      case 8: { LAUNCH(8); } break;      \
      default: { LAUNCH(16); } break;    \
    }                                    \
  } while (0)

// Work-queue counters are per-launch state. Eager launches take the next of kEagerSlots pairs round-robin; a
// launch recorded into a CUDA graph gets a pair of its own that no other launch (eager or captured) ever uses,
// because the graph replays with the slot baked into its arguments while the host counter moves on - two
// kernels pulling from one counter would skip or split items. The pair is zeroed on the launching stream
// right before the kernel (a memset node in a graph), so a slot left dirty by a faulted kernel heals itself.
static int next_sched_slot(cudaStream_t stream) {
  static std::atomic<unsigned> eager{0}, captured{0};
  static int* base = nullptr;
  if (!base && cudaGetSymbolAddress((void**)&base, g_sched) != cudaSuccess) return -1;
  cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(stream, &st) != cudaSuccess) return -1;
  int slot;
  if (st == cudaStreamCaptureStatusActive) {
    const unsigned k = captured.fetch_add(1, std::memory_order_relaxed);
    if (k >= (unsigned)(kSchedSlots - kEagerSlots)) return -1;
    slot = kEagerSlots + (int)k;
  } else {
    slot = (int)(eager.fetch_add(1, std::memory_order_relaxed) % kEagerSlots);
  }
  if (cudaMemsetAsync(base + 2 * slot, 0, 2 * sizeof(int), stream) != cudaSuccess) return -1;
  return slot;
}

#define DBEV_SCHED_SLOT(var)                                                                              \
  const int var = next_sched_slot(stream);                                                                \
  if (var < 0) {                                                                                          \
    set_last_error("bev_pool: no work-queue slot (more than %d captured launches, or a CUDA error)",      \
                   kSchedSlots - kEagerSlots);                                                            \
    return DBEV_ERR_CUDA;                                                                                 \
  }

static int gather_forward_impl(const float* x, int C, const uint32_t* order, const int* cell_start,
                               const int* cell_end, const int4* items, const int* n_items,
                               int batch, int nz, int nslow, int nfast, long long sB, long long sZ,
                               long long sC, float* out, const LiftArgs* lift,
                               cudaStream_t stream) {
  PoolGeom g;
  int rc = make_pool_geom(C, batch, nz, nslow, nfast, sB, sZ, sC, &g);
  if (rc != DBEV_OK) return rc;
  const int lpr = pick_lpr(C);
  if (lift) {
    DBEV_CHECK_ARG(lpr > 0, "lift_splat: C must be a multiple of 4 (got %d)", C);
    const int cb = lpr * 4, nw = 32 / lpr;
    const size_t smem = (size_t)kPoolWarps * (cb * kTilePitch + nw * 2 * cb) * sizeof(float);
    int grid = 0;
#define LAUNCH(L)                                                                        \
  rc = persistent_grid(bev_pool_gather_fwd_kernel<L, true>, smem, &grid);                \
  if (rc != DBEV_OK) return rc;                                                          \
  bev_pool_gather_fwd_kernel<L, true><<<grid, kPoolBlock, smem, stream>>>(               \
      x, order, cell_start, cell_end, items, n_items, out, g, *lift, sched_slot)
    DBEV_SCHED_SLOT(sched_slot);
    DBEV_DISPATCH_LPR(lpr, LAUNCH);
#undef LAUNCH
    DBEV_CHECK_LAUNCH("lift_splat_fwd_kernel");
    return DBEV_OK;
  }
  const LiftArgs none{nullptr, FastDiv(1), FastDiv(1)};
  if (lpr > 0 && sC == 1) {
    // channels-last output: rows of the cells-major map [batch][nz][nslow][nfast][C]
    DBEV_CHECK_ARG(sZ == (long long)nslow * nfast * C && sB == (long long)nz * nslow * nfast * C,
                   "bev_pool: a channel stride of 1 needs the cells-major layout [B][nz][slow][fast][C]");
    const int cb = lpr * 4, nw = 32 / lpr;
    const size_t smem = (size_t)kPoolWarps * (nw * 2 * cb) * sizeof(float);
    int grid = 0;
    static const int minb = getenv("DBEV_POOL_MINB") ? atoi(getenv("DBEV_POOL_MINB")) : 6;   // experiment switch
#define LAUNCH_MB(L, MB)                                                                  \
  rc = persistent_grid(bev_pool_gather_fwd_kernel<L, false, MB, true>, smem, &grid);     \
  if (rc != DBEV_OK) return rc;                                                          \
  bev_pool_gather_fwd_kernel<L, false, MB, true><<<grid, kPoolBlock, smem, stream>>>(    \
      x, order, cell_start, cell_end, items, n_items, out, g, none, sched_slot)
#define LAUNCH(L)                                                                        \
  if (minb == 8) { LAUNCH_MB(L, 8); } else if (minb == 7) { LAUNCH_MB(L, 7); } else if (minb == 5) { LAUNCH_MB(L, 5); } else { LAUNCH_MB(L, 6); }
    DBEV_SCHED_SLOT(sched_slot);
    DBEV_DISPATCH_LPR(lpr, LAUNCH);
#undef LAUNCH
#undef LAUNCH_MB
    DBEV_CHECK_LAUNCH("bev_pool_gather_fwd_kernel (channels-last)");
    return DBEV_OK;
  }
  if (lpr > 0) {
    const int cb = lpr * 4, nw = 32 / lpr;
    const size_t smem = (size_t)kPoolWarps * (cb * kTilePitch + nw * 2 * cb) * sizeof(float);
    int grid = 0;
#define LAUNCH(L)                                                                        \
  rc = persistent_grid(bev_pool_gather_fwd_kernel<L, false, 6>, smem, &grid);            \
  if (rc != DBEV_OK) return rc;                                                          \
  bev_pool_gather_fwd_kernel<L, false, 6><<<grid, kPoolBlock, smem, stream>>>(           \
      x, order, cell_start, cell_end, items, n_items, out, g, none, sched_slot)
    DBEV_SCHED_SLOT(sched_slot);
    DBEV_DISPATCH_LPR(lpr, LAUNCH);
#undef LAUNCH
  } else {
    const int grid = g.nbz * nslow * g.tiles_per_row;
    const size_t smem = (size_t)C * kTilePitch * sizeof(float);
    rc = set_smem(bev_pool_gather_fwd_generic_kernel, smem);
    if (rc != DBEV_OK) return rc;
    bev_pool_gather_fwd_generic_kernel<<<grid, kPoolBlock, smem, stream>>>(x, order, cell_start,
                                                                           cell_end, out, g);
  }
  DBEV_CHECK_LAUNCH("bev_pool_gather_fwd_kernel");
  return DBEV_OK;
}

int bev_pool_gather_forward(const float* x, int C, const uint32_t* order, const int* cell_start,
                            const int* cell_end, const int4* items, const int* n_items, int batch,
                            int nz, int nslow, int nfast, long long sB, long long sZ, long long sC,
                            float* out, cudaStream_t stream) {
  return gather_forward_impl(x, C, order, cell_start, cell_end, items, n_items, batch, nz, nslow,
                             nfast, sB, sZ, sC, out, nullptr, stream);
}

int lift_splat_forward(const float* depth, const float* feat_cl, int C, int D, int fhw,
                       const uint32_t* order, const int* cell_start, const int* cell_end,
                       const int4* items, const int* n_items, int batch, int nz, int nslow,
                       int nfast, long long sB, long long sZ, long long sC, float* out,
                       cudaStream_t stream) {
  DBEV_CHECK_ARG(D > 0 && fhw > 0, "lift_splat: bad frustum shape D=%d fH*fW=%d", D, fhw);
  DBEV_CHECK_ARG((long long)D * fhw < (1LL << 30), "lift_splat: frustum too large");
  const LiftArgs la{depth, FastDiv((unsigned)(D * fhw)), FastDiv((unsigned)fhw)};
  return gather_forward_impl(feat_cl, C, order, cell_start, cell_end, items, n_items, batch, nz,
                             nslow, nfast, sB, sZ, sC, out, &la, stream);
}

int lift_splat_backward(const float* g_cl, const float* depth, const float* feat_cl,
                        const int* point_cell, long long n_pix, int C, int D, int fhw,
                        float* d_depth, float* d_feat_cl, cudaStream_t stream) {
  DBEV_CHECK_ARG(n_pix >= 0 && D > 0 && fhw > 0 && C > 0, "lift_splat_backward: bad sizes");
  const int lpr = pick_lpr(C);
  DBEV_CHECK_ARG(lpr > 0, "lift_splat: C must be a multiple of 4 (got %d)", C);
  if (n_pix == 0) return DBEV_OK;
#define LAUNCH(L)                                                                         \
  lift_splat_bwd_kernel<L><<<ceil_div(n_pix, 256 / L), 256, 0, stream>>>(                 \
      g_cl, depth, feat_cl, point_cell, n_pix, C, D, fhw, d_depth, d_feat_cl)
  DBEV_DISPATCH_LPR(lpr, LAUNCH);
#undef LAUNCH
  DBEV_CHECK_LAUNCH("lift_splat_bwd_kernel");
  return DBEV_OK;
}

int transpose_batched(const float* in, float* out, int batch, int rows, int cols,
                      cudaStream_t stream) {
  DBEV_CHECK_ARG(batch >= 0 && rows >= 0 && cols >= 0 && batch < 65536, "transpose: bad sizes");
  if (batch == 0 || rows == 0 || cols == 0) return DBEV_OK;
  dim3 grid(ceil_div(cols, 32), ceil_div(rows, 32), batch);
  DBEV_CHECK_ARG(grid.y < 65536, "transpose: too many rows (%d)", rows);
  transpose_batched_kernel<<<grid, 256, 0, stream>>>(in, out, rows, cols);
  DBEV_CHECK_LAUNCH("transpose_batched_kernel");
  return DBEV_OK;
}

// Point-centric backward: x_grad[p, :] = grad_cl[cell(p), :] (zeros for dropped points). The 1 GB of
// row writes is perfectly sequential in p and the cell rows (the BEV gradient transposed once to
// cells-major, 67 MB at configs[1]) are gathered from L2; the cell-centric kernel above scatters
// 256-byte rows through the sorted order instead (0.39 -> 0.2x ms at configs[1]).
__global__ void __launch_bounds__(256)
bev_pool_point_bwd_kernel(const float4* __restrict__ grad_cl, const int* __restrict__ point_cell,
                          long long n_points, int c4, float4* __restrict__ x_grad) {
  const long long total = n_points * c4;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
    const long long p = t / c4;
    const int j = (int)(t - p * c4);
    const int cell = __ldg(point_cell + p);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (cell >= 0) v = __ldg(grad_cl + (long long)cell * c4 + j);
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(x_grad + t), "f"(v.x),
                 "f"(v.y), "f"(v.z), "f"(v.w));
  }
}

// ---------------------------------------------------------------------------------------------
// Sort-free lift + splat (opt-in, `deterministic = false`): one LPR-lane group per image pixel keeps
// the pixel's feature row in registers (read ONCE: 17 MB instead of 0.9 GB of row reads from L2) and
// walks its D depth bins; every kept (pixel, bin) adds depth * feat into the BEV cell's channels-last
// row with 16-byte vector reductions (red.global.add.v4.f32) that resolve in L2, where the 67 MB BEV
// map lives. No plan, no sort, no work list: the cell of a frustum point comes straight from the
// geometry (point_cell). The price is the fp32 summation order (run-to-run differences ~1e-7
// relative; the plan-based kernel above stays the bit-reproducible default).
// ---------------------------------------------------------------------------------------------
template <int LPR>
__global__ void __launch_bounds__(256)
lift_splat_atomic_fwd_kernel(const float* __restrict__ depth, const float4* __restrict__ feat_cl,
                             const int* __restrict__ point_cell, long long n_pixels, int D, int fhw,
                             int c4, float4* __restrict__ out_cl) {
  const long long gid = ((long long)blockIdx.x * blockDim.x + threadIdx.x) / LPR;
  const int sub = threadIdx.x % LPR;
  const unsigned wmask = (LPR == 32) ? 0xffffffffu : (((1u << LPR) - 1u) << ((threadIdx.x & 31) / LPR * LPR));
  const bool pix_ok = gid < n_pixels;
  const long long pix = pix_ok ? gid : 0;
  const long long bn = pix / fhw;
  const long long p0 = bn * D * fhw + (pix - bn * fhw);   // point id of depth bin 0
  const bool lane_ok = sub < c4;
  float4 f = make_float4(0.f, 0.f, 0.f, 0.f);
  if (pix_ok && lane_ok) f = __ldg(feat_cl + pix * c4 + sub);
  for (int d0 = 0; d0 < D; d0 += LPR) {
    // cooperative addressing: lane `sub` resolves bin d0 + sub, the group exchanges by shuffle
    const int dm = d0 + sub;
    int cell = -1;
    float w = 0.f;
    if (pix_ok && dm < D) {
      cell = __ldg(point_cell + p0 + (long long)dm * fhw);
      w = __ldg(depth + p0 + (long long)dm * fhw);
    }
    const int cnt = min(LPR, D - d0);
    for (int j = 0; j < cnt; ++j) {
      const int cj = __shfl_sync(wmask, cell, j, LPR);
      const float wj = __shfl_sync(wmask, w, j, LPR);
      if (cj >= 0 && lane_ok) {
        // rounded product first, like the reference's materialised volume
        const float4 v = make_float4(__fmul_rn(wj, f.x), __fmul_rn(wj, f.y), __fmul_rn(wj, f.z), __fmul_rn(wj, f.w));
        atomicAdd(out_cl + (long long)cj * c4 + sub, v);
      }
    }
  }
}

int bev_pool_gather_backward(const float* out_grad, int C, const uint32_t* order,
                             const int* cell_start, const int* cell_end, const int4* items,
                             const int* n_items, int batch, int nz, int nslow, int nfast,
                             long long sB, long long sZ, long long sC, float* x_grad,
                             cudaStream_t stream) {
  PoolGeom g;
  int rc = make_pool_geom(C, batch, nz, nslow, nfast, sB, sZ, sC, &g);
  if (rc != DBEV_OK) return rc;
  const int lpr = pick_lpr(C);
  if (lpr > 0) {
    const int cb = lpr * 4;
    const size_t smem = (size_t)kPoolWarps * (cb * kTilePitch) * sizeof(float);
    int grid = 0;
#define LAUNCH(L)                                                                        \
  rc = persistent_grid(bev_pool_gather_bwd_kernel<L>, smem, &grid);                      \
  if (rc != DBEV_OK) return rc;                                                          \
  bev_pool_gather_bwd_kernel<L><<<grid, kPoolBlock, smem, stream>>>(                     \
      out_grad, order, cell_start, cell_end, items, n_items, x_grad, g, sched_slot)
    DBEV_SCHED_SLOT(sched_slot);
    DBEV_DISPATCH_LPR(lpr, LAUNCH);
#undef LAUNCH
  } else {
    const int grid = g.nbz * nslow * g.tiles_per_row;
    const size_t smem = (size_t)C * kTilePitch * sizeof(float);
    rc = set_smem(bev_pool_gather_bwd_generic_kernel, smem);
    if (rc != DBEV_OK) return rc;
    bev_pool_gather_bwd_generic_kernel<<<grid, kPoolBlock, smem, stream>>>(
        out_grad, order, cell_start, cell_end, x_grad, g);
  }
  DBEV_CHECK_LAUNCH("bev_pool_gather_bwd_kernel");
  const long long ncells = (long long)g.nbz * nslow * nfast;
  bev_pool_zero_tail_kernel<<<kNumSMs * 2, 256, 0, stream>>>(order, cell_start + ncells,
                                                             cell_end + ncells, x_grad, C);
  DBEV_CHECK_LAUNCH("bev_pool_zero_tail_kernel");
  return DBEV_OK;
}

int bev_pool_point_backward(const float* grad_cl, const int* point_cell, long long n_points, int C,
                            float* x_grad, cudaStream_t stream) {
  DBEV_CHECK_ARG(n_points >= 0 && C > 0 && C % 4 == 0, "bev_pool_point_backward: C must be a multiple of 4");
  DBEV_CHECK_ARG(((uintptr_t)grad_cl & 15) == 0 && ((uintptr_t)x_grad & 15) == 0,
                 "bev_pool_point_backward: pointers must be 16-byte aligned");
  if (n_points == 0) return DBEV_OK;
  const long long total = n_points * (C / 4);
  long long blocks = (total + 255) / 256;
  if (blocks > (long long)kNumSMs * 64) blocks = (long long)kNumSMs * 64;
  bev_pool_point_bwd_kernel<<<(int)blocks, 256, 0, stream>>>((const float4*)grad_cl, point_cell, n_points,
                                                             C / 4, (float4*)x_grad);
  DBEV_CHECK_LAUNCH("bev_pool_point_bwd_kernel");
  return DBEV_OK;
}

int bev_point_cells(const float* geom, long long n_points, int batch, const float* off, const float* dx,
                    const float* nx_f, const int* nx_i, int fast_axis, int* point_cell,
                    cudaStream_t stream, int frames) {
  DBEV_CHECK_ARG(fast_axis == 0 || fast_axis == 1, "bev_point_cells: fast_axis must be 0 or 1");
  DBEV_CHECK_ARG(batch > 0 && n_points >= 0 && n_points % batch == 0, "bev_point_cells: bad sizes");
  DBEV_CHECK_ARG(frames >= 1 && batch % frames == 0, "bev_point_cells: batch must be a multiple of frames");
  const long long ncells = (long long)batch * nx_i[0] * nx_i[1] * nx_i[2];
  DBEV_CHECK_ARG(ncells > 0 && ncells < 0x7fffffffLL, "bev_point_cells: bad grid");
  if (n_points == 0) return DBEV_OK;
  bev_keys_from_geom_kernel<<<ceil_div(n_points, 256), 256, 0, stream>>>(
      geom, n_points, n_points / batch, off[0], off[1], off[2], dx[0], dx[1], dx[2], nx_f[0], nx_f[1],
      nx_f[2], nx_i[0], nx_i[1], nx_i[2], fast_axis, (uint32_t)ncells, nullptr, point_cell, frames);
  DBEV_CHECK_LAUNCH("bev_keys_from_geom_kernel");
  return DBEV_OK;
}

int lift_splat_atomic_forward(const float* depth, const float* feat_cl, const int* point_cell,
                              long long n_pixels, int C, int D, int fhw, long long n_cells, float* out_cl,
                              cudaStream_t stream) {
  DBEV_CHECK_ARG(n_pixels >= 0 && C > 0 && C % 4 == 0 && C <= 128 && D > 0 && fhw > 0 && n_cells > 0,
                 "lift_splat_atomic: C must be a multiple of 4, at most 128");
  DBEV_CHECK_ARG(((uintptr_t)feat_cl & 15) == 0 && ((uintptr_t)out_cl & 15) == 0,
                 "lift_splat_atomic: pointers must be 16-byte aligned");
  DBEV_CUDA(cudaMemsetAsync(out_cl, 0, (size_t)n_cells * C * sizeof(float), stream));
  if (n_pixels == 0) return DBEV_OK;
  const int c4 = C / 4;
  int lpr = 1;
  while (lpr < c4) lpr <<= 1;
  const long long threads = n_pixels * lpr;
#define DBEV_LSA(L)                                                                                  \
  lift_splat_atomic_fwd_kernel<L><<<ceil_div(threads, 256), 256, 0, stream>>>(                       \
      depth, (const float4*)feat_cl, point_cell, n_pixels, D, fhw, c4, (float4*)out_cl)
  switch (lpr) {
    case 1: DBEV_LSA(1); break;
    case 2: DBEV_LSA(2); break;
    case 4: DBEV_LSA(4); break;
    case 8: DBEV_LSA(8); break;
    case 16: DBEV_LSA(16); break;
    default: DBEV_LSA(32); break;
  }
#undef DBEV_LSA
  DBEV_CHECK_LAUNCH("lift_splat_atomic_fwd_kernel");
  return DBEV_OK;
}

int bev_pool_interval_forward(int b, int d, int h, int w, int n, int c, int n_intervals,
                              const float* x, const int* geom_feats, const int* interval_starts,
                              const int* interval_lengths, float* out, int zero_out,
                              cudaStream_t stream) {
  DBEV_CHECK_ARG(b > 0 && d > 0 && h > 0 && w > 0 && c > 0 && n >= 0 && n_intervals >= 0,
                 "bev_pool_forward: bad sizes b=%d d=%d h=%d w=%d n=%d c=%d", b, d, h, w, n, c);
  if (zero_out)
    DBEV_CUDA(cudaMemsetAsync(out, 0, (size_t)b * d * h * w * c * sizeof(float), stream));
  if (n_intervals == 0) return DBEV_OK;
  const int grid = ceil_div(n_intervals, kPoolWarps);
  VecCfg cfg;
  if (pick_vec_cfg(c, &cfg)) {
#define LAUNCH(L, K)                                                      \
  bev_pool_interval_fwd_kernel<L, K><<<grid, kPoolBlock, 0, stream>>>(    \
      d, h, w, c, n_intervals, x, geom_feats, interval_starts, interval_lengths, out)
    DBEV_DISPATCH_VEC(cfg, LAUNCH);
#undef LAUNCH
  } else {
    bev_pool_interval_generic_kernel<true><<<grid, kPoolBlock, 0, stream>>>(
        d, h, w, c, n_intervals, x, geom_feats, interval_starts, interval_lengths, out);
  }
  DBEV_CHECK_LAUNCH("bev_pool_interval_fwd_kernel");
  return DBEV_OK;
}

int bev_pool_interval_backward(int b, int d, int h, int w, int n, int c, int n_intervals,
                               const float* out_grad, const int* geom_feats,
                               const int* interval_starts, const int* interval_lengths,
                               float* x_grad, int zero_x_grad, cudaStream_t stream) {
  DBEV_CHECK_ARG(b > 0 && d > 0 && h > 0 && w > 0 && c > 0 && n >= 0 && n_intervals >= 0,
                 "bev_pool_backward: bad sizes b=%d d=%d h=%d w=%d n=%d c=%d", b, d, h, w, n, c);
  if (zero_x_grad) DBEV_CUDA(cudaMemsetAsync(x_grad, 0, (size_t)n * c * sizeof(float), stream));
  if (n_intervals == 0) return DBEV_OK;
  const int grid = ceil_div(n_intervals, kPoolWarps);
  VecCfg cfg;
  if (pick_vec_cfg(c, &cfg)) {
#define LAUNCH(L, K)                                                      \
  bev_pool_interval_bwd_kernel<L, K><<<grid, kPoolBlock, 0, stream>>>(    \
      d, h, w, c, n_intervals, out_grad, geom_feats, interval_starts, interval_lengths, x_grad)
    DBEV_DISPATCH_VEC(cfg, LAUNCH);
#undef LAUNCH
  } else {
    bev_pool_interval_generic_kernel<false><<<grid, kPoolBlock, 0, stream>>>(
        d, h, w, c, n_intervals, out_grad, geom_feats, interval_starts, interval_lengths, x_grad);
  }
  DBEV_CHECK_LAUNCH("bev_pool_interval_bwd_kernel");
  return DBEV_OK;
}

}  // namespace dbev
