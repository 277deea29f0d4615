// LiDAR voxelization for B200: dynamic voxelize, hard voxelize, dynamic scatter.
//
// Reference behaviour reproduced here (bit-exact integer outputs):
//   dynamic_voxelize        mmdet3d/ops/voxel/src/voxelization_cpu.cpp:8-43,146-171
//                           mmdet3d/ops/voxel/src/voxelization_cuda.cu:25-61,485-528
//   hard_voxelize           mmdet3d/ops/voxel/src/voxelization_cpu.cpp:45-144
//                           mmdet3d/ops/voxel/src/voxelization_cuda.cu:64-180,231-402
//   dynamic_point_to_voxel  mmdet3d/ops/voxel/src/scatter_points_cuda.cu:81-103,183-239 (fwd)
//                           mmdet3d/ops/voxel/src/scatter_points_cuda.cu:106-179,241-308 (bwd)
//
// Design (not a port). The reference's deterministic hard voxelize is an
// O(N^2) pairwise scan plus a single-thread <<<1,1>>> pass with four device
// synchronisations; its dynamic scatter calls at::unique_dim (a sort of Nx3
// rows) and then float atomics. Here everything hangs off ONE stable radix
// sort of (linear voxel key, point id):
//   * equal keys keep ascending point id  -> in-voxel order = point order,
//     rank in voxel = position - segment head,
//   * the head of a segment is the voxel's first point -> an exclusive scan of
//     "is a first point" flags in POINT order numbers the voxels in
//     first-appearance order (the reference's voxel order),
//   * sorted keys are the lexicographic (z, y, x) order at::unique_dim returns.
// No atomics on floats, no device sync, no host round trip except the one
// count the reference API returns to Python. All kernels are HBM/latency
// bound streaming passes over N points.
#include "voxelize.cuh"

#include <math.h>

#include "sort.cuh"

namespace dbev {

namespace {

struct VoxGrid {
  float vx, vy, vz;
  float xmin, ymin, zmin;
  int gx, gy, gz;
};

// floor((p - min) / voxel) in fp32 with IEEE sub/div (no contraction), as the
// reference computes it (voxelization_cpu.cpp:24, voxelization_cuda.cu:37).
__device__ __forceinline__ int vox_coord(float p, float lo, float vs) {
  return (int)floorf(__fdiv_rn(__fsub_rn(p, lo), vs));
}

// coors (z, y, x) or (-1, -1, -1): the CPU semantics (voxelization_cpu.cpp:33-38).
// The reference GPU kernel early-returns and leaves later components at their
// initial 0 (voxelization_cuda.cu:38-54); both mark the point invalid for every
// consumer (any component < 0), the CPU form is the well-defined one.
__global__ void __launch_bounds__(256)
dynamic_voxelize_kernel(const float* __restrict__ points, int n, int nfeat, VoxGrid g,
                        int* __restrict__ coors, uint32_t* __restrict__ keys, uint32_t sentinel) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float* p = points + (size_t)i * nfeat;
  const int cx = vox_coord(p[0], g.xmin, g.vx);
  const int cy = vox_coord(p[1], g.ymin, g.vy);
  const int cz = vox_coord(p[2], g.zmin, g.vz);
  const bool ok = cx >= 0 && cx < g.gx && cy >= 0 && cy < g.gy && cz >= 0 && cz < g.gz;
  if (coors) {
    coors[i * 3 + 0] = ok ? cz : -1;
    coors[i * 3 + 1] = ok ? cy : -1;
    coors[i * 3 + 2] = ok ? cx : -1;
  }
  if (keys) keys[i] = ok ? (uint32_t)(((long long)cz * g.gy + cy) * g.gx + cx) : sentinel;
}

// keys for dynamic scatter: coors[n, ncol] (ncol = 3: z,y,x; ncol = 4: b,z,y,x)
__global__ void __launch_bounds__(256)
coors_to_keys_kernel(const int* __restrict__ coors, int n, int ncol, int d0, int d1, int d2,
                     int d3, uint32_t sentinel, uint32_t* __restrict__ keys) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int* c = coors + (size_t)i * ncol;
  long long key;
  bool ok;
  if (ncol == 3) {
    ok = c[0] >= 0 && c[1] >= 0 && c[2] >= 0 && c[0] < d0 && c[1] < d1 && c[2] < d2;
    key = ((long long)c[0] * d1 + c[1]) * d2 + c[2];
  } else {
    // a negative voxel coordinate invalidates the row (coors.lt(0).any(-1),
    // scatter_points_cuda.cu:202); the batch column takes part in that test too
    ok = c[0] >= 0 && c[1] >= 0 && c[2] >= 0 && c[3] >= 0 && c[0] < d0 && c[1] < d1 &&
         c[2] < d2 && c[3] < d3;
    key = (((long long)c[0] * d1 + c[1]) * d2 + c[2]) * d3 + c[3];
  }
  keys[i] = ok ? (uint32_t)key : sentinel;
}

// head flag of every sorted position (valid keys only)
__global__ void __launch_bounds__(256)
head_flags_kernel(const uint32_t* __restrict__ skeys, int n, uint32_t sentinel,
                  int* __restrict__ flags) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const uint32_t k = skeys[j];
  flags[j] = (k != sentinel && (j == 0 || skeys[j - 1] != k)) ? 1 : 0;
}

// seg_of[j] = exclusive scan of head flags (+flag - 1 = ordinal of j's segment).
// Writes head position of each segment and, for hard voxelize, marks the first
// point (in point order) of every voxel.
__global__ void __launch_bounds__(256)
segment_heads_kernel(const uint32_t* __restrict__ skeys, const uint32_t* __restrict__ sidx,
                     const int* __restrict__ flags, const int* __restrict__ excl, int n,
                     uint32_t sentinel, int* __restrict__ head_pos, int* __restrict__ first_flag) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  if (flags[j]) {
    head_pos[excl[j]] = j;
    if (first_flag) first_flag[sidx[j]] = 1;
  }
  // one-past-the-end marker: position of the first invalid key (or n)
  if (skeys[j] == sentinel && (j == 0 || skeys[j - 1] != sentinel)) head_pos[excl[j]] = j;
}

__global__ void set_tail_marker_kernel(const uint32_t* __restrict__ skeys, int n, uint32_t sentinel,
                                       const int* __restrict__ nseg, int* __restrict__ head_pos) {
  // when no invalid key exists the marker after the last segment is n
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    if (n == 0 || skeys[n - 1] != sentinel) head_pos[*nseg] = n;
  }
}

// hard voxelize scatter: sorted position j -> (voxel id, slot) -> copy the point
__global__ void __launch_bounds__(256)
hard_scatter_kernel(const float* __restrict__ points, int nfeat, const uint32_t* __restrict__ skeys,
                    const uint32_t* __restrict__ sidx, const int* __restrict__ flags,
                    const int* __restrict__ excl, const int* __restrict__ head_pos,
                    const int* __restrict__ vid_of_point, int n, uint32_t sentinel, VoxGrid g,
                    int max_points, int max_voxels, float* __restrict__ voxels,
                    int* __restrict__ coors, int* __restrict__ num_points_per_voxel) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const uint32_t k = skeys[j];
  if (k == sentinel) return;
  const int seg = excl[j] + flags[j] - 1;
  const int head = head_pos[seg];
  const int vid = vid_of_point[sidx[head]];  // first-appearance order
  if (vid >= max_voxels) return;             // voxel_num >= max_voxels -> dropped (cpu.cpp:78)
  const int rank = j - head;
  if (rank == 0) {
    const int len = head_pos[seg + 1] - head;
    num_points_per_voxel[vid] = min(len, max_points);
    const int cx = (int)(k % (uint32_t)g.gx);
    const int cy = (int)((k / (uint32_t)g.gx) % (uint32_t)g.gy);
    const int cz = (int)(k / ((uint32_t)g.gx * (uint32_t)g.gy));
    coors[vid * 3 + 0] = cz;
    coors[vid * 3 + 1] = cy;
    coors[vid * 3 + 2] = cx;
    // zero the unused point slots of this voxel so the caller's buffer need not be cleared
    float* vz = voxels + ((size_t)vid * max_points + len) * nfeat;
    for (int t = 0; t < (max_points - min(len, max_points)) * nfeat; ++t) vz[t] = 0.f;
  }
  if (rank < max_points) {
    const float* src = points + (size_t)sidx[j] * nfeat;
    float* dst = voxels + ((size_t)vid * max_points + rank) * nfeat;
    for (int t = 0; t < nfeat; ++t) dst[t] = src[t];
  }
}

__global__ void clamp_count_kernel(const int* __restrict__ nseg, int max_voxels,
                                   int* __restrict__ out) {
  if (threadIdx.x == 0 && blockIdx.x == 0) *out = (max_voxels >= 0) ? min(*nseg, max_voxels) : *nseg;
}

// dynamic scatter: segment metadata
__global__ void __launch_bounds__(256)
scatter_meta_kernel(const uint32_t* __restrict__ skeys, const uint32_t* __restrict__ sidx,
                    const int* __restrict__ flags, const int* __restrict__ excl,
                    const int* __restrict__ head_pos, int n, uint32_t sentinel, int ncol, int d1,
                    int d2, int d3, int* __restrict__ out_coors, int* __restrict__ coors_map,
                    int* __restrict__ reduce_count) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const uint32_t k = skeys[j];
  if (k == sentinel) {
    coors_map[sidx[j]] = -1;
    return;
  }
  const int seg = excl[j] + flags[j] - 1;
  coors_map[sidx[j]] = seg;
  if (flags[j]) {
    reduce_count[seg] = head_pos[seg + 1] - j;
    uint32_t r = k;
    if (ncol == 3) {
      out_coors[seg * 3 + 2] = (int)(r % (uint32_t)d2); r /= (uint32_t)d2;
      out_coors[seg * 3 + 1] = (int)(r % (uint32_t)d1); r /= (uint32_t)d1;
      out_coors[seg * 3 + 0] = (int)r;
    } else {
      out_coors[seg * 4 + 3] = (int)(r % (uint32_t)d3); r /= (uint32_t)d3;
      out_coors[seg * 4 + 2] = (int)(r % (uint32_t)d2); r /= (uint32_t)d2;
      out_coors[seg * 4 + 1] = (int)(r % (uint32_t)d1); r /= (uint32_t)d1;
      out_coors[seg * 4 + 0] = (int)r;
    }
  }
}

// reduced[seg, c] = max / sum / mean over the segment's rows, summed in point
// order (reproducible; the reference uses float atomics, scatter_points_cuda.cu:95-100)
__global__ void __launch_bounds__(256)
scatter_reduce_kernel(const float* __restrict__ feats, int nfeat, const uint32_t* __restrict__ sidx,
                      const int* __restrict__ head_pos, const int* __restrict__ nseg_ptr,
                      int reduce_type, float* __restrict__ reduced) {
  const long long total = (long long)(*nseg_ptr) * nfeat;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const int seg = (int)(t / nfeat), c = (int)(t % nfeat);
    const int s = head_pos[seg], e = head_pos[seg + 1];
    float acc;
    if (reduce_type == 2) {  // MAX
      acc = -INFINITY;
      for (int j = s; j < e; ++j) acc = fmaxf(acc, feats[(size_t)sidx[j] * nfeat + c]);
    } else {
      acc = 0.f;
      for (int j = s; j < e; ++j) acc += feats[(size_t)sidx[j] * nfeat + c];
      if (reduce_type == 1) acc = acc / (float)(e - s);  // MEAN (reduced_feats /= count, :233-234)
    }
    reduced[t] = acc;
  }
}

// backward, sum / mean: grad_feats[i, c] = grad_reduced[map[i], c] (/ count)
// (add_reduce_traceback_grad_kernel, scatter_points_cuda.cu:106-133)
__global__ void __launch_bounds__(256)
scatter_bwd_add_kernel(const float* __restrict__ grad_reduced, const int* __restrict__ coors_map,
                       const int* __restrict__ reduce_count, long long n, int nfeat, int reduce_type,
                       float* __restrict__ grad_feats) {
  const long long total = n * nfeat;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const long long i = t / nfeat;
    const int c = (int)(t % nfeat);
    const int seg = coors_map[i];
    float gval = 0.f;
    if (seg >= 0) {
      gval = grad_reduced[(size_t)seg * nfeat + c];
      if (reduce_type == 1) gval = gval / (float)reduce_count[seg];
    }
    grad_feats[t] = gval;
  }
}

// backward, max: the gradient goes to the LOWEST point id attaining the max
// (atomicMin trace, scatter_points_cuda.cu:136-179)
__global__ void __launch_bounds__(256)
scatter_bwd_max_trace_kernel(const float* __restrict__ feats, const float* __restrict__ reduced,
                             const int* __restrict__ coors_map, long long n, int nfeat,
                             int* __restrict__ reduce_from) {
  const long long total = n * nfeat;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const long long i = t / nfeat;
    const int c = (int)(t % nfeat);
    const int seg = coors_map[i];
    if (seg >= 0 && feats[t] == reduced[(size_t)seg * nfeat + c])
      atomicMin(&reduce_from[(size_t)seg * nfeat + c], (int)i);
  }
}

__global__ void __launch_bounds__(256)
scatter_bwd_max_apply_kernel(const float* __restrict__ grad_reduced,
                             const int* __restrict__ reduce_from, long long m, int nfeat,
                             long long n, float* __restrict__ grad_feats) {
  const long long total = m * nfeat;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const int from = reduce_from[t];
    const int c = (int)(t % nfeat);
    if (from < n) grad_feats[(size_t)from * nfeat + c] = grad_reduced[t];
  }
}

__global__ void fill_i32_kernel(int* p, long long n, int v) {
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n;
       t += (long long)gridDim.x * blockDim.x)
    p[t] = v;
}

int make_grid(const float* voxel_size, const float* coors_range, VoxGrid* g) {
  DBEV_CHECK_ARG(voxel_size[0] > 0 && voxel_size[1] > 0 && voxel_size[2] > 0,
                 "voxelize: voxel_size must be positive");
  g->vx = voxel_size[0]; g->vy = voxel_size[1]; g->vz = voxel_size[2];
  g->xmin = coors_range[0]; g->ymin = coors_range[1]; g->zmin = coors_range[2];
  // grid = round((max - min) / voxel) in fp32 (voxelization_cpu.cpp:120-123)
  g->gx = (int)roundf((coors_range[3] - coors_range[0]) / voxel_size[0]);
  g->gy = (int)roundf((coors_range[4] - coors_range[1]) / voxel_size[1]);
  g->gz = (int)roundf((coors_range[5] - coors_range[2]) / voxel_size[2]);
  DBEV_CHECK_ARG(g->gx > 0 && g->gy > 0 && g->gz > 0, "voxelize: empty grid %dx%dx%d", g->gx,
                 g->gy, g->gz);
  DBEV_CHECK_ARG((long long)g->gx * g->gy * g->gz < 0xfffffff0LL,
                 "voxelize: grid %dx%dx%d exceeds 32-bit voxel keys", g->gx, g->gy, g->gz);
  return DBEV_OK;
}

struct SortedSegments {
  uint32_t* skeys;
  uint32_t* sidx;
  int* flags;
  int* excl;
  int* head_pos;  // [n + 2]
  int* nseg;      // device scalar: number of valid segments
};

}  // namespace

// sorted keys -> head flags, exclusive scan (segment ordinal), head positions [nseg + 1]
// (last entry = one past the last valid row), *nseg = number of valid segments
int segment_sorted_keys(const uint32_t* skeys, const uint32_t* sidx, int n, uint32_t sentinel,
                        int* flags, int* excl, int* head_pos, int* nseg, void* sub, size_t sub_bytes,
                        cudaStream_t stream) {
  const int grid = ceil_div(n, 256);
  head_flags_kernel<<<grid, 256, 0, stream>>>(skeys, n, sentinel, flags);
  int rc = exclusive_scan_i32(flags, excl, n, nseg, sub, sub_bytes, stream);
  if (rc != DBEV_OK) return rc;
  segment_heads_kernel<<<grid, 256, 0, stream>>>(skeys, sidx, flags, excl, n, sentinel, head_pos,
                                                 nullptr);
  set_tail_marker_kernel<<<1, 32, 0, stream>>>(skeys, n, sentinel, nseg, head_pos);
  DBEV_CHECK_LAUNCH("segment_sorted_keys");
  return DBEV_OK;
}

namespace {

// keys0 holds the unsorted keys (taken from `w` by the caller)
int sort_and_segment(uint32_t* keys0, int n, unsigned long long nkeys, uint32_t sentinel,
                     Workspace& w, void* ws, size_t ws_bytes, cudaStream_t stream,
                     int* first_flag, SortedSegments* out) {
  uint32_t* keys1 = w.take<uint32_t>(n);
  uint32_t* vals0 = w.take<uint32_t>(n);
  uint32_t* vals1 = w.take<uint32_t>(n);
  int* flags = w.take<int>(n);
  int* excl = w.take<int>(n);
  int* head_pos = w.take<int>((size_t)n + 2);
  int* nseg = w.take<int>(1);
  if (!w.ok()) {
    set_last_error("voxelize: workspace too small (%zu given, %zu needed so far)", ws_bytes, w.used);
    return DBEV_ERR_WORKSPACE;
  }
  const size_t consumed = align_up(w.used);
  void* sub = (char*)ws + consumed;
  const size_t sub_bytes = ws_bytes > consumed ? ws_bytes - consumed : 0;
  uint32_t* keys[2] = {keys0, keys1};
  uint32_t* vals[2] = {vals0, vals1};
  int sel = 0;
  int rc = radix_sort_pairs(keys, vals, true, n, bits_for(nkeys + 1), sub, sub_bytes, stream, &sel);
  if (rc != DBEV_OK) return rc;
  const int grid = ceil_div(n, 256);
  head_flags_kernel<<<grid, 256, 0, stream>>>(keys[sel], n, sentinel, flags);
  rc = exclusive_scan_i32(flags, excl, n, nseg, sub, sub_bytes, stream);
  if (rc != DBEV_OK) return rc;
  segment_heads_kernel<<<grid, 256, 0, stream>>>(keys[sel], vals[sel], flags, excl, n, sentinel,
                                                 head_pos, first_flag);
  set_tail_marker_kernel<<<1, 32, 0, stream>>>(keys[sel], n, sentinel, nseg, head_pos);
  DBEV_CHECK_LAUNCH("sort_and_segment");
  out->skeys = keys[sel];
  out->sidx = vals[sel];
  out->flags = flags;
  out->excl = excl;
  out->head_pos = head_pos;
  out->nseg = nseg;
  return DBEV_OK;
}

size_t segment_ws_bytes(long long n) {
  return 7 * align_up((size_t)(n + 2) * 4) + radix_sort_ws_bytes(n) + scan_ws_bytes(n) + 4096;
}

}  // namespace

// ---------------------------------------------------------------------------

int voxel_grid_size(const float* voxel_size, const float* coors_range, int* grid_xyz) {
  VoxGrid g;
  int rc = make_grid(voxel_size, coors_range, &g);
  if (rc != DBEV_OK) return rc;
  grid_xyz[0] = g.gx; grid_xyz[1] = g.gy; grid_xyz[2] = g.gz;
  return DBEV_OK;
}

int dynamic_voxelize(const float* points, int n, int nfeat, const float* voxel_size,
                     const float* coors_range, int* coors, cudaStream_t stream) {
  DBEV_CHECK_ARG(n >= 0 && nfeat >= 3, "dynamic_voxelize: need n >= 0 and >= 3 features (got %d, %d)",
                 n, nfeat);
  VoxGrid g;
  int rc = make_grid(voxel_size, coors_range, &g);
  if (rc != DBEV_OK) return rc;
  if (n == 0) return DBEV_OK;
  dynamic_voxelize_kernel<<<ceil_div(n, 256), 256, 0, stream>>>(points, n, nfeat, g, coors, nullptr, 0);
  DBEV_CHECK_LAUNCH("dynamic_voxelize_kernel");
  return DBEV_OK;
}

size_t hard_voxelize_ws_bytes(long long n) {
  return segment_ws_bytes(n) + 3 * align_up((size_t)(n + 1) * 4);
}

int hard_voxelize(const float* points, int n, int nfeat, const float* voxel_size,
                  const float* coors_range, int max_points, int max_voxels, float* voxels,
                  int* coors, int* num_points_per_voxel, int* voxel_num, void* ws, size_t ws_bytes,
                  cudaStream_t stream) {
  DBEV_CHECK_ARG(n >= 0 && nfeat >= 3, "hard_voxelize: need n >= 0 and >= 3 features");
  DBEV_CHECK_ARG(max_points > 0 && max_voxels > 0,
                 "hard_voxelize: max_points / max_voxels must be positive (use dynamic_voxelize for -1)");
  VoxGrid g;
  int rc = make_grid(voxel_size, coors_range, &g);
  if (rc != DBEV_OK) return rc;
  if (n == 0) {
    DBEV_CUDA(cudaMemsetAsync(voxel_num, 0, sizeof(int), stream));
    return DBEV_OK;
  }
  const unsigned long long nkeys = (unsigned long long)g.gx * g.gy * g.gz;
  const uint32_t sentinel = (uint32_t)nkeys;
  Workspace w(ws, ws_bytes);
  uint32_t* keys0 = w.take<uint32_t>(n);
  int* first_flag = w.take<int>(n);
  int* vid_of_point = w.take<int>(n);
  if (!w.ok()) {
    set_last_error("hard_voxelize: workspace too small");
    return DBEV_ERR_WORKSPACE;
  }
  const int grid = ceil_div(n, 256);
  dynamic_voxelize_kernel<<<grid, 256, 0, stream>>>(points, n, nfeat, g, nullptr, keys0, sentinel);
  DBEV_CUDA(cudaMemsetAsync(first_flag, 0, (size_t)n * sizeof(int), stream));
  SortedSegments seg;
  rc = sort_and_segment(keys0, n, nkeys, sentinel, w, ws, ws_bytes, stream, first_flag, &seg);
  if (rc != DBEV_OK) return rc;
  // voxel id of a voxel = number of voxels whose first point comes earlier
  {
    const size_t consumed = align_up(w.used);
    rc = exclusive_scan_i32(first_flag, vid_of_point, n, nullptr, (char*)ws + consumed,
                            ws_bytes > consumed ? ws_bytes - consumed : 0, stream);
    if (rc != DBEV_OK) return rc;
  }
  hard_scatter_kernel<<<grid, 256, 0, stream>>>(points, nfeat, seg.skeys, seg.sidx, seg.flags,
                                                seg.excl, seg.head_pos, vid_of_point, n, sentinel,
                                                g, max_points, max_voxels, voxels, coors,
                                                num_points_per_voxel);
  clamp_count_kernel<<<1, 32, 0, stream>>>(seg.nseg, max_voxels, voxel_num);
  DBEV_CHECK_LAUNCH("hard_voxelize");
  return DBEV_OK;
}

size_t dynamic_scatter_ws_bytes(long long n) { return segment_ws_bytes(n) + align_up((size_t)(n + 1) * 4); }

int dynamic_scatter_forward(const float* feats, const int* coors, int n, int nfeat, int ncol,
                            const int* dims, int reduce_type, float* reduced_feats, int* out_coors,
                            int* coors_map, int* reduce_count, int* num_out, void* ws,
                            size_t ws_bytes, cudaStream_t stream) {
  DBEV_CHECK_ARG(n >= 0 && nfeat > 0, "dynamic_scatter: bad sizes n=%d nfeat=%d", n, nfeat);
  DBEV_CHECK_ARG(ncol == 3 || ncol == 4, "dynamic_scatter: coors must have 3 or 4 columns (got %d)",
                 ncol);
  DBEV_CHECK_ARG(reduce_type >= 0 && reduce_type <= 2, "dynamic_scatter: reduce_type %d (0 sum, 1 mean, 2 max)",
                 reduce_type);
  unsigned long long nkeys = 1;
  for (int i = 0; i < ncol; ++i) {
    DBEV_CHECK_ARG(dims[i] > 0, "dynamic_scatter: dims[%d] must be positive", i);
    nkeys *= (unsigned long long)dims[i];
  }
  DBEV_CHECK_ARG(nkeys < 0xfffffff0ULL, "dynamic_scatter: coordinate space exceeds 32-bit keys");
  if (n == 0) {
    DBEV_CUDA(cudaMemsetAsync(num_out, 0, sizeof(int), stream));
    return DBEV_OK;
  }
  const uint32_t sentinel = (uint32_t)nkeys;
  Workspace w(ws, ws_bytes);
  uint32_t* keys0 = w.take<uint32_t>(n);
  if (!w.ok()) {
    set_last_error("dynamic_scatter: workspace too small");
    return DBEV_ERR_WORKSPACE;
  }
  const int grid = ceil_div(n, 256);
  coors_to_keys_kernel<<<grid, 256, 0, stream>>>(coors, n, ncol, dims[0], dims[1], dims[2],
                                                 ncol == 4 ? dims[3] : 1, sentinel, keys0);
  SortedSegments seg;
  int rc = sort_and_segment(keys0, n, nkeys, sentinel, w, ws, ws_bytes, stream, nullptr, &seg);
  if (rc != DBEV_OK) return rc;
  scatter_meta_kernel<<<grid, 256, 0, stream>>>(seg.skeys, seg.sidx, seg.flags, seg.excl,
                                                seg.head_pos, n, sentinel, ncol, dims[1], dims[2],
                                                ncol == 4 ? dims[3] : 1, out_coors, coors_map,
                                                reduce_count);
  scatter_reduce_kernel<<<kNumSMs * 8, 256, 0, stream>>>(feats, nfeat, seg.sidx, seg.head_pos,
                                                        seg.nseg, reduce_type, reduced_feats);
  DBEV_CUDA(cudaMemcpyAsync(num_out, seg.nseg, sizeof(int), cudaMemcpyDeviceToDevice, stream));
  DBEV_CHECK_LAUNCH("dynamic_scatter_forward");
  return DBEV_OK;
}

int dynamic_scatter_backward(const float* grad_reduced, const float* feats, const float* reduced,
                             const int* coors_map, const int* reduce_count, long long n,
                             long long m, int nfeat, int reduce_type, float* grad_feats,
                             int* reduce_from_ws, cudaStream_t stream) {
  DBEV_CHECK_ARG(n >= 0 && m >= 0 && nfeat > 0, "dynamic_scatter_backward: bad sizes");
  if (n == 0) return DBEV_OK;
  if (reduce_type == 0 || reduce_type == 1) {
    scatter_bwd_add_kernel<<<kNumSMs * 8, 256, 0, stream>>>(grad_reduced, coors_map, reduce_count,
                                                            n, nfeat, reduce_type, grad_feats);
  } else {
    DBEV_CHECK_ARG(reduce_from_ws != nullptr, "dynamic_scatter_backward(max): needs m*nfeat int workspace");
    DBEV_CUDA(cudaMemsetAsync(grad_feats, 0, (size_t)n * nfeat * sizeof(float), stream));
    if (m > 0) {
      fill_i32_kernel<<<kNumSMs * 4, 256, 0, stream>>>(reduce_from_ws, m * nfeat, (int)n);
      scatter_bwd_max_trace_kernel<<<kNumSMs * 8, 256, 0, stream>>>(feats, reduced, coors_map, n,
                                                                    nfeat, reduce_from_ws);
      scatter_bwd_max_apply_kernel<<<kNumSMs * 8, 256, 0, stream>>>(grad_reduced, reduce_from_ws, m,
                                                                    nfeat, n, grad_feats);
    }
  }
  DBEV_CHECK_LAUNCH("dynamic_scatter_backward");
  return DBEV_OK;
}

}  // namespace dbev
