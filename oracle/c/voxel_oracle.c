/*
 * CPU oracle for the LiDAR voxelization path — plain C restatement.
 *
 * TEST INFRASTRUCTURE ONLY: linked / loaded by tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline leg, never by the product (distill-bev_b200/).
 *
 * Follows, function by function (paths relative to the reference checkout):
 *   vo_dynamic_voxelize  mmdet3d/ops/voxel/src/voxelization_cpu.cpp:8-43   (kernel)
 *                        mmdet3d/ops/voxel/src/voxelization_cpu.cpp:146-171 (grid size)
 *   vo_hard_voxelize     mmdet3d/ops/voxel/src/voxelization_cpu.cpp:45-105,107-144
 *   vo_dynamic_scatter   mmdet3d/ops/voxel/src/scatter_points_cuda.cu:183-239 with
 *                        feats_reduce_kernel :81-103 (the reference has no CPU
 *                        binding for this op, voxelization.h:118; at::unique_dim
 *                        = rows sorted lexicographically, inverse map, counts)
 *   vo_dynamic_scatter_backward  scatter_points_cuda.cu:106-179,241-308
 *
 * Parity pin: tests/test_oracle_voxel.py checks this file against (a) the
 * reference's own CPU extension compiled unmodified into oracle/_ref/ when it
 * is present and (b) fixtures generated from it (tests/golden/voxel_*.npz,
 * tools/make_golden.py).
 *
 * All arithmetic on coordinates is float32, as in the reference (scalar_t =
 * float, voxel_size / coors_range are std::vector<float>).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

static void vo_grid(const float* voxel_size, const float* coors_range, int* grid) {
  for (int i = 0; i < 3; ++i)
    grid[i] = (int)roundf((coors_range[3 + i] - coors_range[i]) / voxel_size[i]);
}

void vo_grid_size(const float* voxel_size, const float* coors_range, int* grid_xyz) {
  vo_grid(voxel_size, coors_range, grid_xyz);
}

/* coors[n][3] = (z, y, x) or (-1, -1, -1) */
void vo_dynamic_voxelize(const float* points, int n, int nfeat, const float* voxel_size,
                         const float* coors_range, int* coors) {
  int grid[3];
  vo_grid(voxel_size, coors_range, grid);
  for (int i = 0; i < n; ++i) {
    int coor[3];
    int failed = 0;
    for (int j = 0; j < 3; ++j) {
      volatile float d = points[(size_t)i * nfeat + j] - coors_range[j];
      volatile float q = d / voxel_size[j];
      int c = (int)floorf(q);
      if (c < 0 || c >= grid[j]) {
        failed = 1;
        break;
      }
      coor[2 - j] = c;
    }
    for (int k = 0; k < 3; ++k) coors[(size_t)i * 3 + k] = failed ? -1 : coor[k];
  }
}

/* returns voxel_num; voxels[max_voxels][max_points][nfeat], coors[max_voxels][3],
 * num_points_per_voxel[max_voxels] must be zero-initialised by the caller
 * (voxelize.py:57-62). */
int vo_hard_voxelize(const float* points, int n, int nfeat, const float* voxel_size,
                     const float* coors_range, int max_points, int max_voxels, float* voxels,
                     int* coors, int* num_points_per_voxel) {
  int grid[3];
  vo_grid(voxel_size, coors_range, grid);
  int* temp = (int*)malloc((size_t)(n > 0 ? n : 1) * 3 * sizeof(int));
  vo_dynamic_voxelize(points, n, nfeat, voxel_size, coors_range, temp);
  size_t cells = (size_t)grid[0] * grid[1] * grid[2];
  int* coor_to_voxelidx = (int*)malloc(cells * sizeof(int));
  for (size_t i = 0; i < cells; ++i) coor_to_voxelidx[i] = -1;
  int voxel_num = 0;
  for (int i = 0; i < n; ++i) {
    const int* c = temp + (size_t)i * 3;
    if (c[0] == -1) continue;
    size_t cell = ((size_t)c[0] * grid[1] + c[1]) * grid[0] + c[2];
    int voxelidx = coor_to_voxelidx[cell];
    if (voxelidx == -1) {
      voxelidx = voxel_num;
      if (max_voxels != -1 && voxel_num >= max_voxels) continue;
      voxel_num += 1;
      coor_to_voxelidx[cell] = voxelidx;
      for (int k = 0; k < 3; ++k) coors[(size_t)voxelidx * 3 + k] = c[k];
    }
    int num = num_points_per_voxel[voxelidx];
    if (max_points == -1 || num < max_points) {
      memcpy(voxels + ((size_t)voxelidx * max_points + num) * nfeat, points + (size_t)i * nfeat,
             (size_t)nfeat * sizeof(float));
      num_points_per_voxel[voxelidx] += 1;
    }
  }
  free(temp);
  free(coor_to_voxelidx);
  return voxel_num;
}

/* ---- dynamic scatter ------------------------------------------------------ */

static int g_ncol;
static const int* g_rows;

static int cmp_rows(const void* a, const void* b) {
  const int* ra = g_rows + (size_t)(*(const int*)a) * g_ncol;
  const int* rb = g_rows + (size_t)(*(const int*)b) * g_ncol;
  for (int k = 0; k < g_ncol; ++k) {
    if (ra[k] != rb[k]) return ra[k] < rb[k] ? -1 : 1;
  }
  /* stable: ties by original index */
  return *(const int*)a - *(const int*)b;
}

/* returns M. reduce_type: 0 sum, 1 mean, 2 max. Outputs sized for n rows.
 * coors rows with any negative component are cleaned to all -1 and dropped.
 * The sum is accumulated in point order (float), the reference's atomics run in
 * arbitrary order (its own docstring: differences ~5e-7, scatter_points.py:59-60). */
int vo_dynamic_scatter(const float* feats, const int* coors, int n, int nfeat, int ncol,
                       int reduce_type, float* reduced, int* out_coors, int* coors_map,
                       int* reduce_count) {
  if (n == 0) return 0;
  int* clean = (int*)malloc((size_t)n * ncol * sizeof(int));
  int* idx = (int*)malloc((size_t)n * sizeof(int));
  for (int i = 0; i < n; ++i) {
    int neg = 0;
    for (int k = 0; k < ncol; ++k) neg |= coors[(size_t)i * ncol + k] < 0;
    for (int k = 0; k < ncol; ++k) clean[(size_t)i * ncol + k] = neg ? -1 : coors[(size_t)i * ncol + k];
    idx[i] = i;
  }
  g_ncol = ncol;
  g_rows = clean;
  qsort(idx, (size_t)n, sizeof(int), cmp_rows);
  int m = 0;
  for (int j = 0; j < n; ++j) {
    const int* row = clean + (size_t)idx[j] * ncol;
    if (row[0] < 0) {
      coors_map[idx[j]] = -1;
      continue;
    }
    int is_new = (m == 0) || memcmp(row, out_coors + (size_t)(m - 1) * ncol, (size_t)ncol * sizeof(int)) != 0;
    if (is_new) {
      memcpy(out_coors + (size_t)m * ncol, row, (size_t)ncol * sizeof(int));
      reduce_count[m] = 0;
      for (int c = 0; c < nfeat; ++c) reduced[(size_t)m * nfeat + c] = (reduce_type == 2) ? -INFINITY : 0.f;
      ++m;
    }
    coors_map[idx[j]] = m - 1;
    reduce_count[m - 1] += 1;
  }
  /* reduce in POINT order (ascending i) */
  for (int i = 0; i < n; ++i) {
    int to = coors_map[i];
    if (to < 0) continue;
    for (int c = 0; c < nfeat; ++c) {
      float v = feats[(size_t)i * nfeat + c];
      float* d = reduced + (size_t)to * nfeat + c;
      if (reduce_type == 2) *d = fmaxf(*d, v);
      else *d = *d + v;
    }
  }
  if (reduce_type == 1)
    for (int s = 0; s < m; ++s)
      for (int c = 0; c < nfeat; ++c) reduced[(size_t)s * nfeat + c] /= (float)reduce_count[s];
  free(clean);
  free(idx);
  return m;
}

void vo_dynamic_scatter_backward(const float* grad_reduced, const float* feats,
                                 const float* reduced, const int* coors_map,
                                 const int* reduce_count, int n, int m, int nfeat,
                                 int reduce_type, float* grad_feats) {
  memset(grad_feats, 0, (size_t)n * nfeat * sizeof(float));
  if (reduce_type == 0 || reduce_type == 1) {
    for (int i = 0; i < n; ++i) {
      int to = coors_map[i];
      if (to < 0) continue;
      for (int c = 0; c < nfeat; ++c) {
        float g = grad_reduced[(size_t)to * nfeat + c];
        if (reduce_type == 1) g = g / (float)reduce_count[to];
        grad_feats[(size_t)i * nfeat + c] = g;
      }
    }
  } else {
    int* from = (int*)malloc((size_t)(m > 0 ? m : 1) * nfeat * sizeof(int));
    for (size_t t = 0; t < (size_t)m * nfeat; ++t) from[t] = n;
    for (int i = 0; i < n; ++i) {
      int to = coors_map[i];
      if (to < 0) continue;
      for (int c = 0; c < nfeat; ++c)
        if (feats[(size_t)i * nfeat + c] == reduced[(size_t)to * nfeat + c] &&
            i < from[(size_t)to * nfeat + c])
          from[(size_t)to * nfeat + c] = i;
    }
    for (int s = 0; s < m; ++s)
      for (int c = 0; c < nfeat; ++c) {
        int f = from[(size_t)s * nfeat + c];
        if (f < n) grad_feats[(size_t)f * nfeat + c] = grad_reduced[(size_t)s * nfeat + c];
      }
    free(from);
  }
}
