// CenterHead training targets on the device.
// Reference: CenterHead.get_targets / get_targets_single
//   mmdet3d/models/dense_heads/centerpoint_head.py:400-445, 447-611
//   gaussian_radius / gaussian_2d / draw_heatmap_gaussian  mmdet3d/core/utils/gaussian.py:6-87
// The reference loops in Python over samples, tasks and up to 500 objects with several tiny tensor
// ops per object (radius, centre, slice max, torch.cat of 10 scalars): host-bound. Here one CTA per
// sample: a thread per object finds its task / slot (class order in the task, then original order,
// as the torch.where + cat sequence does), writes anno_box / ind / mask and draws its Gaussian with
// an integer atomicMax on the non-negative float bits (max-combine is order independent ->
// deterministic). The class heat maps of all tasks are one [B, num_classes, H, W] tensor.
#include "center_targets.cuh"

namespace dbev {

namespace {

struct TargetCfg {
  int num_classes, num_tasks, max_objs, H, W, box_dim, norm_bbox;
  float vx, vy, osf, pc_x, pc_y, min_radius;
  float f_1m, f_1p, f_a3x4, f_nb3, f_m1;  // (1-mo), (1+mo), 4*(4*mo), -2*mo, (mo-1) rounded to fp32
  int class_task[32], class_in_task[32];
};

// gaussian_radius (gaussian.py:51-87) in the reference's fp32 order of operations
__device__ float gaussian_radius_ref(float height, float width, const TargetCfg& c) {
  const float hw = __fadd_rn(height, width);
  const float b1 = hw;
  const float c1 = __fdiv_rn(__fmul_rn(__fmul_rn(width, height), c.f_1m), c.f_1p);
  const float sq1 = __fsqrt_rn(__fsub_rn(__fmul_rn(b1, b1), __fmul_rn(4.f, c1)));
  const float r1 = __fdiv_rn(__fadd_rn(b1, sq1), 2.f);
  const float b2 = __fmul_rn(2.f, hw);
  const float c2 = __fmul_rn(__fmul_rn(c.f_1m, width), height);
  const float sq2 = __fsqrt_rn(__fsub_rn(__fmul_rn(b2, b2), __fmul_rn(16.f, c2)));
  const float r2 = __fdiv_rn(__fadd_rn(b2, sq2), 2.f);
  const float b3 = __fmul_rn(c.f_nb3, hw);
  const float c3 = __fmul_rn(__fmul_rn(c.f_m1, width), height);
  const float sq3 = __fsqrt_rn(__fsub_rn(__fmul_rn(b3, b3), __fmul_rn(c.f_a3x4, c3)));
  const float r3 = __fdiv_rn(__fadd_rn(b3, sq3), 2.f);
  return fminf(r1, fminf(r2, r3));
}

__global__ void __launch_bounds__(256)
center_targets_kernel(const float* __restrict__ boxes, const int* __restrict__ labels,
                      const int* __restrict__ offsets, TargetCfg c, float* __restrict__ heatmap,
                      float* __restrict__ anno, long long* __restrict__ ind, unsigned char* __restrict__ mask) {
  const int b = blockIdx.x;
  const int o0 = offsets[b], m = offsets[b + 1] - o0;
  for (int i = threadIdx.x; i < m; i += blockDim.x) {
    const int lab = labels[o0 + i];
    if (lab < 0 || lab >= c.num_classes) continue;
    const int task = c.class_task[lab], cit = c.class_in_task[lab];
    int k = 0;  // slot: objects of earlier classes of the task first, then original order
    for (int j = 0; j < m; ++j) {
      const int lj = labels[o0 + j];
      if (lj < 0 || lj >= c.num_classes || c.class_task[lj] != task) continue;
      const int cj = c.class_in_task[lj];
      k += (cj < cit || (cj == cit && j < i)) ? 1 : 0;
    }
    if (k >= c.max_objs) continue;
    const float* bx = boxes + (long long)(o0 + i) * c.box_dim;
    const float width = __fdiv_rn(__fdiv_rn(bx[3], c.vx), c.osf);
    const float length = __fdiv_rn(__fdiv_rn(bx[4], c.vy), c.osf);
    if (!(width > 0.f && length > 0.f)) continue;
    int radius = (int)gaussian_radius_ref(length, width, c);
    radius = max((int)c.min_radius, radius);
    const float zc = __fadd_rn(bx[2], __fmul_rn(bx[5], 0.5f));   // gravity centre (lidar_box3d.py)
    const float coor_x = __fdiv_rn(__fdiv_rn(__fsub_rn(bx[0], c.pc_x), c.vx), c.osf);
    const float coor_y = __fdiv_rn(__fdiv_rn(__fsub_rn(bx[1], c.pc_y), c.vy), c.osf);
    const int x = (int)coor_x, y = (int)coor_y;
    if (!(x >= 0 && x < c.W && y >= 0 && y < c.H)) continue;
    // draw_heatmap_gaussian: sigma = diameter / 6 (double), max-combine
    const double sigma = (double)(2 * radius + 1) / 6.0;
    const double inv = 1.0 / (2.0 * sigma * sigma);
    float* hm = heatmap + ((long long)b * c.num_classes + lab) * c.H * c.W;
    const int left = min(x, radius), right = min(c.W - x, radius + 1);
    const int top = min(y, radius), bottom = min(c.H - y, radius + 1);
    for (int dy = -top; dy < bottom; ++dy)
      for (int dx = -left; dx < right; ++dx) {
        const float g = (float)exp(-(double)(dx * dx + dy * dy) * inv);
        atomicMax(reinterpret_cast<int*>(hm + (long long)(y + dy) * c.W + (x + dx)), __float_as_int(g));
      }
    const long long slot = ((long long)b * c.num_tasks + task) * c.max_objs + k;
    ind[slot] = (long long)y * c.W + x;
    mask[slot] = 1;
    float* a = anno + slot * 10;
    a[0] = coor_x - (float)x;
    a[1] = coor_y - (float)y;
    a[2] = zc;
    a[3] = c.norm_bbox ? logf(bx[3]) : bx[3];
    a[4] = c.norm_bbox ? logf(bx[4]) : bx[4];
    a[5] = c.norm_bbox ? logf(bx[5]) : bx[5];
    a[6] = sinf(bx[6]);
    a[7] = cosf(bx[6]);
    a[8] = c.box_dim > 7 ? bx[7] : 0.f;
    a[9] = c.box_dim > 8 ? bx[8] : 0.f;
  }
}

}  // namespace

int center_targets(const float* boxes, int box_dim, const int* labels, const int* offsets, int batch,
                   const int* class_task_host, const int* class_in_task_host, int num_classes,
                   int num_tasks, int max_objs, int H, int W, float voxel_x, float voxel_y,
                   float out_size_factor, float pc_min_x, float pc_min_y, float gaussian_overlap,
                   int min_radius, int norm_bbox, float* heatmap, float* anno_box, long long* ind,
                   unsigned char* mask, cudaStream_t stream) {
  DBEV_CHECK_ARG(batch > 0 && num_classes > 0 && num_classes <= 32 && num_tasks > 0 && max_objs > 0 && H > 0 &&
                     W > 0 && box_dim >= 7,
                 "center_targets: bad sizes (at most 32 classes)");
  TargetCfg c;
  c.num_classes = num_classes, c.num_tasks = num_tasks, c.max_objs = max_objs, c.H = H, c.W = W;
  c.box_dim = box_dim, c.norm_bbox = norm_bbox;
  c.vx = voxel_x, c.vy = voxel_y, c.osf = out_size_factor, c.pc_x = pc_min_x, c.pc_y = pc_min_y;
  c.min_radius = (float)min_radius;
  const double mo = (double)gaussian_overlap;
  c.f_1m = (float)(1.0 - mo), c.f_1p = (float)(1.0 + mo), c.f_a3x4 = (float)(4.0 * (4.0 * mo));
  c.f_nb3 = (float)(-2.0 * mo), c.f_m1 = (float)(mo - 1.0);
  for (int i = 0; i < 32; ++i) {
    c.class_task[i] = i < num_classes ? class_task_host[i] : 0;
    c.class_in_task[i] = i < num_classes ? class_in_task_host[i] : 0;
  }
  DBEV_CUDA(cudaMemsetAsync(heatmap, 0, (size_t)batch * num_classes * H * W * sizeof(float), stream));
  DBEV_CUDA(cudaMemsetAsync(anno_box, 0, (size_t)batch * num_tasks * max_objs * 10 * sizeof(float), stream));
  DBEV_CUDA(cudaMemsetAsync(ind, 0, (size_t)batch * num_tasks * max_objs * sizeof(long long), stream));
  DBEV_CUDA(cudaMemsetAsync(mask, 0, (size_t)batch * num_tasks * max_objs, stream));
  center_targets_kernel<<<batch, 256, 0, stream>>>(boxes, labels, offsets, c, heatmap, anno_box, ind, mask);
  DBEV_CHECK_LAUNCH("center_targets_kernel");
  return DBEV_OK;
}

}  // namespace dbev
