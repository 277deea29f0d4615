// CenterHead training targets on the device; see center_targets.cu.
#pragma once

#include "common.cuh"

namespace dbev {

int center_targets(const float* boxes, int box_dim, const int* labels, const int* offsets, int batch,
                   const int* class_task_host, const int* class_in_task_host, int num_classes,
                   int num_tasks, int max_objs, int H, int W, float voxel_x, float voxel_y,
                   float out_size_factor, float pc_min_x, float pc_min_y, float gaussian_overlap,
                   int min_radius, int norm_bbox, float* heatmap, float* anno_box, long long* ind,
                   unsigned char* mask, cudaStream_t stream);

}  // namespace dbev
