// Device-wide building blocks shared by every many-to-one op on the hot path:
// a stable LSD radix sort of (key, payload) pairs and an int32 exclusive scan.
//
// The reference gets the same effect from library calls on its host path:
//   * ranks.argsort()                    mmdet3d/ops/bev_pool/bev_pool.py:92,
//                                         mmdet3d/models/necks/view_transformer_mine.py:168
//   * at::unique_dim(coors, 0, ...)      mmdet3d/ops/voxel/src/scatter_points_cuda.cu:204-205
//   * the O(N^2) point_to_voxelidx scan  mmdet3d/ops/voxel/src/voxelization_cuda.cu:106-147
// Because the sort here is STABLE and the payload starts as iota, equal keys
// keep ascending point order, which is exactly the first-appearance order the
// reference's deterministic hard-voxelize defines and makes every fp32
// segment sum run in one fixed order (bit-reproducible run to run).
#pragma once

#include "common.cuh"

namespace dbev {

constexpr int kSortBlock = 256;                      // 8 warps
constexpr int kSortItems = 16;                       // keys per thread
constexpr int kSortTile = kSortBlock * kSortItems;   // 4096 keys per CTA
constexpr int kRadixBits = 8;
constexpr int kRadix = 1 << kRadixBits;

// Workspace for radix_sort_pairs over n keys (per-pass tile descriptors + histograms).
size_t radix_sort_ws_bytes(long long n);

// Stable sort by key bits [0, num_bits). Input is keys[0] (+ vals[0] unless
// vals_iota, in which case the payload is the element index). Buffers
// ping-pong; *out_sel tells which of keys[]/vals[] holds the result.
int radix_sort_pairs(uint32_t* keys[2], uint32_t* vals[2], bool vals_iota, int n,
                     int num_bits, void* ws, size_t ws_bytes, cudaStream_t stream,
                     int* out_sel);

size_t scan_ws_bytes(long long n);

// out[i] = sum(in[0..i)), in and out may alias. If total_out != nullptr it
// receives sum(in[0..n)) (device pointer).
int exclusive_scan_i32(const int* in, int* out, int n, int* total_out, void* ws,
                       size_t ws_bytes, cudaStream_t stream);

inline int bits_for(unsigned long long count) {  // bits to represent [0, count)
  int b = 1;
  while ((1ull << b) < count) ++b;
  return b;
}

}  // namespace dbev
