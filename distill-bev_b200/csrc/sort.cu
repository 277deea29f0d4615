// Stable LSD radix sort (8-bit digits) + int32 exclusive scan. See sort.cuh.
//
// One pass = three launches:
//   radix_hist_kernel    per-CTA digit histogram -> table[digit][cta]
//   radix_rowscan_kernel one CTA per digit: exclusive scan along the cta axis,
//                        digit totals to totals[digit]
//   radix_scatter_kernel re-reads the tile, ranks every key among equal
//                        digits (warp match_any + per-warp counters, so the
//                        order inside a CTA is (warp, round, lane) = input
//                        order -> stable), scatters key and payload.
// All traffic is 4/8-byte streaming reads and digit-clustered writes; the
// whole thing is HBM/L2 bound: 2 reads + 1 write of 8 B per key per pass.
#include "sort.cuh"

namespace dbev {

namespace {

constexpr int kWarps = kSortBlock / 32;

__global__ void __launch_bounds__(kSortBlock)
radix_hist_kernel(const uint32_t* __restrict__ keys, int n, int shift, uint32_t mask,
                  uint32_t* __restrict__ table, int nblocks) {
  __shared__ uint32_t sh[kRadix];
  sh[threadIdx.x] = 0;
  __syncthreads();
  const long long base = (long long)blockIdx.x * kSortTile;
#pragma unroll 4
  for (int i = 0; i < kSortItems; ++i) {
    long long idx = base + (long long)i * kSortBlock + threadIdx.x;
    uint32_t d = (idx < n) ? ((keys[idx] >> shift) & mask) : 0xffffffffu;
    unsigned peers = __match_any_sync(0xffffffffu, d);
    if (d != 0xffffffffu && (peers & lanemask_lt()) == 0) atomicAdd(&sh[d], __popc(peers));
  }
  __syncthreads();
  table[(size_t)threadIdx.x * nblocks + blockIdx.x] = sh[threadIdx.x];
}

// grid = kRadix CTAs; CTA d scans table[d][0..nblocks) in place (exclusive).
__global__ void __launch_bounds__(256)
radix_rowscan_kernel(uint32_t* __restrict__ table, uint32_t* __restrict__ totals,
                     int nblocks) {
  __shared__ uint32_t warp_tot[8];
  __shared__ uint32_t carry_sh;
  uint32_t* row = table + (size_t)blockIdx.x * nblocks;
  if (threadIdx.x == 0) carry_sh = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int base = 0; base < nblocks; base += 256) {
    int i = base + threadIdx.x;
    uint32_t v = (i < nblocks) ? row[i] : 0;
    uint32_t incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    uint32_t woff = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w)
      if (w < warp) woff += warp_tot[w];
    uint32_t carry = carry_sh;
    if (i < nblocks) row[i] = carry + woff + incl - v;
    __syncthreads();
    if (threadIdx.x == 255) carry_sh = carry + woff + incl;
    __syncthreads();
  }
  if (threadIdx.x == 0) totals[blockIdx.x] = carry_sh;
}

__global__ void __launch_bounds__(kSortBlock)
radix_scatter_kernel(const uint32_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
                     uint32_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out, int n,
                     int shift, uint32_t mask, const uint32_t* __restrict__ table,
                     const uint32_t* __restrict__ totals, int nblocks) {
  __shared__ uint32_t warp_hist[kWarps][kRadix];
  __shared__ uint32_t digit_base[kRadix];
  __shared__ uint32_t scan_tmp[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

  // digit_base[d] = exclusive_scan(totals)[d] + table[d][cta]
  {
    uint32_t t = totals[threadIdx.x];
    uint32_t incl = t;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t u = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += u;
    }
    if (lane == 31) scan_tmp[warp] = incl;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) warp_hist[w][threadIdx.x] = 0;
    __syncthreads();
    uint32_t woff = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w)
      if (w < warp) woff += scan_tmp[w];
    digit_base[threadIdx.x] =
        woff + incl - t + table[(size_t)threadIdx.x * nblocks + blockIdx.x];
  }
  __syncthreads();

  const long long wbase =
      (long long)blockIdx.x * kSortTile + (long long)warp * (32 * kSortItems);
  uint32_t key[kSortItems];
  uint32_t rank[kSortItems];
#pragma unroll
  for (int r = 0; r < kSortItems; ++r) {
    long long idx = wbase + r * 32 + lane;
    key[r] = (idx < n) ? keys_in[idx] : 0u;
  }
#pragma unroll
  for (int r = 0; r < kSortItems; ++r) {
    long long idx = wbase + r * 32 + lane;
    const bool valid = idx < n;
    uint32_t d = valid ? ((key[r] >> shift) & mask) : 0xffffffffu;
    unsigned peers = __match_any_sync(0xffffffffu, d);
    uint32_t before = valid ? warp_hist[warp][d] : 0;
    __syncwarp();
    if (valid && (peers & lanemask_lt()) == 0) warp_hist[warp][d] = before + __popc(peers);
    __syncwarp();
    rank[r] = before + __popc(peers & lanemask_lt());
  }
  __syncthreads();
  {
    // exclusive prefix over warps for digit = threadIdx.x, seeded by global base
    uint32_t running = digit_base[threadIdx.x];
#pragma unroll
    for (int w = 0; w < kWarps; ++w) {
      uint32_t t = warp_hist[w][threadIdx.x];
      warp_hist[w][threadIdx.x] = running;
      running += t;
    }
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < kSortItems; ++r) {
    long long idx = wbase + r * 32 + lane;
    if (idx < n) {
      uint32_t d = (key[r] >> shift) & mask;
      uint32_t pos = warp_hist[warp][d] + rank[r];
      keys_out[pos] = key[r];
      vals_out[pos] = vals_in ? vals_in[idx] : (uint32_t)idx;
    }
  }
}

// ---------------------------------------------------------------- scan -----

constexpr int kScanBlock = 256;
constexpr int kScanItems = 8;
constexpr int kScanTile = kScanBlock * kScanItems;

__device__ __forceinline__ int block_exclusive_scan(int thread_sum, int* block_total) {
  __shared__ int warp_tot[kScanBlock / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int incl = thread_sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) warp_tot[warp] = incl;
  __syncthreads();
  int woff = 0, tot = 0;
#pragma unroll
  for (int w = 0; w < kScanBlock / 32; ++w) {
    int t = warp_tot[w];
    if (w < warp) woff += t;
    tot += t;
  }
  *block_total = tot;
  return woff + incl - thread_sum;
}

__global__ void __launch_bounds__(kScanBlock)
scan_tile_reduce_kernel(const int* __restrict__ in, int n, int* __restrict__ sums) {
  const long long base = (long long)blockIdx.x * kScanTile + (long long)threadIdx.x * kScanItems;
  int s = 0;
#pragma unroll
  for (int i = 0; i < kScanItems; ++i)
    if (base + i < n) s += in[base + i];
  int tot;
  block_exclusive_scan(s, &tot);
  if (threadIdx.x == 0) sums[blockIdx.x] = tot;
}

__global__ void __launch_bounds__(kScanBlock)
scan_tile_downsweep_kernel(const int* in, int* out, int n,
                           const int* __restrict__ tile_offsets, int* __restrict__ total_out) {
  const long long base = (long long)blockIdx.x * kScanTile + (long long)threadIdx.x * kScanItems;
  int v[kScanItems];
  int s = 0;
#pragma unroll
  for (int i = 0; i < kScanItems; ++i) {
    v[i] = (base + i < n) ? in[base + i] : 0;
    s += v[i];
  }
  int tot;
  int excl = block_exclusive_scan(s, &tot);
  const int off = tile_offsets ? tile_offsets[blockIdx.x] : 0;
  int run = off + excl;
#pragma unroll
  for (int i = 0; i < kScanItems; ++i) {
    if (base + i < n) out[base + i] = run;
    run += v[i];
  }
  if (total_out && blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) *total_out = off + tot;
}

}  // namespace

size_t radix_sort_ws_bytes(long long n) {
  long long nblocks = (n + kSortTile - 1) / kSortTile;
  if (nblocks < 1) nblocks = 1;
  return align_up((size_t)nblocks * kRadix * sizeof(uint32_t)) + align_up(kRadix * sizeof(uint32_t));
}

int radix_sort_pairs(uint32_t* keys[2], uint32_t* vals[2], bool vals_iota, int n, int num_bits,
                     void* ws, size_t ws_bytes, cudaStream_t stream, int* out_sel) {
  *out_sel = 0;
  if (n <= 0) return DBEV_OK;
  const int nblocks = ceil_div(n, kSortTile);
  Workspace w(ws, ws_bytes);
  uint32_t* table = w.take<uint32_t>((size_t)nblocks * kRadix);
  uint32_t* totals = w.take<uint32_t>(kRadix);
  if (!w.ok()) {
    set_last_error("radix_sort_pairs: workspace too small (%zu < %zu)", ws_bytes, w.used);
    return DBEV_ERR_WORKSPACE;
  }
  if (num_bits < 1) num_bits = 1;
  const int passes = (num_bits + kRadixBits - 1) / kRadixBits;
  int sel = 0;
  for (int p = 0; p < passes; ++p) {
    const int shift = p * kRadixBits;
    const int bits = (num_bits - shift < kRadixBits) ? (num_bits - shift) : kRadixBits;
    const uint32_t mask = (1u << bits) - 1u;
    const uint32_t* vin = (p == 0 && vals_iota) ? nullptr : vals[sel];
    radix_hist_kernel<<<nblocks, kSortBlock, 0, stream>>>(keys[sel], n, shift, mask, table, nblocks);
    radix_rowscan_kernel<<<kRadix, 256, 0, stream>>>(table, totals, nblocks);
    radix_scatter_kernel<<<nblocks, kSortBlock, 0, stream>>>(
        keys[sel], vin, keys[sel ^ 1], vals[sel ^ 1], n, shift, mask, table, totals, nblocks);
    sel ^= 1;
  }
  DBEV_CHECK_LAUNCH("radix_sort_pairs");
  *out_sel = sel;
  return DBEV_OK;
}

size_t scan_ws_bytes(long long n) {
  size_t total = 0;
  long long m = n;
  while (m > kScanTile) {
    m = (m + kScanTile - 1) / kScanTile;
    total += align_up((size_t)m * sizeof(int));
  }
  return total + 256;
}

int exclusive_scan_i32(const int* in, int* out, int n, int* total_out, void* ws, size_t ws_bytes,
                       cudaStream_t stream) {
  if (n <= 0) {
    if (total_out) DBEV_CUDA(cudaMemsetAsync(total_out, 0, sizeof(int), stream));
    return DBEV_OK;
  }
  const int ntiles = ceil_div(n, kScanTile);
  if (ntiles == 1) {
    scan_tile_downsweep_kernel<<<1, kScanBlock, 0, stream>>>(in, out, n, nullptr, total_out);
    DBEV_CHECK_LAUNCH("exclusive_scan_i32");
    return DBEV_OK;
  }
  Workspace w(ws, ws_bytes);
  int* sums = w.take<int>(ntiles);
  if (!w.ok()) {
    set_last_error("exclusive_scan_i32: workspace too small");
    return DBEV_ERR_WORKSPACE;
  }
  scan_tile_reduce_kernel<<<ntiles, kScanBlock, 0, stream>>>(in, n, sums);
  size_t consumed = align_up(w.used);
  int rc = exclusive_scan_i32(sums, sums, ntiles, nullptr, (char*)ws + consumed,
                              ws_bytes > consumed ? ws_bytes - consumed : 0, stream);
  if (rc != DBEV_OK) return rc;
  scan_tile_downsweep_kernel<<<ntiles, kScanBlock, 0, stream>>>(in, out, n, sums, total_out);
  DBEV_CHECK_LAUNCH("exclusive_scan_i32");
  return DBEV_OK;
}

}  // namespace dbev
