// Stable LSD radix sort (8-bit digits, "onesweep") + int32 exclusive scan. See sort.cuh.
//
// A sort over P digit passes is P + 1 launches:
//   radix_global_hist_kernel  one read of the keys -> global digit histogram of every pass
//   radix_onesweep_kernel x P  each CTA ranks its 4096-key tile (warp match_any + per-warp
//                              counters: order inside a tile = input order -> stable), publishes
//                              its digit counts and gets the counts of all preceding tiles by
//                              decoupled look-back, stages the tile sorted in shared memory and
//                              writes it out coalesced.
// Traffic per pass: one read + one write of (key, payload) = 16 B per element; the first pass
// generates the iota payload instead of reading it.
#include "sort.cuh"

namespace dbev {

namespace {

constexpr int kWarps = kSortBlock / 32;
constexpr uint32_t kFlagAgg = 1u << 30;    // tile aggregate published
constexpr uint32_t kFlagIncl = 2u << 30;   // inclusive prefix published
constexpr uint32_t kValueMask = (1u << 30) - 1u;

// Global digit histograms of every pass in ONE read of the keys (digit counts do not depend
// on the order of the keys, so they can all be taken from the unsorted input).
__global__ void __launch_bounds__(256)
radix_global_hist_kernel(const uint32_t* __restrict__ keys, int n, int passes, int num_bits,
                         uint32_t* __restrict__ hist /*[passes][kRadix]*/) {
  __shared__ uint32_t sh[4][kRadix];
  for (int i = threadIdx.x; i < 4 * kRadix; i += blockDim.x) (&sh[0][0])[i] = 0;
  __syncthreads();
  const int stride = gridDim.x * blockDim.x;
  for (int base = blockIdx.x * blockDim.x; base < n; base += stride * 4) {
    uint32_t k[4];
    bool ok[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int idx = base + u * stride + threadIdx.x;
      ok[u] = idx < n;
      k[u] = ok[u] ? keys[idx] : 0u;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      for (int p = 0; p < passes; ++p) {
        const int shift = p * kRadixBits;
        const int bits = min(kRadixBits, num_bits - shift);
        // digits of neighbouring keys are mostly distinct: plain shared atomics spread over the
        // banks beat warp aggregation here
        if (ok[u]) atomicAdd(&sh[p][(k[u] >> shift) & ((1u << bits) - 1u)], 1u);
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < passes * kRadix; i += blockDim.x) {
    const uint32_t v = (&sh[0][0])[i];
    if (v) atomicAdd(&hist[i], v);
  }
}

// lanes holding the same 8-bit digit, from kRadixBits ballots (fixed latency; the hardware
// match.any instruction serialises over the distinct values of the warp)
__device__ __forceinline__ unsigned match_digit(uint32_t d, bool valid) {
  unsigned peers = __ballot_sync(0xffffffffu, valid);
#pragma unroll
  for (int b = 0; b < kRadixBits; ++b) {
    const bool bit = (d >> b) & 1u;
    const unsigned m = __ballot_sync(0xffffffffu, bit);
    peers &= bit ? m : ~m;
  }
  return valid ? peers : 0u;
}

// block-wide exclusive scan of one value per thread (256 threads)
__device__ __forceinline__ uint32_t block_excl_scan_256(uint32_t v, uint32_t* tmp8, int lane, int warp) {
  uint32_t incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t u = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += u;
  }
  if (lane == 31) tmp8[warp] = incl;
  __syncthreads();
  uint32_t woff = 0;
#pragma unroll
  for (int w = 0; w < 8; ++w)
    if (w < warp) woff += tmp8[w];
  __syncthreads();
  return woff + incl - v;
}

__device__ __forceinline__ uint32_t ld_volatile_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}

// One radix pass in ONE kernel ("onesweep"): every CTA ranks its tile, publishes its digit
// counts and obtains the counts of all preceding tiles by decoupled look-back, so the keys are
// read once and written once per pass. Tile ids come from a ticket counter, which guarantees
// that every tile a CTA waits for is already running. Stable: tiles in order, inside a tile
// (warp, round, lane) = input order.
__global__ void __launch_bounds__(kSortBlock, 2)
radix_onesweep_kernel(const uint32_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
                      uint32_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out, int n,
                      int shift, uint32_t mask, const uint32_t* __restrict__ hist /*[kRadix]*/,
                      uint32_t* __restrict__ desc /*[ntiles][kRadix]*/, uint32_t* __restrict__ ticket) {
  __shared__ uint32_t warp_hist[kWarps][kRadix];
  __shared__ uint32_t global_base[kRadix];   // first output slot of (digit, this tile)
  __shared__ uint32_t local_start[kRadix];   // first slot of the digit inside the tile
  __shared__ uint32_t scan_tmp[8];
  __shared__ uint32_t s_keys[kSortTile];
  __shared__ uint32_t s_vals[kSortTile];
  __shared__ uint32_t tile_s;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

  if (threadIdx.x == 0) tile_s = atomicAdd(ticket, 1u);
#pragma unroll
  for (int w = 0; w < kWarps; ++w) warp_hist[w][threadIdx.x] = 0;
  // exclusive scan of the global digit histogram = first slot of every digit in the output
  const uint32_t digit_base = block_excl_scan_256(hist[threadIdx.x], scan_tmp, lane, warp);
  const uint32_t tile = tile_s;

  const long long tile0 = (long long)tile * kSortTile;
  const long long wbase = tile0 + (long long)warp * (32 * kSortItems);
  uint32_t key[kSortItems], val[kSortItems], rank[kSortItems];
#pragma unroll
  for (int r = 0; r < kSortItems; ++r) {
    const long long idx = wbase + r * 32 + lane;
    key[r] = (idx < n) ? keys_in[idx] : 0u;
    val[r] = (idx < n) ? (vals_in ? vals_in[idx] : (uint32_t)idx) : 0u;
  }
#pragma unroll
  for (int r = 0; r < kSortItems; ++r) {
    const long long idx = wbase + r * 32 + lane;
    const bool valid = idx < n;
    const uint32_t d = valid ? ((key[r] >> shift) & mask) : 0u;
    const unsigned peers = match_digit(d, valid);
    const uint32_t before = valid ? warp_hist[warp][d] : 0;
    __syncwarp();
    if (valid && (peers & lanemask_lt()) == 0) warp_hist[warp][d] = before + __popc(peers);
    __syncwarp();
    rank[r] = before + __popc(peers & lanemask_lt());
  }
  __syncthreads();
  {
    // digit = threadIdx.x
    uint32_t cnt = 0;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) cnt += warp_hist[w][threadIdx.x];
    uint32_t* my_desc = desc + (size_t)tile * kRadix + threadIdx.x;
    // publish the tile aggregate, then look back over the preceding tiles
    asm volatile("st.volatile.global.u32 [%0], %1;" ::"l"(my_desc), "r"(kFlagAgg | cnt) : "memory");
    // look back in windows of kLook tiles: the kLook loads of one window are independent and
    // in flight together, so the serial latency chain is ~1/kLook of a tile-by-tile walk
    constexpr int kLook = 8;
    uint32_t excl = 0;
    long long t = (long long)tile - 1;
    bool done = t < 0;
    while (!done) {
      uint32_t v[kLook];
#pragma unroll
      for (int j = 0; j < kLook; ++j)
        v[j] = (t - j >= 0) ? ld_volatile_u32(desc + (size_t)(t - j) * kRadix + threadIdx.x) : kFlagIncl;
#pragma unroll
      for (int j = 0; j < kLook; ++j) {
        if (done) break;
        while ((v[j] & (kFlagAgg | kFlagIncl)) == 0)   // predecessor not published yet: poll it
          v[j] = ld_volatile_u32(desc + (size_t)(t - j) * kRadix + threadIdx.x);
        excl += v[j] & kValueMask;
        if (v[j] & kFlagIncl) done = true;
      }
      t -= kLook;
    }
    asm volatile("st.volatile.global.u32 [%0], %1;" ::"l"(my_desc), "r"(kFlagIncl | (excl + cnt))
                 : "memory");
    global_base[threadIdx.x] = digit_base + excl;
    const uint32_t start = block_excl_scan_256(cnt, scan_tmp, lane, warp);
    local_start[threadIdx.x] = start;
    uint32_t running = start;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) {
      const uint32_t t = warp_hist[w][threadIdx.x];
      warp_hist[w][threadIdx.x] = running;
      running += t;
    }
  }
  __syncthreads();
  // stage the tile in sorted order in shared memory ...
#pragma unroll
  for (int r = 0; r < kSortItems; ++r) {
    const long long idx = wbase + r * 32 + lane;
    if (idx < n) {
      const uint32_t d = (key[r] >> shift) & mask;
      const uint32_t pos = warp_hist[warp][d] + rank[r];
      s_keys[pos] = key[r];
      s_vals[pos] = val[r];
    }
  }
  __syncthreads();
  // ... so that consecutive threads write consecutive slots of each digit run (coalesced)
  const int count = (int)min((long long)kSortTile, (long long)n - tile0);
  for (int i = threadIdx.x; i < count; i += kSortBlock) {
    const uint32_t k = s_keys[i];
    const uint32_t d = (k >> shift) & mask;
    const uint32_t pos = global_base[d] + ((uint32_t)i - local_start[d]);
    keys_out[pos] = k;
    vals_out[pos] = s_vals[i];
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

constexpr int kMaxPasses = 4;

size_t radix_sort_ws_bytes(long long n) {
  long long ntiles = (n + kSortTile - 1) / kSortTile;
  if (ntiles < 1) ntiles = 1;
  // per pass: tile descriptors; plus the global histograms and ticket counters
  return align_up((size_t)kMaxPasses * ntiles * kRadix * sizeof(uint32_t)) +
         align_up((size_t)kMaxPasses * (kRadix + 1) * sizeof(uint32_t)) + 512;
}

int radix_sort_pairs(uint32_t* keys[2], uint32_t* vals[2], bool vals_iota, int n, int num_bits,
                     void* ws, size_t ws_bytes, cudaStream_t stream, int* out_sel) {
  *out_sel = 0;
  if (n <= 0) return DBEV_OK;
  DBEV_CHECK_ARG(n < (1 << 30), "radix_sort_pairs: at most 2^30 - 1 keys (got %d)", n);
  const int ntiles = ceil_div(n, kSortTile);
  if (num_bits < 1) num_bits = 1;
  if (num_bits > 32) num_bits = 32;
  const int passes = (num_bits + kRadixBits - 1) / kRadixBits;
  Workspace w(ws, ws_bytes);
  uint32_t* desc = w.take<uint32_t>((size_t)passes * ntiles * kRadix);
  uint32_t* hist = w.take<uint32_t>((size_t)passes * kRadix + kMaxPasses);
  if (!w.ok()) {
    set_last_error("radix_sort_pairs: workspace too small (%zu < %zu)", ws_bytes, w.used);
    return DBEV_ERR_WORKSPACE;
  }
  uint32_t* tickets = hist + (size_t)passes * kRadix;
  // desc and hist/tickets are adjacent allocations of the bump allocator: clear both
  DBEV_CUDA(cudaMemsetAsync(desc, 0, (size_t)passes * ntiles * kRadix * sizeof(uint32_t), stream));
  DBEV_CUDA(cudaMemsetAsync(hist, 0, ((size_t)passes * kRadix + kMaxPasses) * sizeof(uint32_t), stream));
  const int hgrid = min(ceil_div(n, 256 * 4), kNumSMs * 8);
  radix_global_hist_kernel<<<hgrid, 256, 0, stream>>>(keys[0], n, passes, num_bits, hist);
  int sel = 0;
  for (int p = 0; p < passes; ++p) {
    const int shift = p * kRadixBits;
    const int bits = (num_bits - shift < kRadixBits) ? (num_bits - shift) : kRadixBits;
    const uint32_t mask = (1u << bits) - 1u;
    const uint32_t* vin = (p == 0 && vals_iota) ? nullptr : vals[sel];
    radix_onesweep_kernel<<<ntiles, kSortBlock, 0, stream>>>(
        keys[sel], vin, keys[sel ^ 1], vals[sel ^ 1], n, shift, mask, hist + (size_t)p * kRadix,
        desc + (size_t)p * ntiles * kRadix, tickets + p);
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
