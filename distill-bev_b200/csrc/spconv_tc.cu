// Sparse convolution on the 5th-gen tensor cores (tcgen05, 3xTF32) for B200 — the wide layers of
// the sparse LiDAR teacher (C_in, C_out in {32, 64, 128}); SURVEY.md §8 row E4.
//
// Reference behaviour: indiceConv (mmdet3d/ops/spconv/include/spconv/spconv_ops.h:261-361) —
// per kernel offset a gather into an [nHot, C_in] buffer, an fp32 cuBLAS mm and a scatter-add —
// followed in SparseEncoder by BatchNorm1d(eval) / residual / ReLU as separate kernels.
//
// Design: an implicit GEMM per tile of 128 output voxels, D[128 voxels, C_out] accumulated in
// TMEM over all (kernel offset k, 32-channel chunk) steps:
//   A = rows gathered through the output-major neighbour table nbr[k][o] (zeros where the
//       neighbour is absent). The 4 producer warps load them into registers and write them with
//       tcgen05.st straight into TENSOR MEMORY (one thread = one voxel row = one TMEM lane, one
//       column per channel): the MMA reads A from TMEM, so the gathered rows never touch shared
//       memory - with 3xTF32 the shared-memory pipe (STS + TMA writes + operand reads of 12 MMAs
//       per step) was the measured limiter of the first version (profiles/r01_spconv_tc.json);
//   B = W[k]^T chunk [C_out, 32] (K-major), fetched by TMA from the pre-transposed weights.
// Offsets with no neighbour in the whole tile are skipped. fp32 parity with the reference's
// cuBLAS-fp32 path is kept by the 3xTF32 split  a*w ~= a_hi*w_hi + a_lo*w_hi + a_hi*w_lo
// (hi = value rounded to TF32, lo = remainder): the tensor core sees only TF32 operands, the
// result is accurate to ~1e-6 relative. Warp roles: 0-3 gather producers (+ thread 0 issues the
// weight TMA), 4-7 epilogue (tcgen05.ld -> BN scale/shift, residual, ReLU -> row-contiguous
// stores; the TMEM lane is the voxel), warp 8 = MMA issuer + TMEM owner. Two TMEM accumulator
// buffers let the epilogue of tile t overlap the MMAs of tile t+1; persistent over tiles.
// Measured dead ends (B200, round 1): separate accumulators per 3xTF32 term (no gain: the step time
// is not an accumulate-dependency chain), 8 producer warps with half a chunk each (slower: more
// barrier traffic), prefetch distance 1 vs 2 (equal), cvt.rna.tf32 (conversion pipe: -15 %).
#include "spconv.cuh"

#include "umma.cuh"

#include <stdlib.h>

namespace dbev {

namespace {

constexpr int kTileM = 128;                   // output voxels per tile (UMMA M, TMEM lanes)
constexpr int kChunk = 32;                    // input channels per stage (one 128 B swizzle row)
constexpr int kProdWarps = 4;                  // one warp per TMEM lane quarter (8 warps measured slower)
constexpr int kMmaWarp = kProdWarps + 4;       // warp 12: MMA issuer + TMEM owner
constexpr int kTcThreads = (kMmaWarp + 1) * 32;  // 416

// value rounded to the nearest TF32 (10-bit mantissa, ties away from zero); the remainder v - head
// is exact in fp32. Integer add + mask (full-rate ALU) instead of cvt.rna.tf32.f32, which issues at
// the conversion-pipe rate (a quarter of the ALU rate) and made the producers ALU-bound.
__device__ __forceinline__ float tf32_head(float v) {
  return __uint_as_float((__float_as_uint(v) + 0x1000u) & 0xFFFFE000u);
}

struct TcShape {
  int c_in, c_out, n_out, n_tiles, chunks, kvol;
  int debug;  // experiments only (DBEV_TC_DEBUG): 1 = skip row loads, 2 = skip TMEM stores, 4 = skip W TMA
};

// NACC accumulators per tile: the three 3xTF32 terms go to different TMEM accumulators (summed by
// the epilogue) so that consecutive MMAs do not form one read-modify-write dependency chain on the
// same accumulator - with K = 8 per instruction and N <= 128 the chain latency, not the tensor
// throughput, set the step time of the single-accumulator version (same ~1550 clk per step for
// N = 32, 64 and 128). NBUF = accumulator sets (2 = epilogue overlaps the next tile).
template <int COUT, int KVOL, int STAGES, int NACC, int NBUF>
__global__ void __launch_bounds__(kTcThreads, 1)
sp_conv_tc_kernel(const __grid_constant__ CUtensorMap tmap_whi,
                  const __grid_constant__ CUtensorMap tmap_wlo, const float* __restrict__ in_feats,
                  const int* __restrict__ nbr, const float* __restrict__ scale,
                  const float* __restrict__ shift, const float* __restrict__ residual, int relu,
                  float* __restrict__ out, TcShape s) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  constexpr int kWBytes = COUT * kChunk * 4;
  constexpr int kStageBytes = 2 * kWBytes;            // W hi + lo; A lives in TMEM
  constexpr uint32_t kAccCols = NBUF * NACC * COUT;   // accumulator columns
  constexpr uint32_t kAStageCols = 64;                // A hi (32 columns) + A lo (32 columns)
  static_assert(STAGES * kAStageCols + kAccCols <= 512, "TMEM budget");
  uint8_t* base = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ __align__(8) uint64_t full_bar[STAGES], empty_bar[STAGES], tmem_full_bar[NBUF],
      tmem_empty_bar[NBUF];
  __shared__ uint32_t stage_flags[STAGES];  // bit 0: first step of a tile, bit 1: last step
  __shared__ uint32_t mask_s[2];
  __shared__ int idx_s[KVOL * kTileM];  // neighbour rows of the current tile, [k][row]
  __shared__ uint32_t tmem_base_s;

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;  // uniform for the compiler

  if (threadIdx.x == 0) {
#pragma unroll
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], kProdWarps + 1);  // one arrival per producer warp + the expect_tx arrival
      mbar_init(&empty_bar[i], 1);
    }
#pragma unroll
    for (int i = 0; i < NBUF; ++i) {
      mbar_init(&tmem_full_bar[i], 1);
      mbar_init(&tmem_empty_bar[i], 4);
    }
    mask_s[0] = mask_s[1] = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_whi) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_wlo) : "memory");
  }
  if (warp == kMmaWarp) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_addr(&tmem_base_s)),
                 "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_s;

  if (warp < kProdWarps) {
    // ------------------------------------------------------------ gather producers
    // Steps of a tile = (active offset k) x (32-channel chunk kc). The rows of step i+1 are
    // requested from L2 BEFORE the thread waits for / fills the stage of step i, so two steps of
    // loads are in flight per thread (the producers are latency-bound, not bandwidth-bound).
    const int t = threadIdx.x;  // row of the tile = TMEM lane
    uint32_t stage = 0, phase = 0;
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < s.n_tiles; tile += gridDim.x, ++it) {
      const int o = tile * kTileM + t;
      uint32_t bits = 0;
      asm volatile("bar.sync 1, 128;" ::: "memory");  // previous tile's idx_s / mask_s reads are done
#pragma unroll
      for (int k = 0; k < KVOL; ++k) {
        const int v = o < s.n_out ? __ldg(nbr + (long long)k * s.n_out + o) : -1;
        idx_s[k * kTileM + t] = v;
        bits |= (v >= 0 ? 1u : 0u) << k;
      }
      const uint32_t par = it & 1u;
      if (threadIdx.x == 0) mask_s[par ^ 1u] = 0;
      bits = __reduce_or_sync(0xffffffffu, bits);
      if (lane == 0 && bits) atomicOr(&mask_s[par], bits);
      asm volatile("bar.sync 1, 128;" ::: "memory");
      uint32_t mask = mask_s[par];
      if (mask == 0) mask = 1u;  // keep producer and MMA issuer in step on an (impossible) empty tile
      const int n_steps = __popc(mask) * s.chunks;

      // Rows of steps i+1 and i+2 are in flight while step i is converted and stored: the producers
      // are bound by L2 latency x loads in flight (128 threads x 8 x 16 B per step), so the prefetch
      // distance sets the gather bandwidth.
      float4 b0[8], b1[8], b2[8];  // step i is consumed from b[i % 3] while i+1, i+2 are in flight
      uint32_t ld_rest = mask;    // load iterator: runs two steps ahead of the consume iterator
      int ld_k = __ffs(ld_rest) - 1, ld_kc = 0, ld_step = 0;
      ld_rest &= ld_rest - 1;
      auto load_step = [&](float4 (&buf)[8]) {
        if (ld_step < n_steps) {
          int row = idx_s[ld_k * kTileM + t];
          if (s.debug & 1) row = -1;
          if (row >= 0) {
            const float4* src = reinterpret_cast<const float4*>(in_feats + (long long)row * s.c_in +
                                                                ld_kc * kChunk);
#pragma unroll
            for (int c = 0; c < 8; ++c) buf[c] = __ldg(src + c);
          } else {  // absent neighbour: a zero row
#pragma unroll
            for (int c = 0; c < 8; ++c) buf[c] = make_float4(0.f, 0.f, 0.f, 0.f);
          }
          ++ld_step;
          if (++ld_kc == s.chunks) {
            ld_kc = 0;
            ld_k = ld_rest ? __ffs(ld_rest) - 1 : 0;
            ld_rest &= ld_rest - 1;
          }
        }
      };
      uint32_t rest = mask;       // consume iterator
      int k_cur = __ffs(rest) - 1, kc_cur = 0;
      rest &= rest - 1;
      int step = 0;
      // consume `cur` (step `step`), refill `nxt2` with step + 2
      auto do_step = [&](float4 (&cur)[8], float4 (&nxt2)[8]) {
        load_step(nxt2);
        mbar_wait(&empty_bar[stage], phase ^ 1u);
        uint8_t* st = base + (size_t)stage * kStageBytes;
        if (threadIdx.x == 0) {
          stage_flags[stage] = (step == 0 ? 1u : 0u) | (step == n_steps - 1 ? 2u : 0u);
          if (s.debug & 4) {
            mbar_arrive(&full_bar[stage]);
          } else {
            mbar_expect_tx(&full_bar[stage], 2u * kWBytes);
            tma_load_2d(st, &tmap_whi, kc_cur * kChunk, k_cur * COUT, &full_bar[stage]);
            tma_load_2d(st + kWBytes, &tmap_wlo, kc_cur * kChunk, k_cur * COUT, &full_bar[stage]);
          }
        }
        uint32_t hi[32], lo[32];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const float v4[4] = {cur[c].x, cur[c].y, cur[c].z, cur[c].w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float h = tf32_head(v4[e]);
            hi[c * 4 + e] = __float_as_uint(h);
            lo[c * 4 + e] = __float_as_uint(v4[e] - h);
          }
        }
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");  // after the empty-barrier wait
        const uint32_t a_addr = tmem_base + ((uint32_t)(warp * 32) << 16) + kAccCols + stage * kAStageCols;
        if (!(s.debug & 2)) {
          tmem_st32(a_addr, hi);
          tmem_st32(a_addr + 32, lo);
          asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(&full_bar[stage]);  // one arrival per producer warp
        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        if (++kc_cur == s.chunks) {
          kc_cur = 0;
          k_cur = rest ? __ffs(rest) - 1 : 0;
          rest &= rest - 1;
        }
        ++step;
      };
      load_step(b0);
      load_step(b1);
      while (step < n_steps) {   // statically rotated buffers: no register copies per step
        do_step(b0, b2);
        if (step < n_steps) do_step(b1, b0);
        if (step < n_steps) do_step(b2, b1);
      }
    }
  } else if (warp == kMmaWarp) {
    // ------------------------------------------------------------ MMA issuer
    // whole warp in the loop (warp-uniform control flow -> descriptors in uniform registers, MMAs
    // issued back to back), one elected lane issues and commits
    const bool leader = elect_one();
    const uint32_t idesc = umma_idesc_tf32(kTileM, COUT);
    const uint64_t desc0 = umma_desc(0, 16, 1024);
    uint32_t stage = 0, phase = 0;
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < s.n_tiles; tile += gridDim.x, ++it) {
      const uint32_t buf = it % NBUF, use = it / NBUF;
      mbar_wait(&tmem_empty_bar[buf], (use & 1u) ^ 1u);  // epilogue has drained this buffer
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t d0 = tmem_base + buf * (NACC * COUT);
      const uint32_t d1 = d0 + (NACC > 1 ? COUT : 0);
      const uint32_t d2 = d0 + (NACC > 2 ? 2 * COUT : (NACC > 1 ? COUT : 0));
      uint32_t first = 1;                      // no accumulator of this tile has been written yet
      while (true) {
        mbar_wait(&full_bar[stage], phase);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t flags = __shfl_sync(0xffffffffu, stage_flags[stage], 0);
        const uint32_t w_hi = smem_addr(base + (size_t)stage * kStageBytes);
        const uint64_t d_whi = desc0 + (uint64_t)(w_hi >> 4), d_wlo = desc0 + (uint64_t)((w_hi + kWBytes) >> 4);
        const uint32_t a_hi = tmem_base + kAccCols + stage * kAStageCols;  // lane 0, A columns
        const uint32_t a_lo = a_hi + 32;
        if (leader && !(s.debug & 16)) {
#pragma unroll
          for (int kk = 0; kk < kChunk / 8; ++kk) {
            // accumulate flags: with NACC < 3 several products share an accumulator, only the first write clears it
            const uint32_t c0 = (kk == 0 && first) ? 0u : 1u;
            const uint32_t c1 = (NACC > 1) ? c0 : 1u;
            const uint32_t c2 = (NACC > 2) ? c0 : 1u;
            umma_tf32_ts(d0, a_hi + kk * 8, d_whi + (uint64_t)(kk * 2), idesc, c0);
            umma_tf32_ts(d1, a_lo + kk * 8, d_whi + (uint64_t)(kk * 2), idesc, c1);
            umma_tf32_ts(d2, a_hi + kk * 8, d_wlo + (uint64_t)(kk * 2), idesc, c2);
          }
          umma_commit(&empty_bar[stage]);  // frees the stage when these MMAs retire
        }
        __syncwarp();
        first = 0;
        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        if (flags & 2u) break;
      }
      if (leader) umma_commit(&tmem_full_bar[buf]);
      __syncwarp();
    }
  } else {
    // ------------------------------------------------------------ epilogue (4 warps after the producers)
    const int q = warp & 3;  // TMEM lane quarter
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < s.n_tiles; tile += gridDim.x, ++it) {
      const uint32_t buf = it % NBUF, use = it / NBUF;
      const int o = tile * kTileM + q * 32 + lane;
      mbar_wait(&tmem_full_bar[buf], use & 1u);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
      for (int cc = 0; cc < COUT / 32; ++cc) {
        if (s.debug & 32) break;
        uint32_t v[32];
        const uint32_t tcol = tmem_base + ((uint32_t)(q * 32) << 16) + buf * (NACC * COUT) + (uint32_t)(cc * 32);
        tmem_ld32(tcol, v);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int j2 = 1; j2 < NACC; ++j2) {   // sum of the 3xTF32 term accumulators
          uint32_t u[32];
          tmem_ld32(tcol + j2 * COUT, u);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(u[j]));
        }
        if (o < s.n_out) {
          float* orow = out + (long long)o * COUT + cc * 32;
          const float* rrow = residual ? residual + (long long)o * COUT + cc * 32 : nullptr;
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            float4 r = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]),
                                   __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
            if (scale) {
              const float4 sc = __ldg(reinterpret_cast<const float4*>(scale + cc * 32 + j));
              r.x *= sc.x, r.y *= sc.y, r.z *= sc.z, r.w *= sc.w;
            }
            if (shift) {
              const float4 sh = __ldg(reinterpret_cast<const float4*>(shift + cc * 32 + j));
              r.x += sh.x, r.y += sh.y, r.z += sh.z, r.w += sh.w;
            }
            if (rrow) {
              const float4 rr = __ldg(reinterpret_cast<const float4*>(rrow + j));
              r.x += rr.x, r.y += rr.y, r.z += rr.z, r.w += rr.w;
            }
            if (relu) r.x = fmaxf(r.x, 0.f), r.y = fmaxf(r.y, 0.f), r.z = fmaxf(r.z, 0.f), r.w = fmaxf(r.w, 0.f);
            *reinterpret_cast<float4*>(orow + j) = r;
          }
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty_bar[buf]);
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == kMmaWarp) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u)
                 : "memory");
  }
}

int encode_w(CUtensorMap* map, const float* w_t, int kvol, int c_in, int c_out) {
  EncodeTiledFn encode = get_encode_fn();
  if (!encode) {
    set_last_error("spconv_forward_tc: cuTensorMapEncodeTiled not available from the driver");
    return DBEV_ERR_CUDA;
  }
  cuuint64_t dims[2] = {(cuuint64_t)c_in, (cuuint64_t)kvol * c_out};
  cuuint64_t strides[1] = {(cuuint64_t)c_in * 4};
  cuuint32_t box[2] = {(cuuint32_t)kChunk, (cuuint32_t)c_out};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)w_t, dims, strides, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                      CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_last_error("spconv_forward_tc: cuTensorMapEncodeTiled failed (%d)", (int)r);
    return DBEV_ERR_CUDA;
  }
  return DBEV_OK;
}

// wt_hi[k][co][ci] = W[k][ci][co] truncated to TF32, wt_lo = remainder
__global__ void sp_pack_weights_kernel(const float* __restrict__ w, int kvol, int c_in, int c_out,
                                       float* __restrict__ wt_hi, float* __restrict__ wt_lo) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)kvol * c_in * c_out;
  if (t >= total) return;
  const int ci = (int)(t % c_in);
  const int co = (int)((t / c_in) % c_out);
  const int k = (int)(t / ((long long)c_in * c_out));
  const float v = w[((long long)k * c_in + ci) * c_out + co];
  const float hi = tf32_head(v);
  wt_hi[t] = hi;
  wt_lo[t] = v - hi;
}

}  // namespace

bool spconv_tc_supported(int c_in, int c_out, int kvol) {
  return (c_in == 32 || c_in == 64 || c_in == 128) && (c_out == 32 || c_out == 64 || c_out == 128) &&
         (kvol == 27 || kvol == 3);
}

int spconv_pack_weights(const float* weight, int kvol, int c_in, int c_out, float* wt_hi,
                        float* wt_lo, cudaStream_t stream) {
  DBEV_CHECK_ARG(kvol >= 1 && c_in >= 1 && c_out >= 1, "spconv_pack_weights: bad sizes");
  const long long total = (long long)kvol * c_in * c_out;
  sp_pack_weights_kernel<<<ceil_div(total, 256), 256, 0, stream>>>(weight, kvol, c_in, c_out, wt_hi,
                                                                   wt_lo);
  DBEV_CHECK_LAUNCH("sp_pack_weights_kernel");
  return DBEV_OK;
}

int spconv_forward_tc(const float* in_feats, int c_in, const float* wt_hi, const float* wt_lo,
                      int c_out, const int* nbr, int kvol, int n_out, const float* scale,
                      const float* shift, const float* residual, int relu, float* out,
                      cudaStream_t stream) {
  DBEV_CHECK_ARG(spconv_tc_supported(c_in, c_out, kvol),
                 "spconv_forward_tc: needs C_in, C_out in {32, 64, 128} and a 27- or 3-tap kernel "
                 "(got %d -> %d, %d taps)", c_in, c_out, kvol);
  DBEV_CHECK_ARG(n_out >= 0, "spconv_forward_tc: negative n_out");
  DBEV_CHECK_ARG(((uintptr_t)in_feats & 15) == 0 && ((uintptr_t)wt_hi & 15) == 0 &&
                     ((uintptr_t)wt_lo & 15) == 0 && ((uintptr_t)out & 15) == 0 &&
                     ((uintptr_t)residual & 15) == 0 && ((uintptr_t)scale & 15) == 0 &&
                     ((uintptr_t)shift & 15) == 0,
                 "spconv_forward_tc: pointers must be 16-byte aligned");
  if (n_out == 0) return DBEV_OK;
  CUtensorMap map_hi, map_lo;
  int rc = encode_w(&map_hi, wt_hi, kvol, c_in, c_out);
  if (rc != DBEV_OK) return rc;
  rc = encode_w(&map_lo, wt_lo, kvol, c_in, c_out);
  if (rc != DBEV_OK) return rc;
  TcShape s;
  s.c_in = c_in, s.c_out = c_out, s.n_out = n_out, s.kvol = kvol;
  s.n_tiles = ceil_div(n_out, kTileM);
  s.chunks = c_in / kChunk;
  {
    const char* e = getenv("DBEV_TC_DEBUG");
    s.debug = e ? atoi(e) : 0;
  }
  int dev = 0, sms = 0;
  DBEV_CUDA(cudaGetDevice(&dev));
  DBEV_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int grid = s.n_tiles < sms ? s.n_tiles : sms;
#define DBEV_TC_LAUNCH(CO, KV, STG, NA, NB)                                                     \
  do {                                                                                          \
    const size_t smem = (size_t)STG * (2 * CO * kChunk * 4) + 1024;                             \
    DBEV_CUDA(cudaFuncSetAttribute(sp_conv_tc_kernel<CO, KV, STG, NA, NB>,                      \
                                   cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));    \
    sp_conv_tc_kernel<CO, KV, STG, NA, NB><<<grid, kTcThreads, smem, stream>>>(                 \
        map_hi, map_lo, in_feats, nbr, scale, shift, residual, relu, out, s);                   \
  } while (0)
  if (kvol == 27) {
    // TMEM budget (512 columns): NBUF * NACC * C_out accumulator columns + STAGES * 64 A columns
    // (splitting the three 3xTF32 terms over separate accumulators, NACC = 3, was measured: no gain)
    if (c_out == 32) DBEV_TC_LAUNCH(32, 27, 4, 1, 2);
    else if (c_out == 64) DBEV_TC_LAUNCH(64, 27, 4, 1, 2);
    else DBEV_TC_LAUNCH(128, 27, 4, 1, 2);
  } else {
    if (c_out == 32) DBEV_TC_LAUNCH(32, 3, 4, 1, 2);
    else if (c_out == 64) DBEV_TC_LAUNCH(64, 3, 4, 1, 2);
    else DBEV_TC_LAUNCH(128, 3, 4, 1, 2);
  }
#undef DBEV_TC_LAUNCH
  DBEV_CHECK_LAUNCH("sp_conv_tc_kernel");
  return DBEV_OK;
}

}  // namespace dbev
