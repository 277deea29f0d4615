// Weight gradient of a dense 2D convolution on the 5th-gen tensor cores (tcgen05, TF32) for the TRAINED
// student BEV encoder and the adaptation layers:
//   ResNetForBEVDet / BasicBlock   mmdet3d/models/backbones/resnet.py:51-62, bricks/res_block.py:70-99
//   FPN_LSS                        mmdet3d/models/necks/lss_fpn.py:62-72
//   adaptation 1x1 convs           mmdet3d/models/detectors/bevdet_distill.py:261-345
// (the reference gets it from aten::convolution_backward -> cuDNN, TF32 under torch's defaults).
//
//   dW[co][ci][ky][kx] = sum over (n, y, x) of  dy[n, y, x, co] * x[n, y*s + ky - p, x*s + kx - p, ci]
//
// is a GEMM whose contraction index is the PIXEL: with NHWC activations both operands have their M / N index
// (a channel) contiguous in memory, i.e. they are MN-major. tcgen05 reads MN-major TF32 operands only in
// the SWIZZLE_128B_BASE32B shared-memory layout (32-byte swizzle atoms; cute/atom/mma_traits_sm100.hpp
// Layout_MN_SW128_32B_Atom), which is what TMA writes with CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B: one
// 128-byte row per pixel (32 channels), so the same 4-D NHWC boxes the forward kernels load feed it.
//
// A CTA owns a 128 (C_out) x 128 (C_in) block of dW for one kx and all ky (three 128-column fp32
// accumulators in tensor memory) and a slice of the pixels (split-K). Per stage it loads an
//   8-pixel-wide, R-row strip of dy  (4 boxes of 32 channels)          A operand, K = 8 pixels of one row
//   the same strip of x shifted by kx - p, with one halo row above / below  (4 boxes)      B operand
// and issues, for every row r and ky, one M=128 N=128 K=8 MMA whose B descriptor starts at halo row r + ky:
// both operands are read from shared memory three times per fill (the L2 -> SMEM fill rate, not the
// tensor pipe, bounds these kernels otherwise: DESIGN.md §2.10). Stride-2 layers load one box per ky with
// TMA element strides. Partial sums [split][kx][ky][C_out][C_in] are combined in fixed order by
// wgrad_reduce_kernel, which also writes torch's [C_out][C_in][KH][KW] layout (deterministic: no atomics).
// Warp roles: warp 0 TMA producer, warp 1 MMA issuer + TMEM owner, warps 2-5 epilogue.
#include "conv_wgrad_tc.cuh"

#include "umma.cuh"

#include <cstdlib>

namespace dbev {

namespace {

constexpr int kWgThreads = 192;
constexpr int kTileC = 128;        // C_out and C_in block of a CTA (UMMA M and N)
constexpr int kStripW = 8;         // pixels per image row in a strip = K of one tf32 MMA
constexpr int kRowBytes = kStripW * 128;   // one strip row of one 32-channel chunk: 8 px x 128 B
constexpr int kWgStages = 3;

struct WgradShape {
  int n_ky, n_kxg;               // taps accumulated per CTA / kx groups (= CTAs per dW block and split)
  int stride, pad;
  int rows;                      // dy rows per stage
  int x_rows, x_blocks;          // rows per x box, x boxes per 32-channel chunk (1: halo windows, 3: one per ky)
  int ky_step;                   // bytes between the ky windows of the x stage
  int units_x, units_y, n_units; // strips: units_x * units_y per image
  int splits, co_tiles, ci_tiles;
  int c_in, c_out;
  long long part_stride;         // floats between [split][kxg][ky] slabs = c_out * c_in
};

__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, int c3,
                                            uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(smem_addr(dst)),
      "l"(map), "r"(smem_addr(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

// MN-major operand, SWIZZLE_128B_BASE32B: in 16-byte units ((8, n), (4, k)) : ((1, LBO), (8, SBO)) - 32 channels
// per 128-byte row, n 32-channel blocks LBO apart, K (pixels) in groups of four rows SBO apart
__device__ __forceinline__ uint64_t umma_desc_mn32(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= 1ull << 46;                                   // descriptor version 1 (Blackwell)
  d |= 1ull << 61;                                   // layout type SWIZZLE_128B_BASE32B
  return d;
}

__device__ __forceinline__ uint32_t umma_idesc_tf32_mn(int m, int n) {
  uint32_t d = 0;
  d |= 1u << 4;                    // c_format  F32
  d |= 2u << 7;                    // a_format  TF32
  d |= 2u << 10;                   // b_format  TF32
  d |= 1u << 15;                   // a_major   MN
  d |= 1u << 16;                   // b_major   MN
  d |= (uint32_t)(n >> 3) << 17;
  d |= (uint32_t)(m >> 4) << 24;
  return d;
}

__global__ void __launch_bounds__(kWgThreads, 1)
conv_wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmap_dy, const __grid_constant__ CUtensorMap tmap_x,
                     float* __restrict__ partial, WgradShape s) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* base = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ __align__(8) uint64_t full_bar[kWgStages], empty_bar[kWgStages], tmem_full_bar;
  __shared__ uint32_t tmem_base_s;
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
  constexpr uint32_t kTmemCols = 512;

  if (threadIdx.x == 0) {
#pragma unroll
    for (int i = 0; i < kWgStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    mbar_init(&tmem_full_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_dy) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_x) : "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_addr(&tmem_base_s)),
                 "r"(kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_s;

  // work item of this CTA: (C_out block, C_in block, kx group, pixel split)
  int idx = blockIdx.x;
  const int split = idx % s.splits;
  idx /= s.splits;
  const int kxg = idx % s.n_kxg;
  idx /= s.n_kxg;
  const int ci_t = idx % s.ci_tiles, co_t = idx / s.ci_tiles;
  const int u0 = (int)((long long)split * s.n_units / s.splits), u1 = (int)((long long)(split + 1) * s.n_units / s.splits);
  const int dy_chunk = s.rows * kRowBytes;                       // bytes of one 32-channel dy box
  const int x_chunk = s.x_blocks * s.x_rows * kRowBytes;         // bytes of the x boxes of one 32-channel chunk
  const int stage_bytes = 4 * (dy_chunk + x_chunk);
  const int per_img = s.units_x * s.units_y;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int u = u0; u < u1; ++u) {
        const int n = u / per_img, r = u - n * per_img;
        const int y0 = (r / s.units_x) * s.rows, x0 = (r % s.units_x) * kStripW;
        mbar_wait(&empty_bar[stage], phase ^ 1u);
        uint8_t* st = base + (size_t)stage * stage_bytes;
        mbar_expect_tx(&full_bar[stage], (uint32_t)stage_bytes);
#pragma unroll
        for (int c = 0; c < 4; ++c)
          tma_load_4d(st + c * dy_chunk, &tmap_dy, co_t * kTileC + c * 32, x0, y0, n, &full_bar[stage]);
        uint8_t* xs = st + 4 * dy_chunk;
        const int ix = x0 * s.stride + kxg - s.pad;
#pragma unroll
        for (int c = 0; c < 4; ++c)
          for (int b = 0; b < s.x_blocks; ++b)
            tma_load_4d(xs + c * x_chunk + b * s.x_rows * kRowBytes, &tmap_x, ci_t * kTileC + c * 32, ix,
                        y0 * s.stride + b - s.pad, n, &full_bar[stage]);
        if (++stage == kWgStages) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (whole warp walks the loop)
    const bool leader = elect_one();
    const uint32_t idesc = umma_idesc_tf32_mn(kTileC, kTileC);
    const uint64_t desc_a = umma_desc_mn32(0, (uint32_t)dy_chunk, 512), desc_b = umma_desc_mn32(0, (uint32_t)x_chunk, 512);
    uint32_t stage = 0, phase = 0;
    for (int u = u0; u < u1; ++u) {
      mbar_wait(&full_bar[stage], phase);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t a0 = smem_addr(base + (size_t)stage * stage_bytes);
      const uint32_t b0 = a0 + 4u * (uint32_t)dy_chunk;
      const uint32_t first = u == u0 ? 0u : 1u;
      if (leader) {
#pragma unroll 1
        for (int r = 0; r < s.rows; ++r) {
          const uint64_t ad = desc_a + (uint64_t)((a0 + (uint32_t)(r * kRowBytes)) >> 4);
          for (int ky = 0; ky < s.n_ky; ++ky) {
            const uint64_t bd = desc_b + (uint64_t)((b0 + (uint32_t)(ky * s.ky_step + r * kRowBytes)) >> 4);
            umma_tf32(tmem_base + (uint32_t)(ky * kTileC), ad, bd, idesc, r != 0 ? 1u : first);
          }
        }
        umma_commit(&empty_bar[stage]);
      }
      __syncwarp();
      if (++stage == kWgStages) { stage = 0; phase ^= 1u; }
    }
    if (leader) umma_commit(&tmem_full_bar);
    __syncwarp();
  } else {
    // ------------------------------------------------------------ epilogue (warps 2..5): TMEM -> partial sums
    const int q = warp & 3;
    const int co = co_t * kTileC + q * 32 + lane;
    mbar_wait(&tmem_full_bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int ky = 0; ky < s.n_ky; ++ky) {
      float* dst = partial + ((long long)(split * s.n_kxg + kxg) * s.n_ky + ky) * s.part_stride +
                   (long long)co * s.c_in + ci_t * kTileC;
#pragma unroll 1
      for (int cc = 0; cc < kTileC / 32; ++cc) {
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(ky * kTileC + cc * 32), v);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (u1 > u0) {
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            *reinterpret_cast<float4*>(dst + cc * 32 + j) =
                make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
        } else {
#pragma unroll
          for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(dst + cc * 32 + j) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
  }
}

// dW[co][ci][ky][kx] (torch layout) = (accumulate ? dW : 0) + sum over splits of partial[split][kx][ky][co][ci];
// a thread owns one (tap, co, ci): coalesced slab reads, the splits are summed in split order (deterministic)
__global__ void wgrad_reduce_kernel(const float* __restrict__ partial, int splits, int n_kxg, int n_ky, int c_out, int c_in,
                                    float* __restrict__ dw, int accumulate) {
  const long long total = (long long)c_out * c_in;
  const int taps = n_kxg * n_ky;
  const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= total * taps) return;
  const int slab = (int)(j / total);                // = kx * n_ky + ky
  const long long i = j - (long long)slab * total;  // (co, ci), ci fastest
  const int kx = slab / n_ky, ky = slab - kx * n_ky;
  const float* p = partial + (long long)slab * total + i;
  const long long step = (long long)taps * total;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  int sp = 0;
  for (; sp + 3 < splits; sp += 4) {
    a0 += p[(long long)sp * step];
    a1 += p[(long long)(sp + 1) * step];
    a2 += p[(long long)(sp + 2) * step];
    a3 += p[(long long)(sp + 3) * step];
  }
  for (; sp < splits; ++sp) a0 += p[(long long)sp * step];
  const float acc = (a0 + a1) + (a2 + a3);
  float* o = dw + i * taps + ky * n_kxg + kx;
  *o = accumulate ? *o + acc : acc;
}

int pick_splits(int groups, int n_units, int sms, long long dw_bytes) {
  // time model: waves * ceil(units / s) strips of MMAs + s slabs of partial sums written and re-read
  int best = 1;
  double best_t = 1e30;
  for (int sp = 1; sp <= n_units && sp <= 32; ++sp) {
    const int waves = ceil_div((long long)groups * sp, sms);
    const double t = (double)waves * ceil_div(n_units, sp) * 1.0 /* ~1 us per strip unit */ +
                     (double)sp * dw_bytes * 2.0 / 4.0e6 /* us at ~4 TB/s */ + 3.0 * waves;
    if (t < best_t) best_t = t, best = sp;
  }
  return best;
}

}  // namespace

size_t conv_wgrad_tc_workspace_bytes(int n_img, int ho, int wo, int c_in, int c_out, int kh, int kw, int stride) {
  if (c_in % kTileC != 0 || c_out % kTileC != 0) return 0;
  const int rows = stride == 1 ? 8 : 4;
  const int n_units = n_img * ceil_div(wo, kStripW) * ceil_div(ho, rows);
  const int groups = (c_out / kTileC) * (c_in / kTileC) * kw;
  const long long dw_bytes = (long long)c_out * c_in * kh * kw * 4;
  const int sp = pick_splits(groups, n_units, kNumSMs, dw_bytes);
  return (size_t)sp * dw_bytes + 1024;
}

int conv_wgrad_tc(const float* x_nhwc, int n_img, int h, int w, int c_in, int x_ld, const float* dy_nhwc, int ho,
                  int wo, int c_out, int dy_ld, int kh, int kw, int stride, int pad, float* dw, int accumulate,
                  void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  DBEV_CHECK_ARG(n_img > 0 && h > 0 && w > 0 && ho > 0 && wo > 0, "conv_wgrad_tc: empty input");
  DBEV_CHECK_ARG(c_in % kTileC == 0 && c_out % kTileC == 0,
                 "conv_wgrad_tc: C_in and C_out must be multiples of 128 (got %d, %d)", c_in, c_out);
  DBEV_CHECK_ARG((kh == 3 && kw == 3 && pad == 1 && (stride == 1 || stride == 2)) ||
                     (kh == 1 && kw == 1 && pad == 0 && stride == 1),
                 "conv_wgrad_tc: 3x3 / pad 1 / stride 1 or 2, or 1x1 / stride 1");
  DBEV_CHECK_ARG(ho == (h + 2 * pad - kh) / stride + 1 && wo == (w + 2 * pad - kw) / stride + 1,
                 "conv_wgrad_tc: dy size does not match the convolution");
  DBEV_CHECK_ARG(x_ld >= c_in && dy_ld >= c_out && x_ld % 4 == 0 && dy_ld % 4 == 0, "conv_wgrad_tc: bad leading dimensions");
  DBEV_CHECK_ARG(((uintptr_t)x_nhwc & 15) == 0 && ((uintptr_t)dy_nhwc & 15) == 0 && ((uintptr_t)dw & 15) == 0 &&
                     ((uintptr_t)workspace & 255) == 0,
                 "conv_wgrad_tc: pointers must be 16-byte aligned (workspace 256)");
  EncodeTiledFn encode = get_encode_fn();
  if (!encode) {
    set_last_error("conv_wgrad_tc: cuTensorMapEncodeTiled not available from the driver");
    return DBEV_ERR_CUDA;
  }
  WgradShape s;
  s.n_ky = kh, s.n_kxg = kw, s.stride = stride, s.pad = pad;
  s.rows = stride == 1 ? 8 : 4;
  if (stride == 1) {
    s.x_blocks = 1, s.x_rows = s.rows + (kh - 1), s.ky_step = kRowBytes;
  } else {
    s.x_blocks = kh, s.x_rows = s.rows, s.ky_step = s.x_rows * kRowBytes;
  }
  s.units_x = ceil_div(wo, kStripW), s.units_y = ceil_div(ho, s.rows);
  s.n_units = n_img * s.units_x * s.units_y;
  s.co_tiles = c_out / kTileC, s.ci_tiles = c_in / kTileC;
  s.c_in = c_in, s.c_out = c_out;
  s.part_stride = (long long)c_out * c_in;
  const long long dw_bytes = (long long)c_out * c_in * kh * kw * 4;
  const int groups = s.co_tiles * s.ci_tiles * s.n_kxg;
  s.splits = pick_splits(groups, s.n_units, kNumSMs, dw_bytes);
  DBEV_CHECK_ARG(workspace != nullptr && workspace_bytes >= (size_t)s.splits * dw_bytes,
                 "conv_wgrad_tc: workspace too small (%zu < %lld)", workspace_bytes, (long long)s.splits * dw_bytes);

  CUtensorMap tmap_dy, tmap_x;
  {
    cuuint64_t dims[4] = {(cuuint64_t)c_out, (cuuint64_t)wo, (cuuint64_t)ho, (cuuint64_t)n_img};
    cuuint64_t strides[3] = {(cuuint64_t)dy_ld * 4, (cuuint64_t)wo * dy_ld * 4, (cuuint64_t)ho * wo * dy_ld * 4};
    cuuint32_t box[4] = {32, (cuuint32_t)kStripW, (cuuint32_t)s.rows, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = encode(&tmap_dy, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)dy_nhwc, dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_last_error("conv_wgrad_tc: cuTensorMapEncodeTiled(dy) failed (%d)", (int)r);
      return DBEV_ERR_CUDA;
    }
  }
  {
    cuuint64_t dims[4] = {(cuuint64_t)c_in, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)n_img};
    cuuint64_t strides[3] = {(cuuint64_t)x_ld * 4, (cuuint64_t)w * x_ld * 4, (cuuint64_t)h * w * x_ld * 4};
    cuuint32_t box[4] = {32, (cuuint32_t)(kStripW * stride), (cuuint32_t)(s.x_rows * stride), 1};
    cuuint32_t estr[4] = {1, (cuuint32_t)stride, (cuuint32_t)stride, 1};
    CUresult r = encode(&tmap_x, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)x_nhwc, dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_last_error("conv_wgrad_tc: cuTensorMapEncodeTiled(x) failed (%d)", (int)r);
      return DBEV_ERR_CUDA;
    }
  }
  const int stage_bytes = 4 * (s.rows + s.x_blocks * s.x_rows) * kRowBytes;
  const size_t smem = (size_t)kWgStages * stage_bytes + 1024;
  DBEV_CHECK_ARG(smem <= 227 * 1024, "conv_wgrad_tc: stages do not fit shared memory");
  const int grid = groups * s.splits;
  DBEV_CUDA(cudaFuncSetAttribute(conv_wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  conv_wgrad_tc_kernel<<<grid, kWgThreads, smem, stream>>>(tmap_dy, tmap_x, (float*)workspace, s);
  DBEV_CHECK_LAUNCH("conv_wgrad_tc_kernel");
  const long long total = (long long)c_out * c_in;
  wgrad_reduce_kernel<<<ceil_div(total * kh * kw, 256), 256, 0, stream>>>((const float*)workspace, s.splits, s.n_kxg, s.n_ky, c_out,
                                                               c_in, dw, accumulate);
  DBEV_CHECK_LAUNCH("wgrad_reduce_kernel");
  return DBEV_OK;
}

}  // namespace dbev
