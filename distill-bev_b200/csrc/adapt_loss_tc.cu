// The student's '1x1conv' adaptation layer fused with the FGD distillation loss, on the 5th-gen tensor cores (B200).
//
// Reference behaviour: bevdet_distill.py:1004 applies channel_wise_adaptations[index] (nn.Conv2d(Cs, Ct, 1),
// built at :216-351) to the student BEV feature, then :1084-1293 reads the adapted map several times (attention
// sums, masked squared differences, and again in the backward pass). Here the adapted map s = x W^T + bias exists
// only as 128-cell x C_out tiles in tensor memory:
//
//   mode 0 (forward)   per cell: sum_c |s|, sum_c s, sum_c (s-t)^2, sum_c catt_c (s-t)^2; per channel and tile: sum s
//   mode 1 (backward)  the tile is recomputed and  ds = (s - t)(Wa + catt_c Wb) + gc_c + gsp  leaves as channels-last
//                      rows [B, HW, C] (what the tcgen05 input- / weight-gradient kernels read); per channel and tile
//                      sum ds (the conv bias gradient)
//
// so the 201 MB adapted map (B = 8, 384 x 128 x 128) is never written or re-read and the NCHW -> channels-last
// transpose of ds disappears. Same per-element formulas as fgd_student_pass_kernel / fgd_bwd_main_kernel
// (distill_loss.cu), which remain the path for every other adaptation layer.
//
// Layout of one CTA (persistent, one per SM, 7 warps): a work item is (128-cell tile, column part) with
// M = 128 cells (TMEM lane = cell), N = C_out or C_out / 2 columns (<= 256), K = C_in in 32-channel chunks.
//   warp 0   TMA producer of the GEMM operands: x tile {32 k, 128 cells} + W rows {32 k, N} per chunk, 3-stage ring
//   warp 1   tcgen05.mma kind::tf32 issuer; accumulators double-buffered in TMEM (2 x 256 columns)
//   warps 2-5 epilogue: a lane owns one cell; per 32-channel group tcgen05.ld 32 columns, teacher values of the same
//            cells x channels from a TMA-staged box (NCHW teacher: {128 cells, 32 channels}), the loss arithmetic in
//            registers, per-channel sums over the warp's 32 cells by a transposed shuffle butterfly (31 shuffles),
//            mode 1: ds rows staged in shared memory (128-byte swizzle) and written by TMA
//   warp 6   TMA producer of the teacher boxes (own ring: the epilogue frees a slot as soon as it holds the values)
// Everything is HBM-bound by design: x (134 MB) + teacher (201 MB) [+ ds (201 MB)] per call.
#include "adapt_loss_tc.cuh"

#include <cuda.h>

#include "umma.cuh"

namespace dbev {

namespace {

constexpr int kPix = 128;          // cells per tile (UMMA M)
constexpr int kKc = 32;            // input channels per stage (one 128-byte swizzle row of fp32)
constexpr int kThreads = 224;      // 7 warps
constexpr int kATile = kPix * kKc * 4;      // 16 KB
constexpr int kTTile = 32 * kPix * 4;       // teacher box: 32 channels x 128 cells = 16 KB
constexpr int kStages = 3;
constexpr int kMaxTSlots = 3;
constexpr int kBufCols = 256;      // TMEM columns per accumulator buffer

struct FusedShape {
  int batch, c_in, c_out, hw, tiles_per_img, n_tiles, k_chunks;
  int n_part, parts;               // columns per item, items per tile
  int stage_bytes, t_slots;
};

// lane L <- sum over the 32 lanes of v[L] (fixed tree, deterministic); v is destroyed
__device__ __forceinline__ float column_sums32(float (&v)[32], int lane) {
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) {
    const bool up = (lane & s) != 0;
#pragma unroll
    for (int i = 0; i < s; ++i) {
      const float send = up ? v[i] : v[i + s], keep = up ? v[i + s] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, s);
    }
  }
  return v[0];
}

__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

template <int MODE>
__global__ void __launch_bounds__(kThreads, 1)
adapt_fgd_tc_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w,
                    const __grid_constant__ CUtensorMap tmap_t, const __grid_constant__ CUtensorMap tmap_ds,
                    FusedShape s, AdaptFgdArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* base = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* t_ring = base + (size_t)kStages * s.stage_bytes;              // [t_slots][32 ch][128 cells]
  uint8_t* stg_base = t_ring + (size_t)s.t_slots * kTTile;               // [4 warps][2][32 cells x 128 B] (mode 1)
  float* colsum = reinterpret_cast<float*>(stg_base + (MODE == 1 ? 4 * 2 * 4096 : 0));   // [4][256]
  __shared__ __align__(8) uint64_t full_bar[kStages], empty_bar[kStages], t_full[kMaxTSlots], t_empty[kMaxTSlots],
      tmem_full_bar[2], tmem_empty_bar[2];
  __shared__ uint32_t tmem_base_s;

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int i = 0; i < kStages; ++i) mbar_init(&full_bar[i], 1), mbar_init(&empty_bar[i], 1);
    for (int i = 0; i < kMaxTSlots; ++i) mbar_init(&t_full[i], 1), mbar_init(&t_empty[i], 4);
    for (int i = 0; i < 2; ++i) mbar_init(&tmem_full_bar[i], 1), mbar_init(&tmem_empty_bar[i], 4);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_x) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_w) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_t) : "memory");
    if (MODE == 1) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_ds) : "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_addr(&tmem_base_s)),
                 "r"((uint32_t)(2 * kBufCols))
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_s;
  const int groups = s.n_part / 32;

  if (warp == 0) {
    // ------------------------------------------------------------ GEMM operand producer
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int tile = blockIdx.x; tile < s.n_tiles; tile += gridDim.x) {
        const int b = tile / s.tiles_per_img, hw0 = (tile % s.tiles_per_img) * kPix;
        for (int part = 0; part < s.parts; ++part) {
          for (int kc = 0; kc < s.k_chunks; ++kc) {
            mbar_wait(&empty_bar[stage], phase ^ 1u);
            uint8_t* st = base + (size_t)stage * s.stage_bytes;
            mbar_expect_tx(&full_bar[stage], (uint32_t)s.stage_bytes);
            tma_load_3d(st, &tmap_x, kc * kKc, hw0, b, &full_bar[stage]);
            tma_load_2d(st + kATile, &tmap_w, kc * kKc, part * s.n_part, &full_bar[stage]);
            if (++stage == kStages) { stage = 0; phase ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 6) {
    // ------------------------------------------------------------ teacher producer
    if (lane == 0) {
      uint32_t slot = 0, phase = 0;
      for (int tile = blockIdx.x; tile < s.n_tiles; tile += gridDim.x) {
        const int b = tile / s.tiles_per_img, hw0 = (tile % s.tiles_per_img) * kPix;
        for (int part = 0; part < s.parts; ++part) {
          for (int g = 0; g < groups; ++g) {
            mbar_wait(&t_empty[slot], phase ^ 1u);
            mbar_expect_tx(&t_full[slot], (uint32_t)kTTile);
            tma_load_3d(t_ring + (size_t)slot * kTTile, &tmap_t, hw0, part * s.n_part + g * 32, b, &t_full[slot]);
            if (++slot == (uint32_t)s.t_slots) { slot = 0; phase ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (warp-uniform loop, one elected lane issues)
    const bool leader = elect_one();
    const uint32_t idesc = umma_idesc_tf32(kPix, s.n_part);
    const uint64_t desc0 = umma_desc(0, 16, 1024);
    uint32_t stage = 0, phase = 0, it = 0;
    for (int tile = blockIdx.x; tile < s.n_tiles; tile += gridDim.x) {
      for (int part = 0; part < s.parts; ++part, ++it) {
        const uint32_t buf = it & 1u, use = it >> 1;
        mbar_wait(&tmem_empty_bar[buf], (use & 1u) ^ 1u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t d_tmem = tmem_base + buf * kBufCols;
#pragma unroll 1
        for (int kc = 0; kc < s.k_chunks; ++kc) {
          mbar_wait(&full_bar[stage], phase);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t a0 = smem_addr(base + (size_t)stage * s.stage_bytes);
          const uint64_t ad = desc0 + (uint64_t)(a0 >> 4), bd = desc0 + (uint64_t)((a0 + kATile) >> 4);
          if (leader) {
#pragma unroll
            for (int kk = 0; kk < kKc / 8; ++kk)
              umma_tf32(d_tmem, ad + (uint64_t)(kk * 2), bd + (uint64_t)(kk * 2), idesc, (kc | kk) != 0 ? 1u : 0u);
            umma_commit(&empty_bar[stage]);
          }
          __syncwarp();
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
        if (leader) umma_commit(&tmem_full_bar[buf]);
        __syncwarp();
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue (warps 2..5)
    const int q = warp & 3;                         // TMEM lane quarter; cell = hw0 + q * 32 + lane
    const int tid = q * 32 + lane;
    uint8_t* stg = stg_base + (size_t)q * 8192;
    const uint32_t stg_row = smem_addr(stg) + (uint32_t)lane * 128u;
    const uint32_t sw_xor = (uint32_t)(lane & 7);
    uint32_t it = 0, slot = 0, t_phase = 0, sbuf = 0;
    float kf = 0.f, kb = 0.f, kp = 0.f;
    if (MODE == 1) {
      const float ib = 1.f / (float)s.batch;
      kf = 2.f * a.w_fg * ib * __ldg(a.grad_losses + 0), kb = 2.f * a.w_bg * ib * __ldg(a.grad_losses + 1);
      kp = a.use_fp ? 2.f * a.w_fp * ib * __ldg(a.grad_losses + 2) : 0.f;
    }
    for (int tile = blockIdx.x; tile < s.n_tiles; tile += gridDim.x) {
      const int b = tile / s.tiles_per_img, tile_in_img = tile % s.tiles_per_img, hw0 = tile_in_img * kPix;
      const int p = hw0 + tid;
      const bool in = p < s.hw;
      const size_t o = (size_t)b * s.hw + (in ? p : 0);
      float wa = 0.f, wb = 0.f, gs = 0.f;
      float a_abs = 0.f, a_sum = 0.f, a_d1 = 0.f, a_d2 = 0.f;
      if (MODE == 1) {
        const float bs = kf * __ldg(a.fgw + o) + kb * __ldg(a.bgw + o), fpk = kp * __ldg(a.fpw + o);
        wa = a.channel_mask ? 0.f : bs, wb = a.channel_mask ? bs + fpk : fpk;
        gs = __ldg(a.gsp + o);
      }
      for (int part = 0; part < s.parts; ++part, ++it) {
        const uint32_t buf = it & 1u, use = it >> 1;
        mbar_wait(&tmem_full_bar[buf], use & 1u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
        for (int g = 0; g < groups; ++g) {
          const int c0 = part * s.n_part + g * 32;
          uint32_t v[32];
          tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + buf * kBufCols + (uint32_t)(g * 32), v);
          // teacher values of this lane's cell, 32 channels (conflict-free: lanes read consecutive cells)
          mbar_wait(&t_full[slot], t_phase);
          const float* tt = reinterpret_cast<const float*>(t_ring + (size_t)slot * kTTile) + tid;
          float t[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) t[j] = tt[j * kPix];
          __syncwarp();
          if (lane == 0) mbar_arrive(&t_empty[slot]);
          if (++slot == (uint32_t)s.t_slots) { slot = 0; t_phase ^= 1u; }
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          float r[32];
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 bi = a.bias ? __ldg(reinterpret_cast<const float4*>(a.bias + c0 + j)) : make_float4(0.f, 0.f, 0.f, 0.f);
            const float4 ca = __ldg(reinterpret_cast<const float4*>(a.catt + (size_t)b * s.c_out + c0 + j));
            const float bj[4] = {bi.x, bi.y, bi.z, bi.w}, cj[4] = {ca.x, ca.y, ca.z, ca.w};
            float gj[4] = {0.f, 0.f, 0.f, 0.f};
            if (MODE == 1) {
              const float4 gc = __ldg(reinterpret_cast<const float4*>(a.gc + (size_t)b * s.c_out + c0 + j));
              gj[0] = gc.x, gj[1] = gc.y, gj[2] = gc.z, gj[3] = gc.w;
            }
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float sv = __uint_as_float(v[j + e]) + bj[e];
              const float df = sv - t[j + e];
              if (MODE == 0) {
                const float q2 = df * df;
                a_abs += fabsf(sv), a_sum += sv, a_d1 += q2, a_d2 += cj[e] * q2;
                r[j + e] = in ? sv : 0.f;
              } else {
                r[j + e] = in ? df * (wa + cj[e] * wb) + gj[e] + gs : 0.f;
              }
            }
          }
          if (MODE == 1) {
            // the staging buffer written two groups ago must have been read by its TMA store
            if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            __syncwarp();
            const uint32_t dst = stg_row + sbuf * 4096u;
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(dst + ((((uint32_t)j >> 2) ^ sw_xor) << 4)),
                           "f"(r[j]), "f"(r[j + 1]), "f"(r[j + 2]), "f"(r[j + 3])
                           : "memory");
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) {
              if (hw0 + q * 32 < s.hw)
                asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(&tmap_ds),
                             "r"(smem_addr(stg) + sbuf * 4096u), "r"(c0), "r"(hw0 + q * 32), "r"(b)
                             : "memory");
              asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
            sbuf ^= 1u;
          }
          if (a.chan_p) colsum[q * 256 + g * 32 + lane] = column_sums32(r, lane);
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(&tmem_empty_bar[buf]);
        if (a.chan_p) {          // the four warps' partial column sums, added in cell order
          epi_bar();
          for (int c = tid; c < s.n_part; c += 128) {
            const float tot = ((colsum[c] + colsum[256 + c]) + colsum[512 + c]) + colsum[768 + c];
            a.chan_p[((size_t)b * s.c_out + part * s.n_part + c) * s.tiles_per_img + tile_in_img] = tot;
          }
          epi_bar();
        }
      }
      if (MODE == 0 && in) a.sa[o] = a_abs, a.sm[o] = a_sum, a.d1[o] = a_d1, a.d2[o] = a_d2;
    }
    if (MODE == 1) {
      if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
      __syncwarp();
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)(2 * kBufCols))
                 : "memory");
  }
}

bool split_columns(int c_out, int* n_part, int* parts) {
  if (c_out >= 32 && c_out <= 256 && c_out % 32 == 0) { *n_part = c_out, *parts = 1; return true; }
  if (c_out > 256 && c_out <= 512 && c_out % 64 == 0) { *n_part = c_out / 2, *parts = 2; return true; }
  return false;
}

int encode_map(EncodeTiledFn encode, CUtensorMap* map, int rank, const void* ptr, const cuuint64_t* dims, const cuuint64_t* strides,
               const cuuint32_t* box, CUtensorMapSwizzle swz, const char* what) {
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, const_cast<void*>(ptr), dims, strides, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_last_error("adapt_fgd_fused: cuTensorMapEncodeTiled(%s) failed (%d)", what, (int)r);
    return DBEV_ERR_CUDA;
  }
  return DBEV_OK;
}

}  // namespace

bool adapt_fgd_fused_supports(int c_in, int c_out, int hw) {
  int n_part, parts;
  return c_in >= kKc && c_in % kKc == 0 && hw > 0 && hw % 4 == 0 && split_columns(c_out, &n_part, &parts);
}

int adapt_fgd_fused(int mode, const float* x_cl, const float* w, int batch, int c_in, int c_out, int hw,
                    const AdaptFgdArgs& args, cudaStream_t stream) {
  DBEV_CHECK_ARG(mode == 0 || mode == 1, "adapt_fgd_fused: mode 0 (forward) or 1 (backward)");
  DBEV_CHECK_ARG(batch > 0 && adapt_fgd_fused_supports(c_in, c_out, hw),
                 "adapt_fgd_fused: C_in %% 32 == 0, C_out %% 32 == 0 up to 256 or %% 64 == 0 up to 512, H*W %% 4 == 0 (got %d -> %d, %d cells)",
                 c_in, c_out, hw);
  DBEV_CHECK_ARG(x_cl && w && args.teacher && args.catt, "adapt_fgd_fused: null input");
  DBEV_CHECK_ARG(mode == 1 ? (args.fgw && args.bgw && args.fpw && args.gsp && args.gc && args.grad_losses && args.ds_cl)
                           : (args.sa && args.sm && args.d1 && args.d2),
                 "adapt_fgd_fused: null loss-state array");
  DBEV_CHECK_ARG(((uintptr_t)x_cl % 16) == 0 && ((uintptr_t)w % 16) == 0 && ((uintptr_t)args.teacher % 16) == 0 &&
                     ((uintptr_t)args.bias % 16) == 0 && ((uintptr_t)args.catt % 16) == 0 && ((uintptr_t)args.gc % 16) == 0 &&
                     ((uintptr_t)args.ds_cl % 16) == 0,
                 "adapt_fgd_fused: tensors must be 16-byte aligned");
  EncodeTiledFn encode = get_encode_fn();
  if (!encode) {
    set_last_error("adapt_fgd_fused: cuTensorMapEncodeTiled not available from the driver");
    return DBEV_ERR_CUDA;
  }
  FusedShape s;
  s.batch = batch, s.c_in = c_in, s.c_out = c_out, s.hw = hw;
  s.tiles_per_img = ceil_div(hw, kPix), s.n_tiles = s.tiles_per_img * batch, s.k_chunks = c_in / kKc;
  split_columns(c_out, &s.n_part, &s.parts);
  s.stage_bytes = kATile + s.n_part * kKc * 4;
  s.t_slots = s.n_part <= 192 ? 3 : 2;
  CUtensorMap tmap_x, tmap_w, tmap_t, tmap_ds;
  {
    cuuint64_t dims[3] = {(cuuint64_t)c_in, (cuuint64_t)hw, (cuuint64_t)batch};
    cuuint64_t strides[2] = {(cuuint64_t)c_in * 4, (cuuint64_t)hw * c_in * 4};
    cuuint32_t box[3] = {kKc, kPix, 1};
    if (int rc = encode_map(encode, &tmap_x, 3, x_cl, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B, "x")) return rc;
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)c_in, (cuuint64_t)c_out};
    cuuint64_t strides[1] = {(cuuint64_t)c_in * 4};
    cuuint32_t box[2] = {kKc, (cuuint32_t)s.n_part};
    if (int rc = encode_map(encode, &tmap_w, 2, w, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B, "W")) return rc;
  }
  {
    cuuint64_t dims[3] = {(cuuint64_t)hw, (cuuint64_t)c_out, (cuuint64_t)batch};
    cuuint64_t strides[2] = {(cuuint64_t)hw * 4, (cuuint64_t)c_out * hw * 4};
    cuuint32_t box[3] = {kPix, 32, 1};
    if (int rc = encode_map(encode, &tmap_t, 3, args.teacher, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_NONE, "teacher")) return rc;
  }
  if (mode == 1) {
    cuuint64_t dims[3] = {(cuuint64_t)c_out, (cuuint64_t)hw, (cuuint64_t)batch};
    cuuint64_t strides[2] = {(cuuint64_t)c_out * 4, (cuuint64_t)hw * c_out * 4};
    cuuint32_t box[3] = {32, 32, 1};
    if (int rc = encode_map(encode, &tmap_ds, 3, args.ds_cl, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B, "ds")) return rc;
  } else {
    tmap_ds = tmap_x;
  }
  int dev = 0, sms = 0;
  DBEV_CUDA(cudaGetDevice(&dev));
  DBEV_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int grid = s.n_tiles < sms ? s.n_tiles : sms;
  const size_t smem = (size_t)kStages * s.stage_bytes + (size_t)s.t_slots * kTTile + (mode == 1 ? 4 * 2 * 4096 : 0) + 4 * 256 * 4 + 1024;
  if (mode == 0) {
    DBEV_CUDA(cudaFuncSetAttribute(adapt_fgd_tc_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    adapt_fgd_tc_kernel<0><<<grid, kThreads, smem, stream>>>(tmap_x, tmap_w, tmap_t, tmap_ds, s, args);
  } else {
    DBEV_CUDA(cudaFuncSetAttribute(adapt_fgd_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    adapt_fgd_tc_kernel<1><<<grid, kThreads, smem, stream>>>(tmap_x, tmap_w, tmap_t, tmap_ds, s, args);
  }
  DBEV_CHECK_LAUNCH("adapt_fgd_tc_kernel");
  return DBEV_OK;
}

}  // namespace dbev
