// 1x1 "adaptation" convolution on the 5th-gen tensor cores (tcgen05, TF32) for B200.
//
// Reference behaviour reproduced here: the '1x1conv' student adaptation layer
//   nn.Conv2d(student_channel, teacher_channel, kernel_size=1)
//   mmdet3d/models/detectors/bevdet_distill.py:216-351 (applied at :1004 under @force_fp32),
// i.e. Y[b, n, hw] = sum_k W[n, k] * X[b, k, hw] + bias[n] on NCHW fp32 tensors.
//
// Design. The conv is a GEMM per image with M = output channels, N = BEV cells, K = input
// channels, both operands K-major with the 128-byte swizzle the tensor core expects:
//   A = W [M, K] row-major                        TMA box {32 k, 128 m}
//   B = X channels-last [HW, K] (= torch.channels_last memory)  TMA box {32 k, 128 cells}
// (an NCHW-contiguous X is first transposed by transpose_batched_kernel; feeding NCHW directly
// as an MN-major TF32 operand needs the 32-bit-atom swizzle variant and produced zeros with the
// plain 128-byte swizzle on B200 - measured). One persistent CTA per SM loops over
// 128-cell tiles; warp 0 = TMA producer (3-stage ring of 32-channel K chunks), warp 1 = MMA issuer
// (tcgen05.mma kind::tf32, 128x128x8 per instruction, all M tiles of the output channels
// accumulate side by side in TMEM), warps 2-5 = epilogue (tcgen05.ld 32x32b -> + bias -> NCHW
// stores: a TMEM lane is an output channel, so every thread writes contiguous cells).
// fp32 data is consumed as TF32 (what cuDNN does for this conv under torch's default
// allow_tf32): relative error ~1e-3, inside the north_star loss tolerance.
#include "adapt_gemm.cuh"

#include <cuda.h>

#include "umma.cuh"

namespace dbev {

namespace {

constexpr int kBM = 128;      // output channels per M tile (UMMA M)
constexpr int kBN = 128;      // BEV cells per tile (UMMA N)
constexpr int kBK = 32;       // input channels per stage = one 128-byte swizzle row of fp32
constexpr int kUmmaK = 8;     // K per tcgen05.mma for 32-bit inputs
constexpr int kThreads = 192; // 6 warps
constexpr int kATileBytes = kBM * kBK * 4;   // 16 KB per M tile per stage
constexpr int kBTileBytes = kBN * kBK * 4;   // 16 KB per stage

struct GemmShape {
  int batch, c_in, c_out, hw;
  int tiles_per_img, n_tiles, k_chunks;
};

// MT = number of 128-channel M tiles (c_out = MT * 128, TMEM columns = MT * 128)
template <int MT, int STAGES>
__global__ void __launch_bounds__(kThreads, 1)
adapt_gemm_tf32_kernel(const __grid_constant__ CUtensorMap tmap_w,
                       const __grid_constant__ CUtensorMap tmap_x, const float* __restrict__ bias,
                       float* __restrict__ y, GemmShape s) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  constexpr int kStageBytes = MT * kATileBytes + kBTileBytes;
  constexpr int kTmemCols = MT == 1 ? 128 : (MT == 2 ? 256 : 512);
  // 1024-byte aligned carve-up (the 128-byte swizzle pattern repeats every 1024 bytes)
  uint8_t* base = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ __align__(8) uint64_t full_bar[STAGES], empty_bar[STAGES], tmem_full_bar, tmem_empty_bar;
  __shared__ uint32_t tmem_base_s;

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;  // uniform for the compiler

  if (threadIdx.x == 0) {
#pragma unroll
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    mbar_init(&tmem_full_bar, 1);
    mbar_init(&tmem_empty_bar, 4);  // one arrival per epilogue warp
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_w) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_x) : "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_addr(&tmem_base_s)),
                 "r"((uint32_t)kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_s;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int tile = blockIdx.x; tile < s.n_tiles; tile += gridDim.x) {
        const int b = tile / s.tiles_per_img;
        const int hw0 = (tile % s.tiles_per_img) * kBN;
        for (int kc = 0; kc < s.k_chunks; ++kc) {
          mbar_wait(&empty_bar[stage], phase ^ 1u);
          uint8_t* st = base + (size_t)stage * kStageBytes;
          mbar_expect_tx(&full_bar[stage], (uint32_t)kStageBytes);
#pragma unroll
          for (int mt = 0; mt < MT; ++mt)
            tma_load_2d(st + mt * kATileBytes, &tmap_w, kc * kBK, mt * kBM, &full_bar[stage]);
          uint8_t* bt = st + MT * kATileBytes;
          tma_load_3d(bt, &tmap_x, kc * kBK, hw0, b, &full_bar[stage]);  // {32 k, 128 cells}
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    // The whole warp walks the loop and one elected lane issues: with warp-uniform control flow the
    // descriptors live in uniform registers and the tcgen05.mma go out back to back (a lane-0-only
    // loop rebuilt them through R2UR, ~25 instructions per MMA).
    const bool leader = elect_one();
    const uint32_t idesc = umma_idesc_tf32(kBM, kBN);
    const uint64_t desc0 = umma_desc(0, 16, 1024);   // K-major, 8-row groups 1024 B apart; a K step of 8 fp32 = 32 B
    uint32_t stage = 0, phase = 0, acc_phase = 0;
    for (int tile = blockIdx.x; tile < s.n_tiles; tile += gridDim.x) {
      mbar_wait(&tmem_empty_bar, acc_phase ^ 1u);  // epilogue has drained the accumulators
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
      for (int kc = 0; kc < s.k_chunks; ++kc) {
        mbar_wait(&full_bar[stage], phase);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t a0 = smem_addr(base + (size_t)stage * kStageBytes);
        const uint64_t ad = desc0 + (uint64_t)(a0 >> 4), bd = desc0 + (uint64_t)((a0 + MT * kATileBytes) >> 4);
        const uint32_t acc0 = kc != 0 ? 1u : 0u;
        if (leader) {
#pragma unroll
          for (int kk = 0; kk < kBK / kUmmaK; ++kk) {
#pragma unroll
            for (int mt = 0; mt < MT; ++mt)
              umma_tf32(tmem_base + mt * kBN, ad + (uint64_t)((mt * kATileBytes + kk * (kUmmaK * 4)) >> 4),
                        bd + (uint64_t)((kk * (kUmmaK * 4)) >> 4), idesc, kk != 0 ? 1u : acc0);
          }
          umma_commit(&empty_bar[stage]);  // frees the smem stage when these MMAs retire
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
      }
      if (leader) umma_commit(&tmem_full_bar);       // accumulators complete -> epilogue
      __syncwarp();
      acc_phase ^= 1u;
    }
  } else {
    // ------------------------------------------------------------ epilogue (warps 2..5)
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < s.n_tiles; tile += gridDim.x) {
      const int b = tile / s.tiles_per_img;
      const int hw0 = (tile % s.tiles_per_img) * kBN;
      mbar_wait(&tmem_full_bar, acc_phase);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
      for (int mt = 0; mt < MT; ++mt) {
        const int m = mt * kBM + q * 32 + lane;  // output channel of this thread
        const float bm = bias ? bias[m] : 0.f;
        float* yrow = y + ((size_t)b * s.c_out + m) * s.hw + hw0;
#pragma unroll 1
        for (int cc = 0; cc < kBN / 32; ++cc) {
          uint32_t v[32];
          tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(mt * kBN + cc * 32), v);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          if (hw0 + cc * 32 + 32 <= s.hw) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              float4 o = make_float4(__uint_as_float(v[j]) + bm, __uint_as_float(v[j + 1]) + bm,
                                     __uint_as_float(v[j + 2]) + bm, __uint_as_float(v[j + 3]) + bm);
              *reinterpret_cast<float4*>(yrow + cc * 32 + j) = o;
            }
          } else {
            for (int j = 0; j < 32; ++j)
              if (hw0 + cc * 32 + j < s.hw) yrow[cc * 32 + j] = __uint_as_float(v[j]) + bm;
          }
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty_bar);
      acc_phase ^= 1u;
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"((uint32_t)kTmemCols)
                 : "memory");
  }
}

}  // namespace

int adapt_conv1x1_forward(const float* x_cl, const float* w, const float* bias, int batch, int c_in,
                          int c_out, int hw, float* y, cudaStream_t stream) {
  const float* x = x_cl;
  DBEV_CHECK_ARG(batch > 0 && hw > 0, "adapt_conv1x1: empty input");
  DBEV_CHECK_ARG(c_in % kBK == 0 && c_in >= kBK, "adapt_conv1x1: input channels must be a multiple of %d (got %d)",
                 kBK, c_in);
  DBEV_CHECK_ARG(c_out % kBM == 0 && c_out >= kBM && c_out <= 512,
                 "adapt_conv1x1: output channels must be 128, 256, 384 or 512 (got %d)", c_out);
  DBEV_CHECK_ARG(hw % 4 == 0, "adapt_conv1x1: H*W must be a multiple of 4 (got %d)", hw);
  DBEV_CHECK_ARG(((uintptr_t)x % 16) == 0 && ((uintptr_t)w % 16) == 0 && ((uintptr_t)y % 16) == 0,
                 "adapt_conv1x1: tensors must be 16-byte aligned");
  EncodeTiledFn encode = get_encode_fn();
  if (!encode) {
    set_last_error("adapt_conv1x1: cuTensorMapEncodeTiled not available from the driver");
    return DBEV_ERR_CUDA;
  }
  CUtensorMap tmap_w, tmap_x;
  {
    cuuint64_t dims[2] = {(cuuint64_t)c_in, (cuuint64_t)c_out};
    cuuint64_t strides[1] = {(cuuint64_t)c_in * 4};
    cuuint32_t box[2] = {kBK, kBM};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = encode(&tmap_w, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)w, dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_last_error("adapt_conv1x1: cuTensorMapEncodeTiled(W) failed (%d)", (int)r);
      return DBEV_ERR_CUDA;
    }
  }
  {
    cuuint64_t dims[3] = {(cuuint64_t)c_in, (cuuint64_t)hw, (cuuint64_t)batch};
    cuuint64_t strides[2] = {(cuuint64_t)c_in * 4, (cuuint64_t)hw * c_in * 4};
    cuuint32_t box[3] = {kBK, kBN, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = encode(&tmap_x, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)x, dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_last_error("adapt_conv1x1: cuTensorMapEncodeTiled(X) failed (%d)", (int)r);
      return DBEV_ERR_CUDA;
    }
  }
  GemmShape s;
  s.batch = batch; s.c_in = c_in; s.c_out = c_out; s.hw = hw;
  s.tiles_per_img = ceil_div(hw, kBN);
  s.n_tiles = s.tiles_per_img * batch;
  s.k_chunks = c_in / kBK;
  int dev = 0, sms = 0;
  DBEV_CUDA(cudaGetDevice(&dev));
  DBEV_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int grid = s.n_tiles < sms ? s.n_tiles : sms;
  const int mt = c_out / kBM;
#define LAUNCH_GEMM(MTV, STG)                                                                    \
  do {                                                                                           \
    const size_t smem = (size_t)STG * (MTV * kATileBytes + kBTileBytes) + 1024;                  \
    DBEV_CUDA(cudaFuncSetAttribute(adapt_gemm_tf32_kernel<MTV, STG>,                             \
                                   cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));     \
    adapt_gemm_tf32_kernel<MTV, STG><<<grid, kThreads, smem, stream>>>(tmap_w, tmap_x, bias, y, s); \
  } while (0)
  if (mt == 1) LAUNCH_GEMM(1, 4);
  else if (mt == 2) LAUNCH_GEMM(2, 4);
  else if (mt == 3) LAUNCH_GEMM(3, 3);
  else LAUNCH_GEMM(4, 2);
#undef LAUNCH_GEMM
  DBEV_CHECK_LAUNCH("adapt_gemm_tf32_kernel");
  return DBEV_OK;
}

}  // namespace dbev
