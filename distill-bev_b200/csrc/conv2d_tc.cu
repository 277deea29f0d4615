// Dense 2D convolution + folded BatchNorm + ReLU on the 5th-gen tensor cores (tcgen05, TF32) for the
// FROZEN LiDAR teacher's BEV backbone / neck:
//   SECOND.forward      mmdet3d/models/backbones/second.py:80-93   (3x3 conv + BN + ReLU blocks)
//   SECONDFPN.forward   mmdet3d/models/necks/second_fpn.py:77-93   (k2/s2 conv, 1x1 and k2/s2
//                                                                   transposed convs + BN + ReLU, concat)
// The reference runs these through cuDNN (TF32 math under torch's default cudnn.allow_tf32) as
// separate conv, BatchNorm and ReLU kernels. Here one implicit-GEMM kernel per layer, NHWC fp32:
//   D[128 output pixels, C_out] += A[128 pixels, 32 ch] * B[C_out, 32 ch]^T  over (ky, kx, 32-ch chunk)
//   A = the input window of the tile for filter tap (ky, kx): ONE 4-D TMA box {32 ch, TX, TY, 1}
//       of the NHWC tensor at coordinates (c0, x0*s + kx - pad, y0*s + ky - pad, n); the borders
//       are TMA out-of-bounds zero fill, stride-2 layers use the tensor map's element strides -
//       no im2col buffer, no index arithmetic in the kernel;
//   B = weights pre-packed [C_out][(ky*KW + kx)*C_in + ci] (K-major), 2-D TMA box {32, C_out}.
// Accumulators live in TMEM (two buffers: the epilogue of tile t overlaps the MMAs of tile t+1);
// the epilogue applies scale/shift (eval BN folded) + ReLU and writes NHWC rows, optionally into
// a channel slice / strided pixel lattice of a larger tensor (FPN concat, transposed conv).
// Warp roles: 0 = TMA producer, 1 = MMA issuer + TMEM owner, 2-5 = epilogue; persistent CTAs,
// two per SM for C_out <= 128.
#include "conv2d_tc.cuh"

#include "umma.cuh"

namespace dbev {

namespace {

constexpr int kPix = 128;                  // output pixels per tile (UMMA M, TMEM lanes)
constexpr int kKc = 32;                    // channels per stage (one 128 B swizzle row)
constexpr int kATile = kPix * kKc * 4;     // 16 KB
constexpr int kConvThreads = 192;

struct ConvShape {
  int n_img, c_in, c_out, ho, wo, kh, kw, stride, pad;
  int tx, ty, tiles_x, tiles_y, n_tiles, cin_chunks;
  // output placement: pixel (oy*omul + oadd_y, ox*omul + oadd_x) of an [n, H_full, W_full, ld] tensor
  int omul, oadd_y, oadd_x, h_full, w_full, ld, c_off, relu;
  int nchw;  // 1: out is [n, ld, H_full, W_full] (a lane = a pixel: stores of one channel coalesce along x)
};

template <int COUT, int STAGES, int MINB>
__global__ void __launch_bounds__(kConvThreads, MINB)
conv2d_tc_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w,
                 const float* __restrict__ scale, const float* __restrict__ shift,
                 float* __restrict__ out, ConvShape s) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  constexpr int kWTile = COUT * kKc * 4;
  constexpr int kStage = kATile + kWTile;
  constexpr uint32_t kTmemCols = COUT <= 64 ? 128 : (COUT <= 128 ? 256 : 512);  // 2 x COUT, power of 2
  uint8_t* base = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ __align__(8) uint64_t full_bar[STAGES], empty_bar[STAGES], tmem_full_bar[2], tmem_empty_bar[2];
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
#pragma unroll
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    mbar_init(&tmem_full_bar[0], 1);
    mbar_init(&tmem_full_bar[1], 1);
    mbar_init(&tmem_empty_bar[0], 4);
    mbar_init(&tmem_empty_bar[1], 4);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_x) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_w) : "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_addr(&tmem_base_s)),
                 "r"(kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_s;
  const int taps = s.kh * s.kw;
  const int per_img = s.tiles_x * s.tiles_y;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int tile = blockIdx.x; tile < s.n_tiles; tile += gridDim.x) {
        const int n = tile / per_img, r = tile % per_img;
        const int y0 = (r / s.tiles_x) * s.ty, x0 = (r % s.tiles_x) * s.tx;
        for (int tap = 0; tap < taps; ++tap) {
          const int ky = tap / s.kw, kx = tap % s.kw;
          const int ix = x0 * s.stride + kx - s.pad, iy = y0 * s.stride + ky - s.pad;
          for (int cc = 0; cc < s.cin_chunks; ++cc) {
            mbar_wait(&empty_bar[stage], phase ^ 1u);
            uint8_t* st = base + (size_t)stage * kStage;
            mbar_expect_tx(&full_bar[stage], (uint32_t)kStage);
            asm volatile(
                "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
                " [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(smem_addr(st)),
                "l"(&tmap_x), "r"(smem_addr(&full_bar[stage])), "r"(cc * kKc), "r"(ix), "r"(iy), "r"(n)
                : "memory");
            tma_load_2d(st + kATile, &tmap_w, tap * s.c_in + cc * kKc, 0, &full_bar[stage]);
            if (++stage == STAGES) { stage = 0; phase ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_tf32(kPix, COUT);
      const int steps = taps * s.cin_chunks;
      uint32_t stage = 0, phase = 0, it = 0;
      for (int tile = blockIdx.x; tile < s.n_tiles; tile += gridDim.x, ++it) {
        const uint32_t buf = it & 1u, use = it >> 1;
        mbar_wait(&tmem_empty_bar[buf], (use & 1u) ^ 1u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t d_tmem = tmem_base + buf * COUT;
        for (int step = 0; step < steps; ++step) {
          mbar_wait(&full_bar[stage], phase);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t a0 = smem_addr(base + (size_t)stage * kStage);
          const uint32_t b0 = a0 + kATile;
#pragma unroll
          for (int kk = 0; kk < kKc / 8; ++kk) {
            const uint64_t adesc = umma_desc(a0 + kk * 32, 16, 1024);
            const uint64_t bdesc = umma_desc(b0 + kk * 32, 16, 1024);
            umma_tf32(d_tmem, adesc, bdesc, idesc, (step | kk) != 0 ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
        umma_commit(&tmem_full_bar[buf]);
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue (warps 2..5)
    const int q = warp & 3;
    const int row = q * 32 + lane;           // pixel of the tile = TMEM lane
    const int py = row / s.tx, px = row % s.tx;
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < s.n_tiles; tile += gridDim.x, ++it) {
      const uint32_t buf = it & 1u, use = it >> 1;
      const int n = tile / per_img, r = tile % per_img;
      const int oy = (r / s.tiles_x) * s.ty + py, ox = (r % s.tiles_x) * s.tx + px;
      const bool valid = oy < s.ho && ox < s.wo;
      const int fy = oy * s.omul + s.oadd_y, fx = ox * s.omul + s.oadd_x;
      float* orow = out + (((long long)n * s.h_full + fy) * s.w_full + fx) * s.ld + s.c_off;
      const long long plane = (long long)s.h_full * s.w_full;
      float* ocol = out + ((long long)n * s.ld + s.c_off) * plane + (long long)fy * s.w_full + fx;
      mbar_wait(&tmem_full_bar[buf], use & 1u);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
      for (int cc = 0; cc < COUT / 32; ++cc) {
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + buf * COUT + (uint32_t)(cc * 32), v);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (valid) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            float4 o = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]),
                                   __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
            if (scale) {
              const float4 sc = __ldg(reinterpret_cast<const float4*>(scale + cc * 32 + j));
              o.x *= sc.x, o.y *= sc.y, o.z *= sc.z, o.w *= sc.w;
            }
            if (shift) {
              const float4 sh = __ldg(reinterpret_cast<const float4*>(shift + cc * 32 + j));
              o.x += sh.x, o.y += sh.y, o.z += sh.z, o.w += sh.w;
            }
            if (s.relu) o.x = fmaxf(o.x, 0.f), o.y = fmaxf(o.y, 0.f), o.z = fmaxf(o.z, 0.f), o.w = fmaxf(o.w, 0.f);
            if (s.nchw) {
              float* oc = ocol + (long long)(cc * 32 + j) * plane;
              oc[0] = o.x, oc[plane] = o.y, oc[2 * plane] = o.z, oc[3 * plane] = o.w;
            } else {
              *reinterpret_cast<float4*>(orow + cc * 32 + j) = o;
            }
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
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols)
                 : "memory");
  }
}

}  // namespace

int conv2d_tc_forward(const float* x_nhwc, int n_img, int h, int w, int c_in, const float* w_packed,
                      int c_out, int kh, int kw, int stride, int pad, const float* scale,
                      const float* shift, int relu, float* out, int out_h, int out_w, int out_ld,
                      int out_c_off, int out_mul, int out_add_y, int out_add_x, int out_nchw,
                      cudaStream_t stream) {
  DBEV_CHECK_ARG(n_img > 0 && h > 0 && w > 0, "conv2d_tc: empty input");
  DBEV_CHECK_ARG(c_in % kKc == 0 && c_in >= kKc, "conv2d_tc: C_in must be a multiple of 32 (got %d)", c_in);
  DBEV_CHECK_ARG(c_out == 64 || c_out == 128 || c_out == 256,
                 "conv2d_tc: C_out must be 64, 128 or 256 (got %d)", c_out);
  DBEV_CHECK_ARG(kh >= 1 && kh <= 3 && kw >= 1 && kw <= 3 && (stride == 1 || stride == 2) && pad >= 0 && pad <= 1,
                 "conv2d_tc: kernel 1..3, stride 1 or 2, padding 0 or 1");
  DBEV_CHECK_ARG(((uintptr_t)x_nhwc & 15) == 0 && ((uintptr_t)w_packed & 15) == 0 && ((uintptr_t)out & 15) == 0 &&
                     ((uintptr_t)scale & 15) == 0 && ((uintptr_t)shift & 15) == 0,
                 "conv2d_tc: pointers must be 16-byte aligned");
  DBEV_CHECK_ARG(out_ld % 4 == 0 && out_c_off % 4 == 0 && out_c_off + c_out <= out_ld && out_mul >= 1,
                 "conv2d_tc: bad output placement");
  ConvShape s;
  s.n_img = n_img, s.c_in = c_in, s.c_out = c_out, s.kh = kh, s.kw = kw, s.stride = stride, s.pad = pad;
  s.ho = (h + 2 * pad - kh) / stride + 1;
  s.wo = (w + 2 * pad - kw) / stride + 1;
  DBEV_CHECK_ARG(s.ho >= 1 && s.wo >= 1, "conv2d_tc: empty output");
  DBEV_CHECK_ARG((s.ho - 1) * out_mul + out_add_y < out_h && (s.wo - 1) * out_mul + out_add_x < out_w,
                 "conv2d_tc: output lattice exceeds the output tensor");
  // tile = TX x TY output pixels with TX * TY = 128, TX a power of two <= W_out
  int tx = 128;
  while (tx > s.wo) tx >>= 1;
  if (tx < 8) tx = 8;
  s.tx = tx, s.ty = kPix / tx;
  s.tiles_x = ceil_div(s.wo, s.tx), s.tiles_y = ceil_div(s.ho, s.ty);
  s.n_tiles = s.tiles_x * s.tiles_y * n_img;
  s.cin_chunks = c_in / kKc;
  s.omul = out_mul, s.oadd_y = out_add_y, s.oadd_x = out_add_x, s.h_full = out_h, s.w_full = out_w;
  s.ld = out_ld, s.c_off = out_c_off, s.relu = relu, s.nchw = out_nchw ? 1 : 0;
  DBEV_CHECK_ARG(s.tx * stride <= 256 && s.ty * stride <= 256, "conv2d_tc: tile too large for a TMA box");

  EncodeTiledFn encode = get_encode_fn();
  if (!encode) {
    set_last_error("conv2d_tc: cuTensorMapEncodeTiled not available from the driver");
    return DBEV_ERR_CUDA;
  }
  CUtensorMap tmap_x, tmap_w;
  {
    // NHWC input as a 4-D tensor (C, W, H, N); box {32, TX*s, TY*s, 1} traversed with element strides
    // {1, s, s, 1} loads TX x TY pixels; out-of-bounds coordinates (the padding) are zero-filled
    cuuint64_t dims[4] = {(cuuint64_t)c_in, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)n_img};
    cuuint64_t strides[3] = {(cuuint64_t)c_in * 4, (cuuint64_t)w * c_in * 4, (cuuint64_t)h * w * c_in * 4};
    cuuint32_t box[4] = {(cuuint32_t)kKc, (cuuint32_t)(s.tx * stride), (cuuint32_t)(s.ty * stride), 1};
    cuuint32_t estr[4] = {1, (cuuint32_t)stride, (cuuint32_t)stride, 1};
    CUresult r = encode(&tmap_x, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)x_nhwc, dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_last_error("conv2d_tc: cuTensorMapEncodeTiled(x) failed (%d)", (int)r);
      return DBEV_ERR_CUDA;
    }
  }
  {
    const int k_total = kh * kw * c_in;
    cuuint64_t dims[2] = {(cuuint64_t)k_total, (cuuint64_t)c_out};
    cuuint64_t strides[1] = {(cuuint64_t)k_total * 4};
    cuuint32_t box[2] = {(cuuint32_t)kKc, (cuuint32_t)c_out};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = encode(&tmap_w, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)w_packed, dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_last_error("conv2d_tc: cuTensorMapEncodeTiled(w) failed (%d)", (int)r);
      return DBEV_ERR_CUDA;
    }
  }
  int dev = 0, sms = 0;
  DBEV_CUDA(cudaGetDevice(&dev));
  DBEV_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
#define DBEV_CONV_LAUNCH(CO, STG, MB)                                                            \
  do {                                                                                           \
    const size_t smem = (size_t)STG * (kATile + CO * kKc * 4) + 1024;                            \
    const int grid = s.n_tiles < sms * MB ? s.n_tiles : sms * MB;                                \
    DBEV_CUDA(cudaFuncSetAttribute(conv2d_tc_kernel<CO, STG, MB>,                                \
                                   cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));     \
    conv2d_tc_kernel<CO, STG, MB><<<grid, kConvThreads, smem, stream>>>(tmap_x, tmap_w, scale, shift, out, s); \
  } while (0)
  // two CTAs per SM where shared memory and TMEM allow it: one CTA's epilogue / TMA latency hides
  // behind the other's MMAs (1.68 -> 1.56 ms for the whole SECOND + SECONDFPN stack); three CTAs of
  // the 64-channel kernel or deeper stage rings gave nothing
  if (c_out == 64) DBEV_CONV_LAUNCH(64, 4, 2);
  else if (c_out == 128) DBEV_CONV_LAUNCH(128, 3, 2);
  else DBEV_CONV_LAUNCH(256, 4, 1);
#undef DBEV_CONV_LAUNCH
  DBEV_CHECK_LAUNCH("conv2d_tc_kernel");
  return DBEV_OK;
}

}  // namespace dbev
