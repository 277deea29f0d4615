// Dense 2D convolution + folded BatchNorm + ReLU on the 5th-gen tensor cores (tcgen05, TF32) for the
// FROZEN LiDAR teacher's BEV backbone / neck:
//   SECOND.forward      mmdet3d/models/backbones/second.py:80-93   (3x3 conv + BN + ReLU blocks)
//   SECONDFPN.forward   mmdet3d/models/necks/second_fpn.py:77-93   (k2/s2 conv, 1x1 and k2/s2
//                                                                   transposed convs + BN + ReLU, concat)
// The reference runs these through cuDNN (TF32 math under torch's default cudnn.allow_tf32) as
// separate conv, BatchNorm and ReLU kernels. Here one implicit-GEMM kernel per layer, NHWC fp32:
//   D[128 output pixels, C_out] += A[128 pixels, 32 ch] * B[C_out, 32 ch]^T  over (ky, kx, 32-ch chunk)
// Two kernels share the structure (warp 0 TMA producer, warp 1 MMA issuer + TMEM owner, warps 2-5
// epilogue; persistent; accumulators double-buffered in TMEM; programmatic dependent launch):
//   * conv2d_tc_kernel    - any 1..3 x 1..3 filter, stride 1 or 2: the A operand of tap (ky, kx) is ONE
//       4-D TMA box {32 ch, TX, TY, 1} of the NHWC tensor at (c0, x0*s + kx - pad, y0*s + ky - pad, n);
//       borders are TMA out-of-bounds zero fill, stride 2 uses the tensor map's element strides - no
//       im2col buffer, no index arithmetic. Used by the three stride-2 layers and the FPN branches
//       (NCHW / strided-lattice stores, optional column groups for the transposed conv's x taps).
//   * conv3x3_halo_kernel - 3x3 / stride 1 / pad 1 (13 of SECOND's 16 convs): the halo of a 32 x 8 pixel
//       tile is loaded once per 32-channel chunk and the nine taps are shifted descriptor windows of it.
//   B = weights pre-packed [C_out][(ky*KW + kx)*C_in + ci] (K-major), 2-D TMA box {32, C_out}.
// The epilogue applies scale/shift (eval BN folded) + ReLU; NHWC outputs leave through a swizzled
// shared-memory staging tile and a TMA store. Measurements and the order in which the bottlenecks
// were found (MMA issue -> L2 fill -> epilogue): DESIGN.md §2.10, profiles/r01_conv3x3_halo.json.
#include "conv2d_tc.cuh"

#include "umma.cuh"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <utility>

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
  int tma_store;  // 1: NHWC output on the plain lattice, written by TMA from a shared-memory staging tile
  int group_cols; // > 0: column block g = col / group_cols goes to lattice x + g, channel col % group_cols
                  // (both x taps of a stride-2 transposed conv in one launch)
  int ncb;        // column blocks: C_out total = ncb * COUT; a work item = (pixel tile, column block), block fastest
  int accumulate; // 1: out += result (TMA reduce-add store / read-modify-write) - gradient accumulation of the
                  // training path (residual branches of BasicBlock, res_block.py:70-99)
  // n_cls > 1: several convolutions of the SAME input with different filters / output lattices in one launch (the
  // four parity classes of a stride-2 input gradient): a work item = (class, pixel tile, column block); class c has
  // a ckh[c] x ckw[c] filter (its own weight matrix) and writes lattice offset (coy[c], cox[c]). n_cls == 1: the
  // scalars above.
  int n_cls, ckh[4], ckw[4], coy[4], cox[4];
};

// weight / output tensor maps of classes 1..3 (class 0 uses tmap_w / tmap_o)
struct ClassMaps {
  CUtensorMap w[3], o[3];
};

template <int COUT, int STAGES, int MINB>
__global__ void __launch_bounds__(kConvThreads, MINB)
conv2d_tc_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w,
                 const __grid_constant__ CUtensorMap tmap_o, const float* __restrict__ scale, const float* __restrict__ shift,
                 float* __restrict__ out, ConvShape s, const __grid_constant__ ClassMaps cmaps) {
  // programmatic dependent launch: let the next kernel of the stream start its prologue on SMs this grid has left
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  constexpr int kWTile = COUT * kKc * 4;
  constexpr int kStage = kATile + kWTile;
  constexpr uint32_t kTmemCols = COUT <= 64 ? 128 : (COUT <= 128 ? 256 : 512);  // 2 x COUT, power of 2
  uint8_t* base = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ __align__(8) uint64_t full_bar[STAGES], empty_bar[STAGES], tmem_full_bar[2], tmem_empty_bar[2];
  __shared__ uint32_t tmem_base_s;
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;  // uniform for the compiler

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
  // everything above (barriers, TMEM, tensor-map prefetch) overlapped the previous kernel's tail; its results
  // (this kernel's input) are complete and visible after this point, and nothing is written before it
  asm volatile("griddepcontrol.wait;" ::: "memory");
  const int per_img = s.tiles_x * s.tiles_y;
  const int per_cls = s.n_tiles * s.ncb, n_items = per_cls * s.n_cls;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int cls = item / per_cls, rem = item - cls * per_cls;
        const int tile = rem / s.ncb, cb = rem - tile * s.ncb;
        const int n = tile / per_img, r = tile % per_img;
        const int y0 = (r / s.tiles_x) * s.ty, x0 = (r % s.tiles_x) * s.tx;
        const int kw_c = s.ckw[cls], taps = s.ckh[cls] * kw_c;
        const CUtensorMap* tw = cls == 0 ? &tmap_w : &cmaps.w[cls - 1];
        for (int tap = 0; tap < taps; ++tap) {
          const int ky = tap / kw_c, kx = tap % kw_c;
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
            tma_load_2d(st + kATile, tw, tap * s.c_in + cc * kKc, cb * COUT, &full_bar[stage]);
            if (++stage == STAGES) { stage = 0; phase ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    // whole warp in the loop, one elected lane issues (see conv3x3_halo_kernel)
    const bool leader = elect_one();
    const uint32_t idesc = umma_idesc_tf32(kPix, COUT);
    const uint64_t desc0 = umma_desc(0, 16, 1024);
    uint32_t stage = 0, phase = 0, it = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
      const int cls_m = item / per_cls;
      const int steps = s.ckh[cls_m] * s.ckw[cls_m] * s.cin_chunks;
      const uint32_t buf = it & 1u, use = it >> 1;
      mbar_wait(&tmem_empty_bar[buf], (use & 1u) ^ 1u);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t d_tmem = tmem_base + buf * COUT;
#pragma unroll 1
      for (int step = 0; step < steps; ++step) {
        mbar_wait(&full_bar[stage], phase);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t a0 = smem_addr(base + (size_t)stage * kStage);
        const uint64_t ad = desc0 + (uint64_t)(a0 >> 4), bd = desc0 + (uint64_t)((a0 + kATile) >> 4);
        const uint32_t acc0 = step != 0 ? 1u : 0u;
        if (leader) {
#pragma unroll
          for (int kk = 0; kk < kKc / 8; ++kk)
            umma_tf32(d_tmem, ad + (uint64_t)(kk * 2), bd + (uint64_t)(kk * 2), idesc, kk != 0 ? 1u : acc0);
          umma_commit(&empty_bar[stage]);
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
      }
      if (leader) umma_commit(&tmem_full_bar[buf]);
      __syncwarp();
    }
  } else {
    // ------------------------------------------------------------ epilogue (warps 2..5)
    // NHWC outputs on the plain pixel lattice leave through shared memory + TMA store (see
    // conv3x3_halo_kernel); NCHW / strided-lattice outputs (FPN) are written directly, a lane = a pixel.
    const int q = warp & 3;
    const int row = q * 32 + lane;           // pixel of the tile = TMEM lane
    const int py = row / s.tx, px = row % s.tx;
    const int by0 = (q * 32) / s.tx, bx0 = (q * 32) % s.tx;   // this warp's store box inside the tile
    uint8_t* stg = base + (size_t)STAGES * kStage + (size_t)q * 8192;
    const uint32_t stg_row = smem_addr(stg) + (uint32_t)lane * 128u;
    const uint32_t sw_xor = (uint32_t)(lane & 7);
    uint32_t it = 0, sbuf = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
      const uint32_t buf = it & 1u, use = it >> 1;
      const int cls = item / per_cls, rem = item - cls * per_cls;
      const int tile = rem / s.ncb, cb = rem - tile * s.ncb;
      const CUtensorMap* to = cls == 0 ? &tmap_o : &cmaps.o[cls - 1];
      const int n = tile / per_img, r = tile % per_img;
      const int ty0 = (r / s.tiles_x) * s.ty, tx0 = (r % s.tiles_x) * s.tx;
      const int oy = ty0 + py, ox = tx0 + px;
      const bool valid = oy < s.ho && ox < s.wo;
      const int fy = oy * s.omul + s.coy[cls], fx = ox * s.omul + s.cox[cls];
      const long long plane = (long long)s.h_full * s.w_full;
      float* orow = out + (((long long)n * s.h_full + fy) * s.w_full + fx) * s.ld + s.c_off;
      float* ocol = out + ((long long)n * s.ld + s.c_off) * plane + (long long)fy * s.w_full + fx;
      mbar_wait(&tmem_full_bar[buf], use & 1u);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
      for (int cc = 0; cc < COUT / 32; ++cc) {
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + buf * COUT + (uint32_t)(cc * 32), v);
        if (s.tma_store) {
          if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
          __syncwarp();
        }
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        const uint32_t dst = stg_row + sbuf * 4096u;
        int grp = 0, ch0 = cb * COUT + cc * 32;   // lattice x offset and first output channel of this column chunk
        if (s.group_cols > 0) grp = ch0 / s.group_cols, ch0 -= grp * s.group_cols;
        if (valid || s.tma_store) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            float4 o = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]),
                                   __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
            if (scale) {
              const float4 sc = __ldg(reinterpret_cast<const float4*>(scale + ch0 + j));
              o.x *= sc.x, o.y *= sc.y, o.z *= sc.z, o.w *= sc.w;
            }
            if (shift) {
              const float4 sh = __ldg(reinterpret_cast<const float4*>(shift + ch0 + j));
              o.x += sh.x, o.y += sh.y, o.z += sh.z, o.w += sh.w;
            }
            if (s.relu) o.x = fmaxf(o.x, 0.f), o.y = fmaxf(o.y, 0.f), o.z = fmaxf(o.z, 0.f), o.w = fmaxf(o.w, 0.f);
            if (s.tma_store) {
              asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(dst + ((((uint32_t)j >> 2) ^ sw_xor) << 4)),
                           "f"(o.x), "f"(o.y), "f"(o.z), "f"(o.w)
                           : "memory");
            } else if (s.nchw) {
              float* oc = ocol + grp + (long long)(ch0 + j) * plane;
              if (s.accumulate) o.x += oc[0], o.y += oc[plane], o.z += oc[2 * plane], o.w += oc[3 * plane];
              oc[0] = o.x, oc[plane] = o.y, oc[2 * plane] = o.z, oc[3 * plane] = o.w;
            } else {
              float4* op = reinterpret_cast<float4*>(orow + (long long)grp * s.ld + ch0 + j);
              if (s.accumulate) {
                const float4 prev = *op;
                o.x += prev.x, o.y += prev.y, o.z += prev.z, o.w += prev.w;
              }
              *op = o;
            }
          }
        }
        if (s.tma_store) {
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          __syncwarp();
          if (lane == 0) {
            if (ty0 + by0 < s.ho && tx0 + bx0 < s.wo) {
              if (s.accumulate)
                asm volatile(
                    "cp.reduce.async.bulk.tensor.4d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(to),
                    "r"(smem_addr(stg) + sbuf * 4096u), "r"(s.c_off + ch0), "r"(tx0 + bx0), "r"(ty0 + by0), "r"(n)
                    : "memory");
              else
                asm volatile(
                    "cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(to),
                    "r"(smem_addr(stg) + sbuf * 4096u), "r"(s.c_off + ch0), "r"(tx0 + bx0), "r"(ty0 + by0), "r"(n)
                    : "memory");
            }
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
          sbuf ^= 1u;
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty_bar[buf]);
    }
    if (s.tma_store) {
      if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
      __syncwarp();
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols)
                 : "memory");
  }
}

// ---------------------------------------------------------------------------------------------
// 3x3 / stride 1 / pad 1 layers (13 of SECOND's 16 convs): HALO-REUSE variant.
// The per-tap kernel above re-reads the input window from L2 once per filter tap and the weights
// once per 128-pixel tile; at ~60 B/clk/SM of TMA fill the whole chip sits on the L2 slice
// throughput cap (~6300 B/clk), the tensor pipe idles (27 / 50 / 67 % busy at 64 / 128 / 256 ch).
// Here one TMA box brings the (32+2) x (8+2)-pixel halo of a 32 x 8-pixel output tile for a 32-channel
// chunk ONCE, and the nine taps are nine shifted WINDOWS of it: the A descriptor of tap (ky, kx) starts
// at pixel row (ky*HX + kx) of the halo and strides HX*128 B between 8-pixel groups (the SWIZZLE_128B
// XOR is a function of the shared-memory address bits, so a window that starts mid-pattern reads what
// TMA wrote). The tile is two M=128 halves (16 image rows each) that share every weight tile, so the
// weight traffic per pixel halves as well. L2->SMEM bytes per 128 pixels (C_in = C_out):
//   64 ch: 435 KB -> 117 KB, 128 ch: 1166 KB -> 387 KB, 256 ch: 3.5 MB -> 1.36 MB.
// Warp roles as above; one CTA per SM; accumulators 2 halves x 2 buffers (x 1 at C_out = 256).
constexpr int kHaloTx = 8;                        // tile width; a tile is one or two M=128 halves of 16 image rows

struct HaloShape {
  int n_img, c_in, ho, wo, tiles_x, tiles_y, cin_chunks;
  // work items: the first n_full items are whole tiles (item_halves halves of 16 rows each); the tiles of
  // the last, partial round are split into single halves so that the round lasts half as long
  int n_full, n_items, item_halves, tile_rows;
  int hx;            // halo pitch in pixels (box width), >= kHaloTx + 2
  int a_stage;       // bytes of one halo stage (multiple of 1024)
  int w_stages;      // weight ring depth
  int a_stages;      // halo ring depth (2; 3 for the 64-column tiles whose chunks last only ~1700 clk)
  int c_off, relu, accumulate;
  int ncb;           // column blocks (see ConvShape)
  int ksplit, cpk;   // split-K: a work item = (tile item, column block, K part), K part fastest; each part walks cpk =
                     // cin_chunks / ksplit input-channel chunks and adds its result to the (zeroed) output with a TMA
                     // reduce-add - two parts commute, so the sum is bit-reproducible. Small maps only: 16 x 16 maps have
                     // 16 pixel tiles, and 128-column tiles (full-rate MMAs) would otherwise leave half of the SMs idle.
};

__device__ __forceinline__ void halo_item(const HaloShape& s, int i, int& n, int& y0, int& x0, int& halves) {
  int t = i, yoff = 0;
  halves = s.item_halves;
  if (i >= s.n_full) {
    const int j = i - s.n_full;
    t = s.n_full + (j >> 1), yoff = (j & 1) * 16, halves = 1;
  }
  const int per_img = s.tiles_x * s.tiles_y;
  n = t / per_img;
  const int r = t - n * per_img;
  y0 = (r / s.tiles_x) * s.tile_rows + yoff, x0 = (r % s.tiles_x) * kHaloTx;
}

template <int COUT>
__global__ void __launch_bounds__(kConvThreads, 1)
conv3x3_halo_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w,
                    const __grid_constant__ CUtensorMap tmap_o, const float* __restrict__ scale, const float* __restrict__ shift,
                    float* __restrict__ out, HaloShape s, long long* __restrict__ prof) {
  // programmatic dependent launch: let the next kernel of the stream start its prologue on SMs this grid has left
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  constexpr int kWTile = COUT * kKc * 4;
  constexpr int kNBuf = COUT <= 128 ? 2 : 1;
  constexpr uint32_t kTmemCols = COUT <= 64 ? 256 : 512;     // 2 halves x kNBuf x COUT
  constexpr int kMaxW = 16;
  constexpr int kMaxA = 3;
  uint8_t* base = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* wbase = base + (size_t)s.a_stages * s.a_stage;
  __shared__ __align__(8) uint64_t a_full[kMaxA], a_empty[kMaxA], w_full[kMaxW], w_empty[kMaxW], tmem_full_bar[2],
      tmem_empty_bar[2];
  __shared__ uint32_t tmem_base_s;
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;  // uniform for the compiler

  if (threadIdx.x == 0) {
    for (int i = 0; i < kMaxA; ++i) {
      mbar_init(&a_full[i], 1);
      mbar_init(&a_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full_bar[i], 1);
      mbar_init(&tmem_empty_bar[i], 4);
    }
    for (int i = 0; i < kMaxW; ++i) {
      mbar_init(&w_full[i], 1);
      mbar_init(&w_empty[i], 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_x) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_w) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_o) : "memory");
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
  // everything above (barriers, TMEM, tensor-map prefetch) overlapped the previous kernel's tail; its results
  // (this kernel's input) are complete and visible after this point, and nothing is written before it
  asm volatile("griddepcontrol.wait;" ::: "memory");
  const uint32_t a_bytes = (uint32_t)((s.tile_rows + 2) * s.hx * kKc * 4);

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      uint32_t sa = 0, pa = 0, sw = 0, pw = 0;
      long long pw_a = 0, pw_w = 0;
      const long long pt0 = prof ? clock64() : 0;
      // the halo of chunk i+1 is requested while the weights of chunk i stream: at tap `ahead` the MMA
      // warp (w_stages taps behind) has left chunk i-1, so its halo stage is free without waiting
      int t_n = blockIdx.x, c_n = 0;
      const int n_work = s.n_items * s.ncb * s.ksplit;
      auto issue_halo = [&]() {
        if (t_n >= n_work) return;
        int n, y0, x0, halves;
        halo_item(s, (t_n / s.ksplit) / s.ncb, n, y0, x0, halves);
        const int chunk = (t_n % s.ksplit) * s.cpk + c_n;
        { const long long t = prof ? clock64() : 0; mbar_wait(&a_empty[sa], pa ^ 1u); if (prof) pw_a += clock64() - t; }
        mbar_expect_tx(&a_full[sa], a_bytes);
        asm volatile(
            "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
            " [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(smem_addr(base + (size_t)sa * s.a_stage)),
            "l"(&tmap_x), "r"(smem_addr(&a_full[sa])), "r"(chunk * kKc), "r"(x0 - 1), "r"(y0 - 1), "r"(n)
            : "memory");
        if (++sa == (uint32_t)s.a_stages) { sa = 0; pa ^= 1u; }
        if (++c_n == s.cpk) { c_n = 0; t_n += gridDim.x; }
      };
      const int ahead = s.w_stages < 8 ? s.w_stages : 8;
      for (int i = 0; i < s.a_stages - 1; ++i) issue_halo();
      for (int item = blockIdx.x; item < n_work; item += gridDim.x) {
        const int cb = (item / s.ksplit) % s.ncb, c0 = (item % s.ksplit) * s.cpk;
        for (int cc = c0; cc < c0 + s.cpk; ++cc) {
#pragma unroll 1
          for (int tap = 0; tap < 9; ++tap) {
            if (tap == ahead) issue_halo();
            { const long long t = prof ? clock64() : 0; mbar_wait(&w_empty[sw], pw ^ 1u); if (prof) pw_w += clock64() - t; }
            mbar_expect_tx(&w_full[sw], (uint32_t)kWTile);
            tma_load_2d(wbase + (size_t)sw * kWTile, &tmap_w, tap * s.c_in + cc * kKc, cb * COUT, &w_full[sw]);
            if (++sw == (uint32_t)s.w_stages) { sw = 0; pw ^= 1u; }
          }
        }
      }
      if (prof) { prof[blockIdx.x * 16 + 0] = clock64() - pt0; prof[blockIdx.x * 16 + 1] = pw_a; prof[blockIdx.x * 16 + 2] = pw_w; }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    // The whole warp walks the loop (warp-uniform control flow keeps the descriptors in uniform
    // registers: a tcgen05.mma is issued every few instructions instead of every ~25); one elected
    // lane issues the MMAs and commits. At N = 64 an MMA lasts ~45 clk, so issue cost is what bounds it.
    const bool leader = elect_one();
    const uint32_t idesc = umma_idesc_tf32(kPix, COUT);
    const uint64_t desc_a = umma_desc(0, 16, (uint32_t)s.hx * 128u), desc_b = umma_desc(0, 16, 1024);
    const uint32_t half_units = (uint32_t)(16 * s.hx * 8);   // 16 halo rows, in 16-byte units
    uint32_t sa = 0, pa = 0, sw = 0, pw = 0, it = 0;
    long long mw_t = 0, mw_a = 0, mw_w = 0;
    const long long mt0 = prof ? clock64() : 0;
    for (int item = blockIdx.x; item < s.n_items * s.ncb * s.ksplit; item += gridDim.x, ++it) {
      const uint32_t buf = kNBuf == 2 ? (it & 1u) : 0u, use = kNBuf == 2 ? (it >> 1) : it;
      const int halves = (item / s.ksplit) / s.ncb < s.n_full ? s.item_halves : 1;
      { const long long t = prof ? clock64() : 0; mbar_wait(&tmem_empty_bar[buf], (use & 1u) ^ 1u); if (prof) mw_t += clock64() - t; }
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t d_tmem = tmem_base + buf * 2 * COUT;
      for (int cc = 0; cc < s.cpk; ++cc) {
        { const long long t = prof ? clock64() : 0; mbar_wait(&a_full[sa], pa); if (prof) mw_a += clock64() - t; }
        const uint32_t a_units = smem_addr(base + (size_t)sa * s.a_stage) >> 4;
#pragma unroll 1
        for (int tap = 0; tap < 9; ++tap) {
          const int ky = tap / 3, kx = tap - ky * 3;
          { const long long t = prof ? clock64() : 0; mbar_wait(&w_full[sw], pw); if (prof) mw_w += clock64() - t; }
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint64_t ad = desc_a + (uint64_t)(a_units + (uint32_t)((ky * s.hx + kx) * 8));
          const uint64_t bd = desc_b + (uint64_t)(smem_addr(wbase + (size_t)sw * kWTile) >> 4);
          const uint32_t acc0 = (cc | tap) != 0 ? 1u : 0u;
          if (leader) {
#pragma unroll
            for (int kk = 0; kk < kKc / 8; ++kk)
              umma_tf32(d_tmem, ad + (uint64_t)(kk * 2), bd + (uint64_t)(kk * 2), idesc, kk != 0 ? 1u : acc0);
            if (halves == 2) {
#pragma unroll
              for (int kk = 0; kk < kKc / 8; ++kk)
                umma_tf32(d_tmem + COUT, ad + (uint64_t)(half_units + kk * 2), bd + (uint64_t)(kk * 2), idesc,
                          kk != 0 ? 1u : acc0);
            }
            umma_commit(&w_empty[sw]);
          }
          __syncwarp();
          if (++sw == (uint32_t)s.w_stages) { sw = 0; pw ^= 1u; }
        }
        if (leader) umma_commit(&a_empty[sa]);
        __syncwarp();
        if (++sa == (uint32_t)s.a_stages) { sa = 0; pa ^= 1u; }
      }
      if (leader) umma_commit(&tmem_full_bar[buf]);
      __syncwarp();
    }
    if (prof && leader) { prof[blockIdx.x * 16 + 4] = clock64() - mt0; prof[blockIdx.x * 16 + 5] = mw_t; prof[blockIdx.x * 16 + 6] = mw_a; prof[blockIdx.x * 16 + 7] = mw_w; }
  } else {
    // ------------------------------------------------------------ epilogue (warps 2..5)
    // A lane owns one pixel (TMEM lane) and 32 channels per pass. Writing those straight to NHWC
    // global memory is a 16-byte store per lane at pixel stride: 32 partial sectors per instruction,
    // measured 25k clk per 128 KB tile (the MMAs of that tile: 18k) - the whole kernel ran at the
    // speed of its epilogue. Instead each warp stages its 32 pixels x 32 channels in shared memory
    // (128-byte swizzle, conflict-free) and one lane hands the box {32 ch, 8 px, 4 rows} to TMA.
    const int q = warp & 3;                  // TMEM lane q*32 + lane = pixel (lane / 8, lane % 8) of the warp's box
    uint8_t* stg = base + (size_t)s.a_stages * s.a_stage + (size_t)s.w_stages * kWTile + (size_t)q * 8192;
    const uint32_t stg_row = smem_addr(stg) + (uint32_t)lane * 128u;
    const uint32_t sw_xor = (uint32_t)(lane & 7);
    uint32_t it = 0, sbuf = 0;
    long long ew_f = 0;
    const long long et0 = prof ? clock64() : 0;
    for (int item = blockIdx.x; item < s.n_items * s.ncb * s.ksplit; item += gridDim.x, ++it) {
      const uint32_t buf = kNBuf == 2 ? (it & 1u) : 0u, use = kNBuf == 2 ? (it >> 1) : it;
      int n, ty0, tx0, halves;
      const int cb = (item / s.ksplit) % s.ncb;
      const bool first_part = item % s.ksplit == 0;      // the bias is added by one K part only
      halo_item(s, (item / s.ksplit) / s.ncb, n, ty0, tx0, halves);
      { const long long t = prof ? clock64() : 0; mbar_wait(&tmem_full_bar[buf], use & 1u); if (prof) ew_f += clock64() - t; }
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
      for (int h = 0; h < halves; ++h) {
        const int oy0 = ty0 + h * 16 + q * 4;      // first image row of this warp's 4 x 8 pixel box
#pragma unroll 1
        for (int cc = 0; cc < COUT / 32; ++cc) {
          uint32_t v[32];
          tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (buf * 2 + h) * COUT + (uint32_t)(cc * 32), v);
          // the staging buffer written two passes ago must have been read by its TMA store
          if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
          __syncwarp();
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          const uint32_t dst = stg_row + sbuf * 4096u;
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            float4 o = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]),
                                   __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
            if (scale) {
              const float4 sc = __ldg(reinterpret_cast<const float4*>(scale + cb * COUT + cc * 32 + j));
              o.x *= sc.x, o.y *= sc.y, o.z *= sc.z, o.w *= sc.w;
            }
            if (shift && first_part) {
              const float4 sh = __ldg(reinterpret_cast<const float4*>(shift + cb * COUT + cc * 32 + j));
              o.x += sh.x, o.y += sh.y, o.z += sh.z, o.w += sh.w;
            }
            if (s.relu) o.x = fmaxf(o.x, 0.f), o.y = fmaxf(o.y, 0.f), o.z = fmaxf(o.z, 0.f), o.w = fmaxf(o.w, 0.f);
            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(dst + ((((uint32_t)j >> 2) ^ sw_xor) << 4)),
                         "f"(o.x), "f"(o.y), "f"(o.z), "f"(o.w)
                         : "memory");
          }
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          __syncwarp();
          if (lane == 0 && oy0 < s.ho) {
            if (s.accumulate)
              asm volatile(
                  "cp.reduce.async.bulk.tensor.4d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(&tmap_o),
                  "r"(smem_addr(stg) + sbuf * 4096u), "r"(s.c_off + cb * COUT + cc * 32), "r"(tx0), "r"(oy0), "r"(n)
                  : "memory");
            else
              asm volatile(
                  "cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(&tmap_o),
                  "r"(smem_addr(stg) + sbuf * 4096u), "r"(s.c_off + cb * COUT + cc * 32), "r"(tx0), "r"(oy0), "r"(n)
                  : "memory");
          }
          if (lane == 0) asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          sbuf ^= 1u;
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty_bar[buf]);
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    __syncwarp();
    if (prof && threadIdx.x == 64) { prof[blockIdx.x * 16 + 8] = clock64() - et0; prof[blockIdx.x * 16 + 9] = ew_f; }
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols)
                 : "memory");
  }
}

// Launch with the programmatic-stream-serialization attribute (PDL): the kernel may become resident while
// its predecessor in the stream drains; it blocks in griddepcontrol.wait before touching memory.
// DBEV_CONV_PDL=0 falls back to plain launches.
template <typename... KArgs, typename... Args>
cudaError_t launch_conv(void (*kernel)(KArgs...), int grid, size_t smem, cudaStream_t stream, Args&&... args) {
  static const bool pdl = !(getenv("DBEV_CONV_PDL") && atoi(getenv("DBEV_CONV_PDL")) == 0);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid), cfg.blockDim = dim3(kConvThreads), cfg.dynamicSmemBytes = smem, cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr, cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

// Experiment switch: shared memory the conv kernels leave free on every SM (DBEV_SMEM_RESERVE bytes, default 0) so that
// blocks of the memory-bound kernels of OTHER streams (BatchNorm statistics / apply, <= 5 KB each) can be co-resident
// with a persistent conv CTA instead of waiting for the whole conv kernel to retire. Measured on the bench step:
// 0 -> 7.70 ms, 10 KB -> 7.72 ms, 24 KB -> 7.88 ms (the conv rings get shallower, the overlap gains nothing because the
// tensor kernels already cover 76 % of the step and the rest is dependency-bound): DESIGN.md 7.1.
int smem_reserve() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("DBEV_SMEM_RESERVE");
    v = e ? atoi(e) : 0;
    if (v < 0) v = 0;
    if (v > 65536) v = 65536;
  }
  return v;
}

// A/B switch: DBEV_HALO_KSPLIT=0 disables split-K in the halo kernel.
bool halo_ksplit() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("DBEV_HALO_KSPLIT");
    v = e ? (atoi(e) != 0) : 1;
  }
  return v != 0;
}

// A/B switch: DBEV_CONV_LATTICE_TMA=0 keeps the direct stores for strided-lattice NHWC outputs.
bool tma_lattice_store() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("DBEV_CONV_LATTICE_TMA");
    v = e ? (atoi(e) != 0) : 1;
  }
  return v != 0;
}

// A/B switch: DBEV_CONV_HALO = 0 (per-tap kernel for every layer), 1 (default: halo kernel, pitch 10),
// 2 (pitch 16), 4 (no split of the last round). Measured: both pitches are bit-identical to the per-tap kernel (descriptor base-offset 0);
// setting the descriptor's base-offset field to (start >> 7) & 7 gives wrong results.
int conv_halo_mode() {
  static int mode = -1;
  if (mode < 0) {
    const char* e = getenv("DBEV_CONV_HALO");
    mode = e ? atoi(e) : 1;
  }
  return mode;
}

}  // namespace

int conv2d_tc_forward(const float* x_nhwc, int n_img, int h, int w, int c_in, const float* w_packed,
                      int c_out, int kh, int kw, int stride, int pad, const float* scale,
                      const float* shift, int relu, float* out, int out_h, int out_w, int out_ld,
                      int out_c_off, int out_mul, int out_add_y, int out_add_x, int out_nchw,
                      int out_groups, cudaStream_t stream) {
  return conv2d_tc_forward_ex(x_nhwc, n_img, h, w, c_in, c_in, w_packed, c_out, 1, kh, kw, stride, pad, scale, shift, relu, out,
                              out_h, out_w, out_ld, out_c_off, out_mul, out_add_y, out_add_x, out_nchw, out_groups, 0, 0,
                              0, stream);
}

// x_ld: channel stride of the input rows (>= c_in: the input may be a channel slice of a wider NHWC tensor);
// force_ho / force_wo > 0 override the output size (taps that fall outside the input read zeros: the four
// output-parity classes of a stride-2 input gradient are stride-1 convolutions of dy padded on ONE side);
// accumulate != 0 adds to the output instead of overwriting it; n_col_blocks > 1: the layer has
// n_col_blocks * c_out output channels (w_packed / scale / shift / the output slice cover all of them) and a work
// item is (pixel tile, block of c_out columns) - layers with few pixel tiles (16 x 16 .. 64 x 64 BEV maps) fill
// the SMs with column blocks instead of running one under-filled launch per block.
namespace {

// Classes of a multi-class launch (ConvShape::n_cls): filter size, weight matrix and lattice offset per class.
struct ConvClasses {
  int n;
  int kh[4], kw[4], add_y[4], add_x[4];
  const float* w[4];
};

int conv2d_tc_impl(const float* x_nhwc, int n_img, int h, int w, int c_in, int x_ld, const float* w_packed,
                   int c_out, int n_col_blocks, int kh, int kw, int stride, int pad, const float* scale,
                   const float* shift, int relu, float* out, int out_h, int out_w, int out_ld,
                   int out_c_off, int out_mul, int out_add_y, int out_add_x, int out_nchw,
                   int out_groups, int force_ho, int force_wo, int accumulate, const ConvClasses* classes, cudaStream_t stream) {
  const int n_cls = classes ? classes->n : 1;
  DBEV_CHECK_ARG(n_cls >= 1 && n_cls <= 4, "conv2d_tc: 1..4 classes");
  DBEV_CHECK_ARG(n_cls == 1 || (out_groups == 1 && !out_nchw && force_ho > 0 && force_wo > 0 && stride == 1),
                 "conv2d_tc: multi-class launches are stride-1 convolutions with a forced output size and NHWC lattice outputs");
  DBEV_CHECK_ARG(x_ld >= c_in && x_ld % 4 == 0, "conv2d_tc: bad input channel stride");
  DBEV_CHECK_ARG(n_col_blocks >= 1 && (n_col_blocks == 1 || out_groups == 1), "conv2d_tc: column blocks need out_groups == 1");
  const int ncb = n_col_blocks;
  DBEV_CHECK_ARG(n_img > 0 && h > 0 && w > 0, "conv2d_tc: empty input");
  DBEV_CHECK_ARG(c_in % kKc == 0 && c_in >= kKc, "conv2d_tc: C_in must be a multiple of 32 (got %d)", c_in);
  DBEV_CHECK_ARG(c_out == 64 || c_out == 128 || c_out == 256,
                 "conv2d_tc: C_out must be 64, 128 or 256 (got %d)", c_out);
  DBEV_CHECK_ARG(kh >= 1 && kh <= 3 && kw >= 1 && kw <= 3 && (stride == 1 || stride == 2) && pad >= 0 && pad <= 1,
                 "conv2d_tc: kernel 1..3, stride 1 or 2, padding 0 or 1");
  DBEV_CHECK_ARG(((uintptr_t)x_nhwc & 15) == 0 && ((uintptr_t)w_packed & 15) == 0 && ((uintptr_t)out & 15) == 0 &&
                     ((uintptr_t)scale & 15) == 0 && ((uintptr_t)shift & 15) == 0,
                 "conv2d_tc: pointers must be 16-byte aligned");
  DBEV_CHECK_ARG(out_groups >= 1 && out_groups <= out_mul && c_out % out_groups == 0 && (c_out / out_groups) % 32 == 0,
                 "conv2d_tc: out_groups must divide C_out into multiples of 32 columns and fit the lattice step");
  DBEV_CHECK_ARG(out_ld % 4 == 0 && out_c_off % 4 == 0 && out_c_off + c_out * ncb / out_groups <= out_ld && out_mul >= 1,
                 "conv2d_tc: bad output placement");
  EncodeTiledFn encode = get_encode_fn();
  if (!encode) {
    set_last_error("conv2d_tc: cuTensorMapEncodeTiled not available from the driver");
    return DBEV_ERR_CUDA;
  }
  int dev = 0, sms = 0;
  DBEV_CUDA(cudaGetDevice(&dev));
  DBEV_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int halo_mode = conv_halo_mode();
  if (n_cls == 1 && halo_mode > 0 && out_groups == 1 && kh == 3 && kw == 3 && stride == 1 && pad == 1 && h >= 16 && w >= kHaloTx && out_mul == 1 &&
      out_add_y == 0 && out_add_x == 0 && !out_nchw && out_h == h && out_w == w && force_ho == 0 && force_wo == 0) {
    HaloShape hs;
    hs.n_img = n_img, hs.c_in = c_in, hs.ho = h, hs.wo = w;
    hs.cin_chunks = c_in / kKc;
    hs.hx = halo_mode == 2 ? 16 : kHaloTx + 2;
    const int w_tile = c_out * kKc * 4;
    const int stage_out = 4 * 2 * 4096;       // 4 epilogue warps x 2 staging buffers x (32 px x 128 B)
    const int smem_max = 227 * 1024 - 2048 - smem_reserve();
    // (keeping all 9 * C_in/32 weight tiles of a 64 -> 64 layer resident in shared memory was tried: only
    // single-half tiles fit beside them, whose halo prefetch distance is too short - 0.132 vs 0.110 ms)
    hs.item_halves = h > 16 ? 2 : 1;      // 16-row images (the student's 512-channel stage): one M=128 half per item
    // small maps: if 256-pixel items would leave more than half of the SMs idle, use 128-pixel items (twice as many):
    // the caller can then take wider column tiles for the same item count (N = 128 MMAs run at the full tensor rate,
    // N = 64 at ~45 % - 71 vs 32 clk per MMA, the shared-memory operand floor)
    if (hs.item_halves == 2 && 2LL * ceil_div(w, kHaloTx) * ceil_div(h, 32) * n_img * ncb <= (long long)sms + sms / 8)
      hs.item_halves = 1;
    hs.tile_rows = 16 * hs.item_halves;
    hs.tiles_x = ceil_div(w, kHaloTx), hs.tiles_y = ceil_div(h, hs.tile_rows);
    const int n_tiles = hs.tiles_x * hs.tiles_y * n_img;
    hs.n_full = n_tiles, hs.n_items = n_tiles;
    hs.ncb = ncb;
    // split-K in two when even the 128-pixel items leave half of the SMs idle and K is long enough to share
    // (HaloShape::ksplit; the 512-channel stage of the student encoder on 16 x 16 maps)
    hs.ksplit = 1;
    if (halo_ksplit() && !relu && !scale && hs.item_halves == 1 && hs.cin_chunks % 2 == 0 && hs.cin_chunks >= 8 &&
        2LL * n_tiles * ncb <= (long long)sms + sms / 8)
      hs.ksplit = 2;
    hs.cpk = hs.cin_chunks / hs.ksplit;
    if (halo_mode != 4 && hs.item_halves == 2 && ncb == 1) {
      const int rem = n_tiles % sms;
      if (rem > 0 && 2 * rem <= sms) hs.n_full = n_tiles - rem, hs.n_items = hs.n_full + 2 * rem;
    }
    hs.a_stage = ((hs.tile_rows + 2) * hs.hx * kKc * 4 + 1023) / 1024 * 1024;
    // 64-column tiles: a weight tile lasts ~190 clk and a halo chunk ~1700 clk against ~2000 clk of TMA latency -> three
    // halo stages and a deep weight ring (small-map layers of the student encoder were latency-bound with 2 + 6)
    hs.a_stages = c_out == 64 ? 3 : 2;
    int w_stages = (smem_max - hs.a_stages * hs.a_stage - stage_out) / w_tile;
    const int w_cap = c_out == 64 ? 14 : 6;
    if (w_stages > w_cap) w_stages = w_cap;
    DBEV_CHECK_ARG(w_stages >= 2, "conv2d_tc: halo tile does not fit shared memory");
    hs.w_stages = w_stages;
    hs.c_off = out_c_off, hs.relu = relu, hs.accumulate = (accumulate || hs.ksplit > 1) ? 1 : 0;
    CUtensorMap tmap_x, tmap_w, tmap_o;
    {
      // NHWC output [n, H, W, ld] as (C, W, H, N); a store box = 32 channels x 8 px x 4 rows, clipped at the borders
      cuuint64_t dims[4] = {(cuuint64_t)out_ld, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)n_img};
      cuuint64_t strides[3] = {(cuuint64_t)out_ld * 4, (cuuint64_t)w * out_ld * 4, (cuuint64_t)h * w * out_ld * 4};
      cuuint32_t box[4] = {(cuuint32_t)kKc, (cuuint32_t)kHaloTx, 4, 1};
      cuuint32_t estr[4] = {1, 1, 1, 1};
      CUresult r = encode(&tmap_o, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)out, dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                          CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) {
        set_last_error("conv2d_tc: cuTensorMapEncodeTiled(out) failed (%d)", (int)r);
        return DBEV_ERR_CUDA;
      }
    }
    {
      cuuint64_t dims[4] = {(cuuint64_t)c_in, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)n_img};
      cuuint64_t strides[3] = {(cuuint64_t)x_ld * 4, (cuuint64_t)w * x_ld * 4, (cuuint64_t)h * w * x_ld * 4};
      cuuint32_t box[4] = {(cuuint32_t)kKc, (cuuint32_t)hs.hx, (cuuint32_t)(hs.tile_rows + 2), 1};
      cuuint32_t estr[4] = {1, 1, 1, 1};
      CUresult r = encode(&tmap_x, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)x_nhwc, dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) {
        set_last_error("conv2d_tc: cuTensorMapEncodeTiled(x halo) failed (%d)", (int)r);
        return DBEV_ERR_CUDA;
      }
    }
    {
      const int k_total = 9 * c_in;
      cuuint64_t dims[2] = {(cuuint64_t)k_total, (cuuint64_t)c_out * ncb};
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
    const size_t smem = (size_t)hs.a_stages * hs.a_stage + (size_t)w_stages * w_tile + stage_out + 1024;
    const int grid = hs.n_items * ncb * hs.ksplit < sms ? hs.n_items * ncb * hs.ksplit : sms;
    if (hs.ksplit > 1 && !accumulate)    // the K parts reduce-add into a zeroed output
      DBEV_CUDA(cudaMemset2DAsync(out + out_c_off, (size_t)out_ld * 4, 0, (size_t)c_out * ncb * 4, (size_t)n_img * h * w, stream));
    // DBEV_CONV_PROF=1: per-role wait cycles (debug only: synchronises and prints after every launch)
    static long long* prof_buf = nullptr;
    long long* prof = nullptr;
    static const bool prof_on = getenv("DBEV_CONV_PROF") != nullptr;
    if (prof_on) {
      if (!prof_buf) DBEV_CUDA(cudaMalloc(&prof_buf, sizeof(long long) * 16 * 1024));
      DBEV_CUDA(cudaMemsetAsync(prof_buf, 0, sizeof(long long) * 16 * 1024, stream));
      prof = prof_buf;
    }
#define DBEV_HALO_LAUNCH(CO)                                                                                 \
  do {                                                                                                       \
    DBEV_CUDA(cudaFuncSetAttribute(conv3x3_halo_kernel<CO>, cudaFuncAttributeMaxDynamicSharedMemorySize,     \
                                   (int)smem));                                                              \
    DBEV_CUDA(launch_conv(conv3x3_halo_kernel<CO>, grid, smem, stream, tmap_x, tmap_w, tmap_o, scale, shift, out, hs, prof)); \
  } while (0)
    if (c_out == 64) DBEV_HALO_LAUNCH(64);
    else if (c_out == 128) DBEV_HALO_LAUNCH(128);
    else DBEV_HALO_LAUNCH(256);
#undef DBEV_HALO_LAUNCH
    DBEV_CHECK_LAUNCH("conv3x3_halo_kernel");
    if (prof) {
      static long long hbuf[16 * 1024];
      DBEV_CUDA(cudaStreamSynchronize(stream));
      DBEV_CUDA(cudaMemcpy(hbuf, prof, sizeof(long long) * 16 * grid, cudaMemcpyDeviceToHost));
      double a[16] = {0};
      for (int b = 0; b < grid; ++b)
        for (int j = 0; j < 16; ++j) a[j] += (double)hbuf[b * 16 + j] / grid;
      fprintf(stderr, "halo prof c=%d->%d %dx%d grid=%d tiles=%d | producer total %.0f wait_a %.0f wait_w %.0f | mma total %.0f "
                      "wait_tmem %.0f wait_a %.0f wait_w %.0f | epi total %.0f wait_full %.0f\n",
              c_in, c_out, h, w, grid, hs.n_items, a[0], a[1], a[2], a[4], a[5], a[6], a[7], a[8], a[9]);
    }
    return DBEV_OK;
  }
  ConvShape s;
  s.n_img = n_img, s.c_in = c_in, s.c_out = c_out, s.kh = kh, s.kw = kw, s.stride = stride, s.pad = pad;
  s.ho = force_ho > 0 ? force_ho : (h + 2 * pad - kh) / stride + 1;
  s.wo = force_wo > 0 ? force_wo : (w + 2 * pad - kw) / stride + 1;
  s.accumulate = accumulate ? 1 : 0;
  DBEV_CHECK_ARG(s.ho >= 1 && s.wo >= 1, "conv2d_tc: empty output");
  s.n_cls = n_cls;
  for (int c = 0; c < 4; ++c) {
    const bool used = c < n_cls;
    s.ckh[c] = classes && used ? classes->kh[c] : kh, s.ckw[c] = classes && used ? classes->kw[c] : kw;
    s.coy[c] = classes && used ? classes->add_y[c] : out_add_y, s.cox[c] = classes && used ? classes->add_x[c] : out_add_x;
    DBEV_CHECK_ARG(s.ckh[c] >= 1 && s.ckh[c] <= 3 && s.ckw[c] >= 1 && s.ckw[c] <= 3, "conv2d_tc: class filters are 1..3 wide");
    DBEV_CHECK_ARG((s.ho - 1) * out_mul + s.coy[c] < out_h && (s.wo - 1) * out_mul + s.cox[c] + out_groups - 1 < out_w,
                   "conv2d_tc: output lattice exceeds the output tensor");
  }
  // tile = TX x TY output pixels with TX * TY = 128, TX a power of two <= W_out
  int tx = 128;
  while (tx > s.wo) tx >>= 1;
  if (tx < 8) tx = 8;
  s.tx = tx, s.ty = kPix / tx;
  s.tiles_x = ceil_div(s.wo, s.tx), s.tiles_y = ceil_div(s.ho, s.ty);
  s.n_tiles = s.tiles_x * s.tiles_y * n_img;
  s.ncb = ncb;
  s.cin_chunks = c_in / kKc;
  s.omul = out_mul, s.oadd_y = out_add_y, s.oadd_x = out_add_x, s.h_full = out_h, s.w_full = out_w;
  s.ld = out_ld, s.c_off = out_c_off, s.relu = relu, s.nchw = out_nchw ? 1 : 0;
  DBEV_CHECK_ARG(s.tx * stride <= 256 && s.ty * stride <= 256, "conv2d_tc: tile too large for a TMA box");

  s.group_cols = out_groups > 1 ? c_out / out_groups : 0;
  // NHWC outputs leave through a shared-memory staging tile + TMA store, also on a strided lattice (the four parity
  // classes of a stride-2 input gradient): the lattice of class (a, b) is itself a 4-D tensor (C, W_out, H_out, N) with
  // pixel strides out_mul * ld and base pixel (a, b) - the direct 16-byte stores at pixel stride it replaces touch 32
  // partial sectors per instruction and made those launches epilogue-bound.
  s.tma_store = (out_groups == 1 && !out_nchw && tma_lattice_store()) || (out_groups == 1 && !out_nchw && out_mul == 1 && out_add_y == 0 &&
                out_add_x == 0 && out_h == s.ho && out_w == s.wo) ? 1 : 0;
  CUtensorMap tmap_x, tmap_w, tmap_o;
  tmap_o = CUtensorMap();
  ClassMaps cmaps;
  memset(&cmaps, 0, sizeof(cmaps));
  DBEV_CHECK_ARG(n_cls == 1 || s.tma_store, "conv2d_tc: multi-class launches need the TMA-store epilogue");
  for (int c = 0; c < n_cls && s.tma_store; ++c) {
    // a warp's 32 pixels of the tile: min(TX, 32) px x 32 / min(TX, 32) rows
    const int bx = s.tx < 32 ? s.tx : 32;
    float* obase = out + ((long long)s.coy[c] * out_w + s.cox[c]) * out_ld;
    cuuint64_t dims[4] = {(cuuint64_t)out_ld, (cuuint64_t)s.wo, (cuuint64_t)s.ho, (cuuint64_t)n_img};
    cuuint64_t strides[3] = {(cuuint64_t)out_mul * out_ld * 4, (cuuint64_t)out_mul * out_w * out_ld * 4,
                             (cuuint64_t)out_h * out_w * out_ld * 4};
    cuuint32_t box[4] = {(cuuint32_t)kKc, (cuuint32_t)bx, (cuuint32_t)(32 / bx), 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = encode(c == 0 ? &tmap_o : &cmaps.o[c - 1], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)obase, dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                        CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_last_error("conv2d_tc: cuTensorMapEncodeTiled(out) failed (%d)", (int)r);
      return DBEV_ERR_CUDA;
    }
  }
  {
    // NHWC input as a 4-D tensor (C, W, H, N); box {32, TX*s, TY*s, 1} traversed with element strides
    // {1, s, s, 1} loads TX x TY pixels; out-of-bounds coordinates (the padding) are zero-filled
    cuuint64_t dims[4] = {(cuuint64_t)c_in, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)n_img};
    cuuint64_t strides[3] = {(cuuint64_t)x_ld * 4, (cuuint64_t)w * x_ld * 4, (cuuint64_t)h * w * x_ld * 4};
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
  for (int c = 0; c < n_cls; ++c) {
    const int k_total = s.ckh[c] * s.ckw[c] * c_in;
    const float* wc = classes ? classes->w[c] : w_packed;
    DBEV_CHECK_ARG(((uintptr_t)wc & 15) == 0, "conv2d_tc: weight matrices must be 16-byte aligned");
    cuuint64_t dims[2] = {(cuuint64_t)k_total, (cuuint64_t)c_out * ncb};
    cuuint64_t strides[1] = {(cuuint64_t)k_total * 4};
    cuuint32_t box[2] = {(cuuint32_t)kKc, (cuuint32_t)c_out};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = encode(c == 0 ? &tmap_w : &cmaps.w[c - 1], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)wc, dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_last_error("conv2d_tc: cuTensorMapEncodeTiled(w) failed (%d)", (int)r);
      return DBEV_ERR_CUDA;
    }
  }
#define DBEV_CONV_LAUNCH(CO, STG, MB)                                                            \
  do {                                                                                           \
    const size_t smem = (size_t)STG * (kATile + CO * kKc * 4) + (s.tma_store ? 4 * 8192 : 0) + 1024; \
    const int grid = s.n_tiles * ncb * n_cls < sms * MB ? s.n_tiles * ncb * n_cls : sms * MB;    \
    DBEV_CUDA(cudaFuncSetAttribute(conv2d_tc_kernel<CO, STG, MB>,                                \
                                   cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));     \
    DBEV_CUDA(launch_conv(conv2d_tc_kernel<CO, STG, MB>, grid, smem, stream, tmap_x, tmap_w, tmap_o, scale, shift, out, s, cmaps)); \
  } while (0)
  // two CTAs per SM where shared memory and TMEM allow it: one CTA's epilogue / TMA latency hides
  // behind the other's MMAs (1.68 -> 1.56 ms for the whole SECOND + SECONDFPN stack); three CTAs of
  // the 64-channel kernel or deeper stage rings gave nothing
  // (+ 32 KB of output staging per CTA: 3 x 24 KB / 2 x 32 KB stage rings keep two CTAs resident)
  // direct-store outputs (FPN branches) need no staging and keep the deeper rings
  // grids that do not fill the SMs twice (small BEV maps of the student encoder): one CTA per SM with a ring deep
  // enough to cover the TMA latency (a 64-column stage lasts ~190 clk, a 128-column stage ~260 clk)
  const bool small_grid = (long long)s.n_tiles * ncb * n_cls <= (long long)sms;
  if (c_out == 64) {
    if (small_grid) DBEV_CONV_LAUNCH(64, 7, 1);
    else if (s.tma_store) DBEV_CONV_LAUNCH(64, 3, 2);
    else DBEV_CONV_LAUNCH(64, 4, 2);
  } else if (c_out == 128) {
    if (small_grid) DBEV_CONV_LAUNCH(128, 5, 1);
    else if (s.tma_store) DBEV_CONV_LAUNCH(128, 2, 2);
    else DBEV_CONV_LAUNCH(128, 3, 2);
  } else if (smem_reserve() > 0 && s.tma_store) DBEV_CONV_LAUNCH(256, 3, 1);
  else DBEV_CONV_LAUNCH(256, 4, 1);
#undef DBEV_CONV_LAUNCH
  DBEV_CHECK_LAUNCH("conv2d_tc_kernel");
  return DBEV_OK;
}

}  // namespace

int conv2d_tc_forward_ex(const float* x_nhwc, int n_img, int h, int w, int c_in, int x_ld, const float* w_packed,
                         int c_out, int n_col_blocks, int kh, int kw, int stride, int pad, const float* scale,
                         const float* shift, int relu, float* out, int out_h, int out_w, int out_ld,
                         int out_c_off, int out_mul, int out_add_y, int out_add_x, int out_nchw,
                         int out_groups, int force_ho, int force_wo, int accumulate, cudaStream_t stream) {
  return conv2d_tc_impl(x_nhwc, n_img, h, w, c_in, x_ld, w_packed, c_out, n_col_blocks, kh, kw, stride, pad, scale, shift, relu, out,
                        out_h, out_w, out_ld, out_c_off, out_mul, out_add_y, out_add_x, out_nchw, out_groups, force_ho, force_wo,
                        accumulate, nullptr, stream);
}

// Input gradient of a 3x3 / stride 2 / pad 1 convolution in ONE launch: dx[n, 2*ho, 2*wo, :] (+)= conv_transpose(dy, W).
// The four parity classes (a, b) of the input pixels are stride-1 convolutions of dy with (1+a) x (1+b) taps written on
// the 2x lattice at offset (a, b) (pack_conv_weights mode 2 holds their matrices at float offsets {0, 1, 3, 5} * C_in *
// C_out); a work item is (class, pixel tile, column block), the 4-tap class first. Four separate launches left most SMs
// idle on the 16^2 .. 64^2 maps (each has a quarter of the pixels and 1-4 taps of K).
int conv2d_tc_dgrad_s2(const float* dy_nhwc, int n_img, int ho, int wo, int c_out_fwd, int dy_ld, const float* w_mode2,
                       int c_in_fwd_total, int col_width, int n_col_blocks, float* dx, int dx_ld, int dx_c_off, int accumulate,
                       cudaStream_t stream) {
  DBEV_CHECK_ARG(col_width * n_col_blocks <= c_in_fwd_total, "conv2d_tc_dgrad_s2: column blocks exceed the input channels");
  const long long per = (long long)c_in_fwd_total * c_out_fwd;
  ConvClasses cls;
  cls.n = 4;
  const int order[4] = {3, 1, 2, 0};                 // heaviest class first
  const int off[4] = {0, 1, 3, 5};
  for (int i = 0; i < 4; ++i) {
    const int c = order[i], a = c >> 1, b = c & 1;
    cls.kh[i] = 1 + a, cls.kw[i] = 1 + b, cls.add_y[i] = a, cls.add_x[i] = b;
    // rows of the class matrix = dx channels; this launch covers rows [dx_c_off, dx_c_off + col_width * n_col_blocks)
    cls.w[i] = w_mode2 + off[c] * per + (long long)dx_c_off * (1 + a) * (1 + b) * c_out_fwd;
  }
  return conv2d_tc_impl(dy_nhwc, n_img, ho, wo, c_out_fwd, dy_ld, cls.w[0], col_width, n_col_blocks, cls.kh[0], cls.kw[0], 1, 0, nullptr,
                        nullptr, 0, dx, 2 * ho, 2 * wo, dx_ld, dx_c_off, 2, cls.add_y[0], cls.add_x[0], 0, 1, ho, wo, accumulate,
                        &cls, stream);
}

}  // namespace dbev
