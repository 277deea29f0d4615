// Shared helpers for the distill-bev B200 kernels (sm_100a only).
//
// Every entry point of the C-ABI (include/distill_bev_b200.h) returns an int
// status and records a message retrievable through dbev_last_error(); no
// entry point synchronises the device or touches a stream other than the one
// it is handed (SURVEY.md §8b "Threading / streams").
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define DBEV_OK 0
#define DBEV_ERR_INVALID 1
#define DBEV_ERR_CUDA 2
#define DBEV_ERR_WORKSPACE 3

namespace dbev {

void set_last_error(const char* fmt, ...);

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs

inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

inline size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

// Bump allocator over a caller-provided workspace; all slices 256 B aligned.
struct Workspace {
  char* base;
  size_t size;
  size_t used;
  Workspace(void* p, size_t bytes) : base((char*)p), size(bytes), used(0) {}
  template <typename T>
  T* take(size_t count) {
    size_t off = align_up(used);
    size_t end = off + count * sizeof(T);
    used = end;
    if (base == nullptr || end > size) return nullptr;
    return (T*)(base + off);
  }
  bool ok() const { return base != nullptr && used <= size; }
};

// Exact unsigned division by a runtime constant for numerators n < 2^31
// (Granlund-Montgomery round-up method): with l = ceil(log2 d), shift = 31 + l and
// mul = ceil(2^shift / d) < 2^32, floor(n / d) == (n * mul) >> shift for every n < 2^31
// because mul * d - 2^shift <= 2^l. One IMAD.WIDE + one funnel shift instead of ~25
// instructions for a hardware-less 32-bit division.
struct FastDiv {
  unsigned mul;
  int shift;
  unsigned d;
  FastDiv() : mul(1u << 31), shift(31), d(1) {}
  explicit FastDiv(unsigned div) : d(div) {
    int l = 0;
    while ((1ull << l) < div) ++l;
    shift = 31 + l;
    mul = (unsigned)(((1ull << shift) + div - 1) / div);
  }
  __host__ __device__ __forceinline__ unsigned div(unsigned n) const {
    return (unsigned)(((unsigned long long)n * mul) >> shift);
  }
  __host__ __device__ __forceinline__ unsigned mod(unsigned n) const { return n - div(n) * d; }
};

__device__ __forceinline__ unsigned lane_id() { return threadIdx.x & 31; }

__device__ __forceinline__ unsigned lanemask_lt() {
  unsigned m;
  asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
  return m;
}

// Streaming 128-bit load: data is touched once, keep it out of L1.
__device__ __forceinline__ float4 ld_stream_f4(const float* p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p));
  return v;
}

__device__ __forceinline__ void st_stream_f4(float* p, const float4& v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p),
               "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w));
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

}  // namespace dbev

#define DBEV_CHECK_ARG(cond, ...)        \
  do {                                   \
    if (!(cond)) {                       \
      dbev::set_last_error(__VA_ARGS__); \
      return DBEV_ERR_INVALID;           \
    }                                    \
  } while (0)

#define DBEV_CHECK_LAUNCH(name)                                                   \
  do {                                                                            \
    cudaError_t e__ = cudaGetLastError();                                         \
    if (e__ != cudaSuccess) {                                                     \
      dbev::set_last_error("%s: CUDA launch failed: %s", name, cudaGetErrorString(e__)); \
      return DBEV_ERR_CUDA;                                                       \
    }                                                                             \
  } while (0)

#define DBEV_CUDA(call)                                                          \
  do {                                                                           \
    cudaError_t e__ = (call);                                                    \
    if (e__ != cudaSuccess) {                                                    \
      dbev::set_last_error("%s failed: %s", #call, cudaGetErrorString(e__));      \
      return DBEV_ERR_CUDA;                                                      \
    }                                                                            \
  } while (0)
