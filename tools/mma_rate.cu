// tcgen05.mma kind::tf32 SS issue/throughput microbenchmark: clk per MMA for M=128, N in {64,128,256},
// operands resident in shared memory (no TMA traffic), SBO 1024 vs 1280 (the halo kernel's window pitch).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -I distill-bev_b200/csrc -o tools/_bin/mma_rate tools/mma_rate.cu -lcuda
#include <cstdio>
#include <cstdlib>
#include "umma.cuh"
using namespace dbev;

template <int N>
__global__ void __launch_bounds__(64, 1) rate_kernel(int iters, int sbo, int n_per_commit, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* base = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) ((float*)base)[i] = 0.f;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_addr(&tmem_base_s)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_s;
  if (warp == 1) {
    const bool leader = elect_one();
    const uint32_t idesc = umma_idesc_tf32(128, N);
    const uint64_t ad0 = umma_desc(smem_addr(base), 16, (uint32_t)sbo);
    const uint64_t bd0 = umma_desc(smem_addr(base) + 48 * 1024, 16, 1024);
    uint32_t phase = 0;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      if (leader) {
        for (int j = 0; j < n_per_commit; j += 4) {
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            umma_tf32(tmem_base + (j & 4 ? N : 0) % 512, ad0 + (uint64_t)(kk * 2 + (j & 8) * 8), bd0 + (uint64_t)(kk * 2), idesc, 1u);
        }
        umma_commit(&bar);
      }
      __syncwarp();
      mbar_wait(&bar, phase);
      phase ^= 1u;
    }
    long long t1 = clock64();
    if (leader) out[blockIdx.x] = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
}

template <int N>
void run(int sbo, int npc, int grid) {
  long long* d;
  cudaMalloc(&d, sizeof(long long) * grid);
  const int iters = 200;
  const size_t smem = 97 * 1024 + 1024;
  cudaFuncSetAttribute(rate_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  rate_kernel<N><<<grid, 64, smem>>>(iters, sbo, npc, d);
  rate_kernel<N><<<grid, 64, smem>>>(iters, sbo, npc, d);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[256];
  cudaMemcpy(h, d, sizeof(long long) * grid, cudaMemcpyDeviceToHost);
  long long mx = 0;
  for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
  printf("N=%d sbo=%d mma_per_commit=%d grid=%d: %.1f clk/MMA (ideal %d) %s\n", N, sbo, npc, grid,
         (double)mx / ((double)iters * npc), N / 2, cudaGetErrorString(e));
  cudaFree(d);
}

int main() {
  for (int grid : {1, 148}) {
    for (int sbo : {1024, 1280}) {
      run<64>(sbo, 8, grid);
      run<64>(sbo, 32, grid);
      run<128>(sbo, 8, grid);
      run<128>(sbo, 32, grid);
      run<256>(sbo, 8, grid);
      run<256>(sbo, 32, grid);
    }
  }
  return 0;
}
