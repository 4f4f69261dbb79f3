// Microbenchmark: issue-to-completion cost of tcgen05.mma (kind::f16, bf16 x bf16 -> fp32, cta_group::1, M = 128,
// K = 16, both operands in shared memory) as a function of N and of the operand swizzle.  One CTA, one issuing thread,
// ITER back-to-back accumulating MMAs on fixed operands, timed with clock64 around issue ... commit ... mbarrier wait.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_rate umma_rate.cu && ./umma_rate
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, int sw) {
  const uint64_t layout = sw == 128 ? 2ull : (sw == 64 ? 4ull : 6ull);
  const uint64_t sbo = (uint64_t)(8 * sw) >> 4;
  return (uint64_t)((addr >> 4) & 0x3FFFu) | (1ull << 16) | (sbo << 32) | (1ull << 46) | (layout << 61);
}
__global__ void __launch_bounds__(128, 1) k(int N, int sw, int iters, int nctas_dummy, long long* out, int a_stride_rows, int two_cta_sim) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tbase;
  const uint32_t base = (smem_u32(smem) + 1023u) & ~1023u;
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) ((uint32_t*)(smem + (base - smem_u32(smem))))[i] = 0;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tbase)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tm = tbase;
  if (threadIdx.x == 0) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint64_t ad0 = make_desc(base, sw), bd0 = make_desc(base + 24 * 1024, sw);
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      // A window shifted by a few rows per MMA (like the conv taps), B tile rotates over 8 tiles
      const uint64_t ad = ad0 + (uint64_t)(((i % 9) * a_stride_rows * sw) >> 4) + 2 * (i & 1);
      const uint64_t bd = bd0 + (uint64_t)(((i % 8) * 256) >> 4) + 2 * (i & 1);
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                   ::"r"(tm), "l"(ad), "l"(bd), "r"(idesc) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    uint32_t done = 0;
    while (!done)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(&bar)) : "memory");
    long long t1 = clock64();
    out[blockIdx.x] = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tm));
}
int main() {
  long long* d; cudaMalloc(&d, 8 * 148);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  const int iters = 4096;
  printf("cycles per tcgen05.mma M=128 K=16 (SS), %d back-to-back, 1 CTA / SM x grid\n", iters);
  for (int grid : {1, 148})
    for (int sw : {64, 128})
      for (int N : {16, 32, 48, 64, 96, 128, 144, 160, 192, 256}) {
        k<<<grid, 128, 100 * 1024>>>(N, sw, iters, 0, d, 3, 0);
        k<<<grid, 128, 100 * 1024>>>(N, sw, iters, 0, d, 3, 0);
        cudaError_t e = cudaDeviceSynchronize();
        long long h[148]; cudaMemcpy(h, d, 8 * grid, cudaMemcpyDeviceToHost);
        long long mx = 0; for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
        printf("grid %3d  swizzle %3d  N %3d : %7.1f cycles/MMA  (%s) -> %5.1f %% of 4096 MAC/clk/SM\n", grid, sw, N, (double)mx / iters,
               cudaGetErrorString(e), 100.0 * (128.0 * N * 16) / ((double)mx / iters) / 4096.0);
      }
  return 0;
}
