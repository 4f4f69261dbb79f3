// Microbenchmark / semantics probe: tcgen05.mma.cta_group::2 (M = 256 over a CTA pair, bf16 x bf16 -> fp32, K = 16,
// both operands in shared memory).  Checks on the device what the conv kernels need to know before pairing CTAs:
//   * each CTA supplies its own 128 A rows and HALF of the B rows (CTA rank r holds B rows [r*N/2, (r+1)*N/2)) at
//     the SAME shared-memory offset; D rows 0..127 land in rank 0's TMEM, rows 128..255 in rank 1's, same columns;
//   * one warp per CTA allocates with cta_group::2; the leader's elected thread issues the MMA and a multicast commit;
//   * cycles per MMA as a function of N (the per-SM B traffic is halved).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_2cta umma_2cta.cu && ./umma_2cta
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdint>
#include <cstdio>
#include <vector>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, int sw) {
  const uint64_t layout = sw == 128 ? 2ull : (sw == 64 ? 4ull : 6ull);
  const uint64_t sbo = (uint64_t)(8 * sw) >> 4;
  return (uint64_t)((addr >> 4) & 0x3FFFu) | (1ull << 16) | (sbo << 32) | (1ull << 46) | (layout << 61);
}
__device__ __forceinline__ uint32_t swz64(uint32_t row, uint32_t chunk) {
  const uint32_t a = row * 64u + chunk * 16u;
  return a ^ (((a >> 7) & 3u) << 4);
}
__host__ __device__ inline int aval(int cta, int r, int k) { return ((r * 3 + k * 5 + cta * 7) % 5) - 2; }
__host__ __device__ inline int bval(int n, int k) { return ((n * 7 + k * 3) % 7) - 3; }
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
k2(int N, int iters, float* dout, long long* cyc, int sw = 64, int kslices = 1, int shift_rows = 0) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tbase;
  uint32_t rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  const uint32_t base = (smem_u32(smem) + 1023u) & ~1023u;
  uint8_t* sp = smem + (base - smem_u32(smem));
  uint8_t* sa = sp;                 // A: 128 rows x 64 B (K = 32, only the first K step is used)
  uint8_t* sb = sp + 16 * 1024;     // B: N/2 rows x 64 B
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) ((uint32_t*)sp)[i] = 0;
  __syncthreads();
  for (int i = threadIdx.x; i < 128 * 32; i += blockDim.x) {
    const int r = i / 32, kk = i % 32;
    *(__nv_bfloat16*)(sa + swz64(r, kk / 8) + (kk % 8) * 2) = __float2bfloat16((float)aval(rank, r, kk));
  }
  for (int i = threadIdx.x; i < (N / 2) * 32; i += blockDim.x) {
    const int r = i / 32, kk = i % 32;
    *(__nv_bfloat16*)(sb + swz64(r, kk / 8) + (kk % 8) * 2) = __float2bfloat16((float)bval((int)rank * (N / 2) + r, kk));
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tbase)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  cluster_sync();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tm = tbase;
  long long t0 = 0, t1 = 0;
  if (rank == 0 && threadIdx.x == 0) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
    const uint64_t ad0 = make_desc(smem_u32(sa), sw), bd0 = make_desc(smem_u32(sb), sw);
    t0 = clock64();
    // rate runs: rotate over the K slices of a row (2 x 16-byte units each) and shift the A window by rows, like conv
    // taps (incremental counters: divisions in the issuing thread would make the loop issue-bound)
    uint32_t ks = 0, tap_off = 0, tap = 0;
    const uint32_t tap_step = (uint32_t)(shift_rows * sw) >> 4, ks_end = 2u * (uint32_t)kslices;
    for (int i = 0; i < iters; ++i) {
      const uint64_t ad = ad0 + ks + tap_off;
      const uint64_t bd = bd0 + ks;
      ks += 2u;
      if (ks == ks_end) { ks = 0; tap_off += tap_step; if (++tap == 9u) { tap = 0; tap_off = 0; } }
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                   ::"r"(tm), "l"(ad), "l"(bd), "r"(idesc), "r"(i) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(&bar)), "h"((uint16_t)3) : "memory");
  }
  // every thread of both CTAs waits on its own CTA's barrier (the commit is multicast to both)
  {
    uint32_t done = 0;
    long long spin = 0;
    while (!done) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(&bar)) : "memory");
      if (++spin > 200000000ll) { if (threadIdx.x == 0) printf("timeout rank %u\n", rank); __trap(); }
    }
  }
  if (rank == 0 && threadIdx.x == 0) { t1 = clock64(); cyc[blockIdx.x / 2] = t1 - t0; }
  asm volatile("tcgen05.fence::after_thread_sync;");
  // D: this CTA's 128 rows, N columns; with `iters` accumulating MMAs of identical operands D = iters * A*B^T
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (dout) {
    for (int c0 = 0; c0 < N; c0 += 16) {
      uint32_t v[16];
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                   : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                     "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                   : "r"(tm + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0) : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      for (int j = 0; j < 16; ++j)
        dout[((size_t)(blockIdx.x) * 128 + warp * 32 + lane) * 256 + c0 + j] = __uint_as_float(v[j]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  cluster_sync();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tm));
}
int main() {
  float* d; long long* c;
  cudaMalloc(&d, sizeof(float) * 2 * 128 * 256); cudaMalloc(&c, 8 * 148);
  cudaFuncSetAttribute(k2, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  // ---- semantics: one pair, iters MMAs (first overwrites, the rest accumulate)
  for (int N : {32, 96, 144, 192, 256}) {
    const int iters = 3;
    cudaMemset(d, 0, sizeof(float) * 2 * 128 * 256);
    k2<<<2, 128, 100 * 1024>>>(N, iters, d, c);
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<float> h(2 * 128 * 256);
    cudaMemcpy(h.data(), d, sizeof(float) * h.size(), cudaMemcpyDeviceToHost);
    long bad = 0; int fr = -1, fn = -1; float got = 0, want = 0;
    for (int cta = 0; cta < 2; ++cta)
      for (int r = 0; r < 128; ++r)
        for (int n = 0; n < N; ++n) {
          int s = 0;
          for (int kk = 0; kk < 16; ++kk) s += aval(cta, r, kk) * bval(n, kk);
          const float w = (float)(s * iters), g = h[((size_t)cta * 128 + r) * 256 + n];
          if (g != w) { if (!bad) { fr = cta * 128 + r; fn = n; got = g; want = w; } ++bad; }
        }
    printf("semantics N %3d: %s, mismatches %ld", N, cudaGetErrorString(e), bad);
    if (bad) printf(" (first at row %d col %d: got %g want %g)", fr, fn, got, want);
    printf("\n");
    if (e != cudaSuccess) return 1;
  }
  // ---- rate: all pairs
  const int iters = 4096;
  for (int grid : {2, 148})
    for (int N : {64, 96, 128, 144, 192, 256}) {
      k2<<<grid, 128, 100 * 1024>>>(N, iters, nullptr, c);
      k2<<<grid, 128, 100 * 1024>>>(N, iters, nullptr, c);
      cudaError_t e = cudaDeviceSynchronize();
      long long h[74]; cudaMemcpy(h, c, 8 * (grid / 2), cudaMemcpyDeviceToHost);
      long long mx = 0; for (int i = 0; i < grid / 2; ++i) mx = h[i] > mx ? h[i] : mx;
      printf("grid %3d  N %3d : %7.1f cycles/MMA (M=256 over the pair)  (%s) -> %5.1f %% of 4096 MAC/clk/SM\n", grid, N,
             (double)mx / iters, cudaGetErrorString(e), 100.0 * (128.0 * N * 16) / ((double)mx / iters) / 4096.0);
    }
  // ---- rate vs operand layout: swizzle mode (row bytes), K slices per row used in rotation, row-shifted A windows
  for (int sw : {32, 64, 128})
    for (int shift : {0, 3})
      for (int N : {96, 192}) {
        const int ksl = sw / 32;
        k2<<<148, 128, 100 * 1024>>>(N, iters, nullptr, c, sw, ksl, shift);
        k2<<<148, 128, 100 * 1024>>>(N, iters, nullptr, c, sw, ksl, shift);
        cudaError_t e = cudaDeviceSynchronize();
        long long h[74]; cudaMemcpy(h, c, 8 * 74, cudaMemcpyDeviceToHost);
        long long mx = 0; for (int i = 0; i < 74; ++i) mx = h[i] > mx ? h[i] : mx;
        printf("layout: swizzle %3d  K slices %d  A row shift %d  N %3d : %7.1f cycles/MMA  (%s)\n", sw, ksl, shift, N,
               (double)mx / iters, cudaGetErrorString(e));
      }
  return 0;
}
