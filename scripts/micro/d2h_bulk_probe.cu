// Probe: device -> pinned-host return of variable-length contiguous segments (the in-frustum span of every voxel line),
//   (a) lanes storing 16 bytes each straight to the host mapping (what snvc_masked_rows_to_host does), against
//   (b) TMA bulk copies: global -> shared (cp.async.bulk + mbarrier), shared -> host (cp.async.bulk.global.shared::cta),
// to see whether bulk stores reach the copy engine's PCIe rate (57 GB/s dense) instead of ~47 GB/s.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/d2h_bulk scripts/micro/d2h_bulk_probe.cu && /tmp/d2h_bulk
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s failed: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)
struct Seg { uint64_t off; uint32_t bytes; uint32_t pad; };

__global__ void lane_copy(const uint4* __restrict__ src, uint4* __restrict__ dst, const Seg* __restrict__ segs, int nseg) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31, nw = (gridDim.x * blockDim.x) >> 5;
  for (int s = warp; s < nseg; s += nw) {
    const uint64_t o = segs[s].off >> 4; const uint32_t n = segs[s].bytes >> 4;
    for (uint32_t i = lane; i < n; i += 32) dst[o + i] = __ldcs(src + o + i);
  }
}

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0; long long t0 = clock64();
  while (!done) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    if (clock64() - t0 > 4000000000ll) __trap();
  }
}
constexpr int NBUF = 4, BUFB = 16384;
__global__ void __launch_bounds__(32) bulk_copy(const char* __restrict__ src, char* __restrict__ dst, const Seg* __restrict__ segs, int nseg) {
  extern __shared__ __align__(128) char smem[];
  __shared__ __align__(8) uint64_t bar[NBUF];
  if (threadIdx.x != 0) return;
  for (int b = 0; b < NBUF; ++b) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar[b])) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  uint32_t it = 0;
  for (int s = blockIdx.x; s < nseg; s += gridDim.x) {
    uint64_t o = segs[s].off; uint32_t left = segs[s].bytes;
    while (left) {
      const uint32_t n = left < (uint32_t)BUFB ? left : (uint32_t)BUFB;
      const uint32_t b = it % NBUF, ph = (it / NBUF) & 1u;
      if (it >= NBUF) asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(NBUF - 1) : "memory");   // buffer b's previous store has read it
      const uint32_t sb = s32(smem + b * BUFB), mb = s32(&bar[b]);
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"(n) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(sb), "l"(src + o), "r"(n), "r"(mb) : "memory");
      mbar_wait(mb, ph);
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + o), "r"(sb), "r"(n) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      o += n; left -= n; ++it;
    }
  }
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

int main() {
  const int B = 8, Z = 192, Y = 20, X = 304, ROW = 64;
  const size_t total = (size_t)B * Z * Y * X * ROW;
  char *d, *h;
  CK(cudaMalloc(&d, total)); CK(cudaHostAlloc(&h, total, cudaHostAllocDefault));
  CK(cudaMemset(d, 0x5a, total)); memset(h, 0, total);
  std::vector<Seg> segs; size_t bytes = 0;
  for (int n = 0; n < B; ++n) for (int z = 0; z < Z; ++z) {
    int hw = 10 + (int)(z * 0.83f); if (hw > X / 2) hw = X / 2;              // half-width of the frustum at this depth
    int ny = 5 + z / 6; if (ny > Y) ny = Y;
    for (int y = (Y - ny) / 2; y < (Y - ny) / 2 + ny; ++y) {
      Seg s; s.off = ((((size_t)n * Z + z) * Y + y) * X + (X / 2 - hw)) * ROW; s.bytes = 2 * hw * ROW; s.pad = 0; segs.push_back(s); bytes += s.bytes;
    }
  }
  printf("%zu segments, %.1f MB of %.1f MB (%.3f)\n", segs.size(), bytes / 1e6, total / 1e6, (double)bytes / total);
  Seg* ds; CK(cudaMalloc(&ds, segs.size() * sizeof(Seg))); CK(cudaMemcpy(ds, segs.data(), segs.size() * sizeof(Seg), cudaMemcpyHostToDevice));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1); float ms;
  CK(cudaMemcpy(h, d, total, cudaMemcpyDeviceToHost));
  cudaEventRecord(e0); for (int i = 0; i < 3; ++i) cudaMemcpyAsync(h, d, total, cudaMemcpyDeviceToHost); cudaEventRecord(e1); CK(cudaDeviceSynchronize());
  cudaEventElapsedTime(&ms, e0, e1); printf("dense DMA: %.2f ms = %.1f GB/s\n", ms / 3, total / 1e6 / (ms / 3));
  for (int blocks : {16, 32, 64, 148}) {
    lane_copy<<<blocks, 256>>>((const uint4*)d, (uint4*)h, ds, (int)segs.size()); CK(cudaDeviceSynchronize());
    cudaEventRecord(e0); for (int i = 0; i < 3; ++i) lane_copy<<<blocks, 256>>>((const uint4*)d, (uint4*)h, ds, (int)segs.size()); cudaEventRecord(e1); CK(cudaDeviceSynchronize());
    cudaEventElapsedTime(&ms, e0, e1); printf("lane stores, %3d blocks: %.2f ms = %.1f GB/s\n", blocks, ms / 3, bytes / 1e6 / (ms / 3));
  }
  CK(cudaFuncSetAttribute(bulk_copy, cudaFuncAttributeMaxDynamicSharedMemorySize, NBUF * BUFB));
  for (int blocks : {8, 16, 32, 64, 148, 296}) {
    memset(h, 0, 1 << 20);
    bulk_copy<<<blocks, 32, NBUF * BUFB>>>(d, h, ds, (int)segs.size()); CK(cudaDeviceSynchronize());
    cudaEventRecord(e0); for (int i = 0; i < 3; ++i) bulk_copy<<<blocks, 32, NBUF * BUFB>>>(d, h, ds, (int)segs.size()); cudaEventRecord(e1); CK(cudaDeviceSynchronize());
    cudaEventElapsedTime(&ms, e0, e1); printf("TMA bulk,    %3d blocks: %.2f ms = %.1f GB/s\n", blocks, ms / 3, bytes / 1e6 / (ms / 3));
  }
  size_t bad = 0; for (auto& s : segs) for (uint32_t i = 0; i < s.bytes; i += 4096) if ((unsigned char)h[s.off + i] != 0x5a) ++bad;
  printf("check: %zu bad samples\n", bad);
  return 0;
}
