// Device -> host return of the lifted voxel grid, in-frustum rows only.
//
// The reference moves results with a blocking `.cpu()` of the dense tensor (tools/inference_agnostic.py:396).  The
// lifted grid of the global branch is 42 % exact zeros at the KITTI geometry (voxels outside the camera frustum, the
// `valid` mask of snvc_frustum_lift_fwd), and the 598 MB per batch of 8 pairs is what bounds the end-to-end rate
// (PCIe, 55.8 GB/s measured).  This kernel writes the voxel rows STRAIGHT INTO THE CALLER'S PINNED HOST BUFFER at
// their dense positions (pinned memory is device-addressable under UVA), and only the rows that can differ from
// what the buffer already holds:
//     valid[r]                      -> the row's `row_bytes` bytes are copied;
//     !valid[r] && prev_valid[r]    -> the row is zero-filled (it held data from the previous batch in this buffer);
//     !valid[r] && !prev_valid[r]   -> nothing moves: the buffer already holds zeros there.
// prev_valid (device, one byte per row, per host buffer) is updated in place; a first use of a host buffer passes
// prev_valid = all ones, which writes every row.  After the launch the host buffer equals the dense tensor bit for bit
// (tests/test_gpu_host_return.py).  A lane group of row_bytes/16 threads moves one row with 16-byte accesses, so the
// PCIe writes are full 64-byte (C = 32 bf16) segments, contiguous across the in-frustum span of a voxel line.
#include "common.cuh"

namespace snvc {
namespace {

template <int LPR>   // lanes per row: row_bytes == 16 * LPR
__global__ void __launch_bounds__(256)
masked_rows_to_host_kernel(const uint4* __restrict__ src, const uint8_t* __restrict__ valid, uint8_t* __restrict__ prev_valid,
                           uint4* __restrict__ dst, int64_t rows, unsigned long long* __restrict__ moved) {
  constexpr int RPT = 4;                                   // rows in flight per lane group (independent loads)
  const int64_t groups_per_block = 256 / LPR;
  const int g = threadIdx.x / LPR, l = threadIdx.x % LPR;
  unsigned long long mine = 0;
  for (int64_t r0 = ((int64_t)blockIdx.x * groups_per_block + g) * RPT; r0 < rows; r0 += (int64_t)gridDim.x * groups_per_block * RPT) {
    uint4 v[RPT];
    uint8_t va[RPT], pv[RPT];
#pragma unroll
    for (int j = 0; j < RPT; ++j) {
      const int64_t r = r0 + j;
      va[j] = r < rows ? valid[r] : 0;
      pv[j] = r < rows ? prev_valid[r] : 0;
      v[j] = make_uint4(0u, 0u, 0u, 0u);
      if (va[j]) v[j] = __ldcs(src + r * LPR + l);
    }
#pragma unroll
    for (int j = 0; j < RPT; ++j) {
      const int64_t r = r0 + j;
      if (va[j] | pv[j]) {
        dst[r * LPR + l] = v[j];                            // host-mapped address: posted PCIe write
        mine += 16;
        if (l == 0 && va[j] != pv[j]) prev_valid[r] = va[j];
      }
    }
  }
  if (moved) {
    for (int o = 16; o; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o);
    if ((threadIdx.x & 31) == 0 && mine) atomicAdd(moved, mine);
  }
}

}  // namespace
}  // namespace snvc

using namespace snvc;

extern "C" int snvc_masked_rows_to_host(const void* src, const uint8_t* valid, uint8_t* prev_valid, void* dst_host_mapped,
                                        int64_t rows, int32_t row_bytes, int32_t max_blocks, unsigned long long* moved_bytes,
                                        void* stream) {
  if (rows == 0) return 0;
  SNVC_CHECK_ARG(src && valid && prev_valid && dst_host_mapped, "null pointer");
  SNVC_CHECK_ARG(row_bytes == 64 || row_bytes == 128 || row_bytes == 32 || row_bytes == 16, "row_bytes must be 16, 32, 64 or 128");
  SNVC_CHECK_ARG(((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst_host_mapped)) & 15) == 0,
                 "src and dst must be 16-byte aligned");
  // the destination must be device-addressable host memory (cudaHostAlloc / cudaHostRegister under UVA)
  cudaPointerAttributes at;
  SNVC_CUDA_OK(cudaPointerGetAttributes(&at, dst_host_mapped));
  SNVC_CHECK_ARG(at.type == cudaMemoryTypeHost && at.devicePointer != nullptr,
                 "dst_host_mapped must be pinned (page-locked) host memory");
  const int lpr = row_bytes / 16;
  const int64_t groups_per_block = 256 / lpr;
  int64_t want = ceil_div(rows, groups_per_block * 4);
  int blocks = (int)std::min<int64_t>(want, max_blocks > 0 ? max_blocks : 2 * (int64_t)sm_count());
  if (blocks < 1) blocks = 1;
  auto* s = static_cast<const uint4*>(src);
  auto* d = static_cast<uint4*>(at.devicePointer);
  cudaStream_t st = (cudaStream_t)stream;
  switch (lpr) {
    case 1: masked_rows_to_host_kernel<1><<<blocks, 256, 0, st>>>(s, valid, prev_valid, d, rows, moved_bytes); break;
    case 2: masked_rows_to_host_kernel<2><<<blocks, 256, 0, st>>>(s, valid, prev_valid, d, rows, moved_bytes); break;
    case 4: masked_rows_to_host_kernel<4><<<blocks, 256, 0, st>>>(s, valid, prev_valid, d, rows, moved_bytes); break;
    default: masked_rows_to_host_kernel<8><<<blocks, 256, 0, st>>>(s, valid, prev_valid, d, rows, moved_bytes); break;
  }
  return launch_status("masked_rows_to_host_kernel");
}
