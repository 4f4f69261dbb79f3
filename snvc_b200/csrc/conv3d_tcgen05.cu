// A2 -- 3-D convolution / transposed convolution as a tcgen05 + TMEM implicit GEMM (sm_100a).
//
// Replaces the cuDNN calls behind nn.Conv3d / nn.ConvTranspose3d (+ eval BatchNorm3d, ReLU,
// residual adds, sigmoid) in snvc/models/submodule.py:32-50,85-168,170-268 and
// snvc/models/vernier.py:250-289,414-438.
//
// GEMM view, per filter tap t = (kd,kh,kw):   D[128 voxels, Cout] += A_t[128 voxels, Cin] * W_t[Cin, Cout]
//   * activations are NDHWC bf16; one 5-D TMA box (Cin x TW x TH x TD x 1) per tap lands the
//     128 x Cin operand tile in shared memory already in the canonical K-major swizzled UMMA
//     layout (box rows = voxels, row = Cin bf16 = 64 or 128 B -> SWIZZLE_64B / SWIZZLE_128B);
//     TMA out-of-bounds zero fill IS the convolution padding, TMA elementStrides IS the stride;
//   * W_t (Cout x Cin, K-major) is a 2-D TMA box from the tap-major packed weights;
//   * one elected thread issues tcgen05.mma (M=128, N=Cout, K=16) Cin/16 times per tap, all taps
//     accumulate in TMEM (fp32); two TMEM accumulators so the epilogue of tile i overlaps the MMAs
//     of tile i+1;
//   * 4 epilogue warps read TMEM (tcgen05.ld 32x32b), apply folded-BN scale/bias, residual, ReLU,
//     sigmoid in fp32 and store bf16 (or fp32) NDHWC rows with 128-bit stores.
// Transposed conv (k3,s2,p1,op1) is decomposed into its 8 output-parity classes (1,2,2,2,4,4,4,8
// taps) -- no zero insertion; each class is the same kernel with a different tap table and an
// output stride of 2.
// Warp roles: warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2..5 = epilogue.
// Persistent: grid = min(#tiles, #SMs), static round-robin tile schedule.
//
// The text above describes the first (per-tap) kernel, still the fallback for 1^3, stride-2 Cin = 64 and odd shapes.
// The layers that carry the FLOPs run on the plane-march kernels further down, each introduced by its own header:
//   v2 conv3d_halo_kernel     one TMA load per input plane, taps = shifted windows of the same tile
//   v3 conv3d_kdfuse_kernel   + the 3 depth taps fused into N (TMEM accumulator ring)
//   v4 conv3d_deconv_kernel   transposed conv, 8 parity classes in one march, staged TMA epilogue
//   v5 conv3d_s2_kernel       stride 2 on "pair rows"
//   v6 conv3d_bigk_kernel     5^3 / 7^3 with streamed weights
//   v7 conv3d_kwfuse_kernel   + the 3 kw taps fused into N, row shift undone by shuffles in the epilogue
//   v8 conv3d_kdpair_kernel   v3 over a CTA pair (tcgen05.mma.cta_group::2): the default for 32-channel output slices
// snvc_conv3d_fwd (bottom of the file) picks the most specific eligible kernel.
#include <cuda.h>

#include <algorithm>
#include <cstdlib>
#include <mutex>
#include <tuple>
#include <vector>

#include "common.cuh"
#include "tcgen05.cuh"

namespace snvc {
namespace {

constexpr int kThreads = 192;
constexpr int kMaxStages = 8;
constexpr int kTileM = 128;

struct ConvParams {
  int N, Cin, Cout, CoutPad;
  int Do, Ho, Wo;              // output tensor extent
  int Dj, Hj, Wj;              // iteration space of this launch
  int TD, TH, TW;              // tile, TD*TH*TW == 128
  int tiles_d, tiles_h, tiles_w;
  int num_tiles;
  int in_stride;               // input coord = j*in_stride + off[tap]
  int out_stride, out_off_d, out_off_h, out_off_w;   // output coord = j*out_stride + out_off
  int K;                       // cubic kernel size (weight slot = (kd*K + kh)*K + kw)
  int nk[3];                   // taps per dim (d,h,w) in this launch
  signed char off[3][8];       // input offset per dim-tap
  signed char kid[3][8];       // kernel index per dim-tap
  int stages, a_bytes, b_bytes;
  int swizzle_bytes;           // 32 / 64 / 128 == Cin*2
  int relu, residual_mode, sigmoid, out_f32;
  int out_cstride, out_coffset, res_cstride, res_coffset;
  const float* scale;
  const float* bias;
  const __nv_bfloat16* residual;
  void* y;
};


struct EpiParams {
  int Cout, CoutPad;
  int relu, residual_mode, sigmoid, out_f32;
  int out_cstride, out_coffset, res_cstride, res_coffset;
  const __nv_bfloat16* residual;
  void* y;
};

// 256-bit global accesses (sm_100: LDG/STG.E.ENL2.256).  A warp's epilogue access touches 32 different voxel rows,
// so its L1 data-pipe cost is one wavefront per row per instruction whatever the width: 32 bytes per instruction
// halve the LSU wavefronts of the 128-bit form.  ncu on the kw-fused kernel: the L1 data pipe, which also feeds the
// UMMA operand reads, is the binding unit (tensor-core reads 53 % + LSU 33 % of its cycles).
__device__ __forceinline__ void stg256(void* p, uint4 a, uint4 b) {
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w),
               "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w)
               : "memory");
}
__device__ __forceinline__ void ldg256(const void* p, uint4& a, uint4& b) {
  asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w)
               : "l"(p));
}
__device__ __forceinline__ bool aligned32(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 31) == 0; }

// Residual row prefetch: issued BEFORE waiting for the accumulator so the global-load latency
// overlaps the MMAs of this tile (up to 64 channels = 8 x 16 B per row).
struct ResidualRow { uint4 q[8]; };
__device__ __forceinline__ void residual_prefetch(const EpiParams& e, bool in_range, int64_t vox, ResidualRow& rr) {
#pragma unroll
  for (int i = 0; i < 8; ++i) rr.q[i] = make_uint4(0u, 0u, 0u, 0u);
  if (!e.residual_mode || !in_range) return;
  const __nv_bfloat16* rp = e.residual + vox * e.res_cstride + e.res_coffset;
  // host guarantees: Cout % 16 == 0 and 16-byte aligned rows whenever a residual is given
  if (aligned32(rp)) {                                   // (Cout % 16 == 0: whole 32-byte pieces)
#pragma unroll
    for (int i = 0; i < 8; i += 2)
      if (i * 8 < e.Cout) ldg256(reinterpret_cast<const uint4*>(rp) + i, rr.q[i], rr.q[i + 1]);
    return;
  }
#pragma unroll
  for (int i = 0; i < 8; ++i)
    if (i * 8 < e.Cout) rr.q[i] = __ldg(reinterpret_cast<const uint4*>(rp) + i);
}

// ---- lean epilogue: the common case (bf16 out, Cout == CoutPad, 16-byte aligned channel rows, no
// sigmoid) as ONE branch-free code path:  x = acc*scale + bias;  x += r*m1;  x = max(x, lo);  x += r*m2
// with (m1, m2, lo) = (1,0,-inf | 0) residual-before-ReLU, (0,1,..) residual-after-ReLU, (0,0,..) none.
// ncu showed the flag-driven generic version costing ~1800 issue slots per warp per 128-row tile:
// the epilogue, not the tensor pipe, bounded every kernel variant.
struct EpiFast { float m1, m2, lo; };

__device__ __forceinline__ void epilogue_chunk16_floats(const uint32_t* acc, int cc, const float* s_scale,
                                                        const float* s_bias, const ResidualRow& rr, const EpiFast f,
                                                        float (&v)[16]) {
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const float4 sc = *reinterpret_cast<const float4*>(s_scale + cc + 4 * q);
    const float4 bi = *reinterpret_cast<const float4*>(s_bias + cc + 4 * q);
    v[4 * q + 0] = fmaf(__uint_as_float(acc[4 * q + 0]), sc.x, bi.x);
    v[4 * q + 1] = fmaf(__uint_as_float(acc[4 * q + 1]), sc.y, bi.y);
    v[4 * q + 2] = fmaf(__uint_as_float(acc[4 * q + 2]), sc.z, bi.z);
    v[4 * q + 3] = fmaf(__uint_as_float(acc[4 * q + 3]), sc.w, bi.w);
  }
  const uint4 q0 = rr.q[cc >> 3], q1 = rr.q[(cc >> 3) + 1];
  const uint32_t w[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float r0 = bf16_lo(w[j]), r1 = bf16_hi(w[j]);
    v[2 * j] = fmaf(r0, f.m2, fmaxf(fmaf(r0, f.m1, v[2 * j]), f.lo));
    v[2 * j + 1] = fmaf(r1, f.m2, fmaxf(fmaf(r1, f.m1, v[2 * j + 1]), f.lo));
  }
}

__device__ __forceinline__ void epilogue_chunk16_vals(const uint32_t* acc, int cc, const float* s_scale, const float* s_bias,
                                                      const ResidualRow& rr, const EpiFast f, uint4& o0, uint4& o1) {
  float v[16];
  epilogue_chunk16_floats(acc, cc, s_scale, s_bias, rr, f, v);
  o0 = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
  o1 = make_uint4(pack_bf16x2(v[8], v[9]), pack_bf16x2(v[10], v[11]), pack_bf16x2(v[12], v[13]), pack_bf16x2(v[14], v[15]));
}

__device__ __forceinline__ void epilogue_chunk16(const uint32_t* acc, int cc, const float* s_scale, const float* s_bias,
                                                 const ResidualRow& rr, const EpiFast f, __nv_bfloat16* yrow) {
  uint4 o0, o1;
  epilogue_chunk16_vals(acc, cc, s_scale, s_bias, rr, f, o0, o1);
  uint4* o = reinterpret_cast<uint4*>(yrow + cc);
  if (aligned32(o)) { stg256(o, o0, o1); return; }
  o[0] = o0;
  o[1] = o1;
}

// ---- staged epilogue: rows go to a shared-memory tile in the TMA swizzle pattern and ONE bulk tensor store
// writes the tile.  Why: a direct store has every lane write 16 B of its own 64-byte voxel row, so each STG.128
// touches 32 different sectors; ncu (profiles/r01_step_v5_summary.txt) showed the L1 data pipe -- which also
// feeds the UMMA operand reads -- 87 % busy, 47 % of it LSU wavefronts of exactly these stores.  Staged, a row costs
// conflict-free STS.128s and the TMA engine reads the tile at full width.
__device__ __forceinline__ void sts_v4(uint32_t addr, uint4 v) {
  asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
// byte offset of 16-byte chunk `c` of row `rho` in a tile of RB-byte rows swizzled SWIZZLE_<RB>B (tile base 1024-aligned)
template <int RB>
__device__ __forceinline__ uint32_t swz(uint32_t rho, uint32_t c) {
  const uint32_t a = rho * (uint32_t)RB + c * 16u;
  return a ^ (((a >> 7) & (uint32_t)(RB / 16 - 1)) << 4);
}
template <int CP>
__device__ __forceinline__ void epilogue_row_fast_smem(uint32_t taddr, bool store, uint32_t tile, uint32_t rho,
                                                       const float* s_scale, const float* s_bias, const ResidualRow& rr,
                                                       const EpiFast f) {
#pragma unroll
  for (int c0 = 0; c0 < CP; c0 += 32) {
    uint32_t acc[32];
    tmem_ld16(taddr + (uint32_t)c0, *reinterpret_cast<uint32_t(*)[16]>(acc));
    if (CP > 16) tmem_ld16(taddr + (uint32_t)c0 + 16u, *reinterpret_cast<uint32_t(*)[16]>(acc + 16));
    tmem_ld_wait();
    if (store) {
      uint4 o0, o1;
      epilogue_chunk16_vals(acc, c0, s_scale, s_bias, rr, f, o0, o1);
      sts_v4(tile + swz<CP * 2>(rho, c0 / 8), o0);
      sts_v4(tile + swz<CP * 2>(rho, c0 / 8 + 1), o1);
      if (CP > 16) {
        epilogue_chunk16_vals(acc + 16, c0 + 16, s_scale, s_bias, rr, f, o0, o1);
        sts_v4(tile + swz<CP * 2>(rho, c0 / 8 + 2), o0);
        sts_v4(tile + swz<CP * 2>(rho, c0 / 8 + 3), o1);
      }
    }
  }
}
__device__ __forceinline__ uint4 lds_v4(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }
__device__ __forceinline__ void tma_store_5d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2, int c3, int c4) {
  asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];"
               ::"l"(map), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

__device__ __forceinline__ void epilogue_row_fast(int CoutPad, uint32_t taddr, bool in_range, __nv_bfloat16* yrow,
                                                  const float* s_scale, const float* s_bias, const ResidualRow& rr,
                                                  const EpiFast f) {
#pragma unroll
  for (int ci = 0; ci < 2; ++ci) {
    const int c0 = ci * 32;
    if (c0 >= CoutPad) break;
    uint32_t acc[32];
    const bool two = c0 + 16 < CoutPad;
    tmem_ld16(taddr + (uint32_t)c0, *reinterpret_cast<uint32_t(*)[16]>(acc));
    if (two) tmem_ld16(taddr + (uint32_t)c0 + 16u, *reinterpret_cast<uint32_t(*)[16]>(acc + 16));
    tmem_ld_wait();
    if (in_range) {
      epilogue_chunk16(acc, c0, s_scale, s_bias, rr, f, yrow);
      if (two) epilogue_chunk16(acc + 16, c0 + 16, s_scale, s_bias, rr, f, yrow);
    }
  }
}

// generic (slow) path: partial channel counts (Cout = 1), sigmoid, unaligned rows
__device__ __noinline__ void epilogue_row_generic(const EpiParams& e, uint32_t taddr, bool in_range, int64_t vox,
                                                  const float* s_scale, const float* s_bias) {
  for (int c0 = 0; c0 < e.CoutPad; c0 += 16) {
    uint32_t acc[16];
    tmem_ld16(taddr + (uint32_t)c0, acc);
    tmem_ld_wait();
    if (!in_range) continue;
    for (int j = 0; j < 16; ++j) {
      const int c = c0 + j;
      if (c >= e.Cout) break;
      float x = fmaf(__uint_as_float(acc[j]), s_scale[c & 63], s_bias[c & 63]);
      float r = 0.f;
      if (e.residual_mode) r = __bfloat162float(e.residual[vox * e.res_cstride + e.res_coffset + c]);
      if (e.residual_mode == 1) x += r;
      if (e.relu) x = fmaxf(x, 0.f);
      if (e.residual_mode == 2) x += r;
      if (e.sigmoid) x = 1.f / (1.f + __expf(-x));
      if (e.out_f32) reinterpret_cast<float*>(e.y)[vox * e.out_cstride + e.out_coffset + c] = x;
      else reinterpret_cast<__nv_bfloat16*>(e.y)[vox * e.out_cstride + e.out_coffset + c] = __float2bfloat16_rn(x);
    }
  }
}

// variant: 1 = fast path eligible
__device__ __forceinline__ int epilogue_variant(const EpiParams& e) {
  return (e.Cout == e.CoutPad && !e.sigmoid && !e.out_f32 && ((e.out_cstride | e.out_coffset) & 7) == 0) ? 1 : 0;
}

// One accumulator row (this thread's TMEM lane) -> global memory.  taddr: lane/column base.
__device__ __forceinline__ void epilogue_row(const EpiParams& e, int variant, uint32_t taddr, bool in_range, int64_t vox,
                                             const float* s_scale, const float* s_bias, const ResidualRow& rr) {
  if (!variant) { epilogue_row_generic(e, taddr, in_range, vox, s_scale, s_bias); return; }
  EpiFast f;
  f.m1 = e.residual_mode == 1 ? 1.f : 0.f;
  f.m2 = e.residual_mode == 2 ? 1.f : 0.f;
  f.lo = e.relu ? 0.f : -INFINITY;
  epilogue_row_fast(e.CoutPad, taddr, in_range, reinterpret_cast<__nv_bfloat16*>(e.y) + vox * e.out_cstride + e.out_coffset,
                    s_scale, s_bias, rr, f);
}

struct TileCoord { int n, jd, jh, jw; };
__device__ __forceinline__ TileCoord decode_tile(const ConvParams& p, int tile) {
  TileCoord t;
  int tw = tile % p.tiles_w; tile /= p.tiles_w;
  int th = tile % p.tiles_h; tile /= p.tiles_h;
  int td = tile % p.tiles_d; tile /= p.tiles_d;
  t.n = tile; t.jd = td * p.TD; t.jh = th * p.TH; t.jw = tw * p.TW;
  return t;
}

__global__ void __launch_bounds__(kThreads, 1)
conv3d_tcgen05_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w,
                      const __grid_constant__ ConvParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t full_bar[kMaxStages];
  __shared__ __align__(8) uint64_t empty_bar[kMaxStages];
  __shared__ __align__(8) uint64_t tmem_full_bar[2];
  __shared__ __align__(8) uint64_t tmem_empty_bar[2];
  __shared__ uint32_t tmem_base_smem;
  __shared__ __align__(16) float s_scale[64], s_bias[64];

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // dynamic smem base rounded up to 1024 B (swizzle-128B atoms are 1024 B)
  const uint32_t smem_base = (smem_u32(smem) + 1023u) & ~1023u;
  const int stage_bytes = p.a_bytes + p.b_bytes;
  const int ntaps = p.nk[0] * p.nk[1] * p.nk[2];
  const uint32_t tmem_cols = p.CoutPad * 2 <= 32 ? 32u : (p.CoutPad * 2 <= 64 ? 64u : 128u);

  if (threadIdx.x < 64) {
    s_scale[threadIdx.x] = (p.scale && threadIdx.x < p.Cout) ? p.scale[threadIdx.x] : 1.f;
    s_bias[threadIdx.x] = (p.bias && threadIdx.x < p.Cout) ? p.bias[threadIdx.x] : 0.f;
  }
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_x) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(smem_u32(&full_bar[s]), 1);
      mbar_init(smem_u32(&empty_bar[s]), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(smem_u32(&tmem_full_bar[b]), 1);
      mbar_init(smem_u32(&tmem_empty_bar[b]), 4);   // one arrive per epilogue warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)),
                 "r"(tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = tmem_base_smem;

  if (warp == 0) {
    // ===================== TMA producer (warp-wide loop, elected lane issues) =====================
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      const TileCoord tc = decode_tile(p, tile);
      const int cd = tc.jd * p.in_stride, ch = tc.jh * p.in_stride, cw = tc.jw * p.in_stride;
      for (int id = 0; id < p.nk[0]; ++id)
        for (int ih = 0; ih < p.nk[1]; ++ih)
          for (int iw = 0; iw < p.nk[2]; ++iw) {
            mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1u);
            if (elect_one()) {
              const uint32_t fb = smem_u32(&full_bar[stage]);
              const uint32_t sa = smem_base + stage * stage_bytes;
              mbar_expect_tx(fb, (uint32_t)stage_bytes);
              tma_load_5d(sa, &map_x, fb, 0, cw + p.off[2][iw], ch + p.off[1][ih], cd + p.off[0][id], tc.n);
              const int slot = (p.kid[0][id] * p.K + p.kid[1][ih]) * p.K + p.kid[2][iw];
              tma_load_2d(sa + p.a_bytes, &map_w, fb, 0, slot * p.CoutPad);
            }
            __syncwarp();
            if (++stage == p.stages) { stage = 0; phase ^= 1u; }
          }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (warp-wide loop, elected lane issues) =====================
    const uint32_t idesc = make_idesc(kTileM, p.CoutPad);
    const uint32_t desc_hi = (uint32_t)(make_smem_desc(0, p.swizzle_bytes) >> 32);
    const int ksteps = p.Cin >> 4;
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
      const int buf = it & 1;
      const uint32_t use = (uint32_t)(it >> 1);
      mbar_wait(smem_u32(&tmem_empty_bar[buf]), (use & 1u) ^ 1u);
      tcgen05_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(buf * p.CoutPad);
      for (int t = 0; t < ntaps; ++t) {
        mbar_wait(smem_u32(&full_bar[stage]), phase);
        tcgen05_fence_after();
        const uint32_t sa = smem_base + stage * stage_bytes;
        const uint32_t a_lo = ((sa >> 4) & 0x3FFFu) | (1u << 16);
        const uint32_t b_lo = (((sa + p.a_bytes) >> 4) & 0x3FFFu) | (1u << 16);
        if (elect_one()) {
          for (int k = 0; k < ksteps; ++k)   // +32 B per K=16 step inside the swizzled row: +2 in the >>4 field
            umma_bf16(d_tmem, desc64(desc_hi, a_lo + 2 * k), desc64(desc_hi, b_lo + 2 * k), idesc, (t | k) ? 1u : 0u);
          umma_commit(smem_u32(&empty_bar[stage]));          // frees the smem slot once these MMAs retire
          if (t == ntaps - 1) umma_commit(smem_u32(&tmem_full_bar[buf]));   // accumulator complete -> epilogue
        }
        __syncwarp();
        if (++stage == p.stages) { stage = 0; phase ^= 1u; }
      }
    }
  } else {
    // ===================== epilogue (warps 2..5) =====================
    const int quad = warp & 3;                 // TMEM lane quadrant this warp may access
    const int row = quad * 32 + lane;          // accumulator row == box-linear voxel index
    EpiParams epi;
    epi.Cout = p.Cout; epi.CoutPad = p.CoutPad; epi.relu = p.relu; epi.residual_mode = p.residual_mode;
    epi.sigmoid = p.sigmoid; epi.out_f32 = p.out_f32; epi.out_cstride = p.out_cstride; epi.out_coffset = p.out_coffset;
    epi.res_cstride = p.res_cstride; epi.res_coffset = p.res_coffset; epi.residual = p.residual; epi.y = p.y;
    const int variant = epilogue_variant(epi);
    const int r_w = row % p.TW, r_h = (row / p.TW) % p.TH, r_d = row / (p.TW * p.TH);
    int it = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
      const int buf = it & 1;
      const uint32_t use = (uint32_t)(it >> 1);
      const TileCoord tc = decode_tile(p, tile);
      const int jd = tc.jd + r_d, jh = tc.jh + r_h, jw = tc.jw + r_w;
      const bool in_range = jd < p.Dj && jh < p.Hj && jw < p.Wj;
      const int od = jd * p.out_stride + p.out_off_d, oh = jh * p.out_stride + p.out_off_h,
                ow = jw * p.out_stride + p.out_off_w;
      const int64_t vox = (((int64_t)tc.n * p.Do + od) * p.Ho + oh) * p.Wo + ow;
      ResidualRow rr;
      residual_prefetch(epi, in_range, vox, rr);
      mbar_wait(smem_u32(&tmem_full_bar[buf]), use & 1u);
      tcgen05_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(buf * p.CoutPad);
      epilogue_row(epi, variant, taddr, in_range, vox, s_scale, s_bias, rr);
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&tmem_empty_bar[buf]));
    }
  }

  // teardown: everyone done with TMEM before the allocating warp frees it
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
  }
}


// ==========================================================================================
// v2: plane-march kernel for stride-1 "same" convolutions (the layers that carry the FLOPs).
//
// The per-tap kernel above re-fetches every activation k^3 times through TMA (one 128-row box per
// tap); the launch list shows it bound by the TMA row rate (~640 cycles per tap), not by bytes.
// Here a CTA owns a column of the volume -- an (TH x TWv) patch of (h, w), all depth planes --
// and marches along d:
//   * each INPUT plane of the patch (+halo) is TMA-loaded ONCE into a ring of smem slots as a
//     dense [(TH+hw) x WP] array of voxel rows (WP = row pitch 16/32/64, hw = (k-1)*dil);
//   * because the tile width equals the pitch, output row r = h*WP + w of the 128-row MMA tile
//     needs input row r + (kh*dil*WP + kw*dil) of plane d+kd*dil: every filter tap is the SAME
//     smem tile read through a UMMA descriptor whose start address is shifted by whole rows --
//     no data movement per tap at all.  Columns w >= TWv of each row wrap into the next row and
//     are discarded by the epilogue (WP-hw of WP columns useful);
//   * all k^3 weight tiles stay resident in smem for the CTA's lifetime (persistent grid);
//   * swizzle phase of shifted windows: TMA and UMMA both swizzle on ABSOLUTE smem address bits
//     (measured: descriptors with base offset 0 and a start address shifted by any number of
//     rows reproduce the oracle for SWIZZLE_32B/64B/128B; tests/test_gpu_conv3d.py).
// ==========================================================================================
constexpr int kMaxSlots = 12;

struct HaloParams {
  int N, Cin;
  int D, H, W;                 // input == output extent
  int K, dil, pad;             // pad == dil*(K-1)/2
  int WP, TH, TWv;             // row pitch, tile rows (WP*TH == 128), valid columns per row = WP - hw
  int tiles_h, tiles_w, num_cols;
  int plane_bytes;             // TMA bytes per plane: (TH+hw)*WP*Cin*2
  int slot_bytes, nslots;
  int w_tap_bytes;             // CoutPad*Cin*2
  // K-split: a plane / weight tile is stored as `nsub` sub-tiles whose rows are `sub_row_bytes` long
  // (Cin=64: one 128-B-row SWIZZLE_128B tile; Cin=32: two 32-B-row SWIZZLE_32B tiles -- a 64-B-row
  // SWIZZLE_64B tile read in 32-byte K-slices is 2-way bank conflicted, measured 89 vs 57 cycles/MMA)
  int sub_row_bytes, nsub, sub_tile_bytes, w_sub_bytes;
  int bo_mode;
  int w_rows_per_tap, w_row0;  // packed-weight rows per tap (full CoutPad) and first row of this launch's Cout slice
  int tma_store, stage_bytes;  // staged epilogue (kd-fused kernel): two swizzled output tiles of stage_bytes each
  const float* scale;
  const float* bias;
  const float* addend;         // v8 only: fp32 [N,3,H,W,Cout] added to the accumulator (first / interior / last plane)
  int add_lo, add_hi;          // output planes that take addend plane 0 / 2 (default 0 / D-1; -1 = none: depth slabs)
  EpiParams epi;
};

template <int K, int KSTEPS, int SUBROW>
__global__ void __launch_bounds__(kThreads, 1)
conv3d_halo_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w,
                   const __grid_constant__ HaloParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t full_bar[kMaxSlots];
  __shared__ __align__(8) uint64_t empty_bar[kMaxSlots];
  __shared__ __align__(8) uint64_t w_bar;
  __shared__ __align__(8) uint64_t tmem_full_bar[2];
  __shared__ __align__(8) uint64_t tmem_empty_bar[2];
  __shared__ uint32_t tmem_base_smem;
  __shared__ __align__(16) float s_scale[64], s_bias[64];

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int K3 = p.K * p.K * p.K;
  const int hw = (p.K - 1) * p.dil;
  const uint32_t w_base = (smem_u32(smem) + 1023u) & ~1023u;
  const uint32_t slots_base = w_base + (((uint32_t)(K3 * p.w_tap_bytes) + 1023u) & ~1023u);
  const uint32_t tmem_cols = p.epi.CoutPad * 2 <= 32 ? 32u : (p.epi.CoutPad * 2 <= 64 ? 64u : 128u);
  const int planes_per_col = p.D + hw;
  constexpr int KPS = SUBROW / 32;            // K=16 steps per sub-tile row

  if (threadIdx.x < 64) {
    s_scale[threadIdx.x] = (p.scale && threadIdx.x < p.epi.Cout) ? p.scale[threadIdx.x] : 1.f;
    s_bias[threadIdx.x] = (p.bias && threadIdx.x < p.epi.Cout) ? p.bias[threadIdx.x] : 0.f;
  }
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_x) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
    for (int s = 0; s < p.nslots; ++s) {
      mbar_init(smem_u32(&full_bar[s]), 1);
      mbar_init(smem_u32(&empty_bar[s]), 1);
    }
    mbar_init(smem_u32(&w_bar), 1);
    for (int b = 0; b < 2; ++b) {
      mbar_init(smem_u32(&tmem_full_bar[b]), 1);
      mbar_init(smem_u32(&tmem_empty_bar[b]), 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)),
                 "r"(tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = tmem_base_smem;

  if (warp == 0) {
    // ===================== TMA producer (warp-wide loop, elected lane issues) =====================
    const uint32_t wb = smem_u32(&w_bar);
    if (elect_one()) {
      mbar_expect_tx(wb, (uint32_t)(K3 * p.w_tap_bytes));
      for (int t = 0; t < K3; ++t)
        for (int sb = 0; sb < p.nsub; ++sb)
          tma_load_2d(w_base + t * p.w_tap_bytes + sb * p.w_sub_bytes, &map_w, wb, sb * (SUBROW / 2), t * p.epi.CoutPad);
    }
    __syncwarp();
    uint32_t q = 0;
    for (int col = blockIdx.x; col < p.num_cols; col += gridDim.x) {
      int tw = col % p.tiles_w, rest = col / p.tiles_w;
      int th = rest % p.tiles_h, n = rest / p.tiles_h;
      const int w0 = tw * p.TWv - p.pad, h0 = th * p.TH - p.pad;
      for (int ip = -p.pad; ip < p.D + hw - p.pad; ++ip, ++q) {
        const uint32_t slot = q % (uint32_t)p.nslots, phase = (q / (uint32_t)p.nslots) & 1u;
        mbar_wait(smem_u32(&empty_bar[slot]), phase ^ 1u);
        if (elect_one()) {
          const uint32_t fb = smem_u32(&full_bar[slot]);
          mbar_expect_tx(fb, (uint32_t)p.plane_bytes);
          for (int sb = 0; sb < p.nsub; ++sb)
            tma_load_5d(slots_base + slot * p.slot_bytes + sb * p.sub_tile_bytes, &map_x, fb, sb * (SUBROW / 2), w0, h0, ip, n);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (warp-wide loop, elected lane issues) =====================
    const uint32_t idesc = make_idesc(kTileM, p.epi.CoutPad);
    const uint32_t desc_hi = (uint32_t)(make_smem_desc(0, SUBROW) >> 32);
    const uint32_t lo_flags = 1u << 16;                                   // LBO field (ignored for swizzled K-major)
    const uint32_t a_sub = (uint32_t)p.sub_tile_bytes >> 4, b_sub = (uint32_t)p.w_sub_bytes >> 4;
    // descriptor offsets, in 16-byte units: filter tap (kh,kw) = whole-row shift of the plane tile
    uint32_t off_hw[K * K];
#pragma unroll
    for (int kh = 0; kh < K; ++kh)
#pragma unroll
      for (int kw = 0; kw < K; ++kw) off_hw[kh * K + kw] = (uint32_t)((kh * p.dil * p.WP + kw * p.dil) * SUBROW) >> 4;
    const uint32_t b_lo0 = ((w_base >> 4) & 0x3FFFu) | lo_flags;
    const uint32_t b_step = (uint32_t)p.w_tap_bytes >> 4;
    mbar_wait(smem_u32(&w_bar), 0);
    uint32_t q0 = 0, it = 0;
    for (int col = blockIdx.x; col < p.num_cols; col += gridDim.x) {
      for (int d = 0; d < p.D; ++d, ++it) {
        for (int pl = (d == 0 ? 0 : hw); pl <= hw; ++pl) {     // planes newly needed by this output plane
          const uint32_t qq = q0 + d + pl;
          mbar_wait(smem_u32(&full_bar[qq % (uint32_t)p.nslots]), (qq / (uint32_t)p.nslots) & 1u);
        }
        const uint32_t buf = it & 1u;
        mbar_wait(smem_u32(&tmem_empty_bar[buf]), ((it >> 1) & 1u) ^ 1u);
        tcgen05_fence_after();
        const uint32_t d_tmem = tmem_base + buf * (uint32_t)p.epi.CoutPad;
        uint32_t plane_lo[K];
#pragma unroll
        for (int kd = 0; kd < K; ++kd) {
          const uint32_t qq = q0 + d + kd * p.dil;
          plane_lo[kd] = (((slots_base + (qq % (uint32_t)p.nslots) * p.slot_bytes) >> 4) & 0x3FFFu) | lo_flags;
        }
        if (elect_one()) {
#pragma unroll
          for (int kd = 0; kd < K; ++kd)
#pragma unroll
            for (int t2 = 0; t2 < K * K; ++t2) {
              const uint32_t a_lo = plane_lo[kd] + off_hw[t2];
              const uint32_t b_lo = b_lo0 + (uint32_t)(kd * K * K + t2) * b_step;
#pragma unroll
              for (int k = 0; k < KSTEPS; ++k) {
                const uint32_t ka = (uint32_t)(k / KPS) * a_sub + 2u * (uint32_t)(k % KPS);
                const uint32_t kb = (uint32_t)(k / KPS) * b_sub + 2u * (uint32_t)(k % KPS);
                umma_bf16(d_tmem, desc64(desc_hi, a_lo + ka), desc64(desc_hi, b_lo + kb), idesc, (kd | t2 | k) ? 1u : 0u);
              }
            }
          umma_commit(smem_u32(&tmem_full_bar[buf]));
          umma_commit(smem_u32(&empty_bar[(q0 + d) % (uint32_t)p.nslots]));   // input plane d is done
        }
        __syncwarp();
      }
      if (elect_one())
        for (int pl = 0; pl < hw; ++pl) umma_commit(smem_u32(&empty_bar[(q0 + p.D + pl) % (uint32_t)p.nslots]));
      __syncwarp();
      q0 += (uint32_t)planes_per_col;
    }
  } else {
    // ===================== epilogue (warps 2..5) =====================
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const int r_w = row % p.WP, r_h = row / p.WP;
    const int variant = epilogue_variant(p.epi);
    uint32_t it = 0;
    for (int col = blockIdx.x; col < p.num_cols; col += gridDim.x) {
      int tw = col % p.tiles_w, rest = col / p.tiles_w;
      int th = rest % p.tiles_h, n = rest / p.tiles_h;
      const int ow = tw * p.TWv + r_w, oh = th * p.TH + r_h;
      const bool in_range = r_w < p.TWv && ow < p.W && oh < p.H;
      const int64_t vox0 = (((int64_t)n * p.D) * p.H + oh) * p.W + ow;
      for (int d = 0; d < p.D; ++d, ++it) {
        const uint32_t buf = it & 1u;
        const int64_t vox = vox0 + (int64_t)d * p.H * p.W;
        ResidualRow rr;
        residual_prefetch(p.epi, in_range, vox, rr);
        mbar_wait(smem_u32(&tmem_full_bar[buf]), (it >> 1) & 1u);
        tcgen05_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + buf * (uint32_t)p.epi.CoutPad;
        epilogue_row(p.epi, variant, taddr, in_range, vox, s_scale, s_bias, rr);
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&tmem_empty_bar[buf]));
      }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
  }
}


// ==========================================================================================
// v3: kd-fused plane march (3x3x3, stride 1, "same" padding) -- the kernel the trunk runs on.
//
// Measured with ncu on v2 (profiles/r01_conv_halo_v2.ncu-rep): every M=128,N=32,K=16 MMA occupies
// the tensor pipe for 64 cycles (16 would be peak) -- with both operands in shared memory the
// 128x16 A slice (4 KB) is the cost, whatever N is.  So N must grow.  Here the three depth taps
// are fused into ONE instruction: for input plane p and in-plane tap (kh,kw)
//     D[128, 3*Cout] += A_p(kh,kw)[128, Cin] * [W(0,kh,kw) | W(1,kh,kw) | W(2,kh,kw)]
// whose three column blocks are the accumulators of output planes p+1, p, p-1.  Accumulators
// live in a RING of R = 512/Cout TMEM blocks, block(g) = (-g) mod R for accumulator plane g, so
// the three blocks an input plane updates are always adjacent columns (one MMA; split in two
// where the ring wraps).  Each input plane is read from HBM/L2 once, read from smem 9x (not 27x)
// and each output plane has R-2 planes of slack before its TMEM block is reused, so the
// epilogue (which drains a block, stores it, and zero-fills it with tcgen05.st for its next
// use; all MMAs accumulate) is off the critical path.
// ==========================================================================================
constexpr int kMaxBlocks = 32;

__device__ __forceinline__ void tmem_st16_zero(uint32_t taddr) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1};"
      ::"r"(taddr), "r"(0u)
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// TCOLS = TMEM columns this CTA allocates: 512 (one CTA per SM) or 256 (two co-resident CTAs per SM:
// while one CTA's MMA warp does its per-plane bookkeeping the other CTA's MMAs keep the tensor pipe busy).
template <int KSTEPS, int SUBROW, int CP, int TCOLS, bool STAGED>
__global__ void __launch_bounds__(kThreads, TCOLS == 512 ? 1 : 2)
conv3d_kdfuse_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w,
                     const __grid_constant__ CUtensorMap map_y, const __grid_constant__ HaloParams p) {
  constexpr int K = 3;
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t full_bar[kMaxSlots];
  __shared__ __align__(8) uint64_t empty_bar[kMaxSlots];
  __shared__ __align__(8) uint64_t w_bar;
  __shared__ __align__(8) uint64_t acc_full_bar[kMaxBlocks];
  __shared__ __align__(8) uint64_t acc_empty_bar[kMaxBlocks];
  __shared__ uint32_t tmem_base_smem;
  __shared__ __align__(16) float s_scale[64], s_bias[64];

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  constexpr int K3 = 27;
  constexpr uint32_t R = (uint32_t)TCOLS / (uint32_t)CP;  // accumulator blocks in the TMEM ring (power of two, >= 8)
  constexpr uint32_t RMASK = R - 1u;
  constexpr uint32_t LOGR = R == 32u ? 5u : (R == 16u ? 4u : 3u);
  static_assert(CP == 16 || CP == 32 || CP == 64, "CoutPad must be 16, 32 or 64");
  static_assert(R == 8u || R == 16u || R == 32u, "accumulator ring must hold 8, 16 or 32 blocks");
  const uint32_t w_base = (smem_u32(smem) + 1023u) & ~1023u;
  const uint32_t stage_base = w_base + (((uint32_t)(K3 * p.w_tap_bytes) + 1023u) & ~1023u);   // 2 output tiles (staged epilogue)
  const uint32_t slots_base = stage_base + 2u * (uint32_t)p.stage_bytes;
  const uint32_t acc_per_col = (uint32_t)p.D + 2u;        // accumulator planes per column: out[-1] .. out[D]

  if (threadIdx.x < 64) {
    s_scale[threadIdx.x] = (p.scale && threadIdx.x < p.epi.Cout) ? p.scale[threadIdx.x] : 1.f;
    s_bias[threadIdx.x] = (p.bias && threadIdx.x < p.epi.Cout) ? p.bias[threadIdx.x] : 0.f;
  }
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_x) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
    if (STAGED) asm volatile("prefetch.tensormap [%0];" ::"l"(&map_y) : "memory");
    for (int s = 0; s < p.nslots; ++s) {
      mbar_init(smem_u32(&full_bar[s]), 1);
      mbar_init(smem_u32(&empty_bar[s]), 1);
    }
    mbar_init(smem_u32(&w_bar), 1);
    for (uint32_t b = 0; b < R; ++b) {
      mbar_init(smem_u32(&acc_full_bar[b]), 1);
      mbar_init(smem_u32(&acc_empty_bar[b]), 4);          // one arrive per epilogue warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)),
                 "r"((uint32_t)TCOLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  if (warp >= 2) {                                        // zero the whole accumulator ring once
    const uint32_t lane_base = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    for (uint32_t c = 0; c < (uint32_t)TCOLS; c += 16u) tmem_st16_zero(lane_base + c);
    tmem_st_wait();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();

  // All ring positions below are carried incrementally (wrap by compare / power-of-two mask): the
  // first version recomputed `q % nslots`, `g % R`, `g / R` per plane with run-time divisors, and the
  // ncu source page showed the MMA warp spending ~75 % of its issue slots in that integer code
  // (MUFU.RCP division sequences) while the tensor-pipe queue (about 6 UTCHMMA deep) ran dry:
  // ~1200 idle cycles per plane on top of the MMA time.
  if (warp == 0) {
    // ===================== TMA producer =====================
    const uint32_t wb = smem_u32(&w_bar);
    if (elect_one()) {
      mbar_expect_tx(wb, (uint32_t)(K3 * p.w_tap_bytes));
      // smem order [(kh,kw)][kd]: the three depth taps of one in-plane tap are adjacent -> one B operand of 3*Cout rows
      for (int t2 = 0; t2 < K * K; ++t2)
        for (int kd = 0; kd < K; ++kd)
          tma_load_2d(w_base + (t2 * K + kd) * p.w_tap_bytes, &map_w, wb, 0, (kd * K * K + t2) * p.w_rows_per_tap + p.w_row0);
    }
    __syncwarp();
    uint32_t slot = 0, phase = 0;
    uint32_t slot_addr = slots_base;
    int tw = blockIdx.x % p.tiles_w, rest = blockIdx.x / p.tiles_w;
    for (int col = blockIdx.x; col < p.num_cols; col += gridDim.x) {
      const int th = rest % p.tiles_h, n = rest / p.tiles_h;
      const int w0 = tw * p.TWv - p.pad, h0 = th * p.TH - p.pad;
      for (int ip = 0; ip < p.D; ++ip) {                  // only real planes: the zero planes -1 and D contribute nothing
        mbar_wait(smem_u32(&empty_bar[slot]), phase ^ 1u);
        if (elect_one()) {
          const uint32_t fb = smem_u32(&full_bar[slot]);
          mbar_expect_tx(fb, (uint32_t)p.plane_bytes);
          tma_load_5d(slot_addr, &map_x, fb, 0, w0, h0, ip, n);
        }
        __syncwarp();
        slot_addr += (uint32_t)p.slot_bytes;
        if (++slot == (uint32_t)p.nslots) { slot = 0; phase ^= 1u; slot_addr = slots_base; }
      }
      const int nc = col + (int)gridDim.x;                // next column's (tw, rest) -- one division per column
      tw = nc % p.tiles_w; rest = nc / p.tiles_w;
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc1 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(CP >> 3) << 17) | ((uint32_t)(kTileM >> 4) << 24);
    constexpr uint32_t idesc2 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)((2 * CP) >> 3) << 17) | ((uint32_t)(kTileM >> 4) << 24);
    constexpr uint32_t idesc3 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)((3 * CP) >> 3) << 17) | ((uint32_t)(kTileM >> 4) << 24);
    const uint32_t desc_hi = (uint32_t)(make_smem_desc(0, SUBROW) >> 32);
    constexpr uint32_t lo_flags = 1u << 16;
    uint32_t off_hw[K * K];
#pragma unroll
    for (int kh = 0; kh < K; ++kh)
#pragma unroll
      for (int kw = 0; kw < K; ++kw) off_hw[kh * K + kw] = (uint32_t)((kh * p.WP + kw) * SUBROW) >> 4;
    const uint32_t b_lo0 = ((w_base >> 4) & 0x3FFFu) | lo_flags;
    constexpr uint32_t b_tap = (uint32_t)(CP * KSTEPS * 32) >> 4;     // one (kd) tile: CP rows x Cin*2 bytes
    const uint32_t a_lo0 = ((slots_base >> 4) & 0x3FFFu) | lo_flags;
    const uint32_t a_step = (uint32_t)p.slot_bytes >> 4;
    mbar_wait(smem_u32(&w_bar), 0);
    uint32_t slot = 0, phase = 0, a_plane = a_lo0;
    uint32_t g = 0;                                       // accumulator plane index of out[pl-1] (global over columns)
    for (int col = blockIdx.x; col < p.num_cols; col += gridDim.x) {
      for (int pl = 0; pl < p.D; ++pl, ++g) {
        mbar_wait(smem_u32(&full_bar[slot]), phase);
        // accumulator planes touched: g+2 (kd=0, out[pl+1]), g+1 (kd=1), g (kd=2, out[pl-1])
        if (pl == 0) {
          mbar_wait(smem_u32(&acc_empty_bar[(0u - g) & RMASK]), ((g >> LOGR) & 1u) ^ 1u);
          mbar_wait(smem_u32(&acc_empty_bar[(0u - (g + 1u)) & RMASK]), (((g + 1u) >> LOGR) & 1u) ^ 1u);
        }
        mbar_wait(smem_u32(&acc_empty_bar[(0u - (g + 2u)) & RMASK]), (((g + 2u) >> LOGR) & 1u) ^ 1u);
        tcgen05_fence_after();
        const uint32_t b0 = (0u - (g + 2u)) & RMASK;                           // block of kd = 0
        const uint32_t n0 = min(3u, R - b0);                                   // blocks before the ring wraps
        const uint32_t d0 = tmem_base + b0 * (uint32_t)CP;
        if (elect_one()) {
          if (n0 == 3u) {
#pragma unroll
            for (int t2 = 0; t2 < K * K; ++t2)
#pragma unroll
              for (int k = 0; k < KSTEPS; ++k)
                umma_bf16(d0, desc64(desc_hi, a_plane + off_hw[t2] + 2u * k),
                          desc64(desc_hi, b_lo0 + (uint32_t)(t2 * K) * b_tap + 2u * k), idesc3, 1u);
          } else {
            const uint32_t ia = n0 == 1u ? idesc1 : idesc2, ib = n0 == 1u ? idesc2 : idesc1;
#pragma unroll
            for (int t2 = 0; t2 < K * K; ++t2)
#pragma unroll
              for (int k = 0; k < KSTEPS; ++k) {
                const uint64_t ad = desc64(desc_hi, a_plane + off_hw[t2] + 2u * k);
                const uint32_t bl = b_lo0 + (uint32_t)(t2 * K) * b_tap + 2u * k;
                umma_bf16(d0, ad, desc64(desc_hi, bl), ia, 1u);                            // kd in [0, n0)
                umma_bf16(tmem_base, ad, desc64(desc_hi, bl + n0 * b_tap), ib, 1u);        // kd in [n0, 3) at block 0
              }
          }
          umma_commit(smem_u32(&empty_bar[slot]));                                          // plane consumed
          umma_commit(smem_u32(&acc_full_bar[(0u - g) & RMASK]));                           // out[pl-1] complete
          if (pl == p.D - 1) {                                                              // column tail: out[D-1], out[D]
            umma_commit(smem_u32(&acc_full_bar[(0u - (g + 1u)) & RMASK]));
            umma_commit(smem_u32(&acc_full_bar[(0u - (g + 2u)) & RMASK]));
          }
        }
        __syncwarp();
        a_plane += a_step;
        if (++slot == (uint32_t)p.nslots) { slot = 0; phase ^= 1u; a_plane = a_lo0; }
      }
      g += 2u;                                            // acc_per_col = D + 2
    }
  } else {
    // ===================== epilogue (warps 2..5) =====================
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const int r_w = row % p.WP, r_h = row / p.WP;
    const int variant = epilogue_variant(p.epi);
    const bool staged = STAGED && variant == 1;           // uniform over the CTA (host launches STAGED only then)
    const bool issuer = row == 0;                          // the thread that owns the bulk-store groups
    const uint32_t rho = (uint32_t)(r_h * p.TWv + r_w);    // row of this thread in the compacted output tile
    EpiFast f;
    f.m1 = p.epi.residual_mode == 1 ? 1.f : 0.f;
    f.m2 = p.epi.residual_mode == 2 ? 1.f : 0.f;
    f.lo = p.epi.relu ? 0.f : -INFINITY;
    const uint32_t lane_base = tmem_base + ((uint32_t)(quad * 32) << 16);
    const int64_t plane_vox = (int64_t)p.H * p.W;
    uint32_t g = 0, sbuf = 0;
    int tw = blockIdx.x % p.tiles_w, rest = blockIdx.x / p.tiles_w;
    for (int col = blockIdx.x; col < p.num_cols; col += gridDim.x) {
      const int th = rest % p.tiles_h, n = rest / p.tiles_h;
      const int ow = tw * p.TWv + r_w, oh = th * p.TH + r_h;
      const bool in_range = r_w < p.TWv && ow < p.W && oh < p.H;
      int64_t vox = (((int64_t)n * p.D) * p.H + oh) * p.W + ow - plane_vox;       // accumulator plane a <-> output plane a - 1
      for (uint32_t a = 0; a < acc_per_col; ++a, ++g, vox += plane_vox) {
        const uint32_t blk = (0u - g) & RMASK;
        const bool real = a >= 1u && a <= (uint32_t)p.D;
        ResidualRow rr;
        residual_prefetch(p.epi, in_range && real, vox, rr);
        mbar_wait(smem_u32(&acc_full_bar[blk]), (g >> LOGR) & 1u);
        tcgen05_fence_after();
        const uint32_t taddr = lane_base + blk * (uint32_t)CP;
        if (real && staged) epilogue_row_fast_smem<CP>(taddr, r_w < p.TWv, stage_base + sbuf * (uint32_t)p.stage_bytes, rho,
                                                       s_scale, s_bias, rr, f);
        else if (real) epilogue_row(p.epi, variant, taddr, in_range, vox, s_scale, s_bias, rr);
#pragma unroll
        for (int c = 0; c < CP; c += 16) tmem_st16_zero(taddr + (uint32_t)c);   // ready for its next output plane
        tmem_st_wait();
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&acc_empty_bar[blk]));
        if (real && staged) {
          // tile complete -> one bulk tensor store.  The issuer first waits until the PREVIOUS store has finished
          // reading the other buffer, so after this barrier every thread may overwrite that buffer (next plane).
          fence_proxy_async_smem();
          if (issuer) tma_store_wait_read0();
          epi_bar_sync();
          if (issuer) {
            tma_store_5d(&map_y, stage_base + sbuf * (uint32_t)p.stage_bytes, 0, tw * p.TWv, th * p.TH, (int)a - 1, n);
            tma_store_commit();
          }
          sbuf ^= 1u;
        }
      }
      const int nc = col + (int)gridDim.x;
      tw = nc % p.tiles_w; rest = nc / p.tiles_w;
    }
    if (staged && issuer) tma_store_wait_all();
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TCOLS) : "memory");
  }
}

// ==========================================================================================
// v7: kw+kd-fused plane march (3x3x3, stride 1, Cout = 32) -- lifts the N = 96 MMA-issue floor of v3.
//
// v3 issues one N = 3*Cout = 96 MMA per in-plane tap (kh,kw): 71.6 cycles each where the tensor pipe needs 48
// (profiles/r01_umma_rate.txt: an SS-mode M=128,K=16 MMA costs max(71.6, N/2) cycles), so the Cout = 32 layers
// -- 80 % of the trunk's FLOPs -- cannot pass 67 % of the tensor peak.  Here the three kw taps are fused into N as
// well.  Because the tile width equals the row pitch WP, the A window of tap (kh,kw) is the window of (kh,0)
// shifted by kw rows, so with
//     P_kw[m] = sum_{kd,kh} X_p(kh)[m] * W(kd,kh,kw)        (un-shifted windows: 3 per plane instead of 9)
// the convolution is  out[m] = P_0[m] + P_1[m+1] + P_2[m+2]:  the kw shift moves from the A operand to the
// accumulator ROW, i.e. to the TMEM lane, and is undone in the epilogue with two warp shuffles per channel
// (rows m+1, m+2 of a valid output column w < WP-2 are in the same tile row, hence -- for WP <= 32 -- in the
// same warp).  One input plane and kh now update a [128 x 288] slab: the accumulator blocks (3 kw x 32 columns
// each) of output planes p-1, p, p+1, adjacent in a ring of 5 blocks (480 TMEM columns), issued as two MMAs
// of N = 144 (72 cycles each = the tensor-pipe time), or N = 192 + 96 where the ring wraps (2 planes in 5):
// 6*KSTEPS MMAs and ~153 cycles per (kh, K step) on average instead of 9*KSTEPS MMAs and 215 cycles.
// ==========================================================================================
constexpr uint32_t kKwBlocks = 5;       // accumulator blocks in the TMEM ring
constexpr uint32_t kKwBlkCols = 96;     // 3 kw x 32 output channels

__device__ __forceinline__ constexpr uint32_t idesc_bf16_m128(uint32_t n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((n >> 3) << 17) | ((uint32_t)(kTileM >> 4) << 24);
}
// ring position (block, use-parity) of an accumulator plane, carried incrementally (5 is not a power of two)
struct KwRing { uint32_t b, ph; };
__device__ __forceinline__ KwRing kw_next(KwRing r) {
  KwRing n{r.b + 1u, r.ph};
  if (n.b == kKwBlocks) { n.b = 0u; n.ph ^= 1u; }
  return n;
}

// Warps: 0 = TMA producer, 1 = MMA issuer, 2..9 = epilogue, two groups of four.  ncu on the first version (one group):
// the epilogue, not the tensor pipe, set the pace -- 421 instructions per plane per warp with ONE warp per scheduler
// (SHFL / LDS round trips and dependent FP32 chains fully exposed: 3.9 cycles per instruction, tensor pipe 37 % busy).
// Two groups drain alternate accumulator planes, so every scheduler has two epilogue warps to interleave, and the
// per-channel scale / bias live in registers.
constexpr int kKwThreads = 320;

template <int KSTEPS, int SUBROW, bool RES>
__global__ void __launch_bounds__(kKwThreads, 1)
conv3d_kwfuse_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w,
                     const __grid_constant__ HaloParams p) {
  constexpr int K = 3, CP = 32, K3 = 27;
  constexpr uint32_t TCOLS = 512;
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t full_bar[kMaxSlots];
  __shared__ __align__(8) uint64_t empty_bar[kMaxSlots];
  __shared__ __align__(8) uint64_t w_bar;
  __shared__ __align__(8) uint64_t acc_full_bar[kKwBlocks];
  __shared__ __align__(8) uint64_t acc_empty_bar[kKwBlocks];
  __shared__ uint32_t tmem_base_smem;
  __shared__ __align__(16) float s_scale[64], s_bias[64];

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t w_base = (smem_u32(smem) + 1023u) & ~1023u;
  const uint32_t slots_base = w_base + (((uint32_t)(K3 * p.w_tap_bytes) + 1023u) & ~1023u);
  const uint32_t acc_per_col = (uint32_t)p.D + 2u;        // accumulator planes per column: out[-1] .. out[D]

  if (threadIdx.x < 64) {
    s_scale[threadIdx.x] = (p.scale && threadIdx.x < p.epi.Cout) ? p.scale[threadIdx.x] : 1.f;
    s_bias[threadIdx.x] = (p.bias && threadIdx.x < p.epi.Cout) ? p.bias[threadIdx.x] : 0.f;
  }
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_x) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
    for (int s = 0; s < p.nslots; ++s) {
      mbar_init(smem_u32(&full_bar[s]), 1);
      mbar_init(smem_u32(&empty_bar[s]), 1);
    }
    mbar_init(smem_u32(&w_bar), 1);
    for (uint32_t b = 0; b < kKwBlocks; ++b) {
      mbar_init(smem_u32(&acc_full_bar[b]), 1);
      mbar_init(smem_u32(&acc_empty_bar[b]), 4);          // one arrive per epilogue warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)),
                 "r"(TCOLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  if (warp >= 2 && warp < 6) {                            // zero the whole accumulator ring once
    const uint32_t lane_base = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    for (uint32_t c = 0; c < TCOLS; c += 16u) tmem_st16_zero(lane_base + c);
    tmem_st_wait();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();

  if (warp == 0) {
    // ===================== TMA producer =====================
    const uint32_t wb = smem_u32(&w_bar);
    if (elect_one()) {
      mbar_expect_tx(wb, (uint32_t)(K3 * p.w_tap_bytes));
      // smem order [kh][kd = 2,1,0][kw]: the 9 tiles of one kh form ONE B operand of 288 rows whose row order is
      // the column order of the accumulator slab (blocks of planes p-1, p, p+1; kw-major inside a block)
      for (int kh = 0; kh < K; ++kh)
        for (int j = 0; j < K; ++j)
          for (int kw = 0; kw < K; ++kw)
            tma_load_2d(w_base + ((kh * K + j) * K + kw) * p.w_tap_bytes, &map_w, wb, 0,
                        (((K - 1 - j) * K + kh) * K + kw) * p.w_rows_per_tap + p.w_row0);
    }
    __syncwarp();
    uint32_t slot = 0, phase = 0;
    uint32_t slot_addr = slots_base;
    int tw = blockIdx.x % p.tiles_w, rest = blockIdx.x / p.tiles_w;
    for (int col = blockIdx.x; col < p.num_cols; col += gridDim.x) {
      const int th = rest % p.tiles_h, n = rest / p.tiles_h;
      const int w0 = tw * p.TWv - p.pad, h0 = th * p.TH - p.pad;
      for (int ip = 0; ip < p.D; ++ip) {
        mbar_wait(smem_u32(&empty_bar[slot]), phase ^ 1u);
        if (elect_one()) {
          const uint32_t fb = smem_u32(&full_bar[slot]);
          mbar_expect_tx(fb, (uint32_t)p.plane_bytes);
          tma_load_5d(slot_addr, &map_x, fb, 0, w0, h0, ip, n);
        }
        __syncwarp();
        slot_addr += (uint32_t)p.slot_bytes;
        if (++slot == (uint32_t)p.nslots) { slot = 0; phase ^= 1u; slot_addr = slots_base; }
      }
      const int nc = col + (int)gridDim.x;
      tw = nc % p.tiles_w; rest = nc / p.tiles_w;
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    const uint32_t desc_hi = (uint32_t)(make_smem_desc(0, SUBROW) >> 32);
    constexpr uint32_t lo_flags = 1u << 16;
    constexpr uint32_t row16 = (uint32_t)SUBROW >> 4;                  // descriptor units (16 B) per operand row
    const uint32_t a_kh = (uint32_t)p.WP * row16;                      // A window of kh starts kh*WP rows further
    const uint32_t b_lo0 = ((w_base >> 4) & 0x3FFFu) | lo_flags;
    constexpr uint32_t b_kh = 9u * (uint32_t)CP * row16;               // 288 weight rows per kh
    const uint32_t a_lo0 = ((slots_base >> 4) & 0x3FFFu) | lo_flags;
    const uint32_t a_step = (uint32_t)p.slot_bytes >> 4;
    mbar_wait(smem_u32(&w_bar), 0);
    uint32_t slot = 0, phase = 0, a_plane = a_lo0;
    KwRing r0{0u, 0u};                                    // ring position of accumulator plane g = out[pl-1]
    for (int col = blockIdx.x; col < p.num_cols; col += gridDim.x) {
      for (int pl = 0; pl < p.D; ++pl) {
        const KwRing r1 = kw_next(r0), r2 = kw_next(r1);
        mbar_wait(smem_u32(&full_bar[slot]), phase);
        if (pl == 0) {
          mbar_wait(smem_u32(&acc_empty_bar[r0.b]), r0.ph ^ 1u);
          mbar_wait(smem_u32(&acc_empty_bar[r1.b]), r1.ph ^ 1u);
        }
        mbar_wait(smem_u32(&acc_empty_bar[r2.b]), r2.ph ^ 1u);
        tcgen05_fence_after();
        // slab columns: [block r0 (kd=2) | r1 (kd=1) | r2 (kd=0)], contiguous unless the ring wraps after 1 or 2 blocks
        const uint32_t nb = kKwBlocks - r0.b;             // blocks before the wrap (>= 3: none)
        const uint32_t n1 = nb >= 3u ? 144u : nb * kKwBlkCols;         // 144 | 192 | 96
        const uint32_t n2 = 288u - n1;
        const uint32_t d1 = tmem_base + r0.b * kKwBlkCols;
        const uint32_t d2 = nb >= 3u ? d1 + 144u : tmem_base;
        const uint32_t i1 = idesc_bf16_m128(n1), i2 = idesc_bf16_m128(n2);
        const uint32_t b2 = n1 * row16;
        if (elect_one()) {
#pragma unroll
          for (int kh = 0; kh < K; ++kh)
#pragma unroll
            for (int k = 0; k < KSTEPS; ++k) {
              const uint64_t ad = desc64(desc_hi, a_plane + (uint32_t)kh * a_kh + 2u * k);
              const uint32_t bl = b_lo0 + (uint32_t)kh * b_kh + 2u * k;
              umma_bf16(d1, ad, desc64(desc_hi, bl), i1, 1u);
              umma_bf16(d2, ad, desc64(desc_hi, bl + b2), i2, 1u);
            }
          umma_commit(smem_u32(&empty_bar[slot]));                     // plane consumed
          umma_commit(smem_u32(&acc_full_bar[r0.b]));                  // out[pl-1] complete
          if (pl == p.D - 1) {                                         // column tail: out[D-1], out[D]
            umma_commit(smem_u32(&acc_full_bar[r1.b]));
            umma_commit(smem_u32(&acc_full_bar[r2.b]));
          }
        }
        __syncwarp();
        a_plane += a_step;
        if (++slot == (uint32_t)p.nslots) { slot = 0; phase ^= 1u; a_plane = a_lo0; }
        r0 = r1;
      }
      r0 = kw_next(kw_next(r0));                          // acc_per_col = D + 2
    }
  } else {
    // ===================== epilogue (warps 2..9) =====================
    const int quad = warp & 3;                            // TMEM lane quadrant this warp may access
    const uint32_t grp = (uint32_t)(warp - 2) >> 2;       // drains accumulator planes with (global index & 1) == grp
    const int row = quad * 32 + lane;
    const int r_w = row % p.WP, r_h = row / p.WP;
    const float m1 = p.epi.residual_mode == 1 ? 1.f : 0.f, m2 = p.epi.residual_mode == 2 ? 1.f : 0.f;
    const float lo = p.epi.relu ? 0.f : -INFINITY;
    float2 sc[CP / 2], bi[CP / 2];
#pragma unroll
    for (int j = 0; j < CP / 2; ++j) {
      sc[j] = make_float2(s_scale[2 * j], s_scale[2 * j + 1]);
      bi[j] = make_float2(s_bias[2 * j], s_bias[2 * j + 1]);
    }
    const uint32_t lane_base = tmem_base + ((uint32_t)(quad * 32) << 16);
    const int64_t plane_vox = (int64_t)p.H * p.W;
    KwRing rg{0u, 0u};
    uint32_t par = 0;                                     // parity of the global accumulator-plane index
    const bool out_f32 = p.epi.out_f32 != 0;
    int tw = blockIdx.x % p.tiles_w, rest = blockIdx.x / p.tiles_w;
    for (int col = blockIdx.x; col < p.num_cols; col += gridDim.x) {
      const int th = rest % p.tiles_h, n = rest / p.tiles_h;
      const int ow = tw * p.TWv + r_w, oh = th * p.TH + r_h;
      const bool in_range = r_w < p.TWv && ow < p.W && oh < p.H;
      int64_t vox = (((int64_t)n * p.D) * p.H + oh) * p.W + ow - plane_vox;       // accumulator plane a <-> output plane a - 1
      for (uint32_t a = 0; a < acc_per_col; ++a, vox += plane_vox, rg = kw_next(rg), par ^= 1u) {
        if (par != grp) continue;
        const bool real = a >= 1u && a <= (uint32_t)p.D;
        uint4 rq[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) rq[i] = make_uint4(0u, 0u, 0u, 0u);
        if (RES && in_range && real) {                    // issued before the wait: overlaps the MMAs of this plane
          const uint4* rp = reinterpret_cast<const uint4*>(p.epi.residual + vox * p.epi.res_cstride + p.epi.res_coffset);
          if (aligned32(rp)) { ldg256(rp, rq[0], rq[1]); ldg256(rp + 2, rq[2], rq[3]); }
          else {
#pragma unroll
            for (int i = 0; i < 4; ++i) rq[i] = __ldg(rp + i);
          }
        }
        mbar_wait(smem_u32(&acc_full_bar[rg.b]), rg.ph);
        tcgen05_fence_after();
        const uint32_t taddr = lane_base + rg.b * kKwBlkCols;
        if (real) {
#pragma unroll
          for (int c0 = 0; c0 < CP; c0 += 16) {
            uint32_t q0[16], q1[16], q2[16];
            tmem_ld16(taddr + (uint32_t)c0, q0);
            tmem_ld16(taddr + (uint32_t)(CP + c0), q1);
            tmem_ld16(taddr + (uint32_t)(2 * CP + c0), q2);
            tmem_ld_wait();
            // out[m] = P_0[m] + P_1[m+1] + P_2[m+2]  (rows = lanes; the upper lanes of a tile row are not output columns)
            float v[16];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float2 s1 = make_float2(__shfl_down_sync(0xffffffffu, __uint_as_float(q1[2 * j]), 1),
                                            __shfl_down_sync(0xffffffffu, __uint_as_float(q1[2 * j + 1]), 1));
              const float2 s2 = make_float2(__shfl_down_sync(0xffffffffu, __uint_as_float(q2[2 * j]), 2),
                                            __shfl_down_sync(0xffffffffu, __uint_as_float(q2[2 * j + 1]), 2));
              float2 t = __fadd2_rn(make_float2(__uint_as_float(q0[2 * j]), __uint_as_float(q0[2 * j + 1])), s1);
              t = __fadd2_rn(t, s2);
              t = __ffma2_rn(t, sc[c0 / 2 + j], bi[c0 / 2 + j]);
              if (RES) {                                  // x += r*m1; x = max(x, lo); x += r*m2   (EpiFast semantics)
                const uint4 rv = rq[c0 / 8 + (j >> 2)];   // channels c0 + 2j, c0 + 2j + 1
                const uint32_t w = (j & 3) == 0 ? rv.x : ((j & 3) == 1 ? rv.y : ((j & 3) == 2 ? rv.z : rv.w));
                const float2 r = make_float2(bf16_lo(w), bf16_hi(w));
                t = __ffma2_rn(r, make_float2(m1, m1), t);
                t = make_float2(fmaxf(t.x, lo), fmaxf(t.y, lo));
                t = __ffma2_rn(r, make_float2(m2, m2), t);
              } else {
                t = make_float2(fmaxf(t.x, lo), fmaxf(t.y, lo));
              }
              v[2 * j] = t.x;
              v[2 * j + 1] = t.y;
            }
            if (in_range && !out_f32) {
              uint4* o = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.epi.y) + vox * p.epi.out_cstride +
                                                  p.epi.out_coffset + c0);
              const uint4 o0 = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]),
                                          pack_bf16x2(v[6], v[7]));
              const uint4 o1 = make_uint4(pack_bf16x2(v[8], v[9]), pack_bf16x2(v[10], v[11]), pack_bf16x2(v[12], v[13]),
                                          pack_bf16x2(v[14], v[15]));
              if (aligned32(o)) stg256(o, o0, o1);
              else { o[0] = o0; o[1] = o1; }
            }
            if (in_range && out_f32) {                    // fp32 rows (tests, module boundaries): same arithmetic, no rounding
              float4* o = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.epi.y) + vox * p.epi.out_cstride +
                                                    p.epi.out_coffset + c0);
#pragma unroll
              for (int j = 0; j < 4; ++j) o[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            }
          }
        }
#pragma unroll
        for (uint32_t c = 0; c < kKwBlkCols; c += 16u) tmem_st16_zero(taddr + c);   // ready for its next output plane
        tmem_st_wait();
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&acc_empty_bar[rg.b]));
      }
      const int nc = col + (int)gridDim.x;
      tw = nc % p.tiles_w; rest = nc / p.tiles_w;
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TCOLS) : "memory");
  }
}

// ==========================================================================================
// v8: CTA-pair kd-fused plane march (3x3x3, stride 1, 32-channel output slice) -- tcgen05.mma.cta_group::2.
//
// Measured (scripts/micro/umma_2cta.cu, profiles/r01_umma_2cta.txt): issued over a CTA pair (M = 256, each CTA
// supplies its own 128 A rows and HALF of the B rows), an SS-mode MMA costs max(51.3, N/2) cycles instead of
// max(71.6, N/2) -- N = 96 runs at 94 % of the tensor peak instead of 67 % -- and each SM reads only N/2 weight rows
// per instruction.  That removes both limits of the Cout = 32 layers at once: the issue floor that v7 attacked by
// fusing kw into N (paying 64 shuffles and 3x the TMEM traffic per output row in the epilogue), and the L1 data pipe
// that then bound v7 (operand reads 5.6 KB per 51-cycle MMA instead of 8.7 KB per 72).
//
// Structure: a cluster of two CTAs; each CTA marches its OWN tile column (its own plane ring, TMA loads, TMEM
// accumulators and epilogue, exactly as v3) and holds half of every weight tile: rank r keeps rows [48r, 48r+48) of
// the 96-row [kd=2 | kd=1 | kd=0] slab of each in-plane tap at the same shared-memory offset.  The leader's MMA warp
// issues for both; every TMA load (either CTA) completes on the LEADER's full barrier (cta_group::2 form), commits
// are multicast to both CTAs' barriers, and the follower's epilogue warps release accumulator blocks with remote
// arrives on the leader's barriers.
// Accumulator ring without instruction variants (a split MMA would need differently shifted weight halves): 14 ring
// blocks + 2 MIRROR blocks (positions 14, 15 alias ring indices 0, 1), so the three blocks an input plane updates are
// always the contiguous positions i, i+1, i+2; a plane whose ring index is 0 or 1 may hold partial sums in both its
// primary and its mirror block, and the epilogue adds the two.
// ==========================================================================================
constexpr int kPairThreads = 320;
constexpr uint32_t kPairRing = 14;

struct PairRing { uint32_t i, ph; };
__device__ __forceinline__ PairRing pr_next(PairRing r) {
  PairRing n{r.i + 1u, r.ph};
  if (n.i == kPairRing) { n.i = 0u; n.ph ^= 1u; }
  return n;
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `addr` (a shared::cta address of this CTA's window) in the cluster's rank-0 CTA
__device__ __forceinline__ uint32_t leader_addr(uint32_t addr) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, 0;" : "=r"(r) : "r"(addr));
  return r;
}
// (relaxed: the producer has nothing to publish, and a cluster-scope release is MEMBAR.ALL.GPU + ERRBAR + CGAERRBAR --
// ~1000 cycles per plane in the producer warp, which capped the Cin = 32 layers at 1565 cycles per plane: ncu showed the
// MMA warp waiting for the plane's full barrier and the tensor pipe 55 % busy)
__device__ __forceinline__ void mbar_expect_tx_cluster(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.relaxed.cluster.shared::cluster.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// Relaxed on purpose: a release at cluster scope also waits for the thread's outstanding GLOBAL stores (the output
// rows just written), which put ~700 cycles per plane on the epilogue's critical path (32->32 layers: 631 us with
// .release against 578 us for v7).  What the arrive must order -- the tcgen05.st zero fill of the drained block -- is
// already complete (tcgen05.wait::st) and fenced (tcgen05.fence::before_thread_sync) when the arrive is issued.
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar) {
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tma_load_5d_pair(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2,
                                                 int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void umma_bf16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, 1, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc)
      : "memory");
}
__device__ __forceinline__ void umma_commit_pair(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"((uint16_t)3)
               : "memory");
}

// WPT = row pitch as a template parameter: every operand-descriptor offset of the 9*KSTEPS MMAs of a plane is then an
// immediate added to two uniform registers.  With the pitch a run-time value the 36 descriptors of a Cin = 64 plane
// did not fit the uniform register file; ptxas built them in vector registers and moved them over with R2UR, ~20
// instructions and 70-90 cycles per UTCHMMA.2CTA -- more than the 51-cycle MMA itself (first version: 32->32 layers
// 620 us, slower than v7).
// ADD: a per-(h, w, channel) fp32 addend joins the accumulator before scale / bias -- the contribution of input
// channels that do not vary with depth (the left half of the plane-sweep cost volume), computed once as a 3-plane
// convolution: plane 0 / 1 / 2 of `addend` = the sums an output plane at depth 0 / interior / D-1 needs (the depth
// padding removes one kd tap at either end).  The interior rows stay in registers for the whole column.
// Work units of the CTA-pair kernel.  A unit is a pair of tile columns and a range of output planes.  Whole columns are
// dealt round-robin to the clusters; the columns of the last, partial round (2112 columns on 148 SMs: 14.27 rounds,
// i.e. 5 % of every large layer spent with 108 SMs idle) are cut into `parts` depth ranges so that the round is
// shared by (almost) all clusters.  A range [d0, d1) loads input planes [max(d0-1, 0), min(d1+1, D)) and, as at the
// ends of a whole column, the first and last accumulator planes of the march are not outputs of this unit.
struct PairUnit { int q, d0, d1, ip0, np; };
struct PairSchedule {
  int nclusters, full_units, parts, total, D;
  __device__ __forceinline__ void init(int npairs, int nclusters_, int D_) {
    nclusters = nclusters_; D = D_;
    const int rounds = npairs / nclusters, rem = npairs - rounds * nclusters;
    full_units = rounds * nclusters;
    // depth ranges per leftover pair-column: the count that makes the leftover cheapest, in units of one whole round:
    // ceil(rem * parts / nclusters) rounds of ranges that each cost (D / parts + 2) / D of a column (two extra planes of
    // halo); ranges keep at least 4 output planes.  (20 leftovers on 74 clusters -> 3 ranges, 0.375 round instead of 1;
    // 46 leftovers -> 3 ranges in two rounds, 0.75 instead of 1.)
    parts = 1;
    if (rem > 0) {
      float best = 1e30f;
      for (int c = 1; c <= 4 && D / c >= 4; ++c) {
        const float cost = (float)((rem * c + nclusters - 1) / nclusters) * ((float)(D / c + 2) / (float)D);
        if (cost < best * 0.999f) { best = cost; parts = c; }
      }
    }
    total = full_units + rem * parts;
  }
  __device__ __forceinline__ PairUnit unit(int u) const {
    PairUnit r;
    if (u < full_units) { r.q = u; r.d0 = 0; r.d1 = D; }
    else {
      const int t = u - full_units, part = t % parts;
      r.q = full_units + t / parts;
      r.d0 = (int)((int64_t)D * part / parts); r.d1 = (int)((int64_t)D * (part + 1) / parts);
    }
    r.ip0 = max(r.d0 - 1, 0);
    r.np = min(r.d1 + 1, D) - r.ip0;
    return r;
  }
};

template <int KSTEPS, int SUBROW, bool RES, int WPT, bool ADD>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kPairThreads, 1)
conv3d_kdpair_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w,
                     const __grid_constant__ CUtensorMap map_r, const __grid_constant__ HaloParams p) {
  constexpr int K = 3, CP = 32;
  constexpr uint32_t TCOLS = 512;
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t full_bar[kMaxSlots];
  __shared__ __align__(8) uint64_t empty_bar[kMaxSlots];
  __shared__ __align__(8) uint64_t w_bar;
  __shared__ __align__(8) uint64_t acc_full_bar[kPairRing];
  __shared__ __align__(8) uint64_t acc_empty_bar[kPairRing];
  __shared__ uint32_t tmem_base_smem;
  __shared__ __align__(16) float s_scale[64], s_bias[64];
  // RES: the residual tile of every output plane (the 128 accumulator rows x 64 bytes, same row order as the TMEM lanes)
  // arrives by TMA in a ring of kResSlots swizzled tiles and is read with conflict-free LDS.128.  Fetched by the
  // epilogue threads themselves (two LDG.256 of 32 different rows per warp) the residual rows cost 234 L1 wavefronts
  // per plane on the pipe that also feeds the UMMA operands: ncu had the residual layer at 89 % of that pipe, tensor
  // pipe 65 %, 0.56 ms against 0.43 ms for the same layer without a residual.
  constexpr uint32_t kResSlots = 4, kResTile = 128u * 64u;
  __shared__ __align__(8) uint64_t res_full_bar[kResSlots];
  __shared__ __align__(8) uint64_t res_empty_bar[kResSlots];

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const uint32_t w_base = (smem_u32(smem) + 1023u) & ~1023u;
  const uint32_t res_base = w_base + (((uint32_t)(K * K * p.w_tap_bytes) + 1023u) & ~1023u);
  const uint32_t slots_base = res_base + (RES ? kResSlots * kResTile : 0u);
  const int npairs = (p.num_cols + 1) >> 1;
  const int pair0 = (int)(blockIdx.x >> 1), pstep = (int)(gridDim.x >> 1);
  PairSchedule sched;
  sched.init(npairs, pstep, p.D);

  if (threadIdx.x < 64) {
    s_scale[threadIdx.x] = (p.scale && threadIdx.x < p.epi.Cout) ? p.scale[threadIdx.x] : 1.f;
    s_bias[threadIdx.x] = (p.bias && threadIdx.x < p.epi.Cout) ? p.bias[threadIdx.x] : 0.f;
  }
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_x) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
    if (RES) asm volatile("prefetch.tensormap [%0];" ::"l"(&map_r) : "memory");
    for (int s = 0; s < p.nslots; ++s) {
      mbar_init(smem_u32(&full_bar[s]), 2);               // leader's copy: one arrive + tx per CTA
      mbar_init(smem_u32(&empty_bar[s]), 1);
    }
    mbar_init(smem_u32(&w_bar), 2);
    for (uint32_t b = 0; b < kResSlots; ++b) {
      mbar_init(smem_u32(&res_full_bar[b]), 1);
      mbar_init(smem_u32(&res_empty_bar[b]), 4);          // the four warps of the group that drains the plane
    }
    for (uint32_t b = 0; b < kPairRing; ++b) {
      mbar_init(smem_u32(&acc_full_bar[b]), 1);
      mbar_init(smem_u32(&acc_empty_bar[b]), 8);          // leader's copy: one arrive per epilogue warp of the draining group, both CTAs
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)),
                 "r"(TCOLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  if (warp >= 2 && warp < 6) {                            // zero this CTA's whole accumulator ring once
    const uint32_t lane_base = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    for (uint32_t c = 0; c < TCOLS; c += 16u) tmem_st16_zero(lane_base + c);
    tmem_st_wait();
  }
  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();                                     // both CTAs: barriers initialised, TMEM allocated and zeroed
  tcgen05_fence_after();

  if (warp == 0) {
    // ===================== TMA producer (both CTAs; every load completes on the LEADER's barrier) =====================
    const uint32_t wb = leader_addr(smem_u32(&w_bar));
    if (elect_one()) {
      mbar_expect_tx_cluster(wb, (uint32_t)(K * K * p.w_tap_bytes));
      // this CTA's half of every tap's 96-row slab [kd=2 | kd=1 | kd=0]: slab rows [48*rank, 48*rank + 48), 16 at a time
      for (int t2 = 0; t2 < K * K; ++t2)
        for (int j = 0; j < 3; ++j) {
          const int srow = 48 * (int)rank + 16 * j;
          const int kd = 2 - (srow >> 5), c0 = srow & 31;
          tma_load_2d_pair(w_base + t2 * p.w_tap_bytes + j * 16 * SUBROW, &map_w, wb, 0,
                           (kd * K * K + t2) * p.w_rows_per_tap + p.w_row0 + c0);
        }
    }
    __syncwarp();
    uint32_t slot = 0, phase = 0;
    uint32_t slot_addr = slots_base;
    uint32_t rslot = 0, rphase = 0;
    for (int u = pair0; u < sched.total; u += pstep) {
      const PairUnit un = sched.unit(u);
      const int col = min(2 * un.q + (int)rank, p.num_cols - 1);        // (odd column count: the last follower re-reads a column)
      const int tw = col % p.tiles_w, rest = col / p.tiles_w;
      const int th = rest % p.tiles_h, n = rest / p.tiles_h;
      const int w0 = tw * p.TWv - p.pad, h0 = th * p.TH - p.pad;
      for (int ip = un.ip0; ip < un.ip0 + un.np; ++ip) {
        mbar_wait(smem_u32(&empty_bar[slot]), phase ^ 1u);
        if (elect_one()) {
          const uint32_t fb = leader_addr(smem_u32(&full_bar[slot]));
          mbar_expect_tx_cluster(fb, (uint32_t)p.plane_bytes);
          tma_load_5d_pair(slot_addr, &map_x, fb, 0, w0, h0, ip, n);
        }
        __syncwarp();
        slot_addr += (uint32_t)p.slot_bytes;
        if (++slot == (uint32_t)p.nslots) { slot = 0; phase ^= 1u; slot_addr = slots_base; }
        if (RES && ip >= un.d0 && ip < un.d1) {           // residual tile of OUTPUT plane ip (this CTA's own barriers)
          mbar_wait(smem_u32(&res_empty_bar[rslot]), rphase ^ 1u);
          if (elect_one()) {
            const uint32_t rb = smem_u32(&res_full_bar[rslot]);
            mbar_expect_tx(rb, (uint32_t)(p.WP * p.TH * 64));       // the box: WP x TH rows (126 of the 128 for pitch 42)
            tma_load_5d(res_base + rslot * kResTile, &map_r, rb, 0, tw * p.TWv, th * p.TH, ip, n);
          }
          __syncwarp();
          if (++rslot == kResSlots) { rslot = 0; rphase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (rank == 0) {
      constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)((3 * CP) >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
      const uint32_t desc_hi = (uint32_t)(make_smem_desc(0, SUBROW) >> 32);
      constexpr uint32_t lo_flags = 1u << 16;
      const uint32_t b_lo0 = ((w_base >> 4) & 0x3FFFu) | lo_flags;
      constexpr uint32_t b_tap = (uint32_t)(48 * SUBROW) >> 4;
      const uint32_t a_lo0 = ((slots_base >> 4) & 0x3FFFu) | lo_flags;
      const uint32_t a_step = (uint32_t)p.slot_bytes >> 4;
      mbar_wait(smem_u32(&w_bar), 0);
      uint32_t slot = 0, phase = 0, a_plane = a_lo0;
      PairRing r0{0u, 0u};                                // ring position of accumulator plane g = out[pl-1]
      for (int u = pair0; u < sched.total; u += pstep) {
        const int np = sched.unit(u).np;
        for (int pl = 0; pl < np; ++pl) {
          const PairRing r1 = pr_next(r0), r2 = pr_next(r1);
          mbar_wait(smem_u32(&full_bar[slot]), phase);
          if (pl == 0) {
            mbar_wait(smem_u32(&acc_empty_bar[r0.i]), r0.ph ^ 1u);
            mbar_wait(smem_u32(&acc_empty_bar[r1.i]), r1.ph ^ 1u);
          }
          mbar_wait(smem_u32(&acc_empty_bar[r2.i]), r2.ph ^ 1u);
          tcgen05_fence_after();
          const uint32_t d0 = tmem_base + r0.i * (uint32_t)CP;          // positions i, i+1, i+2 (14, 15 = mirrors of 0, 1)
          if (elect_one()) {
#pragma unroll
            for (int t2 = 0; t2 < K * K; ++t2)
#pragma unroll
              for (int k = 0; k < KSTEPS; ++k)
                umma_bf16_pair(d0, desc64(desc_hi, a_plane + (uint32_t)((((t2 / K) * WPT + (t2 % K)) * SUBROW) >> 4) + 2u * k),
                               desc64(desc_hi, b_lo0 + (uint32_t)t2 * b_tap + 2u * k), idesc);
            umma_commit_pair(smem_u32(&empty_bar[slot]));                 // plane consumed (both CTAs)
            umma_commit_pair(smem_u32(&acc_full_bar[r0.i]));              // out[pl-1] complete
            if (pl == np - 1) {                                           // end of the march: the last two accumulator planes
              umma_commit_pair(smem_u32(&acc_full_bar[r1.i]));
              umma_commit_pair(smem_u32(&acc_full_bar[r2.i]));
            }
          }
          __syncwarp();
          a_plane += a_step;
          if (++slot == (uint32_t)p.nslots) { slot = 0; phase ^= 1u; a_plane = a_lo0; }
          r0 = r1;
        }
        r0 = pr_next(pr_next(r0));                        // np + 2 accumulator planes per unit
      }
    }
  } else {
    // ===================== epilogue (warps 2..9 of both CTAs; two groups drain alternate planes) =====================
    const int quad = warp & 3;
    const uint32_t grp = (uint32_t)(warp - 2) >> 2;
    const int row = quad * 32 + lane;
    const int r_w = row % p.WP, r_h = row / p.WP;
    const float m1 = p.epi.residual_mode == 1 ? 1.f : 0.f, m2 = p.epi.residual_mode == 2 ? 1.f : 0.f;
    const float lo = p.epi.relu ? 0.f : -INFINITY;
    float2 sc[CP / 2], bi[CP / 2];
#pragma unroll
    for (int j = 0; j < CP / 2; ++j) {
      sc[j] = make_float2(s_scale[2 * j], s_scale[2 * j + 1]);
      bi[j] = make_float2(s_bias[2 * j], s_bias[2 * j + 1]);
    }
    const uint32_t lane_base = tmem_base + ((uint32_t)(quad * 32) << 16);
    const uint32_t empty0 = leader_addr(smem_u32(&acc_empty_bar[0]));
    const int64_t plane_vox = (int64_t)p.H * p.W;
    PairRing rg{0u, 0u};
    uint32_t par = 0;
    const bool out_f32 = p.epi.out_f32 != 0;
    uint32_t rslot = 0, rphase = 0;                       // residual ring position of the next REAL plane (both groups count all)
    for (int u = pair0; u < sched.total; u += pstep) {
      const PairUnit un = sched.unit(u);
      const uint32_t acc_per_unit = (uint32_t)un.np + 2u;
      const int col = 2 * un.q + (int)rank;
      const bool ghost = col >= p.num_cols;               // odd column count: the last follower's results are dropped
      const int colc = ghost ? p.num_cols - 1 : col;
      const int tw = colc % p.tiles_w, rest = colc / p.tiles_w;
      const int th = rest % p.tiles_h, n = rest / p.tiles_h;
      const int ow = tw * p.TWv + r_w, oh = th * p.TH + r_h;
      const bool in_range = !ghost && r_w < p.TWv && r_h < p.TH && ow < p.W && oh < p.H;   // (r_h < TH: 126-row tiles, pitch 42)
      int64_t vox = (((int64_t)n * p.D + un.ip0) * p.H + oh) * p.W + ow - plane_vox;   // accumulator plane a <-> output plane ip0 + a - 1
      float addm[ADD ? CP : 1];
      const float* arow = nullptr;                        // this thread's row of addend plane 0 (planes are plane_vox*CP apart)
      if (ADD) {
        arow = p.addend + ((((int64_t)n * 3) * p.H + oh) * p.W + ow) * CP;
#pragma unroll
        for (int j = 0; j < CP; j += 4) {
          const float4 t = in_range ? __ldg(reinterpret_cast<const float4*>(arow + plane_vox * CP + j)) : make_float4(0.f, 0.f, 0.f, 0.f);
          addm[j] = t.x; addm[j + 1] = t.y; addm[j + 2] = t.z; addm[j + 3] = t.w;
        }
      }
      for (uint32_t a = 0; a < acc_per_unit; ++a, vox += plane_vox, rg = pr_next(rg), par ^= 1u) {
        const int od = un.ip0 + (int)a - 1;               // output plane of this accumulator plane
        const bool real = od >= un.d0 && od < un.d1;
        const uint32_t my_rslot = rslot, my_rphase = rphase;
        if (RES && real) { if (++rslot == kResSlots) { rslot = 0; rphase ^= 1u; } }
        if (par != grp) continue;
        uint4 rq[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) rq[i] = make_uint4(0u, 0u, 0u, 0u);
        if (RES && real) {                                // this thread's row of the plane's residual tile
          mbar_wait(smem_u32(&res_full_bar[my_rslot]), my_rphase);
          const uint32_t tile = res_base + my_rslot * kResTile;
#pragma unroll
          for (int i = 0; i < 4; ++i) rq[i] = lds_v4(tile + swz<64>((uint32_t)row, (uint32_t)i));
        }
        mbar_wait(smem_u32(&acc_full_bar[rg.i]), rg.ph);
        tcgen05_fence_after();
        const uint32_t taddr = lane_base + rg.i * (uint32_t)CP;
        const bool mirrored = rg.i < 2u;                  // partial sums may also sit in the mirror block (position 14 + i)
        const uint32_t maddr = lane_base + (kPairRing + rg.i) * (uint32_t)CP;
        if (real) {
#pragma unroll
          for (int c0 = 0; c0 < CP; c0 += 16) {
            uint32_t q0[16];
            tmem_ld16(taddr + (uint32_t)c0, q0);
            if (mirrored) {
              uint32_t q1[16];
              tmem_ld16(maddr + (uint32_t)c0, q1);
              tmem_ld_wait();
#pragma unroll
              for (int j = 0; j < 16; ++j) q0[j] = __float_as_uint(__uint_as_float(q0[j]) + __uint_as_float(q1[j]));
            } else {
              tmem_ld_wait();
            }
            if (ADD) {
              const bool edge = od == p.add_lo || od == p.add_hi;      // output plane 0 / D-1 of the VOLUME: their own addend planes
              const float* ep = arow + (od == p.add_lo ? 0 : 2 * plane_vox * CP) + c0;
#pragma unroll
              for (int j = 0; j < 16; j += 4) {
                float4 t = make_float4(addm[c0 + j], addm[c0 + j + 1], addm[c0 + j + 2], addm[c0 + j + 3]);
                if (edge) t = in_range ? __ldg(reinterpret_cast<const float4*>(ep + j)) : make_float4(0.f, 0.f, 0.f, 0.f);
                q0[j] = __float_as_uint(__uint_as_float(q0[j]) + t.x);
                q0[j + 1] = __float_as_uint(__uint_as_float(q0[j + 1]) + t.y);
                q0[j + 2] = __float_as_uint(__uint_as_float(q0[j + 2]) + t.z);
                q0[j + 3] = __float_as_uint(__uint_as_float(q0[j + 3]) + t.w);
              }
            }
            float v[16];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              float2 t = __ffma2_rn(make_float2(__uint_as_float(q0[2 * j]), __uint_as_float(q0[2 * j + 1])), sc[c0 / 2 + j],
                                    bi[c0 / 2 + j]);
              if (RES) {                                  // x += r*m1; x = max(x, lo); x += r*m2   (EpiFast semantics)
                const uint4 rv = rq[c0 / 8 + (j >> 2)];   // channels c0 + 2j, c0 + 2j + 1
                const uint32_t w = (j & 3) == 0 ? rv.x : ((j & 3) == 1 ? rv.y : ((j & 3) == 2 ? rv.z : rv.w));
                const float2 r = make_float2(bf16_lo(w), bf16_hi(w));
                t = __ffma2_rn(r, make_float2(m1, m1), t);
                t = make_float2(fmaxf(t.x, lo), fmaxf(t.y, lo));
                t = __ffma2_rn(r, make_float2(m2, m2), t);
              } else {
                t = make_float2(fmaxf(t.x, lo), fmaxf(t.y, lo));
              }
              v[2 * j] = t.x;
              v[2 * j + 1] = t.y;
            }
            if (in_range && !out_f32) {
              uint4* o = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.epi.y) + vox * p.epi.out_cstride +
                                                  p.epi.out_coffset + c0);
              const uint4 o0 = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]),
                                          pack_bf16x2(v[6], v[7]));
              const uint4 o1 = make_uint4(pack_bf16x2(v[8], v[9]), pack_bf16x2(v[10], v[11]), pack_bf16x2(v[12], v[13]),
                                          pack_bf16x2(v[14], v[15]));
              if (aligned32(o)) stg256(o, o0, o1);
              else { o[0] = o0; o[1] = o1; }
            }
            if (in_range && out_f32) {                    // fp32 rows (tests, module boundaries): same arithmetic, no rounding
              float4* o = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.epi.y) + vox * p.epi.out_cstride +
                                                    p.epi.out_coffset + c0);
#pragma unroll
              for (int j = 0; j < 4; ++j) o[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            }
          }
        }
#pragma unroll
        for (uint32_t c = 0; c < (uint32_t)CP; c += 16u) tmem_st16_zero(taddr + c);   // ready for its next output plane
        if (mirrored) {
#pragma unroll
          for (uint32_t c = 0; c < (uint32_t)CP; c += 16u) tmem_st16_zero(maddr + c);
        }
        tmem_st_wait();
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(empty0 + rg.i * 8u);
        // The residual slot is released only here, after every lane has USED its rows: an arrive issued right behind the
        // LDS can overtake them (neither the arrive's release nor __syncwarp waits for another lane's outstanding
        // shared-memory loads), and a producer that is blocked on exactly this slot -- as it is for the first tiles of a
        // launch -- then overwrites rows that are still being read (seen as stale 1 KB box rows on cold first launches).
        if (RES && real && lane == 0) mbar_arrive(smem_u32(&res_empty_bar[my_rslot]));
      }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();                                     // the peer may still be reading / signalling this CTA
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TCOLS) : "memory");
  }
}

// ==========================================================================================
// v4: fused transposed convolution (k3, s2, p1, output_padding 1) -- all 8 output-parity classes
// of an input tile in ONE kernel (the per-tap path above launches 8 kernels, each re-reading the
// input through per-tap TMA boxes and writing a stride-2 quarter of the output).
//
//   out[2j+p] (per dim) = p == 0 ?  x[j] * W[1]  :  x[j] * W[2] + x[j+1] * W[0]
//
// A CTA marches along the INPUT depth of a (TH x TWv) patch; plane j (and j+1) sit in the same
// kind of dense-row shared-memory ring as the plane-march conv kernels, so the +1 shifts in h / w
// are UMMA descriptor row offsets (TWv = WP - 1 valid columns) and the +1 shift in d is "the next
// ring slot".  For a shift s = (sd,sh,sw) the classes p >= s (componentwise) all read the same A
// window; classes that are adjacent in TMEM (class c = pd*4 + ph*2 + pw at columns c*32) are fused
// into one instruction: N = 256 (s=000), 128 (s=100), 64, 32 ... -- 14 MMAs per K=16 step instead of
// 27.  All 27 weight tiles (32 output channels) stay resident; 2 x 256 TMEM columns double-buffer
// the 8 class accumulators so the epilogue of tile-step j overlaps the MMAs of j+1.  Eight
// epilogue warps (two per TMEM lane quadrant, four classes each) prefetch their residual rows
// before waiting for the accumulator and write each output voxel row (64 B) exactly once.
// Cout = 64 runs as two 32-channel output slices (weights 2 x 110 KB).
// ==========================================================================================
constexpr int kDeconvThreads = 384;      // producer, MMA issuer, 8 epilogue warps, 2 staged-tile managers
constexpr int kDeconvCP = 32;

struct DeconvParams {
  int N, Cin;
  int Di, Hi, Wi;              // input extent (output is 2x)
  int WP, TH, TWv;             // row pitch, tile rows (WP*TH == 128), valid columns = WP - 1
  int tiles_h, tiles_w;
  int DC, nchunk;              // depth chunk per work unit
  int num_units;
  int plane_bytes, slot_bytes, nslots;
  int w_tap_bytes;             // 32 * Cin * 2
  int w_rows_per_tap, w_row0;
  int stage_bytes;             // one staged output tile: TH x 2*TWv voxel rows of 64 B (4 tiles follow the weights)
  const float* scale;
  const float* bias;
  EpiParams epi;
};

// static fusion schedule: shift s = sd*4 + sh*2 + sw; classes(s) = {c : (c & s) == s} ascending
struct DeconvRun { unsigned char s, c0, len, tile0; };
__device__ constexpr DeconvRun kDeconvRuns[14] = {
    {0, 0, 8, 0},
    {1, 1, 1, 8}, {1, 3, 1, 9}, {1, 5, 1, 10}, {1, 7, 1, 11},
    {2, 2, 2, 12}, {2, 6, 2, 14},
    {3, 3, 1, 16}, {3, 7, 1, 17},
    {4, 4, 4, 18},
    {5, 5, 1, 22}, {5, 7, 1, 23},
    {6, 6, 2, 24},
    {7, 7, 1, 26}};
// smem tile t -> (s, c); kernel tap per dim: shift 1 -> k = 0; shift 0 -> k = (class bit ? 2 : 1)
__device__ constexpr unsigned char kDeconvTileS[27] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3,
                                                       4, 4, 4, 4, 5, 5, 6, 6, 7};
__device__ constexpr unsigned char kDeconvTileC[27] = {0, 1, 2, 3, 4, 5, 6, 7, 1, 3, 5, 7, 2, 3, 6, 7, 3, 7,
                                                       4, 5, 6, 7, 5, 7, 6, 7, 7};
__device__ __forceinline__ int deconv_tap_of_tile(int t) {
  const int s = kDeconvTileS[t], c = kDeconvTileC[t];
  int tap = 0;
#pragma unroll
  for (int dim = 2; dim >= 0; --dim) {        // dim 2 = d (bit 2), 1 = h, 0 = w
    const int sb = (s >> dim) & 1, cb = (c >> dim) & 1;
    tap = tap * 3 + (sb ? 0 : (cb ? 2 : 1));
  }
  return tap;
}

struct ResidualRow32 { uint4 q[4]; };

__device__ __forceinline__ void epilogue_chunk16_r32_vals(const uint32_t* acc, int cc, const float* s_scale,
                                                          const float* s_bias, const ResidualRow32& rr, const EpiFast f,
                                                          uint4& o0, uint4& o1) {
  float v[16];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const float4 sc = *reinterpret_cast<const float4*>(s_scale + cc + 4 * q);
    const float4 bi = *reinterpret_cast<const float4*>(s_bias + cc + 4 * q);
    v[4 * q + 0] = fmaf(__uint_as_float(acc[4 * q + 0]), sc.x, bi.x);
    v[4 * q + 1] = fmaf(__uint_as_float(acc[4 * q + 1]), sc.y, bi.y);
    v[4 * q + 2] = fmaf(__uint_as_float(acc[4 * q + 2]), sc.z, bi.z);
    v[4 * q + 3] = fmaf(__uint_as_float(acc[4 * q + 3]), sc.w, bi.w);
  }
  const uint4 q0 = rr.q[cc >> 3], q1 = rr.q[(cc >> 3) + 1];
  const uint32_t w[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float r0 = bf16_lo(w[j]), r1 = bf16_hi(w[j]);
    v[2 * j] = fmaf(r0, f.m2, fmaxf(fmaf(r0, f.m1, v[2 * j]), f.lo));
    v[2 * j + 1] = fmaf(r1, f.m2, fmaxf(fmaf(r1, f.m1, v[2 * j + 1]), f.lo));
  }
  o0 = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
  o1 = make_uint4(pack_bf16x2(v[8], v[9]), pack_bf16x2(v[10], v[11]), pack_bf16x2(v[12], v[13]), pack_bf16x2(v[14], v[15]));
}

template <int KSTEPS, int SUBROW>
__global__ void __launch_bounds__(kDeconvThreads, 1)
conv3d_deconv_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w,
                     const __grid_constant__ CUtensorMap map_y, const __grid_constant__ CUtensorMap map_r,
                     const __grid_constant__ DeconvParams p) {
  constexpr int CP = kDeconvCP;
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t full_bar[kMaxSlots];
  __shared__ __align__(8) uint64_t empty_bar[kMaxSlots];
  __shared__ __align__(8) uint64_t w_bar;
  __shared__ __align__(8) uint64_t tmem_full_bar[2];
  __shared__ __align__(8) uint64_t tmem_empty_bar[2];
  __shared__ __align__(8) uint64_t tile_bar[4];            // staged tile b holds its residual (or is simply free again)
  __shared__ __align__(8) uint64_t ready_bar[4];           // staged tile b has been computed by its four warps
  __shared__ uint32_t tmem_base_smem;
  __shared__ __align__(16) float s_scale[32], s_bias[32];

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t w_base = (smem_u32(smem) + 1023u) & ~1023u;
  const uint32_t stage_base = w_base + (((uint32_t)(27 * p.w_tap_bytes) + 1023u) & ~1023u);
  const uint32_t slots_base = stage_base + 4u * (uint32_t)p.stage_bytes;

  if (threadIdx.x < 32) {
    s_scale[threadIdx.x] = p.scale ? p.scale[threadIdx.x] : 1.f;
    s_bias[threadIdx.x] = p.bias ? p.bias[threadIdx.x] : 0.f;
  }
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_x) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_y) : "memory");
    if (p.epi.residual_mode) asm volatile("prefetch.tensormap [%0];" ::"l"(&map_r) : "memory");
    for (int b = 0; b < 4; ++b) {
      mbar_init(smem_u32(&tile_bar[b]), 1);
      mbar_init(smem_u32(&ready_bar[b]), 4);               // one arrive per epilogue warp of the group
    }
    for (int s = 0; s < p.nslots; ++s) {
      mbar_init(smem_u32(&full_bar[s]), 1);
      mbar_init(smem_u32(&empty_bar[s]), 1);
    }
    mbar_init(smem_u32(&w_bar), 1);
    for (int b = 0; b < 2; ++b) {
      mbar_init(smem_u32(&tmem_full_bar[b]), 1);
      mbar_init(smem_u32(&tmem_empty_bar[b]), 8);          // one arrive per epilogue warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)),
                 "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = tmem_base_smem;

  // work unit -> (n, th, tw, depth chunk); planes j0 .. j0+nj-1 are tile-steps, plane j0+nj is loaded if it exists
  auto decode = [&](int unit, int& n, int& h0, int& w0, int& j0, int& nj) {
    const int ch = unit % p.nchunk; unit /= p.nchunk;
    const int tw = unit % p.tiles_w; unit /= p.tiles_w;
    const int th = unit % p.tiles_h; n = unit / p.tiles_h;
    h0 = th * p.TH; w0 = tw * p.TWv; j0 = ch * p.DC; nj = min(p.DC, p.Di - j0);
  };

  if (warp == 0) {
    // ===================== TMA producer =====================
    const uint32_t wb = smem_u32(&w_bar);
    if (elect_one()) {
      mbar_expect_tx(wb, (uint32_t)(27 * p.w_tap_bytes));
      for (int t = 0; t < 27; ++t)
        tma_load_2d(w_base + t * p.w_tap_bytes, &map_w, wb, 0, deconv_tap_of_tile(t) * p.w_rows_per_tap + p.w_row0);
    }
    __syncwarp();
    uint32_t slot = 0, phase = 0, slot_addr = slots_base;
    for (int unit = blockIdx.x; unit < p.num_units; unit += gridDim.x) {
      int n, h0, w0, j0, nj;
      decode(unit, n, h0, w0, j0, nj);
      const int nload = nj + ((j0 + nj < p.Di) ? 1 : 0);
      for (int i = 0; i < nload; ++i) {
        mbar_wait(smem_u32(&empty_bar[slot]), phase ^ 1u);
        if (elect_one()) {
          const uint32_t fb = smem_u32(&full_bar[slot]);
          mbar_expect_tx(fb, (uint32_t)p.plane_bytes);
          tma_load_5d(slot_addr, &map_x, fb, 0, w0, h0, j0 + i, n);
        }
        __syncwarp();
        slot_addr += (uint32_t)p.slot_bytes;
        if (++slot == (uint32_t)p.nslots) { slot = 0; phase ^= 1u; slot_addr = slots_base; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    const uint32_t desc_hi = (uint32_t)(make_smem_desc(0, SUBROW) >> 32);
    constexpr uint32_t lo_flags = 1u << 16;
    const uint32_t b_lo0 = ((w_base >> 4) & 0x3FFFu) | lo_flags;
    constexpr uint32_t b_tap = (uint32_t)(CP * KSTEPS * 32) >> 4;
    const uint32_t a_lo0 = ((slots_base >> 4) & 0x3FFFu) | lo_flags;
    const uint32_t a_step = (uint32_t)p.slot_bytes >> 4;
    const uint32_t row16 = (uint32_t)SUBROW >> 4;                 // one voxel row in 16-byte units
    const uint32_t sh_off = (uint32_t)p.WP * row16, sw_off = row16;
    mbar_wait(smem_u32(&w_bar), 0);
    uint32_t slot = 0, phase = 0, a_cur = a_lo0;
    uint32_t it = 0;
    for (int unit = blockIdx.x; unit < p.num_units; unit += gridDim.x) {
      int n, h0, w0, j0, nj;
      decode(unit, n, h0, w0, j0, nj);
      const bool tail_plane = j0 + nj < p.Di;                    // an extra plane follows the last tile-step
      mbar_wait(smem_u32(&full_bar[slot]), phase);               // plane j0
      for (int i = 0; i < nj; ++i, ++it) {
        const bool has_next = (i + 1 < nj) || tail_plane;
        uint32_t nslot = slot + 1, nphase = phase, a_next = a_cur + a_step;
        if (nslot == (uint32_t)p.nslots) { nslot = 0; nphase ^= 1u; a_next = a_lo0; }
        if (has_next) mbar_wait(smem_u32(&full_bar[nslot]), nphase);
        const uint32_t buf = it & 1u;
        mbar_wait(smem_u32(&tmem_empty_bar[buf]), ((it >> 1) & 1u) ^ 1u);
        tcgen05_fence_after();
        const uint32_t d_tmem = tmem_base + buf * 256u;
        if (elect_one()) {
#pragma unroll
          for (int r = 0; r < 14; ++r) {
            const DeconvRun run = kDeconvRuns[r];
            const bool sd = (run.s & 4) != 0;
            if (sd && !has_next) continue;                       // plane j+1 is beyond the volume: zero contribution
            const uint32_t a_base = (sd ? a_next : a_cur) + ((run.s & 2) ? sh_off : 0u) + ((run.s & 1) ? sw_off : 0u);
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)((run.len * CP) >> 3) << 17) |
                                   ((uint32_t)(kTileM >> 4) << 24);
#pragma unroll
            for (int k = 0; k < KSTEPS; ++k)
              umma_bf16(d_tmem + (uint32_t)run.c0 * CP, desc64(desc_hi, a_base + 2u * k),
                        desc64(desc_hi, b_lo0 + (uint32_t)run.tile0 * b_tap + 2u * k), idesc, (r | k) ? 1u : 0u);
          }
          umma_commit(smem_u32(&tmem_full_bar[buf]));
          umma_commit(smem_u32(&empty_bar[slot]));               // plane j is not needed by later tile-steps
          if (i == nj - 1 && tail_plane) umma_commit(smem_u32(&empty_bar[nslot]));
        }
        __syncwarp();
        slot = nslot; phase = nphase; a_cur = a_next;
      }
      if (tail_plane) {                                          // skip the extra plane's slot
        a_cur += a_step;
        if (++slot == (uint32_t)p.nslots) { slot = 0; phase ^= 1u; a_cur = a_lo0; }
      }
    }
  } else {
    // ===================== epilogue (warps 2..9) + staged-tile managers (warps 10, 11) =====================
    // Two groups of four epilogue warps; group g owns the output depth parity pd = g (classes 4g .. 4g+3).  A
    // direct store would have every lane write 16 B of its own voxel row with 128 B between lanes (stride-2
    // output): 32 L1 wavefronts per instruction, and ncu showed exactly that -- LSU wavefronts 72 % of the cycles,
    // tensor pipe 9 % (profiles/r01_step_v5_summary.txt).  Instead, per (tile-step, ph) the group works IN PLACE on
    // a staged tile of TH x 2*TWv output voxel rows (both pw classes interleaved = contiguous in W), swizzled
    // SWIZZLE_64B: the residual rows arrive by one TMA tensor load, each thread updates its two rows with
    // conflict-free LDS/STS.128, one TMA tensor store writes the tile (image edges are clipped by the TMA unit).
    // Each group double-buffers its tile.  The loads / stores are issued by one manager thread per group (its own
    // warp, so that waiting for "the store has read the tile" never stalls an epilogue warp); epilogue warps and
    // manager talk through two mbarriers per tile (armed: residual landed or tile free; ready: tile computed).
    const bool manager = warp >= 10;
    const int quad = warp & 3;
    const int grp = manager ? warp - 10 : (warp - 2) >> 2;        // output depth parity handled by this group
    const int row = quad * 32 + lane;
    const int r_w = row % p.WP, r_h = row / p.WP;
    const bool has_res = p.epi.residual_mode != 0;
    const int Do = 2 * p.Di;
    const uint32_t tile0 = stage_base + (uint32_t)(2 * grp) * (uint32_t)p.stage_bytes;
    if (manager) {
      if (lane == 0) {
        // hand tile `b` to item (n, h0, w0, j, ph): residual load, or a plain "tile is free" arrive
        auto arm_tile = [&](uint32_t b, bool valid_item, int n, int h0, int w0, int j, int ph) {
          const uint32_t bar = smem_u32(&tile_bar[2 * grp + b]);
          if (valid_item && has_res) {
            mbar_expect_tx(bar, (uint32_t)p.stage_bytes);
            tma_load_5d(tile0 + b * (uint32_t)p.stage_bytes, &map_r, bar, 0, 2 * w0, ph, h0, n * Do + 2 * j + grp);
          } else {
            mbar_arrive(bar);
          }
        };
        uint32_t item = 0;
        int unit = blockIdx.x;
        if (unit < p.num_units) {                            // prologue: the first tile-step's two items
          int n, h0, w0, j0, nj;
          decode(unit, n, h0, w0, j0, nj);
          arm_tile(0u, true, n, h0, w0, j0, 0);
          arm_tile(1u, true, n, h0, w0, j0, 1);
        }
        for (; unit < p.num_units; unit += gridDim.x) {
          int n, h0, w0, j0, nj;
          decode(unit, n, h0, w0, j0, nj);
          for (int i = 0; i < nj; ++i) {
            // the tile-step after this one (what the freed tiles are armed for)
            int nn = n, nh0 = h0, nw0 = w0, j_next = j0 + i + 1;
            bool next_valid = true;
            if (i + 1 >= nj) {
              const int nu = unit + (int)gridDim.x;
              next_valid = nu < p.num_units;
              if (next_valid) { int t1; decode(nu, nn, nh0, nw0, j_next, t1); }
            }
            for (int ph = 0; ph < 2; ++ph, ++item) {
              const uint32_t b = item & 1u;
              mbar_wait(smem_u32(&ready_bar[2 * grp + b]), (item >> 1) & 1u);
              tma_store_5d(&map_y, tile0 + b * (uint32_t)p.stage_bytes, 0, 2 * w0, ph, h0, n * Do + 2 * (j0 + i) + grp);
              tma_store_commit();
              tma_store_wait_read0();                        // the tile has been read: re-arm it for item + 2
              arm_tile(b, next_valid, nn, nh0, nw0, j_next, ph);
            }
          }
        }
        tma_store_wait_all();
      }
    } else {
      EpiFast f;
      f.m1 = p.epi.residual_mode == 1 ? 1.f : 0.f;
      f.m2 = p.epi.residual_mode == 2 ? 1.f : 0.f;
      f.lo = p.epi.relu ? 0.f : -INFINITY;
      const uint32_t lane_base = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(grp * 4 * CP);
      const uint32_t rho0 = (uint32_t)(r_h * 2 * p.TWv + 2 * r_w);   // this thread's pw = 0 row in the staged tile
      uint32_t it = 0, item = 0;                     // tile-steps / staged items processed by this group
      for (int unit = blockIdx.x; unit < p.num_units; unit += gridDim.x) {
        int n, h0, w0, j0, nj;
        decode(unit, n, h0, w0, j0, nj);
        for (int i = 0; i < nj; ++i, ++it) {
          const uint32_t abuf = it & 1u;
#pragma unroll
          for (int ph = 0; ph < 2; ++ph, ++item) {
            const uint32_t b = item & 1u;
            const uint32_t tile = tile0 + b * (uint32_t)p.stage_bytes;
            mbar_wait(smem_u32(&tile_bar[2 * grp + b]), (item >> 1) & 1u);          // residual landed / tile free
            if (ph == 0) {
              mbar_wait(smem_u32(&tmem_full_bar[abuf]), (it >> 1) & 1u);
              tcgen05_fence_after();
            }
#pragma unroll
            for (int pw = 0; pw < 2; ++pw) {
              uint32_t acc[32];
              const uint32_t taddr = lane_base + abuf * 256u + (uint32_t)((ph * 2 + pw) * CP);
              tmem_ld16(taddr, *reinterpret_cast<uint32_t(*)[16]>(acc));
              tmem_ld16(taddr + 16u, *reinterpret_cast<uint32_t(*)[16]>(acc + 16));
              tmem_ld_wait();
              if (r_w < p.TWv) {
                const uint32_t rho = rho0 + (uint32_t)pw;
                ResidualRow32 rr;
#pragma unroll
                for (int q = 0; q < 4; ++q)
                  rr.q[q] = has_res ? lds_v4(tile + swz<64>(rho, (uint32_t)q)) : make_uint4(0u, 0u, 0u, 0u);
                uint4 o0, o1, o2, o3;
                epilogue_chunk16_r32_vals(acc, 0, s_scale, s_bias, rr, f, o0, o1);
                epilogue_chunk16_r32_vals(acc + 16, 16, s_scale, s_bias, rr, f, o2, o3);
                sts_v4(tile + swz<64>(rho, 0u), o0);
                sts_v4(tile + swz<64>(rho, 1u), o1);
                sts_v4(tile + swz<64>(rho, 2u), o2);
                sts_v4(tile + swz<64>(rho, 3u), o3);
              }
            }
            if (ph == 1) {                                   // all eight classes of this tile-step have left TMEM
              tcgen05_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(smem_u32(&tmem_empty_bar[abuf]));
            }
            fence_proxy_async_smem();                        // generic-proxy writes -> visible to the TMA store
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&ready_bar[2 * grp + b]));
          }
        }
      }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// ==========================================================================================
// v5: stride-2 plane march (3x3x3, stride 2, pad 1, Cin = 32) -- hourglass conv1 (32 -> 64 at full
// resolution), which the per-tap kernel ran at 258 TFLOP/s (element-strided TMA boxes are gathers: one
// 64-byte row per TMA row slot; ncu: 0.615 ms, tensor pipe 7 % busy).
//
// The problem with stride 2 in the dense-row scheme is that the A operand of a tap would be every
// SECOND row of the shared-memory tile, which no UMMA descriptor can express.  Two re-indexings make
// every tap a dense, row-shifted window again:
//   * w: two neighbouring voxels (2j, 2j+1) of 32 channels are ONE 128-byte row ("pair row", exactly a
//     SWIZZLE_128B row).  Output column ow reads input w = 2ow-1, 2ow, 2ow+1 = the upper K half of
//     pair ow-1, the lower K half of pair ow, the upper K half of pair ow: kw selects a 64-byte K
//     slice (descriptor start + 64 B) and a shift of 0 / 1 pair rows -- consecutive ow are
//     consecutive rows;
//   * h: each input plane is loaded as two sub-tiles, its even rows (E) and its odd rows (O), by two
//     TMA boxes over the tensor viewed as [N*D][H/2][2][W/2][64]: kh = 1 reads E, kh = 0 / 2 read O
//     shifted by 0 / one tile row.
//   * d: input plane 2o feeds output plane o (kd = 1); input plane 2o+1 feeds o (kd = 2) and o+1
//     (kd = 0) -- fused into one N = 2*Cout instruction on adjacent TMEM accumulator blocks, as in the
//     kd-fused stride-1 kernel.  Every input plane is read from HBM once and consumed by one batch of
//     18 MMAs, so three ring slots suffice next to the 27 resident weight tiles.
// ==========================================================================================
struct S2Params {
  int N, D, H, W;              // input extent (all even); output is D/2 x H/2 x W/2
  int WP, TH, TWv;             // pair-row pitch, tile rows (WP*TH == 128), valid output columns = WP - 1
  int tiles_h, tiles_w;
  int DC, nchunk;              // output-depth chunk per work unit
  int num_units;
  int e_bytes, o_bytes, slot_bytes, nslots;
  const float* scale;
  const float* bias;
  EpiParams epi;
};

template <int CP>
__global__ void __launch_bounds__(kThreads, 1)
conv3d_s2_kernel(const __grid_constant__ CUtensorMap map_xe, const __grid_constant__ CUtensorMap map_xo,
                 const __grid_constant__ CUtensorMap map_w, const __grid_constant__ S2Params p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t full_bar[kMaxSlots];
  __shared__ __align__(8) uint64_t empty_bar[kMaxSlots];
  __shared__ __align__(8) uint64_t w_bar;
  __shared__ __align__(8) uint64_t acc_full_bar[kMaxBlocks];
  __shared__ __align__(8) uint64_t acc_empty_bar[kMaxBlocks];
  __shared__ uint32_t tmem_base_smem;
  __shared__ __align__(16) float s_scale[64], s_bias[64];

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  constexpr uint32_t R = 512u / (uint32_t)CP;             // accumulator blocks in the TMEM ring
  constexpr uint32_t RMASK = R - 1u;
  constexpr uint32_t LOGR = R == 16u ? 4u : 3u;
  static_assert(CP == 32 || CP == 64, "Cout must be 32 or 64");
  constexpr uint32_t kTapBytes = (uint32_t)CP * 64u;      // one weight tile: CP rows x 32 ci x 2 B
  const uint32_t w_base = (smem_u32(smem) + 1023u) & ~1023u;
  const uint32_t slots_base = w_base + ((27u * kTapBytes + 1023u) & ~1023u);
  const int Do = p.D >> 1;

  if (threadIdx.x < 64) {
    s_scale[threadIdx.x] = (p.scale && threadIdx.x < p.epi.Cout) ? p.scale[threadIdx.x] : 1.f;
    s_bias[threadIdx.x] = (p.bias && threadIdx.x < p.epi.Cout) ? p.bias[threadIdx.x] : 0.f;
  }
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_xe) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_xo) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
    for (int s = 0; s < p.nslots; ++s) {
      mbar_init(smem_u32(&full_bar[s]), 1);
      mbar_init(smem_u32(&empty_bar[s]), 1);
    }
    mbar_init(smem_u32(&w_bar), 1);
    for (uint32_t b = 0; b < R; ++b) {
      mbar_init(smem_u32(&acc_full_bar[b]), 1);
      mbar_init(smem_u32(&acc_empty_bar[b]), 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)),
                 "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  if (warp >= 2) {                                        // zero the accumulator ring once (all MMAs accumulate)
    const uint32_t lane_base = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    for (uint32_t c = 0; c < 512u; c += 16u) tmem_st16_zero(lane_base + c);
    tmem_st_wait();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();

  // work unit -> (n, tile, output-depth chunk [o0, o1)); input planes dlo .. 2*o1-1
  auto decode = [&](int unit, int& n, int& oh0, int& ow0, int& o0, int& o1) {
    const int ch = unit % p.nchunk; unit /= p.nchunk;
    const int tw = unit % p.tiles_w; unit /= p.tiles_w;
    const int th = unit % p.tiles_h; n = unit / p.tiles_h;
    oh0 = th * p.TH; ow0 = tw * p.TWv; o0 = ch * p.DC; o1 = min(Do, o0 + p.DC);
  };

  if (warp == 0) {
    // ===================== TMA producer =====================
    const uint32_t wb = smem_u32(&w_bar);
    if (elect_one()) {
      mbar_expect_tx(wb, 27u * kTapBytes);
      // smem order [(kh,kw)][kd = 2, 0, 1]: the kd = 2 / kd = 0 tiles of one in-plane tap are adjacent (one
      // B operand of 2*Cout rows for the odd input planes), the kd = 1 tile (even planes) follows
      for (int t2 = 0; t2 < 9; ++t2)
        for (int j = 0; j < 3; ++j) {
          const int kd = j == 0 ? 2 : (j == 1 ? 0 : 1);
          tma_load_2d(w_base + (uint32_t)(t2 * 3 + j) * kTapBytes, &map_w, wb, 0, (kd * 9 + t2) * CP);
        }
    }
    __syncwarp();
    uint32_t slot = 0, phase = 0, slot_addr = slots_base;
    for (int unit = blockIdx.x; unit < p.num_units; unit += gridDim.x) {
      int n, oh0, ow0, o0, o1;
      decode(unit, n, oh0, ow0, o0, o1);
      const int dlo = max(0, 2 * o0 - 1), dhi = 2 * o1 - 1;
      for (int d = dlo; d <= dhi; ++d) {
        mbar_wait(smem_u32(&empty_bar[slot]), phase ^ 1u);
        if (elect_one()) {
          const uint32_t fb = smem_u32(&full_bar[slot]);
          mbar_expect_tx(fb, (uint32_t)(p.e_bytes + p.o_bytes));
          tma_load_5d(slot_addr, &map_xe, fb, 0, ow0 - 1, 0, oh0, n * p.D + d);                   // even rows 2*oh
          tma_load_5d(slot_addr + (uint32_t)p.e_bytes, &map_xo, fb, 0, ow0 - 1, 1, oh0 - 1, n * p.D + d);   // odd rows 2*oh-1 ..
        }
        __syncwarp();
        slot_addr += (uint32_t)p.slot_bytes;
        if (++slot == (uint32_t)p.nslots) { slot = 0; phase ^= 1u; slot_addr = slots_base; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc1 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(CP >> 3) << 17) | ((uint32_t)(kTileM >> 4) << 24);
    constexpr uint32_t idesc2 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)((2 * CP) >> 3) << 17) | ((uint32_t)(kTileM >> 4) << 24);
    const uint32_t a_hi = (uint32_t)(make_smem_desc(0, 128) >> 32);     // pair rows: SWIZZLE_128B
    const uint32_t b_hi = (uint32_t)(make_smem_desc(0, 64) >> 32);      // weight tiles: 64-byte rows, SWIZZLE_64B
    constexpr uint32_t lo_flags = 1u << 16;
    constexpr uint32_t b_tap = kTapBytes >> 4;
    // A window of tap (kh,kw), in 16-byte units from the slot base: sub-tile (E for kh = 1, O otherwise), row shift
    // (kh = 2: one tile row; kw > 0: one pair row), K slice (kw = 1: lower 64 B, kw = 0 / 2: upper 64 B)
    uint32_t a_off[9];
#pragma unroll
    for (int kh = 0; kh < 3; ++kh)
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        const uint32_t sub = kh == 1 ? 0u : (uint32_t)p.e_bytes;
        const uint32_t rows = (kh == 2 ? (uint32_t)p.WP : 0u) + (kw > 0 ? 1u : 0u);
        a_off[kh * 3 + kw] = (sub + rows * 128u + (kw == 1 ? 0u : 64u)) >> 4;
      }
    const uint32_t b_lo0 = ((w_base >> 4) & 0x3FFFu) | lo_flags;
    const uint32_t a_lo0 = ((slots_base >> 4) & 0x3FFFu) | lo_flags;
    const uint32_t a_step = (uint32_t)p.slot_bytes >> 4;
    mbar_wait(smem_u32(&w_bar), 0);
    uint32_t slot = 0, phase = 0, a_plane = a_lo0;
    uint32_t g0 = 0;                                      // accumulator index of output plane o0 (global over units)
    for (int unit = blockIdx.x; unit < p.num_units; unit += gridDim.x) {
      int n, oh0, ow0, o0, o1;
      decode(unit, n, oh0, ow0, o0, o1);
      const int dlo = max(0, 2 * o0 - 1), dhi = 2 * o1 - 1;
      if ((dlo & 1) == 0) mbar_wait(smem_u32(&acc_empty_bar[g0 & RMASK]), ((g0 >> LOGR) & 1u) ^ 1u);   // o0 == 0
      for (int d = dlo; d <= dhi; ++d) {
        mbar_wait(smem_u32(&full_bar[slot]), phase);
        const int o = d >> 1;
        const bool odd = d & 1;
        const bool lo = o >= o0, hi = odd && (o + 1 < o1);         // targets: acc(o) [kd = 1 or 2], acc(o+1) [kd = 0]
        const uint32_t g = g0 + (uint32_t)(o - o0);                // accumulator index of output plane o
        if (hi) mbar_wait(smem_u32(&acc_empty_bar[(g + 1u) & RMASK]), (((g + 1u) >> LOGR) & 1u) ^ 1u);
        tcgen05_fence_after();
        const uint32_t blk = g & RMASK, blk1 = (g + 1u) & RMASK;
        if (elect_one()) {
          if (!odd) {
#pragma unroll
            for (int t2 = 0; t2 < 9; ++t2)
#pragma unroll
              for (int k = 0; k < 2; ++k)
                umma_bf16(tmem_base + blk * (uint32_t)CP, desc64(a_hi, a_plane + a_off[t2] + 2u * k),
                          desc64(b_hi, b_lo0 + (uint32_t)(t2 * 3 + 2) * b_tap + 2u * k), idesc1, 1u);
          } else if (lo && hi && blk1 == blk + 1u) {
#pragma unroll
            for (int t2 = 0; t2 < 9; ++t2)
#pragma unroll
              for (int k = 0; k < 2; ++k)
                umma_bf16(tmem_base + blk * (uint32_t)CP, desc64(a_hi, a_plane + a_off[t2] + 2u * k),
                          desc64(b_hi, b_lo0 + (uint32_t)(t2 * 3) * b_tap + 2u * k), idesc2, 1u);
          } else {
#pragma unroll
            for (int t2 = 0; t2 < 9; ++t2)
#pragma unroll
              for (int k = 0; k < 2; ++k) {
                const uint64_t ad = desc64(a_hi, a_plane + a_off[t2] + 2u * k);
                const uint32_t bl = b_lo0 + (uint32_t)(t2 * 3) * b_tap + 2u * k;
                if (lo) umma_bf16(tmem_base + blk * (uint32_t)CP, ad, desc64(b_hi, bl), idesc1, 1u);            // kd = 2
                if (hi) umma_bf16(tmem_base + blk1 * (uint32_t)CP, ad, desc64(b_hi, bl + b_tap), idesc1, 1u);   // kd = 0
              }
          }
          umma_commit(smem_u32(&empty_bar[slot]));                           // plane consumed
          if (odd && lo) umma_commit(smem_u32(&acc_full_bar[blk]));          // output plane o complete
        }
        __syncwarp();
        a_plane += a_step;
        if (++slot == (uint32_t)p.nslots) { slot = 0; phase ^= 1u; a_plane = a_lo0; }
      }
      g0 += (uint32_t)(o1 - o0);
    }
  } else {
    // ===================== epilogue (warps 2..5) =====================
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const int r_w = row % p.WP, r_h = row / p.WP;
    const int variant = epilogue_variant(p.epi);
    const uint32_t lane_base = tmem_base + ((uint32_t)(quad * 32) << 16);
    const int Ho = p.H >> 1, Wo = p.W >> 1;
    const int64_t plane_vox = (int64_t)Ho * Wo;
    uint32_t g = 0;
    for (int unit = blockIdx.x; unit < p.num_units; unit += gridDim.x) {
      int n, oh0, ow0, o0, o1;
      decode(unit, n, oh0, ow0, o0, o1);
      const int ow = ow0 + r_w, oh = oh0 + r_h;
      const bool in_range = r_w < p.TWv && ow < Wo && oh < Ho;
      int64_t vox = (((int64_t)n * Do + o0) * Ho + oh) * Wo + ow;
      for (int o = o0; o < o1; ++o, ++g, vox += plane_vox) {
        const uint32_t blk = g & RMASK;
        ResidualRow rr;
        residual_prefetch(p.epi, in_range, vox, rr);
        mbar_wait(smem_u32(&acc_full_bar[blk]), (g >> LOGR) & 1u);
        tcgen05_fence_after();
        const uint32_t taddr = lane_base + blk * (uint32_t)CP;
        epilogue_row(p.epi, variant, taddr, in_range, vox, s_scale, s_bias, rr);
#pragma unroll
        for (int c = 0; c < CP; c += 16) tmem_st16_zero(taddr + (uint32_t)c);
        tmem_st_wait();
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&acc_empty_bar[blk]));
      }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}


// ==========================================================================================
// v6: large-kernel plane march (K = 5 / 7, stride 1, "same" padding, dilation 1 or 2, Cout = 32) -- the
// instance branch's conv1 (7^3, 64 -> 32: 65 % of its FLOPs), conv2 (5^3) and conv3 (5^3, dilation 2),
// vernier.py:252-263, which the per-tap kernel ran at 271 / 140 TFLOP/s (profiles/r01_instance_v1.txt).
//
// Same ingredients as the 3^3 kd-fused kernel -- dense-row plane tiles whose in-plane taps are row-shifted
// UMMA windows, all K depth taps fused into ONE instruction (N = K*32 = 160 / 224 columns: above the
// 144-column break-even of the 71.6-cycle SS-mode MMA floor, so the tensor pipe, not operand fetch, is the
// bound), accumulators of the output planes in a ring of 16 TMEM blocks -- but K^3 weight tiles (250 KB /
// 1.4 MB) cannot stay in shared memory.  They are STREAMED from L2, one K*32-row tile per in-plane tap, and
// the loop order is chosen so that each streamed tile is used 4*Cin/16 times: a CTA keeps a GROUP of four
// consecutive input planes resident and applies every weight tile to all four before moving on (16 bytes of
// weights per cycle and SM instead of 64).  Warp roles: plane producer, MMA issuer, 4 epilogue warps, weight
// producer.  With dilation 2 the accumulator ring is split by output-plane parity so that the K targets of an
// input plane (o = p + pad - 2 kd) are still adjacent TMEM blocks.
// ==========================================================================================
constexpr int kBigThreads = 224;
constexpr int kBigP = 4;                 // input planes per group
constexpr int kBigMaxW = 8;              // weight ring slots

struct BigKParams {
  int N, D, H, W;
  int dil, pad;                // pad == dil*(K-1)/2
  int WP, TH, TWv;
  int tiles_h, tiles_w, num_cols;
  int plane_bytes, plane_slot_bytes;
  int w_bytes, nw;             // bytes of one streamed weight tile (K*32 rows), ring slots
  const float* scale;
  const float* bias;
  EpiParams epi;
};

__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
// TMEM block of accumulator index G (output planes numbered consecutively over this CTA's columns)
__device__ __forceinline__ uint32_t bigk_block(uint32_t G, int dil) {
  return dil == 1 ? ((0u - G) & 15u) : (((G & 1u) << 3) | ((0u - (G >> 1)) & 7u));
}

template <int K, int KSTEPS, int SUBROW>
__global__ void __launch_bounds__(kBigThreads, 1)
conv3d_bigk_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w,
                   const __grid_constant__ BigKParams p) {
  constexpr int CP = 32;
  constexpr int P = kBigP;
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t pfull_bar[P];
  __shared__ __align__(8) uint64_t planes_empty_bar;
  __shared__ __align__(8) uint64_t wfull_bar[kBigMaxW];
  __shared__ __align__(8) uint64_t wempty_bar[kBigMaxW];
  __shared__ __align__(8) uint64_t acc_full_bar[16];
  __shared__ __align__(8) uint64_t acc_empty_bar[16];
  __shared__ uint32_t tmem_base_smem;
  __shared__ __align__(16) float s_scale[64], s_bias[64];

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t planes_base = (smem_u32(smem) + 1023u) & ~1023u;
  const uint32_t w_base = planes_base + (uint32_t)(P * p.plane_slot_bytes);
  const int D = p.D;

  if (threadIdx.x < 64) {
    s_scale[threadIdx.x] = (p.scale && threadIdx.x < p.epi.Cout) ? p.scale[threadIdx.x] : 1.f;
    s_bias[threadIdx.x] = (p.bias && threadIdx.x < p.epi.Cout) ? p.bias[threadIdx.x] : 0.f;
  }
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_x) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
    for (int j = 0; j < P; ++j) mbar_init(smem_u32(&pfull_bar[j]), 1);
    mbar_init(smem_u32(&planes_empty_bar), 1);
    for (int s = 0; s < p.nw; ++s) {
      mbar_init(smem_u32(&wfull_bar[s]), 1);
      mbar_init(smem_u32(&wempty_bar[s]), 1);
    }
    for (int b = 0; b < 16; ++b) {
      mbar_init(smem_u32(&acc_full_bar[b]), 1);
      mbar_init(smem_u32(&acc_empty_bar[b]), 4);          // one arrive per epilogue warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)),
                 "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  if (warp >= 2 && warp < 6) {                            // zero the accumulator ring once (all MMAs accumulate)
    const uint32_t lane_base = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    for (uint32_t c = 0; c < 512u; c += 16u) tmem_st16_zero(lane_base + c);
    tmem_st_wait();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();

  if (warp == 0) {
    // ===================== plane producer: one group of P input planes at a time =====================
    uint32_t gq = 0;
    int tw = blockIdx.x % p.tiles_w, rest = blockIdx.x / p.tiles_w;
    for (int col = blockIdx.x; col < p.num_cols; col += gridDim.x) {
      const int th = rest % p.tiles_h, n = rest / p.tiles_h;
      const int w0 = tw * p.TWv - p.pad, h0 = th * p.TH - p.pad;
      for (int p0 = 0; p0 < D; p0 += P, ++gq) {
        mbar_wait(smem_u32(&planes_empty_bar), (gq & 1u) ^ 1u);       // previous group's MMAs have retired
        if (elect_one()) {
#pragma unroll
          for (int j = 0; j < P; ++j) {
            const uint32_t fb = smem_u32(&pfull_bar[j]);
            if (p0 + j < D) {
              mbar_expect_tx(fb, (uint32_t)p.plane_bytes);
              tma_load_5d(planes_base + (uint32_t)(j * p.plane_slot_bytes), &map_x, fb, 0, w0, h0, p0 + j, n);
            } else {
              mbar_arrive(fb);                                         // keeps the barrier phases in step
            }
          }
        }
        __syncwarp();
      }
      const int nc = col + (int)gridDim.x;
      tw = nc % p.tiles_w; rest = nc / p.tiles_w;
    }
  } else if (warp == 6) {
    // ===================== weight producer: one K*32-row tile per (group, in-plane tap) =====================
    uint32_t slot = 0, phase = 0, slot_addr = w_base;
    for (int col = blockIdx.x; col < p.num_cols; col += gridDim.x) {
      for (int p0 = 0; p0 < D; p0 += P) {
        for (int t2 = 0; t2 < K * K; ++t2) {
          mbar_wait(smem_u32(&wempty_bar[slot]), phase ^ 1u);
          if (elect_one()) {
            const uint32_t fb = smem_u32(&wfull_bar[slot]);
            mbar_expect_tx(fb, (uint32_t)p.w_bytes);
            tma_load_3d(slot_addr, &map_w, fb, 0, t2 * CP, 0);         // rows [kd][co] of tap (., kh, kw)
          }
          __syncwarp();
          slot_addr += (uint32_t)p.w_bytes;
          if (++slot == (uint32_t)p.nw) { slot = 0; phase ^= 1u; slot_addr = w_base; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc0 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kTileM >> 4) << 24);
    const uint32_t desc_hi = (uint32_t)(make_smem_desc(0, SUBROW) >> 32);
    constexpr uint32_t lo_flags = 1u << 16;
    constexpr uint32_t kd_rows = (uint32_t)(CP * SUBROW) >> 4;         // one kd block of the weight tile, 16-byte units
    const uint32_t a_lo0 = ((planes_base >> 4) & 0x3FFFu) | lo_flags;
    const uint32_t a_step = (uint32_t)p.plane_slot_bytes >> 4;
    const uint32_t b_lo0 = ((w_base >> 4) & 0x3FFFu) | lo_flags;
    const uint32_t b_step = (uint32_t)p.w_bytes >> 4;
    const int dil = p.dil, pad = p.pad;
    uint32_t wslot = 0, wphase = 0, b_cur = b_lo0;
    uint32_t gq = 0, it = 0;
    for (int col = blockIdx.x; col < p.num_cols; col += gridDim.x, ++it) {
      const uint32_t Gcol = it * (uint32_t)D;
      int o_waited = -1, o_committed = -1;
      for (int p0 = 0; p0 < D; p0 += P, ++gq) {
        const int np = min(P, D - p0), p_last = p0 + np - 1;
        // accumulator blocks this group touches for the first time
        const int o_need = min(D - 1, p_last + pad);
        for (int o = o_waited + 1; o <= o_need; ++o) {
          const uint32_t G = Gcol + (uint32_t)o;
          mbar_wait(smem_u32(&acc_empty_bar[bigk_block(G, dil)]), ((G >> 4) & 1u) ^ 1u);
        }
        o_waited = max(o_waited, o_need);
        tcgen05_fence_after();
        // per plane: valid depth taps [kd_lo, kd_hi], first TMEM block, blocks before the ring wraps
        uint32_t d_first[P], d_wrap[P], n_first[P], n_total[P], b_off[P];
#pragma unroll
        for (int j = 0; j < P; ++j) {
          const int pp = p0 + j;
          const int kd_lo = max(0, (pp + pad - (D - 1) + dil - 1) / dil);
          const int kd_hi = min(K - 1, (pp + pad) / dil);
          const uint32_t G = Gcol + (uint32_t)(pp + pad - kd_lo * dil);          // largest output plane = lowest kd
          const uint32_t blk = bigk_block(G, dil);
          const uint32_t room = dil == 1 ? 16u - blk : 8u - (blk & 7u);
          n_total[j] = (uint32_t)max(0, kd_hi - kd_lo + 1);
          n_first[j] = min(n_total[j], room);
          d_first[j] = tmem_base + blk * (uint32_t)CP;
          d_wrap[j] = tmem_base + (dil == 1 ? 0u : (blk & 8u)) * (uint32_t)CP;
          b_off[j] = (uint32_t)kd_lo * kd_rows;
        }
        for (int t2 = 0; t2 < K * K; ++t2) {
          mbar_wait(smem_u32(&wfull_bar[wslot]), wphase);
          if (t2 == 0) {
#pragma unroll
            for (int j = 0; j < P; ++j) mbar_wait(smem_u32(&pfull_bar[j]), gq & 1u);
          }
          tcgen05_fence_after();
          const int kh = t2 / K, kw = t2 - kh * K;
          const uint32_t a_off = (uint32_t)((kh * dil * p.WP + kw * dil) * SUBROW) >> 4;
          if (elect_one()) {
#pragma unroll
            for (int j = 0; j < P; ++j) {
              if (j < np) {
                const uint32_t a_lo = a_lo0 + (uint32_t)j * a_step + a_off;
                const uint32_t b_lo = b_cur + b_off[j];
                const uint32_t id1 = idesc0 | (((n_first[j] * CP) >> 3) << 17);
#pragma unroll
                for (int k = 0; k < KSTEPS; ++k)
                  umma_bf16(d_first[j], desc64(desc_hi, a_lo + 2u * k), desc64(desc_hi, b_lo + 2u * k), id1, 1u);
                if (n_first[j] < n_total[j]) {                                     // ring wrap: the remaining depth taps
                  const uint32_t id2 = idesc0 | ((((n_total[j] - n_first[j]) * CP) >> 3) << 17);
                  const uint32_t b2 = b_lo + n_first[j] * kd_rows;
#pragma unroll
                  for (int k = 0; k < KSTEPS; ++k)
                    umma_bf16(d_wrap[j], desc64(desc_hi, a_lo + 2u * k), desc64(desc_hi, b2 + 2u * k), id2, 1u);
                }
              }
            }
            umma_commit(smem_u32(&wempty_bar[wslot]));                             // weight tile consumed
          }
          __syncwarp();
          b_cur += b_step;
          if (++wslot == (uint32_t)p.nw) { wslot = 0; wphase ^= 1u; b_cur = b_lo0; }
        }
        // group done: planes may be overwritten, output planes whose last input plane was in this group are complete
        const int o_done = (p_last == D - 1) ? D - 1 : p_last - pad;
        if (elect_one()) {
          umma_commit(smem_u32(&planes_empty_bar));
          for (int o = o_committed + 1; o <= o_done; ++o) umma_commit(smem_u32(&acc_full_bar[bigk_block(Gcol + (uint32_t)o, dil)]));
        }
        __syncwarp();
        o_committed = max(o_committed, o_done);
      }
    }
  } else {
    // ===================== epilogue (warps 2..5) =====================
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const int r_w = row % p.WP, r_h = row / p.WP;
    const int variant = epilogue_variant(p.epi);
    const uint32_t lane_base = tmem_base + ((uint32_t)(quad * 32) << 16);
    const int64_t plane_vox = (int64_t)p.H * p.W;
    uint32_t G = 0;
    int tw = blockIdx.x % p.tiles_w, rest = blockIdx.x / p.tiles_w;
    for (int col = blockIdx.x; col < p.num_cols; col += gridDim.x) {
      const int th = rest % p.tiles_h, n = rest / p.tiles_h;
      const int ow = tw * p.TWv + r_w, oh = th * p.TH + r_h;
      const bool in_range = r_w < p.TWv && ow < p.W && oh < p.H;
      int64_t vox = (((int64_t)n * D) * p.H + oh) * p.W + ow;
      for (int o = 0; o < D; ++o, ++G, vox += plane_vox) {
        const uint32_t blk = bigk_block(G, p.dil);
        ResidualRow rr;
        residual_prefetch(p.epi, in_range, vox, rr);
        mbar_wait(smem_u32(&acc_full_bar[blk]), (G >> 4) & 1u);
        tcgen05_fence_after();
        const uint32_t taddr = lane_base + blk * (uint32_t)CP;
        epilogue_row(p.epi, variant, taddr, in_range, vox, s_scale, s_bias, rr);
        tmem_st16_zero(taddr);
        tmem_st16_zero(taddr + 16u);
        tmem_st_wait();
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&acc_empty_bar[blk]));
      }
      const int nc = col + (int)gridDim.x;
      tw = nc % p.tiles_w; rest = nc / p.tiles_w;
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// ------------------------------------------------------------------ weight packing
// w fp32: conv [Cout,Cin,k,k,k] / deconv [Cin,Cout,k,k,k]  ->  packed bf16 [k^3][CoutPad][Cin]
__global__ void pack_weights_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out, int Cin, int Cout,
                                    int CoutPad, int K3, int transposed, int64_t total) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int ci = (int)(i % Cin);
    int64_t t = i / Cin;
    int co = (int)(t % CoutPad);
    int tap = (int)(t / CoutPad);
    float v = 0.f;
    if (co < Cout) v = transposed ? w[((int64_t)ci * Cout + co) * K3 + tap] : w[((int64_t)co * Cin + ci) * K3 + tap];
    out[i] = __float2bfloat16_rn(v);
  }
}

// ------------------------------------------------------------------ host side

// pick (TD,TH,TW), product 128, minimising padded work; ties -> wider W (longer contiguous runs)
void pick_tile(int Dj, int Hj, int Wj, int& TD, int& TH, int& TW) {
  double best = -1;
  for (int tw = 128; tw >= 1; tw >>= 1)
    for (int th = 128 / tw; th >= 1; th >>= 1) {
      int td = 128 / (tw * th);
      double padded = (double)round_up(Dj, td) * round_up(Hj, th) * round_up(Wj, tw);
      double eff = (double)Dj * Hj * Wj / padded;
      // prefer tiles that are not degenerate slivers: mild bonus for tw >= 8
      double score = eff + (tw >= 8 ? 1e-3 : 0) + (tw >= 4 ? 1e-4 : 0) + 1e-6 * tw;
      if (score > best) { best = score; TD = td; TH = th; TW = tw; }
    }
}

int launch_conv(const void* x, const void* w_packed, const float* scale, const float* bias, const void* residual, void* y,
                const snvc_conv3d_desc& d, ConvParams p, cudaStream_t stream) {
  EncodeTiledFn enc = encode_fn();
  if (!enc) return fail(SNVC_E_DRIVER, "cuTensorMapEncodeTiled is not available from the CUDA driver");
  p.scale = scale; p.bias = bias; p.residual = (const __nv_bfloat16*)residual; p.y = y;
  pick_tile(p.Dj, p.Hj, p.Wj, p.TD, p.TH, p.TW);
  p.tiles_d = (int)ceil_div(p.Dj, p.TD); p.tiles_h = (int)ceil_div(p.Hj, p.TH); p.tiles_w = (int)ceil_div(p.Wj, p.TW);
  int64_t nt = (int64_t)p.N * p.tiles_d * p.tiles_h * p.tiles_w;
  SNVC_CHECK_ARG(nt < (1ll << 31), "too many tiles");
  p.num_tiles = (int)nt;
  if (p.num_tiles == 0) return 0;
  p.swizzle_bytes = p.Cin * 2;
  p.a_bytes = kTileM * p.Cin * 2;
  p.b_bytes = p.CoutPad * p.Cin * 2;
  const int stage_bytes = p.a_bytes + p.b_bytes;
  p.stages = std::max(2, std::min(kMaxStages, (200 * 1024) / stage_bytes));
  const size_t smem = (size_t)p.stages * stage_bytes + 1024;

  CUtensorMap map_x, map_w;
  {
    cuuint64_t dims[5] = {(cuuint64_t)p.Cin, (cuuint64_t)d.Wi, (cuuint64_t)d.Hi, (cuuint64_t)d.Di, (cuuint64_t)p.N};
    const cuuint64_t cs = (cuuint64_t)(d.in_cstride ? d.in_cstride : p.Cin) * 2;   // bytes between voxels
    cuuint64_t strides[4] = {cs, (cuuint64_t)d.Wi * cs, (cuuint64_t)d.Hi * d.Wi * cs,
                             (cuuint64_t)d.Di * d.Hi * d.Wi * cs};
    const int s = p.in_stride;
    cuuint32_t box[5] = {(cuuint32_t)p.Cin, (cuuint32_t)((p.TW - 1) * s + 1), (cuuint32_t)((p.TH - 1) * s + 1),
                         (cuuint32_t)((p.TD - 1) * s + 1), 1};
    cuuint32_t estr[5] = {1, (cuuint32_t)s, (cuuint32_t)s, (cuuint32_t)s, 1};
    const void* xbase = static_cast<const char*>(x) + (size_t)d.in_coffset * 2;
    CUresult r = enc(&map_x, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(xbase), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_mode(p.swizzle_bytes), CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail((int)r, "cuTensorMapEncodeTiled(x) failed with CUresult %d", (int)r);
  }
  {
    const int K3 = p.K * p.K * p.K;
    cuuint64_t dims[2] = {(cuuint64_t)p.Cin, (cuuint64_t)K3 * p.CoutPad};
    cuuint64_t strides[1] = {(cuuint64_t)p.Cin * 2};
    cuuint32_t box[2] = {(cuuint32_t)p.Cin, (cuuint32_t)p.CoutPad};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&map_w, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(w_packed), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_mode(p.swizzle_bytes), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail((int)r, "cuTensorMapEncodeTiled(w) failed with CUresult %d", (int)r);
  }
  // opt in to the dynamic shared memory this launch needs (static smem counts against the 227 KB cap)
  SNVC_CUDA_OK(cudaFuncSetAttribute(conv3d_tcgen05_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int grid = std::min(p.num_tiles, sm_count());
  conv3d_tcgen05_kernel<<<grid, kThreads, smem, stream>>>(map_x, map_w, p);
  return launch_status("conv3d_tcgen05_kernel");
}


// ---- v2 host side: returns 1 if this conv is not eligible (caller falls back to the per-tap kernel)
// [cout0, cout0 + ncout) : the slice of output channels this launch computes (ncout = 0: all of them)
int launch_halo(const void* x, const void* w_packed, const float* scale, const float* bias, const void* residual,
                void* y, const snvc_conv3d_desc& d, const ConvParams& cp_full, cudaStream_t stream, int bo_mode,
                int cout0 = 0, int ncout = 0) {
  ConvParams cp = cp_full;
  if (ncout > 0) {
    cp.Cout = ncout; cp.CoutPad = round_up(ncout, 16);
    cp.out_coffset += cout0; cp.res_coffset += cout0;
    if (scale) scale += cout0;
    if (bias) bias += cout0;
  }
  const int hw = (d.kernel - 1) * d.dilation;
  if (d.transposed || d.stride != 1 || d.kernel != 3 || 2 * d.pad != hw) return 1;
  if (d.Do != d.Di || d.Ho != d.Hi || d.Wo != d.Wi) return 1;
  EncodeTiledFn enc = encode_fn();
  if (!enc) return fail(SNVC_E_DRIVER, "cuTensorMapEncodeTiled is not available from the CUDA driver");
  HaloParams p{};
  p.N = d.N; p.Cin = d.Cin; p.D = d.Di; p.H = d.Hi; p.W = d.Wi;
  p.K = d.kernel; p.dil = d.dilation; p.pad = d.pad;
  p.scale = scale; p.bias = bias;
  p.epi.Cout = cp.Cout; p.epi.CoutPad = cp.CoutPad; p.epi.relu = cp.relu; p.epi.residual_mode = cp.residual_mode;
  p.epi.sigmoid = cp.sigmoid; p.epi.out_f32 = cp.out_f32; p.epi.out_cstride = cp.out_cstride;
  p.epi.out_coffset = cp.out_coffset; p.epi.res_cstride = cp.res_cstride; p.epi.res_coffset = cp.res_coffset;
  p.epi.residual = (const __nv_bfloat16*)residual; p.epi.y = y;
  p.bo_mode = bo_mode;
  p.w_rows_per_tap = cp_full.CoutPad; p.w_row0 = cout0;
  p.sub_row_bytes = d.Cin * 2;   // single sub-tile (a 2 x SWIZZLE_32B K-split measured no faster than SWIZZLE_64B)
  p.nsub = d.Cin * 2 / p.sub_row_bytes;
  const int row_bytes = d.Cin * 2;
  const int K3 = d.kernel * d.kernel * d.kernel;
  p.w_tap_bytes = cp.CoutPad * d.Cin * 2;
  const int w_total = round_up(K3 * p.w_tap_bytes, 1024);
  const char* mode = opt(OPT_CONV_MODE);
  const bool kdfuse = d.dilation == 1 && !(mode && mode[0] == 'h');     // SNVC_CONV_MODE=halo: v2 (A/B runs)
  if (!kdfuse && ncout > 0) return 1;                                   // Cout slicing is implemented by the kd-fused kernel only
  // staged epilogue (bulk tensor store of a swizzled smem tile), opt-in with SNVC_CONV_STORE=staged.  Measured
  // (profiles/r01_umma_rate.txt, r01_layer_times_staged_vs_direct.txt): an SS-mode M=128,K=16 MMA costs
  // max(71.6, N/2) cycles, so at N = 96 this kernel is bound by MMA issue (18 x 71.6 = 1289 of the 1319 cycles
  // per plane tile), not by the L1 data pipe: removing the store wavefronts changed nothing (615 vs 614 us) and
  // the smaller plane ring made the residual layer slower (773 vs 684 us).  Direct stores stay the default.
  const char* smode = opt(OPT_CONV_STORE);
  const bool staged = kdfuse && cp.Cout == cp.CoutPad && !cp.sigmoid && !cp.out_f32 &&
                      ((cp.out_cstride | cp.out_coffset) & 7) == 0 &&
                      (cp.CoutPad == 16 || cp.CoutPad == 32 || cp.CoutPad == 64) && (smode && smode[0] == 's');
  // pick the row pitch: maximise useful MMA rows, subject to the ring fitting (>= hw + 2 slots) in `budget` bytes
  auto pick = [&](int budget, int min_slots) {
    double best = -1;
    for (int wp = 16; wp <= 64; wp <<= 1) {
      const int twv = wp - hw, th = 128 / wp;
      if (twv <= 0) continue;
      const int slot = p.nsub * round_up(((th + hw) * wp + 16) * p.sub_row_bytes, 1024);
      const int stage = staged ? round_up(th * twv * cp.CoutPad * 2, 1024) : 0;
      if (budget - 2 * stage < slot * min_slots) continue;
      double eff = ((double)d.Wi / (ceil_div(d.Wi, twv) * wp)) * ((double)d.Hi / (ceil_div(d.Hi, th) * th));
      if (eff > best) { best = eff; p.WP = wp; p.TH = th; p.TWv = twv; p.slot_bytes = slot; p.stage_bytes = stage; }
    }
    if (best < 0) return false;
    p.nslots = std::min(kMaxSlots, (budget - 2 * p.stage_bytes) / p.slot_bytes);
    return true;
  };
  if (!pick(225 * 1024 - 1024 - w_total, hw + 2)) return 1;            // weights + ring do not fit: per-tap kernel
  int ctas_per_sm = 1;
  if (kdfuse && cp.CoutPad <= 32) {
    // two CTAs per SM when weights + staging + a 3-slot plane ring fit in half the shared memory (Cin = Cout = 32
    // does): while one CTA's MMA warp does its per-plane bookkeeping the other CTA's MMAs keep the tensor pipe busy
    const int half_budget = (233472 - 2 * 1024) / 2 - 2048 /* static */ - 1024 /* alignment */;
    const char* occ = opt(OPT_CONV_OCC);
    HaloParams keep = p;
    if (!(occ && occ[0] == '1') && pick(half_budget - w_total, staged ? 3 : 4)) ctas_per_sm = 2;
    else p = keep;
  }
  p.tma_store = staged ? 1 : 0;
  p.sub_tile_bytes = p.slot_bytes / p.nsub;
  p.w_sub_bytes = cp.CoutPad * p.sub_row_bytes;
  p.plane_bytes = (p.TH + hw) * p.WP * row_bytes;
  p.tiles_h = (int)ceil_div(d.Hi, p.TH); p.tiles_w = (int)ceil_div(d.Wi, p.TWv);
  const int64_t ncols = (int64_t)d.N * p.tiles_h * p.tiles_w;
  SNVC_CHECK_ARG(ncols < (1ll << 31), "too many tile columns");
  p.num_cols = (int)ncols;
  const size_t smem = (size_t)w_total + 2 * (size_t)p.stage_bytes + (size_t)p.nslots * p.slot_bytes + 1024;

  CUtensorMap map_x, map_w;
  {
    const cuuint64_t cs = (cuuint64_t)(d.in_cstride ? d.in_cstride : d.Cin) * 2;
    cuuint64_t dims[5] = {(cuuint64_t)d.Cin, (cuuint64_t)d.Wi, (cuuint64_t)d.Hi, (cuuint64_t)d.Di, (cuuint64_t)d.N};
    cuuint64_t strides[4] = {cs, (cuuint64_t)d.Wi * cs, (cuuint64_t)d.Hi * d.Wi * cs, (cuuint64_t)d.Di * d.Hi * d.Wi * cs};
    cuuint32_t box[5] = {(cuuint32_t)(p.sub_row_bytes / 2), (cuuint32_t)p.WP, (cuuint32_t)(p.TH + hw), 1, 1};
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    const void* xbase = static_cast<const char*>(x) + (size_t)d.in_coffset * 2;
    CUresult r = enc(&map_x, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(xbase), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_mode(p.sub_row_bytes), CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail((int)r, "cuTensorMapEncodeTiled(x, halo) failed with CUresult %d", (int)r);
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)d.Cin, (cuuint64_t)K3 * cp_full.CoutPad};
    cuuint64_t strides[1] = {(cuuint64_t)d.Cin * 2};
    cuuint32_t box[2] = {(cuuint32_t)(p.sub_row_bytes / 2), (cuuint32_t)cp.CoutPad};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&map_w, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(w_packed), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_mode(p.sub_row_bytes), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail((int)r, "cuTensorMapEncodeTiled(w, halo) failed with CUresult %d", (int)r);
  }
  if (!kdfuse) {
    void (*kern)(const CUtensorMap, const CUtensorMap, const HaloParams) = nullptr;
    switch (d.Cin) {
      case 16: kern = conv3d_halo_kernel<3, 1, 32>; break;
      case 32: kern = conv3d_halo_kernel<3, 2, 64>; break;
      case 64: kern = conv3d_halo_kernel<3, 4, 128>; break;
    }
    SNVC_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int grid = std::min(p.num_cols, sm_count());
    kern<<<grid, kThreads, smem, stream>>>(map_x, map_w, p);
    return launch_status("conv3d_halo_kernel");
  }
  CUtensorMap map_y = map_w;                             // (unused unless staged)
  if (staged) {
    // output slice viewed as (Cout, W, H, D, N); box = the compacted tile (Cout, TWv, TH); edges are clipped by TMA
    const cuuint64_t cs = (cuuint64_t)cp.out_cstride * 2;
    cuuint64_t dims[5] = {(cuuint64_t)cp.Cout, (cuuint64_t)d.Wi, (cuuint64_t)d.Hi, (cuuint64_t)d.Di, (cuuint64_t)d.N};
    cuuint64_t strides[4] = {cs, (cuuint64_t)d.Wi * cs, (cuuint64_t)d.Hi * d.Wi * cs, (cuuint64_t)d.Di * d.Hi * d.Wi * cs};
    cuuint32_t box[5] = {(cuuint32_t)cp.Cout, (cuuint32_t)p.TWv, (cuuint32_t)p.TH, 1, 1};
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    void* ybase = static_cast<char*>(y) + (size_t)cp.out_coffset * 2;
    CUresult r = enc(&map_y, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, ybase, dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_mode(cp.Cout * 2), CU_TENSOR_MAP_L2_PROMOTION_NONE,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail((int)r, "cuTensorMapEncodeTiled(y, staged store) failed with CUresult %d", (int)r);
  }
  void (*kern)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const HaloParams) = nullptr;
#define SNVC_KDFUSE_S(KS, SR, C, T) (staged ? conv3d_kdfuse_kernel<KS, SR, C, T, true> : conv3d_kdfuse_kernel<KS, SR, C, T, false>)
#define SNVC_KDFUSE_T(KS, SR, T)                                                                         \
  kern = cp.CoutPad == 16 ? SNVC_KDFUSE_S(KS, SR, 16, T) : (cp.CoutPad == 32 ? SNVC_KDFUSE_S(KS, SR, 32, T) : nullptr)
#define SNVC_KDFUSE(KS, SR)                                                                              \
  if (ctas_per_sm == 2) { SNVC_KDFUSE_T(KS, SR, 256); }                                                  \
  else if (cp.CoutPad == 64) kern = SNVC_KDFUSE_S(KS, SR, 64, 512);                                      \
  else { SNVC_KDFUSE_T(KS, SR, 512); }
  switch (d.Cin) {
    case 16: SNVC_KDFUSE(1, 32); break;
    case 32: SNVC_KDFUSE(2, 64); break;
    case 64: SNVC_KDFUSE(4, 128); break;
  }
#undef SNVC_KDFUSE
#undef SNVC_KDFUSE_T
#undef SNVC_KDFUSE_S
  if (!kern) return 1;                                  // CoutPad 48: per-tap kernel
  SNVC_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int grid = std::min(p.num_cols, ctas_per_sm * sm_count());
  if (const char* mg = opt(OPT_CONV_MAXGRID)) grid = std::max(1, std::min(grid, atoi(mg)));   // tests: force ring wrap-around
  kern<<<grid, kThreads, smem, stream>>>(map_x, map_w, map_y, p);
  return launch_status("conv3d_kdfuse_kernel");
}

// ---- v8 host side: CTA-pair kd-fused plane march; returns 1 when not eligible (caller tries v7, then v3).
// Eligible: 3x3x3, stride 1, dilation 1, "same" padding, a 32-channel output (slice) on the lean epilogue path.
int launch_kdpair(const void* x, const void* w_packed, const float* scale, const float* bias, const void* residual,
                  void* y, const snvc_conv3d_desc& d, const ConvParams& cp_full, cudaStream_t stream, int cout0 = 0,
                  int ncout = 0, const float* addend = nullptr) {
  const char* mode = opt(OPT_CONV_MODE);
  if (!addend && mode && (mode[0] == 'h' || mode[0] == 'k')) return 1;  // SNVC_CONV_MODE=kw / kd / halo: v7 / v3 / v2 (A/B runs)
  if (addend && (d.Cin != 32 || cp_full.residual_mode || ncout > 0 || d.Di < 2)) return 1;
  ConvParams cp = cp_full;
  if (ncout > 0) {
    cp.Cout = ncout; cp.CoutPad = round_up(ncout, 16);
    cp.out_coffset += cout0; cp.res_coffset += cout0;
    if (scale) scale += cout0;
    if (bias) bias += cout0;
  }
  if (d.transposed || d.stride != 1 || d.kernel != 3 || d.dilation != 1 || d.pad != 1) return 1;
  if (d.Do != d.Di || d.Ho != d.Hi || d.Wo != d.Wi) return 1;
  if (cp.Cout != 32 || cp.CoutPad != 32 || cp.sigmoid || ((cp.out_cstride | cp.out_coffset) & 7) != 0) return 1;
  if (d.Cin != 32 && d.Cin != 64) return 1;
  if (sm_count() < 2) return 1;
  EncodeTiledFn enc = encode_fn();
  if (!enc) return fail(SNVC_E_DRIVER, "cuTensorMapEncodeTiled is not available from the CUDA driver");
  HaloParams p{};
  p.N = d.N; p.Cin = d.Cin; p.D = d.Di; p.H = d.Hi; p.W = d.Wi;
  p.K = 3; p.dil = 1; p.pad = 1;
  p.scale = scale; p.bias = bias; p.addend = addend;
  // addend edge planes: desc.addend_edge_lo / _hi = 0 -> the tensor's own first / last plane; k > 0 -> plane k / Do-1-k
  // (a depth slab whose view starts k planes before the volume); < 0 -> this slab does not hold that edge of the volume
  p.add_lo = d.addend_edge_lo < 0 ? -1 : d.addend_edge_lo;
  p.add_hi = d.addend_edge_hi < 0 ? -1 : d.Do - 1 - d.addend_edge_hi;
  p.epi.Cout = cp.Cout; p.epi.CoutPad = cp.CoutPad; p.epi.relu = cp.relu; p.epi.residual_mode = cp.residual_mode;
  p.epi.sigmoid = 0; p.epi.out_f32 = cp.out_f32; p.epi.out_cstride = cp.out_cstride;
  p.epi.out_coffset = cp.out_coffset; p.epi.res_cstride = cp.res_cstride; p.epi.res_coffset = cp.res_coffset;
  p.epi.residual = (const __nv_bfloat16*)residual; p.epi.y = y;
  p.w_rows_per_tap = cp_full.CoutPad; p.w_row0 = cout0;
  p.sub_row_bytes = d.Cin * 2; p.nsub = 1;
  const int row_bytes = d.Cin * 2, hw = 2;
  p.w_tap_bytes = 48 * row_bytes;                        // this CTA's half of a tap's 96-row [kd=2|kd=1|kd=0] slab
  const int w_total = round_up(9 * p.w_tap_bytes, 1024);
  const int res_total = cp.residual_mode ? 4 * 128 * 64 : 0;           // residual ring: 4 tiles of 128 rows x 64 bytes
  if (cp.residual_mode && (cp.res_cstride % 8 || cp.res_coffset % 8 || (reinterpret_cast<uintptr_t>(residual) & 15))) return 1;
  const int budget = 225 * 1024 - 1024 - w_total - res_total;
  // tile = TH rows of pitch WP (TH * WP <= 128 MMA rows, WP - 2 useful columns per row).  Besides the power-of-two
  // pitches, 42 x 3 (126 rows): at W = 312 / 156 / 78 it uses 91 % of the MMA rows against 89 / 81 / 81 % for pitch 32
  double best = -1;
  const int cand[4][2] = {{16, 8}, {32, 4}, {64, 2}, {42, 3}};
  for (int ci = 0; ci < 4; ++ci) {
    const int wp = cand[ci][0], th = cand[ci][1], twv = wp - hw;
    const int slot = round_up(((th + hw) * wp + 16) * row_bytes, 1024);   // + 16 rows: the kw-shifted windows of the last rows
    if (budget < 4 * slot) continue;
    // useful voxels per issued MMA row
    const double eff = ((double)d.Wi * d.Hi) / ((double)ceil_div(d.Wi, twv) * ceil_div(d.Hi, th) * 128.0);
    if (eff > best * 1.005) { best = eff; p.WP = wp; p.TH = th; p.TWv = twv; p.slot_bytes = slot; }
  }
  if (best < 0) return 1;
  p.nslots = std::min(kMaxSlots, budget / p.slot_bytes);
  p.sub_tile_bytes = p.slot_bytes;
  p.w_sub_bytes = p.w_tap_bytes;
  p.plane_bytes = (p.TH + hw) * p.WP * row_bytes;
  p.tiles_h = (int)ceil_div(d.Hi, p.TH); p.tiles_w = (int)ceil_div(d.Wi, p.TWv);
  const int64_t ncols = (int64_t)d.N * p.tiles_h * p.tiles_w;
  SNVC_CHECK_ARG(ncols < (1ll << 31), "too many tile columns");
  p.num_cols = (int)ncols;
  const size_t smem = (size_t)w_total + (size_t)res_total + (size_t)p.nslots * p.slot_bytes + 1024;

  CUtensorMap map_x, map_w, map_r;
  {
    const cuuint64_t cs = (cuuint64_t)(d.in_cstride ? d.in_cstride : d.Cin) * 2;
    cuuint64_t dims[5] = {(cuuint64_t)d.Cin, (cuuint64_t)d.Wi, (cuuint64_t)d.Hi, (cuuint64_t)d.Di, (cuuint64_t)d.N};
    cuuint64_t strides[4] = {cs, (cuuint64_t)d.Wi * cs, (cuuint64_t)d.Hi * d.Wi * cs, (cuuint64_t)d.Di * d.Hi * d.Wi * cs};
    cuuint32_t box[5] = {(cuuint32_t)d.Cin, (cuuint32_t)p.WP, (cuuint32_t)(p.TH + hw), 1, 1};
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    const void* xbase = static_cast<const char*>(x) + (size_t)d.in_coffset * 2;
    CUresult r = enc(&map_x, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(xbase), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_mode(row_bytes), CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail((int)r, "cuTensorMapEncodeTiled(x, kdpair) failed with CUresult %d", (int)r);
  }
  if (cp.residual_mode) {
    // the residual's 32-channel slice viewed as (C, W, H, D, N); box = the accumulator rows of a plane tile (WP x TH, the
    // wasted columns included, so that tile row == TMEM lane); clipped edges are zero-filled and never stored
    const cuuint64_t cs = (cuuint64_t)cp.res_cstride * 2;
    cuuint64_t dims[5] = {32, (cuuint64_t)d.Wi, (cuuint64_t)d.Hi, (cuuint64_t)d.Di, (cuuint64_t)d.N};
    cuuint64_t strides[4] = {cs, (cuuint64_t)d.Wi * cs, (cuuint64_t)d.Hi * d.Wi * cs, (cuuint64_t)d.Di * d.Hi * d.Wi * cs};
    cuuint32_t box[5] = {32, (cuuint32_t)p.WP, (cuuint32_t)p.TH, 1, 1};
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    const void* rbase = static_cast<const char*>(residual) + (size_t)cp.res_coffset * 2;
    CUresult r = enc(&map_r, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(rbase), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail((int)r, "cuTensorMapEncodeTiled(residual, kdpair) failed with CUresult %d", (int)r);
  } else {
    map_r = map_x;                                       // (unused)
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)d.Cin, (cuuint64_t)27 * cp_full.CoutPad};
    cuuint64_t strides[1] = {(cuuint64_t)row_bytes};
    cuuint32_t box[2] = {(cuuint32_t)d.Cin, 16};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&map_w, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(w_packed), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_mode(row_bytes), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail((int)r, "cuTensorMapEncodeTiled(w, kdpair) failed with CUresult %d", (int)r);
  }
  void (*kern)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const HaloParams) = nullptr;
#define SNVC_PAIR_W(KS, SR, RS, AD)                                                                               \
  (p.WP == 16 ? conv3d_kdpair_kernel<KS, SR, RS, 16, AD>                                                          \
              : (p.WP == 32 ? conv3d_kdpair_kernel<KS, SR, RS, 32, AD>                                            \
                            : (p.WP == 42 ? conv3d_kdpair_kernel<KS, SR, RS, 42, AD> : conv3d_kdpair_kernel<KS, SR, RS, 64, AD>)))
  if (addend) kern = SNVC_PAIR_W(2, 64, false, true);
  else if (d.Cin == 32) kern = cp.residual_mode ? SNVC_PAIR_W(2, 64, true, false) : SNVC_PAIR_W(2, 64, false, false);
  else kern = cp.residual_mode ? SNVC_PAIR_W(4, 128, true, false) : SNVC_PAIR_W(4, 128, false, false);
#undef SNVC_PAIR_W
  SNVC_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  // how many CTA pairs the device can hold at once for this kernel / shared-memory size (cached: the query is slow);
  // a device that cannot co-schedule a pair (MIG slice, SM count 1) falls back to the single-CTA kernels
  static std::mutex mu;
  static std::vector<std::tuple<const void*, size_t, int, int>> cache;   // (kernel, smem, device, clusters)
  int dev = 0, max_clusters = -1;
  SNVC_CUDA_OK(cudaGetDevice(&dev));
  {
    std::lock_guard<std::mutex> lk(mu);
    for (auto& e : cache)
      if (std::get<0>(e) == (const void*)kern && std::get<1>(e) == smem && std::get<2>(e) == dev) max_clusters = std::get<3>(e);
    if (max_clusters < 0) {
      cudaLaunchConfig_t cfg{};
      cfg.gridDim = dim3(2 * (sm_count() / 2)); cfg.blockDim = dim3(kPairThreads); cfg.dynamicSmemBytes = smem;
      cudaLaunchAttribute at{};
      at.id = cudaLaunchAttributeClusterDimension; at.val.clusterDim.x = 2; at.val.clusterDim.y = 1; at.val.clusterDim.z = 1;
      cfg.attrs = &at; cfg.numAttrs = 1;
      int n = 0;
      if (cudaOccupancyMaxActiveClusters(&n, (const void*)kern, &cfg) != cudaSuccess) { n = 0; (void)cudaGetLastError(); }
      max_clusters = n;
      cache.emplace_back((const void*)kern, smem, dev, n);
    }
  }
  if (max_clusters < 1) return 1;
  int pairs = std::min((p.num_cols + 1) / 2, max_clusters);
  if (const char* mg = opt(OPT_CONV_MAXGRID)) pairs = std::max(1, std::min(pairs, atoi(mg)));   // tests: force ring wrap-around
  kern<<<2 * pairs, kPairThreads, smem, stream>>>(map_x, map_w, map_r, p);
  return launch_status("conv3d_kdpair_kernel");
}

// ---- v7 host side: kw+kd-fused plane march; returns 1 when not eligible (caller uses the kd-fused kernel).
// Eligible: 3x3x3, stride 1, dilation 1, "same" padding, a 32-channel output (slice) on the lean epilogue path.
int launch_kwfuse(const void* x, const void* w_packed, const float* scale, const float* bias, const void* residual,
                  void* y, const snvc_conv3d_desc& d, const ConvParams& cp_full, cudaStream_t stream, int cout0 = 0,
                  int ncout = 0) {
  const char* mode = opt(OPT_CONV_MODE);
  if (mode && (mode[0] == 'h' || (mode[0] == 'k' && mode[1] == 'd'))) return 1;   // SNVC_CONV_MODE=kd / halo: v3 / v2 (A/B runs)
  ConvParams cp = cp_full;
  if (ncout > 0) {
    cp.Cout = ncout; cp.CoutPad = round_up(ncout, 16);
    cp.out_coffset += cout0; cp.res_coffset += cout0;
    if (scale) scale += cout0;
    if (bias) bias += cout0;
  }
  if (d.transposed || d.stride != 1 || d.kernel != 3 || d.dilation != 1 || d.pad != 1) return 1;
  if (d.Do != d.Di || d.Ho != d.Hi || d.Wo != d.Wi) return 1;
  if (cp.Cout != 32 || cp.CoutPad != 32 || cp.sigmoid || ((cp.out_cstride | cp.out_coffset) & 7) != 0) return 1;
  if (d.Cin != 32 && d.Cin != 64) return 1;
  EncodeTiledFn enc = encode_fn();
  if (!enc) return fail(SNVC_E_DRIVER, "cuTensorMapEncodeTiled is not available from the CUDA driver");
  HaloParams p{};
  p.N = d.N; p.Cin = d.Cin; p.D = d.Di; p.H = d.Hi; p.W = d.Wi;
  p.K = 3; p.dil = 1; p.pad = 1;
  p.scale = scale; p.bias = bias;
  p.epi.Cout = cp.Cout; p.epi.CoutPad = cp.CoutPad; p.epi.relu = cp.relu; p.epi.residual_mode = cp.residual_mode;
  p.epi.sigmoid = 0; p.epi.out_f32 = cp.out_f32; p.epi.out_cstride = cp.out_cstride;
  p.epi.out_coffset = cp.out_coffset; p.epi.res_cstride = cp.res_cstride; p.epi.res_coffset = cp.res_coffset;
  p.epi.residual = (const __nv_bfloat16*)residual; p.epi.y = y;
  p.w_rows_per_tap = cp_full.CoutPad; p.w_row0 = cout0;
  p.sub_row_bytes = d.Cin * 2; p.nsub = 1;
  const int row_bytes = d.Cin * 2, hw = 2;
  p.w_tap_bytes = cp.CoutPad * row_bytes;
  const int w_total = round_up(27 * p.w_tap_bytes, 1024);
  // row pitch 16 or 32 only: output column w and its partial sums at w+1, w+2 must sit in the same warp
  double best = -1;
  for (int wp = 32; wp >= 16; wp >>= 1) {
    const int twv = wp - hw, th = 128 / wp;
    const double eff = ((double)d.Wi / (ceil_div(d.Wi, twv) * wp)) * ((double)d.Hi / (ceil_div(d.Hi, th) * th));
    if (eff > best) { best = eff; p.WP = wp; p.TH = th; p.TWv = twv; }
  }
  p.slot_bytes = round_up((p.TH + hw) * p.WP * row_bytes, 1024);
  const int budget = 225 * 1024 - 1024 - w_total;
  p.nslots = std::min(kMaxSlots, budget / p.slot_bytes);
  if (p.nslots < 3) return 1;
  p.sub_tile_bytes = p.slot_bytes;
  p.w_sub_bytes = p.w_tap_bytes;
  p.plane_bytes = (p.TH + hw) * p.WP * row_bytes;
  p.tiles_h = (int)ceil_div(d.Hi, p.TH); p.tiles_w = (int)ceil_div(d.Wi, p.TWv);
  const int64_t ncols = (int64_t)d.N * p.tiles_h * p.tiles_w;
  SNVC_CHECK_ARG(ncols < (1ll << 31), "too many tile columns");
  p.num_cols = (int)ncols;
  const size_t smem = (size_t)w_total + (size_t)p.nslots * p.slot_bytes + 1024;

  CUtensorMap map_x, map_w;
  {
    const cuuint64_t cs = (cuuint64_t)(d.in_cstride ? d.in_cstride : d.Cin) * 2;
    cuuint64_t dims[5] = {(cuuint64_t)d.Cin, (cuuint64_t)d.Wi, (cuuint64_t)d.Hi, (cuuint64_t)d.Di, (cuuint64_t)d.N};
    cuuint64_t strides[4] = {cs, (cuuint64_t)d.Wi * cs, (cuuint64_t)d.Hi * d.Wi * cs, (cuuint64_t)d.Di * d.Hi * d.Wi * cs};
    cuuint32_t box[5] = {(cuuint32_t)d.Cin, (cuuint32_t)p.WP, (cuuint32_t)(p.TH + hw), 1, 1};
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    const void* xbase = static_cast<const char*>(x) + (size_t)d.in_coffset * 2;
    CUresult r = enc(&map_x, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(xbase), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_mode(row_bytes), CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail((int)r, "cuTensorMapEncodeTiled(x, kwfuse) failed with CUresult %d", (int)r);
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)d.Cin, (cuuint64_t)27 * cp_full.CoutPad};
    cuuint64_t strides[1] = {(cuuint64_t)row_bytes};
    cuuint32_t box[2] = {(cuuint32_t)d.Cin, (cuuint32_t)cp.CoutPad};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&map_w, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(w_packed), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_mode(row_bytes), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail((int)r, "cuTensorMapEncodeTiled(w, kwfuse) failed with CUresult %d", (int)r);
  }
  void (*kern)(const CUtensorMap, const CUtensorMap, const HaloParams) =
      d.Cin == 32 ? (cp.residual_mode ? conv3d_kwfuse_kernel<2, 64, true> : conv3d_kwfuse_kernel<2, 64, false>)
                  : (cp.residual_mode ? conv3d_kwfuse_kernel<4, 128, true> : conv3d_kwfuse_kernel<4, 128, false>);
  SNVC_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int grid = std::min(p.num_cols, sm_count());
  if (const char* mg = opt(OPT_CONV_MAXGRID)) grid = std::max(1, std::min(grid, atoi(mg)));   // tests: force ring wrap-around
  kern<<<grid, kKwThreads, smem, stream>>>(map_x, map_w, p);
  return launch_status("conv3d_kwfuse_kernel");
}

// ---- v4 host side: fused transposed conv; returns 1 when not eligible (caller uses the per-class launches)
int launch_deconv(const void* x, const void* w_packed, const float* scale, const float* bias, const void* residual,
                  void* y, const snvc_conv3d_desc& d, const ConvParams& cp, cudaStream_t stream, int cout0) {
  if (!d.transposed || cp.sigmoid || cp.out_f32 || (cp.Cout % kDeconvCP) != 0) return 1;
  if (((cp.out_cstride | cp.out_coffset) & 7) != 0) return 1;
  if (!(d.Cin == 16 || d.Cin == 32 || d.Cin == 64)) return 1;
  EncodeTiledFn enc = encode_fn();
  if (!enc) return fail(SNVC_E_DRIVER, "cuTensorMapEncodeTiled is not available from the CUDA driver");
  DeconvParams p{};
  p.N = d.N; p.Cin = d.Cin; p.Di = d.Di; p.Hi = d.Hi; p.Wi = d.Wi;
  p.scale = scale ? scale + cout0 : nullptr;
  p.bias = bias ? bias + cout0 : nullptr;
  p.epi.Cout = kDeconvCP; p.epi.CoutPad = kDeconvCP; p.epi.relu = cp.relu; p.epi.residual_mode = cp.residual_mode;
  p.epi.sigmoid = 0; p.epi.out_f32 = 0; p.epi.out_cstride = cp.out_cstride; p.epi.out_coffset = cp.out_coffset + cout0;
  p.epi.res_cstride = cp.res_cstride; p.epi.res_coffset = cp.res_coffset + cout0;
  p.epi.residual = (const __nv_bfloat16*)residual; p.epi.y = y;
  p.w_rows_per_tap = cp.CoutPad; p.w_row0 = cout0;
  const int row_bytes = d.Cin * 2;
  p.w_tap_bytes = kDeconvCP * row_bytes;
  const int w_total = round_up(27 * p.w_tap_bytes, 1024);
  const int budget = 226 * 1024 - 1024 - w_total;          // 227 KB per CTA minus static smem and alignment
  double best = -1;
  for (int wp = 16; wp <= 64; wp <<= 1) {
    const int twv = wp - 1, th = 128 / wp;
    // a plane slot holds the (TH+1) x WP box (= 128 + WP rows) plus the one row the most shifted MMA window
    // (WP + 1 rows down) reads past it
    const int slot = round_up((128 + wp + 1) * row_bytes, 1024);
    const int stage = round_up(th * 2 * twv * kDeconvCP * 2, 1024);           // one staged output tile (64-byte rows)
    if (budget - 4 * stage < slot * 3) continue;
    const double eff = ((double)d.Wi / (ceil_div(d.Wi, twv) * wp)) * ((double)d.Hi / (ceil_div(d.Hi, th) * th));
    if (eff > best) { best = eff; p.WP = wp; p.TH = th; p.TWv = twv; p.slot_bytes = slot; p.stage_bytes = stage; }
  }
  if (best < 0) return 1;
  p.nslots = std::min(kMaxSlots, (budget - 4 * p.stage_bytes) / p.slot_bytes);
  p.plane_bytes = (p.TH + 1) * p.WP * row_bytes;
  p.tiles_h = (int)ceil_div(d.Hi, p.TH); p.tiles_w = (int)ceil_div(d.Wi, p.TWv);
  // depth chunks: enough work units for ~6 waves when the volume allows it (each chunk re-loads one plane)
  const int64_t cols = (int64_t)d.N * p.tiles_h * p.tiles_w;
  int nchunk = 1;
  while (cols * nchunk < 6ll * sm_count() && d.Di / (nchunk + 1) >= 4) ++nchunk;
  p.DC = (int)ceil_div(d.Di, nchunk);
  p.nchunk = (int)ceil_div(d.Di, p.DC);
  const int64_t units = cols * p.nchunk;
  SNVC_CHECK_ARG(units < (1ll << 31) && (int64_t)d.N * d.Di * 2 < (1ll << 31), "too many work units");
  p.num_units = (int)units;
  const size_t smem = (size_t)w_total + 4 * (size_t)p.stage_bytes + (size_t)p.nslots * p.slot_bytes + 1024;

  CUtensorMap map_x, map_w;
  {
    const cuuint64_t cs = (cuuint64_t)(d.in_cstride ? d.in_cstride : d.Cin) * 2;
    cuuint64_t dims[5] = {(cuuint64_t)d.Cin, (cuuint64_t)d.Wi, (cuuint64_t)d.Hi, (cuuint64_t)d.Di, (cuuint64_t)d.N};
    cuuint64_t strides[4] = {cs, (cuuint64_t)d.Wi * cs, (cuuint64_t)d.Hi * d.Wi * cs, (cuuint64_t)d.Di * d.Hi * d.Wi * cs};
    cuuint32_t box[5] = {(cuuint32_t)d.Cin, (cuuint32_t)p.WP, (cuuint32_t)(p.TH + 1), 1, 1};
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    const void* xbase = static_cast<const char*>(x) + (size_t)d.in_coffset * 2;
    CUresult r = enc(&map_x, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(xbase), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_mode(row_bytes), CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail((int)r, "cuTensorMapEncodeTiled(x, deconv) failed with CUresult %d", (int)r);
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)d.Cin, (cuuint64_t)27 * cp.CoutPad};
    cuuint64_t strides[1] = {(cuuint64_t)d.Cin * 2};
    cuuint32_t box[2] = {(cuuint32_t)d.Cin, (cuuint32_t)kDeconvCP};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&map_w, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(w_packed), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_mode(row_bytes), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail((int)r, "cuTensorMapEncodeTiled(w, deconv) failed with CUresult %d", (int)r);
  }
  // staged tiles: output (and residual) viewed as (32 ch, Wo, ph, Hi, N*Do planes); one box = TH rows x 2*TWv voxels
  CUtensorMap map_y, map_r;
  for (int which = 0; which < 2; ++which) {
    const bool res = which == 1;
    if (res && !cp.residual_mode) { map_r = map_y; break; }
    const cuuint64_t cs = (cuuint64_t)(res ? cp.res_cstride : cp.out_cstride) * 2;
    const int coff = (res ? cp.res_coffset : cp.out_coffset) + cout0;
    const int Wo = 2 * d.Wi;
    cuuint64_t dims[5] = {(cuuint64_t)kDeconvCP, (cuuint64_t)Wo, 2, (cuuint64_t)d.Hi, (cuuint64_t)d.N * 2 * d.Di};
    cuuint64_t strides[4] = {cs, (cuuint64_t)Wo * cs, (cuuint64_t)2 * Wo * cs, (cuuint64_t)2 * d.Hi * Wo * cs};
    cuuint32_t box[5] = {(cuuint32_t)kDeconvCP, (cuuint32_t)(2 * p.TWv), 1, (cuuint32_t)p.TH, 1};
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    char* base = (res ? static_cast<char*>(const_cast<void*>(residual)) : static_cast<char*>(y)) + (size_t)coff * 2;
    CUresult r = enc(res ? &map_r : &map_y, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, base, dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
                     res ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_NONE,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail((int)r, "cuTensorMapEncodeTiled(%s, deconv) failed with CUresult %d", res ? "residual" : "y", (int)r);
  }
  void (*kern)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap, const DeconvParams) = nullptr;
  switch (d.Cin) {
    case 16: kern = conv3d_deconv_kernel<1, 32>; break;
    case 32: kern = conv3d_deconv_kernel<2, 64>; break;
    case 64: kern = conv3d_deconv_kernel<4, 128>; break;
  }
  SNVC_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int grid = std::min(p.num_units, sm_count());
  if (const char* mg = opt(OPT_CONV_MAXGRID)) grid = std::max(1, std::min(grid, atoi(mg)));   // tests: force ring wrap-around
  kern<<<grid, kDeconvThreads, smem, stream>>>(map_x, map_w, map_y, map_r, p);
  return launch_status("conv3d_deconv_kernel");
}

// ---- v5 host side: stride-2 plane march; returns 1 when not eligible (caller uses the per-tap kernel)
int launch_s2(const void* x, const void* w_packed, const float* scale, const float* bias, const void* residual, void* y,
              const snvc_conv3d_desc& d, const ConvParams& cp, cudaStream_t stream) {
  if (d.transposed || d.kernel != 3 || d.stride != 2 || d.pad != 1 || d.dilation != 1 || d.Cin != 32) return 1;
  if (!(cp.Cout == 32 || cp.Cout == 64) || cp.CoutPad != cp.Cout) return 1;
  if ((d.in_cstride != 0 && d.in_cstride != d.Cin) || d.in_coffset != 0) return 1;   // pair rows need dense voxel rows
  if ((d.Di | d.Hi | d.Wi) & 1) return 1;
  EncodeTiledFn enc = encode_fn();
  if (!enc) return fail(SNVC_E_DRIVER, "cuTensorMapEncodeTiled is not available from the CUDA driver");
  S2Params p{};
  p.N = d.N; p.D = d.Di; p.H = d.Hi; p.W = d.Wi;
  p.scale = scale; p.bias = bias;
  p.epi.Cout = cp.Cout; p.epi.CoutPad = cp.CoutPad; p.epi.relu = cp.relu; p.epi.residual_mode = cp.residual_mode;
  p.epi.sigmoid = cp.sigmoid; p.epi.out_f32 = cp.out_f32; p.epi.out_cstride = cp.out_cstride;
  p.epi.out_coffset = cp.out_coffset; p.epi.res_cstride = cp.res_cstride; p.epi.res_coffset = cp.res_coffset;
  p.epi.residual = (const __nv_bfloat16*)residual; p.epi.y = y;
  const int Do = d.Di / 2, Ho = d.Hi / 2, Wo = d.Wi / 2;
  const int w_total = round_up(27 * cp.Cout * 64, 1024);
  const int budget = 225 * 1024 - 1024 - w_total;
  double best = -1;
  for (int wp = 16; wp <= 64; wp <<= 1) {
    const int twv = wp - 1, th = 128 / wp;
    const int e_bytes = 128 * 128;                                     // TH * WP pair rows of 128 B
    const int slot = e_bytes + round_up(((th + 1) * wp + 16) * 128, 1024);
    if (budget < slot * 3) continue;
    const double eff = ((double)Wo / (ceil_div(Wo, twv) * wp)) * ((double)Ho / (ceil_div(Ho, th) * th));
    if (eff > best) { best = eff; p.WP = wp; p.TH = th; p.TWv = twv; p.slot_bytes = slot; p.e_bytes = e_bytes; }
  }
  if (best < 0) return 1;
  p.o_bytes = (p.TH + 1) * p.WP * 128;
  p.nslots = std::min(kMaxSlots, budget / p.slot_bytes);
  p.tiles_h = (int)ceil_div(Ho, p.TH); p.tiles_w = (int)ceil_div(Wo, p.TWv);
  const int64_t cols = (int64_t)d.N * p.tiles_h * p.tiles_w;
  // output-depth chunks: balance the persistent grid (each chunk re-loads one input plane)
  int nchunk = 1;
  double best_eff = -1;
  for (int nc = 1; nc <= 6 && Do / nc >= 2; ++nc) {
    const int dc = (int)ceil_div(Do, nc);
    const int64_t units = cols * ceil_div(Do, dc);
    const double eff = (double)units / (double)(ceil_div(units, sm_count()) * sm_count()) * (2.0 * dc) / (2.0 * dc + 1.0);
    if (eff > best_eff + 1e-9) { best_eff = eff; nchunk = nc; }
  }
  p.DC = (int)ceil_div(Do, nchunk);
  p.nchunk = (int)ceil_div(Do, p.DC);
  const int64_t units = cols * p.nchunk;
  SNVC_CHECK_ARG(units < (1ll << 31) && (int64_t)d.N * d.Di < (1ll << 31), "too many work units");
  p.num_units = (int)units;
  const size_t smem = (size_t)w_total + (size_t)p.nslots * p.slot_bytes + 1024;

  CUtensorMap map_xe, map_xo, map_w;
  for (int odd = 0; odd < 2; ++odd) {
    // input viewed as [N*D][H/2][2][W/2][2 voxels x 32 ch]
    cuuint64_t dims[5] = {64, (cuuint64_t)d.Wi / 2, 2, (cuuint64_t)d.Hi / 2, (cuuint64_t)d.N * d.Di};
    cuuint64_t strides[4] = {128, (cuuint64_t)d.Wi * 64, (cuuint64_t)d.Wi * 128, (cuuint64_t)d.Hi * d.Wi * 64};
    cuuint32_t box[5] = {64, (cuuint32_t)p.WP, 1, (cuuint32_t)(p.TH + odd), 1};
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = enc(odd ? &map_xo : &map_xe, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(x), dims, strides, box,
                     estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail((int)r, "cuTensorMapEncodeTiled(x, s2) failed with CUresult %d", (int)r);
  }
  {
    cuuint64_t dims[2] = {32, (cuuint64_t)27 * cp.CoutPad};
    cuuint64_t strides[1] = {64};
    cuuint32_t box[2] = {32, (cuuint32_t)cp.Cout};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&map_w, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(w_packed), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail((int)r, "cuTensorMapEncodeTiled(w, s2) failed with CUresult %d", (int)r);
  }
  void (*kern)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const S2Params) =
      cp.Cout == 64 ? conv3d_s2_kernel<64> : conv3d_s2_kernel<32>;
  SNVC_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int grid = std::min(p.num_units, sm_count());
  if (const char* mg = opt(OPT_CONV_MAXGRID)) grid = std::max(1, std::min(grid, atoi(mg)));   // tests: force ring wrap-around
  kern<<<grid, kThreads, smem, stream>>>(map_xe, map_xo, map_w, p);
  return launch_status("conv3d_s2_kernel");
}

// ---- v6 host side: large-kernel plane march; returns 1 when not eligible (caller uses the per-tap kernel)
int launch_bigk(const void* x, const void* w_packed, const float* scale, const float* bias, const void* residual, void* y,
                const snvc_conv3d_desc& d, const ConvParams& cp, cudaStream_t stream) {
  const int K = d.kernel;
  const int hw = (K - 1) * d.dilation;
  if (d.transposed || d.stride != 1 || !(K == 5 || K == 7) || 2 * d.pad != hw || !(d.dilation == 1 || d.dilation == 2)) return 1;
  if (d.Do != d.Di || d.Ho != d.Hi || d.Wo != d.Wi) return 1;
  if (cp.CoutPad != 32 || !(d.Cin == 32 || d.Cin == 64)) return 1;
  if (d.dilation == 2 && (d.Di & 1)) return 1;            // the parity-split accumulator ring needs an even depth
  const char* mode = opt(OPT_CONV_MODE);
  if (mode && mode[0] == 't') return 1;
  EncodeTiledFn enc = encode_fn();
  if (!enc) return fail(SNVC_E_DRIVER, "cuTensorMapEncodeTiled is not available from the CUDA driver");
  BigKParams p{};
  p.N = d.N; p.D = d.Di; p.H = d.Hi; p.W = d.Wi; p.dil = d.dilation; p.pad = d.pad;
  p.scale = scale; p.bias = bias;
  p.epi.Cout = cp.Cout; p.epi.CoutPad = cp.CoutPad; p.epi.relu = cp.relu; p.epi.residual_mode = cp.residual_mode;
  p.epi.sigmoid = cp.sigmoid; p.epi.out_f32 = cp.out_f32; p.epi.out_cstride = cp.out_cstride;
  p.epi.out_coffset = cp.out_coffset; p.epi.res_cstride = cp.res_cstride; p.epi.res_coffset = cp.res_coffset;
  p.epi.residual = (const __nv_bfloat16*)residual; p.epi.y = y;
  const int row_bytes = d.Cin * 2;
  p.w_bytes = K * 32 * row_bytes;
  const int budget = 226 * 1024 - 1024;
  double best = -1;
  for (int wp = 16; wp <= 64; wp <<= 1) {
    const int twv = wp - hw, th = 128 / wp;
    if (twv <= 0) continue;
    // box rows plus the rows the most shifted window (hw*WP + hw rows down) reads past the 128-row tile
    const int slot = round_up(((th + hw) * wp + hw + 1) * row_bytes, 1024);
    if (budget < kBigP * slot + 2 * p.w_bytes) continue;
    const double eff = ((double)d.Wi / (ceil_div(d.Wi, twv) * wp)) * ((double)d.Hi / (ceil_div(d.Hi, th) * th));
    if (eff > best) { best = eff; p.WP = wp; p.TH = th; p.TWv = twv; p.plane_slot_bytes = slot; }
  }
  if (best < 0) return 1;
  p.plane_bytes = (p.TH + hw) * p.WP * row_bytes;
  p.nw = std::min(kBigMaxW, (budget - kBigP * p.plane_slot_bytes) / p.w_bytes);
  p.tiles_h = (int)ceil_div(d.Hi, p.TH); p.tiles_w = (int)ceil_div(d.Wi, p.TWv);
  const int64_t ncols = (int64_t)d.N * p.tiles_h * p.tiles_w;
  SNVC_CHECK_ARG(ncols < (1ll << 31) && ncols * d.Di < (1ll << 31), "too many tile columns");
  p.num_cols = (int)ncols;
  const size_t smem = (size_t)kBigP * p.plane_slot_bytes + (size_t)p.nw * p.w_bytes + 1024;

  CUtensorMap map_x, map_w;
  {
    const cuuint64_t cs = (cuuint64_t)(d.in_cstride ? d.in_cstride : d.Cin) * 2;
    cuuint64_t dims[5] = {(cuuint64_t)d.Cin, (cuuint64_t)d.Wi, (cuuint64_t)d.Hi, (cuuint64_t)d.Di, (cuuint64_t)d.N};
    cuuint64_t strides[4] = {cs, (cuuint64_t)d.Wi * cs, (cuuint64_t)d.Hi * d.Wi * cs, (cuuint64_t)d.Di * d.Hi * d.Wi * cs};
    cuuint32_t box[5] = {(cuuint32_t)d.Cin, (cuuint32_t)p.WP, (cuuint32_t)(p.TH + hw), 1, 1};
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    const void* xbase = static_cast<const char*>(x) + (size_t)d.in_coffset * 2;
    CUresult r = enc(&map_x, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(xbase), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_mode(row_bytes), CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail((int)r, "cuTensorMapEncodeTiled(x, bigk) failed with CUresult %d", (int)r);
  }
  {
    // packed weights [kd][kh*K+kw][co][ci] viewed as (ci, (kh,kw,co), kd); one box = the K depth taps of one (kh,kw)
    cuuint64_t dims[3] = {(cuuint64_t)d.Cin, (cuuint64_t)K * K * 32, (cuuint64_t)K};
    cuuint64_t strides[2] = {(cuuint64_t)row_bytes, (cuuint64_t)K * K * 32 * row_bytes};
    cuuint32_t box[3] = {(cuuint32_t)d.Cin, 32, (cuuint32_t)K};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(&map_w, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(w_packed), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_mode(row_bytes), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail((int)r, "cuTensorMapEncodeTiled(w, bigk) failed with CUresult %d", (int)r);
  }
  void (*kern)(const CUtensorMap, const CUtensorMap, const BigKParams) = nullptr;
  if (K == 7 && d.Cin == 64) kern = conv3d_bigk_kernel<7, 4, 128>;
  else if (K == 7 && d.Cin == 32) kern = conv3d_bigk_kernel<7, 2, 64>;
  else if (K == 5 && d.Cin == 64) kern = conv3d_bigk_kernel<5, 4, 128>;
  else kern = conv3d_bigk_kernel<5, 2, 64>;
  SNVC_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int grid = std::min(p.num_cols, sm_count());
  if (const char* mg = opt(OPT_CONV_MAXGRID)) grid = std::max(1, std::min(grid, atoi(mg)));   // tests: force ring wrap-around
  kern<<<grid, kBigThreads, smem, stream>>>(map_x, map_w, p);
  return launch_status("conv3d_bigk_kernel");
}

}  // namespace
}  // namespace snvc

using namespace snvc;

extern "C" int64_t snvc_conv3d_packed_weight_bytes(int32_t Cin, int32_t Cout, int32_t kernel) {
  return (int64_t)kernel * kernel * kernel * round_up(Cout, 16) * Cin * 2;
}

extern "C" int snvc_conv3d_pack_weights(const float* w, void* w_packed, int32_t Cin, int32_t Cout, int32_t kernel,
                                        int32_t transposed, void* stream) {
  SNVC_CHECK_ARG(w && w_packed, "null pointer");
  SNVC_CHECK_ARG(Cin > 0 && Cout > 0 && kernel > 0, "bad dimensions");
  const int CoutPad = round_up(Cout, 16);
  const int K3 = kernel * kernel * kernel;
  const int64_t total = (int64_t)K3 * CoutPad * Cin;
  int blocks = (int)std::min<int64_t>(ceil_div(total, 256), 1024);
  pack_weights_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(w, (__nv_bfloat16*)w_packed, Cin, Cout, CoutPad, K3,
                                                                transposed, total);
  return launch_status("pack_weights_kernel");
}

static int conv3d_fwd_impl(const void* x, const void* w_packed, const float* scale, const float* bias, const void* residual,
                          const float* addend, void* y, const snvc_conv3d_desc* dp, void* stream_);

extern "C" int snvc_conv3d_fwd(const void* x, const void* w_packed, const float* scale, const float* bias,
                               const void* residual, void* y, const snvc_conv3d_desc* dp, void* stream_) {
  return conv3d_fwd_impl(x, w_packed, scale, bias, residual, nullptr, y, dp, stream_);
}

extern "C" int snvc_conv3d_fwd_addend(const void* x, const void* w_packed, const float* scale, const float* bias,
                                      const float* addend, void* y, const snvc_conv3d_desc* dp, void* stream_) {
  SNVC_CHECK_ARG(addend != nullptr && (reinterpret_cast<uintptr_t>(addend) & 15) == 0, "addend must be a 16-byte aligned pointer");
  return conv3d_fwd_impl(x, w_packed, scale, bias, nullptr, addend, y, dp, stream_);
}

static int conv3d_fwd_impl(const void* x, const void* w_packed, const float* scale, const float* bias, const void* residual,
                           const float* addend, void* y, const snvc_conv3d_desc* dp, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  SNVC_CHECK_ARG(dp != nullptr, "desc is null");
  const snvc_conv3d_desc& d = *dp;
  SNVC_CHECK_ARG(x && w_packed && y, "null pointer");
  SNVC_CHECK_ARG(d.Cin == 16 || d.Cin == 32 || d.Cin == 64, "Cin must be 16, 32 or 64 (got %d)", d.Cin);
  SNVC_CHECK_ARG(d.Cout >= 1 && d.Cout <= 64, "Cout must be in [1, 64] (got %d)", d.Cout);
  SNVC_CHECK_ARG(d.N >= 0 && d.Di > 0 && d.Hi > 0 && d.Wi > 0, "bad input extent");
  SNVC_CHECK_ARG(d.out_dtype == SNVC_BF16 || d.out_dtype == SNVC_F32, "out_dtype must be bf16 or f32");
  SNVC_CHECK_ARG(d.residual_mode == 0 || residual != nullptr, "residual_mode set but residual is null");
  SNVC_CHECK_ARG(d.residual_mode == 0 || (d.Cout % 16 == 0 && (d.res_cstride % 8) == 0 && (d.res_coffset % 8) == 0 &&
                                          (reinterpret_cast<uintptr_t>(residual) & 15) == 0),
                 "a residual needs Cout %% 16 == 0 and 16-byte aligned channel rows");
  SNVC_CHECK_ARG(d.in_cstride == 0 || (d.in_cstride % 8 == 0 && d.in_coffset % 8 == 0 && d.in_coffset + d.Cin <= d.in_cstride),
                 "bad input channel slice (in_cstride %d, in_coffset %d)", d.in_cstride, d.in_coffset);
  SNVC_CHECK_ARG((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(w_packed) & 15) == 0 &&
                     (reinterpret_cast<uintptr_t>(y) & 15) == 0,
                 "x, w_packed and y must be 16-byte aligned");
  if (d.N == 0) return 0;

  ConvParams p{};
  p.N = d.N; p.Cin = d.Cin; p.Cout = d.Cout; p.CoutPad = round_up(d.Cout, 16);
  p.Do = d.Do; p.Ho = d.Ho; p.Wo = d.Wo;
  p.K = d.kernel;
  p.relu = d.relu; p.residual_mode = d.residual_mode; p.sigmoid = d.sigmoid; p.out_f32 = d.out_dtype == SNVC_F32;
  p.out_cstride = d.out_cstride ? d.out_cstride : d.Cout;
  p.out_coffset = d.out_coffset;
  p.res_cstride = d.res_cstride ? d.res_cstride : d.Cout;
  p.res_coffset = d.res_coffset;
  SNVC_CHECK_ARG(p.out_coffset + d.Cout <= p.out_cstride, "output channel slice out of range");

  if (!d.transposed) {
    SNVC_CHECK_ARG(d.kernel >= 1 && d.kernel <= 7, "kernel must be in [1,7]");
    SNVC_CHECK_ARG(d.stride == 1 || d.stride == 2, "stride must be 1 or 2");
    SNVC_CHECK_ARG(d.dilation >= 1 && d.pad >= 0, "bad dilation / pad");
    const int ext = d.dilation * (d.kernel - 1) + 1;
    SNVC_CHECK_ARG(d.Do == (d.Di + 2 * d.pad - ext) / d.stride + 1 && d.Ho == (d.Hi + 2 * d.pad - ext) / d.stride + 1 &&
                       d.Wo == (d.Wi + 2 * d.pad - ext) / d.stride + 1,
                   "output extent does not match conv geometry");
    SNVC_CHECK_ARG(d.dilation * (d.kernel - 1) - d.pad <= 127 && d.pad <= 127, "tap offset out of range");
    p.Dj = d.Do; p.Hj = d.Ho; p.Wj = d.Wo;
    p.in_stride = d.stride;
    p.out_stride = 1; p.out_off_d = p.out_off_h = p.out_off_w = 0;
    for (int a = 0; a < 3; ++a) {
      p.nk[a] = d.kernel;
      for (int j = 0; j < d.kernel; ++j) {
        p.off[a][j] = (signed char)(j * d.dilation - d.pad);
        p.kid[a][j] = (signed char)j;
      }
    }
    // SNVC_CONV_MODE=tap forces the per-tap kernel (A/B measurements).  Descriptor base offset stays 0:
    // measured on B200, UMMA swizzles on absolute smem address bits, so row-shifted windows of a
    // TMA-written tile need no base-offset correction (base offset = (addr>>7)&7 gives wrong results).
    const char* mode = opt(OPT_CONV_MODE);
    if (addend) {
      // depth-invariant addend: implemented by the CTA-pair kernel only (3x3x3 s1, Cin = Cout = 32, D >= 2, no residual)
      const int r = launch_kdpair(x, w_packed, scale, bias, nullptr, y, d, p, stream, 0, 0, addend);
      if (r == 1) return fail(SNVC_E_UNSUPPORTED, "snvc_conv3d_fwd_addend: needs a 3x3x3 stride-1 conv with Cin = Cout = 32, D >= 2");
      return r;
    }
    if (!(mode && mode[0] == 't')) {
      int r = launch_s2(x, w_packed, scale, bias, residual, y, d, p, stream);
      if (r != 1) return r;
      r = launch_bigk(x, w_packed, scale, bias, residual, y, d, p, stream);
      if (r != 1) return r;
      r = launch_kdpair(x, w_packed, scale, bias, residual, y, d, p, stream);
      if (r != 1) return r;
      r = launch_kwfuse(x, w_packed, scale, bias, residual, y, d, p, stream);
      if (r != 1) return r;
      r = launch_halo(x, w_packed, scale, bias, residual, y, d, p, stream, 0);
      if (r != 1) return r;
      // 64 -> 64: all 27 weight tiles (221 KB) do not fit next to the plane ring; run the plane march twice on
      // 32-channel output slices (the input, at most half resolution on this path, is read twice from L2/HBM)
      if (p.CoutPad == 64 && d.Cout == 64 && d.dilation == 1 && d.kernel == 3 && d.stride == 1 &&
          ((p.out_cstride | p.out_coffset) & 7) == 0) {
        r = launch_kdpair(x, w_packed, scale, bias, residual, y, d, p, stream, 0, 32);
        if (r == 0) r = launch_kdpair(x, w_packed, scale, bias, residual, y, d, p, stream, 32, 32);
        if (r != 1) return r;
        r = launch_kwfuse(x, w_packed, scale, bias, residual, y, d, p, stream, 0, 32);
        if (r == 0) r = launch_kwfuse(x, w_packed, scale, bias, residual, y, d, p, stream, 32, 32);
        if (r != 1) return r;
        r = launch_halo(x, w_packed, scale, bias, residual, y, d, p, stream, 0, 0, 32);
        if (r == 0) r = launch_halo(x, w_packed, scale, bias, residual, y, d, p, stream, 0, 32, 32);
        if (r != 1) return r;
      }
    }
    return launch_conv(x, w_packed, scale, bias, residual, y, d, p, stream);
  }

  if (addend) return fail(SNVC_E_UNSUPPORTED, "snvc_conv3d_fwd_addend: transposed convolutions are not supported");
  // ConvTranspose3d(k=3, s=2, p=1, output_padding=1): out[o] += x[i] * W[k], o = 2i - 1 + k.
  //   even o = 2j   : k=1, i=j
  //   odd  o = 2j+1 : k=2, i=j   and   k=0, i=j+1
  SNVC_CHECK_ARG(d.kernel == 3 && d.stride == 2 && d.pad == 1 && d.dilation == 1,
                 "transposed conv supports k=3, s=2, p=1, output_padding=1 only");
  SNVC_CHECK_ARG(d.Do == 2 * d.Di && d.Ho == 2 * d.Hi && d.Wo == 2 * d.Wi, "transposed conv output must be 2x input");
  {
    const char* mode = opt(OPT_CONV_MODE);
    if (!(mode && mode[0] == 't')) {                     // SNVC_CONV_MODE=tap: per-class launches (A/B runs)
      int r = 1;
      for (int c0 = 0; c0 < d.Cout; c0 += kDeconvCP) {
        r = launch_deconv(x, w_packed, scale, bias, residual, y, d, p, stream, c0);
        if (r != 0) break;
      }
      if (r != 1) return r;
    }
  }
  p.Dj = d.Di; p.Hj = d.Hi; p.Wj = d.Wi;
  p.in_stride = 1;
  p.out_stride = 2;
  for (int cls = 0; cls < 8; ++cls) {
    const int par[3] = {(cls >> 2) & 1, (cls >> 1) & 1, cls & 1};
    for (int a = 0; a < 3; ++a) {
      if (par[a] == 0) {
        p.nk[a] = 1; p.off[a][0] = 0; p.kid[a][0] = 1;
      } else {
        p.nk[a] = 2; p.off[a][0] = 0; p.kid[a][0] = 2; p.off[a][1] = 1; p.kid[a][1] = 0;
      }
    }
    p.out_off_d = par[0]; p.out_off_h = par[1]; p.out_off_w = par[2];
    if (int e = launch_conv(x, w_packed, scale, bias, residual, y, d, p, stream)) return e;
  }
  return 0;
}
