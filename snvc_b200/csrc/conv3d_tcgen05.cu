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
// The layers that carry the FLOPs run on the plane-march kernels, ONE FILE PER GENERATION (conv3d_v*.cuh, included below
// inside this translation unit's anonymous namespace; their host-side launchers and the dispatch stay in this file).  v2 and
// v7 are reachable only through the SNVC_CONV_MODE debug switch (A/B runs, exact depth-slab tests):
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


#include "conv3d_v2_halo.cuh"   // v2: A/B only (SNVC_CONV_MODE=halo) and the HaloParams struct shared by v3 / v7 / v8

#include "conv3d_v3_kdfuse.cuh"   // v3: product path for Cout = 16 / 64 slices, sigmoid / Cout = 1 epilogues; SNVC_CONV_MODE=kd

#include "conv3d_v7_kwfuse.cuh"   // v7: A/B only (SNVC_CONV_MODE=kw): geometry-independent summation order, used by the exact depth-slab tests

#include "conv3d_v8_kdpair.cuh"   // v8: product path: the default for every 3x3x3 stride-1 layer with 32-channel output slices

#include "conv3d_v4_deconv.cuh"   // v4: product path: every transposed convolution

#include "conv3d_v5_s2.cuh"   // v5: product path: stride-2 3x3x3 with Cin = 32

#include "conv3d_v6_bigk.cuh"   // v6: product path: the 5^3 / 7^3 layers of the instance branch

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
