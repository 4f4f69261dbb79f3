// conv3d_v7_kwfuse.cuh -- one kernel generation of csrc/conv3d_tcgen05.cu (A/B only (SNVC_CONV_MODE=kw): geometry-independent summation order, used by the exact depth-slab tests).
// Included by conv3d_tcgen05.cu INSIDE `namespace snvc { namespace {`, after the shared parameter structs, the PTX
// wrappers (tcgen05.cuh) and the per-tap kernel; the host-side launcher of this kernel stays in conv3d_tcgen05.cu.
#pragma once

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
