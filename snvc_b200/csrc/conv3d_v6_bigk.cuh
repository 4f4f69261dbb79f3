// conv3d_v6_bigk.cuh -- one kernel generation of csrc/conv3d_tcgen05.cu (product path: the 5^3 / 7^3 layers of the instance branch).
// Included by conv3d_tcgen05.cu INSIDE `namespace snvc { namespace {`, after the shared parameter structs, the PTX
// wrappers (tcgen05.cuh) and the per-tap kernel; the host-side launcher of this kernel stays in conv3d_tcgen05.cu.
#pragma once

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
