// conv3d_v5_s2.cuh -- one kernel generation of csrc/conv3d_tcgen05.cu (product path: stride-2 3x3x3 with Cin = 32).
// Included by conv3d_tcgen05.cu INSIDE `namespace snvc { namespace {`, after the shared parameter structs, the PTX
// wrappers (tcgen05.cuh) and the per-tap kernel; the host-side launcher of this kernel stays in conv3d_tcgen05.cu.
#pragma once

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
