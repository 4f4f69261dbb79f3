// conv3d_v3_kdfuse.cuh -- one kernel generation of csrc/conv3d_tcgen05.cu (product path for Cout = 16 / 64 slices, sigmoid / Cout = 1 epilogues; SNVC_CONV_MODE=kd).
// Included by conv3d_tcgen05.cu INSIDE `namespace snvc { namespace {`, after the shared parameter structs, the PTX
// wrappers (tcgen05.cuh) and the per-tap kernel; the host-side launcher of this kernel stays in conv3d_tcgen05.cu.
#pragma once

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
