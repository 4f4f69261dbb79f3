// conv3d_v2_halo.cuh -- one kernel generation of csrc/conv3d_tcgen05.cu (A/B only (SNVC_CONV_MODE=halo) and the HaloParams struct shared by v3 / v7 / v8).
// Included by conv3d_tcgen05.cu INSIDE `namespace snvc { namespace {`, after the shared parameter structs, the PTX
// wrappers (tcgen05.cuh) and the per-tap kernel; the host-side launcher of this kernel stays in conv3d_tcgen05.cu.
#pragma once

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
