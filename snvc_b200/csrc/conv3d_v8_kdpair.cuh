// conv3d_v8_kdpair.cuh -- one kernel generation of csrc/conv3d_tcgen05.cu (product path: the default for every 3x3x3 stride-1 layer with 32-channel output slices).
// Included by conv3d_tcgen05.cu INSIDE `namespace snvc { namespace {`, after the shared parameter structs, the PTX
// wrappers (tcgen05.cuh) and the per-tap kernel; the host-side launcher of this kernel stays in conv3d_tcgen05.cu.
#pragma once

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
