// conv3d_v4_deconv.cuh -- one kernel generation of csrc/conv3d_tcgen05.cu (product path: every transposed convolution).
// Included by conv3d_tcgen05.cu INSIDE `namespace snvc { namespace {`, after the shared parameter structs, the PTX
// wrappers (tcgen05.cuh) and the per-tap kernel; the host-side launcher of this kernel stays in conv3d_tcgen05.cu.
#pragma once

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
