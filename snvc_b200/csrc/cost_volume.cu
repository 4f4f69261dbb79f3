// A1 / A1b -- plane-sweep cost volume for sm_100a.
//
// Semantics: snvc/extension/build_cost_volume/src/BuildCostVolume_cuda.cu:63-98 (forward),
// :15-61 (bilinear rule), :152-205 (backward).  Design (not the reference's one-thread-per-
// element gather): every CTA stages the few input rows it needs in shared memory ONCE and emits
// all depth bins from there, so HBM sees one read of the features and one fully coalesced
// 128-bit write stream of the volume.  Two layouts:
//   * NCDHW fp32/fp64 -- the reference's output layout (drop-in for build_cost_volume);
//   * NDHWC bf16      -- channels-last, exactly what the tcgen05 conv3d TMA-loads
//                        (halves the dominant write stream).
// Arithmetic is written with explicit round-to-nearest intrinsics so that the result is
// bit-identical to the nvcc-compiled reference expression  w1*v1 + w2*v2 + w3*v3 + w4*v4
// as nvcc contracts it with its default -fmad=true: fma(w1, v1, w2*v2) -- the SECOND product is rounded, the first
// is fused (read off the SASS of the reference kernel compiled by oracle/build_ref.py: FMUL v2*w2; FFMA v1*w1 + .;
// tests/test_gpu_ref_pin.py compares against that build bit for bit); the w3/w4 terms are exact zeros because y is
// an integer.
#include <cmath>

#include "common.cuh"

namespace snvc {
namespace {

template <typename T> struct Arith;
template <> struct Arith<float> {
  static __device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
  static __device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
  static __device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
  static __device__ __forceinline__ float fma(float a, float b, float c) { return __fmaf_rn(a, b, c); }
};
template <> struct Arith<double> {
  static __device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
  static __device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
  static __device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
  static __device__ __forceinline__ double fma(double a, double b, double c) { return __fma_rn(a, b, c); }
};

// Sample position for output column pw at depth shift s (.cu:84-96, :26-49).
// Returns false when the sample is outside the image (value 0).  x_low/x_high/lx as the reference.
template <typename T>
__device__ __forceinline__ bool sample_pos(int iw, T neg_shift, int img_w, int& x_low, int& x_high, T& lx) {
  T x = Arith<T>::add((T)iw, neg_shift);
  if (!(x >= (T)0 && x <= (T)(img_w - 1))) return false;
  if (x <= (T)0) x = (T)0;
  x_low = (int)x;
  if (x_low >= img_w - 1) {
    x_high = x_low = img_w - 1;
    x = (T)x_low;
  } else {
    x_high = x_low + 1;
  }
  lx = Arith<T>::sub(x, (T)x_low);
  return true;
}

template <typename T>
__device__ __forceinline__ T interp(T lx, T v1, T v2) {
  T hx = Arith<T>::sub((T)1, lx);
  return Arith<T>::fma(hx, v1, Arith<T>::mul(lx, v2));
}

// ------------------------------------------------------------------------------------------
// NCDHW.  grid = (h-tiles, N*C, d-splits);  smem: TH rows of left and right (+ shifts).
// ------------------------------------------------------------------------------------------
template <typename T, bool VEC4>
__global__ void __launch_bounds__(256)
cv_ncdhw_kernel(const T* __restrict__ left, const T* __restrict__ right, const T* __restrict__ shift,
                T* __restrict__ cost, int C, int img_h, int img_w, int D, int H, int W, int ds, int TH,
                int d_per_cta) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* sL = reinterpret_cast<T*>(smem_raw);
  T* sR = sL + (size_t)TH * img_w;
  T* sS = sR + (size_t)TH * img_w;

  const int nc = blockIdx.y;
  const int n = nc / C;
  const int c = nc - n * C;
  const int ph0 = blockIdx.x * TH;
  const int th = min(TH, H - ph0);
  const int d0 = blockIdx.z * d_per_cta;
  const int dn = min(d_per_cta, D - d0);
  if (dn <= 0) return;

  const T* lplane = left + (int64_t)nc * img_h * img_w;
  const T* rplane = right + (int64_t)nc * img_h * img_w;
  for (int i = threadIdx.x; i < th * img_w; i += blockDim.x) {
    int r = i / img_w, col = i - r * img_w;
    int64_t g = (int64_t)(ph0 + r) * ds * img_w + col;
    sL[r * img_w + col] = lplane[g];
    sR[r * img_w + col] = rplane[g];
  }
  for (int i = threadIdx.x; i < dn; i += blockDim.x) sS[i] = -shift[(int64_t)n * D + d0 + i];
  __syncthreads();

  const int64_t plane = (int64_t)H * W;
  T* outL = cost + (((int64_t)n * 2 * C + c) * D + d0) * plane + (int64_t)ph0 * W;
  T* outR = outL + (int64_t)C * D * plane;

  if (VEC4) {
    // W % 4 == 0 and T == float: one float4 per thread-iteration
    const int W4 = W >> 2;
    const int per_d = th * W4;
    const int total = dn * per_d;
    for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
      int dd = idx / per_d;
      int rem = idx - dd * per_d;
      int r = rem / W4;
      int pw = (rem - r * W4) << 2;
      const T ns = sS[dd];
      const T* rowL = sL + r * img_w;
      const T* rowR = sR + r * img_w;
      float4 vl, vr;
      float* pl = reinterpret_cast<float*>(&vl);
      float* pr = reinterpret_cast<float*>(&vr);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        int iw = (pw + j) * ds;
        pl[j] = (float)rowL[iw];
        int xl, xh;
        T lx;
        pr[j] = sample_pos<T>(iw, ns, img_w, xl, xh, lx) ? (float)interp<T>(lx, rowR[xl], rowR[xh]) : 0.f;
      }
      int64_t o = (int64_t)dd * plane + (int64_t)r * W + pw;
      st_cs_f4(reinterpret_cast<float*>(outL + o), vl);
      st_cs_f4(reinterpret_cast<float*>(outR + o), vr);
    }
  } else {
    const int per_d = th * W;
    const int total = dn * per_d;
    for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
      int dd = idx / per_d;
      int rem = idx - dd * per_d;
      int r = rem / W;
      int pw = rem - r * W;
      int iw = pw * ds;
      int xl, xh;
      T lx;
      int64_t o = (int64_t)dd * plane + (int64_t)r * W + pw;
      outL[o] = sL[r * img_w + iw];
      outR[o] = sample_pos<T>(iw, sS[dd], img_w, xl, xh, lx)
                    ? interp<T>(lx, sR[r * img_w + xl], sR[r * img_w + xh])
                    : (T)0;
    }
  }
}

// ------------------------------------------------------------------------------------------
// NDHWC bf16 (fp32 in).  grid = (H, N, d-splits * w-tiles).
// smem holds one input row of every channel, transposed to [col][C] with a 16-byte XOR swizzle
// (conflict-free float4 reads by 8-channel lanes, <=4-way conflicts on the one-off staging).
// ------------------------------------------------------------------------------------------
// 16-byte chunk q of column `col` lives at chunk (q ^ (col & mask)); mask = 2^k - 1 with 2^k | C/4.
__device__ __forceinline__ int swz(int col, int c, int C, int mask) {
  return col * C + ((((c >> 2) ^ (col & mask)) << 2) | (c & 3));
}
__device__ __forceinline__ const float4* chunk_ptr(const float* base, int col, int q, int C, int mask) {
  return reinterpret_cast<const float4*>(base + col * C + ((q ^ (col & mask)) << 2));
}

__device__ __forceinline__ void cp_async_4(void* smem_dst, const void* gmem_src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)),
               "l"(gmem_src)
               : "memory");
}

// Each thread owns (output column, 8-channel group) pairs and walks the depth bins: the left half is
// packed once and re-stored per bin, the right half costs one sample_pos + 4 LDS.128 + 8 FMAs per bin.
// Staging uses 4-byte cp.async (no register round trip, all loads in flight at once).
__global__ void __launch_bounds__(512)
cv_ndhwc_bf16_kernel(const float* __restrict__ left, const float* __restrict__ right,
                     const float* __restrict__ shift, __nv_bfloat16* __restrict__ cost,
                     __nv_bfloat16* __restrict__ left_planes /* split form: [N,3,H,W,C], else nullptr */, int C, int img_h,
                     int img_w, int D, int H, int W, int ds, int d_per_cta, int n_dsplit, int TW, int S, int mask) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int ph = blockIdx.x;
  const int n = blockIdx.y;
  const int dsplit = blockIdx.z % n_dsplit;
  const int wtile = blockIdx.z / n_dsplit;
  const int d0 = dsplit * d_per_cta;
  const int dn = min(d_per_cta, D - d0);
  const int pw0 = wtile * TW;
  const int tw = min(TW, W - pw0);
  if (dn <= 0 || tw <= 0) return;

  // input column windows staged in smem
  const int llo = pw0 * ds, lhi = (pw0 + tw - 1) * ds + 1;              // left: [llo, lhi)
  const int rlo = max(0, llo - S), rhi = min(img_w, lhi + 1);            // right: [rlo, rhi)
  const int lw = lhi - llo, rw = rhi - rlo;
  float* sL = reinterpret_cast<float*>(smem_raw);
  float* sR = sL + (size_t)lw * C;
  float* sS = sR + (size_t)rw * C;

  const int ih = ph * ds;
  const float* lrow = left + ((int64_t)n * C * img_h + ih) * img_w;      // + c*img_h*img_w + col
  const float* rrow = right + ((int64_t)n * C * img_h + ih) * img_w;
  const int64_t cstride = (int64_t)img_h * img_w;
  for (int i = threadIdx.x; i < C * lw; i += blockDim.x) {
    int c = i / lw, col = i - c * lw;
    cp_async_4(&sL[swz(col, c, C, mask)], lrow + c * cstride + llo + col);
  }
  for (int i = threadIdx.x; i < C * rw; i += blockDim.x) {
    int c = i / rw, col = i - c * rw;
    cp_async_4(&sR[swz(col, c, C, mask)], rrow + c * cstride + rlo + col);
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  for (int i = threadIdx.x; i < dn; i += blockDim.x) sS[i] = -shift[(int64_t)n * D + d0 + i];
  // Sample positions of every (column, bin) of this CTA, computed ONCE: x_low / x_high / lx depend on the column and
  // the bin only, but the C/8 lanes of a column each recomputed them per bin (~10 of the ~21 instructions per lane and
  // bin; with the left half gone the kernel was issue-bound at 48 % of the HBM peak).  Entry = {code, lx}; code = -1
  // outside the image, else x_low | (x_high != x_low) << 30.
  int2* sT = reinterpret_cast<int2*>(sS + ((d_per_cta + 1) & ~1));
  for (int i = threadIdx.x; i < dn * tw; i += blockDim.x) {
    const int dd = i / tw, wl = i - dd * tw;
    int xl, xh;
    float lx;
    const bool ok = sample_pos<float>((pw0 + wl) * ds, -shift[(int64_t)n * D + d0 + dd], img_w, xl, xh, lx);
    sT[i] = make_int2(ok ? (xl | (xh != xl ? 0x40000000 : 0)) : -1, __float_as_int(lx));
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();

  const int CG = C >> 3;                // 8-channel groups per view
  // split form (snvc_cost_volume_split_fwd): the volume holds the right half only (C channels per voxel) and the
  // left half -- a pure broadcast over depth, BuildCostVolume_cuda.cu:84-86 -- is written once, as three planes
  const bool split = left_planes != nullptr;
  const int C2 = split ? C : 2 * C;
  const int roff = split ? 0 : C;
  const int64_t dstride = (int64_t)H * W * C2;
  for (int pair = threadIdx.x; pair < tw * CG; pair += blockDim.x) {
    const int wl = pair / CG, cg = pair - wl * CG;
    const int pw = pw0 + wl, iw = pw * ds;
    __nv_bfloat16* o = cost + ((((int64_t)n * D + d0) * H + ph) * W + pw) * C2 + cg * 8;
    uint4 vl;
    {
      const int col = iw - llo;
      const float4 a = *chunk_ptr(sL, col, 2 * cg, C, mask);
      const float4 b = *chunk_ptr(sL, col, 2 * cg + 1, C, mask);
      vl = make_uint4(pack_bf16x2(a.x, a.y), pack_bf16x2(a.z, a.w), pack_bf16x2(b.x, b.y), pack_bf16x2(b.z, b.w));
    }
    // The two right-image samples of a bin are kept in registers: the shift shrinks with depth, so from one bin to
    // the next x_low stays or advances by one column for most bins (all lanes of a warp alike: iw is an integer),
    // and 0 or 2 instead of 4 LDS.128 are needed -- the shared-memory pipe, not HBM, was the busiest unit (ncu: 88 %).
    int cur_xl = -0x40000000, cur_xh = -0x40000000;
    float r0[8], r1[8];
    if (split && dsplit == 0) {
      __nv_bfloat16* lo = left_planes + ((((int64_t)n * 3) * H + ph) * W + pw) * C + cg * 8;
#pragma unroll
      for (int v = 0; v < 3; ++v) *reinterpret_cast<uint4*>(lo + (int64_t)v * H * W * C) = vl;
    }
    for (int dd = 0; dd < dn; ++dd, o += dstride) {
      if (!split) *reinterpret_cast<uint4*>(o) = vl;           // left half: broadcast over depth
      const int2 te = sT[dd * tw + wl];
      const int xl = te.x & 0x3fffffff, xh = xl + ((te.x >> 30) & 1);
      const float lx = __int_as_float(te.y);
      uint4 v = make_uint4(0u, 0u, 0u, 0u);
      if (te.x >= 0) {                                         // right half: 1-D interpolation along the row
        if (xl >= rlo && xh < rhi) {
          if (xl != cur_xl || xh != cur_xh) {
            if (xl == cur_xh) {
#pragma unroll
              for (int j = 0; j < 8; ++j) r0[j] = r1[j];
            } else {
              const int c0 = xl - rlo;
              *reinterpret_cast<float4*>(r0) = *chunk_ptr(sR, c0, 2 * cg, C, mask);
              *reinterpret_cast<float4*>(r0 + 4) = *chunk_ptr(sR, c0, 2 * cg + 1, C, mask);
            }
            if (xh == xl) {
#pragma unroll
              for (int j = 0; j < 8; ++j) r1[j] = r0[j];
            } else {
              const int c1 = xh - rlo;
              *reinterpret_cast<float4*>(r1) = *chunk_ptr(sR, c1, 2 * cg, C, mask);
              *reinterpret_cast<float4*>(r1 + 4) = *chunk_ptr(sR, c1, 2 * cg + 1, C, mask);
            }
            cur_xl = xl; cur_xh = xh;
          }
        } else {
          // sample outside the staged window (very large shift on a w-tiled row -> left of it; a NEGATIVE shift, which the
          // reference's Python layer rejects but this entry point cannot see, -> right of it): read HBM directly
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            r0[j] = rrow[(int64_t)(cg * 8 + j) * cstride + xl];
            r1[j] = rrow[(int64_t)(cg * 8 + j) * cstride + xh];
          }
          cur_xl = xl; cur_xh = xh;
        }
        const float hx = __fsub_rn(1.f, lx);
        float r[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) r[j] = __fmaf_rn(hx, r0[j], __fmul_rn(lx, r1[j]));
        v = make_uint4(pack_bf16x2(r[0], r[1]), pack_bf16x2(r[2], r[3]), pack_bf16x2(r[4], r[5]), pack_bf16x2(r[6], r[7]));
      }
      *reinterpret_cast<uint4*>(o + roff) = v;
    }
  }
}

// Left planes of the split form: left_planes[n, v, ph, pw, :] = bf16(left[n, :, ph*ds, pw*ds]) for v = 0, 1, 2 -- a
// transposition of 30 MB.  Lanes run over (pw, 8-channel group) with the group fastest: a warp reads, per channel of the
// group, four 32-byte runs of the NCHW rows and writes 512 contiguous bytes of each plane; no shared memory.  (Round 2:
// this was a second launch of the staged split kernel, 19 us of mostly cp.async latency for 0.1 % of the step's bytes.)
__global__ void __launch_bounds__(256)
cv_left_planes_kernel(const float* __restrict__ left, __nv_bfloat16* __restrict__ left_planes, int C, int img_h, int img_w,
                      int H, int W, int ds, int64_t total /* N*H*W*C/8 */) {
  const int CG = C >> 3;
  const int64_t cstride = (int64_t)img_h * img_w;
  const int64_t plane = (int64_t)H * W * C;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int cg = (int)(i % CG);
    int64_t t = i / CG;
    const int pw = (int)(t % W); t /= W;
    const int ph = (int)(t % H);
    const int64_t n = t / H;
    const float* src = left + ((n * C + cg * 8) * img_h + (int64_t)ph * ds) * img_w + (int64_t)pw * ds;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = __ldg(src + j * cstride);
    const uint4 q = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
    __nv_bfloat16* o = left_planes + (n * 3 * H + ph) * (int64_t)W * C + (int64_t)pw * C + cg * 8;
#pragma unroll
    for (int p3 = 0; p3 < 3; ++p3) *reinterpret_cast<uint4*>(o + p3 * plane) = q;
  }
}

// ------------------------------------------------------------------------------------------
// Split form (snvc_cost_volume_split_fwd), whole rows in shared memory.  ncu on the kernel above in split mode: 248 M
// warp instructions for 0.84 GB (issue slots 83 % busy, 48 % of the HBM peak) -- a lane owns a (column, 8 channels)
// and walks the depth bins, so for every bin it re-derives which two right-image columns it needs, compares them with
// the ones it holds, and re-loads or moves 16 registers on one of several paths.
// Here a lane owns a (depth bin, 8 channels) and walks a STRIP of consecutive columns: for a fixed bin the sample
// position advances by exactly one column per step, so the "high" sample of one step is the "low" sample of the
// next and each step costs ONE shared-memory column (2 LDS.128), 8 FMUL + 8 FFMA, 4 packs and a 16-byte store.  The
// per-(bin, column) positions (x_low, x_high, lx -- fp32 arithmetic in the reference's order, so the result stays
// bit-identical) come from a table built once per CTA; the left planes are written by the CTAs of the first depth
// split.  grid = (H, N, depth splits).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(512)
cv_split_bf16_kernel(const float* __restrict__ right, const float* __restrict__ shift, __nv_bfloat16* __restrict__ right_vol,
                     int C, int img_h, int img_w, int D, int H, int W, int ds, int d_per_cta, int mask, int general_only) {
  // right half only (grid.z = depth splits; shared memory = one right row + the sample table -> 3 CTAs per SM); the left
  // planes are cv_left_planes_kernel's, so that no left row sits in the shared memory of every depth split
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int ph = blockIdx.x, n = blockIdx.y, dsplit = blockIdx.z;
  const int d0 = dsplit * d_per_cta;
  const int dn = min(d_per_cta, D - d0);
  if (dn <= 0) return;
  float* sR = reinterpret_cast<float*>(smem_raw);       // [img_w][C], 16-byte chunks XOR-swizzled by column
  int2* sT = reinterpret_cast<int2*>(sR + (size_t)img_w * C);   // [dn][W]: {code, lx}
  const int ih = ph * ds;
  const int64_t cstride = (int64_t)img_h * img_w;
  const float* rrow = right + ((int64_t)n * C * img_h + ih) * img_w;
  // staging: a warp takes one channel at a time and runs along the row (coalesced 4-byte cp.async, no divisions).  The
  // transposing scatter is a 4-way bank conflict (a fifth of the kernel's shared-memory wavefronts, ncu r02); the
  // conflict-free mapping -- 8 columns x the 4 channels of a chunk per warp instruction -- was measured SLOWER (180 vs
  // 156 us: every lane quartet then pulls a different 32-byte sector, four times the L2 requests) and dropped.
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
  for (int c = warp; c < C; c += nwarp)
    for (int col = lane; col < img_w; col += 32) cp_async_4(&sR[swz(col, c, C, mask)], rrow + c * cstride + col);
  asm volatile("cp.async.commit_group;" ::: "memory");
  // sReg[dd] = first column from which the bin is REGULAR up to the end of the row: every sample is inside the image, reads two
  // different columns and sits exactly one column right of its predecessor's (true right of the zero region x < 0, unless an
  // fp32 rounding of pw - shift jumps an integer or the last column is clamped).  Strips inside that range take the
  // branch-free walk below; the table decides, so the result stays bit-identical.
  int* sReg = reinterpret_cast<int*>(sT + (size_t)dn * W);
  for (int dd = warp; dd < dn; dd += nwarp) {
    const float ns = -shift[(int64_t)n * D + d0 + dd];
    int reg_lo = 0, prev_last = -0x40000000;              // prev_last: x_low of column (chunk base - 1)
    for (int pw0 = 0; pw0 < W; pw0 += 32) {
      const int pw = pw0 + lane;
      int xl = -0x40000000, xh = 0;
      float lx = 0.f;
      const bool ok = pw < W && sample_pos<float>(pw * ds, ns, img_w, xl, xh, lx);
      if (pw < W) sT[dd * W + pw] = make_int2(ok ? (xl | (xh != xl ? 0x40000000 : 0)) : -1, __float_as_int(lx));
      const int xl_ok = ok ? xl : -0x40000000;
      int before = __shfl_up_sync(0xffffffffu, xl_ok, 1);
      if (lane == 0) before = prev_last;
      const bool regular = ok && xh == xl + 1 && (pw == 0 || xl == before + 1);
      const unsigned irr = __ballot_sync(0xffffffffu, pw < W && !regular);
      if (irr) reg_lo = pw0 + 32 - __clz(irr);            // one past the last irregular column so far
      prev_last = __shfl_sync(0xffffffffu, xl_ok, 31);
    }
    if (lane == 0) sReg[dd] = (ds == 1 && !general_only) ? reg_lo : W;   // (SNVC_CV_WALK=general: A/B switch)
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();

  const int CG = C >> 3;                                // 8-channel groups
  // right half: item = (bin, strip, channel group); the CG lanes of an item's pixel are adjacent (64-byte segments)
  const int nstrips = max(1, (int)blockDim.x / (dn * CG));
  // The row is walked in two regions so that a warp never mixes the two kinds of step: [0, R0) holds the irregular columns
  // of every bin of this CTA (the zero region left of x = 0: up to max shift columns) and takes the general step; [R0, W)
  // is regular for all bins and takes the branch-free step.  (First version: per-strip choice -- the 8 strips of a bin
  // share a warp, so every warp ran both paths and the kernel got SLOWER, 0.22 vs 0.185 ms.)
  int R0 = 0;
  for (int dd = 0; dd < dn; ++dd) R0 = max(R0, sReg[dd]);
  R0 = min(R0, W);
  auto load_col_g = [&](float2 (&dst)[4], int col, int cg) {
    const float4 a = *chunk_ptr(sR, col, 2 * cg, C, mask), b = *chunk_ptr(sR, col, 2 * cg + 1, C, mask);
    dst[0] = make_float2(a.x, a.y); dst[1] = make_float2(a.z, a.w); dst[2] = make_float2(b.x, b.y); dst[3] = make_float2(b.z, b.w);
  };
  // ---- region A = [R0, W): regular.  Column x_low advances by one per step, so a step is ONE new column (2 LDS.128), the
  // interpolation weight from the table, 4 FMUL2 + 4 FFMA2, 4 packs and the store -- no position decode, no branches.
  // odd strip length: the 8 lane groups of a warp then sit on columns with 8 different (col & 7), i.e. 8 different
  // swizzle phases -> every LDS.128 is served in the minimum 4 wavefronts (an even length -- 26 for KITTI -- measured
  // 27 M bank conflicts per launch, L1 data pipe 85 % busy)
  if (R0 < W) {
    const int LA = ((W - R0 + nstrips - 1) / nstrips) | 1;
    const int items = dn * nstrips * CG;
    for (int it = threadIdx.x; it < items; it += blockDim.x) {
      const int cg = it % CG;
      const int t2 = it / CG;
      const int strip = t2 % nstrips, dd = t2 / nstrips;
      const int w_begin = R0 + strip * LA, w_end = min(W, w_begin + LA);
      if (w_begin >= w_end) continue;
      const int2* tab = sT + dd * W;
      __nv_bfloat16* o = right_vol + ((((int64_t)n * D + d0 + dd) * H + ph) * W + w_begin) * C + cg * 8;
      // two register sets that swap roles every step (unrolled by two): the column shared by consecutive steps never moves
      float2 ra[4], rb[4];
      const int x0 = tab[w_begin].x & 0x3fffffff;
      load_col_g(ra, x0, cg);
      auto fstep = [&](const float2 (&lo)[4], float2 (&hi)[4], int col_hi, int p, __nv_bfloat16* op) {
        load_col_g(hi, col_hi, cg);
        const float lx = __int_as_float(tab[p].y);
        const float hx = __fsub_rn(1.f, lx);
        const float2 lx2 = make_float2(lx, lx), hx2 = make_float2(hx, hx);
        float2 r[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) r[j] = __ffma2_rn(hx2, lo[j], __fmul2_rn(lx2, hi[j]));
        *reinterpret_cast<uint4*>(op) = make_uint4(pack_bf16x2(r[0].x, r[0].y), pack_bf16x2(r[1].x, r[1].y),
                                                   pack_bf16x2(r[2].x, r[2].y), pack_bf16x2(r[3].x, r[3].y));
      };
      int pw = w_begin, xh = x0 + 1;
      for (; pw + 1 < w_end; pw += 2, xh += 2, o += 2 * C) {
        fstep(ra, rb, xh, pw, o);
        fstep(rb, ra, xh + 1, pw + 1, o + C);
      }
      if (pw < w_end) fstep(ra, rb, xh, pw, o);
    }
  }
  // ---- region B = [0, R0): the general step (values as packed fp32 pairs: one FMUL2 + one FFMA2 per two channels, the same
  // roundings as the scalar lx*v2 then fma(hx, v1, .)); short odd strips over all threads
  if (R0 > 0) {
    const int LB = ((R0 + nstrips - 1) / nstrips) | 1;
    const int items = dn * nstrips * CG;
    for (int it = threadIdx.x; it < items; it += blockDim.x) {
      const int cg = it % CG;
      const int t2 = it / CG;
      const int strip = t2 % nstrips, dd = t2 / nstrips;
      const int w_begin = strip * LB, w_end = min(R0, w_begin + LB);
      if (w_begin >= w_end) continue;
      const int2* tab = sT + dd * W;
      __nv_bfloat16* o = right_vol + ((((int64_t)n * D + d0 + dd) * H + ph) * W + w_begin) * C + cg * 8;
      int held = -0x40000000;                             // column held in the set that is "high" after the last step
      float2 ra[4], rb[4];
      // one step: `lo` must end up holding column xl, `hi` column xh; on entry `lo` holds column `held` (the previous
      // step's high column), `hi` is free
      auto step = [&](float2 (&lo)[4], float2 (&hi)[4], int pw, __nv_bfloat16* op) {
        const int2 te = tab[pw];
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (te.x >= 0) {
          const int xl = te.x & 0x3fffffff, xh = xl + ((te.x >> 30) & 1);
          const float lx = __int_as_float(te.y);
          if (xl != held) load_col_g(lo, xl, cg);
          if (xh != xl) load_col_g(hi, xh, cg);
          else {
#pragma unroll
            for (int j = 0; j < 4; ++j) hi[j] = lo[j];
          }
          held = xh;
          const float hx = __fsub_rn(1.f, lx);
          const float2 lx2 = make_float2(lx, lx), hx2 = make_float2(hx, hx);
          float2 r[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) r[j] = __ffma2_rn(hx2, lo[j], __fmul2_rn(lx2, hi[j]));
          v = make_uint4(pack_bf16x2(r[0].x, r[0].y), pack_bf16x2(r[1].x, r[1].y), pack_bf16x2(r[2].x, r[2].y),
                         pack_bf16x2(r[3].x, r[3].y));
        } else {
          held = -0x40000000;                             // nothing usable is held (hi was not written)
        }
        *reinterpret_cast<uint4*>(op) = v;
      };
      int pw = w_begin;
      for (; pw + 1 < w_end; pw += 2, o += 2 * C) {
        step(ra, rb, pw, o);                              // after it: rb holds the high column
        step(rb, ra, pw + 1, o + C);                      // rb is the low set now; after it: ra holds the high column
      }
      if (pw < w_end) step(ra, rb, pw, o);
    }
  }
}

// ------------------------------------------------------------------------------------------
// Backward: gather formulation, one thread per input pixel, fixed summation order
// (d ascending, then the <=2 output columns that touch this pixel) -> deterministic, no atomics.
//   grad_left [n,c,ih,iw]  = sum_d grad[n, c, d, ih/ds, iw/ds]           (iw, ih multiples of ds)
//   grad_right[n,c,ih,x]   = sum_d sum_{pw : x_low==x or x_high==x} w * grad[n, C+c, d, ih/ds, pw]
// Semantics .cu:152-205 (weights below 1e-10 are skipped as in :195-202).
// ------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
cv_bwd_kernel(const T* __restrict__ grad, const T* __restrict__ shift, T* __restrict__ gl, T* __restrict__ gr,
              int C, int H, int W, int D, int ds, int64_t total) {
  const int img_h = H * ds, img_w = W * ds;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int x = (int)(i % img_w);
    int64_t t = i / img_w;
    int ih = (int)(t % img_h);
    int64_t nc = t / img_h;
    int n = (int)(nc / C), c = (int)(nc - (int64_t)n * C);
    T accL = (T)0, accR = (T)0;
    if (ih % ds == 0) {
      const int ph = ih / ds;
      const T* gL = grad + (((int64_t)n * 2 * C + c) * D) * H * W + (int64_t)ph * W;
      const T* gR = gL + (int64_t)C * D * H * W;
      const int64_t plane = (int64_t)H * W;
      const bool on_grid = (x % ds) == 0;
      for (int d = 0; d < D; ++d) {
        if (on_grid) accL += gL[d * plane + x / ds];
        const T s = shift[(int64_t)n * D + d];
        const T ns = -s;
        // candidate output columns: iw - s in (x-1, x+1)  <=>  iw in (x-1+s, x+1+s)
        T lo = (T)x - (T)1 + s, hi = (T)x + (T)1 + s;
        int pw_lo = (int)floor((double)lo / ds) - 1;
        int pw_hi = (int)ceil((double)hi / ds) + 1;
        pw_lo = max(pw_lo, 0);
        pw_hi = min(pw_hi, W - 1);
        for (int pw = pw_lo; pw <= pw_hi; ++pw) {
          int xl, xh;
          T lx;
          if (!sample_pos<T>(pw * ds, ns, img_w, xl, xh, lx)) continue;
          if (xl != x && xh != x) continue;
          T hx = Arith<T>::sub((T)1, lx);
          // hy = 1, ly = 0 (integer y): w1 = hx, w2 = lx, w3 = w4 = 0
          T g = gR[d * plane + pw];
          if (xl == x && hx >= (T)1e-10) accR += g * hx;
          if (xh == x && lx >= (T)1e-10) accR += g * lx;
        }
      }
    }
    gl[i] = accL;
    gr[i] = accR;
  }
}

__global__ void cv_xlow_kernel(const float* __restrict__ shift, int32_t* __restrict__ xlow, int64_t total, int D,
                               int W, int ds) {
  const int img_w = W * ds;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int pw = (int)(i % W);
    int64_t nd = i / W;
    int xl, xh;
    float lx;
    xlow[i] = sample_pos<float>(pw * ds, -shift[nd], img_w, xl, xh, lx) ? xl : -1;
  }
}

int launch_cv_ndhwc(const void* left, const void* right, const void* shift, void* cost, void* left_planes, int64_t N,
                    int64_t C, int64_t IH, int64_t IW, int64_t D, int ds, int64_t H, int64_t W, cudaStream_t stream,
                    int parts = 3);

}  // namespace
}  // namespace snvc

using namespace snvc;

extern "C" int snvc_cost_volume_fwd(const void* left, const void* right, const void* shift, void* cost, int64_t N,
                                    int64_t C, int64_t IH, int64_t IW, int64_t D, int32_t ds, int32_t dtype,
                                    int32_t out_dtype, int32_t out_layout, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  SNVC_CHECK_ARG(ds >= 1, "downsample must be >= 1 (got %d)", ds);
  SNVC_CHECK_ARG(N >= 0 && C >= 0 && IH >= 0 && IW >= 0 && D >= 0, "negative dimension");
  SNVC_CHECK_ARG(IH % ds == 0 && IW % ds == 0, "IH (%lld) and IW (%lld) must be multiples of downsample (%d)",
                 (long long)IH, (long long)IW, ds);
  const int64_t H = IH / ds, W = IW / ds;
  if (N * C * D * H * W == 0) return 0;  // empty output returns early (.cu:235-238)
  SNVC_CHECK_ARG(left && right && shift && cost, "null pointer");
  SNVC_CHECK_ARG(IW < (1 << 24) && IH < (1 << 24) && D < (1 << 24) && N * C < (1ll << 31) && N < 65536,
                 "dimension too large");

  if (out_layout == SNVC_NCDHW) {
    const bool f32 = dtype == SNVC_F32 && out_dtype == SNVC_F32;
    const bool f64 = dtype == SNVC_F64 && out_dtype == SNVC_F64;
    if (!f32 && !f64) return fail(SNVC_E_UNSUPPORTED, "NCDHW cost volume: supported types are f32->f32 and f64->f64");
    const size_t es = f32 ? 4 : 8;
    int TH = 4;
    while (TH > 1 && (2 * TH * IW + D) * es > 96 * 1024) TH >>= 1;
    size_t smem = (2 * (size_t)TH * IW + D) * es;
    if (smem > 200 * 1024) return fail(SNVC_E_UNSUPPORTED, "image row too wide for the staged kernel (IW=%lld)", (long long)IW);
    const int64_t htiles = ceil_div(H, TH);
    SNVC_CHECK_ARG(N * C <= 65535 * 1ll, "N*C too large for one launch (%lld)", (long long)(N * C));
    // enough CTAs for >= ~8 waves: split D when the (h-tile, n, c) grid is small
    int64_t base = htiles * N * C;
    int dsplit = 1;
    const int64_t want = 8ll * 4 * sm_count();
    while (base * dsplit < want && dsplit * 2 <= D) dsplit *= 2;
    int d_per = (int)ceil_div(D, dsplit);
    dsplit = (int)ceil_div(D, d_per);
    dim3 grid((unsigned)htiles, (unsigned)(N * C), (unsigned)dsplit);
    if (f32) {
      const bool vec = (W % 4 == 0) && ((reinterpret_cast<uintptr_t>(cost) & 15) == 0);
      auto k = vec ? cv_ncdhw_kernel<float, true> : cv_ncdhw_kernel<float, false>;
      if (smem > 48 * 1024) SNVC_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      k<<<grid, 256, smem, stream>>>((const float*)left, (const float*)right, (const float*)shift, (float*)cost,
                                     (int)C, (int)IH, (int)IW, (int)D, (int)H, (int)W, ds, TH, d_per);
    } else {
      auto k = cv_ncdhw_kernel<double, false>;
      if (smem > 48 * 1024) SNVC_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      k<<<grid, 256, smem, stream>>>((const double*)left, (const double*)right, (const double*)shift, (double*)cost,
                                     (int)C, (int)IH, (int)IW, (int)D, (int)H, (int)W, ds, TH, d_per);
    }
    return launch_status("cv_ncdhw_kernel");
  }

  if (out_layout == SNVC_NDHWC) {
    if (!(dtype == SNVC_F32 && out_dtype == SNVC_BF16))
      return fail(SNVC_E_UNSUPPORTED, "NDHWC cost volume: supported types are f32 -> bf16");
    return launch_cv_ndhwc(left, right, shift, cost, nullptr, N, C, IH, IW, D, ds, H, W, stream);
  }
  return fail(SNVC_E_BADARG, "unknown out_layout %d", out_layout);
}

namespace snvc {
namespace {
// parts (split form only): bit 0 = right-half volume, bit 1 = left planes (two independent launches)
int launch_cv_ndhwc(const void* left, const void* right, const void* shift, void* cost, void* left_planes, int64_t N,
                    int64_t C, int64_t IH, int64_t IW, int64_t D, int ds, int64_t H, int64_t W, cudaStream_t stream,
                    int parts) {
  SNVC_CHECK_ARG(C % 8 == 0, "NDHWC cost volume needs C %% 8 == 0 (got %lld)", (long long)C);
  const bool split_form = left_planes != nullptr || parts != 3;
  if (split_form && H <= 2147483647ll && N <= 65535 && (!opt(OPT_CV_SPLIT_OLD) || parts != 3)) {
    // split form, whole rows staged: the right row [img_w][C] fp32 + the sample table of one depth split (the left planes
    // are cv_left_planes_kernel, a plain transposition)
    const size_t rows = (size_t)IW * C * 4;
    const size_t budget = 74 * 1024;                                    // three CTAs per SM
    if (rows + (size_t)W * 8 <= 225 * 1024) {
      int d_per = (int)D;
      if (rows + (size_t)W * D * 8 > budget) d_per = (int)std::max<int64_t>(1, ((int64_t)std::max<size_t>(budget, rows + W * 8) - (int64_t)rows) / (W * 8));
      // a row wider than the 3-CTA budget (full-resolution stress volume: 1248 x 32 x 4 B = 160 KB) runs one CTA per SM
      // anyway: give that CTA as many bins as the 225 KB allow, instead of staging 160 KB for ONE 80 KB output row
      if (rows + (size_t)W * 8 > budget)
        d_per = (int)std::max<int64_t>(1, std::min<int64_t>(D, (int64_t)(225 * 1024 - rows) / (W * 8)));
      // fill whole waves: more depth splits when the grid would leave SMs idle
      const double wave = (rows + (size_t)W * d_per * 8 <= budget ? 3.0 : 2.0) * sm_count();
      double best = -1;
      int pick = (int)ceil_div(D, d_per);
      for (int cand = (int)ceil_div(D, d_per); cand <= std::min<int64_t>(D, ceil_div(D, d_per) + 6); ++cand) {
        const double x = (double)H * N * ceil_div(D, ceil_div(D, cand)) / wave;
        const double eff = x / ceil(x) - 0.01 * cand;
        if (eff > best + 1e-9) { best = eff; pick = cand; }
      }
      d_per = (int)ceil_div(D, pick);
      const int dsplit = (int)ceil_div(D, d_per);
      if (dsplit <= 65535) {
        const size_t smem = rows + (size_t)W * d_per * 8 + (size_t)d_per * 4;   // row + sample table + regular-range starts
        if (smem > 48 * 1024)
          SNVC_CUDA_OK(cudaFuncSetAttribute(cv_split_bf16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int mask = 1;  // largest 2^k - 1 (k <= 3) with 2^k | C/4
        while (mask < 7 && ((C / 4) % (2 * (mask + 1))) == 0) mask = 2 * mask + 1;
        // (ncu: with the default carve-out only two of the 70 KB CTAs were resident per SM)
        SNVC_CUDA_OK(cudaFuncSetAttribute(cv_split_bf16_kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                                          (int)cudaSharedmemCarveoutMaxShared));
        if (parts & 2) {
          const int64_t items = N * H * W * (C / 8);
          const int lblocks = (int)std::max<int64_t>(1, std::min<int64_t>(ceil_div(items, 256), (int64_t)sm_count() * 8));
          cv_left_planes_kernel<<<lblocks, 256, 0, stream>>>((const float*)left, (__nv_bfloat16*)left_planes, (int)C, (int)IH,
                                                           (int)IW, (int)H, (int)W, ds, items);
          if (int e = launch_status("cv_left_planes_kernel")) return e;
        }
        if (!(parts & 1)) return 0;
        dim3 grid((unsigned)H, (unsigned)N, (unsigned)dsplit);
        // threads: a lane owns (bin, strip of columns, 8 channels); pick the block size (a multiple of 32, <= 512) whose
        // strips -- odd length, see the kernel -- cover the row with the least padding (KITTI: 12 bins x 4 groups x 8
        // strips of 39 columns = 384 threads, no padding; 512 threads = 10 strips of 33 = 5.8 % padded + an idle warp)
        int threads = 512;
        {
            const int per_strip = std::min(d_per, (int)D) * (int)(C / 8);
            double best_w = 1e30;
            for (int t = 512; t >= 256; t -= 32) {
              const int ns = std::max(1, t / per_strip);
              const int L = (int)(((W + ns - 1) / ns) | 1);
              const double waste = (double)ns * L / (double)W * ((double)t / (double)(ns * per_strip > t ? t : ns * per_strip));
              if (waste < best_w - 1e-9) { best_w = waste; threads = t; }
            }
            if (const char* o = opt(OPT_CV_THREADS)) threads = std::max(32, std::min(512, atoi(o) / 32 * 32));
        }
        cv_split_bf16_kernel<<<grid, threads, smem, stream>>>((const float*)right, (const float*)shift, (__nv_bfloat16*)cost,
                                                           (int)C, (int)IH, (int)IW, (int)D, (int)H, (int)W, ds, d_per, mask,
                                                           opt(OPT_CV_WALK) ? 1 : 0);
        return launch_status("cv_split_bf16_kernel");
      }
    }
  }
  if (parts != 3) return fail(SNVC_E_UNSUPPORTED, "split cost volume: a single part needs rows that fit shared memory");
  {
    SNVC_CHECK_ARG((reinterpret_cast<uintptr_t>(cost) & 15) == 0, "cost must be 16-byte aligned");
    SNVC_CHECK_ARG(H <= 2147483647ll && N <= 65535, "H/N too large");
    // w-tiling: whole row when it fits (KITTI 1/4-res: 2*312*32*4 = 78 KB -> 2 CTAs/SM)
    const size_t budget = 100 * 1024, hard = 200 * 1024;
    int TW = (int)W, S = 0;
    size_t row_bytes = (size_t)C * 4;
    // staged rows + the shifts (the per-(column, bin) sample table, 8 bytes per entry, is added once the depth split is known)
    auto need = [&](int tw, int s) { return ((size_t)((tw - 1) * ds + 1) + (size_t)((tw - 1) * ds + 2 + s)) * row_bytes + (size_t)(D + 2) * 4; };
    if (need(TW, 0) > budget) {
      int ntile = 2;
      while (need((int)ceil_div(W, ntile), 0) > budget / 2 && ntile < W) ++ntile;
      TW = (int)ceil_div(W, ntile);
      // spend what is left of the hard budget on the left extension of the right window
      size_t rest = hard - need(TW, 0);
      S = (int)(rest / row_bytes);
    }
    const int wtiles = (int)ceil_div(W, TW);
    size_t smem = need(TW, S);
    if (smem > 220 * 1024) return fail(SNVC_E_UNSUPPORTED, "cost volume tile does not fit shared memory");
    // split D so the grid fills whole waves (2 CTAs/SM at ~80 KB smem): few splits = little re-staging
    const int64_t base = H * N * wtiles;
    const double wave = 2.0 * sm_count();
    int dsplit = 1;
    double best_eff = -1;
    for (int cand = 1; cand <= 8 && cand <= D; ++cand) {
      const double x = base * (double)ceil_div(D, ceil_div(D, cand)) / wave;
      const double eff = x / ceil(x) - 0.01 * cand;
      if (eff > best_eff + 1e-9) { best_eff = eff; dsplit = cand; }
    }
    int d_per = (int)ceil_div(D, dsplit);
    // the sample table must fit next to the staged rows: split the depth further if it does not
    while (d_per > 1 && smem + (size_t)TW * d_per * 8 > 225 * 1024) d_per = (d_per + 1) / 2;
    dsplit = (int)ceil_div(D, d_per);
    smem += (size_t)TW * d_per * 8;
    SNVC_CHECK_ARG((int64_t)dsplit * wtiles <= 65535, "grid.z too large");
    if (smem > 225 * 1024) return fail(SNVC_E_UNSUPPORTED, "cost volume tile does not fit shared memory");
    if (smem > 48 * 1024)
      SNVC_CUDA_OK(cudaFuncSetAttribute(cv_ndhwc_bf16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int mask = 1;  // largest 2^k - 1 (k <= 3) with 2^k | C/4
    while (mask < 7 && ((C / 4) % (2 * (mask + 1))) == 0) mask = 2 * mask + 1;
    dim3 grid((unsigned)H, (unsigned)N, (unsigned)(dsplit * wtiles));
    cv_ndhwc_bf16_kernel<<<grid, 512, smem, stream>>>((const float*)left, (const float*)right, (const float*)shift,
                                                       (__nv_bfloat16*)cost, (__nv_bfloat16*)left_planes, (int)C, (int)IH,
                                                       (int)IW, (int)D, (int)H, (int)W, ds, d_per, dsplit, TW, S, mask);
    return launch_status("cv_ndhwc_bf16_kernel");
  }
}
}  // namespace
}  // namespace snvc

extern "C" int snvc_cost_volume_split_fwd(const void* left, const void* right, const void* shift, void* right_vol,
                                          void* left_planes, int64_t N, int64_t C, int64_t IH, int64_t IW, int64_t D,
                                          int32_t ds, void* stream_) {
  SNVC_CHECK_ARG(ds >= 1, "downsample must be >= 1");
  SNVC_CHECK_ARG(N >= 0 && C > 0 && IH >= 0 && IW >= 0 && D >= 0, "negative dimension");
  SNVC_CHECK_ARG(IH % ds == 0 && IW % ds == 0, "IH and IW must be multiples of downsample");
  const int64_t H = IH / ds, W = IW / ds;
  if (N * C * D * H * W == 0) return 0;
  SNVC_CHECK_ARG(left && right && shift && (right_vol || left_planes), "null pointer");
  SNVC_CHECK_ARG(((reinterpret_cast<uintptr_t>(left_planes) | reinterpret_cast<uintptr_t>(right_vol)) & 15) == 0,
                 "right_vol and left_planes must be 16-byte aligned");
  // either output may be NULL: the two halves are independent launches, so a caller can enqueue them on different
  // streams (the left planes feed the small addend convolution, which then overlaps the right-half build)
  const int parts = (right_vol ? 1 : 0) | (left_planes ? 2 : 0);
  return launch_cv_ndhwc(left, right, shift, right_vol, left_planes, N, C, IH, IW, D, ds, H, W, (cudaStream_t)stream_, parts);
}

extern "C" int snvc_cost_volume_bwd(const void* grad, const void* shift, void* grad_left, void* grad_right, int64_t N,
                                    int64_t C, int64_t H, int64_t W, int64_t D, int32_t ds, int32_t dtype,
                                    void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  SNVC_CHECK_ARG(ds >= 1, "downsample must be >= 1");
  SNVC_CHECK_ARG(dtype == SNVC_F32 || dtype == SNVC_F64, "dtype must be f32 or f64");
  const int64_t total = N * C * H * ds * W * ds;
  if (total == 0) return 0;
  SNVC_CHECK_ARG(grad_left && grad_right && shift, "null pointer");
  const size_t es = dtype == SNVC_F32 ? 4 : 8;
  if (D == 0 || grad == nullptr) {  // empty gradient: zeros (.cu:270-285)
    SNVC_CUDA_OK(cudaMemsetAsync(grad_left, 0, total * es, stream));
    SNVC_CUDA_OK(cudaMemsetAsync(grad_right, 0, total * es, stream));
    return 0;
  }
  int blocks = (int)std::min<int64_t>(ceil_div(total, 256), (int64_t)sm_count() * 16);
  if (dtype == SNVC_F32)
    cv_bwd_kernel<float><<<blocks, 256, 0, stream>>>((const float*)grad, (const float*)shift, (float*)grad_left,
                                                     (float*)grad_right, (int)C, (int)H, (int)W, (int)D, ds, total);
  else
    cv_bwd_kernel<double><<<blocks, 256, 0, stream>>>((const double*)grad, (const double*)shift, (double*)grad_left,
                                                      (double*)grad_right, (int)C, (int)H, (int)W, (int)D, ds, total);
  return launch_status("cv_bwd_kernel");
}

extern "C" int snvc_cost_volume_xlow(const float* shift, int32_t* xlow, int64_t N, int64_t IW, int64_t D, int32_t ds,
                                     void* stream_) {
  SNVC_CHECK_ARG(ds >= 1 && IW % ds == 0, "bad downsample");
  const int64_t W = IW / ds, total = N * D * W;
  if (total == 0) return 0;
  int blocks = (int)std::min<int64_t>(ceil_div(total, 256), (int64_t)sm_count() * 8);
  cv_xlow_kernel<<<blocks, 256, 0, (cudaStream_t)stream_>>>(shift, xlow, total, (int)D, (int)W, ds);
  return launch_status("cv_xlow_kernel");
}
