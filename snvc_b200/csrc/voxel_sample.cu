// A3 / A4 -- gather kernels: instance ROI voxel sampling and the global frustum-to-voxel lift.
//
// Both reproduce torch's zero-padded (bi|tri)linear grid_sample (ATen/native/GridSampler.h:27-36
// `grid_sampler_unnormalize`, GridSampler.cpp corner order and weight products) with explicit
// round-to-nearest fp32 intrinsics and no FMA contraction, so corner indices and in-bounds masks
// are bit-identical to oracle/grid_sample.py.  What changes vs the reference's call sequence
// (vernier.py:332-346: permute+reshape copy, 6 elementwise normalisation kernels, 2 grid_sample,
// cat) is the data movement: features are re-laid out channels-last once (tiny), every output
// element is written exactly once, in the layout its consumer reads (NDHWC bf16 for the tcgen05
// conv3d, or the reference's NCDHW fp32), with 128-bit accesses.
#include <algorithm>
#include <cmath>
#include <cstdlib>

#include "common.cuh"

namespace snvc {
namespace {

// ---- shared coordinate arithmetic (bit-exact contract; see oracle/grid_sample.py) ---------
__device__ __forceinline__ float unnormalize(float g, int size, bool align_corners) {
  // x / 2 == x * 0.5 exactly in binary fp (no denormals here): saves two IEEE divisions per axis
  if (align_corners) return __fmul_rn(__fmul_rn(__fadd_rn(g, 1.f), 0.5f), (float)(size - 1));
  return __fmul_rn(__fsub_rn(__fmul_rn(__fadd_rn(g, 1.f), (float)size), 1.f), 0.5f);
}
// vernier.py:335-338:  p / resolution * 2 - 1
__device__ __forceinline__ float roi_normalize(float p, float res) {
  return __fsub_rn(__fmul_rn(__fdiv_rn(p, res), 2.f), 1.f);
}

struct Bilinear {
  int x0, y0;          // floor corner (nw)
  float w[4];          // nw, ne, sw, se
  unsigned mask;       // bit k set <=> corner k inside the feature map
};

// qx, qy = p / resolution (the first of the three rounded operations of vernier.py:335-338)
__device__ __forceinline__ Bilinear bilinear_setup_norm(float qx, float qy, int Wf, int Hf) {
  Bilinear b;
  float ix = unnormalize(__fsub_rn(__fmul_rn(qx, 2.f), 1.f), Wf, false);
  float iy = unnormalize(__fsub_rn(__fmul_rn(qy, 2.f), 1.f), Hf, false);
  float fx0 = floorf(ix), fy0 = floorf(iy);
  // non-finite coordinates: no corner is in bounds (torch compares the float->int cast; here the
  // range test on the float itself rejects NaN/inf before the cast)
  bool finite = (fabsf(ix) < 1e9f) && (fabsf(iy) < 1e9f);
  b.x0 = finite ? (int)fx0 : -2;
  b.y0 = finite ? (int)fy0 : -2;
  float fx1 = __fadd_rn(fx0, 1.f), fy1 = __fadd_rn(fy0, 1.f);
  float dx1 = __fsub_rn(fx1, ix), dx0 = __fsub_rn(ix, fx0);
  float dy1 = __fsub_rn(fy1, iy), dy0 = __fsub_rn(iy, fy0);
  b.w[0] = __fmul_rn(dx1, dy1);
  b.w[1] = __fmul_rn(dx0, dy1);
  b.w[2] = __fmul_rn(dx1, dy0);
  b.w[3] = __fmul_rn(dx0, dy0);
  bool xin0 = b.x0 >= 0 && b.x0 < Wf, xin1 = b.x0 + 1 >= 0 && b.x0 + 1 < Wf;
  bool yin0 = b.y0 >= 0 && b.y0 < Hf, yin1 = b.y0 + 1 >= 0 && b.y0 + 1 < Hf;
  b.mask = (xin0 && yin0 ? 1u : 0u) | (xin1 && yin0 ? 2u : 0u) | (xin0 && yin1 ? 4u : 0u) | (xin1 && yin1 ? 8u : 0u);
  return b;
}
__device__ __forceinline__ Bilinear bilinear_setup(float px, float py, float res_x, float res_y, int Wf, int Hf) {
  return bilinear_setup_norm(__fdiv_rn(px, res_x), __fdiv_rn(py, res_y), Wf, Hf);
}

// out = (((0 + v_nw*w_nw) + v_ne*w_ne) + v_sw*w_sw) + v_se*w_se, out-of-bounds corners skipped
__device__ __forceinline__ float acc4(const Bilinear& b, float v0, float v1, float v2, float v3) {
  float o = 0.f;
  if (b.mask & 1u) o = __fadd_rn(o, __fmul_rn(v0, b.w[0]));
  if (b.mask & 2u) o = __fadd_rn(o, __fmul_rn(v1, b.w[1]));
  if (b.mask & 4u) o = __fadd_rn(o, __fmul_rn(v2, b.w[2]));
  if (b.mask & 8u) o = __fadd_rn(o, __fmul_rn(v3, b.w[3]));
  return o;
}

// NCHW fp32 -> NHWC fp32 (both views at once); tiny: N*C*Hf*Wf elements
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ oa,
                                    float* __restrict__ ob, int C, int HW, int64_t total) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    // i indexes the NHWC output; read is strided but the whole map is L2 resident
    int c = (int)(i % C);
    int64_t t = i / C;
    int p = (int)(t % HW);
    int64_t n = t / HW;
    int64_t src = (n * C + c) * HW + p;
    oa[i] = a[src];
    ob[i] = b[src];
  }
}

// NCHW fp32 -> NHWC bf16 (both views at once): the feature copy of the bf16 product path (v4 kernel below).  A thread
// writes 8 channels of one pixel (16 bytes); reads are strided but the maps are L2 resident.
__global__ void nchw_to_nhwc_bf16_kernel(const float* __restrict__ a, const float* __restrict__ b, uint4* __restrict__ oa,
                                         uint4* __restrict__ ob, int C, int HW, int64_t total /* N*HW*C/8 */) {
  const int CG = C >> 3;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int cg = (int)(i % CG);
    const int64_t t = i / CG;
    const int p = (int)(t % HW);
    const int64_t n = t / HW;
    const int64_t src = (n * C + cg * 8) * HW + p;
    float va[8], vb[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { va[j] = a[src + (int64_t)j * HW]; vb[j] = b[src + (int64_t)j * HW]; }
    oa[i] = make_uint4(pack_bf16x2(va[0], va[1]), pack_bf16x2(va[2], va[3]), pack_bf16x2(va[4], va[5]), pack_bf16x2(va[6], va[7]));
    ob[i] = make_uint4(pack_bf16x2(vb[0], vb[1]), pack_bf16x2(vb[2], vb[3]), pack_bf16x2(vb[4], vb[5]), pack_bf16x2(vb[6], vb[7]));
  }
}

// ---- A3, NDHWC output: 8 channels (16/32 B) per thread, lanes = (point, view, channel group) ----
template <typename OutT>
__global__ void __launch_bounds__(256)
roi_sample_ndhwc_kernel(const float* __restrict__ fl, const float* __restrict__ fr, const float* __restrict__ pl,
                        const float* __restrict__ pr, OutT* __restrict__ out, int C, int Hf, int Wf, int64_t P,
                        float res_x, float res_y, int64_t total /* N*P*(2C/8) */) {
  const int CG = C >> 3;          // 8-channel groups per view
  const int G = 2 * CG;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int g = (int)(i % G);
    int64_t np = i / G;
    int64_t n = np / P;
    int64_t p = np - n * P;
    const bool right = g >= CG;
    const int cg = right ? g - CG : g;
    const float* pts = (right ? pr : pl) + n * 2 * P;
    const float* feat = (right ? fr : fl) + n * (int64_t)Hf * Wf * C + cg * 8;
    Bilinear b = bilinear_setup(__ldg(pts + p), __ldg(pts + P + p), res_x, res_y, Wf, Hf);
    float v[4][8];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (b.mask & (1u << k)) {
        const float* src = feat + ((int64_t)(b.y0 + (k >> 1)) * Wf + (b.x0 + (k & 1))) * C;
        *reinterpret_cast<float4*>(v[k]) = __ldg(reinterpret_cast<const float4*>(src));
        *reinterpret_cast<float4*>(v[k] + 4) = __ldg(reinterpret_cast<const float4*>(src + 4));
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[k][j] = 0.f;
      }
    }
    float r[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = acc4(b, v[0][j], v[1][j], v[2][j], v[3][j]);
    OutT* o = out + np * (2 * C) + g * 8;
    if (sizeof(OutT) == 2) {
      uint4 q = {pack_bf16x2(r[0], r[1]), pack_bf16x2(r[2], r[3]), pack_bf16x2(r[4], r[5]), pack_bf16x2(r[6], r[7])};
      *reinterpret_cast<uint4*>(o) = q;
    } else {
      *reinterpret_cast<float4*>(o) = make_float4(r[0], r[1], r[2], r[3]);
      *reinterpret_cast<float4*>(reinterpret_cast<float*>(o) + 4) = make_float4(r[4], r[5], r[6], r[7]);
    }
  }
}

// Division of a 32-bit index by a run-time constant (Granlund-Montgomery): the flat voxel index is split into
// (n, z, y, x) with three of these instead of four 64-bit software divisions (ncu: those were 250 of the
// 393 set-up instructions per voxel, and the kernel was issue-bound).
struct FastDiv { uint32_t m, s1, s2, d; };
__device__ __forceinline__ uint32_t fdiv(uint32_t n, const FastDiv& f) {
  const uint32_t t = __umulhi(f.m, n);
  return (t + ((n - t) >> f.s1)) >> f.s2;
}

// ---- A3, NDHWC output, v2: cooperative gather.  The kernel above recomputes the bilinear set-up (two IEEE
// divisions per axis pair) in every one of the C/8 threads of a (point, view); ncu / bench_instance: 1.5 TB/s,
// issue-bound.  Here a warp owns 32 consecutive points: lane i does the set-up of point i for BOTH views once,
// then C/8 adjacent lanes gather each (point, view) -- 8 channels (two 16-byte loads per corner) per lane -- with
// the set-up handed over by warp shuffles, and write 16 bytes of the bf16 row each.  Arithmetic per channel is
// the separately rounded mul / add chain of acc4(), two channels per instruction (see mul2_exact / add2_exact),
// so the output stays bit-identical to the oracle.
// Exact packed product / sum built from the packed FMA only: RN(a*b + 0) == RN(a*b) and RN(b*1 + a) == RN(a + b).
// (mul.rn.f32x2 followed by add.rn.f32x2 is NOT usable: ptxas 12.9 contracts the pair into one FFMA2 even with
// explicit .rn and -fmad=false -- seen in the SASS -- which changes the rounding.)
__device__ __forceinline__ float2 mul2_exact(float2 a, float2 b) { return __ffma2_rn(a, b, make_float2(0.f, 0.f)); }
__device__ __forceinline__ float2 add2_exact(float2 a, float2 b) { return __ffma2_rn(b, make_float2(1.f, 1.f), a); }

template <typename OutT>
__global__ void __launch_bounds__(256, 4)
roi_sample_coop_kernel(const float* __restrict__ fl, const float* __restrict__ fr, const float* __restrict__ pl,
                       const float* __restrict__ pr, OutT* __restrict__ out, int C, int log_lpv, int Hf, int Wf,
                       FastDiv divP, float res_x, float res_y, uint32_t total /* N*P < 2^31 */) {
  const int lane = threadIdx.x & 31;
  const int lpv = 1 << log_lpv;                 // lanes per (point, view) = C / 8
  const int ppr = 32 >> log_lpv;                // points per round
  const int cg = lane & (lpv - 1);
  const int psel = lane >> log_lpv;
  const uint32_t P = divP.d;
  const uint32_t warp0 = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t nwarps = (gridDim.x * blockDim.x) >> 5;
  for (uint32_t chunk = warp0; chunk * 32u < total; chunk += nwarps) {
    // ---- phase A: lane i sets up point chunk*32 + i for both views
    const uint32_t np = chunk * 32u + lane;
    int base[2] = {0, 0};
    unsigned msk[2] = {0u, 0u};
    float w[2][4];
#pragma unroll
    for (int v = 0; v < 2; ++v)
#pragma unroll
      for (int k = 0; k < 4; ++k) w[v][k] = 0.f;
    if (np < total) {
      const uint32_t n = fdiv(np, divP), pi = np - n * P;
#pragma unroll
      for (int v = 0; v < 2; ++v) {
        const float* pts = (v ? pr : pl) + (size_t)n * 2 * P;
        const Bilinear b = bilinear_setup(__ldg(pts + pi), __ldg(pts + P + pi), res_x, res_y, Wf, Hf);
        msk[v] = b.mask;
        base[v] = b.mask ? (int)n * Hf * Wf + b.y0 * Wf + b.x0 : 0;
        if (b.mask) {                              // (all corners outside: weights may be non-finite, keep zeros)
#pragma unroll
          for (int k = 0; k < 4; ++k) w[v][k] = b.w[k];
        }
      }
    }
    // ---- phase B: per view, lpv rounds of ppr points; lane cg owns channels [8cg, 8cg+8) of the view.
    // (Measured alternatives, bench_instance.py on B200: float4 #cg and #(cg+lpv) per lane 0.055 ms / proposal,
    // 4 channels per lane with C/4 lanes per point 0.059 ms; this form 0.050 ms.)
#pragma unroll
    for (int v = 0; v < 2; ++v) {
      const float* feat = v ? fr : fl;
      for (int r = 0; r < lpv; ++r) {
        const int src = r * ppr + psel;
        const unsigned mm = __shfl_sync(0xffffffffu, msk[v], src);
        const int b = __shfl_sync(0xffffffffu, base[v], src);
        float wk[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) wk[k] = __shfl_sync(0xffffffffu, w[v][k], src);
        const uint32_t onp = chunk * 32u + src;
        float2 acc2[4];
        float4 q[4][2];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          q[k][0] = q[k][1] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (mm & (1u << k)) {
            const float4* sp = reinterpret_cast<const float4*>(feat + (size_t)(b + (k >> 1) * Wf + (k & 1)) * C) + cg * 2;
            q[k][0] = __ldg(sp);
            q[k][1] = __ldg(sp + 1);
          }
        }
        // o = (((0 + v_nw*w_nw) + v_ne*w_ne) + v_sw*w_sw) + v_se*w_se; a masked corner contributes +0 (x + 0 == x)
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float2 wv = make_float2(wk[k], wk[k]);
          const float2 p0 = mul2_exact(make_float2(q[k][0].x, q[k][0].y), wv), p1 = mul2_exact(make_float2(q[k][0].z, q[k][0].w), wv);
          const float2 p2 = mul2_exact(make_float2(q[k][1].x, q[k][1].y), wv), p3 = mul2_exact(make_float2(q[k][1].z, q[k][1].w), wv);
          if (k == 0) {          // 0 + p == p (the product already carries the "+ 0")
            acc2[0] = p0; acc2[1] = p1; acc2[2] = p2; acc2[3] = p3;
          } else {
            acc2[0] = add2_exact(acc2[0], p0); acc2[1] = add2_exact(acc2[1], p1);
            acc2[2] = add2_exact(acc2[2], p2); acc2[3] = add2_exact(acc2[3], p3);
          }
        }
        if (onp < total) {
          OutT* o = out + (size_t)onp * (2 * C) + v * C + cg * 8;
          if (sizeof(OutT) == 2) {
            *reinterpret_cast<uint4*>(o) = make_uint4(pack_bf16x2(acc2[0].x, acc2[0].y), pack_bf16x2(acc2[1].x, acc2[1].y),
                                                      pack_bf16x2(acc2[2].x, acc2[2].y), pack_bf16x2(acc2[3].x, acc2[3].y));
          } else {
            *reinterpret_cast<float4*>(o) = make_float4(acc2[0].x, acc2[0].y, acc2[1].x, acc2[1].y);
            *reinterpret_cast<float4*>(reinterpret_cast<float*>(o) + 4) = make_float4(acc2[2].x, acc2[2].y, acc2[3].x, acc2[3].y);
          }
        }
      }
    }
  }
}

// ---- A3, NDHWC output, v3.  ncu / arithmetic on v2: 17.6 cycles per point per SM = its L1 wavefronts.  Every corner
// fetch of v2 is two LDG.128 whose four lanes per (point, view) read ALTERNATE 16-byte pieces of the 128-byte feature
// row, so each request touches 8 rows and uses half of every 128-byte wavefront (16 wavefronts per point instead of
// 8), and the two views of a point are stored by different requests (2 half-row wavefronts per point).  Here
//   * a corner is ONE 256-bit load per lane (sm_100 LDG.E.256): the lanes of a (point, view) cover its feature row
//     contiguously -> one full wavefront per corner row;
//   * the 2*C/8 lanes of a point gather BOTH views in the same round, so one store request writes whole [L|R] rows;
//   * the set-up is one (point, view) per lane (16 points per warp pass), handed over with the same 6 shuffles.
// Arithmetic per channel is unchanged (mul2_exact / add2_exact chain) -> bit-identical to v2 and to the oracle.
__device__ __forceinline__ void ldg256_f4(const float* p, float4& a, float4& b) {
  asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w)
               : "l"(p));
}
__device__ __forceinline__ void stg256_f4(float* p, float4 a, float4 b) {
  asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(a.x), "f"(a.y), "f"(a.z), "f"(a.w),
               "f"(b.x), "f"(b.y), "f"(b.z), "f"(b.w)
               : "memory");
}

template <typename OutT>
__global__ void __launch_bounds__(256, 3)
roi_sample_coop2_kernel(const float* __restrict__ fl, const float* __restrict__ fr, const float* __restrict__ pl,
                        const float* __restrict__ pr, OutT* __restrict__ out, int C, int log_lpv, int Hf, int Wf,
                        FastDiv divP, float res_x, float res_y, uint32_t total /* N*P < 2^30 */) {
  const int lane = threadIdx.x & 31;
  const int lpv = 1 << log_lpv;                 // lanes per (point, view) = C / 8
  const int log_lpp = log_lpv + 1;              // lanes per point = 2 * lpv
  const int ppr = 32 >> log_lpp;                // points per round
  const int cg = lane & (lpv - 1);
  const int view = (lane >> log_lpv) & 1;
  const int psel = lane >> log_lpp;
  const float* feat = view ? fr : fl;
  const uint32_t P = divP.d;
  // Grid-stride over 16-point chunks: at any moment the whole grid writes one contiguous window of the output.
  // [Block-contiguous ranges raised the L1 hit rate but ran slower, 0.046 vs 0.043 ms / proposal: 444 separate
  // write streams instead of one.]
  const uint32_t nchunks = (total + 15u) >> 4;
  const uint32_t c_begin = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, c_end = nchunks;
  const uint32_t wstep = (gridDim.x * blockDim.x) >> 5;
  // the coordinates of the NEXT chunk are loaded one iteration ahead (their latency was 20 % of the samples)
  auto load_pts = [&](uint32_t chunk, float& u, float& v, uint32_t& n_out) {
    const uint32_t np = chunk * 16u + (uint32_t)(lane >> 1);
    u = v = 0.f; n_out = 0u;
    if (chunk < c_end && np < total) {
      const uint32_t n = fdiv(np, divP), pi = np - n * P;
      const float* pts = ((lane & 1) ? pr : pl) + (size_t)n * 2 * P;
      u = __ldg(pts + pi); v = __ldg(pts + P + pi); n_out = n;
    }
  };
  float nu, nv;
  uint32_t nn;
  load_pts(c_begin, nu, nv, nn);
  for (uint32_t chunk = c_begin; chunk < c_end; chunk += wstep) {
    // ---- phase A: lane i sets up (point chunk*16 + i/2, view i&1)
    const uint32_t np = chunk * 16u + (uint32_t)(lane >> 1);
    const float pu = nu, pv = nv;
    const uint32_t n = nn;
    load_pts(chunk + wstep, nu, nv, nn);
    int base = 0;
    unsigned msk = 0u;
    float w[4] = {0.f, 0.f, 0.f, 0.f};
    if (np < total) {
      const Bilinear b = bilinear_setup(pu, pv, res_x, res_y, Wf, Hf);
      msk = b.mask;
      if (b.mask) {                                // (all corners outside: weights may be non-finite, keep zeros)
        base = (int)n * Hf * Wf + b.y0 * Wf + b.x0;
#pragma unroll
        for (int k = 0; k < 4; ++k) w[k] = b.w[k];
      }
    }
    // ---- phase B: lpv rounds of ppr points, both views of a point in the same round.  A lane group walks
    // CONSECUTIVE points (pt = psel*lpv + r): voxel spacing is a fraction of a feature pixel, so successive points
    // often hit the same 2x2 corner block; its rows stay in registers and are fetched only when the block changes
    // (invariant: q[k] == the feature row if corner k is in bounds, else 0).  [A version that also shifted rows
    // between neighbouring blocks ran 0.070 ms / proposal against 0.045: five divergent paths per round.]
    float4 q[4][2];
    int pb = -(1 << 30);                            // never a neighbour of a real base
    unsigned pm = 0xffffffffu;
    auto fetch = [&](int k, int b, unsigned mm) {
      q[k][0] = q[k][1] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (mm & (1u << k)) ldg256_f4(feat + (size_t)(b + (k >> 1) * Wf + (k & 1)) * C + cg * 8, q[k][0], q[k][1]);
    };
    for (int r = 0; r < lpv; ++r) {
      const int pt = psel * lpv + r;
      const int src = pt * 2 + view;
      const unsigned mm = __shfl_sync(0xffffffffu, msk, src);
      const int b = __shfl_sync(0xffffffffu, base, src);
      float wk[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) wk[k] = __shfl_sync(0xffffffffu, w[k], src);
      const uint32_t onp = chunk * 16u + (uint32_t)pt;
      // same 2x2 corner block as the previous point of this lane group: nothing to fetch (predicated, branch-free)
      const bool same = b == pb && mm == pm;
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (!same) fetch(k, b, mm);
      pb = b; pm = mm;
      // o = (((0 + v_nw*w_nw) + v_ne*w_ne) + v_sw*w_sw) + v_se*w_se; a masked corner contributes +0 (x + 0 == x)
      float2 acc2[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 wv = make_float2(wk[k], wk[k]);
        const float2 p0 = mul2_exact(make_float2(q[k][0].x, q[k][0].y), wv), p1 = mul2_exact(make_float2(q[k][0].z, q[k][0].w), wv);
        const float2 p2 = mul2_exact(make_float2(q[k][1].x, q[k][1].y), wv), p3 = mul2_exact(make_float2(q[k][1].z, q[k][1].w), wv);
        if (k == 0) {          // 0 + p == p (the product already carries the "+ 0")
          acc2[0] = p0; acc2[1] = p1; acc2[2] = p2; acc2[3] = p3;
        } else {
          acc2[0] = add2_exact(acc2[0], p0); acc2[1] = add2_exact(acc2[1], p1);
          acc2[2] = add2_exact(acc2[2], p2); acc2[3] = add2_exact(acc2[3], p3);
        }
      }
      if (onp < total) {
        OutT* o = out + (size_t)onp * (2 * C) + view * C + cg * 8;
        if (sizeof(OutT) == 2) {
          *reinterpret_cast<uint4*>(o) = make_uint4(pack_bf16x2(acc2[0].x, acc2[0].y), pack_bf16x2(acc2[1].x, acc2[1].y),
                                                    pack_bf16x2(acc2[2].x, acc2[2].y), pack_bf16x2(acc2[3].x, acc2[3].y));
        } else {
          stg256_f4(reinterpret_cast<float*>(o), make_float4(acc2[0].x, acc2[0].y, acc2[1].x, acc2[1].y),
                    make_float4(acc2[2].x, acc2[2].y, acc2[3].x, acc2[3].y));
        }
      }
    }
  }
}

// ---- A3, NDHWC bf16 output, v4 (the product path of the instance branch).  ncu on v3 (profiles/r02_hbm_kernels_summary.txt):
// nothing saturated -- DRAM 30 %, L1 63 %, issue 65 % -- but 38.8 warp instructions per point, of which only 18 % are the
// interpolation arithmetic (FFMA2 / FADD2): the rest is 64-bit address arithmetic, zeroing of masked corner registers
// (CS2R 10 %), constant-bank reloads, predicates and the branches of the corner-block reuse, and the six shuffles per
// round cost as many L1 data-pipe wavefronts as all the feature loads.  The bf16 output is compared at the 1e-2 bar (the
// fp32 outputs stay on the exact kernels above; corner indices and masks are the same bit-exact set-up), which allows:
//   * one fused FFMA2 per (corner, channel pair) instead of the separately rounded FMUL2 + FADD2 (<= 1 fp32 ulp before the
//     bf16 rounding);
//   * masked corners as ZERO WEIGHTS on a clamped, always valid address: no per-corner predicate, no register zeroing
//     (the accumulator starts from +0, so fully masked points give +0 like the oracle);
//   * the per-(point, view) set-up handed over through shared memory (one STS.128 + one STS.32 per lane and chunk, one
//     LDS.128 + one LDS.32 per round, both broadcast reads) instead of six shuffles per round;
//   * the lane groups of a round take CONSECUTIVE points (voxel spacing is a fraction of a feature pixel, so the 2*PPR
//     corner rows of a request fall into one or two 128-byte lines and the L1 serves them in as many wavefronts): the reuse
//     that v3 got from registers and branches comes from request coalescing, branch-free.
// base + 4 * off in ONE instruction (IMAD.WIDE.U32); plain pointer arithmetic compiles to a 4-instruction carry chain
__device__ __forceinline__ const float* ptr_plus_u32x4(const float* base, uint32_t off) {
  uint64_t r;
  asm("mad.wide.u32 %0, %1, 4, %2;" : "=l"(r) : "r"(off), "l"(reinterpret_cast<uint64_t>(base)));
  return reinterpret_cast<const float*>(r);
}

template <int LOG_LPV, bool FEAT16>   // lanes per (point, view): C / 8 == 1 << LOG_LPV; features NHWC bf16 | fp32
__global__ void __launch_bounds__(256, 4)
roi_sample_fast_bf16_kernel(const void* __restrict__ fl_, const void* __restrict__ fr_, const float* __restrict__ pl,
                            const float* __restrict__ pr, __nv_bfloat16* __restrict__ out, int Hf, int Wf, FastDiv divP,
                            float res_x, float res_y, float inv_res_x, float inv_res_y,
                            uint32_t total /* N*P < 2^30, N*Hf*Wf*C < 2^31 */) {
  constexpr int LPV = 1 << LOG_LPV, PPR = 16 >> LOG_LPV, ROUNDS = 16 / PPR, C = 8 * LPV;
  __shared__ float4 s_w[8][2][32];      // corner weights (0 for a corner outside the map)
  __shared__ uint4 s_o[8][2][32];       // corner offsets in floats from the view's feature base (clamped into the map)
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  // Lane roles.  Set-up: lane i owns (point i & 15, view i >> 4) of the chunk.  Gather: the half-warp is the VIEW, inside it
  // PPR consecutive points x LPV channel groups -- a quarter-warp (the unit the L1 serves per wavefront) then reads the
  // rows of 8 / LPV consecutive points of one view, i.e. one or two 128-byte lines.  [With the two views of a point in the
  // same quarter every 16-byte load cost 8 wavefronts (2 lines x 4 quarters) and the kernel sat at 97 % of the L1 data pipe.]
  const int cg = lane & (LPV - 1);
  const int view = lane >> 4;
  const int psel = (lane >> LOG_LPV) & (PPR - 1);
  // lane's channel group inside a feature row: 8 channels = 16 B (bf16 copy of the features) or 32 B (fp32)
  const float* const feat = reinterpret_cast<const float*>(view ? fr_ : fl_) + cg * (FEAT16 ? 4 : 8);
  const float* const my_pts = view ? pr : pl;
  const uint32_t P = divP.d;
  const uint32_t nchunks = (total + 15u) >> 4;
  const uint32_t c_begin = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t wstep = (gridDim.x * blockDim.x) >> 5;
  auto load_pts = [&](uint32_t chunk, float& u, float& v, uint32_t& n_out) {
    const uint32_t np = chunk * 16u + (uint32_t)(lane & 15);
    u = v = 0.f; n_out = 0u;
    if (np < total) {                                      // (chunk >= nchunks implies np >= total)
      const uint32_t n = fdiv(np, divP), i = np + n * P;   // n * 2P + (np - n * P)
      u = __ldg(ptr_plus_u32x4(my_pts, i)); v = __ldg(ptr_plus_u32x4(my_pts, i + P)); n_out = n;
    }
  };
  float nu, nv;
  uint32_t nn;
  load_pts(c_begin, nu, nv, nn);
  int buf = 0;
  // this lane's output position inside a chunk's 16 rows of 2C bf16: point psel (+ r * PPR), view, channel group
  __nv_bfloat16* const out_lane = out + (size_t)psel * (2 * C) + view * C + cg * 8;
  const float4* const sw = &s_w[wib][0][view * 16 + psel];
  const uint4* const so = &s_o[wib][0][view * 16 + psel];
  for (uint32_t chunk = c_begin; chunk < nchunks; chunk += wstep, buf ^= 1) {
    // ---- phase A: lane i sets up (point chunk*16 + (i & 15), view i >> 4): the arithmetic of bilinear_setup() (bit-exact corner
    // indices and weights), with the in-bounds tests folded into per-axis weights (w = 0 for a corner outside the map)
    // and corner addresses clamped into the map.  Lanes past the end compute a harmless dummy.
    const float pu = nu, pv = nv;
    const uint32_t n = nn;
    load_pts(chunk + wstep, nu, nv, nn);
    // p / res == p * (1 / res) bit for bit when res is a power of two (the host passes inv_res != 0 only then)
    const float qx = inv_res_x != 0.f ? __fmul_rn(pu, inv_res_x) : __fdiv_rn(pu, res_x);
    const float qy = inv_res_x != 0.f ? __fmul_rn(pv, inv_res_y) : __fdiv_rn(pv, res_y);
    const float ix = unnormalize(__fsub_rn(__fmul_rn(qx, 2.f), 1.f), Wf, false);
    const float iy = unnormalize(__fsub_rn(__fmul_rn(qy, 2.f), 1.f), Hf, false);
    const float fx0 = floorf(ix), fy0 = floorf(iy);
    const bool finite = (fabsf(ix) < 1e9f) && (fabsf(iy) < 1e9f);
    const int x0 = finite ? (int)fx0 : -2, y0 = finite ? (int)fy0 : -2;
    const float dx1 = (unsigned)x0 < (unsigned)Wf ? __fsub_rn(__fadd_rn(fx0, 1.f), ix) : 0.f;
    const float dx0 = (unsigned)(x0 + 1) < (unsigned)Wf ? __fsub_rn(ix, fx0) : 0.f;
    const float dy1 = (unsigned)y0 < (unsigned)Hf ? __fsub_rn(__fadd_rn(fy0, 1.f), iy) : 0.f;
    const float dy0 = (unsigned)(y0 + 1) < (unsigned)Hf ? __fsub_rn(iy, fy0) : 0.f;
    const float4 w = make_float4(__fmul_rn(dx1, dy1), __fmul_rn(dx0, dy1), __fmul_rn(dx1, dy0), __fmul_rn(dx0, dy0));
    constexpr uint32_t RW = FEAT16 ? C / 2 : C;            // 32-bit words per feature row
    const uint32_t x0c = (uint32_t)min(max(x0, 0), Wf - 1) * RW, x1c = (uint32_t)min(max(x0 + 1, 0), Wf - 1) * RW;
    const uint32_t r0 = ((n * Hf + (uint32_t)min(max(y0, 0), Hf - 1)) * Wf) * RW;
    const uint32_t r1 = ((n * Hf + (uint32_t)min(max(y0 + 1, 0), Hf - 1)) * Wf) * RW;
    const uint4 off = make_uint4(r0 + x0c, r0 + x1c, r1 + x0c, r1 + x1c);
    s_w[wib][buf][lane] = w;
    s_o[wib][buf][lane] = off;
    __syncwarp();
    // ---- phase B: ROUNDS rounds of PPR consecutive points, both views of a point in the same round
    __nv_bfloat16* const o = out_lane + (size_t)chunk * 16u * (2 * C);
#pragma unroll
    for (int r = 0; r < ROUNDS; ++r) {
      const float4 wk = sw[buf * 32 + r * PPR];
      const uint4 ok = so[buf * 32 + r * PPR];
      float2 acc[4] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
      const float wv[4] = {wk.x, wk.y, wk.z, wk.w};
      const uint32_t ov[4] = {ok.x, ok.y, ok.z, ok.w};
      if (FEAT16) {
        uint4 q[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) q[k] = __ldg(reinterpret_cast<const uint4*>(ptr_plus_u32x4(feat, ov[k])));
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float2 w2 = make_float2(wv[k], wv[k]);
          const uint32_t u[4] = {q[k].x, q[k].y, q[k].z, q[k].w};
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[j] = __ffma2_rn(make_float2(bf16_lo(u[j]), bf16_hi(u[j])), w2, acc[j]);
        }
      } else {
        float4 q[4][2];
#pragma unroll
        for (int k = 0; k < 4; ++k) ldg256_f4(ptr_plus_u32x4(feat, ov[k]), q[k][0], q[k][1]);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float2 w2 = make_float2(wv[k], wv[k]);
          acc[0] = __ffma2_rn(make_float2(q[k][0].x, q[k][0].y), w2, acc[0]);
          acc[1] = __ffma2_rn(make_float2(q[k][0].z, q[k][0].w), w2, acc[1]);
          acc[2] = __ffma2_rn(make_float2(q[k][1].x, q[k][1].y), w2, acc[2]);
          acc[3] = __ffma2_rn(make_float2(q[k][1].z, q[k][1].w), w2, acc[3]);
        }
      }
      if (chunk * 16u + (uint32_t)(r * PPR + psel) < total)
        *reinterpret_cast<uint4*>(o + (size_t)r * PPR * (2 * C)) =
            make_uint4(pack_bf16x2(acc[0].x, acc[0].y), pack_bf16x2(acc[1].x, acc[1].y), pack_bf16x2(acc[2].x, acc[2].y),
                       pack_bf16x2(acc[3].x, acc[3].y));
    }
  }
}

// ---- A3, NCDHW fp32 output (the reference's layout): one thread per (point, view); channel loop;
//      stores are coalesced across the warp's consecutive points. -------------------------------
__global__ void __launch_bounds__(256)
roi_sample_ncdhw_kernel(const float* __restrict__ fl, const float* __restrict__ fr, const float* __restrict__ pl,
                        const float* __restrict__ pr, float* __restrict__ out, int C, int Hf, int Wf, int64_t P,
                        float res_x, float res_y) {
  const int64_t n = blockIdx.z;
  const bool right = blockIdx.y == 1;
  const float* pts = (right ? pr : pl) + n * 2 * P;
  const float* feat = (right ? fr : fl) + n * (int64_t)Hf * Wf * C;
  float* o = out + (n * 2 * C + (right ? C : 0)) * P;
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < P; p += (int64_t)gridDim.x * blockDim.x) {
    Bilinear b = bilinear_setup(__ldg(pts + p), __ldg(pts + P + p), res_x, res_y, Wf, Hf);
    const float* src[4];
#pragma unroll
    for (int k = 0; k < 4; ++k)
      src[k] = feat + ((int64_t)(b.y0 + (k >> 1)) * Wf + (b.x0 + (k & 1))) * C;
    for (int c = 0; c < C; c += 4) {
      float4 v[4];
#pragma unroll
      for (int k = 0; k < 4; ++k)
        v[k] = (b.mask & (1u << k)) ? __ldg(reinterpret_cast<const float4*>(src[k] + c)) : make_float4(0, 0, 0, 0);
      __stcs(o + (int64_t)(c + 0) * P + p, acc4(b, v[0].x, v[1].x, v[2].x, v[3].x));
      __stcs(o + (int64_t)(c + 1) * P + p, acc4(b, v[0].y, v[1].y, v[2].y, v[3].y));
      __stcs(o + (int64_t)(c + 2) * P + p, acc4(b, v[0].z, v[1].z, v[2].z, v[3].z));
      __stcs(o + (int64_t)(c + 3) * P + p, acc4(b, v[0].w, v[1].w, v[2].w, v[3].w));
    }
  }
}

__global__ void roi_indices_kernel(const float* __restrict__ pts, int32_t* __restrict__ idx, uint8_t* __restrict__ mask,
                                   int64_t N, int64_t P, int Hf, int Wf, float res_x, float res_y) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N * P; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t n = i / P, p = i - n * P;
    Bilinear b = bilinear_setup(pts[n * 2 * P + p], pts[n * 2 * P + P + p], res_x, res_y, Wf, Hf);
    idx[2 * i] = b.x0;
    idx[2 * i + 1] = b.y0;
    mask[i] = (uint8_t)b.mask;
  }
}

// ---- A4: frustum-to-voxel lift --------------------------------------------------------------
struct LiftGeom {
  int range_ordered;   // all three CV ranges are given low-to-high (enables the conservative early-out)
  float cv[6];   // CV_X_MIN, CV_X_MAX, CV_Y_MIN, CV_Y_MAX, CV_Z_MIN, CV_Z_MAX
  int D, H, W;   // extent of the volume tensor passed in (z-bins, rows, cols)
  int Dt, d_base;// depth-slab mode: the tensor holds planes [d_base, d_base + D) of a Dt-plane volume (else Dt = D, 0)
  int Z, Y, X;   // voxel grid extent
  int align_corners;
};

struct Trilinear {
  int x0, y0, z0;
  float wx[2], wy[2], wz[2];   // weight of the low / high corner per axis
  bool valid;                  // all three normalised coordinates inside [-1, 1]
};

// oracle/global_branch.py `lift_grid`: 4-term dots left to right, then divide, normalise
__device__ __forceinline__ void projection_dots(const float* __restrict__ Pm, float x, float y, float z, float& uh,
                                                float& vh, float& wh) {
  auto row = [&](int r) {
    float a = __fmul_rn(Pm[4 * r + 0], x);
    a = __fadd_rn(a, __fmul_rn(Pm[4 * r + 1], y));
    a = __fadd_rn(a, __fmul_rn(Pm[4 * r + 2], z));
    return __fadd_rn(a, Pm[4 * r + 3]);
  };
  uh = row(0); vh = row(1); wh = row(2);
}
__device__ __forceinline__ Trilinear trilinear_finish(float uh, float vh, float wh, float z, const LiftGeom& g) {
  float u = __fdiv_rn(uh, wh), v = __fdiv_rn(vh, wh);
  auto norm = [](float c, float lo, float hi) {
    return __fsub_rn(__fmul_rn(__fdiv_rn(__fsub_rn(c, lo), __fsub_rn(hi, lo)), 2.f), 1.f);
  };
  float gx = norm(u, g.cv[0], g.cv[1]), gy = norm(v, g.cv[2], g.cv[3]), gz = norm(z, g.cv[4], g.cv[5]);
  Trilinear t;
  t.valid = (gx >= -1.f && gx <= 1.f && gy >= -1.f && gy <= 1.f && gz >= -1.f && gz <= 1.f);
  if (!(fabsf(gx) < 1e9f) || !(fabsf(gy) < 1e9f) || !(fabsf(gz) < 1e9f)) { gx = gy = gz = -2.f; }  // oracle: non-finite -> -2
  float ix = unnormalize(gx, g.W, g.align_corners);
  float iy = unnormalize(gy, g.H, g.align_corners);
  float iz = unnormalize(gz, g.Dt, g.align_corners);
  float fx0 = floorf(ix), fy0 = floorf(iy), fz0 = floorf(iz);
  t.x0 = (int)fx0; t.y0 = (int)fy0; t.z0 = (int)fz0;
  t.wx[0] = __fsub_rn(__fadd_rn(fx0, 1.f), ix); t.wx[1] = __fsub_rn(ix, fx0);
  t.wy[0] = __fsub_rn(__fadd_rn(fy0, 1.f), iy); t.wy[1] = __fsub_rn(iy, fy0);
  t.wz[0] = __fsub_rn(__fadd_rn(fz0, 1.f), iz); t.wz[1] = __fsub_rn(iz, fz0);
  return t;
}
__device__ __forceinline__ Trilinear trilinear_setup(const float* __restrict__ Pm, float x, float y, float z,
                                                     const LiftGeom& g) {
  float uh, vh, wh;
  projection_dots(Pm, x, y, z, uh, vh, wh);
  return trilinear_finish(uh, vh, wh, z, g);
}
// Conservative early-out of the cooperative lift: true only if the exact computation above is certain to give
// valid == false (image coordinate outside the cost-volume range by a margin ~1e-3 of the range, four orders of
// magnitude above the fp32 rounding of the quotient).  ~42 % of the KITTI voxel grid is outside the frustum, in
// runs much longer than a warp, so most warps skip the five IEEE divisions and the corner set-up entirely.
__device__ __forceinline__ bool clearly_outside(float uh, float vh, float wh, float z, const LiftGeom& g) {
  if (!(wh > 1e-6f) || !(wh < 1e30f)) return false;
  const float mx = 1e-3f * (g.cv[1] - g.cv[0]) + 1e-3f, my = 1e-3f * (g.cv[3] - g.cv[2]) + 1e-3f,
              mz = 1e-3f * (g.cv[5] - g.cv[4]) + 1e-3f;
  // fabs() of the products: the ranges are given low-to-high (checked on the host); written for wh > 0
  return uh < (g.cv[0] - mx) * wh || uh > (g.cv[1] + mx) * wh || vh < (g.cv[2] - my) * wh || vh > (g.cv[3] + my) * wh ||
         z < g.cv[4] - mz || z > g.cv[5] + mz;
}

// NDHWC bf16 volume -> NDHWC (bf16|f32) or NCDHW f32.  One thread per voxel: the projection / index /
// weight computation (5 IEEE divisions) is done once and amortised over all channels, which are
// processed 16 at a time (2 x 16-byte loads per corner).  EXACT = separately rounded mul/add in the
// oracle's corner order (bit-exact vs oracle/grid_sample.py; used for fp32 outputs); !EXACT = fused
// multiply-add (bf16 product path: half the math instructions, differs by <= 1 bf16 ulp).
// (The first version used 4 lanes per voxel with the setup replicated: ncu showed 619 instructions
// per thread and the SM issue-bound at 75 % while DRAM sat at 17 %.)
template <typename OutT, bool OUT_NDHWC, bool EXACT>
__global__ void __launch_bounds__(256)
lift_ndhwc_kernel(const __nv_bfloat16* __restrict__ vol, const float* __restrict__ proj, const float* __restrict__ zs,
                  const float* __restrict__ ys, const float* __restrict__ xs, OutT* __restrict__ out,
                  uint8_t* __restrict__ valid, int C, LiftGeom g, int64_t total /* N*Z*Y*X */) {
  const int64_t ZYX = (int64_t)g.Z * g.Y * g.X;
  for (int64_t nv = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; nv < total; nv += (int64_t)gridDim.x * blockDim.x) {
    const int64_t n = nv / ZYX;
    const int64_t vox = nv - n * ZYX;
    const int xi = (int)(vox % g.X);
    const int yi = (int)((vox / g.X) % g.Y);
    const int zi = (int)(vox / ((int64_t)g.X * g.Y));
    const Trilinear t = trilinear_setup(proj + n * 12, __ldg(xs + xi), __ldg(ys + yi), __ldg(zs + zi), g);
    if (valid) valid[nv] = t.valid ? 1 : 0;
    float ww[8];
    int64_t off[8];
    unsigned m = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {   // tnw,tne,tsw,tse,bnw,bne,bsw,bse
      const int xx = t.x0 + (k & 1), yy = t.y0 + ((k >> 1) & 1), zz = t.z0 + (k >> 2) - g.d_base;
      const bool in = t.valid && xx >= 0 && xx < g.W && yy >= 0 && yy < g.H && zz >= 0 && zz < g.D;
      m |= in ? (1u << k) : 0u;
      off[k] = in ? (((int64_t)zz * g.H + yy) * g.W + xx) * C : 0;
      ww[k] = __fmul_rn(__fmul_rn(t.wx[k & 1], t.wy[(k >> 1) & 1]), t.wz[k >> 2]);
    }
    const __nv_bfloat16* base = vol + n * (int64_t)g.D * g.H * g.W * C;
    for (int c0 = 0; c0 < C; c0 += 16) {
      float acc[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) acc[j] = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        if (m & (1u << k)) {
          const uint4* src = reinterpret_cast<const uint4*>(base + off[k] + c0);
          const uint4 q0 = __ldg(src), q1 = __ldg(src + 1);
          const uint32_t w[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            if (EXACT) {
              acc[2 * j] = __fadd_rn(acc[2 * j], __fmul_rn(bf16_lo(w[j]), ww[k]));
              acc[2 * j + 1] = __fadd_rn(acc[2 * j + 1], __fmul_rn(bf16_hi(w[j]), ww[k]));
            } else {
              acc[2 * j] = fmaf(bf16_lo(w[j]), ww[k], acc[2 * j]);
              acc[2 * j + 1] = fmaf(bf16_hi(w[j]), ww[k], acc[2 * j + 1]);
            }
          }
        }
      }
      if (OUT_NDHWC) {
        OutT* o = out + nv * C + c0;
        if (sizeof(OutT) == 2) {
          st_cs_v4(o, make_uint4(pack_bf16x2(acc[0], acc[1]), pack_bf16x2(acc[2], acc[3]), pack_bf16x2(acc[4], acc[5]),
                                 pack_bf16x2(acc[6], acc[7])));
          st_cs_v4(reinterpret_cast<__nv_bfloat16*>(o) + 8,
                   make_uint4(pack_bf16x2(acc[8], acc[9]), pack_bf16x2(acc[10], acc[11]), pack_bf16x2(acc[12], acc[13]),
                              pack_bf16x2(acc[14], acc[15])));
        } else {
#pragma unroll
          for (int j = 0; j < 16; j += 4)
            st_cs_f4(reinterpret_cast<float*>(o) + j, make_float4(acc[j], acc[j + 1], acc[j + 2], acc[j + 3]));
        }
      } else {
        float* o = reinterpret_cast<float*>(out) + (n * C + c0) * ZYX + vox;
#pragma unroll
        for (int j = 0; j < 16; ++j) __stcs(o + j * ZYX, acc[j]);
      }
    }
  }
}

// v2 of the NDHWC lift (the product path): the one-thread-per-voxel kernel above is bound by the L1
// tag stage, not by HBM -- ncu (profiles/r01_step_v5.*): every LDG.128 of a warp touches 32 different
// voxel rows (16.7 sectors per request, l1tex 94 % busy, DRAM 20 %).  Here a warp still owns 32
// consecutive voxels and every lane still does the projection / floor / weight set-up of ONE voxel
// (so the 5 IEEE divisions are never replicated), but the gather is done by LPV = C/8 adjacent lanes
// per voxel, 16 bytes of the channel row each: one warp-wide LDG.128 now covers 32/LPV whole rows
// (8 rows of 64 B for C = 32) and one STG.128 writes 32/LPV whole output rows -- 4x fewer L1 tag
// look-ups per byte.  The per-voxel set-up (8 corner weights, base voxel index, in-bounds mask) moves
// from the owning lane to the gathering lanes with warp shuffles; rounds whose voxels are all outside
// the frustum are skipped on a ballot.  Arithmetic per channel is the same instruction sequence as the
// kernel above, so results are bit-identical to it (and EXACT = bit-identical to the oracle).
struct LiftDivs { FastDiv X, Y, Z; };

template <typename OutT, bool EXACT>
__global__ void __launch_bounds__(256, 4)
lift_ndhwc_coop_kernel(const __nv_bfloat16* __restrict__ vol, const float* __restrict__ proj,
                       const float* __restrict__ zs, const float* __restrict__ ys, const float* __restrict__ xs,
                       OutT* __restrict__ out, uint8_t* __restrict__ valid, int C, int log_lpv, LiftGeom g,
                       LiftDivs dv, uint32_t total /* N*Z*Y*X < 2^31 */) {
  const int lane = threadIdx.x & 31;
  const int HW = g.H * g.W;
  const int DHW = g.D * HW;
  const int lpv = 1 << log_lpv;                 // lanes per voxel
  const int vpr = 32 >> log_lpv;                // voxels per round
  const int sub = lane & (lpv - 1);             // which 8-channel group of the row this lane gathers
  const int vsel = lane >> log_lpv;             // which voxel of the round
  const uint32_t warp0 = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t nwarps = (gridDim.x * blockDim.x) >> 5;
  const uint32_t row16 = (uint32_t)C >> 3;      // 16-byte pieces per voxel row
  for (uint32_t chunk = warp0; chunk * 32u < total; chunk += nwarps) {
    // ---- phase A: lane i sets up voxel chunk*32 + i
    const uint32_t nv = chunk * 32u + lane;
    float ww[8];
    int base = 0;
    unsigned m = 0;
    if (nv < total) {
      const uint32_t q1 = fdiv(nv, dv.X), xi = nv - q1 * dv.X.d;
      const uint32_t q2 = fdiv(q1, dv.Y), yi = q1 - q2 * dv.Y.d;
      const uint32_t n = fdiv(q2, dv.Z), zi = q2 - n * dv.Z.d;
      const float zc = __ldg(zs + zi);
      float uh, vh, wh;
      projection_dots(proj + n * 12, __ldg(xs + xi), __ldg(ys + yi), zc, uh, vh, wh);
      bool tvalid = false;
#pragma unroll
      for (int k = 0; k < 8; ++k) ww[k] = 0.f;
      if (!(g.range_ordered && clearly_outside(uh, vh, wh, zc, g))) {
        const Trilinear t = trilinear_finish(uh, vh, wh, zc, g);
        tvalid = t.valid;
        // per-axis in-bounds bits of the low / high corner (unsigned compare = both bounds)
        const int z0 = t.z0 - g.d_base;
        const bool x0i = (unsigned)t.x0 < (unsigned)g.W, x1i = (unsigned)(t.x0 + 1) < (unsigned)g.W;
        const bool y0i = (unsigned)t.y0 < (unsigned)g.H, y1i = (unsigned)(t.y0 + 1) < (unsigned)g.H;
        const bool z0i = (unsigned)z0 < (unsigned)g.D, z1i = (unsigned)(z0 + 1) < (unsigned)g.D;
        const unsigned mxy = (x0i && y0i ? 1u : 0u) | (x1i && y0i ? 2u : 0u) | (x0i && y1i ? 4u : 0u) | (x1i && y1i ? 8u : 0u);
        m = t.valid ? ((z0i ? mxy : 0u) | (z1i ? mxy << 4 : 0u)) : 0u;   // tnw,tne,tsw,tse,bnw,bne,bsw,bse
        const float wxy[4] = {__fmul_rn(t.wx[0], t.wy[0]), __fmul_rn(t.wx[1], t.wy[0]), __fmul_rn(t.wx[0], t.wy[1]),
                              __fmul_rn(t.wx[1], t.wy[1])};
#pragma unroll
        for (int k = 0; k < 8; ++k) ww[k] = __fmul_rn(wxy[k & 3], t.wz[k >> 2]);
        // voxel index of corner 0 in the whole batch (may be "negative" when corner 0 itself is outside; only
        // in-bounds corners are dereferenced).  Host guarantees N*D*H*W < 2^31.
        base = m ? (int)n * DHW + (z0 * g.H + t.y0) * g.W + t.x0 : 0;
      }
      if (valid) valid[nv] = tvalid ? 1 : 0;
    } else {
#pragma unroll
      for (int k = 0; k < 8; ++k) ww[k] = 0.f;
    }
    // ---- whole chunk outside the frustum (42 % of the KITTI grid, in long runs): write its zero rows and move on
    if (__ballot_sync(0xffffffffu, m != 0u) == 0u && chunk * 32u + 32u <= total) {
      const int64_t pieces = (int64_t)chunk * 32 * row16;          // 16-byte pieces (bf16) of the chunk's output rows
      for (uint32_t i = lane; i < 32u * row16; i += 32u) {
        if (sizeof(OutT) == 2) {
          st_cs_v4(reinterpret_cast<uint4*>(out) + pieces + i, make_uint4(0u, 0u, 0u, 0u));
        } else {
          st_cs_f4(reinterpret_cast<float*>(out) + (pieces + i) * 8, make_float4(0.f, 0.f, 0.f, 0.f));
          st_cs_f4(reinterpret_cast<float*>(out) + (pieces + i) * 8 + 4, make_float4(0.f, 0.f, 0.f, 0.f));
        }
      }
      continue;
    }
    // ---- phase B: lpv rounds of vpr voxels; lane group `vsel` gathers voxel r*vpr + vsel
    for (int r = 0; r < lpv; ++r) {
      const int src = r * vpr + vsel;
      const unsigned mm = __shfl_sync(0xffffffffu, m, src);
      const uint32_t onv = chunk * 32u + src;
      const bool live = onv < total;              // tail chunk only; dead voxels carry mm == 0
      // accumulators as fp32 pairs: sm_100 has packed FFMA2 / FMUL2 / FADD2 (two IEEE fp32 operations per issue
      // slot, results identical to the scalar instructions) -- ncu showed this kernel bound by instruction issue
      float2 acc2[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) acc2[j] = make_float2(0.f, 0.f);
      if (__ballot_sync(0xffffffffu, mm != 0u) != 0u) {
        const int b = __shfl_sync(0xffffffffu, base, src);
        float w[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) w[k] = __shfl_sync(0xffffffffu, ww[k], src);
        uint4 q[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          q[k] = make_uint4(0u, 0u, 0u, 0u);
          if (mm & (1u << k)) {
            const uint32_t v = (uint32_t)(b + (k >> 2) * HW + ((k >> 1) & 1) * g.W + (k & 1));
            q[k] = __ldg(reinterpret_cast<const uint4*>(vol) + (v * row16 + sub));      // host: volume < 2^32 x 16 B
          }
        }
        // masked corners were loaded as zeros and the weights are finite: x + 0 * w == x, so no per-corner branch
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const uint32_t u[4] = {q[k].x, q[k].y, q[k].z, q[k].w};
          const float2 wk = make_float2(w[k], w[k]);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float2 v = make_float2(bf16_lo(u[j]), bf16_hi(u[j]));
            if (EXACT) {   // scalar intrinsics: nvcc contracts __fadd2_rn(__fmul2_rn()) into FFMA2 (seen in the SASS)
              acc2[j].x = __fadd_rn(acc2[j].x, __fmul_rn(v.x, wk.x));
              acc2[j].y = __fadd_rn(acc2[j].y, __fmul_rn(v.y, wk.y));
            } else {
              acc2[j] = __ffma2_rn(v, wk, acc2[j]);
            }
          }
        }
      }
      const float acc[8] = {acc2[0].x, acc2[0].y, acc2[1].x, acc2[1].y, acc2[2].x, acc2[2].y, acc2[3].x, acc2[3].y};
      if (!live) continue;
      OutT* o = out + (int64_t)onv * C + sub * 8;
      if (sizeof(OutT) == 2) {
        st_cs_v4(o, make_uint4(pack_bf16x2(acc[0], acc[1]), pack_bf16x2(acc[2], acc[3]), pack_bf16x2(acc[4], acc[5]),
                               pack_bf16x2(acc[6], acc[7])));
      } else {
        st_cs_f4(reinterpret_cast<float*>(o), make_float4(acc[0], acc[1], acc[2], acc[3]));
        st_cs_f4(reinterpret_cast<float*>(o) + 4, make_float4(acc[4], acc[5], acc[6], acc[7]));
      }
    }
  }
}

// v4 of the NDHWC lift, bf16 output (the product path).  ncu on the cooperative kernel above (r02): issue slots 73 % busy,
// but only 10 % of the executed instructions are the FFMA2 of the interpolation and ~20 % the bf16 -> fp32 unpacks; the
// rest is ten set-up shuffles per round, per-corner predicates with zero-initialised registers (CS2R 7 %), 64-bit address
// chains, constant reloads.  Same recipe as the ROI sampler v4:
//   * the set-up (same arithmetic: validity and corner indices stay bit-exact) leaves 8 corner weights that are ZERO for a
//     corner outside the volume (or an invalid voxel) and a clamped, always valid corner-0 index + per-axis step flags;
//     every load is unconditional, fma(v, 0, acc) == acc for the finite values of the trunk, so the result is
//     bit-identical to the kernel above (tests/test_gpu_voxel_sample.py);
//   * the set-up travels through shared memory (2 STS.128 + 1 STS.64 per lane and chunk; 2 LDS.128 + 1 LDS.64 per round);
//   * one IMAD.WIDE per corner address, offsets in 16-byte units.
template <int LOG_LPV>   // lanes per voxel: C / 8 == 1 << LOG_LPV
__global__ void __launch_bounds__(256, 4)
lift_fast_bf16_kernel(const __nv_bfloat16* __restrict__ vol, const float* __restrict__ proj, const float* __restrict__ zs,
                      const float* __restrict__ ys, const float* __restrict__ xs, __nv_bfloat16* __restrict__ out,
                      uint8_t* __restrict__ valid, LiftGeom g, LiftDivs dv, uint32_t total /* N*Z*Y*X < 2^31 */) {
  constexpr int LPV = 1 << LOG_LPV, VPR = 32 >> LOG_LPV, ROUNDS = LPV, C = 8 * LPV;
  constexpr uint32_t ROW16 = LPV;                                // 16-byte pieces per voxel row
  __shared__ float4 s_w[8][2][2][32];                            // [warp][buffer][weights 0-3 | 4-7][voxel of the chunk]
  __shared__ uint2 s_i[8][2][32];                                // corner-0 voxel index (clamped), step flags | valid bit
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int sub = lane & (LPV - 1);
  const int vsel = lane >> LOG_LPV;
  const int HW = g.H * g.W;
  const int DHW = g.D * HW;
  const uint32_t warp0 = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t nwarps = (gridDim.x * blockDim.x) >> 5;
  const uint4* const vol16 = reinterpret_cast<const uint4*>(vol) + sub;
  uint4* const out16 = reinterpret_cast<uint4*>(out) + (size_t)vsel * ROW16 + sub;
  int buf = 0;
  for (uint32_t chunk = warp0; chunk * 32u < total; chunk += nwarps, buf ^= 1) {
    // ---- phase A: lane i sets up voxel chunk*32 + i (arithmetic of lift_ndhwc_coop_kernel, bit-exact indices / validity)
    const uint32_t nv = chunk * 32u + lane;
    float4 wa = make_float4(0.f, 0.f, 0.f, 0.f), wb = wa;
    uint2 id = make_uint2(0u, 0u);
    if (nv < total) {
      const uint32_t q1 = fdiv(nv, dv.X), xi = nv - q1 * dv.X.d;
      const uint32_t q2 = fdiv(q1, dv.Y), yi = q1 - q2 * dv.Y.d;
      const uint32_t n = fdiv(q2, dv.Z), zi = q2 - n * dv.Z.d;
      const float zc = __ldg(zs + zi);
      float uh, vh, wh;
      projection_dots(proj + n * 12, __ldg(xs + xi), __ldg(ys + yi), zc, uh, vh, wh);
      bool tvalid = false;
      if (!(g.range_ordered && clearly_outside(uh, vh, wh, zc, g))) {
        const Trilinear t = trilinear_finish(uh, vh, wh, zc, g);
        tvalid = t.valid;
        if (t.valid) {
          const int z0 = t.z0 - g.d_base;
          // per-axis weights, zero where that corner coordinate is outside the volume
          const float wx0 = (unsigned)t.x0 < (unsigned)g.W ? t.wx[0] : 0.f, wx1 = (unsigned)(t.x0 + 1) < (unsigned)g.W ? t.wx[1] : 0.f;
          const float wy0 = (unsigned)t.y0 < (unsigned)g.H ? t.wy[0] : 0.f, wy1 = (unsigned)(t.y0 + 1) < (unsigned)g.H ? t.wy[1] : 0.f;
          const float wz0 = (unsigned)z0 < (unsigned)g.D ? t.wz[0] : 0.f, wz1 = (unsigned)(z0 + 1) < (unsigned)g.D ? t.wz[1] : 0.f;
          const float w00 = __fmul_rn(wx0, wy0), w10 = __fmul_rn(wx1, wy0), w01 = __fmul_rn(wx0, wy1), w11 = __fmul_rn(wx1, wy1);
          wa = make_float4(__fmul_rn(w00, wz0), __fmul_rn(w10, wz0), __fmul_rn(w01, wz0), __fmul_rn(w11, wz0));
          wb = make_float4(__fmul_rn(w00, wz1), __fmul_rn(w10, wz1), __fmul_rn(w01, wz1), __fmul_rn(w11, wz1));
          const int x0c = min(max(t.x0, 0), g.W - 1), x1c = min(max(t.x0 + 1, 0), g.W - 1);
          const int y0c = min(max(t.y0, 0), g.H - 1), y1c = min(max(t.y0 + 1, 0), g.H - 1);
          const int z0c = min(max(z0, 0), g.D - 1), z1c = min(max(z0 + 1, 0), g.D - 1);
          id.x = (uint32_t)((int)n * DHW + (z0c * g.H + y0c) * g.W + x0c);
          id.y = 8u | (x1c != x0c ? 1u : 0u) | (y1c != y0c ? 2u : 0u) | (z1c != z0c ? 4u : 0u);
        }
      }
      if (valid) valid[nv] = tvalid ? 1 : 0;
    }
    // ---- whole chunk outside the frustum (42 % of the KITTI grid, in long runs): write its zero rows and move on
    if (__ballot_sync(0xffffffffu, id.y != 0u) == 0u && chunk * 32u + 32u <= total) {
      uint4* z = reinterpret_cast<uint4*>(out) + (size_t)chunk * 32u * ROW16;
      for (uint32_t i = lane; i < 32u * ROW16; i += 32u) st_cs_v4(z + i, make_uint4(0u, 0u, 0u, 0u));
      continue;
    }
    s_w[wib][buf][0][lane] = wa;
    s_w[wib][buf][1][lane] = wb;
    s_i[wib][buf][lane] = id;
    __syncwarp();
    // ---- phase B: LPV rounds of VPR consecutive voxels; lane group `vsel` gathers voxel r*VPR + vsel
    uint4* o = out16 + (size_t)chunk * 32u * ROW16;
#pragma unroll
    for (int r = 0; r < ROUNDS; ++r) {
      const int src = r * VPR + vsel;
      const uint2 idr = s_i[wib][buf][src];
      const bool live = chunk * 32u + (uint32_t)src < total;
      float2 acc[4] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
      if (__ballot_sync(0xffffffffu, idr.y != 0u) != 0u) {   // (a round of voxels outside the frustum writes zeros)
        const float4 w0 = s_w[wib][buf][0][src], w1 = s_w[wib][buf][1][src];
        const float w[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
        const uint32_t o000 = idr.x * ROW16;
        const uint32_t dx = (idr.y & 1u) ? ROW16 : 0u, dy = (idr.y & 2u) ? (uint32_t)g.W * ROW16 : 0u,
                       dz = (idr.y & 4u) ? (uint32_t)HW * ROW16 : 0u;
        uint4 q[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const uint32_t off = o000 + ((k & 1) ? dx : 0u) + ((k & 2) ? dy : 0u) + ((k & 4) ? dz : 0u);
          uint64_t a;
          asm("mad.wide.u32 %0, %1, 16, %2;" : "=l"(a) : "r"(off), "l"(reinterpret_cast<uint64_t>(vol16)));
          q[k] = __ldg(reinterpret_cast<const uint4*>(a));
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const uint32_t u[4] = {q[k].x, q[k].y, q[k].z, q[k].w};
          const float2 wk = make_float2(w[k], w[k]);
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[j] = __ffma2_rn(make_float2(bf16_lo(u[j]), bf16_hi(u[j])), wk, acc[j]);
        }
      }
      if (live)
        st_cs_v4(o + (size_t)r * VPR * ROW16,
                 make_uint4(pack_bf16x2(acc[0].x, acc[0].y), pack_bf16x2(acc[1].x, acc[1].y), pack_bf16x2(acc[2].x, acc[2].y),
                            pack_bf16x2(acc[3].x, acc[3].y)));
    }
  }
}

// NCDHW fp32 volume (the reference's layout) -> NCDHW fp32; one thread per voxel, channel loop.
__global__ void __launch_bounds__(256)
lift_ncdhw_kernel(const float* __restrict__ vol, const float* __restrict__ proj, const float* __restrict__ zs,
                  const float* __restrict__ ys, const float* __restrict__ xs, float* __restrict__ out,
                  uint8_t* __restrict__ valid, int C, LiftGeom g, int64_t total /* N*Z*Y*X */) {
  const int64_t ZYX = (int64_t)g.Z * g.Y * g.X;
  const int64_t DHW = (int64_t)g.D * g.H * g.W;
  for (int64_t nv = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; nv < total; nv += (int64_t)gridDim.x * blockDim.x) {
    int64_t n = nv / ZYX;
    int64_t vox = nv - n * ZYX;
    int xi = (int)(vox % g.X);
    int yi = (int)((vox / g.X) % g.Y);
    int zi = (int)(vox / ((int64_t)g.X * g.Y));
    Trilinear t = trilinear_setup(proj + n * 12, __ldg(xs + xi), __ldg(ys + yi), __ldg(zs + zi), g);
    int64_t off[8];
    float ww[8];
    unsigned m = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      int xx = t.x0 + (k & 1), yy = t.y0 + ((k >> 1) & 1), zz = t.z0 + (k >> 2) - g.d_base;
      bool in = t.valid && xx >= 0 && xx < g.W && yy >= 0 && yy < g.H && zz >= 0 && zz < g.D;
      m |= in ? (1u << k) : 0u;
      off[k] = in ? ((int64_t)zz * g.H + yy) * g.W + xx : 0;
      ww[k] = __fmul_rn(__fmul_rn(t.wx[k & 1], t.wy[(k >> 1) & 1]), t.wz[k >> 2]);
    }
    if (valid) valid[nv] = t.valid ? 1 : 0;
    const float* vb = vol + n * C * DHW;
    float* o = out + n * C * ZYX + vox;
    for (int c = 0; c < C; ++c) {
      float a = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k)
        if (m & (1u << k)) a = __fadd_rn(a, __fmul_rn(__ldg(vb + c * DHW + off[k]), ww[k]));
      __stcs(o + c * ZYX, a);
    }
  }
}

__global__ void lift_indices_kernel(const float* __restrict__ proj, const float* __restrict__ zs,
                                    const float* __restrict__ ys, const float* __restrict__ xs, int32_t* __restrict__ idx,
                                    uint8_t* __restrict__ valid, LiftGeom g, int64_t total) {
  const int64_t ZYX = (int64_t)g.Z * g.Y * g.X;
  for (int64_t nv = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; nv < total; nv += (int64_t)gridDim.x * blockDim.x) {
    int64_t n = nv / ZYX;
    int64_t vox = nv - n * ZYX;
    int xi = (int)(vox % g.X);
    int yi = (int)((vox / g.X) % g.Y);
    int zi = (int)(vox / ((int64_t)g.X * g.Y));
    Trilinear t = trilinear_setup(proj + n * 12, xs[xi], ys[yi], zs[zi], g);
    idx[3 * nv] = t.x0;
    idx[3 * nv + 1] = t.y0;
    idx[3 * nv + 2] = t.z0;
    valid[nv] = t.valid ? 1 : 0;
  }
}

FastDiv make_fastdiv(int64_t d) {
  FastDiv f;
  uint32_t l = 0;
  while ((1ull << l) < (uint64_t)d) ++l;
  f.m = (uint32_t)((((uint64_t)1 << 32) * (((uint64_t)1 << l) - (uint64_t)d)) / (uint64_t)d + 1);
  f.s1 = l < 1 ? l : 1; f.s2 = l > 0 ? l - 1 : 0; f.d = (uint32_t)d;
  return f;
}

int grid_for(int64_t total, int per_sm = 8) {
  return (int)std::max<int64_t>(1, std::min<int64_t>(ceil_div(total, 256), (int64_t)sm_count() * per_sm));
}

}  // namespace
}  // namespace snvc

using namespace snvc;

extern "C" int64_t snvc_roi_voxel_sample_workspace_bytes(int64_t N, int64_t C, int64_t Hf, int64_t Wf) {
  return 2 * N * C * Hf * Wf * 4;
}

extern "C" int snvc_roi_voxel_sample_fwd(const float* feat_l, const float* feat_r, const float* pts_l,
                                         const float* pts_r, void* out, void* workspace, int64_t N, int64_t C,
                                         int64_t Hf, int64_t Wf, int64_t P, float res_x, float res_y,
                                         int32_t out_dtype, int32_t out_layout, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  SNVC_CHECK_ARG(N >= 0 && C > 0 && Hf > 0 && Wf > 0 && P >= 0, "bad dimensions");
  if (N * P == 0) return 0;
  SNVC_CHECK_ARG(feat_l && feat_r && pts_l && pts_r && out && workspace, "null pointer");
  SNVC_CHECK_ARG(C % 4 == 0, "C must be a multiple of 4 (got %lld)", (long long)C);
  SNVC_CHECK_ARG((reinterpret_cast<uintptr_t>(workspace) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0,
                 "out / workspace must be 16-byte aligned");
  float* wl = (float*)workspace;
  float* wr = wl + N * C * Hf * Wf;
  const int64_t nfeat = N * C * Hf * Wf;
  const char* rmode = opt(OPT_ROI_MODE);
  const int lpv = (int)(C / 8);
  // bf16 NDHWC output (the instance branch's product path): v4 kernel on a bf16 channels-last copy of the features (fused
  // FMA, zero-weight masking, set-up through shared memory).  SNVC_ROI_MODE=fast32 keeps fp32 features in the same kernel,
  // =v3 / coop1 / thread select the bit-exact kernels (A/B runs, tests).
  const bool fast32 = rmode && rmode[0] == 'f';
  if (out_layout == SNVC_NDHWC && out_dtype == SNVC_BF16 && C % 8 == 0 && lpv <= 16 && (lpv & (lpv - 1)) == 0 &&
      N * P < (1ll << 30) && N * Hf * Wf * C < (1ll << 31) && (reinterpret_cast<uintptr_t>(workspace) & 31) == 0 &&
      (!rmode || fast32)) {
    const void *fa, *fb;
    if (fast32) {
      nchw_to_nhwc_kernel<<<grid_for(nfeat), 256, 0, stream>>>(feat_l, feat_r, wl, wr, (int)C, (int)(Hf * Wf), nfeat);
      fa = wl; fb = wr;
    } else {
      uint4* ha = (uint4*)workspace;
      uint4* hb = ha + nfeat / 8;
      nchw_to_nhwc_bf16_kernel<<<grid_for(nfeat / 8), 256, 0, stream>>>(feat_l, feat_r, ha, hb, (int)C, (int)(Hf * Wf), nfeat / 8);
      fa = ha; fb = hb;
    }
    if (int e = launch_status("nchw_to_nhwc_kernel")) return e;
    // division by a power-of-two resolution (256 in the shipped configuration) is an exact multiplication
    int ex, ey;
    const bool pow2 = res_x > 0.f && res_y > 0.f && std::frexp(res_x, &ex) == 0.5f && std::frexp(res_y, &ey) == 0.5f &&
                      ex > -100 && ex < 100 && ey > -100 && ey < 100;
    const float inv_x = pow2 ? 1.f / res_x : 0.f, inv_y = pow2 ? 1.f / res_y : 0.f;
    const int blocks = (int)std::max<int64_t>(1, std::min<int64_t>(ceil_div(N * P, 128), (int64_t)sm_count() * 4));
#define SNVC_ROI_FAST(L)                                                                                                      \
  do {                                                                                                                        \
    if (fast32)                                                                                                               \
      roi_sample_fast_bf16_kernel<L, false><<<blocks, 256, 0, stream>>>(fa, fb, pts_l, pts_r, (__nv_bfloat16*)out, (int)Hf,   \
                                                                        (int)Wf, make_fastdiv(P), res_x, res_y, inv_x, inv_y, \
                                                                        (uint32_t)(N * P));                                   \
    else                                                                                                                      \
      roi_sample_fast_bf16_kernel<L, true><<<blocks, 256, 0, stream>>>(fa, fb, pts_l, pts_r, (__nv_bfloat16*)out, (int)Hf,    \
                                                                       (int)Wf, make_fastdiv(P), res_x, res_y, inv_x, inv_y,  \
                                                                       (uint32_t)(N * P));                                    \
  } while (0)
    switch (lpv) {
      case 1: SNVC_ROI_FAST(0); break;
      case 2: SNVC_ROI_FAST(1); break;
      case 4: SNVC_ROI_FAST(2); break;
      case 8: SNVC_ROI_FAST(3); break;
      default: SNVC_ROI_FAST(4); break;
    }
#undef SNVC_ROI_FAST
    return launch_status("roi_sample_fast_bf16_kernel");
  }
  nchw_to_nhwc_kernel<<<grid_for(nfeat), 256, 0, stream>>>(feat_l, feat_r, wl, wr, (int)C, (int)(Hf * Wf), nfeat);
  if (int e = launch_status("nchw_to_nhwc_kernel")) return e;
  if (out_layout == SNVC_NDHWC) {
    SNVC_CHECK_ARG(C % 8 == 0, "NDHWC output needs C %% 8 == 0");
    const int64_t total = N * P * (2 * C / 8);
    // cooperative kernels (set-up once per point, C/8 lanes per (point, view)); SNVC_ROI_MODE=thread keeps v1 (A/B runs)
    // v3 (both views per round, 256-bit corner loads; bit-exact); SNVC_ROI_MODE=coop1 keeps v2 (A/B runs)
    if (lpv <= 16 && (lpv & (lpv - 1)) == 0 && N * P < (1ll << 30) && N * Hf * Wf < (1ll << 31) &&
        (reinterpret_cast<uintptr_t>(workspace) & 31) == 0 && (reinterpret_cast<uintptr_t>(out) & 31) == 0 &&
        (out_dtype == SNVC_BF16 || out_dtype == SNVC_F32) && !(rmode && (rmode[0] == 't' || rmode[0] == 'c'))) {
      int log_lpv = 0;
      while ((1 << log_lpv) < lpv) ++log_lpv;
      const int blocks = (int)std::max<int64_t>(1, std::min<int64_t>(ceil_div(N * P, 128), (int64_t)sm_count() * 3));
      if (out_dtype == SNVC_BF16)
        roi_sample_coop2_kernel<__nv_bfloat16><<<blocks, 256, 0, stream>>>(wl, wr, pts_l, pts_r, (__nv_bfloat16*)out, (int)C,
                                                                           log_lpv, (int)Hf, (int)Wf, make_fastdiv(P), res_x,
                                                                           res_y, (uint32_t)(N * P));
      else
        roi_sample_coop2_kernel<float><<<blocks, 256, 0, stream>>>(wl, wr, pts_l, pts_r, (float*)out, (int)C, log_lpv, (int)Hf,
                                                                   (int)Wf, make_fastdiv(P), res_x, res_y, (uint32_t)(N * P));
      return launch_status("roi_sample_coop2_kernel");
    }
    if (lpv <= 32 && (lpv & (lpv - 1)) == 0 && N * P < (1ll << 31) && N * Hf * Wf < (1ll << 31) &&
        (out_dtype == SNVC_BF16 || out_dtype == SNVC_F32) && !(rmode && rmode[0] == 't')) {
      int log_lpv = 0;
      while ((1 << log_lpv) < lpv) ++log_lpv;
      const int blocks = (int)std::max<int64_t>(1, std::min<int64_t>(ceil_div(N * P, 256), (int64_t)sm_count() * 4));
      if (out_dtype == SNVC_BF16)
        roi_sample_coop_kernel<__nv_bfloat16><<<blocks, 256, 0, stream>>>(wl, wr, pts_l, pts_r, (__nv_bfloat16*)out, (int)C, log_lpv,
                                                                          (int)Hf, (int)Wf, make_fastdiv(P), res_x, res_y,
                                                                          (uint32_t)(N * P));
      else
        roi_sample_coop_kernel<float><<<blocks, 256, 0, stream>>>(wl, wr, pts_l, pts_r, (float*)out, (int)C, log_lpv, (int)Hf,
                                                                  (int)Wf, make_fastdiv(P), res_x, res_y, (uint32_t)(N * P));
      return launch_status("roi_sample_coop_kernel");
    }
    if (out_dtype == SNVC_BF16)
      roi_sample_ndhwc_kernel<__nv_bfloat16><<<grid_for(total, 16), 256, 0, stream>>>(
          wl, wr, pts_l, pts_r, (__nv_bfloat16*)out, (int)C, (int)Hf, (int)Wf, P, res_x, res_y, total);
    else if (out_dtype == SNVC_F32)
      roi_sample_ndhwc_kernel<float><<<grid_for(total, 16), 256, 0, stream>>>(wl, wr, pts_l, pts_r, (float*)out, (int)C,
                                                                              (int)Hf, (int)Wf, P, res_x, res_y, total);
    else
      return fail(SNVC_E_UNSUPPORTED, "roi sample: out_dtype must be bf16 or f32");
    return launch_status("roi_sample_ndhwc_kernel");
  }
  if (out_layout == SNVC_NCDHW) {
    if (out_dtype != SNVC_F32) return fail(SNVC_E_UNSUPPORTED, "roi sample NCDHW: out_dtype must be f32");
    SNVC_CHECK_ARG(N <= 65535, "N too large");
    dim3 grid((unsigned)std::min<int64_t>(ceil_div(P, 256), 4096), 2, (unsigned)N);
    roi_sample_ncdhw_kernel<<<grid, 256, 0, stream>>>(wl, wr, pts_l, pts_r, (float*)out, (int)C, (int)Hf, (int)Wf, P,
                                                      res_x, res_y);
    return launch_status("roi_sample_ncdhw_kernel");
  }
  return fail(SNVC_E_BADARG, "unknown out_layout %d", out_layout);
}

extern "C" int snvc_roi_voxel_sample_indices(const float* pts, int32_t* idx, uint8_t* mask, int64_t N, int64_t P,
                                             int64_t Hf, int64_t Wf, float res_x, float res_y, void* stream_) {
  if (N * P == 0) return 0;
  SNVC_CHECK_ARG(pts && idx && mask, "null pointer");
  roi_indices_kernel<<<grid_for(N * P), 256, 0, (cudaStream_t)stream_>>>(pts, idx, mask, N, P, (int)Hf, (int)Wf, res_x,
                                                                        res_y);
  return launch_status("roi_indices_kernel");
}

static int make_geom(LiftGeom& g, const float* cv, int64_t D, int64_t H, int64_t W, int64_t Z, int64_t Y, int64_t X,
                     int ac) {
  SNVC_CHECK_ARG(cv != nullptr, "cv_range_host is null");
  SNVC_CHECK_ARG(D > 0 && H > 0 && W > 0 && Z > 0 && Y > 0 && X > 0, "bad dimensions");
  SNVC_CHECK_ARG(D < (1 << 20) && H < (1 << 20) && W < (1 << 20) && Z < (1 << 20) && Y < (1 << 20) && X < (1 << 20),
                 "dimension too large");
  for (int i = 0; i < 6; ++i) g.cv[i] = cv[i];
  g.range_ordered = (cv[0] < cv[1] && cv[2] < cv[3] && cv[4] < cv[5]) ? 1 : 0;
  g.D = (int)D; g.H = (int)H; g.W = (int)W; g.Z = (int)Z; g.Y = (int)Y; g.X = (int)X;
  g.Dt = (int)D; g.d_base = 0;
  g.align_corners = ac ? 1 : 0;
  return 0;
}

static int lift_fwd_impl(const void* vol, const float* proj, const float* zs, const float* ys, const float* xs,
                         const float* cv_range_host, void* out, uint8_t* valid, int64_t N, int64_t C, int64_t D,
                         int64_t H, int64_t W, int64_t Z, int64_t Y, int64_t X, int32_t align_corners, int32_t in_dtype,
                         int32_t in_layout, int32_t out_dtype, int32_t out_layout, int64_t D_total, int64_t d_base,
                         void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (N == 0 || Z == 0) return 0;
  LiftGeom g;
  if (int e = make_geom(g, cv_range_host, D, H, W, Z, Y, X, align_corners)) return e;
  SNVC_CHECK_ARG(D_total > 0 && D_total < (1 << 20) && d_base > -(1 << 20) && d_base < (1 << 20), "bad depth-slab range");
  g.Dt = (int)D_total; g.d_base = (int)d_base;
  SNVC_CHECK_ARG(vol && proj && zs && ys && xs && out, "null pointer");
  SNVC_CHECK_ARG(N > 0 && C > 0, "bad N / C");
  const int64_t nvox = N * Z * Y * X;
  if (in_layout == SNVC_NDHWC && in_dtype == SNVC_BF16) {
    SNVC_CHECK_ARG(C % 16 == 0, "NDHWC lift needs C %% 16 == 0");
    SNVC_CHECK_ARG((reinterpret_cast<uintptr_t>(vol) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0,
                   "vol / out must be 16-byte aligned");
    const int blocks = grid_for(nvox, 16);
    // cooperative gather (C/8 lanes per voxel) whenever the row splits into a power-of-two number of 16-byte
    // pieces; SNVC_LIFT_MODE=thread keeps the one-thread-per-voxel kernel (A/B runs)
    const int lpv = (int)(C / 8);
    const char* lmode = opt(OPT_LIFT_MODE);
    if (out_layout == SNVC_NDHWC && C % 8 == 0 && lpv <= 32 && (lpv & (lpv - 1)) == 0 && N * D * H * W < (1ll << 31) &&
        nvox < (1ll << 31) && N * D * H * W * lpv < (1ll << 32) && !(lmode && lmode[0] == 't')) {
      int log_lpv = 0;
      while ((1 << log_lpv) < lpv) ++log_lpv;
      LiftDivs dv{make_fastdiv(X), make_fastdiv(Y), make_fastdiv(Z)};
      const int cblocks = (int)std::max<int64_t>(1, std::min<int64_t>(ceil_div(nvox, 256), (int64_t)sm_count() * 4));   // 4 resident blocks / SM, one wave
      if (out_dtype == SNVC_BF16 && log_lpv >= 1 && log_lpv <= 3 && !lmode) {     // v4 (SNVC_LIFT_MODE=coop keeps v3)
        if (log_lpv == 1)
          lift_fast_bf16_kernel<1><<<cblocks, 256, 0, stream>>>((const __nv_bfloat16*)vol, proj, zs, ys, xs, (__nv_bfloat16*)out, valid, g, dv, (uint32_t)nvox);
        else if (log_lpv == 2)
          lift_fast_bf16_kernel<2><<<cblocks, 256, 0, stream>>>((const __nv_bfloat16*)vol, proj, zs, ys, xs, (__nv_bfloat16*)out, valid, g, dv, (uint32_t)nvox);
        else
          lift_fast_bf16_kernel<3><<<cblocks, 256, 0, stream>>>((const __nv_bfloat16*)vol, proj, zs, ys, xs, (__nv_bfloat16*)out, valid, g, dv, (uint32_t)nvox);
        return launch_status("lift_fast_bf16_kernel");
      }
      if (out_dtype == SNVC_BF16)
        lift_ndhwc_coop_kernel<__nv_bfloat16, false><<<cblocks, 256, 0, stream>>>(
            (const __nv_bfloat16*)vol, proj, zs, ys, xs, (__nv_bfloat16*)out, valid, (int)C, log_lpv, g, dv, (uint32_t)nvox);
      else if (out_dtype == SNVC_F32)
        lift_ndhwc_coop_kernel<float, true><<<cblocks, 256, 0, stream>>>((const __nv_bfloat16*)vol, proj, zs, ys, xs,
                                                                         (float*)out, valid, (int)C, log_lpv, g, dv, (uint32_t)nvox);
      else
        return fail(SNVC_E_UNSUPPORTED, "lift: unsupported output type for an NDHWC bf16 volume");
      return launch_status("lift_ndhwc_coop_kernel");
    }
    if (out_layout == SNVC_NDHWC && out_dtype == SNVC_BF16)
      lift_ndhwc_kernel<__nv_bfloat16, true, false><<<blocks, 256, 0, stream>>>(
          (const __nv_bfloat16*)vol, proj, zs, ys, xs, (__nv_bfloat16*)out, valid, (int)C, g, nvox);
    else if (out_layout == SNVC_NDHWC && out_dtype == SNVC_F32)
      lift_ndhwc_kernel<float, true, true><<<blocks, 256, 0, stream>>>((const __nv_bfloat16*)vol, proj, zs, ys, xs,
                                                                       (float*)out, valid, (int)C, g, nvox);
    else if (out_layout == SNVC_NCDHW && out_dtype == SNVC_F32)
      lift_ndhwc_kernel<float, false, true><<<blocks, 256, 0, stream>>>((const __nv_bfloat16*)vol, proj, zs, ys, xs,
                                                                        (float*)out, valid, (int)C, g, nvox);
    else
      return fail(SNVC_E_UNSUPPORTED, "lift: unsupported output type/layout for an NDHWC bf16 volume");
    return launch_status("lift_ndhwc_kernel");
  }
  if (in_layout == SNVC_NCDHW && in_dtype == SNVC_F32) {
    if (!(out_layout == SNVC_NCDHW && out_dtype == SNVC_F32))
      return fail(SNVC_E_UNSUPPORTED, "lift: an NCDHW f32 volume lifts to NCDHW f32 only");
    lift_ncdhw_kernel<<<grid_for(nvox, 16), 256, 0, stream>>>((const float*)vol, proj, zs, ys, xs, (float*)out, valid,
                                                              (int)C, g, nvox);
    return launch_status("lift_ncdhw_kernel");
  }
  return fail(SNVC_E_UNSUPPORTED, "lift: supported inputs are NDHWC bf16 and NCDHW f32");
}

extern "C" int snvc_frustum_lift_fwd(const void* vol, const float* proj, const float* zs, const float* ys,
                                     const float* xs, const float* cv_range_host, void* out, uint8_t* valid, int64_t N,
                                     int64_t C, int64_t D, int64_t H, int64_t W, int64_t Z, int64_t Y, int64_t X,
                                     int32_t align_corners, int32_t in_dtype, int32_t in_layout, int32_t out_dtype,
                                     int32_t out_layout, void* stream_) {
  return lift_fwd_impl(vol, proj, zs, ys, xs, cv_range_host, out, valid, N, C, D, H, W, Z, Y, X, align_corners, in_dtype,
                       in_layout, out_dtype, out_layout, D, 0, stream_);
}

extern "C" int snvc_frustum_lift_slab_fwd(const void* vol, const float* proj, const float* zs, const float* ys,
                                          const float* xs, const float* cv_range_host, void* out, uint8_t* valid,
                                          int64_t N, int64_t C, int64_t D, int64_t H, int64_t W, int64_t Z, int64_t Y,
                                          int64_t X, int32_t align_corners, int32_t in_dtype, int32_t in_layout,
                                          int32_t out_dtype, int32_t out_layout, int64_t D_total, int64_t d_base,
                                          void* stream_) {
  return lift_fwd_impl(vol, proj, zs, ys, xs, cv_range_host, out, valid, N, C, D, H, W, Z, Y, X, align_corners, in_dtype,
                       in_layout, out_dtype, out_layout, D_total, d_base, stream_);
}

extern "C" int snvc_frustum_lift_indices(const float* proj, const float* zs, const float* ys, const float* xs,
                                         const float* cv_range_host, int32_t* idx, uint8_t* valid, int64_t N, int64_t D,
                                         int64_t H, int64_t W, int64_t Z, int64_t Y, int64_t X, int32_t align_corners,
                                         void* stream_) {
  if (N == 0) return 0;
  LiftGeom g;
  if (int e = make_geom(g, cv_range_host, D, H, W, Z, Y, X, align_corners)) return e;
  SNVC_CHECK_ARG(proj && zs && ys && xs && idx && valid, "null pointer");
  const int64_t nvox = N * Z * Y * X;
  lift_indices_kernel<<<grid_for(nvox), 256, 0, (cudaStream_t)stream_>>>(proj, zs, ys, xs, idx, valid, g, nvox);
  return launch_status("lift_indices_kernel");
}
