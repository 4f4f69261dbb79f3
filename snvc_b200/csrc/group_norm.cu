// GroupNorm (+ skip add, ReLU, sigmoid) on channels-last activations -- the `gn=True` variant of convbn_3d / convbn
// (nn.GroupNorm(32, C) after the bias-free convolution: snvc/models/submodule.py:28,49,135,146,195,207,220).
//
// Unlike eval-mode BatchNorm, GroupNorm cannot be folded into the convolution's epilogue: its statistics are taken over
// the whole (S x C/G) extent of every sample.  The convolution therefore writes its fp32 result once and this file does
//   1. gn_stats_kernel     per-(sample, block, channel) partial (sum, sum of squares), fp32 lanes -> fp64 partials,
//                          no atomics (deterministic);
//   2. gn_finalize_kernel  partials -> per-(sample, channel) affine  a = rstd * gamma,  b = beta - mean * rstd * gamma;
//   3. gn_apply_kernel     y = act(a * x + b [+ residual]) [+ residual], written bf16 (or fp32) into a channel slice.
// All three are HBM-bound streaming passes (x is read twice, 4 B per element; y written once).
#include <algorithm>

#include "common.cuh"

namespace snvc {
namespace {

constexpr int kGnThreads = 256;

// x [N, S, C] fp32, C <= 256 and C % 4 == 0.  grid (B, N): block b reduces rows [b*rows_per_block, ...).
// A thread owns the float4 channel group (tid % (C/4)) and strides over rows.
__global__ void __launch_bounds__(kGnThreads)
gn_stats_kernel(const float* __restrict__ x, double* __restrict__ partial, int64_t S, int C, int64_t rows_per_block) {
  const int CG = C >> 2;
  const int lanes = kGnThreads / CG;                 // rows processed concurrently by the block
  const int cg = threadIdx.x % CG, rl = threadIdx.x / CG;
  const int64_t n = blockIdx.y;
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_block, r1 = min(S, r0 + rows_per_block);
  float s[4] = {0.f, 0.f, 0.f, 0.f}, q[4] = {0.f, 0.f, 0.f, 0.f};
  double ds[4] = {0, 0, 0, 0}, dq[4] = {0, 0, 0, 0};
  int cnt = 0;
  if (rl < lanes) {
    for (int64_t r = r0 + rl; r < r1; r += lanes) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(x + (n * S + r) * C) + cg);
      s[0] += v.x; s[1] += v.y; s[2] += v.z; s[3] += v.w;
      q[0] = fmaf(v.x, v.x, q[0]); q[1] = fmaf(v.y, v.y, q[1]); q[2] = fmaf(v.z, v.z, q[2]); q[3] = fmaf(v.w, v.w, q[3]);
      if (++cnt == 64) {                               // flush the fp32 partials into fp64 every 64 rows
#pragma unroll
        for (int j = 0; j < 4; ++j) { ds[j] += s[j]; dq[j] += q[j]; s[j] = 0.f; q[j] = 0.f; }
        cnt = 0;
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) { ds[j] += s[j]; dq[j] += q[j]; }
  __shared__ double sh[kGnThreads][8];
#pragma unroll
  for (int j = 0; j < 4; ++j) { sh[threadIdx.x][j] = ds[j]; sh[threadIdx.x][4 + j] = dq[j]; }
  __syncthreads();
  if (threadIdx.x < CG) {                              // fixed-order reduction over the row lanes
    double a[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int l = 0; l < lanes; ++l)
#pragma unroll
      for (int j = 0; j < 8; ++j) a[j] += sh[l * CG + threadIdx.x][j];
    double* o = partial + ((n * gridDim.x + blockIdx.x) * C + threadIdx.x * 4) * 2;
#pragma unroll
    for (int j = 0; j < 4; ++j) { o[2 * j] = a[j]; o[2 * j + 1] = a[4 + j]; }
  }
}

// one block per sample: sums the partials per channel, then per group; writes ab[n, c] = (a, b)
__global__ void __launch_bounds__(256)
gn_finalize_kernel(const double* __restrict__ partial, const float* __restrict__ gamma, const float* __restrict__ beta,
                   float2* __restrict__ ab, int B, int C, int G, int64_t S, float eps) {
  __shared__ double cs[256], cq[256];
  const int64_t n = blockIdx.x;
  const int c = threadIdx.x;
  if (c < C) {
    double s = 0, q = 0;
    for (int b = 0; b < B; ++b) {
      const double* p = partial + ((n * B + b) * C + c) * 2;
      s += p[0]; q += p[1];
    }
    cs[c] = s; cq[c] = q;
  }
  __syncthreads();
  if (c < C) {
    const int cpg = C / G, g0 = (c / cpg) * cpg;
    double s = 0, q = 0;
    for (int j = 0; j < cpg; ++j) { s += cs[g0 + j]; q += cq[g0 + j]; }
    const double cnt = (double)S * cpg;
    const double mean = s / cnt;
    const double var = fmax(q / cnt - mean * mean, 0.0);
    const float rstd = (float)(1.0 / sqrt(var + (double)eps));
    const float a = rstd * (gamma ? gamma[c] : 1.f);
    ab[n * C + c] = make_float2(a, (beta ? beta[c] : 0.f) - (float)mean * a);
  }
}

struct GnApply {
  int C, relu, residual_mode, sigmoid, out_f32, out_cstride, out_coffset, res_cstride, res_coffset;
};

// one thread per (row, 8-channel group)
__global__ void __launch_bounds__(256)
gn_apply_kernel(const float* __restrict__ x, const float2* __restrict__ ab, const __nv_bfloat16* __restrict__ residual,
                void* __restrict__ y, int64_t total /* N*S*C/8 */, int64_t S, GnApply p) {
  const int CG = p.C >> 3;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int cg = (int)(i % CG);
    const int64_t row = i / CG;
    const int64_t n = row / S;
    const float4 v0 = __ldcs(reinterpret_cast<const float4*>(x + row * p.C + cg * 8));
    const float4 v1 = __ldcs(reinterpret_cast<const float4*>(x + row * p.C + cg * 8) + 1);
    float v[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
    float r[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (p.residual_mode) {
      const uint4 q = __ldg(reinterpret_cast<const uint4*>(residual + row * p.res_cstride + p.res_coffset + cg * 8));
      const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) { r[2 * j] = bf16_lo(w[j]); r[2 * j + 1] = bf16_hi(w[j]); }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float2 c = __ldg(ab + n * p.C + cg * 8 + j);
      float t = fmaf(v[j], c.x, c.y);
      if (p.residual_mode == 1) t += r[j];
      if (p.relu) t = fmaxf(t, 0.f);
      if (p.residual_mode == 2) t += r[j];
      if (p.sigmoid) t = 1.f / (1.f + __expf(-t));
      v[j] = t;
    }
    if (p.out_f32) {
      float* o = reinterpret_cast<float*>(y) + row * p.out_cstride + p.out_coffset + cg * 8;
      reinterpret_cast<float4*>(o)[0] = make_float4(v[0], v[1], v[2], v[3]);
      reinterpret_cast<float4*>(o)[1] = make_float4(v[4], v[5], v[6], v[7]);
    } else {
      __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(y) + row * p.out_cstride + p.out_coffset + cg * 8;
      *reinterpret_cast<uint4*>(o) = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]),
                                                 pack_bf16x2(v[6], v[7]));
    }
  }
}

int gn_blocks(int64_t S) { return (int)std::max<int64_t>(1, std::min<int64_t>(ceil_div(S, 512), 4ll * sm_count())); }

}  // namespace
}  // namespace snvc

using namespace snvc;

extern "C" int64_t snvc_group_norm_workspace_bytes(int64_t N, int64_t S, int32_t C) {
  return N * gn_blocks(S) * (int64_t)C * 2 * 8 + N * (int64_t)C * 8;
}

extern "C" int snvc_group_norm_fwd(const float* x, const float* gamma, const float* beta, const void* residual, void* y,
                                   void* workspace, int64_t N, int64_t S, int32_t C, int32_t groups, float eps, int32_t relu,
                                   int32_t residual_mode, int32_t sigmoid, int32_t out_dtype, int32_t out_cstride,
                                   int32_t out_coffset, int32_t res_cstride, int32_t res_coffset, void* stream_) {
  if (N * S == 0) return 0;
  cudaStream_t stream = (cudaStream_t)stream_;
  SNVC_CHECK_ARG(x && y && workspace, "null pointer");
  SNVC_CHECK_ARG(C >= 8 && C <= 256 && C % 8 == 0, "C must be a multiple of 8 in [8, 256] (got %d)", C);
  SNVC_CHECK_ARG(groups >= 1 && C % groups == 0, "C must be a multiple of groups");
  SNVC_CHECK_ARG(out_dtype == SNVC_BF16 || out_dtype == SNVC_F32, "out_dtype must be bf16 or f32");
  SNVC_CHECK_ARG(residual_mode == 0 || residual != nullptr, "residual_mode set but residual is null");
  const int ocs = out_cstride ? out_cstride : C, rcs = res_cstride ? res_cstride : C;
  SNVC_CHECK_ARG(ocs % 8 == 0 && out_coffset % 8 == 0 && out_coffset + C <= ocs, "bad output channel slice");
  SNVC_CHECK_ARG(!residual_mode || (rcs % 8 == 0 && res_coffset % 8 == 0 && res_coffset + C <= rcs), "bad residual channel slice");
  SNVC_CHECK_ARG(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(workspace) |
                   reinterpret_cast<uintptr_t>(residual)) & 15) == 0, "pointers must be 16-byte aligned");
  SNVC_CHECK_ARG(N <= 65535, "N too large");
  const int B = gn_blocks(S);
  double* partial = static_cast<double*>(workspace);
  float2* ab = reinterpret_cast<float2*>(partial + N * B * (int64_t)C * 2);
  gn_stats_kernel<<<dim3(B, (unsigned)N), kGnThreads, 0, stream>>>(x, partial, S, C, ceil_div(S, B));
  if (int e = launch_status("gn_stats_kernel")) return e;
  gn_finalize_kernel<<<(unsigned)N, 256, 0, stream>>>(partial, gamma, beta, ab, B, C, groups, S, eps);
  if (int e = launch_status("gn_finalize_kernel")) return e;
  GnApply p{C, relu, residual_mode, sigmoid, out_dtype == SNVC_F32, ocs, out_coffset, rcs, res_coffset};
  const int64_t total = N * S * (C / 8);
  const int blocks = (int)std::min<int64_t>(ceil_div(total, 256), (int64_t)sm_count() * 16);
  gn_apply_kernel<<<blocks, 256, 0, stream>>>(x, ab, (const __nv_bfloat16*)residual, y, total, S, p);
  return launch_status("gn_apply_kernel");
}
