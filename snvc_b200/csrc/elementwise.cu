// Layout conversions and the small elementwise steps between the 3-D convolutions.
#include "common.cuh"

namespace snvc {
namespace {

// [N,C,S] fp32 -> [N,S,C] bf16 through a 32x33 shared tile (coalesced both ways)
__global__ void __launch_bounds__(256)
ncdhw_to_ndhwc_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int C, int64_t S) {
  __shared__ float tile[32][33];
  const int64_t n = blockIdx.z;
  const int64_t s0 = (int64_t)blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
  for (int j = ty; j < 32; j += 8) {
    int c = c0 + j;
    int64_t s = s0 + tx;
    tile[j][tx] = (c < C && s < S) ? src[(n * C + c) * S + s] : 0.f;
  }
  __syncthreads();
  for (int j = ty; j < 32; j += 8) {
    int64_t s = s0 + j;
    int c = c0 + tx;
    if (s < S && c < C) dst[(n * S + s) * C + c] = __float2bfloat16_rn(tile[tx][j]);
  }
}

__global__ void __launch_bounds__(256)
ndhwc_to_ncdhw_kernel(const __nv_bfloat16* __restrict__ src, float* __restrict__ dst, int C, int64_t S) {
  __shared__ float tile[32][33];
  const int64_t n = blockIdx.z;
  const int64_t s0 = (int64_t)blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int j = ty; j < 32; j += 8) {
    int64_t s = s0 + j;
    int c = c0 + tx;
    tile[j][tx] = (s < S && c < C) ? __bfloat162float(src[(n * S + s) * C + c]) : 0.f;
  }
  __syncthreads();
  for (int j = ty; j < 32; j += 8) {
    int c = c0 + j;
    int64_t s = s0 + tx;
    if (c < C && s < S) dst[(n * C + c) * S + s] = tile[tx][j];
  }
}

// dst[ns, coffset + c] = bf16( vimg[ns, c] * occ[ns] ), 8 channels per thread
__global__ void __launch_bounds__(256)
scale_occ_kernel(const __nv_bfloat16* __restrict__ vimg, const float* __restrict__ occ, __nv_bfloat16* __restrict__ dst,
                 int64_t total /* NS*C/8 */, int C, int cstride, int coffset) {
  const int CG = C >> 3;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int cg = (int)(i % CG);
    int64_t ns = i / CG;
    float o = __ldg(occ + ns);
    uint4 q = __ldg(reinterpret_cast<const uint4*>(vimg + ns * C + cg * 8));
    uint4 r = {pack_bf16x2(bf16_lo(q.x) * o, bf16_hi(q.x) * o), pack_bf16x2(bf16_lo(q.y) * o, bf16_hi(q.y) * o),
               pack_bf16x2(bf16_lo(q.z) * o, bf16_hi(q.z) * o), pack_bf16x2(bf16_lo(q.w) * o, bf16_hi(q.w) * o)};
    *reinterpret_cast<uint4*>(dst + ns * cstride + coffset + cg * 8) = r;
  }
}

// x [N,Dh,H,W,C] bf16 -> bev [N, C*(Dh/pool), H, W] fp32; channel = c*(Dh/pool) + dh  (vernier.py:436-438)
__global__ void __launch_bounds__(256)
avgpool_bev_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ bev, int64_t total, int Dh, int64_t HW, int C,
                   int pool) {
  const int Dp = Dh / pool;
  const float inv = 1.f / (float)pool;
  // one thread per (n, c, dp, hw); hw fastest -> coalesced fp32 stores, reads hit 2-byte strided (L2 resident)
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t hw = i % HW;
    int64_t t = i / HW;
    int dp = (int)(t % Dp);
    t /= Dp;
    int c = (int)(t % C);
    int64_t n = t / C;
    float a = 0.f;
    for (int k = 0; k < pool; ++k)
      a += __bfloat162float(x[((n * Dh + (int64_t)dp * pool + k) * HW + hw) * C + c]);
    bev[i] = a * inv;
  }
}

// x [N,S0,S1,S2,C] bf16 -> bev [N, R0, R2, C*Q] bf16 (NHWC for the 2-D tensor-core convs), Q = S_axis / pool,
// channel index c*Q + q (the reference's reshape of the pooled NCDHW tensor: vernier.py:436-438).
// One thread per (n, r0, r2, 8-channel group): `pool` 16-byte loads per q, Q results per channel kept in registers and
// written as one run of Q consecutive bf16 per channel.
template <int AXIS, int Q>
__global__ void __launch_bounds__(256)
avgpool_bev_nhwc_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ bev, int64_t total, int S0, int S1,
                        int S2, int C, int pool) {
  const int CG = C >> 3;
  const int R0 = AXIS == 0 ? S1 : S0;
  const float inv = 1.f / (float)pool;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int cg = (int)(i % CG);
    int64_t t = i / CG;
    const int r2 = (int)(t % S2); t /= S2;
    const int r0 = (int)(t % R0);
    const int64_t n = t / R0;
    float acc[Q][8];
#pragma unroll
    for (int q = 0; q < Q; ++q) {
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[q][j] = 0.f;
      for (int k = 0; k < pool; ++k) {
        const int a = q * pool + k;
        const int64_t vox = AXIS == 0 ? (((n * S0 + a) * S1 + r0) * S2 + r2) : (((n * S0 + r0) * S1 + a) * S2 + r2);
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(x + vox * C + cg * 8));
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) { acc[q][2 * j] += bf16_lo(w[j]); acc[q][2 * j + 1] += bf16_hi(w[j]); }
      }
    }
    // this thread's 8 channels x Q windows are 8*Q consecutive output channels (index j*Q + q): Q 16-byte stores
    uint32_t words[4 * Q];
#pragma unroll
    for (int e = 0; e < 4 * Q; ++e) {
      const int e0 = 2 * e, e1 = 2 * e + 1;
      words[e] = pack_bf16x2(acc[e0 % Q][e0 / Q] * inv, acc[e1 % Q][e1 / Q] * inv);
    }
    uint4* o = reinterpret_cast<uint4*>(bev + ((n * R0 + r0) * S2 + r2) * ((int64_t)C * Q) + (int64_t)cg * 8 * Q);
#pragma unroll
    for (int e = 0; e < Q; ++e) o[e] = make_uint4(words[4 * e], words[4 * e + 1], words[4 * e + 2], words[4 * e + 3]);
  }
}

template <int AXIS>
int launch_avgpool_nhwc(int Q, int blocks, cudaStream_t st, const __nv_bfloat16* x, __nv_bfloat16* bev, int64_t total, int S0,
                        int S1, int S2, int C, int pool) {
  switch (Q) {
#define SNVC_POOL_CASE(q) case q: avgpool_bev_nhwc_kernel<AXIS, q><<<blocks, 256, 0, st>>>(x, bev, total, S0, S1, S2, C, pool); break;
    SNVC_POOL_CASE(1) SNVC_POOL_CASE(2) SNVC_POOL_CASE(3) SNVC_POOL_CASE(4) SNVC_POOL_CASE(5) SNVC_POOL_CASE(6) SNVC_POOL_CASE(7)
    SNVC_POOL_CASE(8)
#undef SNVC_POOL_CASE
    default: return fail(SNVC_E_UNSUPPORTED, "at most 8 pooling windows");
  }
  return launch_status("avgpool_bev_nhwc_kernel<>");
}

}  // namespace
}  // namespace snvc

using namespace snvc;

extern "C" int snvc_avgpool_to_bev_nhwc(const void* x, void* bev, int64_t N, int64_t S0, int64_t S1, int64_t S2, int32_t C,
                                        int32_t pool, int32_t axis, void* stream) {
  if (N * S0 * S1 * S2 == 0) return 0;
  SNVC_CHECK_ARG(x && bev, "null pointer");
  SNVC_CHECK_ARG(axis == 0 || axis == 1, "axis must be 0 or 1");
  const int64_t Sax = axis == 0 ? S0 : S1;
  SNVC_CHECK_ARG(pool >= 1 && Sax % pool == 0 && Sax / pool <= 8, "pooled axis must be a multiple of pool with at most 8 windows");
  SNVC_CHECK_ARG(C % 8 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0, "C must be a multiple of 8 and x 16-byte aligned");
  const int64_t total = N * (axis == 0 ? S1 : S0) * S2 * (C / 8);
  const int blocks = (int)std::min<int64_t>(ceil_div(total, 256), (int64_t)sm_count() * 16);
  SNVC_CHECK_ARG((reinterpret_cast<uintptr_t>(bev) & 15) == 0, "bev must be 16-byte aligned");
  const int Q = (int)(Sax / pool);
  if (axis == 0)
    return launch_avgpool_nhwc<0>(Q, blocks, (cudaStream_t)stream, (const __nv_bfloat16*)x, (__nv_bfloat16*)bev, total, (int)S0, (int)S1,
                                  (int)S2, C, pool);
  return launch_avgpool_nhwc<1>(Q, blocks, (cudaStream_t)stream, (const __nv_bfloat16*)x, (__nv_bfloat16*)bev, total, (int)S0, (int)S1,
                                (int)S2, C, pool);
}

extern "C" int snvc_ncdhw_f32_to_ndhwc_bf16(const float* src, void* dst, int64_t N, int64_t C, int64_t S, void* stream) {
  if (N * C * S == 0) return 0;
  SNVC_CHECK_ARG(src && dst, "null pointer");
  SNVC_CHECK_ARG(N <= 65535 && C <= 65535 * 32ll, "N or C too large");
  dim3 grid((unsigned)ceil_div(S, 32), (unsigned)ceil_div(C, 32), (unsigned)N);
  ncdhw_to_ndhwc_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(src, (__nv_bfloat16*)dst, (int)C, S);
  return launch_status("ncdhw_to_ndhwc_kernel");
}

extern "C" int snvc_ndhwc_bf16_to_ncdhw_f32(const void* src, float* dst, int64_t N, int64_t C, int64_t S, void* stream) {
  if (N * C * S == 0) return 0;
  SNVC_CHECK_ARG(src && dst, "null pointer");
  SNVC_CHECK_ARG(N <= 65535 && C <= 65535 * 32ll, "N or C too large");
  dim3 grid((unsigned)ceil_div(S, 32), (unsigned)ceil_div(C, 32), (unsigned)N);
  ndhwc_to_ncdhw_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)src, dst, (int)C, S);
  return launch_status("ndhwc_to_ncdhw_kernel");
}

extern "C" int snvc_scale_by_occupancy(const void* vimg, const float* occ, void* dst, int64_t NS, int32_t C,
                                       int32_t cstride, int32_t coffset, void* stream) {
  if (NS == 0) return 0;
  SNVC_CHECK_ARG(vimg && occ && dst, "null pointer");
  SNVC_CHECK_ARG(C % 8 == 0 && cstride % 8 == 0 && coffset % 8 == 0 && coffset + C <= cstride, "bad channel geometry");
  const int64_t total = NS * (C / 8);
  int blocks = (int)std::min<int64_t>(ceil_div(total, 256), (int64_t)sm_count() * 16);
  scale_occ_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)vimg, occ, (__nv_bfloat16*)dst, total,
                                                             C, cstride, coffset);
  return launch_status("scale_occ_kernel");
}

extern "C" int snvc_avgpool_to_bev(const void* x, float* bev, int64_t N, int64_t Dh, int64_t H, int64_t W, int32_t C,
                                   int32_t pool, void* stream) {
  if (N * Dh * H * W == 0) return 0;
  SNVC_CHECK_ARG(x && bev, "null pointer");
  SNVC_CHECK_ARG(pool >= 1 && Dh % pool == 0, "Dh must be a multiple of pool");
  const int64_t total = N * C * (Dh / pool) * H * W;
  int blocks = (int)std::min<int64_t>(ceil_div(total, 256), (int64_t)sm_count() * 16);
  avgpool_bev_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x, bev, total, (int)Dh, H * W, C, pool);
  return launch_status("avgpool_bev_kernel");
}
