// N2 / A5 -- the depth head that follows the 3-D trunk.
//
//   snvc_disparity_regression     `disparityregression.forward` (snvc/models/submodule.py:76-83):
//                                 out[n,h,w] = sum_k prob[n,k,h,w] * depth[k]
//   snvc_depth_regression_fwd     the whole tail the DSGN-lineage global branch puts in front of it (restated wiring,
//                                 SURVEY.md 3.4): F.interpolate(logits [N,1,D,H,W] -> [Dout,Hout,Wout], 'trilinear',
//                                 align_corners) -> softmax over depth -> disparityregression, FUSED: the up-sampled
//                                 probability volume (368 MB per pair at 192 x 384 x 1248 fp32, written and read twice
//                                 by the reference's three ops) is never materialised.
// Up-sampling follows ATen UpSample.h `area_pixel_compute_source_index` / `compute_source_index_and_lambda`
// (index = (int)src, lambda1 = src - index, lambda0 = 1 - lambda1, upper index clamped) in fp32, and the trilinear
// value is composed as in UpSampleTrilinear3d: d0l*(h0l*(w0l*x + w1l*x) + h1l*(..)) + d1l*(..).
#include <algorithm>

#include "common.cuh"

namespace snvc {
namespace {

struct Axis { int i0, i1; float l0, l1; };

__device__ __forceinline__ Axis source_index(int dst, int in_size, float scale, bool align_corners) {
  float src;
  if (align_corners) {
    src = scale * (float)dst;
  } else {
    src = scale * ((float)dst + 0.5f) - 0.5f;
    src = src < 0.f ? 0.f : src;
  }
  Axis a;
  a.i0 = min((int)src, in_size - 1);
  a.i1 = a.i0 + (a.i0 < in_size - 1 ? 1 : 0);
  a.l1 = fminf(fmaxf(src - (float)a.i0, 0.f), 1.f);
  a.l0 = 1.f - a.l1;
  return a;
}

// one thread per output pixel; the D bilinearly up-sampled logits of the pixel live in shared memory [d][thread]
__global__ void __launch_bounds__(128)
depth_regression_kernel(const float* __restrict__ logits, const float* __restrict__ depth, float* __restrict__ out, int D,
                        int H, int W, int Dout, int Hout, int Wout, float sd, float sh, float sw, int align_corners,
                        int64_t total /* N*Hout*Wout */) {
  extern __shared__ float sm[];
  const int tid = threadIdx.x, nt = blockDim.x;
  const int64_t pix = (int64_t)blockIdx.x * nt + tid;
  const bool live = pix < total;
  const int64_t HWo = (int64_t)Hout * Wout;
  const int64_t n = live ? pix / HWo : 0;
  const int64_t r = live ? pix - n * HWo : 0;
  const int oh = (int)(r / Wout), ow = (int)(r - (int64_t)oh * Wout);
  const Axis ah = source_index(oh, H, sh, align_corners), aw = source_index(ow, W, sw, align_corners);
  const float* base = logits + n * (int64_t)D * H * W;
  const int64_t HW = (int64_t)H * W;
  const int64_t o00 = (int64_t)ah.i0 * W + aw.i0, o01 = (int64_t)ah.i0 * W + aw.i1, o10 = (int64_t)ah.i1 * W + aw.i0,
                o11 = (int64_t)ah.i1 * W + aw.i1;
  float m = -INFINITY;
  for (int d = 0; d < D; ++d) {
    const float* pl = base + d * HW;
    const float v = ah.l0 * (aw.l0 * __ldg(pl + o00) + aw.l1 * __ldg(pl + o01)) +
                    ah.l1 * (aw.l0 * __ldg(pl + o10) + aw.l1 * __ldg(pl + o11));
    sm[d * nt + tid] = v;
    m = fmaxf(m, v);                       // an upper bound of every interpolated logit: a valid softmax shift
  }
  float s = 0.f, acc = 0.f;
  for (int k = 0; k < Dout; ++k) {
    const Axis ad = source_index(k, D, sd, align_corners);
    const float v = ad.l0 * sm[ad.i0 * nt + tid] + ad.l1 * sm[ad.i1 * nt + tid];
    const float e = exp2f((v - m) * 1.4426950408889634f);
    s += e;
    acc = fmaf(e, __ldg(depth + k), acc);
  }
  if (live) out[pix] = acc / s;
}

// out[n,h,w] = sum_k prob[n,k,h,w] * depth[k]; 4 pixels per thread (float4), k-loop streams the volume once
__global__ void __launch_bounds__(256)
disparity_regression_kernel(const float* __restrict__ prob, const float* __restrict__ depth, float* __restrict__ out,
                            int K, int64_t HW, int64_t total4 /* N*HW/4 */) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t hw4 = HW / 4;
    const int64_t n = i / hw4, q = i - n * hw4;
    const float4* p = reinterpret_cast<const float4*>(prob + n * K * HW) + q;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int k = 0; k < K; ++k) {
      const float4 v = __ldcs(p + (int64_t)k * hw4);
      const float dk = __ldg(depth + k);
      a.x += v.x * dk; a.y += v.y * dk; a.z += v.z * dk; a.w += v.w * dk;      // torch.sum order: k ascending
    }
    reinterpret_cast<float4*>(out + n * HW)[q] = a;
  }
}

// any H*W, any alignment: one pixel per thread (the reference's torch.sum accepts every shape, submodule.py:82-83)
__global__ void __launch_bounds__(256)
disparity_regression_scalar_kernel(const float* __restrict__ prob, const float* __restrict__ depth, float* __restrict__ out,
                                   int K, int64_t HW, int64_t total /* N*HW */) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t n = i / HW, q = i - n * HW;
    const float* p = prob + n * K * HW + q;
    float a = 0.f;
    for (int k = 0; k < K; ++k) a += __ldcs(p + (int64_t)k * HW) * __ldg(depth + k);
    out[i] = a;
  }
}

}  // namespace
}  // namespace snvc

using namespace snvc;

static float scale_of(int64_t in, int64_t out, int align_corners) {
  // ATen area_pixel_compute_scale<float>
  if (align_corners) return out > 1 ? (float)(in - 1) / (float)(out - 1) : 0.f;
  return (float)in / (float)out;
}

extern "C" int snvc_depth_regression_fwd(const float* logits, const float* depth_values, float* out, int64_t N, int64_t D,
                                         int64_t H, int64_t W, int64_t Dout, int64_t Hout, int64_t Wout,
                                         int32_t align_corners, void* stream) {
  SNVC_CHECK_ARG(N >= 0 && D > 0 && H > 0 && W > 0 && Dout > 0 && Hout > 0 && Wout > 0, "bad dimensions");
  if (N == 0) return 0;
  SNVC_CHECK_ARG(logits && depth_values && out, "null pointer");
  SNVC_CHECK_ARG(D <= 400 && Dout < (1 << 20) && H < (1 << 20) && W < (1 << 20), "dimension too large (D <= 400)");
  const int64_t total = N * Hout * Wout;
  const size_t smem = (size_t)D * 128 * sizeof(float);
  SNVC_CUDA_OK(cudaFuncSetAttribute(depth_regression_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int64_t blocks = ceil_div(total, 128);
  SNVC_CHECK_ARG(blocks < (1ll << 31), "too many output pixels");
  depth_regression_kernel<<<(unsigned)blocks, 128, smem, (cudaStream_t)stream>>>(
      logits, depth_values, out, (int)D, (int)H, (int)W, (int)Dout, (int)Hout, (int)Wout, scale_of(D, Dout, align_corners),
      scale_of(H, Hout, align_corners), scale_of(W, Wout, align_corners), align_corners ? 1 : 0, total);
  return launch_status("depth_regression_kernel");
}

extern "C" int snvc_disparity_regression(const float* prob, const float* depth_values, float* out, int64_t N, int64_t K,
                                         int64_t HW, void* stream) {
  SNVC_CHECK_ARG(N >= 0 && K > 0 && HW > 0, "bad dimensions");
  if (N == 0) return 0;
  SNVC_CHECK_ARG(prob && depth_values && out, "null pointer");
  if (HW % 4 != 0 || (reinterpret_cast<uintptr_t>(prob) & 15) != 0 || (reinterpret_cast<uintptr_t>(out) & 15) != 0) {
    const int64_t total = N * HW;
    const int blocks = (int)std::max<int64_t>(1, std::min<int64_t>(ceil_div(total, 256), (int64_t)sm_count() * 8));
    disparity_regression_scalar_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(prob, depth_values, out, (int)K, HW, total);
    return launch_status("disparity_regression_scalar_kernel");
  }
  const int64_t total4 = N * HW / 4;
  const int blocks = (int)std::max<int64_t>(1, std::min<int64_t>(ceil_div(total4, 256), (int64_t)sm_count() * 8));
  disparity_regression_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(prob, depth_values, out, (int)K, HW, total4);
  return launch_status("disparity_regression_kernel");
}
