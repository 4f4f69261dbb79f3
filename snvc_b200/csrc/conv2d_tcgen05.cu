// N2 -- 2-D convolution / transposed convolution for the BEV tails, as a tcgen05 + TMEM implicit GEMM (sm_100a).
//
// Replaces the cuDNN calls behind nn.Conv2d / nn.ConvTranspose2d (+ eval BatchNorm2d, ReLU, skip adds) of
//   convbn                      snvc/models/submodule.py:11-29
//   hourglass2d                 snvc/models/submodule.py:317-361
//   hourglass2d_downsample_16   snvc/models/submodule.py:270-315 (helpers :183-195, :210-221)
//   conv5 / hm1 / hm2           snvc/models/vernier.py:289-314, 440-445
// and the restated RPN-side BEV convolutions of the global branch.
//
// GEMM view, per filter tap t = (kh,kw) and 64-channel chunk c of Cin:
//     D[128 pixels, Cout] += A_{t,c}[128 pixels, 64] * W_{t,c}[64, Cout]
//   * activations are NHWC bf16; one 4-D TMA box (64 ch x TW x TH x 1) per (tap, chunk) lands the 128 x 64 operand tile
//     in shared memory in the canonical K-major SWIZZLE_128B UMMA layout; TMA out-of-bounds zero fill is the padding
//     (and the channel tail when Cin is not a multiple of 64), TMA elementStrides is the stride;
//   * the K loop runs over taps x chunks, which lifts the Cin <= 64 limit of the 3-D kernels (the BEV tensors have
//     Cin = 32 * nh/4 = 256 channels, vernier.py:290-295);
//   * Cout up to 256 in ONE MMA (N = Cout, M = 128, cta_group::1), two TMEM accumulators so the epilogue of tile i
//     overlaps the MMAs of tile i+1; epilogue = folded BN scale/bias, skip add, ReLU, sigmoid, bf16 or fp32 NHWC rows;
//   * ConvTranspose2d(k3,s2,p1,op1) = its 4 output-parity classes (1,2,2,4 taps), all in one launch: a tile belongs to
//     a class, classes differ in their tap table and output offset.
// Warp roles: warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2..5 = epilogue.  Persistent grid.
// These layers are 13 of the instance branch's 1708 GFLOP per proposal: the kernel is built for coverage (any Cin,
// Cout <= 256, stride 1/2, transposed) rather than for the last percent of the tensor pipe.
#include <algorithm>
#include <cstdlib>

#include "common.cuh"
#include "tcgen05.cuh"

namespace snvc {
namespace {

constexpr int k2Threads = 192;
constexpr int k2MaxStages = 8;
constexpr int k2TileM = 128;
constexpr int k2MaxTaps = 9;
constexpr int k2MaxCout = 256;

struct Conv2dClass {
  int ntaps, oh, ow;                     // output coord = j*out_stride + (oh, ow)
  signed char dy[k2MaxTaps], dx[k2MaxTaps], widx[k2MaxTaps];   // input coord = j*in_stride + (dy, dx); weight tap slot
};

struct Conv2dParams {
  int N, Cin, Cout, CoutPad;
  int Ho, Wo;                            // output tensor extent
  int Hj, Wj;                            // iteration space of one class
  int TH, TW, tiles_h, tiles_w, tiles_per_class, num_tiles;
  int in_stride, out_stride;
  int nclasses, nchunks, chunk_c;        // Cin is walked in nchunks pieces of chunk_c channels
  Conv2dClass cls[4];
  int stages, a_bytes, b_bytes, swizzle_bytes;
  int relu, residual_mode, sigmoid, out_f32;
  int out_cstride, out_coffset, res_cstride, res_coffset;
  const float* scale;
  const float* bias;
  const __nv_bfloat16* residual;
  void* y;
};

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

struct Tile2d { int cls, n, jh, jw; };
__device__ __forceinline__ Tile2d decode_tile2d(const Conv2dParams& p, int tile) {
  Tile2d t;
  t.cls = tile / p.tiles_per_class;
  int r = tile - t.cls * p.tiles_per_class;
  const int tw = r % p.tiles_w; r /= p.tiles_w;
  const int th = r % p.tiles_h; r /= p.tiles_h;
  t.n = r; t.jh = th * p.TH; t.jw = tw * p.TW;
  return t;
}

__global__ void __launch_bounds__(k2Threads, 1)
conv2d_tcgen05_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w,
                      const __grid_constant__ Conv2dParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t full_bar[k2MaxStages];
  __shared__ __align__(8) uint64_t empty_bar[k2MaxStages];
  __shared__ __align__(8) uint64_t tmem_full_bar[2];
  __shared__ __align__(8) uint64_t tmem_empty_bar[2];
  __shared__ uint32_t tmem_base_smem;
  __shared__ __align__(16) float s_scale[k2MaxCout], s_bias[k2MaxCout];

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t smem_base = (smem_u32(smem) + 1023u) & ~1023u;
  const int stage_bytes = p.a_bytes + p.b_bytes;
  uint32_t tmem_cols = 32;
  while ((int)tmem_cols < 2 * p.CoutPad) tmem_cols <<= 1;

  for (int i = threadIdx.x; i < k2MaxCout; i += k2Threads) {
    s_scale[i] = (p.scale && i < p.Cout) ? p.scale[i] : 1.f;
    s_bias[i] = (p.bias && i < p.Cout) ? p.bias[i] : 0.f;
  }
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_x) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(smem_u32(&full_bar[s]), 1);
      mbar_init(smem_u32(&empty_bar[s]), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(smem_u32(&tmem_full_bar[b]), 1);
      mbar_init(smem_u32(&tmem_empty_bar[b]), 4);   // one arrive per epilogue warp
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
    // ===================== TMA producer =====================
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      const Tile2d tc = decode_tile2d(p, tile);
      const Conv2dClass& k = p.cls[tc.cls];
      const int ch = tc.jh * p.in_stride, cw = tc.jw * p.in_stride;
      for (int t = 0; t < k.ntaps; ++t)
        for (int c = 0; c < p.nchunks; ++c) {
          mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1u);
          if (elect_one()) {
            const uint32_t fb = smem_u32(&full_bar[stage]);
            const uint32_t sa = smem_base + stage * stage_bytes;
            mbar_expect_tx(fb, (uint32_t)stage_bytes);
            tma_load_4d(sa, &map_x, fb, c * p.chunk_c, cw + k.dx[t], ch + k.dy[t], tc.n);
            tma_load_2d(sa + p.a_bytes, &map_w, fb, 0, (k.widx[t] * p.nchunks + c) * p.CoutPad);
          }
          __syncwarp();
          if (++stage == p.stages) { stage = 0; phase ^= 1u; }
        }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    const uint32_t idesc = make_idesc(k2TileM, p.CoutPad);
    const uint32_t desc_hi = (uint32_t)(make_smem_desc(0, p.swizzle_bytes) >> 32);
    const int ksteps = p.chunk_c >> 4;
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
      const int buf = it & 1;
      const uint32_t use = (uint32_t)(it >> 1);
      const int nsteps = p.cls[tile / p.tiles_per_class].ntaps * p.nchunks;
      mbar_wait(smem_u32(&tmem_empty_bar[buf]), (use & 1u) ^ 1u);
      tcgen05_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(buf * p.CoutPad);
      for (int t = 0; t < nsteps; ++t) {
        mbar_wait(smem_u32(&full_bar[stage]), phase);
        tcgen05_fence_after();
        const uint32_t sa = smem_base + stage * stage_bytes;
        const uint32_t a_lo = ((sa >> 4) & 0x3FFFu) | (1u << 16);
        const uint32_t b_lo = (((sa + p.a_bytes) >> 4) & 0x3FFFu) | (1u << 16);
        if (elect_one()) {
          for (int k = 0; k < ksteps; ++k)
            umma_bf16(d_tmem, desc64(desc_hi, a_lo + 2 * k), desc64(desc_hi, b_lo + 2 * k), idesc, (t | k) ? 1u : 0u);
          umma_commit(smem_u32(&empty_bar[stage]));
          if (t == nsteps - 1) umma_commit(smem_u32(&tmem_full_bar[buf]));
        }
        __syncwarp();
        if (++stage == p.stages) { stage = 0; phase ^= 1u; }
      }
    }
  } else {
    // ===================== epilogue (warps 2..5) =====================
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const int r_w = row % p.TW, r_h = row / p.TW;
    const bool vec_ok = ((p.out_cstride | p.out_coffset) & 7) == 0 && (!p.residual_mode || ((p.res_cstride | p.res_coffset) & 7) == 0);
    int it = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
      const int buf = it & 1;
      const uint32_t use = (uint32_t)(it >> 1);
      const Tile2d tc = decode_tile2d(p, tile);
      const Conv2dClass& k = p.cls[tc.cls];
      const int jh = tc.jh + r_h, jw = tc.jw + r_w;
      const bool in_range = jh < p.Hj && jw < p.Wj;
      const int oh = jh * p.out_stride + k.oh, ow = jw * p.out_stride + k.ow;
      const int64_t pix = ((int64_t)tc.n * p.Ho + oh) * p.Wo + ow;
      mbar_wait(smem_u32(&tmem_full_bar[buf]), use & 1u);
      tcgen05_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(buf * p.CoutPad);
      for (int c0 = 0; c0 < p.CoutPad; c0 += 16) {
        uint32_t acc[16];
        tmem_ld16(taddr + (uint32_t)c0, acc);
        tmem_ld_wait();
        if (!in_range) continue;
        float v[16];
        const bool full = c0 + 16 <= p.Cout;
        uint32_t rw[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
        if (p.residual_mode) {
          const __nv_bfloat16* rp = p.residual + pix * p.res_cstride + p.res_coffset + c0;
          if (full && vec_ok) {
            const uint4 q0 = __ldg(reinterpret_cast<const uint4*>(rp)), q1 = __ldg(reinterpret_cast<const uint4*>(rp) + 1);
            rw[0] = q0.x; rw[1] = q0.y; rw[2] = q0.z; rw[3] = q0.w; rw[4] = q1.x; rw[5] = q1.y; rw[6] = q1.z; rw[7] = q1.w;
          } else {
            for (int j = 0; j < 16; ++j)
              if (c0 + j < p.Cout) {
                const uint32_t b = (uint32_t)__bfloat16_as_ushort(rp[j]);
                rw[j >> 1] |= (j & 1) ? (b << 16) : b;
              }
          }
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          float x = fmaf(__uint_as_float(acc[j]), s_scale[c0 + j], s_bias[c0 + j]);
          const float r = (j & 1) ? bf16_hi(rw[j >> 1]) : bf16_lo(rw[j >> 1]);
          if (p.residual_mode == 1) x += r;
          if (p.relu) x = fmaxf(x, 0.f);
          if (p.residual_mode == 2) x += r;
          if (p.sigmoid) x = 1.f / (1.f + __expf(-x));
          v[j] = x;
        }
        if (p.out_f32) {
          float* yp = reinterpret_cast<float*>(p.y) + pix * p.out_cstride + p.out_coffset + c0;
          if (full && ((p.out_cstride | p.out_coffset) & 3) == 0) {
#pragma unroll
            for (int q = 0; q < 4; ++q)
              reinterpret_cast<float4*>(yp)[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
          } else {
            for (int j = 0; j < 16; ++j)
              if (c0 + j < p.Cout) yp[j] = v[j];
          }
        } else {
          __nv_bfloat16* yp = reinterpret_cast<__nv_bfloat16*>(p.y) + pix * p.out_cstride + p.out_coffset + c0;
          if (full && vec_ok) {
            reinterpret_cast<uint4*>(yp)[0] = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]),
                                                         pack_bf16x2(v[6], v[7]));
            reinterpret_cast<uint4*>(yp)[1] = make_uint4(pack_bf16x2(v[8], v[9]), pack_bf16x2(v[10], v[11]),
                                                         pack_bf16x2(v[12], v[13]), pack_bf16x2(v[14], v[15]));
          } else {
            for (int j = 0; j < 16; ++j)
              if (c0 + j < p.Cout) yp[j] = __float2bfloat16_rn(v[j]);
          }
        }
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&tmem_empty_bar[buf]));
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
  }
}

// w fp32: conv [Cout,Cin,kh,kw] / deconv [Cin,Cout,kh,kw]  ->  packed bf16 [kh*kw][nchunks][CoutPad][chunk_c]
__global__ void pack_weights2d_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out, int Cin, int Cout,
                                      int CoutPad, int taps, int nchunks, int chunk_c, int transposed, int64_t total) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int cc = (int)(i % chunk_c);
    int64_t t = i / chunk_c;
    const int co = (int)(t % CoutPad); t /= CoutPad;
    const int ch = (int)(t % nchunks);
    const int tap = (int)(t / nchunks);
    const int ci = ch * chunk_c + cc;
    float v = 0.f;
    if (co < Cout && ci < Cin)
      v = transposed ? w[((int64_t)ci * Cout + co) * taps + tap] : w[((int64_t)co * Cin + ci) * taps + tap];
    out[i] = __float2bfloat16_rn(v);
  }
}

int chunk_of(int Cin) { return Cin >= 64 ? 64 : (Cin >= 32 ? 32 : 16); }

void pick_tile2d(int Hj, int Wj, int& TH, int& TW) {
  double best = -1;
  for (int tw = 128; tw >= 1; tw >>= 1) {
    const int th = 128 / tw;
    const double padded = (double)round_up(Hj, th) * round_up(Wj, tw);
    const double score = (double)Hj * Wj / padded + (tw >= 8 ? 1e-3 : 0) + 1e-6 * tw;
    if (score > best) { best = score; TH = th; TW = tw; }
  }
}

}  // namespace
}  // namespace snvc

using namespace snvc;

extern "C" int64_t snvc_conv2d_packed_weight_bytes(int32_t Cin, int32_t Cout, int32_t kernel) {
  const int cc = chunk_of(Cin);
  return (int64_t)kernel * kernel * ceil_div(Cin, cc) * round_up(Cout, 16) * cc * 2;
}

extern "C" int snvc_conv2d_pack_weights(const float* w, void* w_packed, int32_t Cin, int32_t Cout, int32_t kernel,
                                        int32_t transposed, void* stream) {
  SNVC_CHECK_ARG(w && w_packed, "null pointer");
  SNVC_CHECK_ARG(Cin > 0 && Cout > 0 && (kernel == 1 || kernel == 3), "bad dimensions (kernel must be 1 or 3)");
  const int cc = chunk_of(Cin), nch = (int)ceil_div(Cin, cc), CoutPad = round_up(Cout, 16), taps = kernel * kernel;
  const int64_t total = (int64_t)taps * nch * CoutPad * cc;
  const int blocks = (int)std::min<int64_t>(ceil_div(total, 256), 1024);
  pack_weights2d_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(w, (__nv_bfloat16*)w_packed, Cin, Cout, CoutPad, taps, nch, cc,
                                                                  transposed, total);
  return launch_status("pack_weights2d_kernel");
}

extern "C" int snvc_conv2d_fwd(const void* x, const void* w_packed, const float* scale, const float* bias, const void* residual,
                               void* y, const snvc_conv2d_desc* dp, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  SNVC_CHECK_ARG(dp != nullptr, "desc is null");
  const snvc_conv2d_desc& d = *dp;
  SNVC_CHECK_ARG(x && w_packed && y, "null pointer");
  SNVC_CHECK_ARG(d.Cin >= 16 && d.Cin % 8 == 0, "Cin must be a multiple of 8 and >= 16 (got %d)", d.Cin);
  SNVC_CHECK_ARG(d.Cin % 16 == 0 || d.Cin > 64, "Cin below 64 must be 16, 32 or 48");
  SNVC_CHECK_ARG(d.Cout >= 1 && d.Cout <= k2MaxCout, "Cout must be in [1, 256] (got %d)", d.Cout);
  SNVC_CHECK_ARG(d.N >= 0 && d.Hi > 0 && d.Wi > 0, "bad input extent");
  SNVC_CHECK_ARG(d.kernel == 1 || d.kernel == 3, "kernel must be 1 or 3");
  SNVC_CHECK_ARG(d.out_dtype == SNVC_BF16 || d.out_dtype == SNVC_F32, "out_dtype must be bf16 or f32");
  SNVC_CHECK_ARG(d.residual_mode == 0 || residual != nullptr, "residual_mode set but residual is null");
  SNVC_CHECK_ARG(d.in_cstride == 0 || (d.in_cstride % 8 == 0 && d.in_coffset % 8 == 0 && d.in_coffset + d.Cin <= d.in_cstride),
                 "bad input channel slice (in_cstride %d, in_coffset %d)", d.in_cstride, d.in_coffset);
  SNVC_CHECK_ARG((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(w_packed) & 15) == 0 &&
                     (reinterpret_cast<uintptr_t>(y) & 15) == 0 && (reinterpret_cast<uintptr_t>(residual) & 15) == 0,
                 "x, w_packed, y and residual must be 16-byte aligned");
  if (d.N == 0) return 0;
  EncodeTiledFn enc = encode_fn();
  if (!enc) return fail(SNVC_E_DRIVER, "cuTensorMapEncodeTiled is not available from the CUDA driver");

  Conv2dParams p{};
  p.N = d.N; p.Cin = d.Cin; p.Cout = d.Cout; p.CoutPad = round_up(d.Cout, 16);
  p.Ho = d.Ho; p.Wo = d.Wo;
  p.relu = d.relu; p.residual_mode = d.residual_mode; p.sigmoid = d.sigmoid; p.out_f32 = d.out_dtype == SNVC_F32;
  p.out_cstride = d.out_cstride ? d.out_cstride : d.Cout;
  p.out_coffset = d.out_coffset;
  p.res_cstride = d.res_cstride ? d.res_cstride : d.Cout;
  p.res_coffset = d.res_coffset;
  SNVC_CHECK_ARG(p.out_coffset + d.Cout <= p.out_cstride, "output channel slice out of range");
  SNVC_CHECK_ARG(!d.residual_mode || p.res_coffset + d.Cout <= p.res_cstride, "residual channel slice out of range");
  p.scale = scale; p.bias = bias; p.residual = (const __nv_bfloat16*)residual; p.y = y;
  p.chunk_c = chunk_of(d.Cin);
  p.nchunks = (int)ceil_div(d.Cin, p.chunk_c);
  p.swizzle_bytes = p.chunk_c * 2;

  if (!d.transposed) {
    SNVC_CHECK_ARG(d.stride == 1 || d.stride == 2, "stride must be 1 or 2");
    SNVC_CHECK_ARG(d.dilation >= 1 && d.pad >= 0 && d.dilation * (d.kernel - 1) <= 127 && d.pad <= 127, "bad dilation / pad");
    const int ext = d.dilation * (d.kernel - 1) + 1;
    SNVC_CHECK_ARG(d.Ho == (d.Hi + 2 * d.pad - ext) / d.stride + 1 && d.Wo == (d.Wi + 2 * d.pad - ext) / d.stride + 1,
                   "output extent does not match conv geometry");
    p.Hj = d.Ho; p.Wj = d.Wo; p.in_stride = d.stride; p.out_stride = 1; p.nclasses = 1;
    Conv2dClass& c = p.cls[0];
    c.oh = c.ow = 0; c.ntaps = 0;
    for (int kh = 0; kh < d.kernel; ++kh)
      for (int kw = 0; kw < d.kernel; ++kw) {
        c.dy[c.ntaps] = (signed char)(kh * d.dilation - d.pad);
        c.dx[c.ntaps] = (signed char)(kw * d.dilation - d.pad);
        c.widx[c.ntaps] = (signed char)(kh * d.kernel + kw);
        ++c.ntaps;
      }
  } else {
    // ConvTranspose2d(k=3, s=2, p=1, output_padding=1): out[o] += x[i] * W[k], o = 2i - 1 + k
    //   even o = 2j: k = 1, i = j;   odd o = 2j + 1: k = 2, i = j  and  k = 0, i = j + 1
    SNVC_CHECK_ARG(d.kernel == 3 && d.stride == 2 && d.pad == 1 && d.dilation == 1,
                   "transposed conv supports k=3, s=2, p=1, output_padding=1 only");
    SNVC_CHECK_ARG(d.Ho == 2 * d.Hi && d.Wo == 2 * d.Wi, "transposed conv output must be 2x input");
    p.Hj = d.Hi; p.Wj = d.Wi; p.in_stride = 1; p.out_stride = 2; p.nclasses = 4;
    for (int cls = 0; cls < 4; ++cls) {
      const int ph = cls >> 1, pw = cls & 1;
      Conv2dClass& c = p.cls[cls];
      c.oh = ph; c.ow = pw; c.ntaps = 0;
      const int nh = ph ? 2 : 1, nw = pw ? 2 : 1;
      const int offh[2] = {0, 1}, kh_[2] = {ph ? 2 : 1, 0};
      const int offw[2] = {0, 1}, kw_[2] = {pw ? 2 : 1, 0};
      for (int a = 0; a < nh; ++a)
        for (int b = 0; b < nw; ++b) {
          c.dy[c.ntaps] = (signed char)offh[a];
          c.dx[c.ntaps] = (signed char)offw[b];
          c.widx[c.ntaps] = (signed char)(kh_[a] * 3 + kw_[b]);
          ++c.ntaps;
        }
    }
  }
  pick_tile2d(p.Hj, p.Wj, p.TH, p.TW);
  p.tiles_h = (int)ceil_div(p.Hj, p.TH); p.tiles_w = (int)ceil_div(p.Wj, p.TW);
  const int64_t tpc = (int64_t)p.N * p.tiles_h * p.tiles_w;
  SNVC_CHECK_ARG(tpc * p.nclasses < (1ll << 31), "too many tiles");
  p.tiles_per_class = (int)tpc;
  p.num_tiles = (int)(tpc * p.nclasses);
  if (p.num_tiles == 0) return 0;
  p.a_bytes = k2TileM * p.chunk_c * 2;
  p.b_bytes = p.CoutPad * p.chunk_c * 2;
  const int stage_bytes = p.a_bytes + p.b_bytes;
  p.stages = std::max(2, std::min(k2MaxStages, (196 * 1024) / stage_bytes));
  const size_t smem = (size_t)p.stages * stage_bytes + 1024;

  CUtensorMap map_x, map_w;
  {
    const cuuint64_t cs = (cuuint64_t)(d.in_cstride ? d.in_cstride : d.Cin) * 2;   // bytes between pixels
    cuuint64_t dims[4] = {(cuuint64_t)d.Cin, (cuuint64_t)d.Wi, (cuuint64_t)d.Hi, (cuuint64_t)d.N};
    cuuint64_t strides[3] = {cs, (cuuint64_t)d.Wi * cs, (cuuint64_t)d.Hi * d.Wi * cs};
    const int s = p.in_stride;
    cuuint32_t box[4] = {(cuuint32_t)p.chunk_c, (cuuint32_t)((p.TW - 1) * s + 1), (cuuint32_t)((p.TH - 1) * s + 1), 1};
    cuuint32_t estr[4] = {1, (cuuint32_t)s, (cuuint32_t)s, 1};
    const void* xbase = static_cast<const char*>(x) + (size_t)d.in_coffset * 2;
    CUresult r = enc(&map_x, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(xbase), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_mode(p.swizzle_bytes), CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail((int)r, "cuTensorMapEncodeTiled(x, 2-D conv) failed with CUresult %d", (int)r);
  }
  {
    const int taps = d.kernel * d.kernel;
    cuuint64_t dims[2] = {(cuuint64_t)p.chunk_c, (cuuint64_t)taps * p.nchunks * p.CoutPad};
    cuuint64_t strides[1] = {(cuuint64_t)p.chunk_c * 2};
    cuuint32_t box[2] = {(cuuint32_t)p.chunk_c, (cuuint32_t)p.CoutPad};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&map_w, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(w_packed), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_mode(p.swizzle_bytes), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail((int)r, "cuTensorMapEncodeTiled(w, 2-D conv) failed with CUresult %d", (int)r);
  }
  SNVC_CUDA_OK(cudaFuncSetAttribute(conv2d_tcgen05_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int grid = std::min(p.num_tiles, sm_count());
  conv2d_tcgen05_kernel<<<grid, k2Threads, smem, stream>>>(map_x, map_w, p);
  return launch_status("conv2d_tcgen05_kernel");
}
