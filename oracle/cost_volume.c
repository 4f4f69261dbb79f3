/*
 * ORACLE (test infrastructure, NOT product code).
 *
 * Scalar CPU restatement of the reference's plane-sweep cost-volume operator.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load this file's shared object.  The product path
 * (snvc_b200/) never links or calls it.
 *
 * Follows, statement by statement:
 *   forward  : snvc/extension/build_cost_volume/src/BuildCostVolume_cuda.cu:63-98
 *   bilinear : snvc/extension/build_cost_volume/src/BuildCostVolume_cuda.cu:15-61
 *   backward : snvc/extension/build_cost_volume/src/BuildCostVolume_cuda.cu:152-205
 *              (weights: :101-150)
 *   shapes   : snvc/extension/build_cost_volume/src/BuildCostVolume_cuda.cu:222-229
 *
 * Parity status: the reference ships no test vectors for this op and has no CPU
 * path (BuildCostVolume.cpp:26), so this restatement is pinned only by (i) the
 * hand-computed known-answer cases in tests/test_oracle_cost_volume.py and
 * (ii) an independent vectorised numpy restatement (oracle/cost_volume.py).
 * See DESIGN.md "parity pinning".
 *
 * `fma_mode`: the reference is compiled by nvcc with the default -fmad=true, so
 * `w1*v1 + w2*v2 + w3*v3 + w4*v4` becomes  fma(w4,v4, fma(w3,v3, fma(w1,v1, w2*v2))): the SECOND product is the
 * rounded one (SASS of the reference kernel as built by oracle/build_ref.py: FMUL v2*w2; FFMA v1,w1; FFMA v3,w3;
 * FFMA v4,w4 -- same order in the fp64 instantiation).  fma_mode=1 reproduces that contraction and is checked bit
 * for bit against the reference op on the GPU (tests/test_gpu_ref_pin.py);
 * fma_mode=0 rounds every product and sum separately (what a plain C/NumPy
 * transcription gives; differs by <= 1 ulp).
 *
 * 64-bit indexing throughout (the reference's int32 indexing overflows above
 * 2^31 output elements).
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

#define GEN_BILINEAR(T, SUF, FMA)                                                        \
  static T bilinear_##SUF(const T* bottom, int64_t height, int64_t width, T y, T x,      \
                          int fma_mode) {                                                 \
    /* .cu:21-24 */                                                                       \
    if (y < (T)-1.0 || y > (T)height || x < (T)-1.0 || x > (T)width) return (T)0;         \
    /* .cu:26-27 */                                                                       \
    if (y <= 0) y = 0;                                                                    \
    if (x <= 0) x = 0;                                                                    \
    /* .cu:29-46 */                                                                       \
    int64_t y_low = (int64_t)y, x_low = (int64_t)x, y_high, x_high;                       \
    if (y_low >= height - 1) { y_high = y_low = height - 1; y = (T)y_low; }               \
    else { y_high = y_low + 1; }                                                          \
    if (x_low >= width - 1) { x_high = x_low = width - 1; x = (T)x_low; }                 \
    else { x_high = x_low + 1; }                                                          \
    /* .cu:48-56 */                                                                       \
    volatile T ly = y - (T)y_low;                                                         \
    volatile T lx = x - (T)x_low;                                                         \
    volatile T hy = (T)1. - ly, hx = (T)1. - lx;                                          \
    T v1 = bottom[y_low * width + x_low];                                                 \
    T v2 = bottom[y_low * width + x_high];                                                \
    T v3 = bottom[y_high * width + x_low];                                                \
    T v4 = bottom[y_high * width + x_high];                                               \
    volatile T w1 = hy * hx, w2 = hy * lx, w3 = ly * hx, w4 = ly * lx;                    \
    /* .cu:58 */                                                                          \
    if (fma_mode) {                                                                       \
      volatile T t = w2 * v2;                                                             \
      t = FMA(w1, v1, t);                                                                 \
      t = FMA(w3, v3, t);                                                                 \
      t = FMA(w4, v4, t);                                                                 \
      return t;                                                                           \
    } else {                                                                              \
      volatile T p1 = w1 * v1, p2 = w2 * v2, p3 = w3 * v3, p4 = w4 * v4;                  \
      volatile T s = p1 + p2;                                                             \
      s = s + p3;                                                                         \
      s = s + p4;                                                                         \
      return s;                                                                           \
    }                                                                                     \
  }

GEN_BILINEAR(float, f32, fmaf)
GEN_BILINEAR(double, f64, fma)

/* cost: [N, 2C, D, H, W] with H = IH/ds, W = IW/ds (.cu:224-228). */
#define GEN_FORWARD(T, SUF)                                                               \
  void oracle_cost_volume_fwd_##SUF(const T* left, const T* right, const T* shift,        \
                                    T* cost, int64_t N, int64_t C, int64_t IH,            \
                                    int64_t IW, int64_t D, int64_t ds, int fma_mode) {    \
    const int64_t H = IH / ds, W = IW / ds;                                               \
    /* the reference uses height*downsample as image size (.cu:78-79) */                  \
    const int64_t img_h = H * ds, img_w = W * ds;                                         \
    (void)IH; (void)IW;                                                                   \
    for (int64_t n = 0; n < N; ++n)                                                       \
      for (int64_t c = 0; c < C; ++c)                                                     \
        for (int64_t pd = 0; pd < D; ++pd) {                                              \
          const T shift_pd = -shift[n * D + pd]; /* .cu:84 */                             \
          for (int64_t ph = 0; ph < H; ++ph)                                              \
            for (int64_t pw = 0; pw < W; ++pw) {                                          \
              const int64_t iw = pw * ds, ih = ph * ds;                                   \
              const int64_t index_L = (((n * (2 * C) + c) * D + pd) * H + ph) * W + pw;   \
              const int64_t index_R = index_L + C * D * H * W;                            \
              /* NB: the reference indexes the *input* with img_h/img_w strides */        \
              cost[index_L] = left[((n * C + c) * img_h + ih) * img_w + iw];              \
              volatile T xs = (T)iw + shift_pd; /* .cu:88 */                              \
              if (xs >= (T)0. && xs <= (T)(img_w - 1)) {                                  \
                const T* off = right + (n * C + c) * img_h * img_w;                       \
                cost[index_R] = bilinear_##SUF(off, img_h, img_w, (T)ih, xs, fma_mode);   \
              } else {                                                                    \
                cost[index_R] = (T)0.;                                                    \
              }                                                                           \
            }                                                                             \
        }                                                                                 \
  }

GEN_FORWARD(float, f32)
GEN_FORWARD(double, f64)

/*
 * Backward (.cu:152-205).  The reference accumulates with atomicAdd in an
 * unspecified order; this restatement accumulates in the kernel's linear index
 * order (n, c, pd, ph, pw), in the element type.  Tests therefore compare with
 * a tolerance, or against the f64 instantiation.
 * grad: [N, 2C, D, H, W];  grad_left/right: [N, C, H*ds, W*ds], zero-filled here
 * (at::zeros, .cu:270-271).
 */
#define GEN_BACKWARD(T, SUF)                                                              \
  void oracle_cost_volume_bwd_##SUF(const T* grad, const T* shift, T* grad_left,          \
                                    T* grad_right, int64_t N, int64_t C, int64_t H,       \
                                    int64_t W, int64_t D, int64_t ds) {                   \
    const int64_t img_h = H * ds, img_w = W * ds;                                         \
    memset(grad_left, 0, sizeof(T) * (size_t)(N * C * img_h * img_w));                    \
    memset(grad_right, 0, sizeof(T) * (size_t)(N * C * img_h * img_w));                   \
    for (int64_t n = 0; n < N; ++n)                                                       \
      for (int64_t c = 0; c < C; ++c)                                                     \
        for (int64_t pd = 0; pd < D; ++pd) {                                              \
          const T shift_pd = -shift[n * D + pd];                                          \
          for (int64_t ph = 0; ph < H; ++ph)                                              \
            for (int64_t pw = 0; pw < W; ++pw) {                                          \
              const int64_t iw = pw * ds, ih = ph * ds;                                   \
              const int64_t index_L = (((n * 2 * C + c) * D + pd) * H + ph) * W + pw;     \
              const int64_t index_R = index_L + C * D * H * W;                            \
              grad_left[((n * C + c) * img_h + ih) * img_w + iw] += grad[index_L];        \
              volatile T x = (T)iw + shift_pd;                                            \
              if (!(x >= (T)0. && x <= (T)(img_w - 1))) continue;                         \
              /* bilinear_interpolate_gradient, .cu:101-150 (guards unreachable) */       \
              T y = (T)ih;                                                                \
              if (y <= 0) y = 0;                                                          \
              if (x <= 0) x = 0;                                                          \
              int64_t y_low = (int64_t)y, x_low = (int64_t)x, y_high, x_high;             \
              if (y_low >= img_h - 1) { y_high = y_low = img_h - 1; y = (T)y_low; }       \
              else y_high = y_low + 1;                                                    \
              if (x_low >= img_w - 1) { x_high = x_low = img_w - 1; x = (T)x_low; }       \
              else x_high = x_low + 1;                                                    \
              volatile T ly = y - (T)y_low, lx = x - (T)x_low;                            \
              volatile T hy = (T)1. - ly, hx = (T)1. - lx;                                \
              volatile T w1 = hy * hx, w2 = hy * lx, w3 = ly * hx, w4 = ly * lx;          \
              const T g = grad[index_R];                                                  \
              T* gr = grad_right + (n * C + c) * img_h * img_w;                           \
              if (w1 >= (T)1e-10) gr[y_low * img_w + x_low] += g * w1;                    \
              if (w2 >= (T)1e-10) gr[y_low * img_w + x_high] += g * w2;                   \
              if (w3 >= (T)1e-10) gr[y_high * img_w + x_low] += g * w3;                   \
              if (w4 >= (T)1e-10) gr[y_high * img_w + x_high] += g * w4;                  \
            }                                                                             \
        }                                                                                 \
  }

GEN_BACKWARD(float, f32)
GEN_BACKWARD(double, f64)

/* Index/validity dump used by the bit-exact index parity tests:
 * for every (n, pd, pw): x_low (or -1 when the sample is outside the image). */
void oracle_cost_volume_xlow_f32(const float* shift, int32_t* xlow, int64_t N, int64_t IW,
                                 int64_t D, int64_t ds) {
  const int64_t W = IW / ds, img_w = W * ds;
  for (int64_t n = 0; n < N; ++n)
    for (int64_t pd = 0; pd < D; ++pd) {
      const float s = -shift[n * D + pd];
      for (int64_t pw = 0; pw < W; ++pw) {
        volatile float x = (float)(pw * ds) + s;
        int32_t r = -1;
        if (x >= 0.f && x <= (float)(img_w - 1)) {
          float xx = x;
          if (xx <= 0) xx = 0;
          int64_t xl = (int64_t)xx;
          if (xl >= img_w - 1) xl = img_w - 1;
          r = (int32_t)xl;
        }
        xlow[(n * D + pd) * W + pw] = r;
      }
    }
}
