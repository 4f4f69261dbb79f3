// N4 -- rotated BEV IoU and NMS, entirely on the GPU.
//
// Replaces snvc/extension/iou3d_nms (iou3d_nms_kernel.cu:36-235 geometry, :296-336 nms_kernel, iou3d_nms.cpp:131-177
// nms_gpu).  The pairwise geometry is the reference's algorithm in fp32 (segment intersections + contained corners with
// its 1e-2 margin, ordering around the centroid, fan area) so that keep / suppress decisions agree; what changes is the
// control flow: the reference copies the N x N/64 suppression masks to the HOST and runs the greedy keep loop on the
// CPU (a cudaMalloc, a blocking cudaMemcpy and a cudaFree per call).  Here the keep loop is a second, single-CTA kernel
// -- per 64-box block one thread resolves the diagonal word, then all threads OR the kept rows into the running
// "removed" words in shared memory -- and the kept indices and their count stay on the device; nothing synchronises.
#include <algorithm>

#include "common.cuh"

namespace snvc {
namespace {

constexpr float kEps = 1e-8f;
struct P2 { float x, y; };

__device__ __forceinline__ float cross3(P2 p1, P2 p2, P2 p0) {
  return __fsub_rn(__fmul_rn(__fsub_rn(p1.x, p0.x), __fsub_rn(p2.y, p0.y)), __fmul_rn(__fsub_rn(p2.x, p0.x), __fsub_rn(p1.y, p0.y)));
}
__device__ __forceinline__ bool rect_cross(P2 p1, P2 p2, P2 q1, P2 q2) {
  return fminf(p1.x, p2.x) <= fmaxf(q1.x, q2.x) && fminf(q1.x, q2.x) <= fmaxf(p1.x, p2.x) &&
         fminf(p1.y, p2.y) <= fmaxf(q1.y, q2.y) && fminf(q1.y, q2.y) <= fmaxf(p1.y, p2.y);
}
__device__ __forceinline__ bool in_box(const float* box, P2 p) {
  const float c = cosf(-box[6]), s = sinf(-box[6]);
  const float rx = __fadd_rn(__fmul_rn(__fsub_rn(p.x, box[0]), c), __fmul_rn(__fsub_rn(p.y, box[1]), -s));
  const float ry = __fadd_rn(__fmul_rn(__fsub_rn(p.x, box[0]), s), __fmul_rn(__fsub_rn(p.y, box[1]), c));
  return fabsf(rx) < __fadd_rn(__fdiv_rn(box[3], 2.f), 1e-2f) && fabsf(ry) < __fadd_rn(__fdiv_rn(box[4], 2.f), 1e-2f);
}
__device__ __forceinline__ bool seg_intersection(P2 p1, P2 p0, P2 q1, P2 q0, P2& ans) {
  if (!rect_cross(p0, p1, q0, q1)) return false;
  const float s1 = cross3(q0, p1, p0), s2 = cross3(p1, q1, p0), s3 = cross3(p0, q1, q0), s4 = cross3(q1, p1, q0);
  if (!(__fmul_rn(s1, s2) > 0.f && __fmul_rn(s3, s4) > 0.f)) return false;
  const float s5 = cross3(q1, p1, p0);
  if (fabsf(__fsub_rn(s5, s1)) > kEps) {
    const float d = __fsub_rn(s5, s1);
    ans.x = __fdiv_rn(__fsub_rn(__fmul_rn(s5, q0.x), __fmul_rn(s1, q1.x)), d);
    ans.y = __fdiv_rn(__fsub_rn(__fmul_rn(s5, q0.y), __fmul_rn(s1, q1.y)), d);
  } else {
    const float a0 = __fsub_rn(p0.y, p1.y), b0 = __fsub_rn(p1.x, p0.x), c0 = __fsub_rn(__fmul_rn(p0.x, p1.y), __fmul_rn(p1.x, p0.y));
    const float a1 = __fsub_rn(q0.y, q1.y), b1 = __fsub_rn(q1.x, q0.x), c1 = __fsub_rn(__fmul_rn(q0.x, q1.y), __fmul_rn(q1.x, q0.y));
    const float D = __fsub_rn(__fmul_rn(a0, b1), __fmul_rn(a1, b0));
    ans.x = __fdiv_rn(__fsub_rn(__fmul_rn(b0, c1), __fmul_rn(b1, c0)), D);
    ans.y = __fdiv_rn(__fsub_rn(__fmul_rn(a1, c0), __fmul_rn(a0, c1)), D);
  }
  return true;
}
__device__ __forceinline__ void corners(const float* box, P2* c) {
  const float hx = __fdiv_rn(box[3], 2.f), hy = __fdiv_rn(box[4], 2.f);
  const float px[4] = {__fsub_rn(box[0], hx), __fadd_rn(box[0], hx), __fadd_rn(box[0], hx), __fsub_rn(box[0], hx)};
  const float py[4] = {__fsub_rn(box[1], hy), __fsub_rn(box[1], hy), __fadd_rn(box[1], hy), __fadd_rn(box[1], hy)};
  const float cs = cosf(box[6]), sn = sinf(box[6]);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float dx = __fsub_rn(px[k], box[0]), dy = __fsub_rn(py[k], box[1]);
    c[k].x = __fadd_rn(__fadd_rn(__fmul_rn(dx, cs), __fmul_rn(dy, -sn)), box[0]);
    c[k].y = __fadd_rn(__fadd_rn(__fmul_rn(dx, sn), __fmul_rn(dy, cs)), box[1]);
  }
  c[4] = c[0];
}

__device__ float box_overlap(const float* a, const float* b) {
  P2 ca[5], cb[5], pts[16];
  corners(a, ca);
  corners(b, cb);
  int cnt = 0;
  float cx = 0.f, cy = 0.f;
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) {
      P2 p;
      if (seg_intersection(ca[i + 1], ca[i], cb[j + 1], cb[j], p)) {
        cx = __fadd_rn(cx, p.x); cy = __fadd_rn(cy, p.y);
        pts[cnt++] = p;
      }
    }
  for (int k = 0; k < 4; ++k) {
    if (in_box(a, cb[k])) { cx = __fadd_rn(cx, cb[k].x); cy = __fadd_rn(cy, cb[k].y); pts[cnt++] = cb[k]; }
    if (in_box(b, ca[k])) { cx = __fadd_rn(cx, ca[k].x); cy = __fadd_rn(cy, ca[k].y); pts[cnt++] = ca[k]; }
  }
  if (cnt == 0) return 0.f;
  cx = __fdiv_rn(cx, (float)cnt); cy = __fdiv_rn(cy, (float)cnt);
  float ang[16];
  for (int k = 0; k < cnt; ++k) ang[k] = atan2f(__fsub_rn(pts[k].y, cy), __fsub_rn(pts[k].x, cx));
  for (int j = 0; j < cnt - 1; ++j)                 // the reference's bubble sort (same comparison, angles cached)
    for (int i = 0; i < cnt - j - 1; ++i)
      if (ang[i] > ang[i + 1]) {
        const P2 t = pts[i]; pts[i] = pts[i + 1]; pts[i + 1] = t;
        const float ta = ang[i]; ang[i] = ang[i + 1]; ang[i + 1] = ta;
      }
  float area = 0.f;
  for (int k = 0; k < cnt - 1; ++k) {
    const float ax = __fsub_rn(pts[k].x, pts[0].x), ay = __fsub_rn(pts[k].y, pts[0].y);
    const float bx = __fsub_rn(pts[k + 1].x, pts[0].x), by = __fsub_rn(pts[k + 1].y, pts[0].y);
    area = __fadd_rn(area, __fsub_rn(__fmul_rn(ax, by), __fmul_rn(ay, bx)));
  }
  return __fdiv_rn(fabsf(area), 2.f);
}
__device__ __forceinline__ float iou_bev(const float* a, const float* b) {
  const float sa = __fmul_rn(a[3], a[4]), sb = __fmul_rn(b[3], b[4]);
  const float so = box_overlap(a, b);
  return __fdiv_rn(so, fmaxf(__fsub_rn(__fadd_rn(sa, sb), so), kEps));
}

__global__ void __launch_bounds__(256)
boxes_iou_bev_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ out, int64_t N, int64_t M) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N * M; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t a = i / M, b = i - a * M;
    out[i] = iou_bev(A + a * 7, B + b * 7);
  }
}

// suppression masks, upper triangle only: mask[i][cb] bit t <=> iou(box i, box 64*cb + t) > thresh, 64*cb + t > i.
// One CTA per 64 x 64 block of (row box, column box) pairs, 512 threads: thread (r, s) tests row r against the columns
// s, s + 8, ... and the eight partial words of a row are OR-ed through shared memory (the first version ran the 64
// columns of a row serially in one thread: 230 us of latency for 256 boxes on 10 CTAs of 64 threads).  grid.z = set index
// (independent box sets of N boxes each, e.g. the pairs of a batch).
constexpr int kNmsSplit = 8;
__global__ void __launch_bounds__(64 * kNmsSplit)
nms_mask_kernel(const float* __restrict__ boxes_all, unsigned long long* __restrict__ mask_all, int N, float thresh) {
  const int rb = blockIdx.y, cb = blockIdx.x;
  if (rb > cb) return;
  const int cols = (N + 63) / 64;
  const float* boxes = boxes_all + (int64_t)blockIdx.z * N * 7;
  unsigned long long* mask = mask_all + (int64_t)blockIdx.z * N * cols;
  __shared__ float sb[64 * 7];
  __shared__ unsigned long long part[kNmsSplit][64];
  const int r = threadIdx.x & 63, sp = threadIdx.x >> 6;
  const int col_size = min(N - cb * 64, 64), row_size = min(N - rb * 64, 64);
  for (int k = threadIdx.x; k < col_size * 7; k += blockDim.x) sb[k] = boxes[(int64_t)cb * 64 * 7 + k];
  __syncthreads();
  unsigned long long t = 0;
  if (r < row_size) {
    const float* cur = boxes + (int64_t)(rb * 64 + r) * 7;
    const int j0 = rb == cb ? r + 1 : 0;
    for (int j = sp; j < col_size; j += kNmsSplit)
      if (j >= j0 && iou_bev(cur, sb + j * 7) > thresh) t |= 1ull << j;
  }
  part[sp][r] = t;
  __syncthreads();
  if (sp == 0 && r < row_size) {
#pragma unroll
    for (int k = 1; k < kNmsSplit; ++k) t |= part[k][r];
    mask[(int64_t)(rb * 64 + r) * cols + cb] = t;
  }
}

// greedy keep loop over score-sorted boxes; keep[] receives the kept positions in order, *num_keep their count.
// One CTA per box set (blockIdx.x).
__global__ void __launch_bounds__(256)
nms_keep_kernel(const unsigned long long* __restrict__ mask_all, int64_t* __restrict__ keep_all, int32_t* __restrict__ num_keep, int N) {
  extern __shared__ unsigned long long remv[];
  __shared__ unsigned long long s_keepbits;
  __shared__ int s_count;
  const int cols = (N + 63) / 64;
  const unsigned long long* mask = mask_all + (int64_t)blockIdx.x * N * cols;
  int64_t* keep = keep_all + (int64_t)blockIdx.x * N;
  for (int j = threadIdx.x; j < cols; j += blockDim.x) remv[j] = 0ull;
  if (threadIdx.x == 0) s_count = 0;
  __syncthreads();
  for (int b = 0; b < cols; ++b) {
    if (threadIdx.x == 0) {
      unsigned long long cur = remv[b], kb = 0ull;
      const int n = min(64, N - b * 64);
      int cnt = s_count;
      for (int t = 0; t < n; ++t)
        if (!((cur >> t) & 1ull)) {
          kb |= 1ull << t;
          cur |= mask[(int64_t)(b * 64 + t) * cols + b];
          keep[cnt++] = b * 64 + t;
        }
      s_keepbits = kb;
      s_count = cnt;
    }
    __syncthreads();
    const unsigned long long kb = s_keepbits;
    for (int j = b + 1 + threadIdx.x; j < cols; j += blockDim.x) {
      unsigned long long acc = remv[j];
      for (unsigned long long w = kb; w; w &= w - 1) acc |= mask[(int64_t)(b * 64 + __ffsll((long long)w) - 1) * cols + j];
      remv[j] = acc;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) num_keep[blockIdx.x] = s_count;
}

}  // namespace
}  // namespace snvc

using namespace snvc;

extern "C" int snvc_boxes_iou_bev(const float* boxes_a, const float* boxes_b, float* iou, int64_t N, int64_t M, void* stream) {
  SNVC_CHECK_ARG(N >= 0 && M >= 0, "bad dimensions");
  if (N * M == 0) return 0;
  SNVC_CHECK_ARG(boxes_a && boxes_b && iou, "null pointer");
  const int blocks = (int)std::max<int64_t>(1, std::min<int64_t>(ceil_div(N * M, 256), (int64_t)sm_count() * 8));
  boxes_iou_bev_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(boxes_a, boxes_b, iou, N, M);
  return launch_status("boxes_iou_bev_kernel");
}

extern "C" int64_t snvc_nms_bev_workspace_bytes(int64_t N) { return N * ((N + 63) / 64) * 8; }

extern "C" int snvc_nms_bev_batched(const float* boxes_sorted, void* workspace, int64_t* keep, int32_t* num_keep, int64_t B,
                                    int64_t N, float thresh, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  SNVC_CHECK_ARG(N >= 0 && N <= 60000, "N must be in [0, 60000] (got %lld)", (long long)N);
  SNVC_CHECK_ARG(B >= 0 && B <= 65535, "B must be in [0, 65535]");
  if (B == 0) return 0;
  SNVC_CHECK_ARG(num_keep != nullptr, "null pointer");
  if (N == 0) { SNVC_CUDA_OK(cudaMemsetAsync(num_keep, 0, sizeof(int32_t) * B, stream)); return 0; }
  SNVC_CHECK_ARG(boxes_sorted && workspace && keep, "null pointer");
  const int cols = (int)((N + 63) / 64);
  nms_mask_kernel<<<dim3(cols, cols, (unsigned)B), 64 * kNmsSplit, 0, stream>>>(boxes_sorted, (unsigned long long*)workspace, (int)N, thresh);
  if (int e = launch_status("nms_mask_kernel")) return e;
  nms_keep_kernel<<<(unsigned)B, 256, cols * sizeof(unsigned long long), stream>>>((const unsigned long long*)workspace, keep, num_keep, (int)N);
  return launch_status("nms_keep_kernel");
}

extern "C" int snvc_nms_bev(const float* boxes_sorted, void* workspace, int64_t* keep, int32_t* num_keep, int64_t N,
                            float thresh, void* stream_) {
  return snvc_nms_bev_batched(boxes_sorted, workspace, keep, num_keep, 1, N, thresh, stream_);
}
