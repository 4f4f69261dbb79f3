// N1 -- projection of the instance sampling grid into the left / right ROI frames on the GPU.
//
// Replaces the per-proposal numpy float64 loop of refinementDataset._generate_grid_proj
// (snvc/dataset/KITTIRefinement_dataset.py:847-868: _to_cam :828-845, kitti_util.Calibration.project_rect_to_image
// snvc/dataset/kitti_util.py:282-293, img_proc.affine_transform snvc/utils/img_proc.py:71-74), which projects
// 786 432 grid points x 2 views per proposal on the CPU and ships 12.6 MB of coordinates per proposal to the GPU.
// Here only the pose (5 doubles), two 3x4 projections and two 2x3 affines per proposal cross PCIe; the coordinates
// are produced where the ROI voxel sampling kernel consumes them.
//
// Arithmetic contract: float64 throughout, one cast to float32 at the end, like the reference.  The reference's
// matrix products go through BLAS dgemm (k = 3 or 4 terms); here every dot product is the FMA chain
// acc = a0*b0; acc = fma(a_k, b_k, acc) in k order.  The two can differ in the last float64 bit, which survives
// the float32 cast only at rounding ties: tests/test_gpu_grid_proj.py asserts <= 1 float32 ulp everywhere and
// > 99.99 % bit-identical values against oracle/grid_proj.py (itself bit-exact against the reference's outputs).
// cos / sin of the heading are computed on the host (numpy), not here: device libm differs from glibc in the last ulp.
#include <algorithm>

#include "common.cuh"

namespace snvc {
namespace {

__device__ __forceinline__ double dot3(double a0, double b0, double a1, double b1, double a2, double b2) {
  return fma(a2, b2, fma(a1, b1, a0 * b0));
}

__global__ void __launch_bounds__(256)
roi_grid_project_kernel(const double* __restrict__ pose, const double* __restrict__ Pl, const double* __restrict__ Pr,
                        const double* __restrict__ tl, const double* __restrict__ tr, const double* __restrict__ xp,
                        const double* __restrict__ yp, const double* __restrict__ zp, float* __restrict__ cl,
                        float* __restrict__ cr, float* __restrict__ cam, int nh, int nw, int nl, int64_t P) {
  const int64_t n = blockIdx.y;
  const double c = pose[n * 5 + 0], s = pose[n * 5 + 1];
  const double tx = pose[n * 5 + 2], ty = pose[n * 5 + 3], tz = pose[n * 5 + 4];
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < P; p += (int64_t)gridDim.x * blockDim.x) {
    const int il = (int)(p % nl);
    const int iw = (int)((p / nl) % nw);
    const int ih = (int)(p / ((int64_t)nl * nw));
    const double x = xp[iw], y = yp[ih], z = zp[il];
    // rot_maty @ pts + t   (rows [c 0 s], [0 1 0], [-s 0 c]; the zero terms of the dgemm are exact)
    const double X = dot3(c, x, 0.0, y, s, z) + tx;
    const double Y = dot3(0.0, x, 1.0, y, 0.0, z) + ty;
    const double Z = dot3(-s, x, 0.0, y, c, z) + tz;
    if (cam) {
      cam[(n * P + p) * 3 + 0] = (float)X; cam[(n * P + p) * 3 + 1] = (float)Y; cam[(n * P + p) * 3 + 2] = (float)Z;
    }
#pragma unroll
    for (int v = 0; v < 2; ++v) {
      const double* Pm = (v ? Pr : Pl) + n * 12;
      const double* T = (v ? tr : tl) + n * 6;
      double h[3];
#pragma unroll
      for (int r = 0; r < 3; ++r) h[r] = fma(1.0, Pm[4 * r + 3], dot3(X, Pm[4 * r], Y, Pm[4 * r + 1], Z, Pm[4 * r + 2]));
      const double u = h[0] / h[2], w = h[1] / h[2];
      const double ou = dot3(T[0], u, T[1], w, T[2], 1.0);
      const double ov = dot3(T[3], u, T[4], w, T[5], 1.0);
      float* o = (v ? cr : cl) + n * 2 * P;
      __stcs(o + p, (float)ou);
      __stcs(o + P + p, (float)ov);
    }
  }
}

}  // namespace
}  // namespace snvc

using namespace snvc;

extern "C" int snvc_roi_grid_project(const double* pose, const double* P_left, const double* P_right,
                                     const double* trans_l, const double* trans_r, const double* x_pts,
                                     const double* y_pts, const double* z_pts, float* coord_l, float* coord_r,
                                     float* grid_cam, int64_t N, int64_t nh, int64_t nw, int64_t nl, void* stream) {
  SNVC_CHECK_ARG(N >= 0 && nh > 0 && nw > 0 && nl > 0, "bad dimensions");
  if (N == 0) return 0;
  SNVC_CHECK_ARG(pose && P_left && P_right && trans_l && trans_r && x_pts && y_pts && z_pts && coord_l && coord_r,
                 "null pointer");
  SNVC_CHECK_ARG(N <= 65535 && nh < (1 << 20) && nw < (1 << 20) && nl < (1 << 20), "dimension too large");
  const int64_t P = nh * nw * nl;
  dim3 grid((unsigned)std::max<int64_t>(1, std::min<int64_t>(ceil_div(P, 256), 2048)), (unsigned)N);
  roi_grid_project_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(pose, P_left, P_right, trans_l, trans_r, x_pts, y_pts,
                                                                  z_pts, coord_l, coord_r, grid_cam, (int)nh, (int)nw,
                                                                  (int)nl, P);
  return launch_status("roi_grid_project_kernel");
}
