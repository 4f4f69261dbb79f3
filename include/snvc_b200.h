/*
 * snvc_b200 -- C ABI of the B200-native (sm_100a) dense stereo-to-voxel hot path.
 *
 * This is the drop-in boundary for the path named in BASELINE.json:north_star.  Each entry
 * point below replaces one native / library call of the reference (Nicholasli1995/SNVC); the
 * reference interface it replaces is cited as file:line into the reference tree.
 * INTEGRATION.md shows the reference-side (ctypes) binding.
 *
 * Conventions (all entry points)
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless its name ends in
 *     `_host`; all buffers (outputs, workspaces) are allocated by the caller;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream); work is enqueued
 *     asynchronously, the library never synchronises and never allocates device memory;
 *   - return value: 0 = ok, < 0 = argument error (SNVC_E_*), > 0 = a cudaError_t / CUresult;
 *     snvc_last_error() returns a thread-local human-readable message for the last non-zero return;
 *   - 64-bit element offsets throughout (the reference kernel's int32 indexing overflows above
 *     2^31 outputs, BuildCostVolume_cuda.cu:69-82);
 *   - thread-safe / re-entrant: no mutable global state (the reference is called from one Python
 *     thread per device under nn.DataParallel, tools/inference_agnostic.py:472).
 */
#ifndef SNVC_B200_H_
#define SNVC_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SNVC_ABI_VERSION 1

/* element types */
#define SNVC_F32 0
#define SNVC_BF16 1
#define SNVC_F64 2

/* volume layouts: NCDHW is the reference's (torch contiguous); NDHWC (channels-last) is what
 * the tcgen05 conv3d consumes */
#define SNVC_NCDHW 0
#define SNVC_NDHWC 1

/* argument errors */
#define SNVC_E_BADARG (-1)
#define SNVC_E_UNSUPPORTED (-2)
#define SNVC_E_NOTCUDA (-3)
#define SNVC_E_DRIVER (-4)

int snvc_version(void);
const char* snvc_last_error(void);
/* Number of CUDA kernels this library has launched in this process so far (monotonic counter; the
 * only process-wide state besides cached device attributes).  bench.py reports the difference over
 * its timed region as `gpu_launches`. */
int64_t snvc_launch_count(void);
/* Debug / A-B switches (kernel-generation selection, grid clamps for the ring wrap-around tests): SNVC_CONV_MODE,
 * SNVC_CONV_STORE, SNVC_CONV_OCC, SNVC_CONV_MAXGRID, SNVC_CV_SPLIT_OLD, SNVC_CV_THREADS, SNVC_CV_WALK, SNVC_ROI_MODE, SNVC_LIFT_MODE.  Each is read from
 * the environment once, when the library is loaded; snvc_set_option changes one afterwards (value NULL or "" = unset;
 * name NULL = unset all).  Not synchronised with concurrent launches: set options before starting work.  Launches never
 * call getenv. */
int snvc_set_option(const char* name, const char* value);
const char* snvc_get_option(const char* name);

/* ---------------------------------------------------------------------------------------------
 * A1  plane-sweep cost volume, forward.
 * Replaces build_cost_volume_forward(left, right, shift, downsample)
 *   snvc/extension/build_cost_volume/src/BuildCostVolume.cpp:13-27,46
 *   snvc/extension/build_cost_volume/src/BuildCostVolume_cuda.cu:208-256 (host), :63-98 (kernel)
 * left,right : [N, C, IH, IW] contiguous, element type `dtype` (SNVC_F32 | SNVC_F64)
 * shift      : [N, D] same element type, >= 0 (checked by the Python layer, __init__.py:12)
 * cost       : out_layout NCDHW -> [N, 2C, D, IH/ds, IW/ds];  NDHWC -> [N, D, IH/ds, IW/ds, 2C]
 * supported (dtype, out_dtype, out_layout): (F32,F32,NCDHW) (F64,F64,NCDHW) (F32,BF16,NDHWC)
 * IH, IW must be multiples of `downsample` (the reference mis-indexes otherwise, .cu:78-86).
 */
int snvc_cost_volume_fwd(const void* left, const void* right, const void* shift, void* cost,
                         int64_t N, int64_t C, int64_t IH, int64_t IW, int64_t D, int32_t downsample,
                         int32_t dtype, int32_t out_dtype, int32_t out_layout, void* stream);

/* A1, split form for the tcgen05 trunk (f32 in, bf16 NDHWC out).  The left half of the cost volume is a pure
 * broadcast over depth (BuildCostVolume_cuda.cu:84-86: cost[n,c,d,ph,pw] = left[n,c,ih,iw]), so it is written ONCE:
 * right_vol   : [N, D, IH/ds, IW/ds, C]  = channels [C, 2C) of the NDHWC volume above (the shifted right features)
 * left_planes : [N, 3, IH/ds, IW/ds, C]  = the left features, channels-last bf16, on three identical planes -- the
 *               input of the 3-plane convolution that yields the depth-invariant addend of snvc_conv3d_fwd_addend.
 * Together they carry exactly the information of the [N,D,H,W,2C] volume; half the bytes are written.
 * Either output pointer may be NULL: the halves are independent launches, so a caller can enqueue them on different
 * streams (GlobalHotPath overlaps left planes -> addend convolution with the right-half build). */
int snvc_cost_volume_split_fwd(const void* left, const void* right, const void* shift, void* right_vol,
                               void* left_planes, int64_t N, int64_t C, int64_t IH, int64_t IW, int64_t D,
                               int32_t downsample, void* stream);

/* A1b  cost volume, backward (deterministic, atomics-free).
 * Replaces build_cost_volume_backward(grad, shift, downsample)
 *   BuildCostVolume.cpp:29-43,47;  BuildCostVolume_cuda.cu:259-303 (host), :152-205 (kernel)
 * grad : [N, 2C, D, H, W] (NCDHW);  grad_left, grad_right : [N, C, H*ds, W*ds] (fully written).
 * dtype: SNVC_F32 | SNVC_F64.
 */
int snvc_cost_volume_bwd(const void* grad, const void* shift, void* grad_left, void* grad_right,
                         int64_t N, int64_t C, int64_t H, int64_t W, int64_t D, int32_t downsample,
                         int32_t dtype, void* stream);

/* Debug: x_low per (n, d, pw) (or -1 when the right sample is outside the image), computed by the
 * same device code as the forward kernel; used by the bit-exact index parity tests. */
int snvc_cost_volume_xlow(const float* shift, int32_t* xlow, int64_t N, int64_t IW, int64_t D,
                          int32_t downsample, void* stream);

/* ---------------------------------------------------------------------------------------------
 * A3  instance branch: high-resolution local ROI voxel sampling.
 * Replaces VernierScale._sample_2d_feat / construct_voxel (snvc/models/vernier.py:323-360):
 * coordinate normalisation (:335-338), 2 x torch.nn.functional.grid_sample (bilinear, zeros,
 * align_corners=False; :339-340) and torch.cat (:346), in one pass.
 * feat_l, feat_r : [N, C, Hf, Wf] fp32 contiguous
 * pts_l, pts_r   : [N, 2, P] fp32 pixel coordinates (x row, y row), P = nh*nw*nl
 * res_x, res_y   : cfg.resolution[1], cfg.resolution[0]  (divisors of x and y, :335-336)
 * out            : NCDHW -> [N, 2C, P] (F32);  NDHWC -> [N, P, 2C] (BF16 or F32)
 * workspace      : >= snvc_roi_voxel_sample_workspace_bytes(N, C, Hf, Wf) bytes (NHWC copies)
 */
int64_t snvc_roi_voxel_sample_workspace_bytes(int64_t N, int64_t C, int64_t Hf, int64_t Wf);
int snvc_roi_voxel_sample_fwd(const float* feat_l, const float* feat_r, const float* pts_l,
                              const float* pts_r, void* out, void* workspace, int64_t N, int64_t C,
                              int64_t Hf, int64_t Wf, int64_t P, float res_x, float res_y,
                              int32_t out_dtype, int32_t out_layout, void* stream);
/* Debug: floor corner indices and in-bounds masks per point: idx [N, P, 2] int32 (x_nw, y_nw),
 * mask [N, P] uint8 (bit0 nw, bit1 ne, bit2 sw, bit3 se). */
int snvc_roi_voxel_sample_indices(const float* pts, int32_t* idx, uint8_t* mask, int64_t N, int64_t P,
                                  int64_t Hf, int64_t Wf, float res_x, float res_y, void* stream);

/* ---------------------------------------------------------------------------------------------
 * A4  global branch: trilinear frustum-to-voxel lift (grid computed in-kernel from P).
 * Replaces (restated composition): project_rect_to_image
 * (snvc/utils/torch_utils.py:36-45) + normalisation by CV_{X,Y,Z}_{MIN,MAX} (key names
 * snvc/models/loss3d.py:15-17) + 5-D torch.nn.functional.grid_sample + validity mask.
 * vol   : NCDHW [N,C,D,H,W] (F32) or NDHWC [N,D,H,W,C] (BF16)
 * proj  : [N, 3, 4] fp32 projection matrices
 * zs,ys,xs : voxel-centre coordinates, fp32, lengths Z, Y, X (torch_utils.py:85-94 convention)
 * cv_range_host : HOST pointer, 6 floats {CV_X_MIN, CV_X_MAX, CV_Y_MIN, CV_Y_MAX, CV_Z_MIN, CV_Z_MAX}
 * out   : NCDHW [N,C,Z,Y,X] (F32) or NDHWC [N,Z,Y,X,C] (BF16 | F32)
 * valid : optional [N,Z,Y,X] uint8 (may be NULL)
 */
int snvc_frustum_lift_fwd(const void* vol, const float* proj, const float* zs, const float* ys,
                          const float* xs, const float* cv_range_host, void* out, uint8_t* valid,
                          int64_t N, int64_t C, int64_t D, int64_t H, int64_t W, int64_t Z, int64_t Y,
                          int64_t X, int32_t align_corners, int32_t in_dtype, int32_t in_layout,
                          int32_t out_dtype, int32_t out_layout, void* stream);
/* Depth-slab variant (multi-GPU stress configuration): `vol` holds only planes
 * [d_base, d_base + D) of a D_total-plane volume (d_base may be negative: leading halo planes);
 * coordinates are normalised against D_total, corners outside the slab contribute zero -- the
 * caller partitions the voxel z-range so that every needed plane is inside its slab. */
int snvc_frustum_lift_slab_fwd(const void* vol, const float* proj, const float* zs, const float* ys,
                               const float* xs, const float* cv_range_host, void* out, uint8_t* valid,
                               int64_t N, int64_t C, int64_t D, int64_t H, int64_t W, int64_t Z, int64_t Y,
                               int64_t X, int32_t align_corners, int32_t in_dtype, int32_t in_layout,
                               int32_t out_dtype, int32_t out_layout, int64_t D_total, int64_t d_base,
                               void* stream);
/* Debug: floor corner (x0,y0,z0) per voxel: idx [N,Z,Y,X,3] int32, valid [N,Z,Y,X] uint8. */
int snvc_frustum_lift_indices(const float* proj, const float* zs, const float* ys, const float* xs,
                              const float* cv_range_host, int32_t* idx, uint8_t* valid, int64_t N,
                              int64_t D, int64_t H, int64_t W, int64_t Z, int64_t Y, int64_t X,
                              int32_t align_corners, void* stream);

/* ---------------------------------------------------------------------------------------------
 * A2  3-D convolution stack (tcgen05 / TMEM implicit GEMM, bf16 x bf16 -> fp32).
 * Replaces the cuDNN calls behind nn.Conv3d / nn.ConvTranspose3d + BatchNorm3d(eval) + ReLU +
 * residual adds of convbn_3d / hourglass / hourglass_downsample_16
 * (snvc/models/submodule.py:32-50, 85-168, 170-268) and the instance 3-D CNN
 * (snvc/models/vernier.py:250-289, 414-438).
 *
 * Activations are NDHWC bf16.  Weights are pre-packed by snvc_conv3d_pack_weights into
 * [taps][Cout_pad][Cin] bf16 (tap-major, K-contiguous).
 */
typedef struct snvc_conv3d_desc {
  int32_t N;                 /* batch */
  int32_t Cin, Cout;         /* Cin in {16,32,64}; Cout <= 64 (padded to a multiple of 16 inside) */
  int32_t Di, Hi, Wi;        /* input spatial extent */
  int32_t Do, Ho, Wo;        /* output spatial extent */
  int32_t kernel;            /* cubic kernel size k (1,3,5,7) */
  int32_t stride;            /* 1 or 2 */
  int32_t pad;
  int32_t dilation;
  int32_t transposed;        /* 1: ConvTranspose3d(k=3, s=2, p=1, output_padding=1) */
  int32_t relu;              /* apply ReLU */
  int32_t residual_mode;     /* 0 none, 1 add before ReLU, 2 add after ReLU */
  int32_t sigmoid;           /* apply sigmoid last (fg_cls_head) */
  int32_t out_dtype;         /* SNVC_BF16 | SNVC_F32 */
  int32_t out_cstride;       /* channel stride of y's innermost dim (>= Cout); 0 -> Cout */
  int32_t out_coffset;       /* first channel written inside that stride */
  int32_t res_cstride;       /* same for the residual tensor; 0 -> Cout */
  int32_t res_coffset;
  int32_t in_cstride;        /* channel stride of x's innermost dim (>= Cin); 0 -> Cin */
  int32_t in_coffset;        /* first channel read inside that stride (multiple of 8) */
  int32_t addend_edge_lo;    /* snvc_conv3d_fwd_addend only: which output planes are the FIRST / LAST plane of the whole   */
  int32_t addend_edge_hi;    /* volume (they take addend plane 0 / 2): 0 = this tensor's own plane 0 / Do-1 (default);     */
                             /* k > 0 = plane k / Do-1-k (a depth slab whose view starts k planes outside the volume);     */
                             /* < 0 = none (an interior depth slab).  Ignored (keep 0) by snvc_conv3d_fwd.                 */
} snvc_conv3d_desc;

/* w: Conv3d [Cout,Cin,k,k,k] fp32 (transposed=0) or ConvTranspose3d [Cin,Cout,k,k,k] fp32
 * (transposed=1), DEVICE pointer.  w_packed: k^3 * Cout_pad * Cin bf16, Cout_pad = roundup(Cout,16). */
int64_t snvc_conv3d_packed_weight_bytes(int32_t Cin, int32_t Cout, int32_t kernel);
int snvc_conv3d_pack_weights(const float* w, void* w_packed, int32_t Cin, int32_t Cout, int32_t kernel,
                             int32_t transposed, void* stream);
/* x [N,Di,Hi,Wi,Cin] bf16; scale,bias [Cout] fp32 (folded eval-mode BatchNorm; NULL = 1 / 0);
 * residual: NDHWC bf16 with the output's spatial shape, or NULL; y: NDHWC.
 * y = act( scale * conv(x) + bias [+ residual] ) [+ residual] */
int snvc_conv3d_fwd(const void* x, const void* w_packed, const float* scale, const float* bias,
                    const void* residual, void* y, const snvc_conv3d_desc* desc, void* stream);
/* Same, with a depth-invariant addend:  y = act( scale * (conv(x) + addend[n, v(d), h, w, :]) + bias ),
 * v(d) = 0 for output plane d = 0, 2 for d = Do-1, 1 otherwise;  addend: fp32 [N, 3, Ho, Wo, Cout].
 * This is how the first trunk layer (convbn_3d(2C, C) on the cost volume, submodule.py:32-50 applied to
 * build_cost_volume's output) consumes the SPLIT cost volume: conv over all 2C channels = conv of the right half
 * (x = right_vol, w = the [C, 2C) input-channel slice of the weights) + the convolution of the depth-constant left
 * half, which is the same for every interior output plane and is computed once by a 3-plane snvc_conv3d_fwd
 * (zero padding in depth gives plane 0 / 1 / 2 the tap sets of d = 0 / interior / D-1).  Exact algebra; only the
 * fp32 summation order differs.  Supported: 3x3x3, stride 1, pad 1, Cin = Cout = 32, Di >= 2, no residual. */
int snvc_conv3d_fwd_addend(const void* x, const void* w_packed, const float* scale, const float* bias,
                           const float* addend, void* y, const snvc_conv3d_desc* desc, void* stream);

/* ---------------------------------------------------------------------------------------------
 * A2a  GroupNorm variant of convbn_3d / convbn (`gn=True` -> nn.GroupNorm(32, C): snvc/models/submodule.py:28,49,135,146).
 * x [N, S, C] fp32 channels-last (the convolution's fp32 result, S = D*H*W or H*W); gamma, beta [C] (NULL = 1 / 0);
 * y = act( GN(x) [+ residual] ) [+ residual] written as bf16 | fp32 into the channel slice [out_coffset, +C) of rows of
 * out_cstride channels; residual: bf16 rows of res_cstride channels.  workspace: snvc_group_norm_workspace_bytes bytes.
 * Three launches (partial sums without atomics, finalize, apply): deterministic. */
int64_t snvc_group_norm_workspace_bytes(int64_t N, int64_t S, int32_t C);
int snvc_group_norm_fwd(const float* x, const float* gamma, const float* beta, const void* residual, void* y,
                        void* workspace, int64_t N, int64_t S, int32_t C, int32_t groups, float eps, int32_t relu,
                        int32_t residual_mode, int32_t sigmoid, int32_t out_dtype, int32_t out_cstride,
                        int32_t out_coffset, int32_t res_cstride, int32_t res_coffset, void* stream);

/* ---------------------------------------------------------------------------------------------
 * N2  2-D convolutions of the BEV tails (tcgen05 / TMEM implicit GEMM with a K loop over taps x 64-channel chunks).
 * Replaces the cuDNN calls behind nn.Conv2d / nn.ConvTranspose2d + BatchNorm2d(eval) + ReLU + skip adds of
 *   convbn                      snvc/models/submodule.py:11-29
 *   hourglass2d                 snvc/models/submodule.py:317-361
 *   hourglass2d_downsample_16   snvc/models/submodule.py:270-315 (helpers :183-195, :210-221)
 *   conv5 / hm1 / hm2           snvc/models/vernier.py:289-314, 440-445
 * Activations NHWC bf16; Cin any multiple of 8 >= 64 (or 16 / 32 / 48), Cout <= 256.
 */
typedef struct snvc_conv2d_desc {
  int32_t N;
  int32_t Cin, Cout;
  int32_t Hi, Wi;            /* input spatial extent */
  int32_t Ho, Wo;            /* output spatial extent */
  int32_t kernel;            /* square kernel size k (1 or 3) */
  int32_t stride;            /* 1 or 2 */
  int32_t pad;
  int32_t dilation;
  int32_t transposed;        /* 1: ConvTranspose2d(k=3, s=2, p=1, output_padding=1) */
  int32_t relu;
  int32_t residual_mode;     /* 0 none, 1 add before ReLU, 2 add after ReLU */
  int32_t sigmoid;
  int32_t out_dtype;         /* SNVC_BF16 | SNVC_F32 */
  int32_t out_cstride;       /* channel stride of y's innermost dim (>= Cout); 0 -> Cout */
  int32_t out_coffset;
  int32_t res_cstride;       /* same for the residual tensor; 0 -> Cout */
  int32_t res_coffset;
  int32_t in_cstride;        /* channel stride of x's innermost dim (>= Cin); 0 -> Cin */
  int32_t in_coffset;        /* first channel read inside that stride (multiple of 8) */
  int32_t reserved[2];
} snvc_conv2d_desc;

/* w: Conv2d [Cout,Cin,k,k] fp32 (transposed=0) or ConvTranspose2d [Cin,Cout,k,k] fp32 (transposed=1), DEVICE pointer.
 * w_packed: [k*k][ceil(Cin/cc)][Cout_pad][cc] bf16, cc = 64 (Cin >= 64) | 32 | 16, Cout_pad = roundup(Cout,16),
 * zero-padded in both channel dimensions. */
int64_t snvc_conv2d_packed_weight_bytes(int32_t Cin, int32_t Cout, int32_t kernel);
int snvc_conv2d_pack_weights(const float* w, void* w_packed, int32_t Cin, int32_t Cout, int32_t kernel,
                             int32_t transposed, void* stream);
/* x [N,Hi,Wi,Cin] bf16; scale,bias [Cout] fp32 (folded eval-mode BatchNorm2d; NULL = 1 / 0); residual: NHWC bf16 with the
 * output's spatial shape, or NULL; y: NHWC.   y = act( scale * conv(x) + bias [+ residual] ) [+ residual] */
int snvc_conv2d_fwd(const void* x, const void* w_packed, const float* scale, const float* bias,
                    const void* residual, void* y, const snvc_conv2d_desc* desc, void* stream);

/* ---------------------------------------------------------------------------------------------
 * N1  projection of the instance sampling grid into the left / right ROI frames.
 * Replaces refinementDataset._generate_grid_proj, the per-proposal numpy float64 loop upstream of A3
 *   snvc/dataset/KITTIRefinement_dataset.py:847-868 (_to_cam :828-845, _init_3d_grid :267-282)
 *   snvc/dataset/kitti_util.py:282-293 (project_rect_to_image), snvc/utils/img_proc.py:71-74 (affine_transform)
 * pose     [N,5] float64: cos(ry+pi/2), sin(ry+pi/2), x, y - h/2, z   (computed on the host as the reference does)
 * P_left/P_right [N,3,4] float64 camera projections; trans_l/trans_r [N,2,3] float64 ROI affines
 * x_pts [nw], y_pts [nh], z_pts [nl] float64: np.linspace of the grid ranges
 * coord_l/coord_r [N,2,P] float32 (P = nh*nw*nl, point index (ih*nw+iw)*nl+il) -- the l_pts / r_pts of A3;
 * grid_cam [N,P,3] float32 camera-frame grid points, or NULL.
 * float64 arithmetic, one cast to float32; dot products are FMA chains in k order (<= 1 float32 ulp vs the
 * reference's dgemm, bit-identical for > 99.99 % of the values). */
int snvc_roi_grid_project(const double* pose, const double* P_left, const double* P_right, const double* trans_l,
                          const double* trans_r, const double* x_pts, const double* y_pts, const double* z_pts,
                          float* coord_l, float* coord_r, float* grid_cam, int64_t N, int64_t nh, int64_t nw,
                          int64_t nl, void* stream);

/* ---------------------------------------------------------------------------------------------
 * A5 / N2  depth head behind the trunk.
 * snvc_disparity_regression replaces `disparityregression.forward` (snvc/models/submodule.py:76-83):
 *   prob [N,K,H*W] fp32 (softmaxed volume), depth_values [K] fp32 -> out [N,H*W] fp32 = sum_k prob * depth[k].
 * snvc_depth_regression_fwd fuses what the (restated, DSGN-lineage) global branch runs in front of it:
 *   F.interpolate(logits [N,1,D,H,W] -> (Dout,Hout,Wout), mode='trilinear', align_corners) -> softmax(dim=depth)
 *   -> disparityregression; logits [N,D,H,W] fp32, depth_values [Dout] fp32 -> out [N,Hout,Wout] fp32.  The
 *   up-sampled probability volume is never written.  fp32; within 1e-5 relative of the torch ops (tests). */
int snvc_disparity_regression(const float* prob, const float* depth_values, float* out, int64_t N, int64_t K,
                              int64_t HW, void* stream);
int snvc_depth_regression_fwd(const float* logits, const float* depth_values, float* out, int64_t N, int64_t D,
                              int64_t H, int64_t W, int64_t Dout, int64_t Hout, int64_t Wout, int32_t align_corners,
                              void* stream);

/* ---------------------------------------------------------------------------------------------
 * N4  rotated BEV IoU and NMS.  Replaces snvc/extension/iou3d_nms:
 *   boxes_iou_bev_gpu   iou3d_nms.cpp:74-95,   iou3d_nms_kernel.cu:228-282
 *   nms_gpu             iou3d_nms.cpp:131-177, iou3d_nms_kernel.cu:296-336 (+ the host keep loop)
 * Boxes are [x, y, z, dx, dy, dz, heading] fp32 rows.
 * snvc_nms_bev: boxes_sorted [N,7] in descending score order (iou3d_nms_utils.py:93-98 sorts before the call);
 *   workspace: snvc_nms_bev_workspace_bytes(N) bytes; keep [N] int64 receives the kept positions (ascending), the
 *   count goes to *num_keep (device int32).  Unlike the reference, nothing is copied to the host and nothing
 *   synchronises: the greedy keep loop runs in a second kernel. */
int snvc_boxes_iou_bev(const float* boxes_a, const float* boxes_b, float* iou, int64_t N, int64_t M, void* stream);
int64_t snvc_nms_bev_workspace_bytes(int64_t N);
int snvc_nms_bev(const float* boxes_sorted, void* workspace, int64_t* keep, int32_t* num_keep, int64_t N,
                 float thresh, void* stream);
/* B independent box sets of N boxes each in two launches (the pairs of a batch): boxes_sorted [B,N,7], workspace
 * B * snvc_nms_bev_workspace_bytes(N) bytes, keep [B,N], num_keep [B]. */
int snvc_nms_bev_batched(const float* boxes_sorted, void* workspace, int64_t* keep, int32_t* num_keep, int64_t B,
                         int64_t N, float thresh, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Layout / elementwise helpers on the path.
 */
/* NCDHW fp32 [N,C,S] <-> NDHWC bf16 [N,S,C]  (S = D*H*W) */
int snvc_ncdhw_f32_to_ndhwc_bf16(const float* src, void* dst, int64_t N, int64_t C, int64_t S, void* stream);
int snvc_ndhwc_bf16_to_ncdhw_f32(const void* src, float* dst, int64_t N, int64_t C, int64_t S, void* stream);
/* vernier.py:433: cat([voxel, vimg * occupancy], dim=1) -- writes channels [coffset, coffset+C) of a
 * [N,S,cstride] bf16 buffer with vimg[N,S,C] * occ[N,S] (occ fp32). */
int snvc_scale_by_occupancy(const void* vimg, const float* occ, void* dst, int64_t NS, int32_t C,
                            int32_t cstride, int32_t coffset, void* stream);
/* vernier.py:435-438: AvgPool3d((pool,1,1)) over the first spatial axis + reshape to BEV:
 * x [N,Dh,H,W,C] bf16 NDHWC -> bev [N, C*(Dh/pool), H, W] fp32 (channel index c*(Dh/pool)+dh). */
int snvc_avgpool_to_bev(const void* x, float* bev, int64_t N, int64_t Dh, int64_t H, int64_t W, int32_t C,
                        int32_t pool, void* stream);
/* Same pooling, emitted channels-last in bf16 for the 2-D tensor-core convs, over either of the first two spatial axes:
 * x [N,S0,S1,S2,C] bf16 -> bev [N, R0, S2, C*Q] bf16 with Q = S_axis/pool (<= 8), R0 = the other of (S0, S1), channel index
 * c*Q + q.  axis 0: the instance branch (pool the height axis nh, vernier.py:436-438); axis 1: the global branch's lifted
 * grid [N,Z,Y,X,C] pooled over Y (restated RPN side). */
int snvc_avgpool_to_bev_nhwc(const void* x, void* bev, int64_t N, int64_t S0, int64_t S1, int64_t S2, int32_t C,
                             int32_t pool, int32_t axis, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Depth-slab halo exchange of the single-volume stress configuration.  No reference
 * counterpart (the reference only has nn.DataParallel over the batch, tools/inference_agnostic.py:472).
 * One volume is split along depth over `world` ranks; an extended slab holds `halo` planes below and above its real
 * planes.  snvc_halo_exchange sends the slab's first / last REAL plane to rank-1 / rank+1 and receives their last /
 * first real plane into the INNER halo planes (index halo-1 and planes_ext-halo) -- one ncclGroup of point-to-point
 * transfers over NVLink on `stream`, no host synchronisation; at the global boundary the inner halo plane is zero-filled
 * (the convolution's zero padding).  x: device pointer to planes_ext planes of plane_bytes bytes each.
 * The communicator is an ncclComm_t (passed as void*): snvc_halo_unique_id on one rank, distribute the 128 bytes out of
 * band (torch.distributed broadcast), snvc_halo_comm_create on every rank.  NCCL is resolved with dlopen at first use. */
int snvc_halo_unique_id(void* id128);
int snvc_halo_comm_create(const void* id128, int32_t world, int32_t rank, void** comm);
int snvc_halo_comm_destroy(void* comm);
int snvc_halo_exchange(void* comm, void* x, int64_t planes_ext, int64_t plane_bytes, int32_t halo, int32_t rank,
                       int32_t world, void* stream);

/* The same exchange over NVLink / NVSwitch PEER MEMORY, without NCCL on the data path (the product path of the stress
 * configuration; snvc_halo_exchange measures 146 GB/s per direction and neighbour on an 8 x B200 box, this 4-5 x that).
 * Every rank keeps its slab buffers in an arena from snvc_peer_alloc (a cudaMalloc block, zero-filled); the 64-byte CUDA
 * IPC handle from snvc_peer_export is distributed out of band and the two neighbours map the arena with snvc_peer_open.
 * Because all ranks carve the arena identically, a slab at offset o of the local arena is at offset o of a neighbour's
 * mapping.  The first snvc_peer_ctl_bytes() bytes of an arena are its control block (epoch words; keep them zero).
 * snvc_halo_push(x, ...): ONE kernel that stores the slab's first / last real plane into the lower / upper neighbour's
 * inner halo plane (x_in_lo_peer / x_in_hi_peer = the address of the same slab in that neighbour's mapped arena; NULL at
 * an end of the volume, where the local inner halo plane is zero-filled instead) and then acts as the neighbour barrier:
 * it returns on the stream once both neighbours' planes have landed in x.  ctl / ctl_lo_peer / ctl_hi_peer: the control
 * blocks of the local arena and of the two mappings.  Every rank must call it for the same sequence of slabs.  A wait
 * longer than ~4 s traps (a lost neighbour must not hang the device).  Capturable in a CUDA graph. */
int snvc_peer_alloc(int64_t bytes, void** ptr);
int snvc_peer_free(void* ptr);
int snvc_peer_export(void* ptr, void* handle64);
int snvc_peer_open(const void* handle64, void** peer_ptr);
int snvc_peer_close(void* peer_ptr);
int64_t snvc_peer_ctl_bytes(void);
int snvc_halo_push(void* x, void* x_in_lo_peer, void* x_in_hi_peer, int64_t planes_ext, int64_t plane_bytes, int32_t halo,
                   void* ctl, void* ctl_lo_peer, void* ctl_hi_peer, int32_t max_blocks, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Host return of a row-masked volume (end-to-end path of the global branch).
 * Replaces the blocking dense `.cpu()` the reference's driver does on its outputs
 * (tools/inference_agnostic.py:396) for the lifted voxel grid: rows with valid[r] != 0 are copied to their dense
 * position in `dst_host_mapped`, rows that were valid in the batch this host buffer held before (prev_valid[r] != 0)
 * and are not now are zero-filled, all other rows are already zero in the buffer and do not move.  prev_valid is
 * updated to valid.  A first use of a host buffer passes prev_valid = all ones (every row is written).
 * src [rows, row_bytes] device; valid, prev_valid [rows] uint8 device; dst_host_mapped: PINNED host memory
 * (cudaHostAlloc / cudaHostRegister; device-addressable under UVA), same shape as src; row_bytes in {16,32,64,128};
 * max_blocks: grid size cap (0 = 2 per SM) -- a small grid leaves the SMs to the kernels of the next batch;
 * moved_bytes: optional device counter (atomicAdd of the bytes written to the host), may be NULL. */
int snvc_masked_rows_to_host(const void* src, const uint8_t* valid, uint8_t* prev_valid, void* dst_host_mapped,
                             int64_t rows, int32_t row_bytes, int32_t max_blocks, unsigned long long* moved_bytes,
                             void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SNVC_B200_H_ */
