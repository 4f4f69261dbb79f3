"""Functional wrappers (torch tensors in/out) over the C ABI for the gather / layout kernels.

Every function allocates its outputs with torch (the library never allocates), launches on
torch's current stream of the tensors' device and raises RuntimeError on any failure."""
import ctypes

import numpy as np
import torch

from snvc_b200 import _lib

_TORCH_DT = {_lib.F32: torch.float32, _lib.BF16: torch.bfloat16, _lib.F64: torch.float64}


def _dt(t):
    return {torch.float32: _lib.F32, torch.bfloat16: _lib.BF16, torch.float64: _lib.F64}[t]


# ------------------------------------------------------------------------------------ A3
def roi_voxel_sample(left, right, l_pts, r_pts, resolution, out_dtype=torch.float32, layout="NCDHW"):
    """VernierScale._sample_2d_feat(aggregate='concat') in one pass (vernier.py:323-349).

    left, right [N,F,Hf,Wf] fp32; l_pts, r_pts [N,2,P] fp32 pixel coords; resolution = cfg.resolution.
    Returns [N, 2F, P] (layout 'NCDHW', fp32) or [N, P, 2F] (layout 'NDHWC', bf16/fp32); the
    caller reshapes P -> (nh, nw, nl).  The inputs are not modified."""
    _lib.require_cuda(left, right, l_pts, r_pts)
    left, right = left.contiguous().float(), right.contiguous().float()
    l_pts, r_pts = l_pts.contiguous().float(), r_pts.contiguous().float()
    N, C, Hf, Wf = left.shape
    if right.shape != left.shape or l_pts.shape != r_pts.shape or l_pts.dim() != 3 or l_pts.size(1) != 2 \
            or l_pts.size(0) != N:
        raise RuntimeError("roi_voxel_sample: shape mismatch")
    P = l_pts.size(2)
    lay = _lib.NCDHW if layout == "NCDHW" else _lib.NDHWC
    shape = (N, 2 * C, P) if lay == _lib.NCDHW else (N, P, 2 * C)
    out = torch.empty(shape, dtype=out_dtype, device=left.device)
    L = _lib.lib()
    ws = torch.empty(max(16, L.snvc_roi_voxel_sample_workspace_bytes(N, C, Hf, Wf)), dtype=torch.uint8,
                     device=left.device)
    with torch.cuda.device(left.device):
        st = L.snvc_roi_voxel_sample_fwd(left.data_ptr(), right.data_ptr(), l_pts.data_ptr(), r_pts.data_ptr(),
                                         out.data_ptr(), ws.data_ptr(), N, C, Hf, Wf, P, float(resolution[1]),
                                         float(resolution[0]), _dt(out_dtype), lay, _lib.stream_ptr())
    _lib.check(st, "snvc_roi_voxel_sample_fwd")
    return out


def roi_voxel_sample_indices(pts, Hf, Wf, resolution):
    """Debug: (idx [N,P,2] int32 (x_nw,y_nw), mask [N,P] uint8) from the kernel's own device code."""
    _lib.require_cuda(pts)
    pts = pts.contiguous().float()
    N, _, P = pts.shape
    idx = torch.empty((N, P, 2), dtype=torch.int32, device=pts.device)
    mask = torch.empty((N, P), dtype=torch.uint8, device=pts.device)
    with torch.cuda.device(pts.device):
        st = _lib.lib().snvc_roi_voxel_sample_indices(pts.data_ptr(), idx.data_ptr(), mask.data_ptr(), N, P, Hf, Wf,
                                                      float(resolution[1]), float(resolution[0]), _lib.stream_ptr())
    _lib.check(st, "snvc_roi_voxel_sample_indices")
    return idx, mask


# ------------------------------------------------------------------------------------ A4
def _cv(cv_range):
    arr = (ctypes.c_float * 6)(*[float(v) for v in cv_range])
    return arr


def frustum_lift(vol, proj, zs, ys, xs, cv_range, align_corners=True, layout_in="NCDHW", out_dtype=None,
                 layout_out=None, return_valid=False, d_total=None, d_base=0):
    """Trilinear frustum -> world-voxel lift with the sampling grid computed in-kernel.

    vol: [N,C,D,H,W] fp32 (layout_in 'NCDHW') or [N,D,H,W,C] bf16 ('NDHWC'); proj [N,3,4] fp32;
    zs/ys/xs voxel-centre vectors; cv_range = (CV_X_MIN, CV_X_MAX, CV_Y_MIN, CV_Y_MAX, CV_Z_MIN, CV_Z_MAX).
    Returns [N,C,Z,Y,X] ('NCDHW') or [N,Z,Y,X,C] ('NDHWC').
    Depth-slab mode: `vol` holds planes [d_base, d_base + D) of a `d_total`-plane volume."""
    _lib.require_cuda(vol, proj, zs, ys, xs)
    vol = vol.contiguous()
    proj, zs, ys, xs = (t.contiguous().float() for t in (proj, zs, ys, xs))
    lin = _lib.NCDHW if layout_in == "NCDHW" else _lib.NDHWC
    layout_out = layout_out or layout_in
    lout = _lib.NCDHW if layout_out == "NCDHW" else _lib.NDHWC
    if lin == _lib.NCDHW:
        N, C, D, H, W = vol.shape
    else:
        N, D, H, W, C = vol.shape
    out_dtype = out_dtype or (torch.float32 if lout == _lib.NCDHW else vol.dtype)
    Z, Y, X = zs.numel(), ys.numel(), xs.numel()
    shape = (N, C, Z, Y, X) if lout == _lib.NCDHW else (N, Z, Y, X, C)
    out = torch.empty(shape, dtype=out_dtype, device=vol.device)
    valid = torch.empty((N, Z, Y, X), dtype=torch.uint8, device=vol.device) if return_valid else None
    with torch.cuda.device(vol.device):
        st = _lib.lib().snvc_frustum_lift_slab_fwd(vol.data_ptr(), proj.data_ptr(), zs.data_ptr(), ys.data_ptr(),
                                                   xs.data_ptr(), _cv(cv_range), out.data_ptr(),
                                                   valid.data_ptr() if valid is not None else None, N, C, D, H, W, Z,
                                                   Y, X, int(bool(align_corners)), _dt(vol.dtype), lin, _dt(out_dtype),
                                                   lout, D if d_total is None else int(d_total), int(d_base),
                                                   _lib.stream_ptr())
    _lib.check(st, "snvc_frustum_lift_slab_fwd")
    return (out, valid) if return_valid else out


def frustum_lift_indices(proj, zs, ys, xs, cv_range, vol_dhw, align_corners=True):
    """Debug: (idx [N,Z,Y,X,3] int32 floor corners (x0,y0,z0), valid [N,Z,Y,X] uint8)."""
    _lib.require_cuda(proj, zs, ys, xs)
    proj, zs, ys, xs = (t.contiguous().float() for t in (proj, zs, ys, xs))
    N = proj.size(0)
    D, H, W = vol_dhw
    Z, Y, X = zs.numel(), ys.numel(), xs.numel()
    idx = torch.empty((N, Z, Y, X, 3), dtype=torch.int32, device=proj.device)
    valid = torch.empty((N, Z, Y, X), dtype=torch.uint8, device=proj.device)
    with torch.cuda.device(proj.device):
        st = _lib.lib().snvc_frustum_lift_indices(proj.data_ptr(), zs.data_ptr(), ys.data_ptr(), xs.data_ptr(),
                                                  _cv(cv_range), idx.data_ptr(), valid.data_ptr(), N, D, H, W, Z, Y, X,
                                                  int(bool(align_corners)), _lib.stream_ptr())
    _lib.check(st, "snvc_frustum_lift_indices")
    return idx, valid


def roi_grid_project(boxes, P_left, P_right, trans_l, trans_r, x_range, y_range, z_range, grid_resolution, device=None,
                     return_grid=False):
    """GPU replacement of refinementDataset._generate_grid_proj (KITTIRefinement_dataset.py:847-868).

    boxes [N,7] = [h,w,l,x,y,z,ry] (host array-like), P_left / P_right [3,4] or [N,3,4], trans_l / trans_r [N,2,3]
    (host), ranges + grid_resolution = cfg.x_range / y_range / z_range / grid_resolution.  Returns device tensors
    coord_l, coord_r [N,2,P] float32 (what VernierScale.forward takes as grid_proj_left / grid_proj_right) and, if
    asked, the camera-frame grid [N,P,3] float32.  Only O(N) numbers are prepared on the host (pose, linspace)."""
    import numpy as np
    L = _lib.lib()
    dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
    boxes = np.asarray(boxes, dtype=np.float64).reshape(-1, 7)
    N = boxes.shape[0]
    nh, nw, nl = (int(v) for v in grid_resolution)
    ry = boxes[:, 6] + 0.5 * np.pi                                           # _to_cam: heading + pi/2
    pose = np.stack([np.cos(ry), np.sin(ry), boxes[:, 3], boxes[:, 4] - boxes[:, 0] * 0.5, boxes[:, 5]], axis=1)

    def per_proposal(a, shape):
        a = np.asarray(a, dtype=np.float64)
        if a.shape == shape:
            a = np.broadcast_to(a, (N,) + shape)
        if a.shape != (N,) + shape:
            raise RuntimeError(f"roi_grid_project: expected {shape} or {(N,) + shape}, got {a.shape}")
        return np.ascontiguousarray(a)

    host = [pose, per_proposal(P_left, (3, 4)), per_proposal(P_right, (3, 4)), per_proposal(trans_l, (2, 3)),
            per_proposal(trans_r, (2, 3)), np.linspace(x_range[0], x_range[1], nw), np.linspace(y_range[0], y_range[1], nh),
            np.linspace(z_range[0], z_range[1], nl)]
    d = [torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(dev) for a in host]
    P = nh * nw * nl
    cl = torch.empty((N, 2, P), dtype=torch.float32, device=dev)
    cr = torch.empty((N, 2, P), dtype=torch.float32, device=dev)
    cam = torch.empty((N, P, 3), dtype=torch.float32, device=dev) if return_grid else None
    with torch.cuda.device(dev):
        st = L.snvc_roi_grid_project(*[t.data_ptr() for t in d], cl.data_ptr(), cr.data_ptr(),
                                     cam.data_ptr() if cam is not None else None, N, nh, nw, nl, _lib.stream_ptr())
    _lib.check(st, "snvc_roi_grid_project")
    return (cl, cr, cam) if return_grid else (cl, cr)


# ------------------------------------------------------------------------------------ rotated NMS (N4)
def boxes_iou_bev(boxes_a, boxes_b):
    """iou3d_nms_utils.boxes_iou_bev: [N,7], [M,7] fp32 ([x,y,z,dx,dy,dz,heading]) -> IoU [N,M]."""
    _lib.require_cuda(boxes_a, boxes_b)
    a, b = boxes_a.float().contiguous(), boxes_b.float().contiguous()
    if a.shape[-1] != 7 or b.shape[-1] != 7:
        raise RuntimeError("boxes_iou_bev: boxes must be [*, 7]")
    out = torch.empty((a.shape[0], b.shape[0]), dtype=torch.float32, device=a.device)
    with torch.cuda.device(a.device):
        st = _lib.lib().snvc_boxes_iou_bev(a.data_ptr(), b.data_ptr(), out.data_ptr(), a.shape[0], b.shape[0],
                                           _lib.stream_ptr())
    _lib.check(st, "snvc_boxes_iou_bev")
    return out


def nms_gpu_device(boxes, scores, thresh, pre_maxsize=None):
    """Sync-free rotated NMS: returns (selected [n] int64 indices into `boxes`, padded with -1 after the kept ones;
    num_kept int32 device scalar).  The greedy keep loop runs on the GPU (the reference runs it on the host)."""
    _lib.require_cuda(boxes, scores)
    if boxes.shape[-1] != 7:
        raise RuntimeError("nms: boxes must be [N, 7]")
    order = scores.sort(0, descending=True)[1]
    if pre_maxsize is not None:
        order = order[:pre_maxsize]
    b = boxes.float()[order].contiguous()
    n = b.shape[0]
    L = _lib.lib()
    ws = torch.empty(max(8, L.snvc_nms_bev_workspace_bytes(n)), dtype=torch.uint8, device=b.device)
    keep = torch.full((n,), -1, dtype=torch.int64, device=b.device)
    num = torch.zeros((), dtype=torch.int32, device=b.device)
    with torch.cuda.device(b.device):
        st = L.snvc_nms_bev(b.data_ptr(), ws.data_ptr(), keep.data_ptr(), num.data_ptr(), n, float(thresh), _lib.stream_ptr())
    _lib.check(st, "snvc_nms_bev")
    sel = torch.where(keep >= 0, order[keep.clamp(min=0)], keep)
    return sel, num


def nms_gpu_device_batched(boxes, scores, thresh):
    """Sync-free rotated NMS of B independent box sets in two launches: boxes [B,N,7], scores [B,N] -> (selected [B,N] int64
    indices into each set in score order, padded with -1; num_kept [B] int32)."""
    _lib.require_cuda(boxes, scores)
    if boxes.dim() != 3 or boxes.shape[-1] != 7 or tuple(scores.shape) != tuple(boxes.shape[:2]):
        raise RuntimeError("nms: boxes must be [B, N, 7] and scores [B, N]")
    B, n = scores.shape
    order = scores.sort(1, descending=True)[1]
    b = torch.gather(boxes.float(), 1, order[..., None].expand(B, n, 7)).contiguous()
    L = _lib.lib()
    ws = torch.empty(max(8, B * L.snvc_nms_bev_workspace_bytes(n)), dtype=torch.uint8, device=b.device)
    keep = torch.full((B, n), -1, dtype=torch.int64, device=b.device)
    num = torch.zeros((B,), dtype=torch.int32, device=b.device)
    with torch.cuda.device(b.device):
        st = L.snvc_nms_bev_batched(b.data_ptr(), ws.data_ptr(), keep.data_ptr(), num.data_ptr(), B, n, float(thresh), _lib.stream_ptr())
    _lib.check(st, "snvc_nms_bev_batched")
    sel = torch.where(keep >= 0, torch.gather(order, 1, keep.clamp(min=0)), keep)
    return sel, num


def nms_gpu(boxes, scores, thresh, pre_maxsize=None, **kwargs):
    """Drop-in for iou3d_nms_utils.nms_gpu (iou3d_nms_utils.py:86-102): -> (indices of the kept boxes, None)."""
    sel, num = nms_gpu_device(boxes, scores, thresh, pre_maxsize)
    return sel[:int(num.item())].contiguous(), None


# ------------------------------------------------------------------------------------ depth head
def disparity_regression(prob, depth):
    """`disparityregression.forward` (submodule.py:76-83): prob [N,K,H,W] fp32, depth [K] fp32 -> [N,H,W]."""
    _lib.require_cuda(prob, depth)
    if prob.dtype != torch.float32 or depth.dtype != torch.float32:
        raise RuntimeError("disparity_regression: fp32 tensors expected")
    prob, depth = prob.contiguous(), depth.contiguous()
    N, K, H, W = prob.shape
    out = torch.empty((N, H, W), dtype=torch.float32, device=prob.device)
    with torch.cuda.device(prob.device):
        st = _lib.lib().snvc_disparity_regression(prob.data_ptr(), depth.data_ptr(), out.data_ptr(), N, K, H * W,
                                                  _lib.stream_ptr())
    _lib.check(st, "snvc_disparity_regression")
    return out


def depth_regression_from_logits(logits, depth_values, out_size, align_corners=True):
    """Fused F.interpolate(trilinear) -> softmax(depth) -> disparityregression.
    logits [N,D,H,W] (or [N,1,D,H,W]) fp32, depth_values [Dout] fp32, out_size (Dout,Hout,Wout) -> [N,Hout,Wout]."""
    _lib.require_cuda(logits, depth_values)
    if logits.dim() == 5:
        logits = logits[:, 0]
    logits, depth_values = logits.float().contiguous(), depth_values.float().contiguous()
    N, D, H, W = logits.shape
    Dout, Hout, Wout = (int(v) for v in out_size)
    if depth_values.numel() != Dout:
        raise RuntimeError(f"depth_regression_from_logits: {Dout} depth values expected, got {depth_values.numel()}")
    out = torch.empty((N, Hout, Wout), dtype=torch.float32, device=logits.device)
    with torch.cuda.device(logits.device):
        st = _lib.lib().snvc_depth_regression_fwd(logits.data_ptr(), depth_values.data_ptr(), out.data_ptr(), N, D, H, W,
                                                  Dout, Hout, Wout, int(bool(align_corners)), _lib.stream_ptr())
    _lib.check(st, "snvc_depth_regression_fwd")
    return out


# ------------------------------------------------------------------------------------ GroupNorm (gn=True blocks)
def group_norm_act(x, norm, *, relu=False, residual=None, residual_mode=0, sigmoid=False, out_dtype=torch.bfloat16,
                   out=None, out_coffset=0, res_coffset=0):
    """nn.GroupNorm `norm` (+ skip add, ReLU, sigmoid) on a channels-last fp32 tensor x [N, ..., C] (the fp32 result of a
    bias-free conv: convbn_3d / convbn with gn=True, submodule.py:28,49) -> [N, ..., C] `out_dtype`, or the channel slice
    [out_coffset, +C) of `out`.  residual: bf16 channels-last; mode 1 = add before ReLU, 2 = after."""
    _lib.require_cuda(x)
    if x.dtype != torch.float32 or not x.is_contiguous():
        raise RuntimeError("group_norm_act: contiguous channels-last fp32 input expected")
    N, C = x.shape[0], x.shape[-1]
    S = x.numel() // (N * C) if N * C else 0
    if out is None:
        out = torch.empty(x.shape, dtype=out_dtype, device=x.device)
    elif tuple(out.shape[:-1]) != tuple(x.shape[:-1]) or not out.is_contiguous() or out.shape[-1] < out_coffset + C \
            or out.dtype not in (torch.bfloat16, torch.float32):
        raise RuntimeError("group_norm_act: bad `out` tensor")
    if residual is not None:
        if residual_mode == 0:
            residual_mode = 1
        if residual.dtype != torch.bfloat16 or tuple(residual.shape[:-1]) != tuple(x.shape[:-1]) or not residual.is_contiguous() \
                or residual.shape[-1] < res_coffset + C:
            raise RuntimeError("group_norm_act: residual must be contiguous channels-last bf16 with the input's spatial shape")
    L = _lib.lib()
    ws = torch.empty(max(16, L.snvc_group_norm_workspace_bytes(N, S, C)), dtype=torch.uint8, device=x.device)
    g = norm.weight.detach().float().contiguous() if norm.weight is not None else None
    b = norm.bias.detach().float().contiguous() if norm.bias is not None else None
    with torch.cuda.device(x.device):
        st = L.snvc_group_norm_fwd(x.data_ptr(), g.data_ptr() if g is not None else None, b.data_ptr() if b is not None else None,
                                   residual.data_ptr() if residual is not None else None, out.data_ptr(), ws.data_ptr(), N, S, C,
                                   int(norm.num_groups), float(norm.eps), int(relu),
                                   int(residual_mode if residual is not None else 0), int(sigmoid),
                                   _lib.BF16 if out.dtype == torch.bfloat16 else _lib.F32, out.shape[-1], out_coffset,
                                   residual.shape[-1] if residual is not None else 0, res_coffset, _lib.stream_ptr())
    _lib.check(st, "snvc_group_norm_fwd")
    return out


# ------------------------------------------------------------------------------------ layouts
def to_ndhwc_bf16(x):
    """[N,C,D,H,W] fp32 -> [N,D,H,W,C] bf16."""
    _lib.require_cuda(x)
    x = x.contiguous().float()
    N, C = x.shape[:2]
    sp = tuple(x.shape[2:])
    S = int(np.prod(sp))
    out = torch.empty((N,) + sp + (C,), dtype=torch.bfloat16, device=x.device)
    with torch.cuda.device(x.device):
        st = _lib.lib().snvc_ncdhw_f32_to_ndhwc_bf16(x.data_ptr(), out.data_ptr(), N, C, S, _lib.stream_ptr())
    _lib.check(st, "snvc_ncdhw_f32_to_ndhwc_bf16")
    return out


def to_ncdhw_f32(x):
    """[N,D,H,W,C] bf16 -> [N,C,D,H,W] fp32."""
    _lib.require_cuda(x)
    x = x.contiguous()
    assert x.dtype == torch.bfloat16
    N, C = x.shape[0], x.shape[-1]
    sp = tuple(x.shape[1:-1])
    S = int(np.prod(sp))
    out = torch.empty((N, C) + sp, dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        st = _lib.lib().snvc_ndhwc_bf16_to_ncdhw_f32(x.data_ptr(), out.data_ptr(), N, C, S, _lib.stream_ptr())
    _lib.check(st, "snvc_ndhwc_bf16_to_ncdhw_f32")
    return out
