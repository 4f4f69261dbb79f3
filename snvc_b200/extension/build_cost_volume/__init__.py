"""Drop-in for snvc/extension/build_cost_volume/__init__.py.

    from snvc_b200.extension.build_cost_volume import build_cost_volume
    cost = build_cost_volume(left, right, shift, downsample)      # [N, 2C, D, H/ds, W/ds]

Same positional call, same autograd behaviour (grads for left/right, None for shift and
downsample, __init__.py:17-23 of the reference), same error behaviour (CPU tensors ->
RuntimeError "Not implemented on the CPU", BuildCostVolume.cpp:26; shape mismatches ->
RuntimeError, BuildCostVolume_cuda.cu:216-220).  The one deliberate difference: the reference's
`assert torch.all(shift >= 0.)` (__init__.py:12) forces a device->host sync on every call; here it
is kept by default for parity and can be disabled with SNVC_B200_SKIP_SHIFT_CHECK=1.

`build_cost_volume_ndhwc_bf16` is the product fast path: the same volume emitted channels-last
in bf16, the layout the tcgen05 conv3d consumes.
"""
import ctypes
import os

import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from snvc_b200 import _lib

_DT = {torch.float32: _lib.F32, torch.float64: _lib.F64}


def _check_inputs(left, right, shift):
    _lib.require_cuda(left, right, shift)
    if left.dim() != 4 or tuple(left.shape) != tuple(right.shape):
        raise RuntimeError("Left image and right image should match their size.")
    if shift.dim() != 2 or left.size(0) != shift.size(0):
        raise RuntimeError("Image and shift should of same batch.")
    if left.dtype not in _DT or right.dtype != left.dtype or shift.dtype != left.dtype:
        raise RuntimeError("build_cost_volume: left/right/shift must share a dtype of float32 or float64")


def _forward(left, right, shift, downsample, out_dtype, layout):
    _check_inputs(left, right, shift)
    ds = int(downsample)
    left, right, shift = left.contiguous(), right.contiguous(), shift.contiguous()   # .cu:243-245
    N, C, IH, IW = left.shape
    D = shift.size(1)
    H, W = IH // ds, IW // ds
    shape = (N, 2 * C, D, H, W) if layout == _lib.NCDHW else (N, D, H, W, 2 * C)
    tdt = {_lib.F32: torch.float32, _lib.F64: torch.float64, _lib.BF16: torch.bfloat16}[out_dtype]
    out = torch.empty(shape, dtype=tdt, device=left.device)                           # .cu:228
    with torch.cuda.device(left.device):
        st = _lib.lib().snvc_cost_volume_fwd(left.data_ptr(), right.data_ptr(), shift.data_ptr(), out.data_ptr(),
                                             N, C, IH, IW, D, ds, _DT[left.dtype], out_dtype, layout,
                                             _lib.stream_ptr())
    _lib.check(st, "snvc_cost_volume_fwd")
    return out


class _BuildCostVolume(Function):
    @staticmethod
    def forward(ctx, left, right, shift, downsample):
        ctx.save_for_backward(shift)
        ctx.downsample = downsample
        if os.environ.get("SNVC_B200_SKIP_SHIFT_CHECK", "0") != "1":
            assert torch.all(shift >= 0.)
        return _forward(left, right, shift, downsample, _DT.get(left.dtype, _lib.F32), _lib.NCDHW)

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        shift, = ctx.saved_tensors
        ds = int(ctx.downsample)
        grad = grad_output.contiguous()
        N, C2, D, H, W = grad.shape
        C = C2 // 2
        gl = torch.empty((N, C, H * ds, W * ds), dtype=grad.dtype, device=grad.device)
        gr = torch.empty_like(gl)
        with torch.cuda.device(grad.device):
            st = _lib.lib().snvc_cost_volume_bwd(grad.data_ptr(), shift.contiguous().data_ptr(), gl.data_ptr(),
                                                 gr.data_ptr(), N, C, H, W, D, ds, _DT[grad.dtype], _lib.stream_ptr())
        _lib.check(st, "snvc_cost_volume_bwd")
        return gl, gr, None, None


build_cost_volume = _BuildCostVolume.apply


def build_cost_volume_ndhwc_bf16(left, right, shift, downsample=1):
    """[N,C,IH,IW] fp32 x2 -> [N, D, H, W, 2C] bf16 (inference fast path; no autograd)."""
    return _forward(left, right, shift, downsample, _lib.BF16, _lib.NDHWC)


def build_cost_volume_split_bf16(left, right, shift, downsample=1, parts="both", out_right=None):
    """Split form of the NDHWC bf16 cost volume (snvc_cost_volume_split_fwd): the left half is a pure broadcast over
    depth (BuildCostVolume_cuda.cu:84-86), so it is written once.  Returns
    (right_vol [N,D,H,W,C] = channels [C,2C) of the full volume, left_planes [N,3,H,W,C] = the left features on three
    identical planes, the input of the depth-invariant part of the first trunk convolution).
    `parts` = "right" / "left" builds (and returns) only that half: the two are independent launches.
    `out_right`: a contiguous [N,D,H,W,C] bf16 tensor (e.g. the real planes of a depth slab's extended buffer, N = 1) that
    receives the right half instead of a fresh allocation."""
    _check_inputs(left, right, shift)
    if left.dtype != torch.float32:
        raise RuntimeError("build_cost_volume_split_bf16: fp32 features only")
    ds = int(downsample)
    left, right, shift = left.contiguous(), right.contiguous(), shift.contiguous()
    N, C, IH, IW = left.shape
    D = shift.size(1)
    H, W = IH // ds, IW // ds
    right_vol = left_planes = None
    if parts in ("both", "right"):
        if out_right is not None:
            if tuple(out_right.shape) != (N, D, H, W, C) or out_right.dtype != torch.bfloat16 or not out_right.is_contiguous() \
                    or out_right.device != left.device:
                raise RuntimeError("build_cost_volume_split_bf16: out_right must be a contiguous [N,D,H,W,C] bf16 tensor")
            right_vol = out_right
        else:
            right_vol = torch.empty((N, D, H, W, C), dtype=torch.bfloat16, device=left.device)
    if parts in ("both", "left"):
        left_planes = torch.empty((N, 3, H, W, C), dtype=torch.bfloat16, device=left.device)
    with torch.cuda.device(left.device):
        st = _lib.lib().snvc_cost_volume_split_fwd(left.data_ptr(), right.data_ptr(), shift.data_ptr(),
                                                   right_vol.data_ptr() if right_vol is not None else None,
                                                   left_planes.data_ptr() if left_planes is not None else None,
                                                   N, C, IH, IW, D, ds, _lib.stream_ptr())
    _lib.check(st, "snvc_cost_volume_split_fwd")
    if parts == "right":
        return right_vol
    if parts == "left":
        return left_planes
    return right_vol, left_planes


def cost_volume_xlow(shift, IW, downsample=1):
    """Debug: x_low per (n, d, pw) computed by the kernel's own device code (-1 = outside)."""
    _lib.require_cuda(shift)
    shift = shift.contiguous().float()
    N, D = shift.shape
    out = torch.empty((N, D, IW // downsample), dtype=torch.int32, device=shift.device)
    with torch.cuda.device(shift.device):
        st = _lib.lib().snvc_cost_volume_xlow(shift.data_ptr(), out.data_ptr(), N, IW, D, int(downsample),
                                              _lib.stream_ptr())
    _lib.check(st, "snvc_cost_volume_xlow")
    return out
