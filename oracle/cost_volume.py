"""ORACLE (test infrastructure): plane-sweep cost volume on the CPU.

Two independent restatements of
snvc/extension/build_cost_volume/src/BuildCostVolume_cuda.cu:15-98 (forward) and
:152-205 (backward):

* ``forward_c`` / ``backward_c``  -- the literal scalar loop in ``cost_volume.c``
  (fma_mode=1 reproduces nvcc's default FMA contraction of `.cu:58`).
* ``forward_np``                  -- a vectorised numpy recipe,
  every op rounded in the element type (== fma_mode 0).

They are cross-checked against each other and against hand-computed cases in
tests/test_oracle_cost_volume.py.  Parity vs the reference *binary* is unpinned
(the op is CUDA-only and ships no vectors); see DESIGN.md.
"""
import numpy as np

from . import cbuild


def _chk(left, right, shift):
    assert left.shape == right.shape and left.ndim == 4, "Left image and right image should match their size."
    assert left.shape[0] == shift.shape[0], "Image and shift should of same batch."
    assert left.dtype == right.dtype == shift.dtype and left.dtype in (np.float32, np.float64)


def forward_c(left, right, shift, downsample=1, fma_mode=1):
    """[N,C,IH,IW] x2, shift [N,D] -> cost [N,2C,D,IH//ds,IW//ds] (element type preserved)."""
    left, right, shift = (np.ascontiguousarray(a) for a in (left, right, shift))
    _chk(left, right, shift)
    assert np.all(shift >= 0.0)  # build_cost_volume/__init__.py:12
    N, C, IH, IW = left.shape
    D = shift.shape[1]
    ds = int(downsample)
    assert IH % ds == 0 and IW % ds == 0
    out = np.empty((N, 2 * C, D, IH // ds, IW // ds), dtype=left.dtype)
    if out.size == 0:
        return out
    suf = "f32" if left.dtype == np.float32 else "f64"
    getattr(cbuild.lib(), f"oracle_cost_volume_fwd_{suf}")(
        left.ctypes.data, right.ctypes.data, shift.ctypes.data, out.ctypes.data,
        N, C, IH, IW, D, ds, int(fma_mode))
    return out


def backward_c(grad, shift, downsample=1):
    """grad [N,2C,D,H,W], shift [N,D] -> (grad_left, grad_right) [N,C,H*ds,W*ds]."""
    grad, shift = np.ascontiguousarray(grad), np.ascontiguousarray(shift)
    N, C2, D, H, W = grad.shape
    C, ds = C2 // 2, int(downsample)
    gl = np.zeros((N, C, H * ds, W * ds), dtype=grad.dtype)
    gr = np.zeros_like(gl)
    if grad.size:
        suf = "f32" if grad.dtype == np.float32 else "f64"
        getattr(cbuild.lib(), f"oracle_cost_volume_bwd_{suf}")(
            grad.ctypes.data, shift.ctypes.data, gl.ctypes.data, gr.ctypes.data, N, C, H, W, D, ds)
    return gl, gr


def xlow_c(shift, IW, downsample=1):
    """x_low per (n, d, pw), -1 where the right sample falls outside the image (int32)."""
    shift = np.ascontiguousarray(shift, dtype=np.float32)
    N, D = shift.shape
    out = np.empty((N, D, IW // downsample), dtype=np.int32)
    cbuild.lib().oracle_cost_volume_xlow_f32(shift.ctypes.data, out.ctypes.data, N, IW, D, int(downsample))
    return out


def forward_np(left, right, shift, downsample=1):
    """Vectorised restatement; all arithmetic in the element type, no FMA."""
    _chk(left, right, shift)
    T = left.dtype.type
    N, C, IH, IW = left.shape
    ds = int(downsample)
    H, W = IH // ds, IW // ds
    img_w = W * ds
    iw = (np.arange(W) * ds).astype(left.dtype)
    x = iw[None, None, :] + (-shift)[:, :, None]                     # .cu:84,88  [N,D,W]
    valid = (x >= T(0)) & (x <= T(img_w - 1))
    xc = np.where(x <= T(0), T(0), x)                                # .cu:27
    x0 = xc.astype(np.int64)                                         # .cu:30 (trunc; xc >= 0)
    edge = x0 >= img_w - 1                                           # .cu:41-46
    x0 = np.where(edge, img_w - 1, x0)
    x1 = np.where(edge, img_w - 1, x0 + 1)
    xc = np.where(edge, x0.astype(left.dtype), xc)
    lx = (xc - x0.astype(left.dtype)).astype(left.dtype)             # .cu:49
    hx = (T(1.0) - lx).astype(left.dtype)                            # .cu:50 (hy = 1, ly = 0)
    x0 = np.where(valid, x0, 0)
    x1 = np.where(valid, x1, 0)
    rows = right[:, :, 0:H * ds:ds, :]                               # y_low = ih (integer y, .cu:29)
    idx0 = np.broadcast_to(x0[:, None, :, None, :], (N, C, shift.shape[1], H, W))
    idx1 = np.broadcast_to(x1[:, None, :, None, :], (N, C, shift.shape[1], H, W))
    rows_b = np.broadcast_to(rows[:, :, None, :, :], (N, C, shift.shape[1], H, IW))
    v1 = np.take_along_axis(rows_b, idx0, axis=4)
    v2 = np.take_along_axis(rows_b, idx1, axis=4)
    w1 = hx[:, None, :, None, :]
    w2 = lx[:, None, :, None, :]
    r = ((w1 * v1).astype(left.dtype) + (w2 * v2).astype(left.dtype)).astype(left.dtype)
    r = np.where(valid[:, None, :, None, :], r, T(0))
    l = np.broadcast_to(left[:, :, None, 0:H * ds:ds, 0:W * ds:ds], r.shape)
    return np.concatenate([l, r], axis=1)
