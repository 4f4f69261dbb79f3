"""Host-side plan for one fused conv3d (+ eval BatchNorm, residual, ReLU, sigmoid) launch.

`PackedConv3d` owns the tap-major bf16 weights and the folded BN scale/bias of one
nn.Conv3d / nn.ConvTranspose3d (+ BatchNorm3d) pair of the reference
(convbn_3d, snvc/models/submodule.py:32-50; deconv pairs :127-147,197-208) and launches
snvc_conv3d_fwd on NDHWC bf16 activations."""
import ctypes

import torch

from snvc_b200 import _lib


def fold_bn(bn, cout, device):
    """Eval-mode BatchNorm3d -> per-channel (scale, bias): y = scale * x + bias."""
    if bn is None:
        return None, None
    scale = (bn.weight.detach().float() / torch.sqrt(bn.running_var.detach().float() + bn.eps))
    bias = bn.bias.detach().float() - bn.running_mean.detach().float() * scale
    return scale.to(device).contiguous(), bias.to(device).contiguous()


class PackedConv3d:
    def __init__(self, weight, bn=None, *, transposed=False, stride=1, pad=0, dilation=1):
        _lib.require_cuda(weight)
        w = weight.detach().float().contiguous()
        if transposed:
            self.cin, self.cout = w.shape[0], w.shape[1]
        else:
            self.cout, self.cin = w.shape[0], w.shape[1]
        k = w.shape[2]
        if not (w.shape[2] == w.shape[3] == w.shape[4]):
            raise RuntimeError("PackedConv3d: cubic kernels only")
        self.kernel, self.stride, self.pad, self.dilation, self.transposed = k, stride, pad, dilation, bool(transposed)
        L = _lib.lib()
        nbytes = L.snvc_conv3d_packed_weight_bytes(self.cin, self.cout, k)
        self.packed = torch.empty(nbytes, dtype=torch.uint8, device=w.device)
        with torch.cuda.device(w.device):
            st = L.snvc_conv3d_pack_weights(w.data_ptr(), self.packed.data_ptr(), self.cin, self.cout, k,
                                            int(self.transposed), _lib.stream_ptr())
        _lib.check(st, "snvc_conv3d_pack_weights")
        self.scale, self.bias = fold_bn(bn, self.cout, w.device)

    def out_shape(self, x):
        N, Di, Hi, Wi, _ = x.shape
        if self.transposed:
            return N, 2 * Di, 2 * Hi, 2 * Wi
        ext = self.dilation * (self.kernel - 1) + 1
        f = lambda v: (v + 2 * self.pad - ext) // self.stride + 1
        return N, f(Di), f(Hi), f(Wi)

    def __call__(self, x, *, relu=False, residual=None, residual_mode=0, sigmoid=False, out_dtype=torch.bfloat16,
                 out=None, out_coffset=0, res_coffset=0, in_coffset=0, addend=None, addend_edges=(0, 0)):
        """x [N,D,H,W,Cin] bf16 -> y [N,Do,Ho,Wo,Cout] (or the channel slice [out_coffset, +Cout) of `out`).
        `in_coffset` / `res_coffset` select channel slices of wider x / residual buffers.
        `addend` (fp32 [N,3,Ho,Wo,Cout]): depth-invariant term added to the accumulator before scale / bias
        (snvc_conv3d_fwd_addend; plane 0 / 1 / 2 for output depth 0 / interior / last).  `addend_edges` = (lo, hi) says
        which output planes are the volume's first / last plane when x is a depth slab: 0 = the tensor's own edge plane,
        k > 0 = plane k / Do-1-k, -1 = the slab does not contain that edge."""
        _lib.require_cuda(x)
        if x.dtype != torch.bfloat16 or not x.is_contiguous() or x.shape[-1] < in_coffset + self.cin:
            raise RuntimeError(f"conv3d: x must be contiguous NDHWC bf16 with >= {in_coffset + self.cin} channels, "
                               f"got {tuple(x.shape)} {x.dtype}")
        N, Do, Ho, Wo = self.out_shape(x)
        if out is None:
            out = torch.empty((N, Do, Ho, Wo, self.cout), dtype=out_dtype, device=x.device)
        elif tuple(out.shape[:4]) != (N, Do, Ho, Wo) or not out.is_contiguous() or out.dtype not in (torch.bfloat16, torch.float32) \
                or out.shape[-1] < out_coffset + self.cout:
            raise RuntimeError("conv3d: bad `out` tensor (contiguous NDHWC bf16 / fp32 with room for the channel slice)")
        if residual is not None:
            if residual_mode == 0:
                residual_mode = 1
            if residual.dtype != torch.bfloat16 or tuple(residual.shape[:4]) != (N, Do, Ho, Wo) \
                    or not residual.is_contiguous() or residual.shape[-1] < res_coffset + self.cout:
                raise RuntimeError("conv3d: residual must be contiguous NDHWC bf16 with the output's spatial shape")
        d = _lib.ConvDesc(N=N, Cin=self.cin, Cout=self.cout, Di=x.shape[1], Hi=x.shape[2], Wi=x.shape[3],
                          Do=Do, Ho=Ho, Wo=Wo, kernel=self.kernel, stride=self.stride, pad=self.pad,
                          dilation=self.dilation, transposed=int(self.transposed), relu=int(relu),
                          residual_mode=int(residual_mode if residual is not None else 0), sigmoid=int(sigmoid),
                          out_dtype=_lib.BF16 if out.dtype == torch.bfloat16 else _lib.F32,
                          out_cstride=out.shape[-1], out_coffset=out_coffset,
                          res_cstride=residual.shape[-1] if residual is not None else 0, res_coffset=res_coffset,
                          in_cstride=x.shape[-1], in_coffset=in_coffset,
                          addend_edge_lo=int(addend_edges[0]) if addend is not None else 0,
                          addend_edge_hi=int(addend_edges[1]) if addend is not None else 0)
        if addend is not None:
            if residual is not None or sigmoid:
                raise RuntimeError("conv3d: addend cannot be combined with a residual or sigmoid")
            if addend.dtype != torch.float32 or tuple(addend.shape) != (N, 3, Ho, Wo, self.cout) or not addend.is_contiguous():
                raise RuntimeError(f"conv3d: addend must be contiguous fp32 [N,3,Ho,Wo,Cout], got {tuple(addend.shape)} {addend.dtype}")
            with torch.cuda.device(x.device):
                st = _lib.lib().snvc_conv3d_fwd_addend(x.data_ptr(), self.packed.data_ptr(),
                                                       self.scale.data_ptr() if self.scale is not None else None,
                                                       self.bias.data_ptr() if self.bias is not None else None,
                                                       addend.data_ptr(), out.data_ptr(), ctypes.byref(d), _lib.stream_ptr())
            _lib.check(st, "snvc_conv3d_fwd_addend")
            return out
        with torch.cuda.device(x.device):
            st = _lib.lib().snvc_conv3d_fwd(x.data_ptr(), self.packed.data_ptr(),
                                            self.scale.data_ptr() if self.scale is not None else None,
                                            self.bias.data_ptr() if self.bias is not None else None,
                                            residual.data_ptr() if residual is not None else None,
                                            out.data_ptr(), ctypes.byref(d), _lib.stream_ptr())
        _lib.check(st, "snvc_conv3d_fwd")
        return out


class PackedConv2d:
    """Host-side plan for one fused 2-D conv launch (snvc_conv2d_fwd) on NHWC bf16 activations: the packed bf16 weights
    and folded BatchNorm2d scale / bias (or the conv's own bias) of one nn.Conv2d / nn.ConvTranspose2d of the
    reference's BEV tails (convbn submodule.py:11-29; hourglass2d :317-361; conv5 / hm2 vernier.py:289-314)."""

    def __init__(self, weight, bn=None, *, bias=None, transposed=False, stride=1, pad=0, dilation=1):
        _lib.require_cuda(weight)
        w = weight.detach().float().contiguous()
        if transposed:
            self.cin, self.cout = w.shape[0], w.shape[1]
        else:
            self.cout, self.cin = w.shape[0], w.shape[1]
        if w.dim() != 4 or w.shape[2] != w.shape[3]:
            raise RuntimeError("PackedConv2d: square kernels only")
        self.kernel, self.stride, self.pad, self.dilation, self.transposed = w.shape[2], stride, pad, dilation, bool(transposed)
        L = _lib.lib()
        self.packed = torch.empty(L.snvc_conv2d_packed_weight_bytes(self.cin, self.cout, self.kernel), dtype=torch.uint8,
                                  device=w.device)
        with torch.cuda.device(w.device):
            st = L.snvc_conv2d_pack_weights(w.data_ptr(), self.packed.data_ptr(), self.cin, self.cout, self.kernel,
                                            int(self.transposed), _lib.stream_ptr())
        _lib.check(st, "snvc_conv2d_pack_weights")
        self.scale, self.bias = fold_bn(bn, self.cout, w.device)
        if bn is None and bias is not None:
            self.bias = bias.detach().float().to(w.device).contiguous()

    def out_shape(self, x):
        N, Hi, Wi, _ = x.shape
        if self.transposed:
            return N, 2 * Hi, 2 * Wi
        ext = self.dilation * (self.kernel - 1) + 1
        f = lambda v: (v + 2 * self.pad - ext) // self.stride + 1
        return N, f(Hi), f(Wi)

    def __call__(self, x, *, relu=False, residual=None, residual_mode=0, sigmoid=False, out_dtype=torch.bfloat16, out=None,
                 out_coffset=0, res_coffset=0, in_coffset=0):
        """x [N,H,W,>=Cin] bf16 -> y [N,Ho,Wo,Cout] (or the channel slice [out_coffset, +Cout) of `out`)."""
        _lib.require_cuda(x)
        if x.dtype != torch.bfloat16 or not x.is_contiguous() or x.dim() != 4 or x.shape[-1] < in_coffset + self.cin:
            raise RuntimeError(f"conv2d: x must be contiguous NHWC bf16 with >= {in_coffset + self.cin} channels, "
                               f"got {tuple(x.shape)} {x.dtype}")
        N, Ho, Wo = self.out_shape(x)
        if out is None:
            out = torch.empty((N, Ho, Wo, self.cout), dtype=out_dtype, device=x.device)
        elif tuple(out.shape[:3]) != (N, Ho, Wo) or not out.is_contiguous() or out.dtype not in (torch.bfloat16, torch.float32):
            raise RuntimeError("conv2d: bad `out` tensor")
        if residual is not None:
            if residual_mode == 0:
                residual_mode = 1
            if residual.dtype != torch.bfloat16 or tuple(residual.shape[:3]) != (N, Ho, Wo) or not residual.is_contiguous() \
                    or residual.shape[-1] < res_coffset + self.cout:
                raise RuntimeError("conv2d: residual must be contiguous NHWC bf16 with the output's spatial shape")
        d = _lib.Conv2dDesc(N=N, Cin=self.cin, Cout=self.cout, Hi=x.shape[1], Wi=x.shape[2], Ho=Ho, Wo=Wo, kernel=self.kernel,
                            stride=self.stride, pad=self.pad, dilation=self.dilation, transposed=int(self.transposed),
                            relu=int(relu), residual_mode=int(residual_mode if residual is not None else 0),
                            sigmoid=int(sigmoid), out_dtype=_lib.BF16 if out.dtype == torch.bfloat16 else _lib.F32,
                            out_cstride=out.shape[-1], out_coffset=out_coffset,
                            res_cstride=residual.shape[-1] if residual is not None else 0, res_coffset=res_coffset,
                            in_cstride=x.shape[-1], in_coffset=in_coffset)
        with torch.cuda.device(x.device):
            st = _lib.lib().snvc_conv2d_fwd(x.data_ptr(), self.packed.data_ptr(),
                                            self.scale.data_ptr() if self.scale is not None else None,
                                            self.bias.data_ptr() if self.bias is not None else None,
                                            residual.data_ptr() if residual is not None else None, out.data_ptr(),
                                            ctypes.byref(d), _lib.stream_ptr())
        _lib.check(st, "snvc_conv2d_fwd")
        return out
