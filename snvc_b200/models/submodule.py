"""Drop-in for the 3-D blocks of snvc/models/submodule.py, running on the sm_100a kernels.

Same factory / class names, constructor arguments, forward signatures and state_dict keys as the
reference (`convbn_3d` :32-50, `hourglass` :85-168, `get_hg_down_sample` :170-181,
`get_hg_up_sample` :197-208, `hourglass_downsample_16` :223-268), so a reference checkpoint loads
with strict=True and `vernier.py:20`-style imports keep working.

What differs is execution: every Conv3d/ConvTranspose3d + eval-BatchNorm3d (+ReLU, + residual)
group is ONE launch of the tcgen05 implicit-GEMM kernel (snvc_conv3d_fwd) on channels-last bf16
activations.  Tensor kinds at module boundaries:
  * fp32 / bf16 NCDHW-contiguous input  -> converted once, output returned as fp32 NCDHW
    (exactly what the reference returns);
  * bf16 `torch.channels_last_3d` input -> consumed zero-copy, output is bf16 channels_last_3d
    (chain modules this way to stay in the kernel layout).
Inference only (BatchNorm must be in eval mode, as in tools/inference_agnostic.py:471).
GroupNorm (`gn=True`): conv on the tensor cores with an fp32 result, then the native GroupNorm pass
(snvc_group_norm_fwd: statistics, normalisation, skip add, ReLU, bf16 store).
"""
import torch
import torch.nn as nn

from snvc_b200 import functional as SF
from snvc_b200.conv import PackedConv2d, PackedConv3d


# ----------------------------------------------------------------------------- tensor kinds
def _to_ndhwc(x):
    """-> (NDHWC bf16 contiguous tensor, kind) with kind in {'cl', 'f32'}."""
    if x.dim() != 5:
        raise RuntimeError(f"expected a 5-D volume, got {tuple(x.shape)}")
    if not x.is_cuda:
        raise RuntimeError("Not implemented on the CPU")
    if x.dtype == torch.bfloat16 and x.is_contiguous(memory_format=torch.channels_last_3d):
        return x.permute(0, 2, 3, 4, 1), "cl"          # zero-copy view, NDHWC contiguous
    return SF.to_ndhwc_bf16(x.float()), "f32"


def _from_ndhwc(y, kind):
    if kind == "cl":
        return y.permute(0, 4, 1, 2, 3)                 # logical NCDHW, channels_last_3d strides
    return SF.to_ncdhw_f32(y)


def _opt_ndhwc(t):
    return None if t is None else _to_ndhwc(t)[0]


# ----------------------------------------------------------------------------- conv + norm groups
class _ConvNorm3d(nn.Sequential):
    """nn.Sequential(conv, norm) (same children / keys as the reference) with a fused forward."""

    transposed = False

    def _plan(self):
        conv, norm = self[0], self[1]
        gn = isinstance(norm, nn.GroupNorm)
        if not gn and norm.training:
            raise RuntimeError("snvc_b200 conv blocks are inference-only: call .eval() (BatchNorm uses running stats)")
        vers = (conv.weight.data_ptr(), conv.weight._version, str(conv.weight.device))
        if not gn:
            vers += (norm.weight._version, norm.bias._version, norm.running_mean._version, norm.running_var._version)
        plan = getattr(self, "_snvc_plan", None)
        if plan is None or plan[0] != vers:
            st = conv.stride[0]
            p = PackedConv3d(conv.weight, None if gn else norm, transposed=self.transposed, stride=st,
                             pad=conv.padding[0], dilation=conv.dilation[0])
            plan = (vers, p)
            object.__setattr__(self, "_snvc_plan", plan)
        return plan[1]

    def fused(self, x, *, relu=False, residual=None, residual_mode=0, sigmoid=False, out_dtype=torch.bfloat16,
              out=None, out_coffset=0, in_coffset=0, res_coffset=0):
        """x: NDHWC bf16.  Returns NDHWC."""
        plan = self._plan()
        norm = self[1]
        if isinstance(norm, nn.GroupNorm):
            # conv on the tensor cores (fp32 out), then the native GroupNorm pass: statistics, normalisation, skip add, ReLU
            # and the bf16 store in snvc_group_norm_fwd (no ATen kernels on the path)
            y = plan(x, out_dtype=torch.float32, in_coffset=in_coffset)
            return SF.group_norm_act(y, norm, relu=relu, residual=residual, residual_mode=residual_mode, sigmoid=sigmoid,
                                     out_dtype=out_dtype, out=out, out_coffset=out_coffset, res_coffset=res_coffset)
        return plan(x, relu=relu, residual=residual, residual_mode=residual_mode, sigmoid=sigmoid,
                    out_dtype=out_dtype, out=out, out_coffset=out_coffset, in_coffset=in_coffset,
                    res_coffset=res_coffset)

    def forward(self, x):
        xin, kind = _to_ndhwc(x)
        return _from_ndhwc(self.fused(xin), kind)


class _DeconvNorm3d(_ConvNorm3d):
    transposed = True


def _norm3d(ch, gn, groups=32):
    return nn.GroupNorm(groups, ch) if gn else nn.BatchNorm3d(ch)


def convbn_3d(in_planes, out_planes, kernel_size, stride, pad, dilation=1, gn=False, groups=32):
    """submodule.py:32-50."""
    return _ConvNorm3d(nn.Conv3d(in_planes, out_planes, kernel_size=kernel_size, padding=pad, dilation=dilation,
                                 stride=stride, bias=False),
                       _norm3d(out_planes, gn, groups))


def _deconvbn_3d(cin, cout, gn):
    return _DeconvNorm3d(nn.ConvTranspose3d(cin, cout, kernel_size=3, padding=1, output_padding=1, stride=2,
                                            bias=False),
                         _norm3d(cout, gn))


class _ConvNormReLU(nn.Sequential):
    """nn.Sequential(convbn_3d(...), nn.ReLU(inplace=True)) with one fused launch."""

    def fused(self, x, **kw):
        kw.setdefault("relu", True)
        return self[0].fused(x, **kw)

    def forward(self, x):
        xin, kind = _to_ndhwc(x)
        return _from_ndhwc(self.fused(xin), kind)


def _cbr(cin, cout, k, s, p, d=1, gn=False):
    return _ConvNormReLU(convbn_3d(cin, cout, k, s, p, d, gn=gn), nn.ReLU(inplace=True))


def get_hg_down_sample(channel_in, channel_out, gn, downsample=True):
    """submodule.py:170-181."""
    return _cbr(channel_in, channel_out, 3, 2 if downsample else 1, 1, gn=gn)


def get_hg_up_sample(channel_in, channel_out, gn):
    """submodule.py:197-208."""
    return _deconvbn_3d(channel_in, channel_out, gn)


class disparityregression(nn.Module):
    """submodule.py:76-83: forward(x, depth) = sum(x * depth[None, :, None, None], 1) on the GPU kernel.
    (The reference's constructor calls `.cuda()` on an unused `arange(maxdisp)` buffer, :79; kept as a buffer.)"""

    def __init__(self, maxdisp, cfg=None):
        super().__init__()
        self.register_buffer("disp", torch.arange(maxdisp, dtype=torch.float32), persistent=False)

    def forward(self, x, depth):
        return SF.disparity_regression(x.float(), depth.float())


class hourglass(nn.Module):
    """submodule.py:85-168.  forward(x, presqu, postsqu) -> (out, pre, post); the caller adds the
    residual to `out` (as in the reference)."""

    def __init__(self, inplanes, gn=False):
        super().__init__()
        c2 = inplanes * 2
        self.conv1 = _cbr(inplanes, c2, 3, 2, 1, gn=gn)
        self.conv2 = convbn_3d(c2, c2, kernel_size=3, stride=1, pad=1, gn=gn)
        self.conv3 = _cbr(c2, c2, 3, 2, 1, gn=gn)
        self.conv4 = _cbr(c2, c2, 3, 1, 1, gn=gn)
        self.conv5 = _deconvbn_3d(c2, c2, gn)
        self.conv6 = _deconvbn_3d(c2, inplanes, gn)

    def fused(self, x, presqu=None, postsqu=None, out_residual=None, dst=None, dst_coffset=0):
        """NDHWC bf16 in/out.  `out_residual` fuses the caller's `out + x` into conv6's epilogue."""
        out = self.conv1.fused(x)
        pre = self.conv2.fused(out, relu=True, residual=postsqu, residual_mode=1 if postsqu is not None else 0)
        out = self.conv4.fused(self.conv3.fused(pre))
        post = self.conv5.fused(out, relu=True, residual=presqu if presqu is not None else pre, residual_mode=1)
        out = self.conv6.fused(post, residual=out_residual, residual_mode=1 if out_residual is not None else 0,
                               out=dst, out_coffset=dst_coffset)
        return out, pre, post

    def forward(self, x, presqu, postsqu):
        xin, kind = _to_ndhwc(x)
        out, pre, post = self.fused(xin, _opt_ndhwc(presqu), _opt_ndhwc(postsqu))
        return _from_ndhwc(out, kind), _from_ndhwc(pre, kind), _from_ndhwc(post, kind)


class hourglass_downsample_16(nn.Module):
    """submodule.py:223-268."""

    def __init__(self, inplanes, gn=False):
        super().__init__()
        c2 = inplanes * 2
        self.conv1 = get_hg_down_sample(inplanes, c2, gn)
        self.conv2 = get_hg_down_sample(c2, c2, gn, False)
        self.conv3 = get_hg_down_sample(c2, c2, gn)
        self.conv4 = get_hg_down_sample(c2, c2, gn, False)
        self.conv5 = get_hg_down_sample(c2, c2, gn)
        self.conv6 = get_hg_down_sample(c2, c2, gn, False)
        self.conv7 = get_hg_down_sample(c2, c2, gn)
        self.conv8 = get_hg_down_sample(c2, c2, gn, False)
        self.conv9 = get_hg_up_sample(c2, c2, gn)
        self.conv10 = get_hg_up_sample(c2, c2, gn)
        self.conv11 = get_hg_up_sample(c2, c2, gn)
        self.conv12 = get_hg_up_sample(c2, inplanes, gn)

    def fused(self, x, out_residual=None, dst=None, dst_coffset=0):
        o2 = self.conv2.fused(self.conv1.fused(x))
        o4 = self.conv4.fused(self.conv3.fused(o2))
        o6 = self.conv6.fused(self.conv5.fused(o4))
        o8 = self.conv8.fused(self.conv7.fused(o6))
        i10 = self.conv9.fused(o8, residual=o6, residual_mode=1)       # out_conv9 + out_conv6 (no ReLU, :258-259)
        i11 = self.conv10.fused(i10, residual=o4, residual_mode=1)
        i12 = self.conv11.fused(i11, residual=o2, residual_mode=1)
        return self.conv12.fused(i12, residual=out_residual, residual_mode=1 if out_residual is not None else 0,
                                 out=dst, out_coffset=dst_coffset)

    def forward(self, x):
        xin, kind = _to_ndhwc(x)
        return _from_ndhwc(self.fused(xin), kind)


# ============================================================================================ 2-D BEV blocks (N2)
# Drop-ins for `convbn` (submodule.py:11-29), `get_hg_down_sample_2d` (:183-195), `get_hg_up_sample_2d` (:210-221),
# `hourglass2d` (:317-361) and `hourglass2d_downsample_16` (:270-315) on the 2-D tensor-core kernel (snvc_conv2d_fwd).
# Same names, constructor arguments, forward signatures and state_dict keys.  Tensor kinds as for the 3-D blocks:
# fp32 NCHW in -> fp32 NCHW out; bf16 torch.channels_last in -> bf16 channels_last out (zero-copy).
def _to_nhwc(x):
    if x.dim() != 4:
        raise RuntimeError(f"expected a 4-D feature map, got {tuple(x.shape)}")
    if not x.is_cuda:
        raise RuntimeError("Not implemented on the CPU")
    if x.dtype == torch.bfloat16 and x.is_contiguous(memory_format=torch.channels_last):
        return x.permute(0, 2, 3, 1), "cl"
    return SF.to_ndhwc_bf16(x.float()[:, :, None])[:, 0], "f32"


def _from_nhwc(y, kind):
    if kind == "cl":
        return y.permute(0, 3, 1, 2)
    return SF.to_ncdhw_f32(y[:, None])[:, :, 0]


def _opt_nhwc(t):
    return None if t is None else _to_nhwc(t)[0]


class _ConvNorm2d(nn.Sequential):
    """nn.Sequential(Conv2d | ConvTranspose2d, BatchNorm2d | GroupNorm) with a fused forward."""

    transposed = False

    def _plan(self):
        conv, norm = self[0], self[1]
        gn = isinstance(norm, nn.GroupNorm)
        if not gn and norm.training:
            raise RuntimeError("snvc_b200 conv blocks are inference-only: call .eval() (BatchNorm uses running stats)")
        vers = (conv.weight.data_ptr(), conv.weight._version, str(conv.weight.device))
        if not gn:
            vers += (norm.weight._version, norm.bias._version, norm.running_mean._version, norm.running_var._version)
        plan = getattr(self, "_snvc_plan", None)
        if plan is None or plan[0] != vers:
            p = PackedConv2d(conv.weight, None if gn else norm, transposed=self.transposed, stride=conv.stride[0],
                             pad=conv.padding[0], dilation=conv.dilation[0])
            plan = (vers, p)
            object.__setattr__(self, "_snvc_plan", plan)
        return plan[1]

    def fused(self, x, *, relu=False, residual=None, residual_mode=0, out_dtype=torch.bfloat16):
        """x: NHWC bf16.  Returns NHWC."""
        plan, norm = self._plan(), self[1]
        if isinstance(norm, nn.GroupNorm):
            y = plan(x, out_dtype=torch.float32)
            return SF.group_norm_act(y, norm, relu=relu, residual=residual, residual_mode=residual_mode, out_dtype=out_dtype)
        return plan(x, relu=relu, residual=residual, residual_mode=residual_mode, out_dtype=out_dtype)

    def forward(self, x):
        xin, kind = _to_nhwc(x)
        return _from_nhwc(self.fused(xin), kind)


class _DeconvNorm2d(_ConvNorm2d):
    transposed = True


def _norm2d(ch, gn, groups=32):
    return nn.GroupNorm(groups, ch) if gn else nn.BatchNorm2d(ch)


def convbn(in_planes, out_planes, kernel_size, stride, pad, dilation, gn=False, groups=32):
    """submodule.py:11-29."""
    return _ConvNorm2d(nn.Conv2d(in_planes, out_planes, kernel_size=kernel_size, stride=stride,
                                 padding=dilation if dilation > 1 else pad, dilation=dilation, bias=False),
                       _norm2d(out_planes, gn, groups))


class _ConvNormReLU2d(nn.Sequential):
    def fused(self, x, **kw):
        kw.setdefault("relu", True)
        return self[0].fused(x, **kw)

    def forward(self, x):
        xin, kind = _to_nhwc(x)
        return _from_nhwc(self.fused(xin), kind)


def _cbr2d(cin, cout, stride, gn=False):
    return _ConvNormReLU2d(convbn(cin, cout, 3, stride, 1, 1, gn=gn), nn.ReLU(inplace=True))


def get_hg_down_sample_2d(channel_in, channel_out, gn, downsample=True):
    """submodule.py:183-195."""
    return _cbr2d(channel_in, channel_out, 2 if downsample else 1, gn=gn)


def get_hg_up_sample_2d(channel_in, channel_out, gn):
    """submodule.py:210-221."""
    return _DeconvNorm2d(nn.ConvTranspose2d(channel_in, channel_out, kernel_size=3, padding=1, output_padding=1, stride=2,
                                            bias=False),
                         _norm2d(channel_out, gn))


class hourglass2d(nn.Module):
    """submodule.py:317-361.  forward(x, presqu, postsqu) -> (out, pre, post)."""

    def __init__(self, inplanes, gn=False):
        super().__init__()
        c2 = inplanes * 2
        self.conv1 = _cbr2d(inplanes, c2, 2, gn)
        self.conv2 = convbn(c2, c2, kernel_size=3, stride=1, pad=1, dilation=1, gn=gn)
        self.conv3 = _cbr2d(c2, c2, 2, gn)
        self.conv4 = _cbr2d(c2, c2, 1, gn)
        self.conv5 = get_hg_up_sample_2d(c2, c2, gn)
        self.conv6 = get_hg_up_sample_2d(c2, inplanes, gn)

    def fused(self, x, presqu=None, postsqu=None, out_residual=None, out_dtype=torch.bfloat16):
        out = self.conv1.fused(x)
        pre = self.conv2.fused(out, relu=True, residual=postsqu, residual_mode=1 if postsqu is not None else 0)
        out = self.conv4.fused(self.conv3.fused(pre))
        post = self.conv5.fused(out, relu=True, residual=presqu if presqu is not None else pre, residual_mode=1)
        out = self.conv6.fused(post, residual=out_residual, residual_mode=1 if out_residual is not None else 0,
                               out_dtype=out_dtype)
        return out, pre, post

    def forward(self, x, presqu, postsqu):
        xin, kind = _to_nhwc(x)
        out, pre, post = self.fused(xin, _opt_nhwc(presqu), _opt_nhwc(postsqu))
        return _from_nhwc(out, kind), _from_nhwc(pre, kind), _from_nhwc(post, kind)


class hourglass2d_downsample_16(nn.Module):
    """submodule.py:270-315."""

    def __init__(self, inplanes, gn=False):
        super().__init__()
        c2 = inplanes * 2
        self.conv1 = get_hg_down_sample_2d(inplanes, c2, gn)
        self.conv2 = get_hg_down_sample_2d(c2, c2, gn, False)
        self.conv3 = get_hg_down_sample_2d(c2, c2, gn)
        self.conv4 = get_hg_down_sample_2d(c2, c2, gn, False)
        self.conv5 = get_hg_down_sample_2d(c2, c2, gn)
        self.conv6 = get_hg_down_sample_2d(c2, c2, gn, False)
        self.conv7 = get_hg_down_sample_2d(c2, c2, gn)
        self.conv8 = get_hg_down_sample_2d(c2, c2, gn, False)
        self.conv9 = get_hg_up_sample_2d(c2, c2, gn)
        self.conv10 = get_hg_up_sample_2d(c2, c2, gn)
        self.conv11 = get_hg_up_sample_2d(c2, c2, gn)
        self.conv12 = get_hg_up_sample_2d(c2, inplanes, gn)

    def fused(self, x, out_residual=None, out_dtype=torch.bfloat16):
        o2 = self.conv2.fused(self.conv1.fused(x))
        o4 = self.conv4.fused(self.conv3.fused(o2))
        o6 = self.conv6.fused(self.conv5.fused(o4))
        o8 = self.conv8.fused(self.conv7.fused(o6))
        i10 = self.conv9.fused(o8, residual=o6, residual_mode=1)
        i11 = self.conv10.fused(i10, residual=o4, residual_mode=1)
        i12 = self.conv11.fused(i11, residual=o2, residual_mode=1)
        return self.conv12.fused(i12, residual=out_residual, residual_mode=1 if out_residual is not None else 0,
                                 out_dtype=out_dtype)

    def forward(self, x):
        xin, kind = _to_nhwc(x)
        return _from_nhwc(self.fused(xin), kind)


class BasicBlock(nn.Module):
    """submodule.py:52-74 (2-D residual block of the global backbone): conv1 = convbn + ReLU, conv2 = convbn,
    out = conv2(conv1(x)) + (downsample(x) if downsample else x).  The skip add is fused into conv2's epilogue.
    Supports the kernel's geometry (3x3, stride 1 / 2, dilation 1 with pad 1 or dilation == pad); channel counts
    must be multiples of 8 and >= 16, which holds from the backbone's second stage on."""
    expansion = 1

    def __init__(self, inplanes, planes, stride, downsample, pad, dilation, gn=False):
        super().__init__()
        self.conv1 = _ConvNormReLU2d(convbn(inplanes, planes, 3, stride, pad, dilation, gn=gn), nn.ReLU(inplace=True))
        self.conv2 = convbn(planes, planes, 3, 1, pad, dilation, gn=gn)
        self.downsample = downsample
        self.stride = stride

    def forward(self, x):
        xin, kind = _to_nhwc(x)
        skip = xin
        if self.downsample is not None:
            ds = self.downsample
            skip = ds.fused(xin) if hasattr(ds, "fused") else _to_nhwc(ds(x))[0]
        out = self.conv2.fused(self.conv1.fused(xin), residual=skip, residual_mode=1)
        return _from_nhwc(out, kind)
