"""GPU parity: depth head (A5 / N2) -- disparityregression and the fused up-sample + softmax + regression kernel vs
the torch CPU ops the reference calls (F.interpolate 'trilinear', F.softmax, torch.sum; fp32, tolerance 1e-5)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as TF

import synth

pytestmark = pytest.mark.gpu


def _relerr(a, b):
    return float(np.max(np.abs(a - b)) / max(1e-30, np.max(np.abs(b))))


def test_disparity_regression_matches_reference_module_semantics():
    from snvc_b200.models.submodule import disparityregression
    x = torch.softmax(torch.from_numpy(synth.det_uniform((2, 24, 6, 20), 3, -3, 3, bf16=False)), dim=1)
    depth = torch.linspace(2.0, 40.4, 24)
    want = torch.sum(x * depth[None, :, None, None], 1).numpy()               # submodule.py:82
    got = disparityregression(24, None).cuda()(x.cuda(), depth.cuda()).cpu().numpy()
    assert got.shape == want.shape and _relerr(got, want) <= 1e-6


def test_disparity_regression_any_shape_and_alignment():
    """The reference's torch.sum accepts every shape (submodule.py:82-83): odd H*W and a non-16-byte-aligned view."""
    from snvc_b200.models.submodule import disparityregression
    m = disparityregression(7, None).cuda()
    depth = torch.linspace(1.0, 9.0, 7)
    x = torch.softmax(torch.from_numpy(synth.det_uniform((3, 7, 5, 9), 4, -3, 3, bf16=False)), dim=1)     # H*W = 45
    want = torch.sum(x * depth[None, :, None, None], 1).numpy()
    assert _relerr(m(x.cuda(), depth.cuda()).cpu().numpy(), want) <= 1e-6
    big = torch.zeros(3 * 7 * 5 * 9 + 1, device="cuda")
    view = big[1:].view(3, 7, 5, 9)                                           # 4-byte aligned only
    view.copy_(x)
    assert _relerr(m(view, depth.cuda()).cpu().numpy(), want) <= 1e-6


@pytest.mark.parametrize("ac", [True, False])
@pytest.mark.parametrize("shape", [((12, 6, 10), (48, 24, 40)), ((48, 24, 78), (192, 96, 312)), ((5, 7, 9), (11, 13, 30))])
def test_fused_depth_regression_vs_torch_ops(ac, shape):
    from snvc_b200 import functional as F
    (D, H, W), (Do, Ho, Wo) = shape
    logits = torch.from_numpy(synth.det_uniform((2, 1, D, H, W), 7, -4, 4, bf16=False))
    depth = torch.linspace(2.0, 40.4, Do)
    cost = TF.interpolate(logits, [Do, Ho, Wo], mode="trilinear", align_corners=ac)
    want = torch.sum(torch.softmax(cost[:, 0], dim=1) * depth[None, :, None, None], 1).numpy()
    got = F.depth_regression_from_logits(logits.cuda(), depth.cuda(), (Do, Ho, Wo), ac).cpu().numpy()
    assert got.shape == want.shape == (2, Ho, Wo)
    assert _relerr(got, want) <= 1e-5
    assert 2.0 <= got.min() and got.max() <= 40.4


def test_depth_head_module_vs_oracle():
    """classif convs (bf16 tensor cores) + fused regression vs the fp32 torch restatement."""
    import types
    from oracle import blocks as oblocks, global_branch as ogb
    from snvc_b200.models.stereonet import DepthHead
    from snvc_b200 import functional as F
    cfg = types.SimpleNamespace(GN=False, align_corners=True)
    N, C, D, H, W = 1, 32, 8, 12, 20
    head = DepthHead(cfg, C, maxdisp=32).eval()
    sd = synth.det_state_dict(head, 61)
    head.load_state_dict(sd, strict=True)
    ref = torch.nn.Sequential(oblocks.convbn_3d(C, C, 3, 1, 1), torch.nn.ReLU(), torch.nn.Conv3d(C, 1, 3, 1, 1, bias=False)).eval()
    ref.load_state_dict({k.replace("classif.", ""): v for k, v in sd.items()}, strict=True)
    vol = torch.from_numpy(synth.det_uniform((N, C, D, H, W), 62))
    depth = torch.linspace(2.0, 40.4, 32)
    want = ogb.depth_head(vol, ref, depth, (32, 4 * H, 4 * W), True)
    got = head.cuda()(F.to_ndhwc_bf16(vol.cuda()), depth.cuda(), (4 * H, 4 * W)).cpu().numpy()
    assert got.shape == want.shape
    assert _relerr(got, want) <= 1e-2                                          # bf16 convs in front of it
