"""GPU parity of the 2-D tensor-core convolution (snvc_conv2d_fwd), the native GroupNorm pass and the BEV tails built
on them against plain torch fp32 and the reference-generated goldens.
bf16 operands, fp32 accumulation: bar max|a-b| / max|b| <= 1e-2 (observed ~3e-3)."""
import types

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import synth
from oracle import blocks as oblocks

pytestmark = pytest.mark.gpu
TOL = 1e-2


@pytest.fixture(autouse=True)
def _no_grad():
    with torch.no_grad():
        yield


def _relerr(a, b):
    a, b = np.asarray(a, np.float32), np.asarray(b, np.float32)
    assert a.shape == b.shape, (a.shape, b.shape)
    return float(np.max(np.abs(a - b)) / max(1e-30, np.max(np.abs(b))))


def _nhwc(x):
    return torch.from_numpy(x).cuda().permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)


@pytest.mark.parametrize("Cin,Cout,k,stride,hw,N", [(256, 64, 3, 1, (24, 40), 2), (160, 64, 3, 1, (16, 24), 1), (64, 128, 3, 2, (13, 27), 2),
                                                    (128, 128, 3, 1, (12, 8), 1), (128, 128, 3, 2, (24, 16), 1), (64, 9, 3, 1, (48, 32), 2),
                                                    (32, 32, 1, 1, (9, 11), 1), (128, 256, 3, 1, (10, 14), 1), (16, 48, 3, 1, (8, 8), 1),
                                                    (72, 40, 3, 1, (11, 9), 1), (256, 64, 3, 1, (128, 192), 1)])
def test_conv2d_vs_torch(Cin, Cout, k, stride, hw, N):
    from snvc_b200.conv import PackedConv2d
    H, W = hw
    x = synth.det_uniform((N, Cin, H, W), 11)
    a = float(np.sqrt(3.0 / (Cin * k * k)))
    w = synth.det_uniform((Cout, Cin, k, k), 12, -a, a)
    bias = synth.det_uniform((Cout,), 13, -0.3, 0.3, bf16=False)
    want = F.conv2d(torch.from_numpy(x), torch.from_numpy(w), torch.from_numpy(bias), stride=stride, padding=k // 2)
    conv = PackedConv2d(torch.from_numpy(w).cuda(), None, bias=torch.from_numpy(bias).cuda(), stride=stride, pad=k // 2)
    got = conv(_nhwc(x), out_dtype=torch.float32)
    assert _relerr(got.permute(0, 3, 1, 2).cpu().numpy(), want.numpy()) <= TOL
    got16 = conv(_nhwc(x), relu=True)                                        # bf16 output + ReLU
    assert got16.dtype == torch.bfloat16
    assert _relerr(got16.float().permute(0, 3, 1, 2).cpu().numpy(), torch.relu(want).numpy()) <= TOL


@pytest.mark.parametrize("Cin,Cout,hw,mode", [(128, 128, (6, 4), 1), (128, 64, (12, 8), 1), (64, 64, (7, 5), 2), (128, 128, (24, 16), 0)])
def test_deconv2d_with_skip_vs_torch(Cin, Cout, hw, mode):
    from snvc_b200.conv import PackedConv2d
    H, W = hw
    x = synth.det_uniform((2, Cin, H, W), 21)
    a = float(np.sqrt(3.0 / (Cin * 9 / 4)))
    w = synth.det_uniform((Cin, Cout, 3, 3), 22, -a, a)
    res = synth.det_uniform((2, Cout, 2 * H, 2 * W), 23)
    bn = torch.nn.BatchNorm2d(Cout).eval()
    bn.load_state_dict(synth.det_state_dict(bn, 24))
    y = bn(F.conv_transpose2d(torch.from_numpy(x), torch.from_numpy(w), stride=2, padding=1, output_padding=1))
    r = torch.from_numpy(res)
    want = torch.relu(y + r) if mode == 1 else (torch.relu(y) + r if mode == 2 else torch.relu(y))
    conv = PackedConv2d(torch.from_numpy(w).cuda(), bn.cuda(), transposed=True, stride=2, pad=1)
    got = conv(_nhwc(x), relu=True, residual=_nhwc(res) if mode else None, residual_mode=mode, out_dtype=torch.float32)
    assert _relerr(got.permute(0, 3, 1, 2).cpu().numpy(), want.numpy()) <= TOL


def test_conv2d_channel_slices_and_errors():
    from snvc_b200.conv import PackedConv2d
    x = synth.det_uniform((1, 96, 10, 12), 31)
    w = synth.det_uniform((32, 64, 3, 3), 32, -0.07, 0.07)
    want = F.conv2d(torch.from_numpy(x[:, 32:]), torch.from_numpy(w), padding=1)
    conv = PackedConv2d(torch.from_numpy(w).cuda(), None, stride=1, pad=1)
    out = torch.zeros((1, 10, 12, 64), dtype=torch.float32, device="cuda")
    conv(_nhwc(x), in_coffset=32, out=out, out_coffset=32)
    assert _relerr(out[..., 32:].permute(0, 3, 1, 2).cpu().numpy(), want.numpy()) <= TOL and float(out[..., :32].abs().max()) == 0
    with pytest.raises(RuntimeError):
        conv(torch.zeros((1, 10, 12, 64), device="cuda"))                   # fp32 input
    with pytest.raises(RuntimeError):
        PackedConv2d(torch.zeros((300, 64, 3, 3), device="cuda"))(torch.zeros((1, 4, 4, 64), dtype=torch.bfloat16, device="cuda"))


@pytest.mark.parametrize("axis,shape,pool", [(0, (2, 16, 12, 20, 32), 4), (1, (1, 10, 20, 14, 32), 4), (0, (1, 32, 6, 8, 32), 4), (1, (2, 6, 8, 5, 64), 2)])
def test_avgpool_to_bev_nhwc(axis, shape, pool):
    from snvc_b200 import _lib
    N, S0, S1, S2, C = shape
    x = torch.from_numpy(synth.det_uniform(shape, 41)).cuda().to(torch.bfloat16)
    Q = (S0 if axis == 0 else S1) // pool
    R0 = S1 if axis == 0 else S0
    bev = torch.empty((N, R0, S2, C * Q), dtype=torch.bfloat16, device="cuda")
    st = _lib.lib().snvc_avgpool_to_bev_nhwc(x.data_ptr(), bev.data_ptr(), N, S0, S1, S2, C, pool, axis, _lib.stream_ptr())
    _lib.check(st, "snvc_avgpool_to_bev_nhwc")
    v = x.float().permute(0, 4, 1, 2, 3)                                      # [N,C,S0,S1,S2]
    if axis == 1:
        v = v.permute(0, 1, 3, 2, 4)                                          # pooled axis first
    p = F.avg_pool3d(v, (pool, 1, 1), stride=(pool, 1, 1))                    # [N,C,Q,R0,S2]
    want = p.reshape(N, C * Q, R0, S2).permute(0, 2, 3, 1).to(torch.bfloat16)
    assert torch.equal(bev, want)


@pytest.mark.parametrize("C,G,spatial,mode", [(32, 32, (6, 10, 14), 0), (64, 32, (4, 6, 10), 1), (64, 32, (24, 16), 2), (128, 32, (12, 8), 0),
                                              (160, 32, (5, 7), 1)])
def test_native_group_norm_vs_torch(C, G, spatial, mode):
    from snvc_b200 import functional as SF
    N = 2
    x = synth.det_uniform((N,) + spatial + (C,), 51, -2.0, 3.0, bf16=False)
    res = synth.det_uniform((N,) + spatial + (C,), 52)
    gn = torch.nn.GroupNorm(G, C)
    gn.load_state_dict(synth.det_state_dict(gn, 53))
    perm_in = (0, len(spatial) + 1) + tuple(range(1, len(spatial) + 1))
    perm_out = (0,) + tuple(range(2, len(spatial) + 2)) + (1,)
    y = gn(torch.from_numpy(x).permute(*perm_in)).permute(*perm_out)
    r = torch.from_numpy(res)
    want = torch.relu(y + r) if mode == 1 else (torch.relu(y) + r if mode == 2 else torch.relu(y))
    got = SF.group_norm_act(torch.from_numpy(x).cuda(), gn.cuda(), relu=True,
                            residual=torch.from_numpy(res).cuda().to(torch.bfloat16) if mode else None, residual_mode=mode,
                            out_dtype=torch.float32)
    assert _relerr(got.cpu().numpy(), want.numpy()) <= 1e-5


def test_hourglass2d_blocks_match_reference_golden(golden):
    from snvc_b200.models.submodule import hourglass2d, hourglass2d_downsample_16
    g = golden("hourglass2d_bn")
    m = hourglass2d(64).eval()
    m.load_state_dict(synth.det_state_dict(m, 51), strict=True)               # same keys as the reference module
    out, pre, post = m.cuda()(torch.from_numpy(synth.det_uniform((1, 64, 24, 16), 103)).cuda(), None, None)
    assert out.dtype == torch.float32
    for name, t in (("out", out), ("pre", pre), ("post", post)):
        assert _relerr(t.cpu().numpy(), g[name]) <= TOL, name
    m16 = hourglass2d_downsample_16(64).eval()
    m16.load_state_dict(synth.det_state_dict(m16, 52), strict=True)
    out16 = m16.cuda()(torch.from_numpy(synth.det_uniform((1, 64, 48, 32), 104)).cuda())
    assert _relerr(out16.cpu().numpy(), g["out16"]) <= TOL


def _vernier_cfg(grid=(16, 32, 48), **kw):
    ns = types.SimpleNamespace
    return ns(vernier_type="BEV_type3", gn=False, hrfeat=ns(output_channel=32), num_parts=9, grid_resolution=list(grid),
              n_sample_h=grid[0], n_sample_w=grid[1], n_sample_l=grid[2], resolution=[64, 64], **kw)


def _vernier_inputs(P):
    lf, rf = synth.det_uniform((1, 32, 16, 16), 201), synth.det_uniform((1, 32, 16, 16), 202)
    gl = synth.det_uniform((1, 2, P), 203, -6.4, 70.4, bf16=False)
    gr = synth.det_uniform((1, 2, P), 204, -6.4, 70.4, bf16=False)
    return [torch.from_numpy(a).cuda() for a in (lf, rf, gl, gr)]


def test_vernier_heatmaps_match_reference_golden(golden):
    """construct_voxel -> 3-D CNN -> BEV -> conv5 / hm1 / hm2, all on the GPU kernels, vs the reference's own
    predict_3d_heatmaps output `ncf` (tests/golden/make_golden.py)."""
    from snvc_b200.models.vernier import VernierHotPath
    g = golden("vernier_bev3")
    cfg = _vernier_cfg()
    m = VernierHotPath(cfg, bev_tail=True).eval()
    sd = synth.det_state_dict(oblocks.Vernier3D(32, n_sample_w=32), 31)
    sd.update(synth.det_state_dict(oblocks.VernierBevTail(128, 9, n_sample_w=32), 33))
    m.load_state_dict(sd, strict=True)
    m = m.cuda()
    vox = m.construct_voxel(*_vernier_inputs(16 * 32 * 48))
    heat, occ, offset = m.predict_heatmaps(vox)
    assert offset is None and tuple(heat.shape) == g["ncf"].shape
    assert _relerr(heat.cpu().numpy(), g["ncf"]) <= TOL
    assert _relerr(occ.cpu().numpy(), g["occupancy"]) <= TOL


class _FakeVernierScale(torch.nn.Module):
    """Stand-in with the reference VernierScale's attribute names and forward structure (vernier.py:26-55,460-555) built
    from the oracle's plain-torch blocks: what `accelerate` is handed on a real installation."""

    def __init__(self, cfg):
        super().__init__()
        self.cfg = cfg
        ref3d = oblocks.Vernier3D(32, n_sample_w=cfg.n_sample_w)
        for name in ("vimg_feat", "conv1", "conv2", "conv3", "conv4", "hg_conv3d", "fg_cls_head", "pool_3d"):
            setattr(self, name, getattr(ref3d, name))
        tail = oblocks.VernierBevTail(32 * cfg.grid_resolution[0] // 4, 9, n_sample_w=cfg.n_sample_w)
        self.conv5, self.hm1, self.hm2 = tail.conv5, tail.hm1, tail.hm2
        self.coord_head = torch.nn.Sequential(torch.nn.Conv2d(11, 18, 3, 2, 1), torch.nn.AdaptiveAvgPool2d(1), torch.nn.Sigmoid())
        self.coor_maps = torch.zeros((1, 2, cfg.grid_resolution[2], cfg.grid_resolution[1]))


def test_accelerate_keeps_keys_and_dispatches_through_self(golden):
    from snvc_b200.models.vernier import accelerate
    g = golden("vernier_bev3")
    cfg = _vernier_cfg()
    model = _FakeVernierScale(cfg).eval()
    sd = model.state_dict()
    sd.update(synth.det_state_dict(oblocks.Vernier3D(32, n_sample_w=32), 31))
    sd.update(synth.det_state_dict(oblocks.VernierBevTail(128, 9, n_sample_w=32), 33))
    model.load_state_dict(sd, strict=True)
    keys = list(model.state_dict().keys())
    model = accelerate(model.cuda())
    assert list(model.state_dict().keys()) == keys                           # a reference checkpoint still loads strict=True
    model.load_state_dict({k: v for k, v in sd.items()}, strict=True)
    args = _vernier_inputs(16 * 32 * 48)
    vox = model.construct_voxel(*args)
    heat, occ, offset, coords, bbox = model.predict_3d_heatmaps(vox)
    assert _relerr(heat.cpu().numpy(), g["ncf"]) <= TOL and _relerr(occ.cpu().numpy(), g["occupancy"]) <= TOL
    assert tuple(coords.shape) == (1, 9, 2) and bbox is None and offset is None
    # nn.DataParallel replicas are shallow copies that keep the class: the fused stages must run on the replica's own
    # sub-modules (bound methods captured the device-0 model)
    rep = model._replicate_for_data_parallel()
    assert type(rep) is type(model) and rep.predict_3d_heatmaps.__self__ is rep
    heat2 = rep.predict_3d_heatmaps(rep.construct_voxel(*args))[0]
    assert torch.equal(heat2, heat)


def test_part_reg_head_variant():
    from snvc_b200.models.vernier import VernierHotPath
    cfg = _vernier_cfg(grid=(16, 16, 24), use_part_reg_head=True)
    m = VernierHotPath(cfg, bev_tail=True).eval()
    m.load_state_dict(synth.det_state_dict(m, 61), strict=True)
    m = m.cuda()
    P = 16 * 16 * 24
    vox = m.construct_voxel(*_vernier_inputs(P))
    heat, occ, offset = m.predict_heatmaps(vox)
    assert tuple(offset.shape) == (1, 27, 16, 16, 24) and tuple(heat.shape) == (1, 9, 24, 16)
    # oracle: the same layers in plain torch (vernier.py:279-288,428-431)
    x = vox.float().permute(0, 4, 1, 2, 3).cpu()
    ref = oblocks.Vernier3D(32, n_sample_w=16).eval()
    ref.load_state_dict({k: v for k, v in m.state_dict().items() if k in ref.state_dict()}, strict=True)
    head = torch.nn.Sequential(oblocks.convbn_3d(32, 32, 3, 1, 1), torch.nn.ReLU(), torch.nn.Conv3d(32, 27, 1, 1, 0, bias=False)).eval()
    head.load_state_dict({k[len("part_reg_head."):]: v for k, v in m.state_dict().items() if k.startswith("part_reg_head.")})
    v = ref.conv1(x); v = ref.conv2(v) + v; v = ref.conv3(v) + v
    v = ref.hg_conv3d(v, None, None)[0] + v
    assert _relerr(offset.cpu().numpy(), head(v).numpy()) <= TOL


@pytest.mark.parametrize("n_y", [8, 20])
def test_rpn3d_head_vs_oracle_and_device_nms(n_y):
    """Lifted grid -> rpn3d convs -> Y-pool -> BEV hourglass -> heads on the GPU kernels vs the plain-torch oracle, then
    the device-side decode + rotated NMS chain against the oracle NMS on the same boxes."""
    from oracle import iou3d_nms as onms
    from snvc_b200.models.stereonet import RPN3DHead, decode_proposals
    Z, X = 16, 24
    cfg = types.SimpleNamespace(GN=False, RPN_CONVDIM=32, num_angles=4, num_classes=1, box_corner_parameters=False,
                                X_MIN=-2.4, X_MAX=2.4, Z_MIN=2.0, Z_MAX=5.2, VOXEL_X_SIZE=0.2, VOXEL_Z_SIZE=0.2)
    m = RPN3DHead(cfg, channels=32, n_y=n_y).eval()
    sd = synth.det_state_dict(m, 71)
    m.load_state_dict(sd, strict=True)
    m = m.cuda()
    ref = oblocks.RPN3DHead(32, n_y=n_y).eval()
    ref.load_state_dict(sd, strict=True)
    vox = synth.det_uniform((2, 32, Z, n_y, X), 72)
    want = ref(torch.from_numpy(vox))
    got = m(torch.from_numpy(vox).cuda().permute(0, 2, 3, 4, 1).contiguous().to(torch.bfloat16))
    for name, a, b in zip(("cls", "reg", "ctr"), got, want):
        assert tuple(a.shape) == tuple(b.shape), name
        assert _relerr(a.cpu().numpy(), b.numpy()) <= TOL, name
    boxes, scores, keep, num = decode_proposals(*got, cfg, pre_nms=128, iou_thresh=0.1)
    assert tuple(boxes.shape) == (2, 128, 7) and bool((scores[:, :-1] >= scores[:, 1:]).all())
    for n in range(2):
        b = boxes[n].cpu().numpy()
        bev = np.stack([b[:, 0], b[:, 2], b[:, 1], b[:, 3], b[:, 4], b[:, 5], b[:, 6]], axis=1).astype(np.float32)
        want_keep = onms.nms(bev, scores[n].cpu().numpy(), 0.1)
        k = int(num[n].item())
        assert k == len(want_keep) and np.array_equal(keep[n, :k].cpu().numpy(), want_keep) and bool((keep[n, k:] == -1).all())


@pytest.mark.parametrize("stride,dil", [(1, 1), (2, 1), (1, 2)])
def test_basic_block_vs_torch(stride, dil):
    """submodule.py:52-74 with the kernel-backed convbn: same keys, same numbers as the plain-torch block."""
    from snvc_b200.models.submodule import BasicBlock, convbn
    cin, planes = 32, 64
    ds = convbn(cin, planes, 1, stride, 0, 1)                                 # the 1x1 projection the reference passes in (:391-397)
    m = BasicBlock(cin, planes, stride, ds, 1, dil).eval()
    m.load_state_dict(synth.det_state_dict(m, 81), strict=True)
    ref_ds = oblocks.convbn(cin, planes, 1, stride, 0, 1)
    ref = torch.nn.ModuleDict(dict(conv1=torch.nn.Sequential(oblocks.convbn(cin, planes, 3, stride, 1, dil), torch.nn.ReLU()),
                                   conv2=oblocks.convbn(planes, planes, 3, 1, 1, dil), downsample=ref_ds)).eval()
    ref.load_state_dict(m.state_dict(), strict=True)
    x = synth.det_uniform((2, cin, 20, 28), 82)
    want = ref["conv2"](ref["conv1"](torch.from_numpy(x))) + ref["downsample"](torch.from_numpy(x))
    got = m.cuda()(torch.from_numpy(x).cuda())
    assert _relerr(got.cpu().numpy(), want.numpy()) <= TOL
