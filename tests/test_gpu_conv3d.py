"""GPU parity: the tcgen05 implicit-GEMM conv3d (C ABI snvc_conv3d_fwd) vs the oracle arithmetic
(torch CPU fp32 conv3d / conv_transpose3d + eval-BN affine, as in oracle/blocks.py).

Inputs and weights are exactly representable in bf16 (tests/golden/synth.py), so with fp32
output the only difference is fp32 summation order (tolerance 1e-4 of max|ref|); with bf16 output
the additional error is one bf16 rounding (<= 2^-8 relative; north_star bar: 1e-2)."""
import numpy as np
import pytest
from conftest import set_opt
import torch
import torch.nn.functional as TF

import synth

pytestmark = pytest.mark.gpu


def _relerr(a, b):
    return float(np.max(np.abs(a - b)) / max(1e-30, np.max(np.abs(b))))


def _case(Cin, Cout, k, stride, pad, dil, transposed, dhw, N=1, relu=False, residual_mode=0, sigmoid=False, seed=0):
    from snvc_b200 import functional as F
    from snvc_b200.conv import PackedConv3d
    D, H, W = dhw
    x = synth.det_uniform((N, Cin, D, H, W), seed + 1)
    wshape = (Cin, Cout, k, k, k) if transposed else (Cout, Cin, k, k, k)
    taps = (k ** 3) / (8.0 if transposed else 1.0)
    a = float(np.sqrt(3.0 / (Cin * taps)))
    w = synth.det_uniform(wshape, seed + 2, -a, a)
    scale = synth.det_uniform((Cout,), seed + 3, 0.6, 1.4, bf16=False)
    bias = synth.det_uniform((Cout,), seed + 4, -0.2, 0.2, bf16=False)
    tx, tw = torch.from_numpy(x), torch.from_numpy(w)
    if transposed:
        ref = TF.conv_transpose3d(tx, tw, stride=2, padding=1, output_padding=1)
    else:
        ref = TF.conv3d(tx, tw, stride=stride, padding=pad, dilation=dil)
    ref = ref * torch.from_numpy(scale).view(1, -1, 1, 1, 1) + torch.from_numpy(bias).view(1, -1, 1, 1, 1)
    res = None
    if residual_mode:
        res = synth.det_uniform(tuple(ref.shape), seed + 5)
        if residual_mode == 1:
            ref = ref + torch.from_numpy(res)
    if relu:
        ref = torch.relu(ref)
    if residual_mode == 2:
        ref = ref + torch.from_numpy(res)
    if sigmoid:
        ref = torch.sigmoid(ref)
    ref = ref.numpy()

    import types
    # folded affine expressed as an eval BatchNorm (weight=scale, bias=bias, mean=0, var=1, eps=0)
    bn = types.SimpleNamespace(eps=0.0, weight=torch.from_numpy(scale).cuda(), bias=torch.from_numpy(bias).cuda(),
                               running_mean=torch.zeros(Cout, device="cuda"), running_var=torch.ones(Cout, device="cuda"))
    conv = PackedConv3d(tw.cuda(), bn, transposed=transposed, stride=stride, pad=pad, dilation=dil)
    x16 = F.to_ndhwc_bf16(tx.cuda())
    r16 = F.to_ndhwc_bf16(torch.from_numpy(res).cuda()) if res is not None else None
    outs = {}
    for dt in (torch.float32, torch.bfloat16):
        y = conv(x16, relu=relu, residual=r16, residual_mode=residual_mode, sigmoid=sigmoid, out_dtype=dt)
        torch.cuda.synchronize()
        outs[dt] = y.float().permute(0, 4, 1, 2, 3).cpu().numpy()
    assert outs[torch.float32].shape == ref.shape
    e32, e16 = _relerr(outs[torch.float32], ref), _relerr(outs[torch.bfloat16], ref)
    assert e32 <= 1e-4, f"fp32-out rel err {e32}"
    assert e16 <= 1e-2 and e16 <= 6e-3, f"bf16-out rel err {e16}"
    return e32, e16


def test_conv_tiny_first():
    """Smallest possible launch: a protocol bug traps here (bounded waits) instead of later."""
    _case(32, 32, 1, 1, 0, 1, False, (2, 8, 8))


@pytest.mark.parametrize("Cin,Cout", [(32, 32), (64, 32), (64, 64), (32, 64), (16, 16)])
def test_conv_3x3x3_s1(Cin, Cout):
    _case(Cin, Cout, 3, 1, 1, 1, False, (6, 10, 20), N=2, relu=True)


@pytest.mark.parametrize("dhw", [(8, 16, 24), (5, 7, 9), (12, 24, 78)])
def test_conv_3x3x3_s2(dhw):
    _case(32, 64, 3, 2, 1, 1, False, dhw, relu=True)
    _case(64, 64, 3, 2, 1, 1, False, dhw, relu=True)


def test_conv_s2_plane_march_variants(monkeypatch):
    """Stride-2 pair-row kernel (Cin = 32, even extents): both Cout variants, batch, residual epilogues, a real
    KITTI-width slab, and -- with the grid clamped to 2 CTAs -- many work units per CTA so that the TMEM accumulator
    ring and the plane ring wrap around several times."""
    _case(32, 32, 3, 2, 1, 1, False, (8, 16, 24), N=2, relu=True)
    _case(32, 64, 3, 2, 1, 1, False, (6, 20, 64), N=2, relu=True, residual_mode=1)
    _case(32, 64, 3, 2, 1, 1, False, (4, 96, 312), relu=True)
    set_opt(monkeypatch, "SNVC_CONV_MAXGRID", "2")
    _case(32, 64, 3, 2, 1, 1, False, (40, 16, 62), relu=True)
    _case(32, 32, 3, 2, 1, 1, False, (72, 16, 30), relu=True, residual_mode=2)


@pytest.mark.parametrize("Cin,Cout,mode", [(64, 64, 1), (64, 32, 1), (64, 64, 0)])
def test_deconv_k3_s2(Cin, Cout, mode):
    _case(Cin, Cout, 3, 2, 1, 1, True, (3, 6, 10), N=2, relu=(mode == 1 and Cout == 64), residual_mode=mode)


def test_deconv_staged_tiles_many_units_per_cta(monkeypatch):
    """Fused transposed conv with its staged (TMA load residual -> in-place update -> TMA store) epilogue: tiles
    clipped at the H / W edges, several work units per CTA (grid clamped) so that both tile buffers, the plane ring
    and the TMEM double buffer wrap, every residual mode."""
    set_opt(monkeypatch, "SNVC_CONV_MAXGRID", "3")
    _case(64, 32, 3, 2, 1, 1, True, (10, 20, 40), N=2, residual_mode=1)                 # hourglass conv6 (+ out residual)
    _case(64, 64, 3, 2, 1, 1, True, (6, 12, 39), relu=True, residual_mode=1)            # hourglass conv5: relu(bn + pre)
    _case(64, 32, 3, 2, 1, 1, True, (5, 9, 17), N=3)                                    # no residual
    _case(32, 32, 3, 2, 1, 1, True, (4, 8, 16), relu=True, residual_mode=2)


def test_conv_staged_tma_store_epilogue(monkeypatch):
    """Opt-in epilogue of the kd-fused kernel: rows staged in a swizzled shared-memory tile, one bulk tensor store."""
    set_opt(monkeypatch, "SNVC_CONV_STORE", "staged")
    _case(32, 32, 3, 1, 1, 1, False, (6, 10, 40), N=2, relu=True, residual_mode=1)
    _case(64, 32, 3, 1, 1, 1, False, (4, 96, 312), relu=True)
    _case(64, 64, 3, 1, 1, 1, False, (5, 12, 70), relu=True)
    _case(16, 16, 3, 1, 1, 1, False, (6, 10, 20), N=2, relu=True)


def test_conv_residual_modes_and_sigmoid():
    _case(32, 32, 3, 1, 1, 1, False, (4, 8, 16), residual_mode=1)                      # dres1.1: bn + x
    _case(32, 32, 3, 1, 1, 1, False, (4, 8, 16), relu=True, residual_mode=1)           # hourglass conv2 (+postsqu)
    _case(32, 32, 5, 1, 2, 1, False, (6, 10, 12), relu=True, residual_mode=2)          # vernier conv2: relu(bn) + x
    _case(32, 1, 3, 1, 1, 1, False, (4, 8, 16), sigmoid=True)                          # fg_cls_head.2 (Cout = 1)


def test_conv_instance_kernels():
    _case(64, 32, 1, 1, 0, 1, False, (4, 8, 16), relu=True)                            # vimg_feat 1^3
    _case(64, 32, 7, 1, 3, 1, False, (8, 12, 16), relu=True)                           # conv1 7^3
    _case(32, 32, 5, 1, 4, 2, False, (10, 12, 16), relu=True, residual_mode=2)         # conv3 5^3 dilation 2


def test_conv_large_kernel_plane_march(monkeypatch):
    """5^3 / 7^3 plane march with streamed weights (instance conv1-conv3): more than 16 planes (the TMEM accumulator
    ring wraps), depths that are not multiples of the 4-plane group, several tile columns per CTA (grid clamped), both
    dilations (dilation 2 uses the parity-split ring; an odd depth falls back to the per-tap kernel)."""
    set_opt(monkeypatch, "SNVC_CONV_MAXGRID", "2")
    _case(64, 32, 7, 1, 3, 1, False, (18, 9, 40), relu=True)                               # conv1
    _case(32, 32, 5, 1, 2, 1, False, (21, 12, 60), N=2, relu=True, residual_mode=2)        # conv2
    _case(32, 32, 5, 1, 4, 2, False, (22, 10, 50), relu=True, residual_mode=2)             # conv3 (dilation 2)
    _case(32, 32, 5, 1, 4, 2, False, (7, 10, 20), relu=True)                               # odd depth -> per-tap path
    _case(32, 32, 7, 1, 3, 1, False, (5, 8, 30), relu=True)
    _case(64, 32, 5, 1, 2, 1, False, (9, 8, 24), relu=True)


def test_conv_kw_kd_fused_plane_march(monkeypatch):
    """v7 kernel (kw and kd taps fused into N, 5-block TMEM ring, shuffle epilogue): both row pitches (W = 60 -> 32,
    W = 40 / 14 -> 16), Cin 32 / 64, the 64 -> 64 output slices, every residual mode, ragged H / W edges, depths 1 and 2,
    and -- grid clamped -- several tile columns per CTA with depths that are not multiples of 5, so the accumulator
    ring wraps at every phase."""
    set_opt(monkeypatch, "SNVC_CONV_MODE", "kw")
    set_opt(monkeypatch, "SNVC_CONV_MAXGRID", "2")
    _case(32, 32, 3, 1, 1, 1, False, (13, 10, 60), N=2, relu=True, residual_mode=1)
    _case(64, 32, 3, 1, 1, 1, False, (7, 16, 40), relu=True)
    _case(64, 64, 3, 1, 1, 1, False, (6, 12, 44), relu=True, residual_mode=1)
    _case(32, 32, 3, 1, 1, 1, False, (9, 9, 14), N=3, residual_mode=2, relu=True)
    _case(32, 32, 3, 1, 1, 1, False, (1, 8, 30))
    _case(64, 32, 3, 1, 1, 1, False, (2, 5, 33), residual_mode=1)
    _case(32, 32, 3, 1, 1, 1, False, (24, 4, 31), relu=True)


def test_conv_cta_pair_plane_march(monkeypatch):
    """v8 kernel (the default for 32-channel output slices): tcgen05.mma.cta_group::2 over a cluster of two CTAs, each
    marching its own tile column with half of every weight slab; 14-block accumulator ring with two mirror blocks.
    Depths beyond 14 (the ring and its mirrors wrap), an ODD number of tile columns (the last follower is a ghost),
    all three row pitches (W = 60 -> 32, W = 40 -> 16, W = 124 -> 64), Cin 32 / 64, the 64 -> 64 output slices, every
    residual mode, ragged edges, depth 1 / 2; first with the natural grid, then clamped to 1 and 2 pairs so that each
    pair walks many columns."""
    cases = [dict(a=(32, 32, 3, 1, 1, 1, False, (24, 4, 30)), k=dict(N=3, relu=True)),                     # 3 columns
             dict(a=(32, 32, 3, 1, 1, 1, False, (17, 10, 60)), k=dict(N=2, relu=True, residual_mode=1)),
             dict(a=(64, 32, 3, 1, 1, 1, False, (15, 16, 40)), k=dict(relu=True)),
             dict(a=(64, 64, 3, 1, 1, 1, False, (6, 12, 44)), k=dict(relu=True, residual_mode=1)),
             dict(a=(32, 32, 3, 1, 1, 1, False, (9, 9, 14)), k=dict(N=3, residual_mode=2, relu=True)),
             dict(a=(64, 32, 3, 1, 1, 1, False, (30, 3, 124)), k=dict(relu=True, residual_mode=1)),
             dict(a=(32, 32, 3, 1, 1, 1, False, (1, 8, 30)), k=dict()),
             dict(a=(64, 32, 3, 1, 1, 1, False, (2, 5, 33)), k=dict(residual_mode=1))]
    for c in cases[:3]:
        _case(*c["a"], **c["k"])
    for clamp in ("1", "2"):
        set_opt(monkeypatch, "SNVC_CONV_MAXGRID", clamp)
        for c in cases:
            _case(*c["a"], **c["k"])
    # pitch-42 tiles (3 rows x 42 columns = 126 of the 128 MMA rows; chosen when they waste fewer rows): W = 40 / 78 with
    # heights that are not multiples of 3, both Cin, residual, several columns per pair
    _case(32, 32, 3, 1, 1, 1, False, (17, 9, 40), N=2, relu=True, residual_mode=1)
    _case(64, 32, 3, 1, 1, 1, False, (5, 10, 78), relu=True)
    _case(64, 64, 3, 1, 1, 1, False, (6, 7, 118), N=2, relu=True, residual_mode=2)
    # depth-split tail units: the pair-columns left over after the whole rounds are cut into 2 / 3 / 4 depth ranges
    set_opt(monkeypatch, "SNVC_CONV_MAXGRID", "2")                                            # 3 pairs on 2 clusters -> 2 ranges
    _case(64, 32, 3, 1, 1, 1, False, (15, 16, 40), relu=True, residual_mode=1)
    set_opt(monkeypatch, "SNVC_CONV_MAXGRID", "3")                                            # 4 pairs on 3 clusters -> 3 ranges
    _case(32, 32, 3, 1, 1, 1, False, (13, 8, 60), N=2, relu=True, residual_mode=2)
    set_opt(monkeypatch, "SNVC_CONV_MAXGRID", "4")                                            # 5 pairs on 4 clusters -> 4 ranges
    _case(32, 32, 3, 1, 1, 1, False, (16, 4, 300), relu=True, residual_mode=1)
    _case(64, 64, 3, 1, 1, 1, False, (18, 4, 270), relu=True)                               # 9 columns: ghost follower + ranges


def _split_case(dhw, N=1, seed=0):
    """First trunk layer on the split cost volume: a 64 -> 32 3x3x3 conv whose first 32 input channels do not vary
    with depth == conv of the other 32 channels + a depth-invariant addend from a 3-plane conv of the constant half."""
    from snvc_b200 import functional as F
    from snvc_b200.conv import PackedConv3d
    import types
    D, H, W = dhw
    left = synth.det_uniform((N, 32, 1, H, W), seed + 1)
    right = synth.det_uniform((N, 32, D, H, W), seed + 2)
    a = float(np.sqrt(3.0 / (64 * 27)))
    w = synth.det_uniform((32, 64, 3, 3, 3), seed + 3, -a, a)
    scale = synth.det_uniform((32,), seed + 4, 0.6, 1.4, bf16=False)
    bias = synth.det_uniform((32,), seed + 5, -0.2, 0.2, bf16=False)
    x = torch.cat([torch.from_numpy(left).expand(N, 32, D, H, W), torch.from_numpy(right)], 1)
    ref = TF.conv3d(x, torch.from_numpy(w), padding=1)
    ref = torch.relu(ref * torch.from_numpy(scale).view(1, -1, 1, 1, 1) + torch.from_numpy(bias).view(1, -1, 1, 1, 1)).numpy()
    bn = types.SimpleNamespace(eps=0.0, weight=torch.from_numpy(scale).cuda(), bias=torch.from_numpy(bias).cuda(),
                               running_mean=torch.zeros(32, device="cuda"), running_var=torch.ones(32, device="cuda"))
    tw = torch.from_numpy(w).cuda()
    p_left = PackedConv3d(tw[:, :32].contiguous(), None, stride=1, pad=1)
    p_right = PackedConv3d(tw[:, 32:].contiguous(), bn, stride=1, pad=1)
    left3 = F.to_ndhwc_bf16(torch.from_numpy(left).expand(N, 32, 3, H, W).contiguous().cuda())
    xr = F.to_ndhwc_bf16(torch.from_numpy(right).cuda())
    addend = p_left(left3, out_dtype=torch.float32)
    for dt, tol in ((torch.float32, 1e-4), (torch.bfloat16, 6e-3)):
        y = p_right(xr, relu=True, addend=addend, out_dtype=dt)
        torch.cuda.synchronize()
        e = _relerr(y.float().permute(0, 4, 1, 2, 3).cpu().numpy(), ref)
        assert e <= tol, f"{dt}: rel err {e}"


def test_conv_depth_invariant_addend(monkeypatch):
    """snvc_conv3d_fwd_addend (CTA-pair kernel): depths 2 and 3 (every plane is an edge plane or the only interior one),
    a depth beyond the 14-block ring, an odd column count, all three row pitches, ragged edges, clamped grid."""
    _split_case((2, 8, 30))
    _split_case((3, 5, 33), N=2)
    _split_case((17, 10, 60), N=2)
    _split_case((6, 9, 40), N=2)                                                            # pitch-42 tiles
    set_opt(monkeypatch, "SNVC_CONV_MAXGRID", "3")                                            # 4 pairs on 3 clusters -> 3 depth ranges
    _split_case((13, 8, 60), N=2)
    set_opt(monkeypatch, "SNVC_CONV_MAXGRID", "4")                                            # 5 pairs on 4 clusters -> 4 depth ranges
    _split_case((16, 4, 300))
    set_opt(monkeypatch, "SNVC_CONV_MAXGRID", "1")
    _split_case((20, 4, 30), N=3)
    _split_case((9, 16, 40))
    _split_case((6, 3, 124))
    from snvc_b200.conv import PackedConv3d
    p1 = PackedConv3d(torch.zeros(32, 32, 3, 3, 3, device="cuda"), None, stride=1, pad=1)
    x = torch.zeros(1, 1, 4, 8, 32, dtype=torch.bfloat16, device="cuda")
    with pytest.raises(RuntimeError, match="D >= 2"):                               # a single depth plane is not supported
        p1(x, addend=torch.zeros(1, 3, 4, 8, 32, device="cuda"))


def test_conv_kd_fused_kernel_still_green(monkeypatch):
    """SNVC_CONV_MODE=kd keeps the v3 kernel (kd taps only) reachable for A/B runs; it also serves Cout = 16 / 64."""
    set_opt(monkeypatch, "SNVC_CONV_MODE", "kd")
    set_opt(monkeypatch, "SNVC_CONV_MAXGRID", "2")
    _case(32, 32, 3, 1, 1, 1, False, (13, 10, 60), N=2, relu=True, residual_mode=1)
    _case(64, 32, 3, 1, 1, 1, False, (7, 16, 40), relu=True)


def test_conv_kitti_level_shapes():
    """One slab of the global trunk's real W/H (W=312 is not a multiple of the tile)."""
    _case(64, 32, 3, 1, 1, 1, False, (4, 96, 312), relu=True)
    _case(32, 32, 3, 1, 1, 1, False, (3, 96, 312), relu=True, residual_mode=1)
