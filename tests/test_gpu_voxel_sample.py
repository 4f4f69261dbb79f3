"""GPU parity: ROI voxel sampling (A3) and frustum-to-voxel lift (A4) vs the CPU oracle."""
import numpy as np
import pytest
from conftest import set_opt
import torch

import synth
from oracle import global_branch as ogb
from oracle import grid_sample as ogs

pytestmark = pytest.mark.gpu


def _F():
    from snvc_b200 import functional
    return functional


def _relerr(a, b):
    return float(np.max(np.abs(a - b)) / max(1e-30, np.max(np.abs(b))))


def _roi_case(N=2, C=32, Hf=16, Wf=16, grid=(4, 6, 10), res=(64, 64), seed=0):
    nh, nw, nl = grid
    P = nh * nw * nl
    lf = synth.det_uniform((N, C, Hf, Wf), seed + 1)
    rf = synth.det_uniform((N, C, Hf, Wf), seed + 2)
    gl = synth.det_uniform((N, 2, P), seed + 3, -0.1 * res[1], 1.1 * res[1], bf16=False)
    gr = synth.det_uniform((N, 2, P), seed + 4, -0.1 * res[1], 1.1 * res[1], bf16=False)
    return lf, rf, gl, gr, (nh, nw, nl), res


@pytest.mark.parametrize("kw", [dict(), dict(N=1, C=8, Hf=9, Wf=12, grid=(3, 5, 7), res=(48, 36)),
                                dict(N=3, C=64, Hf=16, Wf=16, grid=(2, 4, 8))])
def test_roi_sample_vs_oracle(kw):
    lf, rf, gl, gr, (nh, nw, nl), res = _roi_case(**kw)
    want = ogs.roi_voxel_sample(lf, rf, gl, gr, nh, nw, nl, res)             # [N,2C,nh,nw,nl]
    t = [torch.from_numpy(a).cuda() for a in (lf, rf, gl, gr)]
    keep = [a.clone() for a in t]
    got = _F().roi_voxel_sample(*t, res).reshape(want.shape).cpu().numpy()
    assert _relerr(got, want) <= 1e-5                                         # fp32 bar (north_star)
    assert np.array_equal(got, want)                                          # and in fact bit-exact
    for a, b in zip(t, keep):
        assert torch.equal(a, b)                                              # inputs untouched
    got16 = _F().roi_voxel_sample(*t, res, out_dtype=torch.bfloat16, layout="NDHWC")
    N = lf.shape[0]
    got16 = got16.float().reshape(N, nh, nw, nl, -1).permute(0, 4, 1, 2, 3).cpu().numpy()
    # bf16 product path (v4 kernel): one fused FMA per corner instead of separately rounded mul + add -> within one bf16
    # ulp of the rounded oracle, identical zero support (masked corners are zero weights, fully masked points give +0)
    r16 = synth.bf16_round(want)
    assert np.max(np.abs(got16 - r16)) <= 2.0 ** -7 * np.max(np.abs(want))
    assert np.mean(got16 == r16) > 0.99 and np.array_equal(got16 == 0, r16 == 0)
    got32 = _F().roi_voxel_sample(*t, res, out_dtype=torch.float32, layout="NDHWC")
    assert np.array_equal(got32.reshape(N, nh, nw, nl, -1).permute(0, 4, 1, 2, 3).cpu().numpy(), want)


def test_roi_bf16_product_path_rounds_features_once():
    """The bf16-output product kernel samples a bf16 channels-last COPY of the features (half the L1 traffic): with
    features that are not bf16-representable the result is the sampling of the rounded features (within one bf16 ulp of
    that oracle) and stays far inside the 1e-2 bar against the unrounded fp32 oracle; fp32 outputs are untouched by this."""
    N, C, Hf, Wf, (nh, nw, nl), res = 2, 32, 16, 16, (4, 6, 10), (64, 64)
    P = nh * nw * nl
    lf, rf = synth.det_uniform((N, C, Hf, Wf), 61, bf16=False), synth.det_uniform((N, C, Hf, Wf), 62, bf16=False)
    gl = synth.det_uniform((N, 2, P), 63, -6.4, 70.4, bf16=False)
    gr = synth.det_uniform((N, 2, P), 64, -6.4, 70.4, bf16=False)
    assert not np.array_equal(lf, synth.bf16_round(lf))
    want = ogs.roi_voxel_sample(lf, rf, gl, gr, nh, nw, nl, res)
    want_r = ogs.roi_voxel_sample(synth.bf16_round(lf), synth.bf16_round(rf), gl, gr, nh, nw, nl, res)
    t = [torch.from_numpy(a).cuda() for a in (lf, rf, gl, gr)]
    got = _F().roi_voxel_sample(*t, res, out_dtype=torch.bfloat16, layout="NDHWC")
    got = got.float().reshape(N, nh, nw, nl, -1).permute(0, 4, 1, 2, 3).cpu().numpy()
    r16 = synth.bf16_round(want_r)
    assert np.max(np.abs(got - r16)) <= 2.0 ** -7 * np.max(np.abs(want_r)) and np.mean(got == r16) > 0.99
    assert _relerr(got, want) <= 1e-2 and np.array_equal(got == 0, want == 0)
    exact = _F().roi_voxel_sample(*t, res).reshape(want.shape).cpu().numpy()
    assert np.array_equal(exact, want)


def test_roi_indices_bit_exact():
    lf, rf, gl, gr, grid, res = _roi_case(N=2, grid=(8, 16, 24))
    Hf = Wf = 16
    idx, mask = _F().roi_voxel_sample_indices(torch.from_numpy(gl).cuda(), Hf, Wf, res)
    gx = ogs.roi_normalize(gl[:, 0], res[1])
    gy = ogs.roi_normalize(gl[:, 1], res[0])
    _, _, x0, y0 = ogs.corners_2d(gx, gy, Wf, Hf, False)
    assert np.array_equal(idx.cpu().numpy()[..., 0], x0.astype(np.int32))
    assert np.array_equal(idx.cpu().numpy()[..., 1], y0.astype(np.int32))
    inb = lambda x, y: ((x >= 0) & (x < Wf) & (y >= 0) & (y < Hf)).astype(np.uint8)
    want_mask = inb(x0, y0) | (inb(x0 + 1, y0) << 1) | (inb(x0, y0 + 1) << 2) | (inb(x0 + 1, y0 + 1) << 3)
    assert np.array_equal(mask.cpu().numpy(), want_mask)
    assert 0 < (want_mask == 15).mean() < 1 and (want_mask == 0).any()        # exercises the zero padding


def test_roi_sample_matches_reference_golden(golden):
    """Against outputs of the reference's own VernierScale.construct_voxel (tests/golden)."""
    g = golden("vernier_bev3")
    nh, nw, nl = 16, 32, 48
    P = nh * nw * nl
    lf, rf = synth.det_uniform((1, 32, 16, 16), 201), synth.det_uniform((1, 32, 16, 16), 202)
    gl = synth.det_uniform((1, 2, P), 203, -6.4, 70.4, bf16=False)
    gr = synth.det_uniform((1, 2, P), 204, -6.4, 70.4, bf16=False)
    vox = _F().roi_voxel_sample(*[torch.from_numpy(a).cuda() for a in (lf, rf, gl, gr)], (64, 64))
    vox = vox.reshape(1, 64, nh, nw, nl).cpu().numpy()
    assert _relerr(vox[:, :, ::2, ::4, ::4], g["voxel_sub"]) <= 1e-5
    np.testing.assert_allclose(vox.astype(np.float64).sum(axis=(0, 2, 3, 4)), g["voxel_chan_sum"], rtol=0, atol=2e-2)


def _small_geom(ac):
    return ogb.GlobalGeometry(IH=48, IW=160, D=12, depth_min=2.0, depth_max=21.2, X_MIN=-6.0, X_MAX=6.0, Y_MIN=-1.0,
                              Y_MAX=2.0, Z_MIN=2.0, Z_MAX=20.0, VOXEL_X_SIZE=0.4, VOXEL_Y_SIZE=0.5, VOXEL_Z_SIZE=0.6,
                              align_corners=ac,
                              P=np.array([[90.0, 0, 80.0, 5.6], [0, 90.0, 22.0, 0.03], [0, 0, 1.0, 0.0003]], np.float32))


@pytest.mark.parametrize("ac", [True, False])
def test_lift_vs_oracle(ac):
    geom = _small_geom(ac)
    N, C = 2, 16
    D, H, W = geom.D, geom.IH // 4, geom.IW // 4
    vol = synth.det_uniform((N, C, D, H, W), 7)
    Ps = np.stack([geom.P, geom.P * np.float32([[1.0], [1.01], [1.0]])]).astype(np.float32)
    want, wvalid = ogb.frustum_lift(vol, Ps, geom)
    zs, ys, xs = (torch.from_numpy(a).cuda() for a in ogb.voxel_centres(geom))
    tv, tp = torch.from_numpy(vol).cuda(), torch.from_numpy(Ps).cuda()
    F = _F()
    got, valid = F.frustum_lift(tv, tp, zs, ys, xs, geom.cv_ranges(), ac, return_valid=True)
    assert 0.05 < wvalid.mean() < 0.95
    assert np.array_equal(valid.cpu().numpy().astype(bool), wvalid)
    assert _relerr(got.cpu().numpy(), want) <= 1e-5
    assert np.array_equal(got.cpu().numpy(), want)
    # channels-last bf16 volume (exactly representable here) -> same numbers
    v16 = F.to_ndhwc_bf16(tv)
    assert torch.equal(F.to_ncdhw_f32(v16), tv)
    got2 = F.frustum_lift(v16, tp, zs, ys, xs, geom.cv_ranges(), ac, layout_in="NDHWC", out_dtype=torch.float32,
                          layout_out="NCDHW")
    assert np.array_equal(got2.cpu().numpy(), want)
    got3 = F.frustum_lift(v16, tp, zs, ys, xs, geom.cv_ranges(), ac, layout_in="NDHWC", out_dtype=torch.float32,
                          layout_out="NDHWC")
    assert np.array_equal(got3.permute(0, 4, 1, 2, 3).cpu().numpy(), want)
    got4 = F.frustum_lift(v16, tp, zs, ys, xs, geom.cv_ranges(), ac, layout_in="NDHWC")
    assert got4.dtype == torch.bfloat16          # product path: FMA accumulation, then one bf16 rounding
    g4 = got4.float().permute(0, 4, 1, 2, 3).cpu().numpy()
    assert np.max(np.abs(g4 - want)) <= 2.0 ** -8 * np.max(np.abs(want)) + 1e-6
    assert np.array_equal(g4 == 0, want == 0)    # identical support (validity mask + zero padding)


@pytest.mark.parametrize("ac", [True, False])
def test_lift_indices_bit_exact_kitti_geometry(ac):
    geom = ogb.GlobalGeometry(align_corners=ac)
    zs, ys, xs = ogb.voxel_centres(geom)
    cv = geom.cv_ranges()
    grid, wvalid = ogb.lift_grid(zs, ys, xs, geom.P, cv)
    D, H, W = geom.D, geom.IH // 4, geom.IW // 4
    _, _, _, x0, y0, z0 = ogs.corners_3d(grid[..., 0], grid[..., 1], grid[..., 2], W, H, D, ac)
    idx, valid = _F().frustum_lift_indices(torch.from_numpy(geom.P[None]).cuda(), *[torch.from_numpy(a).cuda() for a in (zs, ys, xs)],
                                           cv, (D, H, W), ac)
    idx, valid = idx.cpu().numpy()[0], valid.cpu().numpy()[0].astype(bool)
    assert np.array_equal(valid, wvalid)
    assert 0.5 < wvalid.mean() < 0.65                     # ~58 % of centres in frustum
    for k, ref in enumerate((x0, y0, z0)):
        assert np.array_equal(idx[..., k][wvalid], ref.astype(np.int32)[wvalid])


@pytest.mark.parametrize("C", [8, 16, 32, 64, 128])
def test_roi_sample_kernel_generations_agree(C, monkeypatch):
    """The exact ROI sampler (v3: both views per round, 256-bit corner loads) must be bit-identical to v2
    (SNVC_ROI_MODE=coop1) and to the one-thread-per-(point, view, 8 channels) kernel (SNVC_ROI_MODE=thread), with a
    point count that is not a multiple of 16."""
    lf, rf, gl, gr, (nh, nw, nl), res = _roi_case(N=2, C=C, Hf=12, Wf=20, grid=(3, 7, 11), res=(48, 80), seed=C)
    assert (2 * nh * nw * nl) % 16 != 0
    t = [torch.from_numpy(a).cuda() for a in (lf, rf, gl, gr)]
    fast = _F().roi_voxel_sample(*t, res, out_dtype=torch.bfloat16, layout="NDHWC")      # default: the v4 product kernel
    for od in (torch.bfloat16, torch.float32):
        outs = []
        for mode in ("thread", "coop1", "v3"):
            set_opt(monkeypatch, "SNVC_ROI_MODE", mode)
            outs.append(_F().roi_voxel_sample(*t, res, out_dtype=od, layout="NDHWC"))
        view = torch.int16 if od == torch.bfloat16 else torch.int32
        assert torch.equal(outs[0].view(view), outs[1].view(view))
        assert torch.equal(outs[0].view(view), outs[2].view(view))
        assert outs[2].float().abs().max().item() > 0
        if od == torch.bfloat16:        # v4 (fused FMA) vs the exact kernels: <= 1 bf16 ulp, same zero support
            a, b = fast.float(), outs[2].float()
            assert (a - b).abs().max().item() <= 2.0 ** -7 * b.abs().max().item()
            assert (a == b).float().mean().item() > 0.99 and torch.equal(a == 0, b == 0)


@pytest.mark.parametrize("C", [16, 32, 64])
def test_lift_cooperative_kernel_matches_thread_per_voxel_kernel(C, monkeypatch):
    """The product lift (C/8 lanes per voxel, set-up shared by warp shuffles) must be bit-identical to the
    one-thread-per-voxel kernel (SNVC_LIFT_MODE=thread), incl. a voxel count that is not a multiple of 32."""
    geom = ogb.GlobalGeometry(IH=48, IW=160, D=12, depth_min=2.0, depth_max=21.2, X_MIN=-6.2, X_MAX=6.2, Y_MIN=-1.0,
                              Y_MAX=2.0, Z_MIN=2.0, Z_MAX=20.0, VOXEL_X_SIZE=0.4, VOXEL_Y_SIZE=0.6, VOXEL_Z_SIZE=0.6,
                              P=np.array([[90.0, 0, 80.0, 5.6], [0, 90.0, 22.0, 0.03], [0, 0, 1.0, 0.0003]], np.float32))
    N = 3
    D, H, W = geom.D, geom.IH // 4, geom.IW // 4
    F = _F()
    vol = F.to_ndhwc_bf16(torch.from_numpy(synth.det_uniform((N, C, D, H, W), 11)).cuda())
    zs, ys, xs = (torch.from_numpy(a).cuda() for a in ogb.voxel_centres(geom))
    assert (N * zs.numel() * ys.numel() * xs.numel()) % 32 != 0
    Ps = torch.from_numpy(np.stack([geom.P] * N)).cuda()
    # v4, the default bf16 kernel (zero-weight masking, set-up through shared memory): same numbers as the others
    fast, fv = F.frustum_lift(vol, Ps, zs, ys, xs, geom.cv_ranges(), True, layout_in="NDHWC", out_dtype=torch.bfloat16,
                              return_valid=True)
    for od in (torch.bfloat16, torch.float32):
        set_opt(monkeypatch, "SNVC_LIFT_MODE", "thread")
        want, wv = F.frustum_lift(vol, Ps, zs, ys, xs, geom.cv_ranges(), True, layout_in="NDHWC", out_dtype=od,
                                  return_valid=True)
        if od == torch.bfloat16:
            assert torch.equal(fv, wv) and torch.equal(fast.view(torch.int16), want.view(torch.int16))
        set_opt(monkeypatch, "SNVC_LIFT_MODE", "coop")
        got, gv = F.frustum_lift(vol, Ps, zs, ys, xs, geom.cv_ranges(), True, layout_in="NDHWC", out_dtype=od,
                                 return_valid=True)
        assert torch.equal(gv, wv) and 0.05 < wv.float().mean().item() < 0.95
        assert torch.equal(got.view(torch.int16 if od == torch.bfloat16 else torch.int32),
                           want.view(torch.int16 if od == torch.bfloat16 else torch.int32))
