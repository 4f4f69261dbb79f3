"""GPU parity: snvc_b200 cost-volume kernels (through the C ABI) vs the CPU oracle."""
import numpy as np
import pytest
import torch

import synth
from oracle import cost_volume as ocv

pytestmark = pytest.mark.gpu


def _bcv():
    from snvc_b200.extension import build_cost_volume as m
    return m


def _rand(shape, seed, dtype=np.float32):
    return np.random.default_rng(seed).standard_normal(shape).astype(dtype)


EDGE_SHIFTS = [0.0, 0.25, 1.0, 8.999, 9.0, 12.0, 1e-4, 3.5, 2.0, 100.0]


@pytest.mark.parametrize("shape,ds", [((2, 3, 6, 10), 1), ((2, 3, 6, 10), 2), ((1, 5, 7, 13), 1), ((3, 8, 12, 40), 1),
                                      ((1, 4, 12, 24), 4)])
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_ncdhw_bit_exact_vs_oracle(shape, ds, dtype):
    N, C, H, W = shape
    l, r = _rand(shape, 1, dtype), _rand(shape, 2, dtype)
    s = np.tile(np.array(EDGE_SHIFTS, dtype=dtype)[None], (N, 1))
    s[-1] = s[-1][::-1]
    want = ocv.forward_c(l, r, s, ds, fma_mode=1)
    got = _bcv().build_cost_volume(torch.from_numpy(l).cuda(), torch.from_numpy(r).cuda(), torch.from_numpy(s).cuda(), ds)
    assert got.shape == want.shape and got.dtype == torch.from_numpy(want).dtype
    assert np.array_equal(got.cpu().numpy(), want)          # bit-exact (integer-index + fma contract)


@pytest.mark.parametrize("shape,ds", [((2, 8, 6, 16), 1), ((1, 32, 5, 40), 1), ((2, 16, 8, 24), 2), ((1, 64, 3, 20), 1),
                                      ((1, 24, 4, 18), 1)])
def test_ndhwc_bf16_equals_rounded_oracle(shape, ds):
    N, C, H, W = shape
    l, r = _rand(shape, 3), _rand(shape, 4)
    s = np.tile(np.float32(EDGE_SHIFTS)[None], (N, 1))
    want = synth.bf16_round(ocv.forward_c(l, r, s, ds, fma_mode=1))          # [N,2C,D,H,W]
    got = _bcv().build_cost_volume_ndhwc_bf16(torch.from_numpy(l).cuda(), torch.from_numpy(r).cuda(),
                                              torch.from_numpy(s).cuda(), ds)
    assert got.dtype == torch.bfloat16 and got.shape == (N, len(EDGE_SHIFTS), H // ds, W // ds, 2 * C)
    got = got.float().permute(0, 4, 1, 2, 3).cpu().numpy()
    assert np.array_equal(got, want)


@pytest.mark.parametrize("shape,ds", [((2, 8, 6, 16), 1), ((1, 32, 5, 40), 1), ((2, 16, 8, 24), 2), ((1, 32, 96, 312), 1),
                                      ((1, 32, 3, 1248), 1)])      # last: the stress volume's 160 KB rows (7 bins per CTA)
def test_split_form_equals_rounded_oracle(shape, ds):
    """snvc_cost_volume_split_fwd: right_vol = channels [C, 2C) of the volume, left_planes = the depth-invariant left
    half on three identical planes; both equal to the bf16-rounded oracle volume (and hence to the unsplit kernel)."""
    N, C, H, W = shape
    l, r = _rand(shape, 5), _rand(shape, 6)
    s = np.tile(np.float32(EDGE_SHIFTS + [40.6, 17.3])[None], (N, 1))
    want = synth.bf16_round(ocv.forward_c(l, r, s, ds, fma_mode=1))          # [N,2C,D,H,W]
    rv, lp = _bcv().build_cost_volume_split_bf16(torch.from_numpy(l).cuda(), torch.from_numpy(r).cuda(),
                                                 torch.from_numpy(s).cuda(), ds)
    D = s.shape[1]
    assert rv.shape == (N, D, H // ds, W // ds, C) and lp.shape == (N, 3, H // ds, W // ds, C)
    assert np.array_equal(rv.float().permute(0, 4, 1, 2, 3).cpu().numpy(), want[:, C:])
    lpn = lp.float().permute(0, 4, 1, 2, 3).cpu().numpy()
    for v in range(3):
        assert np.array_equal(lpn[:, :, v], want[:, :C, 0])
    assert np.array_equal(want[:, :C, 0], want[:, :C, D - 1])                # the oracle's left half is a broadcast over depth


def test_xlow_indices_bit_exact():
    s = np.float32([EDGE_SHIFTS, EDGE_SHIFTS[::-1]])
    for IW, ds in ((10, 1), (312, 1), (24, 2)):
        want = ocv.xlow_c(s, IW, ds)
        got = _bcv().cost_volume_xlow(torch.from_numpy(s).cuda(), IW, ds).cpu().numpy()
        assert np.array_equal(got, want)


def test_kitti_shape_single_pair_vs_oracle():
    """BASELINE.json configs[0] shape: features [1,32,96,312], 48 depth bins."""
    from oracle.global_branch import GlobalGeometry
    g = GlobalGeometry()
    l, r = _rand((1, 32, 96, 312), 10), _rand((1, 32, 96, 312), 11)
    s = g.shifts(1)
    want = ocv.forward_c(l, r, s, 1, fma_mode=1)
    tl, tr, ts = (torch.from_numpy(a).cuda() for a in (l, r, s))
    got = _bcv().build_cost_volume(tl, tr, ts, 1)
    assert np.array_equal(got.cpu().numpy(), want)
    assert np.array_equal(_bcv().cost_volume_xlow(ts, 312).cpu().numpy(), ocv.xlow_c(s, 312))
    got16 = _bcv().build_cost_volume_ndhwc_bf16(tl, tr, ts, 1).float().permute(0, 4, 1, 2, 3).cpu().numpy()
    assert np.array_equal(got16, synth.bf16_round(want))


def test_full_batch_properties():
    """configs[1] size (batch 8): size-independent properties instead of an element-wise oracle."""
    N, C, H, W, D = 8, 32, 96, 312, 48
    gen = torch.Generator(device="cuda").manual_seed(0)
    l = torch.randn((N, C, H, W), device="cuda", generator=gen)
    r = torch.randn((N, C, H, W), device="cuda", generator=gen)
    m = _bcv()
    z = m.build_cost_volume(l, r, torch.zeros((N, D), device="cuda"), 1)
    assert torch.equal(z[:, :C], l[:, :, None].expand(N, C, D, H, W))
    assert torch.equal(z[:, C:], r[:, :, None].expand(N, C, D, H, W))            # shift 0 -> copy
    del z
    k = 7
    i7 = m.build_cost_volume(l, r, torch.full((N, D), float(k), device="cuda"), 1)
    assert torch.equal(i7[:, C:, :, :, k:], r[:, :, None, :, :W - k].expand(N, C, D, H, W - k))
    assert torch.count_nonzero(i7[:, C:, :, :, :k]) == 0                           # integer shift -> exact copy
    del i7
    far = m.build_cost_volume(l, r, torch.full((N, D), W - 0.5, device="cuda"), 1)
    assert torch.count_nonzero(far[:, C:]) == 0                                    # beyond the image -> zeros
    del far
    # linearity in (left, right) and agreement of the two layouts
    s = torch.rand((N, D), device="cuda") * 40
    a = m.build_cost_volume_ndhwc_bf16(l, r, s, 1)
    b = m.build_cost_volume(l, r, s, 1)
    assert torch.equal(a.permute(0, 4, 1, 2, 3), b.to(torch.bfloat16))


def test_backward_vs_oracle():
    N, C, H, W, D = 2, 3, 5, 12, 6
    for ds in (1, 2):
        for dtype in (np.float32, np.float64):
            g = _rand((N, 2 * C, D, H, W), 5, dtype)
            s = (np.random.default_rng(6).random((N, D)) * 9).astype(dtype)
            s[0, 0], s[0, 1] = 0.0, 3.0
            wl, wr = ocv.backward_c(g.astype(np.float64), s.astype(np.float64), ds)
            l = torch.zeros((N, C, H * ds, W * ds), dtype=torch.from_numpy(g).dtype, device="cuda", requires_grad=True)
            r = torch.zeros_like(l, requires_grad=True)
            out = _bcv().build_cost_volume(l, r, torch.from_numpy(s).cuda(), ds)
            out.backward(torch.from_numpy(g).cuda())
            tol = 1e-5 if dtype == np.float32 else 1e-12
            # double shifts differ from float shifts only for the f32 run; compare against the matching oracle
            wl32, wr32 = ocv.backward_c(g, s, ds)
            assert np.max(np.abs(l.grad.cpu().numpy() - wl32)) <= tol * max(1, np.abs(wl32).max())
            assert np.max(np.abs(r.grad.cpu().numpy() - wr32)) <= tol * max(1, np.abs(wr32).max())


def test_error_behaviour():
    m = _bcv()
    l = torch.zeros((1, 2, 4, 4))
    with pytest.raises(RuntimeError, match="Not implemented on the CPU"):        # BuildCostVolume.cpp:26
        m.build_cost_volume(l, l, torch.zeros((1, 3)), 1)
    lc = l.cuda()
    with pytest.raises(RuntimeError):                                             # .cu:216-218
        m.build_cost_volume(lc, torch.zeros((1, 2, 4, 5), device="cuda"), torch.zeros((1, 3), device="cuda"), 1)
    with pytest.raises(RuntimeError):                                             # .cu:219-220
        m.build_cost_volume(lc, lc, torch.zeros((2, 3), device="cuda"), 1)
    with pytest.raises(AssertionError):                                           # __init__.py:12
        m.build_cost_volume(lc, lc, -torch.ones((1, 3), device="cuda"), 1)
    empty = m.build_cost_volume(lc[:0], lc[:0], torch.zeros((0, 3), device="cuda"), 1)   # .cu:235-238
    assert empty.shape == (0, 4, 3, 4, 4)
