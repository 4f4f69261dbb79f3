"""GPU parity PIN: snvc_b200 kernels (through the C ABI) vs the REFERENCE's own CUDA ops, compiled from
/root/reference by oracle/build_ref.py into oracle/_ref/*.so (torch extension modules; the .so files travel to the
GPU box, the sources never enter the repo).

  build_cost_volume_cuda.build_cost_volume_forward / _backward
        snvc/extension/build_cost_volume/src/BuildCostVolume.cpp:13-48, BuildCostVolume_cuda.cu:63-98,152-205,208-303
  iou3d_nms_cuda.boxes_iou_bev_gpu / nms_gpu
        snvc/extension/iou3d_nms/src/iou3d_nms.cpp:94-177, iou3d_nms_kernel.cu:36-336

Bars: cost-volume forward bit-exact (fp32, fp64, downsample 2, KITTI shape); backward within the reference's own
atomicAdd ordering noise; the C oracle (oracle/cost_volume.c, fma_mode=1) bit-exact against the reference op, which is
what pins the oracle every other cost-volume test uses; rotated IoU within 1e-5; NMS keep lists identical."""
import os
import sys

import numpy as np
import pytest
import torch

from oracle import build_ref, cost_volume as ocv, iou3d_nms as onms

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not build_ref.available(), reason="oracle/_ref/*.so not built (python oracle/build_ref.py)")]

EDGE_SHIFTS = [0.0, 0.25, 1.0, 8.999, 9.0, 12.0, 1e-4, 3.5, 2.0, 100.0]


@pytest.fixture(scope="module")
def ref_cv():
    return build_ref.load("build_cost_volume_cuda")


@pytest.fixture(scope="module")
def ref_nms():
    return build_ref.load("iou3d_nms_cuda")


def _rand(shape, seed, dtype=np.float32):
    return np.random.default_rng(seed).standard_normal(shape).astype(dtype)


def _bcv():
    from snvc_b200.extension import build_cost_volume as m
    return m


@pytest.mark.parametrize("shape,ds", [((2, 3, 6, 10), 1), ((2, 3, 6, 10), 2), ((1, 5, 7, 13), 1), ((3, 8, 12, 40), 1),
                                      ((1, 4, 12, 24), 4), ((2, 32, 24, 78), 1)])
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_forward_bit_exact_vs_reference_op(ref_cv, shape, ds, dtype):
    N = shape[0]
    l, r = _rand(shape, 1, dtype), _rand(shape, 2, dtype)
    s = np.tile(np.array(EDGE_SHIFTS, dtype=dtype)[None], (N, 1))
    s[-1] = s[-1][::-1]
    tl, tr, ts = (torch.from_numpy(a).cuda() for a in (l, r, s))
    want = ref_cv.build_cost_volume_forward(tl, tr, ts, ds)
    got = _bcv().build_cost_volume(tl, tr, ts, ds)
    assert got.shape == want.shape and got.dtype == want.dtype
    assert torch.equal(got, want)
    # and the C oracle itself is pinned to the reference op, bit for bit
    assert np.array_equal(ocv.forward_c(l, r, s, ds, fma_mode=1), want.cpu().numpy())


def test_forward_kitti_shape_bit_exact_vs_reference_op(ref_cv):
    """configs[0]/[1] geometry: features [2,32,96,312], 48 plane-sweep shifts (40.6 .. 2.4 px)."""
    from snvc_b200.utils.geometry import kitti_global_cfg, plane_sweep_shifts
    g = torch.Generator(device="cuda").manual_seed(10)
    l = torch.randn((2, 32, 96, 312), device="cuda", generator=g)
    r = torch.randn((2, 32, 96, 312), device="cuda", generator=g)
    s = torch.from_numpy(plane_sweep_shifts(kitti_global_cfg(), 2)).cuda()
    want = ref_cv.build_cost_volume_forward(l, r, s, 1)
    assert torch.equal(_bcv().build_cost_volume(l, r, s, 1), want)
    # product layouts = the bf16 rounding of the reference volume, channels-last
    want_cl = want.to(torch.bfloat16).permute(0, 2, 3, 4, 1)
    assert torch.equal(_bcv().build_cost_volume_ndhwc_bf16(l, r, s, 1), want_cl)
    rv, lp = _bcv().build_cost_volume_split_bf16(l, r, s, 1)
    assert torch.equal(rv, want_cl[..., 32:])
    for k in range(3):
        assert torch.equal(lp[:, k], want_cl[:, 0, ..., :32])


@pytest.mark.parametrize("shape,ds,dtype", [((2, 3, 6, 10), 1, np.float32), ((2, 3, 6, 10), 2, np.float32),
                                            ((2, 8, 12, 40), 1, np.float64), ((1, 32, 24, 78), 1, np.float32)])
def test_backward_vs_reference_op(ref_cv, shape, ds, dtype, monkeypatch):
    """The reference scatters with atomicAdd (sum order varies run to run); ours is a deterministic gather."""
    N, C, IH, IW = shape
    D = len(EDGE_SHIFTS)
    s = np.tile(np.array(EDGE_SHIFTS, dtype=dtype)[None], (N, 1))
    grad = _rand((N, 2 * C, D, IH // ds, IW // ds), 9, dtype)
    tg, ts = torch.from_numpy(grad).cuda(), torch.from_numpy(s).cuda()
    wl, wr = ref_cv.build_cost_volume_backward(tg, ts, ds)
    l = torch.zeros(shape, dtype=tg.dtype, device="cuda", requires_grad=True)
    r = torch.zeros(shape, dtype=tg.dtype, device="cuda", requires_grad=True)
    monkeypatch.setenv("SNVC_B200_SKIP_SHIFT_CHECK", "1")
    _bcv().build_cost_volume(l, r, ts, ds).backward(tg)
    tol = 1e-5 if dtype == np.float32 else 1e-13
    for got, want in ((l.grad, wl), (r.grad, wr)):
        assert got.shape == want.shape
        assert float((got - want).abs().max()) <= tol * float(want.abs().max())
    assert float(wl.abs().max()) > 0 and float(wr.abs().max()) > 0


def test_reference_op_error_behaviour_matches(ref_cv):
    l = torch.zeros((1, 2, 4, 8))
    with pytest.raises(RuntimeError, match="Not implemented on the CPU"):
        ref_cv.build_cost_volume_forward(l, l, torch.zeros((1, 3)), 1)
    with pytest.raises(RuntimeError, match="Not implemented on the CPU"):
        _bcv().build_cost_volume(l, l, torch.zeros((1, 3)), 1)
    a, b = torch.zeros((1, 2, 4, 8), device="cuda"), torch.zeros((1, 2, 4, 6), device="cuda")
    with pytest.raises(RuntimeError, match="should match their size"):
        ref_cv.build_cost_volume_forward(a, b, torch.zeros((1, 3), device="cuda"), 1)
    with pytest.raises(RuntimeError, match="should match their size"):
        _bcv().build_cost_volume(a, b, torch.zeros((1, 3), device="cuda"), 1)


# ------------------------------------------------------------------------------------------------- N4
def test_boxes_iou_bev_vs_reference_op(ref_nms):
    from snvc_b200 import functional as F
    a, _ = onms.synthetic_boxes(64, seed=7)
    b = a[5:45].copy()
    b[:, :2] += 0.3
    b[:, 6] += 0.2
    ta, tb = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
    want = torch.zeros((a.shape[0], b.shape[0]), device="cuda")
    ref_nms.boxes_iou_bev_gpu(ta, tb, want)
    got = F.boxes_iou_bev(ta, tb)
    assert int((want > 0.05).sum()) > 20
    assert float((got - want).abs().max()) <= 1e-5
    # the numpy oracle against the reference op as well (pins oracle/iou3d_nms.py)
    assert np.max(np.abs(onms.boxes_iou_bev(a, b) - want.cpu().numpy())) <= 1e-5


@pytest.mark.parametrize("n,thresh", [(96, 0.1), (200, 0.25), (64, 0.01), (1, 0.1), (1000, 0.3)])
def test_nms_keep_list_identical_to_reference_op(ref_nms, n, thresh):
    from snvc_b200 import functional as F
    boxes, scores = onms.synthetic_boxes(n, seed=11 + n)
    tb, tsc = torch.from_numpy(boxes).cuda(), torch.from_numpy(scores).cuda()
    order = tsc.sort(0, descending=True)[1]                     # iou3d_nms_utils.py:93-98
    sorted_boxes = tb[order].contiguous()
    keep = torch.zeros(n, dtype=torch.int64)                    # host tensor, as iou3d_nms_utils.py:99 (LongTensor)
    num = ref_nms.nms_gpu(sorted_boxes, keep, float(thresh))
    want = order[keep[:num].cuda()].contiguous()
    got, _ = F.nms_gpu(tb, tsc, thresh)
    assert torch.equal(got, want)
    if n <= 200:
        assert np.array_equal(onms.nms(boxes, scores, thresh), want.cpu().numpy())
