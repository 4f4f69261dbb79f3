"""GPU parity: rotated BEV IoU + NMS vs oracle/iou3d_nms.py."""
import numpy as np
import pytest
import torch

from oracle import iou3d_nms as o

pytestmark = pytest.mark.gpu


def test_boxes_iou_bev_vs_oracle():
    from snvc_b200 import functional as F
    a, _ = o.synthetic_boxes(40, seed=7)
    b = a[5:30].copy()
    b[:, :2] += 0.3                                       # shifted copies: plenty of partial overlaps
    b[:, 6] += 0.2
    want = o.boxes_iou_bev(a, b)
    got = F.boxes_iou_bev(torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()).cpu().numpy()
    assert got.shape == want.shape and (want > 0.05).sum() > 10
    assert np.max(np.abs(got - want)) <= 1e-5           # device sinf / cosf / atan2f vs libm: a few ulp in the corners


@pytest.mark.parametrize("n,thresh", [(96, 0.1), (200, 0.25), (64, 0.01), (1, 0.1)])
def test_nms_identical_keep_set_and_order(n, thresh):
    from snvc_b200 import functional as F
    boxes, scores = o.synthetic_boxes(n, seed=11 + n)
    want = o.nms(boxes, scores, thresh)
    # the comparison is only meaningful if no pair sits on the threshold (it does not, by a wide margin)
    order = np.argsort(-scores, kind="stable")
    iou = F.boxes_iou_bev(torch.from_numpy(boxes[order]).cuda(), torch.from_numpy(boxes[order]).cuda()).cpu().numpy()
    assert np.min(np.abs(iou - thresh)) > 2e-5
    got, none = F.nms_gpu(torch.from_numpy(boxes).cuda(), torch.from_numpy(scores).cuda(), thresh)
    assert none is None and got.dtype == torch.int64
    assert np.array_equal(got.cpu().numpy(), want)      # identical selection, identical (score) order
    sel, num = F.nms_gpu_device(torch.from_numpy(boxes).cuda(), torch.from_numpy(scores).cuda(), thresh, pre_maxsize=50)
    w2 = o.nms(boxes, scores, thresh, pre_maxsize=50)
    assert int(num.item()) == len(w2) and np.array_equal(sel[:len(w2)].cpu().numpy(), w2) and bool((sel[len(w2):] == -1).all())


def test_nms_empty_and_bad_input():
    from snvc_b200 import functional as F
    got, _ = F.nms_gpu(torch.zeros((0, 7), device="cuda"), torch.zeros((0,), device="cuda"), 0.1)
    assert got.numel() == 0
    with pytest.raises(RuntimeError):
        F.nms_gpu(torch.zeros((4, 5), device="cuda"), torch.zeros((4,), device="cuda"), 0.1)


def test_batched_nms_equals_per_set_nms():
    """snvc_nms_bev_batched (all box sets of a batch in two launches; 512-thread mask blocks) == the single-set entry point."""
    from snvc_b200 import functional as F
    sets = [o.synthetic_boxes(200, seed=40 + i) for i in range(5)]
    boxes = torch.from_numpy(np.stack([b for b, _ in sets])).cuda()
    scores = torch.from_numpy(np.stack([s for _, s in sets])).cuda()
    sel, num = F.nms_gpu_device_batched(boxes, scores, 0.2)
    for i in range(5):
        s1, n1 = F.nms_gpu_device(boxes[i], scores[i], 0.2)
        assert int(num[i].item()) == int(n1.item()) and torch.equal(sel[i], s1)
        want = o.nms(sets[i][0], sets[i][1], 0.2)
        assert np.array_equal(sel[i, :len(want)].cpu().numpy(), want)
