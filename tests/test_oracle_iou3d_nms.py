"""Oracle sanity (CPU): the float32 restatement of the reference's rotated-IoU algorithm vs an independent float64
polygon clipper (the reference's corner test has a 1e-2 margin, so the two agree to ~2e-3, not to rounding)."""
import numpy as np

from oracle import iou3d_nms as o


def test_rotated_iou_restatement_vs_exact_clipping():
    boxes, _ = o.synthetic_boxes(36, seed=5)
    worst, nz = 0.0, 0
    for a in boxes:
        for b in boxes:
            r, e = float(o.iou_bev(a, b)), o.iou_bev_exact(a, b)
            worst, nz = max(worst, abs(r - e)), nz + (e > 0.05)
    assert worst < 5e-3 and nz > 60
    assert abs(float(o.iou_bev(boxes[0], boxes[0])) - 1.0) < 1e-5
    far = boxes[0].copy(); far[0] += 100
    assert float(o.iou_bev(boxes[0], far)) == 0.0


def test_nms_keeps_highest_scores_and_suppresses_overlaps():
    boxes, scores = o.synthetic_boxes(48, seed=6)
    keep = o.nms(boxes, scores, 0.1)
    assert keep[0] == int(np.argmax(scores)) and len(set(keep.tolist())) == len(keep)
    iou = o.boxes_iou_bev(boxes[keep], boxes[keep])
    assert np.all(iou[np.triu_indices(len(keep), 1)] <= 0.1)
    assert len(o.nms(boxes, scores, 0.1, pre_maxsize=10)) <= 10


def test_restatement_vs_reference_cpu_op():
    """PIN: the reference ships a CPU build of the same algorithm (iou3d_cpu.cpp:232-252 `boxes_iou_bev_cpu`);
    oracle/build_ref.py compiles it from /root/reference into oracle/_ref/ (build container only)."""
    import pytest
    from oracle import build_ref
    if not build_ref.available():
        pytest.skip("oracle/_ref not built")
    import torch
    ref = build_ref.load("iou3d_nms_cuda")
    a, _ = o.synthetic_boxes(48, seed=7)
    b = a[5:40].copy()
    b[:, :2] += 0.3
    b[:, 6] += 0.2
    want = torch.zeros((a.shape[0], b.shape[0]))
    ref.boxes_iou_bev_cpu(torch.from_numpy(a), torch.from_numpy(b), want)
    got = o.boxes_iou_bev(a, b)
    assert (want.numpy() > 0.05).sum() > 20
    assert np.max(np.abs(got - want.numpy())) <= 1e-5
