"""GPU: the in-frustum-only host return (snvc_masked_rows_to_host, HostPipeline sparse_return) delivers the DENSE
tensor into the pinned host buffer bit for bit, while moving only the rows that can differ from what the buffer holds."""
import numpy as np
import pytest
import torch

import synth

pytestmark = pytest.mark.gpu


def _call(src, valid, prev, h_out, moved, blocks=0):
    from snvc_b200 import _lib
    st = _lib.lib().snvc_masked_rows_to_host(src.data_ptr(), valid.data_ptr(), prev.data_ptr(), h_out.data_ptr(), valid.numel(),
                                             src.shape[-1] * src.element_size(), blocks, moved.data_ptr(), _lib.stream_ptr())
    _lib.check(st, "snvc_masked_rows_to_host")


@pytest.mark.parametrize("C,rows,blocks", [(32, 100003, 0), (64, 4099, 3), (8, 777, 1), (16, 50000, 16)])
def test_masked_rows_to_host_equals_dense_copy(C, rows, blocks):
    g = torch.Generator(device="cuda").manual_seed(C + rows)
    h_out = torch.full((rows, C), 7.0, dtype=torch.bfloat16).pin_memory()          # garbage: first use must overwrite all of it
    prev = torch.ones(rows, dtype=torch.uint8, device="cuda")
    moved = torch.zeros((), dtype=torch.int64, device="cuda")
    last_valid = torch.ones(rows, dtype=torch.bool, device="cuda")
    for it in range(4):
        valid = (torch.rand(rows, device="cuda", generator=g) < (0.2 + 0.2 * it)).to(torch.uint8)
        src = torch.randn((rows, C), device="cuda", generator=g).to(torch.bfloat16)
        src = torch.where(valid[:, None].bool(), src, torch.zeros_like(src))          # masked rows are +0, as the lift writes them
        moved.zero_()
        _call(src, valid, prev, h_out, moved, blocks)
        torch.cuda.synchronize()
        assert torch.equal(h_out.view(torch.int16), src.cpu().view(torch.int16))
        assert torch.equal(prev, valid)
        touched = int((valid.bool() | last_valid).sum())
        assert int(moved.item()) == touched * C * 2                                   # only rows that could differ moved
        last_valid = valid.bool()


def test_masked_rows_to_host_rejects_pageable_memory():
    from snvc_b200 import _lib
    src = torch.zeros((64, 32), dtype=torch.bfloat16, device="cuda")
    valid = torch.ones(64, dtype=torch.uint8, device="cuda")
    h = torch.zeros((64, 32), dtype=torch.bfloat16)                                   # not pinned
    st = _lib.lib().snvc_masked_rows_to_host(src.data_ptr(), valid.data_ptr(), valid.data_ptr(), h.data_ptr(), 64, 64, 0, None,
                                             _lib.stream_ptr())
    assert st != 0


@pytest.mark.parametrize("graphed", [True, False])
def test_pipeline_sparse_return_equals_dense_with_changing_calibration(graphed):
    """Batches with DIFFERENT projection matrices (different frustum footprints) through the same two host buffers:
    every delivered buffer equals the dense forward bit for bit, and fewer bytes than the dense tensor crossed PCIe."""
    import types
    from oracle import global_branch as ogb
    from snvc_b200.models.stereonet import GlobalHotPath, HostPipeline
    geom = ogb.GlobalGeometry(IH=64, IW=192, D=8, depth_min=2.0, depth_max=14.8, X_MIN=-6.0, X_MAX=6.0, Y_MIN=-1.0, Y_MAX=2.0,
                              Z_MIN=2.0, Z_MAX=14.0, VOXEL_X_SIZE=0.4, VOXEL_Y_SIZE=0.5, VOXEL_Z_SIZE=0.5, align_corners=True,
                              P=np.array([[110.0, 0, 96.0, 6.0], [0, 110.0, 30.0, 0.03], [0, 0, 1.0, 0.0003]], np.float32))
    cv = geom.cv_ranges()
    cfg = types.SimpleNamespace(X_MIN=geom.X_MIN, X_MAX=geom.X_MAX, Y_MIN=geom.Y_MIN, Y_MAX=geom.Y_MAX, Z_MIN=geom.Z_MIN,
                                Z_MAX=geom.Z_MAX, VOXEL_X_SIZE=geom.VOXEL_X_SIZE, VOXEL_Y_SIZE=geom.VOXEL_Y_SIZE,
                                VOXEL_Z_SIZE=geom.VOXEL_Z_SIZE, CV_X_MIN=cv[0], CV_X_MAX=cv[1], CV_Y_MIN=cv[2], CV_Y_MAX=cv[3],
                                CV_Z_MIN=cv[4], CV_Z_MAX=cv[5], align_corners=True, GN=False)
    N, Fc, H, W = 2, 32, geom.IH // 4, geom.IW // 4
    m = GlobalHotPath(cfg).eval()
    m.load_state_dict(synth.det_state_dict(m, 41), strict=True)
    m = m.cuda()
    shift = torch.from_numpy(np.ascontiguousarray(geom.shifts(N)))
    scales = [1.0, 1.6, 0.7, 1.0, 2.2, 0.5]                                            # focal length changes the footprint
    batches = []
    for i, sc in enumerate(scales):
        Ps = np.stack([geom.P * np.float32([[sc], [sc], [1.0]]), geom.P * np.float32([[1.0], [sc], [1.0]])]).astype(np.float32)
        batches.append((torch.from_numpy(synth.det_uniform((N, Fc, H, W), 600 + i)), torch.from_numpy(synth.det_uniform((N, Fc, H, W), 700 + i)),
                        torch.from_numpy(Ps)))
    with torch.no_grad():
        want = [m(l.cuda(), r.cuda(), shift.cuda(), P.cuda(), torch.bfloat16, "NDHWC").cpu() for l, r, P in batches]
        pipe = HostPipeline(m, depth=2, graphed=graphed)
        assert pipe.sparse_return
        bufs = [torch.full(want[0].shape, 3.0, dtype=torch.bfloat16).pin_memory() for _ in range(2)]
        for i, (l, r, P) in enumerate(batches):
            ev = pipe.submit(l.pin_memory(), r.pin_memory(), shift.pin_memory(), P.pin_memory(), bufs[i % 2])
            ev.synchronize()
            assert torch.equal(bufs[i % 2].view(torch.int16), want[i].view(torch.int16)), i
        dense = len(batches) * want[0].numel() * 2
        moved = int(pipe.moved_bytes.item())
        assert 0 < moved < dense, (moved, dense)
        # a buffer somebody else wrote to must be forgotten, then it is rewritten completely
        bufs[0].fill_(5.0)
        pipe.forget(bufs[0])
        l, r, P = batches[2]
        pipe.submit(l.pin_memory(), r.pin_memory(), shift.pin_memory(), P.pin_memory(), bufs[0]).synchronize()
        assert torch.equal(bufs[0].view(torch.int16), want[2].view(torch.int16))
