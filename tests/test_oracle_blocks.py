"""Pins oracle/blocks.py to the reference: the golden outputs were produced by the reference's own
modules (tests/golden/make_golden.py, run against /root/reference); here the restated classes
load the same deterministic weights with strict=True and must reproduce them on the CPU."""
import numpy as np
import pytest
import torch

import synth
from oracle import blocks, grid_sample as gs



@pytest.fixture(autouse=True)
def _no_grad():
    with torch.no_grad():
        yield


def _close(a, b, tol=2e-6):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape
    assert np.max(np.abs(a - b)) <= tol * max(1.0, np.max(np.abs(b))), np.max(np.abs(a - b))


@pytest.mark.parametrize("gn", [False, True])
def test_hourglass_matches_reference(golden, gn):
    g = golden("hourglass_gn" if gn else "hourglass_bn")
    m = blocks.Hourglass(32, gn=gn).eval()
    m.load_state_dict(synth.det_state_dict(m, 11 + int(gn)), strict=True)
    x = torch.from_numpy(synth.det_uniform((1, 32, 8, 16, 16), 101))
    out, pre, post = m(x, None, None)
    _close(out, g["out"]); _close(pre, g["pre"]); _close(post, g["post"])
    out2, pre2, post2 = m(x, pre.clone(), post.clone())
    _close(out2, g["out2"]); _close(pre2, g["pre2"]); _close(post2, g["post2"])


def test_hg16_matches_reference(golden):
    g = golden("hg16_bn")
    m = blocks.HourglassDownsample16(32).eval()
    m.load_state_dict(synth.det_state_dict(m, 21), strict=True)
    _close(m(torch.from_numpy(synth.det_uniform((1, 32, 16, 16, 16), 102))), g["out"])


def _vernier_inputs():
    nh, nw, nl = 16, 32, 48
    P = nh * nw * nl
    lf, rf = synth.det_uniform((1, 32, 16, 16), 201), synth.det_uniform((1, 32, 16, 16), 202)
    gl = synth.det_uniform((1, 2, P), 203, -6.4, 70.4, bf16=False)
    gr = synth.det_uniform((1, 2, P), 204, -6.4, 70.4, bf16=False)
    return lf, rf, gl, gr, (nh, nw, nl)


def test_roi_voxel_sample_matches_reference(golden):
    g = golden("vernier_bev3")
    lf, rf, gl, gr, (nh, nw, nl) = _vernier_inputs()
    vox = gs.roi_voxel_sample(lf, rf, gl, gr, nh, nw, nl, (64, 64))
    assert vox.shape == (1, 64, nh, nw, nl)
    _close(vox[:, :, ::2, ::4, ::4], g["voxel_sub"], 1e-5)
    np.testing.assert_allclose(vox.astype(np.float64).sum(axis=(0, 2, 3, 4)), g["voxel_chan_sum"], rtol=0, atol=2e-2)


def test_vernier3d_matches_reference(golden):
    g = golden("vernier_bev3")
    lf, rf, gl, gr, (nh, nw, nl) = _vernier_inputs()
    vox = torch.from_numpy(gs.roi_voxel_sample(lf, rf, gl, gr, nh, nw, nl, (64, 64)))
    m = blocks.Vernier3D(32, n_sample_w=nw).eval()
    m.load_state_dict(synth.det_state_dict(m, 31), strict=True)
    bev, occ = m(vox)
    _close(bev, g["voxel_bev"], 2e-5)
    _close(occ, g["occupancy"], 2e-5)


def test_hourglass2d_matches_reference(golden):
    g = golden("hourglass2d_bn")
    m = blocks.Hourglass2d(64).eval()
    m.load_state_dict(synth.det_state_dict(m, 51), strict=True)
    out, pre, post = m(torch.from_numpy(synth.det_uniform((1, 64, 24, 16), 103)), None, None)
    _close(out, g["out"]); _close(pre, g["pre"]); _close(post, g["post"])
    m16 = blocks.Hourglass2dDownsample16(64).eval()
    m16.load_state_dict(synth.det_state_dict(m16, 52), strict=True)
    _close(m16(torch.from_numpy(synth.det_uniform((1, 64, 48, 32), 104))), g["out16"])


def test_vernier_bev_tail_matches_reference(golden):
    """conv5 -> hm1 -> permute -> hm2 on the reference's own voxel_BEV reproduces the reference's heatmaps (`ncf`)."""
    g = golden("vernier_bev3")
    tail = blocks.VernierBevTail(128, 9, n_sample_w=32).eval()
    tail.load_state_dict(synth.det_state_dict(tail, 33), strict=True)
    _close(tail(torch.from_numpy(g["voxel_bev"])), g["ncf"], 2e-5)
