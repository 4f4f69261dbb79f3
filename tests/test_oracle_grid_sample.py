"""Pins oracle/grid_sample.py against the third-party arithmetic the reference actually calls
(torch.nn.functional.grid_sample, vernier.py:339-340) executed on the CPU."""
import warnings

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import grid_sample as gs


@pytest.mark.parametrize("ac", [False, True])
def test_2d_vs_torch(ac):
    rng = np.random.default_rng(1)
    inp = rng.standard_normal((2, 5, 7, 9)).astype(np.float32)
    grid = (rng.random((2, 4, 11, 2)).astype(np.float32) * 2.6 - 1.3)
    a = gs.grid_sample_2d(inp, grid, ac)
    b = F.grid_sample(torch.from_numpy(inp), torch.from_numpy(grid), align_corners=ac).numpy()
    assert np.max(np.abs(a - b)) <= 1e-5 * np.max(np.abs(b))


@pytest.mark.parametrize("ac", [False, True])
def test_3d_vs_torch(ac):
    rng = np.random.default_rng(2)
    inp = rng.standard_normal((2, 3, 4, 6, 5)).astype(np.float32)
    grid = (rng.random((2, 3, 4, 5, 3)).astype(np.float32) * 2.6 - 1.3)
    a = gs.grid_sample_3d(inp, grid, ac)
    b = F.grid_sample(torch.from_numpy(inp), torch.from_numpy(grid), align_corners=ac).numpy()
    assert np.max(np.abs(a - b)) <= 1e-5 * np.max(np.abs(b))


def test_unnormalize_forms_agree_bitwise():
    # ((g+1)*W-1)/2 (CUDA header form) == (g+1)*(W/2)-0.5 (CPU vectorised form)
    g = (np.random.default_rng(3).random(200000).astype(np.float32) * 2.4 - 1.2)
    for W in (16, 48, 63, 64, 77, 96, 312):
        a = gs.unnormalize(g, W, False)
        b = ((g + np.float32(1)).astype(np.float32) * np.float32(W / 2)).astype(np.float32) - np.float32(0.5)
        assert np.array_equal(a, b.astype(np.float32))


def test_roi_voxel_sample_vs_reference_recipe():
    """Same call sequence as vernier.py:332-346, executed with torch on the CPU."""
    rng = np.random.default_rng(4)
    N, Fc, Hf, Wf, nh, nw, nl, res = 2, 6, 16, 16, 4, 6, 5, (64, 64)
    lf, rf = (rng.standard_normal((N, Fc, Hf, Wf)).astype(np.float32) for _ in range(2))
    lp, rp = (rng.uniform(-6, 70, (N, 2, nh * nw * nl)).astype(np.float32) for _ in range(2))
    got = gs.roi_voxel_sample(lf, rf, lp, rp, nh, nw, nl, res)
    outs = []
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for f, p in ((lf, lp), (rf, rp)):
            g = torch.from_numpy(p).permute(0, 2, 1).reshape(N, nh, nw * nl, 2)
            g[:, :, :, 0] = g[:, :, :, 0] / res[1] * 2 - 1
            g[:, :, :, 1] = g[:, :, :, 1] / res[0] * 2 - 1
            outs.append(F.grid_sample(torch.from_numpy(f), g).reshape(N, Fc, nh, nw, nl))
    want = torch.cat(outs, 1).numpy()
    assert got.shape == want.shape
    assert np.max(np.abs(got - want)) <= 1e-5 * np.max(np.abs(want))
