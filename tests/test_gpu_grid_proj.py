"""GPU parity: on-device projection of the instance sampling grid vs oracle/grid_proj.py
(which is bit-exact against the reference's own _generate_grid_proj, tests/test_oracle_grid_proj.py)."""
import numpy as np
import pytest
import torch

from oracle import grid_proj as ogp

pytestmark = pytest.mark.gpu


def _ulp_diff(a, b):
    ia, ib = a.view(np.int32).astype(np.int64), b.view(np.int32).astype(np.int64)
    ia = np.where(ia < 0, -(ia & 0x7FFFFFFF), ia)
    ib = np.where(ib < 0, -(ib & 0x7FFFFFFF), ib)
    return np.abs(ia - ib)


@pytest.mark.parametrize("n,grid", [(3, (8, 16, 24)), (5, (4, 7, 9)), (2, (32, 128, 192))])
def test_grid_project_vs_oracle(n, grid):
    from snvc_b200 import functional as F
    c = ogp.synthetic_case(n=n, grid_resolution=grid)
    wl, wr, wg = ogp.generate_grid_proj(c["samples"], c["P_left"], c["P_right"], c["trans_l"], c["trans_r"],
                                        c["x_range"], c["y_range"], c["z_range"], c["grid_resolution"])
    gl, gr, gg = F.roi_grid_project(c["samples"], c["P_left"], c["P_right"], c["trans_l"], c["trans_r"], c["x_range"],
                                    c["y_range"], c["z_range"], c["grid_resolution"], return_grid=True)
    for got, want in ((gl, wl), (gr, wr), (gg, wg.astype(np.float32))):
        got = got.cpu().numpy()
        assert got.shape == want.shape and got.dtype == np.float32
        d = _ulp_diff(got, np.ascontiguousarray(want))
        assert d.max() <= 1                                   # float64 FMA chain vs dgemm: ties of the float32 cast only
        assert (d == 0).mean() > 0.9999


def test_grid_project_feeds_roi_sampling_like_host_coordinates():
    """End of the N1 -> A3 chain: sampling with device-generated coordinates == sampling with the oracle's coordinates,
    except at the (rare) points whose coordinate differs by one ulp."""
    import synth
    from snvc_b200 import functional as F
    c = ogp.synthetic_case(n=2, grid_resolution=(8, 16, 24))
    wl, wr, _ = ogp.generate_grid_proj(c["samples"], c["P_left"], c["P_right"], c["trans_l"], c["trans_r"],
                                       c["x_range"], c["y_range"], c["z_range"], c["grid_resolution"])
    gl, gr = F.roi_grid_project(c["samples"], c["P_left"], c["P_right"], c["trans_l"], c["trans_r"], c["x_range"],
                                c["y_range"], c["z_range"], c["grid_resolution"])
    lf = torch.from_numpy(synth.det_uniform((2, 32, 64, 64), 1)).cuda()
    rf = torch.from_numpy(synth.det_uniform((2, 32, 64, 64), 2)).cuda()
    a = F.roi_voxel_sample(lf, rf, gl, gr, (256, 256))
    b = F.roi_voxel_sample(lf, rf, torch.from_numpy(wl).cuda(), torch.from_numpy(wr).cuda(), (256, 256))
    same = (gl.cpu().numpy() == wl).all(axis=1) & (gr.cpu().numpy() == wr).all(axis=1)        # [N, P]
    assert same.mean() > 0.999
    a, b = a.cpu().numpy(), b.cpu().numpy()
    assert all(np.array_equal(a[i][:, same[i]], b[i][:, same[i]]) for i in range(2))
    assert np.abs(a).max() > 0


def test_grid_project_rejects_bad_shapes():
    from snvc_b200 import functional as F
    c = ogp.synthetic_case(n=2)
    with pytest.raises(RuntimeError):
        F.roi_grid_project(c["samples"], c["P_left"][:2], c["P_right"], c["trans_l"], c["trans_r"], c["x_range"],
                           c["y_range"], c["z_range"], c["grid_resolution"])
