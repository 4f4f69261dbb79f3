"""The multi-threaded torch CPU baseline (oracle/torch_path.py) agrees with the scalar/numpy oracle."""
import numpy as np
import torch

import synth
from oracle import blocks, cost_volume as ocv, global_branch as ogb, torch_path


def test_cost_volume_torch_matches_c_oracle():
    rng = np.random.default_rng(0)
    l, r = rng.standard_normal((2, 3, 6, 10)).astype(np.float32), rng.standard_normal((2, 3, 6, 10)).astype(np.float32)
    s = np.float32([[0, 0.25, 1.0, 8.999, 9.0, 12.0, 1e-4], [0.5, 2.5, 3.75, 9.0, 8.5, 0.0, 100.0]])
    for ds in (1, 2):
        a = torch_path.cost_volume_torch(torch.from_numpy(l), torch.from_numpy(r), torch.from_numpy(s), ds).numpy()
        assert np.array_equal(a, ocv.forward_c(l, r, s, ds, fma_mode=0))


def test_global_hot_path_cpu_matches_numpy_oracle():
    geom = ogb.GlobalGeometry(IH=32, IW=96, D=4, depth_min=2.0, depth_max=8.4, X_MIN=-3.0, X_MAX=3.0, Y_MIN=-1.0,
                              Y_MAX=1.0, Z_MIN=2.0, Z_MAX=8.0, VOXEL_X_SIZE=0.5, VOXEL_Y_SIZE=0.5, VOXEL_Z_SIZE=0.5,
                              P=np.array([[60.0, 0, 48.0, 3.0], [0, 60.0, 15.0, 0.03], [0, 0, 1.0, 0.0003]], np.float32))
    N, Fc, H, W = 1, 8, geom.IH // 4, geom.IW // 4
    lf, rf = synth.det_uniform((N, Fc, H, W), 1), synth.det_uniform((N, Fc, H, W), 2)
    shift, Ps = geom.shifts(N), geom.P[None].copy()
    trunk = blocks.GlobalTrunk(2 * Fc, 8).eval()
    trunk.load_state_dict(synth.det_state_dict(trunk, 3))
    got = torch_path.GlobalHotPathCPU(trunk, geom)(torch.from_numpy(lf), torch.from_numpy(rf), torch.from_numpy(shift),
                                                   torch.from_numpy(Ps)).numpy()
    with torch.no_grad():
        feat = trunk(torch.from_numpy(ocv.forward_c(lf, rf, shift, 1, fma_mode=0))).numpy()
    want, _ = ogb.frustum_lift(feat, Ps, geom)
    assert got.shape == want.shape
    assert np.max(np.abs(got - want)) <= 1e-5 * max(1.0, np.max(np.abs(want)))
