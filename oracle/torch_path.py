"""ORACLE / CPU baseline (test infrastructure): the global hot path as the reference would run it
on the CPU -- multi-threaded torch ops -- used by bench.py's `cpu_baseline` and `--impl reference`.

The reference has no CPU implementation of its cost-volume op (BuildCostVolume.cpp:26), so that
stage is the vectorised torch recipe of SURVEY.md Appendix A (restating
BuildCostVolume_cuda.cu:15-98); the trunk is oracle.blocks.GlobalTrunk (nn.Conv3d /
ConvTranspose3d / BatchNorm3d exactly as snvc/models/submodule.py:32-50,85-168 builds them) and
the lift is torch.nn.functional.grid_sample on the grid of oracle.global_branch.lift_grid
(projection helper snvc/utils/torch_utils.py:36-45).  tests/test_oracle_torch_path.py checks it
against the scalar / numpy oracle."""
import numpy as np
import torch
import torch.nn.functional as F

from . import global_branch as gb


def cost_volume_torch(left, right, shift, downsample=1):
    """[N,C,IH,IW] x2, shift [N,D] (>= 0) -> [N,2C,D,H,W]; separately rounded (no FMA)."""
    ds = int(downsample)
    N, C, IH, IW = left.shape
    H, W = IH // ds, IW // ds
    img_w = W * ds
    D = shift.shape[1]
    iw = (torch.arange(W, dtype=left.dtype) * ds)
    x = iw[None, None, :] - shift[:, :, None]                          # [N,D,W]
    valid = (x >= 0) & (x <= img_w - 1)
    xc = x.clamp(min=0)
    x0 = xc.floor().clamp(max=img_w - 1)
    edge = x0 >= img_w - 1
    x1 = torch.where(edge, x0, x0 + 1)
    lx = torch.where(edge, torch.zeros_like(xc), xc - x0)
    hx = 1 - lx
    x0i = torch.where(valid, x0, torch.zeros_like(x0)).long()
    x1i = torch.where(valid, x1, torch.zeros_like(x1)).long()
    rows = right[:, :, 0:H * ds:ds, :]                                 # [N,C,H,IW]
    idx0 = x0i[:, None, :, None, :].expand(N, C, D, H, W)
    idx1 = x1i[:, None, :, None, :].expand(N, C, D, H, W)
    rows_b = rows[:, :, None].expand(N, C, D, H, IW)
    v1 = torch.gather(rows_b, 4, idx0)
    v2 = torch.gather(rows_b, 4, idx1)
    r = (hx[:, None, :, None, :] * v1 + lx[:, None, :, None, :] * v2) * valid[:, None, :, None, :].to(left.dtype)
    l = left[:, :, None, 0:H * ds:ds, 0:W * ds:ds].expand(N, C, D, H, W)
    return torch.cat([l, r], dim=1)


class GlobalHotPathCPU:
    def __init__(self, trunk, geom: gb.GlobalGeometry):
        self.trunk = trunk.eval()
        self.geom = geom
        self.zs, self.ys, self.xs = gb.voxel_centres(geom)
        self.cv = geom.cv_ranges()

    @torch.no_grad()
    def __call__(self, left, right, shift, Ps):
        cost = cost_volume_torch(left, right, shift, 1)
        feat = self.trunk(cost)
        outs = []
        for n in range(feat.shape[0]):
            g, valid = gb.lift_grid(self.zs, self.ys, self.xs, Ps[n].numpy(), self.cv)
            g = np.where(np.isfinite(g), g, np.float32(-2))
            o = F.grid_sample(feat[n:n + 1], torch.from_numpy(g[None]), mode="bilinear", padding_mode="zeros",
                              align_corners=self.geom.align_corners)
            outs.append(o[0] * torch.from_numpy(valid[None].astype(np.float32)))
        return torch.stack(outs)
