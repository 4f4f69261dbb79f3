"""ORACLE (test infrastructure): the restated global branch hot path on the CPU.

cost volume (oracle.cost_volume) -> GlobalTrunk (oracle.blocks) -> frustum-to-voxel lift.

The reference does not ship the global model class (snvc/models/__init__.py:1-2 are
commented-out imports); per SURVEY.md section 3.4 the composition is restated from the
DSGN lineage (README.md:68) with the shipped helpers as the spec for each step:
  * projection          snvc/utils/torch_utils.py:36-45  (project_rect_to_image)
  * voxel-centre grid   snvc/utils/torch_utils.py:77-98  (arange(MIN, MAX-eps, step)+step/2)
  * range key names     snvc/models/loss3d.py:15-20      (CV_*_MIN/MAX, *_MIN/MAX, VOXEL_*_SIZE)
The numeric geometry below (KITTI-typical P2, 0.2 m voxels, ...) is benchmark
parametrisation, not reference content.

Bit-exact contract for the lift (the CUDA kernel follows the same fp32 op order, no FMA):
  uh = ((P00*x + P01*y) + P02*z) + P03      (likewise vh, wh)          4-term dot, left to right
  u = uh / wh ; v = vh / wh                                            torch_utils.py:43-44
  g = ((c - CV_MIN) / (CV_MAX - CV_MIN)) * 2 - 1      for c in (u, v, z)
  valid = all(-1 <= g <= 1)
  out = grid_sample(volume, g, align_corners) * valid
"""
from dataclasses import dataclass, field
import numpy as np

from . import grid_sample as gs

F32 = np.float32


@dataclass
class GlobalGeometry:
    """Benchmark geometry."""
    IH: int = 384
    IW: int = 1248
    feat_stride: int = 4
    D: int = 48                       # maxdisp 192 / downsample_disp 4
    fu: float = 721.5377
    baseline: float = 0.54            # torch_utils.py:25
    depth_min: float = 2.0
    depth_max: float = 40.4
    X_MIN: float = -30.4
    X_MAX: float = 30.4
    Y_MIN: float = -1.0
    Y_MAX: float = 3.0
    Z_MIN: float = 2.0
    Z_MAX: float = 40.4
    VOXEL_X_SIZE: float = 0.2
    VOXEL_Y_SIZE: float = 0.2
    VOXEL_Z_SIZE: float = 0.2
    align_corners: bool = True        # cfg key `align_corners` (submodule.py:375)
    P: np.ndarray = field(default_factory=lambda: np.array(
        [[721.5377, 0.0, 609.5593, 44.85728],
         [0.0, 721.5377, 172.854, 0.2163791],
         [0.0, 0.0, 1.0, 0.002745884]], dtype=np.float32))

    @property
    def depth_interval(self):
        return (self.depth_max - self.depth_min) / self.D

    def depth_bins(self):
        """z_d = depth_min + (d + 0.5) * interval, fp32."""
        return (F32(self.depth_min) + (np.arange(self.D, dtype=F32) + F32(0.5)) * F32(self.depth_interval)).astype(F32)

    def shifts(self, n=1):
        """shift[n,d] = fu * baseline / z_d / feat_stride  (px at feature resolution, all >= 0)."""
        s = (F32(self.fu) * F32(self.baseline) / self.depth_bins() / F32(self.feat_stride)).astype(F32)
        return np.tile(s[None], (n, 1))

    def cv_ranges(self):
        """(CV_X_MIN, CV_X_MAX, CV_Y_MIN, CV_Y_MAX, CV_Z_MIN, CV_Z_MAX) (loss3d.py:15-17 key names).

        With align_corners=True, -1/+1 map to the first/last feature sample: image px 0 and
        feat_stride*(W-1); depth-bin centres z_0 and z_{D-1}.  With align_corners=False they map
        to the outer edges of the first/last cell."""
        z = self.depth_bins()
        Wf, Hf = self.IW // self.feat_stride, self.IH // self.feat_stride
        if self.align_corners:
            return (0.0, float(self.feat_stride * (Wf - 1)), 0.0, float(self.feat_stride * (Hf - 1)),
                    float(z[0]), float(z[-1]))
        h = self.feat_stride / 2.0
        return (-h, self.feat_stride * (Wf - 1) + h, -h, self.feat_stride * (Hf - 1) + h,
                float(self.depth_min), float(self.depth_max))

    def voxel_dims(self):
        return (len(_centres(self.Z_MIN, self.Z_MAX, self.VOXEL_Z_SIZE)),
                len(_centres(self.Y_MIN, self.Y_MAX, self.VOXEL_Y_SIZE)),
                len(_centres(self.X_MIN, self.X_MAX, self.VOXEL_X_SIZE)))


def _centres(lo, hi, step):
    """torch_utils.py:85-94: arange(MIN, MAX - sign(step)*1e-10, step, float32) + step/2."""
    import torch
    a = torch.arange(lo, hi - np.sign(step) * 1e-10, step=step, dtype=torch.float32) + step / 2.0
    return a.numpy().astype(F32)


def voxel_centres(geom: GlobalGeometry):
    return (_centres(geom.Z_MIN, geom.Z_MAX, geom.VOXEL_Z_SIZE),
            _centres(geom.Y_MIN, geom.Y_MAX, geom.VOXEL_Y_SIZE),
            _centres(geom.X_MIN, geom.X_MAX, geom.VOXEL_X_SIZE))


def lift_grid(zs, ys, xs, P, cv):
    """Normalised sampling grid [Z,Y,X,3] and validity mask [Z,Y,X] for one projection matrix."""
    P = P.astype(F32)
    z, y, x = np.meshgrid(zs.astype(F32), ys.astype(F32), xs.astype(F32), indexing="ij")

    def row(r):
        acc = (P[r, 0] * x).astype(F32)
        acc = (acc + (P[r, 1] * y).astype(F32)).astype(F32)
        acc = (acc + (P[r, 2] * z).astype(F32)).astype(F32)
        return (acc + P[r, 3]).astype(F32)

    uh, vh, wh = row(0), row(1), row(2)
    with np.errstate(divide="ignore", invalid="ignore"):
        u = (uh / wh).astype(F32)
        v = (vh / wh).astype(F32)

    def norm(c, lo, hi):
        return ((((c - F32(lo)).astype(F32) / F32(F32(hi) - F32(lo))).astype(F32) * F32(2)).astype(F32) - F32(1)).astype(F32)

    gx, gy, gz = norm(u, cv[0], cv[1]), norm(v, cv[2], cv[3]), norm(z, cv[4], cv[5])
    valid = ((gx >= -1) & (gx <= 1) & (gy >= -1) & (gy <= 1) & (gz >= -1) & (gz <= 1))
    return np.stack([gx, gy, gz], -1), valid


def frustum_lift(volume, Ps, geom: GlobalGeometry):
    """volume [N,C,D,H,W] fp32, Ps [N,3,4] -> voxels [N,C,Z,Y,X] fp32 (+ valid [N,Z,Y,X])."""
    zs, ys, xs = voxel_centres(geom)
    cv = geom.cv_ranges()
    outs, valids = [], []
    for n in range(volume.shape[0]):
        g, valid = lift_grid(zs, ys, xs, Ps[n], cv)
        g = np.where(np.isfinite(g), g, F32(-2))          # NaN/inf (wh == 0) can only be invalid
        o = gs.grid_sample_3d(volume[n:n + 1], g[None], geom.align_corners)[0]
        outs.append(o * valid[None].astype(F32))
        valids.append(valid)
    return np.stack(outs), np.stack(valids)


def depth_head(vol, classif, depth_values, out_size, align_corners=True):
    """Restated depth head (SURVEY.md 3.4 'adjacent'; blocks pinned by submodule.py:32-50,76-83): plain torch fp32.
    vol [N,C,D,H,W] torch tensor, classif = Sequential(convbn_3d, ReLU, Conv3d(C,1)), depth_values [Dout] tensor,
    out_size (Dout,Hout,Wout) -> depth [N,Hout,Wout] (numpy)."""
    import torch
    import torch.nn.functional as F
    with torch.no_grad():
        cost = classif(vol)                                                        # [N,1,D,H,W]
        cost = F.interpolate(cost, list(out_size), mode="trilinear", align_corners=align_corners)
        prob = F.softmax(torch.squeeze(cost, 1), dim=1)
        return torch.sum(prob * depth_values[None, :, None, None], 1).numpy()      # disparityregression.forward
