"""Host-side geometry helpers for the global branch (benchmark parametrisation).

Conventions follow the reference helpers: voxel centres `arange(MIN, MAX-eps, step) + step/2`
(snvc/utils/torch_utils.py:85-94), stereo baseline 0.54 m (torch_utils.py:25), range key names
CV_*_MIN/MAX, *_MIN/MAX, VOXEL_*_SIZE (snvc/models/loss3d.py:15-20).  The numeric defaults
(KITTI-typical P2, 0.2 m voxels, depth 2..40.4 m in 48 bins) are benchmark parameters, not
reference content (the reference ships no config)."""
import types

import numpy as np

KITTI_P2 = np.array([[721.5377, 0.0, 609.5593, 44.85728],
                     [0.0, 721.5377, 172.854, 0.2163791],
                     [0.0, 0.0, 1.0, 0.002745884]], dtype=np.float32)


def kitti_global_cfg(IH=384, IW=1248, feat_stride=4, D=48, depth_min=2.0, depth_max=40.4, align_corners=True, GN=False):
    """Attribute-style cfg for snvc_b200.models.stereonet.GlobalHotPath (+ shift helper inputs)."""
    interval = (depth_max - depth_min) / D
    z = (np.float32(depth_min) + (np.arange(D, dtype=np.float32) + np.float32(0.5)) * np.float32(interval)).astype(np.float32)
    Wf, Hf = IW // feat_stride, IH // feat_stride
    if align_corners:      # -1/+1 <-> first/last feature sample and first/last depth-bin centre
        cv = (0.0, float(feat_stride * (Wf - 1)), 0.0, float(feat_stride * (Hf - 1)), float(z[0]), float(z[-1]))
    else:                  # -1/+1 <-> outer edges of the first/last cell
        h = feat_stride / 2.0
        cv = (-h, feat_stride * (Wf - 1) + h, -h, feat_stride * (Hf - 1) + h, float(depth_min), float(depth_max))
    return types.SimpleNamespace(
        IH=IH, IW=IW, feat_stride=feat_stride, D=D, depth_bins=z, fu=float(KITTI_P2[0, 0]), baseline=0.54,
        X_MIN=-30.4, X_MAX=30.4, Y_MIN=-1.0, Y_MAX=3.0, Z_MIN=2.0, Z_MAX=40.4,
        VOXEL_X_SIZE=0.2, VOXEL_Y_SIZE=0.2, VOXEL_Z_SIZE=0.2,
        CV_X_MIN=cv[0], CV_X_MAX=cv[1], CV_Y_MIN=cv[2], CV_Y_MAX=cv[3], CV_Z_MIN=cv[4], CV_Z_MAX=cv[5],
        align_corners=align_corners, GN=GN)


def plane_sweep_shifts(cfg, n=1):
    """shift[n, d] = f_u * baseline / z_d / feat_stride  (pixels at feature resolution, all >= 0)."""
    s = (np.float32(cfg.fu) * np.float32(cfg.baseline) / cfg.depth_bins / np.float32(cfg.feat_stride)).astype(np.float32)
    return np.tile(s[None], (n, 1))
