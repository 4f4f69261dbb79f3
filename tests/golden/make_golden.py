"""Generate tests/golden/*.npz by RUNNING THE REFERENCE'S OWN MODULES (CPU, this container).

    PYTHONPATH=/root/reference python tests/golden/make_golden.py

Needs /root/reference (absent on the GPU box) -> the outputs are committed.  Inputs and weights
are not stored: they are regenerated bit-identically from tests/golden/synth.py.
Reference entry points exercised:
  snvc.models.submodule.hourglass                 (submodule.py:85-168)   bn and gn
  snvc.models.submodule.hourglass_downsample_16   (submodule.py:223-268)
  snvc.models.submodule.hourglass2d / hourglass2d_downsample_16   (submodule.py:317-361, 270-315)
  snvc.models.vernier.VernierScale.construct_voxel        (vernier.py:323-360)
  snvc.models.vernier.VernierScale.predict_3d_heatmaps    (vernier.py:414-458), BEV_type3
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, "/root/reference")
import synth  # noqa: E402

# matplotlib is imported at module import time by vernier.py:11 / visualization/points.py:8
for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.cm", "mpl_toolkits", "mpl_toolkits.mplot3d"):
    sys.modules.setdefault(name, types.ModuleType(name))
sys.modules["mpl_toolkits.mplot3d"].Axes3D = object

from snvc.models import submodule as ref_sub  # noqa: E402
from snvc.models import vernier as ref_vernier  # noqa: E402

torch.set_grad_enabled(False)


def case_hourglass(gn):
    m = ref_sub.hourglass(32, gn=gn).eval()
    m.load_state_dict(synth.det_state_dict(m, 11 + int(gn)), strict=True)
    x = torch.from_numpy(synth.det_uniform((1, 32, 8, 16, 16), 101))
    out, pre, post = m(x, None, None)
    # second call exercises the presqu / postsqu arguments (submodule.py:152-164)
    out2, pre2, post2 = m(x, pre.clone(), post.clone())
    return dict(out=out.numpy(), pre=pre.numpy(), post=post.numpy(),
                out2=out2.numpy(), pre2=pre2.numpy(), post2=post2.numpy())


def case_hourglass2d():
    """submodule.hourglass2d (:317-361) and hourglass2d_downsample_16 (:270-315), the 2-D BEV hourglasses."""
    m = ref_sub.hourglass2d(64, gn=False).eval()
    m.load_state_dict(synth.det_state_dict(m, 51), strict=True)
    x = torch.from_numpy(synth.det_uniform((1, 64, 24, 16), 103))
    out, pre, post = m(x, None, None)
    m16 = ref_sub.hourglass2d_downsample_16(64, gn=False).eval()
    m16.load_state_dict(synth.det_state_dict(m16, 52), strict=True)
    x16 = torch.from_numpy(synth.det_uniform((1, 64, 48, 32), 104))
    return dict(out=out.numpy(), pre=pre.numpy(), post=post.numpy(), out16=m16(x16).numpy())


def case_hg16():
    m = ref_sub.hourglass_downsample_16(32, gn=False).eval()
    m.load_state_dict(synth.det_state_dict(m, 21), strict=True)
    x = torch.from_numpy(synth.det_uniform((1, 32, 16, 16, 16), 102))
    return dict(out=m(x).numpy())


def vernier_cfg(grid=(16, 32, 48), gn=False):
    ns = types.SimpleNamespace
    st = lambda nm, nb, blk, nbl, nch: ns(num_modules=nm, num_branches=nb, block=blk, num_blocks=nbl,
                                          num_channels=nch, fuse_method="SUM")
    hr = ns(name="hrnet-w32", head_type="default", init_weights=False, pre_trained_path="",
            extra=ns(stage1=st(1, 1, "bottleneck", [4], [64]), stage2=st(1, 2, "basic", [4, 4], [32, 64]),
                     stage3=st(4, 3, "basic", [4, 4, 4], [32, 64, 128]),
                     stage4=st(3, 4, "basic", [4, 4, 4, 4], [32, 64, 128, 256])))
    return ns(vernier_type="BEV_type3", gn=gn, backbone="hrnet", hrnet=hr, hrfeat=ns(output_channel=32),
              num_parts=9, grid_resolution=list(grid), n_sample_h=grid[0], n_sample_w=grid[1],
              n_sample_l=grid[2], x_range=[-3.2, 3.2], y_range=[-1.6, 1.6], z_range=[-4.8, 4.8],
              resolution=[64, 64])


VERNIER_3D_PREFIXES = ("vimg_feat.", "conv1.", "conv2.", "conv3.", "conv4.", "hg_conv3d.", "fg_cls_head.")


def synth_roi_points(N, P, res, seed):
    """Pixel coordinates in [-0.1*res, 1.1*res): mostly inside the ROI, some outside (zeros padding)."""
    return synth.det_uniform((N, 2, P), seed, -0.1 * res, 1.1 * res, bf16=False)


def case_vernier():
    cfg = vernier_cfg()
    torch.manual_seed(0)
    m = ref_vernier.VernierScale(cfg).eval()
    nh, nw, nl = cfg.grid_resolution
    sd = m.state_dict()
    sub = torch.nn.Module()
    for p in ("vimg_feat", "conv1", "conv2", "conv3", "conv4", "hg_conv3d", "fg_cls_head"):
        setattr(sub, p, getattr(m, p))
    new = synth.det_state_dict(sub, 31)
    sd.update(new)
    # the 2-D BEV tail (conv5 / hm1 / hm2, vernier.py:289-314) and the coordinate head (:68-93) get deterministic
    # weights as well, so that `ncf` / `coordinates` below are reproducible from synth.py alone
    tail = torch.nn.Module()
    for p in ("conv5", "hm1", "hm2"):
        setattr(tail, p, getattr(m, p))
    sd.update(synth.det_state_dict(tail, 33))
    head = torch.nn.Module()
    head.coord_head = m.coord_head
    sd.update(synth.det_state_dict(head, 35))
    m.load_state_dict(sd, strict=True)
    N, P = 1, nh * nw * nl
    lf = torch.from_numpy(synth.det_uniform((N, 32, 16, 16), 201))
    rf = torch.from_numpy(synth.det_uniform((N, 32, 16, 16), 202))
    gl = torch.from_numpy(synth_roi_points(N, P, 64, 203))
    gr = torch.from_numpy(synth_roi_points(N, P, 64, 204))
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        vox = m.construct_voxel(lf, rf, gl.clone(), gr.clone())          # [1,64,16,32,48]
    cap = {}
    h = m.pool_3d.register_forward_hook(lambda mod, i, o: cap.__setitem__("pooled", o))
    ncf, occ, _, coords, _ = m.predict_3d_heatmaps(vox)
    h.remove()
    pooled = cap["pooled"]
    bev = pooled.reshape(pooled.shape[0], -1, pooled.shape[3], pooled.shape[4])
    v = vox.numpy()
    return dict(voxel_sub=v[:, :, ::2, ::4, ::4].copy(),
                voxel_chan_sum=v.astype(np.float64).sum(axis=(0, 2, 3, 4)),
                voxel_abs_sum=np.abs(v.astype(np.float64)).sum(),
                voxel_bev=bev.numpy(), occupancy=occ.numpy(), ncf=ncf.numpy(), coordinates=coords.numpy())


def case_grid_proj():
    """The reference's own `refinementDataset._generate_grid_proj` (KITTIRefinement_dataset.py:847-868, with
    `_init_3d_grid` :267-282, `_to_cam` :828-845, kitti_util.Calibration.project_rect_to_image, img_proc.affine_transform)
    on synthetic proposals (oracle.grid_proj.synthetic_case: inputs are regenerated from the seed, outputs stored)."""
    for name in ("imageio", "tensorboardX", "matplotlib.patches", "matplotlib.lines"):
        sys.modules.setdefault(name, types.ModuleType(name))
    from snvc.dataset import KITTIRefinement_dataset as ref_ds
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    from oracle import grid_proj as ogp
    c = ogp.synthetic_case()
    ds = object.__new__(ref_ds.refinementDataset)
    ds.cfg = types.SimpleNamespace(x_range=c["x_range"], y_range=c["y_range"], z_range=c["z_range"],
                                   grid_resolution=list(c["grid_resolution"]))
    ds._init_3d_grid()
    calib = lambda P: types.SimpleNamespace(project_rect_to_image=lambda pts, P=P: _project(pts, P))

    def _project(pts, P):
        from snvc.dataset import kitti_util
        cal = object.__new__(kitti_util.Calibration)
        cal.P = P
        return cal.project_rect_to_image(pts)

    meta = {"trans_l": c["trans_l"], "trans_r": c["trans_r"]}
    coord_l, coord_r, grid_3d = ds._generate_grid_proj(c["samples"], calib(c["P_left"]), calib(c["P_right"]), meta)
    return dict(coord_l=coord_l, coord_r=coord_r, grid_3d=grid_3d.astype(np.float32))


if __name__ == "__main__":
    cases = {"hourglass_bn": lambda: case_hourglass(False), "hourglass_gn": lambda: case_hourglass(True),
             "hg16_bn": case_hg16, "hourglass2d_bn": case_hourglass2d, "vernier_bev3": case_vernier, "grid_proj": case_grid_proj}
    if len(sys.argv) > 1:
        cases = {k: v for k, v in cases.items() if k in sys.argv[1:]}
    for name, fn in cases.items():
        d = fn()
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **{k: np.asarray(v) for k, v in d.items()})
        print(name, {k: (np.asarray(v).shape, float(np.abs(v).max())) for k, v in d.items()})
