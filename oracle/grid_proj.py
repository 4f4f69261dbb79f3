"""ORACLE (test infrastructure): projection of the instance sampling grid ("Vernier scale") into the
left / right ROI feature frames -- SURVEY.md 8(f) N1, the CPU loop directly upstream of the ROI voxel
sampling kernel.

Restates, in numpy float64 exactly as the reference does, then one cast to float32:
  _init_3d_grid          snvc/dataset/KITTIRefinement_dataset.py:267-282   linspace / meshgrid(indexing='xy')
  _to_cam                KITTIRefinement_dataset.py:828-845                R_y(ry + pi/2) @ p + (x, y - h/2, z)
  project_rect_to_image  snvc/dataset/kitti_util.py:282-293                [p,1] @ P^T, divide by the third row
  affine_transform       snvc/utils/img_proc.py:71-74                      (trans @ [u,v,1]^T).astype(float32)
  _generate_grid_proj    KITTIRefinement_dataset.py:847-868                loop over proposals, concatenate

Pinned in tests/test_oracle_grid_proj.py against tests/golden/grid_proj.npz, which holds the outputs of the
reference's own `_generate_grid_proj` (imported from /root/reference by tests/golden/make_golden.py).
"""
import numpy as np


def grid_points(x_range, y_range, z_range, grid_resolution):
    """-> pts [3, P] float64 with P = nh*nw*nl, point index (ih*nw + iw)*nl + il  (_init_3d_grid + reshape(3,-1))."""
    nh, nw, nl = grid_resolution
    x_pts = np.linspace(x_range[0], x_range[1], nw)
    y_pts = np.linspace(y_range[0], y_range[1], nh)
    z_pts = np.linspace(z_range[0], z_range[1], nl)
    grid_x, grid_y, grid_z = np.meshgrid(x_pts, y_pts, z_pts, indexing='xy')     # each [nh, nw, nl]
    return np.concatenate([grid_x[None, :], grid_y[None, :], grid_z[None, :]]).reshape(3, -1)


def to_cam(pts_3d, sample):
    """sample = [h, w, l, x, y, z, ry]  (_to_cam)."""
    ry = sample[6] + 0.5 * np.pi
    rot_maty = np.array([[np.cos(ry), 0, np.sin(ry)],
                         [0, 1, 0],
                         [-np.sin(ry), 0, np.cos(ry)]])
    x, y, z = sample[3:6]
    center_y = y - sample[0] * 0.5
    return rot_maty @ pts_3d + np.array([[x], [center_y], [z]])


def project_rect_to_image(pts_3d_rect, P):
    """[n,3] float64, P [3,4] -> [n,2]  (kitti_util.Calibration.project_rect_to_image)."""
    n = pts_3d_rect.shape[0]
    hom = np.hstack((pts_3d_rect, np.ones((n, 1))))
    pts_2d = np.dot(hom, np.transpose(P))
    pts_2d[:, 0] /= pts_2d[:, 2]
    pts_2d[:, 1] /= pts_2d[:, 2]
    return pts_2d[:, 0:2]


def affine_transform(kpts_2d, trans, dtype=np.float32):
    """img_proc.affine_transform: [n,2], [2,3] -> [2,n] cast to float32."""
    homo = np.concatenate([kpts_2d, np.ones((len(kpts_2d), 1))], axis=1)
    return (trans @ homo.T).astype(dtype)


def generate_grid_proj(samples, P_left, P_right, trans_l, trans_r, x_range, y_range, z_range, grid_resolution):
    """-> coord_l [N,2,P] f32, coord_r [N,2,P] f32, grid_3d_cam [N,P,3] f64  (_generate_grid_proj)."""
    pts_3d = grid_points(x_range, y_range, z_range, grid_resolution)
    coord_l, coord_r, grid_3d = [], [], []
    for idx, sample in enumerate(samples):
        pts_cam = to_cam(pts_3d, np.asarray(sample, dtype=np.float64)).T
        grid_3d.append(pts_cam[None, :, :])
        coord_l.append(affine_transform(project_rect_to_image(pts_cam, P_left), trans_l[idx])[None, :, :])
        coord_r.append(affine_transform(project_rect_to_image(pts_cam, P_right), trans_r[idx])[None, :, :])
    return np.concatenate(coord_l), np.concatenate(coord_r), np.concatenate(grid_3d)


def synthetic_case(n=3, seed=20, grid_resolution=(8, 16, 24), resolution=(256, 256)):
    """Synthetic proposals in the spirit of SURVEY.md 8(d) cfg-4: boxes [h,w,l,x,y,z,ry] with x~U(-10,10), y=1.6,
    z~U(8,40), ry~U(-pi,pi); KITTI-typical P2 / P3; centre-crop affine of a square ROI around the projected centre."""
    rng = np.random.RandomState(seed)
    P2 = np.array([[721.5377, 0.0, 609.5593, 44.85728], [0.0, 721.5377, 172.854, 0.2163791], [0.0, 0.0, 1.0, 0.002745884]])
    P3 = P2.copy()
    P3[0, 3] = 44.85728 - 721.5377 * 0.54
    samples, tl, tr = [], [], []
    for _ in range(n):
        h, w, l = 1.5 + 0.2 * rng.rand(), 1.6 + 0.2 * rng.rand(), 3.8 + 0.6 * rng.rand()
        x, y, z, ry = rng.uniform(-10, 10), 1.6, rng.uniform(8, 40), rng.uniform(-np.pi, np.pi)
        samples.append([h, w, l, x, y, z, ry])
        for P, dst in ((P2, tl), (P3, tr)):
            c = project_rect_to_image(np.array([[x, y - h / 2, z]]), P)[0]
            half = 0.5 * 721.5377 * 6.4 / z                       # ROI covering the 6.4 m wide sampling grid
            x1, y1, x2, y2 = c[0] - half, c[1] - half, c[0] + half, c[1] + half
            m = np.zeros((2, 3), dtype=np.float32)                # img_proc.get_affine_trans_center_crop
            m[0, 0] = resolution[0] / (x2 - x1); m[0, 2] = -m[0, 0] * x1
            m[1, 1] = resolution[1] / (y2 - y1); m[1, 2] = -m[1, 1] * y1
            dst.append(m)
    return dict(samples=np.array(samples), P_left=P2, P_right=P3, trans_l=np.stack(tl), trans_r=np.stack(tr),
                x_range=(-3.2, 3.2), y_range=(-1.6, 1.6), z_range=(-4.8, 4.8), grid_resolution=tuple(grid_resolution))
