"""ORACLE (test infrastructure): zero-padded bilinear / trilinear grid sampling on the CPU.

The reference calls ``torch.nn.functional.grid_sample`` with its defaults
(snvc/models/vernier.py:339-340: mode='bilinear', padding_mode='zeros',
align_corners=False) -- the arithmetic lives in the third-party dependency torch
(pinned by the reference to pytorch-1.9.0, spec-file.txt:253; 2.11.0 here), whose
published algorithm (ATen/native/GridSampler.h:27-36 ``grid_sampler_unnormalize``
and GridSampler.cpp ``grid_sampler_{2d,3d}_cpu_impl``) is restated below in
numpy, one fp32 rounding per written operation, no FMA.

The restatement is pinned in tests/test_oracle_grid_sample.py against torch's own
CPU kernel run in this container (values) and by construction defines the
bit-exact corner indices / in-bounds masks the CUDA kernels must reproduce.
"""
import numpy as np

F32 = np.float32


def unnormalize(coord, size, align_corners):
    """ATen GridSampler.h:27-36."""
    coord = coord.astype(F32)
    if align_corners:
        return (((coord + F32(1)) / F32(2)).astype(F32) * F32(size - 1)).astype(F32)
    return ((((coord + F32(1)).astype(F32) * F32(size)).astype(F32) - F32(1)).astype(F32) / F32(2)).astype(F32)


def roi_normalize(px, res):
    """vernier.py:335-338: ``p / resolution * 2 - 1`` (three fp32 ops)."""
    px = px.astype(F32)
    return (((px / F32(res)).astype(F32) * F32(2)).astype(F32) - F32(1)).astype(F32)


def corners_2d(gx, gy, W, H, align_corners=False):
    """Returns ix, iy (fp32 source coords), ix_nw, iy_nw (int64 floor) for a normalised grid."""
    ix = unnormalize(gx, W, align_corners)
    iy = unnormalize(gy, H, align_corners)
    return ix, iy, np.floor(ix).astype(np.int64), np.floor(iy).astype(np.int64)


def grid_sample_2d(inp, grid, align_corners=False):
    """inp [N,C,H,W] fp32, grid [N,Ho,Wo,2] (x,y normalised) -> [N,C,Ho,Wo].

    Accumulation order nw, ne, sw, se (GridSampler.cpp grid_sampler_2d_cpu_impl)."""
    inp = np.ascontiguousarray(inp, dtype=F32)
    N, C, H, W = inp.shape
    gx, gy = grid[..., 0], grid[..., 1]
    ix, iy, x0, y0 = corners_2d(gx, gy, W, H, align_corners)
    x1, y1 = x0 + 1, y0 + 1
    fx0, fy0, fx1, fy1 = (a.astype(F32) for a in (x0, y0, x1, y1))
    w_nw = ((fx1 - ix).astype(F32) * (fy1 - iy).astype(F32)).astype(F32)
    w_ne = ((ix - fx0).astype(F32) * (fy1 - iy).astype(F32)).astype(F32)
    w_sw = ((fx1 - ix).astype(F32) * (iy - fy0).astype(F32)).astype(F32)
    w_se = ((ix - fx0).astype(F32) * (iy - fy0).astype(F32)).astype(F32)
    out = np.zeros((N, C) + gx.shape[1:], dtype=F32)
    n_idx = np.arange(N).reshape(N, *([1] * (gx.ndim - 1)))
    for (yy, xx, ww) in ((y0, x0, w_nw), (y0, x1, w_ne), (y1, x0, w_sw), (y1, x1, w_se)):
        ok = (xx >= 0) & (xx < W) & (yy >= 0) & (yy < H)
        xs, ys = np.where(ok, xx, 0), np.where(ok, yy, 0)
        v = inp[n_idx, :, ys, xs]                      # [N,Ho,Wo,C]
        v = np.moveaxis(v, -1, 1)
        term = (v * ww[:, None]).astype(F32)
        out = np.where(ok[:, None], (out + term).astype(F32), out)
    return out


def corners_3d(gx, gy, gz, W, H, D, align_corners=False):
    ix = unnormalize(gx, W, align_corners)
    iy = unnormalize(gy, H, align_corners)
    iz = unnormalize(gz, D, align_corners)
    return ix, iy, iz, np.floor(ix).astype(np.int64), np.floor(iy).astype(np.int64), np.floor(iz).astype(np.int64)


def grid_sample_3d(inp, grid, align_corners=False):
    """inp [N,C,D,H,W] fp32, grid [N,Do,Ho,Wo,3] (x,y,z) -> [N,C,Do,Ho,Wo].

    Corner order tnw,tne,tsw,tse,bnw,bne,bsw,bse; weight = (dx*dy)*dz
    (GridSampler.cpp grid_sampler_3d_cpu_impl)."""
    inp = np.ascontiguousarray(inp, dtype=F32)
    N, C, D, H, W = inp.shape
    gx, gy, gz = grid[..., 0], grid[..., 1], grid[..., 2]
    ix, iy, iz, x0, y0, z0 = corners_3d(gx, gy, gz, W, H, D, align_corners)
    x1, y1, z1 = x0 + 1, y0 + 1, z0 + 1
    f = lambda a: a.astype(F32)
    dx1, dx0 = (f(x1) - ix).astype(F32), (ix - f(x0)).astype(F32)   # weight of low / high corner
    dy1, dy0 = (f(y1) - iy).astype(F32), (iy - f(y0)).astype(F32)
    dz1, dz0 = (f(z1) - iz).astype(F32), (iz - f(z0)).astype(F32)
    out = np.zeros((N, C) + gx.shape[1:], dtype=F32)
    n_idx = np.arange(N).reshape(N, *([1] * (gx.ndim - 1)))
    for (zz, wz) in ((z0, dz1), (z1, dz0)):
        for (yy, wy) in ((y0, dy1), (y1, dy0)):
            for (xx, wx) in ((x0, dx1), (x1, dx0)):
                ww = ((wx * wy).astype(F32) * wz).astype(F32)
                ok = (xx >= 0) & (xx < W) & (yy >= 0) & (yy < H) & (zz >= 0) & (zz < D)
                xs, ys, zs = np.where(ok, xx, 0), np.where(ok, yy, 0), np.where(ok, zz, 0)
                v = np.moveaxis(inp[n_idx, :, zs, ys, xs], -1, 1)
                term = (v * ww[:, None]).astype(F32)
                out = np.where(ok[:, None], (out + term).astype(F32), out)
    return out


def roi_voxel_sample(left, right, l_pts, r_pts, nh, nw, nl, resolution):
    """vernier.py:323-349 (aggregate='concat').

    left,right [N,F,Hf,Wf]; l_pts,r_pts [N,2,P] pixel coords (P = nh*nw*nl, (h,w,l) C-order);
    resolution = cfg.resolution -- x is divided by resolution[1], y by resolution[0]
    (vernier.py:335-338, quirk preserved).  Returns [N,2F,nh,nw,nl]."""
    N, Fc = left.shape[:2]
    outs = []
    for feat, pts in ((left, l_pts), (right, r_pts)):
        g = np.transpose(pts, (0, 2, 1)).reshape(N, nh, nw * nl, 2)          # :332
        gx = roi_normalize(g[..., 0], resolution[1])                         # :335
        gy = roi_normalize(g[..., 1], resolution[0])                         # :336
        s = grid_sample_2d(feat, np.stack([gx, gy], -1), align_corners=False)  # :339
        outs.append(s.reshape(N, Fc, nh, nw, nl))
    return np.concatenate(outs, axis=1)                                      # :346
