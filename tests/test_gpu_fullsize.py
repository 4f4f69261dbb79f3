"""GPU parity AT THE SHAPES THE BENCH RUNS (BASELINE.json configs[1] and configs[3]) against the CPU oracle.

  * one full-size global pair -- features [32,96,312], 48 depth bins, 192x20x304 voxels -- computed INSIDE the
    graph-replayed batch of 8 (the launch geometry bench.py times: pitch-42 tiles, the TMEM ring wrapping, the
    depth-split tail units on 148 SMs, the split first layer with its addend) vs oracle.torch_path.GlobalHotPathCPU
    (fp32 torch CPU ops); bar max|a-b|/max|b| <= 1e-2 and the top-100 BEV ordering of SURVEY.md 8(d);
  * one full-size instance proposal -- voxel grid [32,128,192], 64 channels -- through VernierHotPath vs
    oracle.blocks.Vernier3D; bar 1e-2 on the BEV features and the occupancy;
  * the instance ordering test of SURVEY.md 8(d) on 64 proposals at grid [16,64,96]: ordering of the per-proposal
    confidences and the arg-max cells identical to the fp32 oracle wherever the oracle's own margins exceed the
    bf16 noise.
The oracle costs a few seconds of host CPU per unit, which bounds the number of units compared."""
import types

import numpy as np
import pytest
import torch

import synth
from oracle import blocks as oblocks
from oracle import global_branch as ogb
from oracle import grid_sample as ogs
from oracle import torch_path

pytestmark = pytest.mark.gpu
TOL = 1e-2


@pytest.fixture(autouse=True)
def _no_grad():
    with torch.no_grad():
        yield


def _relerr(a, b):
    a, b = np.asarray(a, np.float32), np.asarray(b, np.float32)
    assert a.shape == b.shape, (a.shape, b.shape)
    return float(np.max(np.abs(a - b)) / max(1e-30, np.max(np.abs(b))))


def _topk_check(want, got, k=100):
    """Top-k BEV cells of a fixed random 1x1 head over the Y-pooled lifted volume (want/got [C,Z,Y,X] fp32)."""
    head = synth.det_uniform((want.shape[0],), 55, bf16=False)
    s_ref = np.einsum("czyx,c->zx", want, head).reshape(-1) / want.shape[2]
    s_got = np.einsum("czyx,c->zx", got, head).reshape(-1) / want.shape[2]
    top_ref = np.argsort(-s_ref, kind="stable")[:k]
    top_got = np.argsort(-s_got, kind="stable")[:k]
    assert set(top_ref[:k // 2]) <= set(top_got) and set(top_got[:k // 2]) <= set(top_ref)
    gaps = np.abs(np.diff(s_ref[top_ref]))
    noise = 2 * np.max(np.abs(s_ref - s_got))
    j = 0
    while j < k - 1 and gaps[j] > noise:
        j += 1
    assert np.array_equal(top_ref[:j], top_got[:j])          # identical order wherever the oracle separates the cells
    return j


def test_configs1_pairs_inside_batch_of_8_vs_oracle():
    from snvc_b200.models.stereonet import GlobalHotPath, GraphedHotPath
    from snvc_b200.utils.geometry import KITTI_P2, kitti_global_cfg, plane_sweep_shifts
    cfg = kitti_global_cfg()
    m = GlobalHotPath(cfg).eval()
    sd = synth.det_state_dict(m, 41)
    m.load_state_dict(sd, strict=True)
    m = m.cuda()
    B, C, H, W, D = 8, 32, 96, 312, 48
    g = torch.Generator().manual_seed(10)
    lf, rf = torch.randn((B, C, H, W), generator=g), torch.randn((B, C, H, W), generator=g)
    shift = torch.from_numpy(plane_sweep_shifts(cfg, B))
    Ps = np.stack([KITTI_P2 * np.float32([[1.0], [1.0 + 0.004 * i], [1.0]]) for i in range(B)]).astype(np.float32)
    gp = GraphedHotPath(m, B, C, (H, W), D, torch.bfloat16, "NDHWC", stages=True)
    assert gp.split                                                             # the path bench.py times
    vox = gp(lf.cuda(), rf.cuda(), shift.cuda(), torch.from_numpy(Ps).cuda())
    assert tuple(vox.shape) == (B, 192, 20, 304, 32)
    # oracle: independent geometry object + plain torch blocks with the same state_dict keys
    geom = ogb.GlobalGeometry()
    trunk = oblocks.GlobalTrunk(2 * C, 32).eval()
    trunk.load_state_dict(sd, strict=True)
    cpu = torch_path.GlobalHotPathCPU(trunk, geom)
    for n in (0, 5):                                                            # first pair and a tail-unit pair
        want = cpu(lf[n:n + 1], rf[n:n + 1], shift[n:n + 1], torch.from_numpy(Ps[n:n + 1]))[0].numpy()   # [32,Z,Y,X]
        got = vox[n].float().permute(3, 0, 1, 2).cpu().numpy()
        err = _relerr(got, want)
        assert err <= TOL, (n, err)
        # voxels the oracle leaves exactly zero on every channel (outside the frustum) are exactly zero here too
        outside = np.all(want == 0, axis=0)
        assert outside.mean() > 0.3 and np.all(got[:, outside] == 0)
        _topk_check(want, got)


def _vernier_cfg(grid):
    ns = types.SimpleNamespace
    return ns(vernier_type="BEV_type3", gn=False, hrfeat=ns(output_channel=32), num_parts=9, grid_resolution=list(grid),
              n_sample_h=grid[0], n_sample_w=grid[1], n_sample_l=grid[2], resolution=[256, 256])


def _instance_inputs(N, grid, seed, Hf=64):
    P = grid[0] * grid[1] * grid[2]
    lf, rf = synth.det_uniform((N, 32, Hf, Hf), seed), synth.det_uniform((N, 32, Hf, Hf), seed + 1)
    gl = synth.det_uniform((N, 2, P), seed + 2, -25.6, 281.6, bf16=False)       # ROI 256 px: some points outside
    gr = synth.det_uniform((N, 2, P), seed + 3, -25.6, 281.6, bf16=False)
    return lf, rf, gl, gr


def test_configs3_full_size_proposal_vs_oracle():
    """[32,128,192] grid, 64-channel voxels (201 MB fp32 in the reference): sampling + 3-D CNN vs the fp32 oracle."""
    from snvc_b200.models.vernier import VernierHotPath
    grid = (32, 128, 192)
    cfg = _vernier_cfg(grid)
    m = VernierHotPath(cfg).eval()
    sd = synth.det_state_dict(m, 31)
    m.load_state_dict(sd, strict=True)
    m = m.cuda()
    lf, rf, gl, gr = _instance_inputs(1, grid, 700)
    bev, occ = m(*[torch.from_numpy(a).cuda() for a in (lf, rf, gl, gr)])
    ref = oblocks.Vernier3D(32, n_sample_w=grid[1]).eval()
    ref.load_state_dict(sd, strict=True)
    vox = ogs.roi_voxel_sample(lf, rf, gl, gr, *grid, (256, 256))               # [1,64,32,128,192] fp32
    want_bev, want_occ = ref(torch.from_numpy(vox))
    assert _relerr(bev.cpu().numpy(), want_bev.numpy()) <= TOL
    assert _relerr(occ.cpu().numpy(), want_occ.numpy()) <= TOL


def test_instance_confidence_ordering_and_argmax_vs_oracle():
    """the ordering of the per-proposal confidences (mean over parts of max(ncf), vernier.py:683-686) over
    64 proposals matches the fp32 oracle's (>= 90 % of the ranks identical; swaps only between proposals the oracle separates
    by less than the bf16 noise), and so does the arg-max cell of the part heatmaps.  The proposals differ
    in feature amplitude (as real ROIs do), which spreads their confidences.  The 2-D tail (conv5 / hm1 / hm2,
    vernier.py:440-455) is evaluated by the SAME fp32 torch modules on both sides, so the comparison isolates the bf16
    3-D path under test."""
    from snvc_b200.models.vernier import VernierHotPath
    grid = (16, 64, 96)
    cfg = _vernier_cfg(grid)
    m = VernierHotPath(cfg).eval()
    sd = synth.det_state_dict(m, 31)
    m.load_state_dict(sd, strict=True)
    m = m.cuda()
    ref = oblocks.Vernier3D(32, n_sample_w=grid[1]).eval()
    ref.load_state_dict(sd, strict=True)
    tail = oblocks.VernierBevTail(128, 9, n_sample_w=grid[1]).eval()
    tail.load_state_dict(synth.det_state_dict(tail, 33), strict=True)
    NP, chunk = 64, 8
    amp = 0.3 + 1.4 * np.random.default_rng(4).permutation(NP).astype(np.float32) / (NP - 1)
    conf_ref, conf_got, arg_ref, arg_got, margin = [], [], [], [], []
    for c0 in range(0, NP, chunk):
        lf, rf, gl, gr = _instance_inputs(chunk, grid, 900 + 10 * c0, Hf=64)
        a = amp[c0:c0 + chunk, None, None, None]
        lf, rf = synth.bf16_round(lf * a), synth.bf16_round(rf * a)
        bev, _ = m(*[torch.from_numpy(x).cuda() for x in (lf, rf, gl, gr)])
        vox = ogs.roi_voxel_sample(lf, rf, gl, gr, *grid, (256, 256))
        want_bev, _ = ref(torch.from_numpy(vox))
        h_ref, h_got = tail(want_bev).numpy(), tail(bev.cpu()).numpy()          # [chunk,9,L,W]
        for i in range(chunk):
            fr, fg = h_ref[i].reshape(9, -1), h_got[i].reshape(9, -1)
            conf_ref.append(fr.max(1).mean()); conf_got.append(fg.max(1).mean())
            arg_ref.append(fr.argmax(1)); arg_got.append(fg.argmax(1))
            top2 = np.sort(fr, axis=1)[:, -2:]
            margin.append((top2[:, 1] - top2[:, 0], np.abs(fr - fg).max(1)))
    conf_ref, conf_got = np.array(conf_ref), np.array(conf_got)
    err = np.max(np.abs(conf_ref - conf_got))
    assert err <= TOL * np.max(np.abs(conf_ref)), err
    order_ref = np.argsort(-conf_ref, kind="stable")
    order_got = np.argsort(-conf_got, kind="stable")
    same = int((order_ref == order_got).sum())
    assert same >= int(0.9 * NP), (same, err)                                  # the ranking is the oracle's ...
    rank_got = np.empty(NP, np.int64)
    rank_got[order_got] = np.arange(NP)
    for a in range(NP):                                                        # ... and a pair is only ever swapped when the
        for b in range(a + 1, NP):                                             # oracle itself separates it by less than the noise
            i, j = order_ref[a], order_ref[b]
            if rank_got[i] > rank_got[j]:
                assert conf_ref[i] - conf_ref[j] <= 2 * err, (i, j, conf_ref[i] - conf_ref[j], err)
    agree = sum(int(arg_ref[i][p] == arg_got[i][p]) for i in range(NP) for p in range(9))
    assert agree >= 0.97 * NP * 9, agree
    for i in range(NP):
        gap, e = margin[i]
        for p in range(9):
            if gap[p] > 2 * e[p]:                                              # the oracle's own arg-max is unambiguous
                assert arg_ref[i][p] == arg_got[i][p]
