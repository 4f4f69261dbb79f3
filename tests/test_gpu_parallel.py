"""GPU parity of the depth-slab-parallel global hot path (stress configuration) against
the unsplit single-GPU path: same kernels, slabs + halo exchange + z-partitioned lift.

world 1 runs in-process (exercises the extended-slab layout and the d_base/d_total lift); world 2 runs
two processes -- NCCL on two GPUs when the box has them, otherwise gloo (host-staged halos) with both
ranks on cuda:0, so the exchange path is exercised on a single-GPU box as well."""
import os
import socket
import sys
import types

import numpy as np
import pytest
from conftest import set_opt
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests", "golden")):
    if p not in sys.path:
        sys.path.insert(0, p)

pytestmark = pytest.mark.gpu


def _cfg_and_inputs(D=16):
    import synth
    from oracle import global_branch as ogb
    geom = ogb.GlobalGeometry(IH=64, IW=192, D=D, depth_min=2.0, depth_max=14.8, X_MIN=-6.0, X_MAX=6.0, Y_MIN=-1.0,
                              Y_MAX=2.0, Z_MIN=2.0, Z_MAX=14.0, VOXEL_X_SIZE=0.4, VOXEL_Y_SIZE=0.5, VOXEL_Z_SIZE=0.25,
                              align_corners=True,
                              P=np.array([[110.0, 0, 96.0, 6.0], [0, 110.0, 30.0, 0.03], [0, 0, 1.0, 0.0003]], np.float32))
    cv = geom.cv_ranges()
    cfg = types.SimpleNamespace(X_MIN=geom.X_MIN, X_MAX=geom.X_MAX, Y_MIN=geom.Y_MIN, Y_MAX=geom.Y_MAX, Z_MIN=geom.Z_MIN,
                                Z_MAX=geom.Z_MAX, VOXEL_X_SIZE=geom.VOXEL_X_SIZE, VOXEL_Y_SIZE=geom.VOXEL_Y_SIZE,
                                VOXEL_Z_SIZE=geom.VOXEL_Z_SIZE, CV_X_MIN=cv[0], CV_X_MAX=cv[1], CV_Y_MIN=cv[2],
                                CV_Y_MAX=cv[3], CV_Z_MIN=cv[4], CV_Z_MAX=cv[5], align_corners=True, GN=False)
    H, W = geom.IH // 4, geom.IW // 4
    lf, rf = synth.det_uniform((1, 32, H, W), 301), synth.det_uniform((1, 32, H, W), 302)
    return cfg, lf, rf, np.ascontiguousarray(geom.shifts(1)), geom.P[None].copy()


def _model(cfg, dev):
    import synth
    from snvc_b200.models.stereonet import GlobalHotPath
    m = GlobalHotPath(cfg).eval()
    m.load_state_dict(synth.det_state_dict(m, 41), strict=True)
    return m.to(dev)


def _slab_run(rank, world, dev, use_comm=False, use_arena=False):
    from snvc_b200 import parallel as par
    cfg, lf, rf, shift, P = _cfg_and_inputs()
    with torch.no_grad():
        m = _model(cfg, dev)
        args = [torch.from_numpy(a).to(dev) for a in (lf, rf, shift, P)]
        full = m(*args)                                                   # [1, C, Z, Y, X] fp32
        slab = par.DepthSlab(shift.shape[1], world, rank)
        comm = par.HaloComm(world, rank, dev) if use_comm else None      # C-ABI snvc_halo_exchange over its own NCCL communicator
        # peer-memory path: slabs in a CUDA-IPC-mapped arena, snvc_halo_push (NVLink peer stores + neighbour barrier)
        arena = par.PeerArena(world, rank, dev, par.slab_arena_bytes(slab, lf.shape[2], lf.shape[3])) if use_arena else None
        for _ in range(3 if use_arena else 1):                            # repeated forwards reuse the arena (epochs advance)
            part, (zlo, zhi) = par.slab_global_forward(m, *args, slab, comm=comm, arena=arena)
        part = part.clone()
        torch.cuda.synchronize(dev)
        if comm is not None:
            comm.close()
        if arena is not None:
            arena.close()
    ref = full[:, :, zlo:zhi]
    err = float((part - ref).abs().max() / full.abs().max()) if zhi > zlo else 0.0
    return zlo, zhi, err, int(full.shape[2])


# Two conv paths.  "kw" (SNVC_CONV_MODE=kw, the single-CTA kernels): every accumulator receives its taps in the same
# order whatever the launch geometry, so slab and unsplit runs are bit-identical and the bookkeeping is checked exactly.
# Default (CTA-pair kernel): planes whose accumulator-ring index is 0 or 1 are summed as primary + mirror block, and
# which planes those are depends on the plane's position in the CTA's march -- an fp32 re-association that flips a few
# bf16 roundings (one ulp = 2^-8) after five layers; the bar is the bf16 tolerance of the path (1e-2).
MODES = [("kw", 1e-6), (None, 1e-2)]


@pytest.mark.parametrize("mode,tol", MODES)
def test_slab_world1_matches_unsplit(mode, tol, monkeypatch):
    if mode:
        set_opt(monkeypatch, "SNVC_CONV_MODE", mode)
    zlo, zhi, err, Z = _slab_run(0, 1, torch.device("cuda", 0))
    assert (zlo, zhi) == (0, Z)
    assert err <= tol, err           # same planes, same weights: only the slab bookkeeping differs


def test_slab_world1_peer_arena_matches_unsplit(monkeypatch):
    """World 1 through the peer-memory halo path: the slabs live in a `PeerArena` (library-allocated, carved identically
    on every forward) and every layer's exchange is `snvc_halo_push` -- without neighbours it zero-fills both inner halo
    planes, the convolution's depth padding.  (World 2 on two GPUs: test_slab_peer_memory_halo_matches_unsplit.)"""
    set_opt(monkeypatch, "SNVC_CONV_MODE", "kw")
    zlo, zhi, err, Z = _slab_run(0, 1, torch.device("cuda", 0), use_arena=True)
    assert (zlo, zhi) == (0, Z) and err <= 1e-6, err


def test_graphed_slab_forward_world1_replays_eager():
    """`GraphedSlabForward` (the stress leg's launch mode): the captured slab forward replays to the eager result, also after
    the inputs were reloaded."""
    from snvc_b200 import parallel as par
    dev = torch.device("cuda", 0)
    cfg, lf, rf, shift, P = _cfg_and_inputs()
    with torch.no_grad():
        m = _model(cfg, dev)
        args = [torch.from_numpy(a).to(dev) for a in (lf, rf, shift, P)]
        slab = par.DepthSlab(shift.shape[1], 1, 0)
        want, zr = par.slab_global_forward(m, *args, slab, out_dtype=torch.bfloat16, layout_out="NDHWC")
        g = par.GraphedSlabForward(m, *[torch.zeros_like(a) for a in args[:2]], args[2], args[3], slab)
        g.load(*args)
        got, zr2 = g.replay()
        torch.cuda.synchronize()
        assert zr == zr2 and torch.equal(got.view(torch.int16), want.view(torch.int16))
        g.load(args[1], args[0], args[2], args[3])           # swapped views: a different result, then back
        other = g.replay()[0].clone()
        g.load(*args)
        again = g.replay()[0]
        torch.cuda.synchronize()
        assert not torch.equal(other, want) and torch.equal(again.view(torch.int16), want.view(torch.int16))
        g.close()


def _worker(rank, world, port, use_nccl, q, use_arena=False):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dev = torch.device("cuda", rank if use_nccl else 0)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl" if use_nccl else "gloo", rank=rank, world_size=world)
    try:
        q.put((rank,) + _slab_run(rank, world, dev, use_comm=use_nccl and not use_arena, use_arena=use_arena))
    except Exception as e:                                   # noqa: BLE001 -- the parent must not wait for a dead rank
        q.put((rank, "error", repr(e)))
        raise
    finally:
        dist.destroy_process_group()


def _collect(q, procs, world):
    res = [q.get(timeout=150) for _ in range(world)]
    bad = [r for r in res if len(r) > 1 and r[1] == "error"]
    for p in procs:
        p.join(timeout=30)
        if p.is_alive():
            p.kill()
    assert not bad, bad
    assert all(p.exitcode == 0 for p in procs)
    return sorted(res)


@pytest.mark.parametrize("mode,tol", MODES)
def test_slab_world2_matches_unsplit(mode, tol, monkeypatch):
    import torch.multiprocessing as mp
    if mode:
        set_opt(monkeypatch, "SNVC_CONV_MODE", mode)       # inherited by the spawned ranks
    use_nccl = torch.cuda.device_count() >= 2
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, use_nccl, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = _collect(q, procs, 2)
    (_, lo0, hi0, e0, Z), (_, lo1, hi1, e1, _) = res
    assert lo0 == 0 and hi0 == lo1 and hi1 == Z and hi0 > 0 and hi1 > lo1
    # identical inputs; slabs see identical planes after the exchange ("kw": bf16-identical features)
    assert e0 <= tol and e1 <= tol, (e0, e1)


@pytest.mark.parametrize("world", [2])                     # (the 16-plane test volume holds two slabs of >= 4*HALO planes)
def test_slab_peer_memory_halo_matches_unsplit(world, monkeypatch):
    """The product halo path of the stress configuration: slabs in CUDA-IPC-mapped peer memory, snvc_halo_push (peer stores
    over NVLink + neighbour barrier in one kernel) instead of NCCL.  Needs one GPU per rank (the push kernels of the
    ranks wait for each other on the device)."""
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    set_opt(monkeypatch, "SNVC_CONV_MODE", "kw")           # geometry-independent summation order: exact bookkeeping check
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, True, q, True)) for r in range(world)]
    for p in procs:
        p.start()
    res = _collect(q, procs, world)
    Z = res[0][4]
    assert res[0][1] == 0 and res[-1][2] == Z and all(res[i][2] == res[i + 1][1] for i in range(world - 1))
    assert all(r[3] <= 1e-6 for r in res), [r[3] for r in res]

