"""Host-side multi-GPU logic on CPU: world_size-2 `gloo` process groups.

* shard_range / gather_outputs: pair / proposal sharding is a partition, the gather restores order;
* DepthSlab + exchange_depth_halo + SlabTrunk: the depth-slab split of the global trunk with a halo
  exchange after every layer reproduces the unsplit trunk.  The layer executor here is plain torch on
  the CPU (oracle.blocks modules driven through the product's `.fused(...)` layer interface on NDHWC
  tensors) -- the slab algebra, views and exchange pattern are the product code under test
  (snvc_b200/parallel.py); the CUDA kernels are covered by the `-m gpu` tests.
"""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests", "golden")):
    if p not in sys.path:
        sys.path.insert(0, p)

from snvc_b200 import parallel as par   # noqa: E402  (pure host logic: importable without the CUDA library)


def test_shard_range_is_a_partition():
    for n in (0, 1, 7, 8, 64, 65):
        for world in (1, 2, 3, 8):
            spans = [par.shard_range(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for (a, b), (c, d) in zip(spans, spans[1:]):
                assert b == c and a <= b and c <= d
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        par.shard_range(4, 2, 2)


def test_depth_slab_geometry():
    s = par.DepthSlab(96, 8, 3)
    assert (s.d0, s.Dl) == (36, 12) and not s.first and not s.last
    assert s.ext_bins() == list(range(34, 50))
    with pytest.raises(ValueError):
        par.DepthSlab(48, 8, 0)        # 6-plane slabs: not a multiple of 4
    with pytest.raises(ValueError):
        par.DepthSlab(32, 8, 0)        # 4-plane slabs: thinner than 4 * HALO


def test_slab_z_range_partitions_the_voxel_grid():
    zs = np.arange(2.0, 40.4 - 1e-10, 0.2, dtype=np.float32) + np.float32(0.1)
    spans = [par.slab_z_range(zs, 2.4, 40.0, 48, par.DepthSlab(48, 4, r)) for r in range(4)]
    assert spans[0][0] == 0 and spans[-1][1] == len(zs)
    for (a, b), (c, d) in zip(spans, spans[1:]):
        assert b == c


# ------------------------------------------------------------------------------------ workers
def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _init(rank, world, port):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)


def _gather_worker(rank, world, port, q):
    _init(rank, world, port)
    try:
        n_total = 5
        full = torch.arange(n_total * 3, dtype=torch.float32).view(n_total, 3)
        lo, hi = par.shard_range(n_total, world, rank)
        local, = par.shard_batch([full], world, rank)
        assert local.shape[0] == hi - lo
        out = par.gather_outputs(local * 2, n_total)
        q.put((rank, bool(torch.equal(out, full * 2))))
    finally:
        dist.destroy_process_group()


class _TorchLayer:
    """oracle.blocks conv(+bn)(+relu) group behind the product's fused-layer interface (NDHWC fp32, CPU)."""

    def __init__(self, mod, transposed=False, relu_default=False):
        seq = mod[0] if isinstance(mod[0], torch.nn.Sequential) else mod    # _cbr -> (convbn_3d, ReLU)
        self.conv, self.bn = seq[0], seq[1]
        self.transposed, self.relu_default = transposed, relu_default
        self.cout = self.conv.out_channels

    def fused(self, x, relu=None, residual=None, residual_mode=0, out=None):
        relu = self.relu_default if relu is None else relu
        y = self.bn(self.conv(x.permute(0, 4, 1, 2, 3))).permute(0, 2, 3, 4, 1)
        if residual is not None and residual_mode in (0, 1):
            y = y + residual
        if relu:
            y = torch.relu(y)
        if residual is not None and residual_mode == 2:
            y = y + residual
        if out is not None:
            out.copy_(y)
            return out
        return y.contiguous()


def _wrap_trunk(trunk):
    import types
    m = types.SimpleNamespace()
    m.dres0 = [_TorchLayer(trunk.dres0[0], relu_default=True), _TorchLayer(trunk.dres0[1], relu_default=True)]
    m.dres1 = [_TorchLayer(trunk.dres1[0], relu_default=True), _TorchLayer(trunk.dres1[1])]
    hg = trunk.hg
    m.hg = types.SimpleNamespace(
        conv1=_TorchLayer(hg.conv1, relu_default=True), conv2=_TorchLayer(hg.conv2),
        conv3=_TorchLayer(hg.conv3, relu_default=True), conv4=_TorchLayer(hg.conv4, relu_default=True),
        conv5=_TorchLayer(hg.conv5, transposed=True), conv6=_TorchLayer(hg.conv6, transposed=True))
    return m


class _CountingArena:
    """Stand-in for PeerArena on the CPU: hands out ordinary tensors and records what SlabTrunk carves, in order."""

    ALIGN = par.PeerArena.ALIGN

    def __init__(self):
        self.shapes, self.elements = [], 0

    def reset(self):
        self.shapes, self.elements = [], 0

    def empty(self, shape, dtype=torch.float32):
        self.shapes.append(tuple(int(v) for v in shape))
        n = int(np.prod(shape))
        self.elements += (n + self.ALIGN - 1) // self.ALIGN * self.ALIGN     # (alignment slack counted in elements: >= bytes / 2)
        return torch.empty(tuple(shape), dtype=dtype)

    def owns(self, x):
        return False                                   # -> the exchange takes the torch.distributed / boundary path


@pytest.mark.parametrize("world,rank", [(1, 0), (2, 0), (4, 1), (4, 3)])
def test_slab_arena_bytes_bounds_what_the_trunk_carves(world, rank, monkeypatch):
    """`slab_arena_bytes` must bound the slabs one forward carves from a PeerArena (the arena is sized with it before any
    rank has run a layer), and every rank must carve the same sequence -- that is what makes `offset here == offset in the
    neighbour's mapping` true.  CPU: the trunk runs plain torch layers, the exchange is stubbed out."""
    import synth
    from oracle import blocks as oblocks
    monkeypatch.setattr(par, "exchange_depth_halo", lambda x, slab, group=None, comm=None: x)
    D, H, W, Cin, ch = 32, 8, 12, 8, 4
    with torch.no_grad():
        trunk = oblocks.GlobalTrunk(Cin, ch).eval()
        trunk.load_state_dict(synth.det_state_dict(trunk, 7))
        slab = par.DepthSlab(D, world, rank)
        arena = _CountingArena()
        ext = torch.zeros((1, slab.Dl + 2 * par.HALO, H, W, Cin))
        out = par.SlabTrunk(_wrap_trunk(trunk), slab, arena=arena)(ext)
    assert tuple(out.shape) == (1, slab.Dl + 2 * par.HALO, H, W, ch)
    assert len(arena.shapes) == 10                                      # one extended slab per layer of the trunk
    assert 2 * arena.elements <= par.slab_arena_bytes(slab, H, W, ch)    # bf16 bytes incl. alignment slack
    other = _CountingArena()                                             # the carving depends on the slab thickness only
    with torch.no_grad():
        par.SlabTrunk(_wrap_trunk(trunk), par.DepthSlab(D, world, (rank + 1) % world), arena=other)(ext)
    assert other.shapes == arena.shapes


def _slab_worker(rank, world, port, q):
    _init(rank, world, port)
    try:
        import synth
        from oracle import blocks as oblocks
        D, H, W, Cin = 16, 8, 12, 8
        with torch.no_grad():
            trunk = oblocks.GlobalTrunk(Cin, 4).eval()
            trunk.load_state_dict(synth.det_state_dict(trunk, 7))
            cost = torch.from_numpy(synth.det_uniform((1, Cin, D, H, W), 3, bf16=False))
            want = trunk(cost).permute(0, 2, 3, 4, 1)                       # [1, D, H, W, ch]
            slab = par.DepthSlab(D, world, rank)
            bins = slab.ext_bins()
            ext = torch.zeros((1, len(bins), H, W, Cin))
            for i, b in enumerate(bins):                                    # cost-volume bins are independent
                if 0 <= b < D:
                    ext[:, i] = cost[:, :, b].permute(0, 2, 3, 1)
            got = par.SlabTrunk(_wrap_trunk(trunk), slab)(ext)
            h = par.HALO
            mine = got[:, h:h + slab.Dl]
            ref = want[:, slab.d0:slab.d0 + slab.Dl]
            err = float((mine - ref).abs().max() / ref.abs().max())
            # the INNER halo planes hold the neighbours' adjacent planes (or zeros at the global boundary); the outer
            # ones only keep stride-2 levels aligned and are never read for a real output
            lo_ok = bool((got[:, h - 1] == 0).all()) if slab.first else \
                float((got[:, h - 1] - want[:, slab.d0 - 1]).abs().max()) < 1e-5
            hi_ok = bool((got[:, -h] == 0).all()) if slab.last else \
                float((got[:, -h] - want[:, slab.d0 + slab.Dl]).abs().max()) < 1e-5
        q.put((rank, err, lo_ok, hi_ok))
    finally:
        dist.destroy_process_group()


def _run(worker, world=2):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return sorted(res)


def test_gather_outputs_world2():
    assert _run(_gather_worker) == [(0, True), (1, True)]


def test_depth_slab_trunk_matches_unsplit_world2():
    res = _run(_slab_worker)
    for rank, err, lo_ok, hi_ok in res:
        assert err <= 1e-5, (rank, err)
        assert lo_ok and hi_ok, (rank, lo_ok, hi_ok)
