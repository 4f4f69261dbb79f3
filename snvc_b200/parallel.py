"""Multi-GPU partitioning of the hot path: one process per GPU (torchrun), torch.distributed plumbing.

The reference's only parallelism is single-process `nn.DataParallel` over the batch dimension
(tools/inference_agnostic.py:472: scatter proposals, replicate the model, gather on GPU 0).  Here:

* pairs (global branch) and proposals (instance branch) are independent units -> each rank takes a
  contiguous block (`shard_range`), weights are replicated, NO collective on the data path;
  `gather_outputs` reproduces DataParallel's gather for callers that want it;
* the single-volume stress case splits the DEPTH axis into slabs (`DepthSlab`): cost-volume bins are
  independent (BuildCostVolume_cuda.cu:84), every 3x3x3 conv needs neighbour planes, exchanged with
  point-to-point sends between ranks r-1 / r+1 (`exchange_depth_halo`; on GPUs the C-ABI `snvc_halo_exchange` over the
  library's own NCCL communicator, `HaloComm`; torch.distributed / gloo in the CPU tests), the lift is partitioned by
  the world-z range each slab covers.

`SlabTrunk` drives any module tree with the layer interface of snvc_b200.models.submodule
(`.fused(x, relu=, residual=, residual_mode=, out=)` on NDHWC tensors), so the slab algebra is
testable on CPU with a plain-torch layer executor (tests/test_parallel_gloo.py).
"""
import torch
import torch.distributed as dist

HALO = 2   # halo planes kept on each side of a slab (1 is needed by a 3^3 conv; 2 keeps stride-2 levels aligned)


# ------------------------------------------------------------------------------------ batch sharding
def shard_range(n, world, rank):
    """Contiguous block [lo, hi) of `n` units for `rank`: the first n % world ranks get one extra."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"bad world/rank {world}/{rank}")
    base, extra = divmod(int(n), world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_batch(tensors, world, rank):
    """Slice every tensor of `tensors` (all with the same leading dim) to this rank's block."""
    n = tensors[0].shape[0]
    lo, hi = shard_range(n, world, rank)
    return [t[lo:hi] for t in tensors]


def gather_outputs(local, n_total, group=None):
    """all_gather ragged leading-dim shards (as produced by `shard_range`) -> [n_total, ...] on every rank."""
    world = dist.get_world_size(group)
    sizes = [shard_range(n_total, world, r) for r in range(world)]
    width = max(hi - lo for lo, hi in sizes)
    pad = torch.zeros((width,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[:local.shape[0]] = local
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    return torch.cat([b[:hi - lo] for b, (lo, hi) in zip(bufs, sizes)], dim=0)


# ------------------------------------------------------------------------------------ depth slabs
class DepthSlab:
    """Rank `rank`'s slab of a D-plane volume split over `world` ranks.

    Slab boundaries are multiples of 4 (two stride-2 levels below the full resolution), every slab
    has at least 4*HALO planes so the 1/4-resolution level still owns HALO real planes to send."""

    def __init__(self, D, world, rank):
        if D % (4 * world) != 0:
            raise ValueError(f"depth {D} must be a multiple of 4*world ({4 * world})")
        self.D, self.world, self.rank = D, world, rank
        self.Dl = D // world
        if world > 1 and self.Dl < 4 * HALO:
            raise ValueError(f"slab of {self.Dl} planes is thinner than {4 * HALO}")
        self.d0 = rank * self.Dl

    @property
    def first(self):
        return self.rank == 0

    @property
    def last(self):
        return self.rank == self.world - 1

    def ext_bins(self):
        """Global plane indices of the extended slab [d0-HALO, d0+Dl+HALO) (may fall outside [0, D))."""
        return list(range(self.d0 - HALO, self.d0 + self.Dl + HALO))


class HaloComm:
    """An NCCL communicator of the library's own (snvc_halo_comm_create) for the depth-slab halo exchange: the 128-byte
    ncclUniqueId is created on rank 0 and broadcast through torch.distributed; afterwards an exchange is ONE C call
    (ncclGroup of <= 2 sends + 2 receives on the caller's stream) instead of four P2POp objects and a
    batch_isend_irecv per layer -- at 8 ranks the 10 exchanges of a volume otherwise cost more host time than the
    3 ms of kernels they separate."""

    def __init__(self, world, rank, device, group=None):
        import ctypes
        from snvc_b200 import _lib
        self.world, self.rank, self.device = world, rank, device
        self.comm = ctypes.c_void_p()
        if world == 1:
            return
        L = _lib.lib()
        buf = (ctypes.c_ubyte * 128)()
        if rank == 0:
            _lib.check(L.snvc_halo_unique_id(buf), "snvc_halo_unique_id")
        t = torch.tensor(list(buf), dtype=torch.uint8, device=device)
        dist.broadcast(t, src=_peer(0, group), group=group)
        host = t.cpu().tolist()
        for i, v in enumerate(host):
            buf[i] = v
        with torch.cuda.device(device):
            _lib.check(L.snvc_halo_comm_create(buf, world, rank, ctypes.byref(self.comm)), "snvc_halo_comm_create")

    def exchange(self, x):
        from snvc_b200 import _lib
        planes = x.shape[1]
        plane_bytes = x[0, 0].numel() * x.element_size()
        with torch.cuda.device(x.device):
            st = _lib.lib().snvc_halo_exchange(self.comm, x.data_ptr(), planes, plane_bytes, HALO, self.rank, self.world,
                                               _lib.stream_ptr())
        _lib.check(st, "snvc_halo_exchange")
        return x

    def close(self):
        from snvc_b200 import _lib
        if self.comm:
            _lib.lib().snvc_halo_comm_destroy(self.comm)
            self.comm = None


class _RawCuda:
    """`__cuda_array_interface__` view of raw device memory (torch.as_tensor wraps it without a copy)."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(ptr), False), "version": 2}


class PeerArena:
    """Slab buffers in NVLink / NVSwitch PEER MEMORY: the product path of the depth-slab halo exchange (snvc_halo_push).

    Every rank allocates one arena with the library (snvc_peer_alloc: a cudaMalloc block, so it can be exported through
    CUDA IPC), the 64-byte handles travel through `torch.distributed.all_gather`, and each rank maps the arenas of ranks
    r-1 and r+1.  All ranks carve their arena identically (`reset()` at the start of a forward, then the same sequence of
    `empty()` calls), so a slab at offset o here is at offset o in a neighbour's mapping, and `push(x)` can store this
    slab's first / last real plane straight into the neighbours' inner halo planes -- one kernel that is also the neighbour
    barrier -- instead of an ncclSend / ncclRecv group (146 GB/s per direction on the 8 x B200 box)."""

    ALIGN = 256

    def __init__(self, world, rank, device, nbytes, group=None):
        import ctypes
        from snvc_b200 import _lib
        self.world, self.rank, self.device, self.group = world, rank, device, group
        L = _lib.lib()
        self.ctl_bytes = int(L.snvc_peer_ctl_bytes())
        self.nbytes = int(nbytes) + self.ctl_bytes + self.ALIGN
        self.base, self.lo, self.hi = ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_void_p()
        on_cpu = world > 1 and dist.get_backend(group) == "gloo"       # (CPU-side tests: both ranks on one GPU, host tensors)
        flag_dev = "cpu" if on_cpu else device

        def agree(err):
            """Collective: every rank learns whether ANY rank failed the local step just done (a rank that raised on its own
            would leave the others waiting in the next collective).  Raises on all ranks alike."""
            bad = torch.tensor([1 if err else 0], dtype=torch.int32, device=flag_dev)
            if world > 1:
                dist.all_reduce(bad, op=dist.ReduceOp.MAX, group=group)
            if int(bad.item()):
                self._release(L)
                raise RuntimeError(f"PeerArena: peer-memory set-up failed on some rank ({err or 'another rank'})")

        handle = (ctypes.c_ubyte * 64)()
        err = None
        try:                                                    # local: allocate + export
            with torch.cuda.device(device):
                _lib.check(L.snvc_peer_alloc(self.nbytes, ctypes.byref(self.base)), "snvc_peer_alloc")
                _lib.check(L.snvc_peer_export(self.base, handle), "snvc_peer_export")
        except Exception as e:                                  # noqa: BLE001
            err = str(e)[:200]
        agree(err)
        every = []
        if world > 1:
            mine = torch.tensor(list(handle), dtype=torch.uint8, device=flag_dev)
            every = [torch.empty_like(mine) for _ in range(world)]
            dist.all_gather(every, mine, group=group)
        try:                                                    # local: map the neighbours' arenas
            with torch.cuda.device(device):
                for attr, peer in (("lo", rank - 1), ("hi", rank + 1)):
                    if 0 <= peer < world:
                        raw = (ctypes.c_ubyte * 64)(*every[peer].cpu().tolist())
                        _lib.check(L.snvc_peer_open(raw, ctypes.byref(getattr(self, attr))), "snvc_peer_open")
            self.mem = torch.as_tensor(_RawCuda(self.base.value, self.nbytes), device=device)
        except Exception as e:                                  # noqa: BLE001
            err = str(e)[:200]
        agree(err)                                              # also the barrier: every mapping exists before anyone pushes
        self.off = self.ctl_bytes

    def _release(self, L):
        with torch.cuda.device(self.device):
            for h in (self.lo, self.hi):
                if h:
                    L.snvc_peer_close(h)
            if self.base:
                L.snvc_peer_free(self.base)
        self.mem = None
        self.base = self.lo = self.hi = None

    def reset(self):
        self.off = self.ctl_bytes

    def empty(self, shape, dtype=torch.bfloat16):
        n = 1
        for v in shape:
            n *= int(v)
        nb = n * torch.empty((), dtype=dtype).element_size()
        start = (self.off + self.ALIGN - 1) // self.ALIGN * self.ALIGN
        if start + nb > self.nbytes:
            raise RuntimeError(f"PeerArena: {start + nb} bytes needed, {self.nbytes} allocated")
        self.off = start + nb
        return self.mem[start:start + nb].view(dtype).view(tuple(shape))

    def owns(self, x):
        return self.base.value <= x.data_ptr() < self.base.value + self.nbytes

    def push(self, x):
        """Halo exchange of the extended slab x [1, planes, H, W, C] (allocated by `empty`), in place, on the current stream."""
        from snvc_b200 import _lib
        o = x.data_ptr() - self.base.value
        planes = x.shape[1]
        plane_bytes = x[0, 0].numel() * x.element_size()
        lo = self.lo.value + o if self.lo else None
        hi = self.hi.value + o if self.hi else None
        with torch.cuda.device(x.device):
            st = _lib.lib().snvc_halo_push(x.data_ptr(), lo, hi, planes, plane_bytes, HALO, self.base.value,
                                           self.lo.value if self.lo else None, self.hi.value if self.hi else None, 0,
                                           _lib.stream_ptr())
        _lib.check(st, "snvc_halo_push")
        return x

    def close(self):
        from snvc_b200 import _lib
        if self.base:
            torch.cuda.synchronize(self.device)
            sync = (lambda: dist.barrier(group=self.group)) if self.world > 1 else (lambda: None)
            sync()                                              # nobody unmaps an arena a neighbour may still push into
            L = _lib.lib()
            self.mem = None
            with torch.cuda.device(self.device):
                for h in (self.lo, self.hi):
                    if h:
                        L.snvc_peer_close(h)
                sync()
                L.snvc_peer_free(self.base)
            self.base = self.lo = self.hi = None


def slab_arena_bytes(slab, H, W, ch=32):
    """Upper bound of the arena bytes one `slab_global_forward` carves: five full-resolution ch-channel slabs (dres0 x 2,
    dres1 x 2, hourglass output), three half-resolution and two quarter-resolution 2*ch-channel slabs, each with 2*HALO
    extra planes, + alignment slack."""
    full = (slab.Dl + 2 * HALO) * H * W * ch * 2
    half = (slab.Dl // 2 + 2 * HALO) * ((H + 1) // 2) * ((W + 1) // 2) * 2 * ch * 2
    quarter = (slab.Dl // 4 + 2 * HALO) * ((H + 3) // 4) * ((W + 3) // 4) * 2 * ch * 2
    return 5 * full + 3 * half + 2 * quarter + 16 * PeerArena.ALIGN


def exchange_depth_halo(x, slab, group=None, comm=None):
    """x: [1, Dl_level + 2*HALO, H, W, C] extended slab at any pyramid level (in place).

    Every consumer of a slab reads at most ONE plane beyond the real ones -- a 3^3 stride-1 conv one on each side, a
    stride-2 conv (pad 1) only the lower one, a transposed conv (k3,s2,p1,op1) only the upper one -- so the exchange
    moves one plane per direction: the first / last real plane goes to rank r-1 / r+1 and lands in their INNER halo
    plane (index HALO-1 / -HALO); at the global boundary the inner halo plane is zeroed (the convolution's zero
    padding).  The outer halo planes only keep stride-2 levels aligned and are never read for a real output.
    `comm` (HaloComm): the C-ABI NCCL path (GPU); otherwise torch.distributed point-to-point ops (gloo in the CPU tests)."""
    h = HALO
    if x.shape[0] != 1 or x.shape[1] < 2 * h + 1:
        raise ValueError("slab too thin for the halo exchange (or batch != 1: depth slices must be contiguous)")
    if comm is not None and x.is_cuda:
        return comm.exchange(x)
    lo_halo, hi_halo = x[:, h - 1:h], x[:, -h:x.shape[1] - h + 1]            # contiguous views (batch 1)
    ops, stage_lo, stage_hi = [], None, None
    # gloo moves host memory only: stage through the CPU there (tests); NCCL sends device buffers over NVLink
    stage = slab.world > 1 and x.is_cuda and dist.get_backend(group) == "gloo"
    if slab.world > 1:
        if not slab.first:
            send_lo = x[:, h:h + 1]
            recv_lo = lo_halo
            if stage:
                send_lo, stage_lo = send_lo.cpu(), torch.empty(lo_halo.shape, dtype=x.dtype)
                recv_lo = stage_lo
            ops += [dist.P2POp(dist.isend, send_lo, _peer(slab.rank - 1, group), group),
                    dist.P2POp(dist.irecv, recv_lo, _peer(slab.rank - 1, group), group)]
        if not slab.last:
            send_hi = x[:, -h - 1:x.shape[1] - h]
            recv_hi = hi_halo
            if stage:
                send_hi, stage_hi = send_hi.cpu(), torch.empty(hi_halo.shape, dtype=x.dtype)
                recv_hi = stage_hi
            ops += [dist.P2POp(dist.isend, send_hi, _peer(slab.rank + 1, group), group),
                    dist.P2POp(dist.irecv, recv_hi, _peer(slab.rank + 1, group), group)]
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    if slab.first or slab.world == 1:
        lo_halo.zero_()
    elif stage_lo is not None:
        lo_halo.copy_(stage_lo)
    if slab.last or slab.world == 1:
        hi_halo.zero_()
    elif stage_hi is not None:
        hi_halo.copy_(stage_hi)
    return x


def _peer(rank_in_group, group):
    return dist.get_global_rank(group, rank_in_group) if group is not None else rank_in_group


class SlabTrunk:
    """The global trunk (dres0 / dres1 / hourglass, models/stereonet.py) on one depth slab, with a halo
    exchange after every layer.  `model` needs attributes dres0, dres1, hg with the fused-layer
    interface; tensors are NDHWC with N == 1 (depth slices of an N=1 volume are contiguous views)."""

    def __init__(self, model, slab, group=None, comm=None, arena=None):
        self.m, self.slab, self.group, self.comm, self.arena = model, slab, group, comm, arena

    def _xchg(self, x):
        if self.arena is not None and x.is_cuda and self.arena.owns(x):
            return self.arena.push(x)                            # NVLink peer stores + neighbour barrier in one kernel
        return exchange_depth_halo(x, self.slab, self.group, self.comm)

    def _empty(self, shape, like):
        """Every layer's extended output slab; from the arena when there is one (same carving order on every rank)."""
        if self.arena is not None:
            return self.arena.empty(shape, like.dtype)
        return torch.empty(shape, dtype=like.dtype, device=like.device)

    @staticmethod
    def _cout(layer):
        """Output channels of a fused layer: test executors carry `.cout`; the product's modules are
        Sequential(conv, norm) or Sequential(Sequential(conv, norm), ReLU)."""
        c = getattr(layer, "cout", None)
        if c:
            return c
        first = layer[0]
        return first.out_channels if hasattr(first, "out_channels") else first[0].out_channels

    def _s1(self, layer, x, residual=None, **kw):
        """stride-1 conv on an ext slab [1, Dl+2*HALO, ...]: only the real planes and the two inner halo planes are
        convolved (input view x[:, 1:-1], zero padding beyond it), so the slab costs (Dl + 2) planes of work instead of
        (Dl + 4); the two inner halo outputs lack a depth tap and are overwritten by the exchange."""
        cout = self._cout(layer)
        out = self._empty(tuple(x.shape[:-1]) + (cout,), x)
        out[:, 0].zero_()
        out[:, -1].zero_()
        if residual is not None:
            kw["residual"] = residual[:, 1:-1]
        layer.fused(x[:, 1:-1], out=out[:, 1:-1], **kw)
        return self._xchg(out)

    def _s2(self, layer, x, **kw):
        """stride-2 conv: ext [1, Dl+4, ...] -> ext [1, Dl/2+4, ...]; the conv output (Dl/2+2 planes, plane o
        centred on ext input plane 2o) is written into planes [1, -1) of the half-resolution slab."""
        N, De, H, W, _ = x.shape
        assert N == 1 and (De - 2 * HALO) % 2 == 0
        Do = (De - 2 * HALO) // 2 + 2 * HALO
        cout = self._cout(layer)
        out = self._empty((1, Do, (H + 1) // 2, (W + 1) // 2, cout), x)
        out[:, 0].zero_()
        out[:, -1].zero_()
        layer.fused(x, out=out[:, 1:-1], **kw)
        return self._xchg(out)

    def _up(self, layer, x, residual=None, **kw):
        """transposed conv (k3,s2,p1,op1): input planes [1, -1) of the ext slab -> ext slab at 2x resolution."""
        if residual is not None:
            kw["residual"] = residual
        if self.arena is not None:
            N, De, H, W, _ = x.shape
            kw["out"] = self._empty((N, 2 * (De - 2), 2 * H, 2 * W, self._cout(layer)), x)
        return self._xchg(layer.fused(x[:, 1:-1], **kw))

    def head_split(self, right_ext, addend):
        """dres0.conv1 on the SPLIT cost volume of a slab (models/stereonet.py:trunk_head_split): `right_ext` is the extended
        right-half slab [1, Dl+2*HALO, H, W, F], `addend` the fp32 [1,3,H,W,ch] share of the depth-constant left half.  The
        addend's edge variants belong to the first / last plane of the WHOLE volume, which sit one plane inside the
        convolved view on the first / last rank and nowhere on interior ranks."""
        plan = self.m._split_plans()[1]
        out = self._empty(tuple(right_ext.shape[:-1]) + (plan.cout,), right_ext)
        out[:, 0].zero_()
        out[:, -1].zero_()
        edges = (1 if self.slab.first else -1, 1 if self.slab.last else -1)
        plan(right_ext[:, 1:-1], relu=True, addend=addend, addend_edges=edges, out=out[:, 1:-1])
        return self._xchg(out)

    def __call__(self, cost=None, head=None):
        """cost: extended cost-volume slab [1, Dl+2*HALO, H, W, 2F] (inner halo planes valid, or zero at the boundary);
        or `head` = the output of `head_split` (dres0.conv1 already applied to the split volume)."""
        m = self.m
        x = head if head is not None else self._s1(m.dres0[0], cost)
        x = self._s1(m.dres0[1], x)
        y = self._s1(m.dres1[0], x)
        x = self._s1(m.dres1[1], y, residual=x, residual_mode=1)
        hg = m.hg
        o = self._s2(hg.conv1, x)
        pre = self._s1(hg.conv2, o, relu=True)
        o = self._s2(hg.conv3, pre)
        o = self._s1(hg.conv4, o)
        post = self._up(hg.conv5, o, relu=True, residual=pre, residual_mode=1)
        return self._up(hg.conv6, post, residual=x, residual_mode=1)


def slab_z_range(zs, cv_z_min, cv_z_max, D, slab, align_corners=True):
    """Voxel z-indices [zlo, zhi) whose lower depth-plane index floor(iz) falls into this slab
    (host float64 estimate; the slab's HALO planes absorb any fp32 disagreement with the kernel).
    Voxels in front of / behind the volume go to the first / last slab."""
    import numpy as np
    z = np.asarray(zs, dtype=np.float64)
    g = (z - cv_z_min) / (cv_z_max - cv_z_min) * 2 - 1
    iz = (g + 1) / 2 * (D - 1) if align_corners else ((g + 1) * D - 1) / 2
    plane = np.clip(np.floor(iz), 0, D - 1).astype(np.int64)
    owner = np.minimum(plane // slab.Dl, slab.world - 1)
    idx = np.nonzero(owner == slab.rank)[0]
    if idx.size == 0:
        return 0, 0
    assert np.all(np.diff(idx) == 1), "voxel z centres must be monotonic in depth"
    return int(idx[0]), int(idx[-1]) + 1


def _cached_z_range(model, D, slab):
    """slab_z_range of the model's voxel grid, computed once per (model, D, world, rank): it reads `model.zs` back to the
    host, which would otherwise synchronise every forward."""
    cache = model.__dict__.setdefault("_snvc_slab_z", {})
    key = (D, slab.world, slab.rank, model.zs.data_ptr(), model.zs._version)
    if key not in cache:
        cache[key] = slab_z_range(model.zs.cpu().numpy(), model.cv_range[4], model.cv_range[5], D, slab, model.align_corners)
    return cache[key]


def slab_global_forward(model, left_feat, right_feat, shift, proj, slab, group=None, out_dtype=torch.float32,
                        layout_out="NCDHW", comm=None, arena=None):
    """Depth-slab-parallel GlobalHotPath.forward for ONE pair (N == 1).

    Every rank passes the same (replicated) inputs and returns (voxels[:, zlo:zhi] slice, (zlo, zhi)):
    its slice of the lifted voxel grid along Z (layout as `GlobalHotPath.forward`).
    Halo exchange: `arena` (PeerArena: NVLink peer stores, the product path) > `comm` (HaloComm: NCCL point-to-point through
    the C ABI) > torch.distributed point-to-point ops (gloo in the CPU tests)."""
    from snvc_b200 import functional as SF
    from snvc_b200.extension.build_cost_volume import build_cost_volume_ndhwc_bf16, build_cost_volume_split_bf16
    if left_feat.shape[0] != 1:
        raise RuntimeError("slab_global_forward handles one pair (the stress configuration)")
    D = shift.shape[1]
    if D != slab.D:
        raise RuntimeError("shift / slab depth mismatch")
    bins = slab.ext_bins()
    keep = [b for b in bins if 0 <= b < D]
    lo_pad, hi_pad = keep[0] - bins[0], bins[-1] - keep[-1]
    sh = shift[:, keep[0]:keep[-1] + 1].contiguous()
    if arena is not None:
        arena.reset()                                                 # same carving on every rank and in every forward
    trunk = SlabTrunk(model, slab, group, comm, arena)
    F, H, W = left_feat.shape[1], left_feat.shape[2], left_feat.shape[3]
    split = hasattr(model, "split_supported") and model.split_supported(D)

    def ext_buffer(ch):
        """extended slab [1, Dl+2*HALO, H, W, ch]; planes outside the volume are the convolution's zero padding.  The
        volume kernels write the slab's real planes straight into it (batch 1: a depth range is a contiguous view)."""
        buf = torch.empty((1, len(bins), H, W, ch), dtype=torch.bfloat16, device=left_feat.device)
        if lo_pad:
            buf[:, :lo_pad].zero_()
        if hi_pad:
            buf[:, len(bins) - hi_pad:].zero_()
        return buf, buf[:, lo_pad:lo_pad + len(keep)]

    feat = None
    if split:
        # split cost volume (models/stereonet.py): right half per slab, the depth-constant left half as a 3-plane addend
        right_ext, real = ext_buffer(F)
        _, left_planes = build_cost_volume_split_bf16(left_feat, right_feat, sh, 1, out_right=real)
        try:
            feat = trunk(head=trunk.head_split(right_ext, model.trunk_head_addend(left_planes)))
        except RuntimeError as e:
            if "(-2)" not in str(e):          # SNVC_E_UNSUPPORTED: no cluster launch on this device -> unsplit volume
                raise
    if feat is None:
        cost_in = build_cost_volume_ndhwc_bf16(left_feat, right_feat, sh, 1)
        if lo_pad or hi_pad:
            cost, real = ext_buffer(2 * F)
            real.copy_(cost_in)
        else:
            cost = cost_in
        feat = trunk(cost)                                            # [1, Dl+2*HALO, H, W, ch]
    zlo, zhi = _cached_z_range(model, D, slab)
    vox = SF.frustum_lift(feat, proj, model.zs[zlo:zhi].contiguous(), model.ys, model.xs, model.cv_range,
                          model.align_corners, layout_in="NDHWC", out_dtype=out_dtype, layout_out=layout_out,
                          d_total=D, d_base=bins[0])
    return vox, (zlo, zhi)


class GraphedSlabForward:
    """CUDA-graph replay of `slab_global_forward` for fixed inputs shapes: the ~45 kernel launches, 15 allocations and 10
    halo exchanges (ncclSend / ncclRecv groups are capturable) of one slab forward become one `cudaGraphLaunch`.  With 8
    ranks a slab forward is ~2.5 ms of GPU work, less than the host needs to enqueue it eagerly from Python.
    Every rank must construct and replay it collectively (the captured graphs contain matching sends / receives)."""

    def __init__(self, model, left_feat, right_feat, shift, proj, slab, comm=None, out_dtype=torch.bfloat16,
                 layout_out="NDHWC", warmup=2, arena=None):
        self.inputs = tuple(t.clone() for t in (left_feat, right_feat, shift, proj))
        self.args = (model, slab, comm, out_dtype, layout_out, arena)
        dev = left_feat.device
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(warmup):                              # library / NCCL connection set-up happens outside the capture
                self._run()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph), torch.no_grad():
            self.vox, self.z_range = self._run()

    def _run(self):
        model, slab, comm, out_dtype, layout_out, arena = self.args
        l, r, sh, pr = self.inputs
        return slab_global_forward(model, l, r, sh, pr, slab, out_dtype=out_dtype, layout_out=layout_out, comm=comm,
                                   arena=arena)

    def load(self, left_feat, right_feat, shift, proj):
        for d, s_ in zip(self.inputs, (left_feat, right_feat, shift, proj)):
            d.copy_(s_, non_blocking=True)

    def replay(self):
        self.graph.replay()
        return self.vox, self.z_range

    def close(self):
        """Destroy the captured graph.  Must happen BEFORE the HaloComm is closed: NCCL keeps a communicator alive while a
        captured graph references it, and ncclCommDestroy blocks until that graph is gone (measured: a 60 s hang)."""
        if self.graph is not None:
            self.graph.reset()
            self.graph = None
        self.vox = None

