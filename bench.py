#!/usr/bin/env python
"""Benchmark of the SNVC dense stereo-to-voxel hot path on B200 (see BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one pass of the global-branch hot path (plane-sweep cost volume -> 3-D trunk
dres0/dres1/hourglass -> frustum-to-voxel lift) over one batch of 8 synthetic KITTI-shaped stereo
pairs (BASELINE.json configs[1]; features [8,32,96,312] fp32, 48 depth bins, voxel grid
[192,20,304]) per GPU.  Pairs are independent, so ranks shard them with no data-path collective
("scaling": "weak", 8 pairs per rank per step).  Prints ONE JSON line on rank 0.

`--impl reference` times the reference's CPU path for the same stages (multi-threaded torch ops,
oracle/torch_path.py; the cost-volume op has no CPU implementation in the reference so that stage
is the torch restatement) on the host cores, one pair per step.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time
import types

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))

PAIRS_PER_GPU = 8
FEAT_C, FEAT_H, FEAT_W, DEPTH_BINS = 32, 96, 312, 48
TRUNK_GFLOP_PER_PAIR = 491.90          # BASELINE.md section 3 (sum 2*k^3*Cin*Cout*V_out)
CV_BYTES_PER_PAIR_BF16 = 191_692_992   # fp32 in -> bf16 NDHWC out (full 64-channel volume)
# split form (the product path): features in, right-half volume + three left planes out
CV_BYTES_PER_PAIR_SPLIT = 2 * FEAT_C * FEAT_H * FEAT_W * 4 + DEPTH_BINS * 4 + (DEPTH_BINS + 3) * FEAT_H * FEAT_W * FEAT_C * 2
CONV1_GFLOP_FULL = 2 * 27 * 64 * 32 * DEPTH_BINS * FEAT_H * FEAT_W * 1e-9      # dres0.conv1 on 64 channels, per pair (159.0)
CONV1_GFLOP_RIGHT = CONV1_GFLOP_FULL / 2                                        # its right-half launch (79.5)
ADDEND_GFLOP = 2 * 27 * 32 * 32 * 3 * FEAT_H * FEAT_W * 1e-9                   # the 3-plane addend convolution (5.0)
LIFT_VOX = 192 * 20 * 304


def lift_bytes_per_pair(out_bytes):
    # read the bf16 trunk output once (32*48*96*312*2) + write the lifted voxels
    return 32 * 48 * 96 * 312 * 2 + LIFT_VOX * 32 * out_bytes


def geometry():
    """CPU arm only: the oracle's own geometry object (independent of the product's)."""
    from oracle.global_branch import GlobalGeometry
    return GlobalGeometry()


def workload_config(world, eager=False):
    """`config` of the JSON line; both arms print the same one (the reference arm adds its sample)."""
    return {"workload": "global branch hot path, batch 8 synthetic KITTI pairs per GPU (feat 8x32x96x312 "
                        "fp32 -> cost volume 64x48x96x312 bf16 -> dres0/dres1/hourglass bf16 -> lift to "
                        "192x20x304 voxels), random init, BN eval folded",
            "pairs_per_gpu_per_step": PAIRS_PER_GPU, "parallelism": f"pair-sharded x{world}, no collective",
            "launch": "eager (one Python launch per kernel)" if eager else
                      "CUDA-graph replay (GraphedHotPath, one graph per stage; one graph set per resident input set, the graphs read "
                      "the inputs in place)",
            "cost_volume_form": "split (default): the depth-invariant left half of the 64-channel volume is kept as 3 planes "
                                "and enters dres0.conv1 as an addend; SNVC_SPLIT_CV=0 materialises the full volume",
            "l2": "inputs rotate over 4 sets (245 MB) and each step streams >3 GB of intermediates; "
                  "both exceed the 126 MB L2"}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sustained=d["bf16_tflops_sustained"],
                    source="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, source="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "20", "-i", str(self.index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except Exception:
                continue
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7),
                              ("sw_power_cap", 8)):
                if len(r) > col and r[col].lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------
def cpu_reference_pairs_per_s(steps, warmup, threads=None):
    """The reference's CPU path (torch ops on the host cores), one pair per step."""
    import torch
    from oracle import blocks, torch_path
    import synth
    threads = threads or len(os.sched_getaffinity(0)) or 1
    torch.set_num_threads(threads)
    geom = geometry()
    trunk = blocks.GlobalTrunk(2 * FEAT_C, 32).eval()
    trunk.load_state_dict(synth.det_state_dict(trunk, 41))
    path = torch_path.GlobalHotPathCPU(trunk, geom)
    g = torch.Generator().manual_seed(10)
    left = torch.randn((1, FEAT_C, FEAT_H, FEAT_W), generator=g)
    right = torch.randn((1, FEAT_C, FEAT_H, FEAT_W), generator=g)
    shift = torch.from_numpy(geom.shifts(1))
    Ps = torch.from_numpy(geom.P[None].copy())
    for _ in range(warmup):
        path(left, right, shift, Ps)
    t0 = time.perf_counter()
    for _ in range(steps):
        path(left, right, shift, Ps)
    dt = time.perf_counter() - t0
    return steps / dt, dt / steps, threads


def run_reference(args, rank, world):
    if rank != 0:
        return
    steps, warmup = max(1, min(args.steps, 3)), max(1, min(args.warmup, 1))
    v, spp, threads = cpu_reference_pairs_per_s(steps, warmup)
    sample = f"{steps} timed + {warmup} warm-up passes of 1 synthetic pair (fp32, torch CPU ops)"
    line = {"impl": "reference", "metric": "stereo pairs/s (cost volume + 3D trunk + voxel lift)", "value": v,
            "unit": "pairs/s", "n_gpus": args.gpus, "steps": steps, "warmup": warmup, "ms_per_step": spp * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": dict(workload_config(world), sample="each step = 1 pair of that workload on the host cores"),
            "cpu_baseline": {"value": v, "unit": "pairs/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------
def roofline_traffic():
    """DRAM bytes per launch per kernel, from the committed `ncu --set full` capture of this build
    (profiles/roofline_traffic.json, regenerated by scripts/ncu_traffic.py on the GPU box)."""
    p = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f)
    return {}


def gpu_baseline_leg(B):
    """Same-box library baselines for the three stages: the reference's own cost-volume kernel
    (oracle/_ref, if it was built), cuDNN bf16 channels_last_3d Conv3d trunk, ATen grid_sample lift.  Baseline code only."""
    sys.path.insert(0, os.path.join(ROOT, "scripts"))
    import gpu_baselines
    try:
        g = gpu_baselines.global_baselines(B, iters=3, with_fp32=False)
    except Exception as e:            # a baseline must never take the benchmark down
        return {"error": str(e)[:200]}
    keep = ("cost_volume_reference_kernel_f32_ms", "trunk_cudnn_bf16_channels_last_ms", "lift_aten_grid_sample_bf16_ms",
            "lift_aten_grid_sample_f32_ms", "sum_library_best_ms", "pairs_per_s_library_best", "trunk_ours_vs_cudnn_bf16_relerr")
    out = {k: g[k] for k in keep if k in g}
    out["what"] = ("per batch of %d pairs on this GPU: reference BuildCostVolumeForward kernel compiled for sm_100 (fp32 NCDHW), "
                   "oracle.blocks.GlobalTrunk through cuDNN (bf16, channels_last_3d, cudnn.benchmark), F.grid_sample (bf16)" % B)
    return out


def instance_leg(dev, rank, world, dist, peaks, frame_proposals=64, chunk=8, steps=2):
    """BASELINE.json configs[3]: 64 proposals per frame sharded over the ranks (contiguous blocks, no collective), each
    rank runs ROI voxel sampling + the BEV_type3 3-D CNN + the BEV tail on its block in chunks of <= 8 proposals."""
    import torch
    import synth
    from snvc_b200.models.vernier import VernierHotPath
    from snvc_b200.parallel import shard_range
    ns = types.SimpleNamespace
    grid = (32, 128, 192)
    cfg = ns(vernier_type="BEV_type3", gn=False, hrfeat=ns(output_channel=32), num_parts=9, grid_resolution=list(grid),
             n_sample_h=grid[0], n_sample_w=grid[1], n_sample_l=grid[2], resolution=[256, 256])
    m = VernierHotPath(cfg, bev_tail=True).eval()
    m.load_state_dict(synth.det_state_dict(m, 31), strict=True)
    m = m.to(dev)
    lo, hi = shard_range(frame_proposals, world, rank)
    mine = hi - lo
    P = min(chunk, max(mine, 1))
    g = torch.Generator(device=dev).manual_seed(500 + rank)
    lf = torch.randn((P, 32, 64, 64), device=dev, generator=g)
    rf = torch.randn((P, 32, 64, 64), device=dev, generator=g)
    nh, nw, nl = grid
    hh, ww, ll = torch.meshgrid(torch.linspace(0, 1, nh), torch.linspace(0, 1, nw), torch.linspace(0, 1, nl), indexing="ij")

    def coords(shift):       # coherent projections of the grid into the 256 x 256 ROI, partly outside it
        u = (-12.0 + 280.0 * ll + 25.0 * ww + shift).reshape(-1)
        v = (-8.0 + 270.0 * hh + 18.0 * ww).reshape(-1)
        return torch.stack([u, v])[None].repeat(P, 1, 1).contiguous().to(dev)
    gl, gr = coords(0.0), coords(-9.0)
    ev = lambda: torch.cuda.Event(enable_timing=True)
    nchunks = (mine + P - 1) // P if mine else 0
    with torch.no_grad():
        vox = m.construct_voxel(lf, rf, gl, gr)
        m.predict_heatmaps(vox)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0, t1, t2 = ev(), ev(), ev()
        samp = cnn = 0.0
        a, b = ev(), ev()
        a.record()
        for _ in range(steps):
            for _ in range(nchunks):
                t0.record()
                vox = m.construct_voxel(lf, rf, gl, gr)
                t1.record()
                m.predict_heatmaps(vox)
                t2.record()
        b.record()
        torch.cuda.synchronize()
        # per-stage split (events inside the loop above would need a sync per chunk): REP back-to-back sampling launches, then
        # REP CNN passes, so that neither number carries the other's allocator / clock transient
        REP = 4
        del vox
        t0.record()
        for _ in range(REP):
            vox = m.construct_voxel(lf, rf, gl, gr)
        t1.record()
        for _ in range(REP):
            m.predict_heatmaps(vox)
        t2.record()
        torch.cuda.synchronize()
        samp, cnn = t0.elapsed_time(t1) / (P * REP), t1.elapsed_time(t2) / (P * REP)
    t = torch.tensor([a.elapsed_time(b) / steps, samp, cnn], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    frame_ms, samp, cnn = t.tolist()
    gflop = 1708.36
    return {"workload": "configs[3]: 64 proposals/frame, ROI voxel sampling (grid 32x128x192, 64 ch) + BEV_type3 3-D CNN + BEV tail, bf16",
            "proposals_per_frame": frame_proposals, "parallelism": f"proposal-sharded x{world}, no collective",
            "ms_per_frame": frame_ms, "proposals_per_s": frame_proposals / (frame_ms * 1e-3), "frames_per_s": 1e3 / frame_ms,
            "roi_sampling": {"ms_per_proposal": samp, "achieved_gbs": 114_294_784 / (samp * 1e-3) / 1e9,
                             "frac_hbm": 114_294_784 / (samp * 1e-3) / 1e9 / peaks["hbm"], "bytes_per_proposal": 114_294_784},
            "cnn": {"ms_per_proposal": cnn, "achieved_tflops": gflop / cnn, "frac_tensor": gflop / cnn / peaks["tf_sustained"],
                    "gflop_per_proposal": gflop}}


def stress_leg(dev, rank, world, dist, steps=5):
    """BASELINE.json configs[4]: ONE volume, 2x depth bins (D = 96) at full resolution (features 32 x 384 x 1248: cost
    volume 5.9 GB bf16, trunk 15.7 TFLOP), split into depth slabs over the ranks with a conv3d halo exchange (NCCL P2P over
    NVLink) after every layer and a z-partitioned lift.  Strong scaling: the work is fixed, ranks divide it."""
    import torch
    import synth
    from snvc_b200 import parallel as par
    from snvc_b200.models.stereonet import GlobalHotPath
    from snvc_b200.utils.geometry import KITTI_P2, kitti_global_cfg, plane_sweep_shifts
    D, H, W = 96, 384, 1248
    if D % (4 * world) != 0:
        return {"skipped": f"D={D} is not a multiple of 4*world"}
    cfg = kitti_global_cfg(IH=H, IW=W, feat_stride=1, D=D)
    m = GlobalHotPath(cfg).eval()
    m.load_state_dict(synth.det_state_dict(m, 41), strict=True)
    m = m.to(dev)
    g = torch.Generator(device=dev).manual_seed(7)                    # same seed on every rank: replicated inputs
    lf = torch.randn((1, 32, H, W), device=dev, generator=g)
    rf = torch.randn((1, 32, H, W), device=dev, generator=g)
    shift = torch.from_numpy(plane_sweep_shifts(cfg, 1)).to(dev)
    proj = torch.from_numpy(KITTI_P2[None].copy()).to(dev)
    slab = par.DepthSlab(D, world, rank)
    # halo exchange: peer-memory push (snvc_halo_push over a CUDA-IPC-mapped arena; default) or NCCL point-to-point through
    # the C ABI (SNVC_STRESS_HALO=nccl, the A/B baseline: 146 GB/s per direction and neighbour on the 8-GPU box)
    use_peer = world > 1 and os.environ.get("SNVC_STRESS_HALO", "peer") != "nccl"
    arena = None
    if use_peer:
        try:                                                   # (PeerArena fails on ALL ranks alike or on none)
            arena = par.PeerArena(world, rank, dev, par.slab_arena_bytes(slab, H, W))
        except RuntimeError as e:
            sys.stderr.write(f"[rank {rank}] stress: {str(e)[:160]}; NCCL halo exchange instead\n")
            use_peer = False
    comm = par.HaloComm(world, rank, dev) if (world > 1 and not use_peer) else None
    ev = lambda: torch.cuda.Event(enable_timing=True)
    eager = lambda: par.slab_global_forward(m, lf, rf, shift, proj, slab, out_dtype=torch.bfloat16, layout_out="NDHWC", comm=comm,
                                            arena=arena)
    with torch.no_grad():
        for _ in range(2):
            eager()
        torch.cuda.synchronize()
        run, launch = eager, "eager (one Python -> C-ABI launch per kernel / exchange)"
        g = None
        if os.environ.get("SNVC_STRESS_GRAPH", "1") != "0":
            # one graph per rank holding the whole slab forward incl. the 10 NCCL halo exchanges; all ranks agree on
            # whether the capture worked before anyone replays (the graphs contain matching sends / receives)
            ok = torch.ones(1, device=dev)
            try:
                g = par.GraphedSlabForward(m, lf, rf, shift, proj, slab, comm=comm, warmup=0, arena=arena)
            except Exception as e:                                     # noqa: BLE001 -- fall back to eager launches
                ok.zero_()
                g = None
                sys.stderr.write(f"[rank {rank}] stress: graph capture failed ({str(e)[:120]}); eager launches\n")
            if world > 1:
                dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            if ok.item() > 0:
                run, launch = g.replay, "CUDA-graph replay (one graph per rank: kernels + halo exchanges)"
        for _ in range(2):
            run()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = ev(), ev()
        e0.record()
        for _ in range(steps):
            run()
        e1.record()
        torch.cuda.synchronize()
    if g is not None:
        g.close()                                           # before the communicator: ncclCommDestroy waits for graphs that captured it
    if comm is not None:
        comm.close()
    if arena is not None:
        arena.close()
    ms = torch.tensor([e0.elapsed_time(e1) / steps], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    gflop = 491.90 * 2 * 16
    halo_mb = 384 * 1248 * 2 / 1e6 * (32 * 5 + 64 * (2 / 4 + 3 / 16))        # one plane per direction per layer, summed over the 10 layers
    torch.cuda.empty_cache()
    return {"workload": "configs[4] stress: 1 volume, D=96, features 32x384x1248 (cost volume 5.9 GB bf16, trunk 15.7 TFLOP)",
            "parallelism": f"depth slabs x{world}" + ((", conv3d halo exchange after each of the 10 layers: " +
                                                               ("snvc_halo_push (one kernel: 16-byte peer stores of the boundary planes into the neighbours' CUDA-IPC-mapped slabs over NVLink + epoch barrier)"
                                                                if use_peer else "snvc_halo_exchange (one ncclGroup of send/recv with ranks r-1 / r+1)")) if world > 1 else ""),
            "scaling": "strong", "launch": launch, "cost_volume_form": "split" if m.split_supported(D) else "full",
            "ms_per_volume": ms.item(), "volumes_per_s": 1e3 / ms.item(),
            "aggregate_tflops": gflop / ms.item(), "slab_planes": slab.Dl,
            "halo_mb_per_rank_per_direction": halo_mb if world > 1 else 0.0}


def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    import synth
    from snvc_b200 import _lib
    from snvc_b200.models.stereonet import GlobalHotPath, GraphedHotPath, HostPipeline, ProposalDecoder, RPN3DHead
    from snvc_b200.extension.build_cost_volume import build_cost_volume_ndhwc_bf16
    from snvc_b200.utils import numa
    from snvc_b200.utils.geometry import KITTI_P2, kitti_global_cfg, plane_sweep_shifts

    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device (no CPU fallback in snvc_b200)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    # host side of this rank next to its GPU: cores + (first-touch) pinned buffers on the GPU's NUMA node
    all_cpus = os.sched_getaffinity(0)
    numa_info = numa.bind_to_gpu_node(local_rank) if os.environ.get("SNVC_NUMA_BIND", "1") != "0" else {"bound": False}
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    L = _lib.lib()
    cfg = kitti_global_cfg()
    model = GlobalHotPath(cfg).eval()
    model.load_state_dict(synth.det_state_dict(model, 41), strict=True)   # deterministic random init (tests/golden/synth.py)
    model = model.to(dev)
    B, K, W = PAIRS_PER_GPU, args.steps, args.warmup
    NSETS = 4   # inputs rotate over 4 sets (4 x 61 MB > 126 MB L2); intermediates (>3 GB / step) thrash L2 anyway
    gen = torch.Generator(device=dev).manual_seed(100 + rank)
    lefts = [torch.randn((B, FEAT_C, FEAT_H, FEAT_W), device=dev, generator=gen) for _ in range(NSETS)]
    rights = [torch.randn((B, FEAT_C, FEAT_H, FEAT_W), device=dev, generator=gen) for _ in range(NSETS)]
    shift = torch.from_numpy(plane_sweep_shifts(cfg, B)).to(dev)
    proj = torch.from_numpy(KITTI_P2[None].repeat(B, 0).copy()).to(dev)
    out_dtype, layout_out = torch.bfloat16, "NDHWC"
    os.environ["SNVC_B200_SKIP_SHIFT_CHECK"] = "1"

    ev = lambda: torch.cuda.Event(enable_timing=True)
    stage_ms = {"cost_volume": 0.0, "conv1": 0.0, "trunk": 0.0, "lift": 0.0}
    # The step is replayed from CUDA graphs (snvc_b200.models.stereonet.GraphedHotPath), one graph per stage -- cost volume |
    # 3-plane addend conv | dres0.conv1 | rest of the trunk | lift -- so that events between them time each stage inside
    # the timed region.
    # --eager launches the same kernels one by one from Python (host launch overhead then bounds the step).
    # One captured graph set PER INPUT SET: the graphs read the resident inputs in place (a producer -- the 2-D backbone -- would
    # write its features into a graph's static input buffers), so the timed region holds no input copy.
    graphs = None
    if not args.eager:
        graphs = [GraphedHotPath(model, B, FEAT_C, (FEAT_H, FEAT_W), DEPTH_BINS, out_dtype, layout_out, stages=True)
                  for _ in range(NSETS)]
        for g_, l_, r_ in zip(graphs, lefts, rights):
            g_.load(l_, r_, shift, proj)
        torch.cuda.synchronize()
        lefts = [g_.inputs[0] for g_ in graphs]               # (the staging copies are dropped: 4 x 61 MB stay resident)
        rights = [g_.inputs[1] for g_ in graphs]
    graphed = graphs[0] if graphs else None

    def step(i, timed):
        l, r = lefts[i % NSETS], rights[i % NSETS]
        e = [ev() for _ in range(6)] if timed else None
        if graphed is not None:
            g_ = graphs[i % NSETS]                          # the graph set captured on this input set
            if timed:
                e[0].record()
            # events: e0 start | e1 after the volume build | e5 after the addend conv (split form) | e2 after conv1 |
            # e3 after the rest of the trunk | e4 after the lift
            order = {"cost_volume": 1, "conv1_addend": 5, "conv1": 2, "trunk_rest": 3, "lift": 4}
            names = graphed.stage_names
            vox = g_.replay((lambda k: e[order[names[k]]].record()) if timed else None)
            return vox, e
        if timed:
            e[0].record()
        cost = build_cost_volume_ndhwc_bf16(l, r, shift, 1)
        if timed:
            e[1].record()
        feat = model.trunk(cost, mark=(lambda name: e[2].record()) if timed else None)
        if timed:
            e[3].record()
        vox = model.lift(feat, proj, out_dtype, layout_out)
        if timed:
            e[4].record()
        return vox, e

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    with torch.no_grad():
        for i in range(W):
            step(i, False)
        sync_all()
        sampler = ClockSampler(local_rank)
        if rank == 0:
            sampler.start()
        t_start, t_stop = ev(), ev()
        evs = []
        launches0 = L.snvc_launch_count()
        t_start.record()
        for i in range(K):
            _, e = step(W + i, True)
            evs.append(e)
        t_stop.record()
        launches = L.snvc_launch_count() - launches0
        if graphed is not None:                             # replayed kernels do not pass the library's counter
            launches = K * graphed.launches_per_replay
        sync_all()
        elapsed_ms = t_start.elapsed_time(t_stop)
        has_addend_stage = graphed is not None and "conv1_addend" in graphed.stage_names
        for e in evs:
            stage_ms["cost_volume"] += e[0].elapsed_time(e[1])
            stage_ms["conv1"] += (e[5] if has_addend_stage else e[1]).elapsed_time(e[2])
            stage_ms["trunk"] += e[1].elapsed_time(e[3])        # (split form: includes the 3-plane addend convolution)
            stage_ms["lift"] += e[3].elapsed_time(e[4])

        # ---- end to end through the public host-buffer API: pinned host inputs -> H2D -> hot path -> the lifted voxels
        # into pinned host memory, every step, transfers overlapped with compute on separate streams
        # (snvc_b200.models.stereonet.HostPipeline).  The return moves only the in-frustum voxel rows (the dense tensor
        # is what the host buffer holds afterwards; tests/test_gpu_host_return.py) -------------------------------------
        NH = 2
        h_in = [(torch.randn((B, FEAT_C, FEAT_H, FEAT_W)).pin_memory(), torch.randn((B, FEAT_C, FEAT_H, FEAT_W)).pin_memory(),
                 torch.from_numpy(plane_sweep_shifts(cfg, B)).pin_memory(),
                 torch.from_numpy(KITTI_P2[None].repeat(B, 0).copy()).pin_memory()) for _ in range(NH)]
        Z, Y, X = model.zs.numel(), model.ys.numel(), model.xs.numel()
        h_out = [torch.empty((B, Z, Y, X, 32), dtype=out_dtype).pin_memory() for _ in range(NH)]
        pipe = HostPipeline(model, depth=2, out_dtype=out_dtype, layout_out=layout_out, graphed=not args.eager,
                            sparse_return=not args.dense_return, return_blocks=args.return_blocks)
        e2e_steps = max(4, min(K, 20))
        for i in range(3):
            pipe.submit(*h_in[i % NH], h_out[i % NH])
        pipe.drain()
        sync_all()
        pipe.moved_bytes.zero_()
        e0, e1 = ev(), ev()
        cur = torch.cuda.current_stream()
        e0.record()
        for i in range(e2e_steps):
            pipe.submit(*h_in[i % NH], h_out[i % NH])
        cur.wait_stream(pipe.s_out)
        cur.wait_stream(pipe.s_in)
        e1.record()
        torch.cuda.synchronize()
        e2e_ms = e0.elapsed_time(e1)
        dense_bytes = int(B * LIFT_VOX * 32 * 2)
        d2h_bytes = int(pipe.moved_bytes.item()) // e2e_steps if pipe.sparse_return else dense_bytes

        # ---- second end-to-end record: the volume's real consumer is on the device.  hot path -> RPN3DHead (3-D convs on
        # the lifted grid, Y-pool -> BEV, 2-D hourglass, heads) -> decode + rotated BEV NMS, all on the GPU; only the
        # proposals (boxes, scores, keep lists) return to the host -----------------------------------------------------
        e2e_prop = None
        if not args.eager and not args.no_proposals:
            rcfg = types.SimpleNamespace(**vars(cfg), RPN_CONVDIM=32, num_angles=4, num_classes=1, box_corner_parameters=False)
            rpn = RPN3DHead(rcfg, channels=32, n_y=Y).eval()
            rpn.load_state_dict(synth.det_state_dict(rpn, 71), strict=True)
            rpn = rpn.to(dev)
            PRE = 256
            h_prop = [(torch.empty((B, PRE, 7)).pin_memory(), torch.empty((B, PRE)).pin_memory(),
                       torch.empty((B, PRE), dtype=torch.int64).pin_memory(), torch.empty((B,), dtype=torch.int32).pin_memory())
                      for _ in range(NH)]
            decoder = ProposalDecoder(rcfg, dev, pre_nms=PRE, iou_thresh=0.25)
            consumer = lambda v: decoder(*rpn(v))
            g2 = [GraphedHotPath(model, B, FEAT_C, (FEAT_H, FEAT_W), DEPTH_BINS, out_dtype, layout_out, consumer=consumer)
                  for _ in range(NH)]                              # hot path + RPN head + decode + NMS: ONE graph per slot
            s_in, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
            done = [None] * NH

            def submit_prop(i):
                k = i % NH
                with torch.cuda.stream(s_in):
                    if done[k] is not None:
                        s_in.wait_event(done[k][0])
                    for d, h in zip(g2[k].inputs, h_in[k]):
                        d.copy_(h, non_blocking=True)
                    ready = torch.cuda.Event()
                    ready.record(s_in)
                cur.wait_event(ready)
                if done[k] is not None:
                    cur.wait_event(done[k][1])                 # the slot's previous proposals must have left the device
                g2[k].replay()
                computed = torch.cuda.Event()
                computed.record(cur)
                with torch.cuda.stream(s_out):
                    s_out.wait_event(computed)
                    for h, d in zip(h_prop[k], g2[k].consumed):
                        h.copy_(d, non_blocking=True)
                    copied = torch.cuda.Event()
                    copied.record(s_out)
                done[k] = (computed, copied)

            for i in range(3):
                submit_prop(i)
            sync_all()
            p0, p1 = ev(), ev()
            p0.record()
            for i in range(e2e_steps):
                submit_prop(i)
            cur.wait_stream(s_out)
            cur.wait_stream(s_in)
            p1.record()
            torch.cuda.synchronize()
            e2e_prop = p0.elapsed_time(p1)
            del g2
        clocks = sampler.stop() if rank == 0 else None
        del pipe, h_out
        torch.cuda.empty_cache()

    t = torch.tensor([elapsed_ms, e2e_ms, stage_ms["cost_volume"], stage_ms["trunk"], stage_ms["lift"], stage_ms["conv1"],
                      e2e_prop if e2e_prop is not None else 0.0], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    elapsed_ms, e2e_ms, cv_ms, trunk_ms, lift_ms, conv1_ms, e2e_prop = t.tolist()

    peaks = measured_peaks()
    instance = None if args.no_instance else instance_leg(dev, rank, world, dist, peaks)
    stress = None if args.no_stress else stress_leg(dev, rank, world, dist)
    gpu_base = gpu_baseline_leg(B) if (rank == 0 and world == 1 and not args.no_gpu_baseline) else None
    cpu_v = cpu_threads = None
    if rank == 0 and not args.no_cpu_baseline:
        os.sched_setaffinity(0, all_cpus)                   # the CPU arm gets every host core, not just the GPU-local node
        cpu_v, _, cpu_threads = cpu_reference_pairs_per_s(2, 1)

    if rank == 0:
        pairs = world * B * K
        value = pairs / (elapsed_ms * 1e-3)
        split = graphed is not None and graphed.split
        # executed FLOPs: the split first layer convolves the depth-constant left half once (3 planes) instead of 48 times
        trunk_gflop = TRUNK_GFLOP_PER_PAIR - (CONV1_GFLOP_FULL - CONV1_GFLOP_RIGHT) + ADDEND_GFLOP if split else TRUNK_GFLOP_PER_PAIR
        trunk_tflops = trunk_gflop * 1e-3 * B * K / (trunk_ms * 1e-3)
        conv1_gflop = CONV1_GFLOP_RIGHT if split else CONV1_GFLOP_FULL
        conv1_tflops = conv1_gflop * 1e-3 * B * K / (conv1_ms * 1e-3)
        cv_bytes = CV_BYTES_PER_PAIR_SPLIT if split else CV_BYTES_PER_PAIR_BF16
        cv_gbs = cv_bytes * B * K / (cv_ms * 1e-3) / 1e9
        lift_gbs = lift_bytes_per_pair(2) * B * K / (lift_ms * 1e-3) / 1e9
        traffic = roofline_traffic()
        # per-stage kernels against the roofline that bounds each; the dominant one (largest share of the step among the
        # single-kernel stages) is the headline `roofline`, the others ride along in roofline.stages
        conv1_name = ("conv3d_kdpair_kernel<2,64,addend> (dres0.conv1 3x3x3 on the split cost volume: right half 32->32 + "
                      "depth-invariant addend, 1 launch / step)") if split else "conv3d_kdpair_kernel<4,128> (dres0.conv1 3x3x3 64->32, 1 launch / step)"
        stages = {
            "conv1": {"bound": "tensor", "kernel": conv1_name, "ms_per_step": conv1_ms / K, "achieved": conv1_tflops,
                      "peak": peaks["tf_sustained"], "unit": "TFLOP/s", "frac": conv1_tflops / peaks["tf_sustained"],
                      "frac_of_burst_peak": conv1_tflops / peaks["tf_burst"],
                      "traffic": traffic.get("dres0.conv1_split_dram_bytes_per_launch" if split else "dres0.conv1_dram_bytes_per_launch"),
                      "algorithmic_flop_per_launch": conv1_gflop * 1e9 * B, "share_of_step": conv1_ms / elapsed_ms},
            "cost_volume": {"bound": "hbm", "kernel": "cv_split_bf16_kernel (right-half volume) + cv_left_planes_kernel (2 launches / step)" if split
                            else "cv_ndhwc_bf16_kernel", "ms_per_step": cv_ms / K, "achieved": cv_gbs, "peak": peaks["hbm"],
                            "unit": "GB/s", "frac": cv_gbs / peaks["hbm"], "traffic": traffic.get("cost_volume_dram_bytes_per_step"),
                            "algorithmic_bytes_per_launch": cv_bytes * B, "share_of_step": cv_ms / elapsed_ms},
            "lift": {"bound": "hbm", "kernel": "lift_fast_bf16_kernel (1 launch / step)", "ms_per_step": lift_ms / K,
                     "achieved": lift_gbs, "peak": peaks["hbm"], "unit": "GB/s", "frac": lift_gbs / peaks["hbm"],
                     "traffic": traffic.get("lift_dram_bytes_per_launch"), "algorithmic_bytes_per_launch": lift_bytes_per_pair(2) * B,
                     "share_of_step": lift_ms / elapsed_ms},
            "trunk": {"bound": "tensor", "kernel": "all 15 conv launches of dres0 / dres1 / hourglass", "ms_per_step": trunk_ms / K,
                      "achieved": trunk_tflops, "peak": peaks["tf_sustained"], "unit": "TFLOP/s",
                      "frac": trunk_tflops / peaks["tf_sustained"], "share_of_step": trunk_ms / elapsed_ms,
                      "executed_gflop_per_pair": trunk_gflop, "reference_gflop_per_pair": TRUNK_GFLOP_PER_PAIR},
        }
        dom = max(("conv1", "cost_volume", "lift"), key=lambda k: stages[k]["ms_per_step"])
        roof = dict(stages[dom])
        roof["peak_source"] = peaks["source"] + (" (sustained bf16: kernel timed inside a long step)" if roof["bound"] == "tensor" else " (HBM copy bandwidth)")
        roof["stages"] = {k: v for k, v in stages.items() if k != dom}
        line = {
            "metric": "stereo pairs/s (cost volume + 3D trunk + voxel lift)", "value": value, "unit": "pairs/s",
            "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": elapsed_ms / K, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": workload_config(world, args.eager),
            "clocks": clocks,
            "e2e": {"value": world * B * e2e_steps / (e2e_ms * 1e-3), "unit": "pairs/s",
                    "h2d_bytes_per_step": int(2 * B * FEAT_C * FEAT_H * FEAT_W * 4 + B * DEPTH_BINS * 4 + B * 48),
                    "d2h_bytes_per_step": d2h_bytes, "dense_result_bytes_per_step": dense_bytes, "steps": e2e_steps,
                    "return": ("in-frustum voxel rows only, written by snvc_masked_rows_to_host straight into the pinned host "
                               "buffer at their dense positions (the buffer holds the dense tensor afterwards); "
                               "d2h_bytes_per_step is counted by the kernel") if pipe_sparse(args) else "dense cudaMemcpyAsync",
                    "numa": numa_info,
                    "proposals_on_device": ({"value": world * B * e2e_steps / (e2e_prop * 1e-3), "unit": "pairs/s",
                                             "d2h_bytes_per_step": int(B * 256 * (7 * 4 + 4 + 8) + B * 4),
                                             "what": "same host inputs and H2D traffic; the lifted volume feeds RPN3DHead + decode + rotated "
                                                     "BEV NMS on the device and only the proposals return (random-init heads)"}
                                            if e2e_prop else None),
                    "api": "snvc_b200.models.stereonet.HostPipeline.submit (pinned host buffers; H2D / compute / return on three "
                           "streams, 2 slots" + ("" if args.eager else ", one CUDA-graph replay per batch") + ")"},
            "gpu_launches": int(launches),
            "roofline": roof,
        }
        # configs[3] / configs[4] sub-records: top-level keys for readers of the raw line, and inside `roofline` (next to
        # the other per-stage fractions) because consumers that keep only the contract's keys drop unknown top-level ones
        if instance is not None:
            line["instance"] = instance
            roof["stages"]["instance_roi_sampling"] = {
                "bound": "hbm", "kernel": "roi_sample_fast_bf16_kernel (+ feature transposition), per proposal",
                "ms_per_proposal": instance["roi_sampling"]["ms_per_proposal"], "achieved": instance["roi_sampling"]["achieved_gbs"],
                "peak": peaks["hbm"], "unit": "GB/s", "frac": instance["roi_sampling"]["frac_hbm"],
                "traffic": (traffic["roi_dram_bytes_per_launch"] // 8) if "roi_dram_bytes_per_launch" in traffic else None,
                "algorithmic_bytes_per_launch": instance["roi_sampling"]["bytes_per_proposal"]}
            roof["stages"]["instance_cnn"] = {
                "bound": "tensor", "kernel": "instance 3-D CNN + BEV tail (conv3d_bigk / kdpair / conv2d kernels), per proposal",
                "ms_per_proposal": instance["cnn"]["ms_per_proposal"], "achieved": instance["cnn"]["achieved_tflops"],
                "peak": peaks["tf_sustained"], "unit": "TFLOP/s", "frac": instance["cnn"]["frac_tensor"],
                "proposals_per_s": instance["proposals_per_s"], "parallelism": instance["parallelism"]}
        if stress is not None:
            line["stress"] = stress
            if "ms_per_volume" in stress:
                roof["stages"]["stress_volume"] = {
                    "bound": "tensor", "kernel": "configs[4]: whole depth-slab forward of one D=96 full-resolution volume",
                    "ms_per_volume": stress["ms_per_volume"], "achieved": stress["aggregate_tflops"],
                    "peak": peaks["tf_sustained"] * world, "unit": "TFLOP/s",
                    "frac": stress["aggregate_tflops"] / (peaks["tf_sustained"] * world), "scaling": "strong",
                    "parallelism": stress["parallelism"]}
        if gpu_base is not None:
            line["gpu_baseline"] = gpu_base
        if cpu_v is not None:
            line["cpu_baseline"] = {"value": cpu_v, "unit": "pairs/s", "cores": cpu_threads, "kind": "port",
                                    "sample": "2 timed + 1 warm-up passes of 1 synthetic pair, same stages, fp32 torch "
                                              "CPU ops (oracle/torch_path.py), rank 0's host cores"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def pipe_sparse(args):
    return not args.dense_return


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--eager", action="store_true", help="launch kernel by kernel from Python instead of CUDA-graph replay")
    ap.add_argument("--dense-return", action="store_true", help="e2e: plain dense cudaMemcpyAsync of the voxels instead of the in-frustum-only return")
    ap.add_argument("--return-blocks", type=int, default=32, help="e2e: grid size of the host-return kernel")
    ap.add_argument("--no-instance", action="store_true", help="skip the configs[3] (instance branch) sub-record")
    ap.add_argument("--no-stress", action="store_true", help="skip the configs[4] (depth-slab stress volume) sub-record")
    ap.add_argument("--no-gpu-baseline", action="store_true", help="skip the same-box library baselines (N = 1 only)")
    ap.add_argument("--no-proposals", action="store_true", help="skip the proposals-on-device end-to-end record")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
