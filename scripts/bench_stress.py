"""Stress configuration (BASELINE.json configs[4]): ONE volume with 2x depth bins at full
resolution (features 32 x 384 x 1248, D = 96: cost volume 64 x 96 x 384 x 1248 = 5.9 GB bf16, trunk 15.7 TFLOP),
split into depth slabs over the ranks with a conv3d halo exchange (NCCL P2P over NVLink) after every layer and a
z-partitioned lift.  Launch with torchrun:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 scripts/bench_stress.py [steps]

Rank 0 prints ONE JSON line (volumes/s, max-over-ranks CUDA-event time, trunk TFLOP/s aggregate)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import torch
import torch.distributed as dist
import synth
from snvc_b200 import parallel as par
from snvc_b200.models.stereonet import GlobalHotPath
from snvc_b200.utils.geometry import KITTI_P2, kitti_global_cfg, plane_sweep_shifts

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); os.environ.setdefault("MASTER_PORT", "29533")
os.environ["SNVC_B200_SKIP_SHIFT_CHECK"] = "1"
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
D, H, W = 96, 384, 1248
cfg = kitti_global_cfg(IH=H, IW=W, feat_stride=1, D=D)
m = GlobalHotPath(cfg).eval(); m.load_state_dict(synth.det_state_dict(m, 41), strict=True); m = m.to(dev)
g = torch.Generator(device=dev).manual_seed(7)                    # same seed on every rank: replicated inputs
lf = torch.randn((1, 32, H, W), device=dev, generator=g); rf = torch.randn((1, 32, H, W), device=dev, generator=g)
shift = torch.from_numpy(plane_sweep_shifts(cfg, 1)).to(dev)
proj = torch.from_numpy(KITTI_P2[None].copy()).to(dev)
slab = par.DepthSlab(D, world, rank)
comm = par.HaloComm(world, rank, dev) if world > 1 else None
ev = lambda: torch.cuda.Event(enable_timing=True)
with torch.no_grad():
    for _ in range(2):
        par.slab_global_forward(m, lf, rf, shift, proj, slab, out_dtype=torch.bfloat16, layout_out="NDHWC", comm=comm)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = ev(), ev()
    e0.record()
    for _ in range(steps):
        vox, (zlo, zhi) = par.slab_global_forward(m, lf, rf, shift, proj, slab, out_dtype=torch.bfloat16, layout_out="NDHWC", comm=comm)
    e1.record()
    torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1) / steps], device=dev, dtype=torch.float64)
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
if rank == 0:
    gflop = 491.90 * 2 * 16          # 2x depth bins, 16x the pixels of the 1/4-resolution volume
    print(json.dumps({"workload": "stress: 1 volume, D=96, 384x1248 features, depth slabs + conv3d halo exchange",
                      "n_gpus": world, "ms_per_volume": ms.item(), "volumes_per_s": 1e3 / ms.item(),
                      "trunk_gflop_per_volume": gflop, "aggregate_tflops_incl_cv_and_lift": gflop / ms.item(),
                      "slab_planes": slab.Dl, "z_slice_rank0": [zlo, zhi], "peak_mem_gb_rank0": torch.cuda.max_memory_allocated() / 1e9}))
if world > 1:
    dist.destroy_process_group()
