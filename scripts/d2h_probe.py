"""Device->host return probe (run on the GPU box): topology, dense cudaMemcpyAsync GB/s, and the masked zero-copy
return kernel at several grid sizes, alone and under a concurrent hot-path replay.  With torchrun: every rank prints."""
import os, sys, glob, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
rank = int(os.environ.get("LOCAL_RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
bind = os.environ.get("SNVC_NUMA_BIND", "1") != "0"
import torch
from snvc_b200.utils import numa
torch.cuda.set_device(rank)
info = numa.bind_to_gpu_node(rank) if bind else {"bound": False}
if rank == 0:
    print("cpus", os.cpu_count(), "nodes", sorted(glob.glob("/sys/devices/system/node/node[0-9]*")))
    try:
        print(subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True, timeout=30).stdout)
    except Exception as e:
        print("topo failed", e)
print(f"[rank {rank}] numa: {info}", flush=True)
import synth
from snvc_b200 import _lib
from snvc_b200.models.stereonet import GlobalHotPath, GraphedHotPath
from snvc_b200.utils.geometry import KITTI_P2, kitti_global_cfg, plane_sweep_shifts
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
dev = torch.device("cuda", rank)
cfg = kitti_global_cfg()
m = GlobalHotPath(cfg).eval(); m.load_state_dict(synth.det_state_dict(m, 41)); m = m.to(dev)
B = 8
ev = lambda: torch.cuda.Event(enable_timing=True)
def timeit(fn, n=5, pre=None):
    fn(); torch.cuda.synchronize()
    if world > 1: dist.barrier(); torch.cuda.synchronize()
    a, b = ev(), ev(); a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize(); return a.elapsed_time(b) / n
with torch.no_grad():
    g = GraphedHotPath(m, B, 32, (96, 312), 48, torch.bfloat16, "NDHWC")
    g_other = GraphedHotPath(m, B, 32, (96, 312), 48, torch.bfloat16, "NDHWC")   # the "next batch" (own memory pool, as in HostPipeline)
    gen = torch.Generator(device=dev).manual_seed(1)
    g(torch.randn((B, 32, 96, 312), device=dev, generator=gen), torch.randn((B, 32, 96, 312), device=dev, generator=gen),
      torch.from_numpy(plane_sweep_shifts(cfg, B)).to(dev), torch.from_numpy(KITTI_P2[None].repeat(B, 0).copy()).to(dev))
    torch.cuda.synchronize()
    vox, valid = g.vox, g.valid
    frac = valid.float().mean().item()
    ho = torch.empty(vox.shape, dtype=torch.bfloat16).pin_memory()
    mb = vox.numel() * 2 / 1e6
    t = timeit(lambda: ho.copy_(vox, non_blocking=True)); print(f"[rank {rank}] dense cudaMemcpyAsync D2H {mb:.0f} MB: {t:.2f} ms = {mb/t:.1f} GB/s", flush=True)
    L = _lib.lib()
    moved = torch.zeros((), dtype=torch.int64, device=dev)
    ones = torch.ones(valid.numel(), dtype=torch.uint8, device=dev)
    s2 = torch.cuda.Stream()
    for blocks in ((8, 16, 32, 64, 148) if world == 1 else (32,)):
        def run(v=valid):
            prev = v.reshape(-1).clone()
            L.snvc_masked_rows_to_host(vox.data_ptr(), v.data_ptr(), prev.data_ptr(), ho.data_ptr(), v.numel(), 64, blocks, moved.data_ptr(), _lib.stream_ptr())
        t = timeit(run)
        t1 = timeit(lambda: run(ones.view(valid.shape)))
        def with_compute():
            with torch.cuda.stream(s2):
                s2.wait_stream(torch.cuda.current_stream()); run()
            g_other.replay(); torch.cuda.current_stream().wait_stream(s2)
        t2 = timeit(with_compute)
        print(f"[rank {rank}] masked return, {blocks:3d} blocks: valid {frac:.3f} -> {mb*frac:.0f} MB in {t:.2f} ms = {mb*frac/t:.1f} GB/s | all rows {mb:.0f} MB in {t1:.2f} ms = {mb/t1:.1f} GB/s | with a concurrent step: {t2:.2f} ms", flush=True)
    tc = timeit(lambda: g_other.replay()); print(f"[rank {rank}] compute only: {tc:.2f} ms", flush=True)
    assert torch.equal(ho.view(torch.int16), vox.cpu().view(torch.int16))
    def dense_with_compute():
        with torch.cuda.stream(s2):
            s2.wait_stream(torch.cuda.current_stream()); ho.copy_(vox, non_blocking=True)
        g_other.replay(); torch.cuda.current_stream().wait_stream(s2)
    t3 = timeit(dense_with_compute); print(f"[rank {rank}] dense cudaMemcpyAsync with a concurrent step: {t3:.2f} ms", flush=True)
if world > 1: dist.destroy_process_group()
