"""Time the depth-slab halo exchange alone: N ranks, one full-resolution 32-channel plane (30.7 MB) per direction.
torchrun --nproc-per-node N scripts/halo_bw_probe.py"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
from snvc_b200 import parallel as par
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
comm = par.HaloComm(world, rank, dev)
for C in (32, 64):
    x = torch.zeros((1, 16, 384, 1248, C) if C == 32 else (1, 10, 192, 624, C), dtype=torch.bfloat16, device=dev)
    for _ in range(3):
        comm.exchange(x)
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        comm.exchange(x)
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / 20], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    pb = x[0, 0].numel() * 2
    if rank == 0:
        print(f"world {world} plane {pb/1e6:.1f} MB: {t.item()*1e3:.1f} us per exchange = {pb/t.item()/1e6:.0f} GB/s per direction and neighbour "
              f"[NCCL_MIN_P2P_NCHANNELS={os.environ.get('NCCL_MIN_P2P_NCHANNELS')} NCCL_MAX_P2P_NCHANNELS={os.environ.get('NCCL_MAX_P2P_NCHANNELS')}]", flush=True)
comm.close(); dist.destroy_process_group()
