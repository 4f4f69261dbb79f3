"""Probe: is snvc_halo_exchange (ncclSend / ncclRecv group on the library's own communicator) capturable in a CUDA graph?
torchrun --nproc-per-node 2 scripts/halo_graph_probe.py     (prints a line per stage; run under a short timeout)"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
from snvc_b200 import parallel as par
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
say = lambda *a: (print(f"[{rank}] {time.time():.1f}", *a, flush=True))
comm = par.HaloComm(world, rank, dev)
x = torch.full((1, 8, 64, 64, 32), float(rank + 1), dtype=torch.bfloat16, device=dev)
for _ in range(2):
    comm.exchange(x)
torch.cuda.synchronize(); say("eager ok", x[0, 1, 0, 0, 0].item(), x[0, -2, 0, 0, 0].item())
mode = sys.argv[1] if len(sys.argv) > 1 else "global"
g = torch.cuda.CUDAGraph()
say("capture begin", mode)
with torch.cuda.graph(g, capture_error_mode=mode):
    comm.exchange(x)
    y = x * 2
say("capture end")
dist.barrier(); say("barrier ok")
for i in range(3):
    g.replay()
    torch.cuda.synchronize(); say("replay", i, y[0, 1, 0, 0, 0].item())
g.reset(); comm.close()
dist.destroy_process_group(); say("done")
