"""Pure compute of one depth slab of the stress volume on ONE GPU (halo exchange stubbed out): how much of the N-rank
stress time is kernels on a thin slab, how much is exchange / skew.   python scripts/slab_compute_probe.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import torch, synth
from snvc_b200 import parallel as par
from snvc_b200.models.stereonet import GlobalHotPath
from snvc_b200.utils.geometry import KITTI_P2, kitti_global_cfg, plane_sweep_shifts
dev = torch.device("cuda", 0)
D, H, W = 96, 384, 1248
cfg = kitti_global_cfg(IH=H, IW=W, feat_stride=1, D=D)
m = GlobalHotPath(cfg).eval(); m.load_state_dict(synth.det_state_dict(m, 41), strict=True); m = m.to(dev)
g = torch.Generator(device=dev).manual_seed(7)
lf = torch.randn((1, 32, H, W), device=dev, generator=g); rf = torch.randn((1, 32, H, W), device=dev, generator=g)
shift = torch.from_numpy(plane_sweep_shifts(cfg, 1)).to(dev); proj = torch.from_numpy(KITTI_P2[None].copy()).to(dev)
times = []
orig = par.exchange_depth_halo
def fake(x, slab, group=None, comm=None):
    e = torch.cuda.Event(enable_timing=True); e.record(); times.append(e); return x
par.exchange_depth_halo = fake
for world, rank in ((1, 0), (2, 0), (4, 1), (8, 3), (8, 0)):
    slab = par.DepthSlab(D, world, rank)
    with torch.no_grad():
        for _ in range(2):
            par.slab_global_forward(m, lf, rf, shift, proj, slab, out_dtype=torch.bfloat16, layout_out="NDHWC")
        torch.cuda.synchronize()
        times.clear()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        par.slab_global_forward(m, lf, rf, shift, proj, slab, out_dtype=torch.bfloat16, layout_out="NDHWC")
        e1.record(); torch.cuda.synchronize()
    marks = [e0] + times + [e1]
    seg = [marks[i].elapsed_time(marks[i + 1]) for i in range(len(marks) - 1)]
    print(f"world {world} rank {rank}: slab {slab.Dl} planes, total {e0.elapsed_time(e1):.3f} ms (ideal {14.8/world:.2f}); "
          "cv+addend+conv1 | dres0.2 | dres1.1 | dres1.2 | hg1 | hg2 | hg3 | hg4 | hg5 | hg6 | lift: " + " ".join(f"{s:.3f}" for s in seg), flush=True)
