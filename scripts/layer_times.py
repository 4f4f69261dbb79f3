"""Per-layer CUDA-event timings of the global trunk in a normal (non-profiler) run."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import torch
import synth
from snvc_b200 import conv as C
from snvc_b200.models.stereonet import GlobalHotPath
from snvc_b200.extension.build_cost_volume import build_cost_volume_ndhwc_bf16
from snvc_b200.utils.geometry import KITTI_P2, kitti_global_cfg, plane_sweep_shifts

os.environ["SNVC_B200_SKIP_SHIFT_CHECK"] = "1"
dev = torch.device("cuda", 0)
cfg = kitti_global_cfg()
m = GlobalHotPath(cfg).eval(); m.load_state_dict(synth.det_state_dict(m, 41)); m = m.to(dev)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
g = torch.Generator(device=dev).manual_seed(1)
l = torch.randn((B, 32, 96, 312), device=dev, generator=g); r = torch.randn((B, 32, 96, 312), device=dev, generator=g)
shift = torch.from_numpy(plane_sweep_shifts(cfg, B)).to(dev)
proj = torch.from_numpy(KITTI_P2[None].repeat(B, 0).copy()).to(dev)
recs = []
orig = C.PackedConv3d.__call__
def timed(self, x, **kw):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); y = orig(self, x, **kw); e1.record()
    recs.append((f"{'deconv' if self.transposed else 'conv'} k{self.kernel} s{self.stride} {self.cin}->{self.cout} in{tuple(x.shape[1:4])}", e0, e1))
    return y
with torch.no_grad():
    for it in range(4):
        if it == 3:
            C.PackedConv3d.__call__ = timed
        cost = build_cost_volume_ndhwc_bf16(l, r, shift, 1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); feat = m.trunk(cost); e1.record()
        vox = m.lift(feat, proj, torch.bfloat16, "NDHWC")
    torch.cuda.synchronize()
tot = 0
for name, a, b in recs:
    t = a.elapsed_time(b) * 1e3; tot += t
    print(f"{t:9.1f} us  {name}")
print(f"sum {tot:.1f} us   trunk (events) {e0.elapsed_time(e1)*1e3:.1f} us")
