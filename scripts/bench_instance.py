"""Instance branch (BASELINE.json configs[3]): high-res local ROI voxel sampling + refinement 3-D CNN
(vernier.py:323-360, 414-438; vernier_type 'BEV_type3', grid [32,128,192], ROI features 32 x 64 x 64 per view).
Per-layer CUDA-event timings and proposals/s for a batch of P proposals (default 8; 64 proposals = 1 frame).

    python scripts/bench_instance.py [proposals_per_batch] [steps]

Coordinates are synthetic coherent projections (a tilted plane sweep across the ROI, some points outside it) rather
than uniform noise, so the gather pattern is realistic.  Prints ONE JSON line."""
import json, os, sys, types
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import numpy as np
import torch
import synth
from snvc_b200 import conv as C
from snvc_b200.models.vernier import VernierHotPath

P = int(sys.argv[1]) if len(sys.argv) > 1 else 8
STEPS = int(sys.argv[2]) if len(sys.argv) > 2 else 5
ns = types.SimpleNamespace
grid = (32, 128, 192)
cfg = ns(vernier_type="BEV_type3", gn=False, hrfeat=ns(output_channel=32), num_parts=9, grid_resolution=list(grid),
         n_sample_h=grid[0], n_sample_w=grid[1], n_sample_l=grid[2], resolution=[256, 256])
dev = torch.device("cuda", 0)
m = VernierHotPath(cfg).eval(); m.load_state_dict(synth.det_state_dict(m, 31), strict=True); m = m.to(dev)
g = torch.Generator(device=dev).manual_seed(5)
lf = torch.randn((P, 32, 64, 64), device=dev, generator=g); rf = torch.randn((P, 32, 64, 64), device=dev, generator=g)
nh, nw, nl = grid
hh, ww, ll = torch.meshgrid(torch.linspace(0, 1, nh), torch.linspace(0, 1, nw), torch.linspace(0, 1, nl), indexing="ij")
def coords(shift):
    u = (-12.0 + 280.0 * ll + 25.0 * ww + shift).reshape(-1)         # pixel coords in the 256 x 256 ROI, partly outside
    v = (-8.0 + 270.0 * hh + 18.0 * ww).reshape(-1)
    return torch.stack([u, v])[None].repeat(P, 1, 1).contiguous().to(dev)
gl, gr = coords(0.0), coords(-9.0)
recs = []
orig = C.PackedConv3d.__call__
def timed(self, x, **kw):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); y = orig(self, x, **kw); e1.record()
    recs.append((f"{'deconv' if self.transposed else 'conv'} k{self.kernel} s{self.stride} d{self.dilation} {self.cin}->{self.cout} in{tuple(x.shape[1:4])}",
                 2.0 * self.kernel ** 3 * self.cin * self.cout * y.shape[0] * y.shape[1] * y.shape[2] * y.shape[3] / (8.0 if self.transposed else 1.0), e0, e1))
    return y
ev = lambda: torch.cuda.Event(enable_timing=True)
with torch.no_grad():
    for _ in range(2):
        m(lf, rf, gl, gr)
    torch.cuda.synchronize()
    t0, t1, t2 = ev(), ev(), ev()
    t0.record()
    for _ in range(STEPS):
        vox = m.construct_voxel(lf, rf, gl, gr)
    t1.record()
    for _ in range(STEPS):
        m.predict_3d(vox)
    t2.record()
    torch.cuda.synchronize()
    C.PackedConv3d.__call__ = timed
    m.predict_3d(vox)
    torch.cuda.synchronize()
layers = [(n, fl, a.elapsed_time(b)) for n, fl, a, b in recs]
for n, fl, ms in layers:
    print(f"{ms*1e3:10.1f} us  {fl/ms/1e9:8.1f} TFLOP/s  {n}", file=sys.stderr)
samp_ms, cnn_ms = t0.elapsed_time(t1) / STEPS, t1.elapsed_time(t2) / STEPS
gflop = sum(fl for _, fl, _ in layers) / 1e9 / P
line = {"workload": "instance branch hot path: ROI voxel sampling + BEV_type3 3-D CNN, grid 32x128x192, bf16", "proposals_per_batch": P,
        "proposals_per_s": P / ((samp_ms + cnn_ms) * 1e-3), "frames_per_s_64_proposals": P / ((samp_ms + cnn_ms) * 1e-3) / 64,
        "roi_sampling_ms_per_proposal": samp_ms / P, "roi_sampling_gbs": 114_294_784 * P / (samp_ms * 1e-3) / 1e9,
        "cnn3d_ms_per_proposal": cnn_ms / P, "cnn3d_gflop_per_proposal": gflop, "cnn3d_tflops": gflop * P / cnn_ms,
        "layers": [{"layer": n, "ms": ms, "tflops": fl / ms / 1e9} for n, fl, ms in layers]}
print(json.dumps(line))
