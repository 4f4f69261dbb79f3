"""Where does the proposals-on-device path spend its time?  (GPU box)"""
import os, sys, types
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import torch, synth
from snvc_b200 import conv as C
from snvc_b200.models.stereonet import RPN3DHead, decode_proposals, ProposalDecoder
from snvc_b200.utils.geometry import kitti_global_cfg
dev = torch.device("cuda", 0)
cfg = kitti_global_cfg()
rcfg = types.SimpleNamespace(**vars(cfg), RPN_CONVDIM=32, num_angles=4, num_classes=1, box_corner_parameters=False)
rpn = RPN3DHead(rcfg, channels=32, n_y=20).eval(); rpn.load_state_dict(synth.det_state_dict(rpn, 71)); rpn = rpn.to(dev)
B = 8
vox = (torch.randn((B, 192, 20, 304, 32), device=dev) * 0.5).to(torch.bfloat16)
ev = lambda: torch.cuda.Event(enable_timing=True)
recs = []
o3, o2 = C.PackedConv3d.__call__, C.PackedConv2d.__call__
def t3(self, x, **kw):
    a, b = ev(), ev(); a.record(); y = o3(self, x, **kw); b.record()
    recs.append((f"3d {'deconv' if self.transposed else 'conv'} k{self.kernel} s{self.stride} {self.cin}->{self.cout} in{tuple(x.shape[1:4])}", a, b)); return y
def t2(self, x, **kw):
    a, b = ev(), ev(); a.record(); y = o2(self, x, **kw); b.record()
    recs.append((f"2d {'deconv' if self.transposed else 'conv'} k{self.kernel} s{self.stride} {self.cin}->{self.cout} in{tuple(x.shape[1:3])}", a, b)); return y
with torch.no_grad():
    for _ in range(2): out = rpn(vox); res = decode_proposals(*out, rcfg, pre_nms=256)
    torch.cuda.synchronize()
    a, b, c = ev(), ev(), ev()
    a.record(); out = rpn(vox); b.record(); res = decode_proposals(*out, rcfg, pre_nms=256); c.record(); torch.cuda.synchronize()
    print(f"rpn head {a.elapsed_time(b):.3f} ms, decode+nms {b.elapsed_time(c):.3f} ms (eager, batch {B})")
    C.PackedConv3d.__call__, C.PackedConv2d.__call__ = t3, t2
    rpn(vox); torch.cuda.synchronize()
    for n, x, y in recs: print(f"{x.elapsed_time(y)*1e3:9.1f} us  {n}")
    C.PackedConv3d.__call__, C.PackedConv2d.__call__ = o3, o2
    dec = ProposalDecoder(rcfg, dev, pre_nms=256)
    out = rpn(vox)
    for name, fn in (("rpn", lambda: rpn(vox)), ("decode+nms", lambda: dec(*out)), ("rpn+decode+nms", lambda: dec(*rpn(vox)))):
        g = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            fn()
            with torch.cuda.graph(g):
                res = fn()
        torch.cuda.synchronize()
        a.record()
        for _ in range(5): g.replay()
        b.record(); torch.cuda.synchronize()
        print(f"graph replay {name}: {a.elapsed_time(b)/5:.3f} ms")
    print("kept", res[3].tolist())
    import torch.profiler as tp
    with tp.profile(activities=[tp.ProfilerActivity.CUDA]) as prof:
        g.replay(); torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=14, max_name_column_width=60))
