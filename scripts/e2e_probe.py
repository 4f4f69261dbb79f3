"""Where does the end-to-end time go?  Pure PCIe copies, sequential e2e, pipelined e2e (run on the GPU box)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import torch, synth
from snvc_b200.models.stereonet import GlobalHotPath, HostPipeline
from snvc_b200.utils.geometry import KITTI_P2, kitti_global_cfg, plane_sweep_shifts
os.environ["SNVC_B200_SKIP_SHIFT_CHECK"] = "1"
dev = torch.device("cuda", 0)
cfg = kitti_global_cfg()
m = GlobalHotPath(cfg).eval(); m.load_state_dict(synth.det_state_dict(m, 41)); m = m.to(dev)
B = 8
hl = torch.randn(B, 32, 96, 312).pin_memory(); hr = torch.randn(B, 32, 96, 312).pin_memory()
hs = torch.from_numpy(plane_sweep_shifts(cfg, B)).pin_memory(); hp = torch.from_numpy(KITTI_P2[None].repeat(B, 0).copy()).pin_memory()
Z, Y, X = m.zs.numel(), m.ys.numel(), m.xs.numel()
ho = [torch.empty((B, Z, Y, X, 32), dtype=torch.bfloat16).pin_memory() for _ in range(2)]
ev = lambda: torch.cuda.Event(enable_timing=True)
def timeit(fn, n=5):
    fn(); torch.cuda.synchronize()
    a, b = ev(), ev(); a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize(); return a.elapsed_time(b) / n
with torch.no_grad():
    d_out = torch.empty((B, Z, Y, X, 32), dtype=torch.bfloat16, device=dev)
    t = timeit(lambda: ho[0].copy_(d_out, non_blocking=True)); print(f"D2H 597.7 MB: {t:.2f} ms = {597.7/t:.1f} GB/s")
    dl = torch.empty_like(hl, device=dev)
    t = timeit(lambda: dl.copy_(hl, non_blocking=True)); print(f"H2D 30.7 MB: {t:.2f} ms = {30.67/t:.1f} GB/s")
    s2 = torch.cuda.Stream()
    def both():
        with torch.cuda.stream(s2):
            s2.wait_stream(torch.cuda.current_stream()); dl.copy_(hl, non_blocking=True)
        ho[0].copy_(d_out, non_blocking=True); torch.cuda.current_stream().wait_stream(s2)
    t = timeit(both); print(f"D2H + concurrent H2D: {t:.2f} ms")
    args = [x.to(dev) for x in (hl, hr, hs, hp)]
    t = timeit(lambda: m(*args, torch.bfloat16, "NDHWC")); print(f"compute only: {t:.2f} ms")
    def d2h_while_compute():
        with torch.cuda.stream(s2):
            s2.wait_stream(torch.cuda.current_stream()); ho[0].copy_(d_out, non_blocking=True)
        m(*args, torch.bfloat16, "NDHWC"); torch.cuda.current_stream().wait_stream(s2)
    t = timeit(d2h_while_compute); print(f"compute with a concurrent D2H: {t:.2f} ms")
    def seq():
        a = [x.to(dev, non_blocking=True) for x in (hl, hr, hs, hp)]
        ho[0].copy_(m(*a, torch.bfloat16, "NDHWC"), non_blocking=True)
    t = timeit(seq); print(f"sequential e2e: {t:.2f} ms/step = {B/t*1e3:.0f} pairs/s")
    for depth in (2, 3):
        pipe = HostPipeline(m, depth=depth)
        def run(n=8):
            for i in range(n): pipe.submit(hl, hr, hs, hp, ho[i % 2])
            torch.cuda.current_stream().wait_stream(pipe.s_out)
        t0 = time.perf_counter(); t = timeit(run, 2) / 8; host = (time.perf_counter() - t0) / 24
        print(f"pipelined depth {depth}: {t:.2f} ms/step = {B/t*1e3:.0f} pairs/s (host {host*1e3:.2f} ms/step wall)")
    print("mem GB", torch.cuda.max_memory_allocated() / 1e9, torch.cuda.memory_reserved() / 1e9)
