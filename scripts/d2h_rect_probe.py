"""Probe: return the in-frustum part of the lifted grid with the COPY ENGINE -- per (pair, z) the valid voxels form a
rectangle of (y, x) rows (pinhole projection), i.e. one cudaMemcpy2DAsync each -- against the zero-copy kernel
(snvc_masked_rows_to_host) and the dense copy.  python scripts/d2h_rect_probe.py"""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import torch, synth
from snvc_b200 import _lib
from snvc_b200.models.stereonet import GlobalHotPath, GraphedHotPath
from snvc_b200.utils.geometry import KITTI_P2, kitti_global_cfg, plane_sweep_shifts
dev = torch.device("cuda", 0)
cfg = kitti_global_cfg()
m = GlobalHotPath(cfg).eval(); m.load_state_dict(synth.det_state_dict(m, 41)); m = m.to(dev)
B = 8
rt = ctypes.CDLL("libcudart.so.12")
rt.cudaMemcpy2DAsync.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p]
ev = lambda: torch.cuda.Event(enable_timing=True)
def timeit(fn, n=5):
    fn(); torch.cuda.synchronize()
    a, b = ev(), ev(); a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize(); return a.elapsed_time(b) / n
with torch.no_grad():
    g = GraphedHotPath(m, B, 32, (96, 312), 48, torch.bfloat16, "NDHWC")
    g2 = GraphedHotPath(m, B, 32, (96, 312), 48, torch.bfloat16, "NDHWC")
    gen = torch.Generator(device=dev).manual_seed(1)
    g(torch.randn((B, 32, 96, 312), device=dev, generator=gen), torch.randn((B, 32, 96, 312), device=dev, generator=gen),
      torch.from_numpy(plane_sweep_shifts(cfg, B)).to(dev), torch.from_numpy(KITTI_P2[None].repeat(B, 0).copy()).to(dev))
    torch.cuda.synchronize()
    vox, valid = g.vox, g.valid                      # [B,Z,Y,X,32] bf16, [B,Z,Y,X] u8
    Bn, Z, Y, X, C = vox.shape
    v = valid.cpu().numpy().astype(bool)
    rects, nbytes, nvalid = [], 0, int(v.sum()) * 64
    for n in range(Bn):
        for z in range(Z):
            p = v[n, z]
            if not p.any(): continue
            ys = p.any(axis=1).nonzero()[0]; xs = p.any(axis=0).nonzero()[0]
            y0, y1, x0, x1 = ys[0], ys[-1] + 1, xs[0], xs[-1] + 1
            rects.append((n, z, y0, y1, x0, x1)); nbytes += (y1 - y0) * (x1 - x0) * 64
    print(f"{len(rects)} rectangles, {nbytes/1e6:.1f} MB (valid rows {nvalid/1e6:.1f} MB, dense {vox.numel()*2/1e6:.1f} MB)", flush=True)
    ho = torch.zeros(vox.shape, dtype=torch.bfloat16).pin_memory()
    pitch = X * 64
    def rect_copy():
        st = _lib.stream_ptr()
        for (n, z, y0, y1, x0, x1) in rects:
            off = int((((n * Z + z) * Y + y0) * X + x0) * 64)
            rt.cudaMemcpy2DAsync(ho.data_ptr() + off, pitch, vox.data_ptr() + off, pitch, int((x1 - x0) * 64), int(y1 - y0), 2, st)
    t = timeit(rect_copy); print(f"rectangle DMA: {t:.2f} ms = {nbytes/1e6/t:.1f} GB/s", flush=True)
    assert torch.equal(ho.view(torch.int16), vox.cpu().view(torch.int16)), "rectangles do not cover the valid rows"
    t = timeit(lambda: ho.copy_(vox, non_blocking=True)); print(f"dense DMA: {t:.2f} ms = {vox.numel()*2/1e6/t:.1f} GB/s", flush=True)
    L = _lib.lib(); moved = torch.zeros((), dtype=torch.int64, device=dev)
    def kern():
        prev = valid.reshape(-1).clone()
        L.snvc_masked_rows_to_host(vox.data_ptr(), valid.data_ptr(), prev.data_ptr(), ho.data_ptr(), valid.numel(), 64, 32, moved.data_ptr(), _lib.stream_ptr())
    t = timeit(kern); print(f"zero-copy kernel: {t:.2f} ms = {nvalid/1e6/t:.1f} GB/s", flush=True)
    s2 = torch.cuda.Stream()
    def overl(fn):
        def f():
            with torch.cuda.stream(s2):
                s2.wait_stream(torch.cuda.current_stream()); fn()
            g2.replay(); torch.cuda.current_stream().wait_stream(s2)
        return f
    print(f"with a concurrent step: rectangle DMA {timeit(overl(rect_copy)):.2f} ms | kernel {timeit(overl(kern)):.2f} ms | compute alone {timeit(lambda: g2.replay()):.2f} ms", flush=True)
    import time
    t0 = time.perf_counter(); rect_copy(); t1 = time.perf_counter(); torch.cuda.synchronize()
    print(f"host time to enqueue the rectangles: {(t1-t0)*1e3:.2f} ms", flush=True)
    # ---- hybrid: the far z-planes of every pair (almost fully inside the frustum) as ONE dense DMA per pair on a second
    # stream, the rest through the zero-copy kernel: do SM stores and the copy engine add up on PCIe?
    s3 = torch.cuda.Stream()
    for zs in (192, 176, 160, 144, 128, 112, 96):
        vk = valid.clone(); vk[:, zs:] = 0
        kb = int(vk.sum().item()) * 64; db = Bn * (Z - zs) * Y * X * 64
        def hybrid():
            cur = torch.cuda.current_stream()
            with torch.cuda.stream(s3):
                s3.wait_stream(cur)
                if zs < Z:
                    for n in range(Bn):
                        ho[n, zs:].copy_(vox[n, zs:], non_blocking=True)
            prev = vk.reshape(-1).clone()
            L.snvc_masked_rows_to_host(vox.data_ptr(), vk.data_ptr(), prev.data_ptr(), ho.data_ptr(), vk.numel(), 64, 32, moved.data_ptr(), _lib.stream_ptr())
            cur.wait_stream(s3)
        t = timeit(hybrid)
        t2 = timeit(overl(hybrid))
        print(f"hybrid z*={zs}: kernel {kb/1e6:.0f} MB + DMA {db/1e6:.0f} MB = {(kb+db)/1e6:.0f} MB in {t:.2f} ms = {(kb+db)/1e6/t:.1f} GB/s -> {8/t*1e3:.0f} pairs/s bound | with a concurrent step {t2:.2f} ms", flush=True)
    assert torch.equal(ho.view(torch.int16), vox.cpu().view(torch.int16))
