"""Split cost-volume kernel: block-size sweep at the benchmark shape (GPU box)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from snvc_b200 import _lib
from snvc_b200.extension.build_cost_volume import build_cost_volume_split_bf16
from snvc_b200.utils.geometry import kitti_global_cfg, plane_sweep_shifts
dev = torch.device("cuda", 0)
cfg = kitti_global_cfg(); B = 8
g = torch.Generator(device=dev).manual_seed(1)
sets = [(torch.randn((B, 32, 96, 312), device=dev, generator=g), torch.randn((B, 32, 96, 312), device=dev, generator=g)) for _ in range(4)]
sh = torch.from_numpy(plane_sweep_shifts(cfg, B)).to(dev)
ev = lambda: torch.cuda.Event(enable_timing=True)
def t(parts, n=20):
    for i in range(3): build_cost_volume_split_bf16(*sets[i % 4], sh, 1, parts=parts)
    torch.cuda.synchronize(); a, b = ev(), ev(); a.record()
    for i in range(n): build_cost_volume_split_bf16(*sets[i % 4], sh, 1, parts=parts)
    b.record(); torch.cuda.synchronize(); return a.elapsed_time(b) / n
ref = None
for walk in (None, "general", None, "general"):
    _lib.set_option("SNVC_CV_WALK", walk)
    r = build_cost_volume_split_bf16(*sets[0], sh, 1, parts="right")
    if ref is None: ref = r.clone()
    print(f"walk {walk or 'two-region'}: right {t('right')*1e3:.1f} us, both {t('both')*1e3:.1f} us, identical {torch.equal(r, ref)}", flush=True)
_lib.set_option("SNVC_CV_WALK", None)
for th in (None, 384):
    _lib.set_option("SNVC_CV_THREADS", th)
    r = build_cost_volume_split_bf16(*sets[0], sh, 1, parts="right")
    if ref is None: ref = r.clone()
    ok = torch.equal(r, ref)
    print(f"threads {th}: right {t('right')*1e3:.1f} us, both {t('both')*1e3:.1f} us, identical {ok}", flush=True)
