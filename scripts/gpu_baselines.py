#!/usr/bin/env python
"""Same-box GPU baselines: what the REFERENCE's own code path costs on this B200, stage by
stage, next to the snvc_b200 kernel for the same stage.  Baseline code only -- none of it is on the product path.

  cost volume   the reference's BuildCostVolumeForward kernel (BuildCostVolume_cuda.cu:63-98) compiled for sm_100
                by oracle/build_ref.py (fp32 NCDHW, the only form it has)
  trunk         nn.Conv3d / ConvTranspose3d / BatchNorm3d (oracle.blocks.GlobalTrunk = submodule.py:32-50,85-168 wiring)
                through cuDNN: fp32 NCDHW as the reference runs it (tools/inference_agnostic.py:18 cudnn.benchmark=True),
                and bf16 channels_last_3d, the strongest library configuration
  lift          ATen grid_sampler_3d (F.grid_sample, the call the restated global lift makes) fp32 and bf16
  ROI sampling  VernierScale._sample_2d_feat as written (vernier.py:323-349): normalise, 2 x F.grid_sample, cat
  instance CNN  oracle.blocks.Vernier3D (vernier.py:250-289,414-438) through cuDNN, bf16 channels_last_3d

    python scripts/gpu_baselines.py [--out gpurun_out/gpu_baselines.json] [--pairs 8] [--proposals 2]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))


def _time(fn, warm=2, iters=5):
    import torch
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def global_baselines(B=8, iters=5, with_fp32=True):
    """-> dict of ms per batch of B pairs for the library / reference implementations and ours, stage by stage."""
    import numpy as np
    import torch
    import torch.nn.functional as F
    import synth
    from oracle import blocks as oblocks, build_ref, global_branch as ogb
    from snvc_b200.extension import build_cost_volume as bcv
    from snvc_b200.models.stereonet import GlobalHotPath
    from snvc_b200.utils.geometry import KITTI_P2, kitti_global_cfg, plane_sweep_shifts
    torch.backends.cudnn.benchmark = True                       # tools/inference_agnostic.py:18
    dev = torch.device("cuda")
    cfg = kitti_global_cfg()
    C, H, W, D = 32, 96, 312, 48
    g = torch.Generator(device=dev).manual_seed(3)
    lf = torch.randn((B, C, H, W), device=dev, generator=g)
    rf = torch.randn((B, C, H, W), device=dev, generator=g)
    shift = torch.from_numpy(plane_sweep_shifts(cfg, B)).to(dev)
    proj = torch.from_numpy(KITTI_P2[None].repeat(B, 0).copy()).to(dev)
    os.environ["SNVC_B200_SKIP_SHIFT_CHECK"] = "1"
    out = {"pairs": B}
    with torch.no_grad():
        # ---- cost volume
        if build_ref.available():
            ref = build_ref.load("build_cost_volume_cuda")
            out["cost_volume_reference_kernel_f32_ms"] = _time(lambda: ref.build_cost_volume_forward(lf, rf, shift, 1), iters=iters)
        out["cost_volume_ours_f32_ncdhw_ms"] = _time(lambda: bcv.build_cost_volume(lf, rf, shift, 1), iters=iters)
        out["cost_volume_ours_bf16_ndhwc_ms"] = _time(lambda: bcv.build_cost_volume_ndhwc_bf16(lf, rf, shift, 1), iters=iters)
        out["cost_volume_ours_bf16_split_ms"] = _time(lambda: bcv.build_cost_volume_split_bf16(lf, rf, shift, 1), iters=iters)
        # ---- trunk through cuDNN
        model = GlobalHotPath(cfg).eval()
        sd = synth.det_state_dict(model, 41)
        model.load_state_dict(sd, strict=True)
        model = model.to(dev)
        trunk = oblocks.GlobalTrunk(2 * C, 32).eval()
        trunk.load_state_dict(sd, strict=True)
        cost_cl = bcv.build_cost_volume_ndhwc_bf16(lf, rf, shift, 1)                     # [B,D,H,W,64] bf16
        x_bf16 = cost_cl.permute(0, 4, 1, 2, 3)                                          # NCDHW view, channels_last_3d strides
        t_bf16 = trunk.to(dev).to(torch.bfloat16).to(memory_format=torch.channels_last_3d)
        out["trunk_cudnn_bf16_channels_last_ms"] = _time(lambda: t_bf16(x_bf16), iters=iters)
        feat_lib = t_bf16(x_bf16)
        ours_feat = model.trunk(cost_cl)
        out["trunk_ours_unsplit_ms"] = _time(lambda: model.trunk(cost_cl), iters=iters)
        rv, lp = bcv.build_cost_volume_split_bf16(lf, rf, shift, 1)
        out["trunk_ours_split_ms"] = _time(lambda: model.trunk_tail(model.trunk_head_split(rv, lp)), iters=iters)
        a, b = ours_feat.float(), feat_lib.permute(0, 2, 3, 4, 1).float()
        out["trunk_ours_vs_cudnn_bf16_relerr"] = float((a - b).abs().max() / b.abs().max())
        del a, b
        if with_fp32:
            t_f32 = oblocks.GlobalTrunk(2 * C, 32).eval()
            t_f32.load_state_dict(sd, strict=True)
            t_f32 = t_f32.to(dev)
            nb = max(1, B // 4)                                                           # 2 pairs: fp32 NCDHW activations are 4x larger
            x_f32 = x_bf16[:nb].float().contiguous()
            out["trunk_cudnn_fp32_ncdhw_ms"] = _time(lambda: t_f32(x_f32), warm=1, iters=2) * (B / nb)
            out["trunk_cudnn_fp32_ncdhw_note"] = f"timed on {nb} pairs, scaled to {B}; allow_tf32={torch.backends.cudnn.allow_tf32}"
            del x_f32, t_f32
        # ---- lift through ATen grid_sample
        geom = ogb.GlobalGeometry()
        zs, ys, xs = ogb.voxel_centres(geom)
        grid, valid = ogb.lift_grid(zs, ys, xs, KITTI_P2, geom.cv_ranges())
        grid = np.where(np.isfinite(grid), grid, np.float32(-2))
        gt = torch.from_numpy(grid[None]).to(dev).expand(B, -1, -1, -1, -1).contiguous()
        vm = torch.from_numpy(valid[None, None].astype(np.float32)).to(dev)
        vol_f32 = feat_lib.float().contiguous()                                          # [B,32,D,H,W] NCDHW fp32
        out["lift_aten_grid_sample_f32_ms"] = _time(lambda: F.grid_sample(vol_f32, gt, mode="bilinear", padding_mode="zeros",
                                                                         align_corners=True) * vm, iters=iters)
        vol_b = feat_lib.contiguous()
        gb = gt.to(torch.bfloat16)
        out["lift_aten_grid_sample_bf16_ms"] = _time(lambda: F.grid_sample(vol_b, gb, mode="bilinear", padding_mode="zeros",
                                                                          align_corners=True) * vm.to(torch.bfloat16), iters=iters)
        out["lift_ours_bf16_ms"] = _time(lambda: model.lift(ours_feat, proj, torch.bfloat16, "NDHWC"), iters=iters)
        out["lift_ours_f32_ncdhw_ms"] = _time(lambda: model.lift(ours_feat, proj, torch.float32, "NCDHW"), iters=iters)
    lib = out.get("cost_volume_reference_kernel_f32_ms", 0.0) + out["trunk_cudnn_bf16_channels_last_ms"] + out["lift_aten_grid_sample_bf16_ms"]
    ours = out["cost_volume_ours_bf16_split_ms"] + out["trunk_ours_split_ms"] + out["lift_ours_bf16_ms"]
    out["sum_library_best_ms"] = lib
    out["sum_ours_ms"] = ours
    out["pairs_per_s_library_best"] = B / lib * 1e3
    out["pairs_per_s_ours_stage_sum"] = B / ours * 1e3
    return out


def instance_baselines(NP=2, iters=3):
    import torch
    import torch.nn.functional as F
    import synth
    from oracle import blocks as oblocks
    from snvc_b200 import functional as SF
    torch.backends.cudnn.benchmark = True
    dev = torch.device("cuda")
    nh, nw, nl, Fc, Hf = 32, 128, 192, 32, 64
    P = nh * nw * nl
    g = torch.Generator(device=dev).manual_seed(5)
    lf = torch.randn((NP, Fc, Hf, Hf), device=dev, generator=g)
    rf = torch.randn((NP, Fc, Hf, Hf), device=dev, generator=g)
    gl = torch.rand((NP, 2, P), device=dev, generator=g) * 300 - 22
    gr = torch.rand((NP, 2, P), device=dev, generator=g) * 300 - 22
    res = (256, 256)

    def sample_ref():                                        # vernier.py:323-349, statement by statement
        outs = []
        for feat, pts in ((lf, gl), (rf, gr)):
            p = pts.permute(0, 2, 1).reshape(NP, nh, nw * nl, 2)
            p[:, :, :, 0] = p[:, :, :, 0] / res[1] * 2 - 1
            p[:, :, :, 1] = p[:, :, :, 1] / res[0] * 2 - 1
            outs.append(F.grid_sample(feat, p, align_corners=False).reshape(NP, Fc, nh, nw, nl))
        return torch.cat(outs, dim=1)

    out = {"proposals": NP}
    with torch.no_grad():
        out["roi_sample_aten_as_reference_f32_ms"] = _time(sample_ref, iters=iters)
        out["roi_sample_ours_bf16_ms"] = _time(lambda: SF.roi_voxel_sample(lf, rf, gl, gr, res, torch.bfloat16, "NDHWC"), iters=iters)
        out["roi_sample_ours_f32_ncdhw_ms"] = _time(lambda: SF.roi_voxel_sample(lf, rf, gl, gr, res), iters=iters)
        import types
        from snvc_b200.models.vernier import VernierHotPath
        ns = types.SimpleNamespace
        cfg = ns(vernier_type="BEV_type3", gn=False, hrfeat=ns(output_channel=32), num_parts=9, grid_resolution=[nh, nw, nl],
                 n_sample_h=nh, n_sample_w=nw, n_sample_l=nl, resolution=list(res))
        m = VernierHotPath(cfg).eval()
        sd = synth.det_state_dict(m, 31)
        m.load_state_dict(sd, strict=True)
        m = m.to(dev)
        vox_cl = m.construct_voxel(lf, rf, gl, gr)                                       # [NP,nh,nw,nl,64] bf16
        out["instance_cnn_ours_ms"] = _time(lambda: m.predict_3d(vox_cl), iters=iters)
        ref = oblocks.Vernier3D(32, n_sample_w=nw).eval()
        ref.load_state_dict(sd, strict=True)
        ref = ref.to(dev).to(torch.bfloat16).to(memory_format=torch.channels_last_3d)
        x = vox_cl.permute(0, 4, 1, 2, 3)
        out["instance_cnn_cudnn_bf16_channels_last_ms"] = _time(lambda: ref(x), warm=2, iters=iters)
    for k in list(out):
        if k.endswith("_ms"):
            out[k.replace("_ms", "_ms_per_proposal")] = out[k] / NP
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "gpu_baselines.json"))
    ap.add_argument("--pairs", type=int, default=8)
    ap.add_argument("--proposals", type=int, default=2)
    ap.add_argument("--skip-instance", action="store_true")
    a = ap.parse_args()
    import torch
    res = {"gpu": torch.cuda.get_device_name(0), "torch": torch.__version__, "cudnn": torch.backends.cudnn.version(),
           "global": global_baselines(a.pairs)}
    if not a.skip_instance:
        res["instance"] = instance_baselines(a.proposals)
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    with open(a.out, "w") as f:
        json.dump(res, f, indent=1)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
