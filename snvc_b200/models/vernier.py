"""Instance branch hot path of snvc/models/vernier.py on the sm_100a kernels.

`VernierHotPath` holds the reference's 3-D sub-modules and the 2-D BEV tail under the reference's attribute names
(vimg_feat, conv1..conv4, hg_conv3d, fg_cls_head, [part_reg_head], pool_3d, conv5, hm1, hm2; vernier.py:250-314), so the
matching slice of a `VernierScale` state_dict loads with strict=True, and implements
  construct_voxel        vernier.py:351-360 (-> _sample_2d_feat :323-349)
  predict_3d_heatmaps    vernier.py:414-445: the 3-D part of vernier_type 'BEV_type3' / 'BEV_type2' on the tcgen05 3-D
                         kernels, AvgPool3d + reshape -> BEV, and conv5 / hm1 / hm2 on the 2-D tensor-core kernel
`accelerate(model)` converts a reference `VernierScale` instance in place: its 3-D / BEV sub-modules are replaced by the
drop-in modules of snvc_b200.models.submodule (same parameters, same state_dict keys) and its class by a subclass whose
`construct_voxel` / `predict_3d_heatmaps` run them fused.  Everything dispatches through `self`, so `nn.DataParallel`
replicas (tools/inference_agnostic.py:472) use their own device's parameters.  The 2-D ROI backbone (hrnet.py) and the
coordinate head (vernier.py:68-93, a few 18-channel 2-D blocks) stay the reference's torch modules.
"""
import torch
import torch.nn as nn

from snvc_b200 import _lib
from snvc_b200 import functional as SF
from snvc_b200.models.submodule import (_cbr, _cbr2d, convbn, convbn_3d, hourglass, hourglass2d, hourglass2d_downsample_16,
                                        hourglass_downsample_16)

SUPPORTED_TYPES = ("BEV_type3", "BEV_type2")      # identical 3-D / BEV stages (vernier.py:190-314); type3 adds the coord head


def _packed(owner, key, conv, ctor):
    """Per-module cache of a packed bare conv (no norm), invalidated by parameter version / device."""
    vers = (conv.weight.data_ptr(), conv.weight._version, str(conv.weight.device))
    cache = owner.__dict__.get("_snvc_bare", {})
    hit = cache.get(key)
    if hit is None or hit[0] != vers:
        hit = (vers, ctor(conv))
        cache = dict(cache)
        cache[key] = hit
        object.__setattr__(owner, "_snvc_bare", cache)
    return hit[1]


class _VernierOps:
    """The fused stages, written against `self.<reference attribute names>` so that both `VernierHotPath` and an
    accelerated reference `VernierScale` (and its DataParallel replicas) run them on their own parameters."""

    # ---- A3 ------------------------------------------------------------------------------
    def construct_voxel_cl(self, left, right, grid_proj_left, grid_proj_right):
        """-> [N, nh, nw, nl, 2F] bf16 channels-last (the kernel layout)."""
        cfg = self.cfg
        v = SF.roi_voxel_sample(left, right, grid_proj_left, grid_proj_right, cfg.resolution, out_dtype=torch.bfloat16,
                                layout="NDHWC")
        return v.view(left.shape[0], cfg.n_sample_h, cfg.n_sample_w, cfg.n_sample_l, -1)

    # ---- A2e -----------------------------------------------------------------------------
    def _core_3d(self, voxel):
        """voxel [N,nh,nw,nl,2*dim] bf16 -> (conv4 output [N,nh,nw,nl,dim] bf16, occupancy [N,nh,nw,nl] fp32,
        offset [N,27,nh,nw,nl] fp32 or None)   (vernier.py:415-435)."""
        from snvc_b200.conv import PackedConv3d
        dim = self.conv4[0][0].out_channels
        N, nh, nw, nl, _ = voxel.shape
        vimg = self.vimg_feat.fused(voxel)                                   # :415
        v = self.conv1.fused(voxel)                                          # :417
        v = self.conv2.fused(v, residual=v, residual_mode=2)                 # :418  relu(bn(conv)) + v
        v = self.conv3.fused(v, residual=v, residual_mode=2)                 # :419
        cat = torch.empty((N, nh, nw, nl, 2 * dim), dtype=torch.bfloat16, device=voxel.device)
        if isinstance(self.hg_conv3d, hourglass):
            self.hg_conv3d.fused(v, out_residual=v, dst=cat)                 # :420-421, written into cat[..., :dim]
        else:
            self.hg_conv3d.fused(v, out_residual=v, dst=cat)                 # :422-423
        h = self.fg_cls_head[0].fused(cat, relu=True, in_coffset=0)          # :427 (reads cat[..., :dim])
        occ_conv = _packed(self, "occ", self.fg_cls_head[2], lambda c: PackedConv3d(c.weight, None, stride=1, pad=1))
        occ = occ_conv(h, sigmoid=True, out_dtype=torch.float32)             # [N,nh,nw,nl,1] fp32
        offset = None
        if hasattr(self, "part_reg_head"):                                   # :428-431
            hp = self.part_reg_head[0].fused(cat, relu=True, in_coffset=0)
            reg_conv = _packed(self, "reg", self.part_reg_head[2], lambda c: PackedConv3d(c.weight, None, stride=1, pad=0))
            offset = reg_conv(hp, out_dtype=torch.float32).permute(0, 4, 1, 2, 3)
        with torch.cuda.device(voxel.device):
            st = _lib.lib().snvc_scale_by_occupancy(vimg.data_ptr(), occ.data_ptr(), cat.data_ptr(), N * nh * nw * nl,
                                                    dim, 2 * dim, dim, _lib.stream_ptr())   # :433
        _lib.check(st, "snvc_scale_by_occupancy")
        return self.conv4.fused(cat), occ.view(N, nh, nw, nl), offset        # :435

    def predict_3d(self, voxel):
        """-> (voxel_BEV [N, dim*nh/4, nw, nl] fp32 NCHW as the reference builds it, occupancy [N,nh,nw,nl] fp32)."""
        v, occ, _ = self._core_3d(voxel)
        N, nh, nw, nl, dim = v.shape
        pool = 4
        bev = torch.empty((N, dim * (nh // pool), nw, nl), dtype=torch.float32, device=voxel.device)
        with torch.cuda.device(voxel.device):
            st = _lib.lib().snvc_avgpool_to_bev(v.data_ptr(), bev.data_ptr(), N, nh, nw, nl, dim, pool, _lib.stream_ptr())   # :436-438
        _lib.check(st, "snvc_avgpool_to_bev")
        return bev, occ

    def predict_heatmaps(self, voxel):
        """-> (heatmaps [N, num_parts, nl, nw] fp32, occupancy, offset): vernier.py:414-445 end to end on the GPU kernels
        (pooled BEV channels-last bf16 -> conv5 -> hm1 -> hm2 on the 2-D tensor-core kernel)."""
        from snvc_b200.conv import PackedConv2d
        v, occ, offset = self._core_3d(voxel)
        N, nh, nw, nl, dim = v.shape
        pool = 4
        bev = torch.empty((N, nw, nl, dim * (nh // pool)), dtype=torch.bfloat16, device=voxel.device)
        with torch.cuda.device(voxel.device):
            st = _lib.lib().snvc_avgpool_to_bev_nhwc(v.data_ptr(), bev.data_ptr(), N, nh, nw, nl, dim, pool, 0, _lib.stream_ptr())
        _lib.check(st, "snvc_avgpool_to_bev_nhwc")
        x = self.conv5.fused(bev)                                            # :440
        if isinstance(self.hm1, hourglass2d):
            x = self.hm1.fused(x)[0]                                         # :442
        else:
            x = self.hm1.fused(x)                                            # :444
        # hm2 is applied to the (W, L)-transposed feature map (:442-445).  conv(x^T, K) = conv(x, K^T)^T, so the
        # transposition moves to the 3x3 kernel (swapped once at pack time) and to a view of the 9-channel result.
        hm2 = _packed(self, "hm2", self.hm2, lambda c: PackedConv2d(c.weight.transpose(2, 3), None, bias=c.bias,
                                                                     stride=1, pad=c.padding[0]))
        y = hm2(x, out_dtype=torch.float32)                                  # [N, nw, nl, parts]
        return y.permute(0, 3, 2, 1), occ, offset


class VernierHotPath(nn.Module, _VernierOps):
    def __init__(self, cfg, bev_tail=False):
        super().__init__()
        if cfg.vernier_type not in SUPPORTED_TYPES:
            raise NotImplementedError(f"snvc_b200 implements vernier_type in {SUPPORTED_TYPES} (BEV_type3 is the shipped one)")
        self.cfg = cfg
        dim, gn = cfg.hrfeat.output_channel, cfg.gn
        self.vimg_feat = _cbr(2 * dim, dim, 1, 1, 0, gn=gn)
        self.conv1 = _cbr(2 * dim, dim, 7, 1, 3, gn=gn)
        self.conv2 = _cbr(dim, dim, 5, 1, 2, gn=gn)
        self.conv3 = _cbr(dim, dim, 5, 1, 4, d=2, gn=gn)
        self.conv4 = _cbr(2 * dim, dim, 3, 1, 1, gn=gn)
        self.hg_conv3d = hourglass(dim, gn=gn) if cfg.n_sample_w <= 16 else hourglass_downsample_16(dim, gn=gn)
        self.fg_cls_head = nn.Sequential(convbn_3d(dim, dim, 3, 1, 1, gn=gn), nn.ReLU(inplace=True),
                                         nn.Conv3d(dim, 1, 3, 1, 1, bias=False), nn.Sigmoid())
        if getattr(cfg, "use_part_reg_head", False):                         # vernier.py:279-288
            self.part_reg_head = nn.Sequential(convbn_3d(dim, dim, 3, 1, 1, gn=gn), nn.ReLU(inplace=True),
                                               nn.Conv3d(dim, 27, 1, 1, 0, bias=False))
        self.pool_3d = nn.AvgPool3d((4, 1, 1), stride=(4, 1, 1))
        if bev_tail:                                                          # vernier.py:289-314
            dim_height = dim * cfg.grid_resolution[0] // 4
            self.conv5 = _ConvNormReLU2dSeq(dim_height, 64, gn)
            self.hm1 = hourglass2d(64, gn=gn) if cfg.n_sample_w <= 16 else hourglass2d_downsample_16(64, gn=gn)
            self.hm2 = nn.Conv2d(64, getattr(cfg, "num_parts", 9), 3, 1, 1, bias=False)
        self.dim = dim

    def construct_voxel(self, left, right, grid_proj_left, grid_proj_right, channels_last=True):
        """-> [N, nh, nw, nl, 2F] bf16 (channels_last, the kernel layout) or the reference's [N, 2F, nh, nw, nl] fp32."""
        if channels_last:
            return self.construct_voxel_cl(left, right, grid_proj_left, grid_proj_right)
        cfg = self.cfg
        v = SF.roi_voxel_sample(left, right, grid_proj_left, grid_proj_right, cfg.resolution)
        return v.view(left.shape[0], -1, cfg.n_sample_h, cfg.n_sample_w, cfg.n_sample_l)

    def forward(self, left_feat, right_feat, grid_proj_left, grid_proj_right):
        with torch.cuda.device(left_feat.device):
            return self.predict_3d(self.construct_voxel(left_feat, right_feat, grid_proj_left, grid_proj_right))


def _ConvNormReLU2dSeq(cin, cout, gn):
    """nn.Sequential(convbn(cin, cout, 3, 1, 1, 1), nn.ReLU(inplace=True)) -- conv5 (vernier.py:296-298)."""
    return _cbr2d(cin, cout, 1, gn)


_PREFIXES_3D = ("vimg_feat", "conv1", "conv2", "conv3", "conv4", "hg_conv3d", "fg_cls_head", "part_reg_head")
_PREFIXES_2D = ("conv5", "hm1", "hm2")


class _AcceleratedForward(_VernierOps):
    """Overrides of VernierScale.construct_voxel (vernier.py:351-360) and predict_3d_heatmaps (:362-458)."""

    def construct_voxel(self, left, right, grid_proj_left, grid_proj_right):
        with torch.cuda.device(left.device):
            return self.construct_voxel_cl(left, right, grid_proj_left, grid_proj_right)

    def predict_3d_heatmaps(self, voxel, depth=None):
        if depth is not None:
            raise NotImplementedError
        with torch.cuda.device(voxel.device):
            heatmaps, occupancy, offset = self.predict_heatmaps(voxel)
        coordinates, bbox = None, None
        if self.cfg.vernier_type == "BEV_type3":                             # vernier.py:446-455
            n = len(heatmaps)
            coor_maps = self.coor_maps.repeat(n, 1, 1, 1).to(heatmaps.device)
            coordinates = self.coord_head(torch.cat([heatmaps, coor_maps], dim=1)).view(n, -1, 2)
            bbox = self.bbox_head(coordinates.reshape(n, -1)) if hasattr(self, "bbox_head") else None
        return heatmaps, occupancy, offset, coordinates, bbox


def accelerate(model):
    """Convert a reference `VernierScale` (vernier_type 'BEV_type3' / 'BEV_type2') in place and return it.  The model
    keeps its state_dict keys and parameter values; loading new weights afterwards works as before (packed weights are
    re-derived when a parameter changes)."""
    cfg = model.cfg
    hot = VernierHotPath(cfg, bev_tail=True)
    want = set(hot.state_dict().keys())
    sd = {k: v for k, v in model.state_dict().items() if k in want}
    hot.load_state_dict(sd, strict=True)
    ref_p = next(model.parameters())
    hot = hot.to(ref_p.device).eval()
    for name in _PREFIXES_3D + _PREFIXES_2D:
        if hasattr(hot, name):
            setattr(model, name, getattr(hot, name))
    if not isinstance(model, _AcceleratedForward):
        model.__class__ = type("Accelerated" + type(model).__name__, (_AcceleratedForward, type(model)), {})
    return model
