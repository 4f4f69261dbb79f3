"""Instance branch hot path of snvc/models/vernier.py on the sm_100a kernels.

`VernierHotPath` holds exactly the reference's 3-D sub-modules under the reference's attribute
names (vimg_feat, conv1..conv4, hg_conv3d, fg_cls_head, pool_3d; vernier.py:250-289), so the
matching slice of a `VernierScale` state_dict loads with strict=True, and implements
  construct_voxel        vernier.py:351-360 (-> _sample_2d_feat :323-349)
  predict_3d_heatmaps    vernier.py:414-438, the 3-D part of vernier_type == 'BEV_type3'
`accelerate(model)` patches a reference `VernierScale` instance in place so that its own
`forward` (vernier.py:460-555) runs these stages on the B200 kernels; the 2-D ROI backbone
(hrnet.py) and the 2-D BEV tail (conv5 / hm1 / hm2 / coord_head, vernier.py:440-455) stay the
reference's torch modules.
"""
import types

import torch
import torch.nn as nn

from snvc_b200 import _lib
from snvc_b200 import functional as SF
from snvc_b200.models.submodule import _ConvNorm3d, _cbr, convbn_3d, hourglass, hourglass_downsample_16


class VernierHotPath(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        if cfg.vernier_type != "BEV_type3":
            raise NotImplementedError("snvc_b200 implements vernier_type='BEV_type3' (the shipped configuration)")
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
        self.pool_3d = nn.AvgPool3d((4, 1, 1), stride=(4, 1, 1))
        self.dim = dim

    # ---- A3 ------------------------------------------------------------------------------
    def construct_voxel(self, left, right, grid_proj_left, grid_proj_right, channels_last=True):
        """-> [N, nh, nw, nl, 2F] bf16 (channels_last, the kernel layout) or the reference's
        [N, 2F, nh, nw, nl] fp32."""
        nh, nw, nl = self.cfg.n_sample_h, self.cfg.n_sample_w, self.cfg.n_sample_l
        N = left.shape[0]
        if channels_last:
            v = SF.roi_voxel_sample(left, right, grid_proj_left, grid_proj_right, self.cfg.resolution,
                                    out_dtype=torch.bfloat16, layout="NDHWC")
            return v.view(N, nh, nw, nl, -1)
        v = SF.roi_voxel_sample(left, right, grid_proj_left, grid_proj_right, self.cfg.resolution)
        return v.view(N, -1, nh, nw, nl)

    # ---- A2e -----------------------------------------------------------------------------
    def _occupancy_conv(self):
        conv = self.fg_cls_head[2]
        vers = (conv.weight.data_ptr(), conv.weight._version)
        plan = getattr(self, "_occ_plan", None)
        if plan is None or plan[0] != vers:
            from snvc_b200.conv import PackedConv3d
            plan = (vers, PackedConv3d(conv.weight, None, stride=1, pad=1))
            object.__setattr__(self, "_occ_plan", plan)
        return plan[1]

    def predict_3d(self, voxel):
        """voxel [N,nh,nw,nl,2*dim] bf16 channels-last -> (voxel_BEV [N, dim*nh/4, nw, nl] fp32,
        occupancy [N, nh, nw, nl] fp32)   (vernier.py:414-438)."""
        dim = self.dim
        N, nh, nw, nl, _ = voxel.shape
        vimg = self.vimg_feat.fused(voxel)                                   # :415
        v = self.conv1.fused(voxel)                                          # :417
        v = self.conv2.fused(v, residual=v, residual_mode=2)                 # :418  relu(bn(conv)) + v
        v = self.conv3.fused(v, residual=v, residual_mode=2)                 # :419
        cat = torch.empty((N, nh, nw, nl, 2 * dim), dtype=torch.bfloat16, device=voxel.device)
        self.hg_conv3d.fused(v, out_residual=v, dst=cat)                     # :420-423, written into cat[..., :dim]
        h = self.fg_cls_head[0].fused(cat, relu=True, in_coffset=0)          # :427 (reads cat[..., :dim])
        occ = self._occupancy_conv()(h, sigmoid=True, out_dtype=torch.float32)   # [N,nh,nw,nl,1] fp32
        st = _lib.lib().snvc_scale_by_occupancy(vimg.data_ptr(), occ.data_ptr(), cat.data_ptr(), N * nh * nw * nl,
                                                dim, 2 * dim, dim, _lib.stream_ptr())   # :433
        _lib.check(st, "snvc_scale_by_occupancy")
        v = self.conv4.fused(cat)                                            # :435
        pool = 4
        bev = torch.empty((N, dim * (nh // pool), nw, nl), dtype=torch.float32, device=voxel.device)
        st = _lib.lib().snvc_avgpool_to_bev(v.data_ptr(), bev.data_ptr(), N, nh, nw, nl, dim, pool,
                                            _lib.stream_ptr())              # :436-438
        _lib.check(st, "snvc_avgpool_to_bev")
        return bev, occ.view(N, nh, nw, nl)

    def forward(self, left_feat, right_feat, grid_proj_left, grid_proj_right):
        with torch.cuda.device(left_feat.device):
            return self.predict_3d(self.construct_voxel(left_feat, right_feat, grid_proj_left, grid_proj_right))


_PREFIXES = ("vimg_feat", "conv1", "conv2", "conv3", "conv4", "hg_conv3d", "fg_cls_head")


def accelerate(model):
    """Patch a reference `VernierScale` (vernier_type 'BEV_type3') in place: its 3-D stages run on
    snvc_b200 with a snapshot of the model's 3-D parameters taken now (call again after loading
    new weights).  Returns the model."""
    hot = VernierHotPath(model.cfg)
    sd = {k: v for k, v in model.state_dict().items() if k.split(".")[0] in _PREFIXES}
    hot.load_state_dict(sd, strict=True)
    hot = hot.to(next(model.parameters()).device).eval()
    object.__setattr__(model, "_snvc_b200_hot", hot)

    def construct_voxel(self, left, right, grid_proj_left, grid_proj_right):
        return self._snvc_b200_hot.construct_voxel(left, right, grid_proj_left, grid_proj_right)

    def predict_3d_heatmaps(self, voxel, depth=None):
        if depth is not None:
            raise NotImplementedError
        voxel_BEV, occupancy = self._snvc_b200_hot.predict_3d(voxel)
        voxel_BEV = self.conv5(voxel_BEV)                                    # vernier.py:440
        if self.cfg.n_sample_w <= 16:
            heatmap_feats = self.hm1(voxel_BEV, None, None)[0].permute(0, 1, 3, 2)
        else:
            heatmap_feats = self.hm1(voxel_BEV).permute(0, 1, 3, 2)
        heatmaps = self.hm2(heatmap_feats)
        n = len(heatmaps)
        coor_maps = self.coor_maps.repeat(n, 1, 1, 1).to(heatmaps.device)
        coordinates = self.coord_head(torch.cat([heatmaps, coor_maps], dim=1)).view(n, -1, 2)
        bbox = self.bbox_head(coordinates.reshape(n, -1)) if hasattr(self, "bbox_head") else None
        return heatmaps, occupancy, None, coordinates, bbox

    model.construct_voxel = types.MethodType(construct_voxel, model)
    model.predict_3d_heatmaps = types.MethodType(predict_3d_heatmaps, model)
    return model
