"""Global branch hot path: plane-sweep cost volume -> 3-D trunk -> frustum-to-voxel lift.

The reference does not ship its global model class (snvc/models/__init__.py:1-2 are commented-out
imports of `StereoNet` / `NewStereoNet`); per SURVEY.md section 3.4 the composition below is
restated from the DSGN lineage (README.md:68) out of the blocks the reference does ship:
  build_cost_volume           snvc/extension/build_cost_volume/__init__.py:26
  convbn_3d / hourglass       snvc/models/submodule.py:32-50, 85-168
  projection / voxel centres  snvc/utils/torch_utils.py:36-45, 77-98
  range key names             snvc/models/loss3d.py:15-20 (CV_*_MIN/MAX, *_MIN/MAX, VOXEL_*_SIZE)
Wiring: dres0 = 2 x (convbn_3d + ReLU); dres1 = convbn_3d + ReLU + convbn_3d, out = dres1(x) + x;
out = hourglass(out, None, None)[0] + out; voxels = grid_sample(out, project(voxel centres)) * valid.
"""
import numpy as np
import torch
import torch.nn as nn

from snvc_b200 import functional as SF
from snvc_b200.extension.build_cost_volume import build_cost_volume_ndhwc_bf16, build_cost_volume_split_bf16
from snvc_b200.models.submodule import _cbr, convbn_3d, hourglass


def voxel_centres(lo, hi, step):
    """snvc/utils/torch_utils.py:85-94: arange(MIN, MAX - sign(step)*1e-10, step) + step/2 (float32)."""
    return torch.arange(lo, hi - np.sign(step) * 1e-10, step=step, dtype=torch.float32) + step / 2.0


class GlobalHotPath(nn.Module):
    """cfg: attribute-style object with the reference's key names (loss3d.py:15-20):
    X_MIN, X_MAX, Y_MIN, Y_MAX, Z_MIN, Z_MAX, VOXEL_{X,Y,Z}_SIZE, CV_{X,Y,Z}_{MIN,MAX}; optional
    `GN` (submodule.py:372) and `align_corners` (:375)."""

    def __init__(self, cfg, feat_channels=32, trunk_channels=32):
        super().__init__()
        gn = bool(getattr(cfg, "GN", False))
        cin, ch = 2 * feat_channels, trunk_channels
        self.dres0 = nn.Sequential(_cbr(cin, ch, 3, 1, 1, gn=gn), _cbr(ch, ch, 3, 1, 1, gn=gn))
        self.dres1 = nn.Sequential(_cbr(ch, ch, 3, 1, 1, gn=gn), convbn_3d(ch, ch, 3, 1, 1, gn=gn))
        self.hg = hourglass(ch, gn=gn)
        self.align_corners = bool(getattr(cfg, "align_corners", True))
        self.cv_range = tuple(float(getattr(cfg, k)) for k in
                              ("CV_X_MIN", "CV_X_MAX", "CV_Y_MIN", "CV_Y_MAX", "CV_Z_MIN", "CV_Z_MAX"))
        self.register_buffer("zs", voxel_centres(cfg.Z_MIN, cfg.Z_MAX, cfg.VOXEL_Z_SIZE), persistent=False)
        self.register_buffer("ys", voxel_centres(cfg.Y_MIN, cfg.Y_MAX, cfg.VOXEL_Y_SIZE), persistent=False)
        self.register_buffer("xs", voxel_centres(cfg.X_MIN, cfg.X_MAX, cfg.VOXEL_X_SIZE), persistent=False)

    # ---- stages on channels-last bf16 -----------------------------------------------------
    def trunk(self, cost, mark=None):
        """cost [N,D,H,W,2F] bf16 -> [N,D,H,W,ch] bf16 (one fused conv launch per layer; 64-channel layers
        run as two output slices).  `mark(name)`, if given, is called right after the first layer has been
        enqueued (bench.py records a CUDA event there to time the dominant kernel on its own)."""
        x = self.trunk_head(cost)
        if mark is not None:
            mark("dres0.conv1")
        return self.trunk_tail(x)

    def trunk_head(self, cost):
        """dres0.conv1 (3x3x3, 2F -> ch): the single largest kernel of the path."""
        return self.dres0[0].fused(cost)

    # ---- first layer on the SPLIT cost volume ------------------------------------------------------------------
    def _split_plans(self):
        """dres0.conv1 = convbn_3d(2F, ch) + ReLU split by input channel: (left part, no norm, fp32 out) and
        (right part + folded BatchNorm).  Cached like _ConvNorm3d._plan; None when the layer is not eligible."""
        from snvc_b200.conv import PackedConv3d
        cn = self.dres0[0][0]                            # _ConvNorm3d(conv, norm)
        conv, norm = cn[0], cn[1]
        F2 = conv.weight.shape[1]
        if isinstance(norm, nn.GroupNorm) or conv.weight.shape[0] != 32 or F2 != 64 or conv.kernel_size[0] != 3:
            return None
        if norm.training:
            raise RuntimeError("snvc_b200 conv blocks are inference-only: call .eval() (BatchNorm uses running stats)")
        vers = (conv.weight.data_ptr(), conv.weight._version, str(conv.weight.device), norm.weight._version,
                norm.bias._version, norm.running_mean._version, norm.running_var._version)
        plan = getattr(self, "_snvc_split_plan", None)
        if plan is None or plan[0] != vers:
            F = F2 // 2
            plan = (vers, PackedConv3d(conv.weight[:, :F].contiguous(), None, stride=1, pad=1),
                    PackedConv3d(conv.weight[:, F:].contiguous(), norm, stride=1, pad=1))
            object.__setattr__(self, "_snvc_split_plan", plan)
        return plan[1], plan[2]

    def split_supported(self, depth_bins):
        """The split first layer needs the CTA-pair conv kernel (default conv mode), >= 2 depth bins, 32 + 32 channels."""
        import os
        from snvc_b200 import _lib
        return depth_bins >= 2 and not _lib.get_option("SNVC_CONV_MODE") and os.environ.get("SNVC_SPLIT_CV", "1") != "0" \
            and self._split_plans() is not None

    def trunk_head_split(self, right_vol, left_planes):
        """dres0.conv1 on the split cost volume (build_cost_volume_split_bf16).  The left half of the volume does not
        vary with depth, so its share of the 3x3x3 convolution is the same for every interior output plane: it is
        computed ONCE as a 3-plane convolution of `left_planes` (zero padding in depth gives planes 0 / 1 / 2 the tap
        sets of output depth 0 / interior / D-1) and enters the right half's convolution as an fp32 addend before the
        folded BatchNorm and ReLU.  Same algebra as the 64-channel layer, half its FLOPs and half the volume bytes."""
        return self.trunk_head_right(right_vol, self.trunk_head_addend(left_planes))

    def trunk_head_addend(self, left_planes):
        """[N,3,H,W,F] bf16 -> fp32 [N,3,H,W,ch]: the left half's share of dres0.conv1 for output depth 0 / interior / D-1."""
        return self._split_plans()[0](left_planes, out_dtype=torch.float32)

    def trunk_head_right(self, right_vol, addend):
        return self._split_plans()[1](right_vol, relu=True, addend=addend)

    def trunk_tail(self, x):
        x = self.dres0[1].fused(x)
        x = self.dres1[1].fused(self.dres1[0].fused(x), residual=x, residual_mode=1)
        return self.hg.fused(x, out_residual=x)[0]

    def lift(self, vol, proj, out_dtype=torch.float32, layout_out="NCDHW", return_valid=False):
        return SF.frustum_lift(vol, proj, self.zs, self.ys, self.xs, self.cv_range, self.align_corners,
                               layout_in="NDHWC", out_dtype=out_dtype, layout_out=layout_out, return_valid=return_valid)

    def forward(self, left_feat, right_feat, shift, proj, out_dtype=torch.float32, layout_out="NCDHW", return_valid=False):
        """left_feat/right_feat [N,F,H,W] fp32, shift [N,D] fp32 (>= 0), proj [N,3,4] fp32
        -> lifted voxels [N,ch,Z,Y,X] (layout_out 'NCDHW') or [N,Z,Y,X,ch] ('NDHWC'); with `return_valid` also the
        in-frustum mask [N,Z,Y,X] uint8 (voxels with valid == 0 are exactly zero on every channel)."""
        if self.split_supported(shift.shape[1]):
            try:
                right_vol, left_planes = build_cost_volume_split_bf16(left_feat, right_feat, shift, 1)
                feat = self.trunk_tail(self.trunk_head_split(right_vol, left_planes))
            except RuntimeError as e:
                # the addend form exists on the CTA-pair kernel only; where cluster launches are unavailable (MIG slice,
                # < 2 SMs) it reports SNVC_E_UNSUPPORTED (-2) and the materialised 64-channel volume is used instead
                if "(-2)" not in str(e):
                    raise
                feat = self.trunk(build_cost_volume_ndhwc_bf16(left_feat, right_feat, shift, 1))
        else:
            feat = self.trunk(build_cost_volume_ndhwc_bf16(left_feat, right_feat, shift, 1))
        return self.lift(feat, proj, out_dtype, layout_out, return_valid)


class GraphedHotPath:
    """CUDA-graph replay of `GlobalHotPath` for one fixed batch shape.

    The eager path costs the host ~15 Python -> ctypes launches per batch (each with two tensor-map encodes and an
    output allocation); at 4-5 ms of GPU work per 8 pairs that host work is on the critical path whenever the CPU is
    busy or slow (bench.py measured 3.9 - 14 ms per step for the same 3.9 ms of kernels).  Here the launches of one
    batch are captured once -- the C ABI launches on the caller's stream, keeps no host state and never synchronises,
    so it is capturable as is -- and replayed with one `cudaGraphLaunch` per stage.

    `stages=True` captures one graph per stage (`stage_names`: cost volume | [3-plane addend convolution of the split
    first layer] | dres0.conv1 | rest of the trunk | lift) sharing one memory pool, so a caller can record events between them (bench.py); `stages=False` captures one graph.
    Inputs are copied into the graph's static buffers (`self.inputs`); the result is the static tensor `self.vox`
    (valid until the next replay).  `launches_per_replay` = kernels captured, from the library's launch counter."""

    def __init__(self, model, batch, feat_channels, feat_hw, depth_bins, out_dtype=torch.bfloat16, layout_out="NDHWC",
                 stages=False, consumer=None):
        """`consumer(vox)`: an on-device consumer of the lifted volume (e.g. RPN3DHead + decode_proposals) captured as one
        more stage behind the lift; its return value is kept in `self.consumed` (static tensors, valid until the next replay)."""
        from snvc_b200 import _lib
        self.model = model
        dev = next(model.parameters()).device
        self.dev = dev
        H, W = feat_hw
        self.inputs = (torch.zeros((batch, feat_channels, H, W), device=dev), torch.zeros((batch, feat_channels, H, W), device=dev),
                       torch.zeros((batch, depth_bins), device=dev), torch.zeros((batch, 3, 4), device=dev))
        self.inputs[3][:, 2, 2] = 1.0                       # a harmless projection for the warm-up pass
        l, r, sh, pr = self.inputs
        self.split = model.split_supported(depth_bins)
        tail = [lambda: setattr(self, "_feat", model.trunk_tail(self._x1)),
                lambda: self._set_lift(model.lift(self._feat, pr, out_dtype, layout_out, return_valid=True))]
        if self.split:
            # five stages: the 3-plane addend convolution is part of the first layer but gets its own graph, so that the
            # "conv1" stage is the single large launch a caller may want to time on its own.  (Forking left planes ->
            # addend convolution onto a second stream to hide its ~50 us under the right-half build was measured and
            # dropped: either kernel fills every SM's shared memory, so the two branches serialise anyway --
            # profiles/r02_notes.txt.)
            fns = [lambda: setattr(self, "_cost", build_cost_volume_split_bf16(l, r, sh, 1)),
                   lambda: setattr(self, "_addend", model.trunk_head_addend(self._cost[1])),
                   lambda: setattr(self, "_x1", model.trunk_head_right(self._cost[0], self._addend))] + tail
            self.stage_names = ["cost_volume", "conv1_addend", "conv1", "trunk_rest", "lift"]
        else:
            fns = [lambda: setattr(self, "_cost", build_cost_volume_ndhwc_bf16(l, r, sh, 1)),
                   lambda: setattr(self, "_x1", model.trunk_head(self._cost))] + tail
            self.stage_names = ["cost_volume", "conv1", "trunk_rest", "lift"]
        self.consumed = None
        if consumer is not None:
            fns.append(lambda: setattr(self, "consumed", consumer(self.vox)))
            self.stage_names.append("consumer")
        if not stages:
            parts = list(fns)
            fns = [lambda: [f() for f in parts]]
            self.stage_names = ["all"]
        L = _lib.lib()
        with torch.no_grad():
            side = torch.cuda.Stream(dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):                  # eager warm-up: weight packing, lazy module state
                for f in fns:
                    f()
            torch.cuda.current_stream(dev).wait_stream(side)
            torch.cuda.synchronize(dev)
            self.graphs, pool = [], None
            n0 = L.snvc_launch_count()
            for f in fns:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, pool=pool):
                    f()
                pool = g.pool()
                self.graphs.append(g)
            self.launches_per_replay = int(L.snvc_launch_count() - n0)

    def _set_lift(self, res):
        self.vox, self.valid = res                          # voxels + in-frustum mask [N,Z,Y,X] uint8 (static tensors)

    def load(self, left_feat, right_feat, shift, proj):
        for d, s in zip(self.inputs, (left_feat, right_feat, shift, proj)):
            d.copy_(s, non_blocking=True)

    def replay(self, between=None):
        """Replays the captured stages in order on the current stream; `between(i)` is called after stage i."""
        for i, g in enumerate(self.graphs):
            g.replay()
            if between is not None:
                between(i)
        return self.vox

    def __call__(self, left_feat, right_feat, shift, proj):
        self.load(left_feat, right_feat, shift, proj)
        return self.replay()


class DepthHead(nn.Module):
    """Depth head behind the trunk (restated wiring; blocks: convbn_3d submodule.py:32-50,
    disparityregression :76-83): classif = convbn_3d(ch, ch, 3, 1, 1) + ReLU + Conv3d(ch, 1, 3, 1, 1, bias=False);
    depth = disparityregression(softmax(F.interpolate(classif(x), [maxdisp, H, W], 'trilinear')), depth_values).
    The two convs run on the tcgen05 kernels, the rest as ONE fused kernel (snvc_depth_regression_fwd)."""

    def __init__(self, cfg, channels=32, maxdisp=192):
        super().__init__()
        self.classif = nn.Sequential(convbn_3d(channels, channels, 3, 1, 1, gn=bool(getattr(cfg, "GN", False))),
                                     nn.ReLU(inplace=True), nn.Conv3d(channels, 1, 3, 1, 1, bias=False))
        self.align_corners = bool(getattr(cfg, "align_corners", True))
        self.maxdisp = maxdisp

    def _logit_conv(self):
        from snvc_b200.conv import PackedConv3d
        conv = self.classif[2]
        vers = (conv.weight.data_ptr(), conv.weight._version)
        plan = getattr(self, "_plan", None)
        if plan is None or plan[0] != vers:
            plan = (vers, PackedConv3d(conv.weight, None, stride=1, pad=1))
            object.__setattr__(self, "_plan", plan)
        return plan[1]

    def forward(self, vol, depth_values, out_hw):
        """vol [N,D,H,W,ch] bf16 channels-last (the trunk output), depth_values [maxdisp] fp32 -> depth [N,Hout,Wout]."""
        h = self.classif[0].fused(vol, relu=True)
        logits = self._logit_conv()(h, out_dtype=torch.float32)              # [N,D,H,W,1] fp32
        return SF.depth_regression_from_logits(logits[..., 0], depth_values, (self.maxdisp, out_hw[0], out_hw[1]),
                                               self.align_corners)


class RPN3DHead(nn.Module):
    """RPN side of the global branch, downstream of the lift (restated wiring of the DSGN lineage,
    README.md:68, out of blocks the reference ships):
        rpn3d_conv   = convbn_3d(C, C, 3, 1, 1) + ReLU          submodule.py:32-50          on the lifted grid [N,Z,Y,X,C]
        rpn3d_conv2  = hourglass(C)(x, None, None)[0] + x       submodule.py:85-168
        rpn3d_pool   = AvgPool3d((4,1,1)) over Y, reshape to BEV [N, C*Y/4, Z, X]   (the op of vernier.py:289,436-438)
        rpn3d_conv3  = convbn(C*Y/4, B, 3, 1, 1, 1) + ReLU      submodule.py:11-29
        rpn3d_conv4  = hourglass2d(B)(x, None, None)[0] + x     submodule.py:317-361
        cls / reg towers = convbn(B, B) + ReLU; heads bbox_cls [A*K], bbox_reg [A*R], bbox_centerness [A*K] = Conv2d 3x3
    with A = cfg.num_angles, K = cfg.num_classes, R = 24 | 7 (cfg.box_corner_parameters), the output conventions RPN3DLoss
    consumes (loss3d.py:84-103,253-275); BEV cell (z, x) <-> compute_locations_bev (torch_utils.py:77-98).
    Everything runs on the tcgen05 kernels: 3-D convs, the Y-pool into channels-last BEV, 2-D convs with the K loop over
    the 160 BEV channels."""

    def __init__(self, cfg, channels=32, n_y=20, pool=4):
        super().__init__()
        from snvc_b200.models.submodule import _cbr2d, hourglass2d
        gn = bool(getattr(cfg, "GN", False))
        B = 2 * int(getattr(cfg, "RPN_CONVDIM", 32))
        self.pool = pool
        self.num_angles, self.num_classes = int(getattr(cfg, "num_angles", 4)), int(getattr(cfg, "num_classes", 1))
        self.reg_dim = 24 if getattr(cfg, "box_corner_parameters", False) else 7
        A = self.num_angles * self.num_classes
        self.rpn3d_conv = _cbr(channels, channels, 3, 1, 1, gn=gn)
        self.rpn3d_conv2 = hourglass(channels, gn=gn)
        self.rpn3d_conv3 = _cbr2d(channels * (n_y // pool), B, 1, gn)
        self.rpn3d_conv4 = hourglass2d(B, gn=gn)
        self.rpn3d_cls_convs = _cbr2d(B, B, 1, gn)
        self.rpn3d_bbox_convs = _cbr2d(B, B, 1, gn)
        self.bbox_cls = nn.Conv2d(B, A, 3, 1, 1)
        self.bbox_reg = nn.Conv2d(B, A * self.reg_dim, 3, 1, 1)
        self.bbox_centerness = nn.Conv2d(B, A, 3, 1, 1)

    def _head(self, conv):
        from snvc_b200.conv import PackedConv2d
        vers = (conv.weight.data_ptr(), conv.weight._version, conv.bias._version, str(conv.weight.device))
        cache = self.__dict__.get("_heads", {})
        hit = cache.get(id(conv))
        if hit is None or hit[0] != vers:
            hit = (vers, PackedConv2d(conv.weight, None, bias=conv.bias, stride=1, pad=1))
            cache = dict(cache)
            cache[id(conv)] = hit
            object.__setattr__(self, "_heads", cache)
        return hit[1]

    def bev_features(self, vox):
        """vox [N,Z,Y,X,C] bf16 channels-last (the lifted grid) -> BEV features [N,Z,X,B] bf16 channels-last."""
        from snvc_b200 import _lib
        x = self.rpn3d_conv.fused(vox)
        x = self.rpn3d_conv2.fused(x, out_residual=x)[0]
        N, Z, Y, X, C = x.shape
        bev = torch.empty((N, Z, X, C * (Y // self.pool)), dtype=torch.bfloat16, device=x.device)
        with torch.cuda.device(x.device):
            st = _lib.lib().snvc_avgpool_to_bev_nhwc(x.data_ptr(), bev.data_ptr(), N, Z, Y, X, C, self.pool, 1, _lib.stream_ptr())
        _lib.check(st, "snvc_avgpool_to_bev_nhwc")
        b = self.rpn3d_conv3.fused(bev)
        return self.rpn3d_conv4.fused(b, out_residual=b)[0]

    def forward(self, vox):
        """-> (bbox_cls [N,A*K,Z,X], bbox_reg [N,A*R,Z,X], bbox_centerness [N,A*K,Z,X]) fp32, NCHW views of channels-last
        results (what RPN3DLoss / the proposal decoder index)."""
        b = self.bev_features(vox)
        c, r = self.rpn3d_cls_convs.fused(b), self.rpn3d_bbox_convs.fused(b)
        cls = self._head(self.bbox_cls)(c, out_dtype=torch.float32)
        reg = self._head(self.bbox_reg)(r, out_dtype=torch.float32)
        ctr = self._head(self.bbox_centerness)(r, out_dtype=torch.float32)
        return cls.permute(0, 3, 1, 2), reg.permute(0, 3, 1, 2), ctr.permute(0, 3, 1, 2)


class ProposalDecoder:
    """BEV head outputs -> rotated-NMS'd proposals per pair, entirely on the device (no host synchronisation, CUDA-graph
    capturable): scores sigmoid(cls) * sigmoid(centerness) per (cell, angle anchor), the `pre_nms` best are decoded as
    [x + dx, y_a + dy, z + dz, l * e^dl, w * e^dw, h * e^dh, angle_a + dtheta] around compute_locations_bev
    (torch_utils.py:77-98; 7-parameter regression, loss3d.py:101) and passed to the rotated BEV NMS (snvc_nms_bev, N4).
    Restated decoder: the reference ships the heads' loss, not a decoder."""

    def __init__(self, cfg, device, anchor_size=(1.56, 1.6, 3.9), anchor_y=1.0, pre_nms=512, iou_thresh=0.25):
        self.zs = voxel_centres(cfg.Z_MIN, cfg.Z_MAX, cfg.VOXEL_Z_SIZE).to(device)
        self.xs = voxel_centres(cfg.X_MIN, cfg.X_MAX, cfg.VOXEL_X_SIZE).to(device)
        self.anchor_size, self.anchor_y, self.pre_nms, self.iou_thresh = anchor_size, anchor_y, pre_nms, iou_thresh

    def __call__(self, bbox_cls, bbox_reg, bbox_centerness):
        """-> (boxes [N, k, 7] in score order, scores [N, k], keep [N, k] int64 kept positions padded with -1,
        num_keep [N] int32), k = min(pre_nms, cells * anchors)."""
        N, A, Z, X = bbox_cls.shape
        dev = bbox_cls.device
        zs, xs = self.zs, self.xs
        angles = torch.arange(A, device=dev, dtype=torch.float32) * (np.pi / A)
        score = (torch.sigmoid(bbox_cls) * torch.sigmoid(bbox_centerness)).reshape(N, -1)          # [N, A*Z*X]
        k = min(self.pre_nms, score.shape[1])
        top, idx = score.topk(k, dim=1)
        a = idx // (Z * X)
        cell = idx % (Z * X)
        zi, xi = cell // X, cell % X
        reg = bbox_reg.reshape(N, A, -1, Z * X)                                                      # [N, A, R, Z*X]
        if reg.shape[2] != 7:
            raise RuntimeError("ProposalDecoder: 7-parameter regression expected (cfg.box_corner_parameters = False)")
        sel = reg[torch.arange(N, device=dev)[:, None], a, :, cell]                                  # [N, k, 7]
        h0, w0, l0 = self.anchor_size
        boxes = torch.stack([xs[xi] + sel[..., 0], self.anchor_y + sel[..., 1], zs[zi] + sel[..., 2],
                             l0 * torch.exp(sel[..., 5].clamp(-2, 2)), w0 * torch.exp(sel[..., 4].clamp(-2, 2)),
                             h0 * torch.exp(sel[..., 3].clamp(-2, 2)), angles[a] + sel[..., 6]], dim=-1)
        # BEV NMS works on [x, y(bev) = z, z, dx, dy, dz, heading]: put the ground-plane axes first
        bev_boxes = torch.stack([boxes[..., 0], boxes[..., 2], boxes[..., 1], boxes[..., 3], boxes[..., 4], boxes[..., 5],
                                 boxes[..., 6]], dim=-1).contiguous()
        keep, num = SF.nms_gpu_device_batched(bev_boxes, top, self.iou_thresh)       # all pairs in two launches
        return boxes, top, keep, num


def decode_proposals(bbox_cls, bbox_reg, bbox_centerness, cfg, anchor_size=(1.56, 1.6, 3.9), anchor_y=1.0, pre_nms=512,
                     iou_thresh=0.25):
    """One-shot form of `ProposalDecoder` (builds the cell-centre vectors on every call; not graph-capturable)."""
    return ProposalDecoder(cfg, bbox_cls.device, anchor_size, anchor_y, pre_nms, iou_thresh)(bbox_cls, bbox_reg, bbox_centerness)


class HostPipeline:
    """Host-buffer front end of `GlobalHotPath`: batches come from / go back to PINNED host memory.

    The reference moves every batch with blocking `.cuda()` / `.cpu()` calls around the model
    (tools/inference_agnostic.py:389-395, 605-640).  On B200 the lifted volume of 8 pairs is 0.6 GB, so
    the device-to-host transfer (PCIe) takes longer than the whole hot path; this front end therefore runs
    three streams -- host-to-device, compute (the caller's current stream), device-to-host -- over
    `depth` slots so that the transfers of batch i-1 / i+1 overlap the kernels of batch i (PCIe is full
    duplex).  `submit` enqueues one batch and returns immediately; `drain` waits for everything.

    Return path (`sparse_return=True`, the default): 42 % of the KITTI voxel grid lies outside the camera frustum and is
    exactly zero, so only the in-frustum voxel rows cross PCIe: `snvc_masked_rows_to_host` writes them straight into the
    caller's pinned buffer at their dense positions and zero-fills only rows that held data from the batch the buffer
    received before (a per-buffer device-side mask remembers which rows are non-zero; the first use of a buffer writes
    every row).  After `drain()` / the returned event the host buffer equals the dense tensor bit for bit.  Contract:
    a host output buffer that has been submitted must not be written by anyone else between submits; call
    `forget(h_out)` (or pass a fresh buffer) if it was.  `sparse_return=False` is the plain dense `cudaMemcpyAsync`."""

    def __init__(self, model, depth=2, out_dtype=torch.bfloat16, layout_out="NDHWC", graphed=True, sparse_return=True,
                 return_blocks=0):
        """`graphed`: every slot owns a `GraphedHotPath` (captured at the first batch of a new shape); the host then
        issues three copies and one graph launch per batch instead of ~16 kernel launches, which keeps the pipeline
        PCIe-bound when the host is busy (eager: 240 - 680 pairs/s for the same work, profiles/r01_ab_kw_kd.txt).
        `return_blocks`: grid size of the return kernel (0 = library default)."""
        self.model, self.depth, self.out_dtype, self.layout_out = model, depth, out_dtype, layout_out
        dev = next(model.parameters()).device
        self.dev = dev
        self.graphed = graphed
        self.sparse_return = bool(sparse_return) and layout_out == "NDHWC"
        self.return_blocks = int(return_blocks)
        self.s_in, self.s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
        self.slots = [dict(inputs=None, computed=None, copied_out=None, graph=None, shape=None) for _ in range(depth)]
        self.n = 0
        self._prev_valid = {}                                # host buffer (data_ptr, numel) -> device mask of its non-zero rows
        self.moved_bytes = torch.zeros((), dtype=torch.int64, device=dev)   # bytes the return kernel wrote to the host

    def forget(self, h_out=None):
        """Drop what the pipeline knows about the contents of `h_out` (all buffers if None): its next use rewrites every row."""
        if h_out is None:
            self._prev_valid.clear()
        else:
            self._prev_valid.pop((h_out.data_ptr(), h_out.numel()), None)

    def _return(self, vox, valid, h_out):
        """Enqueue the device-to-host return of `vox` on the current stream (the pipeline's output stream)."""
        from snvc_b200 import _lib
        if h_out.numel() != vox.numel():
            # result stays on the device (valid until the slot's next batch): only a digest -- the first
            # h_out.numel() values of every pair -- goes back, enough for the host to observe completion
            h_out.copy_(vox.reshape(vox.shape[0], -1)[:, :h_out.numel() // vox.shape[0]].reshape(h_out.shape), non_blocking=True)
            return
        row_bytes = vox.shape[-1] * vox.element_size()
        if not self.sparse_return or valid is None or row_bytes not in (16, 32, 64, 128) or h_out.dtype != vox.dtype:
            h_out.copy_(vox, non_blocking=True)
            return
        key = (h_out.data_ptr(), h_out.numel())
        prev = self._prev_valid.get(key)
        if prev is None or prev.numel() != valid.numel():
            prev = torch.ones(valid.numel(), dtype=torch.uint8, device=self.dev)     # unknown contents: write every row
            self._prev_valid[key] = prev
        with torch.cuda.device(self.dev):
            st = _lib.lib().snvc_masked_rows_to_host(vox.data_ptr(), valid.data_ptr(), prev.data_ptr(), h_out.data_ptr(),
                                                     valid.numel(), row_bytes, self.return_blocks,
                                                     self.moved_bytes.data_ptr(), _lib.stream_ptr())
        _lib.check(st, "snvc_masked_rows_to_host")

    def _submit_graphed(self, slot, h_left, h_right, h_shift, h_proj, h_out):
        cur = torch.cuda.current_stream(self.dev)
        shape = (tuple(h_left.shape), tuple(h_shift.shape))
        if slot["graph"] is None or slot["shape"] != shape:
            self.drain()
            N, C, H, W = h_left.shape
            slot["graph"] = GraphedHotPath(self.model, N, C, (H, W), h_shift.shape[1], self.out_dtype, self.layout_out)
            slot["shape"] = shape
            slot["computed"] = slot["copied_out"] = None
        g = slot["graph"]
        with torch.cuda.stream(self.s_in):
            if slot["computed"] is not None:          # the slot's previous inputs must have been consumed
                self.s_in.wait_event(slot["computed"])
            for d, h in zip(g.inputs, (h_left, h_right, h_shift, h_proj)):
                d.copy_(h, non_blocking=True)
            ready = torch.cuda.Event()
            ready.record(self.s_in)
        cur.wait_event(ready)
        if slot["copied_out"] is not None:            # the slot's previous result must have left the device
            cur.wait_event(slot["copied_out"])
        vox = g.replay()
        slot["computed"] = torch.cuda.Event()
        slot["computed"].record(cur)
        with torch.cuda.stream(self.s_out):
            self.s_out.wait_event(slot["computed"])
            self._return(vox, g.valid, h_out)
            slot["copied_out"] = torch.cuda.Event()
            slot["copied_out"].record(self.s_out)
        return slot["copied_out"]

    def submit(self, h_left, h_right, h_shift, h_proj, h_out):
        """All arguments are pinned host tensors; `h_out` receives the lifted voxels.  (An `h_out` smaller than the voxel
        volume receives a per-pair digest instead and the volume stays on the device for an on-device consumer -- the
        reference's own pipeline feeds it to the RPN without leaving the GPU.)"""
        for t in (h_left, h_right, h_shift, h_proj, h_out):
            if t.is_cuda or not t.is_pinned():
                raise RuntimeError("HostPipeline.submit expects pinned host tensors")
        slot = self.slots[self.n % self.depth]
        self.n += 1
        if self.graphed:
            return self._submit_graphed(slot, h_left, h_right, h_shift, h_proj, h_out)
        cur = torch.cuda.current_stream(self.dev)
        with torch.cuda.stream(self.s_in):
            if slot["computed"] is not None:          # the slot's previous inputs must have been consumed
                self.s_in.wait_event(slot["computed"])
            if slot["inputs"] is None or slot["inputs"][0].shape != h_left.shape:
                slot["inputs"] = tuple(torch.empty(t.shape, dtype=t.dtype, device=self.dev)
                                       for t in (h_left, h_right, h_shift, h_proj))
            for d, h in zip(slot["inputs"], (h_left, h_right, h_shift, h_proj)):
                d.copy_(h, non_blocking=True)
            ready = torch.cuda.Event()
            ready.record(self.s_in)
        cur.wait_event(ready)
        vox, valid = self.model(*slot["inputs"], self.out_dtype, self.layout_out, return_valid=True)
        slot["computed"] = torch.cuda.Event()
        slot["computed"].record(cur)
        with torch.cuda.stream(self.s_out):
            self.s_out.wait_event(slot["computed"])
            self._return(vox, valid, h_out)
            vox.record_stream(self.s_out)
            valid.record_stream(self.s_out)
            slot["copied_out"] = torch.cuda.Event()
            slot["copied_out"].record(self.s_out)
        return slot["copied_out"]

    def drain(self):
        self.s_in.synchronize()
        torch.cuda.current_stream(self.dev).synchronize()
        self.s_out.synchronize()
