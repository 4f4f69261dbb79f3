"""ORACLE (test infrastructure): plain-torch CPU restatement of the reference's 3-D conv blocks.

Each builder cites the reference lines it follows.  The arithmetic itself
(Conv3d / ConvTranspose3d / BatchNorm3d / GroupNorm) is the third-party dependency
torch (reference pin: pytorch-1.9.0, spec-file.txt:253; 2.11.0 here).  state_dict
key names and shapes equal the reference's, which is how
tests/test_oracle_blocks.py pins these classes: the reference modules' weights are
loaded with strict=True and outputs compared on the golden fixtures
(tests/golden/make_golden.py, generated from /root/reference in the build container).
"""
import torch
import torch.nn as nn
import torch.nn.functional as F


def _norm3d(ch, gn, groups=32):
    return nn.GroupNorm(groups, ch) if gn else nn.BatchNorm3d(ch)


def convbn_3d(cin, cout, kernel_size, stride, pad, dilation=1, gn=False, groups=32):
    """submodule.py:32-50."""
    return nn.Sequential(
        nn.Conv3d(cin, cout, kernel_size=kernel_size, padding=pad, dilation=dilation, stride=stride, bias=False),
        _norm3d(cout, gn, groups))


def _deconvbn_3d(cin, cout, gn):
    """submodule.py:127-147,197-208: ConvTranspose3d(k3,s2,p1,op1,bias=False) + norm."""
    return nn.Sequential(
        nn.ConvTranspose3d(cin, cout, kernel_size=3, padding=1, output_padding=1, stride=2, bias=False),
        _norm3d(cout, gn))


def _cbr(cin, cout, k, s, p, d=1, gn=False):
    return nn.Sequential(convbn_3d(cin, cout, k, s, p, d, gn=gn), nn.ReLU(inplace=True))


class Hourglass(nn.Module):
    """submodule.py:85-168 (class ``hourglass``)."""

    def __init__(self, inplanes, gn=False):
        super().__init__()
        c2 = inplanes * 2
        self.conv1 = _cbr(inplanes, c2, 3, 2, 1, gn=gn)           # :89-97
        self.conv2 = convbn_3d(c2, c2, 3, 1, 1, gn=gn)            # :99-105
        self.conv3 = _cbr(c2, c2, 3, 2, 1, gn=gn)                 # :107-115
        self.conv4 = _cbr(c2, c2, 3, 1, 1, gn=gn)                 # :117-125
        self.conv5 = _deconvbn_3d(c2, c2, gn)                     # :127-136
        self.conv6 = _deconvbn_3d(c2, inplanes, gn)               # :138-147

    def forward(self, x, presqu, postsqu):                        # :149-168
        out = self.conv1(x)
        pre = self.conv2(out)
        pre = F.relu(pre + postsqu) if postsqu is not None else F.relu(pre)
        out = self.conv4(self.conv3(pre))
        post = F.relu(self.conv5(out) + (presqu if presqu is not None else pre))
        return self.conv6(post), pre, post


class HourglassDownsample16(nn.Module):
    """submodule.py:223-268 (class ``hourglass_downsample_16``; helpers :170-208)."""

    def __init__(self, inplanes, gn=False):
        super().__init__()
        c2 = inplanes * 2
        self.conv1 = _cbr(inplanes, c2, 3, 2, 1, gn=gn)
        self.conv2 = _cbr(c2, c2, 3, 1, 1, gn=gn)
        for i in (3, 5, 7):
            setattr(self, f"conv{i}", _cbr(c2, c2, 3, 2, 1, gn=gn))
            setattr(self, f"conv{i + 1}", _cbr(c2, c2, 3, 1, 1, gn=gn))
        for i in (9, 10, 11):
            setattr(self, f"conv{i}", _deconvbn_3d(c2, c2, gn))
        self.conv12 = _deconvbn_3d(c2, inplanes, gn)

    def forward(self, x):                                         # :245-268
        o2 = self.conv2(self.conv1(x))
        o4 = self.conv4(self.conv3(o2))
        o6 = self.conv6(self.conv5(o4))
        o8 = self.conv8(self.conv7(o6))
        o10 = self.conv10(self.conv9(o8) + o6)
        o11 = self.conv11(o10 + o4)
        return self.conv12(o11 + o2)


class GlobalTrunk(nn.Module):
    """Restated global-branch 3-D trunk (DSGN lineage, README.md:68).

    dres0 = 2 x (convbn_3d 3^3 + ReLU)  64->32->32
    dres1 = convbn_3d+ReLU, convbn_3d ; out = dres1(x) + x
    hg    = hourglass(32)(x, None, None)[0] + x
    Blocks are the reference's (submodule.py:32-50,85-168); the wiring is restated
    because the reference does not ship the global model class (models/__init__.py:1-2).
    """

    def __init__(self, cin=64, ch=32, gn=False):
        super().__init__()
        self.dres0 = nn.Sequential(_cbr(cin, ch, 3, 1, 1, gn=gn), _cbr(ch, ch, 3, 1, 1, gn=gn))
        self.dres1 = nn.Sequential(_cbr(ch, ch, 3, 1, 1, gn=gn), convbn_3d(ch, ch, 3, 1, 1, gn=gn))
        self.hg = Hourglass(ch, gn=gn)

    def forward(self, cost):
        x = self.dres0(cost)
        x = self.dres1(x) + x
        return self.hg(x, None, None)[0] + x


class Vernier3D(nn.Module):
    """The 3-D part of VernierScale (vernier_type='BEV_type3'): layers vernier.py:250-289,
    forward vernier.py:414-438.  Attribute names equal the reference's so a VernierScale
    state_dict (filtered to these prefixes) loads with strict=True."""

    def __init__(self, dim=32, n_sample_w=128, gn=False, pool=4):
        super().__init__()
        self.n_sample_w = n_sample_w
        self.vimg_feat = _cbr(2 * dim, dim, 1, 1, 0, gn=gn)                 # :250-252
        self.conv1 = _cbr(2 * dim, dim, 7, 1, 3, gn=gn)                     # :253-255
        self.conv2 = _cbr(dim, dim, 5, 1, 2, gn=gn)                         # :256-258
        self.conv3 = _cbr(dim, dim, 5, 1, 4, d=2, gn=gn)                    # :259-261
        self.conv4 = _cbr(2 * dim, dim, 3, 1, 1, gn=gn)                     # :262-264
        self.hg_conv3d = Hourglass(dim, gn) if n_sample_w <= 16 else HourglassDownsample16(dim, gn)  # :265-268
        self.fg_cls_head = nn.Sequential(convbn_3d(dim, dim, 3, 1, 1, gn=gn), nn.ReLU(inplace=True),
                                         nn.Conv3d(dim, 1, 3, 1, 1, bias=False), nn.Sigmoid())      # :269-278
        self.pool_3d = nn.AvgPool3d((pool, 1, 1), stride=(pool, 1, 1))      # :289

    def forward(self, voxel):
        """-> (voxel_BEV [N, dim*nh/4, nw, nl], occupancy [N, nh, nw, nl])  (vernier.py:414-438)."""
        vimg = self.vimg_feat(voxel)
        voxel = self.conv1(voxel)
        voxel = self.conv2(voxel) + voxel
        voxel = self.conv3(voxel) + voxel
        if self.n_sample_w <= 16:
            voxel = self.hg_conv3d(voxel, None, None)[0] + voxel
        else:
            voxel = self.hg_conv3d(voxel) + voxel
        occupancy = self.fg_cls_head(voxel)
        voxel = torch.cat([voxel, vimg * occupancy], dim=1)
        voxel = self.pool_3d(self.conv4(voxel))
        N, Fc, H, W, L = voxel.shape
        return voxel.reshape(N, -1, W, L), occupancy.squeeze(1)


# ------------------------------------------------------------------------------------------------ 2-D BEV blocks (N2)
def _norm2d(ch, gn, groups=32):
    return nn.GroupNorm(groups, ch) if gn else nn.BatchNorm2d(ch)


def convbn(cin, cout, kernel_size, stride, pad, dilation, gn=False, groups=32):
    """submodule.py:11-29."""
    return nn.Sequential(
        nn.Conv2d(cin, cout, kernel_size=kernel_size, stride=stride, padding=dilation if dilation > 1 else pad,
                  dilation=dilation, bias=False),
        _norm2d(cout, gn, groups))


def _cbr2d(cin, cout, stride, gn=False):
    """submodule.py:183-195 (get_hg_down_sample_2d) / :321-322: convbn 3x3 + ReLU."""
    return nn.Sequential(convbn(cin, cout, 3, stride, 1, 1, gn=gn), nn.ReLU(inplace=True))


def _deconvbn_2d(cin, cout, gn):
    """submodule.py:210-221, 335-345: ConvTranspose2d(k3,s2,p1,op1,bias=False) + norm."""
    return nn.Sequential(nn.ConvTranspose2d(cin, cout, kernel_size=3, padding=1, output_padding=1, stride=2, bias=False),
                         _norm2d(cout, gn))


class Hourglass2d(nn.Module):
    """submodule.py:317-361 (class ``hourglass2d``)."""

    def __init__(self, inplanes, gn=False):
        super().__init__()
        c2 = inplanes * 2
        self.conv1 = _cbr2d(inplanes, c2, 2, gn)
        self.conv2 = convbn(c2, c2, 3, 1, 1, 1, gn=gn)
        self.conv3 = _cbr2d(c2, c2, 2, gn)
        self.conv4 = _cbr2d(c2, c2, 1, gn)
        self.conv5 = _deconvbn_2d(c2, c2, gn)
        self.conv6 = _deconvbn_2d(c2, inplanes, gn)

    def forward(self, x, presqu, postsqu):
        out = self.conv1(x)
        pre = self.conv2(out)
        pre = F.relu(pre + postsqu) if postsqu is not None else F.relu(pre)
        out = self.conv4(self.conv3(pre))
        post = F.relu(self.conv5(out) + (presqu if presqu is not None else pre))
        return self.conv6(post), pre, post


class Hourglass2dDownsample16(nn.Module):
    """submodule.py:270-315 (class ``hourglass2d_downsample_16``)."""

    def __init__(self, inplanes, gn=False):
        super().__init__()
        c2 = inplanes * 2
        self.conv1 = _cbr2d(inplanes, c2, 2, gn)
        self.conv2 = _cbr2d(c2, c2, 1, gn)
        for i in (3, 5, 7):
            setattr(self, f"conv{i}", _cbr2d(c2, c2, 2, gn))
            setattr(self, f"conv{i + 1}", _cbr2d(c2, c2, 1, gn))
        for i in (9, 10, 11):
            setattr(self, f"conv{i}", _deconvbn_2d(c2, c2, gn))
        self.conv12 = _deconvbn_2d(c2, inplanes, gn)

    def forward(self, x):
        o2 = self.conv2(self.conv1(x))
        o4 = self.conv4(self.conv3(o2))
        o6 = self.conv6(self.conv5(o4))
        o8 = self.conv8(self.conv7(o6))
        o10 = self.conv10(self.conv9(o8) + o6)
        o11 = self.conv11(o10 + o4)
        return self.conv12(o11 + o2)


class VernierBevTail(nn.Module):
    """The 2-D BEV tail of VernierScale (BEV_type2 / BEV_type3) up to the part heatmaps: layers vernier.py:289-314,
    forward vernier.py:440-445: conv5 (convbn 3x3 + ReLU) -> hm1 (hourglass2d[_downsample_16]) -> permute(0,1,3,2) -> hm2.
    Attribute names equal the reference's (conv5, hm1, hm2)."""

    def __init__(self, dim_height=256, num_parts=9, n_sample_w=128, gn=False):
        super().__init__()
        self.n_sample_w = n_sample_w
        self.conv5 = nn.Sequential(convbn(dim_height, 64, 3, 1, 1, 1, gn=gn), nn.ReLU(inplace=True))
        self.hm1 = Hourglass2d(64, gn) if n_sample_w <= 16 else Hourglass2dDownsample16(64, gn)
        self.hm2 = nn.Conv2d(64, num_parts, 3, 1, 1, bias=False)

    def forward(self, voxel_bev):
        x = self.conv5(voxel_bev)
        x = self.hm1(x, None, None)[0] if self.n_sample_w <= 16 else self.hm1(x)
        return self.hm2(x.permute(0, 1, 3, 2))


class RPN3DHead(nn.Module):
    """Restated RPN side of the global branch, plain torch: the shipped blocks convbn_3d
    (submodule.py:32-50), hourglass (:85-168), AvgPool3d + reshape to BEV (vernier.py:289,436-438), convbn (:11-29),
    hourglass2d (:317-361) and 3x3 Conv2d heads with the output conventions of RPN3DLoss (loss3d.py:84-103,253-275).
    Input: lifted voxels [N, C, Z, Y, X]; the pooled axis is Y."""

    def __init__(self, channels=32, n_y=20, pool=4, bev_dim=64, num_angles=4, num_classes=1, reg_dim=7, gn=False):
        super().__init__()
        A = num_angles * num_classes
        self.pool = pool
        self.rpn3d_conv = _cbr(channels, channels, 3, 1, 1, gn=gn)
        self.rpn3d_conv2 = Hourglass(channels, gn=gn)
        self.rpn3d_conv3 = _cbr2d(channels * (n_y // pool), bev_dim, 1, gn)
        self.rpn3d_conv4 = Hourglass2d(bev_dim, gn=gn)
        self.rpn3d_cls_convs = _cbr2d(bev_dim, bev_dim, 1, gn)
        self.rpn3d_bbox_convs = _cbr2d(bev_dim, bev_dim, 1, gn)
        self.bbox_cls = nn.Conv2d(bev_dim, A, 3, 1, 1)
        self.bbox_reg = nn.Conv2d(bev_dim, A * reg_dim, 3, 1, 1)
        self.bbox_centerness = nn.Conv2d(bev_dim, A, 3, 1, 1)

    def forward(self, vox):
        x = self.rpn3d_conv(vox)
        x = self.rpn3d_conv2(x, None, None)[0] + x
        x = x.permute(0, 1, 3, 2, 4)                                   # [N, C, Y, Z, X]: pool the height axis
        x = F.avg_pool3d(x, (self.pool, 1, 1), stride=(self.pool, 1, 1))
        N, C, Yq, Z, X = x.shape
        b = self.rpn3d_conv3(x.reshape(N, C * Yq, Z, X))
        b = self.rpn3d_conv4(b, None, None)[0] + b
        c, r = self.rpn3d_cls_convs(b), self.rpn3d_bbox_convs(b)
        return self.bbox_cls(c), self.bbox_reg(r), self.bbox_centerness(r)
