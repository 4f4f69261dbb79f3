"""CPU check of the algebra behind the product path's first layer: the reference's cost volume is
`cost[n, c, d, h, w] = left[n, c, h, w]` for c < C (BuildCostVolume_cuda.cu:84-86), so a 3x3x3 convolution over the
2C-channel volume equals the convolution of the right half plus a depth-invariant term that a 3-plane convolution of
the left features yields for output depth 0 / interior / D-1.  Pure torch fp32 on the oracle's own cost volume."""
import numpy as np
import torch
import torch.nn.functional as TF

from oracle import cost_volume as ocv


def test_first_layer_equals_right_half_conv_plus_three_plane_addend():
    rng = np.random.default_rng(0)
    N, C, H, W, D, Co = 2, 4, 6, 9, 5, 3
    left = rng.standard_normal((N, C, H, W)).astype(np.float32)
    right = rng.standard_normal((N, C, H, W)).astype(np.float32)
    shift = np.tile(np.float32([0.0, 0.5, 1.25, 3.0, 7.75])[None], (N, 1))
    cost = torch.from_numpy(ocv.forward_c(left, right, shift, 1, fma_mode=1))          # [N, 2C, D, H, W]
    for d in range(D):                                                                 # the left half is a broadcast over depth
        assert torch.equal(cost[:, :C, d], torch.from_numpy(left))
    w = torch.from_numpy(rng.standard_normal((Co, 2 * C, 3, 3, 3)).astype(np.float32))
    want = TF.conv3d(cost, w, padding=1)
    left3 = torch.from_numpy(left)[:, :, None].expand(N, C, 3, H, W).contiguous()
    addend = TF.conv3d(left3, w[:, :C].contiguous(), padding=1)                        # planes: depth 0 / interior / D-1
    got = TF.conv3d(cost[:, C:].contiguous(), w[:, C:].contiguous(), padding=1)
    variant = torch.tensor([0] + [1] * (D - 2) + [2])
    got = got + addend[:, :, variant]
    assert torch.allclose(got, want, rtol=0, atol=2e-5 * float(want.abs().max()))
