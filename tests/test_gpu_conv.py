"""tcgen05 implicit-GEMM convolution (vo_conv2d) against torch.nn.functional.conv2d in fp32 (TF32 disabled): the layer
shapes of the R2D2 network (3x3 and 2x2 taps, dilations 1..8, 32/64/128 channels), zero padding at every border."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("H,W,cin,cout,k,dil,relu", [
    (20, 150, 32, 32, 3, 1, True),       # W not a multiple of the 128-pixel tile
    (17, 128, 32, 64, 3, 1, True),
    (33, 300, 64, 64, 3, 1, False),
    (24, 260, 64, 128, 3, 1, True),
    (40, 200, 128, 128, 3, 2, True),     # dilated
    (40, 200, 128, 128, 2, 2, False),    # 2x2 taps, padding 1
    (40, 200, 128, 128, 2, 4, False),
    (48, 140, 128, 128, 2, 8, False),
    (5, 9, 32, 32, 3, 1, False),         # tile mostly out of bounds
])
def test_conv_matches_torch_fp32(H, W, cin, cout, k, dil, relu):
    import torch
    import torch.nn.functional as F
    from vo_b200 import ops
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    g = torch.Generator(device="cuda").manual_seed(H * 1000 + W + cin + cout + k + dil)
    x = torch.randn(H, W, cin, device="cuda", generator=g)
    w = torch.randn(cout, k, k, cin, device="cuda", generator=g) / np.sqrt(k * k * cin)
    scale = torch.rand(cout, device="cuda", generator=g) + 0.5
    shift = torch.randn(cout, device="cuda", generator=g)
    got = ops.conv2d(x, w, scale, shift, k, dil, relu)
    pad = ((k - 1) * dil) // 2
    ref = F.conv2d(x.permute(2, 0, 1)[None].double(), w.permute(0, 3, 1, 2).double(), padding=pad, dilation=dil)[0]
    ref = ref * scale.double()[:, None, None] + shift.double()[:, None, None]
    if relu:
        ref = ref.clamp_min(0)
    ref = ref.permute(1, 2, 0)
    assert got.shape == ref.shape
    err = (got.double() - ref).abs().max().item()
    assert err < 2e-5 * max(1.0, ref.abs().max().item()), err      # 3xTF32: fp32-grade
