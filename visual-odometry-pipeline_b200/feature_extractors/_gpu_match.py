"""Shared glue of the SIFT / ORB plug-ins: numpy descriptors in, (K,2) int64 matches out, all distance work and
the acceptance rule on the GPU through libvo_b200 (vo_match_u8 / vo_match_f32)."""
import numpy as np
import torch

import _bootstrap  # noqa: F401
from vo_b200 import ops


def _dev(a, dtype):
    if isinstance(a, torch.Tensor):
        return a.to(device="cuda", dtype=dtype).contiguous()
    return torch.from_numpy(np.ascontiguousarray(a, dtype={torch.uint8: np.uint8, torch.float32: np.float32}[dtype])).cuda()


def knn_ratio_u8(ref_desc, cur_desc, ratio=0.85, norm=ops.VO_NORM_L2_U8):
    res = ops.match_u8(_dev(ref_desc, torch.uint8), _dev(cur_desc, torch.uint8), norm, ops.VO_MODE_RATIO, ratio,
                       want_dist=False)
    return res.numpy()


def mutual_u8(ref_desc, cur_desc, norm=ops.VO_NORM_HAMMING):
    res = ops.match_u8(_dev(ref_desc, torch.uint8), _dev(cur_desc, torch.uint8), norm, ops.VO_MODE_MUTUAL, 0.0,
                       want_dist=False)
    return res.numpy()


def knn_ratio_f32(ref_desc, cur_desc, ratio=0.85, tag="sift"):
    a, b = _dev(ref_desc, torch.float32), _dev(cur_desc, torch.float32)
    # OpenCV SIFT descriptors are integer-valued in [0,255]: one fp16 pass (11 significant bits) is then exact, and so is
    # the fp32 accumulation (128 x 255 x 510 < 2^24).  Checked on EVERY call, on both frames (one fused reduction each):
    # RootSIFT / normalised descriptors or another extractor under the same tag take the split 3xTF32 pass instead.
    exact = a.numel() > 0 and b.numel() > 0 and all(
        bool(((t == t.round()) & (t >= 0) & (t <= 255)).all().item()) for t in (a, b))
    prec = ops.VO_PREC_F16X1 if exact else ops.VO_PREC_TF32X3
    if a.shape[-1] != 128:
        prec = ops.VO_PREC_FP32_SIMT
    res = ops.match_f32(a, b, ops.VO_METRIC_L2, ops.VO_MODE_RATIO, ratio, precision=prec, want_dist=False)
    return res.numpy()
