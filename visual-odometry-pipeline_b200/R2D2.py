"""R2D2 plug-in — the reference's R2D2.py (matchers and feature extraction) on the tensor cores.

Interface kept: `mnn_matcher` (:29-37), `similarity_matcher` (:40-51), `ratio_mutual_nn_matcher` (:53-66),
`get_matches(ref_kp, ref_desc, cur_kp, cur_desc, imgshape)` (:234-236) and `extract_features_and_desc(image)`
(:202-232).  Descriptors are L2-normalised fp32 CUDA tensors, exactly what the reference's network emits.

The reference materialises `sim = d1 @ d2.t()` (N x M fp32 in HBM) and runs topk / max over it; here one fused
tcgen05 kernel (3xTF32, accumulator in TMEM) produces the row top-2 and the column arg-max directly and a small
finalize kernel applies `ratio <= 0.90 and mutual` (or the similarity threshold) — see csrc/match_f32_tc.cu.

`extract_features_and_desc` runs the R2D2 network on the tensor cores too (csrc/conv_tc.cu, csrc/r2d2_net.cu); the
weights are read from the user's naver/r2d2 checkpoint file: the product ships none (they are CC BY-NC-SA 3.0, (c) NAVER;
the only copy in this repository is the parity-test fixture tests/golden/r2d2_net.npz, see tests/golden/NOTICE.md).
"""
import os
import sys

import numpy as np
import torch

import _bootstrap  # noqa: F401
from vo_b200 import ops


def _prep(d):
    if not isinstance(d, torch.Tensor):
        d = torch.from_numpy(np.ascontiguousarray(d, dtype=np.float32))
    return d.to(device="cuda", dtype=torch.float32).contiguous()


def _run(d1, d2, mode, param):
    a, b = _prep(d1), _prep(d2)
    # 3xTF32 by default (the north-star GEMM); VO_MATCH_PRECISION=f16x3 selects the split-fp16 pass: the same 22 operand
    # bits at twice the MMA rate, valid for |x| < 255 (R2D2 descriptors are unit-norm)
    prec = ops.VO_PREC_F16X3 if os.environ.get("VO_MATCH_PRECISION", "tf32x3").lower() == "f16x3" else ops.VO_PREC_TF32X3
    if a.shape[-1] != 128:
        prec = ops.VO_PREC_FP32_SIMT
    return ops.match_f32(a, b, ops.VO_METRIC_COSINE, mode, param, precision=prec, want_dist=True)


def mnn_matcher(descriptors_a, descriptors_b, threshold=0.9):
    """Mutual nearest neighbours with similarity >= threshold -> numpy int64 (K,2)."""
    return _run(descriptors_a, descriptors_b, ops.VO_MODE_THRESH_MUTUAL, threshold).numpy()


def similarity_matcher(descriptors1, descriptors2, threshold=0.9):
    """Nearest neighbour with similarity >= threshold -> (matches tensor (K,2), distances tensor (K,))."""
    res = _run(descriptors1, descriptors2, ops.VO_MODE_THRESH, threshold)
    k = int(res.count[0].item())
    return res.pairs[0, :k].to(torch.int64), res.dist[0, :k]


def ratio_mutual_nn_matcher(descriptors1, descriptors2, ratio=0.90):
    """Lowe ratio (<= ratio, on sqrt(2-2 sim)) AND mutual NN -> (numpy int64 (K,2), distances tensor (K,))."""
    res = _run(descriptors1, descriptors2, ops.VO_MODE_RATIO_MUTUAL, ratio)
    k = int(res.count[0].item())
    return res.pairs[0, :k].to(torch.int64).cpu().numpy(), res.dist[0, :k]


def get_matches(ref_kp, ref_desc, cur_kp, cur_desc, imgshape):
    return ratio_mutual_nn_matcher(ref_desc, cur_desc)[0]


# ---------------------------------------------------------------------------------------------------------
# Front-end: the network, the heads, NMS and the score filter run in libvo_b200.so (vo_r2d2_*, csrc/r2d2_net.cu +
# csrc/conv_tc.cu).  Only the checkpoint FILE is the user's: the package ships no weights (CC BY-NC-SA 3.0); point
# args['model'] at a naver/r2d2 checkpoint (the reference's default path is kept).
args = {"model": "feature_extractors/r2d2/models/faster2d2_WASF_N16.pt", "scale_f": 2 ** 0.25, "min_size": 256,
        "max_size": 1380, "min_scale": 0, "max_scale": 1, "reliability_thr": 0.7, "repeatability_thr": 0.7,
        "score_thr": 0.85, "gpu": [0]}
_nets = {}          # (H, W) -> R2D2Net
_weights = None


def _frontend(H, W):
    global _weights
    from vo_b200 import r2d2_frontend
    if _weights is None:
        if not os.path.exists(args["model"]):
            raise RuntimeError(f"R2D2 checkpoint {args['model']!r} not found: set R2D2.args['model'] to a naver/r2d2 model file "
                               "(e.g. faster2d2_WASF_N16.pt).  The matchers in this module work on any L2-normalised "
                               "128-d descriptors.")
        _weights = r2d2_frontend.load_checkpoint(args["model"])
    if (H, W) not in _nets:
        _nets[(H, W)] = r2d2_frontend.R2D2Net(_weights[0], _weights[1], H, W)
    return _nets[(H, W)]


def extract_features_and_desc(image, trt=False):
    """image: HxWx3 uint8 BGR (OpenCV) -> (keypoints (N,3) [x, y, scale] numpy float32, descriptors (N,128) CUDA
    tensor), as the reference returns them (R2D2.py:202-232).  Extraction happens at scale 1 only, like the
    reference (its multi-scale loop leaves after the first iteration, :133-135)."""
    rgb = np.ascontiguousarray(image[:, :, ::-1])                      # cv2.cvtColor(image, COLOR_BGR2RGB)
    net = _frontend(rgb.shape[0], rgb.shape[1])
    xys, desc, _ = net.extract(rgb, args["reliability_thr"], args["repeatability_thr"], args["score_thr"])
    return xys.cpu().numpy(), desc.clone()
