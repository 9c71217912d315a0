"""R2D2 plug-in — the matcher half of the reference's R2D2.py on the tensor cores.

Interface kept: `mnn_matcher` (:29-37), `similarity_matcher` (:40-51), `ratio_mutual_nn_matcher` (:53-66),
`get_matches(ref_kp, ref_desc, cur_kp, cur_desc, imgshape)` (:234-236) and `extract_features_and_desc(image)`
(:202-232).  Descriptors are L2-normalised fp32 CUDA tensors, exactly what the reference's network emits.

The reference materialises `sim = d1 @ d2.t()` (N x M fp32 in HBM) and runs topk / max over it; here one fused
tcgen05 kernel (3xTF32, accumulator in TMEM) produces the row top-2 and the column arg-max directly and a small
finalize kernel applies `ratio <= 0.90 and mutual` (or the similarity threshold) — see csrc/match_f32_tc.cu.

The R2D2 network itself is a front-end outside the accelerated path and is not redistributed
(CC BY-NC-SA): `extract_features_and_desc` loads it from a naver/r2d2 checkout at feature_extractors/r2d2 if the
user has placed one there, and raises otherwise.
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
    prec = ops.VO_PREC_TF32X3 if a.shape[-1] == 128 else ops.VO_PREC_FP32_SIMT
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
# Front-end (not accelerated here): thin loader around a user-supplied naver/r2d2 checkout.
args = {"model": "feature_extractors/r2d2/models/faster2d2_WASF_N16.pt", "reliability_thr": 0.7,
        "repeatability_thr": 0.7, "score_thr": 0.85}
_net = None


def _load_frontend():
    global _net
    if _net is not None:
        return _net
    root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "feature_extractors", "r2d2")
    if not os.path.isdir(root) or not os.path.exists(args["model"]):
        raise RuntimeError(
            "R2D2 feature extraction needs a naver/r2d2 checkout at feature_extractors/r2d2 (with "
            "models/faster2d2_WASF_N16.pt); it is a third-party front-end outside this package's scope. "
            "The matchers in this module work on any L2-normalised 128-d descriptors.")
    sys.path.insert(1, root)
    from nets.patchnet import Fast_Quad_L2Net_ConfCFS  # noqa: F401  (names resolved by eval below)
    import nets.patchnet as patchnet
    ckpt = torch.load(args["model"], map_location="cpu")
    net = eval("patchnet." + ckpt["net"])
    net.load_state_dict({k.replace("module.", ""): v for k, v in ckpt["state_dict"].items()})
    _net = net.eval().cuda()
    return _net


def extract_features_and_desc(image, trt=False):
    """image: HxWx3 uint8 -> (keypoints (N,3) [x, y, scale] numpy, descriptors (N,128) CUDA tensor)."""
    import cv2
    net = _load_frontend()
    rgb = cv2.cvtColor(image, cv2.COLOR_BGR2RGB).astype(np.float32) / 255.0
    mean, std = np.array([0.485, 0.456, 0.406], np.float32), np.array([0.229, 0.224, 0.225], np.float32)
    x = torch.from_numpy(((rgb - mean) / std).transpose(2, 0, 1))[None].cuda()
    with torch.no_grad():
        out = net(imgs=[x])
    desc, rel, rep = out["descriptors"][0], out["reliability"][0], out["repeatability"][0]
    peak = rep == torch.nn.functional.max_pool2d(rep, 3, 1, 1)
    keep = peak & (rep >= args["repeatability_thr"]) & (rel >= args["reliability_thr"])
    ys, xs = keep[0, 0].nonzero(as_tuple=True)
    score = rel[0, 0, ys, xs] * rep[0, 0, ys, xs]
    sel = score > args["score_thr"]
    ys, xs = ys[sel], xs[sel]
    d = desc[0, :, ys, xs].t().contiguous()
    kp = torch.stack([xs.float(), ys.float(), torch.full_like(xs, 32.0, dtype=torch.float32)], 1)
    return kp.cpu().numpy(), d
