"""ORB plug-in — interface of the reference's feature_extractors/ORB.py (:10-21, :23-32).

Default matcher = what the reference builds: `cv2.BFMatcher()` is NORM_L2, i.e. L2 over the 32 byte VALUES of
each descriptor, 2-NN + 0.85 ratio.  Set `orb_matcher: hamming_mutual` in config/vo_params.yaml (or
MATCHER below) for the north-star semantics: 256-bit Hamming + mutual nearest neighbour.
"""
import cv2
import numpy as np

from feature_extractors import _gpu_match

import os

MATCHER = "l2_ratio"  # or "hamming_mutual"
# Extraction: "gpu" (vo_orb_extract, csrc/orb.cu: the same keypoint set, bit-identical pt / angle / response /
# descriptors as cv2.ORB_create().detectAndCompute — tests/test_gpu_orb_frontend.py on a B200 — in level-major /
# row-major order instead of OpenCV's unspecified one) or "opencv" (cv2 on the CPU, as the reference; VO_EXTRACTOR=opencv).
EXTRACTOR = os.environ.get("VO_EXTRACTOR", "gpu")
_orb = None
_gpu_orb = {}


def _detector():
    global _orb
    if _orb is None:
        _orb = cv2.ORB_create()
    return _orb


def extract_features_and_desc(image):
    if EXTRACTOR == "gpu":
        from vo_b200.orb_frontend import OrbExtractor
        key = tuple(image.shape[:2])
        if key not in _gpu_orb:
            _gpu_orb[key] = OrbExtractor(*key)
        kp, desc, _ = _gpu_orb[key].extract(np.ascontiguousarray(image))
        return kp.cpu().numpy().astype(np.float64), desc.cpu().numpy()
    gray = cv2.cvtColor(image, cv2.COLOR_BGR2GRAY)
    kps, desc = _detector().detectAndCompute(gray, None)
    return np.asarray([[k.pt[0], k.pt[1]] for k in kps]), desc


def get_matches(ref_kp, ref_desc, cur_kp, cur_desc, img_shape, pix_rad=100, flag=2):
    if flag != 2:
        return None
    if MATCHER == "hamming_mutual":
        return _gpu_match.mutual_u8(ref_desc, cur_desc)
    return _gpu_match.knn_ratio_u8(ref_desc, cur_desc, 0.85)
