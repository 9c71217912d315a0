"""SIFT plug-in — same module-level interface as the reference's feature_extractors/SIFT.py
(`extract_features_and_desc(image) -> (kp, desc)` :14-23, `get_matches(...) -> int[K,2]` :25-34).

Detection / description runs on the GPU (vo_sift_extract, csrc/sift.cu: same keypoints and descriptors as OpenCV to the
tolerance of sift_frontend.py, tests/test_gpu_sift_frontend.py); VO_EXTRACTOR=opencv keeps cv2 on the CPU.  Matching —
brute-force 2-NN in L2 plus Lowe's 0.85 ratio test — runs on the tensor cores through vo_match_f32.
"""
import os

import cv2
import numpy as np

from feature_extractors import _gpu_match

EXTRACTOR = os.environ.get("VO_EXTRACTOR", "gpu")  # or "opencv"
_sift = None
_gpu_sift = {}


def _detector():
    global _sift
    if _sift is None:
        make = getattr(getattr(cv2, "xfeatures2d", None), "SIFT_create", None) or cv2.SIFT_create
        _sift = make()
    return _sift


def extract_features_and_desc(image):
    if EXTRACTOR == "gpu":
        from vo_b200.sift_frontend import SiftExtractor
        key = tuple(image.shape[:2])
        if key not in _gpu_sift:
            _gpu_sift[key] = SiftExtractor(*key)
        kp, desc, _ = _gpu_sift[key].extract(np.ascontiguousarray(image))
        return kp.cpu().numpy().astype(np.float64), desc.cpu().numpy()
    gray = cv2.cvtColor(image, cv2.COLOR_BGR2GRAY)
    kps, desc = _detector().detectAndCompute(gray, None)
    return np.asarray([[k.pt[0], k.pt[1]] for k in kps]), desc


def get_matches(ref_kp, ref_desc, cur_kp, cur_desc, img_shape, pix_rad=100, flag=2):
    if flag != 2:
        return None  # the reference defines flag == 2 only (SIFT.py:26)
    return _gpu_match.knn_ratio_f32(ref_desc, cur_desc, 0.85, tag="sift")
