"""Generates tests/golden/sift_golden.npz (run in the build container): a synthetic BGR image and what the reference's OWN
SIFT plug-in returns for it (feature_extractors/SIFT.py:14-23 imported from /root/reference), plus the cv2.KeyPoint fields of
the same detector, for oracle/sift_frontend.py.  Harness-side shim only: this image's OpenCV has SIFT in the main module,
the reference asks for cv2.xfeatures2d.SIFT_create (opencv-contrib 4.5.4) — same class, same defaults."""
import os
import sys
import types

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_orb_golden import synthetic_bgr  # noqa: E402


def main():
    if not hasattr(cv2, "xfeatures2d"):
        cv2.xfeatures2d = types.SimpleNamespace(SIFT_create=cv2.SIFT_create)
    sys.path.insert(0, "/root/reference")
    from feature_extractors import SIFT as ref_sift            # the reference's plug-in, unmodified
    img = synthetic_bgr(8215, 240, 416)
    kp, desc = ref_sift.extract_features_and_desc(img)
    kps, desc2 = ref_sift.sift.detectAndCompute(cv2.cvtColor(img, cv2.COLOR_BGR2GRAY), None)
    assert np.array_equal(desc, desc2) and np.array_equal(desc, np.rint(desc)) and desc.max() <= 255
    np.savez_compressed(os.path.join(HERE, "sift_golden.npz"), image=img, kp=kp, desc=desc.astype(np.uint8),
                        size=np.array([k.size for k in kps], np.float32), angle=np.array([k.angle for k in kps], np.float32),
                        response=np.array([k.response for k in kps], np.float32),
                        octave=np.array([k.octave for k in kps], np.int64), cv2_version=np.array(cv2.__version__))
    print("sift_golden:", img.shape, len(kps), "keypoints, cv2", cv2.__version__)


if __name__ == "__main__":
    main()
