"""Generates tests/golden/orb_pattern.npy and tests/golden/orb_golden.npz (run in the build container, where
/root/reference and OpenCV are present; the fixtures travel, this script's inputs do not).

orb_pattern.npy  the 256 learned rBRIEF point pairs (`bit_pattern_31_`, 512 points of int8 x, y).  The table is data of
                 OpenCV's features2d module (modules/features2d/src/orb.cpp, Apache-2.0) — the third-party dependency
                 that holds the arithmetic of the reference's ORB plug-in — and is read out of the installed cv2 binary.
orb_golden.npz   a synthetic BGR image and what the reference's OWN plug-in returns for it
                 (feature_extractors/ORB.py:10-21 imported from /root/reference), plus the cv2.KeyPoint fields of the
                 same call, for the oracle in oracle/orb_frontend.py.
"""
import glob
import os
import struct
import sys

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def extract_pattern():
    head = struct.pack("<8i", 8, -3, 9, 5, 4, 2, 7, -12)       # first two point pairs of bit_pattern_31_
    for f in glob.glob(os.path.join(os.path.dirname(cv2.__file__), "*.so")):
        blob = open(f, "rb").read()
        i = blob.find(head)
        if i >= 0:
            assert blob.find(head, i + 1) < 0, "pattern head is not unique in the binary"
            tab = np.frombuffer(blob[i:i + 256 * 4 * 4], dtype="<i4")
            assert tab.min() >= -15 and tab.max() <= 15
            return tab.astype(np.int8).reshape(512, 2)
    raise RuntimeError("bit_pattern_31_ not found in the cv2 binary")


def synthetic_bgr(seed, h, w):
    rng = np.random.default_rng(seed)
    base = rng.integers(0, 256, (h // 9, w // 9, 3), dtype=np.uint8)
    img = cv2.resize(base, (w, h), interpolation=cv2.INTER_CUBIC)
    img = cv2.add(img, rng.integers(0, 25, img.shape, dtype=np.uint8))
    for _ in range(12):                                            # flat rectangles: tie-heavy FAST / Harris scores
        x, y = int(rng.integers(20, w - 60)), int(rng.integers(20, h - 50))
        cv2.rectangle(img, (x, y), (x + int(rng.integers(8, 40)), y + int(rng.integers(8, 30))),
                      tuple(int(v) for v in rng.integers(0, 256, 3)), -1)
    return img


def main():
    np.save(os.path.join(HERE, "orb_pattern.npy"), extract_pattern())
    sys.path.insert(0, "/root/reference")
    from feature_extractors import ORB as ref_orb                 # the reference's plug-in, unmodified
    img = synthetic_bgr(8214, 240, 416)
    kp, desc = ref_orb.extract_features_and_desc(img)
    kps, desc2 = ref_orb.orb.detectAndCompute(cv2.cvtColor(img, cv2.COLOR_BGR2GRAY), None)
    assert np.array_equal(desc, desc2)
    np.savez_compressed(os.path.join(HERE, "orb_golden.npz"), image=img, kp=kp, desc=desc,
                        octave=np.array([k.octave for k in kps], np.int32),
                        angle=np.array([k.angle for k in kps], np.float32),
                        response=np.array([k.response for k in kps], np.float32),
                        size=np.array([k.size for k in kps], np.float32),
                        cv2_version=np.array(cv2.__version__))
    print("orb_golden:", img.shape, len(kps), "keypoints, cv2", cv2.__version__)


if __name__ == "__main__":
    main()
