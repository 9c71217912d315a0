"""The ORB front-end oracle (oracle/orb_frontend.py) against the reference's own plug-in.

Golden: tests/golden/orb_golden.npz holds a synthetic BGR image and what /root/reference/feature_extractors/ORB.py
returned for it (tests/golden/make_orb_golden.py).  Keypoints are compared as SETS (OpenCV's order depends on
std::nth_element inside retainBest); everything attached to a keypoint must agree bit for bit."""
import os

import numpy as np
import pytest

from oracle import orb_frontend as of

HERE = os.path.dirname(os.path.abspath(__file__))


def _by_key(level, x, y, *fields):
    return {(int(l), int(a), int(b)): tuple(f[i] for f in fields) for i, (l, a, b) in enumerate(zip(level, x, y))}


def _cv_keys(pt, octave, scales):
    s = np.array([scales[o] for o in octave], np.float32)
    return np.rint(pt[:, 0] / s).astype(int), np.rint(pt[:, 1] / s).astype(int)


def test_extract_features_and_desc_equals_reference_plugin(golden):
    g = golden("orb_golden.npz")
    kp, desc = of.extract_features_and_desc(g["image"])
    assert kp.dtype == np.float64 and kp.shape == g["kp"].shape and desc.shape == g["desc"].shape
    out = of.detect_and_compute(of.bgr_to_gray(g["image"]))
    scales = of.level_scales()
    gx, gy = _cv_keys(g["kp"].astype(np.float32), g["octave"], scales)
    want = _by_key(g["octave"], gx, gy, g["kp"].astype(np.float32), g["angle"], g["response"], g["size"], g["desc"])
    got = _by_key(out["level"], out["xl"], out["yl"], out["pt"], out["angle"], out["response"], out["size"], out["desc"])
    assert set(want) == set(got) and len(want) == len(g["kp"])          # same keypoints, none duplicated
    for key, (pt, ang, resp, size, d) in want.items():
        p2, a2, r2, s2, d2 = got[key]
        assert np.array_equal(pt, p2) and ang == a2 and resp == r2 and size == s2, key
        assert np.array_equal(d, d2), key
    assert len(set(out["level"])) == 8                                   # every pyramid level contributes


def test_stages_equal_opencv():
    """Stage by stage against the installed OpenCV (skipped where cv2 is absent): gray conversion, INTER_LINEAR_EXACT
    resize cascade, FAST + non-maximum suppression, the float Gaussian, fastAtan2."""
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(8214)
    bgr = rng.integers(0, 256, (97, 163, 3), dtype=np.uint8)
    gray = of.bgr_to_gray(bgr)
    assert np.array_equal(gray, cv2.cvtColor(bgr, cv2.COLOR_BGR2GRAY))
    tex = cv2.add(cv2.resize(rng.integers(0, 256, (30, 52), dtype=np.uint8), (416, 240), interpolation=cv2.INTER_CUBIC),
                  rng.integers(0, 30, (240, 416), dtype=np.uint8))
    cur = tex
    for scale in of.level_scales()[1:]:
        inv = np.float32(1.0) / scale
        dw, dh = int(np.rint(np.float32(416) * inv)), int(np.rint(np.float32(240) * inv))
        ref = cv2.resize(cur, (dw, dh), interpolation=cv2.INTER_LINEAR_EXACT)
        assert np.array_equal(of.resize_linear_exact(cur, dw, dh), ref)
        cur = ref
    kps = cv2.FastFeatureDetector_create(20, True).detect(tex, None)
    xs, ys, sc = of.fast_detect(tex, 20)
    assert len(kps) > 500
    assert np.array_equal(np.stack([xs, ys, sc], 1), np.array([[k.pt[0], k.pt[1], k.response] for k in kps]))
    k = cv2.getGaussianKernel(7, 2, cv2.CV_32F)
    assert np.array_equal(of.gaussian_kernel_7_2(), k[:, 0])
    assert np.array_equal(of.blur_7x7(tex), cv2.sepFilter2D(tex, cv2.CV_8U, k, k, borderType=cv2.BORDER_REFLECT_101))
    for y, x in rng.integers(-5000, 5000, (500, 2)):
        assert of.fast_atan2(y, x) == np.float32(cv2.fastAtan2(float(y), float(x)))


def test_whole_front_end_equals_opencv_on_other_images():
    """Fresh images, live cv2.ORB_create(): noise, a smooth image with few corners (fewer than the per-level quota),
    a flat image with rectangles (tie-heavy scores), a tiny image whose top levels are smaller than the border."""
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(99)
    rect = np.full((200, 330), 90, np.uint8)
    for _ in range(40):
        x, y = int(rng.integers(35, 250)), int(rng.integers(35, 140))
        cv2.rectangle(rect, (x, y), (x + int(rng.integers(8, 50)), y + int(rng.integers(8, 35))), int(rng.integers(0, 256)), -1)
    images = [rng.integers(0, 256, (150, 260), dtype=np.uint8),
              cv2.resize(rng.integers(0, 256, (9, 12), dtype=np.uint8), (320, 240), interpolation=cv2.INTER_CUBIC),
              rect,
              rng.integers(0, 256, (97, 163), dtype=np.uint8)]
    scales = of.level_scales()
    for img in images:
        kps, desc = cv2.ORB_create().detectAndCompute(img, None)
        out = of.detect_and_compute(img)
        assert len(kps) == len(out["level"])
        if not kps:
            continue
        pt = np.array([k.pt for k in kps], np.float32)
        octave = np.array([k.octave for k in kps])
        gx, gy = _cv_keys(pt, octave, scales)
        want = _by_key(octave, gx, gy, pt, np.array([k.angle for k in kps], np.float32),
                       np.array([k.response for k in kps], np.float32), desc)
        got = _by_key(out["level"], out["xl"], out["yl"], out["pt"], out["angle"], out["response"], out["desc"])
        assert set(want) == set(got)
        for key, (p, a, r, d) in want.items():
            p2, a2, r2, d2 = got[key]
            assert np.array_equal(p, p2) and a == a2 and r == r2 and np.array_equal(d, d2), key


def test_non_default_quotas_and_level_counts_equal_opencv():
    """nfeatures / nlevels other than the defaults (the bench's ORB-5k setting among them)."""
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(7)
    img = rng.integers(0, 256, (180, 300), dtype=np.uint8)
    scales = of.level_scales()
    for nfeatures, nlevels in ((5000, 8), (1500, 5), (60, 3)):
        kps, desc = cv2.ORB_create(nfeatures=nfeatures, nlevels=nlevels).detectAndCompute(img, None)
        out = of.detect_and_compute(img, nfeatures=nfeatures, nlevels=nlevels)
        assert len(kps) == len(out["level"]) > 0
        pt = np.array([k.pt for k in kps], np.float32)
        octave = np.array([k.octave for k in kps])
        gx, gy = _cv_keys(pt, octave, scales)
        want = _by_key(octave, gx, gy, pt, np.array([k.angle for k in kps], np.float32),
                       np.array([k.response for k in kps], np.float32), desc)
        got = _by_key(out["level"], out["xl"], out["yl"], out["pt"], out["angle"], out["response"], out["desc"])
        assert set(want) == set(got)
        for key, (p, a, r, d) in want.items():
            p2, a2, r2, d2 = got[key]
            assert np.array_equal(p, p2) and a == a2 and r == r2 and np.array_equal(d, d2), key
