"""GPU SIFT front-end (vo_sift_extract) against the CPU restatement and the golden output of the reference's own plug-in,
to the tolerance of tests/test_oracle_sift.py.  Strict since its first pass on a B200."""
import numpy as np
import pytest

pytestmark = [pytest.mark.gpu]


def _run(image):
    import vo_b200  # noqa: F401
    from vo_b200.sift_frontend import SiftExtractor
    sift = SiftExtractor(*image.shape[:2])
    kp, desc, aux = (t.cpu().numpy() for t in sift.extract(image))
    sift.close()
    return {"pt": kp, "size": aux[:, 0], "angle": aux[:, 1], "desc": desc}


def test_sift_matches_oracle_and_reference_plugin(golden):
    from oracle import sift_frontend as sf
    from test_oracle_sift import _check
    g = golden("sift_golden.npz")
    got = _run(g["image"])
    assert _check(sf.detect_and_compute(sf.bgr_to_gray(g["image"])), got) >= 480
    assert _check({"pt": g["kp"].astype(np.float32), "size": g["size"], "angle": g["angle"], "desc": g["desc"]}, got) >= 480


def _kitti_like(seed=8214, h=376, w=1241):
    """Smooth random texture + noise (no flat regions: exact ties in the DoG stack would make the keypoint set depend on
    the last bit of the blur)."""
    rng = np.random.default_rng(seed)
    coarse = rng.integers(0, 256, (h // 8 + 2, w // 8 + 2)).astype(np.float32)
    ys, xs = np.arange(h) / 8.0, np.arange(w) / 8.0
    y0, x0 = ys.astype(int), xs.astype(int)
    fy, fx = (ys - y0)[:, None], (xs - x0)[None, :]
    img = (coarse[y0][:, x0] * (1 - fy) * (1 - fx) + coarse[y0 + 1][:, x0] * fy * (1 - fx) +
           coarse[y0][:, x0 + 1] * (1 - fy) * fx + coarse[y0 + 1][:, x0 + 1] * fy * fx)
    return np.ascontiguousarray(np.clip(img + rng.integers(0, 20, img.shape), 0, 255).astype(np.uint8))


def test_sift_kitti_shaped_and_flat():
    from oracle import sift_frontend as sf
    from test_oracle_sift import _check
    img = _kitti_like()
    assert _check(sf.detect_and_compute(img), _run(img)) > 500
    assert len(_run(np.full((120, 200), 77, np.uint8))["size"]) == 0
