"""CPU restatement of the reference's ORB front-end — TEST INFRASTRUCTURE (oracle), never on the product path.

Reference: feature_extractors/ORB.py:8 (`orb = cv2.ORB_create()`, all defaults) and :10-21
(`extract_features_and_desc`: BGR -> gray, `orb.detectAndCompute(image, None)`, keypoints as an (N, 2) array of
`pt`).  The arithmetic lives in OpenCV (pinned opencv-python==4.5.4.60 in the reference's requirements.txt:4-5; this
image has 4.13.0), whose sources are not under /root/reference; this file restates the published algorithm
(features2d/src/orb.cpp, fast.cpp, fast_score.cpp; imgproc resize INTER_LINEAR_EXACT, separable float filter) and is
PINNED against `cv2.ORB_create().detectAndCompute` run here: identical keypoint sets (level, pixel), bit-identical
angles, Harris responses, `pt` and 256-bit descriptors (tests/test_oracle_orb.py, tests/golden/orb_golden.npz).

Pipeline (defaults: 500 features, scale 1.2f, 8 levels, edge 31, patch 31, FAST threshold 20, Harris score):
  1. gray = (3735 B + 19235 G + 9798 R + 2^14) >> 15                                   (cvtColor BGR2GRAY, 8U)
  2. pyramid: level l has size cvRound(W / s_l), s_l = float(pow(double(1.2f), l)); each level is resized from the
     PREVIOUS one with INTER_LINEAR_EXACT (8.8 fixed-point coefficients, 16.16 accumulation), reflect-101 border of 32
  3. per level: FAST-9/16 (threshold 20) with 3x3 non-maximum suppression on the corner score; drop points closer
     than 31 px to the border; keep the 2 n_l best by FAST score (+ ties); Harris response (7x7 block of Sobel-like
     integer gradients, k = 0.04, in fp32); keep the n_l best by Harris response (+ ties)
  4. orientation: intensity centroid over the circular patch of radius 15, angle = cv::fastAtan2(m01, m10) (degrees)
  5. every level is smoothed by the 7x7, sigma 2 Gaussian as OpenCV's *float* separable filter computes it there
     (row pass: fp32 FMA chain left to right; column pass: symmetric pairs added first; round half even)
  6. rBRIEF: the 256 learned point pairs (bit_pattern_31_) rotated by the angle in fp32, cvRound of each coordinate,
     bit = I(p0) < I(p1)

Host dependence: every stage is integer / explicitly rounded except the Gaussian, which follows the FMA-chained row pass
of the AVX2 object of OpenCV's filter code (the path every AVX2-capable x86 host takes; an SSE-only host would differ in
rare pixels by one grey level).

Keypoint ORDER: OpenCV's depends on the internals of std::nth_element inside KeyPointsFilter::retainBest; only the SET
is defined by the algorithm.  This oracle emits level-major, then row-major order; comparisons are made as sets.
"""
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PATTERN_PATH = os.path.join(os.path.dirname(_HERE), "tests", "golden", "orb_pattern.npy")

# FAST ring of radius 3, (dx, dy), clockwise from (0, 3) as in fast.cpp makeOffsets
_RING = [(0, 3), (1, 3), (2, 2), (3, 1), (3, 0), (3, -1), (2, -2), (1, -3), (0, -3), (-1, -3), (-2, -2), (-3, -1),
         (-3, 0), (-3, 1), (-2, 2), (-1, 3)]
_BORDER = 32   # max(edgeThreshold, ceil(15 sqrt 2), HARRIS_BLOCK_SIZE / 2) + 1
_F = np.float32


def pattern():
    """(512, 2) int32: bit_pattern_31_ of orb.cpp as (x, y) points; pair i = points 2i, 2i + 1."""
    return np.load(_PATTERN_PATH).astype(np.int32)


def bgr_to_gray(img):
    """cv2.cvtColor(img, COLOR_BGR2GRAY) for 8-bit input: 15-bit fixed-point weights, round to nearest."""
    b, g, r = (img[..., i].astype(np.int64) for i in range(3))
    return ((b * 3735 + g * 19235 + r * 9798 + (1 << 14)) >> 15).astype(np.uint8)


def _linear_exact_coeffs(dst, src):
    """Source offset, 8.8 fixed-point weight of the right neighbour and edge flag per destination index
    (resize.cpp interpolationLinear<ufixedpoint16>::getCoeffs)."""
    scale = np.float64(1.0) / (np.float64(dst) / np.float64(src))
    x = np.arange(dst, dtype=np.float64)
    fval = scale * (x + 0.5) - 0.5
    iv = np.floor(fval).astype(np.int64)
    inside = (iv >= 0) & (iv < src - 1) & (src > 1)
    left = ~((iv >= 0) & (src > 1))
    ofs = np.where(inside, iv, np.where(left, 0, src - 1))
    c1 = np.where(inside, np.rint((fval - iv) * 256.0), 0).astype(np.int64)
    return ofs, c1, inside


def resize_linear_exact(img, dw, dh):
    """cv2.resize(img, (dw, dh), interpolation=INTER_LINEAR_EXACT) for one 8-bit channel."""
    H, W = img.shape
    ox, cx, inx = _linear_exact_coeffs(dw, W)
    oy, cy, iny = _linear_exact_coeffs(dh, H)
    p = img.astype(np.int64)
    ox1, oy1 = np.minimum(ox + 1, W - 1), np.minimum(oy + 1, H - 1)
    h = np.where(inx[None, :], (256 - cx)[None, :] * p[:, ox] + cx[None, :] * p[:, ox1], p[:, ox] * 256)
    v = np.where(iny[:, None], (256 - cy)[:, None] * h[oy, :] + cy[:, None] * h[oy1, :], h[oy, :] * 256)
    return ((v + (1 << 15)) >> 16).astype(np.uint8)


def fast_scores(img, thr=20):
    """Corner-score map of FAST-9/16 (0 = no corner): a pixel is a corner iff 9 contiguous ring pixels are all
    brighter than v + thr or all darker than v - thr; its score is the largest threshold for which it still is one
    (fast_score.cpp cornerScore<16>: max over the 16 arcs of the arc's minimum |difference|, minus 1)."""
    H, W = img.shape
    out = np.zeros((H, W), np.int32)
    if H < 7 or W < 7:
        return out
    p = img.astype(np.int16)
    c = p[3:H - 3, 3:W - 3]
    d = np.stack([c - p[3 + dy:H - 3 + dy, 3 + dx:W - 3 + dx] for dx, dy in _RING], 0)
    d2 = np.concatenate([d, d[:8]], 0)
    best = np.full(c.shape, -32768, np.int16)
    for k in range(16):
        seg = d2[k:k + 9]
        best = np.maximum(best, np.maximum(seg.min(0), (-seg).min(0)))
    out[3:H - 3, 3:W - 3] = np.where(best > thr, best - 1, 0)
    return out


def fast_detect(img, thr=20):
    """cv2.FastFeatureDetector_create(thr, True).detect: (x, y, score), row-major; strict 3x3 maximum of the score."""
    s = fast_scores(img, thr)
    H, W = s.shape
    pad = np.pad(s, 1)
    keep = s > 0
    for dy in (-1, 0, 1):
        for dx in (-1, 0, 1):
            if dx or dy:
                keep &= s > pad[1 + dy:1 + dy + H, 1 + dx:1 + dx + W]
    ys, xs = np.nonzero(keep)
    return xs, ys, s[ys, xs]


def retain_best(resp, n):
    """Indices kept by KeyPointsFilter::retainBest: the n largest responses plus every tie with the n-th."""
    m = len(resp)
    if n >= m:
        return np.arange(m)
    if n <= 0:
        return np.arange(0)
    thr = np.sort(resp)[::-1][n - 1]
    return np.nonzero(resp >= thr)[0]


def umax_table(half=15):
    """Row half-widths of the circular patch (orb.cpp computeKeyPoints)."""
    umax = np.zeros(half + 2, np.int64)
    vmax = int(np.floor(_F(half) * np.sqrt(_F(2.0)) / 2 + 1))
    vmin = int(np.ceil(_F(half) * np.sqrt(_F(2.0)) / 2))
    for v in range(vmax + 1):
        umax[v] = int(np.rint(np.sqrt(float(half * half - v * v))))
    v0 = 0
    for v in range(half, vmin - 1, -1):
        while umax[v0] == umax[v0 + 1]:
            v0 += 1
        umax[v] = v0
        v0 += 1
    return umax


_P1 = _F(0.9997878412794807) * _F(180 / np.pi)
_P3 = _F(-0.3258083974640975) * _F(180 / np.pi)
_P5 = _F(0.1555786518463281) * _F(180 / np.pi)
_P7 = _F(-0.04432655554792128) * _F(180 / np.pi)


def fast_atan2(y, x):
    """cv::fastAtan2 (mathfuncs_core): 7th-order odd polynomial in fp32, degrees in [0, 360)."""
    y, x = _F(y), _F(x)
    ax, ay, eps = np.abs(x), np.abs(y), _F(2.220446049250313e-16)
    if ax >= ay:
        c = ay / (ax + eps)
        c2 = c * c
        a = (((_P7 * c2 + _P5) * c2 + _P3) * c2 + _P1) * c
    else:
        c = ax / (ay + eps)
        c2 = c * c
        a = _F(90.0) - (((_P7 * c2 + _P5) * c2 + _P3) * c2 + _P1) * c
    if x < 0:
        a = _F(180.0) - a
    if y < 0:
        a = _F(360.0) - a
    return _F(a)


def gaussian_kernel_7_2():
    """cv2.getGaussianKernel(7, 2, CV_32F): exp(-x^2 / 8) normalised in double, rounded to fp32."""
    x = np.arange(-3, 4, dtype=np.float64)
    k = np.exp(-(x * x) / 8.0)
    return (k / k.sum()).astype(np.float32)


def _fma(a, b, c):
    return (a.astype(np.float64) * np.float64(b) + c.astype(np.float64)).astype(np.float32)


def blur_7x7(img):
    """The smoothing ORB applies to every pyramid level before sampling descriptors: GaussianBlur(7x7, sigma 2,
    reflect-101) as OpenCV's float separable filter evaluates it (pinned against cv2.sepFilter2D): row pass
    s = k0 p0, then s = fma(p_i, k_i, s) left to right; column pass s = k3 h3 + k4 (h4 + h2) + k5 (h5 + h1) + k6 (h6 + h0);
    result rounded half-to-even and saturated."""
    k = gaussian_kernel_7_2()
    H, W = img.shape
    p = np.pad(img, 3, mode="reflect").astype(np.float32)
    s = (p[:, 0:W] * k[0]).astype(np.float32)
    for i in range(1, 7):
        s = _fma(p[:, i:i + W], k[i], s)
    h = s
    s = (h[3:3 + H] * k[3]).astype(np.float32)
    for j in (1, 2, 3):
        t = (h[3 + j:3 + j + H] + h[3 - j:3 - j + H]).astype(np.float32)
        s = (s + (t * k[3 + j]).astype(np.float32)).astype(np.float32)
    return np.clip(np.rint(s), 0, 255).astype(np.uint8)


def harris_response(ext, X, Y, block=7, k=0.04):
    """orb.cpp HarrisResponses at (X, Y) of the bordered level `ext`: integer gradient sums, fp32 formula."""
    r = block // 2
    blk = ext[Y - r - 1:Y + r + 2, X - r - 1:X + r + 2].astype(np.int64)
    ix = (blk[1:-1, 2:] - blk[1:-1, :-2]) * 2 + (blk[:-2, 2:] - blk[:-2, :-2]) + (blk[2:, 2:] - blk[2:, :-2])
    iy = (blk[2:, 1:-1] - blk[:-2, 1:-1]) * 2 + (blk[2:, :-2] - blk[:-2, :-2]) + (blk[2:, 2:] - blk[:-2, 2:])
    a, b, c = _F(int((ix * ix).sum())), _F(int((iy * iy).sum())), _F(int((ix * iy).sum()))
    scale = _F(1.0) / _F((1 << 2) * block * _F(255.0))
    return (a * b - c * c - _F(k) * (a + b) * (a + b)) * (scale * scale * scale * scale)


def ic_angle(ext, X, Y, umax, half=15):
    """orb.cpp ICAngles: intensity-centroid orientation of the circular patch."""
    m01 = 0
    us = np.arange(-half, half + 1)
    m10 = int((us * ext[Y, X - half:X + half + 1].astype(np.int64)).sum())
    for v in range(1, half + 1):
        d = int(umax[v])
        rp = ext[Y + v, X - d:X + d + 1].astype(np.int64)
        rm = ext[Y - v, X - d:X + d + 1].astype(np.int64)
        m01 += v * int((rp - rm).sum())
        m10 += int((np.arange(-d, d + 1) * (rp + rm)).sum())
    return fast_atan2(m01, m10)


def rotated_pattern(angle_deg, pat):
    """Integer sample offsets (ix, iy) of the 512 pattern points for a keypoint angle (orb.cpp GET_VALUE)."""
    a32 = _F(angle_deg) * _F(np.pi / 180.0)
    ca, sb = _F(np.cos(np.float64(a32))), _F(np.sin(np.float64(a32)))
    px, py = pat[:, 0].astype(np.float32), pat[:, 1].astype(np.float32)
    return np.rint(px * ca - py * sb).astype(np.int64), np.rint(px * sb + py * ca).astype(np.int64)


def level_scales(scale_factor=1.2, nlevels=8):
    sf = np.float64(np.float32(scale_factor))            # ORB_create takes a float
    return [np.float32(np.power(sf, np.float64(l))) for l in range(nlevels)]


def features_per_level(nfeatures=500, scale_factor=1.2, nlevels=8):
    factor = np.float32(1.0 / np.float64(np.float32(scale_factor)))
    nd = _F(nfeatures) * (_F(1) - factor) / (_F(1) - _F(np.power(np.float64(factor), np.float64(nlevels))))
    per, s = [], 0
    for _ in range(nlevels - 1):
        per.append(int(np.rint(nd)))
        s += per[-1]
        nd = _F(nd * factor)
    per.append(max(nfeatures - s, 0))
    return per


def build_pyramid(gray, scale_factor=1.2, nlevels=8):
    """Bordered levels (reflect-101, 32 px) and their scales."""
    H, W = gray.shape
    scales = level_scales(scale_factor, nlevels)
    levels, cur = [], gray
    for l in range(nlevels):
        if l > 0:
            inv = _F(1.0) / scales[l]
            dw, dh = int(np.rint(_F(W) * inv)), int(np.rint(_F(H) * inv))
            cur = resize_linear_exact(cur, dw, dh)
        levels.append(np.pad(cur, _BORDER, mode="reflect"))
    return levels, scales


def detect_and_compute(gray, nfeatures=500, scale_factor=1.2, nlevels=8, edge=31, patch=31, fast_thr=20):
    """cv2.ORB_create(...).detectAndCompute(gray, None).  Returns a dict of arrays, one row per keypoint:
    level, xl, yl (pixel in its level), pt (N, 2) fp32 in level-0 coordinates, size, angle, response (fp32), desc (N, 32) u8."""
    pat = pattern()
    levels, scales = build_pyramid(gray, scale_factor, nlevels)
    per = features_per_level(nfeatures, scale_factor, nlevels)
    umax = umax_table(patch // 2)
    rows = []
    for l in range(nlevels):
        ext = levels[l]
        img = ext[_BORDER:-_BORDER, _BORDER:-_BORDER]
        h, w = img.shape
        xs, ys, sc = fast_detect(img, fast_thr)
        m = (xs >= edge) & (xs < w - edge) & (ys >= edge) & (ys < h - edge)
        xs, ys, sc = xs[m], ys[m], sc[m].astype(np.float32)
        keep = retain_best(sc, 2 * per[l])
        xs, ys = xs[keep], ys[keep]
        resp = np.array([harris_response(ext, x + _BORDER, y + _BORDER) for x, y in zip(xs, ys)], np.float32)
        keep = retain_best(resp, per[l])
        for x, y, r in zip(xs[keep], ys[keep], resp[keep]):
            rows.append((l, int(x), int(y), _F(r), ic_angle(ext, x + _BORDER, y + _BORDER, umax, patch // 2)))
    smooth = [None] * nlevels
    desc = np.zeros((len(rows), 32), np.uint8)
    for i, (l, x, y, r, ang) in enumerate(rows):
        if smooth[l] is None:
            smooth[l] = blur_7x7(levels[l][_BORDER:-_BORDER, _BORDER:-_BORDER])
        ix, iy = rotated_pattern(ang, pat)
        vals = smooth[l][y + iy, x + ix].astype(np.int32)   # >= 31 px from the border: never leaves the level
        bits = (vals[0::2] < vals[1::2]).astype(np.uint8)
        desc[i] = np.packbits(bits.reshape(32, 8)[:, ::-1], axis=1)[:, 0]
    lv = np.array([r[0] for r in rows], np.int32)
    xl = np.array([r[1] for r in rows], np.int32)
    yl = np.array([r[2] for r in rows], np.int32)
    sc = np.array([scales[l] for l in lv], np.float32)
    pt = np.stack([xl.astype(np.float32) * sc, yl.astype(np.float32) * sc], 1) if len(rows) else np.zeros((0, 2), np.float32)
    return {"level": lv, "xl": xl, "yl": yl, "pt": pt.astype(np.float32), "size": (_F(patch) * sc).astype(np.float32),
            "angle": np.array([r[4] for r in rows], np.float32), "response": np.array([r[3] for r in rows], np.float32),
            "desc": desc}


def extract_features_and_desc(image_bgr):
    """feature_extractors/ORB.py:10-21: (kp (N, 2) float64 of pt.x, pt.y; desc (N, 32) uint8)."""
    out = detect_and_compute(bgr_to_gray(image_bgr))
    return out["pt"].astype(np.float64), out["desc"]
