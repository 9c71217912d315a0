"""CPU restatement of the reference's SIFT front-end — TEST INFRASTRUCTURE (oracle), never on the product path.

Reference: feature_extractors/SIFT.py:10 (`cv2.xfeatures2d.SIFT_create()`, all defaults: 3 octave layers, contrast
threshold 0.04, edge threshold 10, sigma 1.6, image doubled) and :14-23 (`extract_features_and_desc`: BGR -> gray,
`sift.detectAndCompute(image, None)`, keypoints as an (N, 2) array of `pt`).  The arithmetic lives in OpenCV (features2d
sift.dispatch.cpp / sift.simd.hpp); this file restates the published algorithm in plain fp32 numpy.

PARITY BAR.  Unlike ORB, SIFT cannot be pinned bit for bit: OpenCV runs AVX2 / AVX-512 objects built with FMA contraction
whenever the host CPU allows, so the low-order bits of the reference itself depend on the host
(tools/probe/sift_detector_probe.py reproduces the detector bit for bit only after modelling that contraction).  This
restatement is therefore pinned against `cv2.SIFT_create().detectAndCompute` to a TOLERANCE (tests/test_oracle_sift.py):
the same keypoints (position within 1e-3 px, size within 1e-3, angle within 0.05 degrees for all but a handful of
decision-boundary cases) and descriptor entries within +-1 of OpenCV's 0..255 values.

Pipeline: gray -> fp32; x2 bilinear up-sampling; Gaussian scale space (sigma 1.6, 3 layers + 3 per octave, every octave
seeded by 2:1 decimation of layer 3); difference of Gaussians; 26-neighbour extrema above floor(0.5 * 0.04 / 3 * 255);
up to 5 Newton steps of the 3-D quadratic fit (adjustLocalExtrema), contrast and edge tests; 36-bin gradient-orientation
histogram (radius 4.5 scl, Gaussian weight 1.5 scl, fastAtan2), smoothed, one keypoint per peak >= 0.8 max with parabolic
refinement; halve the coordinates (first octave is -1); sort and drop duplicates; 4 x 4 x 8 descriptor with trilinear
interpolation, 0.2 clipping, scaled by 512 and saturated to 0..255 (stored as fp32).
"""
import numpy as np

from oracle.orb_frontend import bgr_to_gray

F = np.float32
N_LAYERS, SIGMA, CONTRAST, EDGE = 3, 1.6, 0.04, 10.0
IMG_BORDER, MAX_STEPS, ORI_BINS = 5, 5, 36


def gaussian_kernel(sigma):
    """cv2.getGaussianKernel(ksize, sigma, CV_32F) with the kernel size GaussianBlur derives for fp32 images."""
    ksize = int(np.rint(sigma * 4 * 2 + 1)) | 1
    x = np.arange(ksize, dtype=np.float64) - (ksize - 1) * 0.5
    k = np.exp(-(x * x) / (2.0 * sigma * sigma))
    return (k / k.sum()).astype(np.float32)


def gaussian_blur(img, sigma):
    """cv2.GaussianBlur(img, (0, 0), sigma) on an fp32 image, reflect-101 border (to fp32 rounding)."""
    k = gaussian_kernel(sigma)
    r = len(k) // 2
    H, W = img.shape
    p = np.pad(img, ((0, 0), (r, r)), mode="reflect")
    h = np.zeros((H, W), np.float32)
    for i in range(len(k)):
        h += k[i] * p[:, i:i + W]
    p = np.pad(h, ((r, r), (0, 0)), mode="reflect")
    v = np.zeros((H, W), np.float32)
    for i in range(len(k)):
        v += k[i] * p[i:i + H, :]
    return v


def upsample2(img):
    """cv2.resize(img, (2W, 2H), interpolation=INTER_LINEAR) on an fp32 image."""
    def axis(n):
        f = (np.arange(2 * n, dtype=np.float32) + F(0.5)) * F(0.5) - F(0.5)
        i0 = np.floor(f).astype(np.int64)
        w = (f - i0).astype(np.float32)
        w[i0 < 0] = 0
        i0 = np.clip(i0, 0, n - 1)
        i1 = np.clip(i0 + 1, 0, n - 1)
        return i0, i1, w
    H, W = img.shape
    x0, x1, wx = axis(W)
    y0, y1, wy = axis(H)
    rows = img[:, x0] * (F(1) - wx)[None, :] + img[:, x1] * wx[None, :]
    return (rows[y0, :] * (F(1) - wy)[:, None] + rows[y1, :] * wy[:, None]).astype(np.float32)


def build_pyramids(gray):
    base = gaussian_blur(upsample2(gray.astype(np.float32)), float(np.sqrt(max(SIGMA * SIGMA - 0.5 * 0.5 * 4, 0.01))))
    n_oct = int(np.rint(np.log(float(min(base.shape))) / np.log(2.0) - 2)) + 1
    k = 2.0 ** (1.0 / N_LAYERS)
    sig = [SIGMA]
    for i in range(1, N_LAYERS + 3):
        prev = (k ** (i - 1)) * SIGMA
        sig.append(float(np.sqrt((prev * k) ** 2 - prev ** 2)))
    gauss = []
    for o in range(n_oct):
        for i in range(N_LAYERS + 3):
            if o == 0 and i == 0:
                gauss.append(base)
            elif i == 0:
                src = gauss[(o - 1) * (N_LAYERS + 3) + N_LAYERS]
                gauss.append(np.ascontiguousarray(src[::2, ::2][:src.shape[0] // 2, :src.shape[1] // 2]))
            else:
                gauss.append(gaussian_blur(gauss[-1], sig[i]))
    dog = [gauss[o * (N_LAYERS + 3) + i + 1] - gauss[o * (N_LAYERS + 3) + i] for o in range(n_oct) for i in range(N_LAYERS + 2)]
    return gauss, dog, n_oct


def _solve3(a, b):
    det = a[0, 0] * (a[1, 1] * a[2, 2] - a[1, 2] * a[2, 1]) - a[0, 1] * (a[1, 0] * a[2, 2] - a[1, 2] * a[2, 0]) \
        + a[0, 2] * (a[1, 0] * a[2, 1] - a[1, 1] * a[2, 0])
    if det == 0:
        return np.zeros(3, np.float32)
    d = F(1) / F(det)
    x0 = d * (b[0] * (a[1, 1] * a[2, 2] - a[1, 2] * a[2, 1]) - a[0, 1] * (b[1] * a[2, 2] - a[1, 2] * b[2]) + a[0, 2] * (b[1] * a[2, 1] - a[1, 1] * b[2]))
    x1 = d * (a[0, 0] * (b[1] * a[2, 2] - a[1, 2] * b[2]) - b[0] * (a[1, 0] * a[2, 2] - a[1, 2] * a[2, 0]) + a[0, 2] * (a[1, 0] * b[2] - b[1] * a[2, 0]))
    x2 = d * (a[0, 0] * (a[1, 1] * b[2] - b[1] * a[2, 1]) - a[0, 1] * (a[1, 0] * b[2] - b[1] * a[2, 0]) + b[0] * (a[1, 0] * a[2, 1] - a[1, 1] * a[2, 0]))
    return np.array([x0, x1, x2], np.float32)


def adjust_local_extrema(dog, octv, layer, r, c):
    """sift.simd.hpp adjustLocalExtrema: returns (pt_x, pt_y, octave word, size, response, r, c, layer) or None."""
    img_scale = F(1.0) / F(255)
    ds, ss, cs = img_scale * F(0.5), img_scale, img_scale * F(0.25)
    xi = xr = xc = F(0)
    for _ in range(MAX_STEPS):
        idx = octv * (N_LAYERS + 2) + layer
        img, prv, nxt = dog[idx], dog[idx - 1], dog[idx + 1]
        dD = np.array([(img[r, c + 1] - img[r, c - 1]) * ds, (img[r + 1, c] - img[r - 1, c]) * ds, (nxt[r, c] - prv[r, c]) * ds], np.float32)
        v2 = img[r, c] * F(2)
        dxx = (img[r, c + 1] + img[r, c - 1] - v2) * ss
        dyy = (img[r + 1, c] + img[r - 1, c] - v2) * ss
        dss = (nxt[r, c] + prv[r, c] - v2) * ss
        dxy = (img[r + 1, c + 1] - img[r + 1, c - 1] - img[r - 1, c + 1] + img[r - 1, c - 1]) * cs
        dxs = (nxt[r, c + 1] - nxt[r, c - 1] - prv[r, c + 1] + prv[r, c - 1]) * cs
        dys = (nxt[r + 1, c] - nxt[r - 1, c] - prv[r + 1, c] + prv[r - 1, c]) * cs
        X = _solve3(np.array([[dxx, dxy, dxs], [dxy, dyy, dys], [dxs, dys, dss]], np.float32), dD)
        xi, xr, xc = -X[2], -X[1], -X[0]
        if abs(xi) < 0.5 and abs(xr) < 0.5 and abs(xc) < 0.5:
            break
        if max(abs(xi), abs(xr), abs(xc)) > 2 ** 31 / 3:
            return None
        c += int(np.rint(xc)); r += int(np.rint(xr)); layer += int(np.rint(xi))
        if layer < 1 or layer > N_LAYERS or c < IMG_BORDER or c >= img.shape[1] - IMG_BORDER or r < IMG_BORDER or r >= img.shape[0] - IMG_BORDER:
            return None
    else:
        return None
    idx = octv * (N_LAYERS + 2) + layer
    img, prv, nxt = dog[idx], dog[idx - 1], dog[idx + 1]
    dD = np.array([(img[r, c + 1] - img[r, c - 1]) * ds, (img[r + 1, c] - img[r - 1, c]) * ds, (nxt[r, c] - prv[r, c]) * ds], np.float32)
    contr = img[r, c] * img_scale + (dD[0] * xc + dD[1] * xr + dD[2] * xi) * F(0.5)
    if abs(contr) * N_LAYERS < F(CONTRAST):
        return None
    v2 = img[r, c] * F(2)
    dxx = (img[r, c + 1] + img[r, c - 1] - v2) * ss
    dyy = (img[r + 1, c] + img[r - 1, c] - v2) * ss
    dxy = (img[r + 1, c + 1] - img[r + 1, c - 1] - img[r - 1, c + 1] + img[r - 1, c - 1]) * cs
    tr, det = dxx + dyy, dxx * dyy - dxy * dxy
    if det <= 0 or tr * tr * F(EDGE) >= (F(EDGE) + 1) * (F(EDGE) + 1) * det:
        return None
    octave = octv + (layer << 8) + (int(np.rint((np.float64(xi) + 0.5) * 255)) << 16)
    size = F(SIGMA) * np.float32(np.power(F(2.0), (F(layer) + xi) / F(N_LAYERS))) * F(1 << octv) * F(2)
    return (F(c) + xc) * F(1 << octv), (F(r) + xr) * F(1 << octv), octave, size, abs(contr), r, c, layer


_P1, _P3 = F(0.9997878412794807) * F(180 / np.pi), F(-0.3258083974640975) * F(180 / np.pi)
_P5, _P7 = F(0.1555786518463281) * F(180 / np.pi), F(-0.04432655554792128) * F(180 / np.pi)


def fast_atan2_deg(y, x):
    """cv::hal::fastAtan2 (degrees) on arrays: the 7th-order polynomial OpenCV uses instead of atan2 (~0.3 degree error)."""
    y, x = y.astype(np.float32), x.astype(np.float32)
    ax, ay = np.abs(x), np.abs(y)
    eps = F(2.220446049250313e-16)
    wide = ax >= ay
    c = np.where(wide, ay / (ax + eps), ax / (ay + eps)).astype(np.float32)
    c2 = c * c
    a = (((_P7 * c2 + _P5) * c2 + _P3) * c2 + _P1) * c
    a = np.where(wide, a, F(90.0) - a)
    a = np.where(x < 0, F(180.0) - a, a)
    a = np.where(y < 0, F(360.0) - a, a)
    return a.astype(np.float32)


def orientation_hist(img, px, py, radius, sigma):
    """calcOrientationHist: smoothed 36-bin histogram and its maximum."""
    n = ORI_BINS
    expf_scale = F(-1.0) / (F(2.0) * F(sigma) * F(sigma))
    ii, jj = np.mgrid[-radius:radius + 1, -radius:radius + 1]
    ys, xs = py + ii, px + jj
    ok = (ys > 0) & (ys < img.shape[0] - 1) & (xs > 0) & (xs < img.shape[1] - 1)
    ys, xs, ii, jj = ys[ok], xs[ok], ii[ok], jj[ok]
    dx = img[ys, xs + 1] - img[ys, xs - 1]
    dy = img[ys - 1, xs] - img[ys + 1, xs]
    w = np.exp(((ii * ii + jj * jj).astype(np.float32) * expf_scale).astype(np.float32)).astype(np.float32)
    ori = fast_atan2_deg(dy, dx)
    mag = np.sqrt(dx * dx + dy * dy).astype(np.float32)
    bins = np.rint(F(n / 360.0) * ori).astype(np.int64)
    bins = np.where(bins >= n, bins - n, bins)
    bins = np.where(bins < 0, bins + n, bins)
    temp = np.zeros(n, np.float32)
    np.add.at(temp, bins, (w * mag).astype(np.float32))
    t = np.concatenate([temp[-2:], temp, temp[:2]])
    hist = (t[0:n] + t[4:n + 4]) * F(1.0 / 16.0) + (t[1:n + 1] + t[3:n + 3]) * F(4.0 / 16.0) + t[2:n + 2] * F(6.0 / 16.0)
    return hist.astype(np.float32), hist.max()


def detect(gray):
    """Keypoints before duplicate removal, in full-resolution (doubled-image) coordinates, plus the Gaussian pyramid."""
    gauss, dog, n_oct = build_pyramids(gray)
    thr = int(np.floor(0.5 * CONTRAST / N_LAYERS * 255))
    out = []
    for o in range(n_oct):
        for i in range(1, N_LAYERS + 1):
            idx = o * (N_LAYERS + 2) + i
            cur, prv, nxt = dog[idx], dog[idx - 1], dog[idx + 1]
            H, W = cur.shape
            if H <= 2 * IMG_BORDER or W <= 2 * IMG_BORDER:
                continue
            core = cur[IMG_BORDER:H - IMG_BORDER, IMG_BORDER:W - IMG_BORDER]
            nb = np.stack([im[IMG_BORDER + dy:H - IMG_BORDER + dy, IMG_BORDER + dx:W - IMG_BORDER + dx]
                           for im in (prv, cur, nxt) for dy in (-1, 0, 1) for dx in (-1, 0, 1)], 0)
            cand = (np.abs(core) > thr) & (((core > 0) & (core >= nb.max(0))) | ((core < 0) & (core <= nb.min(0))))
            for r, c in zip(*np.nonzero(cand)):
                k = adjust_local_extrema(dog, o, i, int(r) + IMG_BORDER, int(c) + IMG_BORDER)
                if k is None:
                    continue
                x, y, octave, size, resp, r1, c1, layer = k
                scl_octv = size * F(0.5) / F(1 << o)
                hist, omax = orientation_hist(gauss[o * (N_LAYERS + 3) + layer], c1, r1, int(np.rint(F(4.5) * scl_octv)), F(1.5) * scl_octv)
                mag_thr = F(omax * F(0.8))
                n = ORI_BINS
                for j in range(n):
                    l, r2 = (j - 1) % n, (j + 1) % n
                    if hist[j] > hist[l] and hist[j] > hist[r2] and hist[j] >= mag_thr:
                        b = F(j) + F(0.5) * (hist[l] - hist[r2]) / (hist[l] - F(2) * hist[j] + hist[r2])
                        b = b + n if b < 0 else (b - n if b >= n else b)
                        ang = F(360.0) - F(F(360.0 / n) * b)
                        if abs(ang - F(360.0)) < np.finfo(np.float32).eps:
                            ang = F(0)
                        out.append((x, y, size, ang, resp, octave))
    return out, gauss


def finalize(kps):
    """First octave is -1: halve coordinates and size, rewrite the octave byte; then KeyPointsFilter::removeDuplicatedSorted
    (sort by x, y, size, angle, response, octave — the last four descending — and drop repeats of (x, y, size, angle))."""
    rows = []
    for x, y, size, ang, resp, octave in kps:
        oc = (octave & ~255) | ((octave - 1) & 255)
        rows.append((F(x * F(0.5)), F(y * F(0.5)), F(size * F(0.5)), F(ang), F(resp), oc))
    rows.sort(key=lambda k: (k[0], k[1], -k[2], -k[3], -k[4], -k[5]))
    out = []
    for k in rows:
        if not out or (k[0], k[1], k[2], k[3]) != (out[-1][0], out[-1][1], out[-1][2], out[-1][3]):
            out.append(k)
    return out


def descriptor(img, ptx, pty, ori, scl, d=4, n=8):
    """calcSIFTDescriptor: 128 values in 0..255 (fp32)."""
    px, py = int(np.rint(ptx)), int(np.rint(pty))
    cos_t = F(np.cos(np.float64(F(ori) * F(np.pi / 180))))
    sin_t = F(np.sin(np.float64(F(ori) * F(np.pi / 180))))
    bins_per_rad = F(n / 360.0)
    exp_scale = F(-1.0) / F(d * d * 0.5)
    hist_width = F(3.0) * F(scl)
    radius = int(np.rint(hist_width * F(1.4142135623730951) * F(d + 1) * F(0.5)))
    radius = min(radius, int(np.sqrt(float(img.shape[1]) ** 2 + float(img.shape[0]) ** 2)))
    cos_t, sin_t = cos_t / hist_width, sin_t / hist_width
    ii, jj = np.mgrid[-radius:radius + 1, -radius:radius + 1]
    fi, fj = ii.astype(np.float32), jj.astype(np.float32)
    c_rot = fj * cos_t - fi * sin_t
    r_rot = fj * sin_t + fi * cos_t
    rbin = r_rot + F(d // 2) - F(0.5)
    cbin = c_rot + F(d // 2) - F(0.5)
    rr, cc = py + ii, px + jj
    ok = (rbin > -1) & (rbin < d) & (cbin > -1) & (cbin < d) & (rr > 0) & (rr < img.shape[0] - 1) & (cc > 0) & (cc < img.shape[1] - 1)
    rr, cc, rbin, cbin, c_rot, r_rot = rr[ok], cc[ok], rbin[ok], cbin[ok], c_rot[ok], r_rot[ok]
    dx = img[rr, cc + 1] - img[rr, cc - 1]
    dy = img[rr - 1, cc] - img[rr + 1, cc]
    w = np.exp(((c_rot * c_rot + r_rot * r_rot) * exp_scale).astype(np.float32)).astype(np.float32)
    o = fast_atan2_deg(dy, dx)
    mag = (np.sqrt(dx * dx + dy * dy).astype(np.float32) * w).astype(np.float32)
    obin = ((o - F(ori)) * bins_per_rad).astype(np.float32)
    r0, c0, o0 = np.floor(rbin).astype(np.int64), np.floor(cbin).astype(np.int64), np.floor(obin).astype(np.int64)
    rb, cb, ob = rbin - r0, cbin - c0, obin - o0
    o0 = np.where(o0 < 0, o0 + n, o0)
    o0 = np.where(o0 >= n, o0 - n, o0)
    hist = np.zeros(((d + 2), (d + 2), (n + 2)), np.float32)
    v_r1 = mag * rb; v_r0 = mag - v_r1
    v11 = v_r1 * cb; v10 = v_r1 - v11
    v01 = v_r0 * cb; v00 = v_r0 - v01
    for (dr, dc, v) in ((0, 0, v00), (0, 1, v01), (1, 0, v10), (1, 1, v11)):
        v1 = v * ob
        np.add.at(hist, (r0 + 1 + dr, c0 + 1 + dc, o0), (v - v1).astype(np.float32))
        np.add.at(hist, (r0 + 1 + dr, c0 + 1 + dc, o0 + 1), v1.astype(np.float32))
    core = hist[1:d + 1, 1:d + 1, :].copy()
    core[:, :, 0] += core[:, :, n]
    core[:, :, 1] += core[:, :, n + 1]
    raw = core[:, :, :n].reshape(-1).astype(np.float32)
    thr = F(np.sqrt(F((raw * raw).sum(dtype=np.float32)))) * F(0.2)
    raw = np.minimum(raw, thr)
    nrm = F(512.0) / max(F(np.sqrt(F((raw * raw).sum(dtype=np.float32)))), np.finfo(np.float32).eps)
    return np.clip(np.rint(raw * nrm), 0, 255).astype(np.float32)


def detect_and_compute(gray):
    """cv2.SIFT_create().detectAndCompute(gray, None) -> dict(pt (N,2), size, angle, response, octave, desc (N,128))."""
    raw, gauss = detect(gray)
    kps = finalize(raw)
    desc = np.zeros((len(kps), 128), np.float32)
    for i, (x, y, size, ang, resp, oc) in enumerate(kps):
        octave, layer = oc & 255, (oc >> 8) & 255
        octave = octave if octave < 128 else (-128 | octave)
        scale = F(1.0) / F(1 << octave) if octave >= 0 else F(1 << -octave)
        a = F(360.0) - ang
        if abs(a - F(360.0)) < np.finfo(np.float32).eps:
            a = F(0)
        desc[i] = descriptor(gauss[(octave + 1) * (N_LAYERS + 3) + layer], x * scale, y * scale, a, size * scale * F(0.5))
    arr = lambda j, t: np.array([k[j] for k in kps], t)  # noqa: E731
    return {"pt": np.stack([arr(0, np.float32), arr(1, np.float32)], 1) if kps else np.zeros((0, 2), np.float32),
            "size": arr(2, np.float32), "angle": arr(3, np.float32), "response": arr(4, np.float32), "octave": arr(5, np.int64),
            "desc": desc}


def extract_features_and_desc(image_bgr):
    """feature_extractors/SIFT.py:14-23: (kp (N, 2) float64 of pt.x, pt.y; desc (N, 128) float32)."""
    out = detect_and_compute(bgr_to_gray(image_bgr))
    return out["pt"].astype(np.float64), out["desc"]
