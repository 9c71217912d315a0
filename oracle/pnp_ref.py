"""Reference-sampler PnP-RANSAC ("Mode R") on the CPU — TEST INFRASTRUCTURE ONLY (see oracle.py header).

What VisualOdometry.computepose_3D_2D does (VisualOdometry_Stereo.py:120-135): three bootstrap resamples (np.random.randint,
:122), each through cv2.solvePnPRansac(iterationsCount=100, reprojectionError=1.5), best of three by inlier count (> 20).
`ransac_replica` restates the inside of that OpenCV call from its parts — the multiply-with-carry sample
table, a five-point minimal solve per row, projectPoints + fp32 squared error, the adaptive iteration count, the refit on
the best minimal model's inliers — with the minimal solver as a parameter:
  solver="cv"     cv2.solvePnP(SOLVEPNP_EPNP) per row: reproduces cv2.solvePnPRansac EXACTLY (tests/test_oracle_pnp_ref.py),
                  which pins table, scoring, stopping rule and refit to the reference's own call;
  solver="epnp"   `epnp5` below: the published EPnP algorithm (Lepetit et al. 2009; OpenCV calib3d epnp.cpp) in numpy, with
                  LAPACK eigen-solvers, a canonical basis for the two-dimensional null space of M^T M (five points) and
                  Horn's absolute orientation — the restatement the CUDA kernels (csrc/pnp_ref.cu) are compared with.
OpenCV's own five-point models cannot be reproduced by any restatement: they depend on the arbitrary null-space basis its
SVD returns (tools/probe/epnp_basis_probe.py), so solver="epnp" agrees with solver="cv" statistically, not per hypothesis."""
import numpy as np

F32 = np.float32
MODEL_POINTS = 5
_tables = {}


def mwc_table(n, iters=100, m=MODEL_POINTS):
    """cv::RNG((uint64)-1) re-created on every solvePnPRansac call; `m` distinct indices per iteration (ptsetreg.cpp getSubset)."""
    key = (int(n), int(iters), int(m))
    if key not in _tables:
        state = 0xFFFFFFFFFFFFFFFF
        out = np.zeros((iters, m), np.int32)
        for it in range(iters):
            s = []
            while len(s) < m:
                state = ((state & 0xFFFFFFFF) * 4164903690 + (state >> 32)) & 0xFFFFFFFFFFFFFFFF
                v = (state & 0xFFFFFFFF) % n
                if v not in s:
                    s.append(v)
            out[it] = s
        _tables[key] = out
    return _tables[key]


def _fix_sign(v):
    m = int(np.argmax(np.abs(v)))
    return -v if v[m] < 0 else v


def _eigh_desc(A):
    w, v = np.linalg.eigh(A)
    order = np.argsort(-w, kind="stable")
    return w[order], v[:, order]


def _canonical_null_basis(a, b):
    p = a * a + b * b
    k = int(np.argmax(p))
    nrm = np.hypot(a[k], b[k])
    w1 = (a[k] * a + b[k] * b) / nrm
    w0 = _fix_sign((-b[k] * a + a[k] * b) / nrm)
    return w1, w0


def epnp5(X, uv, K):
    """EPnP on five points.  X (5,3) float32, uv (5,2) float32, K 3x3.  Returns (R 3x3, t 3) in double, or None."""
    n = len(X)
    fu, fv, uc, vc = float(K[0, 0]), float(K[1, 1]), float(K[0, 2]), float(K[1, 2])
    uvd = np.asarray(uv, F32).astype(np.float64)
    xn = ((uvd[:, 0] - uc) * (1.0 / fu)).astype(F32).astype(np.float64)          # undistortPoints -> float32 normalised coords
    yn = ((uvd[:, 1] - vc) * (1.0 / fv)).astype(F32).astype(np.float64)
    us = np.stack([xn * fu + uc, yn * fv + vc], 1)
    pws = np.asarray(X, F32).astype(np.float64)
    cws = np.zeros((4, 3))
    cws[0] = pws.sum(0) / n
    d0 = pws - cws[0]
    dc, uc_ = _eigh_desc(d0.T @ d0)
    axes = np.stack([_fix_sign(uc_[:, i]) for i in range(3)])         # principal directions (rows), orthonormal
    ks = np.sqrt(np.maximum(dc, 0.0) / n)
    for i in range(1, 4):
        cws[i] = cws[0] + ks[i - 1] * axes[i - 1]
    # barycentric coordinates: the control vectors are k_i * (orthonormal axis i), so the inverse of [c1-c0 c2-c0 c3-c0] is
    # axis_i / k_i row by row; a vanishing k_i (planar or collinear points) gives a zero row — the pseudo-inverse
    # cvInvert(CV_SVD) returns in epnp.cpp (singular values <= 2 eps * sum are dropped, SVD::backSubst)
    inv_k = np.where(ks > 2 * np.finfo(np.float64).eps * ks.sum(), 1.0 / np.where(ks > 0, ks, 1.0), 0.0)
    al = np.zeros((n, 4))
    al[:, 1:] = (d0 @ axes.T) * inv_k
    al[:, 0] = 1.0 - al[:, 1] - al[:, 2] - al[:, 3]
    M = np.zeros((2 * n, 12))
    for j in range(4):
        M[0::2, 3 * j] = al[:, j] * fu
        M[0::2, 3 * j + 2] = al[:, j] * (uc - us[:, 0])
        M[1::2, 3 * j + 1] = al[:, j] * fv
        M[1::2, 3 * j + 2] = al[:, j] * (vc - us[:, 1])
    MtM = M.T @ M
    if inv_k[2] == 0.0 and inv_k[1] != 0.0:
        # coplanar points: the fourth control point coincides with the centroid, its three columns of M vanish and e9, e10, e11
        # span an exactly-null eigenspace (OpenCV gets an arbitrary basis of it from its SVD).  Canonical choice: those unit
        # vectors as v[0..2], and the weakest direction of the 9 x 9 block of the three real control points as v[3]
        _, V9 = _eigh_desc(MtM[:9, :9])
        x = np.zeros(12)
        x[:9] = V9[:, 8]
        v = [np.eye(12)[11], np.eye(12)[10], np.eye(12)[9], _fix_sign(x)]
    else:
        _, V = _eigh_desc(MtM)
        v0, v1 = _canonical_null_basis(V[:, 11], V[:, 10])
        v = [v0, v1, _fix_sign(V[:, 9]), _fix_sign(V[:, 8])]
    dv = np.zeros((4, 6, 3))
    pairs = [(0, 1), (0, 2), (0, 3), (1, 2), (1, 3), (2, 3)]
    for i in range(4):
        for j, (a, b) in enumerate(pairs):
            dv[i, j] = v[i][3 * a:3 * a + 3] - v[i][3 * b:3 * b + 3]
    L = np.zeros((6, 10))
    for i in range(6):
        d = dv[:, i]
        L[i] = [d[0] @ d[0], 2 * d[0] @ d[1], d[1] @ d[1], 2 * d[0] @ d[2], 2 * d[1] @ d[2], d[2] @ d[2], 2 * d[0] @ d[3],
                2 * d[1] @ d[3], 2 * d[2] @ d[3], d[3] @ d[3]]
    rho = np.array([((cws[a] - cws[b]) ** 2).sum() for a, b in pairs])

    def lsq(A, b):
        return np.linalg.lstsq(A, b, rcond=None)[0]

    def gauss_newton(be):
        be = np.array(be, np.float64)
        for _ in range(5):
            A = np.stack([2 * L[:, 0] * be[0] + L[:, 1] * be[1] + L[:, 3] * be[2] + L[:, 6] * be[3],
                          L[:, 1] * be[0] + 2 * L[:, 2] * be[1] + L[:, 4] * be[2] + L[:, 7] * be[3],
                          L[:, 3] * be[0] + L[:, 4] * be[1] + 2 * L[:, 5] * be[2] + L[:, 8] * be[3],
                          L[:, 6] * be[0] + L[:, 7] * be[1] + L[:, 8] * be[2] + 2 * L[:, 9] * be[3]], 1)
            b = rho - (L[:, 0] * be[0] ** 2 + L[:, 1] * be[0] * be[1] + L[:, 2] * be[1] ** 2 + L[:, 3] * be[0] * be[2] +
                       L[:, 4] * be[1] * be[2] + L[:, 5] * be[2] ** 2 + L[:, 6] * be[0] * be[3] + L[:, 7] * be[1] * be[3] +
                       L[:, 8] * be[2] * be[3] + L[:, 9] * be[3] ** 2)
            be = be + lsq(A, b)
        return be

    def r_and_t(be):
        ccs = sum(be[i] * v[i].reshape(4, 3) for i in range(4))
        pcs = al @ ccs
        if pcs[0, 2] < 0:
            pcs = -pcs
        pc0, pw0 = pcs.mean(0), pws.mean(0)
        S = (pws - pw0).T @ (pcs - pc0)                       # S[a][b] = sum pw_a pc_b
        Nq = np.array([[S[0, 0] + S[1, 1] + S[2, 2], S[1, 2] - S[2, 1], S[2, 0] - S[0, 2], S[0, 1] - S[1, 0]],
                       [S[1, 2] - S[2, 1], S[0, 0] - S[1, 1] - S[2, 2], S[0, 1] + S[1, 0], S[2, 0] + S[0, 2]],
                       [S[2, 0] - S[0, 2], S[0, 1] + S[1, 0], -S[0, 0] + S[1, 1] - S[2, 2], S[1, 2] + S[2, 1]],
                       [S[0, 1] - S[1, 0], S[2, 0] + S[0, 2], S[1, 2] + S[2, 1], -S[0, 0] - S[1, 1] + S[2, 2]]])
        _, Vq = _eigh_desc(Nq)
        qw, qx, qy, qz = Vq[:, 0]
        R = np.array([[qw * qw + qx * qx - qy * qy - qz * qz, 2 * (qx * qy - qw * qz), 2 * (qx * qz + qw * qy)],
                      [2 * (qx * qy + qw * qz), qw * qw - qx * qx + qy * qy - qz * qz, 2 * (qy * qz - qw * qx)],
                      [2 * (qx * qz - qw * qy), 2 * (qy * qz + qw * qx), qw * qw - qx * qx - qy * qy + qz * qz]])
        t = pc0 - R @ pw0
        Xc = pws @ R.T + t
        with np.errstate(all="ignore"):
            err = np.sqrt((us[:, 0] - (uc + fu * Xc[:, 0] / Xc[:, 2])) ** 2 + (us[:, 1] - (vc + fv * Xc[:, 1] / Xc[:, 2])) ** 2).sum() / n
        return R, t, err

    best = None
    with np.errstate(all="ignore"):
        for cand in range(3):
            if cand == 0:
                x = lsq(L[:, [0, 1, 3, 6]], rho)
                b0 = np.sqrt(abs(x[0]))
                sg = -1.0 if x[0] < 0 else 1.0
                be = [b0, sg * x[1] / b0, sg * x[2] / b0, sg * x[3] / b0]
            elif cand == 1:
                x = lsq(L[:, [0, 1, 2]], rho)
                be = [np.sqrt(-x[0]), np.sqrt(-x[2]) if x[2] < 0 else 0.0, 0.0, 0.0] if x[0] < 0 else \
                     [np.sqrt(x[0]), np.sqrt(x[2]) if x[2] > 0 else 0.0, 0.0, 0.0]
                if x[1] < 0:
                    be[0] = -be[0]
            else:
                x = lsq(L[:, [0, 1, 2, 3, 4]], rho)
                be = [np.sqrt(-x[0]), np.sqrt(-x[2]) if x[2] < 0 else 0.0, 0.0, 0.0] if x[0] < 0 else \
                     [np.sqrt(x[0]), np.sqrt(x[2]) if x[2] > 0 else 0.0, 0.0, 0.0]
                if x[1] < 0:
                    be[0] = -be[0]
                be[2] = x[3] / be[0]
            try:
                R, t, err = r_and_t(gauss_newton(be))
            except np.linalg.LinAlgError:
                continue
            if np.isfinite(err) and np.isfinite(R).all() and np.isfinite(t).all() and (best is None or err < best[2]):
                best = (R, t, err)
    return None if best is None else (best[0], best[1])


def reproj_err2(R, t, K, xyz, uv):
    """PnPRansacCallback::computeError: projectPoints in double -> float32, squared distance in float32."""
    X = np.asarray(xyz, F32).astype(np.float64)
    x = X @ np.asarray(R, np.float64).T + np.asarray(t, np.float64).reshape(3)
    with np.errstate(all="ignore"):
        z = np.where(x[:, 2] != 0, 1.0 / x[:, 2], 1.0)
    pu = (x[:, 0] * z * K[0, 0] + K[0, 2]).astype(F32)
    pv = (x[:, 1] * z * K[1, 1] + K[1, 2]).astype(F32)
    uv = np.asarray(uv, F32)
    du, dv = (uv[:, 0] - pu).astype(F32), (uv[:, 1] - pv).astype(F32)
    return ((du * du).astype(F32) + (dv * dv).astype(F32)).astype(F32)


def ransac_scan(counts, n, iters=100, confidence=0.99):
    """RANSACPointSetRegistrator::run over precomputed inlier counts (counts[h] < 0: the minimal solve gave no model).
    Returns (best index or -1, iterations run, best count)."""
    niters, best, max_good, it = iters, -1, 0, 0
    while it < niters:
        good = int(counts[it])
        it += 1
        if good < 0:
            continue
        if good > max(max_good, MODEL_POINTS - 1):
            max_good, best = good, it - 1
            ep = min(max((n - good) / n, 0.0), 1.0)
            num = max(1.0 - confidence, np.finfo(np.float64).tiny)
            denom = 1.0 - (1.0 - ep) ** MODEL_POINTS
            if denom < np.finfo(np.float64).tiny:
                niters = 0
            else:
                num, denom = np.log(num), np.log(denom)
                if not (denom >= 0 or -num >= niters * (-denom)):
                    niters = int(np.rint(num / denom))
    return (best if max_good > 0 else -1), it, max_good


def ransac_replica(obj, img, K, solver="epnp", iters=100, thr_px=1.5, confidence=0.99, orig=None):
    """cv2.solvePnPRansac(obj, img, K, None, iterationsCount=iters, reprojectionError=thr_px) from its parts.  All `iters`
    minimal models are solved and scored; the sequential loop with its adaptive stop is then replayed over the counts
    (ransac_scan) — the same result as stopping early, since a model's count does not depend on the iterations before it.
    `orig` (solver="epnp"): the original index of every (resampled) point; a sample that repeats one is spent without a model.
    Returns (ok, rvec, tvec, inliers (k,1) int32, counts (iters,), poses [iters], best index, iterations run)."""
    import cv2
    obj, img = np.asarray(obj, F32), np.asarray(img, F32).reshape(-1, 2)
    n = len(obj)
    K = np.asarray(K, np.float64)
    tab = mwc_table(n, iters)
    thr = F32(thr_px * thr_px)
    counts = np.full(iters, -1, np.int32)
    poses = [None] * iters
    for it in range(iters):
        s = tab[it]
        if solver == "cv":
            ok, rv, tv = cv2.solvePnP(obj[s], img[s].reshape(-1, 1, 2), K, None, flags=cv2.SOLVEPNP_EPNP)
            pose = (cv2.Rodrigues(rv)[0], tv.ravel()) if ok else None
        elif orig is not None and len(set(np.asarray(orig)[s].tolist())) < MODEL_POINTS:
            pose = None          # the resample repeats a correspondence inside this sample: four distinct points, no model (csrc/pnp.cu)
        else:
            try:
                pose = epnp5(obj[s], img[s], K)
            except np.linalg.LinAlgError:
                pose = None
        if pose is None:
            continue
        poses[it] = pose
        counts[it] = int((reproj_err2(pose[0], pose[1], K, obj, img) <= thr).sum())
    best, iters_run, _ = ransac_scan(counts, n, iters, confidence)
    if best < 0:
        return False, None, None, None, counts, poses, best, iters_run
    R, t = poses[best]
    inl = np.nonzero(reproj_err2(R, t, K, obj, img) <= thr)[0]
    ok, rv, tv = cv2.solvePnP(obj[inl].astype(np.float64), img[inl].astype(np.float64).reshape(-1, 1, 2), K, None,
                              flags=cv2.SOLVEPNP_ITERATIVE)
    return bool(ok), rv, tv, inl.reshape(-1, 1).astype(np.int32), counts, poses, best, iters_run


def pose_3d_2d_ref(xyz, uv, K, boot, solver="epnp", iters=100, thr_px=1.5, min_inliers=20):
    """Lines :120-144 of computepose_3D_2D on already gathered / gated correspondences: `boot` (3, n) are the bootstrap index
    rows the reference draws with np.random.randint(0, n, n).  Returns dict(ok, T_rel (pose.pose: the inverse of [R|t]), n_inl,
    restart, iteration, rvec, tvec, counts (restarts, iters), inliers of the winning restart (indices into its resample))."""
    import cv2
    best, out = 0, dict(ok=False, T_rel=np.eye(4), n_inl=0, restart=-1, iteration=-1, counts=[])
    for r, order in enumerate(np.asarray(boot)):
        obj, img = np.asarray(xyz, F32)[order], np.asarray(uv, F32)[order]
        ok, rv, tv, inl, counts, poses, bi, _ = ransac_replica(obj, img, K, solver, iters, thr_px, orig=order)
        out["counts"].append(counts)
        if ok and inl.shape[0] > best and inl.shape[0] > min_inliers:
            best = inl.shape[0]
            M = np.eye(4)
            M[:3, :3] = cv2.Rodrigues(rv)[0]
            M[:3, 3] = tv.ravel()
            out.update(ok=True, T_rel=np.linalg.inv(M), n_inl=best, restart=r, iteration=bi, rvec=rv.ravel(), tvec=tv.ravel(),
                       inliers=inl.ravel(), minimal=poses[bi])
    out["counts"] = np.asarray(out["counts"])
    return out
