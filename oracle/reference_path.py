"""The reference's CPU path for one frame pair, restated with the same third-party calls it makes —
TEST / BASELINE INFRASTRUCTURE ONLY (see oracle.py header).  bench.py times this as the CPU baseline
(`cpu_baseline`, `--impl reference`); tests use it for pose-level agreement.

Reference call sites (file:line under the upstream repository):
  knn + ratio        feature_extractors/SIFT.py:25-34, ORB.py:23-32   cv2.BFMatcher().knnMatch(k=2), 0.85
  Hamming mutual     north-star semantics                              cv2.BFMatcher(NORM_HAMMING, crossCheck=True)
  R2D2 matcher       R2D2.py:53-66                                     torch matmul / topk / max
  gather + filter    VisualOdometry_Stereo.py:257-264
  depthTo3d + gate   VisualOdometry_Stereo.py:96-105                   (cv2.rgbd is absent here: numpy restatement)
  3x solvePnPRansac  VisualOdometry_Stereo.py:120-135                  bootstrap order, iterationsCount=100, 1.5 px
  pose assembly      VisualOdometry_Stereo.py:137-144                  Rodrigues, [R|t], inverse
"""
import numpy as np


def set_threads(n):
    import cv2
    cv2.setNumThreads(int(n))
    try:
        import torch
        torch.set_num_threads(int(n))
    except Exception:
        pass
    return cv2.getNumThreads()


_bf_l2 = None
_bf_ham = None


def match_knn_ratio(ref_desc, cur_desc, ratio=0.85):
    import cv2
    global _bf_l2
    if _bf_l2 is None:
        _bf_l2 = cv2.BFMatcher()           # NORM_L2 also for uint8 ORB descriptors
    out = []
    for m, n in _bf_l2.knnMatch(ref_desc, cur_desc, k=2):
        if m.distance < ratio * n.distance:
            out.append((m.queryIdx, m.trainIdx))
    return np.asarray(out, np.int64).reshape(-1, 2)


def match_hamming_mutual(ref_desc, cur_desc):
    import cv2
    global _bf_ham
    if _bf_ham is None:
        _bf_ham = cv2.BFMatcher(cv2.NORM_HAMMING, crossCheck=True)
    ms = _bf_ham.match(ref_desc, cur_desc)
    out = np.asarray([(m.queryIdx, m.trainIdx) for m in ms], np.int64).reshape(-1, 2)
    return out[np.argsort(out[:, 0], kind="stable")]


def match_r2d2(ref_desc, cur_desc, ratio=0.90):
    import torch
    a = torch.as_tensor(ref_desc)
    b = torch.as_tensor(cur_desc)
    sim = a @ b.t()
    top_s, top_i = torch.topk(sim, 2, dim=1)
    dist = torch.sqrt(2 - 2 * top_s)
    back = torch.max(sim, dim=0)[1]
    rows = torch.arange(sim.shape[0])
    keep = (back[top_i[:, 0]] == rows) & ((dist[:, 0] / (dist[:, 1] + 1e-8)) <= ratio)
    return torch.stack([rows[keep], top_i[keep, 0]], 1).numpy().astype(np.int64)


def depth_to_3d(depth, K):
    """cv2.rgbd.depthTo3d for float depth, restated (rgbd/depth_to_3d.hpp): intrinsics cast to fp32."""
    d = depth.astype(np.float32)
    H, W = d.shape
    fx, fy, cx, cy = (np.float32(K[0, 0]), np.float32(K[1, 1]), np.float32(K[0, 2]), np.float32(K[1, 2]))
    xc = (np.arange(W, dtype=np.float32) - cx) * (np.float32(1) / fx)
    yc = (np.arange(H, dtype=np.float32) - cy) * (np.float32(1) / fy)
    out = np.empty((H, W, 3), np.float32)
    out[..., 0] = xc[None, :] * d
    out[..., 1] = yc[:, None] * d
    out[..., 2] = d
    return out


def pose_3d_2d(left_kp, right_kp, depth, K, rng=np.random, restarts=3):
    """VisualOdometry.computepose_3D_2D.  Returns (retval, T_rel 4x4, n_common, best_inlier)."""
    import cv2
    three_d = depth_to_3d(depth.copy(), K)
    pts = three_d[left_kp[:, 1].astype(np.int32), left_kp[:, 0].astype(np.int32)]
    good = (pts[:, 2] > 0) & (pts[:, 2] < 50)
    pts, left_kp, right_kp = pts[good], left_kp[good], right_kp[good]
    best, best_rt = 0, None
    for _ in range(restarts):
        order = rng.randint(0, pts.shape[0], pts.shape[0])
        obj = pts.copy()[order]
        img = np.ascontiguousarray(right_kp.copy()[order]).reshape(-1, 1, 2)
        ok, r, t, inl = cv2.solvePnPRansac(objectPoints=obj, imagePoints=img, cameraMatrix=K, distCoeffs=None,
                                           iterationsCount=100, reprojectionError=1.5)
        if ok and inl.shape[0] > best and inl.shape[0] > 20:
            best, best_rt = inl.shape[0], (r, t)
    T = np.eye(4)
    if best_rt is None:
        return False, T, len(left_kp), best
    M = np.eye(4)
    M[:3, :3] = cv2.Rodrigues(best_rt[0])[0]
    M[:3, 3:] = best_rt[1]
    return True, np.linalg.inv(M), len(left_kp), best


def process_pair(ref_desc, cur_desc, ref_kp, cur_kp, depth, K, matcher="knn_ratio", rng=np.random):
    """match -> gather -> 3 px flow filter -> computepose_3D_2D (VisualOdometry_Stereo.py:256-269)."""
    if matcher == "knn_ratio":
        m = match_knn_ratio(ref_desc, cur_desc)
    elif matcher == "hamming_mutual":
        m = match_hamming_mutual(ref_desc, cur_desc)
    elif matcher == "r2d2":
        m = match_r2d2(ref_desc, cur_desc)
    else:
        raise ValueError(matcher)
    if len(m) == 0:
        return False, np.eye(4), 0, 0
    r = ref_kp[m[:, 0], :2].astype(np.float32)
    c = cur_kp[m[:, 1], :2].astype(np.float32)
    flow = np.linalg.norm(r - c, axis=1)
    r, c = r[flow >= 3], c[flow >= 3]
    try:
        return pose_3d_2d(r, c, depth, K, rng)
    except Exception:
        return False, np.eye(4), 0, 0
