"""The reference's keyframe-based VO loop on the CPU — TEST INFRASTRUCTURE ONLY (see oracle.py header).

Restates VisualOdometry.process_frame (VisualOdometry_Stereo.py:223-297) with the reference's own third-party
calls (cv2.BFMatcher / cv2.solvePnPRansac through oracle/reference_path.py): match against the KEYFRAME (:251),
3 px flow filter (:260-264), computepose_3D_2D (:87-149), 1.5 m x frame-gap gate (:271), chaining (:283) or
identity on failure (:290), keyframe refresh when common_pts < 200, inliers < 100, |t| > 1.5 or > 3 failures
(:285-296).  Feature extraction is replaced by precomputed (keypoints, descriptors) per frame, exactly like the
GPU test stubs it, so both sides see the same inputs.
"""
import numpy as np

from . import reference_path as rp


class ReferenceVO:
    def __init__(self, K, matcher="knn_ratio", seed=8214):
        self.K = np.asarray(K, np.float64)
        self.matcher = matcher
        self.rng = np.random.RandomState(seed)      # vo_stereo_runner.py:20-24
        self.key = None                             # (id, kp, desc, depth, pose)
        self.bad_pnp = 0
        self.poses = []

    def process_frame(self, kp, desc, depth, frame_no):
        if frame_no == 0:
            self.key = (0, kp, desc, depth, np.eye(4))
            self.poses.append(np.eye(4))
            return self.poses[-1]
        kid, kkp, kdesc, kdepth, kpose = self.key
        if self.matcher == "knn_ratio":
            m = rp.match_knn_ratio(kdesc, desc)
        elif self.matcher == "hamming_mutual":
            m = rp.match_hamming_mutual(kdesc, desc)
        else:
            m = rp.match_r2d2(kdesc, desc)
        retval, T, common, inl, dist = False, np.eye(4), 0, 0, 0.0
        try:
            r = kkp[m[:, 0], :2].astype(np.float32)
            c = kp[m[:, 1], :2].astype(np.float32)
            flow = np.linalg.norm(r - c, axis=1)
            r, c = r[flow >= 3], c[flow >= 3]
            retval, T, common, inl = rp.pose_3d_2d(r, c, kdepth, self.K, self.rng)
            if not retval:
                self.bad_pnp += 1
            dist = float(np.linalg.norm(T[:3, 3]))
            if dist > 1.5 * (frame_no - kid):
                retval = False
                self.bad_pnp += 1
        except Exception:
            self.bad_pnp += 1
            retval = False
        to_update = False
        if retval:
            self.bad_pnp = 0
            pose = kpose @ T
            to_update = common < 200 or inl < 100 or dist > 1.5
        else:
            pose = kpose.copy()
        self.poses.append(pose)
        if to_update or self.bad_pnp > 3:
            self.key = (frame_no, kp, desc, depth, pose)
        return pose
