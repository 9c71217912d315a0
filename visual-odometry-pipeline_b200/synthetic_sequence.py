"""Synthetic RGB-D *sequence* with persistent landmarks (for the keyframe-based process_frame loop and the
trajectory-level ATE / RPE comparison).  A corridor of world landmarks, a camera moving forward ~0.67 m per
frame with small rotations (statistics of the reference's shipped KITTI-03 trajectory),
per-frame keypoints = noisy projections, depth maps stamped at the truncated keypoint pixels, descriptors =
landmark prototype + per-view noise, plus unmatched distractors.  numpy only.
"""
import numpy as np

from .synthetic import KITTI_K, KITTI_WH, MASTER_SEED, _rodrigues


def _view_descriptors(rng, kind, proto, n_extra):
    n = proto.shape[0]
    if kind == "orb":
        bits = proto ^ (rng.random(proto.shape) < 0.04).astype(np.uint8)
        extra = rng.integers(0, 2, size=(n_extra, 256), dtype=np.uint8)
        return np.packbits(np.concatenate([bits, extra], 0), axis=1)
    if kind == "sift":
        d = np.abs(proto + 0.08 * rng.standard_normal(proto.shape))
        extra = np.abs(rng.standard_normal((n_extra, 128)))
        d = np.concatenate([d, extra], 0)
        d *= 512.0 / np.linalg.norm(d, axis=1, keepdims=True)
        return np.clip(np.rint(d), 0, 255).astype(np.float32)
    d = proto + 0.05 * rng.standard_normal(proto.shape)
    extra = rng.standard_normal((n_extra, 128))
    d = np.concatenate([d, extra], 0)
    return (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)


def make_sequence(n_frames=30, n_kp=1500, kind="orb", K=KITTI_K, wh=KITTI_WH, noise_px=0.3, seed=MASTER_SEED):
    """Returns (frames, gt_poses): frames[i] = dict(kp (n,2) f64, desc, depth (H,W) f32); gt_poses (n_frames,4,4)
    camera-to-world with frame 0 at the identity."""
    rng = np.random.default_rng(seed)
    W, H = wh
    fx, fy, cx, cy = K[0, 0], K[1, 1], K[0, 2], K[1, 2]
    length = 0.9 * n_frames + 70.0
    n_land = int(n_kp * 14 * length / 70.0)
    P = np.stack([rng.uniform(-28, 28, n_land), rng.uniform(-9, 9, n_land), rng.uniform(3, length, n_land)], 1)
    if kind == "orb":
        proto = rng.integers(0, 2, size=(n_land, 256), dtype=np.uint8)
    elif kind == "sift":
        proto = np.abs(rng.standard_normal((n_land, 128)))
    else:
        proto = rng.standard_normal((n_land, 128))
        proto /= np.linalg.norm(proto, axis=1, keepdims=True)

    poses = [np.eye(4)]
    for _ in range(1, n_frames):
        step = np.eye(4)
        step[:3, :3] = _rodrigues(rng.normal(0, 0.006, 3))
        step[:3, 3] = [rng.normal(0, 0.02), rng.normal(0, 0.01), float(np.clip(rng.normal(0.67, 0.15), 0.2, 1.1))]
        poses.append(poses[-1] @ step)
    poses = np.stack(poses)

    vv, uu = np.mgrid[0:H, 0:W]
    background = (60.0 + 5.0 * np.sin(uu / 97.0) * np.cos(vv / 61.0)).astype(np.float32)  # beyond the 50 m gate
    frames = []
    for T in poses:
        R, c = T[:3, :3], T[:3, 3]
        Xc = (P - c) @ R                      # R^T (P - c)
        z = Xc[:, 2]
        with np.errstate(divide="ignore", invalid="ignore"):
            u = fx * Xc[:, 0] / z + cx
            v = fy * Xc[:, 1] / z + cy
        vis = np.nonzero((z > 3.0) & (z < 48.0) & (u >= 2) & (u < W - 2) & (v >= 2) & (v < H - 2))[0]
        n_l = min(len(vis), int(n_kp * 0.85))
        vis = vis[:n_l]  # lowest landmark ids first: a landmark stays tracked while it is visible (high overlap)
        kp = np.stack([u[vis], v[vis]], 1) + rng.normal(0, noise_px, (n_l, 2))
        kp[:, 0] = np.clip(kp[:, 0], 0, W - 1.001)
        kp[:, 1] = np.clip(kp[:, 1], 0, H - 1.001)
        depth = background.copy()
        depth[kp[:, 1].astype(np.int64), kp[:, 0].astype(np.int64)] = z[vis].astype(np.float32)
        n_extra = n_kp - n_l
        kp_extra = np.stack([rng.uniform(0, W - 1.001, n_extra), rng.uniform(0, H - 1.001, n_extra)], 1)
        desc = _view_descriptors(rng, kind, proto[vis], n_extra)
        order = rng.permutation(n_kp)
        frames.append(dict(kp=np.concatenate([kp, kp_extra], 0)[order], desc=np.ascontiguousarray(desc[order]), depth=depth))
    return frames, poses
