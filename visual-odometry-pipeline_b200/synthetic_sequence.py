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


def make_long_sequence(n_frames=800, n_kp=1500, kind="orb", K=KITTI_K, wh=KITTI_WH, noise_px=0.3, seed=MASTER_SEED):
    """Like make_sequence, for hundreds of frames: the landmarks FOLLOW the camera path (make_sequence scatters them in a
    straight corridor, which a random-walk heading leaves after a few hundred frames).  Every frame adds landmarks in the far
    slab of its own view volume, so the visible set is renewed at the rate the camera advances and a landmark is tracked
    from ~48 m down to ~3 m (about 65 frames).  Same keypoint / depth / descriptor model.  Returns (frames, gt_poses)."""
    rng = np.random.default_rng(seed)
    W, H = wh
    fx, fy, cx, cy = K[0, 0], K[1, 1], K[0, 2], K[1, 2]
    poses = [np.eye(4)]
    for _ in range(1, n_frames):
        step = np.eye(4)
        step[:3, :3] = _rodrigues(rng.normal(0, 0.006, 3))
        step[:3, 3] = [rng.normal(0, 0.02), rng.normal(0, 0.01), float(np.clip(rng.normal(0.67, 0.15), 0.2, 1.1))]
        poses.append(poses[-1] @ step)
    poses = np.stack(poses)

    def box(n, z0, z1):          # points in a camera-frame box that covers the view frustum up to z1
        z = rng.uniform(z0, z1, n)
        return np.stack([rng.uniform(-0.95, 0.95, n) * z * (W / 2) / fx, rng.uniform(-0.95, 0.95, n) * z * (H / 2) / fy + 0.0, z], 1)

    per_frame = int(n_kp * 0.85 * 0.9 / 45.0 * 2.2) + 8           # renewal rate: ~0.67 m of a 45 m deep volume per frame, with margin
    chunks = [box(int(n_kp * 0.85 * 2.2), 3.0, 48.0)]              # prefill of frame 0's volume
    for T in poses[1:]:
        chunks.append(box(per_frame, 46.0, 49.5) @ T[:3, :3].T + T[:3, 3])
    P = np.concatenate(chunks, 0)
    n_land = len(P)
    if kind == "orb":
        proto = rng.integers(0, 2, size=(n_land, 256), dtype=np.uint8)
    elif kind == "sift":
        proto = np.abs(rng.standard_normal((n_land, 128))).astype(np.float32)
    else:
        proto = rng.standard_normal((n_land, 128)).astype(np.float32)
        proto /= np.linalg.norm(proto, axis=1, keepdims=True)

    vv, uu = np.mgrid[0:H, 0:W]
    background = (60.0 + 5.0 * np.sin(uu / 97.0) * np.cos(vv / 61.0)).astype(np.float32)  # beyond the 50 m gate
    frames = []
    born = np.concatenate([np.zeros(len(chunks[0]), np.int64)] + [np.full(len(c), i + 1, np.int64) for i, c in enumerate(chunks[1:])])
    for fi, T in enumerate(poses):
        R, c = T[:3, :3], T[:3, 3]
        alive = np.nonzero(born <= fi)[0]
        Xc = (P[alive] - c) @ R
        z = Xc[:, 2]
        with np.errstate(divide="ignore", invalid="ignore"):
            u = fx * Xc[:, 0] / z + cx
            v = fy * Xc[:, 1] / z + cy
        ok = (z > 3.0) & (z < 48.0) & (u >= 2) & (u < W - 2) & (v >= 2) & (v < H - 2)
        vis = alive[ok]
        n_l = min(len(vis), int(n_kp * 0.85))
        sel = np.nonzero(ok)[0][:n_l]          # oldest landmarks first: a landmark stays tracked while it is visible
        vis = vis[:n_l]
        kp = np.stack([u[sel], v[sel]], 1) + rng.normal(0, noise_px, (n_l, 2))
        kp[:, 0] = np.clip(kp[:, 0], 0, W - 1.001)
        kp[:, 1] = np.clip(kp[:, 1], 0, H - 1.001)
        depth = background.copy()
        depth[kp[:, 1].astype(np.int64), kp[:, 0].astype(np.int64)] = z[sel].astype(np.float32)
        n_extra = n_kp - n_l
        kp_extra = np.stack([rng.uniform(0, W - 1.001, n_extra), rng.uniform(0, H - 1.001, n_extra)], 1)
        desc = _view_descriptors(rng, kind, proto[vis], n_extra)
        order = rng.permutation(n_kp)
        frames.append(dict(kp=np.concatenate([kp, kp_extra], 0)[order], desc=np.ascontiguousarray(desc[order]), depth=depth))
    return frames, poses
