"""Deterministic synthetic frame pairs with ground truth.

KITTI-shaped (1241x376, intrinsics of config/vo_params.yaml:10-18) or ZED-shaped (2208x1242) RGB-D pairs:
landmarks with a known relative camera motion, keypoints with pixel noise, a fraction of
descriptor-consistent but geometrically wrong associations (the matcher pairs them, RANSAC must reject
them), unmatched distractors, far / invalid depth to exercise the 0<Z<50 gate, and SIFT-like
(integer-valued f32, norm ~512), R2D2-like (unit f32) or ORB-like (256-bit) descriptors.

numpy only: the same arrays feed the CUDA path, the CPU oracle and the CPU baseline.
"""
import numpy as np

KITTI_K = np.array([[721.53, 0.0, 609.55], [0.0, 721.53, 172.85], [0.0, 0.0, 1.0]])
KITTI_WH = (1241, 376)
ZED_K = np.array([[1400.0, 0.0, 1104.0], [0.0, 1400.0, 621.0], [0.0, 0.0, 1.0]])
ZED_WH = (2208, 1242)
MASTER_SEED = 8214  # vo_stereo_runner.py:20


def _rodrigues(w):
    th = np.linalg.norm(w)
    if th < 1e-12:
        return np.eye(3)
    k = w / th
    Kx = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    return np.eye(3) + np.sin(th) * Kx + (1 - np.cos(th)) * (Kx @ Kx)


def _descriptors(rng, kind, n_land, n_ref, n_cur):
    """prototype + per-view noise for landmarks, independent rows for distractors."""
    if kind == "orb":
        proto = rng.integers(0, 2, size=(n_land, 256), dtype=np.uint8)

        def view(n_total):
            bits = proto ^ (rng.random((n_land, 256)) < 0.05).astype(np.uint8)
            extra = rng.integers(0, 2, size=(n_total - n_land, 256), dtype=np.uint8)
            return np.packbits(np.concatenate([bits, extra], 0), axis=1)  # (n,32) uint8

        return view(n_ref), view(n_cur)
    if kind == "r2d2":
        proto = rng.standard_normal((n_land, 128))
        proto /= np.linalg.norm(proto, axis=1, keepdims=True)

        def view(n_total):
            d = proto + 0.05 * rng.standard_normal((n_land, 128))
            extra = rng.standard_normal((n_total - n_land, 128))
            d = np.concatenate([d, extra], 0)
            d /= np.linalg.norm(d, axis=1, keepdims=True)
            return d.astype(np.float32)

        return view(n_ref), view(n_cur)
    if kind == "sift":
        proto = np.abs(rng.standard_normal((n_land, 128)))

        def view(n_total):
            d = np.abs(proto + 0.08 * rng.standard_normal((n_land, 128)))
            extra = np.abs(rng.standard_normal((n_total - n_land, 128)))
            d = np.concatenate([d, extra], 0)
            d *= 512.0 / np.linalg.norm(d, axis=1, keepdims=True)
            return np.clip(np.rint(d), 0, 255).astype(np.float32)  # integer-valued, like real SIFT

        return view(n_ref), view(n_cur)
    raise ValueError(f"unknown descriptor kind {kind!r}")


def make_pair(index, n_kp=2000, kind="sift", K=KITTI_K, wh=KITTI_WH, n_cur=None, outlier_frac=0.30,
              land_frac=0.80, noise_px=0.3, seed=MASTER_SEED):
    """One synthetic frame pair.  Returns a dict of numpy arrays:
    ref_desc, cur_desc, ref_kp (N,2) f32, cur_kp (M,2) f32, depth (H,W) f32, T_rel (4,4) f64 ground truth
    (pose of camera 2 in camera 1's frame — what the reference stores), gt_cur_of_ref (N,) int (-1 = none),
    geom_ok (N,) bool (association is geometrically consistent)."""
    rng = np.random.default_rng(seed + int(index))
    W, H = wh
    n_ref = int(n_kp)
    n_cur = int(n_cur if n_cur is not None else n_kp)
    n_land = int(min(n_ref, n_cur) * land_frac)
    fx, fy, cx, cy = K[0, 0], K[1, 1], K[0, 2], K[1, 2]

    # motion: forward translation ~N(0.67,0.2) m, small rotation (statistics of the shipped KITTI-03 run)
    d = float(np.clip(rng.normal(0.67, 0.2), 0.05, 1.2))
    c = np.array([rng.normal(0, 0.02), rng.normal(0, 0.01), d])
    Rc = _rodrigues(rng.normal(0, 0.008, 3))
    T_rel = np.eye(4)
    T_rel[:3, :3], T_rel[:3, 3] = Rc, c

    # landmarks on distinct pixels of frame 1 (margin keeps truncation inside the image)
    flat = rng.choice((W - 8) * (H - 8), size=n_ref, replace=False)
    pu = (flat % (W - 8) + 4).astype(np.int64)
    pv = (flat // (W - 8) + 4).astype(np.int64)
    z = rng.uniform(4.0, 45.0, n_ref)
    far = rng.random(n_ref) < 0.10
    z[far] = rng.uniform(50.0, 80.0, int(far.sum()))
    bad = rng.random(n_ref) < 0.01
    zero = bad & (rng.random(n_ref) < 0.5)
    z_map = z.copy()
    z_map[bad] = np.nan
    z_map[zero] = 0.0

    vv, uu = np.mgrid[0:H, 0:W]
    depth = (22.0 + 9.0 * np.sin(uu / 97.0) * np.cos(vv / 61.0) + 6.0 * (vv / H)).astype(np.float32)
    depth[pv, pu] = z_map.astype(np.float32)

    ref_kp = np.stack([pu + rng.random(n_ref) * 0.999, pv + rng.random(n_ref) * 0.999], 1).astype(np.float32)

    # 3-D point the pipeline will reconstruct: truncated pixel, fp32 depth (VisualOdometry_Stereo.py:96-97)
    z32 = depth[pv, pu].astype(np.float64)
    X1 = np.stack([(pu - cx) / fx * z32, (pv - cy) / fy * z32, z32], 1)
    with np.errstate(invalid="ignore", divide="ignore"):
        X2 = (X1 - c) @ Rc  # = Rc^T (X1 - c)
        proj = np.stack([fx * X2[:, 0] / X2[:, 2] + cx, fy * X2[:, 1] / X2[:, 2] + cy], 1)
    proj += rng.normal(0, noise_px, proj.shape)

    cur_kp_land = proj[:n_land].copy()
    inside = (np.isfinite(cur_kp_land).all(1) & (cur_kp_land[:, 0] >= 1) & (cur_kp_land[:, 0] < W - 1)
              & (cur_kp_land[:, 1] >= 1) & (cur_kp_land[:, 1] < H - 1) & (X2[:n_land, 2] > 0.1))
    wrong = (rng.random(n_land) < outlier_frac) | ~inside
    n_wrong = int(wrong.sum())
    cur_kp_land[wrong] = np.stack([rng.uniform(1, W - 1, n_wrong), rng.uniform(1, H - 1, n_wrong)], 1)
    cur_extra = np.stack([rng.uniform(1, W - 1, n_cur - n_land), rng.uniform(1, H - 1, n_cur - n_land)], 1)
    cur_kp = np.concatenate([cur_kp_land, cur_extra], 0)

    ref_desc, cur_desc = _descriptors(rng, kind, n_land, n_ref, n_cur)

    # shuffle both frames independently so that matches are not the identity map
    perm_r = rng.permutation(n_ref)
    perm_c = rng.permutation(n_cur)
    inv_c = np.empty(n_cur, np.int64)
    inv_c[perm_c] = np.arange(n_cur)
    gt = np.full(n_ref, -1, np.int64)
    gt[:n_land] = inv_c[:n_land]
    geom_ok = np.zeros(n_ref, bool)
    geom_ok[:n_land] = ~wrong & np.isfinite(z_map[:n_land]) & (z_map[:n_land] > 0) & (z_map[:n_land] < 50)
    return dict(
        ref_desc=np.ascontiguousarray(ref_desc[perm_r]), cur_desc=np.ascontiguousarray(cur_desc[perm_c]),
        ref_kp=np.ascontiguousarray(ref_kp[perm_r]), cur_kp=np.ascontiguousarray(cur_kp[perm_c].astype(np.float32)),
        depth=depth, T_rel=T_rel, gt_cur_of_ref=gt[perm_r], geom_ok=geom_ok[perm_r], K=np.asarray(K, np.float64),
    )


def _proto(rng, kind, n):
    if kind == "orb":
        return rng.integers(0, 2, size=(n, 256), dtype=np.uint8)
    if kind == "sift":
        return np.abs(rng.standard_normal((n, 128))).astype(np.float32)
    p = rng.standard_normal((n, 128)).astype(np.float32)
    return p / np.linalg.norm(p, axis=1, keepdims=True)


def _view(rng, kind, proto):
    """One frame's noisy view of its keypoints' descriptor prototypes (same noise model as _descriptors)."""
    if kind == "orb":
        return np.packbits(proto ^ (rng.random(proto.shape, dtype=np.float32) < 0.05).astype(np.uint8), axis=1)
    if kind == "sift":
        d = np.abs(proto + np.float32(0.08) * rng.standard_normal(proto.shape, dtype=np.float32))
        d *= np.float32(512.0) / np.linalg.norm(d, axis=1, keepdims=True)
        return np.clip(np.rint(d), 0, 255).astype(np.float32)
    d = proto + np.float32(0.05) * rng.standard_normal(proto.shape, dtype=np.float32)
    return (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)


def make_chain(first_index, n_pairs, n_kp=2000, kind="sift", K=KITTI_K, wh=KITTI_WH, outlier_frac=0.30, land_frac=0.80,
               noise_px=0.3, seed=MASTER_SEED, out=None):
    """A synthetic SEQUENCE of n_pairs + 1 frames in which every consecutive pair (i, i+1) has the statistics of
    make_pair: frame i+1 is the current frame of pair i and the reference frame of pair i+1, so a frame's descriptors,
    keypoints and depth map exist (and travel to the GPU) once.  Frame j+1 is built from frame j: land_frac of frame
    j's keypoints are re-observed (projection under the pair's motion + pixel noise; outlier_frac of them at a random
    position: descriptor-consistent, geometrically wrong), the rest are new keypoints; every keypoint of every frame
    gets a depth stamp at its truncated pixel (10 % beyond the 50 m gate, 1 % NaN / 0), which defines the 3-D point the
    pipeline reconstructs when the frame acts as the reference.
    Returns dict(desc [F,N,..], kp [F,N,2] f32, depth [F,H,W] f32, T_rel [F-1,4,4] f64 (pose of camera i+1 in camera i),
    gt_cur_of_ref [F-1,N] (index into frame i+1, -1 = none), geom_ok [F-1,N], K).  `out`: optional dict of preallocated
    arrays (e.g. pinned host memory) with those keys and shapes for desc / kp / depth."""
    rng = np.random.default_rng(seed + 7919 * int(first_index))
    W, H = wh
    F, N = int(n_pairs) + 1, int(n_kp)
    n_land = int(N * land_frac)
    fx, fy, cx, cy = K[0, 0], K[1, 1], K[0, 2], K[1, 2]
    ddim = (32,) if kind == "orb" else (128,)
    ddt = np.uint8 if kind == "orb" else np.float32
    desc = out["desc"] if out is not None else np.empty((F, N) + ddim, ddt)
    kp = out["kp"] if out is not None else np.empty((F, N, 2), np.float32)
    depth = out["depth"] if out is not None else np.empty((F, H, W), np.float32)
    T_all = np.empty((F - 1, 4, 4), np.float64)
    gt_all = np.full((F - 1, N), -1, np.int64)
    ok_all = np.zeros((F - 1, N), bool)
    vv, uu = np.mgrid[0:H, 0:W]
    background = (22.0 + 9.0 * np.sin(uu / 97.0) * np.cos(vv / 61.0) + 6.0 * (vv / H)).astype(np.float32)

    def stamp(j, pos):
        """depth map of frame j with a stamp under every keypoint; returns the depth each keypoint will read back."""
        z = rng.uniform(4.0, 45.0, N)
        far = rng.random(N) < 0.10
        z[far] = rng.uniform(50.0, 80.0, int(far.sum()))
        bad = rng.random(N) < 0.01
        z[bad] = np.where(rng.random(int(bad.sum())) < 0.5, 0.0, np.nan)
        pu, pv = pos[:, 0].astype(np.int64), pos[:, 1].astype(np.int64)
        depth[j] = background
        depth[j][pv, pu] = z.astype(np.float32)          # colliding pixels: the last stamp wins, and is what is read back
        return pu, pv, depth[j][pv, pu].astype(np.float64)

    pos = np.stack([rng.uniform(1, W - 1, N), rng.uniform(1, H - 1, N)], 1).astype(np.float32)
    proto = _proto(rng, kind, N)
    kp[0] = pos
    desc[0] = _view(rng, kind, proto)
    for j in range(F - 1):
        pu, pv, z = stamp(j, kp[j])
        d = float(np.clip(rng.normal(0.67, 0.2), 0.05, 1.2))
        c = np.array([rng.normal(0, 0.02), rng.normal(0, 0.01), d])
        Rc = _rodrigues(rng.normal(0, 0.008, 3))
        T_all[j] = np.eye(4)
        T_all[j][:3, :3], T_all[j][:3, 3] = Rc, c
        X1 = np.stack([(pu - cx) / fx * z, (pv - cy) / fy * z, z], 1)
        with np.errstate(invalid="ignore", divide="ignore"):
            X2 = (X1 - c) @ Rc
            proj = np.stack([fx * X2[:, 0] / X2[:, 2] + cx, fy * X2[:, 1] / X2[:, 2] + cy], 1)
        proj += rng.normal(0, noise_px, proj.shape)
        src = rng.permutation(N)[:n_land]                 # keypoints of frame j re-observed in frame j+1
        nxt = proj[src]
        inside = (np.isfinite(nxt).all(1) & (nxt[:, 0] >= 1) & (nxt[:, 0] < W - 1) & (nxt[:, 1] >= 1) & (nxt[:, 1] < H - 1)
                  & (X2[src, 2] > 0.1))
        wrong = (rng.random(n_land) < outlier_frac) | ~inside
        nw = int(wrong.sum())
        nxt[wrong] = np.stack([rng.uniform(1, W - 1, nw), rng.uniform(1, H - 1, nw)], 1)
        new_pos = np.concatenate([nxt, np.stack([rng.uniform(1, W - 1, N - n_land), rng.uniform(1, H - 1, N - n_land)], 1)], 0)
        new_proto = np.concatenate([proto[src], _proto(rng, kind, N - n_land)], 0)
        perm = rng.permutation(N)                         # frame j+1 in its own order: matches are not the identity map
        inv = np.empty(N, np.int64)
        inv[perm] = np.arange(N)
        gt_all[j, src] = inv[:n_land]
        ok_all[j, src] = ~wrong & np.isfinite(z[src]) & (z[src] > 0) & (z[src] < 50)
        kp[j + 1] = new_pos[perm].astype(np.float32)
        proto = new_proto[perm]
        desc[j + 1] = _view(rng, kind, proto)
    stamp(F - 1, kp[F - 1])                                # the last frame is never a reference; its map is a valid image anyway
    return dict(desc=desc, kp=kp, depth=depth, T_rel=T_all, gt_cur_of_ref=gt_all, geom_ok=ok_all, K=np.asarray(K, np.float64))


def chain_pair(chain, i):
    """Pair i of a make_chain sequence in make_pair's layout (views, no copies)."""
    return dict(ref_desc=chain["desc"][i], cur_desc=chain["desc"][i + 1], ref_kp=chain["kp"][i], cur_kp=chain["kp"][i + 1],
                depth=chain["depth"][i], T_rel=chain["T_rel"][i], gt_cur_of_ref=chain["gt_cur_of_ref"][i],
                geom_ok=chain["geom_ok"][i], K=chain["K"])


def make_batch(first_index, count, **kw):
    """Stack `count` pairs (indices first_index ...) into batch arrays [B, ...]."""
    pairs = [make_pair(first_index + i, **kw) for i in range(count)]
    out = {k: np.stack([p[k] for p in pairs], 0) for k in ("ref_desc", "cur_desc", "ref_kp", "cur_kp", "depth", "T_rel")}
    out["K"] = pairs[0]["K"]
    out["pairs"] = pairs
    return out


def pose_errors(T_est, T_gt):
    """(rotation error rad, translation error m) between two 4x4 poses."""
    # ||Ra - Rb||_F = 2 sqrt(2) |sin(theta/2)|: accurate for tiny angles, unlike arccos((tr-1)/2)
    f = np.linalg.norm(T_est[:3, :3] - T_gt[:3, :3])
    ang = 2.0 * np.arcsin(min(1.0, f / (2.0 * np.sqrt(2.0))))
    return float(ang), float(np.linalg.norm(T_est[:3, 3] - T_gt[:3, 3]))
