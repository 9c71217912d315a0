"""CUDA PnP-RANSAC vs the oracle under a shared hypothesis table: table, per-hypothesis inlier counts, winner
and inlier set bit-exact; refined pose within 1e-4 rad / 1e-3 m of the oracle refit and of cv2's."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROT_TOL, TRANS_TOL = 1e-4, 1e-3          # north-star tolerances


def _gpu(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _rot_angle(Ra, Rb):
    return float(2.0 * np.arcsin(min(1.0, np.linalg.norm(Ra - Rb) / (2.0 * np.sqrt(2.0)))))


def _run(xyz, uv, K, H, seed=8214, pair=0, cap=None):
    import torch
    from vo_b200 import ops
    n = len(xyz)
    cap = cap or max(n, 1)
    X = np.zeros((1, cap, 3), np.float32); X[0, :n] = xyz
    U = np.zeros((1, cap, 2), np.float32); U[0, :n] = uv
    n_pts = torch.tensor([n], dtype=torch.int32, device="cuda")
    hyp = ops.hypotheses(n_pts, H, seed, pair)
    res = ops.pnp_ransac(_gpu(X), _gpu(U), n_pts, K, hyp, 1.5, 20, 10, want_counts=True)
    return hyp, res


def test_hypothesis_table_equals_oracle(orc):
    import torch
    from vo_b200 import ops
    n_pts = torch.tensor([5000, 3, 4, 57], dtype=torch.int32, device="cuda")
    hyp = ops.hypotheses(n_pts, 1024, 8214, 10).cpu().numpy()
    for b, n in enumerate((5000, 3, 4, 57)):
        assert np.array_equal(hyp[b], orc.hypotheses(n, 1024, 8214, 10 + b))


@pytest.mark.parametrize("H", [128, 1024])
def test_counts_winner_mask_bit_exact_and_pose(golden, orc, H):
    g = golden("pnp.npz")
    xyz, uv, K = g["xyz"], g["uv"], g["K"]
    hyp, res = _run(xyz, uv, K, H)
    o = orc.pnp_ransac(xyz, uv, K, hyp[0].cpu().numpy())
    assert np.array_equal(res.hyp_counts[0].cpu().numpy(), o["counts"])            # every hypothesis, bit-exact
    assert int(res.best_h.item()) == o["best_h"] and int(res.n_inl.item()) == o["n_inl"]
    assert np.array_equal(res.mask[0, :len(xyz)].cpu().numpy(), o["mask"])         # inlier SET bit-exact
    assert int(res.status.item()) == 0
    rt = res.rt[0].cpu().numpy()
    assert _rot_angle(rt[:9].reshape(3, 3), o["rt"][:9].reshape(3, 3)) < 1e-9
    assert np.linalg.norm(rt[9:] - o["rt"][9:]) < 1e-9
    assert np.allclose(res.T_rel[0].cpu().numpy(), o["T_rel"], atol=1e-9)
    # rvec/tvec output is consistent with R|t
    import cv2
    rv = res.rvec_tvec[0].cpu().numpy()
    assert np.allclose(cv2.Rodrigues(rv[:3])[0], rt[:9].reshape(3, 3), atol=1e-12)
    # same inlier set fed to OpenCV's own refit (SURVEY: pose parity reduces to inlier-set parity)
    sel = np.nonzero(o["mask"])[0]
    _, rv2, tv2 = cv2.solvePnP(xyz[sel].astype(np.float64), uv[sel].astype(np.float64), K, None, flags=cv2.SOLVEPNP_ITERATIVE)
    assert _rot_angle(rt[:9].reshape(3, 3), cv2.Rodrigues(rv2)[0]) < ROT_TOL
    assert np.linalg.norm(rt[9:] - tv2.ravel()) < TRANS_TOL
    # and the reference's own call on the same correspondences (different sampler: its noise floor)
    R = cv2.Rodrigues(g["ransac_rvec"])[0]
    assert _rot_angle(rt[:9].reshape(3, 3), R) < 5e-4 and np.linalg.norm(rt[9:] - g["ransac_tvec"]) < 5e-3


def test_batch_of_scenes_vs_oracle_and_ground_truth(orc):
    import torch
    from vo_b200 import ops, synthetic
    B, N, H = 6, 2500, 512
    ps = [synthetic.make_pair(300 + b, n_kp=N, kind="orb") for b in range(B)]
    corr = []
    for p in ps:
        m, _ = orc.match_u8(p["ref_desc"], p["cur_desc"], orc.NORM_HAMMING, orc.MODE_MUTUAL)
        corr.append(orc.gather_backproject(m, p["ref_kp"], p["cur_kp"], p["depth"], p["K"]))
    cap = max(len(c[0]) for c in corr) + 5
    X = np.zeros((B, cap, 3), np.float32); U = np.zeros((B, cap, 2), np.float32)
    n_pts = np.zeros(B, np.int32)
    for b, c in enumerate(corr):
        n = len(c[0]) if b != 4 else 3                                   # pair 4: too few points
        X[b, :n], U[b, :n], n_pts[b] = c[0][:n], c[2][:n], n
    n_dev = _gpu(n_pts)
    hyp = ops.hypotheses(n_dev, H, 8214, 300)
    res = ops.pnp_ransac(_gpu(X), _gpu(U), n_dev, ps[0]["K"], hyp, 1.5, 20, 10, want_counts=True)
    st = res.status.cpu().numpy()
    for b, p in enumerate(ps):
        n = n_pts[b]
        if b == 4:
            assert st[b] == ops._lib.VO_ST_TOO_FEW_POINTS
            assert np.array_equal(res.T_rel[b].cpu().numpy(), np.eye(4))
            continue
        o = orc.pnp_ransac(X[b, :n], U[b, :n], p["K"], orc.hypotheses(n, H, 8214, 300 + b))
        assert np.array_equal(res.hyp_counts[b].cpu().numpy(), o["counts"])
        assert int(res.best_h[b].item()) == o["best_h"]
        assert np.array_equal(res.mask[b, :n].cpu().numpy(), o["mask"])
        assert st[b] == 0 and o["ok"]
        T = res.T_rel[b].cpu().numpy()
        ang, dt = synthetic.pose_errors(T, o["T_rel"])
        assert ang < 1e-9 and dt < 1e-9
        ang, dt = synthetic.pose_errors(T, p["T_rel"])                   # ground truth, 0.3 px noise
        assert ang < 2e-3 and dt < 2e-2


def test_no_model_on_pure_noise():
    rng = np.random.default_rng(0)
    xyz = rng.uniform(1, 30, (300, 3)).astype(np.float32)
    uv = rng.uniform(0, 1000, (300, 2)).astype(np.float32)
    K = np.array([[700.0, 0, 600], [0, 700.0, 180], [0, 0, 1]])
    hyp, res = _run(xyz, uv, K, 64)
    from vo_b200 import ops
    assert int(res.status.item()) == ops._lib.VO_ST_NO_MODEL
    assert int(res.n_inl.item()) <= 20 and int(res.mask.sum().item()) == 0
    assert np.array_equal(res.T_rel[0].cpu().numpy(), np.eye(4))


def test_points_beyond_one_staging_tile(orc):
    """> 2048 correspondences: the score kernel stages several shared-memory tiles."""
    from vo_b200 import synthetic
    p = synthetic.make_pair(17, n_kp=9000, kind="orb", outlier_frac=0.2)
    m, _ = orc.match_u8(p["ref_desc"], p["cur_desc"], orc.NORM_HAMMING, orc.MODE_MUTUAL)
    xyz, _, cuv, _, _ = orc.gather_backproject(m, p["ref_kp"], p["cur_kp"], p["depth"], p["K"])
    assert len(xyz) > 4200
    hyp, res = _run(xyz, cuv, p["K"], 256)
    o = orc.pnp_ransac(xyz, cuv, p["K"], hyp[0].cpu().numpy())
    assert np.array_equal(res.hyp_counts[0].cpu().numpy(), o["counts"])
    assert np.array_equal(res.mask[0, :len(xyz)].cpu().numpy(), o["mask"])


def _scene(rng, n, K, outliers=0.3):
    """n correspondences of one rigid motion (0.3 px noise) with a fraction of wrong image points."""
    u, v, z = rng.uniform(0, 1241, n), rng.uniform(0, 376, n), rng.uniform(4, 45, n)
    P = np.stack([(u - K[0, 2]) / K[0, 0] * z, (v - K[1, 2]) / K[1, 1] * z, z], 1)
    w = np.array([0.004, -0.006, 0.003])
    th = np.linalg.norm(w)
    k = w / th
    Kx = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    R = np.eye(3) + np.sin(th) * Kx + (1 - np.cos(th)) * Kx @ Kx
    Q = P @ R.T + np.array([0.02, -0.01, -0.6])
    q = np.stack([K[0, 0] * Q[:, 0] / Q[:, 2] + K[0, 2], K[1, 1] * Q[:, 1] / Q[:, 2] + K[1, 2]], 1) + rng.normal(0, 0.3, (n, 2))
    bad = rng.random(n) < outliers
    q[bad] = np.stack([rng.uniform(0, 1241, bad.sum()), rng.uniform(0, 376, bad.sum())], 1)
    return P.astype(np.float32), q.astype(np.float32)


def test_score_tiles_ragged_counts_and_pruning_is_exact(orc):
    """The FFMA2 score kernel reads 1024-point tiles in pairs of points: counts around the tile and pair boundaries
    (0, 3, 4, odd, 1023 / 1024 / 1025, 2049, 3071) in ONE ragged batch with a capacity that is no multiple of the tile.
    Per-hypothesis counts equal the oracle's, and the pruned run (no counts requested: hypotheses that cannot reach the
    running best stop early; hypotheses scored in the order of their tile-0 counts) returns the same winner, inlier
    count, mask and pose as the unpruned one.  A second, 4-tile capacity covers the unsorted pruned path."""
    import torch
    from vo_b200 import ops, synthetic
    K = synthetic.KITTI_K
    rng = np.random.default_rng(11)
    ns = [0, 3, 4, 5, 37, 1023, 1024, 1025, 2049, 3071, 6145]
    cap, H = 6200, 160      # 7 tiles: the pruned run takes the sorted path (tile-0 pass, hypothesis sort, carried counts)
    X = np.zeros((len(ns), cap, 3), np.float32)
    U = np.zeros((len(ns), cap, 2), np.float32)
    for b, n in enumerate(ns):
        X[b, :n], U[b, :n] = _scene(rng, n, K)
        X[b, n:], U[b, n:] = 7.0, 100.0          # garbage past the count must never be read as a point
    n_pts = torch.tensor(ns, dtype=torch.int32, device="cuda")
    hyp = ops.hypotheses(n_pts, H, 8214, 0)
    full = ops.pnp_ransac(_gpu(X), _gpu(U), n_pts, K, hyp, 1.5, 20, 10, want_counts=True)
    fast = ops.pnp_ransac(_gpu(X), _gpu(U), n_pts, K, hyp, 1.5, 20, 10, want_counts=False)
    for b, n in enumerate(ns):
        if n >= 4:
            o = orc.pnp_ransac(X[b, :n], U[b, :n], K, hyp[b].cpu().numpy())
            assert np.array_equal(full.hyp_counts[b].cpu().numpy(), o["counts"]), n
            assert np.array_equal(full.mask[b, :n].cpu().numpy(), o["mask"]), n
        else:
            assert int(full.status[b].item()) == ops._lib.VO_ST_TOO_FEW_POINTS
    for name in ("n_inl", "best_h", "status", "mask", "T_rel", "rt"):
        assert torch.equal(getattr(full, name), getattr(fast, name)), name
    keep = [i for i, n in enumerate(ns) if n <= 3100]
    Xs, Us = np.ascontiguousarray(X[keep, :3100]), np.ascontiguousarray(U[keep, :3100])
    n_s = n_pts[keep].contiguous()
    hyp_s = hyp[keep].contiguous()
    full_s = ops.pnp_ransac(_gpu(Xs), _gpu(Us), n_s, K, hyp_s, 1.5, 20, 10, want_counts=True)
    fast_s = ops.pnp_ransac(_gpu(Xs), _gpu(Us), n_s, K, hyp_s, 1.5, 20, 10, want_counts=False)
    for name in ("n_inl", "best_h", "status", "mask", "T_rel", "rt"):
        assert torch.equal(getattr(full_s, name), getattr(fast_s, name)), name
        assert torch.equal(getattr(full_s, name), getattr(full, name)[keep][..., :3100] if name == "mask" else getattr(full, name)[keep]), name
