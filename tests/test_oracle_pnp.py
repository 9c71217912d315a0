"""PnP oracle (hypothesis table, P3P, fp32 scoring, refit) against cv2 golden vectors."""
import numpy as np


def _rot_angle(Ra, Rb):
    return float(2.0 * np.arcsin(min(1.0, np.linalg.norm(Ra - Rb) / (2.0 * np.sqrt(2.0)))))


def test_hypothesis_table_properties(orc):
    hyp = orc.hypotheses(50, 4096, seed=8214, pair=3)
    assert hyp.min() >= 0 and hyp.max() < 50
    assert all(len(set(r)) == 4 for r in hyp.tolist())
    assert np.array_equal(hyp, orc.hypotheses(50, 4096, seed=8214, pair=3))        # deterministic
    assert not np.array_equal(hyp, orc.hypotheses(50, 4096, seed=8214, pair=4))
    assert abs(np.bincount(hyp.ravel(), minlength=50).std() / (4096 * 4 / 50)) < 0.1  # roughly uniform
    assert (orc.hypotheses(3, 8) == -1).all()                                        # too few points
    assert sorted(orc.hypotheses(4, 1)[0].tolist()) == [0, 1, 2, 3]


def test_p3p_solutions_match_cv2(golden, orc):
    g = golden("pnp.npz")
    xyz, uv, K = g["xyz"].astype(np.float64), g["uv"].astype(np.float64), g["K"]
    found = total = 0
    for q, t in enumerate(g["trip"]):
        ours = orc.p3p(xyz[t], uv[t], K)
        for k in range(int(g["p3p_nsol"][q])):
            Rc, tc = g["p3p_sols"][q, k, :9].reshape(3, 3), g["p3p_sols"][q, k, 9:]
            if not np.isfinite(g["p3p_sols"][q, k]).all():
                continue                       # cv2 emits NaN "solutions" for a degenerate triplet
            total += 1
            if any(_rot_angle(R, Rc) < 1e-6 and np.linalg.norm(tt - tc) < 1e-5 for R, tt in ours):
                found += 1
    assert total >= 50
    assert found >= total - 2          # Grunert vs OpenCV's P3P: identical solution sets up to conditioning


def test_inlier_rule_matches_cv2_projectpoints(golden, orc):
    g = golden("pnp.npz")
    mask = orc.inlier_mask(g["xyz"], g["uv"], g["K"], g["pose0"], 1.5)
    diff = np.nonzero(mask != g["mask_cv"])[0]
    # only points whose fp32 error straddles 2.25 by rounding may differ (projectPoints works in f64)
    assert all(abs(float(g["err_cv"][i]) - 2.25) < 1e-3 for i in diff)
    assert len(diff) <= 2 and mask.sum() > 100


def test_refit_matches_cv2_iterative(golden, orc):
    g = golden("pnp.npz")
    mask = np.zeros(len(g["xyz"]), np.uint8)
    mask[g["ransac_inliers"]] = 1
    rt = orc.refit(g["xyz"], g["uv"], mask, g["K"], g["pose0"], 10)
    want = g["refit_rt"]
    assert _rot_angle(rt[:9].reshape(3, 3), want[:9].reshape(3, 3)) < 1e-6     # north-star: 1e-4 rad
    assert np.linalg.norm(rt[9:] - want[9:]) < 1e-5                             # north-star: 1e-3 m


def test_full_ransac_agrees_with_cv2_solvepnpransac(golden, orc):
    from vo_b200 import synthetic
    g = golden("pnp.npz")
    xyz, uv, K = g["xyz"], g["uv"], g["K"]
    res = orc.pnp_ransac(xyz, uv, K, orc.hypotheses(len(xyz), 512, 8214, 0))
    assert res["ok"]
    R = __import__("cv2").Rodrigues(g["ransac_rvec"])[0]
    T = np.eye(4)
    T[:3, :3], T[:3, 3] = R.T, -R.T @ g["ransac_tvec"]
    ang, dt = synthetic.pose_errors(res["T_rel"], T)
    assert ang < 1e-4 * 5 and dt < 5e-3          # different samplers: the reference's own noise floor
    ang, dt = synthetic.pose_errors(res["T_rel"], g["T_gt"])
    assert ang < 1e-3 and dt < 1e-2
    inl_cv = set(g["ransac_inliers"].tolist())
    inl = set(np.nonzero(res["mask"])[0].tolist())
    assert len(inl & inl_cv) > 0.9 * len(inl_cv)


def test_too_few_points_and_no_model(orc):
    K = np.array([[700.0, 0, 600], [0, 700.0, 180], [0, 0, 1]])
    xyz = np.zeros((3, 3), np.float32)
    uv = np.zeros((3, 2), np.float32)
    assert not orc.pnp_ransac(xyz, uv, K, orc.hypotheses(3, 16))["ok"]
    rng = np.random.default_rng(0)
    xyz = rng.uniform(1, 30, (200, 3)).astype(np.float32)
    uv = rng.uniform(0, 1000, (200, 2)).astype(np.float32)        # pure noise: no model reaches 20 inliers
    res = orc.pnp_ransac(xyz, uv, K, orc.hypotheses(200, 64))
    assert not res["ok"] and res["n_inl"] <= 20
