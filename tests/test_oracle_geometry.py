"""Back-projection / gather restatement against the reference's own unprojection helper and by construction."""
import numpy as np


def test_dense_backprojection_matches_unprojection_kp(golden, orc):
    g = golden("backproject.npz")
    depth, K, kp = g["depth"], g["K"], g["kp"]
    dense = orc.backproject_dense(depth, K)
    ui, vi = kp[:, 0].astype(np.int32), kp[:, 1].astype(np.int32)
    got = dense[vi, ui].astype(np.float64)
    want = g["xyz_f64"]                                   # Utils/geom_utils.py:55-77 in f64
    ok = np.isfinite(want).all(1)
    assert ok.sum() > 150
    assert np.allclose(got[ok], want[ok], rtol=3e-6, atol=1e-6)   # fp32 vs f64: a few ulp
    assert np.array_equal(dense[..., 2], depth, equal_nan=True)
    assert dense.dtype == np.float32 and dense.shape == depth.shape + (3,)


def test_gather_gate_filter_and_order(golden, orc):
    g = golden("backproject.npz")
    depth, K, kp = g["depth"], g["K"], g["kp"]
    n = len(kp)
    rng = np.random.default_rng(1)
    cur = kp + rng.normal(0, 4.0, kp.shape).astype(np.float32)
    cur[:20] = kp[:20] + 0.5                                # flow < 3 px -> dropped (VisualOdometry_Stereo.py:263)
    pairs = np.stack([np.arange(n), rng.permutation(n)], 1).astype(np.int32)
    cur_kp = np.zeros_like(cur)
    cur_kp[pairs[:, 1]] = cur
    xyz, ruv, cuv, src, oob = orc.gather_backproject(pairs, kp, cur_kp, depth, K)
    assert not oob
    # literal numpy restatement of :257-264 and :96-105
    r = kp[pairs[:, 0], :2].astype(np.float32)
    c = cur_kp[pairs[:, 1], :2].astype(np.float32)
    diff = np.linalg.norm(r - c, axis=1)
    r, c = r[diff >= 3], c[diff >= 3]
    idx = np.nonzero(diff >= 3)[0]
    dense = orc.backproject_dense(depth, K)
    p3 = dense[r[:, 1].astype(np.int32), r[:, 0].astype(np.int32)]
    good = (p3[:, 2] > 0) & (p3[:, 2] < 50)
    assert np.array_equal(xyz, p3[good])
    assert np.array_equal(ruv, r[good]) and np.array_equal(cuv, c[good])
    assert np.array_equal(src, idx[good])
    assert (diff[:20] < 3).all() and not np.isin(np.arange(20), src).any()


def test_out_of_image_keypoint_is_flagged(orc):
    depth = np.ones((8, 8), np.float32)
    K = np.array([[10.0, 0, 4], [0, 10.0, 4], [0, 0, 1]])
    kp = np.array([[2.5, 3.5], [8.2, 1.0]], np.float32)
    cur = kp + 5
    pairs = np.array([[0, 0], [1, 1]], np.int32)
    xyz, *_, oob = orc.gather_backproject(pairs, kp, cur, depth, K)
    assert oob and len(xyz) == 1
