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
    import vo_b200  # noqa: F401
    from vo_b200 import dropin
    return dropin.load(tmp_path, extractor, extra)


@pytest.mark.parametrize("kind,extractor,extra,matcher", [
    ("orb", "orb", "", "knn_ratio"),                                   # reference ORB semantics (L2 on bytes + ratio)
    ("orb", "orb", "\norb_matcher: hamming_mutual\n", "hamming_mutual"),   # north-star semantics
    ("sift", "sift", "", "knn_ratio"),
])
def test_process_frame_trajectory_vs_reference_loop(tmp_path, kind, extractor, extra, matcher):
    """Default drop-in (pnp_mode: reference = the reference's sampler, under the SAME numpy bootstrap stream) against the
    reference's loop (oracle/reference_vo.py: pinned pose for pose to the reference's own class by
    tests/golden/make_ref_trajectory.py) on a short sequence: same keyframe decisions, every pose within the reference's own
    per-pose noise floor, evaluator numbers inside the band eight reference seeds span."""
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
        assert vo.pnp_mode == "reference"
        refs = [ReferenceVO(synthetic.KITTI_K, matcher=matcher, seed=s) for s in (8214, 1234, 99, 7, 8, 9, 10, 11)]
        img = np.zeros((synthetic.KITTI_WH[1], synthetic.KITTI_WH[0], 3), np.uint8)
        np.random.seed(8214)                                           # vo_stereo_runner.py:20-24: the stream the bootstrap draws from
        ours, theirs, keys = [], [[] for _ in refs], []
        for i, f in enumerate(frames):
            feed["cur"] = (f["kp"], f["desc"])
            keys.append(vo.ref_data[-1].id if i else 0)
            ours.append(vo.process_frame(img, f["depth"], (100, 100), i).pose.copy())
            for r, out in zip(refs, theirs):
                out.append(r.process_frame(f["kp"], f["desc"], f["depth"], i).copy())
        ours, theirs = np.stack(ours), [np.stack(t) for t in theirs]
        assert len(vo.global_poses) == len(frames)
        e_ours = np.array(kitti_eval.evaluate(gt, ours)[:3])
        e_ref = np.array([kitti_eval.evaluate(gt, t)[:3] for t in theirs])
        assert e_ours[1] < 0.03 and e_ref[0][1] < 0.03                 # absolute accuracy: centimetres over ~16 m
        # per pose: ours vs the reference run on the same bootstrap seed, against how far two reference seeds are apart
        d = np.linalg.norm(ours[:, :3, 3] - theirs[0][:, :3, 3], axis=1)
        d_self = max(np.linalg.norm(theirs[k][:, :3, 3] - theirs[0][:, :3, 3], axis=1).max() for k in range(1, len(refs)))
        assert d.max() <= max(2.0 * d_self, 5e-3), (d.max(), d_self)
        # evaluator numbers: inside the band the reference's eight seeds span, widened by its width on either side (24 frames:
        # the band is wide; tests/test_gpu_trajectory_long.py is the 800-frame version against the reference's own class)
        lo, hi = e_ref.min(0), e_ref.max(0)
        pad = (hi - lo) + 1e-9
        assert np.all(e_ours >= lo - pad) and np.all(e_ours <= hi + pad), (e_ours, e_ref)
        print(f"[{kind}/{matcher}] ours {e_ours} ref seeds {e_ref.tolist()} max|dpos| {d.max():.4f} (ref self {d_self:.4f})")
    finally:
        os.chdir(cwd)


def test_throughput_mode_is_at_least_as_accurate(tmp_path):
    """pnp_mode: throughput (512 counter-based P3P hypotheses on the un-resampled set, refit on all inliers of the winner)
    is a different, lower-variance estimator than the reference's: its errors against ground truth must not exceed the
    reference loop's."""
    import vo_b200  # noqa: F401
    from vo_b200 import synthetic, synthetic_sequence
    from oracle import kitti_eval
    from oracle.reference_vo import ReferenceVO
    cwd = os.getcwd()
    try:
        vos = _load_dropin(tmp_path, "orb", "\npnp_mode: throughput\n")
        frames, gt = synthetic_sequence.make_sequence(n_frames=24, n_kp=1500, kind="orb", seed=77)
        feed = {}
        vos.extract_features_and_desc = lambda img: feed["cur"]
        vo = vos.VisualOdometry(synthetic.KITTI_K, seq=0)
        ref = ReferenceVO(synthetic.KITTI_K, matcher="knn_ratio")
        img = np.zeros((synthetic.KITTI_WH[1], synthetic.KITTI_WH[0], 3), np.uint8)
        ours, theirs = [], []
        for i, f in enumerate(frames):
            feed["cur"] = (f["kp"], f["desc"])
            ours.append(vo.process_frame(img, f["depth"], (100, 100), i).pose.copy())
            theirs.append(ref.process_frame(f["kp"], f["desc"], f["depth"], i).copy())
        e_ours, e_ref = kitti_eval.evaluate(gt, np.stack(ours)), kitti_eval.evaluate(gt, np.stack(theirs))
        assert e_ours[1] <= 1.1 * e_ref[1] and e_ours[0] <= 1.25 * e_ref[0], (e_ours, e_ref)
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
