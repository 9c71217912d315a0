"""The product's evaluator and pose-file formats (plot_utils/) against the reference's known
answer on its shipped KITTI-03 data, against the loop restatements in oracle/kitti_eval.py, and round trips of the
text formats the reference writes (plot_utils/prepare_data.py:8-27)."""
import os

import numpy as np
import pytest


def _dict(P):
    return {i: P[i] for i in range(len(P))}


def _to4(a):
    P = np.tile(np.eye(4), (len(a), 1, 1))
    P[:, :3, :] = np.asarray(a, np.float64).reshape(-1, 3, 4)
    return P


def _random_walk(n, seed, step=0.9):
    from vo_b200.synthetic import _rodrigues
    rng = np.random.default_rng(seed)
    T = [np.eye(4)]
    for _ in range(n - 1):
        s = np.eye(4)
        s[:3, :3] = _rodrigues(rng.normal(0, 0.01, 3))
        s[:3, 3] = [rng.normal(0, 0.03), rng.normal(0, 0.02), step + rng.normal(0, 0.1)]
        T.append(T[-1] @ s)
    return np.stack(T)


def test_eval_known_answer_kitti03(golden):
    import vo_b200  # noqa: F401
    from vo_b200.plot_utils.kittievalodom import KittiEvalOdom
    g = golden("kitti03_eval.npz")
    got = KittiEvalOdom().eval_poses(_dict(_to4(g["gt"])), _dict(_to4(g["pred"])), alignment="6dof")
    assert np.allclose(got, g["expected"], rtol=1e-12, atol=0)


def test_eval_methods_equal_loop_restatement():
    import vo_b200  # noqa: F401
    from vo_b200.plot_utils.kittievalodom import KittiEvalOdom
    from oracle import kitti_eval
    gt = _random_walk(1300, 1)                       # ~1.2 km: every segment length up to 800 m occurs
    noise = _random_walk(1300, 2, step=0.0)
    pred = gt.copy()
    pred[:, :3, 3] += 0.02 * np.cumsum(np.random.default_rng(3).normal(0, 1, (1300, 3)), axis=0)
    pred[:, :3, :3] = noise[:, :3, :3] @ gt[:, :3, :3]
    ev = KittiEvalOdom()
    want = kitti_eval.sequence_errors(gt, pred)
    got = ev.calc_sequence_errors(_dict(gt), _dict(pred))
    assert len(got) == len(want) > 300
    assert np.allclose(np.asarray(got, float), np.asarray(want, float), rtol=1e-9, atol=1e-12)
    assert [r[0] for r in got] == [r[0] for r in want] and [r[3] for r in got] == [r[3] for r in want]
    # helpers
    dist = ev.trajectory_distances(_dict(gt))
    for first, ln in ((0, 100), (10, 800), (1290, 100), (700, 300)):
        i = ev.last_frame_from_segment_length(dist, first, ln)
        ref = next((k for k in range(first, len(dist)) if dist[k] > dist[first] + ln), -1)
        assert i == ref
    # the 4-tuple of eval() equals the oracle restatement
    a = ev.eval_poses(_dict(gt), _dict(pred), alignment="6dof")
    b = kitti_eval.evaluate(gt, pred)
    assert np.allclose(a, b, rtol=1e-10, atol=1e-13)
    t, r = ev.compute_overall_err(got)
    assert t == pytest.approx(np.mean([x[2] for x in want])) and r == pytest.approx(np.mean([x[1] for x in want]))
    seg = ev.compute_segment_error(got)
    for ln in ev.lengths:
        sel = [x for x in want if x[3] == ln]
        assert seg[ln] == pytest.approx([np.mean([x[2] for x in sel]), np.mean([x[1] for x in sel])])
    # partial prediction (shorter than the ground truth) and a scale alignment
    short = _dict(pred[:400])
    assert len(ev.calc_sequence_errors(_dict(gt), short)) == len(kitti_eval.sequence_errors(gt, pred[:400]))
    scaled = {k: v.copy() for k, v in _dict(pred).items()}
    for v in scaled.values():
        v[:3, 3] *= 0.5
    s = ev.eval_poses(_dict(gt), scaled, alignment="scale")
    assert s[1] < 0.2 and ev.eval_poses(_dict(gt), scaled, alignment=None)[1] > 0.4


def test_eval_is_fast_on_a_long_sequence():
    import time
    import vo_b200  # noqa: F401
    from vo_b200.plot_utils.kittievalodom import KittiEvalOdom
    gt = _random_walk(20000, 5)
    t0 = time.perf_counter()
    out = KittiEvalOdom().eval_poses(_dict(gt), _dict(gt))
    assert time.perf_counter() - t0 < 5.0
    assert out[0] == 0.0 and out[1] < 1e-12


def test_pose_text_formats_round_trip(tmp_path):
    import vo_b200  # noqa: F401
    from vo_b200.plot_utils import prepare_data as pd
    from vo_b200.plot_utils.kittievalodom import KittiEvalOdom
    poses = _random_walk(40, 9)
    npy = str(tmp_path / "global_poses.npy")
    np.save(npy, poses)
    pd.prepare_data(npy)
    # byte-identical to the reference's per-pose np.savetxt loop (prepare_data.py:10-13)
    with open(npy + ".ref.txt", "w") as f:
        for p in poses:
            np.savetxt(f, p.reshape(1, 16))
    assert open(npy + ".txt").read() == open(npy + ".ref.txt").read()
    back = KittiEvalOdom().load_poses_from_txt(npy + ".txt")
    assert sorted(back) == list(range(40)) and np.array_equal(np.stack([back[i] for i in range(40)]), poses)
    # KITTI ground truth: 12 floats per line, no trailing newline (the shipped 03.txt) -> *_modified.txt
    gt_txt = str(tmp_path / "03.txt")
    with open(gt_txt, "w") as f:
        f.write("\n".join(" ".join("%.6e" % v for v in p[:3].reshape(-1)) for p in poses))
    pd.prepare_kitti_gt_data(gt_txt)
    mod = str(tmp_path / "03_modified.txt")
    lines = open(mod).read().split("\n")
    assert lines[-1] == "" and all(l.endswith(" 0.00 0.00 0.00 1.00") for l in lines[:-1]) and len(lines) == 41
    P, idx = pd.load_kitti_poses(mod)
    assert np.allclose(P, poses, rtol=1e-6, atol=1e-6) and np.array_equal(idx, np.arange(40))
    out = str(tmp_path / "w.txt")
    pd.write_kitti_poses(out, poses)
    P2, _ = pd.load_kitti_poses(out)
    assert np.array_equal(P2, poses)
    with open(out, "a") as f:
        f.write("1 2 3\n")
    with pytest.raises(ValueError):
        pd.load_kitti_poses(out)


def test_eval_segment_table_equals_reference_on_kitti03(golden):
    """calc_sequence_errors / compute_overall_err / compute_segment_error / compute_ATE / compute_RPE against the
    reference's own methods run on its shipped KITTI-03 data (tests/golden/make_golden.py: gen_kitti_eval)."""
    import vo_b200  # noqa: F401
    from vo_b200.plot_utils.kittievalodom import KittiEvalOdom
    from oracle import kitti_eval
    g, s = golden("kitti03_eval.npz"), golden("kitti03_segments.npz")
    gt, pred = _to4(g["gt"]), _to4(g["pred"])
    gt, pred = np.linalg.inv(gt[0]) @ gt, np.linalg.inv(pred[0]) @ pred
    ev = KittiEvalOdom()
    seq = ev.calc_sequence_errors(_dict(gt), _dict(pred))
    assert np.asarray(seq).shape == s["seq_err"].shape
    assert np.allclose(np.asarray(seq, float), s["seq_err"], rtol=1e-10, atol=1e-14)
    assert np.allclose(ev.compute_overall_err(seq), s["overall"], rtol=1e-12)
    seg = ev.compute_segment_error(seq)
    for ln, row in zip(ev.lengths, s["segments"]):
        assert (seg[ln] == [] and np.isnan(row).all()) or np.allclose(seg[ln], row, rtol=1e-12)
    assert ev.compute_ATE(_dict(gt), _dict(pred)) == pytest.approx(float(s["ate"]), rel=1e-13)
    assert np.allclose(ev.compute_RPE(_dict(gt), _dict(pred)), s["rpe"], rtol=1e-12)
    assert np.allclose(ev.trajectory_distances(_dict(gt)), s["dist"], rtol=0, atol=0)
    # and the oracle's loop restatement is pinned by the same table
    assert np.allclose(np.asarray(kitti_eval.sequence_errors(gt, pred), float), s["seq_err"], rtol=1e-10, atol=1e-14)


def test_frame_prefetcher_equals_sequential_loader(tmp_path):
    import cv2
    import vo_b200  # noqa: F401
    from vo_b200 import frame_io
    rng = np.random.default_rng(0)
    imgs, deps = [], []
    for i in (3, 0, 2, 1, 10, 4):                      # written out of order: the loader sorts by name
        img = rng.integers(0, 256, (24, 40, 3), dtype=np.uint8)
        dep = rng.uniform(0.5, 60.0, (24, 40)).astype(np.float32)
        cv2.imwrite(str(tmp_path / f"{i:06d}.png"), img)
        np.save(str(tmp_path / f"{i:06d}_depth.npy"), dep)
        imgs.append((i, img)); deps.append((i, dep))
    order = sorted(i for i, _ in imgs)
    got = list(frame_io.FramePrefetcher(str(tmp_path), ahead=3, workers=2))
    assert [g[0] for g in got] == list(range(6)) and len(frame_io.FramePrefetcher(str(tmp_path))) == 6
    for (idx, rgb, depth), name in zip(got, order):
        want_bgr = dict(imgs)[name]
        assert np.array_equal(rgb, want_bgr[:, :, ::-1]) and np.array_equal(depth, dict(deps)[name])
    # a broken frame raises at its position, like the sequential loop would
    (tmp_path / "000002.png").write_bytes(b"not a png")
    it = iter(frame_io.FramePrefetcher(str(tmp_path), ahead=2, workers=2))
    assert next(it)[0] == 0 and next(it)[0] == 1
    with pytest.raises(FileNotFoundError):
        next(it)
