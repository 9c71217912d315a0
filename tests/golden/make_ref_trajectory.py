#!/usr/bin/env python
"""Trajectory oracle from the reference's OWN class: /root/reference/VisualOdometry_Stereo.py `VisualOdometry`
(process_frame :223-297, computepose_3D_2D :87-149) and its own feature_extractors/{ORB,SIFT}.get_matches, imported
unmodified from a temporary copy with the three shims of SURVEY 8(c) (skimage / matplotlib stubs, cv2.xfeatures2d,
cv2.rgbd.depthTo3d restatement) and `visualize_results: False`.  Feature extraction is stubbed with the precomputed
keypoints / descriptors of a synthetic sequence (synthetic_sequence.make_long_sequence), exactly as the GPU tests stub it.

Runs the reference loop with S bootstrap seeds (np.random.seed(8214 + s), vo_stereo_runner.py:20-24 seeds 8214) over an
N-frame sequence and records, per seed, the reference evaluator's numbers (plot_utils/kittievalodom.py eval quantities:
ATE, RPE trans, RPE rot, mean segment errors) against the synthetic ground truth, the keyframe ids, and the trajectory
of seed 0.  Also checks that the restated loop oracle/reference_vo.py reproduces the real class pose for pose.

Build container only (needs /root/reference).   python tests/golden/make_ref_trajectory.py [n_frames] [n_seeds]
-> tests/golden/ref_traj_<kind>.npz"""
import importlib
import os
import shutil
import sys
import tempfile
import types

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

import cv2  # noqa: E402

import vo_b200  # noqa: E402,F401
from vo_b200 import synthetic, synthetic_sequence  # noqa: E402
from oracle import kitti_eval, reference_path  # noqa: E402
from oracle.reference_vo import ReferenceVO  # noqa: E402

SEQ = dict(n_kp=1500, seed=303)


def import_reference(extractor):
    """The real reference modules from a writable copy (the class mkdirs in CWD and reads config/vo_params.yaml at import)."""
    tmp = tempfile.mkdtemp(prefix="ref_vo_")
    for name in ("VisualOdometry_Stereo.py", "Utils", "feature_extractors", "config"):
        src = os.path.join(REF, name)
        (shutil.copytree if os.path.isdir(src) else shutil.copy)(src, os.path.join(tmp, name))
    cfg = os.path.join(tmp, "config", "vo_params.yaml")
    txt = open(cfg).read()
    import re
    txt = re.sub(r'feature_extractor:\s*"?\w+"?', f'feature_extractor: "{extractor}"', txt)
    txt = re.sub(r"visualize_results:\s*\w+", "visualize_results: False", txt)
    open(cfg, "w").write(txt)
    for mod in ("skimage", "skimage.exposure", "skimage.util", "skimage.util.shape", "matplotlib", "matplotlib.pyplot"):
        sys.modules.setdefault(mod, types.ModuleType(mod))
    if not hasattr(cv2, "xfeatures2d"):
        cv2.xfeatures2d = types.SimpleNamespace(SIFT_create=cv2.SIFT_create)
    if not hasattr(cv2, "rgbd"):
        cv2.rgbd = types.SimpleNamespace(depthTo3d=lambda depth, K: reference_path.depth_to_3d(depth, np.asarray(K)))
    os.chdir(tmp)
    sys.path.insert(0, tmp)
    for name in list(sys.modules):
        if name.split(".")[0] in ("VisualOdometry_Stereo", "Utils", "feature_extractors"):
            del sys.modules[name]
    return importlib.import_module("VisualOdometry_Stereo"), tmp


def run_reference(vos, frames, seed):
    np.random.seed(seed)                                  # the only RNG stream on the path (bootstrap, :122)
    feed = {}
    vos.extract_features_and_desc = lambda img: feed["cur"]
    vo = vos.VisualOdometry(synthetic.KITTI_K, seq=0)
    img = np.zeros((synthetic.KITTI_WH[1], synthetic.KITTI_WH[0], 3), np.uint8)
    poses, keys = [], []
    for i, f in enumerate(frames):
        feed["cur"] = (f["kp"], f["desc"])
        keys.append(vo.ref_data[-1].id if i else 0)
        poses.append(np.array(vo.process_frame(img, f["depth"], (100, 100), i).pose, np.float64).copy())
    return np.stack(poses), np.asarray(keys)


def main():
    n_frames = int(sys.argv[1]) if len(sys.argv) > 1 else 800
    n_seeds = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    kind = sys.argv[3] if len(sys.argv) > 3 else "orb"
    cwd = os.getcwd()
    frames, gt = synthetic_sequence.make_long_sequence(n_frames=n_frames, kind=kind, **SEQ)
    vos, tmp = import_reference(kind)
    import contextlib
    import io
    evals, keys_all, poses0 = [], [], None
    for s in range(n_seeds):
        with contextlib.redirect_stdout(io.StringIO()):
            poses, keys = run_reference(vos, frames, 8214 + s)
        e = kitti_eval.evaluate(gt, poses)
        evals.append(e[:5])
        keys_all.append(keys)
        if s == 0:
            poses0 = poses
            port = ReferenceVO(synthetic.KITTI_K, matcher="knn_ratio", seed=8214)   # RandomState(8214) == np.random.seed(8214) stream
            pp = np.stack([port.process_frame(f["kp"], f["desc"], f["depth"], i).copy() for i, f in enumerate(frames)])
            print("restated loop vs the real class: max |dpose| =", np.abs(pp - poses).max())
        print(f"seed {8214 + s}: ATE {e[0]:.5f}  RPE_t {e[1]:.6f}  RPE_r {e[2]:.7f}  keyframes {len(set(keys.tolist()))}", flush=True)
    os.chdir(cwd)
    shutil.rmtree(tmp, ignore_errors=True)
    evals = np.asarray(evals, np.float64)
    out = os.path.join(HERE, f"ref_traj_{kind}.npz")
    np.savez_compressed(out, evals=evals, poses_seed0=poses0, keys_seed0=keys_all[0], n_frames=n_frames, n_kp=SEQ["n_kp"],
                        seq_seed=SEQ["seed"], seeds=8214 + np.arange(n_seeds), port_max_dpose=np.abs(pp - poses0).max())
    print("mean", evals.mean(0), "\nstd/mean", evals.std(0) / evals.mean(0), "\nmin", evals.min(0), "\nmax", evals.max(0))
    print(out, os.path.getsize(out) / 1024, "KiB")


if __name__ == "__main__":
    main()
