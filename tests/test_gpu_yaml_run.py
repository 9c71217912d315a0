"""`python3 vo_runner.py` end to end with the ORB and the SIFT plug-ins in their default configuration — GPU extractor
(vo_orb_extract / vo_sift_extract), GPU matcher, GPU back-projection, the reference's own PnP sampler on the GPU — from png +
*_depth.npy files on disk to the (N,4,4) float64 pose file, with the REFERENCE's vo_params.yaml keys (vo_runner.py:6-16,
vo_stereo_runner.py:27-60).  Scene: a textured fronto-parallel plane, camera moving sideways, so the true trajectory is known
in closed form.  (tests/test_r2d2_frontend.py has the same run for feature_extractor: r2d2.)"""
import importlib
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("extractor", ["orb", "sift"])
def test_reference_yaml_offline_run(tmp_path, extractor):
    import cv2
    from vo_b200 import synthetic
    W, H = synthetic.KITTI_WH
    rng = np.random.default_rng(9)
    coarse = rng.integers(0, 256, (H // 6 + 2, (W + 100) // 6 + 2, 3)).astype(np.float32)
    big = cv2.resize(coarse, None, fx=6, fy=6, interpolation=cv2.INTER_CUBIC).clip(0, 255).astype(np.uint8)   # smooth texture: corners and blobs
    data = tmp_path / "frames"
    data.mkdir()
    Z, shift, n = 6.0, 12, 5
    for i in range(n):
        cv2.imwrite(str(data / f"{i:06d}.png"), np.ascontiguousarray(big[:H, i * shift:i * shift + W]))
        np.save(str(data / f"{i:06d}_depth.npy"), np.full((H, W), Z, np.float32))
    (tmp_path / "config").mkdir()
    (tmp_path / "config" / "vo_params.yaml").write_text(
        f'vo_method: "rgbd"\nfeature_extractor: "{extractor}"\n'
        f'image_path: "{data}"\n'
        "camera_intrinsic_matrix:\n" + "".join(f"  - {v}\n" for v in synthetic.KITTI_K.reshape(-1)) +
        f"output_filename: {tmp_path}/global_poses\nvisualize_results: True\n"
        'gt_txt_file_path : "../plot_utils/data/03.txt"\nposes_file_path : "../plot_utils/data/global_poses.npy"\n')
    pkg = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "visual-odometry-pipeline_b200")
    cwd = os.getcwd()
    try:
        os.chdir(tmp_path)
        if pkg not in sys.path:
            sys.path.insert(0, pkg)
        for m in ("VisualOdometry_Stereo", "vo_stereo_runner", "vo_runner", "feature_extractors.ORB", "feature_extractors.SIFT"):
            sys.modules.pop(m, None)
        runner = importlib.import_module("vo_runner")
        plug = sys.modules[f"feature_extractors.{extractor.upper()}"]
        assert plug.EXTRACTOR == "gpu"                       # the plug-ins extract on the device by default
        runner.read_yaml_file()
        poses = np.load(str(tmp_path / "global_poses.npy"))
        assert poses.shape == (n, 4, 4) and poses.dtype == np.float64
        step = shift * Z / synthetic.KITTI_K[0, 0]
        for i in range(n):
            assert np.allclose(poses[i][:3, :3], np.eye(3), atol=5e-3), i
            assert np.allclose(poses[i][:3, 3], [i * step, 0, 0], atol=0.02), (i, poses[i][:3, 3])
    finally:
        os.chdir(cwd)
