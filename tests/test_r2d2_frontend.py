"""R2D2 front-end (vo_r2d2_*) against the reference's own network, NMS and score filter run on CPU
in fp32 with the shipped faster2d2_WASF_N16 weights (tests/golden/r2d2_net.npz, made by make_golden.py: gen_r2d2_net).
Floating point: maps within 2e-4, descriptors within 2e-4; the keypoint set is identical except for pixels the
reference itself decides within that tolerance of a threshold or of a 3x3 tie."""
import numpy as np
import pytest


def _weights(g):
    return str(g["net"]).split("(")[0], {k[3:]: g[k] for k in g.files if k.startswith("w__")}


def test_architecture_table_matches_the_checkpoint(golden):
    import vo_b200  # noqa: F401
    from vo_b200 import r2d2_frontend as rf
    g = golden("r2d2_net.npz")
    name, sd = _weights(g)
    assert name == "Fast_Quad_L2Net_ConfCFS"
    layers = rf.layer_table(name, sd)
    assert [(L["cin"], L["cout"], L["k"], L["dil"], L["bn"], L["relu"], L["pool_after"]) for L in layers] == [
        (3, 32, 3, 1, 1, 1, 0), (32, 32, 3, 1, 1, 1, 0), (32, 64, 3, 1, 1, 1, 2), (64, 64, 3, 1, 1, 1, 0), (64, 128, 3, 1, 1, 1, 0),
        (128, 128, 3, 2, 1, 1, 0), (128, 128, 2, 2, 1, 0, 0), (128, 128, 2, 4, 1, 0, 0), (128, 128, 2, 8, 0, 0, 0)]
    assert layers[0]["w"].shape == (32, 3, 3, 3) and layers[6]["w"].shape == (128, 2, 2, 128)
    # layout: [C_out][ky][kx][C_in] is the transpose of torch's [C_out][C_in][ky][kx]
    assert np.array_equal(layers[5]["w"][7, 2, 0, 33], sd["ops.16.weight"][7, 33, 2, 0])
    with pytest.raises(ValueError):
        rf.layer_table("L2_Net", sd)
    with pytest.raises(ValueError):
        rf.layer_table("Quad_L2Net_ConfCFS", {k: v for k, v in sd.items() if not k.startswith("ops.23")})


@pytest.mark.gpu
def test_maps_keypoints_descriptors_vs_reference_network(golden):
    import torch
    import vo_b200  # noqa: F401
    from vo_b200 import r2d2_frontend as rf
    g = golden("r2d2_net.npz")
    name, sd = _weights(g)
    img = g["image"]
    net = rf.R2D2Net(name, sd, img.shape[0], img.shape[1], max_kp=4096)
    assert (net.Ho, net.Wo) == g["rel"].shape == (96, 162)
    xys, desc, scores, rel, rep = net.extract(img, 0.7, 0.7, 0.85, want_maps=True)
    rel, rep = rel.cpu().numpy(), rep.cpu().numpy()
    assert np.abs(rel - g["rel"]).max() < 2e-4 and np.abs(rep - g["rep"]).max() < 2e-4
    keep = g["keep"]
    want_xy = g["xys_all"][keep][:, :2].astype(np.int64)
    want = {(int(x), int(y)): i for i, (x, y) in zip(keep, want_xy)}
    got_xy = xys.cpu().numpy()
    assert np.all(got_xy[:, 2] == 32.0)
    got = {(int(x), int(y)): i for i, (x, y) in enumerate(got_xy[:, :2])}
    # row-major order, like torch.nonzero
    lin = got_xy[:, 1].astype(np.int64) * net.Wo + got_xy[:, 0].astype(np.int64)
    assert np.all(np.diff(lin) > 0)
    tol = 3e-4
    for (x, y) in set(got) ^ set(want):                    # a differing pixel must sit on a decision boundary
        c, q = float(g["rel"][y, x]), float(g["rep"][y, x])
        nb = g["rep"][max(0, y - 1):y + 2, max(0, x - 1):x + 2]
        second = np.sort(nb.reshape(-1))[-2] if nb.size > 1 else -1
        near = abs(c - 0.7) < tol or abs(q - 0.7) < tol or abs(c * q - 0.85) < tol or abs(q - nb.max()) < tol and abs(q - second) < tol
        assert near, (x, y, c, q)
    common = sorted(set(got) & set(want))
    assert len(common) >= len(want) - 2 and len(common) > 50
    gi = np.array([got[k] for k in common]); wi = np.array([want[k] for k in common])
    d = desc.cpu().numpy()
    assert np.abs(d[gi] - g["desc_all"][wi]).max() < 2e-4
    assert np.abs(np.linalg.norm(d, axis=1) - 1.0).max() < 1e-5
    assert np.abs(scores.cpu().numpy()[gi] - g["scores_all"][wi]).max() < 3e-4
    # lower thresholds: more keypoints, still the reference's rule (checked against its maps)
    xys2, _, sc2 = net.extract(img, 0.5, 0.5, 0.3)
    assert len(xys2) > len(xys)
    # pinned host input gives the same result as a device tensor
    a = net.extract(torch.from_numpy(img).pin_memory(), 0.7, 0.7, 0.85)[0].clone()
    b = net.extract(torch.from_numpy(img).cuda(), 0.7, 0.7, 0.85)[0].clone()
    assert torch.equal(a, b) and torch.equal(a, xys)


@pytest.mark.gpu
def test_kitti_sized_image_runs_and_is_deterministic(golden):
    import torch
    import vo_b200  # noqa: F401
    from vo_b200 import r2d2_frontend as rf
    g = golden("r2d2_net.npz")
    name, sd = _weights(g)
    H, W = 376, 1241
    rng = np.random.default_rng(1)
    img = np.kron(rng.integers(0, 256, (H // 8, W // 8 + 1, 3)), np.ones((8, 8, 1)))[:H, :W].astype(np.uint8)
    net = rf.R2D2Net(name, sd, H, W)
    assert (net.Ho, net.Wo) == (376, 1240)
    xys, desc, scores = net.extract(img, 0.7, 0.7, 0.85)
    n = len(xys)
    first = (xys.clone(), desc.clone())
    xys_b, desc_b, _ = net.extract(img, 0.7, 0.7, 0.85)
    assert len(xys_b) == n and torch.equal(first[0], xys_b) and torch.equal(first[1], desc_b)
    assert n > 100 and float(scores.min()) > 0.85
    assert float((desc.norm(dim=1) - 1).abs().max()) < 1e-5
    assert float(xys[:, 0].max()) < 1240 and float(xys[:, 1].max()) < 376


@pytest.mark.gpu
def test_dropin_module_extracts_and_matches(golden, tmp_path):
    """The reference's module-level interface: R2D2.extract_features_and_desc(image BGR) -> (kps (N,3) numpy, desc (N,128)
    CUDA tensor), fed into R2D2.get_matches — with a checkpoint FILE in the reference's format (R2D2.py:68-79)."""
    import importlib
    import os
    import sys
    import torch
    g = golden("r2d2_net.npz")
    name, sd = _weights(g)
    ckpt = tmp_path / "model.pt"
    state = {"module." + k: torch.from_numpy(v) for k, v in sd.items()}
    state["module.ops.1.num_batches_tracked"] = torch.tensor(0)
    torch.save({"net": name + "()", "state_dict": state}, str(ckpt))
    pkg = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "visual-odometry-pipeline_b200")
    if pkg not in sys.path:
        sys.path.insert(0, pkg)
    sys.modules.pop("R2D2", None)
    r2 = importlib.import_module("R2D2")
    r2.args["model"] = str(ckpt)
    img_rgb = g["image"]
    bgr = np.ascontiguousarray(img_rgb[:, :, ::-1])
    kps, desc = r2.extract_features_and_desc(bgr)
    keep = g["keep"]
    assert kps.dtype == np.float32 and kps.shape[1] == 3 and isinstance(desc, torch.Tensor) and desc.is_cuda
    assert abs(len(kps) - len(keep)) <= 2 and desc.shape == (len(kps), 128)
    # a shifted copy of the image: the plug-in's own matcher pairs most keypoints with their shifted selves
    shifted = np.roll(bgr, 4, axis=1)
    kps2, desc2 = r2.extract_features_and_desc(shifted)
    m = r2.get_matches(kps, desc, kps2, desc2, bgr.shape)
    assert m.dtype == np.int64 and m.shape[1] == 2 and len(m) > 0.5 * len(kps)
    dx = kps2[m[:, 1], 0] - kps[m[:, 0], 0]
    assert np.mean(np.abs(dx - 4) < 1.5) > 0.8
    r2.args["model"] = "/nonexistent/model.pt"
    r2._weights = None; r2._nets.clear()
    with pytest.raises(RuntimeError):
        r2.extract_features_and_desc(bgr)


@pytest.mark.gpu
def test_yaml_run_with_r2d2_recovers_camera_motion(golden, tmp_path):
    """feature_extractor: r2d2 through the drop-in VisualOdometry.process_frame — images in, poses out, every stage on
    the GPU (network, matcher, back-projection, PnP).  Scene: a textured fronto-parallel plane 5 m away, camera moving
    sideways, so consecutive frames are shifted copies and the true trajectory is known in closed form."""
    import os
    import sys
    import torch
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__))))
    from test_gpu_dropin import _load_dropin
    from vo_b200 import synthetic
    g = golden("r2d2_net.npz")
    name, sd = _weights(g)
    ckpt = tmp_path / "model.pt"
    torch.save({"net": name + "()", "state_dict": {"module." + k: torch.from_numpy(v) for k, v in sd.items()}}, str(ckpt))
    cwd = os.getcwd()
    try:
        sys.modules.pop("R2D2", None)
        vos = _load_dropin(tmp_path, "r2d2")
        sys.modules["R2D2"].args["model"] = str(ckpt)
        W, H = synthetic.KITTI_WH
        rng = np.random.default_rng(3)
        big = np.kron(rng.integers(0, 256, (H // 6 + 1, (W + 200) // 6 + 1, 3)), np.ones((6, 6, 1))).astype(np.uint8)
        Z, shift = 5.0, 14                                       # pixels per frame -> metres: shift * Z / fx
        step = shift * Z / synthetic.KITTI_K[0, 0]
        depth = np.full((H, W), Z, np.float32)
        vo = vos.VisualOdometry(synthetic.KITTI_K, seq=0)
        poses = []
        for i in range(5):
            frame = np.ascontiguousarray(big[:H, i * shift:i * shift + W])    # camera moved +x: the scene slides left
            poses.append(vo.process_frame(frame, depth, (100, 100), i).pose.copy())
        poses = np.stack(poses)
        assert vo.bad_pnp == 0
        for i in range(5):
            assert np.allclose(poses[i][:3, :3], np.eye(3), atol=2e-3), i
            assert np.allclose(poses[i][:3, 3], [i * step, 0.0, 0.0], atol=0.01), (i, poses[i][:3, 3])
    finally:
        os.chdir(cwd)


def _torch_net(name, sd, dev):
    """The architecture of nets/patchnet.py rebuilt from plain torch modules (fp32 reference for a floating-point kernel)."""
    import torch
    import torch.nn as nn
    from vo_b200 import r2d2_frontend as rf
    mods = []
    for L in rf.layer_table(name, sd):
        conv = nn.Conv2d(L["cin"], L["cout"], L["k"], padding=((L["k"] - 1) * L["dil"]) // 2, dilation=L["dil"])
        conv.weight.data = torch.from_numpy(L["w"].transpose(0, 3, 1, 2).copy())
        conv.bias.data = torch.from_numpy(L["bias"].copy())
        mods.append(conv)
        if L["bn"]:
            bn = nn.BatchNorm2d(L["cout"], affine=False)
            bn.running_mean.data = torch.from_numpy(L["bn_mean"].copy())
            bn.running_var.data = torch.from_numpy(L["bn_var"].copy())
            mods.append(bn)
        if L["relu"]:
            mods.append(nn.ReLU())
        if L["pool_after"]:
            mods.append(nn.MaxPool2d(2))
    if rf.ARCH[name]["upsample"] == 2:
        mods.append(nn.Upsample(scale_factor=2, mode="bilinear", align_corners=False))
    return nn.Sequential(*mods).eval().double().to(dev)


@pytest.mark.gpu
def test_full_resolution_architecture_vs_torch_fp64():
    """Quad_L2Net_ConfCFS (the non-'faster' models: no pooling, no up-sampling, 2x2 taps up to dilation 16) with random
    weights against the same stack in torch fp64: covers the upsample == 1 head path and the widest dilations."""
    import torch
    import torch.nn.functional as F
    import vo_b200  # noqa: F401
    from vo_b200 import r2d2_frontend as rf
    rng = np.random.default_rng(5)
    name = "Quad_L2Net_ConfCFS"
    chans = [3, 32, 32, 64, 64, 128, 128, 128, 128, 128]
    sd, idx = {}, 0
    for li, (k, dil, bn, relu, pool) in enumerate(rf.ARCH[name]["layers"]):
        cin, cout = chans[li], chans[li + 1]
        sd[f"ops.{idx}.weight"] = (rng.standard_normal((cout, cin, k, k)) / np.sqrt(cin * k * k)).astype(np.float32)
        sd[f"ops.{idx}.bias"] = (0.1 * rng.standard_normal(cout)).astype(np.float32)
        if bn:
            sd[f"ops.{idx + 1}.running_mean"] = (0.1 * rng.standard_normal(cout)).astype(np.float32)
            sd[f"ops.{idx + 1}.running_var"] = rng.uniform(0.5, 1.5, cout).astype(np.float32)
        idx += 1 + bn + relu
    sd["clf.weight"] = (rng.standard_normal((2, 128, 1, 1)) * 0.3).astype(np.float32)
    sd["clf.bias"] = rng.standard_normal(2).astype(np.float32) * 0.1
    sd["sal.weight"] = (rng.standard_normal((1, 128, 1, 1)) * 0.3).astype(np.float32)
    sd["sal.bias"] = rng.standard_normal(1).astype(np.float32) * 0.1
    H, W = 70, 150
    img = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
    net = rf.R2D2Net(name, sd, H, W, max_kp=H * W)
    assert (net.Ho, net.Wo) == (H, W)
    xys, desc, scores, rel, rep = net.extract(img, 0.0, 0.0, -1.0, want_maps=True)     # thresholds off: every local maximum
    dev = torch.device("cuda")
    body = _torch_net(name, sd, dev)
    mean = torch.tensor([0.485, 0.456, 0.406], device=dev).view(1, 3, 1, 1)
    std = torch.tensor([0.229, 0.224, 0.225], device=dev).view(1, 3, 1, 1)
    x = ((torch.from_numpy(img).to(dev).permute(2, 0, 1)[None].float() / 255 - mean) / std).double()
    with torch.no_grad():
        f = body(x)
        u = F.conv2d(f ** 2, torch.from_numpy(sd["clf.weight"]).double().to(dev), torch.from_numpy(sd["clf.bias"]).double().to(dev))
        rel_t = F.softmax(u, dim=1)[0, 1]
        s = F.softplus(F.conv2d(f ** 2, torch.from_numpy(sd["sal.weight"]).double().to(dev), torch.from_numpy(sd["sal.bias"]).double().to(dev)))
        rep_t = (s / (1 + s))[0, 0]
        d_t = F.normalize(f, p=2, dim=1)[0]
    assert float((rel.double() - rel_t).abs().max()) < 2e-4 and float((rep.double() - rep_t).abs().max()) < 2e-4
    xi, yi = xys[:, 0].long(), xys[:, 1].long()
    assert len(xys) > 200
    assert float((desc.double() - d_t[:, yi, xi].t()).abs().max()) < 2e-4
    # every reported keypoint is a 3x3 local maximum of the repeatability map the library itself produced
    mx = F.max_pool2d(rep[None, None], 3, 1, 1)[0, 0]
    assert bool((rep[yi, xi] == mx[yi, xi]).all()) and int((rep == mx).sum()) == len(xys)


@pytest.mark.gpu
def test_capacity_overflow_and_bad_configs(golden):
    """More keypoints than max_kp: the first max_kp in row-major order are kept and the total is still reported; broken
    configurations fail loudly at creation."""
    import torch
    import vo_b200  # noqa: F401
    from vo_b200 import _lib, r2d2_frontend as rf
    g = golden("r2d2_net.npz")
    name, sd = _weights(g)
    img = g["image"]
    big = rf.R2D2Net(name, sd, img.shape[0], img.shape[1], max_kp=4096)
    xa, da, _ = big.extract(img, 0.3, 0.3, 0.1)
    xa, da = xa.clone(), da.clone()
    assert len(xa) > 60
    small = rf.R2D2Net(name, sd, img.shape[0], img.shape[1], max_kp=50)
    xb, db, _ = small.extract(img, 0.3, 0.3, 0.1)
    assert len(xb) == 50 and int(small.count.item()) == len(xa)
    assert torch.equal(xb, xa[:50]) and torch.equal(db, da[:50])
    with pytest.raises(ValueError):
        small.extract(img[:-1])                                   # wrong shape
    bad = dict(sd)
    bad["ops.3.weight"] = sd["ops.3.weight"][:, :16]              # C_in no longer matches the previous C_out
    with pytest.raises(_lib.VoError):
        rf.R2D2Net(name, bad, 64, 64)
    # a 1-row image and an image narrower than one 128-pixel tile still run
    tiny = rf.R2D2Net(name, sd, 2, 40, max_kp=64)
    x, d, s = tiny.extract(np.zeros((2, 40, 3), np.uint8), 0.0, 0.0, -1.0)
    assert tiny.Ho == 2 and tiny.Wo == 40 and d.shape[1] == 128


@pytest.mark.gpu
def test_reference_yaml_offline_run(golden, tmp_path):
    """`python3 vo_runner.py` with the REFERENCE's own vo_params.yaml keys (feature_extractor: r2d2, visualize_results:
    True, ...): png + *_depth.npy files on disk -> vo_runner.read_yaml_file() -> <output>.npy of (N,4,4) float64 poses,
    which then goes through prepare_data and the evaluator like the reference's plot_utils workflow."""
    import importlib
    import os
    import sys
    import cv2
    import torch
    from vo_b200 import synthetic
    g = golden("r2d2_net.npz")
    name, sd = _weights(g)
    ckpt = tmp_path / "model.pt"
    torch.save({"net": name + "()", "state_dict": {"module." + k: torch.from_numpy(v) for k, v in sd.items()}}, str(ckpt))
    W, H = synthetic.KITTI_WH
    rng = np.random.default_rng(9)
    big = np.kron(rng.integers(0, 256, (H // 6 + 1, (W + 100) // 6 + 1, 3)), np.ones((6, 6, 1))).astype(np.uint8)
    data = tmp_path / "frames"
    data.mkdir()
    Z, shift, n = 6.0, 12, 4
    for i in range(n):
        cv2.imwrite(str(data / f"{i:06d}.png"), np.ascontiguousarray(big[:H, i * shift:i * shift + W]))
        np.save(str(data / f"{i:06d}_depth.npy"), np.full((H, W), Z, np.float32))
    (tmp_path / "config").mkdir()
    (tmp_path / "config" / "vo_params.yaml").write_text(
        'vo_method: "rgbd"\nfeature_extractor: "r2d2"\n'
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
        for m in ("VisualOdometry_Stereo", "vo_stereo_runner", "vo_runner", "R2D2"):
            sys.modules.pop(m, None)
        runner = importlib.import_module("vo_runner")
        sys.modules["R2D2"].args["model"] = str(ckpt)
        runner.read_yaml_file()
        poses = np.load(str(tmp_path / "global_poses.npy"))
        assert poses.shape == (n, 4, 4) and poses.dtype == np.float64
        step = shift * Z / synthetic.KITTI_K[0, 0]
        for i in range(n):
            assert np.allclose(poses[i][:3, 3], [i * step, 0, 0], atol=0.01), (i, poses[i][:3, 3])
        # the reference's evaluation workflow on the produced file (plot_utils/prepare_data.py + kittievalodom.py)
        from vo_b200.plot_utils import prepare_data as pd
        from vo_b200.plot_utils.kittievalodom import KittiEvalOdom
        pd.prepare_data(str(tmp_path / "global_poses.npy"))
        ev = KittiEvalOdom()
        pred = ev.load_poses_from_txt(str(tmp_path / "global_poses.npy.txt"))
        gt = {i: np.eye(4) for i in range(n)}
        for i in range(n):
            gt[i][0, 3] = i * step
        ate, rpe, rot, dist = ev.eval_poses(gt, pred, alignment="6dof")
        assert dist == pytest.approx((n - 1) * step) and rpe < 0.05
    finally:
        os.chdir(cwd)
