#!/usr/bin/env python
"""Regenerates tests/golden/*.npz.  Run in the build container only: it imports the reference from
/root/reference (read-only) and cv2 — the third-party module that holds the reference's arithmetic — and
records their outputs on small seeded inputs.  The fixtures travel with the repo; this script does not need
to (and cannot) run on the GPU box.

    python tests/golden/make_golden.py

Sources of each fixture (file:line under /root/reference):
  match_u8.npz        cv2.BFMatcher(NORM_HAMMING) knnMatch / crossCheck; feature_extractors/ORB.py:23-32 get_matches
  match_f32_sift.npz  cv2.BFMatcher().knnMatch(k=2); feature_extractors/SIFT.py:25-34 get_matches
  match_f32_r2d2.npz  R2D2.py:29-66 mnn_matcher / similarity_matcher / ratio_mutual_nn_matcher (CPU tensors)
  backproject.npz     Utils/geom_utils.py:55-77 unprojection_kp (f64 cross-check of the depthTo3d restatement)
  pnp.npz             cv2.solveP3P, cv2.projectPoints, cv2.solvePnP(ITERATIVE), cv2.solvePnPRansac
                      (VisualOdometry_Stereo.py:129 call signature)
  kitti03_eval.npz    plot_utils/kittievalodom.py:513-570 eval() on plot_utils/data (known-answer tuple)
  r2d2_net.npz        the reference's R2D2 network + NMS on a synthetic image, with the shipped faster2d2 weights
  kitti03_segments.npz  the evaluator's per-segment table and reductions on the same data (:181-233, :247-270, :361-469)
"""
import ast
import os
import sys
import tempfile
import types

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

import cv2  # noqa: E402
import torch  # noqa: E402

import vo_b200  # noqa: E402,F401
from vo_b200 import synthetic  # noqa: E402


def save(name, **arrs):
    path = os.path.join(HERE, name)
    np.savez_compressed(path, **arrs)
    print(f"{name}: {os.path.getsize(path)/1024:.1f} KiB")


def import_reference_extractors():
    """feature_extractors.{ORB,SIFT} from the reference, with the SURVEY 8(c) shim for xfeatures2d."""
    if not hasattr(cv2, "xfeatures2d"):
        cv2.xfeatures2d = types.SimpleNamespace(SIFT_create=cv2.SIFT_create)
    sys.path.insert(0, REF)
    import importlib
    orb = importlib.import_module("feature_extractors.ORB")
    sift = importlib.import_module("feature_extractors.SIFT")
    sys.path.remove(REF)
    return orb, sift


def reference_r2d2_matchers():
    """The three matcher functions of R2D2.py, extracted by name from the reference source at run time
    (the module itself asserts a CUDA device at import, R2D2.py:195)."""
    src = open(os.path.join(REF, "R2D2.py")).read()
    tree = ast.parse(src)
    want = {"mnn_matcher", "similarity_matcher", "ratio_mutual_nn_matcher"}
    ns = {"torch": torch}
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in want:
            exec(compile(ast.Module([node], []), "R2D2.py", "exec"), ns)
    return ns


def gen_match_u8(orb_mod):
    rng = np.random.default_rng(101)
    # tie-heavy: 40 prototypes, few flipped bits -> many equal Hamming distances
    proto = rng.integers(0, 256, (40, 32), dtype=np.uint8)
    def view(n):
        d = proto[rng.integers(0, 40, n)].copy()
        flips = rng.integers(0, 256, (n, 3))
        for i in range(n):
            for f in flips[i][: rng.integers(0, 4)]:
                d[i, f // 8] ^= np.uint8(1 << (f % 8))
        return d
    ref, cur = view(300), view(280)
    knn = cv2.BFMatcher(cv2.NORM_HAMMING).knnMatch(ref, cur, k=2)
    ham_idx = np.array([[m.trainIdx, n.trainIdx] for m, n in knn], np.int32)
    ham_dist = np.array([[m.distance, n.distance] for m, n in knn], np.float32)
    cc = cv2.BFMatcher(cv2.NORM_HAMMING, crossCheck=True).match(ref, cur)
    cc_pairs = np.array(sorted((m.queryIdx, m.trainIdx) for m in cc), np.int64)
    knn1_cols = cv2.BFMatcher(cv2.NORM_HAMMING).knnMatch(cur, ref, k=1)
    col_idx = np.array([m[0].trainIdx for m in knn1_cols], np.int32)
    # what the reference really does for ORB: BFMatcher() = L2 over byte values + ratio 0.85
    l2 = cv2.BFMatcher().knnMatch(ref, cur, k=2)
    l2_idx = np.array([[m.trainIdx, n.trainIdx] for m, n in l2], np.int32)
    l2_dist = np.array([[m.distance, n.distance] for m, n in l2], np.float32)
    ref_pairs = np.asarray(orb_mod.get_matches(None, ref, None, cur, None)).reshape(-1, 2).astype(np.int64)
    # a second, realistic (ORB-like noise) case for the ratio test
    p = synthetic.make_pair(7, n_kp=400, kind="orb")
    ref2, cur2 = p["ref_desc"], p["cur_desc"]
    ref_pairs2 = np.asarray(orb_mod.get_matches(None, ref2, None, cur2, None)).reshape(-1, 2).astype(np.int64)
    cc2 = cv2.BFMatcher(cv2.NORM_HAMMING, crossCheck=True).match(ref2, cur2)
    cc_pairs2 = np.array(sorted((m.queryIdx, m.trainIdx) for m in cc2), np.int64)
    save("match_u8.npz", ref=ref, cur=cur, ham_idx=ham_idx, ham_dist=ham_dist, cc_pairs=cc_pairs, col_idx=col_idx,
         l2_idx=l2_idx, l2_dist=l2_dist, ref_orb_pairs=ref_pairs, ref2=ref2, cur2=cur2, ref_orb_pairs2=ref_pairs2,
         cc_pairs2=cc_pairs2)


def gen_match_sift(sift_mod):
    p = synthetic.make_pair(11, n_kp=320, n_cur=300, kind="sift")
    ref, cur = p["ref_desc"], p["cur_desc"]
    knn = cv2.BFMatcher().knnMatch(ref, cur, k=2)
    idx = np.array([[m.trainIdx, n.trainIdx] for m, n in knn], np.int32)
    dist = np.array([[m.distance, n.distance] for m, n in knn], np.float32)
    pairs = np.asarray(sift_mod.get_matches(None, ref, None, cur, None)).reshape(-1, 2).astype(np.int64)
    save("match_f32_sift.npz", ref=ref, cur=cur, knn_idx=idx, knn_dist=dist, ref_sift_pairs=pairs)


def gen_match_r2d2(fns):
    p = synthetic.make_pair(13, n_kp=360, n_cur=330, kind="r2d2")
    ref, cur = p["ref_desc"], p["cur_desc"]
    # a few exact duplicates exercise the sim>1 -> NaN -> reject quirk
    cur = cur.copy()
    cur[5] = ref[17]
    cur[6] = ref[17]
    a, b = torch.from_numpy(ref), torch.from_numpy(cur)
    rm, rm_d = fns["ratio_mutual_nn_matcher"](a, b)
    mnn = fns["mnn_matcher"](a, b)
    sm, sm_d = fns["similarity_matcher"](a, b)
    mnn_t = fns["mnn_matcher"](a, b, threshold=0.7)         # the default 0.9 keeps only the duplicates here
    sm_t, _ = fns["similarity_matcher"](a, b, threshold=0.7)
    save("match_f32_r2d2.npz", ref=ref, cur=cur, ratio_mutual_pairs=np.asarray(rm, np.int64),
         ratio_mutual_dist=rm_d.numpy(), mnn_pairs=np.asarray(mnn, np.int64), sim_pairs=sm.numpy().astype(np.int64),
         sim_dist=sm_d.numpy(), mnn_pairs_t07=np.asarray(mnn_t, np.int64), sim_pairs_t07=sm_t.numpy().astype(np.int64))


def gen_backproject():
    sys.path.insert(0, REF)
    from Utils.geom_utils import unprojection_kp
    sys.path.remove(REF)
    rng = np.random.default_rng(5)
    H, W = 48, 64
    K = np.array([[721.53, 0, 31.55], [0, 721.53, 22.85], [0, 0, 1.0]])
    depth = rng.uniform(0.5, 70.0, (H, W)).astype(np.float32)
    depth[3, 4] = 0.0
    depth[10, 20] = np.nan
    kp = np.stack([rng.uniform(0, W - 0.01, 200), rng.uniform(0, H - 0.01, 200)], 1).astype(np.float32)
    ui, vi = kp[:, 0].astype(np.int32), kp[:, 1].astype(np.int32)
    z = depth[vi, ui].astype(np.float64)
    # the reference gathers at truncated pixels (VisualOdometry_Stereo.py:97); unprojection_kp on those pixels
    xyz64 = unprojection_kp(np.stack([ui, vi], 1).astype(np.float64), z, K)
    save("backproject.npz", depth=depth, K=K, kp=kp, xyz_f64=xyz64)


def gen_pnp():
    p = synthetic.make_pair(21, n_kp=900, kind="orb")
    from oracle import oracle as orc
    pairs, _ = orc.match_u8(p["ref_desc"], p["cur_desc"], orc.NORM_HAMMING, orc.MODE_MUTUAL)
    xyz, ruv, cuv, src, oob = orc.gather_backproject(pairs, p["ref_kp"], p["cur_kp"], p["depth"], p["K"])
    K = p["K"]
    rng = np.random.default_rng(3)
    # (a) P3P solutions of 40 random triplets
    trip = np.stack([rng.choice(len(xyz), 3, replace=False) for _ in range(40)]).astype(np.int32)
    sols = np.full((40, 4, 12), np.nan)
    nsol = np.zeros(40, np.int32)
    for q, t in enumerate(trip):
        n, rvs, tvs = cv2.solveP3P(xyz[t].astype(np.float64).reshape(3, 1, 3), cuv[t].astype(np.float64).reshape(3, 1, 2),
                                   K, None, flags=cv2.SOLVEPNP_P3P)
        nsol[q] = n
        for k in range(n):
            sols[q, k, :9] = cv2.Rodrigues(rvs[k])[0].ravel()
            sols[q, k, 9:] = tvs[k].ravel()
    # (b) the reference call itself (VisualOdometry_Stereo.py:129)
    ok, rvec, tvec, inl = cv2.solvePnPRansac(objectPoints=xyz, imagePoints=np.ascontiguousarray(cuv).reshape(-1, 1, 2),
                                             cameraMatrix=K, distCoeffs=None, iterationsCount=100, reprojectionError=1.5)
    inl = inl.ravel().astype(np.int32)
    # (c) inlier mask of a fixed pose by OpenCV's rule: projectPoints -> fp32 -> squared error <= 2.25
    R0 = cv2.Rodrigues(rvec)[0]
    pose0 = np.concatenate([R0.ravel(), tvec.ravel()]).astype(np.float32)
    proj = cv2.projectPoints(xyz.astype(np.float64), cv2.Rodrigues(pose0[:9].astype(np.float64).reshape(3, 3))[0],
                             pose0[9:].astype(np.float64), K, None)[0].reshape(-1, 2).astype(np.float32)
    err = ((cuv - proj) ** 2).sum(1).astype(np.float32)
    mask_cv = (err <= np.float32(2.25)).astype(np.uint8)
    # (d) OpenCV's refit on a given inlier set
    sel = inl
    _, rv2, tv2 = cv2.solvePnP(xyz[sel].astype(np.float64), cuv[sel].astype(np.float64), K, None, flags=cv2.SOLVEPNP_ITERATIVE)
    refit_rt = np.concatenate([cv2.Rodrigues(rv2)[0].ravel(), tv2.ravel()])
    save("pnp.npz", xyz=xyz, uv=cuv, K=K, trip=trip, p3p_nsol=nsol, p3p_sols=sols, ransac_ok=np.array(ok),
         ransac_rvec=rvec.ravel(), ransac_tvec=tvec.ravel(), ransac_inliers=inl, pose0=pose0, err_cv=err, mask_cv=mask_cv,
         refit_rt=refit_rt, T_gt=p["T_rel"])


def gen_kitti_eval():
    sys.modules.setdefault("matplotlib", types.ModuleType("matplotlib"))
    mpl = sys.modules["matplotlib"]
    mpl.use = lambda *a, **k: None
    plt = types.ModuleType("matplotlib.pyplot")
    sys.modules["matplotlib.pyplot"] = plt
    mpl.pyplot = plt
    sys.path.insert(0, os.path.join(REF, "plot_utils"))
    import kittievalodom
    sys.path.pop(0)
    tmp = tempfile.mkdtemp()
    os.makedirs(os.path.join(tmp, "config"))
    os.makedirs(os.path.join(tmp, "plot_utils"))
    os.symlink(os.path.join(REF, "plot_utils", "data"), os.path.join(tmp, "plot_utils", "data"))
    with open(os.path.join(REF, "config", "vo_params.yaml")) as f, open(os.path.join(tmp, "config", "vo_params.yaml"), "w") as g:
        g.write(f.read())
    cwd = os.getcwd()
    os.chdir(os.path.join(tmp, "plot_utils"))
    try:
        tup = kittievalodom.KittiEvalOdom().eval("resdir", 1, alignment="6dof")
    finally:
        os.chdir(cwd)
    gt = np.loadtxt(os.path.join(REF, "plot_utils", "data", "03_modified.txt"))[:, :12]
    pred = np.loadtxt(os.path.join(REF, "plot_utils", "data", "global_poses.npy.txt"))[:, :12]
    save("kitti03_eval.npz", gt=gt.astype(np.float64), pred=pred.astype(np.float64), expected=np.array(tup, np.float64))
    print("evaluator tuple:", tup)
    # the per-segment table and its reductions on the same data (calc_sequence_errors :181-233, compute_overall_err
    # :247-270, compute_segment_error :361-390, compute_ATE :392-427, compute_RPE :429-469), called directly
    def to_dict(a):
        P = np.tile(np.eye(4), (len(a), 1, 1))
        P[:, :3, :] = a.reshape(-1, 3, 4)
        return {i: P[i] for i in range(len(a))}
    ev = kittievalodom.KittiEvalOdom()
    pg, pr = to_dict(gt), to_dict(pred)
    g0, p0 = np.linalg.inv(pg[0]), np.linalg.inv(pr[0])
    for k in pr:
        pr[k] = p0 @ pr[k]
        pg[k] = g0 @ pg[k]
    seq_err = ev.calc_sequence_errors(pg, pr)
    overall = ev.compute_overall_err(seq_err)
    seg = ev.compute_segment_error(seq_err)
    seg_arr = np.array([seg[l] if len(seg[l]) else [np.nan, np.nan] for l in ev.lengths], np.float64)
    save("kitti03_segments.npz", seq_err=np.asarray(seq_err, np.float64), overall=np.asarray(overall, np.float64),
         segments=seg_arr, ate=np.float64(ev.compute_ATE(pg, pr)), rpe=np.asarray(ev.compute_RPE(pg, pr), np.float64),
         dist=np.asarray(ev.trajectory_distances(pg), np.float64))


def gen_r2d2_net():
    """The reference's R2D2 network (nets/patchnet.py) with its shipped faster2d2_WASF_N16 weights (the model R2D2.py:191
    selects), run on CPU in fp32 on a small synthetic image; NonMaxSuppression and the scale-1 body of extract_multiscale
    are taken from R2D2.py by name (the module itself asserts a CUDA device at import, :195)."""
    import torch.nn as nn
    import torch.nn.functional as F
    r2 = os.path.join(REF, "feature_extractors", "r2d2")
    sys.path.insert(0, r2)
    import nets.patchnet as patchnet
    sys.path.remove(r2)
    ck = torch.load(os.path.join(r2, "models", "faster2d2_WASF_N16.pt"), map_location="cpu", weights_only=False)
    net = eval(ck["net"], vars(patchnet))
    sd = {k.replace("module.", ""): v for k, v in ck["state_dict"].items()}
    net.load_state_dict(sd)
    net.eval()
    tree = ast.parse(open(os.path.join(REF, "R2D2.py")).read())
    ns = dict(torch=torch, nn=nn, F=F, np=np)
    for node in tree.body:
        if isinstance(node, (ast.ClassDef, ast.FunctionDef)) and node.name in ("NonMaxSuppression", "extract_multiscale"):
            exec(compile(ast.Module([node], []), "R2D2.py", "exec"), ns)
    rng = np.random.default_rng(8214)
    H, W = 97, 163                                   # odd sizes: MaxPool2d(2) floors, the up-sampled maps are 96 x 162
    img = np.full((H, W, 3), 110, np.uint8)
    for _ in range(70):
        x0, y0 = int(rng.integers(0, W)), int(rng.integers(0, H))
        w, h = int(rng.integers(4, 40)), int(rng.integers(4, 30))
        col = tuple(int(c) for c in rng.integers(0, 256, 3))
        if rng.random() < 0.5:
            cv2.rectangle(img, (x0, y0), (x0 + w, y0 + h), col, -1)
        else:
            cv2.circle(img, (x0, y0), w // 2 + 2, col, -1)
    img = cv2.GaussianBlur(img, (3, 3), 0.8)
    img = np.clip(img.astype(np.int32) + rng.integers(-6, 7, img.shape), 0, 255).astype(np.uint8)
    # tools/dataloader.py norm_RGB: ToTensor + Normalize(ImageNet mean / std)
    t = torch.from_numpy(img).permute(2, 0, 1).float().div(255)
    mean = torch.tensor([0.485, 0.456, 0.406]).view(3, 1, 1)
    std = torch.tensor([0.229, 0.224, 0.225]).view(3, 1, 1)
    x = ((t - mean) / std)[None]
    with torch.no_grad():
        res = net(imgs=[x])
        rel, rep, desc = res["reliability"][0], res["repeatability"][0], res["descriptors"][0]
        det = ns["NonMaxSuppression"](rel_thr=0.7, rep_thr=0.7)
        xys, D, scores = ns["extract_multiscale"](net, x, det, min_size=0, max_size=9999, trt=False)
    idxs = np.argwhere(scores.numpy() > 0.85)        # extract_keypoints, R2D2.py:186-188
    arrs = {"w__" + k: v.numpy() for k, v in sd.items() if v.ndim > 0}
    save("r2d2_net.npz", net=np.array(ck["net"]), image=img, rel=rel[0, 0].numpy(), rep=rep[0, 0].numpy(),
         xys_all=xys.numpy(), scores_all=scores.numpy(), desc_all=D.numpy(), keep=idxs.reshape(-1),
         desc_map_sample=desc[0, :, ::7, ::11].numpy(), **arrs)
    print("r2d2:", ck["net"], "maps", tuple(rel.shape), "keypoints", len(scores), "score > 0.85:", len(idxs))


if __name__ == "__main__":
    orb_mod, sift_mod = import_reference_extractors()
    gen_match_u8(orb_mod)
    gen_match_sift(sift_mod)
    gen_match_r2d2(reference_r2d2_matchers())
    gen_backproject()
    gen_pnp()
    gen_kitti_eval()
    gen_r2d2_net()
