#!/usr/bin/env python
"""R2D2 front-end at KITTI size on one GPU: this library (tcgen05 3xTF32 convolutions, fused heads / NMS / descriptor
gather) against the reference's own way of running it — the same network as torch modules on the same GPU (cuDNN,
fp32 with and without TF32), followed by the reference's NMS / gather code path.  Weights: tests/golden/r2d2_net.npz.
Prints one JSON line."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def torch_reference(sd, name, dev):
    """The architecture of nets/patchnet.py rebuilt from plain torch modules (test infrastructure: baseline only)."""
    import torch
    import torch.nn as nn
    from vo_b200 import r2d2_frontend as rf
    ops_, layers = [], rf.layer_table(name, sd)
    for L in layers:
        conv = nn.Conv2d(L["cin"], L["cout"], L["k"], padding=((L["k"] - 1) * L["dil"]) // 2, dilation=L["dil"])
        conv.weight.data = torch.from_numpy(L["w"].transpose(0, 3, 1, 2).copy())
        conv.bias.data = torch.from_numpy(L["bias"])
        ops_.append(conv)
        if L["bn"]:
            bn = nn.BatchNorm2d(L["cout"], affine=False)
            bn.running_mean.data = torch.from_numpy(L["bn_mean"]); bn.running_var.data = torch.from_numpy(L["bn_var"])
            ops_.append(bn)
        if L["relu"]:
            ops_.append(nn.ReLU(inplace=True))
        if L["pool_after"]:
            ops_.append(nn.MaxPool2d(2))
    if rf.ARCH[name]["upsample"] == 2:
        ops_.append(nn.Upsample(scale_factor=2, mode="bilinear", align_corners=False))
    body = nn.Sequential(*ops_).eval().to(dev)
    clf = nn.Conv2d(128, 2, 1); clf.weight.data = torch.from_numpy(sd["clf.weight"]); clf.bias.data = torch.from_numpy(sd["clf.bias"])
    sal = nn.Conv2d(128, 1, 1); sal.weight.data = torch.from_numpy(sd["sal.weight"]); sal.bias.data = torch.from_numpy(sd["sal.bias"])
    clf, sal = clf.to(dev), sal.to(dev)
    F = torch.nn.functional

    def run(x):
        with torch.no_grad():
            f = body(x)
            rel = F.softmax(clf(f ** 2), dim=1)[:, 1:2]
            u = F.softplus(sal(f ** 2)); rep = u / (1 + u)
            desc = F.normalize(f, p=2, dim=1)
            maxima = (rep == F.max_pool2d(rep, 3, 1, 1)) & (rep >= 0.7) & (rel >= 0.7)
            y, xx = maxima.nonzero().t()[2:4]
            sc = rel[0, 0, y, xx] * rep[0, 0, y, xx]
            keep = sc > 0.85
            return torch.stack([xx[keep].float(), y[keep].float()], 1), desc[0, :, y[keep], xx[keep]].t(), rel, rep
    return run


def main():
    import torch
    import vo_b200  # noqa: F401
    from vo_b200 import ops, r2d2_frontend as rf
    g = np.load(os.path.join(ROOT, "tests", "golden", "r2d2_net.npz"))
    name, sd = str(g["net"]).split("(")[0], {k[3:]: g[k] for k in g.files if k.startswith("w__")}
    H, W = 376, 1241
    rng = np.random.default_rng(1)
    img = np.kron(rng.integers(0, 256, (H // 8, W // 8 + 1, 3)), np.ones((8, 8, 1)))[:H, :W].astype(np.uint8)
    dev = torch.device("cuda")
    net = rf.R2D2Net(name, sd, H, W)
    img_dev = torch.from_numpy(img).to(dev)
    for _ in range(3):
        xys, desc, scores = net.extract(img_dev)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = ops.launch_count()
    e0.record()
    R = 20
    for _ in range(R):
        xys, desc, scores = net.extract(img_dev)
    e1.record(); torch.cuda.synchronize()
    ours_ms = e0.elapsed_time(e1) / R
    launches = (ops.launch_count() - l0) / R
    layers = rf.layer_table(name, sd)
    flop, h, w = 0.0, H, W
    for L in layers:
        flop += 2.0 * h * w * L["cout"] * L["cin"] * L["k"] ** 2
        if L["pool_after"]:
            h, w = h // 2, w // 2
    mean = torch.tensor([0.485, 0.456, 0.406], device=dev).view(1, 3, 1, 1); std = torch.tensor([0.229, 0.224, 0.225], device=dev).view(1, 3, 1, 1)
    x = (torch.from_numpy(img).to(dev).permute(2, 0, 1)[None].float() / 255 - mean) / std
    out = {"frame": f"{W}x{H}", "model": name, "ours_ms": ours_ms, "ours_fps": 1e3 / ours_ms, "keypoints": int(len(xys)),
           "gpu_launches_per_frame": launches, "conv_gflop_per_frame": flop / 1e9,
           "ours_conv_tflops_algorithmic": flop / (ours_ms * 1e-3) / 1e12, "ours_tflops_issued_3xtf32": 3 * flop / (ours_ms * 1e-3) / 1e12}
    ref = torch_reference(sd, name, dev)
    for tf32 in (False, True):
        torch.backends.cudnn.allow_tf32 = tf32
        torch.backends.cuda.matmul.allow_tf32 = tf32
        torch.backends.cudnn.benchmark = True
        for _ in range(3):
            kp_r, d_r, rel_r, rep_r = ref(x)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(R):
            kp_r, d_r, rel_r, rep_r = ref(x)
        e1.record(); torch.cuda.synchronize()
        out["torch_cudnn_tf32_ms" if tf32 else "torch_cudnn_fp32_ms"] = e0.elapsed_time(e1) / R
        if not tf32:
            _, _, _, rel_o, rep_o = net.extract(img_dev, want_maps=True)
            out["max_abs_rel_diff_vs_torch_fp32"] = float((rel_o - rel_r[0, 0]).abs().max())
            out["max_abs_rep_diff_vs_torch_fp32"] = float((rep_o - rep_r[0, 0]).abs().max())
            out["torch_keypoints"] = int(len(kp_r))
    out["speedup_vs_torch_fp32"] = out["torch_cudnn_fp32_ms"] / ours_ms
    out["speedup_vs_torch_tf32"] = out["torch_cudnn_tf32_ms"] / ours_ms
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
