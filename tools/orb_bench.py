"""GPU-box helper: ORB front-end (vo_orb_extract) on a KITTI-shaped frame — parity against the CPU restatement,
frames/s with the image resident in HBM and from pinned host memory, and the same frame through cv2.ORB_create() on the
host cores.  First thing to run on a B200 for csrc/orb.cu:
    python -m pytest tests/test_gpu_orb_frontend.py -q && python tools/orb_bench.py"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import vo_b200  # noqa: F401
from vo_b200 import ops
from vo_b200.orb_frontend import OrbExtractor


def synthetic_bgr(h, w, seed=8214):
    rng = np.random.default_rng(seed)
    coarse = rng.integers(0, 256, (h // 8 + 2, w // 8 + 2, 3)).astype(np.float32)
    ys, xs = np.arange(h) / 8.0, np.arange(w) / 8.0
    y0, x0 = ys.astype(int), xs.astype(int)
    fy, fx = (ys - y0)[:, None, None], (xs - x0)[None, :, None]
    img = (coarse[y0][:, x0] * (1 - fy) * (1 - fx) + coarse[y0 + 1][:, x0] * fy * (1 - fx) +
           coarse[y0][:, x0 + 1] * (1 - fy) * fx + coarse[y0 + 1][:, x0 + 1] * fy * fx)
    return np.clip(img + rng.integers(0, 25, img.shape), 0, 255).astype(np.uint8)


def main(h=376, w=1241, reps=200):
    img = synthetic_bgr(h, w)
    out = {"frame": f"{w}x{h} BGR", "reps": reps}
    try:   # parity (the oracle is test infrastructure: used here as the checker only)
        from oracle import orb_frontend as of
        want = of.detect_and_compute(of.bgr_to_gray(img))
    except Exception as e:   # noqa: BLE001
        want, out["oracle"] = None, f"unavailable: {e}"
    orb = OrbExtractor(h, w)
    dev_img = torch.from_numpy(img).cuda()
    kp, desc, aux = orb.extract(dev_img)
    torch.cuda.synchronize()
    out["keypoints"] = int(kp.shape[0])
    if want is not None:
        scales = of.level_scales()
        lev = aux[:, 0].cpu().numpy().astype(int)
        s = np.array([scales[l] for l in lev], np.float32)
        k = kp.cpu().numpy()
        got = {(int(l), int(x), int(y)): bytes(d) for l, x, y, d in zip(lev, np.rint(k[:, 0] / s), np.rint(k[:, 1] / s), desc.cpu().numpy())}
        ref = {(int(l), int(x), int(y)): bytes(d) for l, x, y, d in zip(want["level"], want["xl"], want["yl"], want["desc"])}
        out["parity"] = {"same_keypoints": set(got) == set(ref), "descriptors_equal": sum(got.get(q) == d for q, d in ref.items()), "of": len(ref)}
    l0 = ops.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for name, src in (("resident", dev_img), ("pinned_host", torch.from_numpy(img).pin_memory())):
        for _ in range(10):
            orb.extract(src)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(reps):
            orb.extract(src)          # includes the D2H read of the keypoint count, as a caller would do
        e1.record()
        torch.cuda.synchronize()
        out[f"{name}_frames_per_s"] = reps / (e0.elapsed_time(e1) * 1e-3)
    out["launches_per_frame"] = (ops.launch_count() - l0) / (2 * (reps + 10))
    try:
        import cv2
        cv2.setNumThreads(os.cpu_count() or 1)
        o = cv2.ORB_create()
        gray = cv2.cvtColor(img, cv2.COLOR_BGR2GRAY)
        for _ in range(3):
            o.detectAndCompute(gray, None)
        t0 = time.perf_counter()
        for _ in range(20):
            o.detectAndCompute(cv2.cvtColor(img, cv2.COLOR_BGR2GRAY), None)
        out["opencv_cpu_frames_per_s"] = 20 / (time.perf_counter() - t0)
        out["opencv_threads"] = cv2.getNumThreads()
    except Exception as e:   # noqa: BLE001
        out["opencv_cpu"] = f"unavailable: {e}"
    print(json.dumps(out))


if __name__ == "__main__":
    main(*[int(a) for a in sys.argv[1:]])
