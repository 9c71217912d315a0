"""GPU-box helper: SIFT front-end (vo_sift_extract) on a KITTI-shaped frame — parity (to the tolerance of
tests/test_oracle_sift.py) against OpenCV on the host, frames/s resident and from pinned host memory, OpenCV's own time.
    python -m pytest tests/test_gpu_sift_frontend.py -q && python tools/sift_bench.py"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import torch
import vo_b200  # noqa: F401
from vo_b200 import ops
from vo_b200.sift_frontend import SiftExtractor
from orb_bench import synthetic_bgr


def main(h=376, w=1241, reps=50):
    img = synthetic_bgr(h, w)
    out = {"frame": f"{w}x{h} BGR", "reps": reps}
    sift = SiftExtractor(h, w)
    dev_img = torch.from_numpy(img).cuda()
    kp, desc, aux = sift.extract(dev_img)
    torch.cuda.synchronize()
    out["keypoints"] = int(kp.shape[0])
    try:
        import cv2
        from test_oracle_sift import _match     # the checker of the test-suite, not part of the product
        cv2.setNumThreads(os.cpu_count() or 1)
        o = cv2.SIFT_create()
        kps, d = o.detectAndCompute(cv2.cvtColor(img, cv2.COLOR_BGR2GRAY), None)
        ref = {"pt": np.array([k.pt for k in kps], np.float32), "size": np.array([k.size for k in kps], np.float32),
               "angle": np.array([k.angle for k in kps], np.float32), "desc": d}
        got = {"pt": kp.cpu().numpy(), "size": aux[:, 0].cpu().numpy(), "angle": aux[:, 1].cpu().numpy(), "desc": desc.cpu().numpy()}
        pairs, ang, derr = _match(ref, got)
        out["parity_vs_opencv"] = {"opencv_keypoints": len(kps), "paired": len(pairs), "max_angle_err_deg": ang,
                                   "desc_within_1": float((derr <= 1).mean()) if len(derr) else None, "desc_max_err": float(derr.max()) if len(derr) else None}
        t0 = time.perf_counter()
        for _ in range(5):
            o.detectAndCompute(cv2.cvtColor(img, cv2.COLOR_BGR2GRAY), None)
        out["opencv_cpu_frames_per_s"] = 5 / (time.perf_counter() - t0)
        out["opencv_threads"] = cv2.getNumThreads()
    except Exception as e:   # noqa: BLE001
        out["opencv"] = f"unavailable: {e}"
    l0 = ops.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for name, src in (("resident", dev_img), ("pinned_host", torch.from_numpy(img).pin_memory())):
        for _ in range(3):
            sift.extract(src)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(reps):
            sift.extract(src)
        e1.record()
        torch.cuda.synchronize()
        out[f"{name}_frames_per_s"] = reps / (e0.elapsed_time(e1) * 1e-3)
    out["launches_per_frame"] = (ops.launch_count() - l0) / (2 * (reps + 3))
    print(json.dumps(out))


if __name__ == "__main__":
    main(*[int(a) for a in sys.argv[1:]])
