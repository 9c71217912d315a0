"""Device-resident keyframe loop on R2D2-shaped frames (10k keypoints, cosine ratio + mutual, 3xTF32, B = 1): frames/s of repeated
runs and the per-stage GPU time per frame."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import vo_b200  # noqa: E402,F401
from vo_b200 import ops, synthetic, synthetic_sequence  # noqa: E402
from vo_b200.device_loop import DeviceLoop  # noqa: E402

n_kp, n_frames = int(os.environ.get("N_KP", 10000)), 40
frames, gt = synthetic_sequence.make_sequence(n_frames=n_frames, n_kp=n_kp, kind="r2d2", seed=3)
dev = torch.device("cuda", 0)
pinned = [(torch.from_numpy(np.concatenate([f["kp"], np.full((len(f["kp"]), 1), 32.0)], 1).astype(np.float32)).pin_memory(),
           torch.from_numpy(f["desc"]).to(dev), torch.from_numpy(f["depth"]).pin_memory()) for f in frames]


def run_dev():
    loop = DeviceLoop(synthetic.KITTI_K, synthetic.KITTI_WH, n_kp, kind="r2d2", norm_or_metric=ops.VO_METRIC_COSINE, mode=ops.VO_MODE_RATIO_MUTUAL,
                      match_param=0.9, precision=None, n_hyp=1024, kp_stride=3)
    t0 = time.perf_counter()
    for i, (kp, d, z) in enumerate(pinned):
        loop.push(kp, d, z, i)
    t1 = time.perf_counter()
    res = loop.poses()
    t2 = time.perf_counter()
    loop.close()
    return res, t1 - t0, t2 - t1


for it in range(4):
    t0 = time.perf_counter()
    (got, info), t_push, t_sync = run_dev()
    print(f"run {it}: {n_frames / (time.perf_counter() - t0):.1f} fps  (enqueue {t_push * 1e3:.1f} ms, wait {t_sync * 1e3:.1f} ms)", flush=True)
ops.profile_enable(True)
ops.profile_collect()
run_dev()
print(json.dumps({k: round(v[0] / max(v[1], 1), 4) for k, v in ops.profile_collect().items() if v[1]}))
