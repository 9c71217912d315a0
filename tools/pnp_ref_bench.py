"""Latency of the reference-sampler PnP-RANSAC (vo_pnp_ransac_ref, Mode R) against the throughput sampler (vo_pnp_ransac,
512 hypotheses) on one pair's correspondences, and against the reference's CPU call (3 x cv2.solvePnPRansac).  One JSON line."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402
import vo_b200  # noqa: E402,F401
from vo_b200 import ops  # noqa: E402
from test_oracle_pnp_ref import _scene  # noqa: E402


def main():
    import cv2
    out = {}
    for n in (500, 1500, 4000):
        X, uv, K = _scene(3, n, 0.3)
        rng = np.random.RandomState(1)
        boot = np.stack([rng.randint(0, n, n) for _ in range(3)]).astype(np.int32)
        gX, gu, gb = (torch.from_numpy(np.ascontiguousarray(a)).cuda() for a in (X, uv, boot))
        cnt = torch.tensor([n], dtype=torch.int32, device="cuda")
        hyp = ops.hypotheses(cnt, 512)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

        def timeit(fn, reps=50):
            for _ in range(5):
                fn()
            torch.cuda.synchronize()
            e0.record()
            for _ in range(reps):
                fn()
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / reps
        t_ref = timeit(lambda: ops.pnp_ransac_ref(gX, gu, n, K, gb))
        t_thr = timeit(lambda: ops.pnp_ransac(gX[None], gu[None], cnt, K, hyp))
        cv2.setNumThreads(os.cpu_count() or 1)
        t0 = time.perf_counter()
        for _ in range(5):
            for r in range(3):
                cv2.solvePnPRansac(X[boot[r]], uv[boot[r]].reshape(-1, 1, 2), K, None, iterationsCount=100, reprojectionError=1.5)
        t_cv = (time.perf_counter() - t0) / 5 * 1e3
        out[n] = {"mode_r_ms": t_ref, "mode_t_512hyp_ms": t_thr, "cv2_3x_solvePnPRansac_ms": t_cv}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
