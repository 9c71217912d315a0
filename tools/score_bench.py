"""Times the PnP-RANSAC scoring kernel alone (pruned and unpruned) on synthetic correspondences.

  python tools/score_bench.py [--pairs 32 --hyp 16384 --points 13700]

Unpruned (`want_counts=True`: every hypothesis is scored against every point) gives the raw test rate and the FP32-lane
utilisation (15 fp32 lane-operations per test); pruned is what the pipeline runs.
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vo_b200  # noqa: E402
from vo_b200 import ops, synthetic  # noqa: E402


def make(B, n, rng, outliers=0.3):
    K = synthetic.KITTI_K
    W, Hh = synthetic.KITTI_WH
    xyz = np.empty((B, n, 3), np.float32)
    uv = np.empty((B, n, 2), np.float32)
    for b in range(B):
        u = rng.uniform(0, W, n)
        v = rng.uniform(0, Hh, n)
        z = rng.uniform(4, 45, n)
        P = np.stack([(u - K[0, 2]) / K[0, 0] * z, (v - K[1, 2]) / K[1, 1] * z, z], 1)
        w = rng.normal(0, 0.008, 3)
        th = np.linalg.norm(w)
        k = w / th
        Kx = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
        R = np.eye(3) + np.sin(th) * Kx + (1 - np.cos(th)) * Kx @ Kx
        t = np.array([0.02, -0.01, -0.67])
        Q = P @ R.T + t
        q = np.stack([K[0, 0] * Q[:, 0] / Q[:, 2] + K[0, 2], K[1, 1] * Q[:, 1] / Q[:, 2] + K[1, 2]], 1)
        q += rng.normal(0, 0.3, q.shape)
        bad = rng.random(n) < outliers
        q[bad] = np.stack([rng.uniform(0, W, bad.sum()), rng.uniform(0, Hh, bad.sum())], 1)
        xyz[b], uv[b] = P, q
    return xyz, uv


def measure(pairs, hyp_n, points, reps=5, dev=None):
    """{"pruned": {...}, "unpruned": {...}}: best-of-`reps` event time of the score stage (vo_profile stage timer)."""
    dev = dev or torch.device("cuda:0")
    rng = np.random.default_rng(8214)
    xyz, uv = make(pairs, points, rng)
    xyz, uv = torch.from_numpy(xyz).to(dev), torch.from_numpy(uv).to(dev)
    n_pts = torch.full((pairs,), points, dtype=torch.int32, device=dev)
    hyp = ops.hypotheses(n_pts, hyp_n)
    out = {"pairs": pairs, "hyp": hyp_n, "points": points}
    sms = torch.cuda.get_device_properties(dev).multi_processor_count
    for name, counts in (("pruned", False), ("unpruned", True)):
        best = None
        for rep in range(reps + 2):
            ops.profile_enable(True, dev)
            ops.profile_collect(dev)
            r = ops.pnp_ransac(xyz, uv, n_pts, synthetic.KITTI_K, hyp, want_counts=counts)
            ms = ops.profile_collect(dev)["score"][0]
            if rep >= 2:
                best = ms if best is None else min(best, ms)
        tests = pairs * hyp_n * points
        out[name] = {"score_ms": best, "tests_per_s": tests / best * 1e3,
                     "fp32_lane_util_at_1965MHz": tests * 15 / (best * 1e-3) / (sms * 128 * 1.965e9),
                     "median_inliers": float(r.n_inl.float().median())}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=32)
    ap.add_argument("--hyp", type=int, default=16384)
    ap.add_argument("--points", type=int, default=13700)
    ap.add_argument("--reps", type=int, default=5)
    a = ap.parse_args()
    print(json.dumps(measure(a.pairs, a.hyp, a.points, a.reps)))


if __name__ == "__main__":
    main()
