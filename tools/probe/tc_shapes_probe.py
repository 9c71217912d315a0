"""Matcher-only timings of the tensor-core passes at the BASELINE shapes (A/B of two builds: VO_B200_LIB=...)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import vo_b200  # noqa: E402,F401
from vo_b200 import ops  # noqa: E402


def timed(fn, reps):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


g = torch.Generator(device="cuda").manual_seed(1)
out = {}
only = os.environ.get("VO_PROBE_ONLY")
for name, B, N in (("sift2k", 256, 2000), ("sift5k", 64, 5000), ("sift10k", 24, 10000), ("sift14k", 12, 14000), ("sift20k", 8, 20000)):
    if only and name != only:
        continue
    ref = torch.randint(0, 256, (B, N, 128), device="cuda", generator=g).float()
    cur = torch.randint(0, 256, (B, N, 128), device="cuda", generator=g).float()
    ms = timed(lambda: ops.match_f32(ref, cur, ops.VO_METRIC_L2, ops.VO_MODE_RATIO, 0.85, ops.VO_PREC_F16X1), 10)
    out[name] = (round(ms, 3), round(2.0 * B * N * N * 128 / ms / 1e9, 1))
    del ref, cur
for name, B, N in (("bits5k", 250, 5000), ("bits2k", 512, 2000)):
    if only and name != only:
        continue
    ref = torch.randint(0, 256, (B, N, 32), dtype=torch.uint8, device="cuda", generator=g)
    cur = torch.randint(0, 256, (B, N, 32), dtype=torch.uint8, device="cuda", generator=g)
    ms = timed(lambda: ops.match_u8(ref, cur, ops.VO_NORM_HAMMING_TC, ops.VO_MODE_MUTUAL, 0.0), 10)
    out[name] = (round(ms, 3), round(B * N * N / ms / 1e9, 3))
    ms = timed(lambda: ops.match_u8(ref, cur, ops.VO_NORM_L2_U8, ops.VO_MODE_RATIO, 0.85), 10)
    out[name + "_l2"] = (round(ms, 3), round(B * N * N / ms / 1e9, 3))
print(os.environ.get("VO_B200_LIB", "default"), os.environ.get("VO_TC_PERSIST", ""), out, flush=True)
