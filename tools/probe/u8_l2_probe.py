"""Times vo_match_u8 in the reference-faithful ORB mode (byte-wise L2 + ratio) against the Hamming / mutual mode."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import vo_b200
from vo_b200 import ops

B, N = 250, 5000
g = torch.Generator(device="cuda").manual_seed(0)
a = torch.randint(0, 256, (B, N, 32), dtype=torch.uint8, device="cuda", generator=g)
b = torch.randint(0, 256, (B, N, 32), dtype=torch.uint8, device="cuda", generator=g)
for name, norm, mode in (("hamming+mutual", ops.VO_NORM_HAMMING, ops.VO_MODE_MUTUAL), ("l2_u8+ratio", ops.VO_NORM_L2_U8, ops.VO_MODE_RATIO)):
    for _ in range(2):
        ops.match_u8(a, b, norm, mode, 0.85, want_dist=False)
    torch.cuda.synchronize()
    ops.profile_enable(True); ops.profile_collect()
    for _ in range(3):
        ops.match_u8(a, b, norm, mode, 0.85, want_dist=False)
    st = ops.profile_collect(); ops.profile_enable(False)
    ms = st["match"][0] / st["match"][1]
    print(f"{name}: match {ms:.3f} ms per {B} pairs -> {B * N * N / ms / 1e9:.3f} T dist/s; finalize {st['finalize'][0] / st['finalize'][1]:.3f} ms")
