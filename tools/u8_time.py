"""Tuning helper (GPU box): time the Hamming matcher alone on c2-shaped input (250 pairs of 5k x 5k) with the
library's stage events, and print a checksum of the accepted pairs (equal across builds that are bit-exact)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vo_b200  # noqa: F401
from vo_b200 import ops

def main(B=250, n=5000, reps=5):
    g = torch.Generator(device="cuda").manual_seed(0)
    a = torch.randint(0, 256, (B, n, 32), dtype=torch.uint8, device="cuda", generator=g)
    b = torch.randint(0, 256, (B, n, 32), dtype=torch.uint8, device="cuda", generator=g)
    b[:, : n // 2] = a[:, torch.randperm(n, device="cuda", generator=g)[: n // 2]]      # planted associations
    for _ in range(2):
        r = ops.match_u8(a, b, want_dist=False)
    torch.cuda.synchronize()
    ops.profile_enable(True); ops.profile_collect()
    for _ in range(reps):
        r = ops.match_u8(a, b, want_dist=False)
    st = ops.profile_collect(); ops.profile_enable(False)
    ms = st["match"][0] / st["match"][1]
    cnt = r.count.long()
    mask = torch.arange(r.pairs.shape[1], device="cuda")[None, :] < cnt[:, None]
    chk = int((r.pairs.long() * mask[..., None]).sum().item()) + int(cnt.sum().item()) * 1000003
    print(f"match_u8 {ms:.3f} ms per {B} pairs of {n} x {n} -> {B * n * n / ms / 1e9:.3f} Tdist/s; checksum {chk}", flush=True)

if __name__ == "__main__":
    main(*[int(x) for x in sys.argv[1:]])
