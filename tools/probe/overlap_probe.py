"""Does the zero-copy depth sampler run while a matcher launch owns the SMs?  (e2e path: sampling of chunk c+1 is meant
to overlap the matcher of chunk c.)  Times the sampler alone, the matcher alone, and both launched together on two streams."""
import os, sys, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
import vo_b200  # noqa
from vo_b200 import ops

dev = torch.device("cuda")
H, W, F, N = 376, 1241, 125, 5000
host = torch.empty((F, H, W), dtype=torch.float32).pin_memory(); host.uniform_(1, 40)
rng = np.random.default_rng(0)
kp = torch.from_numpy(np.stack([rng.uniform(1, W - 1, (F, N)), rng.uniform(1, H - 1, (F, N))], 2).astype(np.float32)).to(dev)
out = torch.empty((F, N), dtype=torch.float32, device=dev)
desc = torch.randint(0, 256, (F + 1, N, 32), dtype=torch.uint8, device=dev)
sA = torch.cuda.Stream(); sB = torch.cuda.Stream(priority=-1); sC = torch.cuda.Stream()
ev = lambda: torch.cuda.Event(enable_timing=True)

def matcher():
    return ops.match_u8(desc[:F], desc[1:F + 1], ops.VO_NORM_HAMMING, ops.VO_MODE_MUTUAL, 0.0, want_dist=False)

res = {}
for name, sB_ in (("prio", sB), ("noprio", sC)):
    for mode in ("sampler_alone", "matcher_alone", "both"):
        for rep in range(3):
            torch.cuda.synchronize()
            a0, a1, b0, b1, go = ev(), ev(), ev(), ev(), torch.cuda.Event()
            with torch.cuda.stream(sA):
                go.record(sA)
                a0.record(sA)
                if mode != "sampler_alone":
                    matcher()
                a1.record(sA)
            with torch.cuda.stream(sB_):
                sB_.wait_event(go)
                b0.record(sB_)
                if mode != "matcher_alone":
                    ops.sample_depth(kp, host, out=out)
                b1.record(sB_)
            torch.cuda.synchronize()
        res[f"{name}/{mode}"] = {"matcher_ms": a0.elapsed_time(a1), "sampler_ms": b0.elapsed_time(b1)}
print(json.dumps(res, indent=1))
