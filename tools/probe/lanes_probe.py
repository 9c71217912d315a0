"""Resident pipeline rate vs the number of stream lanes of sequence.run_resident (chunks issued round-robin on `lanes` streams,
one vo_ctx each, so that one chunk's small-grid PnP kernels overlap the next chunk's matcher).  One JSON line per workload."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

import vo_b200  # noqa: E402,F401
from vo_b200 import ops, sequence  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    res = {}
    for name in sys.argv[1:] or ["c2", "c3"]:
        wl = bench.WORKLOADS[name]
        P = wl["pairs"]
        host = bench.make_host_chain(wl, P, first_index=0, pinned=True)
        seq = sequence.FrameSequence.from_numpy({**host["_pinned"], "K": host["K"]}, dev)
        cfg = sequence.PipelineConfig(n_hyp=wl["n_hyp"], **bench.matcher_cfg(wl["kind"], ops, wl))
        base = None
        for lanes, chunk in [(1, wl["chunk"]), (2, wl["chunk"]), (2, wl["chunk"] // 2), (3, wl["chunk"] // 2), (4, wl["chunk"] // 4), (1, wl["chunk"] // 2)]:
            out = ops.PipelineBuffers(P, dev)
            for _ in range(2):
                sequence.run_resident(seq, cfg, chunk=chunk, out=out, lanes=lanes)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = max(2, int(600 / P * {"c2": 30, "c3": 8}.get(name, 8)) // 8)
            e0.record()
            for _ in range(reps):
                sequence.run_resident(seq, cfg, chunk=chunk, out=out, lanes=lanes)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / reps
            T = out.T_rel.cpu().numpy()
            if base is None:
                base = T
            res[f"{name}:lanes{lanes}:chunk{chunk}"] = {"pairs_per_s": round(P / ms * 1e3, 1), "ms_per_pass": round(ms, 3),
                                                         "same_poses": bool(np.array_equal(T, base))}
            print(name, lanes, chunk, res[f"{name}:lanes{lanes}:chunk{chunk}"], flush=True)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
