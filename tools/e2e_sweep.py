"""Tuning helper (GPU box): the host-buffer (e2e) leg of the c2 workload alone, swept over the share of depth maps that
is sampled zero-copy (hybrid transport) and the chunk size.  Prints pairs/s per setting.
usage: python tools/e2e_sweep.py [pairs] [frac,frac,...] [chunk,chunk,...]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import vo_b200  # noqa: F401
from vo_b200 import ops, sequence, synthetic

def main():
    P = int(sys.argv[1]) if len(sys.argv) > 1 else 500
    fracs = [float(x) for x in (sys.argv[2] if len(sys.argv) > 2 else "0.2,0.35,0.5").split(",")]
    chunks = [int(x) for x in (sys.argv[3] if len(sys.argv) > 3 else "125").split(",")]
    unique = 20
    host = synthetic.make_batch(0, unique, n_kp=5000, kind="orb", K=synthetic.KITTI_K, wh=synthetic.KITTI_WH)
    reps = (P + unique - 1) // unique
    keys = ("ref_desc", "cur_desc", "ref_kp", "cur_kp", "depth")
    pinned = {k: torch.from_numpy(np.concatenate([host[k]] * reps, 0)[:P]).pin_memory() for k in keys}
    rep = {k: pinned[k].numpy() for k in keys}     # numpy views of the pinned buffers: pinned once for every setting
    rep["K"] = host["K"]
    cfg = sequence.PipelineConfig(n_hyp=1024, norm_or_metric=ops.VO_NORM_HAMMING, mode=ops.VO_MODE_MUTUAL, match_param=0.0, precision=0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dev_batch = sequence.PairBatch.from_numpy({**{k: rep[k] for k in keys}, "K": host["K"]}, "cuda")
    out = ops.PipelineBuffers(P, torch.device("cuda"))
    for chunk in chunks:                         # the same pairs resident in HBM, same chunking: the compute floor
        for _ in range(2):
            sequence.run_resident(dev_batch, cfg, chunk=chunk, out=out)
        torch.cuda.synchronize(); e0.record()
        for _ in range(3):
            sequence.run_resident(dev_batch, cfg, chunk=chunk, out=out)
        e1.record(); torch.cuda.synchronize()
        print(f"resident chunk {chunk:4d}: {P / (e0.elapsed_time(e1) / 3) * 1e3:9.0f} pairs/s", flush=True)
    del dev_batch
    ref_T = None
    for mode, frac_list in (("dense", [0.0]), ("hybrid", fracs), ("sampled", [1.0]), ("matched", [1.0])):
        for frac in frac_list:
            for chunk in chunks:
                r = sequence.HostPairRunner(rep, cfg, chunk=chunk, device="cuda", depth_mode=mode, sampled_frac=frac)
                for _ in range(2):
                    r.run(); torch.cuda.synchronize()
                e0.record()
                for _ in range(3):
                    r.run(); torch.cuda.synchronize()
                e1.record(); torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / 3
                ok = float((r.host_status.numpy() == 0).mean())
                T = r.host_T.numpy().copy()
                ref_T = T if ref_T is None else ref_T
                same = bool(np.array_equal(T, ref_T))
                print(f"{mode:8s} frac {frac:.2f} chunk {chunk:4d}: {P / ms * 1e3:9.0f} pairs/s, {r.count_matched_bytes() / ms / 1e6:6.1f} GB/s H2D, ok {ok:.3f}, poses equal to the first setting: {same}", flush=True)
                del r

if __name__ == "__main__":
    main()
