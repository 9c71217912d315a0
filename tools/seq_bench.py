#!/usr/bin/env python
"""Frames/s of the keyframe loop on one GPU: drop-in VisualOdometry.process_frame (host policy, one D2H round trip
per frame) vs the device-resident loop (vo_seq_*, one synchronisation per sequence).  Features are precomputed
(the front-ends are out of scope).  Prints one JSON line per kind."""
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import vo_b200  # noqa: F401
    from vo_b200 import synthetic, synthetic_sequence
    from vo_b200.device_loop import DeviceLoop
    from test_gpu_dropin import _load_dropin
    import pathlib
    n_frames = int(os.environ.get("SEQ_FRAMES", 60))
    for kind, n_kp, norm, mode, prec, extra in (("orb", 5000, 0, 1, 0, "\norb_matcher: hamming_mutual\n"), ("sift", 2000, 0, 0, 3, "")):
        frames, gt = synthetic_sequence.make_sequence(n_frames=n_frames, n_kp=n_kp, kind=kind, seed=3)
        cwd = os.getcwd()
        tmp = pathlib.Path(tempfile.mkdtemp())
        vos = _load_dropin(tmp, kind, extra + "\npnp_mode: throughput\n")      # the sampler the device loop runs (comparable poses)
        feed = {}
        vos.extract_features_and_desc = lambda img: feed["cur"]
        img = np.zeros((synthetic.KITTI_WH[1], synthetic.KITTI_WH[0], 3), np.uint8)

        def run_host():
            vo = vos.VisualOdometry(synthetic.KITTI_K, seq=0)
            out = []
            for i, f in enumerate(frames):
                feed["cur"] = (f["kp"], f["desc"])
                out.append(vo.process_frame(img, f["depth"], (100, 100), i).pose.copy())
            return np.stack(out), vo
        import contextlib, io
        with contextlib.redirect_stdout(io.StringIO()):
            run_host()
            t0 = time.perf_counter(); want, vo = run_host(); t_host = time.perf_counter() - t0
        os.chdir(cwd)

        pinned = [(torch.from_numpy(np.ascontiguousarray(f["kp"], dtype=np.float32)).pin_memory(),
                   torch.from_numpy(f["desc"]).pin_memory(), torch.from_numpy(f["depth"]).pin_memory()) for f in frames]

        enqueue_s = []

        def run_dev():
            loop = DeviceLoop(synthetic.KITTI_K, synthetic.KITTI_WH, n_kp, kind=kind, norm_or_metric=norm, mode=mode,
                              match_param=0.85, precision=prec, n_hyp=vo.n_hyp, seed=vo.seed)
            torch.cuda.synchronize()
            t_e = time.perf_counter()
            for i, (kp, d, z) in enumerate(pinned):
                loop.push(kp, d, z, i)
            enqueue_s.append(time.perf_counter() - t_e)      # host time to enqueue the whole sequence (no synchronisation inside)
            return loop.poses()
        from vo_b200 import ops
        run_dev()
        torch.cuda.synchronize()
        t_devs = []
        for _ in range(5):
            t0 = time.perf_counter(); got, info = run_dev(); t_devs.append(time.perf_counter() - t0)
        t_dev = min(t_devs)
        ops.profile_enable(True); ops.profile_collect()
        t0 = time.perf_counter(); run_dev(); t_prof = time.perf_counter() - t0
        stages = {k: round(v[0] / max(v[1], 1), 4) for k, v in ops.profile_collect().items() if v[1]}
        ops.profile_enable(False)
        print(json.dumps({"kind": kind, "device_loop_stage_ms_per_frame": stages, "profiled_run_s": t_prof}), flush=True)
        print(json.dumps({"kind": kind, "n_kp": n_kp, "frames": n_frames, "host_policy_fps": n_frames / t_host,
                          "device_loop_fps": n_frames / t_dev, "host_enqueue_us_per_frame": 1e6 * min(enqueue_s) / n_frames,
                          "gpu_us_per_frame": 1e6 * t_dev / n_frames, "device_loop_fps_median_of_5": n_frames / float(np.median(t_devs)), "max_abs_pose_diff": float(np.abs(got - want).max()),
                          "keyframes": int(info[:, 5].sum()), "bad_pnp": int((info[1:, 0] != 0).sum()),
                          "max_pos_err_m": float(np.linalg.norm(got[:, :3, 3] - gt[:, :3, 3], axis=1).max())}), flush=True)


if __name__ == "__main__":
    main()
