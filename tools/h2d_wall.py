"""Host->device wall of the box with N ranks pulling at once: what the e2e leg of bench.py can at
best reach.  Per rank: (a) bulk cudaMemcpyAsync of pinned buffers (GB/s), (b) zero-copy 32-byte sector reads through
vo_sample_depth (M reads/s), (c) both at once.  Max-over-ranks times, aggregate rates.
    python tools/h2d_wall.py                                   # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/h2d_wall.py
Prints one JSON line (rank 0)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import vo_b200  # noqa: E402,F401
from vo_b200 import ops, sequence  # noqa: E402


def main():
    rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    sequence.bind_to_gpu_numa(local) if world > 1 else None
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    H, W, F, N = 376, 1241, 512, 5000
    host = torch.empty((F, H, W), dtype=torch.float32).pin_memory()
    host.uniform_(1.0, 40.0)
    devbuf = torch.empty((64, H, W), dtype=torch.float32, device=dev)
    rng = np.random.default_rng(rank)
    kp = torch.from_numpy(np.stack([rng.uniform(1, W - 1, (F, N)), rng.uniform(1, H - 1, (F, N))], 2).astype(np.float32)).to(dev)
    out = torch.empty((F, N), dtype=torch.float32, device=dev)
    s2 = torch.cuda.Stream()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def bulk():
        for lo in range(0, F, 64):
            devbuf.copy_(host[lo:lo + 64], non_blocking=True)

    def sampled():
        ops.sample_depth(kp, host, out=out)

    def both():
        with torch.cuda.stream(s2):
            ops.sample_depth(kp[F // 2:], host[F // 2:], out=out[F // 2:])
        for lo in range(0, F // 2, 64):
            devbuf.copy_(host[lo:lo + 64], non_blocking=True)
        torch.cuda.current_stream().wait_stream(s2)

    res = {}
    for name, fn in (("bulk", bulk), ("sampled", sampled), ("both", both)):
        fn(); sync()
        best = 1e30
        for _ in range(3):
            sync()
            e0.record(); fn(); e1.record()
            torch.cuda.synchronize()
            t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            best = min(best, float(t.item()))
        res[name] = best
    if rank == 0:
        map_bytes = F * H * W * 4
        print(json.dumps({
            "n_gpus": world, "frames_per_rank": F, "map_bytes_per_rank": map_bytes,
            "bulk_gbs_per_gpu": map_bytes / res["bulk"] / 1e6, "bulk_gbs_aggregate": world * map_bytes / res["bulk"] / 1e6,
            "sampled_mreads_per_gpu": F * N / res["sampled"] / 1e3, "sampled_mreads_aggregate": world * F * N / res["sampled"] / 1e3,
            "sampled_frames_per_s_per_gpu": F / res["sampled"] * 1e3, "bulk_frames_per_s_per_gpu": F / res["bulk"] * 1e3,
            "both_frames_per_s_per_gpu": F / res["both"] * 1e3, "ms": res}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
