"""Host-side throughput layer on CPU: pair sharding, the pose all-gather (gloo, world_size 2) and the prefix-product
pose chaining that mirrors VisualOdometry.process_frame's chaining / failure rule (VisualOdometry_Stereo.py:283,290)."""
import os
import socket

import numpy as np
import pytest
import torch


def _seq():
    import vo_b200  # noqa: F401
    from vo_b200 import sequence
    return sequence


def _random_poses(n, seed=0):
    from vo_b200.synthetic import _rodrigues
    rng = np.random.default_rng(seed)
    T = np.tile(np.eye(4), (n, 1, 1))
    for i in range(n):
        T[i, :3, :3] = _rodrigues(rng.normal(0, 0.01, 3))
        T[i, :3, 3] = rng.normal(0, 0.5, 3)
    return T


def test_shard_range_partitions_the_pair_list():
    seq = _seq()
    for n, w in ((1000, 8), (10, 3), (7, 8), (0, 2), (100000, 8)):
        spans = [seq.shard_range(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        sizes = [hi - lo for lo, hi in spans]
        assert max(sizes) - min(sizes) <= 1


def test_chain_poses_equals_sequential_product_and_failed_pairs_are_identity():
    seq = _seq()
    T = _random_poses(37)
    ok = np.ones(37, bool)
    ok[[3, 4, 20]] = False
    got = seq.chain_poses(T, ok)
    want = [np.eye(4)]
    for i in range(37):
        want.append(want[-1] @ (T[i] if ok[i] else np.eye(4)))   # :283 chain, :290 identity on failure
    assert np.allclose(got, np.stack(want), atol=1e-12)
    assert got.shape == (38, 4, 4)
    assert np.allclose(seq.chain_poses(T[:0]), np.eye(4)[None])


def test_gate_rejects_failed_and_implausible_motion():
    seq = _seq()
    T = _random_poses(5)
    T[2, :3, 3] = [0, 0, 1.6]                    # > 1.5 m per frame (:271)
    status = np.array([0, 1, 0, 0, 2])
    T[0, :3, 3] = [0.1, 0, 0.6]; T[3, :3, 3] = [0, 0, 1.49]
    assert seq.gate_poses(T, status).tolist() == [True, False, False, True, False]


def _worker(rank, world, port, out_dir):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    seq = _seq()
    n = 12
    lo, hi = seq.shard_range(n, rank, world)
    T_all = torch.from_numpy(_random_poses(n, seed=5))
    st_all = torch.tensor([0, 0, 1, 0, 0, 0, 0, 2, 0, 0, 0, 0], dtype=torch.int32)
    T, st = seq.all_gather_poses(T_all[lo:hi].contiguous(), st_all[lo:hi].contiguous(), world)
    assert torch.equal(T, T_all) and torch.equal(st, st_all)          # rank order == pair order
    chain = seq.chain_poses(T.numpy(), seq.gate_poses(T.numpy(), st.numpy(), max_step_m=10.0))
    np.save(os.path.join(out_dir, f"chain{rank}.npy"), chain)
    dist.destroy_process_group()


def test_all_gather_and_chain_world_size_2_gloo(tmp_path):
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    a, b = np.load(tmp_path / "chain0.npy"), np.load(tmp_path / "chain1.npy")
    assert np.array_equal(a, b) and a.shape == (13, 4, 4)
    seq = _seq()
    T = _random_poses(12, seed=5)
    ok = np.ones(12, bool); ok[[2, 7]] = False
    assert np.allclose(a, seq.chain_poses(T, ok), atol=1e-12)


def test_chunk_schedule_covers_all_pairs_and_tapers_the_tail():
    from vo_b200.sequence import chunk_schedule
    for B, chunk in [(1000, 125), (250, 125), (100, 125), (8, 4), (1, 1), (32, 16), (1000, 1000), (7, 3), (5, 2), (40, 5)]:
        sch = chunk_schedule(B, chunk)
        assert sch[0][0] == 0 and sch[-1][1] == B
        assert all(a[1] == b[0] for a, b in zip(sch, sch[1:]))                      # contiguous, in order
        assert all(0 < hi - lo <= chunk for lo, hi in sch)                          # fits the staging buffers
    sizes = [hi - lo for lo, hi in chunk_schedule(1000, 125)]
    assert sizes[-2:] == [62, 31] and max(sizes[:-2]) - min(sizes[:-2]) <= 1       # short drain, equal body
