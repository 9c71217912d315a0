"""Stand-in for compute-sanitizer, which is closed on this GPU pool (profiles/sanitizer/README.md): every kernel family at
small, ragged shapes (tools/sanitize_cases.py runs as a test), the tensor-core matchers against the CUDA-core validation
kernel on the device, and run-to-run bit-identity of every configuration — a data race in the mbarrier ring, the named
barriers, the multicast commits or the 64-bit atomics shows up as a result that flips between runs."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))


def _gpu(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.mark.parametrize("group", ["match_f32", "match_u8", "pipeline", "geometry", "seq", "orb", "sift", "conv"])
def test_every_kernel_family_runs_at_small_ragged_shapes(group):
    import sanitize_cases
    sanitize_cases.GROUPS[group]()


def _snapshot(r, b_count):
    out = []
    for b in range(b_count):
        k = int(r.count[b])
        out.append((r.pairs[b, :k].cpu().numpy().copy(), None if r.dist is None else r.dist[b, :k].cpu().numpy().copy()))
    extra = [t.cpu().numpy().copy() for t in (r.knn_idx, r.knn_val, r.col_idx) if t is not None]
    return out, extra


def _same(a, b):
    for (p0, d0), (p1, d1) in zip(a[0], b[0]):
        if not np.array_equal(p0, p1) or (d0 is not None and not np.array_equal(d0, d1, equal_nan=True)):
            return False
    return all(np.array_equal(x, y, equal_nan=True) for x, y in zip(a[1], b[1]))


def test_matchers_are_bit_identical_run_to_run_and_agree_with_the_cuda_core_kernel():
    import torch
    from vo_b200 import ops, synthetic
    n_ref = torch.tensor([1100, 257, 129], dtype=torch.int32, device="cuda")
    n_cur = torch.tensor([1333, 129, 700], dtype=torch.int32, device="cuda")
    for kind, metric in (("sift", ops.VO_METRIC_L2), ("r2d2", ops.VO_METRIC_COSINE)):
        b = synthetic.make_batch(40, 3, n_kp=1100, kind=kind, n_cur=1333)
        ref, cur = _gpu(b["ref_desc"]), _gpu(b["cur_desc"])
        for mode, param in ((ops.VO_MODE_RATIO, 0.85), (ops.VO_MODE_MUTUAL, 0.0), (ops.VO_MODE_RATIO_MUTUAL, 0.9)):
            want = _snapshot(ops.match_f32(ref, cur, metric, mode, param, precision=ops.VO_PREC_FP32_SIMT, n_ref=n_ref, n_cur=n_cur,
                                           want_knn=True), 3)
            for prec in (ops.VO_PREC_TF32X3, ops.VO_PREC_TF32X1, ops.VO_PREC_F16X1, ops.VO_PREC_F16X3):
                if kind == "r2d2" and prec in (ops.VO_PREC_TF32X1, ops.VO_PREC_F16X1):
                    continue                       # one 11-bit pass is exact on integer-valued descriptors only
                runs = [_snapshot(ops.match_f32(ref, cur, metric, mode, param, precision=prec, n_ref=n_ref, n_cur=n_cur, want_knn=True), 3)
                        for _ in range(6)]
                assert all(_same(runs[0], r) for r in runs[1:]), (kind, mode, prec)
                if kind == "sift":                 # exact arithmetic: every precision equals the fp32 CUDA-core kernel bit for bit
                    assert _same(runs[0], want), (kind, mode, prec)
                else:                              # real-valued: same neighbours except fp32-level near ties
                    idx, widx = runs[0][1][0], want[1][0]
                    assert (idx[..., 0] != widx[..., 0]).mean() < 2e-3
    b = synthetic.make_batch(41, 3, n_kp=1100, kind="orb", n_cur=1333)
    ref, cur = _gpu(b["ref_desc"]), _gpu(b["cur_desc"])
    for norm in (ops.VO_NORM_HAMMING, ops.VO_NORM_HAMMING_TC, ops.VO_NORM_L2_U8):
        for mode, param in ((ops.VO_MODE_RATIO, 0.85), (ops.VO_MODE_MUTUAL, 0.0)):
            runs = [_snapshot(ops.match_u8(ref, cur, norm, mode, param, n_ref=n_ref, n_cur=n_cur, want_knn=True), 3) for _ in range(6)]
            assert all(_same(runs[0], r) for r in runs[1:]), (norm, mode)


def test_pipeline_is_bit_identical_run_to_run():
    import torch
    from vo_b200 import ops, synthetic
    for kind, nm, mode, param, prec, n_kp, n_hyp in (("orb", ops.VO_NORM_HAMMING, ops.VO_MODE_MUTUAL, 0.0, 0, 1500, 512),
                                                     ("sift", ops.VO_METRIC_L2, ops.VO_MODE_RATIO, 0.85, ops.VO_PREC_F16X1, 9000, 512),
                                                     ("r2d2", ops.VO_METRIC_COSINE, ops.VO_MODE_RATIO_MUTUAL, 0.9, ops.VO_PREC_TF32X3, 1500, 512)):
        b = synthetic.make_batch(7, 3, n_kp=n_kp, kind=kind)
        args = [_gpu(b[k]) for k in ("ref_desc", "cur_desc", "ref_kp", "cur_kp", "depth")]
        outs = []
        for _ in range(4):
            r = ops.pipeline(*args, b["K"], norm_or_metric=nm, mode=mode, match_param=param, precision=prec, n_hyp=n_hyp)
            torch.cuda.synchronize()
            outs.append((r.T_rel.cpu().numpy().copy(), r.n_inl.cpu().numpy().copy(), r.n_matches.cpu().numpy().copy(), r.status.cpu().numpy().copy()))
        for o in outs[1:]:
            assert all(np.array_equal(x, y) for x, y in zip(outs[0], o)), kind
