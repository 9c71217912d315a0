"""vo_pipeline (the whole hot path in one C-ABI call) vs the oracle pipeline and ground truth."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _gpu(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.mark.parametrize("kind,norm_or_metric,mode,param,prec", [
    ("orb", 0, 1, 0.0, 0),          # c2: Hamming mutual
    ("orb", 1, 0, 0.85, 0),         # reference ORB semantics: L2 bytes + ratio
    ("sift", 0, 0, 0.85, 2),        # c1: SIFT L2 + ratio, SIMT
    ("sift", 0, 0, 0.85, 1),        # c1 on the tensor-core path
    ("sift", 0, 0, 0.85, 3),        # c1 on the fp16 tensor-core pass (what bench c1 / c4 run)
])
def test_pipeline_vs_oracle(orc, kind, norm_or_metric, mode, param, prec):
    from vo_b200 import ops, synthetic
    B, N, H = 4, 2000, 512
    batch = synthetic.make_batch(500, B, n_kp=N, kind=kind)
    res = ops.pipeline(_gpu(batch["ref_desc"]), _gpu(batch["cur_desc"]), _gpu(batch["ref_kp"]), _gpu(batch["cur_kp"]),
                       _gpu(batch["depth"]), batch["K"], norm_or_metric=norm_or_metric, mode=mode, match_param=param,
                       precision=prec, n_hyp=H, seed=8214, pair0=500)
    T = res.T_rel.cpu().numpy()
    for b, p in enumerate(batch["pairs"]):
        o = orc.pair_pipeline(p["ref_desc"], p["cur_desc"], p["ref_kp"], p["cur_kp"], p["depth"], p["K"],
                              norm_or_metric=norm_or_metric, mode=mode, match_param=param, n_hyp=H, seed=8214, pair=500 + b)
        assert int(res.n_matches[b].item()) == o["n_matches"]
        assert int(res.n_corr[b].item()) == o["n_corr"]
        assert int(res.n_inl[b].item()) == o["n_inl"]
        assert int(res.status[b].item()) == 0 and o["ok"]
        ang, dt = synthetic.pose_errors(T[b], o["T_rel"])
        assert ang < 1e-9 and dt < 1e-9
        ang, dt = synthetic.pose_errors(T[b], p["T_rel"])
        assert ang < 2e-3 and dt < 2e-2


@pytest.mark.parametrize("kind", ["sift", "orb_tc", "orb_l2"])
def test_schedules_of_the_same_work_give_the_same_poses(kind):
    """The same consecutive pairs through (a) the persistent and the plain form of the fp16 / e4m3 matcher passes
    (VO_TC_PERSIST), (b) one or two stream lanes of sequence.run_resident, (c) chunks of different sizes, (d) the frames-
    prepared-once path and the two-sided one (VO_NO_CHAIN_PREP): the schedule never shows in the result."""
    import os
    import torch
    from vo_b200 import ops, sequence, synthetic
    P, N = 12, 1500
    k = "sift" if kind == "sift" else "orb"
    chain = synthetic.make_chain(77, P, n_kp=N, kind=k)
    seq = sequence.FrameSequence.from_numpy(chain, "cuda")
    mc = {"sift": dict(norm_or_metric=ops.VO_METRIC_L2, mode=ops.VO_MODE_RATIO, match_param=0.85, precision=ops.VO_PREC_F16X1),
          "orb_tc": dict(norm_or_metric=ops.VO_NORM_HAMMING_TC, mode=ops.VO_MODE_MUTUAL, match_param=0.0),
          "orb_l2": dict(norm_or_metric=ops.VO_NORM_L2_U8, mode=ops.VO_MODE_RATIO, match_param=0.85)}[kind]
    cfg = sequence.PipelineConfig(n_hyp=256, **mc)

    def run(chunk, lanes=1, **env):
        old = {k_: os.environ.get(k_) for k_ in env}
        os.environ.update(env)
        try:
            out = sequence.run_resident(seq, cfg, pair0=5, chunk=chunk, lanes=lanes)
            torch.cuda.synchronize()
            return out.T_rel.cpu().numpy(), out.status.cpu().numpy(), out.n_inl.cpu().numpy()
        finally:
            for k_, v in old.items():
                if v is None:
                    os.environ.pop(k_, None)
                else:
                    os.environ[k_] = v
    want = run(P)
    assert (want[1] == 0).sum() >= P - 1 and want[2].min() > 20
    for got in (run(P, VO_TC_PERSIST="0"), run(P, VO_TC_PERSIST="1"), run(5, lanes=2), run(3, lanes=3), run(P, VO_NO_CHAIN_PREP="1"), run(1)):
        for a, b in zip(want, got):
            assert np.array_equal(a, b)
