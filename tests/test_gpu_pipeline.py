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
