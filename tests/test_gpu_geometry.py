"""CUDA back-projection / gather vs the oracle: bit-exact fp32, incl. NaN / 0 / far depth and truncation."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _gpu(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.mark.parametrize("wh,K", [((1241, 376), "kitti"), ((2208, 1242), "zed"), ((64, 48), "kitti"), ((7, 3), "kitti")])
def test_dense_bit_exact(orc, wh, K):
    from vo_b200 import ops, synthetic
    Kmat = synthetic.KITTI_K if K == "kitti" else synthetic.ZED_K
    W, H = wh
    rng = np.random.default_rng(W)
    depth = rng.uniform(0.0, 80.0, (H, W)).astype(np.float32)
    depth[rng.random((H, W)) < 0.01] = np.nan
    depth[rng.random((H, W)) < 0.01] = 0.0
    got = ops.backproject_dense(_gpu(depth), Kmat).cpu().numpy()
    want = orc.backproject_dense(depth, Kmat)
    nan = np.isnan(want)
    assert np.array_equal(np.isnan(got), nan)                            # NaN depth -> NaN point, same places
    assert np.array_equal(got.view(np.uint32)[~nan], want.view(np.uint32)[~nan])   # everything else bitwise


def test_dense_batch_and_golden_f64(golden):
    from vo_b200 import ops
    g = golden("backproject.npz")
    depth = np.stack([g["depth"], g["depth"] * 2])
    out = ops.backproject_dense(_gpu(depth), g["K"]).cpu().numpy()
    kp = g["kp"]
    ui, vi = kp[:, 0].astype(np.int32), kp[:, 1].astype(np.int32)
    want = g["xyz_f64"]
    ok = np.isfinite(want).all(1)
    assert np.allclose(out[0][vi, ui][ok], want[ok], rtol=3e-6, atol=1e-6)
    assert np.allclose(out[1][vi, ui][ok], 2 * want[ok], rtol=3e-6, atol=1e-6)      # linearity in z


def test_gather_vs_oracle_batch(orc):
    import torch
    from vo_b200 import ops, synthetic
    B, N = 4, 1500
    ps = [synthetic.make_pair(200 + b, n_kp=N, kind="orb") for b in range(B)]
    cap = N
    pairs = np.zeros((B, cap, 2), np.int32)
    n_pairs = np.zeros(B, np.int32)
    for b, p in enumerate(ps):
        m, _ = orc.match_u8(p["ref_desc"], p["cur_desc"], orc.NORM_HAMMING, orc.MODE_MUTUAL)
        if b == 3:
            m = m[:0]                                                     # a pair with no matches at all
        pairs[b, :len(m)] = m
        n_pairs[b] = len(m)
    ref_kp = np.stack([p["ref_kp"] for p in ps])
    cur_kp = np.stack([p["cur_kp"] for p in ps])
    depth = np.stack([p["depth"] for p in ps])
    c = ops.gather_backproject(_gpu(pairs), _gpu(n_pairs), _gpu(ref_kp), _gpu(cur_kp), _gpu(depth), ps[0]["K"])
    for b, p in enumerate(ps):
        xyz, ruv, cuv, src, oob = orc.gather_backproject(pairs[b, :n_pairs[b]], p["ref_kp"], p["cur_kp"], p["depth"], p["K"])
        k = int(c.count[b].item())
        assert k == len(xyz) and not oob and int(c.status[b].item()) == 0
        assert np.array_equal(c.xyz[b, :k].cpu().numpy().view(np.uint32), xyz.view(np.uint32))
        assert np.array_equal(c.ref_uv[b, :k].cpu().numpy(), ruv) and np.array_equal(c.cur_uv[b, :k].cpu().numpy(), cuv)
        assert np.array_equal(c.src[b, :k].cpu().numpy(), src)
    assert int(c.count[3].item()) == 0


def test_gather_flags_out_of_image_keypoints():
    import torch
    from vo_b200 import ops
    depth = torch.ones((1, 8, 8), device="cuda")
    K = np.array([[10.0, 0, 4], [0, 10.0, 4], [0, 0, 1]])
    kp = torch.tensor([[[2.5, 3.5], [8.2, 1.0]]], device="cuda")
    pairs = torch.tensor([[[0, 0], [1, 1]]], dtype=torch.int32, device="cuda")
    c = ops.gather_backproject(pairs, torch.tensor([2], dtype=torch.int32, device="cuda"), kp, kp + 5, depth, K)
    assert int(c.count.item()) == 1 and int(c.status.item()) == ops._lib.VO_ST_KP_OUT_OF_IMAGE


def test_sampled_depth_equals_dense_lookup_device_and_pinned_host(orc):
    import vo_b200
    """vo_sample_depth + the depth_kp path of vo_pipeline give exactly what the dense-map path gives, whether the map
    lives in HBM or stays in pinned host memory (zero-copy), including NaN / zero depth and the status of a keypoint
    that truncates outside the image."""
    import torch
    from vo_b200 import ops, synthetic
    B, N = 3, 1200
    batch = synthetic.make_batch(700, B, n_kp=N, kind="orb")
    g = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    ref_kp, cur_kp, depth = g(batch["ref_kp"]), g(batch["cur_kp"]), g(batch["depth"])
    z_dev = ops.sample_depth(ref_kp, depth)
    pinned = torch.from_numpy(np.ascontiguousarray(batch["depth"])).pin_memory()
    z_host = ops.sample_depth(ref_kp, pinned)
    torch.cuda.synchronize()
    want = np.stack([batch["depth"][b][batch["ref_kp"][b, :, 1].astype(np.int32), batch["ref_kp"][b, :, 0].astype(np.int32)]
                     for b in range(B)])
    for z in (z_dev, z_host):
        got = z.cpu().numpy()
        assert np.array_equal(np.isnan(got), np.isnan(want))
        assert np.array_equal(got[~np.isnan(want)], want[~np.isnan(want)])
    kw = dict(norm_or_metric=ops.VO_NORM_HAMMING, mode=ops.VO_MODE_MUTUAL, n_hyp=256, pair0=700)
    a = ops.pipeline(g(batch["ref_desc"]), g(batch["cur_desc"]), ref_kp, cur_kp, depth, batch["K"], **kw)
    b = ops.pipeline(g(batch["ref_desc"]), g(batch["cur_desc"]), ref_kp, cur_kp, None, batch["K"], depth_kp=z_host,
                     hw=depth.shape[1:], **kw)
    for f in ("n_matches", "n_corr", "n_inl", "status"):
        assert torch.equal(getattr(a, f), getattr(b, f)), f
    assert torch.equal(a.T_rel, b.T_rel)
    # a matched keypoint outside the image fails the pair on both paths
    bad = batch["ref_kp"].copy()
    bad[1, :, 0] += 5000.0
    zb = ops.sample_depth(g(bad), pinned)
    c = ops.pipeline(g(batch["ref_desc"]), g(batch["cur_desc"]), g(bad), cur_kp, None, batch["K"], depth_kp=zb,
                     hw=depth.shape[1:], **kw)
    d = ops.pipeline(g(batch["ref_desc"]), g(batch["cur_desc"]), g(bad), cur_kp, depth, batch["K"], **kw)
    assert torch.equal(c.status, d.status) and int(c.status[1].item()) & vo_b200.VO_ST_KP_OUT_OF_IMAGE
    assert int(c.status[0].item()) == 0


def test_host_runner_sampled_equals_dense():
    import torch
    from vo_b200 import ops, sequence, synthetic
    B = 12
    batch = synthetic.make_batch(40, B, n_kp=800, kind="orb")
    cfg = sequence.PipelineConfig(ops.VO_NORM_HAMMING, ops.VO_MODE_MUTUAL, 0.0, 0, n_hyp=256)
    outs = []
    for mode in ("dense", "sampled", "hybrid", "matched"):   # matched: two lanes (stream + vo_ctx), maps read in place
        r = sequence.HostPairRunner(batch, cfg, chunk=5, device="cuda", depth_mode=mode)
        for _ in range(2):                                    # second pass: stage buffers and lanes are reused
            T, st, inl = r.run(pair0=40)
            torch.cuda.synchronize()
        outs.append((T.clone(), st.clone(), inl.clone(), r.count_matched_bytes()))
    for o in outs[1:]:
        assert torch.equal(outs[0][0], o[0]) and torch.equal(outs[0][1], o[1]) and torch.equal(outs[0][2], o[2])
    assert outs[1][3] < outs[0][3] / 3 and outs[1][3] < outs[2][3] < outs[0][3]
    assert outs[3][3] < outs[1][3]                            # only the matched keypoints' pixels cross the bus
