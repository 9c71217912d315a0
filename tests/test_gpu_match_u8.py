"""CUDA byte-descriptor matcher (through the C ABI) vs golden vectors and the CPU oracle.  Bit-exact."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _gpu(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _pairs(res, b=0):
    return res.numpy(b)


def test_golden_hamming_knn_and_crosscheck(golden):
    from vo_b200 import ops
    g = golden("match_u8.npz")
    r = ops.match_u8(_gpu(g["ref"]), _gpu(g["cur"]), ops.VO_NORM_HAMMING, ops.VO_MODE_MUTUAL, want_knn=True)
    assert np.array_equal(r.knn_idx[0].cpu().numpy(), g["ham_idx"])
    assert np.array_equal(r.knn_val[0].cpu().numpy(), g["ham_dist"])
    assert np.array_equal(r.col_idx[0].cpu().numpy(), g["col_idx"])
    assert np.array_equal(_pairs(r), g["cc_pairs"])
    r2 = ops.match_u8(_gpu(g["ref2"]), _gpu(g["cur2"]), ops.VO_NORM_HAMMING, ops.VO_MODE_MUTUAL)
    assert np.array_equal(_pairs(r2), g["cc_pairs2"])


def test_golden_reference_orb_get_matches_l2_bytes(golden):
    from vo_b200 import ops
    g = golden("match_u8.npz")
    r = ops.match_u8(_gpu(g["ref"]), _gpu(g["cur"]), ops.VO_NORM_L2_U8, ops.VO_MODE_RATIO, 0.85, want_knn=True)
    assert np.array_equal(r.knn_idx[0].cpu().numpy(), g["l2_idx"])
    assert np.array_equal(r.knn_val[0].cpu().numpy(), g["l2_dist"])
    assert np.array_equal(_pairs(r), g["ref_orb_pairs"].reshape(-1, 2))
    r2 = ops.match_u8(_gpu(g["ref2"]), _gpu(g["cur2"]), ops.VO_NORM_L2_U8, ops.VO_MODE_RATIO, 0.85)
    assert np.array_equal(_pairs(r2), g["ref_orb_pairs2"].reshape(-1, 2))


@pytest.mark.parametrize("n,m", [(5000, 5000), (1237, 3001), (513, 255), (1, 700), (700, 1)])
def test_vs_oracle_sizes(orc, n, m):
    from vo_b200 import ops, synthetic
    p = synthetic.make_pair(n + m, n_kp=max(n, 8), n_cur=max(m, 8), kind="orb")
    ref, cur = p["ref_desc"][:n], p["cur_desc"][:m]
    for norm, mode, param in ((ops.VO_NORM_HAMMING, ops.VO_MODE_MUTUAL, 0.0), (ops.VO_NORM_HAMMING, ops.VO_MODE_RATIO, 0.85),
                              (ops.VO_NORM_L2_U8, ops.VO_MODE_RATIO, 0.85), (ops.VO_NORM_HAMMING, ops.VO_MODE_NN, 0.0)):
        r = ops.match_u8(_gpu(ref), _gpu(cur), norm, mode, param, want_knn=True)
        ridx, rval, cidx = orc.knn_u8(ref, cur, norm)
        assert np.array_equal(r.knn_idx[0].cpu().numpy(), ridx)
        assert np.array_equal(r.knn_val[0].cpu().numpy(), rval)
        assert np.array_equal(r.col_idx[0].cpu().numpy(), cidx)
        want, wd = orc.accept(ridx, rval, cidx, mode, param)
        assert np.array_equal(_pairs(r), want)
        k = len(want)
        assert np.array_equal(r.dist[0, :k].cpu().numpy(), wd)


def test_ragged_batch_with_counts(orc):
    import torch
    from vo_b200 import ops, synthetic
    B, N, M = 5, 640, 600
    ps = [synthetic.make_pair(40 + b, n_kp=N, n_cur=M, kind="orb") for b in range(B)]
    ref = np.stack([p["ref_desc"] for p in ps])
    cur = np.stack([p["cur_desc"] for p in ps])
    n_ref = np.array([640, 1, 333, 0, 512], np.int32)
    n_cur = np.array([600, 600, 17, 50, 0], np.int32)
    r = ops.match_u8(_gpu(ref), _gpu(cur), ops.VO_NORM_HAMMING, ops.VO_MODE_MUTUAL, n_ref=_gpu(n_ref), n_cur=_gpu(n_cur))
    for b in range(B):
        want, _ = orc.match_u8(ref[b, :n_ref[b]], cur[b, :n_cur[b]], orc.NORM_HAMMING, orc.MODE_MUTUAL)
        assert np.array_equal(_pairs(r, b), want), b
    assert r.count.cpu().tolist() == [len(orc.match_u8(ref[b, :n_ref[b]], cur[b, :n_cur[b]], 0, 1)[0]) for b in range(B)]


def test_empty_inputs_and_bad_arguments():
    import torch
    from vo_b200 import ops, _lib
    e = torch.zeros((0, 32), dtype=torch.uint8, device="cuda")
    d = torch.arange(64, dtype=torch.uint8, device="cuda").reshape(2, 32)
    assert ops.match_u8(e, d).count.item() == 0
    assert ops.match_u8(d, e).count.item() == 0
    with pytest.raises(_lib.VoError):
        ops.match_u8(torch.zeros((4, 16), dtype=torch.uint8, device="cuda"), torch.zeros((4, 16), dtype=torch.uint8, device="cuda"))
    with pytest.raises(_lib.VoError):
        ops.match_u8(d, d, mode=ops.VO_MODE_THRESH)


def test_self_match_is_identity_at_full_size():
    """size-independent property at the c2 size: matching a frame against itself gives i -> i with distance 0."""
    import torch
    from vo_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(1)
    d = torch.randint(0, 256, (5000, 32), dtype=torch.uint8, device="cuda", generator=g)
    r = ops.match_u8(d, d, ops.VO_NORM_HAMMING, ops.VO_MODE_MUTUAL)
    pairs = r.numpy()
    assert np.array_equal(pairs[:, 0], np.arange(5000)) and np.array_equal(pairs[:, 1], np.arange(5000))
    assert float(r.dist[0].abs().max()) == 0.0


# ------------------------------------------------------------------------------------------------ byte-L2 on the tensor cores
def test_golden_reference_orb_l2_bytes_tensor_core_pass(golden):
    """The reference's ORB rule (BFMatcher NORM_L2 over byte values + ratio 0.85) without a column arg-min request runs as
    the exact fp16 tensor-core pass: row top-2 (index, distance) equal to cv2's and accepted pairs equal to the reference's
    own get_matches, bit for bit; and equal to the CUDA-core kernel (VO_U8_L2_SIMT=1) on the same input."""
    import os
    from vo_b200 import ops
    g = golden("match_u8.npz")
    r = ops.match_u8(_gpu(g["ref"]), _gpu(g["cur"]), ops.VO_NORM_L2_U8, ops.VO_MODE_RATIO, 0.85, want_knn="rows")
    assert r.col_idx is None
    assert np.array_equal(r.knn_idx[0].cpu().numpy(), g["l2_idx"])
    assert np.array_equal(r.knn_val[0].cpu().numpy(), g["l2_dist"])
    assert np.array_equal(_pairs(r), g["ref_orb_pairs"].reshape(-1, 2))
    os.environ["VO_U8_L2_SIMT"] = "1"
    try:
        s = ops.match_u8(_gpu(g["ref"]), _gpu(g["cur"]), ops.VO_NORM_L2_U8, ops.VO_MODE_RATIO, 0.85, want_knn="rows")
    finally:
        del os.environ["VO_U8_L2_SIMT"]
    assert np.array_equal(r.knn_idx[0].cpu().numpy(), s.knn_idx[0].cpu().numpy())
    assert np.array_equal(r.knn_val[0].cpu().numpy(), s.knn_val[0].cpu().numpy())
    k = int(r.count[0])
    assert k == int(s.count[0]) and np.array_equal(r.dist[0, :k].cpu().numpy(), s.dist[0, :k].cpu().numpy())


@pytest.mark.parametrize("n,m", [(5000, 5000), (1237, 3001), (513, 255), (1, 700), (700, 1), (193, 385)])
def test_l2_bytes_tensor_core_pass_vs_oracle_sizes(orc, n, m):
    from vo_b200 import ops, synthetic
    p = synthetic.make_pair(n + m, n_kp=max(n, 8), n_cur=max(m, 8), kind="orb")
    ref, cur = p["ref_desc"][:n], p["cur_desc"][:m]
    ridx, rval, cidx = orc.knn_u8(ref, cur, ops.VO_NORM_L2_U8)
    for mode, param in ((ops.VO_MODE_RATIO, 0.85), (ops.VO_MODE_NN, 0.0)):
        r = ops.match_u8(_gpu(ref), _gpu(cur), ops.VO_NORM_L2_U8, mode, param, want_knn="rows")
        assert np.array_equal(r.knn_idx[0].cpu().numpy(), ridx)
        assert np.array_equal(r.knn_val[0].cpu().numpy(), rval)
        want, wd = orc.accept(ridx, rval, cidx, mode, param)
        assert np.array_equal(_pairs(r), want)
        assert np.array_equal(r.dist[0, :len(want)].cpu().numpy(), wd)


def test_l2_bytes_tensor_core_pass_ragged_batch(orc):
    from vo_b200 import ops, synthetic
    B, N, M = 4, 640, 600
    ps = [synthetic.make_pair(60 + b, n_kp=N, n_cur=M, kind="orb") for b in range(B)]
    ref = np.stack([p["ref_desc"] for p in ps])
    cur = np.stack([p["cur_desc"] for p in ps])
    n_ref = np.array([640, 1, 333, 512], np.int32)
    n_cur = np.array([600, 600, 17, 2], np.int32)
    r = ops.match_u8(_gpu(ref), _gpu(cur), ops.VO_NORM_L2_U8, ops.VO_MODE_RATIO, 0.85, n_ref=_gpu(n_ref), n_cur=_gpu(n_cur))
    for b in range(B):
        want, _ = orc.match_u8(ref[b, :n_ref[b]], cur[b, :n_cur[b]], orc.NORM_L2_U8, orc.MODE_RATIO, 0.85)
        assert np.array_equal(_pairs(r, b), want), b


def test_l2_bytes_tensor_core_pass_random_shapes(orc):
    """30 random (n, m) shapes around the tile / row-block boundaries, byte-wise L2 + ratio on the tensor-core pass."""
    from vo_b200 import ops, synthetic
    rng = np.random.default_rng(77)
    edges = [1, 2, 47, 48, 49, 127, 128, 129, 191, 192, 193, 255, 256, 257, 383, 384, 385, 576, 577]
    p = synthetic.make_pair(903, n_kp=700, n_cur=700, kind="orb")
    for it in range(30):
        n = int(rng.choice(edges)) if it % 2 == 0 else int(rng.integers(1, 700))
        m = int(rng.choice(edges)) if it % 3 == 0 else int(rng.integers(1, 700))
        ref, cur = p["ref_desc"][:n], p["cur_desc"][:m]
        r = ops.match_u8(_gpu(ref), _gpu(cur), ops.VO_NORM_L2_U8, ops.VO_MODE_RATIO, 0.85, want_knn="rows")
        ridx, rval, cidx = orc.knn_u8(ref, cur, ops.VO_NORM_L2_U8)
        assert np.array_equal(r.knn_idx[0].cpu().numpy(), ridx), (n, m)
        assert np.array_equal(r.knn_val[0].cpu().numpy(), rval), (n, m)
        want, _ = orc.accept(ridx, rval, cidx, ops.VO_MODE_RATIO, 0.85)
        assert np.array_equal(_pairs(r), want), (n, m)


@pytest.mark.parametrize("n,m", [(5000, 5000), (1237, 3001), (513, 255), (129, 700), (700, 1)])
def test_tensor_core_hamming_is_bit_identical_to_xor_popc(orc, golden, n, m):
    """VO_NORM_HAMMING_TC (bits as e4m3 -1 / +1, K = 256 tcgen05 GEMM, row top-2 — or row arg-max only when nothing reads the second best; roles swapped for the column arg-min) against the XOR + POPC
    kernel and the CPU oracle: same neighbours, same distances, same column arg-min, same accepted pairs, every rule."""
    import torch
    from vo_b200 import ops, synthetic
    p = synthetic.make_pair(3 * n + m, n_kp=max(n, 8), n_cur=max(m, 8), kind="orb")
    ref, cur = p["ref_desc"][:n].copy(), p["cur_desc"][:m].copy()
    if n > 40 and m > 40:   # tie-heavy block: duplicated descriptors on both sides
        ref[7:27] = ref[3]
        cur[11:31] = cur[5]
    ridx, rval, cidx = orc.knn_u8(ref, cur, ops.VO_NORM_HAMMING)
    for mode, param in ((ops.VO_MODE_MUTUAL, 0.0), (ops.VO_MODE_RATIO, 0.85), (ops.VO_MODE_NN, 0.0)):
        for knn in (True, "rows", False):
            a = ops.match_u8(_gpu(ref), _gpu(cur), ops.VO_NORM_HAMMING, mode, param, want_knn=knn)
            b = ops.match_u8(_gpu(ref), _gpu(cur), ops.VO_NORM_HAMMING_TC, mode, param, want_knn=knn)
            torch.cuda.synchronize()
            assert np.array_equal(_pairs(a), _pairs(b)), (mode, knn)
            k = int(a.count[0])
            assert np.array_equal(a.dist[0, :k].cpu().numpy(), b.dist[0, :k].cpu().numpy())
            if knn:
                assert np.array_equal(b.knn_idx[0].cpu().numpy(), ridx)
                assert np.array_equal(b.knn_val[0].cpu().numpy(), rval)
            if knn is True:
                assert np.array_equal(b.col_idx[0].cpu().numpy(), cidx)
    g = golden("match_u8.npz")   # cv2.BFMatcher(NORM_HAMMING, crossCheck=True) on the tie-heavy golden pair
    r = ops.match_u8(_gpu(g["ref"]), _gpu(g["cur"]), ops.VO_NORM_HAMMING_TC, ops.VO_MODE_MUTUAL, want_knn=True)
    assert np.array_equal(r.knn_idx[0].cpu().numpy(), g["ham_idx"]) and np.array_equal(r.knn_val[0].cpu().numpy(), g["ham_dist"])
    assert np.array_equal(r.col_idx[0].cpu().numpy(), g["col_idx"]) and np.array_equal(_pairs(r), g["cc_pairs"])


def test_tensor_core_hamming_ragged_batch(orc):
    import torch
    from vo_b200 import ops, synthetic
    b = synthetic.make_batch(5, 3, n_kp=900, kind="orb", n_cur=777)
    n_ref = torch.tensor([900, 513, 1], dtype=torch.int32, device="cuda")
    n_cur = torch.tensor([777, 129, 300], dtype=torch.int32, device="cuda")
    a = ops.match_u8(_gpu(b["ref_desc"]), _gpu(b["cur_desc"]), ops.VO_NORM_HAMMING, ops.VO_MODE_MUTUAL, n_ref=n_ref, n_cur=n_cur, want_knn=True)
    t = ops.match_u8(_gpu(b["ref_desc"]), _gpu(b["cur_desc"]), ops.VO_NORM_HAMMING_TC, ops.VO_MODE_MUTUAL, n_ref=n_ref, n_cur=n_cur, want_knn=True)
    for i in range(3):
        assert np.array_equal(a.numpy(i), t.numpy(i))
        nr, nc = int(n_ref[i]), int(n_cur[i])
        assert np.array_equal(a.knn_idx[i, :nr].cpu().numpy(), t.knn_idx[i, :nr].cpu().numpy())
        assert np.array_equal(a.knn_val[i, :nr].cpu().numpy(), t.knn_val[i, :nr].cpu().numpy())
        assert np.array_equal(a.col_idx[i, :nc].cpu().numpy(), t.col_idx[i, :nc].cpu().numpy())


def test_sift_values_as_uint8_equal_the_float32_path(golden):
    """128-byte rows under VO_NORM_L2_U8 (OpenCV SIFT descriptors are integers 0..255: uint8 holds them exactly, a quarter of
    the bytes over the bus): same 2-NN, same distances, same accepted pairs as the float32 matcher and as cv2.knnMatch."""
    import torch
    from vo_b200 import ops, synthetic
    g = golden("match_f32_sift.npz")
    ref, cur = g["ref"], g["cur"]
    assert np.array_equal(ref, np.rint(ref)) and ref.min() >= 0 and ref.max() <= 255
    f = ops.match_f32(_gpu(ref), _gpu(cur), ops.VO_METRIC_L2, ops.VO_MODE_RATIO, 0.85, precision=ops.VO_PREC_F16X1, want_knn="rows")
    u = ops.match_u8(_gpu(ref.astype(np.uint8)), _gpu(cur.astype(np.uint8)), ops.VO_NORM_L2_U8, ops.VO_MODE_RATIO, 0.85, want_knn="rows")
    assert np.array_equal(_pairs(f), _pairs(u)) and len(_pairs(u)) > 10
    assert np.array_equal(f.knn_idx[0].cpu().numpy(), u.knn_idx[0].cpu().numpy())
    assert np.array_equal(f.knn_val[0].cpu().numpy(), u.knn_val[0].cpu().numpy())
    k = int(u.count[0])
    assert np.array_equal(f.dist[0, :k].cpu().numpy(), u.dist[0, :k].cpu().numpy())
    # ragged batch through the whole pipeline: identical poses
    b = synthetic.make_batch(3, 2, n_kp=1300, kind="sift")
    args = [_gpu(b[k]) for k in ("ref_kp", "cur_kp", "depth")]
    pf = ops.pipeline(_gpu(b["ref_desc"]), _gpu(b["cur_desc"]), *args, b["K"], norm_or_metric=ops.VO_METRIC_L2, mode=ops.VO_MODE_RATIO,
                      match_param=0.85, precision=ops.VO_PREC_F16X1, n_hyp=256)
    pu = ops.pipeline(_gpu(b["ref_desc"].astype(np.uint8)), _gpu(b["cur_desc"].astype(np.uint8)), *args, b["K"],
                      norm_or_metric=ops.VO_NORM_L2_U8, mode=ops.VO_MODE_RATIO, match_param=0.85, precision=0, n_hyp=256)
    torch.cuda.synchronize()
    assert np.array_equal(pf.T_rel.cpu().numpy(), pu.T_rel.cpu().numpy()) and np.array_equal(pf.n_matches.cpu().numpy(), pu.n_matches.cpu().numpy())
