"""The CPU oracle's matchers against the golden vectors (cv2 + the reference's own get_matches / torch matchers)."""
import numpy as np


def test_hamming_knn_matches_cv2(golden, orc):
    g = golden("match_u8.npz")
    ridx, rval, cidx = orc.knn_u8(g["ref"], g["cur"], orc.NORM_HAMMING)
    assert np.array_equal(ridx, g["ham_idx"])          # tie-heavy: lowest index rule, both neighbours
    assert np.array_equal(rval, g["ham_dist"])
    assert np.array_equal(cidx, g["col_idx"])          # knnMatch(cur, ref, k=1)


def test_hamming_mutual_matches_cv2_crosscheck(golden, orc):
    g = golden("match_u8.npz")
    pairs, _ = orc.match_u8(g["ref"], g["cur"], orc.NORM_HAMMING, orc.MODE_MUTUAL)
    assert np.array_equal(pairs, g["cc_pairs"])
    pairs2, _ = orc.match_u8(g["ref2"], g["cur2"], orc.NORM_HAMMING, orc.MODE_MUTUAL)
    assert np.array_equal(pairs2, g["cc_pairs2"])


def test_l2_u8_knn_and_reference_orb_get_matches(golden, orc):
    g = golden("match_u8.npz")
    ridx, rval, _ = orc.knn_u8(g["ref"], g["cur"], orc.NORM_L2_U8)
    assert np.array_equal(ridx, g["l2_idx"])
    assert np.array_equal(rval, g["l2_dist"])          # bit-exact fp32 sqrt of the integer sum
    for r, c, want in (("ref", "cur", "ref_orb_pairs"), ("ref2", "cur2", "ref_orb_pairs2")):
        pairs, _ = orc.match_u8(g[r], g[c], orc.NORM_L2_U8, orc.MODE_RATIO, 0.85)
        assert np.array_equal(pairs, g[want].reshape(-1, 2))


def test_sift_knn_and_reference_get_matches(golden, orc):
    g = golden("match_f32_sift.npz")
    ridx, rval, _ = orc.knn_f32(g["ref"], g["cur"], orc.METRIC_L2)
    assert np.array_equal(ridx, g["knn_idx"])
    assert np.array_equal(rval, g["knn_dist"])         # integer-valued descriptors: every sum exact
    pairs, _ = orc.match_f32(g["ref"], g["cur"], orc.METRIC_L2, orc.MODE_RATIO, 0.85)
    assert np.array_equal(pairs, g["ref_sift_pairs"])
    assert len(pairs) > 50


def _same_or_near_tie(orc, ref, cur, got, want, metric):
    """pairs equal, or every differing row is a near-tie (<1e-5 relative) in fp64."""
    if np.array_equal(got, want):
        return True
    gd, wd = {int(a): int(b) for a, b in got}, {int(a): int(b) for a, b in want}
    for r in set(gd) | set(wd):
        if gd.get(r) == wd.get(r):
            continue
        if r in gd and r in wd:
            s = orc.pair_scores_f64(ref, cur, [r, r], [gd[r], wd[r]], metric)
            if abs(s[0] - s[1]) <= 1e-5 * max(abs(s[0]), abs(s[1])):
                continue
        return False
    return True


def test_r2d2_matchers_match_reference_torch(golden, orc):
    g = golden("match_f32_r2d2.npz")
    ref, cur = g["ref"], g["cur"]
    pairs, dist = orc.match_f32(ref, cur, orc.METRIC_COSINE, orc.MODE_RATIO_MUTUAL, 0.90)
    assert _same_or_near_tie(orc, ref, cur, pairs, g["ratio_mutual_pairs"], orc.METRIC_COSINE)
    if np.array_equal(pairs, g["ratio_mutual_pairs"]):
        assert np.allclose(dist, g["ratio_mutual_dist"], rtol=0, atol=2e-3)  # sqrt(2-2s) amplifies 1-ulp sim noise
    mn, _ = orc.match_f32(ref, cur, orc.METRIC_COSINE, orc.MODE_THRESH_MUTUAL, 0.9)
    assert _same_or_near_tie(orc, ref, cur, mn, g["mnn_pairs"], orc.METRIC_COSINE)
    sm, _ = orc.match_f32(ref, cur, orc.METRIC_COSINE, orc.MODE_THRESH, 0.9)
    assert _same_or_near_tie(orc, ref, cur, sm, g["sim_pairs"], orc.METRIC_COSINE)
    mn7, _ = orc.match_f32(ref, cur, orc.METRIC_COSINE, orc.MODE_THRESH_MUTUAL, 0.7)
    assert _same_or_near_tie(orc, ref, cur, mn7, g["mnn_pairs_t07"], orc.METRIC_COSINE)
    sm7, _ = orc.match_f32(ref, cur, orc.METRIC_COSINE, orc.MODE_THRESH, 0.7)
    assert _same_or_near_tie(orc, ref, cur, sm7, g["sim_pairs_t07"], orc.METRIC_COSINE)
    assert len(pairs) > 50 and len(mn7) > 50 and len(mn) >= 1


def test_empty_and_single_column_inputs(orc):
    e8 = np.zeros((0, 32), np.uint8)
    d8 = np.arange(64, dtype=np.uint8).reshape(2, 32)
    for ref, cur in ((e8, d8), (d8, e8)):
        pairs, _ = orc.match_u8(ref, cur, orc.NORM_HAMMING, orc.MODE_MUTUAL)
        assert pairs.shape == (0, 2)
    # one column: no second neighbour -> ratio test keeps nothing, mutual keeps the best row only
    pairs, _ = orc.match_u8(d8, d8[:1], orc.NORM_HAMMING, orc.MODE_RATIO, 0.85)
    assert pairs.shape == (0, 2)
    pairs, _ = orc.match_u8(d8, d8[:1], orc.NORM_HAMMING, orc.MODE_MUTUAL)
    assert pairs.tolist() == [[0, 0]]
