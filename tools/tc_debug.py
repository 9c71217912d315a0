"""Bring-up helper: tcgen05 matcher vs the CUDA-core FP32 kernel on the same device (run on the GPU box)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import vo_b200
from vo_b200 import ops, synthetic

def run(kind, n, m, prec, metric, B=1):
    ps = [synthetic.make_pair(900 + b, n_kp=max(n, 8), n_cur=max(m, 8), kind=kind) for b in range(B)]
    ref = torch.from_numpy(np.stack([p["ref_desc"][:n] for p in ps])).cuda()
    cur = torch.from_numpy(np.stack([p["cur_desc"][:m] for p in ps])).cuda()
    a = ops.match_f32(ref, cur, metric, ops.VO_MODE_NN, 0.0, precision=ops.VO_PREC_FP32_SIMT, want_knn=True)
    torch.cuda.synchronize()
    try:
        t = ops.match_f32(ref, cur, metric, ops.VO_MODE_NN, 0.0, precision=prec, want_knn=True)
        torch.cuda.synchronize()
    except Exception as e:
        print(f"{kind} {n}x{m} prec={prec}: EXCEPTION {e}")
        return False
    ai, ti = a.knn_idx.cpu().numpy(), t.knn_idx.cpu().numpy()
    av, tv = a.knn_val.cpu().numpy(), t.knn_val.cpu().numpy()
    ac, tcx = a.col_idx.cpu().numpy(), t.col_idx.cpu().numpy()
    bad1 = (ai[..., 0] != ti[..., 0]).sum(); bad2 = (ai[..., 1] != ti[..., 1]).sum(); badc = (ac != tcx).sum()
    dv = np.nanmax(np.abs(av - tv)) if av.size else 0
    print(f"{kind} B={B} {n}x{m} prec={prec}: row1 mismatches {bad1}, row2 {bad2}, col {badc}, max|dval| {dv:.3g}")
    if bad1 and bad1 < 1e9:
        idx = np.argwhere(ai[..., 0] != ti[..., 0])[:5]
        for b_, r in idx:
            print("   row", b_, r, "simt", ai[b_, r], av[b_, r], "tc", ti[b_, r], tv[b_, r])
    return bad1 == 0 and badc == 0

if __name__ == "__main__":
    ok = True
    for kind, metric in (("sift", ops.VO_METRIC_L2), ("r2d2", ops.VO_METRIC_COSINE)):
        for prec in (ops.VO_PREC_TF32X1, ops.VO_PREC_TF32X3):
            for n, m in ((128, 128), (128, 256), (256, 128), (200, 333), (2000, 2000)):
                ok &= run(kind, n, m, prec, metric)
    run("sift", 300, 280, ops.VO_PREC_TF32X1, ops.VO_METRIC_L2, B=3)
    print("ALL OK" if ok else "MISMATCHES")
