"""Stage-by-stage comparison of the GPU ORB front-end (vo_orb_extract, csrc/orb.cu) with the CPU restatement
(oracle/orb_frontend.py), through vo_orb_debug_read.  Prints, per image and pyramid level, the first stage whose output
differs: gray / pyramid level -> FAST score map -> NMS candidate list -> retainBest(2n) -> Harris -> retainBest(n) ->
angle -> Gaussian -> descriptor.  Diagnostic tool, run on the GPU box:  python tools/orb_bisect.py"""
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import vo_b200  # noqa: E402,F401
from vo_b200.orb_frontend import OrbExtractor  # noqa: E402
from vo_b200._lib import check  # noqa: E402
from oracle import orb_frontend as of  # noqa: E402

B = of._BORDER


def read(orb, what, level, dtype, count=None):
    cap = 64 << 20
    buf = np.empty(cap, np.uint8)
    n = ctypes.c_size_t()
    check(orb.ctx.lib.vo_orb_debug_read(orb.handle, what, level, ctypes.c_void_p(buf.ctypes.data), cap, ctypes.byref(n)),
          "vo_orb_debug_read")
    a = buf[:n.value].view(dtype)
    return a.copy() if count is None else a[:count].copy()


def xy_set(xy):
    return set((int(v & 0xffff), int(v >> 16)) for v in xy)


def bisect(name, gray, image=None):
    print(f"== {name}: {gray.shape}")
    H, W = gray.shape
    orb = OrbExtractor(H, W)
    kp, desc, aux = (t.cpu().numpy() for t in orb.extract(gray if image is None else image))
    geo = read(orb, 0, 0, np.int32).reshape(-1, 4)
    counts = read(orb, 10, 0, np.int32)
    cand_n, surv_n, fin_n = counts[0:8], counts[8:16], counts[16:24]
    levels, scales = of.build_pyramid(gray)
    per = of.features_per_level()
    umax = of.umax_table(15)
    pat = of.pattern()
    ok = True
    angles_all = read(orb, 11, 0, np.float32)
    base = 0
    for l in range(8):
        ext = levels[l]
        img = ext[B:-B, B:-B]
        h, w = img.shape
        msg = []
        if (geo[l, 0], geo[l, 1], geo[l, 2]) != (w, h, per[l]):
            msg.append(f"geometry dev {tuple(geo[l])} vs oracle {(w, h, per[l])}")
        pyr = read(orb, 1, l, np.uint8).reshape(geo[l, 1], geo[l, 0])
        if pyr.shape != img.shape or not np.array_equal(pyr, img):
            d = np.argwhere(pyr != img) if pyr.shape == img.shape else []
            msg.append(f"pyramid differs at {len(d)} px, first {d[:3].tolist() if len(d) else '-'}")
        sc_dev = read(orb, 2, l, np.uint8).reshape(h, w)
        sc_ref = of.fast_scores(img, 20)
        if not np.array_equal(sc_dev, sc_ref):
            d = np.argwhere(sc_dev != sc_ref)
            msg.append(f"FAST map differs at {len(d)} px, first {[(int(y), int(x), int(sc_dev[y, x]), int(sc_ref[y, x])) for y, x in d[:4]]}")
        xs, ys, s = of.fast_detect(img, 20)
        m = (xs >= 31) & (xs < w - 31) & (ys >= 31) & (ys < h - 31)
        xs, ys, s = xs[m], ys[m], s[m].astype(np.float32)
        cand = read(orb, 4, l, np.uint32, int(cand_n[l]))
        cand_sc = read(orb, 5, l, np.uint8, int(cand_n[l]))
        ref_c = set(zip(xs.tolist(), ys.tolist()))
        if xy_set(cand) != ref_c:
            a, b = xy_set(cand), ref_c
            msg.append(f"NMS list: dev {len(a)} ref {len(b)}, dev-only {sorted(a - b)[:4]}, ref-only {sorted(b - a)[:4]}, dup {len(cand) - len(a)}")
        else:
            smap = {(int(v & 0xffff), int(v >> 16)): int(c) for v, c in zip(cand, cand_sc)}
            bad = [(x, y) for x, y, q in zip(xs.tolist(), ys.tolist(), s.tolist()) if smap[(x, y)] != int(q)]
            if bad:
                msg.append(f"candidate scores differ at {bad[:4]}")
        keep = of.retain_best(s, 2 * per[l])
        xs2, ys2 = xs[keep], ys[keep]
        surv = read(orb, 6, l, np.uint32, int(surv_n[l]))
        resp_dev = read(orb, 7, l, np.float32, int(surv_n[l]))
        ref_s = set(zip(xs2.tolist(), ys2.tolist()))
        if xy_set(surv) != ref_s:
            a = xy_set(surv)
            msg.append(f"retainBest(2n): dev {len(surv)} (distinct {len(a)}) ref {len(ref_s)}, dev-only {sorted(a - ref_s)[:4]}, ref-only {sorted(ref_s - a)[:4]}")
        resp = np.array([of.harris_response(ext, x + B, y + B) for x, y in zip(xs2, ys2)], np.float32)
        rmap = {(int(v & 0xffff), int(v >> 16)): r for v, r in zip(surv, resp_dev)}
        bad = [(x, y, float(r), float(rmap[(x, y)])) for x, y, r in zip(xs2.tolist(), ys2.tolist(), resp) if (x, y) in rmap and rmap[(x, y)].tobytes() != np.float32(r).tobytes()]
        if bad:
            msg.append(f"Harris differs at {len(bad)} kp, first {bad[:3]}")
        keep = of.retain_best(resp, per[l])
        xs3, ys3, r3 = xs2[keep], ys2[keep], resp[keep]
        fin = read(orb, 8, l, np.uint32, int(fin_n[l]))
        ref_f = set(zip(xs3.tolist(), ys3.tolist()))
        if xy_set(fin) != ref_f:
            a = xy_set(fin)
            msg.append(f"retainBest(n): dev {len(fin)} (distinct {len(a)}) ref {len(ref_f)}, dev-only {sorted(a - ref_f)[:4]}, ref-only {sorted(ref_f - a)[:4]}")
        if len(fin) and not np.all(np.diff(((fin >> 16).astype(np.int64) << 16) | (fin & 0xffff)) > 0):
            msg.append("kept list is not in (y, x) order")
        # angle / blurred / descriptors of the keypoints the device kept
        ang_dev = angles_all[base:base + len(fin)]
        bad = []
        for i, v in enumerate(fin):
            x, y = int(v & 0xffff), int(v >> 16)
            a = of.ic_angle(ext, x + B, y + B, umax, 15)
            if np.float32(a).tobytes() != ang_dev[i].tobytes():
                bad.append((x, y, float(a), float(ang_dev[i])))
        if bad:
            msg.append(f"angle differs at {len(bad)} kp, first {bad[:3]}")
        blur_dev = read(orb, 3, l, np.uint8).reshape(h, w)
        blur_ref = of.blur_7x7(img)
        if not np.array_equal(blur_dev, blur_ref):
            d = np.argwhere(blur_dev != blur_ref)
            msg.append(f"Gaussian differs at {len(d)} px, first {[(int(y), int(x), int(blur_dev[y, x]), int(blur_ref[y, x])) for y, x in d[:4]]}")
        bad = []
        for i, v in enumerate(fin):
            x, y = int(v & 0xffff), int(v >> 16)
            ix, iy = of.rotated_pattern(ang_dev[i], pat)
            vals = blur_dev[y + iy, x + ix].astype(np.int32)
            bits = (vals[0::2] < vals[1::2]).astype(np.uint8)
            want = np.packbits(bits.reshape(32, 8)[:, ::-1], axis=1)[:, 0]
            if not np.array_equal(want, desc[base + i]):
                bad.append((x, y, int(np.unpackbits(want ^ desc[base + i]).sum())))
        if bad:
            msg.append(f"descriptor differs at {len(bad)} kp (given device angle+blur), first {bad[:3]}")
        base += len(fin)
        print(f"  level {l} ({w}x{h}, n={per[l]}): cand {cand_n[l]} surv {surv_n[l]} kept {fin_n[l]}  " + ("OK" if not msg else "MISMATCH"))
        for m_ in msg:
            print("      -", m_)
            ok = False
    print(f"  total {len(kp)} keypoints (sum of kept {int(fin_n.sum())})")
    orb.close()
    return ok


def main():
    rng = np.random.default_rng(8214)
    g = np.load(os.path.join(ROOT, "tests", "golden", "orb_golden.npz"))
    ok = bisect("golden BGR", of.bgr_to_gray(g["image"]), g["image"])
    tex = np.kron(rng.integers(0, 256, (47, 156), dtype=np.uint8), np.ones((8, 8), np.uint8))[:376, :1241]
    tex = (tex.astype(np.int32) + rng.integers(0, 25, tex.shape)).clip(0, 255).astype(np.uint8)
    ok &= bisect("kitti blocky", np.ascontiguousarray(tex))
    ok &= bisect("noise", rng.integers(0, 256, (150, 260), dtype=np.uint8))
    ok &= bisect("tiny", rng.integers(0, 256, (97, 163), dtype=np.uint8))
    torch.cuda.synchronize()
    print("ALL OK" if ok else "MISMATCHES FOUND")


if __name__ == "__main__":
    main()
