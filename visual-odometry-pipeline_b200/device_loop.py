"""Device-resident keyframe loop (vo_seq_*): the control flow of VisualOdometry.process_frame
(VisualOdometry_Stereo.py:232-297) kept on the GPU.

`DeviceLoop.push(kp, desc, depth, frame_no)` only enqueues work (copies into the loop's frame slot, vo_pipeline
against the keyframe slot, a one-thread policy kernel for the gates / bad-PnP counter / pose chaining / keyframe
rule, a conditional slot promotion).  `poses()` synchronises once and returns what the reference keeps in
`global_poses` (:293).  Features come from the caller (the front-ends stay on the host).
"""
import ctypes

import numpy as np
import torch

from . import _lib, ops
from ._lib import SeqConfig, check


class DeviceLoop:
    def __init__(self, K, wh, n_cap, *, kind="orb", kp_stride=2, norm_or_metric=None, mode=None, match_param=0.85,
                 precision=None, allow_inexact=False, n_hyp=512, seed=8214, thr_px=1.5, min_inliers=20, refine_iters=10,
                 min_flow_px=3.0, z_range=(0.0, 50.0), max_step_m=1.5, kf_min_common=200, kf_min_inliers=100,
                 kf_max_dist=1.5, bad_pnp_limit=3, max_frames=4096, device=None):
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.ctx = ops.context(self.device)
        f32 = kind != "orb"
        if norm_or_metric is None:
            norm_or_metric = {"orb": ops.VO_NORM_L2_U8, "sift": ops.VO_METRIC_L2}.get(kind, ops.VO_METRIC_COSINE)
        if mode is None:
            mode = ops.VO_MODE_RATIO_MUTUAL if kind == "r2d2" else ops.VO_MODE_RATIO
        if precision is None:
            # one 11-bit pass is exact only on integer-valued descriptors (OpenCV SIFT); unit-norm R2D2 descriptors need
            # the split passes to reproduce the reference's fp32 `d1 @ d2.t()` decisions (R2D2.py:56-65)
            precision = ops.VO_PREC_F16X1 if kind == "sift" else ops.VO_PREC_TF32X3
        if (f32 and norm_or_metric == ops.VO_METRIC_COSINE and precision in (ops.VO_PREC_TF32X1, ops.VO_PREC_F16X1)
                and not allow_inexact):
            raise ValueError("DeviceLoop: a single 11-bit pass (TF32X1 / F16X1) on cosine similarities of real-valued "
                             "descriptors changes ratio / mutual decisions (~1e-3 similarity error); use TF32X3 / F16X3 "
                             "or pass allow_inexact=True")
        c = SeqConfig()
        c.desc_is_f32, c.n_cap, c.kp_stride, c.H, c.W = int(f32), int(n_cap), int(kp_stride), int(wh[1]), int(wh[0])
        c.K = (ctypes.c_double * 9)(*np.asarray(K, np.float64).reshape(9))
        c.norm_or_metric, c.mode, c.precision, c.match_param = int(norm_or_metric), int(mode), int(precision), float(match_param)
        c.min_flow_px, c.z_min, c.z_max = float(min_flow_px), float(z_range[0]), float(z_range[1])
        c.n_hyp, c.seed, c.thr_px = int(n_hyp), int(seed), float(thr_px)
        c.min_inliers, c.refine_iters = int(min_inliers), int(refine_iters)
        c.max_step_m, c.kf_min_common, c.kf_min_inliers = float(max_step_m), int(kf_min_common), int(kf_min_inliers)
        c.kf_max_dist, c.bad_pnp_limit, c.max_frames = float(kf_max_dist), int(bad_pnp_limit), int(max_frames)
        self.cfg = c
        self.desc_dtype = np.float32 if f32 else np.uint8
        self.desc_cols = 128 if f32 else 32
        h = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            check(self.ctx.lib.vo_seq_create(self.ctx.handle, ctypes.byref(c), ctypes.byref(h)), "vo_seq_create")
        self.handle = h
        self._keep = []          # host staging arrays must outlive the asynchronous copies

    @staticmethod
    def _ptr(a):
        if isinstance(a, torch.Tensor):
            return ctypes.c_void_p(a.data_ptr())
        return a.ctypes.data_as(ctypes.c_void_p)

    def push(self, kp, desc, depth, frame_no):
        """kp (n, >=2) array-like (x, y first), desc (n, 32) uint8 or (n, 128) float32, depth (H, W) float32.
        numpy arrays (copied from host memory) or torch tensors (device or pinned) are accepted."""
        c = self.cfg
        if isinstance(kp, torch.Tensor):
            kp_a = kp[:, :c.kp_stride].to(torch.float32).contiguous()
            n = int(kp_a.shape[0])
        else:
            kp_a = np.ascontiguousarray(np.asarray(kp)[:, :c.kp_stride], dtype=np.float32)
            n = int(kp_a.shape[0])
        if isinstance(desc, torch.Tensor):
            desc_a = desc.contiguous()
        else:
            desc_a = np.ascontiguousarray(desc, dtype=self.desc_dtype)
        if tuple(desc_a.shape) != (n, self.desc_cols):
            raise ValueError(f"DeviceLoop.push: descriptors {tuple(desc_a.shape)} do not match {n} keypoints x {self.desc_cols}")
        if isinstance(depth, torch.Tensor):
            depth_a = depth.to(torch.float32).contiguous()
        else:
            depth_a = np.ascontiguousarray(depth, dtype=np.float32)
        if tuple(depth_a.shape) != (c.H, c.W):
            raise ValueError(f"DeviceLoop.push: depth {tuple(depth_a.shape)} != ({c.H}, {c.W})")
        self._keep.append((kp_a, desc_a, depth_a))
        with torch.cuda.device(self.device):
            stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
            check(self.ctx.lib.vo_seq_push(self.handle, self._ptr(desc_a), self._ptr(kp_a), n, self._ptr(depth_a),
                                           int(frame_no), stream), "vo_seq_push")

    def push_image(self, image, depth, frame_no, nfeatures=500, frontend=None):
        """Extract on the device and push — the image is the only per-frame upload besides the depth map; one host read of
        the keypoint count per frame.  frontend: "orb" (orb_frontend.OrbExtractor, cv2.ORB_create(nfeatures) semantics; the
        default of a byte-descriptor loop) or "sift" (sift_frontend.SiftExtractor; the default of a float-descriptor loop).
        Both extractors are parity-tested on a B200 (tests/test_gpu_orb_frontend.py, tests/test_gpu_sift_frontend.py)."""
        frontend = frontend or ("orb" if self.desc_cols == 32 else "sift")
        if (frontend == "orb") != (self.desc_cols == 32) or frontend not in ("orb", "sift"):
            raise ValueError(f"DeviceLoop.push_image: front-end {frontend!r} does not produce this loop's descriptors")
        if getattr(self, "_extractor", None) is None:
            if frontend == "orb":
                from .orb_frontend import OrbExtractor
                self._extractor = OrbExtractor(self.cfg.H, self.cfg.W, nfeatures=nfeatures, device=self.device)
            else:
                from .sift_frontend import SiftExtractor
                self._extractor = SiftExtractor(self.cfg.H, self.cfg.W, max_keypoints=self.cfg.n_cap, device=self.device)
        kp, desc, _ = self._extractor.extract(image)
        return self.push(kp, desc, depth, frame_no)

    def __len__(self):
        return int(self.ctx.lib.vo_seq_frames(self.handle))

    def poses(self, first=0, count=None):
        """(poses (count,4,4) f64, info (count,6) int32: status, n_matches, n_corr, n_inl, keyframe id, promoted).
        Synchronises the stream."""
        n = len(self)
        count = n - first if count is None else count
        poses = np.empty((count, 4, 4), np.float64)
        info = np.empty((count, 6), np.int32)
        with torch.cuda.device(self.device):
            stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
            check(self.ctx.lib.vo_seq_read(self.handle, int(first), int(count), self._ptr(poses), self._ptr(info), stream),
                  "vo_seq_read")
        self._keep.clear()
        return poses, info

    def close(self):
        if getattr(self, "handle", None):
            self.ctx.lib.vo_seq_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
