"""ORB front-end on the GPU (vo_orb_create / vo_orb_extract): cv2.ORB_create().detectAndCompute with the reference's
default parameters (feature_extractors/ORB.py:8-21).

Bit-identical to the CPU restatement pinned on OpenCV (identical keypoint set, pt, angle, response, descriptors) on a B200:
tests/test_gpu_orb_frontend.py; tools/orb_bisect.py compares every stage.  Default extractor of feature_extractors/ORB.py."""
import ctypes

import numpy as np
import torch

from . import ops
from ._lib import OrbConfig, VoError, check


class OrbExtractor:
    """orb = OrbExtractor(H, W); kp, desc, aux = orb.extract(image_uint8)   (image [H,W] gray or [H,W,3] BGR).

    kp [n,2] float32 = KeyPoint.pt, desc [n,32] uint8, aux [n,4] float32 = (octave, angle, response, size); rows in
    level-major, then row-major order (OpenCV's own order is unspecified)."""

    def __init__(self, H, W, nfeatures=500, nlevels=8, fast_threshold=20, device=None):
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.ctx = ops.context(self.device)
        self.H, self.W = int(H), int(W)
        cfg = OrbConfig(self.H, self.W, int(nfeatures), int(nlevels), int(fast_threshold))
        h = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            check(self.ctx.lib.vo_orb_create(self.ctx.handle, ctypes.byref(cfg), ctypes.byref(h)), "vo_orb_create")
        self.handle = h
        self.cap = int(self.ctx.lib.vo_orb_capacity(h))
        self.kp = torch.empty((self.cap, 2), dtype=torch.float32, device=self.device)
        self.desc = torch.empty((self.cap, 32), dtype=torch.uint8, device=self.device)
        self.aux = torch.empty((self.cap, 4), dtype=torch.float32, device=self.device)
        self.count = torch.zeros((2,), dtype=torch.int32, device=self.device)

    def extract(self, image):
        if isinstance(image, np.ndarray):
            image = torch.from_numpy(np.ascontiguousarray(image)).to(self.device)
        if image.dtype != torch.uint8 or not image.is_contiguous() or not (image.is_cuda or image.is_pinned()):
            raise ValueError("OrbExtractor.extract: image must be contiguous uint8 on the device or in pinned host memory")
        if tuple(image.shape[:2]) != (self.H, self.W) or (image.dim() == 3 and image.shape[2] != 3) or image.dim() not in (2, 3):
            raise ValueError(f"OrbExtractor.extract: expected [{self.H},{self.W}] or [{self.H},{self.W},3], got {tuple(image.shape)}")
        with torch.cuda.device(self.device):
            check(self.ctx.lib.vo_orb_extract(self.handle, ctypes.c_void_p(image.data_ptr()), 1 if image.dim() == 2 else 3,
                                              ctypes.c_void_p(self.kp.data_ptr()), ctypes.c_void_p(self.desc.data_ptr()),
                                              ctypes.c_void_p(self.aux.data_ptr()), ctypes.c_void_p(self.count.data_ptr()),
                                              ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)), "vo_orb_extract")
        n, overflow = (int(v) for v in self.count.cpu())
        if overflow:
            raise VoError("vo_orb_extract: a pyramid level kept more tied keypoints than the output holds")
        return self.kp[:n], self.desc[:n], self.aux[:n]

    def close(self):
        if self.handle:
            self.ctx.lib.vo_orb_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
