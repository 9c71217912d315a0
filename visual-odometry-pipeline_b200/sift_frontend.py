"""SIFT front-end on the GPU (vo_sift_create / vo_sift_extract): cv2.SIFT_create().detectAndCompute with the reference's
default parameters (feature_extractors/SIFT.py:10-23), held to a tolerance (same keypoints to 1e-2 px / 0.25 degrees,
descriptor entries within 1 — OpenCV's own low-order bits depend on the host CPU).

Verified against the CPU restatement under the host emulation (tests/test_sift_emulation.py) and on a B200
(tests/test_gpu_sift_frontend.py, strict).  Default extractor of the drop-in plug-in feature_extractors/SIFT.py."""
import ctypes

import numpy as np
import torch

from . import ops
from ._lib import SiftConfig, VoError, check


class SiftExtractor:
    """sift = SiftExtractor(H, W); kp, desc, aux = sift.extract(image_uint8)   (image [H,W] gray or [H,W,3] BGR).

    kp [n,2] float32 = KeyPoint.pt, desc [n,128] float32 (0..255), aux [n,4] float32 = (size, angle, response, octave word);
    rows in OpenCV's order (sorted by x, y, ...; duplicates removed)."""

    def __init__(self, H, W, max_keypoints=32768, device=None):
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.ctx = ops.context(self.device)
        self.H, self.W, self.cap = int(H), int(W), int(max_keypoints)
        cfg = SiftConfig(self.H, self.W, self.cap)
        h = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            check(self.ctx.lib.vo_sift_create(self.ctx.handle, ctypes.byref(cfg), ctypes.byref(h)), "vo_sift_create")
        self.handle = h
        self.kp = torch.empty((self.cap, 2), dtype=torch.float32, device=self.device)
        self.desc = torch.empty((self.cap, 128), dtype=torch.float32, device=self.device)
        self.aux = torch.empty((self.cap, 4), dtype=torch.float32, device=self.device)
        self.count = torch.zeros((2,), dtype=torch.int32, device=self.device)

    def extract(self, image):
        if isinstance(image, np.ndarray):
            image = torch.from_numpy(np.ascontiguousarray(image)).to(self.device)
        if image.dtype != torch.uint8 or not image.is_contiguous() or not (image.is_cuda or image.is_pinned()):
            raise ValueError("SiftExtractor.extract: image must be contiguous uint8 on the device or in pinned host memory")
        if tuple(image.shape[:2]) != (self.H, self.W) or (image.dim() == 3 and image.shape[2] != 3) or image.dim() not in (2, 3):
            raise ValueError(f"SiftExtractor.extract: expected [{self.H},{self.W}] or [{self.H},{self.W},3], got {tuple(image.shape)}")
        p = lambda t: ctypes.c_void_p(t.data_ptr())  # noqa: E731
        with torch.cuda.device(self.device):
            check(self.ctx.lib.vo_sift_extract(self.handle, p(image), 1 if image.dim() == 2 else 3, p(self.kp), p(self.desc),
                                               p(self.aux), p(self.count),
                                               ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)), "vo_sift_extract")
        n, raw = (int(v) for v in self.count.cpu())
        if raw > self.cap:
            raise VoError(f"vo_sift_extract: {raw} candidates exceed max_keypoints = {self.cap}")
        return self.kp[:n], self.desc[:n], self.aux[:n]

    def close(self):
        if self.handle:
            self.ctx.lib.vo_sift_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
