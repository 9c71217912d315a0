"""R2D2 front-end on the GPU (vo_r2d2_*): `extract_features_and_desc` of the reference's R2D2.py:202-232 — network
forward (feature_extractors/r2d2/nets/patchnet.py), reliability / repeatability heads, 3x3 NMS with both thresholds
(R2D2.py:82-101), score filter (:186-188) — for one image at scale 1, which is all the reference's extract_multiscale
does (its loop `break`s after the first scale, :133-135).

The checkpoint format is the reference's: {'net': "<class name>(...)", 'state_dict': {...}} (R2D2.py:68-79).  The two
architectures the shipped models use are described here as data (non-parametric layers — ReLU, pooling, dilation,
up-sampling — are not in a state_dict).
"""
import ctypes

import numpy as np
import torch

from . import ops
from ._lib import R2d2Config, R2d2Layer, check

# (kernel, dilation, batch-norm, relu, pool_after) per convolution, in order; channel counts come from the weights.
# patchnet.py:56-66: stride-s layers of a dilated net keep stride 1 and multiply the dilation of the FOLLOWING layers.
ARCH = {
    # Quad_L2Net_ConfCFS (:104-118, :121-139): r2d2_WASF_N16.pt, r2d2_WAF_N16.pt, r2d2_WASF_N8_big.pt
    "Quad_L2Net_ConfCFS": dict(upsample=1, layers=[(3, 1, 1, 1, 0), (3, 1, 1, 1, 0), (3, 1, 1, 1, 0), (3, 2, 1, 1, 0), (3, 2, 1, 1, 0),
                                                   (3, 4, 1, 1, 0), (2, 4, 1, 0, 0), (2, 8, 1, 0, 0), (2, 16, 0, 0, 0)]),
    # Fast_Quad_L2Net_ConfCFS (:141-186): faster2d2_WASF_N16.pt — MaxPool2d(2) after the third layer, Upsample x2 at the end
    "Fast_Quad_L2Net_ConfCFS": dict(upsample=2, layers=[(3, 1, 1, 1, 0), (3, 1, 1, 1, 0), (3, 1, 1, 1, 2), (3, 1, 1, 1, 0), (3, 1, 1, 1, 0),
                                                        (3, 2, 1, 1, 0), (2, 2, 1, 0, 0), (2, 4, 1, 0, 0), (2, 8, 0, 0, 0)]),
}


def load_checkpoint(path):
    """The reference's load_network (R2D2.py:68-79) without instantiating its classes: -> (net name, {key: ndarray})."""
    try:
        ck = torch.load(path, map_location="cpu", weights_only=True)
    except Exception:
        ck = torch.load(path, map_location="cpu", weights_only=False)
    name = str(ck["net"]).split("(")[0].strip()
    sd = {k.replace("module.", ""): v.detach().cpu().numpy() for k, v in ck["state_dict"].items() if hasattr(v, "detach")}
    return name, sd


def layer_table(name, sd):
    """[(dict per convolution)] with weights in the library's layout [C_out][k][k][C_in]."""
    if name not in ARCH:
        raise ValueError(f"unknown R2D2 architecture {name!r} (known: {sorted(ARCH)})")
    conv_ids = sorted(int(k.split(".")[1]) for k in sd if k.startswith("ops.") and k.endswith(".weight") and sd[k].ndim == 4)
    spec = ARCH[name]["layers"]
    if len(conv_ids) != len(spec):
        raise ValueError(f"{name}: checkpoint has {len(conv_ids)} convolutions, the architecture {len(spec)}")
    out = []
    for i, (k, dil, bn, relu, pool) in zip(conv_ids, spec):
        w = np.asarray(sd[f"ops.{i}.weight"], np.float32)
        if w.shape[2] != k or w.shape[3] != k:
            raise ValueError(f"{name}: ops.{i} has {w.shape[2]}x{w.shape[3]} taps, expected {k}x{k}")
        d = dict(cout=int(w.shape[0]), cin=int(w.shape[1]), k=k, dil=dil, bn=bn, relu=relu, pool_after=pool,
                 w=np.ascontiguousarray(w.transpose(0, 2, 3, 1)), bias=np.ascontiguousarray(sd[f"ops.{i}.bias"], dtype=np.float32))
        if bn:
            d["bn_mean"] = np.ascontiguousarray(sd[f"ops.{i + 1}.running_mean"], dtype=np.float32)
            d["bn_var"] = np.ascontiguousarray(sd[f"ops.{i + 1}.running_var"], dtype=np.float32)
        out.append(d)
    return out


class R2D2Net:
    """net = R2D2Net(name, state_dict, H, W); xys, desc, scores = net.extract(rgb_uint8)."""

    def __init__(self, name, sd, H, W, max_kp=30000, device=None, bn_eps=1e-5):
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.ctx = ops.context(self.device)
        self.layers = layer_table(name, sd)
        self.H, self.W, self.max_kp = int(H), int(W), int(max_kp)
        arr = (R2d2Layer * len(self.layers))()
        p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
        for i, L in enumerate(self.layers):
            arr[i].cin, arr[i].cout, arr[i].k, arr[i].dil = L["cin"], L["cout"], L["k"], L["dil"]
            arr[i].bn, arr[i].relu, arr[i].pool_after = L["bn"], L["relu"], L["pool_after"]
            arr[i].w, arr[i].bias = p(L["w"]), p(L["bias"])
            arr[i].bn_mean = p(L["bn_mean"]) if L["bn"] else None
            arr[i].bn_var = p(L["bn_var"]) if L["bn"] else None
        C = self.layers[-1]["cout"]
        self.C = C
        self._heads = [np.ascontiguousarray(sd["clf.weight"], dtype=np.float32).reshape(2, C),
                       np.ascontiguousarray(sd["clf.bias"], dtype=np.float32).reshape(2),
                       np.ascontiguousarray(sd["sal.weight"], dtype=np.float32).reshape(C),
                       np.ascontiguousarray(sd["sal.bias"], dtype=np.float32).reshape(1)]
        cfg = R2d2Config()
        cfg.H, cfg.W, cfg.n_layers, cfg.layers = self.H, self.W, len(self.layers), arr
        cfg.upsample = ARCH[name]["upsample"]
        cfg.clf_w, cfg.clf_b, cfg.sal_w, cfg.sal_b = (p(a) for a in self._heads)
        cfg.bn_eps, cfg.max_kp = float(bn_eps), self.max_kp
        h = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            check(self.ctx.lib.vo_r2d2_create(self.ctx.handle, ctypes.byref(cfg), ctypes.byref(h)), "vo_r2d2_create")
        self.handle = h
        ho, wo = ctypes.c_int(), ctypes.c_int()
        check(self.ctx.lib.vo_r2d2_out_shape(h, ctypes.byref(ho), ctypes.byref(wo)), "vo_r2d2_out_shape")
        self.Ho, self.Wo = ho.value, wo.value
        self.xys = torch.empty((self.max_kp, 3), dtype=torch.float32, device=self.device)
        self.desc = torch.empty((self.max_kp, C), dtype=torch.float32, device=self.device)
        self.scores = torch.empty((self.max_kp,), dtype=torch.float32, device=self.device)
        self.count = torch.zeros((1,), dtype=torch.int32, device=self.device)

    @classmethod
    def from_checkpoint(cls, path, H, W, **kw):
        name, sd = load_checkpoint(path)
        return cls(name, sd, H, W, **kw)

    def extract(self, rgb, rel_thr=0.7, rep_thr=0.7, score_thr=0.85, want_maps=False):
        """rgb: uint8 [H,W,3] (numpy, or a torch tensor on the device / in pinned memory).  Returns (xys [n,3], desc
        [n,C], scores [n]) as CUDA tensors (views of the net's output buffers) — plus (rel, rep) maps if asked."""
        if isinstance(rgb, np.ndarray):
            rgb = torch.from_numpy(np.ascontiguousarray(rgb, dtype=np.uint8))
        if rgb.dtype != torch.uint8 or tuple(rgb.shape) != (self.H, self.W, 3) or not rgb.is_contiguous():
            raise ValueError(f"R2D2Net.extract: expected contiguous uint8 [{self.H},{self.W},3], got {rgb.dtype} {tuple(rgb.shape)}")
        rel = rep = None
        if want_maps:
            rel = torch.empty((self.Ho, self.Wo), dtype=torch.float32, device=self.device)
            rep = torch.empty((self.Ho, self.Wo), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
            if not rgb.is_cuda and not rgb.is_pinned():
                torch.cuda.current_stream().synchronize()      # a pageable source is copied synchronously by the runtime
            check(self.ctx.lib.vo_r2d2_extract(self.handle, ctypes.c_void_p(rgb.data_ptr()), float(rel_thr), float(rep_thr),
                                               float(score_thr), ctypes.c_void_p(self.xys.data_ptr()),
                                               ctypes.c_void_p(self.desc.data_ptr()), ctypes.c_void_p(self.scores.data_ptr()),
                                               ctypes.c_void_p(self.count.data_ptr()),
                                               ctypes.c_void_p(rel.data_ptr()) if want_maps else None,
                                               ctypes.c_void_p(rep.data_ptr()) if want_maps else None, stream), "vo_r2d2_extract")
        self._keep = rgb
        n = min(int(self.count.item()), self.max_kp)
        out = (self.xys[:n], self.desc[:n], self.scores[:n])
        return out + (rel, rep) if want_maps else out

    def close(self):
        if getattr(self, "handle", None):
            self.ctx.lib.vo_r2d2_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
