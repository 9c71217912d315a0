"""Loader for the drop-in modules (VisualOdometry_Stereo, feature_extractors.*, R2D2) the way the reference is run: the
current directory holds config/vo_params.yaml, which VisualOdometry_Stereo.py reads at import time to pick the feature
plug-in (reference: VisualOdometry_Stereo.py:16-24).  Used by the tests, tools/seq_bench.py and bench.py's `api_rate` block."""
import importlib
import os
import sys

PKG = os.path.dirname(os.path.abspath(__file__))


def load(workdir, extractor, extra=""):
    """Write config/vo_params.yaml (the shipped one with `feature_extractor` set and `extra` lines appended) under `workdir`,
    chdir there and import a fresh VisualOdometry_Stereo.  The caller restores the working directory."""
    cfg = os.path.join(str(workdir), "config")
    os.makedirs(cfg, exist_ok=True)
    src = open(os.path.join(PKG, "config", "vo_params.yaml")).read().replace('feature_extractor: "orb"', f'feature_extractor: "{extractor}"')
    with open(os.path.join(cfg, "vo_params.yaml"), "w") as f:
        f.write(src + extra)
    os.chdir(str(workdir))
    if PKG not in sys.path:
        sys.path.insert(0, PKG)
    for name in ("VisualOdometry_Stereo", "vo_stereo_runner", "vo_runner", "feature_extractors.ORB", "feature_extractors.SIFT"):
        sys.modules.pop(name, None)
    return importlib.import_module("VisualOdometry_Stereo")
