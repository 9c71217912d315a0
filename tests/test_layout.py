"""The product never touches the oracle, the reference tree or a CPU fallback."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "visual-odometry-pipeline_b200")


def _sources():
    for base, _, files in os.walk(PKG):
        if "build" in base.split(os.sep):
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".sh")):
                yield os.path.join(base, f)


def test_product_does_not_import_oracle_or_reference():
    pat = re.compile(r"(from\s+oracle|import\s+oracle|/root/reference|oracle/)")
    bad = [p for p in _sources() if pat.search(open(p).read())]
    assert not bad, bad


def test_no_compat_layers_in_product():
    pat = re.compile(r"\b(import triton|torch\.compile|tilelang)\b")
    bad = [p for p in _sources() if pat.search(open(p).read())]
    assert not bad, bad


def test_required_files_exist():
    for rel in ("include/vo_b200.h", "oracle/vo_oracle.c", "oracle/oracle.py", "bench.py", "__graft_entry__.py",
                "DESIGN.md", "INTEGRATION.md", "tests/golden/make_golden.py", "oracle/orb_frontend.py", "oracle/sift_frontend.py",
                "tests/golden/make_orb_golden.py", "tests/golden/make_sift_golden.py", "tests/golden/orb_pattern.npy"):
        assert os.path.exists(os.path.join(ROOT, rel)), rel
