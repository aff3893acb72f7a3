"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads and
exports every symbol include/lattigpu.h declares, the ctypes binding covers all
of them, and the product path fails loudly (no CPU fallback) without a GPU.
No compute entry point is called here."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "lattigo-fhe-by-go_b200")


@pytest.fixture(scope="module")
def built():
    import importlib.util

    spec = importlib.util.spec_from_file_location("lg_build", os.path.join(PKG, "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.build()


def test_library_exports_every_header_symbol(built):
    from lattigpu import _lib

    names = _lib.header_symbols()
    assert len(names) > 90
    out = subprocess.check_output(["nm", "-D", "--defined-only", built], text=True)
    exported = set(re.findall(r"\bT (lg_[a-z0-9_]+)", out))
    missing = [n for n in names if n not in exported]
    assert not missing, missing


def test_binding_covers_header(built):
    from lattigpu import _lib

    assert set(_lib.header_symbols()) == set(_lib.SIGNATURES)
    L = _lib.lib()
    assert L.lg_version().startswith(b"lattigpu")


def test_product_does_not_touch_oracle():
    """the oracle is test infrastructure: nothing under the package may import, link or run it"""
    for root, _, files in os.walk(PKG):
        if os.path.basename(root) in ("lib", "__pycache__"):
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                src = open(os.path.join(root, f)).read()
                assert "oracle" not in src.replace("test infrastructure", ""), os.path.join(root, f)


def test_no_cpu_fallback(built):
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from lattigpu import LattigpuError, ring

    with pytest.raises(LattigpuError, match="no CUDA device"):
        ring.NewContextWithParams(8, [576460752303439873])


def test_sass_is_sm100a(built):
    out = subprocess.check_output(["cuobjdump", "-lelf", built], text=True)
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs
