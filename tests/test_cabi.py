"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads and
exports every symbol include/lattigpu.h declares, the ctypes binding covers all
of them, and the product path fails loudly (no CPU fallback) without a GPU.
No compute entry point is called here."""
import ctypes as C
import os
import re
import subprocess
import sys

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


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the CPU arm the driver times beside the GPU arm): one JSON line with the contract's
    keys, no GPU work.  Run on the smallest parameter set so the check takes a second."""
    import json
    import subprocess

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--params", "0",
                          "--cpu-sample", "2", "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "ops/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "ops/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0 and "not the headline set" in d["config"]["workload"]
    # under torchrun only rank 0 works: the other ranks exit 0 without output
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--gpus", "2"],
                         capture_output=True, text=True, timeout=300, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


C_EXAMPLE = os.path.join(ROOT, "examples", "c", "ring_smoke.c")


def _build_c_example(built, out):
    libdir = os.path.dirname(built)
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I" + os.path.join(ROOT, "include"),
                           C_EXAMPLE, "-L" + libdir, "-llattigpu", "-Wl,-rpath," + libdir, "-o", out])


def test_plain_c_program_builds_against_the_header(built, tmp_path):
    """the boundary is usable from plain C (what cgo sees): the header is C99 -pedantic clean and the example links"""
    out = str(tmp_path / "ring_smoke")
    _build_c_example(built, out)
    import torch

    if not torch.cuda.is_available():  # without a GPU it must fail loudly, not compute on the CPU
        res = subprocess.run([out], capture_output=True, text=True)
        assert res.returncode != 0 and "no CUDA device" in res.stderr


@pytest.mark.gpu
def test_plain_c_program_runs(built, tmp_path):
    """NTT round trip, a key switch and an argument check driven from C, no Python in the process"""
    out = str(tmp_path / "ring_smoke")
    _build_c_example(built, out)
    res = subprocess.run([out, "0"], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0 and res.stdout.strip().endswith("ring_smoke: ok"), (res.stdout, res.stderr)


def test_go_shim_calls_only_declared_entry_points():
    """go/ring/gpu.go and go/ckks/evaluator_gpu.go (complete cgo bindings; no Go toolchain in the image) may only call
    functions that include/lattigpu.h declares"""
    from lattigpu import _lib

    declared = set(_lib.header_symbols())
    for rel in ("go/ring/gpu.go", "go/ckks/evaluator_gpu.go"):
        src = open(os.path.join(ROOT, rel)).read()
        # no elisions: "..." may only appear as Go's variadic spread inside append(...)
        assert not [ln for ln in src.splitlines() if "..." in ln and not re.search(r"append\(.*\.\.\.\)", ln)], rel
        used = set(re.findall(r"C\.(lg_[a-z0-9_]+)\(", src)) - {"lg_stream_t"}  # the type conversion C.lg_stream_t(...)
        assert used and not (used - declared), (rel, sorted(used - declared))
