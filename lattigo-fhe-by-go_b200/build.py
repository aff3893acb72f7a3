"""Builds the CUDA library in-tree: lattigo-fhe-by-go_b200/lib/liblattigpu.so (sm_100a only).

    python lattigo-fhe-by-go_b200/build.py [--force] [--verbose]
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
SO = os.path.join(LIBDIR, "liblattigpu.so")
SOURCES = ["capi.cu", "ringext.cu", "capi_ext.cu", "capi_bfv.cu", "capi_shard.cu", "ntt.cu", "elementwise.cu", "permute.cu", "basisext.cu", "keyswitch.cu", "crp.cu", "scaler.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC,-O2", "--fmad=false",
]


def _newest_source():
    t = 0.0
    for root, _, files in os.walk(CSRC):
        for f in files:
            t = max(t, os.path.getmtime(os.path.join(root, f)))
    t = max(t, os.path.getmtime(os.path.join(HERE, "..", "include", "lattigpu.h")))
    return t


def build(force=False, verbose=False):
    os.makedirs(LIBDIR, exist_ok=True)
    if not force and os.path.exists(SO) and os.path.getmtime(SO) >= _newest_source():
        return SO
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    objs = []
    procs = []
    for src in SOURCES:
        obj = os.path.join(LIBDIR, src.replace(".cu", ".o"))
        cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(CSRC, src), "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            sys.stderr.write("== %s\n%s\n" % (src, out))
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed")
    subprocess.check_call([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-Xlinker", "--no-undefined", "-o", SO] + objs + ["-lcudart_static", "-lpthread", "-ldl", "-lrt"])
    return SO


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
