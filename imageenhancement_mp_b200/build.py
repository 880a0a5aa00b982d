"""Builds ``libimgenh_b200.so`` (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

    python -m imageenhancement_mp_b200.build [--force] [--verbose]

nvcc cross-compiles without a GPU, so this runs in the CPU-only build container; the
resulting .so is git-ignored but travels to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import concurrent.futures as cf
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
ROOT = os.path.dirname(HERE)
LIB = os.path.join(HERE, "libimgenh_b200.so")
OBJ = os.path.join(HERE, "build")

SOURCES = ["api.cu", "conv_tcgen05.cu", "debug_conv.cu", "layout.cu", "kpn_apply.cu", "kpn_tcgen05.cu", "metrics.cu", "preprocess.cu"]
HEADERS = ["ie_common.cuh", "ie_ptx.cuh"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
]


def nvcc():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: the CUDA extension cannot be built")
    return exe


def _digest():
    h = hashlib.sha256()
    for f in SOURCES + HEADERS:
        with open(os.path.join(CSRC, f), "rb") as fh:
            h.update(fh.read())
    with open(os.path.join(ROOT, "include", "imgenh_b200.h"), "rb") as fh:
        h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force=False, verbose=False):
    """Compile if sources changed since the last build; returns the path of the .so."""
    os.makedirs(OBJ, exist_ok=True)
    stamp = os.path.join(OBJ, "stamp")
    dig = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read() == dig:
        return LIB
    exe = nvcc()
    extra = ["-Xptxas", "-v"] if verbose else []

    def compile_one(src):
        obj = os.path.join(OBJ, src.replace(".cu", ".o"))
        cmd = [exe] + NVCC_FLAGS + extra + ["-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        return src, obj, r

    objs = []
    with cf.ThreadPoolExecutor(max_workers=min(8, len(SOURCES))) as ex:
        for src, obj, r in ex.map(compile_one, SOURCES):
            if verbose or r.returncode != 0:
                sys.stderr.write(f"--- {src}\n{r.stdout}{r.stderr}\n")
            if r.returncode != 0:
                raise RuntimeError(f"nvcc failed on {src}")
            objs.append(obj)
    cmd = [exe, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("link failed")
    with open(stamp, "w") as fh:
        fh.write(dig)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
