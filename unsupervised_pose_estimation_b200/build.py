"""Build libvsl_b200.so (the C-ABI CUDA library, include/vsl.h) in-tree with nvcc for sm_100a.

    python -m unsupervised_pose_estimation_b200.build [--force]

The library has no torch dependency: plain CUDA runtime, `extern "C"` entry points.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_PATH = os.path.join(PKG_DIR, "libvsl_b200.so")
SOURCES = ["vsl_fused.cu", "vsl_layers.cu", "vsl_input.cu"]
HEADERS = ["vsl_math.cuh", "vsl_tile.cuh", os.path.join("..", "..", "include", "vsl.h")]
NVCC_FLAGS = [
    "-std=c++17", "-O3", "-lineinfo", "--threads", "4",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-Xcompiler", "-fPIC", "-shared", "-cudart", "static",
]


def find_nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    return None


def is_stale():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    """Compile the library if it is missing or older than its sources. Returns the .so path."""
    if not force and not is_stale():
        return LIB_PATH
    nvcc = find_nvcc()
    if nvcc is None:
        raise RuntimeError("nvcc not found: cannot build libvsl_b200.so (there is no CPU fallback)")
    extra = os.environ.get("VSL_NVCC_EXTRA", "").split()  # developer knob, e.g. -DVSL_CTAS_PER_SM=2
    cmd = [nvcc] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + \
        ["-o", LIB_PATH] + [os.path.join(CSRC, f) for f in SOURCES]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n%s\n%s" % (" ".join(cmd), res.stderr))
    if verbose:
        print(res.stderr)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
