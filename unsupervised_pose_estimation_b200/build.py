"""Build libvsl_b200.so (the C-ABI CUDA library, include/vsl.h) in-tree with nvcc for sm_100a.

    python -m unsupervised_pose_estimation_b200.build [--force]

The library has no torch dependency: plain CUDA runtime, `extern "C"` entry points.
"""
from __future__ import annotations

import contextlib
import fcntl
import os
import shutil
import subprocess
import sys
import tempfile

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_PATH = os.path.join(PKG_DIR, "libvsl_b200.so")
SOURCES = ["vsl_fused.cu", "vsl_layers.cu", "vsl_input.cu", "vsl_source_grad.cu", "vsl_metrics.cu", "vsl_augment.cu"]
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


@contextlib.contextmanager
def build_lock():
    """Exclusive inter-process lock around "is it stale? -> build -> load": under torchrun every rank imports
    the package at once, and only one may run nvcc while the others wait for the finished file."""
    fd = os.open(os.path.join(PKG_DIR, ".build.lock"), os.O_CREAT | os.O_RDWR, 0o644)
    try:
        fcntl.flock(fd, fcntl.LOCK_EX)
        yield
    finally:
        fcntl.flock(fd, fcntl.LOCK_UN)
        os.close(fd)


def build(force=False, verbose=False, out_path=None, extra_flags=()):
    """Compile the library if it is missing or older than its sources. Returns the .so path.

    nvcc writes to a temporary file in the same directory which is then renamed onto the target, so a
    concurrent ``ctypes.CDLL`` never sees a half-written ELF.  ``out_path`` / ``extra_flags``: developer
    knobs for building kernel variants next to the product library (tools/)."""
    target = out_path or LIB_PATH
    if out_path is None and not force and not is_stale():
        return LIB_PATH
    nvcc = find_nvcc()
    if nvcc is None:
        raise RuntimeError("nvcc not found: cannot build libvsl_b200.so (there is no CPU fallback)")
    extra = os.environ.get("VSL_NVCC_EXTRA", "").split() + list(extra_flags)  # e.g. -DVSL_EXACT_RCP
    fd, tmp = tempfile.mkstemp(prefix=".libvsl_b200.", suffix=".so.tmp", dir=os.path.dirname(target))
    os.close(fd)
    cmd = [nvcc] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + \
        ["-o", tmp] + [os.path.join(CSRC, f) for f in SOURCES]
    try:
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("nvcc failed:\n%s\n%s" % (" ".join(cmd), res.stderr))
        os.chmod(tmp, 0o755)
        os.replace(tmp, target)
    finally:
        if os.path.exists(tmp):
            os.unlink(tmp)
    if verbose:
        print(res.stderr)
    return target


if __name__ == "__main__":
    with build_lock():
        print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
