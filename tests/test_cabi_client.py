"""The drop-in boundary without Python on the caller's side: tests/cabi/cabi_client.cpp is plain C++ + the CUDA
runtime, includes include/vsl.h and links libvsl_b200.so.  CPU: it compiles and links.  GPU: its outputs equal
the oracles'."""
import os
import shutil
import subprocess

import numpy as np
import pytest

from unsupervised_pose_estimation_b200 import _lib, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cabi", "cabi_client.cpp")
OUT_DIR = os.path.join(ROOT, "tests", "cabi", "_build")


def build_client():
    nvcc = build.find_nvcc()
    if nvcc is None:
        pytest.skip("nvcc not available")
    _lib.load()   # builds the library if needed
    os.makedirs(OUT_DIR, exist_ok=True)
    exe = os.path.join(OUT_DIR, "cabi_client")
    lib_dir = os.path.dirname(_lib.lib_path())
    cmd = [nvcc, "-std=c++17", "-O1", "-x", "cu", SRC, "-o", exe, "-L" + lib_dir, "-lvsl_b200",
           "-Xlinker", "-rpath," + lib_dir, "-cudart", "static"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    return exe


def test_client_compiles_and_links_without_torch():
    exe = build_client()
    ldd = shutil.which("ldd")
    if ldd:
        deps = subprocess.run([ldd, exe], capture_output=True, text=True).stdout
        names = [line.split()[0] for line in deps.splitlines() if line.strip()]
        assert any(n.startswith("libvsl_b200") for n in names)
        assert not any(n.startswith(("libtorch", "libc10", "libpython")) for n in names), names
    res = subprocess.run([exe], capture_output=True, text=True)   # usage error, no CUDA call
    assert res.returncode == 1 and "usage" in res.stderr


@pytest.mark.gpu
def test_client_pyramid_and_ssim_equal_the_oracles(tmp_path):
    import torch
    from oracle import pil_pyramid_oracle as P
    from oracle import vsl_oracle as O
    exe = build_client()
    rng = np.random.RandomState(11)
    B, H, W, L = 2, 48, 80, 4
    frames = rng.randint(0, 256, (B, H, W, 3)).astype(np.uint8)
    fin, fout = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    with open(fin, "wb") as f:
        f.write(np.array([B, H, W, L], np.int32).tobytes())
        f.write(frames.tobytes())
    res = subprocess.run([exe, "pyramid", fin, fout], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    levels, tensors = P.pyramid(frames, L)
    want = b"".join(levels[s].tobytes() for s in range(1, L)) + tensors[L - 1].tobytes()
    assert open(fout, "rb").read() == want

    x, y = torch.rand(2, 3, 20, 36), torch.rand(2, 3, 20, 36)
    with open(fin, "wb") as f:
        f.write(np.array([2, 3, 20, 36], np.int32).tobytes())
        f.write(x.numpy().tobytes())
        f.write(y.numpy().tobytes())
    res = subprocess.run([exe, "ssim", fin, fout], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    got = torch.from_numpy(np.frombuffer(open(fout, "rb").read(), np.float32).reshape(2, 3, 20, 36).copy())
    ref = O.ssim(x.cuda(), y.cuda()).cpu()   # the reference's SSIM arithmetic on the same GPU: bit-exact
    assert torch.equal(got, ref)
