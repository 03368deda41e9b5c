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


@pytest.mark.gpu
def test_client_fused_loss_forward_backward_equals_the_oracle(tmp_path):
    """vsl_loss_forward_backward + vsl_loss_combine_grads called from plain C++ (no Python, no torch on the caller's
    side): auto-masks are the oracle's bits, losses 1e-6, gradients w.r.t. the disparities and poses 5e-5."""
    import torch
    from oracle import vsl_oracle as O
    from unsupervised_pose_estimation_b200 import synthetic
    torch.backends.cuda.matmul.allow_tf32 = False
    exe = build_client()
    B, H, W, frames, S = 2, 64, 96, [0, -1, 1], 4
    F = len(frames) - 1
    opt = O.make_opt(height=H, width=W, batch_size=B, frame_ids=frames)
    inputs, outputs, leaves = synthetic.make_batch(B, H, W, frames, seed=5, family="smooth", device="cuda")
    Ts = {f: O.transformation_from_parameters(leaves[("axisangle", 0, f)][:, 0], leaves[("translation", 0, f)][:, 0], f < 0)
          .detach().requires_grad_(True) for f in frames[1:]}
    out = {("disp", s): leaves[("disp", s)] for s in range(S)}
    out.update({("cam_T_cam", 0, f): Ts[f] for f in frames[1:]})
    gen = torch.Generator().manual_seed(9)
    noise = [torch.randn(B, F, H, W, generator=gen).cuda() for _ in range(S)]
    O.generate_images_pred(opt, inputs, out)
    ref = O.compute_losses(opt, inputs, out, noise)
    wrt = [leaves[("disp", s)] for s in range(S)] + [Ts[f] for f in frames[1:]]
    ref_g = torch.autograd.grad(ref["loss"], wrt)
    arrays = ([inputs[("color", 0, s)] for s in range(S)] + [inputs[("color", f, 0)] for f in frames[1:]]
              + [leaves[("disp", s)].detach() for s in range(S)] + [inputs[("inv_K", 0)], inputs[("K", 0)]]
              + [Ts[f].detach() for f in frames[1:]] + noise)
    upstream = np.zeros(2 * S + 1, np.float32)
    upstream[2 * S] = 1.0   # d/d losses["loss"]
    fin, fout = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    with open(fin, "wb") as f:
        f.write(np.array([B, H, W, S, F], np.int32).tobytes())
        f.write(np.array([np.float32(1 / opt.max_depth), np.float32(1 / opt.min_depth - 1 / opt.max_depth),
                          opt.disparity_smoothness], np.float32).tobytes())
        for a in arrays:
            f.write(a.contiguous().cpu().numpy().astype(np.float32).tobytes())
        f.write(upstream.tobytes())
    res = subprocess.run([exe, "loss", fin, fout], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    got = np.frombuffer(open(fout, "rb").read(), np.float32)
    off = 0

    def take(shape):
        nonlocal off
        n = int(np.prod(shape))
        t = torch.from_numpy(got[off:off + n].reshape(shape).copy())
        off += n
        return t
    losses = take((3 * S + 1,))
    for s in range(S):
        for key, v in (("min_loss/%d" % s, losses[s]), ("loss/%d" % s, losses[S + s])):
            assert abs(v.item() - ref[key].item()) <= 1e-6 * abs(ref[key].item()), key
    assert abs(losses[2 * S].item() - ref["loss"].item()) <= 1e-6 * ref["loss"].item()
    for s in range(S):
        assert torch.equal(take((B, H, W)), out["identity_selection/%d" % s].cpu()), s
    for s in range(S):
        g = take((B, 1, H >> s, W >> s))
        assert ((g - ref_g[s].cpu()).norm() / ref_g[s].cpu().norm()).item() <= 5e-5, s
    gT = take((F, B, 4, 4))
    for i in range(F):
        r = ref_g[S + i].cpu()
        assert ((gT[i] - r).norm() / r.norm()).item() <= 5e-5, i
    assert off == got.size
