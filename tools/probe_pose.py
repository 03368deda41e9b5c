"""Developer probe: bit-exactness of an own axis-angle -> rotation kernel against torch's op sequence
(reference layers.py:133-172) on the GPU."""
import ctypes, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from unsupervised_pose_estimation_b200 import layers as L
lib = ctypes.CDLL(os.path.join(ROOT, "tools", "ubench", "libposetest.so"))
lib.run_rot.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
dev = "cuda"
for scale in (0.01, 0.3, 3.0):
    n = 1 << 20
    v = (scale * torch.randn(n, 1, 3, generator=torch.Generator().manual_seed(1))).to(dev)
    R = L.rot_from_axisangle(v)
    angle = torch.norm(v, 2, 2, True)
    ref = torch.cat([angle.view(n, 1), torch.cos(angle).view(n, 1), torch.sin(angle).view(n, 1), R[:, :3, :3].reshape(n, 9)], 1).contiguous()
    for variant in range(4):
        out = torch.empty(n, 12, device=dev)
        rc = lib.run_rot(n, variant, v.data_ptr(), out.data_ptr(), torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        mism = (out != ref).sum(0).tolist()
        print("scale", scale, "variant", variant, "rc", rc, "mismatches per column [angle, cos, sin, R00..R22]:", mism)
