"""GPU probe: bmm rounding order for batch 1 (cuBLAS may pick another kernel). Developer tool."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import vsl_oracle as O
from unsupervised_pose_estimation_b200 import synthetic, layers as L
from unsupervised_pose_estimation_b200 import functional as VF
torch.backends.cuda.matmul.allow_tf32 = False
dev = "cuda"
for (B, H, W) in [(1, 32, 64), (1, 192, 640), (2, 32, 64), (1, 64, 96), (3, 32, 64), (12, 192, 640), (24, 192, 640), (1, 320, 1024)]:
    inputs, outputs, leaves = synthetic.make_batch(B, H, W, [0, -1, 1], seed=8, family="iid", device=dev,
                                                   pose_fn=L.transformation_from_parameters, requires_grad=False)
    depth = 0.1 + 5 * torch.rand(B, 1, H, W, device=dev)
    K, iK = inputs[("K", 0)], inputs[("inv_K", 0)]
    T = outputs[("cam_T_cam", 0, -1)]
    cam_ref = O.backproject(depth, iK)
    pix_ref = O.project(cam_ref, K, T, H, W)
    P = torch.matmul(K, T)[:, :3, :]
    res = {}
    for name, arith in {"cuda": 0, "nofma": 2, "reverse": 4}.items():
        cam = VF.backproject(depth, iK, arith)
        pix = VF.project(cam_ref, P, H, W, 1e-7, arith)
        res[name] = (int((cam != cam_ref).sum()), int((pix != pix_ref).sum()))
    # raw bmm against explicit formulas in fp64-emulated fma
    c = torch.matmul(P, cam_ref)
    print((B, H, W), res, "numel", cam_ref.numel(), pix_ref.numel())
