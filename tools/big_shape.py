"""Developer check (GPU box): one large image size against the oracle (index arithmetic at scale)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch

import test_gpu_parity as T
from oracle import vsl_oracle as O
from unsupervised_pose_estimation_b200 import synthetic

for (B, H, W, frames) in [(2, 1024, 2048, [0, -1, 1]), (1, 1536, 2560, [0, -1, 1, "s"])]:
    opt = O.make_opt(height=H, width=W, batch_size=B, frame_ids=list(frames))
    inputs, outputs, leaves = synthetic.make_batch(B, H, W, frames, synthetic.K_KITTI, seed=3, family="smooth", device=T.DEV)
    ref_out, ref_losses, ref_g = T.run_oracle(opt, inputs, outputs, leaves, seed=1)
    out, losses, g = T.run_ours(opt, inputs, outputs, leaves, seed=1, side="none")
    lerr = max(abs(losses[k].item() - ref_losses[k].item()) / abs(ref_losses[k].item()) for k in ref_losses)
    mism = sum(int((out["identity_selection/%d" % s] != ref_out["identity_selection/%d" % s]).sum()) for s in range(4))
    gerr = max(((g[k] - ref_g[k]).norm() / ref_g[k].norm()).item() for k in ref_g)
    print(B, H, W, frames, "loss err %.1e | mask mismatches %d | grad err %.1e" % (lerr, mism, gerr))
